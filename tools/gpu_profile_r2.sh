#!/bin/bash
# Round-2 evidence visit (1 GPU): one `ncu --set full` capture per shipped hot kernel + the launch list of the bench.
# Numbers printed under ncu are never bench values.
mkdir -p gpurun_out
cap() {  # name, kernel regex (demangled), bench_configs config
  timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$2" -s 3 -c 1 -o gpurun_out/r2_$1 python bench_configs.py --only $3 --iters 2 > gpurun_out/r2_$1.log 2>&1
  echo "$1 rc=$?"
}
cap planar_fwd 'planar_dev_kernel.*int.8, .int.8, .int.0, .bool.0' C2
cap planar_inv 'planar_dev_kernel.*int.8, .int.8, .int.1, .bool.0' C2
cap planar_logpdf 'planar_dev_kernel.*int.1, .bool.1' C2
cap radial_fwd 'chain_v1_kernel' C3
cap radial_inv 'radial_unrolled' C3
cap rqs_fwd 'rqs_unrolled_kernel.*bool.0' C4
cap rqs_inv 'rqs_unrolled_kernel.*bool.1' C4
cap coupling_tc 'coupling_tc_kernel' C5
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:planar_vjp_kernel" -s 2 -c 1 -o gpurun_out/r2_planar_vjp python tools/bench_vjp.py 128 8 > gpurun_out/r2_planar_vjp.log 2>&1; echo "planar_vjp rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:chain_sample_kernel" -s 2 -c 1 -o gpurun_out/r2_sample python tools/bench_sample.py > gpurun_out/r2_sample.log 2>&1; echo "sample rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_bench.csv python bench.py --steps 5 --warmup 3 --profile > gpurun_out/r2_bench_under_ncu.log 2>&1; echo "launch list rc=$?"
timeout 300 python tools/bench_sample.py > gpurun_out/r2_sample_bench.log 2>&1; tail -n 4 gpurun_out/r2_sample_bench.log
# summaries are made HERE (ncu is on the box); the .ncu-rep files are too large to travel back (64 MiB cap)
for f in gpurun_out/r2_*.ncu-rep; do
  b=$(basename $f .ncu-rep)
  (python tools/summarise_ncu.py $f; python tools/ncu_hot.py $f --top 12) > gpurun_out/${b}_ncu_full.txt 2>&1
  rm -f $f
done
timeout 120 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "rand" 2>&1 | tail -2
