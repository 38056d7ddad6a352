#!/bin/bash
# ncu --set full capture of the reverse-mode kernels (1 GPU); summaries are made on the box, the reports deleted.
mkdir -p gpurun_out
cap() {  # name, kernel regex (demangled)
  timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$2" -s 2 -c 1 -o gpurun_out/r2_$1 python tools/bench_vjp.py 128 8 > gpurun_out/r2_$1.log 2>&1
  echo "$1 rc=$?"
}
cap radial_vjp 'radial_vjp_kernel'
cap rqs_vjp_fwd 'rqs_vjp_kernel.*bool.0, .bool.1'
cap rqs_vjp_inv 'rqs_vjp_kernel.*bool.1, .bool.1'
cap coupling_vjp 'coupling_vjp_fast_kernel'
cap bn_vjp 'bn_eval_vjp_kernel'
for f in gpurun_out/r2_*.ncu-rep; do
  b=$(basename $f .ncu-rep)
  (python tools/summarise_ncu.py $f; python tools/ncu_hot.py $f --top 14) > gpurun_out/${b}_ncu_full.txt 2>&1
  rm -f $f
done
