#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 600 -k "rqs or config4 or (parity and rqs) or fused_chain" > gpurun_out/c2_pytest_rqs.log 2>&1; echo "pytest rc=$?"; tail -n 5 gpurun_out/c2_pytest_rqs.log
timeout 300 python bench_configs.py --only C4 --iters 20 > gpurun_out/c2_c4.log 2>&1; tail -n 3 gpurun_out/c2_c4.log | cut -c1-200
cap() {
  timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$2" -s 3 -c 1 -o gpurun_out/r2_$1 python bench_configs.py --only $3 --iters 2 > gpurun_out/r2_$1.log 2>&1
  echo "$1 rc=$?"
}
cap rqs_fwd 'rqs_unrolled_kernel.*bool.0' C4
cap rqs_inv 'rqs_unrolled_kernel.*bool.1' C4
cap planar_inv 'planar_dev_kernel.*int.1, .bool.0' C2
cap planar_logpdf 'planar_dev_kernel.*int.1, .bool.1' C2
cap planar_fwd 'planar_dev_kernel.*int.8, .int.8, .int.0, .bool.0' C2
