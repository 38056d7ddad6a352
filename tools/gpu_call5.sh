#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/c5_pytest.log 2>&1; echo "pytest rc=$?"; grep -E "^FAILED|^ERROR|passed|failed" gpurun_out/c5_pytest.log | cut -c1-250 | head -40
cap() {
  timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$2" -s 3 -c 1 -o gpurun_out/r2_$1 python bench_configs.py --only $3 --iters 2 > gpurun_out/r2_$1.log 2>&1
  echo "$1 rc=$?"
}
cap rqs_fwd 'rqs_unrolled_kernel.*bool.0' C4
cap rqs_inv 'rqs_unrolled_kernel.*bool.1' C4
cap planar_inv 'planar_dev_kernel.*int.1, .bool.0' C2
cap planar_logpdf 'planar_dev_kernel.*int.1, .bool.1' C2
cap coupling_tc 'coupling_tc_kernel' C5
