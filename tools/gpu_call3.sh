#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 600 -k "rqs_bin_counts" > gpurun_out/c3_pytest_rqs.log 2>&1; echo "pytest rqs rc=$?"; grep -E "AssertionError|assert |passed|failed" gpurun_out/c3_pytest_rqs.log | cut -c1-300 | head -40
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 600 -k "coupling or realnvp or config5 or fused_chain" > gpurun_out/c3_pytest_cpl.log 2>&1; echo "pytest coupling rc=$?"; grep -E "AssertionError|Error|passed|failed" gpurun_out/c3_pytest_cpl.log | cut -c1-300 | head -20
timeout 300 python bench_configs.py --only C5 --iters 20 > gpurun_out/c3_c5.log 2>&1; tail -n 5 gpurun_out/c3_c5.log | cut -c1-250
cap() {
  timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$2" -s 3 -c 1 -o gpurun_out/r2_$1 python bench_configs.py --only $3 --iters 2 > gpurun_out/r2_$1.log 2>&1
  echo "$1 rc=$?"
}
cap rqs_fwd 'rqs_unrolled_kernel.*bool.0' C4
cap rqs_inv 'rqs_unrolled_kernel.*bool.1' C4
cap planar_inv 'planar_dev_kernel.*int.1, .bool.0' C2
cap planar_logpdf 'planar_dev_kernel.*int.1, .bool.1' C2
cap coupling_tc 'coupling_tc_kernel' C5
