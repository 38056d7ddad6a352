#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --timeout 600 -k "coupling_and_batchnorm_vjp or realnvp" > gpurun_out/c12_pytest.log 2>&1; echo "pytest rc=$?"; grep -E "^FAILED|^ERROR|passed|failed" gpurun_out/c12_pytest.log | cut -c1-300 | head
timeout 300 python tools/bench_vjp.py > gpurun_out/c12_vjp.log 2>&1; tail -3 gpurun_out/c12_vjp.log
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:coupling_vjp_fast_kernel" -s 2 -c 1 -o gpurun_out/r2_coupling_vjp python tools/bench_vjp.py 128 8 > gpurun_out/r2_coupling_vjp.log 2>&1
(python tools/summarise_ncu.py gpurun_out/r2_coupling_vjp.ncu-rep; python tools/ncu_hot.py gpurun_out/r2_coupling_vjp.ncu-rep --top 24) > gpurun_out/r2_coupling_vjp_ncu_full.txt 2>&1
rm -f gpurun_out/*.ncu-rep
