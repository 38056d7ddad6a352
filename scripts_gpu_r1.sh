#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 600 -k "tensor_core or realnvp or coupling" > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 900 python bench_configs.py --iters 20 --only C5 --json gpurun_out/configs_c5.json > gpurun_out/configs_c5.log 2>&1; echo "rc=$?" >> gpurun_out/configs_c5.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:coupling_tc_kernel -s 2 -c 1 -o gpurun_out/prof_coupling_tc python bench_configs.py --iters 3 --only C5 > gpurun_out/ncu_tc.log 2>&1
tail -n 4 gpurun_out/pytest_gpu.log; tail -n 6 gpurun_out/configs_c5.log | cut -c1-300
