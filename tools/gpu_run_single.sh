#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 900 python bench_configs.py --iters 20 --json gpurun_out/configs_1gpu.json > gpurun_out/configs_1gpu.log 2>&1; echo "rc=$?" >> gpurun_out/configs_1gpu.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:coupling_tc_kernel -s 2 -c 1 -o gpurun_out/prof_coupling_tc python bench_configs.py --iters 3 --only C5 > gpurun_out/ncu_tc.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches_c5.csv python bench_configs.py --iters 2 --only C5 > gpurun_out/launches_c5.log 2>&1
tail -n 3 gpurun_out/pytest_gpu.log; tail -n 12 gpurun_out/configs_1gpu.log | cut -c1-230
