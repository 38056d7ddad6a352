#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 900 python bench_configs.py --iters 20 --json gpurun_out/configs_1gpu.json > gpurun_out/configs_1gpu.log 2>&1; echo "rc=$?" >> gpurun_out/configs_1gpu.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:chain_v1 -s 4 -c 1 -o gpurun_out/prof_chain_v1 python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_full.log 2>&1
tail -n 4 gpurun_out/pytest_gpu.log; tail -n 12 gpurun_out/configs_1gpu.log | cut -c1-260
