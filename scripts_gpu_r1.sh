#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/bench.log
B2B_V1_NW=12 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/bench_nw12.log 2>&1; echo "bench rc=$?" >> gpurun_out/bench_nw12.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:chain_v1 -s 4 -c 1 -o gpurun_out/prof_chain_v1 python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_full.log 2>&1
tail -n 4 gpurun_out/pytest_gpu.log; tail -n 2 gpurun_out/bench.log | cut -c1-1300; tail -n 2 gpurun_out/bench_nw12.log | cut -c1-300
