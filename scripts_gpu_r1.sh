#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/bench.log 2>&1
B2B_V1_CFG=1118 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/bench_1118.log 2>&1
timeout 600 python bench_configs.py --iters 20 --only C2,C3 > gpurun_out/configs_c23.log 2>&1
B2B_V1_CFG=1128 timeout 600 python bench_configs.py --iters 20 --only C3 > gpurun_out/configs_c3_1128.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:chain_v1 -s 4 -c 1 -o gpurun_out/prof_chain_v1 python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_full.log 2>&1
tail -n 3 gpurun_out/pytest_gpu.log; for f in bench bench_1118; do tail -n 1 gpurun_out/$f.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['roofline']['frac'], d['per_layer_launches']['frac'])"; done; tail -n 4 gpurun_out/configs_c23.log | cut -c1-200; tail -n 2 gpurun_out/configs_c3_1128.log | cut -c1-200
