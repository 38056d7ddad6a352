#!/bin/bash
# 2-GPU box visit: the multi-GPU tests, the per-config numbers and the headline bench at N = 2
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_multigpu.py -m gpu -q --timeout 500 > gpurun_out/pytest_2gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_2gpu.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench_configs.py --iters 10 --json gpurun_out/configs_2gpu.json > gpurun_out/configs_2gpu.log 2>&1; echo "rc=$?" >> gpurun_out/configs_2gpu.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_2gpu.log 2>&1
tail -n 3 gpurun_out/pytest_2gpu.log; tail -n 4 gpurun_out/configs_2gpu.log | cut -c1-160; tail -n 1 gpurun_out/bench_2gpu.log | cut -c1-400
