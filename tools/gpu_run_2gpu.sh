#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 600 python bench_configs.py --iters 20 --only C2 > gpurun_out/configs_c2.log 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench_configs.py --iters 10 --json gpurun_out/configs_2gpu.json > gpurun_out/configs_2gpu.log 2>&1; echo "rc=$?" >> gpurun_out/configs_2gpu.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_2gpu.log 2>&1
tail -n 5 gpurun_out/pytest_gpu.log; tail -n 2 gpurun_out/configs_c2.log | cut -c1-200; tail -n 5 gpurun_out/configs_2gpu.log | cut -c1-200
