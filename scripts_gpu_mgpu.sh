#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus.txt 2>&1
timeout 600 python -m pytest tests/test_multigpu.py -m gpu -q > gpurun_out/pytest_mgpu.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_mgpu.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_2gpu.log 2>&1; echo "rc=$?" >> gpurun_out/bench_2gpu.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench_configs.py --iters 10 --json gpurun_out/configs_2gpu.json > gpurun_out/configs_2gpu.log 2>&1; echo "rc=$?" >> gpurun_out/configs_2gpu.log
timeout 900 python bench_configs.py --iters 20 --json gpurun_out/configs_1gpu.json > gpurun_out/configs_1gpu.log 2>&1; echo "rc=$?" >> gpurun_out/configs_1gpu.log
tail -n 3 gpurun_out/pytest_mgpu.log; tail -n 2 gpurun_out/bench_2gpu.log | cut -c1-600; tail -n 12 gpurun_out/configs_2gpu.log | cut -c1-400; tail -n 12 gpurun_out/configs_1gpu.log | cut -c1-400
