#!/bin/bash
# 2-GPU box visit: the multi-GPU tests and the bench (headline + every BASELINE config, NCCL sum in the timed region) at N = 2
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_multigpu.py -m gpu -q --timeout 500 > gpurun_out/pytest_2gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_2gpu.log
NCCL_DEBUG=INFO NCCL_DEBUG_FILE=gpurun_out/nccl.n2.%p.log timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_2gpu.log 2>&1; echo "bench rc=$?"
tail -n 3 gpurun_out/pytest_2gpu.log; tail -n 1 gpurun_out/bench_2gpu.log | cut -c1-6000; grep -h "Init COMPLETE" gpurun_out/nccl.n2.*.log | cut -c1-200 | head -6
