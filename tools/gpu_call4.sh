#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/c4_pytest.log 2>&1; echo "pytest rc=$?"; grep -E "^FAILED|^ERROR|passed|failed" gpurun_out/c4_pytest.log | cut -c1-250 | head -40
timeout 600 python bench_configs.py --iters 20 --json gpurun_out/c4_configs.json > gpurun_out/c4_configs.log 2>&1; cut -c1-200 gpurun_out/c4_configs.log | tail -n 14
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/c4_bench.log 2>&1; echo "bench rc=$?"; tail -n 2 gpurun_out/c4_bench.log | cut -c1-3000
