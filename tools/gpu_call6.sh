#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 600 -k "rqs or config4 or fused_chain" > gpurun_out/c6_pytest.log 2>&1; echo "pytest rc=$?"; grep -E "^FAILED|^ERROR|passed|failed" gpurun_out/c6_pytest.log | cut -c1-250 | head
timeout 300 python bench_configs.py --only C4 --iters 20 > gpurun_out/c6_c4.log 2>&1; tail -n 2 gpurun_out/c6_c4.log | cut -c1-200
timeout 300 python tools/tune_e2e.py > gpurun_out/c6_tune_e2e.log 2>&1; cat gpurun_out/c6_tune_e2e.log | tail -20
