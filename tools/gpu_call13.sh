#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 -k "planar" > gpurun_out/c13_pytest.log 2>&1; echo "pytest rc=$?"; grep -E "^FAILED|^ERROR|passed|failed|^E  +assert|AssertionError" gpurun_out/c13_pytest.log | cut -c1-300 | head -40
timeout 300 python bench_configs.py --only C2 --iters 10 > gpurun_out/c13_c2.log 2>&1; tail -15 gpurun_out/c13_c2.log
