#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 -k "vjp or rqs_bin or realnvp or autograd or radial" > gpurun_out/c11_pytest.log 2>&1; echo "pytest rc=$?"; grep -E "^FAILED|^ERROR|passed|failed|^E  +assert|AssertionError" gpurun_out/c11_pytest.log | cut -c1-300 | head -40
timeout 300 python tools/bench_vjp.py > gpurun_out/c11_vjp.log 2>&1; cat gpurun_out/c11_vjp.log | tail -12
