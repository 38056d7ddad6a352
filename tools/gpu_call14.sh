#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -k "spline_layer or more_than_two or batchnorm or rqs_vjp" > gpurun_out/c14_pytest.log 2>&1; echo "pytest rc=$?"; grep -E "^FAILED|^ERROR|passed|failed|^E  +assert|AssertionError|Error" gpurun_out/c14_pytest.log | cut -c1-300 | head -30
timeout 300 python tools/bench_vjp.py > gpurun_out/c14_vjp.log 2>&1; tail -3 gpurun_out/c14_vjp.log
