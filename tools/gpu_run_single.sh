#!/bin/bash
# one GPU box visit: the whole GPU test suite, smoke, every BASELINE config, the bench (both arms)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; grep -E "^FAILED|^ERROR|passed|failed" gpurun_out/pytest_gpu.log | cut -c1-250 | head -20
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -n 1 gpurun_out/smoke.log
timeout 600 python bench_configs.py --iters 20 --json gpurun_out/r02_configs_1gpu.json > gpurun_out/r02_configs_1gpu.log 2>&1; cut -c1-170 gpurun_out/r02_configs_1gpu.log | tail -n 12
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_1gpu.log 2>&1; echo "bench rc=$?"
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r02_bench_reference.log 2>&1; tail -n 1 gpurun_out/r02_bench_reference.log | cut -c1-300
tail -n 1 gpurun_out/r02_bench_1gpu.log | python -c "
import sys, json
l = json.loads(sys.stdin.readline())
print('value', l['value'], 'frac', l['roofline']['frac'], 'e2e', l['e2e']['value'], 'cpu', l['cpu_baseline']['value'], l['cpu_baseline']['cores'])
for k, v in l['configs'].items(): print(k, 'ms', round(v['ms'], 4), 'samples/s', '%.3e' % v['samples_per_s'], 'frac', round(v['frac'], 3))
"
