#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/c10_bench.log 2>&1; echo "bench rc=$?"; tail -n 3 gpurun_out/c10_bench.log | cut -c1-1500
tail -n 1 gpurun_out/c10_bench.log | python -c "
import sys, json
l = json.loads(sys.stdin.readline())
print('value', l['value'], 'e2e', l['e2e'], 'numa', l['config']['numa'])
for k, v in l['configs'].items(): print(k, 'ms', round(v['ms'], 4), 'samples/s', '%.3e' % v['samples_per_s'], 'frac', round(v['frac'], 3), v['oracle_check'])
print(l['cpu_baseline'])
"
