#!/bin/bash
# 8-GPU box visit: the bench (headline + every BASELINE config, NCCL sum in the timed region) at N = 8 and N = 4
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo8.txt 2>&1
for n in 8 4; do
NCCL_DEBUG=INFO NCCL_DEBUG_FILE=gpurun_out/nccl.n$n.%p.log timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/bench_${n}gpu.log 2>&1; echo "bench $n rc=$?"
tail -n 1 gpurun_out/bench_${n}gpu.log | python -c "
import sys, json
l = json.loads(sys.stdin.readline())
print('value', l['value'], 'e2e', l['e2e']['value'], 'numa', l['config']['numa'])
for k, v in l['configs'].items(): print(k, 'ms', round(v['ms'], 4), 'samples/s', '%.3e' % v['samples_per_s'], 'frac', round(v['frac'], 3), v['oracle_check'])
"
done
grep -h "Init COMPLETE" gpurun_out/nccl.n8.*.log | cut -c1-160 | head -4
