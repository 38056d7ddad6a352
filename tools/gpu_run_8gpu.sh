#!/bin/bash
# 8-GPU box visit: the multi-GPU tests and the bench (headline + every BASELINE config, NCCL sum in the timed region) at N = 8, 4, 2
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_multigpu.py -m gpu -q --timeout 500 > gpurun_out/pytest_mgpu.log 2>&1; echo "pytest rc=$?"; tail -n 2 gpurun_out/pytest_mgpu.log
for n in 8 4 2; do
NCCL_DEBUG=INFO NCCL_DEBUG_FILE=gpurun_out/nccl.n$n.%p.log timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/r02_bench_${n}gpu.log 2>&1; echo "bench $n rc=$?"
tail -n 1 gpurun_out/r02_bench_${n}gpu.log | python -c "
import sys, json
l = json.loads(sys.stdin.readline())
print('value', '%.4e' % l['value'], 'e2e', '%.4e' % l['e2e']['value'], 'dram', round(l['e2e']['host_dram_gbs'], 1), 'numa', l['config']['numa'], 'dev0', l['config']['cuda_device_of_rank0'])
for k, v in l['configs'].items(): print(k, 'ms', round(v['ms'], 4), 'samples/s', '%.3e' % v['samples_per_s'], 'frac', round(v['frac'], 3), {a: float('%.2g' % b) for a, b in v['oracle_check'].items()})
"
done
rm -f gpurun_out/nccl.n*.log
