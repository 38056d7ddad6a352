#!/bin/bash
# Round-2 profiling visit (1 GPU): ncu --set full of every shipped hot kernel that had no capture, plus the topology the
# e2e NUMA binding needs.  Numbers printed under ncu are never bench values.
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
(lscpu | head -40; numactl -H 2>/dev/null; for d in /sys/bus/pci/devices/*; do if [ -f $d/numa_node ] && grep -q 0x10de $d/vendor 2>/dev/null; then echo "$d $(cat $d/numa_node) $(cat $d/local_cpulist)"; fi; done; nvidia-smi --query-gpu=index,pci.bus_id --format=csv) > gpurun_out/numa.txt 2>&1
cap() {  # name regex only
  timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$2" -s 3 -c 1 -o gpurun_out/r2_$1 python bench_configs.py --only $3 --iters 2 > gpurun_out/r2_$1.log 2>&1
  echo "$1 rc=$?"
}
cap rqs_fwd 'rqs_unrolled_kernel.*false' C4
cap rqs_inv 'rqs_unrolled_kernel.*true' C4
cap planar_inv 'planar_dev_kernel.*, 1, false' C2
cap planar_logpdf 'planar_dev_kernel.*, 1, true' C2
cap coupling_tc 'coupling_tc_kernel' C5
cap radial_fwd 'chain_v1_kernel' C3
cap radial_inv 'radial_unrolled' C3
timeout 900 python bench_configs.py --iters 10 --json gpurun_out/r2_configs_1gpu_base.json > gpurun_out/r2_configs_1gpu_base.log 2>&1
tail -n 14 gpurun_out/r2_configs_1gpu_base.log | cut -c1-220
