#!/bin/bash
# one GPU box visit: tests, bench, ncu launch list + full capture of the headline kernel (numbers under ncu are never bench values)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print(\"smoke ok\")" > gpurun_out/smoke.log 2>&1; tail -n 1 gpurun_out/smoke.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/bench.log
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_reference.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv python bench.py --steps 5 --warmup 3 --profile > gpurun_out/bench_under_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:planar_dev -s 4 -c 1 -o gpurun_out/prof_planar_dev python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_full.log 2>&1
tail -n 3 gpurun_out/pytest_gpu.log; tail -n 2 gpurun_out/bench.log | cut -c1-1800; tail -n 1 gpurun_out/bench_reference.log | cut -c1-300
