#!/bin/bash
# Round-2 GPU session: tests, smoke, bench lines of several workloads.  Usage: gpu_session_r2.sh [tag] [workloads...]
TAG=${1:-a}; shift
WL=${@:-metric}
O=gpurun_out/r2_$TAG; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > $O/smi.txt; nproc >> $O/smi.txt; lscpu | grep "Model name" >> $O/smi.txt
timeout 1500 python -m pytest tests -q -m gpu --durations=8 > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
timeout 200 python __graft_entry__.py smoke > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/smoke.log
for w in $WL; do
  timeout 600 python bench.py --workload $w --no-cpu-baseline --no-side > $O/bench_$w.json 2> $O/bench_$w.err; echo "bench $w rc=$?" >> $O/bench_$w.err
done
tail -5 $O/pytest_gpu.log; cat $O/smoke.log; for w in $WL; do cut -c1-600 $O/bench_$w.json; done
