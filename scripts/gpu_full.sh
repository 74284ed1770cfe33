#!/bin/bash
# what the driver runs at round end: the GPU tests, smoke, the reference arm and the default bench line.  Usage: gpu_full.sh TAG
TAG=${1:-full}
O=gpurun_out/r2_$TAG; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > $O/smi.txt; nproc >> $O/smi.txt; lscpu | grep "Model name" >> $O/smi.txt
timeout 1500 python -m pytest tests -x -q -m gpu --durations=6 > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
timeout 200 python __graft_entry__.py smoke > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/smoke.log
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err
timeout 900 python bench.py > $O/bench.json 2> $O/bench.err; echo "bench rc=$?" >> $O/bench.err
tail -12 $O/pytest_gpu.log; cat $O/smoke.log; tail -3 $O/bench.err; cut -c1-700 $O/bench.json
