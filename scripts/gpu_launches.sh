#!/bin/bash
# ncu launch list only.  Usage: gpu_launches.sh TAG WORKLOAD [count] [skip]
TAG=$1; W=$2; C=${3:-700}; S=${4:-0}
O=gpurun_out/r2_$TAG; mkdir -p $O
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s $S -c $C --csv --log-file $O/launches_$W.csv python bench.py --workload $W --steps 1 --warmup 3 --no-cpu-baseline --no-side > $O/ncu_launch_$W.log 2>&1
tail -2 $O/ncu_launch_$W.log | cut -c1-200
