#!/bin/bash
# ncu launch list + one full capture of a kernel.  Usage: gpu_prof.sh TAG WORKLOAD KERNEL_REGEX [skip]
TAG=$1; W=$2; KR=$3; SKIP=${4:-8}
O=gpurun_out/r2_$TAG; mkdir -p $O
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches_$W.csv python bench.py --workload $W --steps 1 --warmup 3 --no-cpu-baseline --no-side > $O/ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$KR -s $SKIP -c 1 -o $O/prof_$W -f python bench.py --workload $W --steps 1 --warmup 3 --no-cpu-baseline --no-side > $O/ncu_full.log 2>&1
ls -la $O
