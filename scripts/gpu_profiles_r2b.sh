#!/bin/bash
# Round-2 (second half) evidence: bench line + reference arm, launch list of the look-ahead fit, full captures of the chain kernels.
TAG=${1:-d}
O=gpurun_out/r2_$TAG; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/smi.txt
timeout 900 python bench.py > $O/bench.json 2> $O/bench.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 500 -c 520 --csv --log-file $O/launches_fit_n4096.csv python scripts/fit_once.py 4096 8 > $O/ncu1.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:chol_head_kernel -s 40 -c 1 -o $O/prof_chol_head -f python scripts/fit_once.py 4096 8 > $O/ncu2.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:potrf_diag_kernel -s 40 -c 1 -o $O/prof_potrf -f python scripts/fit_once.py 4096 8 > $O/ncu3.log 2>&1
B200BO_CHOL_TRACE=1 B200BO_CHOL_GRAPH=0 timeout 100 python scripts/fit_once.py 4096 8 > $O/trace_fit_n4096.txt 2>&1
timeout 300 python scripts/chol_sched_ab.py > $O/sched_ab.log 2>&1
ls -la $O
tail -c 600 $O/bench.err
