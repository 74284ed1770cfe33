#!/bin/bash
# Round-2 profile session: ncu launch list + full captures of the round's kernels.  Usage: gpu_profiles_r2.sh TAG
TAG=${1:-prof}
O=gpurun_out/r2_$TAG; mkdir -p $O
B="python bench.py --workload metric --steps 1 --warmup 3 --no-cpu-baseline --no-side"
# (1) every launch of the steady state with its device time (skip the fit / warm-up launches)
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 330 -c 130 --csv --log-file $O/launches_metric.csv $B > $O/ncu1.log 2>&1
# (2) full captures: the slice product, the kernel-evaluation pass, the scores
timeout 300 ncu --set full --clock-control none --import-source on -k regex:acq_i8_gemm -s 12 -c 1 -o $O/prof_gemm -f $B > $O/ncu2.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:kstar_slice -s 12 -c 1 -o $O/prof_kstar -f $B > $O/ncu3.log 2>&1
# (3) K4 on tcgen05 and K1 at N = 8192 (D = 16), K1 at the north_star shape
timeout 300 ncu --set full --clock-control none --import-source on -k regex:syrk_i8_kernel -s 9 -c 1 -o $O/prof_syrk_i8_n8192 -f python scripts/fit_once.py 8192 16 > $O/ncu4.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:kmat_kernel -s 2 -c 1 -o $O/prof_kmat_n8192 -f python scripts/fit_once.py 8192 16 > $O/ncu5.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:kmat_kernel -s 2 -c 1 -o $O/prof_kmat_n4096 -f python scripts/fit_once.py 4096 8 > $O/ncu6.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_fit_n4096.csv python scripts/fit_once.py 4096 8 > $O/ncu7.log 2>&1
ls -la $O | head -30
