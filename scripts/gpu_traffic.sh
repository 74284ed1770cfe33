#!/bin/bash
# DRAM bytes per launch of the acquisition-step kernels, one workload each (ncu, two metrics, default cache control).  Usage: gpu_traffic.sh TAG
TAG=${1:-t}
O=gpurun_out/r2_$TAG; mkdir -p $O
for W in metric cfg2 cfg3 cfg5; do
  C=200; [ $W = cfg3 ] && C=700; [ $W = cfg5 ] && C=360
  timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"acq_i8_gemm|kstar_slice|acq_finish|acq_grad|argmax_blocks" -c $C --csv --log-file $O/traffic_$W.csv python bench.py --workload $W --steps 1 --warmup 3 --no-cpu-baseline --no-side > $O/traffic_$W.log 2>&1
  tail -1 $O/traffic_$W.log | cut -c1-200
done
ls -la $O
