#!/bin/bash
# 8-GPU check: in-library sharding test at 2/4/8 replicas + torchrun benches (metric weak scaling, cfg5 = 1048576 TS candidates over 8 ranks)
NG=${1:-8}
O=gpurun_out/r2_m$NG; mkdir -p $O
nvidia-smi -L > $O/smi.txt
timeout 300 python -m pytest tests -q -m gpu -x -k "multi_gpu" > $O/pytest_multi.log 2>&1; echo "pytest rc=$?" >> $O/pytest_multi.log
tail -3 $O/pytest_multi.log
for W in metric cfg5; do
  timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $NG --workload $W --steps 5 --warmup 3 > $O/bench_${W}_${NG}gpu.json 2> $O/bench_${W}_${NG}gpu.err; echo "bench $W rc=$?" >> $O/bench_${W}_${NG}gpu.err
  tail -2 $O/bench_${W}_${NG}gpu.err | cut -c1-200; cut -c100-330 $O/bench_${W}_${NG}gpu.json
done
