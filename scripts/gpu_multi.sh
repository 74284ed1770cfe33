#!/bin/bash
# multi-GPU check: the in-library sharding test + the torchrun bench at N ranks.  Usage: gpu_multi.sh TAG NGPUS [workload]
TAG=$1; NG=$2; W=${3:-metric}
O=gpurun_out/r2_$TAG; mkdir -p $O
nvidia-smi -L > $O/smi.txt
timeout 600 python -m pytest tests -q -m gpu -x -k "multi_gpu" > $O/pytest_multi.log 2>&1; echo "pytest rc=$?" >> $O/pytest_multi.log
tail -5 $O/pytest_multi.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $NG --workload $W --steps 10 --warmup 3 > $O/bench_${W}_${NG}gpu.json 2> $O/bench_${W}_${NG}gpu.err; echo "bench rc=$?" >> $O/bench_${W}_${NG}gpu.err
tail -4 $O/bench_${W}_${NG}gpu.err; cut -c1-900 $O/bench_${W}_${NG}gpu.json
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $NG --workload $W --steps 3 --warmup 1 > $O/bench_ref_${NG}gpu.json 2> $O/bench_ref_${NG}gpu.err
cut -c1-300 $O/bench_ref_${NG}gpu.json
