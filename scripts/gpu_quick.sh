#!/bin/bash
# quick GPU check of a subset of tests + optional bench lines.  Usage: gpu_quick.sh TAG "pytest -k expr" [workloads...]
TAG=$1; K=$2; shift; shift
O=gpurun_out/r2_$TAG; mkdir -p $O
timeout 900 python -m pytest tests -q -m gpu -x -k "$K" > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
tail -25 $O/pytest.log
for w in "$@"; do
  timeout 600 python bench.py --workload $w --no-cpu-baseline --no-side > $O/bench_$w.json 2> $O/bench_$w.err; echo "bench $w rc=$?" >> $O/bench_$w.err
  tail -3 $O/bench_$w.err; cut -c1-400 $O/bench_$w.json
done
