#!/bin/bash
# One GPU session: tests, smoke, bench (+reference arm), ncu launch list and a full capture of the top kernel.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/smi.txt; nproc >> gpurun_out/smi.txt; lscpu | grep "Model name" >> gpurun_out/smi.txt
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 120 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/bench.err
if [ "$1" == "ncu" ]; then
  rm -f gpurun_out/*.ncu-rep
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:acq_fused -s 3 -c 1 -o gpurun_out/prof_acq -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_acq.log 2>&1
fi
tail -3 gpurun_out/pytest_gpu.log; cat gpurun_out/smoke.log; cat gpurun_out/bench.json
