#!/bin/bash
# Final evidence of round 2: smoke, the whole GPU suite, the bench line and its reference arm.  Usage: gpu_final_r2.sh TAG
TAG=${1:-h}
O=gpurun_out/r2_$TAG; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/smi.txt
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; tail -1 $O/smoke.log
timeout 900 python -m pytest tests -m gpu -q --durations=5 > $O/pytest_gpu.log 2>&1; tail -9 $O/pytest_gpu.log
timeout 900 python bench.py > $O/bench.json 2> $O/bench.err; tail -c 300 $O/bench.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err
python - <<PY
import json
d = json.load(open("$O/bench.json"))
print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches")}, "e2e", d["e2e"]["value"], "frac", d["roofline"]["frac"], "traffic", d["roofline"]["traffic"], d["clocks"])
s = d["side_metrics"]
print("fit4096", s["fit_n4096_d8"]); print("cfg4", s["cfg4_map_sweep"]); print("direct", s["direct_l_search"]); print("joint", s["joint_posterior_sample"])
for k in ("workload_cfg3", "workload_cfg5", "workload_cfg2"):
    print(k, s[k]["value"], s[k]["e2e"]["value"], s[k]["roofline"]["frac"])
print("target", s["target_n4096_d8_m262144"]["weak"]["value"], s["target_n4096_d8_m262144"]["weak"]["e2e"]["value"])
print("ref", json.load(open("$O/bench_ref.json"))["value"])
PY
