#!/bin/bash
O=gpurun_out/r2_q; mkdir -p $O
timeout 300 python -m pytest tests -q -m gpu -x -k "trailing_update or factor_many or golden or engines_agree or config3" > $O/pytest.log 2>&1; tail -3 $O/pytest.log
timeout 100 python scripts/fit_once.py 4096 8 | tail -1
timeout 100 python scripts/fit_once.py 8192 16 | tail -1
timeout 200 python bench.py --workload metric --no-cpu-baseline --no-side | cut -c100-260
timeout 300 ncu --set full --clock-control none --import-source on -k regex:syrk_i8_kernel -s 9 -c 1 -o $O/prof_syrk_i8_n8192 -f python scripts/fit_once.py 8192 16 > $O/ncu4.log 2>&1
timeout 300 ncu --set full --clock-control none --cache-control none --import-source on -k regex:acq_i8_gemm -s 12 -c 1 -o $O/prof_gemm_warm -f python bench.py --workload metric --steps 1 --warmup 3 --no-cpu-baseline --no-side > $O/ncu2.log 2>&1
ls $O
