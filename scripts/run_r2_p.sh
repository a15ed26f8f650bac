#!/bin/bash
# round 2, call P: full GPU suite + default bench line + A/B of the default axis-pass choice
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/p_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/p_pytest.log
timeout 300 python scripts/time_corr3.py large_slab > gpurun_out/p_corr.log 2>&1; grep CORR3 gpurun_out/p_corr.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/p_bench_n1.json 2> gpurun_out/p_bench_n1.err; echo "bench rc=$?"; cut -c1-900 gpurun_out/p_bench_n1.json
