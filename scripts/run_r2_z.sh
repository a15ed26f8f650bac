#!/bin/bash
# round 2, call Z: wide pipelined y pass as the default -- parity, A/B, the N = 1 bench line
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "pipelined or variants or goldens or large_slab or config_sizes or edge or tiny or errors" > gpurun_out/z_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/z_pytest.log
timeout 300 python scripts/time_corr3.py large_slab 2>&1 | grep CORR3
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; echo "bench rc=$?"; cut -c1-200 gpurun_out/r2_bench_n1.json
