#!/bin/bash
# round 2, call U: relax parity tests + timing after the fx2 reuse; sweep16 N=1 with the static fallback
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q -k "relax or cli or sweep" > gpurun_out/u_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/u_pytest.log
python scripts/time_relax2.py > gpurun_out/u_relax.log 2>&1; grep RELAX gpurun_out/u_relax.log
timeout 600 python bench.py --workload sweep16 --steps 3 --warmup 3 > gpurun_out/r2_bench_sweep16_n1.json 2> gpurun_out/r2_bench_sweep16_n1.err; echo "sweep n1 rc=$?"; cut -c1-600 gpurun_out/r2_bench_sweep16_n1.json
