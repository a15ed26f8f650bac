#!/bin/bash
# round 2, call O: lean pipelined axis pass (cp.async slots + registers) A/B; relax footprint GPU test
mkdir -p gpurun_out
timeout 600 python scripts/time_corr3.py large_slab > gpurun_out/o_corr.log 2>&1
grep CORR3 gpurun_out/o_corr.log
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "footprint or variants or x_slab" > gpurun_out/o_pytest.log 2>&1; tail -3 gpurun_out/o_pytest.log
