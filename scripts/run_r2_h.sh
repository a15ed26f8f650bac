#!/bin/bash
# round 2, call H: full GPU test suite, direct-path crossover, TMA axis pass after the LDS fix, pow-table A/B
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/h_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/h_pytest.log; tail -14 gpurun_out/h_pytest.log
python scripts/time_direct.py pyridine large_slab > gpurun_out/h_direct.log 2>&1; grep DIRECT gpurun_out/h_direct.log
: > gpurun_out/h_corr.log
CORR_CUFFT=0 CORR_FLAGS=0,4,8 python scripts/time_corr2.py large_slab tma_cu111 triangulene >> gpurun_out/h_corr.log 2>&1
CORR_CUFFT=0 CORR_FLAGS=0 DBSPM_B200_LIB=$PWD/variants/libdbspm_b200_pow1.so python scripts/time_corr2.py large_slab >> gpurun_out/h_corr.log 2>&1
grep "CORR" gpurun_out/h_corr.log
