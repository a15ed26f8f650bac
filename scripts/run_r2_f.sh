#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/f_corr.log
CORR_FLAGS=0,8192 python scripts/time_corr2.py large_slab pyridine tma_cu111 >> gpurun_out/f_corr.log 2>&1
for f in variants/libdbspm_b200_z_*.so; do
  CORR_CUFFT=0 CORR_FLAGS=0,8192 DBSPM_B200_LIB=$PWD/$f python scripts/time_corr2.py large_slab >> gpurun_out/f_corr.log 2>&1
done
grep "CORR\|failed" gpurun_out/f_corr.log
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "overlap or compact or pruned" > gpurun_out/f_pytest.log 2>&1; tail -3 gpurun_out/f_pytest.log
