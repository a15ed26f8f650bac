#!/bin/bash
# round 2, call A: GPU tests, then relax variants (A/B), then a short bench
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/a_gpu.txt 2>&1
free -g >> gpurun_out/a_gpu.txt; nproc >> gpurun_out/a_gpu.txt
timeout 1500 python -m pytest tests -m gpu -x -q --durations=15 > gpurun_out/a_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/a_pytest.log
tail -30 gpurun_out/a_pytest.log
: > gpurun_out/a_relax.log
for L in -1 1; do RELAX_LAYOUT=$L timeout 300 python scripts/time_relax2.py >> gpurun_out/a_relax.log 2>&1; done
for v in d0mb3 d0mb5 d0mb6 d1mb4 d2mb4 dinfmb4 dinfmb5 dinfmb3 dinf_t64mb8 d0_t64mb8 d0_t32mb16 dinf_t32mb16; do
  DBSPM_B200_LIB=$PWD/variants/libdbspm_b200_$v.so timeout 300 python scripts/time_relax2.py >> gpurun_out/a_relax.log 2>&1
done
grep RELAX gpurun_out/a_relax.log
