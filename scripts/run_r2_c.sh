#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/c_relax.log
python scripts/time_relax2.py >> gpurun_out/c_relax.log 2>&1
for v in pf0 pf2 t32mb16 t32mb18 t32mb20 t32mb16pf0 t64mb9 t32mb16d1 t32r112 t32r104 t32r120 t32r144; do
  DBSPM_B200_LIB=$PWD/variants/libdbspm_b200_$v.so timeout 300 python scripts/time_relax2.py >> gpurun_out/c_relax.log 2>&1
done
grep RELAX gpurun_out/c_relax.log | grep bench
timeout 900 python -m pytest tests -m gpu -x -q -k "relax or cli or sweep" > gpurun_out/c_pytest.log 2>&1; tail -3 gpurun_out/c_pytest.log
