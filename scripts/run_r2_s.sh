#!/bin/bash
# round 2, call S: dynamic relax schedule -- parity test, then timing against the static launch and over chunk lengths
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "dynamic_schedule or footprint or x_slab or twin or relax" > gpurun_out/s_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/s_pytest.log
: > gpurun_out/s_relax.log
for c in 0 6 3 4 8 12 2; do
  echo "== chunk $c" >> gpurun_out/s_relax.log
  DBSPM_RELAX_CHUNK=$c timeout 300 python scripts/time_relax2.py >> gpurun_out/s_relax.log 2>&1
done
grep "RELAX\|==" gpurun_out/s_relax.log
