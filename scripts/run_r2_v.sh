#!/bin/bash
# round 2, call V (8 GPUs): where "the rest" of the N = 8 step goes, dynamic against static relax schedule
mkdir -p gpurun_out
for c in dyn 0; do
  if [ "$c" = "0" ]; then export DBSPM_RELAX_CHUNK=0; else unset DBSPM_RELAX_CHUNK; fi
  NCCL_DEBUG=WARN timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 2959$((RANDOM % 9)) bench.py --gpus 8 --steps 5 --warmup 3 --no-e2e --no-compact --no-cufft --no-cpu-baseline > gpurun_out/v_bench_n8_$c.json 2> gpurun_out/v_bench_n8_$c.err
  echo "chunk=$c rc=$?"; grep "rest of the step" gpurun_out/v_bench_n8_$c.err | sort | head -8
  python - <<PY
import json
d=json.loads(open('gpurun_out/v_bench_n8_$c.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['relax']['ms'], d['relax']['in_step_ms_static_relax_gather'], d['relax']['in_step_stages_ms_max_over_ranks'])
PY
done
