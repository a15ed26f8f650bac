#!/bin/bash
# round 2, call Y (2 GPUs): PeerAssembler (copy-engine assembly on rank 0) in the bench step against the NCCL gathers (--no-peer)
mkdir -p gpurun_out
for mode in asm nccl; do
  extra=""; [ "$mode" = "nccl" ] && extra="--no-peer"
  NCCL_DEBUG=WARN timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2956$((RANDOM % 9)) bench.py --gpus 2 --steps 5 --warmup 3 --no-compact --no-cufft --no-cpu-baseline $extra > gpurun_out/y_bench_n2_$mode.json 2> gpurun_out/y_bench_n2_$mode.err
  echo "$mode rc=$?"; tail -4 gpurun_out/y_bench_n2_$mode.err | cut -c1-300
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/y_bench_n2_$mode.json').read().strip().splitlines()[-1])
    print(d['ms_per_step'], d['dist_check_rel_err'], d['dist_relax_tile_bit_identical'], d['relax']['in_step_stages_ms_max_over_ranks'], {k:v for k,v in d['e2e'].items() if k!='note'} if d.get('e2e') else None)
except Exception as e: print('parse failed', e)
PY
done
