#!/bin/bash
# round 2, end-of-round build on 8 GPUs: bench at N = 8 and N = 4 (slab mode)
mkdir -p gpurun_out
nvidia-smi -L | wc -l
for n in 8 4; do
  NCCL_DEBUG=WARN timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2958$n bench.py --gpus $n --steps 5 --warmup 3 > gpurun_out/r2_bench_n$n.json 2> gpurun_out/r2_bench_n$n.err
  echo "bench n$n rc=$?"; python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2_bench_n$n.json').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step','dist_check_rel_err','dist_relax_tile_bit_identical','dist_stages_ms')}); print(d['relax']); print({k:v for k,v in d['roofline'].items() if k in ('frac','ms_sr_plus_es','ms_sr','ms_es')}); print({k:v for k,v in d['e2e'].items() if k!='note'}); print(d['compact_tip'])
except Exception as e: print('parse failed', e)
PY
  tail -3 gpurun_out/r2_bench_n$n.err
done
