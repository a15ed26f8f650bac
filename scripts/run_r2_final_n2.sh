#!/bin/bash
# round 2, end-of-round build on 2 GPUs: the two-process tests, bench N = 2 (slab mode), sharded sweep16
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "two_processes or two_ranks" > gpurun_out/r2_pytest_2gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_pytest_2gpu.log; tail -4 gpurun_out/r2_pytest_2gpu.log
NCCL_DEBUG=WARN timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err
echo "bench n2 rc=$?"; python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2_bench_n2.json').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step','dist_check_rel_err','dist_relax_tile_bit_identical','dist_stages_ms')}); print(d['relax']); print({k:v for k,v in d['roofline'].items() if k in ('frac','ms_sr_plus_es','ms_sr','ms_es')}); print({k:v for k,v in d['e2e'].items() if k!='note'})
except Exception as e: print('parse failed', e)
PY
tail -3 gpurun_out/r2_bench_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus 2 --workload sweep16 --steps 5 --warmup 3 > gpurun_out/r2_bench_sweep16_n2.json 2> gpurun_out/r2_bench_sweep16_n2.err; echo "sweep n2 rc=$?"; cut -c1-500 gpurun_out/r2_bench_sweep16_n2.json
