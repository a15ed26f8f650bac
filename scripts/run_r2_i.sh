#!/bin/bash
# round 2, call I (2 GPUs): sweep16 graph vs eager at N=1, sharded sweep at N=2, sweep graph test, a few A/B timings
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "sweep or direct" > gpurun_out/i_pytest.log 2>&1; tail -5 gpurun_out/i_pytest.log
timeout 600 python bench.py --workload sweep16 --steps 5 --warmup 3 > gpurun_out/i_sweep_n1.json 2> gpurun_out/i_sweep_n1.err; echo "sweep n1 rc=$?"; cut -c1-1800 gpurun_out/i_sweep_n1.json; tail -3 gpurun_out/i_sweep_n1.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus 2 --workload sweep16 --steps 5 --warmup 3 > gpurun_out/i_sweep_n2.json 2> gpurun_out/i_sweep_n2.err; echo "sweep n2 rc=$?"; cut -c1-1200 gpurun_out/i_sweep_n2.json; tail -3 gpurun_out/i_sweep_n2.err
: > gpurun_out/i_corr.log
CORR_CUFFT=0 CORR_FLAGS=0,4,8 python scripts/time_corr2.py pyridine >> gpurun_out/i_corr.log 2>&1
CORR_CUFFT=0 CORR_FLAGS=0 DBSPM_B200_LIB=$PWD/variants/libdbspm_b200_pow1.so python scripts/time_corr2.py large_slab >> gpurun_out/i_corr.log 2>&1
CORR_CUFFT=0 CORR_FLAGS=0 python scripts/time_corr2.py large_slab >> gpurun_out/i_corr.log 2>&1
grep "CORR" gpurun_out/i_corr.log
python scripts/time_direct.py pyridine large_slab > gpurun_out/i_direct.log 2>&1; grep DIRECT gpurun_out/i_direct.log | cut -c1-200
