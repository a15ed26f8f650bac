#!/bin/bash
# round 2, call T (2 GPUs): sweep tests, sharded sweep graph at N = 2, bench N = 2 with the dynamic relax schedule
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "sweep or dynamic" > gpurun_out/t_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/t_pytest.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus 2 --workload sweep16 --steps 5 --warmup 3 > gpurun_out/r2_bench_sweep16_n2.json 2> gpurun_out/r2_bench_sweep16_n2.err; echo "sweep n2 rc=$?"; cut -c1-900 gpurun_out/r2_bench_sweep16_n2.json; tail -3 gpurun_out/r2_bench_sweep16_n2.err
timeout 600 python bench.py --workload sweep16 --steps 3 --warmup 3 > gpurun_out/r2_bench_sweep16_n1.json 2> gpurun_out/r2_bench_sweep16_n1.err; echo "sweep n1 rc=$?"; cut -c1-700 gpurun_out/r2_bench_sweep16_n1.json
NCCL_DEBUG=WARN timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 2 --steps 5 --warmup 3 --no-e2e --no-compact > gpurun_out/t_bench_n2.json 2> gpurun_out/t_bench_n2.err
echo "bench n2 rc=$?"; cut -c1-700 gpurun_out/t_bench_n2.json
