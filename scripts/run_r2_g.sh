#!/bin/bash
# round 2, call G (2 GPUs): the two-process tests, bench N=2 (slab mode), bench N=1
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/g_gpus.txt
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "two_processes or two_ranks or x_slab" > gpurun_out/g_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/g_pytest.log; tail -15 gpurun_out/g_pytest.log
NCCL_DEBUG=WARN timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/g_bench_n2.json 2> gpurun_out/g_bench_n2.err
echo "bench n2 rc=$?"; tail -c 3000 gpurun_out/g_bench_n2.json; tail -5 gpurun_out/g_bench_n2.err
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/g_bench_n1.json 2> gpurun_out/g_bench_n1.err
echo "bench n1 rc=$?"; head -c 2500 gpurun_out/g_bench_n1.json; tail -3 gpurun_out/g_bench_n1.err
