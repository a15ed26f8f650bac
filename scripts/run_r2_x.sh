#!/bin/bash
# round 2, call X: what the driver runs at round end, on the end-of-round build -- GPU suite, smoke(), default bench line
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/x_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/x_pytest.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/x_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/x_smoke.log
timeout 900 python bench.py > gpurun_out/x_bench_default.json 2> gpurun_out/x_bench_default.err; echo "bench rc=$?"; cut -c1-400 gpurun_out/x_bench_default.json
