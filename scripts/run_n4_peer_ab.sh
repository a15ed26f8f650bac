for F in 0 16; do
DBSPM_DIST_FLAGS=$F timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29507 bench.py --gpus 4 --steps 4 --warmup 3 --no-compact --no-e2e > gpurun_out/bench_n4_i$F.json 2> gpurun_out/bench_n4_i$F.err; grep "dist stage\|symmetric\|Error\|error" gpurun_out/bench_n4_i$F.err | head -5
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_n4_i$F.json").read().strip().splitlines()[-1])
print($F, d["value"], d["roofline"]["ms_sr_plus_es"], d["dist_stages_ms"], d["dist_check_rel_err_es"])
PY
done
