#!/usr/bin/env bash
# 2-GPU: the default bench (replicas + C4 leg) with and without per-rank core pinning
set -u
N=${1:-2}
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
nproc
timeout 500 $TR --master-port 29751 bench.py --gpus $N --steps 30 --warmup 5 > gpurun_out/g${N}_pin.json 2> gpurun_out/g${N}_pin.err; echo "rc=$?"
BENCH_NO_PIN=1 timeout 500 $TR --master-port 29752 bench.py --gpus $N --steps 30 --warmup 5 > gpurun_out/g${N}_nopin.json 2> gpurun_out/g${N}_nopin.err; echo "rc=$?"
python - <<PY
import json
for f in ("pin", "nopin"):
    try:
        d = json.load(open("gpurun_out/g${N}_%s.json" % f))
        c = d.get("c4") or {}
        print(f, "replica p50", d["ms_p50"], "mean", d["ms_per_step"], "| c4 p50", c.get("ms_p50"), "mean", c.get("ms_per_step"), "ratio", c.get("sharded_over_unsharded_p50"), "cores", d["detail"]["host_cores_of_this_rank"])
    except Exception as e:
        print(f, "unreadable", e)
PY
