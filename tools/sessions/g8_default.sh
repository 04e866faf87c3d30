#!/usr/bin/env bash
# N-GPU: the default bench alone (replicas + the C4 leg), as the driver runs it
set -u
N=${1:-8}
mkdir -p gpurun_out
cd "$(dirname "$0")/../.."
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29761 bench.py --gpus $N --steps 40 --warmup 5 > gpurun_out/g${N}_default_b.json 2> gpurun_out/g${N}_default_b.err; echo "rc=$?"
python - <<PY
import json
d = json.load(open("gpurun_out/g${N}_default_b.json"))
c = d.get("c4") or {}
print("replica p50", d["ms_p50"], "mean", d["ms_per_step"], "value", round(d["value"] / 1e6), "M | c4 p50", c.get("ms_p50"), "mean", c.get("ms_per_step"), "ratio", c.get("sharded_over_unsharded_p50"), "slow", d["detail"].get("slow_steps"))
PY
