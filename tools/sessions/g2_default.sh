#!/usr/bin/env bash
set -u
N=${1:-2}
mkdir -p gpurun_out
cd "$(dirname "$0")/../.."
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 500 $TR --master-port 29771 bench.py --gpus $N --steps 30 --warmup 5 > gpurun_out/g${N}_default_c.json 2> gpurun_out/g${N}_default_c.err; echo "rc=$?"
timeout 300 $TR --master-port 29772 bench.py --impl reference --gpus $N --steps 3 --warmup 1 > gpurun_out/g${N}_ref.json 2> gpurun_out/g${N}_ref.err; echo "ref rc=$?"
python - <<PY
import json
d = json.load(open("gpurun_out/g${N}_default_c.json"))
c = d.get("c4") or {}
print("replica p50", d["ms_p50"], "mean", d["ms_per_step"], "value", round(d["value"] / 1e6), "M | c4 p50", c.get("ms_p50"), "mean", c.get("ms_per_step"), "ratio", c.get("sharded_over_unsharded_p50"), c.get("order"), c.get("error"))
try:
    r = json.load(open("gpurun_out/g${N}_ref.json"))
    print("reference arm:", r.get("value"), r.get("unit"), r.get("n_gpus"), r.get("cpu_baseline", {}).get("cores"))
except Exception as e:
    print("ref unreadable", e)
PY
