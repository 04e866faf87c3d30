#!/usr/bin/env bash
# 2-GPU: the sharded C4 leg alone, with the device timeline of one scan per rank on stderr (BENCH_C4_TIMELINE)
set -u
N=${1:-2}
mkdir -p gpurun_out
cd "$(dirname "$0")/../.."
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
BENCH_C4_TIMELINE=1 timeout 400 $TR --master-port 29731 bench.py --gpus $N --workload c4 --steps 60 --warmup 5 > gpurun_out/g${N}_c4_peer_b.json 2> gpurun_out/g${N}_c4_peer_b.err; echo "rc=$?"
grep -E "rank [0-9]" gpurun_out/g${N}_c4_peer_b.err | head -80
python - <<PY
import json
d = json.load(open("gpurun_out/g${N}_c4_peer_b.json"))
print("c4 sharded x$N p50", d["ms_p50"], "mean", d["ms_per_step"], "p99", d["ms_p99"], "e2e p50", d["e2e"]["ms_p50"], "launches/scan/rank", d["gpu_launches"] / d["steps"] / d["n_gpus"])
PY
