#!/usr/bin/env bash
# N-GPU: the sharded C4 configuration alone: host loop (default) with the device timeline of one scan per rank, then the
# device-resident loop (device_loop = 1) for comparison.
set -u
N=${1:-8}
mkdir -p gpurun_out
cd "$(dirname "$0")/../.."
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
BENCH_C4_TIMELINE=1 timeout 400 $TR --master-port 29741 bench.py --gpus $N --workload c4 --steps 40 --warmup 5 > gpurun_out/g${N}_c4_host.json 2> gpurun_out/g${N}_c4_host.err; echo "rc=$?"
grep -E "rank [0-9]" gpurun_out/g${N}_c4_host.err | sort -s -k2,2n | head -120 > gpurun_out/g${N}_c4_timeline.txt
head -30 gpurun_out/g${N}_c4_timeline.txt
timeout 400 $TR --master-port 29742 bench.py --gpus $N --workload c4 --steps 40 --warmup 5 --device-loop 1 > gpurun_out/g${N}_c4_dev1.json 2> gpurun_out/g${N}_c4_dev1.err; echo "rc=$?"
python - <<PY
import json
for f in ("host", "dev1"):
    try:
        d = json.load(open("gpurun_out/g${N}_c4_%s.json" % f))
        print(f, "c4 sharded x$N p50", d["ms_p50"], "mean", d["ms_per_step"], "p99", d["ms_p99"], "e2e p50", d["e2e"]["ms_p50"], "launches/scan/rank", d["gpu_launches"] / d["steps"] / d["n_gpus"])
    except Exception as e:
        print(f, "unreadable", e)
PY
