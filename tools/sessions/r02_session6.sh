#!/usr/bin/env bash
# Round-2 single-GPU session 6: where do the rare ~1 ms steps come from (105-scan A/B with the slowest scans printed), the other
# single-GPU workloads on the final code (C1, C3, C4 unsharded).
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/../.."
T=${TAG:-s6}
echo "== 1. A/B over 105 scans: default | r = no reuse | i = synchronous insert"
AB_SCANS=105 timeout 500 python tools/ab_latency.py 0 0r 0i 2>&1 | tee gpurun_out/${T}_ab_latency.log | tail -24
echo "== 2. C1, C3, C4 (unsharded)"
for w in c1 c3 c4; do
  timeout 400 python bench.py --workload $w --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/${T}_bench_$w.json 2> gpurun_out/${T}_bench_$w.err; echo "$w rc=$?"
done
python - <<PY
import json
for w in ("c1", "c3", "c4"):
    try:
        d = json.load(open("gpurun_out/${T}_bench_%s.json" % w))
        print(w, "p50", d.get("ms_p50"), "mean", d.get("ms_per_step"), "p99", d.get("ms_p99"), "pts/s", round(d["value"] / 1e6, 1), "e2e p50", d["e2e"].get("ms_p50"), "slow", (d.get("detail") or {}).get("slow_steps"))
    except Exception as e:
        print(w, "unreadable", e)
PY
