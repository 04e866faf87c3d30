#!/usr/bin/env bash
# quick single-GPU check: GPU suite + a short headline bench
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/../.."
T=${TAG:-q}
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/${T}_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${T}_pytest_gpu.log
tail -4 gpurun_out/${T}_pytest_gpu.log
AB_SCANS=60 timeout 300 python tools/ab_latency.py 0 2>&1 | tee gpurun_out/${T}_ab_latency.log | tail -6
timeout 300 python bench.py --steps 40 --warmup 5 --no-replay > gpurun_out/${T}_bench_c2.json 2> gpurun_out/${T}_bench_c2.err; echo "bench rc=$?"
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${T}_bench_c2.json"))
    print("C2 p50", d.get("ms_p50"), "mean", d.get("ms_per_step"), "p99", d.get("ms_p99"), "e2e p50", d["e2e"].get("ms_p50"), "serial p50", d["e2e"]["serial"].get("ms_p50"), "parity ok", d.get("parity", {}).get("ok"), "host p50", d["detail"]["host_ms_p50"], "slow", d["detail"].get("slow_steps"))
except Exception as e:
    print("bench line unreadable:", e)
PY
