#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/../.."
timeout 400 python bench.py > gpurun_out/fin2_bench_c2.json 2> gpurun_out/fin2_bench_c2.err; echo "bench rc=$?"
python - <<PY
import json
d = json.load(open("gpurun_out/fin2_bench_c2.json"))
print("C2 p50", d.get("ms_p50"), "mean", d.get("ms_per_step"), "p99", d.get("ms_p99"), "max", d["detail"].get("ms_max"), "e2e p50", d["e2e"].get("ms_p50"), "mean", d["e2e"].get("ms_per_step"), "serial p50", d["e2e"]["serial"].get("ms_p50"), "parity ok", d.get("parity", {}).get("ok"), "slow", d["detail"].get("slow_steps"), "value", d["value"])
PY
