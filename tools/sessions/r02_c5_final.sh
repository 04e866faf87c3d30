#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/../.."
timeout 300 python bench.py --workload c5 --seqs-per-gpu 32 --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/fin_bench_c5_s32.json 2> gpurun_out/fin_bench_c5_s32.err; echo "rc=$?"
python - <<PY
import json
d = json.load(open("gpurun_out/fin_bench_c5_s32.json"))
print("scans/s", round(d["scans_per_s"]), "points/s", round(d["value"] / 1e6), "M  e2e", round(d["e2e"]["value"] / 1e6), "M  launches", d["gpu_launches"], d["roofline"].get("frac"))
PY
