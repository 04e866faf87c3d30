#!/usr/bin/env bash
# Round-2 single-GPU session 3: reuse of proven neighbour sets on rematch passes + map_incremental off the critical path.
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/../.."
T=${TAG:-s3}
echo "== 1. GPU test-suite"
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/${T}_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${T}_pytest_gpu.log
tail -4 gpurun_out/${T}_pytest_gpu.log
echo "== 2. A/B: default | r = no reuse | i = synchronous insert | ri = both off"
timeout 400 python tools/ab_latency.py ${AB_MODES:-0 0r 0i 0ri} 2>&1 | tee gpurun_out/${T}_ab_latency.log | tail -10
echo "== 3. headline bench"
timeout 500 python bench.py > gpurun_out/${T}_bench_c2.json 2> gpurun_out/${T}_bench_c2.err; echo "bench rc=$?"
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${T}_bench_c2.json"))
    print("C2 p50", d.get("ms_p50"), "mean", d.get("ms_per_step"), "e2e p50", d["e2e"].get("ms_p50"), "serial p50", d["e2e"]["serial"].get("ms_p50"), "kernels", d["roofline"].get("kernel_ms_per_scan"), "parity ok", d.get("parity", {}).get("ok"), "replay", d.get("replay"))
except Exception as e:
    print("bench line unreadable:", e)
PY
echo "== 4. ncu launch list"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${T}_launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-replay > gpurun_out/${T}_ncu_bench.log 2>&1; echo "ncu rc=$?"
