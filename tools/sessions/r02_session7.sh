#!/usr/bin/env bash
# Round-2 single-GPU session 7: 128-visit tags / 256-bucket ring batches (the ~1 ms exact-fallback step of the 105-scan replay),
# reuse pass as its own kernel.
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/../.."
T=${TAG:-s7}
echo "== 1. GPU test-suite"
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/${T}_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${T}_pytest_gpu.log
tail -4 gpurun_out/${T}_pytest_gpu.log
echo "== 2. A/B over 105 scans: default | r = no reuse"
AB_SCANS=105 timeout 500 python tools/ab_latency.py 0 0r 2>&1 | tee gpurun_out/${T}_ab_latency.log | tail -12
echo "== 3. headline bench"
timeout 500 python bench.py > gpurun_out/${T}_bench_c2.json 2> gpurun_out/${T}_bench_c2.err; echo "bench rc=$?"
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${T}_bench_c2.json"))
    print("C2 p50", d.get("ms_p50"), "mean", d.get("ms_per_step"), "p99", d.get("ms_p99"), "e2e p50", d["e2e"].get("ms_p50"), "serial p50", d["e2e"]["serial"].get("ms_p50"), "kernels", d["roofline"].get("kernel_ms_per_scan"), "parity ok", d.get("parity", {}).get("ok"), "slow", d["detail"].get("slow_steps"))
except Exception as e:
    print("bench line unreadable:", e)
PY
echo "== 4. ncu launch list"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${T}_launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-replay > gpurun_out/${T}_ncu_bench.log 2>&1; echo "ncu rc=$?"
