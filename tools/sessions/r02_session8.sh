#!/usr/bin/env bash
# Round-2 single-GPU session 8: seeded / pruned nearest-point pass for far queries (the 1 ms scan of the 105-scan replay),
# final numbers: GPU suite, 105-scan A/B, headline bench (100 steps), C5 with 32 / 64 sequences, launch list.
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/../.."
T=${TAG:-s8}
echo "== 1. GPU test-suite"
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/${T}_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${T}_pytest_gpu.log
tail -4 gpurun_out/${T}_pytest_gpu.log
echo "== 2. A/B over 105 scans: default | r = no reuse | i = synchronous insert"
AB_SCANS=105 timeout 500 python tools/ab_latency.py 0 0r 0i 2>&1 | tee gpurun_out/${T}_ab_latency.log | tail -14
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
echo "== 4. C5"
for cfg in "32 0" "64 0"; do
  set -- $cfg
  timeout 400 python bench.py --workload c5 --seqs-per-gpu $1 --c5-threads $2 --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_bench_c5_s$1.json 2> gpurun_out/${T}_bench_c5_s$1.err; echo "c5 s=$1 rc=$?"
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${T}_bench_c5_s$1.json"))
    print("  scans/s", round(d["scans_per_s"]), "points/s", round(d["value"] / 1e6), "M  e2e", round(d["e2e"]["value"] / 1e6), "M  launches", d["gpu_launches"])
except Exception as e:
    print("  unreadable:", e)
PY
done
echo "== 5. ncu launch list"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${T}_launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-replay > gpurun_out/${T}_ncu_bench.log 2>&1; echo "ncu rc=$?"
