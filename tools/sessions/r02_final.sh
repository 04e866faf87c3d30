#!/usr/bin/env bash
# Round-2 final single-GPU validation: GPU suite, smoke(), headline bench (both arms), ncu --set full of one scan.
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/../.."
T=${TAG:-fin}
echo "== 1. GPU test-suite"
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/${T}_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${T}_pytest_gpu.log
tail -3 gpurun_out/${T}_pytest_gpu.log
echo "== 2. smoke()"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== 3. headline bench"
timeout 500 python bench.py > gpurun_out/${T}_bench_c2.json 2> gpurun_out/${T}_bench_c2.err; echo "bench rc=$?"
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${T}_bench_c2.json"))
    print("C2 p50", d.get("ms_p50"), "mean", d.get("ms_per_step"), "p99", d.get("ms_p99"), "e2e p50", d["e2e"].get("ms_p50"), "serial p50", d["e2e"]["serial"].get("ms_p50"), "parity ok", d.get("parity", {}).get("ok"), "slow", d["detail"].get("slow_steps"), "cpu", (d.get("cpu_baseline") or {}).get("value"))
    print("roofline", {k: d["roofline"].get(k) for k in ("kernel", "achieved", "frac", "avg_launch_us")}, "issue", (d["roofline"].get("issue") or {}).get("frac"), "step", (d["roofline"].get("step") or {}).get("frac"))
except Exception as e:
    print("bench line unreadable:", e)
PY
echo "== 4. ncu --set full of every kernel of one scan"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'^k_|dlt' -s 68 -c 17 \
    -o gpurun_out/${T}_full -f python tools/prof_replay.py --scans 6 > gpurun_out/${T}_ncu_full.log 2>&1; echo "ncu full rc=$?"
