#!/usr/bin/env bash
# Round-2 single-GPU session: GPU suite, headline bench (both arms), loop placement A/B, k_knn8 variants, ncu launch list and
# one --set full capture of the hot kernels.  Every step is bounded by its own timeout; results land in gpurun_out/.
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/../.."
T=${TAG:-s1}

echo "== 1. GPU test-suite"
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/${T}_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${T}_pytest_gpu.log
tail -5 gpurun_out/${T}_pytest_gpu.log

echo "== 2. headline bench"
timeout 400 python bench.py > gpurun_out/${T}_bench_c2.json 2> gpurun_out/${T}_bench_c2.err; echo "bench rc=$?"
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${T}_bench_c2.json"))
    print("C2 p50", d.get("ms_p50"), "mean", d.get("ms_per_step"), "e2e", d["e2e"], "roofline", d.get("roofline"), "parity", d.get("parity"), "launches", d.get("gpu_launches"))
except Exception as e:
    print("bench line unreadable:", e)
PY

echo "== 3. loop placement A/B"
timeout 300 python tools/ab_latency.py ${AB_MODES:-0 0s 1 1p 2} 2>&1 | tee gpurun_out/${T}_ab_latency.log | tail -12

echo "== 4. k_knn8 variants"
timeout 300 python tools/knn_variants.py 2>&1 | tee gpurun_out/${T}_knn_variants.log | tail -8

echo "== 5. ncu launch list of the bench command"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${T}_launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-replay > gpurun_out/${T}_ncu_bench.log 2>&1; echo "ncu rc=$?"

echo "== 6. ncu --set full of the hot kernels (one scan's worth)"
timeout 500 ncu --set full --clock-control none --import-source on -k regex:'k_knn8|k_residual|k_loop|k_vox|k_scan_deskew|k_incr' -s 60 -c 40 \
    -o gpurun_out/${T}_full -f python tools/prof_replay.py --scans 6 > gpurun_out/${T}_ncu_full.log 2>&1; echo "ncu full rc=$?"
ls -la gpurun_out/ | tail -20
