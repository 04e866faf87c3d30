#!/usr/bin/env bash
# Round-2 single-GPU session 2: k_knn8 with 32-bit keys against the previous build (gpurun_scratch/*.so), the GPU suite, the
# headline bench, C5 with 16 / 64 sequences, C4 unsharded, and an ncu --set full capture of EVERY kernel of one scan.
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/../.."
T=${TAG:-s2}
echo "== 1. k_knn8 A/B against the previous build"
timeout 300 python tools/knn_variants.py 2>&1 | tee gpurun_out/${T}_knn_variants.log | tail -6
echo "== 2. GPU test-suite"
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/${T}_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${T}_pytest_gpu.log
tail -4 gpurun_out/${T}_pytest_gpu.log
echo "== 3. headline bench"
timeout 400 python bench.py > gpurun_out/${T}_bench_c2.json 2> gpurun_out/${T}_bench_c2.err; echo "bench rc=$?"
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${T}_bench_c2.json"))
    print("C2 p50", d.get("ms_p50"), "mean", d.get("ms_per_step"), "e2e p50", d["e2e"].get("ms_p50"), "serial p50", d["e2e"]["serial"].get("ms_p50"), "kernels", d["roofline"].get("kernel_ms_per_scan"), "parity ok", d.get("parity", {}).get("ok"))
except Exception as e:
    print("bench line unreadable:", e)
PY
echo "== 4. C5 (16 and 64 sequences on one GPU), C4 unsharded"
for s in 16 64; do
  timeout 300 python bench.py --workload c5 --seqs-per-gpu $s --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_bench_c5_s$s.json 2> gpurun_out/${T}_bench_c5_s$s.err; echo "c5 s=$s rc=$?"
done
timeout 400 python bench.py --workload c4 --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/${T}_bench_c4_n1.json 2> gpurun_out/${T}_bench_c4_n1.err; echo "c4 rc=$?"
python - <<PY
import json
for f in ("c5_s16", "c5_s64", "c4_n1"):
    try:
        d = json.load(open("gpurun_out/${T}_bench_%s.json" % f))
        print(f, "value", d.get("value"), d.get("unit"), "scans/s", d.get("scans_per_s"), "p50", d.get("ms_p50"), "e2e", d["e2e"].get("value"))
    except Exception as e:
        print(f, "unreadable:", e)
PY
echo "== 5. ncu --set full of every kernel of one scan (scan 5 of a 6-scan replay)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'^k_|dlt' -s 68 -c 17 \
    -o gpurun_out/${T}_full -f python tools/prof_replay.py --scans 6 > gpurun_out/${T}_ncu_full.log 2>&1; echo "ncu full rc=$?"
tail -3 gpurun_out/${T}_ncu_full.log
