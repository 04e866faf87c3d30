#!/usr/bin/env bash
# First GPU call of the next session: everything that was built after the round-1 GPU budget ended,
# measured in one go on ONE B200.  Usage (from the repo root, on the build box):
#   /usr/local/graft/bin/gpurun --timeout 900 -- 'bash tools/gpu_session_next.sh'
# Every step is bounded by its own timeout; results land in gpurun_out/ (merged back by gpurun).
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."

echo "== 1. GPU test-suite (the two test_zz_* files have never run on a GPU)"
timeout 300 python -m pytest tests -m gpu -q > gpurun_out/next_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/next_pytest_gpu.log
tail -3 gpurun_out/next_pytest_gpu.log

echo "== 2. headline bench on the final code (host-side changes of DESIGN section 5 included)"
timeout 240 python bench.py > gpurun_out/next_bench_c2.json 2> gpurun_out/next_bench_c2.err; echo "bench rc=$?"
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/next_bench_c2.json"))
    print("C2 p50", d["ms_p50"], "ms  e2e p50", d["e2e"].get("ms_p50"), " host stages", d["config"].get("host_stage_ms_mean"))
except Exception as e:
    print("bench line unreadable:", e)
PY

echo "== 3. loop placement / zero-copy A/B (0 host loop, 0z + zero-copy result block, 2 / 1 device loop)"
timeout 240 python tools/ab_latency.py 0 0z 2 1 2>&1 | tee gpurun_out/next_ab_latency.log | tail -8

echo "== 4. k_knn8 variants (DLT_KNN8_PRUNE, +MINBLOCKS=5) against the product library"
ls gpurun_scratch/*.so > /dev/null 2>&1 || echo "(no variant libraries: run 'make -C daliti_b200/csrc variants' on the build box first)"
timeout 240 python tools/knn_variants.py 2>&1 | tee gpurun_out/next_knn_variants.log | tail -6

echo "== 5. ncu launch list of the bench command (shares of one scan)"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/next_launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/next_ncu_bench.log 2>&1; echo "ncu rc=$?"

echo "== done.  Multi-GPU follow-ups (separate calls, charged N x):"
echo "   gpurun --gpus 2 -- 'python -m pytest tests/test_sharded.py -m gpu -q; for x in peer nccl; do python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --workload c4 --steps 40 --warmup 5 --shard-exchange \$x > gpurun_out/next_c4_n2_\$x.json; done; python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29613 bench.py --gpus 2 --workload c4 --steps 40 --warmup 5 --device-loop 1 > gpurun_out/next_c4_n2_peer_devfinish.json'"
