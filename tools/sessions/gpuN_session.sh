#!/usr/bin/env bash
# N-GPU session (N = $1, default 2): sharded-map GPU tests (N = 2 only), the default bench at N (replicas + the sharded C4 leg),
# C4 alone with more steps, C5 with 64 / N sequences per GPU.
set -u
N=${1:-2}
mkdir -p gpurun_out
cd "$(dirname "$0")/../.."
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
if [ "$N" = "2" ]; then
  echo "== sharded GPU tests"
  timeout 600 python -m pytest tests/test_sharded.py -m gpu -q > gpurun_out/g${N}_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/g${N}_pytest.log; tail -3 gpurun_out/g${N}_pytest.log
fi
echo "== default bench at N=$N"
timeout 600 $TR --master-port 29701 bench.py --gpus $N --steps 40 --warmup 5 > gpurun_out/g${N}_bench_default.json 2> gpurun_out/g${N}_bench_default.err; echo "rc=$?"
if [ "${SKIP_C4_ALONE:-0}" != "1" ]; then
echo "== c4 alone, peer exchange, host loop"
timeout 400 $TR --master-port 29711 bench.py --gpus $N --workload c4 --steps 60 --warmup 5 > gpurun_out/g${N}_c4_peer.json 2> gpurun_out/g${N}_c4_peer.err; echo "rc=$?"
fi
echo "== c5, 64 sequences over $N GPUs"
timeout 500 $TR --master-port 29721 bench.py --gpus $N --workload c5 --seqs-per-gpu $((64 / N)) --steps 30 --warmup 3 > gpurun_out/g${N}_c5.json 2> gpurun_out/g${N}_c5.err; echo "rc=$?"
python - <<PY
import json
for f in ("bench_default", "c4_peer", "c5"):
    try:
        d = json.load(open("gpurun_out/g${N}_%s.json" % f))
        print(f, "value", round(d["value"] / 1e6, 1), "M pts/s  p50", d.get("ms_p50"), "mean", d.get("ms_per_step"), "scans/s", round(d.get("scans_per_s", 0)), "e2e", round(d["e2e"]["value"] / 1e6, 1),
              "c4", {k: (d.get("c4") or {}).get(k) for k in ("ms_p50", "ms_per_step", "sharded_over_unsharded_p50", "error")} if f == "bench_default" else "")
    except Exception as e:
        print(f, "unreadable", e)
        try:
            print(open("gpurun_out/g${N}_%s.err" % f).read()[-800:])
        except Exception:
            pass
PY
