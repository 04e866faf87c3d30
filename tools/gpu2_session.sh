#!/usr/bin/env bash
# 2-GPU session: sharded-map tests on the final code, default bench at N=2 (replicas + C4 leg), C4 loop placement / transport A/B.
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
echo "== sharded GPU tests"
timeout 600 python -m pytest tests/test_sharded.py -m gpu -q > gpurun_out/g2_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/g2_pytest.log; tail -3 gpurun_out/g2_pytest.log
echo "== default bench at N=2"
timeout 600 $TR --master-port 29701 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/g2_bench_default.json 2> gpurun_out/g2_bench_default.err; echo "rc=$?"
for cfg in "0 peer" "2 peer" "1 peer" "0 nccl" "2 nccl"; do
  set -- $cfg
  echo "== c4 device-loop $1 exchange $2"
  timeout 400 $TR --master-port 2971$1 bench.py --gpus 2 --workload c4 --steps 40 --warmup 5 --device-loop $1 --shard-exchange $2 > gpurun_out/g2_c4_loop$1_$2.json 2> gpurun_out/g2_c4_loop$1_$2.err; echo "rc=$?"
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/g2_*.json')):
    try:
        d=json.load(open(f))
        print(f, 'p50', round(d['ms_p50'],4), 'mean', round(d['ms_per_step'],4), 'e2e p50', round(d['e2e']['ms_p50'],4), 'launches/scan', d['gpu_launches']/d['steps']/d['n_gpus'], (d.get('c4') or {}).get('ms_p50'), (d.get('c4') or {}).get('error'))
    except Exception as e:
        print(f, 'unreadable', e)
PY
