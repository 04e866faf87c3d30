#!/usr/bin/env bash
# Round-2 single-GPU session 9: the far-query scan (scan 86 of the C2 replay) under the device timeline, then the 105-scan A/B.
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/../.."
T=${TAG:-s9}
echo "== 1. timeline of scans 85-87"
TL_SCANS=88 TL_PRINT=85,86 timeout 400 python tools/timeline.py c2 2>&1 | grep -v "k_iekf_step\|gj_warp" | tee gpurun_out/${T}_timeline.log | tail -50
echo "== 2. far-point GPU tests"
timeout 600 python -m pytest tests/test_measure_parity.py tests/test_map_parity.py tests/test_pipeline_parity.py -m gpu -q -x 2>&1 | tail -3
echo "== 3. A/B over 105 scans"
AB_SCANS=105 timeout 500 python tools/ab_latency.py 0 2>&1 | tee gpurun_out/${T}_ab_latency.log | tail -6
