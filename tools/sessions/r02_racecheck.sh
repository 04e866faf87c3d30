#!/usr/bin/env bash
# compute-sanitizer racecheck (shared-memory hazards) over the kernels new in round 2 (bounded: 130 s)
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/../.."
timeout 130 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests/test_variants.py tests/test_measure_parity.py tests/test_map_parity.py -m gpu -q -x -k "reuse or far_points or insert or ties" > gpurun_out/racecheck.log 2>&1; echo "racecheck rc=$?" | tee -a gpurun_out/racecheck.log
grep -E "passed|failed|RACECHECK SUMMARY|hazard|error" gpurun_out/racecheck.log | tail -8
