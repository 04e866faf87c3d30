#!/usr/bin/env bash
# compute-sanitizer memcheck over the GPU tests that run the kernels new in round 2 (bounded: 150 s)
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/../.."
timeout 150 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_variants.py tests/test_measure_parity.py -m gpu -q -x -k "reuse or async or far_points" > gpurun_out/memcheck.log 2>&1; echo "memcheck rc=$?" | tee -a gpurun_out/memcheck.log
grep -E "passed|failed|ERROR SUMMARY|Invalid|error" gpurun_out/memcheck.log | tail -8
