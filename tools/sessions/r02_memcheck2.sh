#!/usr/bin/env bash
# compute-sanitizer memcheck, second batch: pipeline parity, multi-sequence driver, scan kernels (bounded: 170 s)
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/../.."
timeout 170 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_pipeline_parity.py tests/test_sequences.py tests/test_scan_parity.py tests/test_map_parity.py -m gpu -q -x > gpurun_out/memcheck2.log 2>&1; echo "memcheck rc=$?" | tee -a gpurun_out/memcheck2.log
grep -E "passed|failed|ERROR SUMMARY|Invalid|error" gpurun_out/memcheck2.log | tail -8
