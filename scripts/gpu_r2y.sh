#!/bin/bash
# compute-sanitizer memcheck over the smoke run and a small slice of the parity tests (invalid / misaligned / out-of-bounds accesses)
mkdir -p gpurun_out
export PYTHONPATH=$PWD
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_memcheck_smoke.log 2>&1; echo "memcheck smoke rc=$?"; grep -E "ERROR SUMMARY|smoke OK|Invalid|Misaligned" gpurun_out/r02_memcheck_smoke.log | head -8
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 python -m pytest tests/test_gpu_engine3.py tests/test_gpu_parity.py -q --no-header -x -k "small or odd or B0 or engine3_variants or tf_dedup" > gpurun_out/r02_memcheck_tests.log 2>&1; echo "memcheck tests rc=$?"; grep -E "ERROR SUMMARY|passed|failed|Invalid|Misaligned" gpurun_out/r02_memcheck_tests.log | head -8
