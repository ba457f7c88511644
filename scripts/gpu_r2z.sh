#!/bin/bash
# compute-sanitizer: memcheck over the whole GPU suite; racecheck + synccheck over the smoke run
mkdir -p gpurun_out
export PYTHONPATH=$PWD
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 python -m pytest tests -m gpu -q --no-header -x > gpurun_out/r02_memcheck_suite.log 2>&1; echo "memcheck suite rc=$?"; grep -E "ERROR SUMMARY|passed|failed|Invalid|Misaligned" gpurun_out/r02_memcheck_suite.log | head -8
timeout 600 compute-sanitizer --tool synccheck --error-exitcode 9 --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_synccheck_smoke.log 2>&1; echo "synccheck smoke rc=$?"; grep -E "ERROR SUMMARY|smoke OK|Barrier|divergent" gpurun_out/r02_synccheck_smoke.log | head -8
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_racecheck_smoke.log 2>&1; echo "racecheck smoke rc=$?"; grep -E "RACECHECK SUMMARY|ERROR SUMMARY|smoke OK|hazard" gpurun_out/r02_racecheck_smoke.log | head -12
