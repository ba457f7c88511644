#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/sim3_trace.py > gpurun_out/sim3_trace_full.txt 2> gpurun_out/sim3_trace.err; tail -3 gpurun_out/sim3_trace.err
CAPR_SIM3_DEBUG=15 timeout 300 python scripts/sim3_trace.py > gpurun_out/sim3_trace_off.txt 2>> gpurun_out/sim3_trace.err
CAPR_SIM3_DEBUG=1 timeout 300 python scripts/sim3_trace.py > gpurun_out/sim3_trace_nopool.txt 2>> gpurun_out/sim3_trace.err
head -3 gpurun_out/sim3_trace_full.txt; head -2 gpurun_out/sim3_trace_off.txt; head -2 gpurun_out/sim3_trace_nopool.txt
