#!/bin/bash
mkdir -p gpurun_out
CAPR_SIM3_DEBUG=15 timeout 60 python scripts/sim3_trace.py > gpurun_out/sim3_trace_off.txt 2> gpurun_out/sim3_trace.err; head -1 gpurun_out/sim3_trace_off.txt
timeout 60 python scripts/sim3_trace.py > gpurun_out/sim3_trace_full.txt 2>> gpurun_out/sim3_trace.err; head -1 gpurun_out/sim3_trace_full.txt
CAPR_SIM3_DEBUG=4 timeout 60 python scripts/sim3_trace.py > gpurun_out/sim3_trace_nogather.txt 2>> gpurun_out/sim3_trace.err; head -1 gpurun_out/sim3_trace_nogather.txt
