#!/bin/bash
export PYTHONPATH=$PWD
mkdir -p gpurun_out; : > gpurun_out/r02_gather_scaling.jsonl
for p in zipf uniform sequential; do for g in 18 37 74 111 148; do timeout 60 python scripts/gather_scaling.py $p $g | tee -a gpurun_out/r02_gather_scaling.jsonl; done; done
