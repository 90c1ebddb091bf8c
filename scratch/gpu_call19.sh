#!/bin/bash
mkdir -p gpurun_out
GWBSE_NO_FILL_OVERLAP=1 timeout 300 python scratch/debug_fill_overlap.py serial > gpurun_out/c19_serial.log 2>&1; tail -3 gpurun_out/c19_serial.log
for m in A E F; do GWBSE_FILL_OVERLAP_DEBUG=$m timeout 300 python scratch/debug_fill_overlap.py mode$m > gpurun_out/c19_$m.log 2>&1; tail -3 gpurun_out/c19_$m.log; done
