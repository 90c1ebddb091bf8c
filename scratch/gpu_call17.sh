#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scratch/debug_side_invsqrt.py > gpurun_out/c17_debug.log 2>&1; echo "rc=$?"; tail -12 gpurun_out/c17_debug.log
