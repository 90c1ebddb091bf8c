#!/bin/bash
mkdir -p gpurun_out
GWBSE_NO_TMA=1 GWBSE_FILL_OVERLAP_DEBUG=E timeout 300 python scratch/debug_fill_overlap.py modeE_notma > gpurun_out/c20_E_notma.log 2>&1; tail -3 gpurun_out/c20_E_notma.log
GWBSE_FILL_OVERLAP_DEBUG=E timeout 300 python scratch/debug_fill_overlap.py modeE_tma > gpurun_out/c20_E_tma.log 2>&1; tail -3 gpurun_out/c20_E_tma.log
