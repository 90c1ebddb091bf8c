#!/bin/bash
mkdir -p gpurun_out
for d in 1 2 4 8 16; do
  GWBSE_AO3C_LANEDIV=$d timeout 600 python scratch/ao3c_bench.py --system c60-tzvp --reps 1 --check 0 > gpurun_out/c9_ao3c_c60_div$d.log 2>&1; echo "div $d rc=$? $(grep best gpurun_out/c9_ao3c_c60_div$d.log)"
done
GWBSE_AO3C_LANEDIV=4 timeout 600 python -m pytest tests/test_zzz_gpu_ao3c_device.py -m gpu -x -q 2>&1 | tail -2
GWBSE_AO3C_LANEDIV=8 timeout 600 python scratch/ao3c_bench.py --system benzene-tzvp --reps 1 --check 2 2>&1 | grep -E "best|deviation"
