#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_zzz_gpu_ao3c_device.py -m gpu -x -q > gpurun_out/c8_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/c8_pytest.log
GWBSE_AO3C_MINB2=1 timeout 600 python scratch/ao3c_bench.py --system c60-tzvp --reps 1 > gpurun_out/c8_ao3c_c60_minb2.log 2>&1; echo "minb2 rc=$?"
timeout 600 python scratch/ao3c_bench.py --system c60-tzvp --reps 1 > gpurun_out/c8_ao3c_c60_minb4.log 2>&1; echo "minb4 rc=$?"
grep -E "pass|best" gpurun_out/c8_ao3c_c60_minb2.log gpurun_out/c8_ao3c_c60_minb4.log
head -40 gpurun_out/c8_ao3c_c60_minb4.log | tail -34
