#!/bin/bash
# round-2 call 2: TMA-staged GEMM correctness + timing sweep, regression of the kernel tests, ao3c profile
mkdir -p gpurun_out
timeout 600 python scratch/gemm_tma_check.py > gpurun_out/c2_tma_check.log 2>&1; echo "tma check rc=$?"
tail -25 gpurun_out/c2_tma_check.log
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_host.py -m gpu -x -q > gpurun_out/c2_pytest.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/c2_pytest.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ao3c_kernel -s 100 -c 45 -o gpurun_out/r02_ao3c_benzene python scratch/ao3c_bench.py --system benzene-tzvp --reps 0 > gpurun_out/c2_ao3c_ncu.log 2>&1; echo "ao3c ncu rc=$?"
