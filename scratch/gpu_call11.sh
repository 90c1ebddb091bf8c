#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scratch/streaming_roofline.py > gpurun_out/c11_streaming.log 2>&1; echo "streaming rc=$?"; grep -v "^ALG" gpurun_out/c11_streaming.log | tail -22
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:'coldots|axpy_kernel|scale_cols|dpr_kernel|olsen|sigma_multi|tree_|sigma_offdiag_weight|slice_diag|bse_diag|rowdot|rpa_weights|symmetrize|diag_scale' --csv --log-file gpurun_out/c11_streaming_ncu.csv python scratch/streaming_roofline.py > gpurun_out/c11_streaming_under_ncu.log 2>&1; echo "streaming ncu rc=$?"
gzip -f gpurun_out/c11_streaming_ncu.csv
K="not medium_size and not full_size and not large_l"
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_kernels.py tests/test_zzz_gpu_ao3c_device.py -m gpu -x -q -k "$K" > gpurun_out/c11_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -6 gpurun_out/c11_memcheck.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python scratch/gemm_tma_check.py --quick > gpurun_out/c11_memcheck_tma.log 2>&1; echo "memcheck tma rc=$?"; tail -4 gpurun_out/c11_memcheck_tma.log
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_kernels.py tests/test_zzz_gpu_ao3c_device.py -m gpu -x -q -k "test_dgemm_splitk_and_beta or test_sigma_ppm or test_sigma_tree_exact or test_gramschmidt or test_bse_operator_golden or test_ao3c_water or test_sigma_x" > gpurun_out/c11_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -6 gpurun_out/c11_racecheck.log
