#!/bin/bash
# coldots cluster kernel check, streaming table again, full launch list of the current build, ncu --set full of eps SYRK + short-K leg
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "gramschmidt or davidson or bse or dense" > gpurun_out/c12_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/c12_pytest.log
timeout 600 python scratch/streaming_roofline.py > gpurun_out/c12_streaming.log 2>&1; echo "streaming rc=$?"; grep -v "^ALG" gpurun_out/c12_streaming.log | tail -18
timeout 400 python scratch/ncu_eps.py > gpurun_out/c12_eps.log 2>&1; echo "eps rc=$?"; tail -2 gpurun_out/c12_eps.log
timeout 400 ncu --set full --clock-control none --import-source on -k regex:gemm_tma_kernel -s 1 -c 1 -o gpurun_out/r02_gemm_tma_eps_syrk python scratch/ncu_eps.py > gpurun_out/c12_ncu_eps.log 2>&1; echo "ncu eps rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:gemm_dmma_kernel -s 3 -c 1 -o gpurun_out/r02_gemm_shortk_bse_leg python scratch/ncu_gemm_shortk.py > gpurun_out/c12_ncu_shortk.log 2>&1; echo "ncu shortk rc=$?"; grep TFLOP gpurun_out/c12_ncu_shortk.log | head -3
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 60000 --csv --log-file gpurun_out/r02_launches_dcv5t_full.csv python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu --no-parity --also '' > gpurun_out/c12_bench_under_ncu.log 2>&1; echo "launch list rc=$?"
gzip -f gpurun_out/r02_launches_dcv5t_full.csv
du -sh gpurun_out
