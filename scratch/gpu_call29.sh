#!/bin/bash
# build GEMM of the materialised BSE blocks: tile-shape timing, then one ncu --set full capture of the planner's choice
mkdir -p gpurun_out
for cfg in auto 10 11 12 14; do
  if [ "$cfg" = auto ]; then unset GWBSE_DENSE_BUILD_CFG; else export GWBSE_DENSE_BUILD_CFG=$cfg; fi
  timeout 200 python scratch/ncu_bse_dense.py 3 2>&1 | grep -v "^(" | tee -a gpurun_out/c29_build_cfgs.log
done
unset GWBSE_DENSE_BUILD_CFG
timeout 400 ncu --set full --clock-control none --import-source on -k regex:gemm_tma_kernel -c 1 -o gpurun_out/r02_gemm_tma_bse_block_build python scratch/ncu_bse_dense.py 1 > gpurun_out/c29_ncu.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/c29_ncu.log
ls -la gpurun_out/r02_gemm_tma_bse_block_build.ncu-rep
