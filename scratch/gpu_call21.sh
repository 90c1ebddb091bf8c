#!/bin/bash
mkdir -p gpurun_out
for w in none gemm eig; do timeout 200 python scratch/tma_coresidency_repro.py $w 2>&1 | tail -4; done
for v in NO_PREFETCH NO_SETMAXNREG; do echo "== variant $v"; for w in gemm eig; do GWBSE_B200_TEST_MOCK_DIR=$PWD/scratch/libvar_$v timeout 200 python scratch/tma_coresidency_repro.py $w 2>&1 | tail -4; done; done
echo "== TMA off"; GWBSE_NO_TMA=1 timeout 200 python scratch/tma_coresidency_repro.py gemm 2>&1 | tail -2
