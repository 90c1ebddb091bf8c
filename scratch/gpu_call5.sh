#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/c5_probe.log; : > $L
S="N N 65536 3200 3200 1 0 0 3"
for b in probe probe_ORDER_JI probe_ORDER_SNAKE probe_PROBE_NO_LDS; do
  for c in 0 1; do echo "== $b cfg $c" >> $L; timeout 120 scratch/bin/$b $c $S 2>&1 | grep TIME >> $L; done
done
for c in 104 100 103; do echo "== cp.async cfg $c" >> $L; timeout 120 scratch/bin/probe $c $S 2>&1 | grep TIME >> $L; done
echo "== TN" >> $L
for c in 0 1 104; do timeout 120 scratch/bin/probe $c T N 16384 3200 3200 1 0 0 3 2>&1 | grep TIME >> $L; done
echo "== short K" >> $L
for c in 0 1 104; do timeout 120 scratch/bin/probe $c T N 4320 50816 288 1 0 0 3 2>&1 | grep TIME >> $L; done
cat $L
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tma_kernel -s 1 -c 1 -o gpurun_out/r02_gemm_tma_cfg10_mulright scratch/bin/probe 0 N N 65536 3200 3200 1 0 0 1 > gpurun_out/c5_ncu1.log 2>&1; echo "ncu1 rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tma_kernel -s 1 -c 1 -o gpurun_out/r02_gemm_tma_cfg11_shortk scratch/bin/probe 1 T N 4320 50816 288 1 0 0 1 > gpurun_out/c5_ncu2.log 2>&1; echo "ncu2 rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ao3c_kernel -s 100 -c 45 -o gpurun_out/r02_ao3c_benzene python scratch/ao3c_bench.py --system benzene-tzvp --reps 0 > gpurun_out/c5_ao3c_ncu.log 2>&1; echo "ao3c ncu rc=$?"
ls -la gpurun_out/*.ncu-rep
