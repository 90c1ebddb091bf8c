#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/c3_probe.log; : > $L
run() { echo "== $*" >> $L; timeout 60 scratch/bin/probe $* >> $L 2>&1; echo "rc=$?" >> $L; }
run 0 N N 300 200 150 1 0 0
run 0 N N 300 200 150 3 0 0
run 0 N N 300 200 150 1 2 0
run 0 N N 300 200 150 1 2 1
run 0 T N 300 200 150 1 0 0
run 0 N T 300 200 150 1 0 0
run 0 T T 300 200 150 1 0 0
run 0 N N 130 66 18 1 0 0
run 0 N N 1000 38 514 1 0 0
run 0 N N 64 64 16 1 0 0
run 0 N N 16 2 1030 1 0 0
run 0 N N 258 514 34 1 0 0
run 1 N N 300 200 150 3 2 1
run 2 T N 300 200 150 3 2 1
run 3 T T 300 200 150 3 2 1
run 4 N T 300 200 150 3 2 1
grep -B1 -A3 -E "rc=[1-9]" $L | head -60
echo; grep -c "rc=0" $L
