#!/bin/bash
# round-2 call 1: GPU test suite, AO-integral producer timing, ncu launch list (time + DRAM bytes) of the current build
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/c1_gpu.txt; nproc >> gpurun_out/c1_gpu.txt; free -g >> gpurun_out/c1_gpu.txt
timeout 600 python -m pytest tests -m gpu -x -q -rxXs > gpurun_out/c1_pytest.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/c1_pytest.log
timeout 300 python scratch/ao3c_bench.py --system benzene-tzvp --check 2 > gpurun_out/c1_ao3c_benzene.log 2>&1; echo "ao3c benzene rc=$?"
timeout 600 python scratch/ao3c_bench.py --system c60-tzvp --reps 2 > gpurun_out/c1_ao3c_c60.log 2>&1; echo "ao3c c60 rc=$?"
tail -30 gpurun_out/c1_ao3c_c60.log
timeout 1500 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r02_launches_dcv5t.csv python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/c1_bench_under_ncu.log 2>&1; echo "ncu rc=$?"
wc -l gpurun_out/r02_launches_dcv5t.csv
gzip -f gpurun_out/r02_launches_dcv5t.csv
