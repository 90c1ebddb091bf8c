#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_zz_qsgw.py tests/test_zz_gpu_orb_output.py tests/test_gpu_host.py -m gpu -x -q > gpurun_out/c7_pytest.log 2>&1; echo "pytest rc=$?"
tail -8 gpurun_out/c7_pytest.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ao3c_kernel -s 100 -c 14 -o gpurun_out/r02_ao3c_benzene python scratch/ao3c_bench.py --system benzene-tzvp --reps 0 > gpurun_out/c7_ao3c_ncu.log 2>&1; echo "ao3c ncu rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:ao3c_kernel --csv --log-file gpurun_out/c7_ao3c_c60_launches.csv python scratch/ao3c_bench.py --system c60-tzvp --reps 0 --aux-block 64 > gpurun_out/c7_ao3c_c60.log 2>&1; echo "ao3c c60 list rc=$?"
gzip -f gpurun_out/c7_ao3c_c60_launches.csv
du -sh gpurun_out
