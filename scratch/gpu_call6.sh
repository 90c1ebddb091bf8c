#!/bin/bash
mkdir -p gpurun_out
timeout 900 python scratch/gemm_tma_check.py > gpurun_out/c6_tma_check.log 2>&1; echo "tma check rc=$?"
tail -14 gpurun_out/c6_tma_check.log
timeout 1200 python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/c6_pytest.log 2>&1; echo "pytest rc=$?"
tail -22 gpurun_out/c6_pytest.log
GWBSE_PROFILE=gpurun_out/c6_profile_dcv5t.txt timeout 900 python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e > gpurun_out/c6_bench_dcv5t.json 2> gpurun_out/c6_bench_dcv5t.err; echo "bench dcv5t rc=$?"
GWBSE_PROFILE=gpurun_out/c6_profile_c60.txt timeout 1200 python bench.py --workload c60-tzvp --steps 1 --warmup 1 --no-cpu > gpurun_out/c6_bench_c60.json 2> gpurun_out/c6_bench_c60.err; echo "bench c60 rc=$?"
tail -3 gpurun_out/c6_bench_c60.err
python - <<'PY'
import json
for t in ("dcv5t","c60"):
    try:
        d=json.loads(open(f"gpurun_out/c6_bench_{t}.json").read().strip().splitlines()[-1])
        print(t, d["value"], d["roofline"]["achieved"], d["roofline"]["frac"], d["config"]["stage_seconds"], d["config"]["results"], d["config"]["gw_iterations"], d["config"]["davidson_iterations"], d.get("e2e"))
    except Exception as e:
        print(t, "failed", e)
PY
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tma_kernel -s 1 -c 1 -o gpurun_out/r02_gemm_tma_cfg10_mulright scratch/bin/probe 0 N N 65536 3200 3200 1 0 0 1 > gpurun_out/c6_ncu1.log 2>&1; echo "ncu1 rc=$?"
du -sh gpurun_out
