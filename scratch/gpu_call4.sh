#!/bin/bash
mkdir -p gpurun_out
timeout 900 python scratch/gemm_tma_check.py > gpurun_out/c4_tma_check.log 2>&1; echo "tma check rc=$?"
tail -22 gpurun_out/c4_tma_check.log
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/c4_pytest.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/c4_pytest.log
GWBSE_PROFILE=gpurun_out/c4_profile_tma.txt timeout 900 python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e > gpurun_out/c4_bench_tma.json 2> gpurun_out/c4_bench_tma.err; echo "bench tma rc=$?"
GWBSE_NO_TMA=1 GWBSE_PROFILE=gpurun_out/c4_profile_notma.txt timeout 900 python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e > gpurun_out/c4_bench_notma.json 2> gpurun_out/c4_bench_notma.err; echo "bench notma rc=$?"
python - <<'PY'
import json
for t in ("tma","notma"):
    try:
        d=json.loads(open(f"gpurun_out/c4_bench_{t}.json").read().strip().splitlines()[-1])
        print(t, d["value"], d["roofline"]["achieved"], d["roofline"]["frac"], d["config"]["stage_seconds"], d["config"]["results"])
    except Exception as e:
        print(t, "failed", e)
PY
