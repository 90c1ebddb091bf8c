#!/bin/bash
# block build re-oriented (contiguous stores for Hd too) on the 128 x 128 TMA tile: BSE tests, then the C60 step
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_zzz_large_pipeline.py tests/test_gpu_host.py -m gpu -x -q -k "bse or pipeline or config2" > gpurun_out/c30_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/c30_pytest.log
timeout 200 python scratch/ncu_bse_dense.py 3 2>&1 | grep -v "^("
GWBSE_PROFILE=gpurun_out/c30_profile_c60.txt timeout 600 python bench.py --steps 1 --warmup 1 --also '' --no-cpu --no-e2e > gpurun_out/c30_bench.json 2> gpurun_out/c30_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/c30_bench.json").read().strip().splitlines()[-1])
    print("C60", d["value"], "frac", d["roofline"]["frac"], d["run"]["stage_seconds"], d["run"]["results"], d["run"]["bse_direct_terms"]["blocks_built"])
except Exception as e: print("bench parse failed", e)
PY
grep -n "AmBm" gpurun_out/c30_profile_c60.txt | head -5
