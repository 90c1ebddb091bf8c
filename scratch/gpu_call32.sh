#!/bin/bash
# windowed MultiplyRight: full GPU suite, then the C60 step
mkdir -p gpurun_out
timeout 170 python -m pytest tests -m gpu -x -q > gpurun_out/c32_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/c32_pytest.log
GWBSE_PROFILE=gpurun_out/c32_profile_c60.txt timeout 100 python bench.py --steps 1 --warmup 1 --also '' --no-cpu --no-e2e > gpurun_out/c32_bench.json 2> gpurun_out/c32_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/c32_bench.json").read().strip().splitlines()[-1])
    print("C60", d["value"], "frac", d["roofline"]["frac"], d["run"]["stage_seconds"], d["run"]["results"], d["run"]["bse_direct_terms"]["blocks_built"])
except Exception as e: print("bench parse failed", e)
PY
grep -n "mmn_mul_right" gpurun_out/c32_profile_c60.txt | head -3
tail -2 gpurun_out/c32_bench.err | cut -c1-200
