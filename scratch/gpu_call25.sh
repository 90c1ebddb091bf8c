#!/bin/bash
# materialised BSE direct-term blocks: full GPU suite, then the C60 headline step with the region profile
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/c25_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/c25_pytest.log
GWBSE_PROFILE=gpurun_out/c25_profile_c60.txt timeout 600 python bench.py --steps 1 --warmup 1 --also '' --no-cpu --no-e2e > gpurun_out/c25_bench.json 2> gpurun_out/c25_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/c25_bench.err | cut -c1-300
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/c25_bench.json").read().strip().splitlines()[-1])
    print("C60", d["value"], "frac", d["roofline"]["frac"], d["run"]["stage_seconds"], d["run"]["results"], d["run"]["bse_direct_terms"])
except Exception as e: print("bench parse failed", e)
PY
head -14 gpurun_out/c25_profile_c60.txt
grep -n "M=" gpurun_out/c25_profile_c60.txt | sort -k2 | head -5
nvidia-smi --query-gpu=memory.used,memory.total --format=csv
