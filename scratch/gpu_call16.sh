#!/bin/bash
# V^-1/2 beside the fill (side context + host thread), threaded pair lists: correctness + effect on the C60 step and its e2e arm
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/c16_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/c16_pytest.log
GWBSE_PROFILE=gpurun_out/c16_profile_c60.txt timeout 900 python bench.py --steps 2 --warmup 1 --also '' --no-cpu --e2e-steps 2 > gpurun_out/c16_bench.json 2> gpurun_out/c16_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/c16_bench.err | cut -c1-200
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/c16_bench.json").read().strip().splitlines()[-1])
    print("C60", d["value"], "e2e", d["e2e"]["value"], d["e2e"]["stage_seconds"], "frac", d["roofline"]["frac"], d["run"]["stage_seconds"], d["run"]["results"])
except Exception as e: print("bench parse failed", e)
PY
head -12 gpurun_out/c16_profile_c60.txt
GWBSE_NO_FILL_OVERLAP=1 timeout 600 python bench.py --steps 2 --warmup 1 --also '' --no-cpu --no-e2e > gpurun_out/c16_bench_nooverlap.json 2> /dev/null; python -c "
import json; d=json.loads(open('gpurun_out/c16_bench_nooverlap.json').read().strip().splitlines()[-1]); print('no overlap', d['value'], d['run']['stage_seconds'])"
