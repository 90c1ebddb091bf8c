#!/bin/bash
# full GPU suite on the restated root search + BSECoupling + cluster column dots; default bench (C60 headline, DCV5T also); short-K ncu
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --durations=6 > gpurun_out/c13_pytest.log 2>&1; echo "pytest rc=$?"; tail -14 gpurun_out/c13_pytest.log
GWBSE_PROFILE=gpurun_out/c13_profile.txt timeout 1500 python bench.py --steps 2 --warmup 1 > gpurun_out/c13_bench.json 2> gpurun_out/c13_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/c13_bench.err | cut -c1-300
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/c13_bench.json").read().strip().splitlines()[-1])
    print("C60", d["value"], "e2e", d["e2e"]["value"] if d["e2e"] else None, "frac", d["roofline"]["frac"], d["run"]["stage_seconds"], d["run"]["results"])
    a=d.get("also"); print("also", a["value"], a["e2e"]["value"] if a["e2e"] else None, a["gemm_frac_of_peak"], a["config"]["results"]) if a else None
    print("cpu", d.get("cpu_baseline",{}).get("value"))
except Exception as e: print("bench parse failed", e)
PY
timeout 300 python scratch/ncu_gemm_shortk.py 359 > gpurun_out/c13_shortk.log 2>&1; tail -5 gpurun_out/c13_shortk.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_dmma_kernel -s 3 -c 1 -o gpurun_out/r02_gemm_shortk_evenpitch python scratch/ncu_gemm_shortk.py 359 > gpurun_out/c13_ncu_shortk.log 2>&1; echo "ncu shortk rc=$?"
timeout 200 bash -c 'time python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/c13_ref.json 2> gpurun_out/c13_ref.err'; echo "ref rc=$?"; cut -c1-400 gpurun_out/c13_ref.json
