#!/bin/bash
# end-of-round check as the driver runs it: full GPU suite, smoke, default bench line (C60 headline + DCV5T + CPU baseline + e2e)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/c28_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/c28_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c28_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/c28_smoke.log
timeout 1500 python bench.py > gpurun_out/c28_bench_default.json 2> gpurun_out/c28_bench_default.err; echo "bench rc=$?"
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/c28_bench_default.json").read().strip().splitlines()[-1])
    print("C60 value", d["value"], "e2e", d["e2e"]["value"], "frac", d["roofline"]["frac"], "share", d["roofline"]["kernel_share_of_step"], d["run"]["stage_seconds"], d["run"]["bse_direct_terms"]["blocks_built"], d["clocks"])
    a = d.get("also") or {}
    print("also", a.get("value"), (a.get("e2e") or {}).get("value"), a.get("gemm_tflops"))
    print("cpu", d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"])
except Exception as e: print("bench parse failed", e)
PY
tail -3 gpurun_out/c28_bench_default.err | cut -c1-300
