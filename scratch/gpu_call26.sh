#!/bin/bash
# 2 GPUs with the materialised BSE blocks (sharded columns): sharded-vs-single test + short C60 bench with the parity field
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/c26_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/c26_pytest.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 2 --steps 1 --warmup 1 --also '' --no-e2e > gpurun_out/c26_bench_2gpu.json 2> gpurun_out/c26_bench_2gpu.err; echo "bench2 rc=$?"
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/c26_bench_2gpu.json").read().strip().splitlines()[-1])
    print("N=2 C60", d["value"], d["run"]["stage_seconds"], d.get("sharded_vs_single"), d["run"]["results"], d["run"]["bse_direct_terms"])
except Exception as e: print("parse failed", e)
PY
tail -5 gpurun_out/c26_bench_2gpu.err | cut -c1-300
