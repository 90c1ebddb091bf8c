#!/bin/bash
# 8 GPUs, C60 headline with the materialised BSE blocks sharded over the ranks
mkdir -p gpurun_out
GWBSE_PROFILE=gpurun_out/c31_profile_c60_8gpu.txt timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29581 bench.py --gpus 8 --steps 1 --warmup 1 --also '' --no-e2e > gpurun_out/c31_bench_8gpu.json 2> gpurun_out/c31_bench_8gpu.err; echo "bench8 rc=$?"
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/c31_bench_8gpu.json").read().strip().splitlines()[-1])
    print("N=8 C60", d["value"], "frac", d["roofline"]["frac"], d["run"]["stage_seconds"], d.get("sharded_vs_single"), d["run"]["results"], d["run"]["bse_direct_terms"])
except Exception as e: print("bench parse failed", e)
PY
head -16 gpurun_out/c31_profile_c60_8gpu.txt
tail -3 gpurun_out/c31_bench_8gpu.err | cut -c1-300
