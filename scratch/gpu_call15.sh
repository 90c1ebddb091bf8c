#!/bin/bash
# 8 GPUs: C60 headline (BASELINE config 4: Mmn sharded over 8 B200 with NCCL epsilon reduction), region profile of rank 0
mkdir -p gpurun_out
nvidia-smi -L | wc -l
GWBSE_PROFILE=gpurun_out/c15_profile_c60_8gpu.txt timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 8 --steps 2 --warmup 1 --also '' --e2e-steps 1 > gpurun_out/c15_bench_8gpu.json 2> gpurun_out/c15_bench_8gpu.err; echo "bench8 rc=$?"
tail -3 gpurun_out/c15_bench_8gpu.err | cut -c1-300
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/c15_bench_8gpu.json").read().strip().splitlines()[-1])
    print("N=8 C60", d["value"], "e2e", d["e2e"]["value"] if d["e2e"] else None, "frac", d["roofline"]["frac"], "share", d["roofline"]["kernel_share_of_step"], d["run"]["stage_seconds"], d.get("sharded_vs_single"), d["run"]["results"])
except Exception as e: print("bench parse failed", e)
PY
head -24 gpurun_out/c15_profile_c60_8gpu.txt
