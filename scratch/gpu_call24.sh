#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus 4 --steps 2 --warmup 1 --also '' --e2e-steps 1 > gpurun_out/c24_bench_4gpu.json 2> gpurun_out/c24_bench_4gpu.err; echo "bench4 rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/c24_bench_4gpu.json").read().strip().splitlines()[-1])
print("N=4 C60", d["value"], "e2e", d["e2e"]["value"], d["run"]["stage_seconds"], d.get("sharded_vs_single"), d["run"]["results"])
PY
