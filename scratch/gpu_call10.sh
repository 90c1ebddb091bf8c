#!/bin/bash
# 2 GPUs: sharded-vs-single parity test, bench at N=2 (headline + also, with parity field), reference arm under torchrun
mkdir -p gpurun_out
nvidia-smi -L | head -3
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/c10_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/c10_pytest.log
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 2 --warmup 1 > gpurun_out/c10_bench_2gpu.json 2> gpurun_out/c10_bench_2gpu.err; echo "bench2 rc=$?"
tail -3 gpurun_out/c10_bench_2gpu.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 > gpurun_out/c10_ref_2gpu.json 2> gpurun_out/c10_ref_2gpu.err; echo "ref2 rc=$?"
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/c10_bench_2gpu.json").read().strip().splitlines()[-1])
    print("N=2", d["value"], d["e2e"]["value"] if d["e2e"] else None, d["roofline"]["frac"], d.get("sharded_vs_single"), d["config"]["results"])
    a=d.get("also"); print("also", a["value"], a["e2e"], a["config"]["results"], a["config"]["stage_seconds"]) if a else None
except Exception as e: print("bench parse failed", e)
try:
    d=json.loads(open("gpurun_out/c10_ref_2gpu.json").read().strip().splitlines()[-1])
    print("ref", d["value"], d["ms_per_step"], d["cpu_baseline"]["threads"], d["cpu_baseline"]["cores"])
except Exception as e: print("ref parse failed", e)
PY
