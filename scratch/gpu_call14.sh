#!/bin/bash
# full GPU suite with the UKS twins, NVTX ranges in the library
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --durations=5 > gpurun_out/c14_pytest.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/c14_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c14_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/c14_smoke.log
