#!/bin/bash
# unrestricted exact / cda evaluators on the device + regression of the restricted exact / cda paths they were factored out of
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_zz_gpu_uks.py tests/test_gpu_host.py tests/test_gpu_zzz_large_pipeline.py -m gpu -x -q --durations=6 -k "not synthetic_pipeline" > gpurun_out/c27_pytest.log 2>&1; echo "pytest rc=$?"; tail -14 gpurun_out/c27_pytest.log
