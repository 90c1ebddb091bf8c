#!/bin/bash
mkdir -p gpurun_out
timeout 500 python scratch/debug_fill_overlap.py overlap > gpurun_out/c18_overlap.log 2>&1; tail -4 gpurun_out/c18_overlap.log
GWBSE_NO_FILL_OVERLAP=1 timeout 500 python scratch/debug_fill_overlap.py serial > gpurun_out/c18_serial.log 2>&1; tail -4 gpurun_out/c18_serial.log
python - <<'PY'
import numpy as np
a=np.load("gpurun_out/c18_overlap.npz"); b=np.load("gpurun_out/c18_serial.npz")
for r in range(3):
    print("rep",r,"max|dSx|", np.abs(a[f"sx{r}"]-b["sx0"]).max(), "serial self", np.abs(b[f"sx{r}"]-b["sx0"]).max(), "max|dQP|", np.abs(a[f"qp{r}"]-b["qp0"]).max())
d=np.abs(a["sx0"]-b["sx0"]); i=np.unravel_index(d.argmax(), d.shape); print("worst element", i, a["sx0"][i], b["sx0"][i])
print("rows with diff>1e-9:", np.where(d.max(axis=1)>1e-9)[0][:40], "count", (d.max(axis=1)>1e-9).sum())
PY
