"""The weighted RPA SYRK (epsilon(i w), rpa.cc:75-127) at DCV5T size for an ncu capture of the TMA kernel with weights:
   ncu --set full --clock-control none --import-source on -k regex:gemm_tma_kernel -s 1 -c 1 -o gpurun_out/x python scratch/ncu_eps.py"""
import sys
sys.path.insert(0, '.')
import numpy as np
from votca_b200.api import Context
ctx = Context(0)
N, naux, homo = 1249, 3177, 143
rng = np.random.default_rng(5)
ctx.mmn_alloc(naux, 0, homo, 0, N - 1)
slab = np.asfortranarray(rng.standard_normal((N, naux)) * 0.02)
for m in range(homo + 1):
    ctx.mmn_set_slice(m, slab)
e = np.sort(np.concatenate([rng.uniform(-1.2, -0.25, homo + 1), 0.02 + 3.0 * rng.uniform(0, 1, N - homo - 1) ** 2]))
for r in range(2):
    ctx.rpa_epsilon(0, 0.5, 1e-3, e, homo, 0, N - 1, fetch=False)
ctx.sync()
ctx.timer_start()
for r in range(3):
    ctx.rpa_epsilon(0, 0.5, 1e-3, e, homo, 0, N - 1, fetch=False)
ms = ctx.timer_stop_ms() / 3
S = (homo + 1) * (N - homo - 1)
print(f'epsilon(iw) S={S} Naux={naux}: {ms:.3f} ms  {S * naux * (naux + 1) / ms / 1e9:.2f} TFLOP/s (SYRK count)')
ctx.close()
