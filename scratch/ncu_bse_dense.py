"""Build of a materialised BSE block at a mid size (vt = 64, ct = 128, Naux = 4560: one 4096 x 16384 x 4560 GEMM with both
operands M-major and the scattered C addressing of capi_bse.cu dense_build) + its skinny product, for timing and for
   ncu --set full --clock-control none --import-source on -k regex:gemm_tma_kernel -c 1 -o gpurun_out/x python scratch/ncu_bse_dense.py
GWBSE_DENSE_BUILD_CFG=10|11|12|14 forces the tile shape of the build GEMM."""
import os, sys, time
sys.path.insert(0, '.')
import numpy as np
from votca_b200.api import Context
ctx = Context(0)
naux, vt, ct = 4560, 64, 128
mt = vt + ct
rng = np.random.default_rng(0)
ctx.mmn_alloc(naux, 0, mt - 1, 0, mt - 1)
ctx.mmn_set_all(rng.standard_normal((mt, mt, naux)))
ctx.set_option("bse_dense", 2)
Hqp = np.eye(mt)
X1 = rng.standard_normal((vt * ct, 1))
X30 = rng.standard_normal((vt * ct, 30))
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
for kind, co in (("Hd", (0, 0, 1, 0)), ("Hd2", (0, 0, 0, 1))):
    for r in range(reps):
        ctx.bse_configure(vt - 1, 0, 0, mt - 1, rng.uniform(0.3, 1.0, naux), Hqp)  # new screening -> new block
        ctx.sync()
        ctx.timer_start()
        ctx.bse_matmul(co, X1)
        ms = ctx.timer_stop_ms()
        if r == reps - 1:
            B = vt * ct
            print(f"{kind}: build + 1-column product {ms:.2f} ms -> {2.0 * B * B * naux / ms / 1e9:.2f} TFLOP/s "
                  f"(cfg {os.environ.get('GWBSE_DENSE_BUILD_CFG', 'auto')})", flush=True)
    ctx.timer_start()
    for r in range(5):
        ctx.bse_matmul(co, X30)
    print(f"{kind}: 30-column product from the resident block incl. H2D/D2H of the vectors {ctx.timer_stop_ms() / 5:.3f} ms", flush=True)
print(ctx.bse_dense_stats())
