"""Short-K GEMM of the BSE Hd intermediate (M = vt*k, N = Naux*chunk, K = ct), K-major operands at an EVEN pitch as the
library keeps them (trial vectors re-pitched, Mmn rows padded), for ncu:
   ncu --set full --clock-control none --import-source on -k regex:gemm_dmma_kernel -s 3 -c 1 -o gpurun_out/x python scratch/ncu_gemm_shortk.py [ct]"""
import sys
sys.path.insert(0, '.')
from votca_b200.api import Context
ctx = Context(0)
k = int(sys.argv[1]) if len(sys.argv) > 1 else 359
m, n, ld = 5400, 145920, (k + 1) // 2 * 2
A = ctx.malloc(m * ld); B = ctx.malloc(ld * n); C = ctx.malloc(m * n)
for r in range(3):
    ctx.dgemm('T', 'N', m, n, k, 1.0, A, ld, B, ld, 0.0, C, m, -1, 0)
ctx.sync()
ctx.timer_start()
for r in range(5):
    ctx.dgemm('T', 'N', m, n, k, 1.0, A, ld, B, ld, 0.0, C, m, -1, 0)
ms = ctx.timer_stop_ms() / 5
print(f'dgemm TN {m}x{n}x{k} (pitch {ld}): {ms:.3f} ms {2*m*n*k/ms/1e9:.2f} TFLOP/s')
for cfg in (4, 10, 11, 14):
    ctx.dgemm('T', 'N', m, n, k, 1.0, A, ld, B, ld, 0.0, C, m, cfg, 1)
    ctx.sync()
    ctx.timer_start()
    for r in range(5):
        ctx.dgemm('T', 'N', m, n, k, 1.0, A, ld, B, ld, 0.0, C, m, cfg, 1)
    ms = ctx.timer_stop_ms() / 5
    print(f'  cfg{cfg}: {ms:.3f} ms {2*m*n*k/ms/1e9:.2f} TFLOP/s')
