"""Short-K GEMM of the BSE Hd intermediate (M = vt*k, N = Naux*chunk, K = ct = 287), plain addressing, for ncu."""
import sys
sys.path.insert(0, '.')
from votca_b200.api import Context
ctx = Context(0)
m, n, k = 4320, 203328, 287
A = ctx.malloc(m * k); B = ctx.malloc(k * n); C = ctx.malloc(m * n)
for r in range(3):
    ctx.dgemm('T', 'N', m, n, k, 1.0, A, k, B, k, 0.0, C, m, -1, 0)
ctx.sync()
ctx.timer_start()
for r in range(5):
    ctx.dgemm('T', 'N', m, n, k, 1.0, A, k, B, k, 0.0, C, m, -1, 0)
ms = ctx.timer_stop_ms() / 5
print(f'dgemm TN {m}x{n}x{k}: {ms:.3f} ms {2*m*n*k/ms/1e9:.2f} TFLOP/s')
for cfg in (0, 3, 4, 5):
    ctx.dgemm('T', 'N', m, n, k, 1.0, A, k, B, k, 0.0, C, m, cfg, 1)
    ctx.sync()
    ctx.timer_start()
    for r in range(5):
        ctx.dgemm('T', 'N', m, n, k, 1.0, A, k, B, k, 0.0, C, m, cfg, 1)
    ms = ctx.timer_stop_ms() / 5
    print(f'  cfg{cfg}: {ms:.3f} ms {2*m*n*k/ms/1e9:.2f} TFLOP/s')
