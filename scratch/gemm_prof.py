import sys, numpy as np
sys.path.insert(0, '.')
from votca_b200.api import Context
ctx = Context(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
rng = np.random.default_rng(0)
dA = ctx.upload(rng.standard_normal((n, n))); dB = ctx.upload(rng.standard_normal((n, n))); dC = ctx.malloc(n * n)
for (ta, tb) in (('T', 'N'), ('N', 'N')):
    for r in range(2):
        ctx.dgemm(ta, tb, n, n, n, 1.0, dA, n, dB, n, 0.0, dC, n, 0, 1)
    ctx.sync()
    ctx.timer_start()
    for r in range(3):
        ctx.dgemm(ta, tb, n, n, n, 1.0, dA, n, dB, n, 0.0, dC, n, 0, 1)
    ms = ctx.timer_stop_ms() / 3
    print(f'dgemm {ta}{tb} n={n}: {ms:.3f} ms {2*n**3/ms/1e9:.2f} TFLOP/s')
