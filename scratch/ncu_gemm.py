"""One MultiplyRight-shaped DGEMM (M-major A x K-major B, the dominant launch of the GW-BSE path) for an ncu capture:
   ncu --set full --clock-control none --import-source on -k regex:gemm_dmma -s 2 -c 1 -o gpurun_out/x python scratch/ncu_gemm.py"""
import sys
sys.path.insert(0, '.')
from votca_b200.api import Context
ctx = Context(0)
m, n, k = 148 * 2 * 128 * 2, 3177, 3177
A = ctx.malloc(m * k); B = ctx.malloc(k * n); C = ctx.malloc(m * n)
for r in range(3):
    ctx.dgemm('N', 'N', m, n, k, 1.0, A, m, B, k, 0.0, C, m, -1, 0)
ctx.sync()
ctx.timer_start()
for r in range(3):
    ctx.dgemm('N', 'N', m, n, k, 1.0, A, m, B, k, 0.0, C, m, -1, 0)
ms = ctx.timer_stop_ms() / 3
print(f'dgemm NN {m}x{n}x{k}: {ms:.3f} ms {2*m*n*k/ms/1e9:.2f} TFLOP/s; algorithmic bytes {8*(m*k+k*n+m*n)}')
