"""Reproducer: is gemm_tma_kernel correct while another context's grids share the GPU?
Context A repeats a batched fill-shaped GEMM (TMA kernel) and compares every result bitwise with its own serial result;
context B (second host thread) meanwhile runs large GEMMs / symmetric eigensolves.  argv[1]: what B runs
(gemm | eig | none)."""
import sys, threading, time
sys.path.insert(0, '.')
import numpy as np
from votca_b200.api import Context
what = sys.argv[1] if len(sys.argv) > 1 else "gemm"
A, B = Context(0), Context(0)
rng = np.random.default_rng(1)
m, n, k = 1860, 539 * 8, 1860        # 8 "aux functions" side by side
a = A.malloc(k * m); b = A.malloc(k * n); c = A.malloc(m * n)
A.h2d(a, rng.standard_normal((k, m))); A.h2d(b, rng.standard_normal((k, n)))
def run_a():
    A.dgemm('T', 'N', m, n, k, 1.0, a, k, b, k, 0.0, c, m, -1, 0)
    A.sync()
    return A.download(c, (m, n)).copy()
ref = run_a()
assert np.array_equal(ref, run_a())
nb = 4560
x = B.malloc(nb * nb); y = B.malloc(nb * nb); z = B.malloc(nb * nb)
S = rng.standard_normal((nb, nb)); S = S @ S.T / nb + np.eye(nb)
B.h2d(x, S); B.h2d(y, S)
stop = False
def busy():
    while not stop:
        if what == "gemm":
            B.dgemm('N', 'N', nb, nb, nb, 1.0, x, nb, y, nb, 0.0, z, nb, -1, 0)
        elif what == "eig":
            B.h2d(z, S)
            w = np.empty(nb)
            B.call("gwbse_sym_eig_dev", nb, z, nb, w.ctypes.data_as(__import__('ctypes').c_void_p))
        else:
            time.sleep(0.01)
    B.sync()
t = threading.Thread(target=busy); t.start()
time.sleep(0.3)
bad = 0
for rep in range(120):
    got = run_a()
    if not np.array_equal(got, ref):
        d = np.abs(got - ref)
        cols = np.where(d.max(axis=0) > 0)[0]
        bad += 1
        if bad <= 3:
            rows = np.where(d.max(axis=1) > 0)[0]
            flat = np.flatnonzero(d.reshape(-1, order="F") > 0)
            print(f"rep {rep}: MISMATCH max {d.max():.3e}, {len(cols)} bad columns, first {cols[:4]} last {cols[-4:]}; "
                  f"{len(rows)} bad rows [{rows[0]}..{rows[-1]}]; bad elements {len(flat)}, flat range [{flat[0]}, {flat[-1]}] "
                  f"span {flat[-1] - flat[0] + 1}; got sample {got.reshape(-1, order='F')[flat[:3]]} ref {ref.reshape(-1, order='F')[flat[:3]]}", flush=True)
stop = True; t.join()
print(f"B runs {what}: {bad} of 120 repetitions of A differ from the serial result", flush=True)
