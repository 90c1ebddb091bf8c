#!/usr/bin/env python
"""Device check + timing of the TMA-staged DMMA GEMM against NumPy and against the cp.async kernel.
  python scratch/gemm_tma_check.py [--quick]"""
import sys
import numpy as np
sys.path.insert(0, '.')
from votca_b200.api import Context

ctx = Context(0)
rng = np.random.default_rng(5)


import ctypes


def check(ta, tb, m, n, k, cfg, splitk=0, pad=0, shift=0):
    A = rng.standard_normal((m, k))
    B = rng.standard_normal((k, n))
    lda = (k if ta == 'T' else m) + pad
    ldb = (n if tb == 'T' else k) + pad
    Ah = np.zeros((lda, m if ta == 'T' else k), order='F')
    Bh = np.zeros((ldb, k if tb == 'T' else n), order='F')
    if ta == 'T':
        Ah[:k, :] = A.T
    else:
        Ah[:m, :] = A
    if tb == 'T':
        Bh[:n, :] = B.T
    else:
        Bh[:k, :] = B
    if shift:  # operands that start 8 bytes into their (16-byte aligned) allocations
        bA = ctx.upload(np.concatenate([[9.0], Ah.ravel(order='F')]))
        bB = ctx.upload(np.concatenate([[9.0], Bh.ravel(order='F')]))
        dA, dB = ctypes.c_void_p(bA.value + 8), ctypes.c_void_p(bB.value + 8)
    else:
        bA = dA = ctx.upload(Ah)
        bB = dB = ctx.upload(Bh)
    ldc = m + pad
    C0 = np.asfortranarray(rng.standard_normal((ldc, n)))
    dC = ctx.upload(C0)
    ctx.dgemm(ta, tb, m, n, k, 0.7, dA, lda, dB, ldb, 0.3, dC, ldc, cfg, splitk)
    full = ctx.download(dC, (ldc, n))
    got = full[:m]
    ref = 0.7 * A @ B + 0.3 * C0[:m]
    err = np.abs(got - ref).max() / max(1.0, np.abs(ref).max())
    untouched = np.array_equal(full[m:], C0[m:])
    for p in (bA, bB, dC):
        ctx.free(p)
    return err, untouched


def bench(ta, tb, m, n, k, cfg, splitk=0, reps=3):
    A = ctx.malloc(m * k); B = ctx.malloc(k * n); C = ctx.malloc(m * n)
    lda = k if ta == 'T' else m; ldb = n if tb == 'T' else k
    ctx.dgemm(ta, tb, m, n, k, 1.0, A, lda, B, ldb, 0.0, C, m, cfg, splitk); ctx.sync()
    ctx.timer_start()
    for r in range(reps):
        ctx.dgemm(ta, tb, m, n, k, 1.0, A, lda, B, ldb, 0.0, C, m, cfg, splitk)
    ms = ctx.timer_stop_ms() / reps
    for p in (A, B, C):
        ctx.free(p)
    return 2.0 * m * n * k / ms / 1e9


bad = 0
for cfg in (10, 11, 12, 13, 14, -1):
    for ta in 'NT':
        for tb in 'NT':
            for (m, n, k) in ((300, 200, 150), (130, 66, 18), (1000, 38, 514), (64, 64, 16), (16, 2, 1030), (258, 514, 34)):
                for sk in (0, 3):
                    for pad, shift in ((0, 0), (2, 0), (2, 1)):
                        err, unt = check(ta, tb, m, n, k, cfg, sk, pad, shift)
                        if err > 1e-12 or not unt:
                            bad += 1
                            print(f"FAIL cfg={cfg} {ta}{tb} {m}x{n}x{k} sk={sk} pad={pad} shift={shift}: err={err:.2e} untouched={unt}", flush=True)
    print(f"cfg {cfg}: checked, failures so far {bad}", flush=True)
# odd leading dimensions must fall back to the cp.async kernel and still be right
for (m, n, k) in ((301, 201, 151), (129, 65, 17)):
    err, unt = check('N', 'N', m, n, k, -1, 0, 1)
    print(f"odd pitch {m}x{n}x{k}: err={err:.2e}")
    bad += err > 1e-12
print("correctness failures:", bad, flush=True)
if bad or "--quick" in sys.argv:
    sys.exit(1 if bad else 0)

shapes = [("NN", 4096, 4096, 4096), ("TN", 4096, 4096, 4096), ("NN", 75776, 3177 + 1, 3178), ("TN", 4320, 50832, 288),
          ("TN", 3178, 3178, 40000), ("TN", 288, 1920, 114688), ("TN", 144, 1920, 114688), ("NN", 41328, 30, 3178),
          ("TN", 3178, 30, 41328), ("TN", 1250, 432, 1250)]
for cfg in (4, 0, 10, 11, 12, 13, 14, -1):
    row = []
    for (t, m, n, k) in shapes:
        try:
            row.append("%6.2f" % bench(t[0], t[1], m, n, k, cfg))
        except Exception as e:
            row.append("  err ")
    print("cfg %3d: " % cfg + " ".join(row), flush=True)
print("shapes: " + " ".join(f"{t}{m}x{n}x{k}" for (t, m, n, k) in shapes))
