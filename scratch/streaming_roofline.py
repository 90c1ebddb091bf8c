#!/usr/bin/env python
"""Streaming (HBM-bound) kernels of the path at a DCV5T-like size, for the roofline table of DESIGN.md 3.5.

Plain run: times every C-ABI call with CUDA events (gwbse_timer_*) and prints algorithmic bytes, GB/s and the
fraction of the measured HBM copy bandwidth (MEASURED_PEAKS.json).  Under
  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:<names> ...
the same launches give the DRAM traffic per kernel; `ALG` lines name the kernels and their algorithmic bytes so that
scratch/streaming_table.py can join both.  Sizes: B = 41328 (x2 for the full-BSE vectors), k = 30 trial vectors,
Mmn with 144 slices of n = 1249 rows and Naux = 3177 (1/3 of DCV5T's 431 slices: 4.6 GB, the passes are per slice)."""
import ctypes
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from votca_b200.api import Context  # noqa: E402
from votca_b200._capi import ptr  # noqa: E402

try:
    PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    PEAK_SRC = "MEASURED_PEAKS.json"
except Exception:
    PEAK, PEAK_SRC = 6650.0, "fallback of B200_PROFILING.md"

ctx = Context(0)
rng = np.random.default_rng(3)
rows = []


def timed(name, kernels, alg_bytes, fn, reps=3, pre=None):
    """pre: untimed call before every timed one (invalidates a cache the timed call is meant to rebuild)"""
    fn()
    ctx.sync()
    ms = 0.0
    if pre is None:
        ctx.timer_start()
        for _ in range(reps):
            fn()
        ms = ctx.timer_stop_ms() / reps
    else:
        for _ in range(reps):
            pre()
            ctx.sync()
            ctx.timer_start()
            fn()
            ms += ctx.timer_stop_ms() / reps
    gbs = alg_bytes / ms / 1e6
    rows.append((name, kernels, alg_bytes, ms, gbs))
    print(f"ALG {kernels} {alg_bytes:.0f}")
    print(f"{name:38s} {alg_bytes / 1e9:8.3f} GB  {ms:8.3f} ms  {gbs:7.0f} GB/s  {gbs / PEAK:5.2f} of {PEAK:.0f}", flush=True)


# ---------------------------------------------------------------- dense column passes (Davidson vectors)
B2, k = 82656, 30
V = ctx.malloc(B2 * k)
W = ctx.malloc(B2 * k)
ctx.h2d(V, rng.standard_normal((B2, k)))
ctx.h2d(W, rng.standard_normal((B2, k)))
norms = np.empty(k)
ones = np.ones(k)
timed("colnorms  (Gram-Schmidt norms)", "coldots_kernel", 8.0 * B2 * k,
      lambda: ctx.call("gwbse_colnorms_dev", B2, k, V, B2, ptr(norms)))
timed("coldots   (Olsen x.r)", "coldots_kernel", 2 * 8.0 * B2 * k,
      lambda: ctx.call("gwbse_coldots_dev", B2, k, V, B2, W, B2, ptr(norms)))
timed("axpy      (residual r = Aq - l q)", "axpy_kernel", 3 * 8.0 * B2 * k,
      lambda: ctx.call("gwbse_axpy_dev", B2, k, 0.5, V, B2, W, B2))
timed("scale_cols (normalise)", "scale_cols_kernel", 2 * 8.0 * B2 * k,
      lambda: ctx.call("gwbse_scale_cols_dev", B2, k, W, B2, ptr(ones)))
diag = ctx.malloc(B2)
ctx.h2d(diag, rng.uniform(1, 2, B2))
lam = rng.uniform(0.1, 0.2, k)
timed("dpr correction (+ normalise)", "dpr_kernel", 8.0 * (2 * B2 * k + B2) + 3 * 8.0 * B2 * k,
      lambda: ctx.call("gwbse_davidson_correction_dev", B2, k, 0, diag, ptr(lam), V, B2, V, B2, W, B2))
timed("olsen correction (+ normalise)", "olsen_finish_kernel", 8.0 * B2 * k * 12,
      lambda: ctx.call("gwbse_davidson_correction_dev", B2, k, 1, diag, ptr(lam), V, B2, V, B2, W, B2))
for p in (V, W, diag):
    ctx.free(p)

# ---------------------------------------------------------------- Mmn passes
N, naux, homo, nslices = 1249, 3177, 143, 144
ctx.mmn_alloc(naux, 0, nslices - 1, 0, N - 1)
_, mtotal, ntotal, mlocal, npad = ctx.mmn_dims() if hasattr(ctx, "mmn_dims") else (naux, nslices, N, nslices, (N + 15) // 16 * 16)
slab = np.asfortranarray(rng.standard_normal((N, naux)) * 0.02)
for m in range(nslices):
    ctx.mmn_set_slice(m, slab)
mmn_bytes = 8.0 * nslices * npad * naux
timed("mmn_snapshot (D2D copy)", "memcpy", 2 * mmn_bytes, lambda: ctx.call("gwbse_mmn_snapshot"), reps=2)
timed("mmn_restore  (D2D copy)", "memcpy", 2 * mmn_bytes, lambda: ctx.call("gwbse_mmn_restore"), reps=2)

e = np.sort(np.concatenate([rng.uniform(-1.2, -0.25, homo + 1), 0.02 + 3.0 * rng.uniform(0, 1, N - homo - 1) ** 2]))
e[:homo + 1].sort()
weight = rng.uniform(0.1, 1.0, naux)
freq = rng.uniform(0.3, 2.0, naux)
q = nslices
ctx.sigma_ppm_set(weight, freq, e, homo, 0, 0, 1e-3)
levels = np.arange(q, dtype=np.int32)
freqs = e[:q] + 0.01
# term-by-term kernel: one pass over the level's n x Naux block per batch of frequencies
ctx.set_option("sigma_tree_min_terms", 1e18)
timed("sigma_c diag, term by term (1 freq/level)", "sigma_multi_kernel", 8.0 * q * N * naux,
      lambda: ctx.sigma_ppm_eval(levels, freqs), reps=2)
lv8 = np.repeat(levels, 8)
fr8 = np.repeat(e[:q], 8) + np.tile(np.linspace(-0.05, 0.05, 8), q)
gptr = np.arange(0, 8 * q + 1, 8, dtype=np.int32)
out8 = np.empty(8 * q)


def groups():
    ctx.call("gwbse_sigma_eval_groups", 0, q, ptr(levels), ptr(gptr), ptr(fr8), ptr(out8), None)


timed("sigma_c diag, term by term (8 freq/level)", "sigma_multi_kernel", 8.0 * q * N * naux, groups, reps=2)
# treecode: moments are built once per screening update (one pass over the level's block), evaluations walk the tree
ctx.set_option("sigma_tree_min_terms", 0)
ctx.sigma_ppm_set(weight, freq, e, homo, 0, 0, 1e-3)
timed("sigma_c treecode: moments + 8 freq/level", "tree_leaf_moments_kernel|tree_m2m_kernel|tree_eval_kernel",
      8.0 * q * N * naux, groups, reps=2, pre=lambda: ctx.sigma_update_energies(0, e))
timed("sigma_c treecode: 8 freq/level, moments cached", "tree_eval_kernel", 8.0 * q * 8 * 60 * 24, groups, reps=2)
timed("sigma_c offdiag (weights + GEMM; weights alone: ncu list)", "sigma_offdiag_weight_kernel", 2 * 8.0 * q * N * naux,
      lambda: ctx.sigma_ppm_offdiag(e[:q]), reps=1)

# BSE diagonal
vt, ct = 48, 96
Hqp = np.diag(e[:vt + ct])
ctx.bse_configure(vt - 1, 0, 0, vt + ct - 1, rng.uniform(0.3, 1.0, naux), Hqp)
timed("bse diagonal (slice_diag + bse_diag)", "slice_diag_kernel|bse_diag_kernel",
      8.0 * (vt + ct) * naux * 2 + 8.0 * vt * ct * naux, lambda: ctx.bse_diagonal((1, 2, 1, 0)), reps=2)
print("peak source:", PEAK_SRC)
ctx.close()
