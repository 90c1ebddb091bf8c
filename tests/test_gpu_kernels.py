"""GPU parity tests of the C ABI kernels against the CPU oracle (run with -m gpu on the B200).

FP64 contractions are compared at 1e-11 relative Frobenius error (BASELINE.json asks 1e-10 for Mmn);
the only difference to the oracle is summation order.
"""
import os

import numpy as np
import pytest

from oracle import bse_operator as bop
from oracle import rpa as orpa
from oracle import sigma as osig
from oracle import threecenter
from oracle.davidson import DavidsonSolver
from tests.helpers import methane_mmn, rel_frob

pytestmark = pytest.mark.gpu

TOL = 1e-11


@pytest.fixture(scope="module")
def ctx():
    from votca_b200.api import Context
    c = Context(0)
    yield c
    c.close()


def random_tc(rng, naux=70, mtotal=30, ntotal=45):
    tc = threecenter.TCMatrix(naux, 0, mtotal - 1, 0, ntotal - 1)
    tc.M = rng.standard_normal((mtotal, ntotal, naux)) / np.sqrt(naux)
    return tc


def push(ctx, tc):
    ctx.mmn_alloc(tc.naux, tc.mmin, tc.mmax, tc.nmin, tc.nmax)
    ctx.mmn_set_all(tc.M)


# ------------------------------------------------------------------ GEMM
@pytest.mark.parametrize("cfg", [-1, 0, 1, 2])
@pytest.mark.parametrize("ta,tb", [("N", "N"), ("T", "N"), ("N", "T"), ("T", "T")])
def test_dgemm_transposes(ctx, cfg, ta, tb):
    rng = np.random.default_rng(1)
    for (m, n, k) in [(1, 1, 1), (7, 5, 3), (130, 67, 41), (257, 129, 100), (64, 300, 17), (33, 20, 515)]:
        A = rng.standard_normal((k, m) if ta == "T" else (m, k))
        B = rng.standard_normal((n, k) if tb == "T" else (k, n))
        C0 = rng.standard_normal((m, n))
        ref = 0.7 * (A.T if ta == "T" else A) @ (B.T if tb == "T" else B) - 0.3 * C0
        out = ctx.gemm_host(ta, tb, 0.7, A, B, -0.3, C0, cfg=cfg)
        assert rel_frob(ref, out) < TOL, (m, n, k)


@pytest.mark.parametrize("splitk", [1, 2, 5])
def test_dgemm_splitk_and_beta(ctx, splitk):
    rng = np.random.default_rng(2)
    A = rng.standard_normal((150, 900))
    B = rng.standard_normal((900, 140))
    C0 = rng.standard_normal((150, 140))
    for cfg in (0, 1, 2):
        out = ctx.gemm_host("N", "N", 1.0, A, B, 1.0, C0, cfg=cfg, splitk=splitk)
        assert rel_frob(A @ B + C0, out) < TOL


def test_dgemm_empty_k(ctx):
    C0 = np.arange(12.0).reshape(3, 4)
    out = ctx.gemm_host("N", "N", 1.0, np.zeros((3, 0)), np.zeros((0, 4)), 0.0, C0)
    assert np.all(out == 0.0)


def test_dgemm_shape_error(ctx):
    from votca_b200.api import GwbseError
    d = ctx.malloc(16)
    with pytest.raises(GwbseError, match="Shape mismatch"):
        ctx.dgemm("N", "N", 4, 4, 4, 1.0, d, 2, d, 4, 0.0, d, 4)
    ctx.free(d)


# ------------------------------------------------------------------ dense auxiliaries
def test_dense_aux(ctx):
    rng = np.random.default_rng(3)
    A = rng.standard_normal((37, 37))
    S = A @ A.T + 37 * np.eye(37)
    w, V = ctx.sym_eig(S)
    assert np.allclose(w, np.linalg.eigvalsh(S), rtol=1e-12)
    assert rel_frob(S, (V * w) @ V.T) < 1e-12
    assert rel_frob(np.linalg.inv(A), ctx.inverse(A)) < 1e-9
    b = rng.standard_normal((37, 3))
    assert rel_frob(np.linalg.solve(A, b), ctx.lu_solve(A, b)) < 1e-9
    T = rng.standard_normal((12, 12))
    Bm = rng.standard_normal((12, 12)) + 5 * np.eye(12)
    wr, wi, VR = ctx.gen_eig(T, Bm)
    import scipy.linalg
    wref = scipy.linalg.eigvals(T, Bm)
    key = lambda z: (np.round(z.real, 8), np.round(z.imag, 8))  # noqa: E731
    got = np.array(sorted(wr + 1j * wi, key=key))
    assert np.allclose(got, np.array(sorted(wref, key=key)), rtol=1e-9, atol=1e-10)
    for i in range(12):  # real eigenpairs satisfy T x = lambda B x
        if wi[i] == 0.0:
            x = VR[:, i]
            assert np.linalg.norm(T @ x - wr[i] * (Bm @ x)) < 1e-9 * np.linalg.norm(x)


# ------------------------------------------------------------------ Mmn
def test_mmn_fill_and_mulright(ctx):
    rng = np.random.default_rng(4)
    N, naux = 41, 53
    mos = rng.standard_normal((N, N))
    ao = rng.standard_normal((naux, N, N))
    ao = ao + ao.transpose(0, 2, 1)
    tc = threecenter.TCMatrix(naux, 2, 24, 2, N - 1)
    tc.fill_3c_mo(ao, mos)
    ctx.mmn_alloc(naux, 2, 24, 2, N - 1)
    ctx.mmn_set_mos(mos)
    ctx.mmn_fill_block(0, ao[:20])
    ctx.mmn_fill_block(20, ao[20:])
    assert rel_frob(tc.M, ctx.mmn_get_all()) < TOL
    R = rng.standard_normal((naux, naux))
    tc.multiply_right(R)
    ctx.mmn_mul_right(R)
    assert rel_frob(tc.M, ctx.mmn_get_all()) < TOL
    # snapshot / restore (Rebuild)
    ctx.call("gwbse_mmn_snapshot")
    ctx.mmn_mul_right(R)
    ctx.call("gwbse_mmn_restore")
    assert rel_frob(tc.M, ctx.mmn_get_all()) < TOL


def test_mmn_rotate(ctx):
    """TCMatrix_gwbse::Rotate (threecenter.cc:108-131): QP-window rows of the QP-window slices only."""
    rng = np.random.default_rng(41)
    tc = random_tc(rng, naux=37, mtotal=22, ntotal=31)
    tc.mmin, tc.mmax, tc.nmin, tc.nmax = 1, 22, 0, 30
    push(ctx, tc)
    qpmin, qpmax = 3, 17
    U, _ = np.linalg.qr(rng.standard_normal((qpmax - qpmin + 1,) * 2))
    tc.rotate(U, qpmin, qpmax)
    ctx.mmn_rotate(U, qpmin, qpmax)
    assert rel_frob(tc.M, ctx.mmn_get_all()) < TOL
    with pytest.raises(RuntimeError):
        ctx.mmn_rotate(U, 0, 14)  # window below mmin


def test_mmn_golden_threecenter(ctx, golden, methane):
    """test_threecenter_gwbse.cc:36-126 through the CUDA path (AO integrals from the host)."""
    mos = golden["threecenter_gwbse/MOs"]
    ctx.mmn_alloc(17, 0, 5, 0, 7)
    ctx.mmn_set_mos(mos)
    ctx.mmn_fill_block(0, methane["ao3c"])
    L, removed = ctx.pseudo_invsqrt(methane["S"], methane["V"])
    Lref, rref = threecenter.pseudo_invsqrt_gwbse(methane["S"], methane["V"])
    assert removed == rref
    assert rel_frob(Lref, L) < 1e-9
    ctx.mmn_mul_right(L)
    for i, name in [(0, "ref0b"), (2, "ref2b"), (4, "ref4b")]:
        assert rel_frob(golden["threecenter_gwbse/" + name], ctx.mmn_get_slice(i)) < 1e-5
    ctx.mmn_mul_right(np.eye(17))
    for i, name in [(0, "ref0b"), (2, "ref2b"), (4, "ref4b")]:
        assert rel_frob(golden["threecenter_gwbse/" + name], ctx.mmn_get_slice(i)) < 1e-5


# ------------------------------------------------------------------ RPA
def test_rpa_epsilon_golden(ctx, golden):
    tc = methane_mmn(golden["rpa/eigenvectors"])
    push(ctx, tc)
    e = golden["rpa/eigenvals"].ravel()
    assert rel_frob(golden["rpa/i_ref"], ctx.rpa_epsilon(0, 0.5, 1e-4, e, 4, 0, 16)) < 1e-4
    assert rel_frob(golden["rpa/r_ref"], ctx.rpa_epsilon(1, 0.0, 1e-4, e, 4, 0, 16)) < 1e-4
    assert rel_frob(golden["rpa/r_complex_ref"], ctx.rpa_epsilon(2, complex(0.5, 0.5), 1e-4, e, 4, 0, 16)) < 1e-4


@pytest.mark.parametrize("shape", [(70, 30, 45, 10), (200, 40, 161, 23)])
def test_rpa_epsilon_random(ctx, shape):
    naux, mtotal, ntotal, homo = shape
    rng = np.random.default_rng(5)
    tc = random_tc(rng, naux, mtotal, ntotal)
    push(ctx, tc)
    e = np.sort(rng.uniform(-1, 2, ntotal))
    r = orpa.RPA(tc)
    r.configure(homo, 0, ntotal - 1)
    r.set_rpa_input_energies(e)
    assert rel_frob(r.calculate_epsilon_i(0.5), ctx.rpa_epsilon(0, 0.5, r.ETA, e, homo, 0, ntotal - 1)) < TOL
    assert rel_frob(r.calculate_epsilon_r(0.3), ctx.rpa_epsilon(1, 0.3, r.ETA, e, homo, 0, ntotal - 1)) < TOL
    w = complex(0.2, 0.4)
    assert rel_frob(r.calculate_epsilon_r(w), ctx.rpa_epsilon(2, w, r.ETA, e, homo, 0, ntotal - 1)) < TOL
    assert rel_frob(np.tril(r.h2p_apb()), np.tril(ctx.rpa_h2p_apb(e, homo, 0, ntotal - 1))) < TOL


# ------------------------------------------------------------------ Sigma
def _sigma_setup(ctx, rng, kind, naux=60, mtotal=26, ntotal=37, homo=8, qpmin=2, qpmax=22):
    tc = random_tc(rng, naux, mtotal, ntotal)
    e = np.sort(np.concatenate([rng.uniform(-1.2, -0.3, homo + 1), rng.uniform(0.05, 2.5, ntotal - homo - 1)]))
    r = orpa.RPA(tc)
    r.configure(homo, 0, ntotal - 1)
    r.set_rpa_input_energies(e)
    s = osig.create(kind, tc, r)
    s.configure(osig.SigmaOptions(homo=homo, qpmin=qpmin, qpmax=qpmax, rpamin=0, rpamax=ntotal - 1, eta=1e-3))
    return tc, e, r, s


def test_sigma_x(ctx):
    rng = np.random.default_rng(6)
    tc, e, r, s = _sigma_setup(ctx, rng, "ppm")
    push(ctx, tc)
    assert rel_frob(s.calc_exchange_matrix(), ctx.sigma_x(8, 0, 2, 22)) < TOL


def test_sigma_ppm(ctx):
    rng = np.random.default_rng(7)
    tc, e, r, s = _sigma_setup(ctx, rng, "ppm")
    s.prepare_screening()  # rotates tc.M in place; push the rotated tensor
    push(ctx, tc)
    ctx.sigma_ppm_set(s.ppm_weight, s.ppm_freq, e, 8, 0, 2, 1e-3)
    q = s.qptotal
    levels = np.repeat(np.arange(q), 3)
    freqs = np.tile([-0.45, 0.1, 0.77], q) + 0.01 * levels
    sig, dsig = ctx.sigma_ppm_eval(levels, freqs, deriv=True)
    ref = np.array([s.calc_correlation_diag_element(l, w) for l, w in zip(levels, freqs)])
    dref = np.array([s.calc_correlation_diag_element_derivative(l, w) for l, w in zip(levels, freqs)])
    assert rel_frob(ref, sig) < 1e-10
    assert rel_frob(dref, dsig) < 1e-10
    fq = e[2:2 + q] + 0.05
    assert rel_frob(s.calc_correlation_offdiag(fq), ctx.sigma_ppm_offdiag(fq)) < 1e-10


def test_sigma_exact(ctx):
    rng = np.random.default_rng(8)
    tc, e, r, s = _sigma_setup(ctx, rng, "exact", naux=40, mtotal=20, ntotal=20, homo=5, qpmin=1, qpmax=18)
    tc.M *= 0.2  # keep the RPA matrix positive definite
    push(ctx, tc)
    omega, XpY, _ = r.diagonalize_h2p()
    s.rpa_omegas = omega
    s.residues = [s._calc_residues(i, XpY) for i in range(s.qptotal)]
    ctx.sigma_exact_prepare(omega, XpY, e, 5, 0, 19, 1, 18, 1e-3)
    q = s.qptotal
    levels = np.repeat(np.arange(q), 2)
    freqs = np.tile([-0.5, 0.6], q) + 0.013 * levels
    sig, dsig = ctx.sigma_exact_eval(levels, freqs, deriv=True)
    ref = np.array([s.calc_correlation_diag_element(l, w) for l, w in zip(levels, freqs)])
    dref = np.array([s.calc_correlation_diag_element_derivative(l, w) for l, w in zip(levels, freqs)])
    assert rel_frob(ref, sig) < 1e-10
    assert rel_frob(dref, dsig) < 1e-10
    fq = e[1:1 + q] + 0.02
    assert rel_frob(s.calc_correlation_offdiag(fq), ctx.sigma_exact_offdiag(fq)) < 1e-10


def _direct_sigma(M, fac, pole, e, nocc, eta, pref, level, w):
    """Term-by-term sum of sigma_ppm.cc:37-91 / sigma_exact.cc:40-83 in NumPy (M: level x n x poles)."""
    a = np.where(np.arange(M.shape[1])[:, None] < nocc, e[:, None] - pole[None, :], e[:, None] + pole[None, :])
    r = fac[None, :] * M[level] ** 2
    t = w - a
    den = t * t + eta * eta
    return pref * (r * t / den).sum(), pref * (r * (eta * eta - t * t) / den ** 2).sum()


@pytest.mark.parametrize("tree_bytes", [8 << 30, 1 << 20])
def test_sigma_tree_ppm(ctx, tree_bytes):
    """Treecode Sigma_c (sigma_tree.cu) against the term-by-term sum, incl. frequencies on and next to poles,
    and with a moment store too small for all levels (slots recycled)."""
    rng = np.random.default_rng(70)
    naux, mtotal, ntotal, homo = 300, 24, 140, 30
    tc = random_tc(rng, naux, mtotal, ntotal)
    e = np.sort(np.concatenate([rng.uniform(-1.2, -0.3, homo + 1), rng.uniform(0.05, 2.5, ntotal - homo - 1)]))
    weight = rng.uniform(0.05, 1.0, naux)
    weight[::17] = 0.0  # skipped poles (sigma_ppm.cc:47-52)
    freq = rng.uniform(0.3, 6.0, naux)
    push(ctx, tc)
    ctx.set_option("sigma_tree_min_terms", 0)
    ctx.set_option("sigma_tree_bytes", tree_bytes)
    try:
        ctx.sigma_ppm_set(weight, freq, e, homo, 0, 0, 1e-3)
        fac = np.where(weight < 1e-9, 0.0, weight * freq)
        levels = np.repeat(np.arange(mtotal), 5)
        freqs = rng.uniform(-2.0, 4.0, levels.size)
        freqs[0] = e[3] - freq[1]        # exactly on a pole
        freqs[1] = e[50] + freq[2] + 1e-5  # next to one
        freqs[2] = 40.0                   # far outside the pole range
        sig, dsig = ctx.sigma_ppm_eval(levels, freqs, deriv=True)
        ref = np.array([_direct_sigma(tc.M, fac, freq, e, homo + 1, 1e-3, 0.5, l, w) for l, w in zip(levels, freqs)])
        scale = np.abs(ref[:, 0]).max()
        assert np.abs(ref[:, 0] - sig).max() < 1e-11 * scale
        assert np.abs(ref[:, 1] - dsig).max() < 1e-10 * np.abs(ref[:, 1]).max()
        sig2 = ctx.sigma_ppm_eval(levels[::-1].copy(), freqs[::-1].copy())  # batching does not change a result
        assert np.array_equal(sig2[::-1], sig)
        # the term-by-term kernel agrees as well
        ctx.set_option("sigma_tree_min_terms", 1e18)
        sig3 = ctx.sigma_ppm_eval(levels, freqs)
        assert np.abs(sig3 - sig).max() < 1e-11 * scale
        # new energies move the poles: the tree is rebuilt
        ctx.set_option("sigma_tree_min_terms", 0)
        e2 = e + 0.01 * rng.standard_normal(ntotal)
        ctx.sigma_update_energies(0, e2)
        sig4 = ctx.sigma_ppm_eval(levels, freqs)
        ref4 = np.array([_direct_sigma(tc.M, fac, freq, e2, homo + 1, 1e-3, 0.5, l, w)[0] for l, w in zip(levels, freqs)])
        assert np.abs(ref4 - sig4).max() < 1e-11 * np.abs(ref4).max()
    finally:
        ctx.set_option("sigma_tree_min_terms", 1 << 15)
        ctx.set_option("sigma_tree_bytes", 8 << 30)


def test_sigma_tree_exact(ctx):
    rng = np.random.default_rng(8)
    tc, e, r, s = _sigma_setup(ctx, rng, "exact", naux=40, mtotal=20, ntotal=20, homo=5, qpmin=1, qpmax=18)
    tc.M *= 0.2
    push(ctx, tc)
    omega, XpY, _ = r.diagonalize_h2p()
    s.rpa_omegas = omega
    s.residues = [s._calc_residues(i, XpY) for i in range(s.qptotal)]
    ctx.set_option("sigma_tree_min_terms", 0)
    try:
        ctx.sigma_exact_prepare(omega, XpY, e, 5, 0, 19, 1, 18, 1e-3)
        q = s.qptotal
        levels = np.repeat(np.arange(q), 2)
        freqs = np.tile([-0.5, 0.6], q) + 0.013 * levels
        sig, dsig = ctx.sigma_exact_eval(levels, freqs, deriv=True)
        ref = np.array([s.calc_correlation_diag_element(l, w) for l, w in zip(levels, freqs)])
        dref = np.array([s.calc_correlation_diag_element_derivative(l, w) for l, w in zip(levels, freqs)])
        assert rel_frob(ref, sig) < 1e-10
        assert rel_frob(dref, dsig) < 1e-10
    finally:
        ctx.set_option("sigma_tree_min_terms", 1 << 15)


def test_sigma_tree_full_size(ctx):
    """DCV5T-sized pole set (n = 1249, Naux = 3177, 3.97 M terms per level): treecode == term-by-term kernel."""
    rng = np.random.default_rng(71)
    naux, mtotal, ntotal, homo = 3177, 3, 1249, 143
    tc = random_tc(rng, naux, mtotal, ntotal)
    e = np.sort(np.concatenate([rng.uniform(-1.2, -0.25, homo + 1), 0.02 + 3.0 * rng.uniform(0, 1, ntotal - homo - 1) ** 2]))
    weight, freq = rng.uniform(0.05, 1.0, naux), 0.3 + 8.0 * rng.uniform(0, 1, naux) ** 2
    push(ctx, tc)
    ctx.sigma_ppm_set(weight, freq, e, homo, 0, 0, 1e-3)
    levels = np.repeat(np.arange(mtotal), 40)
    freqs = np.tile(np.linspace(-1.5, 3.5, 40), mtotal) + 1e-3 * levels
    sig, dsig = ctx.sigma_ppm_eval(levels, freqs, deriv=True)
    ctx.set_option("sigma_tree_min_terms", 1e18)
    try:
        ref, dref = ctx.sigma_ppm_eval(levels, freqs, deriv=True)
    finally:
        ctx.set_option("sigma_tree_min_terms", 1 << 15)
    assert np.abs(ref - sig).max() < 1e-11 * np.abs(ref).max()
    assert np.abs(dref - dsig).max() < 1e-10 * np.abs(dref).max()


def test_sigma_golden(ctx, golden):
    """test_sigma_ppm.cc / test_sigma_exact.cc reference matrices through the CUDA path."""
    for kind in ("ppm", "exact"):
        d = "sigma_" + kind
        e = golden["inline/sigma_ppm_mo_energy"] if kind == "ppm" else golden["inline/sigma_exact_mo_energy"]
        tc = methane_mmn(golden[d + "/MOs"])
        r = orpa.RPA(tc)
        r.configure(4, 0, 16)
        r.set_rpa_input_energies(e)
        push(ctx, tc)
        assert rel_frob(golden[d + "/x_ref"], ctx.sigma_x(4, 0, 0, 16)) < 1e-5
        if kind == "ppm":
            s = osig.create("ppm", tc, r)
            s.configure(osig.SigmaOptions(homo=4, qpmin=0, qpmax=16, rpamin=0, rpamax=16, eta=1e-3))
            # PPM parameters from the device epsilon matrices (ppm.cc:30-59 on the GPU path)
            w, phi = ctx.sym_eig(ctx.rpa_epsilon(1, 0.0, 1e-4, e, 4, 0, 16))
            weight = 1 - 1 / w
            eps_i = ctx.rpa_epsilon(0, 0.5, 1e-4, e, 4, 0, 16)
            inv = ctx.inverse(phi.T @ eps_i @ phi)
            freq = np.zeros(17)
            for i in range(17):
                if weight[i] < 1e-5:
                    weight[i], freq[i] = 0.0, 0.5
                else:
                    nom = inv[i, i] - 1.0
                    freq[i] = np.sqrt(abs(-nom / (nom + weight[i]) * 0.25))
            ctx.mmn_mul_right(phi)
            ctx.sigma_ppm_set(weight, freq, e, 4, 0, 0, 1e-3)
            c = ctx.sigma_ppm_offdiag(e)
            c[np.diag_indices(17)] = ctx.sigma_ppm_eval(np.arange(17), e)
        else:
            omega, XpY, _ = r.diagonalize_h2p()
            ctx.sigma_exact_prepare(omega, XpY, e, 4, 0, 16, 0, 16, 1e-3)
            c = ctx.sigma_exact_offdiag(e)
            c[np.diag_indices(17)] = ctx.sigma_exact_eval(np.arange(17), e)
        assert rel_frob(golden[d + "/c_ref"], c) < 1e-5


# ------------------------------------------------------------------ BSE operator
OPS = {"singlet_tda": (1, 2, 1, 0), "triplet_tda": (1, 0, 1, 0), "singlet_btda_b": (0, 2, 0, 1),
       "hqp": (1, 0, 0, 0), "hx": (0, 1, 0, 0), "hd": (0, 0, 1, 0), "hd2": (0, 0, 0, 1)}


def test_bse_operator_golden(ctx, golden):
    tc = methane_mmn(golden["bse_operator/MOs"])
    tc.multiply_right(golden["bse_operator/rpa_op"])
    push(ctx, tc)
    eps = golden["inline/bse_operator_epsilon_inv"]
    ctx.bse_configure(4, 0, 0, 8, eps, golden["bse_operator/Hqp"][:9, :9])
    eye = np.eye(20)
    for name in ("hqp", "hx", "hd", "hd2"):
        dense = ctx.bse_matmul(OPS[name], eye)
        assert rel_frob(golden["bse_operator/%s_ref" % name], dense) < 1e-3
        assert np.abs(np.diag(dense) - ctx.bse_diagonal(OPS[name])).max() < 1e-12


@pytest.mark.parametrize("dims", [(50, 30, 33, 9, 3, 27, 6), (90, 40, 47, 15, 0, 39, 21)])
@pytest.mark.parametrize("chunk", [1 << 30, 1 << 16])
def test_bse_operator_random(ctx, dims, chunk):
    """The factorised form of every term (the materialised blocks are switched off here and have their own test)."""
    naux, mtotal, ntotal, homo, vmin, cmax, k = dims
    rng = np.random.default_rng(9)
    tc = random_tc(rng, naux, mtotal, ntotal)
    push(ctx, tc)
    ctx.set_option("bse_chunk_bytes", chunk)
    ctx.set_option("bse_dense", 0)
    vt, ct = homo - vmin + 1, cmax - homo
    Hqp = rng.standard_normal((vt + ct, vt + ct))
    Hqp = Hqp + Hqp.T
    eps = rng.uniform(0.3, 1.0, naux)
    ctx.bse_configure(homo, 0, vmin, cmax, eps, Hqp)
    X = rng.standard_normal((vt * ct, k))
    opt = bop.BSEOperatorOptions(homo=homo, rpamin=0, qpmin=0, vmin=vmin, cmax=cmax)
    for name, co in OPS.items():
        op = bop.BSEOperator(*co, eps, tc, Hqp)
        op.configure(opt)
        ref = op.matmul(X)  # reference formulation (row-by-row rebuild of H)
        assert rel_frob(ref, ctx.bse_matmul(co, X)) < TOL, name
        assert rel_frob(op.diagonal(), ctx.bse_diagonal(co)) < TOL, name
    ctx.set_option("bse_chunk_bytes", 8 << 30)
    ctx.set_option("bse_dense", 1)


@pytest.mark.parametrize("dims", [(50, 30, 33, 9, 3, 27, 6), (90, 40, 47, 15, 0, 39, 21), (800, 24, 40, 7, 1, 22, 5)])
def test_bse_operator_materialised_blocks(ctx, dims):
    """Hd / Hd2 from their resident B x B blocks (option bse_dense = 2: built at the first product) against the
    reference formulation; the first block is parked in the second Mmn buffer, the second one gets its own; a
    MultiplyRight (new Mmn content, buffers trade places) and a new screening each invalidate both."""
    if os.environ.get("GWBSE_B200_TEST_MOCK_DIR"):
        pytest.skip("a device-memory policy of the CUDA library; the CPU stand-in has none")
    naux, mtotal, ntotal, homo, vmin, cmax, k = dims
    rng = np.random.default_rng(19)
    tc = random_tc(rng, naux, mtotal, ntotal)
    push(ctx, tc)
    Q, _ = np.linalg.qr(rng.standard_normal((naux, naux)))
    ctx.mmn_mul_right(Q)  # the second Mmn buffer now exists
    tc.multiply_right(Q)
    ctx.set_option("bse_dense", 2)
    vt, ct = homo - vmin + 1, cmax - homo
    B = vt * ct
    Hqp = rng.standard_normal((vt + ct, vt + ct))
    Hqp = Hqp + Hqp.T
    opt = bop.BSEOperatorOptions(homo=homo, rpamin=0, qpmin=0, vmin=vmin, cmax=cmax)
    X = rng.standard_normal((B, k))

    def check(eps, tensor):
        ctx.bse_configure(homo, 0, vmin, cmax, eps, Hqp)
        for name, co in OPS.items():
            op = bop.BSEOperator(*co, eps, tensor, Hqp)
            op.configure(opt)
            assert rel_frob(op.matmul(X), ctx.bse_matmul(co, X)) < TOL, name
            assert rel_frob(op.matmul(X[:, :1]), ctx.bse_matmul(co, X[:, :1])) < TOL, name

    try:
        b0, c0, _ = ctx.bse_dense_stats()
        eps = rng.uniform(0.3, 1.0, naux)
        check(eps, tc)
        b1, c1, resident = ctx.bse_dense_stats()
        assert b1 - b0 == 2  # one block per direct term, reused by singlet / triplet / bare Hd and by B / bare Hd2
        assert c1 - c0 == 5 * (k + 1)
        ld = B + B % 2
        ld += 2 if ld % 256 == 0 else 0
        assert resident == 2 * 8.0 * ld * B
        check(eps, tc)  # same key: no further build
        assert ctx.bse_dense_stats()[0] == b1
        eps2 = rng.uniform(0.3, 1.0, naux)
        check(eps2, tc)  # new screening
        assert ctx.bse_dense_stats()[0] == b1 + 2
        ctx.mmn_mul_right(Q.T)  # new tensor, and the buffer the first block was parked in is the tensor now
        tc.multiply_right(Q.T)
        check(eps2, tc)
        assert ctx.bse_dense_stats()[0] == b1 + 4
    finally:
        ctx.set_option("bse_dense", 1)


@pytest.mark.parametrize("window", [(6, 40), (3, 39), (0, 47)])
def test_mul_right_on_a_row_window_then_the_rest(ctx, window):
    """gwbse_mmn_mul_right_window_dev: the BSE operator configured inside the window sees the rotated tensor at once
    (both the factorised and the materialised direct terms), every other reader gets the complete product - also
    when a second rotation follows while the first is still pending outside its window."""
    rng = np.random.default_rng(39)
    naux, mtotal, ntotal, homo, vmin, cmax = 60, 40, 47, 15, 6, 38
    tc = random_tc(rng, naux, mtotal, ntotal)
    push(ctx, tc)
    Q1, _ = np.linalg.qr(rng.standard_normal((naux, naux)))
    Q2, _ = np.linalg.qr(rng.standard_normal((naux, naux)))
    vt, ct = homo - vmin + 1, cmax - homo
    Hqp = rng.standard_normal((vt + ct, vt + ct))
    Hqp = Hqp + Hqp.T
    eps = rng.uniform(0.3, 1.0, naux)
    opt = bop.BSEOperatorOptions(homo=homo, rpamin=0, qpmin=0, vmin=vmin, cmax=cmax)
    X = rng.standard_normal((vt * ct, 7))
    try:
        for Q in (Q1, Q2):
            ctx.mmn_mul_right_window(Q, *window)
            tc.multiply_right(Q)
            ctx.bse_configure(homo, 0, vmin, cmax, eps, Hqp)
            for dense in (0, 2):
                ctx.set_option("bse_dense", dense)
                for name, co in OPS.items():
                    op = bop.BSEOperator(*co, eps, tc, Hqp)
                    op.configure(opt)
                    assert rel_frob(op.matmul(X), ctx.bse_matmul(co, X)) < TOL, (name, dense)
                    assert rel_frob(op.diagonal(), ctx.bse_diagonal(co)) < TOL, (name, dense)
        assert rel_frob(tc.M, ctx.mmn_get_all()) < 1e-12
        # epsilon reads every unoccupied row: a pending rotation is completed first
        ctx.mmn_mul_right_window(Q1.T, *window)
        tc.multiply_right(Q1.T)
        e = np.sort(rng.uniform(-1.0, 2.0, ntotal))
        r = orpa.RPA(tc)
        r.configure(homo, 0, ntotal - 1)
        r.set_rpa_input_energies(e)
        assert rel_frob(r.calculate_epsilon_i(0.5), ctx.rpa_epsilon(0, 0.5, 1e-4, e, homo, 0, ntotal - 1)) < TOL
        assert rel_frob(tc.M, ctx.mmn_get_all()) < 1e-12
    finally:
        ctx.set_option("bse_dense", 1)


def test_bse_materialised_blocks_pay_back_rule(ctx):
    """Default policy: single-column products under ever-changing screening (the dynamical-screening loop of
    bse.cc:608-716) never form a block; a many-column product does at once; the result is the same either way."""
    if os.environ.get("GWBSE_B200_TEST_MOCK_DIR"):
        pytest.skip("a device-memory policy of the CUDA library; the CPU stand-in has none")
    rng = np.random.default_rng(29)
    naux, mtotal, ntotal, homo, vmin, cmax = 60, 40, 47, 15, 0, 39
    tc = random_tc(rng, naux, mtotal, ntotal)
    push(ctx, tc)
    vt, ct = homo - vmin + 1, cmax - homo
    Hqp = np.eye(vt + ct)
    X = rng.standard_normal((vt * ct, 64))
    b0 = ctx.bse_dense_stats()[0]
    for it in range(6):
        ctx.bse_configure(homo, 0, vmin, cmax, rng.uniform(0.3, 1.0, naux), Hqp)
        ctx.bse_matmul(OPS["hd"], X[:, :1])
    assert ctx.bse_dense_stats()[0] == b0
    eps = rng.uniform(0.3, 1.0, naux)
    ctx.bse_configure(homo, 0, vmin, cmax, eps, Hqp)
    ctx.set_option("bse_dense", 0)
    fact = ctx.bse_matmul(OPS["hd"], X)
    ctx.set_option("bse_dense", 1)
    dense = ctx.bse_matmul(OPS["hd"], X)
    assert ctx.bse_dense_stats()[0] == b0 + 1
    assert rel_frob(fact, dense) < TOL


def test_properties_medium_size(ctx):
    """Size-independent properties at a size the oracle does not reach in seconds (Naux = 640, 150 x 301 levels):
    orthogonal MultiplyRight round trip, symmetry / definiteness of epsilon and Sigma_x, self-adjointness of the
    BSE blocks for an (m,n)-symmetric tensor, linearity of the operator product."""
    rng = np.random.default_rng(90)
    naux, mtotal, ntotal, homo = 640, 150, 301, 49
    tc = random_tc(rng, naux, mtotal, ntotal)
    sym = 0.5 * (tc.M[:, :mtotal, :] + tc.M[:, :mtotal, :].transpose(1, 0, 2))
    tc.M[:, :mtotal, :] = sym  # M[m][n] = M[n][m] inside the m window, as for real three-centre integrals
    push(ctx, tc)
    Q, _ = np.linalg.qr(rng.standard_normal((naux, naux)))
    ctx.mmn_mul_right(Q)
    assert rel_frob(tc.M @ Q, ctx.mmn_get_all()) < 1e-12
    ctx.mmn_mul_right(Q.T)
    assert rel_frob(tc.M, ctx.mmn_get_all()) < 1e-12
    e = np.sort(np.concatenate([rng.uniform(-1.2, -0.3, homo + 1), rng.uniform(0.05, 2.5, ntotal - homo - 1)]))
    eps = ctx.rpa_epsilon(0, 0.5, 1e-4, e, homo, 0, ntotal - 1)
    assert np.abs(eps - eps.T).max() == 0.0
    assert np.linalg.eigvalsh(eps).min() >= 1.0 - 1e-12  # eps_i(w) - 1 is a sum of weighted Gram matrices
    sx = ctx.sigma_x(homo, 0, 0, mtotal - 1)
    assert np.abs(sx - sx.T).max() < 1e-13 and np.linalg.eigvalsh(sx).max() <= 1e-12
    vmin, cmax, k = 20, 119, 6
    vt, ct = homo - vmin + 1, cmax - homo
    Hqp = rng.standard_normal((vt + ct,) * 2)
    Hqp = Hqp + Hqp.T
    ctx.bse_configure(homo, 0, vmin, cmax, rng.uniform(0.3, 1.0, naux), Hqp)
    X, Y = rng.standard_normal((vt * ct, k)), rng.standard_normal((vt * ct, k))
    for name in ("singlet_tda", "triplet_tda", "singlet_btda_b", "hd2"):
        co = OPS[name]
        HX, HY = ctx.bse_matmul(co, X), ctx.bse_matmul(co, Y)
        a, b = X.T @ HY, (Y.T @ HX).T
        assert np.abs(a - b).max() < 1e-11 * max(1.0, np.abs(a).max()), name
        HZ = ctx.bse_matmul(co, 2.0 * X - 3.0 * Y)
        assert rel_frob(2.0 * HX - 3.0 * HY, HZ) < 1e-12, name


def test_bse_operator_errors(ctx):
    from votca_b200.api import GwbseError
    rng = np.random.default_rng(10)
    push(ctx, random_tc(rng))
    ctx.bse_configure(9, 0, 3, 27, np.ones(70), np.eye(25))
    with pytest.raises(GwbseError, match="Hd and Hd2"):
        ctx.bse_matmul((1, 0, 1, 1), np.zeros((7 * 18, 1)))


# ------------------------------------------------------------------ Davidson helpers
def test_gramschmidt_and_correction(ctx):
    rng = np.random.default_rng(11)
    Q = rng.standard_normal((500, 14))
    Q[:, :6] = np.linalg.qr(Q[:, :6])[0]
    ref = Q.copy()
    DavidsonSolver._gramschmidt(ref, 6)
    out = ctx.gramschmidt(Q, 6)
    assert rel_frob(ref, out) < 1e-11
    assert np.abs(out.T @ out - np.eye(14)).max() < 1e-13
    Q0 = rng.standard_normal((500, 5))
    ref0 = Q0.copy()
    DavidsonSolver._gramschmidt(ref0, 0)
    assert rel_frob(ref0, ctx.gramschmidt(Q0, 0)) < 1e-11
    ds = DavidsonSolver()
    ds.Adiag = rng.uniform(1, 3, 500)
    lam = np.array([0.9, 1.1, 1.2])
    R = rng.standard_normal((500, 3))
    Qv = rng.standard_normal((500, 3))
    for corr in ("DPR", "OLSEN"):
        ds.correction = corr
        refw = np.array([ds._correction(Qv[:, j], lam[j], R[:, j]) for j in range(3)]).T
        refw /= np.linalg.norm(refw, axis=0)
        assert rel_frob(refw, ctx.davidson_correction(ds.Adiag, lam, R, Qv, olsen=(corr == "OLSEN"))) < 1e-11
