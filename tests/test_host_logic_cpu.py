"""CPU tests of the host-side QP root search: the known-answer cases of the reference's
xtp/src/tests/test_qp_solver_utils.cc:87-316, run against BOTH restatements of qp_solver_utils.h -
the C++ host layer that ships (votca_b200/host/qp_rootsearch.h, through a small g++-built harness, no GPU
involved) and the oracle (oracle/qp_solver.py)."""
import ctypes
import os
import subprocess
from dataclasses import dataclass

import numpy as np
import pytest

from oracle import qp_solver as oq

HERE = os.path.dirname(os.path.abspath(__file__))
KTOL = 1e-10


@pytest.fixture(scope="module")
def harness():
    src = os.path.join(HERE, "host_harness", "qp_harness.cc")
    out = os.path.join(HERE, "host_harness", "build", "libqp_harness.so")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    subprocess.run(["g++", "-std=c++17", "-O1", "-fPIC", "-shared", "-o", out, src], check=True)
    lib = ctypes.CDLL(out)
    d, l, i = ctypes.c_double, ctypes.c_long, ctypes.c_int
    lib.qp_normalize.argtypes = [l, d, d, d, d, l, ctypes.c_void_p]
    lib.qp_effective_shell_width.argtypes = [d, d, l]
    lib.qp_effective_shell_width.restype = d
    lib.qp_accept_root.argtypes = [d, d, d, d, d]
    lib.qp_windowed.argtypes = [i, d, d, d, d, d, l, d, d, d, d, i, ctypes.c_void_p]
    lib.anderson_run.argtypes = [l, d, l, l, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
    lib.newton_sqrt.argtypes = [d, d, l, d, ctypes.c_void_p]
    lib.quadrature_points.argtypes = [ctypes.c_char_p, l, ctypes.c_void_p, ctypes.c_void_p]
    lib.vc2index_eval.argtypes = [l, l, l, i, l, l]
    lib.vc2index_eval.restype = l
    return lib


@dataclass
class LegacyOpt:
    qp_grid_steps: int = 0
    qp_grid_spacing: float = 0.0
    qp_full_window_half_width: float = -1.0
    qp_dense_spacing: float = -1.0
    qp_adaptive_shell_width: float = -1.0
    qp_adaptive_shell_count: int = 0


def _normalize_cpp(lib, o):
    out = np.zeros(6)
    rc = lib.qp_normalize(o.qp_grid_steps, o.qp_grid_spacing, o.qp_full_window_half_width, o.qp_dense_spacing,
                          o.qp_adaptive_shell_width, o.qp_adaptive_shell_count, out.ctypes.data)
    assert rc == 0
    return out


def _normalize_py(o):
    legacy = (oq.legacy_full_window_half_width(o), oq.legacy_adaptive_shell_width(o))
    oq.normalize_grid_search_options(o)
    return np.array([o.qp_full_window_half_width, o.qp_dense_spacing, o.qp_adaptive_shell_width,
                     o.qp_adaptive_shell_count, *legacy])


NORMALIZE_CASES = [
    # (options, expected half width, dense spacing, shell width)      test_qp_solver_utils.cc:87-150
    (dict(), 0.75, 0.002, 0.025),
    (dict(qp_grid_steps=201, qp_grid_spacing=0.01), 1.0, 0.01, 2.0 / 49.0),
    (dict(qp_grid_steps=201, qp_grid_spacing=0.01, qp_full_window_half_width=0.75, qp_dense_spacing=0.002,
          qp_adaptive_shell_width=0.03), 0.75, 0.002, 0.03),
]


@pytest.mark.parametrize("kw,half,dense,shell", NORMALIZE_CASES)
def test_normalize_grid_search_options(harness, kw, half, dense, shell):
    for res in (_normalize_cpp(harness, LegacyOpt(**kw)), _normalize_py(LegacyOpt(**kw))):
        assert res[0] == pytest.approx(half, rel=KTOL)
        assert res[1] == pytest.approx(dense, rel=KTOL)
        assert res[2] == pytest.approx(shell, rel=KTOL)
        assert res[3] == 0
    if kw.get("qp_grid_steps"):
        for res in (_normalize_cpp(harness, LegacyOpt(**kw)), _normalize_py(LegacyOpt(**kw))):
            assert res[4] == pytest.approx(1.0, rel=KTOL)          # (201 - 1) * 0.01 / 2
            assert res[5] == pytest.approx(2.0 / 49.0, rel=KTOL)  # full width / (max(21, 201 / 4) - 1)


def test_shell_count_overrides_shell_width(harness):
    assert harness.qp_effective_shell_width(0.75, 0.03, 30) == pytest.approx(0.75 / 30.0, rel=KTOL)
    o = oq.SolverOptions(qp_full_window_half_width=0.75, qp_adaptive_shell_width=0.03, qp_adaptive_shell_count=30)
    assert oq.effective_adaptive_shell_width(o) == pytest.approx(0.75 / 30.0, rel=KTOL)


@pytest.mark.parametrize("residual,Z,ok", [(1e-7, 0.8, True), (1e-3, 0.8, False), (1e-7, 0.01, False),
                                           (1e-7, 2.0, False), (1e-7, -0.5, False)])
def test_accept_root(harness, residual, Z, ok):
    assert bool(harness.qp_accept_root(residual, Z, 1e-5, 0.05, 1.5)) is ok
    cand = oq.RootCandidate(omega=0.12, residual=residual, deriv=-1.25, Z=Z, distance_to_ref=0.12)
    assert oq.accept_root(cand, oq.SolverOptions(g_sc_limit=1e-5, min_accepted_Z=0.05, max_accepted_Z=1.5)) is ok


class _Mock:
    def __init__(self, kind, r1, r2=0.0):
        self.kind, self.r1, self.r2 = kind, r1, r2

    def value(self, w, *_):
        if self.kind == 0:
            return self.r1 - w
        if self.kind == 1:
            return w - self.r1
        return (w - self.r1) * (w - self.r2)

    def deriv(self, w):
        return 1.0 if self.kind == 1 else -1.0


def _windowed_cpp(lib, kind, r1, r2, left, right, half, brent):
    out = np.zeros(10)
    lib.qp_windowed(kind, r1, r2, 0.0, left, right, 1, 1e-8, half, 0.002, 0.05, int(brent), out.ctypes.data)
    return dict(found=bool(out[0]), root=out[1], n_acc=int(out[2]), n_rej=int(out[3]), Z=out[4],
                shells=int(out[5]), first_interval=int(out[6]), first_accepted=int(out[7]), chosen=int(out[8]),
                intervals=int(out[9]))


def _windowed_py(kind, r1, r2, left, right, half, brent):
    opt = oq.SolverOptions(g_sc_limit=1e-8, qp_bisection_max_iter=200, qp_full_window_half_width=half,
                           qp_dense_spacing=0.002, qp_adaptive_shell_width=0.05, qp_adaptive_shell_count=0)
    root, acc, rej, d = oq.solve_qp_grid_windowed(_Mock(kind, r1, r2), 0.0, left, right, 1, opt, use_brent=brent)
    return dict(found=root is not None, root=root, n_acc=len(acc), n_rej=len(rej),
                Z=acc[0].Z if acc else (rej[0].Z if rej else 0.0), shells=d.shells_explored,
                first_interval=d.first_interval_shell, first_accepted=d.first_accepted_shell, chosen=d.chosen_shell,
                intervals=d.intervals_found)


def _both(harness, *a):
    return _windowed_cpp(harness, *a), _windowed_py(*a)


def test_windowed_solver_simple_root_bisection(harness):
    for r in _both(harness, 0, 0.23, 0.0, -0.5, 0.5, 0.5, False):
        assert r["found"] and r["root"] == pytest.approx(0.23, rel=1e-6)
        assert (r["n_acc"], r["n_rej"]) == (1, 0) and r["Z"] == pytest.approx(1.0, rel=KTOL)
        assert (r["first_interval"], r["first_accepted"], r["chosen"], r["intervals"]) == (5, 5, 5, 1)
        assert r["shells"] >= 5


def test_windowed_solver_simple_root_brent(harness):
    for r in _both(harness, 0, 0.23, 0.0, -0.5, 0.5, 0.5, True):
        assert r["found"] and r["root"] == pytest.approx(0.23, rel=1e-10)
        assert (r["n_acc"], r["n_rej"], r["intervals"]) == (1, 0, 1) and r["shells"] >= 5


def test_windowed_solver_returns_nearest_accepted_root(harness):
    for r in _both(harness, 2, 0.12, 0.62, -0.2, 0.8, 0.8, False):
        assert r["found"] and r["root"] == pytest.approx(0.12, rel=1e-6)
        assert (r["n_acc"], r["n_rej"]) == (2, 0) and r["intervals"] >= 2
        assert r["first_accepted"] >= 0 and r["chosen"] >= 0


def test_windowed_solver_returns_rejected_root_without_accepted_one(harness):
    for r in _both(harness, 1, 0.23, 0.0, -0.5, 0.5, 0.5, False):
        assert r["found"] and r["root"] == pytest.approx(0.23, rel=1e-6)
        assert (r["n_acc"], r["n_rej"]) == (0, 1) and r["Z"] < 0.0
        assert (r["first_interval"], r["first_accepted"], r["chosen"], r["intervals"]) == (5, -1, 5, 1)


def test_host_and_oracle_agree(harness):
    """Same root and the same shell bookkeeping from both restatements (the C++ search is the product, the Python
    one the checker) for several window placements, including a pair of roots inside one shell interval that the
    scan cannot see (no sign change between shell points)."""
    for (r1, r2, left, right) in [(0.12, 0.62, -0.2, 0.8), (-0.31, 0.44, -0.6, 0.7), (0.06, 0.08, -0.3, 0.3)]:
        a, b = _both(harness, 2, r1, r2, left, right, 0.8, False)
        assert a["found"] == b["found"]
        if a["found"]:
            assert a["root"] == pytest.approx(b["root"], abs=1e-9)
        else:
            assert (r1, r2) == (0.06, 0.08)
        for k in ("n_acc", "n_rej", "shells", "first_interval", "first_accepted", "chosen", "intervals"):
            assert a[k] == b[k], k


# test_anderson.cc:33-100
ANDERSON_IN1 = np.array([-0.580533, -0.535803, -0.476481, -0.380558, 0.0969526, 0.133036, 0.164243])
ANDERSON_OUT = np.array([[-0.604342, -0.548675, -0.488088, -0.385654, 0.106193, 0.139172, 0.170433],
                         [-0.605576, -0.549458, -0.488876, -0.385821, 0.106788, 0.139509, 0.170768],
                         [-0.606242, -0.549887, -0.489296, -0.385898, 0.107162, 0.139718, 0.170977]])
ANDERSON_REF = np.array([[-0.597199, -0.544813, -0.484606, -0.384126, 0.103421, 0.137331, 0.168576],
                         [-0.606303, -0.549862, -0.489247, -0.385968, 0.10708, 0.139698, 0.170959],
                         [-0.606242, -0.549888, -0.489296, -0.385897, 0.107163, 0.139718, 0.170977]])


def _approx(a, ref, tol):  # Eigen isApprox: relative Frobenius
    return np.linalg.norm(a - ref) <= tol * min(np.linalg.norm(a), np.linalg.norm(ref))


def test_anderson_mixing_reference_vectors(harness):
    """Linear step, 2nd- and 3rd-order Anderson steps of the reference's unit test, host C++ class and oracle."""
    from oracle import gw as ogw
    mixed = np.zeros((3, 7))
    out = np.ascontiguousarray(ANDERSON_OUT)
    harness.anderson_run(3, 0.7, 7, 3, ANDERSON_IN1.ctypes.data, out.ctypes.data, mixed.ctypes.data)
    mix = ogw.Anderson(3, 0.7)
    vin, py = ANDERSON_IN1.copy(), []
    for s in range(3):
        mix.update_input(vin)
        mix.update_output(ANDERSON_OUT[s])
        vin = mix.mix_history()
        py.append(vin)
    for s in range(3):
        assert _approx(mixed[s], ANDERSON_REF[s], 1e-5), s
        assert _approx(py[s], ANDERSON_REF[s], 1e-5), s
        assert np.abs(mixed[s] - py[s]).max() < 1e-9


def test_newton_rapson_reference_case(harness):
    root = ctypes.c_double()
    info = harness.newton_sqrt(612.0, 10.0, 50, 1e-9, ctypes.byref(root))
    assert info == 0 and root.value == pytest.approx(24.738633753, rel=1e-9)

    class F:
        def value(self, x):
            return x * x - 612.0

        def deriv(self, x):
            return 2.0 * x
    x, ok = oq.newton_raphson(F(), 10.0, 50, 1e-9, 1.0)
    assert ok and x == pytest.approx(24.738633753, rel=1e-9)


# xtp/src/tests/DataFiles/gaussian_quadratures/{gauss_legendre,modified_gauss_legendre}.mm: integral of
# exp(-x^2) with orders 8, 10, 12, 14, 16, 18, 20, 40, 100 (test_gaussian_quadratures.cc:36-110, tolerance 1e-10)
QUAD_ORDERS = [8, 10, 12, 14, 16, 18, 20, 40, 100]
QUAD_REF = {
    "legendre": [1.798265329486247, 1.7635524479921414, 1.7755135781619735, 1.7715209261601161, 1.7726581320386712,
                 1.7724628015402868, 1.7724077952532697, 1.772453824202762, 1.7724538509055152],
    "modified_legendre": [1.7666951327337739, 1.7769696007269602, 1.7704741425269137, 1.773217660393972,
                          1.772169769028222, 1.7725597110830396, 1.7724137859609552, 1.7724538549863016,
                          1.7724538509055152],
}


@pytest.mark.parametrize("scheme", ["legendre", "modified_legendre"])
def test_cda_quadratures_reference_integrals(harness, scheme):
    """Gauss-Legendre points/weights (Newton iteration here, 50-digit tables in the reference) and their mapping to
    the integration domain, for the C++ host layer and the oracle."""
    from oracle import sigma as osig
    got_cpp, got_py = [], []
    for order in QUAD_ORDERS:
        pts, wts = np.zeros(order), np.zeros(order)
        sym = harness.quadrature_points(scheme.encode(), order, pts.ctypes.data, wts.ctypes.data)
        assert sym == (1 if scheme == "modified_legendre" else 0)
        got_cpp.append(float(np.sum(wts * (2.0 if sym else 1.0) * np.exp(-pts ** 2))))
        s = osig.SigmaCDA.__new__(osig.SigmaCDA)
        s.opt = osig.SigmaOptions(order=order, quadrature_scheme=scheme)
        p, w, sy = s._quadrature()
        got_py.append(float(np.sum(w * (2.0 if sy else 1.0) * np.exp(-p ** 2))))
    ref = np.array(QUAD_REF[scheme])
    assert _approx(np.array(got_cpp), ref, 1e-10)
    assert _approx(np.array(got_py), ref, 1e-10)
    assert harness.quadrature_points(b"laguerre", 8, None, None) == -1  # not available in this build: loud error


def test_vc2index_known_answers(harness):
    """test_vc2index.cc:33-68: I(3, 12) = 32 for vmin 0, cmin 10, ctotal 10; v(I), c(I) invert it over all pairs."""
    vmin, cmin, ctotal, vtotal = 0, 10, 10, 9
    ev = lambda what, a, b=0: harness.vc2index_eval(vmin, cmin, ctotal, what, a, b)  # noqa: E731
    assert ev(0, 3, 12) == 32 and ev(2, 32) == 12 and ev(1, 32) == 3
    j = 0
    for v2 in range(vtotal):
        for c2 in range(ctotal):
            assert ev(1, j) == vmin + v2 and ev(2, j) == cmin + c2 and ev(0, vmin + v2, cmin + c2) == j
            j += 1
