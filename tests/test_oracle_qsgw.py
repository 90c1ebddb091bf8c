"""QSGW (GW::CalculateQSGW, gw.cc:798-1130) in the oracle: the two cases of the reference's own test suite
(test_gw.cc:342-436 qsgw_ppm, :438-500 qsgw_virtual_threshold) on methane / 3-21G with the reference's MO and Vxc
fixtures.  The reference checks structure, not numbers: convergence from a G0W0 and from an evGW seed to the same
energies (1e-4), a physical gap, a unitary rotation of the right size, and for the trimmed window the seed energy of
the excluded level and an identity block in the rotation.  The GPU path is then held to these oracle numbers
(tests/test_gpu_zz_qsgw.py)."""
import numpy as np
import pytest

from oracle import gw as ogw
from tests.helpers import methane_mmn, rel_frob


def qsgw_options(**kw):
    o = ogw.GWOptions(homo=4, qpmin=0, qpmax=16, rpamin=0, rpamax=16, gw_sc_max_iterations=1, qp_solver="grid", eta=1e-3,
                      sigma_integration="ppm", reset_3c=5, gw_mixing_order=20, gw_mixing_alpha=0.2, g_sc_limit=1e-5,
                      g_sc_max_iterations=50, gw_sc_limit=1e-5, qsgw_max_iterations=50, qsgw_sc_limit=1e-5)
    for k, v in kw.items():
        setattr(o, k, v)
    return o


def run_qsgw(golden, **kw):
    opt = qsgw_options(**kw)
    q = opt.qpmax - opt.qpmin + 1
    tc = methane_mmn(golden["gw/mo_eigenvectors"])
    g = ogw.GW(tc, golden["gw/vxc"][:q, :q], golden["inline/gw_mo_eigenvalues"])
    g.configure(opt)
    g.calculate_gw_perturbation()
    seed = g.get_gwa_results().copy()
    tc.rebuild()  # test_gw.cc:386: Mmn back in the DFT-MO basis before the QSGW loop
    g.calculate_qsgw()
    return g, seed


@pytest.fixture(scope="module")
def qsgw_from_g0w0(golden):
    return run_qsgw(golden, qpmax=13, qsgw_max_virt_correction=0.35)


def test_qsgw_ppm_starting_point_independence(golden, qsgw_from_g0w0):
    g1, _ = qsgw_from_g0w0
    g2, _ = run_qsgw(golden, qpmax=13, qsgw_max_virt_correction=0.35, gw_sc_max_iterations=50)
    e1, e2 = g1.get_gwa_results(), g2.get_gwa_results()
    assert g1.qsgw_iterations < 50 and g2.qsgw_iterations < 50
    assert rel_frob(e1, e2) < 1e-4
    assert e1[4] < 0.0 < e1[5]
    U = g1.qsgw_rotation
    assert U.shape == (14, 14)
    assert rel_frob(np.eye(14), U.T @ U) < 1e-6


def test_qsgw_virtual_threshold(golden):
    g, seed = run_qsgw(golden, qpmax=16, qsgw_max_virt_correction=0.2)
    e = g.get_gwa_results()
    assert abs(e[16] - seed[16]) <= 1e-8 * abs(seed[16])
    U = g.qsgw_rotation
    assert U.shape == (17, 17) and abs(U[16, 16] - 1.0) < 1e-8
    assert g.qsgw_seed_energies.shape == (17,)
    # the rotation registered with the RPA is cleared again (gw.cc:1062-1063)
    assert g.rpa.qsgw_U is None
