"""QSGW through the C++ host layer (GW::CalculateQSGW mirror, votca_b200/host/gw.h) and the CUDA kernels - rotation of
the QP-window rows (gwbse_mmn_rotate), hole slices rotated inside the RPA sums (gwbse_rpa_set_qsgw_rotation), Mmn rebuilt
in the QP basis for the BSE - against the oracle's restatement of gw.cc:798-1130 on the reference's methane fixtures,
with the structural checks of the reference's own tests (test_gw.cc:342-500)."""
import numpy as np
import pytest

from oracle import bse as obse
from oracle import gw as ogw
from tests.helpers import methane_integrals, methane_mmn, rel_frob
from tests.test_oracle_qsgw import qsgw_options

pytestmark = pytest.mark.gpu


def _job(golden, methane, **opts):
    from votca_b200.api import Job
    job = Job(0)
    job.set_scalar("homo", 4)
    job.set_array("mos", golden["gw/mo_eigenvectors"])
    job.set_array("mo_energies", golden["inline/gw_mo_eigenvalues"])
    job.set_ao3c(methane["ao3c"])
    job.set_array("aux_overlap", methane["S"])
    job.set_array("aux_coulomb", methane["V"])
    q = int(opts.pop("q"))
    job.set_array("vxc", golden["gw/vxc"][:q, :q])
    job.set_options(**opts)
    return job


def _oracle(golden, **kw):
    opt = qsgw_options(**kw)
    q = opt.qpmax - opt.qpmin + 1
    tc = methane_mmn(golden["gw/mo_eigenvectors"])
    g = ogw.GW(tc, golden["gw/vxc"][:q, :q], golden["inline/gw_mo_eigenvalues"])
    g.configure(opt)
    g.calculate_gw_perturbation()
    seed = g.get_gwa_results().copy()
    tc.rebuild()
    g.calculate_qsgw()
    return g, seed, tc


@pytest.mark.parametrize("integrator", ["ppm", "exact"])
def test_qsgw_matches_oracle(golden, methane, integrator):
    g, seed, tc = _oracle(golden, qpmax=13, qsgw_max_virt_correction=0.35, sigma_integration=integrator)
    job = _job(golden, methane, q=14, tasks="gw", ranges="explicit", rpamax=16, qpmin=0, qpmax=13, bsemin=0, bsemax=13,
               gw__mode="G0W0", gw__sigma_integrator=integrator, gw__do_qsgw=True, gw__qsgw_max_iterations=50,
               gw__qsgw_sc_limit=1e-5, gw__mixing_order=20, gw__mixing_alpha=0.2, gw__qsgw_max_virt_correction=0.35,
               gw__qp_sc_max_iter=50)
    try:
        job.run()
        assert job.scalar("is_qsgw") == 1.0
        assert int(job.scalar("qsgw_iterations")) == g.qsgw_iterations
        assert np.abs(job.get("QPpert_energies") - seed).max() < 1e-6          # seed energies stay in QPpert
        e = job.get("QPdiag_eigenvalues")
        assert np.abs(e - g.get_gwa_results()).max() < 1e-6
        assert e[4] < 0.0 < e[5]
        U = job.get("QPdiag_eigenvectors")
        assert U.shape == (14, 14) and rel_frob(np.eye(14), U.T @ U) < 1e-6
        # same rotation as the oracle's up to the sign of each column and rotations inside methane's degenerate levels:
        # compare the projectors on groups of (near-)degenerate QP energies
        Uo, eo = g.qsgw_rotation, g.get_gwa_results()
        i = 0
        while i < 14:
            j = i
            while j + 1 < 14 and abs(eo[j + 1] - eo[i]) < 1e-4:
                j += 1
            assert np.abs(U[:, i:j + 1] @ U[:, i:j + 1].T - Uo[:, i:j + 1] @ Uo[:, i:j + 1].T).max() < 1e-4
            i = j + 1
        assert np.abs(job.get("RPA_inputenergies") - g.rpa_input_energies()).max() < 1e-6
    finally:
        job.close()


def test_qsgw_virtual_threshold_and_bse_in_the_qp_basis(golden, methane):
    """test_gw.cc:438-500: level 16 exceeds the correction threshold, keeps its seed energy and an identity block in
    the rotation; then the BSE hookup of gwbse.cc:1032-1075 (Mmn refilled with C U, Hqp = diag(e_QSGW)) against the
    oracle doing the same."""
    from oracle import threecenter
    g, seed, tc = _oracle(golden, qpmax=16, qsgw_max_virt_correction=0.2)
    job = _job(golden, methane, q=17, tasks="gw,singlets", ranges="full", gw__mode="G0W0", gw__do_qsgw=True, gw__qsgw_max_iterations=50,
               gw__qsgw_sc_limit=1e-5, gw__mixing_order=20, gw__mixing_alpha=0.2, gw__qsgw_max_virt_correction=0.2,
               gw__qp_sc_max_iter=50, bse__exctotal=3, bse__useTDA=True, bse__davidson__tolerance="lapack")
    try:
        job.run()
        e = job.get("QPdiag_eigenvalues")
        U = job.get("QPdiag_eigenvectors")
        assert abs(e[16] - seed[16]) <= 1e-8 * abs(seed[16])
        assert U.shape == (17, 17) and abs(U[16, 16] - 1.0) < 1e-8
        assert np.abs(e - g.get_gwa_results()).max() < 1e-6
        C = golden["gw/mo_eigenvectors"] @ g.qsgw_rotation
        m = methane_integrals()
        tq = threecenter.TCMatrix(m["basis"].size, 0, 16, 0, 16)
        tq.fill_from_integrals(m["ao3c"], m["S"], m["V"], C)
        b = obse.BSE(tq, factorised=True)
        b.configure(obse.BSEOptions(useTDA=True, homo=4, rpamin=0, rpamax=16, qpmin=0, qpmax=16, vmin=0, cmax=16, nmax=3,
                                    davidson_tolerance="lapack", use_Hqp_offdiag=False),
                    g.rpa_input_energies(), np.diag(g.get_gwa_results()))
        es = b.solve_singlets()
        assert np.abs(es["eigenvalues"] - job.get("BSE_singlet_eigenvalues")).max() < 1e-6
    finally:
        job.close()
