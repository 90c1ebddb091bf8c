"""GPU parity of the C++ host layer (GWBSE job facade -> GW / BSE / Davidson classes -> CUDA kernels)
against the reference's golden vectors and the CPU oracle.  Mirrors test_gw.cc, test_bse.cc."""
import os
import tempfile

import numpy as np
import pytest

from oracle import bse as obse
from oracle import gw as ogw
from tests.helpers import methane_mmn, rel_frob

pytestmark = pytest.mark.gpu


def make_job(methane, mos, energies, homo=4, **options):
    from votca_b200.api import Job
    job = Job(0)
    job.set_scalar("homo", homo)
    job.set_array("mos", mos)
    job.set_array("mo_energies", energies)
    job.set_ao3c(methane["ao3c"])
    job.set_array("aux_overlap", methane["S"])
    job.set_array("aux_coulomb", methane["V"])
    job.set_options(ranges="full", **options)
    return job


GW_OPTS = dict(tasks="gw", gw__mode="G0W0", gw__sigma_integrator="ppm", gw__eta=1e-3, gw__qp_solver="grid",
               gw__mixing_order=0, gw__mixing_alpha=0.7, gw__qp_sc_limit=1e-5, gw__qp_sc_max_iter=50,
               gw__sc_limit=1e-5)


# test_gw.cc:111-270
@pytest.mark.parametrize("suffix", ["", "2"])
def test_gw_full(golden, methane, suffix):
    e = golden["inline/gw_mo_eigenvalues"]
    job = make_job(methane, golden["gw/mo_eigenvectors" + suffix], e, gw__qp_grid_steps=601,
                   gw__qp_grid_spacing=0.005, **GW_OPTS)
    job.set_array("vxc", golden["gw/vxc" + suffix])
    job.run()
    ref = golden["gw/ref" + suffix]
    assert rel_frob(np.diag(ref), job.get("QPpert_energies")) < 1e-4
    assert rel_frob(ref, job.get("Hqp")) < 1e-4
    # and against the oracle on identical inputs: 1e-6 Ha (BASELINE.json)
    tc = methane_mmn(golden["gw/mo_eigenvectors" + suffix])
    g = ogw.GW(tc, golden["gw/vxc" + suffix], e)
    g.configure(ogw.GWOptions(homo=4, qpmin=0, qpmax=16, rpamin=0, rpamax=16, gw_sc_max_iterations=1, eta=1e-3,
                              sigma_integration="ppm", qp_solver="grid", qp_grid_steps=601, qp_grid_spacing=0.005,
                              gw_mixing_order=0, g_sc_limit=1e-5, g_sc_max_iterations=50))
    g.calculate_gw_perturbation()
    assert np.abs(g.get_gwa_results() - job.get("QPpert_energies")).max() < 1e-6
    g.calculate_hqp()
    assert np.abs(g.get_hqp() - job.get("Hqp")).max() < 1e-6
    assert "GW calculation took" in job.log()
    job.close()


# test_gw.cc:272-340
def test_gw_canonical_and_brent(golden, methane):
    e = golden["inline/gw_mo_eigenvalues"]
    res = {}
    for finder in ("bisection", "brent"):
        job = make_job(methane, golden["gw/mo_eigenvectors2"], e, gw__qp_full_window_half_width=1.5,
                       gw__qp_dense_spacing=0.005, gw__qp_adaptive_shell_width=0.02, gw__qp_root_finder=finder,
                       **GW_OPTS)
        job.set_array("vxc", golden["gw/vxc2"])
        job.run()
        res[finder] = job.get("QPpert_energies")
        job.close()
    assert rel_frob(np.diag(golden["gw/ref2"]), res["bisection"]) < 1e-4
    assert rel_frob(res["bisection"], res["brent"]) < 1e-5


@pytest.mark.parametrize("integrator", ["ppm", "exact", "cda"])
def test_evgw_vs_oracle(golden, methane, integrator):
    """evGW (3 iterations, Anderson mixing) for all three sigma integrators against the oracle."""
    e = golden["inline/gw_mo_eigenvalues"]
    mos, vxc = golden["gw/mo_eigenvectors"], golden["gw/vxc"]
    opts = dict(GW_OPTS)
    opts.update(gw__mode="evGW", gw__sigma_integrator=integrator, gw__sc_max_iter=3, gw__mixing_order=2,
                gw__quadrature_order=12, gw__alpha=1e-3)
    job = make_job(methane, mos, e, **opts)
    job.set_array("vxc", vxc)
    job.run()
    tc = methane_mmn(mos)
    g = ogw.GW(tc, vxc, e)
    g.configure(ogw.GWOptions(homo=4, qpmin=0, qpmax=16, rpamin=0, rpamax=16, gw_sc_max_iterations=3, eta=1e-3,
                              sigma_integration=integrator, qp_solver="grid", gw_mixing_order=2, gw_mixing_alpha=0.7,
                              g_sc_limit=1e-5, g_sc_max_iterations=50, gw_sc_limit=1e-5, order=12, alpha=1e-3))
    g.calculate_gw_perturbation()
    assert np.abs(g.get_gwa_results() - job.get("QPpert_energies")).max() < 1e-6
    assert np.abs(g.rpa_input_energies() - job.get("RPA_inputenergies")).max() < 1e-6
    assert job.scalar("gw_iterations") == g.iterations
    job.close()


def _bse_job(golden, methane, **options):
    Hqp = golden["bse/Hqp"]
    opts = dict(tasks="singlets", bse__exctotal=3, bse__useTDA=True, bse__davidson__correction="DPR",
                bse__davidson__tolerance="lapack", bse__davidson__update="safe", bse__davidson__maxiter=50,
                bse__use_Hqp_offdiag=True, bse__dyn_screen_max_iter=10, bse__dyn_screen_tol=1e-5)
    opts.update(options)
    job = make_job(methane, golden["bse/MOs"], golden["bse/MO_energies"].ravel(), **opts)
    job.set_array("Hqp", Hqp)
    job.set_array("RPA_inputenergies", np.diag(Hqp).copy())
    return job


def _subspace(ref, vec):
    return np.linalg.norm(ref.T @ vec, axis=0)


# test_bse.cc:37-377
def test_bse_singlets_tda(golden, methane):
    job = _bse_job(golden, methane, bse__use_Hqp_offdiag=False, bse__dyn_screen_max_iter=0)
    job.run()
    assert rel_frob(golden["bse/singlets_nooffdiag_tda"].ravel(), job.get("BSE_singlet_eigenvalues")) < 1e-3
    assert np.allclose(_subspace(golden["bse/singlets_psi_nooffdiag_tda"], job.get("BSE_singlet_eigenvectors")), 1,
                       atol=1e-5)
    job.close()
    job = _bse_job(golden, methane)
    job.run()
    assert rel_frob(golden["bse/singlets_tda"].ravel(), job.get("BSE_singlet_eigenvalues")) < 1e-3
    assert np.allclose(_subspace(golden["bse/singlets_psi_tda"], job.get("BSE_singlet_eigenvectors")), 1, atol=1e-5)
    assert rel_frob(golden["bse/singlets_dynamic_TDA"].ravel(), job.get("BSE_singlet_dynamic")) < 5e-3
    assert job.scalar("singlet_converged") == 1.0
    job.close()


def test_bse_singlets_full(golden, methane):
    job = _bse_job(golden, methane, bse__useTDA=False)
    job.run()
    nrm = lambda a: a / np.linalg.norm(a, axis=0)  # noqa: E731
    assert rel_frob(golden["bse/singlets_btda"].ravel(), job.get("BSE_singlet_eigenvalues")) < 1e-3
    assert np.allclose(_subspace(nrm(golden["bse/singlets_psi_btda"]), nrm(job.get("BSE_singlet_eigenvectors"))), 1,
                       atol=1e-5)
    assert np.allclose(_subspace(nrm(golden["bse/singlets_psi_AR_btda"]), nrm(job.get("BSE_singlet_eigenvectors2"))),
                       1, atol=1e-5)
    assert rel_frob(golden["bse/singlets_dynamic_full"].ravel(), job.get("BSE_singlet_dynamic")) < 5e-2
    job.close()


def test_bse_triplets(golden, methane):
    job = _bse_job(golden, methane, tasks="triplets", bse__exctotal=1)
    job.run()
    assert rel_frob(golden["bse/triplets_tda"].ravel(), job.get("BSE_triplet_eigenvalues")) < 1e-3
    assert rel_frob(golden["bse/triplets_dynamic_TDA"].ravel(), job.get("BSE_triplet_dynamic")) < 1e-3
    job.close()


def test_bse_vs_oracle_energies_and_oscillator_strengths(golden, methane):
    """Full GW+BSE chain vs the oracle on identical inputs: energies 1e-6 Ha, f 1e-5 (BASELINE.json)."""
    from oracle import integrals
    e = golden["inline/gw_mo_eigenvalues"]
    mos, vxc = golden["gw/mo_eigenvectors"], golden["gw/vxc"]
    dip_ao = integrals.dipole(methane["basis"])
    inter = obse.free_transition_dipoles(dip_ao, mos, 0, 5, 5, 12)
    for tda in (True, False):
        opts = dict(GW_OPTS)
        opts.update(tasks="gw,singlets,triplets", bse__exctotal=4, bse__useTDA=tda, bse__use_Hqp_offdiag=True,
                    bse__davidson__tolerance="strict")
        job = make_job(methane, mos, e, **opts)
        job.set_array("vxc", vxc)
        for ax, d in zip("xyz", inter):
            job.set_array("dipole_" + ax, d)
        job.run()
        tc = methane_mmn(mos)
        g = ogw.GW(tc, vxc, e)
        g.configure(ogw.GWOptions(homo=4, qpmin=0, qpmax=16, rpamin=0, rpamax=16, gw_sc_max_iterations=1, eta=1e-3,
                                  sigma_integration="ppm", gw_mixing_order=0, g_sc_limit=1e-5,
                                  g_sc_max_iterations=50))
        g.calculate_gw_perturbation()
        g.calculate_hqp()
        b = obse.BSE(tc, factorised=True)
        b.configure(obse.BSEOptions(useTDA=tda, homo=4, rpamin=0, rpamax=16, qpmin=0, qpmax=16, vmin=0, cmax=16,
                                    nmax=4, davidson_tolerance="strict", use_Hqp_offdiag=True),
                    g.rpa_input_energies(), g.get_hqp())
        et = b.solve_triplets()
        es = b.solve_singlets()
        assert np.abs(es["eigenvalues"] - job.get("BSE_singlet_eigenvalues")).max() < 1e-6
        assert np.abs(et["eigenvalues"] - job.get("BSE_triplet_eigenvalues")).max() < 1e-6
        tdip = obse.coupled_transition_dipoles(es, inter, 12, 5, tda)
        f_ref = obse.oscillator_strengths(tdip, es["eigenvalues"])
        # degenerate (T2) levels: compare the sums over each degenerate shell
        f = job.get("oscillator_strengths")
        assert abs(f_ref.sum() - f.sum()) < 1e-5
        job.close()


def test_options_xml_roundtrip(golden, methane):
    """The job accepts the reference's options XML layout (dftgwbse -> gwbse subtree)."""
    xml = """<options><dftgwbse><job_name>methane</job_name><gwbse>
      <tasks>gw</tasks><ranges>full</ranges>
      <gw><mode>G0W0</mode><sigma_integrator>ppm</sigma_integrator><mixing_order>0</mixing_order>
          <qp_grid_steps>601</qp_grid_steps><qp_grid_spacing>0.005</qp_grid_spacing>
          <qp_sc_max_iter>50</qp_sc_max_iter></gw>
    </gwbse></dftgwbse></options>"""
    with tempfile.NamedTemporaryFile("w", suffix=".xml", delete=False) as fh:
        fh.write(xml)
        path = fh.name
    from votca_b200.api import Job
    job = Job(0)
    job.load_options_xml(path)
    os.unlink(path)
    job.set_scalar("homo", 4)
    job.set_array("mos", golden["gw/mo_eigenvectors"])
    job.set_array("mo_energies", golden["inline/gw_mo_eigenvalues"])
    job.set_array("vxc", golden["gw/vxc"])
    job.set_ao3c(methane["ao3c"])
    job.set_array("aux_overlap", methane["S"])
    job.set_array("aux_coulomb", methane["V"])
    job.run()
    assert rel_frob(np.diag(golden["gw/ref"]), job.get("QPpert_energies")) < 1e-4
    job.close()


def test_error_convention():
    from votca_b200.api import GwbseError, Job
    job = Job(0)
    with pytest.raises(GwbseError, match="homo"):
        job.run()
    job.close()


def test_evgw_treecode_sigma(golden, methane):
    """The same evGW run with the treecode Sigma_c evaluator forced on (sigma_tree.cu): energies unchanged."""
    e = golden["inline/gw_mo_eigenvalues"]
    mos, vxc = golden["gw/mo_eigenvectors"], golden["gw/vxc"]
    opts = dict(GW_OPTS)
    opts.update(gw__mode="evGW", gw__sigma_integrator="ppm", gw__sc_max_iter=3)
    res = []
    for min_terms in (1e18, 0):
        job = make_job(methane, mos, e, **opts)
        job.set_array("vxc", vxc)
        job.kernel_ctx().set_option("sigma_tree_min_terms", min_terms)
        job.run()
        res.append((job.get("QPpert_energies"), job.get("RPA_inputenergies")))
        job.close()
    assert np.abs(res[0][0] - res[1][0]).max() < 1e-9
    assert np.abs(res[0][1] - res[1][1]).max() < 1e-9
