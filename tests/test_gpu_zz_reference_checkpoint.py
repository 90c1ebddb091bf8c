"""CUDA path against the reference's OWN outputs: the BSE part of the dftgwbse integration tests (water, 3-21G +
aux-def2-svp; checkpoints molecule_neutral.orb / molecule_neutral_tda.orb, xtp/src/tests/CMakeLists.txt:336-416).
Inputs: the checkpoint's MOs, RPA input energies and Hqp (QPdiag eigendecomposition; Vxc is not stored, so the QP
step cannot be replayed) plus AO integrals from the oracle's integral code (pinned in test_oracle_golden.py).
Everything else - Mmn fill with d/f aux shells, V^-1/2, eps, screening rotation, BSE operator, Davidson, transition
dipoles, dynamical screening - runs through the C++ host layer and the CUDA kernels.  Tolerances: BASELINE.json
(energies 1e-6 Ha, oscillator strengths 1e-5); the reference's own test allows 1e-4."""
import numpy as np
import pytest

from oracle import bse as obse
from tests.helpers import orb_case, water_integrals

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("tag", ["neutral", "neutral_tda"])
def test_bse_matches_reference_checkpoint(tag):
    from votca_b200.api import Job
    w, c = water_integrals(), orb_case(tag)
    ref_e = c["BSE_singlet_eigenvalues"].ravel()
    vt, ct = c["homo"] - c["bse_vmin"] + 1, c["bse_cmax"] - c["homo"]
    inter = obse.free_transition_dipoles(w["dipole"], c["mos"], c["bse_vmin"], vt, c["homo"] + 1, ct)
    job = Job(0)
    job.set_scalar("homo", c["homo"])
    job.set_array("mos", c["mos"])
    job.set_array("mo_energies", c["mo_energies"].ravel())
    job.set_ao3c(w["ao3c"])
    job.set_array("aux_overlap", w["S"])
    job.set_array("aux_coulomb", w["V"])
    job.set_array("Hqp", c["Hqp"])
    job.set_array("RPA_inputenergies", c["RPA_inputenergies"].ravel())
    for ax, d in zip("xyz", inter):
        job.set_array("dipole_" + ax, d)
    job.set_options(ranges="full", tasks="singlets", bse__exctotal=len(ref_e), bse__useTDA=c["useTDA"],
                    bse__use_Hqp_offdiag=c["use_Hqp_offdiag"], bse__dyn_screen_max_iter=5, bse__dyn_screen_tol=1e-5)
    job.run()
    assert (job.scalar("rpamin"), job.scalar("rpamax"), job.scalar("bse_vmin"), job.scalar("bse_cmax")) == \
        (c["rpamin"], c["rpamax"], c["bse_vmin"], c["bse_cmax"])
    assert np.abs(job.get("BSE_singlet_eigenvalues").ravel() - ref_e).max() < 1e-6
    f_ref = obse.oscillator_strengths(c["transition_dipoles"], ref_e)
    assert np.abs(job.get("oscillator_strengths").ravel() - f_ref).max() < 1e-5
    assert np.abs(np.abs(job.get("transition_dipoles").T) - np.abs(c["transition_dipoles"])).max() < 1e-5
    assert np.abs(job.get("BSE_singlet_dynamic").ravel() - c["BSE_singlet_dynamic"].ravel()).max() < 1e-5
    job.close()


def test_ppm_parameters_match_reference_known_answer(golden, methane):
    """test_ppm.cc:36-108 through the CUDA path: Mmn filled on the device from core-Hamiltonian orbitals, eps_r(0)
    and eps_i(0.5) assembled and diagonalised / inverted on the device (ppm.cc:30-59), compared with the inline
    plasmon-pole frequencies and weights of the reference test (1e-4)."""
    from tests.helpers import methane_core_hamiltonian_mos, rel_frob
    from votca_b200.api import Context
    e, C = methane_core_hamiltonian_mos()
    n = methane["basis"].size
    ctx = Context(0)
    try:
        ctx.mmn_alloc(n, 0, 16, 0, 16)
        ctx.mmn_set_mos(C)
        ctx.mmn_fill_block(0, methane["ao3c"])
        L, removed = ctx.pseudo_invsqrt(methane["S"], methane["V"], 5e-7)
        assert removed == 0
        ctx.mmn_mul_right(L)
        w, phi = ctx.sym_eig(ctx.rpa_epsilon(1, 0.0, 1e-4, e, 4, 0, 16))
        weight = 1.0 - 1.0 / w
        eps_i = ctx.rpa_epsilon(0, 0.5, 1e-4, e, 4, 0, 16)
        inv = ctx.inverse(phi.T @ eps_i @ phi)
        freq = np.zeros(n)
        for i in range(n):  # ppm.cc:47-57
            if weight[i] < 1e-5:
                weight[i], freq[i] = 0.0, 0.5
            else:
                nom = inv[i, i] - 1.0
                freq[i] = np.sqrt(abs(-nom / (nom + weight[i]) * 0.25))
        assert rel_frob(golden["inline/ppm_freq"], freq) < 1e-4
        assert rel_frob(golden["inline/ppm_weight"], weight) < 1e-4
    finally:
        ctx.close()


def test_config0_methane_svp_tier_r():
    """BASELINE.json config 0 with own integrals: methane, def2-svp + aux-def2-svp (d shells in the orbital basis, d and
    f in the aux basis), RI-RHF orbitals, default ranges (q = 14), G0W0(ppm) + full BSE singlets and triplets.
    CUDA path against the oracle on identical inputs: QP and BSE energies 1e-6 Ha, oscillator strengths 1e-5."""
    from oracle import gw as ogw
    from oracle import threecenter
    from tests.helpers import methane_svp_case
    from votca_b200.api import Job
    c = methane_svp_case()
    N, q, homo = c["dft"].size, c["q"], c["homo"]
    e, C = c["hf"]["energies"], c["hf"]["mos"]
    vxc = c["hf"]["exchange_mo"][:q, :q]
    vt, ct = homo + 1, q - homo - 1
    inter = obse.free_transition_dipoles(c["dipole"], C, 0, vt, homo + 1, ct)
    job = Job(0)
    job.set_scalar("homo", homo)
    job.set_array("mos", C)
    job.set_array("mo_energies", e)
    job.set_array("vxc", vxc)
    job.set_ao3c(c["ao3c"])
    job.set_array("aux_overlap", c["S"])
    job.set_array("aux_coulomb", c["V"])
    for ax, d in zip("xyz", inter):
        job.set_array("dipole_" + ax, d)
    job.set_options(tasks="gw,singlets,triplets", gw__mode="G0W0", gw__sigma_integrator="ppm", bse__exctotal=5,
                    bse__useTDA=False)
    job.run()
    assert (job.scalar("qpmin"), job.scalar("qpmax"), job.scalar("bse_vmin"), job.scalar("bse_cmax")) == (0, q - 1, 0, q - 1)
    tc = threecenter.TCMatrix(c["aux"].size, 0, q - 1, 0, N - 1)
    tc.fill_from_integrals(c["ao3c"], c["S"], c["V"], C)
    g = ogw.GW(tc, vxc, e)
    g.configure(ogw.GWOptions(homo=homo, qpmin=0, qpmax=q - 1, rpamin=0, rpamax=N - 1, gw_sc_max_iterations=1,
                              sigma_integration="ppm", g_sc_max_iterations=100))
    g.calculate_gw_perturbation()
    g.calculate_hqp()
    assert np.abs(g.get_gwa_results() - job.get("QPpert_energies").ravel()).max() < 1e-6
    b = obse.BSE(tc, factorised=True)
    b.configure(obse.BSEOptions(useTDA=False, homo=homo, rpamin=0, rpamax=N - 1, qpmin=0, qpmax=q - 1, vmin=0,
                                cmax=q - 1, nmax=5, use_Hqp_offdiag=False), g.rpa_input_energies(), g.get_hqp())
    et = b.solve_triplets()
    es = b.solve_singlets()
    assert np.abs(es["eigenvalues"] - job.get("BSE_singlet_eigenvalues").ravel()).max() < 1e-6
    assert np.abs(et["eigenvalues"] - job.get("BSE_triplet_eigenvalues").ravel()).max() < 1e-6
    tdip = obse.coupled_transition_dipoles(es, inter, ct, vt, False)
    f_ref = obse.oscillator_strengths(tdip, es["eigenvalues"])
    # the lowest singlets are the three components of a T2 level: compare the sum over the shell
    assert abs(f_ref.sum() - job.get("oscillator_strengths").sum()) < 1e-5
    job.close()
