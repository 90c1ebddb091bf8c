"""Pins the CPU oracle against the reference's own golden vectors.

Each test mirrors one Boost test of xtp/src/tests and uses the same tolerance
(Eigen isApprox = relative Frobenius norm).
"""
import numpy as np
import pytest

from oracle import bse as obse
from oracle import bse_operator as bop
from oracle import gw as ogw
from oracle import rpa as orpa
from oracle import sigma as osig
from tests.helpers import methane_mmn, rel_frob


# test_threecenter_gwbse.cc:36-126
def test_threecenter_gwbse(golden, methane):
    tc = methane_mmn(golden["threecenter_gwbse/MOs"], mmax=5, nmax=7)
    for i, name in [(0, "ref0b"), (2, "ref2b"), (4, "ref4b")]:
        assert rel_frob(golden["threecenter_gwbse/" + name], tc[i]) < 1e-5
    tc.multiply_right(np.eye(methane["basis"].size))
    for i, name in [(0, "ref0b"), (2, "ref2b"), (4, "ref4b")]:
        assert rel_frob(golden["threecenter_gwbse/" + name], tc[i]) < 1e-5


# test_rpa.cc:41-67
def test_rpa_calcenergies(golden):
    r = orpa.RPA(None)
    r.configure(4, 0, 9)
    r.update_rpa_input_energies(golden["inline/rpa_update_dft"], golden["inline/rpa_update_gw"], 1)
    assert rel_frob(golden["inline/rpa_update_ref"], r.get_rpa_input_energies()) < 1e-4


# test_rpa.cc:69-140
def test_rpa_full(golden):
    tc = methane_mmn(golden["rpa/eigenvectors"])
    r = orpa.RPA(tc)
    r.configure(4, 0, 16)
    r.set_rpa_input_energies(golden["rpa/eigenvals"].ravel())
    assert rel_frob(golden["rpa/i_ref"], r.calculate_epsilon_i(0.5)) < 1e-4
    assert rel_frob(golden["rpa/r_ref"], r.calculate_epsilon_r(0.0)) < 1e-4
    assert rel_frob(golden["rpa/r_complex_ref"], r.calculate_epsilon_r(complex(0.5, 0.5))) < 1e-4


# test_rpa_h2p.cc:36-112
def test_rpa_h2p(golden):
    tc = methane_mmn(golden["rpa/eigenvectors"])
    r = orpa.RPA(tc)
    r.configure(4, 0, 16)
    r.set_rpa_input_energies(golden["rpa/eigenvals"].ravel())
    omega, XpY, erpa = r.diagonalize_h2p()
    assert abs(erpa - float(golden["inline/rpa_h2p_erpa"])) < 1e-6 * abs(erpa) * 100
    assert rel_frob(golden["inline/rpa_h2p_omega"], omega) < 1e-4


@pytest.mark.parametrize("kind,tolx,tolc", [("exact", 1e-5, 1e-5), ("ppm", 1e-5, 1e-5), ("cda", 1e-4, 1e-5)])
def test_sigma(golden, kind, tolx, tolc):
    # test_sigma_exact.cc:43-120, test_sigma_ppm.cc:44-122, test_sigma_cda.cc:41-111
    d = "sigma_" + kind
    e = golden["inline/sigma_ppm_mo_energy"] if kind == "ppm" else golden["inline/sigma_exact_mo_energy"]
    tc = methane_mmn(golden[d + "/MOs"])
    r = orpa.RPA(tc)
    r.configure(4, 0, 16)
    r.set_rpa_input_energies(e)
    s = osig.create(kind, tc, r)
    s.configure(osig.SigmaOptions(homo=4, qpmin=0, qpmax=16, rpamin=0, rpamax=16, eta=1e-3,
                                  quadrature_scheme="legendre", order=100, alpha=1e-3))
    assert rel_frob(golden[d + "/x_ref"], s.calc_exchange_matrix()) < tolx
    s.prepare_screening()
    c = s.calc_correlation_offdiag(e)
    c[np.diag_indices(17)] = s.calc_correlation_diag(e)
    cref = golden[d + "/c_ref"]
    assert rel_frob(np.diag(cref), np.diag(c)) < tolc
    if kind != "cda":  # the reference checks only the diagonal for CDA (off-diagonal returns 0)
        assert rel_frob(cref, c) < tolc


def _gw_options(**kw):
    o = ogw.GWOptions(homo=4, qpmin=0, qpmax=16, rpamin=0, rpamax=16, gw_sc_max_iterations=1, eta=1e-3,
                      sigma_integration="ppm", reset_3c=5, qp_solver="grid", gw_mixing_order=0,
                      gw_mixing_alpha=0.7, g_sc_limit=1e-5, g_sc_max_iterations=50, gw_sc_limit=1e-5)
    for k, v in kw.items():
        setattr(o, k, v)
    return o


# test_gw.cc:111-270
@pytest.mark.parametrize("suffix", ["", "2"])
def test_gw_full(golden, suffix):
    tc = methane_mmn(golden["gw/mo_eigenvectors" + suffix])
    g = ogw.GW(tc, golden["gw/vxc" + suffix], golden["inline/gw_mo_eigenvalues"])
    g.configure(_gw_options(qp_grid_steps=601, qp_grid_spacing=0.005))
    g.calculate_gw_perturbation()
    ref = golden["gw/ref" + suffix]
    assert rel_frob(np.diag(ref), g.get_gwa_results()) < 1e-4
    g.calculate_hqp()
    assert rel_frob(ref, g.get_hqp()) < 1e-4


# test_gw.cc:272-340
def test_gw_canonical_and_brent(golden):
    res = {}
    for finder in ("bisection", "brent"):
        tc = methane_mmn(golden["gw/mo_eigenvectors2"])
        g = ogw.GW(tc, golden["gw/vxc2"], golden["inline/gw_mo_eigenvalues"])
        g.configure(_gw_options(qp_full_window_half_width=1.5, qp_dense_spacing=0.005,
                                qp_adaptive_shell_width=0.02, qp_root_finder=finder))
        g.calculate_gw_perturbation()
        res[finder] = g.get_gwa_results()
    assert rel_frob(np.diag(golden["gw/ref2"]), res["bisection"]) < 1e-4
    assert rel_frob(res["bisection"], res["brent"]) < 1e-5


# test_bse_operator.cc:37-151
def test_bse_operator(golden):
    tc = methane_mmn(golden["bse_operator/MOs"])
    tc.multiply_right(golden["bse_operator/rpa_op"])
    eps = golden["inline/bse_operator_epsilon_inv"]
    opt = bop.BSEOperatorOptions(cmax=8, homo=4, qpmin=0, rpamin=0, vmin=0)
    for name, mk in [("hqp", bop.hqp_op), ("hx", bop.hx_op), ("hd", bop.hd_op), ("hd2", bop.hd2_op)]:
        op = mk(eps, tc, golden["bse_operator/Hqp"])
        op.configure(opt)
        dense = op.dense()
        assert rel_frob(golden["bse_operator/%s_ref" % name], dense) < 1e-3
        assert np.abs(np.diag(dense) - op.diagonal()).max() < 1e-12
        # the factorised formulation (what the GPU path computes) is the same operator
        assert np.abs(op.matmul_factorised(np.eye(op.size)) - dense).max() < 1e-12


def _bse_options(**kw):
    o = obse.BSEOptions(cmax=16, rpamax=16, rpamin=0, vmin=0, nmax=3, useTDA=True, homo=4, qpmin=0, qpmax=16,
                        max_dyn_iter=10, dyn_tolerance=1e-5, davidson_correction="DPR",
                        davidson_tolerance="lapack", davidson_update="safe", davidson_maxiter=50)
    for k, v in kw.items():
        setattr(o, k, v)
    return o


def _subspace(ref, vec):
    return np.linalg.norm(ref.T @ vec, axis=0)


# test_bse.cc:37-377
@pytest.mark.parametrize("factorised", [False, True])
def test_bse(golden, factorised):
    Hqp = golden["bse/Hqp"]
    rpa_e = np.diag(Hqp).copy()
    tc = methane_mmn(golden["bse/MOs"])
    b = obse.BSE(tc, factorised=factorised)
    nrm = lambda a: a / np.linalg.norm(a, axis=0)  # noqa: E731

    b.configure(_bse_options(use_Hqp_offdiag=False), rpa_e, Hqp)
    es = b.solve_singlets()
    assert rel_frob(golden["bse/singlets_nooffdiag_tda"].ravel(), es["eigenvalues"]) < 1e-3
    assert np.allclose(_subspace(golden["bse/singlets_psi_nooffdiag_tda"], es["eigenvectors"]), 1, atol=1e-5)

    b.configure(_bse_options(), rpa_e, Hqp)
    assert rel_frob(Hqp, b.Hqp) < 1e-3
    es = b.solve_singlets()
    assert rel_frob(golden["bse/singlets_tda"].ravel(), es["eigenvalues"]) < 1e-3
    assert np.allclose(_subspace(golden["bse/singlets_psi_tda"], es["eigenvectors"]), 1, atol=1e-5)
    dyn = b.perturbative_dynamical_screening(es, rpa_e)
    assert rel_frob(golden["bse/singlets_dynamic_TDA"].ravel(), dyn) < 5e-3

    b.configure(_bse_options(useTDA=False), rpa_e, Hqp)
    es = b.solve_singlets()
    assert rel_frob(golden["bse/singlets_btda"].ravel(), es["eigenvalues"]) < 1e-3
    assert np.allclose(_subspace(nrm(golden["bse/singlets_psi_btda"]), nrm(es["eigenvectors"])), 1, atol=1e-5)
    assert np.allclose(_subspace(nrm(golden["bse/singlets_psi_AR_btda"]), nrm(es["eigenvectors2"])), 1,
                       atol=1e-5)
    dyn = b.perturbative_dynamical_screening(es, rpa_e)
    assert rel_frob(golden["bse/singlets_dynamic_full"].ravel(), dyn) < 5e-2

    b.configure(_bse_options(nmax=1), rpa_e, Hqp)
    es = b.solve_triplets()
    assert rel_frob(golden["bse/triplets_tda"].ravel(), es["eigenvalues"]) < 1e-3
    dyn = b.perturbative_dynamical_screening(es, rpa_e)
    assert rel_frob(golden["bse/triplets_dynamic_TDA"].ravel(), dyn) < 1e-3

    b.configure(_bse_options(nmax=1, cmax=15, vmin=1), rpa_e, Hqp)
    assert rel_frob(golden["bse/Hqp_cut"], b.Hqp) < 1e-3
    b2 = obse.BSE(tc)
    b2.configure(_bse_options(nmax=1, qpmin=1, qpmax=15), rpa_e, golden["bse/Hqp_cut"])
    assert rel_frob(golden["bse/Hqp_extended"], b2.Hqp) < 1e-3


# ---------------------------------------------------------------------------------------------------------
# AO integrals (the host-side producer of the Fill inputs): test_aomatrix.cc, test_aomatrix3d.cc,
# test_threecenter_dft.cc.  These pin the oracle's McMurchie-Davidson integrals beyond the s,p shells of 3-21G.
# ---------------------------------------------------------------------------------------------------------
def _basis(golden, name, mol):
    import json

    from oracle import basis as obasis
    bs = json.loads(str(golden[f"basis/{name}.json"]))
    bs = {el: [(int(l), [tuple(p) for p in prims]) for l, prims in shells] for el, shells in bs.items()}
    return obasis.AOBasis(bs, [str(e) for e in golden[f"molecule_{mol}/elements"]],
                          golden[f"molecule_{mol}/positions_bohr"])


def _pseudo_invsqrt(V, etol=1e-7):
    w, U = np.linalg.eigh(V)
    d = np.where(w < etol, 0.0, 1.0 / np.sqrt(np.where(w < etol, 1.0, w)))
    return (U * d) @ U.T


# test_aomatrix.cc:42-122 (methane 3-21G: overlap 1e-4, Coulomb 1e-5, Pseudo_InvSqrt_GWBSE 1e-5)
def test_aomatrix_methane(golden, methane):
    from oracle import threecenter
    assert rel_frob(golden["aomatrix/overlap_ref"], methane["S"]) < 1e-4
    assert rel_frob(golden["aomatrix/coulomb_ref"], methane["V"]) < 1e-5
    L, removed = threecenter.pseudo_invsqrt_gwbse(methane["S"], methane["V"], 1e-7)
    assert removed == 0 and rel_frob(golden["aomatrix/coulombinvsqrtgw_ref"], L) < 1e-5


# test_aomatrix.cc:125-149: single-centre overlap of contracted S, P, D, F shells
def test_aomatrix_contracted_spdf(golden):
    from oracle import integrals
    assert rel_frob(golden["aomatrix/overlap_ref_contracted"], integrals.overlap(_basis(golden, "contracted", "C"))) < 1e-4


# test_aomatrix3d.cc:91-123: dipole integrals between G shells on two carbon atoms 1 Angstrom apart
def test_aomatrix3d_g_shell_dipoles(golden):
    from oracle import integrals
    D = integrals.dipole(_basis(golden, "G", "C2"))
    for k in range(3):
        assert rel_frob(golden[f"aomatrix3d/dip_ref_large_{k}"], D[k]) < 1e-4


# test_threecenter_dft.cc:38-73: V^-1/2 (P|mu nu) for methane 3-21G, aux functions 0 and 4
def test_threecenter_dft_small_basis(golden, methane):
    Td = np.einsum("pq,qmn->pmn", _pseudo_invsqrt(methane["V"]), methane["ao3c"])
    assert rel_frob(golden["threecenter_dft/Ref0"], Td[0]) < 1e-5
    assert rel_frob(golden["threecenter_dft/Ref4"], Td[4]) < 1e-5


# test_aomatrix.cc:175-236 and test_threecenter_dft.cc:76-115 ("large_l_test", shipped commented out): I shells
# (l = 6) as aux functions, G shells (l = 4) as orbitals, two carbon atoms.  The shipped data is self-inconsistent
# in ONE function per I shell - index 8 has a self-Coulomb of 0.659 where the other twelve m components have
# 0.17746, and a self-overlap of 0.805 instead of 1 - which is presumably why the test is disabled; everything that
# does not involve that function is a valid known answer and is pinned here at the test's own tolerance (1e-5).
def test_large_l_integrals(golden):
    from oracle import integrals
    aux, dft = _basis(golden, "I", "C2"), _basis(golden, "G", "C2")
    keep = [i for i in range(aux.size) if i % 13 != 8]
    sub = np.ix_(keep, keep)
    bad = golden["aomatrix/coulomb_ref_gi"]
    assert abs(bad[8, 8] - 0.659) < 1e-3 and np.allclose(np.delete(np.diag(bad)[:13], 8), 0.17746, atol=1e-5)
    S, V = integrals.overlap(aux), integrals.coulomb2c(aux)
    assert np.allclose(np.diag(V), V[0, 0], rtol=1e-12)  # every m component of a shell has the same self-Coulomb
    assert rel_frob(golden["aomatrix/overlap_ref_gi"][sub], S[sub]) < 1e-5
    assert rel_frob(bad[sub], V[sub]) < 1e-5
    # three-centre (P | mu nu) in the DFT layout V^-1/2 (P|mu nu): aux rows 1 and 3 do not couple to function 8
    # (odd-m sine components), rows 0 and 12 do and inherit the defect of the shipped data
    Td = np.einsum("pq,qmn->pmn", _pseudo_invsqrt(V), integrals.coulomb3c(aux, dft))
    assert rel_frob(golden["threecenter_dft/RefList1"], Td[1]) < 1e-5
    assert rel_frob(golden["threecenter_dft/RefList2"], Td[3]) < 1e-5
    assert rel_frob(golden["threecenter_dft/RefList0"], Td[0]) > 1e-2  # documents the defect, not a parity claim


# ---------------------------------------------------------------------------------------------------------
# End to end against the reference's own dftgwbse runs (integration-test checkpoints molecule_neutral.orb and
# molecule_neutral_tda.orb, xtp/src/tests/CMakeLists.txt:336-416; the reference compares BSE_singlet/eigenvalues
# at 1e-4 and transition dipoles at 1e-4).  From the checkpoint's MOs and Hqp (its QPdiag eigendecomposition;
# Vxc is not stored, so the QP step itself cannot be replayed) the oracle rebuilds everything else: AO integrals
# with d and f aux shells, Mmn in the GW layout, eps, the BSE operator, Davidson, transition dipoles and the
# perturbative dynamical screening.
# ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("tag", ["neutral", "neutral_tda"])
def test_reference_checkpoint_bse_end_to_end(tag):
    from oracle import threecenter
    from tests.helpers import orb_case, water_integrals
    w, c = water_integrals(), orb_case(tag)
    C = c["mos"]
    assert np.abs(C.T @ w["S_dft"] @ C - np.eye(C.shape[1])).max() < 1e-10  # own overlap vs the reference's MOs
    tc = threecenter.TCMatrix(w["aux"].size, c["rpamin"], max(c["bse_cmax"], c["qpmax"]), c["rpamin"], c["rpamax"])
    tc.fill_from_integrals(w["ao3c"], w["S"], w["V"], C)
    assert np.abs(np.diag(c["Hqp"]) - c["QPpert_energies"].ravel()).max() < 1e-10
    ref_e = c["BSE_singlet_eigenvalues"].ravel()
    b = obse.BSE(tc, factorised=True)
    b.configure(obse.BSEOptions(useTDA=c["useTDA"], homo=c["homo"], rpamin=c["rpamin"], rpamax=c["rpamax"],
                                qpmin=c["qpmin"], qpmax=c["qpmax"], vmin=c["bse_vmin"], cmax=c["bse_cmax"],
                                nmax=len(ref_e), use_Hqp_offdiag=c["use_Hqp_offdiag"], max_dyn_iter=5,
                                dyn_tolerance=1e-5), c["RPA_inputenergies"].ravel(), c["Hqp"])
    es = b.solve_singlets()
    assert np.abs(es["eigenvalues"] - ref_e).max() < 1e-8
    vt, ct = c["homo"] - c["bse_vmin"] + 1, c["bse_cmax"] - c["homo"]
    inter = obse.free_transition_dipoles(w["dipole"], C, c["bse_vmin"], vt, c["homo"] + 1, ct)
    td = obse.coupled_transition_dipoles(es, inter, ct, vt, c["useTDA"])
    assert np.abs(np.abs(td) - np.abs(c["transition_dipoles"])).max() < 1e-8  # sign of an eigenvector is free
    f_ref = obse.oscillator_strengths(c["transition_dipoles"], ref_e)
    assert np.abs(obse.oscillator_strengths(td, es["eigenvalues"]) - f_ref).max() < 1e-8
    dyn = b.perturbative_dynamical_screening(es, c["RPA_inputenergies"].ravel())
    assert np.abs(dyn - c["BSE_singlet_dynamic"].ravel()).max() < 1e-8


# test_aomatrix.cc:62-75 (kinetic, methane 3-21G, 1e-5), :226-235 (kinetic between G shells, shipped disabled),
# test_aopotential.cc:45-56 (nuclear attraction, 1e-5)
def test_kinetic_and_nuclear_attraction(golden, methane):
    from oracle import integrals
    from tests.helpers import NUCLEAR_CHARGE
    assert rel_frob(golden["aomatrix/kinetic_ref"], integrals.kinetic(methane["basis"])) < 1e-5
    assert rel_frob(golden["aomatrix/kinetic_ref_gi"], integrals.kinetic(_basis(golden, "G", "C2"))) < 1e-5
    Z = [NUCLEAR_CHARGE[str(e)] for e in golden["molecule/elements"]]
    V = integrals.nuclear_attraction(methane["basis"], Z, golden["molecule/positions_bohr"])
    assert rel_frob(golden["aopotential/esp_ref"], V) < 1e-5


# test_ppm.cc:36-108: plasmon-pole frequencies and weights for methane with core-Hamiltonian orbitals (1e-4)
def test_ppm_parameters(golden, methane):
    from oracle import threecenter
    from tests.helpers import methane_core_hamiltonian_mos
    e, C = methane_core_hamiltonian_mos()
    tc = threecenter.TCMatrix(methane["basis"].size, 0, 16, 0, 16)
    tc.fill_from_integrals(methane["ao3c"], methane["S"], methane["V"], C)
    r = orpa.RPA(tc)
    r.configure(4, 0, 16)
    r.set_rpa_input_energies(e)
    s = osig.create("ppm", tc, r)
    s.configure(osig.SigmaOptions(homo=4, qpmin=0, qpmax=16, rpamin=0, rpamax=16, eta=1e-3))
    s.prepare_screening()
    assert rel_frob(golden["inline/ppm_freq"], s.ppm_freq) < 1e-4
    assert rel_frob(golden["inline/ppm_weight"], s.ppm_weight) < 1e-4


def test_oracle_boys_function_against_multiprecision():
    """The oracle's Boys function (series + downward recursion, erf + upward recursion beyond x = 35) against
    40-digit arithmetic; it anchors every Coulomb integral above."""
    import mpmath
    from oracle import integrals
    mpmath.mp.dps = 40
    rng = np.random.default_rng(0)
    xs = np.concatenate([[0.0, 1e-9, 0.3, 34.999, 35.0, 35.001, 80.0, 1000.0], rng.uniform(0, 60, 60)])
    F = integrals.boys(16, xs)
    worst = 0.0
    for n in (0, 1, 5, 10, 16):
        for i, x in enumerate(xs):
            ref = mpmath.hyp1f1(n + 0.5, n + 1.5, -mpmath.mpf(float(x))) / (2 * n + 1)
            worst = max(worst, float(abs((mpmath.mpf(float(F[n, i])) - ref) / ref)))
    assert worst < 1e-14, worst


# test_populationanalysis.cc:38-74 (atompop) and :95-180 (fragment_pop)
def test_lowdin_charges_and_fragment_populations_known_answers(golden):
    from oracle import population as opop
    from tests.helpers import methane_integrals
    m = methane_integrals()
    S = m["S"]
    basis_atom = np.concatenate([[sh.atom] * (2 * sh.l + 1) for sh in m["basis"].shells])
    nuc = np.array([6.0, 1.0, 1.0, 1.0, 1.0])
    MOs = golden["populationanalysis/MOs"]
    charges = nuc - opop.lowdin_per_atom(opop.ground_state_density(MOs, 5), S, basis_atom, 5)
    ref = np.array([0.68862, -0.172155, -0.172154, -0.172154, -0.172155])
    assert np.abs(charges - ref).max() < 1e-5 * np.abs(ref).max() * 5
    MOs2, spsi = golden["populationanalysis/MOs2"], golden["populationanalysis/spsi_ref"]
    Gs, H, E = opop.fragment_populations(S, MOs2, 4, 0, 16, basis_atom, nuc, [[0, 1], [2, 3, 4]], spsi)
    assert abs(Gs[0] - 0.5164649) < 1e-5 and abs(Gs[1] + 0.5164628) < 1e-5
    assert np.abs(E[0] - [-0.384176, -0.812396, -0.414518]).max() < 1e-5
    assert np.abs(H[0] - [0.657215, 0.622434, 0.654751]).max() < 1e-5
    assert np.abs(E[1] - [-0.615827, -0.187602, -0.58548]).max() < 1e-5
    assert np.abs(H[1] - [0.342785, 0.377565, 0.34525]).max() < 1e-5
