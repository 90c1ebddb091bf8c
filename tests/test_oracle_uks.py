"""oracle/uks.py (the UKS twins).  The reference has no unit test for these classes, so the restatement is held to the
restricted oracle - itself pinned on the reference's rpa/, sigma_ppm/, gw/ and bse_operator/ fixtures - in the
closed-shell limit, where every UKS formula must collapse to its restricted counterpart."""
import numpy as np

from oracle import bse as obse
from oracle import bse_operator as bop
from oracle import gw as ogw
from oracle import rpa as orpa
from oracle import uks
from tests.helpers import load_golden, methane_mmn, uks_case
from tests.test_oracle_golden import _gw_options


def test_spin_summed_epsilon_equals_restricted_in_the_closed_shell_limit():
    g = load_golden()
    C, e = g["gw/mo_eigenvectors"], g["inline/gw_mo_eigenvalues"]
    r = orpa.RPA(methane_mmn(C))
    r.configure(4, 0, 16)
    r.set_rpa_input_energies(e)
    u = uks.RPAUKS(methane_mmn(C), methane_mmn(C))
    u.configure(4, 4, 0, 16)
    u.set_rpa_input_energies(e, e)
    for f in (0.0, 0.5):
        assert np.abs(u.calculate_epsilon_i(f) - r.calculate_epsilon_i(f)).max() < 1e-12
        assert np.abs(u.calculate_epsilon_r(f) - r.calculate_epsilon_r(f)).max() < 1e-12
    z = complex(0.3, 0.2)
    assert np.abs(u.calculate_epsilon_r(z) - r.calculate_epsilon_r(z)).max() < 1e-12
    # energies update: identical to the restricted rule when rpamin = 0 (rpa.cc:52-62 counts the head from 0)
    gwa = e[2:12] + 0.01 * np.arange(10)
    r.update_rpa_input_energies(e, gwa, 2)
    u.update_rpa_input_energies(e, e, gwa, gwa, 2)
    assert np.abs(u.energies(0) - r.get_rpa_input_energies()).max() < 1e-15
    assert np.abs(u.energies(1) - r.get_rpa_input_energies()).max() < 1e-15


def test_gw_uks_reproduces_restricted_gw_in_the_closed_shell_limit():
    """evGW on both: the restricted result is the reference's gw/ref fixture (tests/test_oracle_golden.py)."""
    g = load_golden()
    C, e, vxc = g["gw/mo_eigenvectors"], g["inline/gw_mo_eigenvalues"], g["gw/vxc"]
    opt = _gw_options(qp_grid_steps=601, qp_grid_spacing=0.005)
    rg = ogw.GW(methane_mmn(C), vxc, e)
    rg.configure(opt)
    rg.calculate_gw_perturbation()
    rg.calculate_hqp()
    ug = uks.GWUKS(methane_mmn(C), methane_mmn(C), vxc, vxc, e, e)
    ug.configure(_gw_options(qp_grid_steps=601, qp_grid_spacing=0.005), 4, 4)
    ug.calculate_gw_perturbation()
    ug.calculate_hqp()
    for s in range(2):
        assert np.abs(ug.get_gwa_results(s) - rg.get_gwa_results()).max() < 1e-10
        assert np.abs(ug.get_hqp(s) - rg.get_hqp()).max() < 1e-10
    assert np.abs(np.diag(g["gw/ref"]) - ug.get_gwa_results(0)).max() / np.abs(np.diag(g["gw/ref"])).max() < 1e-4


def test_uks_operator_blocks_reduce_to_the_restricted_operators():
    """Same-spin blocks of <1,1,1,0> against the sum of the restricted <1,0,0,0>, <0,1,0,0>, <0,0,1,0> operators (pinned on bse_operator/*_ref), the
    cross-spin block against the screened transition-density form, diagonal against the dense matrix, symmetry."""
    g = load_golden()
    C = g["bse_operator/MOs"]
    eps = g["inline/bse_operator_epsilon_inv"]
    Hqp = g["bse_operator/Hqp"]
    Ma = methane_mmn(C)
    Ma.multiply_right(g["bse_operator/rpa_op"])
    Mb = methane_mmn(C)
    Mb.multiply_right(g["bse_operator/rpa_op"])
    op = uks.exciton_uks_tda(eps, Ma, Mb, Hqp, Hqp)
    op.configure(4, 4, 0, 0, 8)
    D = op.dense()
    n = op.blk[0].size
    ropt = bop.BSEOperatorOptions(cmax=8, homo=4, qpmin=0, rpamin=0, vmin=0)
    parts = {}
    for name, mk in (("hqp", bop.hqp_op), ("hx", bop.hx_op), ("hd", bop.hd_op)):
        r = mk(eps, Ma, Hqp)
        r.configure(ropt)
        parts[name] = r.dense()
    same = parts["hqp"] + parts["hx"] + parts["hd"]  # BSE_OPERATOR<0,0,1,0> is -Hd already
    assert np.abs(D[:n, :n] - same).max() < 1e-12 and np.abs(D[n:, n:] - same).max() < 1e-12
    A = np.vstack([Ma[v][5:9, :] for v in range(5)])  # rows (v, c), the transition densities
    cross = -(A * eps) @ A.T
    assert np.abs(D[:n, n:] - cross).max() < 1e-12 and np.abs(D[n:, :n] - cross).max() < 1e-12
    assert np.abs(D - D.T).max() < 1e-12
    assert np.abs(np.diag(D) - op.diagonal()).max() < 1e-12


def test_open_shell_case_runs_and_is_spin_asymmetric():
    c = uks_case()
    ug = uks.GWUKS(methane_mmn(c["Ca"]), methane_mmn(c["Cb"]), c["vxc_a"], c["vxc_b"], c["ea"], c["eb"])
    ug.configure(_gw_options(qp_grid_steps=601, qp_grid_spacing=0.005, gw_sc_max_iterations=1), c["homo_a"], c["homo_b"])
    ug.calculate_gw_perturbation()
    qa, qb = ug.get_gwa_results(0), ug.get_gwa_results(1)
    assert np.all(np.isfinite(qa)) and np.all(np.isfinite(qb)) and np.abs(qa - qb).max() > 1e-3
    ug.calculate_hqp()
    bs = uks.BSEUKS(methane_mmn(c["Ca"]), methane_mmn(c["Cb"]))
    o = obse.BSEOptions(cmax=16, rpamax=16, rpamin=0, vmin=0, nmax=3, useTDA=True, homo=4, qpmin=0, qpmax=16,
                        davidson_tolerance="lapack", davidson_maxiter=50)
    bs.configure(o, c["homo_a"], c["homo_b"], ug.rpa.energies(0), ug.rpa.energies(1), ug.get_hqp(0), ug.get_hqp(1))
    res = bs.solve_excitons_uks_tda()
    H = bs.operator_tda().dense()
    assert H.shape == (5 * 12 + 4 * 13,) * 2
    w = np.linalg.eigvalsh(H)
    assert np.abs(res["eigenvalues"][:3] - w[:3]).max() < 1e-8


def test_uks_b_operator_reduces_to_restricted_blocks_and_full_bse_is_consistent():
    """<0,1,0,1>: same-spin blocks = restricted <0,1,0,0> + <0,0,0,1>; in the closed-shell limit the cross-spin block
    equals the same-spin Hd2 block (both channels' tensors are the same); the dense full-BSE roots are real and positive
    and normalised to |X.X - Y.Y| = 1 (the reference divides by sqrt(abs(norm)), bse_uks.cc:386-392)."""
    g = load_golden()
    C, eps, Hqp = g["bse_operator/MOs"], g["inline/bse_operator_epsilon_inv"], g["bse_operator/Hqp"]
    Ma, Mb = methane_mmn(C), methane_mmn(C)
    Ma.multiply_right(g["bse_operator/rpa_op"])
    Mb.multiply_right(g["bse_operator/rpa_op"])
    op = uks.exciton_uks_btda_b(eps, Ma, Mb, Hqp, Hqp)
    op.configure(4, 4, 0, 0, 8)
    D = op.dense()
    n = op.blk[0].size
    ropt = bop.BSEOperatorOptions(cmax=8, homo=4, qpmin=0, rpamin=0, vmin=0)
    hx, hd2 = bop.hx_op(eps, Ma, Hqp), bop.hd2_op(eps, Ma, Hqp)
    hx.configure(ropt)
    hd2.configure(ropt)
    assert np.abs(D[:n, :n] - (hx.dense() + hd2.dense())).max() < 1e-12
    assert np.abs(D[:n, n:] - hd2.dense()).max() < 1e-12 and np.abs(D[n:, :n] - hd2.dense()).max() < 1e-12
    assert np.abs(np.diag(D) - op.diagonal()).max() < 1e-12
    c = uks_case()
    ug = uks.GWUKS(methane_mmn(c["Ca"]), methane_mmn(c["Cb"]), c["vxc_a"], c["vxc_b"], c["ea"], c["eb"])
    ug.configure(_gw_options(qp_grid_steps=601, qp_grid_spacing=0.005), c["homo_a"], c["homo_b"])
    ug.calculate_gw_perturbation()
    ug.calculate_hqp()
    bs = uks.BSEUKS(methane_mmn(c["Ca"]), methane_mmn(c["Cb"]))
    o = obse.BSEOptions(cmax=16, rpamax=16, rpamin=0, vmin=0, nmax=5, useTDA=False, homo=4, qpmin=0, qpmax=16)
    bs.configure(o, c["homo_a"], c["homo_b"], ug.rpa.energies(0), ug.rpa.energies(1), ug.get_hqp(0), ug.get_hqp(1))
    full = bs.solve_excitons_uks_btda_dense()
    assert np.all(full["eigenvalues"] > 0) and np.all(np.diff(full["eigenvalues"]) >= 0)
    X, Y = full["eigenvectors"], full["eigenvectors2"]
    assert np.abs(np.abs(np.sum(X * X, axis=0) - np.sum(Y * Y, axis=0)) - 1.0).max() < 1e-10


def test_exact_and_cda_uks_evaluators_reduce_to_the_restricted_ones():
    """sigma_exact_uks.cc / sigma_cda_uks.cc in the closed-shell limit: the unrestricted H2p has the restricted
    eigenvalues on its spin-symmetric modes and as many dark (spin-antisymmetric) ones, which the screening-mode
    filter of rpa_uks.cc:135-150 removes; residues grow by sqrt(2) while the closed-shell factor 2 is gone, so every
    element equals the restricted evaluator's (pinned on the reference's sigma_exact/ and sigma_cda/ fixtures)."""
    from oracle import sigma as osig
    g = load_golden()
    C, e = g["gw/mo_eigenvectors"], g["inline/gw_mo_eigenvalues"]
    r = orpa.RPA(methane_mmn(C))
    r.configure(4, 0, 16)
    r.set_rpa_input_energies(e)
    u = uks.RPAUKS(methane_mmn(C), methane_mmn(C))
    u.configure(4, 4, 0, 16)
    u.set_rpa_input_energies(e, e)
    om_r, _, erpa_r = r.diagonalize_h2p()
    om_u, modes = u.screening_modes()
    # twice the states; at least the dark half is filtered (mixtures inside degenerate eigenspaces may survive with
    # partial weight - the pole sums below are invariant under that)
    assert len(u.diagonalize_h2p()[0]) == 2 * len(om_r) and len(om_r) <= len(om_u) < 2 * len(om_r)
    assert set(np.round(om_r, 6)) <= set(np.round(om_u, 6))
    sopt = osig.SigmaOptions(homo=4, qpmin=0, qpmax=16, rpamin=0, rpamax=16, eta=1e-3, quadrature_scheme="legendre",
                             order=12, alpha=1e-3)
    freqs = e[:17] + 0.013
    for name, mk in (("exact", lambda M: uks.SigmaExactUKS(M, u.spin[0], u)), ("cda", lambda M: uks.SigmaCDAUKS(M, u, 0))):
        M1, M2 = methane_mmn(C), methane_mmn(C)
        sr = osig.create(name, M1, r)
        su = mk(M2)
        for s in (sr, su):
            s.configure(sopt)
            s.prepare_screening()
        for lvl in (0, 4, 5, 11):
            a, b = sr.calc_correlation_diag_element(lvl, freqs[lvl]), su.calc_correlation_diag_element(lvl, freqs[lvl])
            assert abs(a - b) < 1e-9 * max(1.0, abs(a)), (name, lvl, a, b)
        if name == "exact":
            assert abs(sr.calc_correlation_diag_element_derivative(3, freqs[3]) -
                       su.calc_correlation_diag_element_derivative(3, freqs[3])) < 1e-8
            assert abs(sr.calc_correlation_offdiag_element(2, 7, freqs[2], freqs[7]) -
                       su.calc_correlation_offdiag_element(2, 7, freqs[2], freqs[7])) < 1e-10


def test_open_shell_gw_with_the_exact_and_cda_evaluators_runs():
    c = uks_case()
    got = {}
    for name in ("exact", "cda"):
        ug = uks.GWUKS(methane_mmn(c["Ca"]), methane_mmn(c["Cb"]), c["vxc_a"], c["vxc_b"], c["ea"], c["eb"])
        ug.configure(_gw_options(qp_grid_steps=201, qp_grid_spacing=0.01, gw_sc_max_iterations=1,
                                 sigma_integration=name), c["homo_a"], c["homo_b"])
        ug.calculate_gw_perturbation()
        got[name] = (ug.get_gwa_results(0), ug.get_gwa_results(1))
        assert np.all(np.isfinite(got[name][0])) and np.abs(got[name][0] - got[name][1]).max() > 1e-3
    # both integrate the same W: around the gap the quasiparticle energies agree to the accuracy of the quadrature
    # (high virtual levels sit among the poles of W, where the two root searches may settle on different roots)
    assert np.abs(got["exact"][0][:8] - got["cda"][0][:8]).max() < 5e-3
