"""Real-chemistry ("tier R", own integrals) case of BASELINE.json config 0 on the CPU: methane, def2-svp +
aux-def2-svp, RI-RHF orbitals (oracle/scf.py), G0W0(ppm) + full BSE with the oracle.  Checks the pieces against
each other and against textbook numbers; the same inputs go through the CUDA path in
tests/test_gpu_zz_reference_checkpoint.py."""
import numpy as np

from oracle import bse as obse
from oracle import gw as ogw
from oracle import rpa as orpa
from oracle import sigma as osig
from oracle import threecenter
from tests.helpers import methane_svp_case

HARTREE_EV = 27.211386


def test_rhf_ri_methane():
    c = methane_svp_case()
    hf = c["hf"]
    assert (c["dft"].size, c["aux"].size, c["homo"], c["q"]) == (34, 104, 4, 14)
    assert abs(hf["total_energy"] + 40.1699) < 2e-3  # RHF/def2-SVP methane
    C = hf["mos"]
    assert np.abs(C.T @ hf["overlap"] @ C - np.eye(34)).max() < 1e-10
    e = hf["energies"]
    assert abs(e[0] + 11.22) < 0.02 and abs(e[4] * HARTREE_EV + 14.85) < 0.1 and e[5] > 0.15


def test_sigma_x_equals_hf_exchange():
    """Sigma_x from the Mmn path (fill + Pseudo_InvSqrt_GWBSE + CalcExchangeMatrix) against -K/2 of the SCF, which
    contracts the same three-centre integrals in the AO basis: an independent check of the fill and of sigma_base.cc:36-52."""
    c = methane_svp_case()
    N, q = c["dft"].size, c["q"]
    tc = threecenter.TCMatrix(c["aux"].size, 0, q - 1, 0, N - 1)
    tc.fill_from_integrals(c["ao3c"], c["S"], c["V"], c["hf"]["mos"])
    r = orpa.RPA(tc)
    r.configure(c["homo"], 0, N - 1)
    r.set_rpa_input_energies(c["hf"]["energies"])
    s = osig.create("ppm", tc, r)
    s.configure(osig.SigmaOptions(homo=c["homo"], qpmin=0, qpmax=q - 1, rpamin=0, rpamax=N - 1))
    assert np.abs(s.calc_exchange_matrix() - c["hf"]["exchange_mo"][:q, :q]).max() < 1e-7


def test_g0w0_bse_methane_is_physical():
    c = methane_svp_case()
    N, q, homo = c["dft"].size, c["q"], c["homo"]
    tc = threecenter.TCMatrix(c["aux"].size, 0, q - 1, 0, N - 1)
    tc.fill_from_integrals(c["ao3c"], c["S"], c["V"], c["hf"]["mos"])
    g = ogw.GW(tc, c["hf"]["exchange_mo"][:q, :q], c["hf"]["energies"])
    g.configure(ogw.GWOptions(homo=homo, qpmin=0, qpmax=q - 1, rpamin=0, rpamax=N - 1, gw_sc_max_iterations=1,
                              sigma_integration="ppm", g_sc_max_iterations=100))
    g.calculate_gw_perturbation()
    g.calculate_hqp()
    ip = -g.get_gwa_results()[homo] * HARTREE_EV
    assert 13.8 < ip < 15.0  # methane vertical IP: 14.35 eV (experiment), G0W0@HF/def2-SVP 14.4 eV
    b = obse.BSE(tc, factorised=True)
    b.configure(obse.BSEOptions(useTDA=False, homo=homo, rpamin=0, rpamax=N - 1, qpmin=0, qpmax=q - 1, vmin=0,
                                cmax=q - 1, nmax=5, use_Hqp_offdiag=False), g.rpa_input_energies(), g.get_hqp())
    es, et = b.solve_singlets(), b.solve_triplets()
    assert np.all(np.diff(es["eigenvalues"]) >= -1e-10) and es["eigenvalues"][0] > 0.3
    assert et["eigenvalues"][0] < es["eigenvalues"][0]  # triplets below singlets
    # the three lowest singlets / triplets are the components of a T2 level
    assert np.ptp(es["eigenvalues"][:3]) < 2e-3 and np.ptp(et["eigenvalues"][:3]) < 4e-3
