"""oracle/bsecoupling.py against the known answers of the reference's test_bsecoupling.cc (coupling_test :41-143,
tb_output_test :145-300): |j_diag| = 23.662750 eV, |j_pert| = 9.529579 eV (BOOST_CHECK_CLOSE 1e-4 %), n_FE = 2,
n_CT = 18."""
import numpy as np

from oracle import bsecoupling as obc
from oracle import threecenter
from tests.helpers import bsecoupling_case


def run_oracle(spin="singlet"):
    c = bsecoupling_case()
    tc = threecenter.TCMatrix(c["basis"].size, 0, 33, 0, 33)
    tc.fill_from_integrals(c["ao3c"], c["S"], c["V"], c["AB_mos"])
    frag = obc.Fragment(mos=c["A_mos"], bse_vmin=0, bse_vmax=4, bse_cmin=5, bse_cmax=16, singlets=c["spsi"],
                        triplets=c["spsi"], singlet_energies=np.array([0.08831]), triplet_energies=np.array([0.08831]))
    coup = obc.BSECoupling(obc.CouplingOptions(spin=spin, levA=1, levB=1, occA=3, unoccA=3, occB=3, unoccB=3))
    coup.calculate_couplings(frag, frag, c["AB_mos"], c["S_dft"], tc, c["Hqp"], c["rpa_energies"], homo=9, rpamin=0,
                             rpamax=33, qpmin=0, qpmax=33, bse_vmin=0, bse_cmax=33, use_Hqp_offdiag=True)
    return c, coup


def test_known_answers_of_the_reference():
    c, coup = run_oracle()
    j_pert = coup.coupling_element("singlet", 0, 0, 0)
    j_diag = coup.coupling_element("singlet", 0, 0, 1)
    ref_diag, ref_pert = c["known_eV"]
    assert abs(abs(j_diag) - ref_diag) / ref_diag < 1e-6 * 100  # BOOST_CHECK_CLOSE(…, 1e-4) is in percent
    assert abs(abs(j_pert) - ref_pert) / ref_pert < 1e-6 * 100
    ch = coup.channels["singlet"]
    assert ch.J_dimer.shape == (20, 20)  # n_FE 2 + n_CT 18
    assert np.allclose(ch.J_dimer, ch.J_dimer.T, atol=1e-10)


def test_limits_are_clamped_and_spin_is_checked():
    import pytest
    with pytest.raises(RuntimeError):
        obc.BSECoupling(obc.CouplingOptions(spin="quintet"))
    c = bsecoupling_case()
    coup = obc.BSECoupling(obc.CouplingOptions(levA=7, levB=7, occA=99, unoccA=-1, occB=2, unoccB=2))
    tc = threecenter.TCMatrix(c["basis"].size, 0, 33, 0, 33)
    tc.fill_from_integrals(c["ao3c"], c["S"], c["V"], c["AB_mos"])
    frag = obc.Fragment(mos=c["A_mos"], bse_vmin=0, bse_vmax=4, bse_cmin=5, bse_cmax=16, singlets=c["spsi"])
    coup.calculate_couplings(frag, frag, c["AB_mos"], c["S_dft"], tc, c["Hqp"], c["rpa_energies"], homo=9, rpamin=0,
                             rpamax=33, qpmin=0, qpmax=33, bse_vmin=0, bse_cmax=33)
    assert (coup.levA, coup.levB, coup.occA, coup.unoccA) == (3, 3, 5, 12)
    n_ct = 5 * 2 + 12 * 2
    assert coup.channels["singlet"].J_dimer.shape == (6 + n_ct, 6 + n_ct)
