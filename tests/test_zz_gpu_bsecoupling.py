"""BSECoupling through the CUDA path: host class votca_b200/host/bsecoupling.h driven through the
job facade, against (a) the known answers of the reference's own test_bsecoupling.cc and (b) the oracle's matrices.
The projection P = [Frenkel | CT], H P, J = P^T H P and S = P^T P are all formed on the device."""
import re

import numpy as np
import pytest

from tests.helpers import bsecoupling_case
from tests.test_oracle_bsecoupling import run_oracle

pytestmark = pytest.mark.gpu


def make_job(c, spin="singlet", tb=False, from_basis=False, **levels):
    from votca_b200.api import Job
    job = Job(0)
    lv = dict(statesA=1, occA=3, unoccA=3, statesB=1, occB=3, unoccB=3)
    lv.update(levels)
    job.set_option("bsecoupling.spin", spin)
    job.set_option("bsecoupling.use_perturbation", True)
    job.set_option("bsecoupling.output_tb", tb)
    for m in "AB":
        job.set_option(f"bsecoupling.molecule{m}.states", lv["states" + m])
        job.set_option(f"bsecoupling.molecule{m}.occLevels", lv["occ" + m])
        job.set_option(f"bsecoupling.molecule{m}.unoccLevels", lv["unocc" + m])
        job.set_array(f"{m}.mos", c["A_mos"])
        job.set_array(f"{m}.BSE_singlet_eigenvectors", c["spsi"])
        job.set_array(f"{m}.BSE_triplet_eigenvectors", c["spsi"])
        job.set_array(f"{m}.BSE_singlet_eigenvalues", np.full(3, 0.08831))
        for k, v in (("bse_vmin", 0), ("bse_vmax", 4), ("bse_cmin", 5), ("bse_cmax", 16)):
            job.set_scalar(f"{m}.{k}", v)
    job.set_array("mos", c["AB_mos"])
    job.set_array("Hqp", c["Hqp"])
    job.set_array("RPA_inputenergies", c["rpa_energies"])
    for k, v in (("homo", 9), ("rpamin", 0), ("rpamax", 33), ("qpmin", 0), ("qpmax", 33), ("bse_vmin", 0),
                 ("bse_cmax", 33), ("use_Hqp_offdiag", 1)):
        job.set_scalar(k, v)
    if from_basis:
        from tests.test_ao3c_core_cpu import pack
        for which in ("dft", "aux"):
            job.set_basis(which, *pack(c["basis"]))
    else:
        job.set_ao3c(c["ao3c"])
        job.set_array("aux_overlap", c["S"])
        job.set_array("aux_coulomb", c["V"])
        job.set_array("dft_overlap", c["S_dft"])
    return job


def test_known_answers_of_the_reference_through_cuda():
    """test_bsecoupling.cc:138-143: |j_diag| = 23.662750 eV, |j_pert| = 9.529579 eV to 1e-4 percent."""
    c = bsecoupling_case()
    job = make_job(c)
    job.run_coupling()
    hrt2ev = 27.21138602
    j_pert = job.get("JAB_singlet_pert")[0, 1] * hrt2ev
    j_diag = job.get("JAB_singlet_diag")[0, 1] * hrt2ev
    ref_diag, ref_pert = c["known_eV"]
    assert abs(abs(j_diag) - ref_diag) / ref_diag < 1e-6
    assert abs(abs(j_pert) - ref_pert) / ref_pert < 1e-6
    xml = job.coupling_xml()
    m = re.search(r'j_pert="([^"]+)" j_diag="([^"]+)"', xml)
    assert abs(abs(float(m.group(1))) - ref_pert) / ref_pert < 1e-6
    assert abs(abs(float(m.group(2))) - ref_diag) / ref_diag < 1e-6
    job.close()


def test_matrices_against_the_oracle_and_tb_output():
    """J_dimer, S_dimer (20 x 20: 2 Frenkel + 18 CT states, test_bsecoupling.cc:296-299) and both coupling matrices
    against the oracle; singlets and triplets; tb_matrices node of the XML."""
    c, coup = run_oracle(spin="all")
    job = make_job(c, spin="all", tb=True)
    job.run_coupling()
    for spin in ("singlet", "triplet"):
        ch = coup.channels[spin]
        assert np.abs(job.get(f"J_dimer_{spin}") - ch.J_dimer).max() < 1e-9
        assert np.abs(job.get(f"S_dimer_{spin}") - ch.S_dimer).max() < 1e-10
        assert np.abs(job.get(f"JAB_{spin}_pert") - ch.JAB[0]).max() < 1e-8
        assert np.abs(job.get(f"JAB_{spin}_diag") - ch.JAB[1]).max() < 1e-8
    xml = job.coupling_xml()
    assert 'n_FE="2"' in xml and 'n_CT="18"' in xml and "<triplet" in xml and "H_CT_CT" in xml
    job.close()


def test_integrals_and_overlap_from_the_device_basis():
    """No integral arrays at all: (P|mu nu), (P|Q), <P|Q> and the dimer's AO overlap come from the device kernels."""
    c = bsecoupling_case()
    job = make_job(c, from_basis=True)
    job.run_coupling()
    hrt2ev = 27.21138602
    ref_diag, ref_pert = c["known_eV"]
    assert abs(abs(job.get("JAB_singlet_diag")[0, 1]) * hrt2ev - ref_diag) / ref_diag < 1e-6
    assert abs(abs(job.get("JAB_singlet_pert")[0, 1]) * hrt2ev - ref_pert) / ref_pert < 1e-6
    job.close()


def test_requests_beyond_what_is_stored_are_clamped():
    c = bsecoupling_case()
    job = make_job(c, statesA=7, statesB=2, occA=99, unoccA=-1, occB=2, unoccB=2)
    job.run_coupling()
    assert job.scalar("levA") == 3 and job.scalar("levB") == 2
    n = 5 + 5 * 2 + 12 * 2
    assert job.get("J_dimer_singlet").shape == (n, n)
    with pytest.raises(Exception, match="not a key of bsecoupling.xml"):
        job.set_option("bsecoupling.moleculeC.states", 1)
    job.close()
    bad = make_job(c, spin="quintet")
    with pytest.raises(Exception, match="not known"):
        bad.run_coupling()
    bad.close()
