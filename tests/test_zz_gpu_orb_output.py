"""GPU: a job writes its results as an .orb (HDF5) checkpoint (gwbse_job_set_orb_output) whose datasets equal
the arrays the job returns, under the names Orbitals::WriteToCpt uses (orbitals.cc:990-1063).  The writer itself
is verified on the CPU (tests/test_checkpoint_writer_cpu.py)."""
import numpy as np
import pytest

from oracle.orbfile import OrbFile


pytestmark = pytest.mark.gpu


def test_job_writes_orb_checkpoint(tmp_path):
    from votca_b200 import synthetic
    from votca_b200.api import Job
    N, naux, homo = synthetic.CONFIGS["tiny"]
    s = synthetic.make_small(N, naux, homo)
    path = tmp_path / "tiny.orb"
    job = Job(0)
    job.set_scalar("homo", homo)
    for name in ("mos", "mo_energies", "vxc", "aux_overlap", "aux_coulomb"):
        job.set_array(name, s[name])
    job.set_ao3c(s["ao3c"])
    job.set_options(tasks="gw,singlets", gw__mode="G0W0", bse__useTDA=False, bse__exctotal=4)
    job.set_orb_output(path)
    job.run()
    f = OrbFile(str(path))
    f.verify_checksums()
    at = f.attrs("/QMdata")
    assert at["occupied_levels"] == homo + 1 and at["useTDA"] == 0 and at["version"] == 9
    for k in ("rpamin", "rpamax", "qpmin", "qpmax", "bse_vmin", "bse_cmax"):
        assert at[k] == int(job.scalar(k))
    assert np.array_equal(f.read("/QMdata/mos/eigenvectors"), s["mos"])
    assert np.array_equal(f.read("/QMdata/QPpert_energies").ravel(), job.get("QPpert_energies").ravel())
    assert np.array_equal(f.read("/QMdata/QPdiag/eigenvectors"), job.get("QPdiag_eigenvectors"))
    assert np.array_equal(f.read("/QMdata/BSE_singlet/eigenvalues").ravel(), job.get("BSE_singlet_eigenvalues").ravel())
    assert np.array_equal(f.read("/QMdata/BSE_singlet/eigenvectors"), job.get("BSE_singlet_eigenvectors"))
    assert np.array_equal(f.read("/QMdata/BSE_singlet/eigenvectors2"), job.get("BSE_singlet_eigenvectors2"))
    assert f.read("/QMdata/BSE_triplet/eigenvalues").shape == (0, 1)
    job.close()


def test_job_writes_summary_xml(tmp_path):
    """<job>_summary.xml as the dftgwbse tool writes it (GWBSE::addoutput, gwbse.cc:580-738; layout of the reference's
    own xtp-tutorials/pyxtp/files_examples/methane_summary.xml): same element tree, attribute order, number formats
    (%+1.6f eV, %+1.4f e*bohr), values equal to the arrays the job returns."""
    import re
    import xml.etree.ElementTree as ET
    from tests.helpers import methane_svp_case
    from oracle import bse as obse
    from votca_b200.api import Job
    c = methane_svp_case()
    q, homo = c["q"], c["homo"]
    vt, ct = homo + 1, q - homo - 1
    job = Job(0)
    job.set_scalar("homo", homo)
    job.set_scalar("dft_total_energy", c["hf"]["total_energy"])
    job.set_array("mos", c["hf"]["mos"])
    job.set_array("mo_energies", c["hf"]["energies"])
    job.set_array("vxc", c["hf"]["exchange_mo"][:q, :q])
    job.set_ao3c(c["ao3c"])
    job.set_array("aux_overlap", c["S"])
    job.set_array("aux_coulomb", c["V"])
    for ax, d in zip("xyz", obse.free_transition_dipoles(c["dipole"], c["hf"]["mos"], 0, vt, homo + 1, ct)):
        job.set_array("dipole_" + ax, d)
    job.set_options(tasks="gw,singlets,triplets", gw__mode="G0W0", bse__exctotal=4, bse__useTDA=True)
    path = tmp_path / "methane_summary.xml"
    job.set_summary_output(path)
    job.run()
    text = path.read_text()
    hrt2ev = 27.21138602
    assert text.startswith('<output>\n\t<GWBSE DFTEnergy="%+1.6f " units="eV">\n\t\t<dft HOMO="4" LUMO="5">\n'
                           % (c["hf"]["total_energy"] * hrt2ev))
    assert text.endswith("\t</GWBSE>\n</output>\n")
    root = ET.fromstring(text)
    levels = root.findall("./GWBSE/dft/level")
    assert [int(lv.get("number")) for lv in levels] == list(range(q))
    for i, lv in enumerate(levels):
        assert lv.find("dft_energy").text == "%+1.6f " % (c["hf"]["energies"][i] * hrt2ev)
        assert lv.find("gw_energy").text == "%+1.6f " % (job.get("QPpert_energies")[i] * hrt2ev)
        assert lv.find("qp_energy").text == "%+1.6f " % (job.get("QPdiag_eigenvalues")[i] * hrt2ev)
    sing = root.findall("./GWBSE/singlets/level")
    assert len(sing) == 4 and [lv.get("number") for lv in sing] == ["1", "2", "3", "4"]
    f, td, es = job.get("oscillator_strengths"), job.get("transition_dipoles"), job.get("BSE_singlet_eigenvalues")
    for i, lv in enumerate(sing):
        assert lv.find("omega").text == "%+1.6f " % (es[i] * hrt2ev)
        assert abs(float(lv.find("f").text) - f[i]) < 1e-6
        assert lv.find("Trdipole").text == "%+1.4f %+1.4f %+1.4f" % tuple(td[:, i])
        assert re.search(r'<Trdipole gauge="length" unit="e\*bohr">', text)
    trip = root.findall("./GWBSE/triplets/level")
    assert len(trip) == 4 and trip[0].find("omega").text == "%+1.6f " % (job.get("BSE_triplet_eigenvalues")[0] * hrt2ev)
    assert trip[0].find("f") is None
    job.close()
