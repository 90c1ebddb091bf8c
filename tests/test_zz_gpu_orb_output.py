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
