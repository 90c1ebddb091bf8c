"""oracle/orbfile.py (minimal HDF5 reader for .orb checkpoints) against the reference's checked-in files.  Needs
/root/reference, so it only runs in the build container; the arrays it extracts travel in tests/golden."""
import os

import numpy as np
import pytest

from oracle.orbfile import OrbFile
from tests.helpers import load_golden

IT = "/root/reference/xtp/src/tests/DataFiles/xtp_tools_integration_tests"
pytestmark = pytest.mark.skipif(not os.path.isdir(IT), reason="reference tree not available")


def test_reads_groups_datasets_and_attributes():
    f = OrbFile(os.path.join(IT, "molecule_neutral.orb"))
    assert f.keys("/") == ["QMdata"]
    top = f.keys("/QMdata")  # densely stored group (fractal heap)
    for k in ("mos", "QPdiag", "BSE_singlet", "RPA_inputenergies", "transition_dipoles", "dft", "aux", "qmmolecule"):
        assert k in top
    assert f.keys("/QMdata/mos") == ["eigenvalues", "eigenvectors", "eigenvectors2"]  # compact link storage
    at = f.attrs("/QMdata")  # densely stored attributes
    assert at["qm_package"] == "xtp" and at["XCFunctional"] == "XC_HYB_GGA_XC_PBEH"
    assert (at["rpamin"], at["rpamax"], at["qpmin"], at["qpmax"], at["bse_vmin"], at["bse_cmax"]) == (0, 12, 0, 12, 0, 12)
    assert at["occupied_levels"] == 5 and at["ScaHFX"] == 0.25 and at["useTDA"] == 0
    assert f.attrs("/QMdata/dft")["name"] == "3-21G" and f.attrs("/QMdata/aux")["basissize"] == 76
    e = f.read("/QMdata/mos/eigenvalues")
    C = f.read("/QMdata/mos/eigenvectors")
    assert e.shape == (13, 1) and C.shape == (13, 13) and np.all(np.diff(e.ravel()) > 0)
    assert f.read("/QMdata/BSE_triplet/eigenvalues").size == 0
    assert len([p for p in f.walk("/QMdata") if p.startswith("/QMdata/transition_dipoles/")]) == 10


@pytest.mark.parametrize("tag", ["neutral", "neutral_tda"])
def test_golden_npz_matches_the_checkpoints(tag):
    f = OrbFile(os.path.join(IT, f"molecule_{tag}.orb"))
    g = load_golden()
    for name, path in (("mos", "mos/eigenvectors"), ("QPdiag_eigenvalues", "QPdiag/eigenvalues"),
                       ("BSE_singlet_eigenvalues", "BSE_singlet/eigenvalues"), ("BSE_singlet_dynamic", "BSE_singlet_dynamic")):
        assert np.array_equal(g[f"orb/{tag}/{name}"], f.read("/QMdata/" + path))


def test_other_checkpoints_parse():
    for fn in ("molecule_ch4.orb", "molecule_cation.orb", "molecule_cation_tda.orb"):
        f = OrbFile(os.path.join(IT, fn))
        assert "QPpert_energies" in f.keys("/QMdata") and f.read("/QMdata/mos/eigenvalues").size > 0
