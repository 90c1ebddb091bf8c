"""CPU checks of the tier-R input plumbing (votca_b200/realsys.py): basis-set data extracted from the reference's
library gives the function counts of SURVEY.md section 8, shell arrays equal the oracle's AOBasis (order and
normalised coefficients), generated geometries are sane."""
import json

import numpy as np

from oracle import basis as obasis
from tests import helpers
from tests.test_ao3c_core_cpu import pack
from votca_b200 import realsys


def test_function_counts_of_the_baseline_configurations():
    for name, (n, naux) in {"methane-svp": (34, 104), "benzene-tzvp": (222, 546), "c60-tzvp": (1860, 4560),
                            "benzene26-svp": (2964, 9672)}.items():
        s = realsys.system(name)
        assert (s["nbasis"], s["naux"]) == (n, naux), name


def test_shell_arrays_equal_oracle_basis():
    g = helpers.load_golden()
    el = [str(e) for e in g["molecule_methane_tutorial/elements"]]
    pos = np.asarray(g["molecule_methane_tutorial/positions_bohr"])
    for ours, key in (("def2-svp", "def2-svp_CH"), ("aux-def2-svp", "aux-def2-svp_CH")):
        bs = json.loads(str(g[f"basis/{key}.json"]))
        bs = {e: [(int(l), [tuple(p) for p in prims]) for l, prims in shells] for e, shells in bs.items()}
        ref = pack(obasis.AOBasis(bs, el, pos))
        got = realsys.shell_arrays(ours, el, pos)
        for a, b in zip(ref, got):
            assert a.shape == b.shape and np.allclose(a, b, rtol=1e-13, atol=0)


def test_geometries():
    el, pos = realsys.c60()
    d = np.linalg.norm(pos[:, None] - pos[None], axis=2) / realsys.ANG2BOHR
    np.fill_diagonal(d, 9.0)
    assert np.allclose(np.sort(d, axis=1)[:, :3], 1.42, atol=1e-6)         # every carbon has three bonds
    assert np.allclose(np.linalg.norm(pos, axis=1), np.linalg.norm(pos[0]))   # all on one sphere
    el, pos = realsys.cluster(realsys.benzene(), 26)
    assert len(el) == 312 and len({tuple(np.round(p, 6)) for p in pos}) == 312
    el, pos = realsys.methane()
    assert np.allclose(np.linalg.norm(pos[1:] - pos[0], axis=1) / realsys.ANG2BOHR, 1.087)
