"""Shared test helpers: golden fixtures + the methane/3-21G system of the reference unit tests."""
import json
import os
from functools import lru_cache

import numpy as np

from oracle import basis as obasis
from oracle import integrals, threecenter

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "votca_fixtures.npz")


@lru_cache(maxsize=None)
def load_golden():
    with np.load(GOLDEN) as z:
        return {k: z[k] for k in z.files}


def rel_frob(ref, val):
    """Eigen isApprox metric: ||a-b|| <= tol * min(||a||, ||b||)."""
    ref, val = np.asarray(ref), np.asarray(val)
    return np.linalg.norm(ref - val) / min(np.linalg.norm(ref), np.linalg.norm(val))


@lru_cache(maxsize=None)
def methane_integrals():
    g = load_golden()
    bs = json.loads(str(g["basis/3-21G.json"]))
    bs = {el: [(int(l), [tuple(p) for p in prims]) for l, prims in shells] for el, shells in bs.items()}
    ao = obasis.AOBasis(bs, [str(e) for e in g["molecule/elements"]], g["molecule/positions_bohr"])
    return {
        "basis": ao,
        "S": integrals.overlap(ao),
        "V": integrals.coulomb2c(ao),
        "ao3c": integrals.coulomb3c(ao, ao),
    }


def methane_mmn(mos, mmax=16, nmax=16):
    m = methane_integrals()
    tc = threecenter.TCMatrix(m["basis"].size, 0, mmax, 0, nmax)
    tc.fill_from_integrals(m["ao3c"], m["S"], m["V"], mos)
    return tc
