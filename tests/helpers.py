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


def _basis_from_golden(g, basis_key, mol):
    bs = json.loads(str(g[basis_key]))
    bs = {el: [(int(l), [tuple(p) for p in prims]) for l, prims in shells] for el, shells in bs.items()}
    return obasis.AOBasis(bs, [str(e) for e in g[f"molecule_{mol}/elements"]], g[f"molecule_{mol}/positions_bohr"])


@lru_cache(maxsize=None)
def water_integrals():
    """Water, 3-21G + aux-def2-svp (s,p,d,f aux shells): the system of the reference's dftgwbse integration tests."""
    g = load_golden()
    dft = _basis_from_golden(g, "basis/water_3-21G.json", "water")
    aux = _basis_from_golden(g, "basis/aux-def2-svp_OH.json", "water")
    return {"dft": dft, "aux": aux, "S": integrals.overlap(aux), "V": integrals.coulomb2c(aux),
            "ao3c": integrals.coulomb3c(aux, dft), "dipole": integrals.dipole(dft), "S_dft": integrals.overlap(dft)}


def orb_case(tag):
    """Inputs and reference outputs of one integration-test checkpoint ('neutral' or 'neutral_tda')."""
    g = load_golden()
    c = {k.split("/", 2)[2]: g[k] for k in g if k.startswith(f"orb/{tag}/")}
    c["homo"] = int(c["attr_occupied_levels"]) - 1
    for k in ("rpamin", "rpamax", "qpmin", "qpmax", "bse_vmin", "bse_cmax"):
        c[k] = int(c["attr_" + k])
    c["useTDA"] = bool(c["attr_useTDA"])
    c["use_Hqp_offdiag"] = bool(c["attr_use_Hqp_offdiag"])
    V = c["QPdiag_eigenvectors"]
    c["Hqp"] = V @ np.diag(c["QPdiag_eigenvalues"].ravel()) @ V.T  # QPdiag is the eigendecomposition of Hqp
    return c


NUCLEAR_CHARGE = {"H": 1, "C": 6, "N": 7, "O": 8}


@lru_cache(maxsize=None)
def methane_core_hamiltonian_mos():
    """'MOs' of test_ppm.cc:38-57: eigenvectors / eigenvalues of the AO core Hamiltonian T + V_nuc (ordinary, not
    generalised, eigenproblem - the test only needs some orthonormal set with methane's symmetry)."""
    g = load_golden()
    m = methane_integrals()
    Z = [NUCLEAR_CHARGE[str(e)] for e in g["molecule/elements"]]
    H = integrals.kinetic(m["basis"]) + integrals.nuclear_attraction(m["basis"], Z, g["molecule/positions_bohr"])
    return np.linalg.eigh(H)


@lru_cache(maxsize=None)
def methane_svp_case():
    """BASELINE.json config 0 with own integrals ("tier R"): methane, def2-svp + aux-def2-svp (N = 34, Naux = 104,
    homo = 4, default ranges q = 14), RI-RHF orbitals from oracle/scf.py; Vxc = HF exchange in the MO basis."""
    from oracle import scf
    g = load_golden()
    dft = _basis_from_golden(g, "basis/def2-svp_CH.json", "methane_tutorial")
    aux = _basis_from_golden(g, "basis/aux-def2-svp_CH.json", "methane_tutorial")
    Z = [NUCLEAR_CHARGE[str(e)] for e in g["molecule_methane_tutorial/elements"]]
    hf = scf.rhf_ri(dft, aux, Z, g["molecule_methane_tutorial/positions_bohr"], sum(Z))
    homo = sum(Z) // 2 - 1
    q = min(3 * homo + 1, dft.size - 1) + 1
    return {"dft": dft, "aux": aux, "hf": hf, "homo": homo, "q": q, "S": integrals.overlap(aux),
            "V": integrals.coulomb2c(aux), "ao3c": integrals.coulomb3c(aux, dft), "dipole": integrals.dipole(dft)}
