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


def oracle_basis(name, elements, positions):
    """oracle AOBasis of a basis set shipped in votca_b200/data/basis_sets.json (extracted from the reference's
    basis-set library by votca_b200/data/make_basis_data.py)."""
    from votca_b200 import realsys
    bs = {el: [(int(l), [tuple(p) for p in prims]) for l, prims in shells]
          for el, shells in realsys.basis_set(name).items()}
    return obasis.AOBasis(bs, elements, positions)


@lru_cache(maxsize=None)
def benzene_tzvp_case():
    """BASELINE.json config 2: benzene, def2-tzvp + aux-def2-tzvp (N = 222, Naux = 546, homo = 20, q = 62, RPA size
    4221).  Inputs (RI-RHF orbitals) and the oracle's evGW(exact) + full-BSE results come from
    tests/golden/benzene_tzvp_evgw_exact.npz (tests/golden/make_benzene_tzvp.py); the AO integrals are recomputed
    with the compiled host harness of the integral code and checked against the sample stored in the fixture."""
    from tests import test_ao3c_core_cpu as hh
    from votca_b200 import realsys
    path = os.path.join(os.path.dirname(GOLDEN), "benzene_tzvp_evgw_exact.npz")
    with np.load(path) as z:
        c = {k: z[k] for k in z.files}
    el, pos = realsys.benzene()
    dft, aux = oracle_basis("def2-tzvp", el, pos), oracle_basis("aux-def2-tzvp", el, pos)
    lib = hh._bind(hh._build("libao3c_host.so", ["-O2"]))
    c["ao3c"] = hh.ao3c(lib, aux, dft)
    assert np.abs(c["ao3c"][::97, ::13, ::7] - c["ao3c_sample"]).max() < 1e-12
    c["S"], c["V"], c["dipole"] = hh.overlap(lib, aux), hh.coulomb2c(lib, aux), hh.dipole(lib, dft)
    c["dft"], c["aux"], c["elements"], c["positions"] = dft, aux, el, pos
    c["homo"], c["q"] = int(c["homo"]), int(c["q"])
    return c


@lru_cache(maxsize=None)
def bsecoupling_case():
    """The system of the reference's test_bsecoupling.cc: methane monomers A and B (B = A shifted by 4 bohr along x),
    3-21G as orbital and auxiliary basis, dimer levels 0..33 for RPA / GW / BSE, homo = 9, Hqp with off-diagonals."""
    g = load_golden()
    bs = json.loads(str(g["basis/3-21G.json"]))
    bs = {el: [(int(l), [tuple(p) for p in prims]) for l, prims in shells] for el, shells in bs.items()}
    el = [str(e) for e in g["bsecoupling/elements"]]
    posA = g["bsecoupling/positions_bohr"]
    posB = posA + np.array([4.0, 0.0, 0.0])
    dimer = obasis.AOBasis(bs, el + el, np.vstack([posA, posB]))
    qp_vec, qp_val = g["bsecoupling/Hqp"], g["bsecoupling/qpdiag_eigenvalues"]
    Hqp = qp_vec @ np.diag(qp_val) @ qp_vec.T
    return {"basis": dimer, "A_mos": g["bsecoupling/A_MOs"], "AB_mos": g["bsecoupling/AB_MOs"], "Hqp": Hqp,
            "rpa_energies": np.diag(Hqp).copy(), "spsi": g["bsecoupling/spsi_ref"], "homo": 9,
            "S_dft": integrals.overlap(dimer), "S": integrals.overlap(dimer), "V": integrals.coulomb2c(dimer),
            "ao3c": integrals.coulomb3c(dimer, dimer), "known_eV": g["bsecoupling/known_answers_eV"]}


@lru_cache(maxsize=None)
def uks_case():
    """Open-shell test system for the UKS twins (no reference fixture exists for them): methane 3-21G of the reference
    unit tests, alpha orbitals = gw/mo_eigenvectors, beta orbitals = the same set rotated among neighbours by a seeded
    orthogonal matrix close to 1, beta energies shifted; doublet occupation homo_alpha = 4, homo_beta = 3."""
    g = load_golden()
    Ca = g["gw/mo_eigenvectors"]
    ea = g["inline/gw_mo_eigenvalues"].copy()
    rng = np.random.default_rng(1917)
    K = 0.08 * rng.standard_normal((17, 17))
    Q, _ = np.linalg.qr(np.eye(17) + K - K.T)
    Cb = Ca @ Q
    eb = ea + 0.03 * np.arange(17) / 17.0 + 0.01
    vxc_a = g["gw/vxc"]
    Rv = 0.01 * rng.standard_normal((17, 17))
    vxc_b = vxc_a + 0.5 * (Rv + Rv.T)
    return {"Ca": Ca, "Cb": Cb, "ea": ea, "eb": eb, "vxc_a": vxc_a, "vxc_b": vxc_b, "homo_a": 4, "homo_b": 3}
