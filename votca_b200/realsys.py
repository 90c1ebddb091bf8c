"""Real-chemistry ("tier R") inputs for the device AO-integral producer: basis-set data, a few
geometries that can be generated without data files, and the flat shell arrays gwbse_basis_create takes.

Plumbing, not product arithmetic: the only math here is geometry; contraction normalisation goes through the
library (gwbse_basis_normalize).  Basis data: votca_b200/data/basis_sets.json (H, C, N, S of def2-svp, def2-tzvp
and their aux sets, extracted from the reference's basis-set library by data/make_basis_data.py).
"""
import ctypes
import json
import os

import numpy as np

ANG2BOHR = 1.8897259886  # tools::conv::ang2bohr (tools/include/votca/tools/constants.h:48)
_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "basis_sets.json")
_cache = {}


def basis_set(name):
    if not _cache:
        with open(_DATA) as fh:
            _cache.update(json.load(fh))
    return _cache[name]


def shell_arrays(basis_name, elements, positions_bohr):
    """(l, nprim, centers, exps, coefs) in AOBasis::Fill order (aobasis.cc:85-105): atoms in input order, shells
    in basis-set order, coefs normalised as AOShell::LibintShell + normalizeContraction do (aoshell.cc:65-89)."""
    from ._capi import capi, ptr
    bs = basis_set(basis_name)
    l, nprim, centers, exps, raw = [], [], [], [], []
    for el, pos in zip(elements, np.asarray(positions_bohr, dtype=np.float64)):
        for sl, prims in bs[el]:
            l.append(sl)
            nprim.append(len(prims))
            centers.append(pos)
            exps += [p[0] for p in prims]
            raw += [p[1] for p in prims]
    l = np.array(l, dtype=np.int32)
    nprim = np.array(nprim, dtype=np.int32)
    centers = np.ascontiguousarray(np.array(centers, dtype=np.float64))
    exps = np.array(exps, dtype=np.float64)
    raw = np.array(raw, dtype=np.float64)
    coefs = np.empty_like(exps)
    if capi().gwbse_basis_normalize(len(l), ptr(l), ptr(nprim), ptr(exps), ptr(raw), ptr(coefs)) != 0:
        raise ValueError("invalid shell in basis set " + basis_name)
    return l, nprim, centers, exps, coefs


def nfunc(l):
    return int((2 * np.asarray(l) + 1).sum())


# ------------------------------------------------------------------------------------------------ geometries (bohr)
def methane():
    d = 1.087 / np.sqrt(3.0)
    pos = np.array([[0, 0, 0], [d, d, d], [-d, -d, d], [-d, d, -d], [d, -d, -d]], dtype=np.float64)
    return ["C", "H", "H", "H", "H"], pos * ANG2BOHR


def benzene(cc=1.397, ch=1.087):
    ang = np.arange(6) * np.pi / 3.0
    ring = np.stack([np.cos(ang), np.sin(ang), np.zeros(6)], axis=1)
    return ["C"] * 6 + ["H"] * 6, np.vstack([ring * cc, ring * (cc + ch)]) * ANG2BOHR


def c60(bond=1.42):
    """Vertices of the truncated icosahedron: even permutations of (0, +-1, +-3phi), (+-1, +-(2+phi), +-2phi),
    (+-phi, +-2, +-(2phi+1)); edge length 2, scaled to the C-C bond length."""
    phi = (1.0 + np.sqrt(5.0)) / 2.0
    pts = set()
    for base in ((0.0, 1.0, 3 * phi), (1.0, 2 + phi, 2 * phi), (phi, 2.0, 2 * phi + 1)):
        for sx in (1, -1):
            for sy in (1, -1):
                for sz in (1, -1):
                    v = (base[0] * sx, base[1] * sy, base[2] * sz)
                    for perm in ((0, 1, 2), (1, 2, 0), (2, 0, 1)):
                        pts.add(tuple(round(v[i], 10) + 0.0 for i in perm))
    pos = np.array(sorted(pts)) * (bond / 2.0)
    assert pos.shape == (60, 3)
    return ["C"] * 60, pos * ANG2BOHR


def cluster(molecule, count, spacing_ang=5.0):
    """count copies of a molecule on a cubic lattice (the ~300-atom case is 26 benzenes at 5 A)."""
    el, pos = molecule
    n = int(np.ceil(count ** (1.0 / 3.0)))
    els, out = [], []
    for i in range(count):
        shift = np.array([i % n, (i // n) % n, i // (n * n)], dtype=np.float64) * spacing_ang * ANG2BOHR
        els += list(el)
        out.append(pos + shift)
    return els, np.vstack(out)


SYSTEMS = {
    # name: (geometry, orbital basis, aux basis)   sizes of the SURVEY.md section 8 table
    "methane-svp": (methane, "def2-svp", "aux-def2-svp"),
    "benzene-tzvp": (benzene, "def2-tzvp", "aux-def2-tzvp"),
    "c60-tzvp": (c60, "def2-tzvp", "aux-def2-tzvp"),
    "benzene26-svp": (lambda: cluster(benzene(), 26), "def2-svp", "aux-def2-svp"),
}


def system(name):
    """{'elements', 'positions', 'dft': shell arrays, 'aux': shell arrays, 'nbasis', 'naux'}"""
    geom, dft, aux = SYSTEMS[name]
    el, pos = geom()
    d, a = shell_arrays(dft, el, pos), shell_arrays(aux, el, pos)
    return {"elements": el, "positions": pos, "dft": d, "aux": a, "nbasis": nfunc(d[0]), "naux": nfunc(a[0])}
