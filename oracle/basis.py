"""AO basis construction for the oracle (test infrastructure).

Restates, for the hot path's inputs only:
  * csg::XYZReader unit handling (csg/include/votca/csg/xyzreader.h:86:
    Angstrom -> bohr with tools::conv::ang2bohr = 1.8897259886,
    tools/include/votca/tools/constants.h:48),
  * BasisSet::Load (xtp/src/libxtp/basisset.cc:150-199): one Shell per letter of
    the shell "type" attribute ("SP" -> S then P), contraction picked by type,
  * AOBasis::Fill (xtp/src/libxtp/aobasis.cc:85-105): atoms in file order,
    shells in element order, functions appended shell by shell,
  * AOShell::LibintShell + normalizeContraction (aoshell.cc:65-89): pure (real
    solid harmonic) shells, libint primitive normalisation embedded, then the
    contraction divided by sqrt(self overlap of the shell's first function).
"""
import math
import xml.etree.ElementTree as ET
from dataclasses import dataclass

import numpy as np

ANG2BOHR = 1.8897259886
L_OF = {"S": 0, "P": 1, "D": 2, "F": 3, "G": 4, "H": 5, "I": 6}


def read_xyz(path):
    """Returns (elements, positions in bohr)."""
    with open(path) as fh:
        lines = fh.read().splitlines()
    nat = int(lines[0].split()[0])
    elems, pos = [], []
    for ln in lines[2:2 + nat]:
        tok = ln.split()
        elems.append(tok[0])
        pos.append([float(tok[1]), float(tok[2]), float(tok[3])])
    return elems, np.array(pos, dtype=np.float64) * ANG2BOHR


def load_basisset(path):
    """{element: [(l, [(decay, contraction), ...]), ...]} in file order."""
    root = ET.parse(path).getroot()
    out = {}
    for el in root.findall("element"):
        shells = []
        for sh in el.findall("shell"):
            for sub in sh.get("type"):
                prims = []
                for const in sh.findall("constant"):
                    decay = float(const.get("decay"))
                    contraction = 0.0
                    for c in const.findall("contractions"):
                        if c.get("type") == sub:
                            contraction = float(c.get("factor"))
                    prims.append((decay, contraction))
                shells.append((L_OF[sub], prims))
        out[el.get("name")] = shells
    return out


def dfact(n):
    """double factorial with (-1)!! = 1."""
    return 1.0 if n <= 0 else float(np.prod(np.arange(n, 0, -2, dtype=np.float64)))


@dataclass
class Shell:
    l: int
    center: np.ndarray
    exps: np.ndarray
    coefs: np.ndarray  # include primitive normalisation and the VOTCA shell norm
    start: int
    atom: int

    @property
    def nfunc(self):
        return 2 * self.l + 1


def _prim_norm(alpha, l):
    # libint2::Shell::renorm: normalisation of x^l exp(-alpha r^2)
    return math.sqrt(2.0 ** l * (2.0 * alpha) ** (l + 1.5) / (math.pi ** 1.5 * dfact(2 * l - 1)))


def _self_overlap_xl(exps, coefs, l):
    # <sum_p c_p x^l e^{-a_p r^2} | sum_q c_q x^l e^{-a_q r^2}>
    s = 0.0
    for a, ca in zip(exps, coefs):
        for b, cb in zip(exps, coefs):
            p = a + b
            s += ca * cb * dfact(2 * l - 1) / (2.0 * p) ** l * (math.pi / p) ** 1.5
    return s


class AOBasis:
    def __init__(self, basisset, elements, positions):
        self.shells = []
        n = 0
        for iat, (el, pos) in enumerate(zip(elements, positions)):
            for l, prims in basisset[el]:
                exps = np.array([p[0] for p in prims])
                raw = np.array([p[1] for p in prims])
                coefs = raw * np.array([_prim_norm(a, l) for a in exps])
                # Racah-normalised solid harmonics share the norm of x^l, so the
                # "first function" self overlap of aoshell.cc:81-89 equals this.
                coefs = coefs / math.sqrt(_self_overlap_xl(exps, coefs, l))
                self.shells.append(Shell(l, np.array(pos, dtype=np.float64), exps, coefs, n, iat))
                n += 2 * l + 1
        self.size = n
        self.maxl = max(s.l for s in self.shells)

    @classmethod
    def from_files(cls, basis_xml, xyz):
        elems, pos = read_xyz(xyz)
        return cls(load_basisset(basis_xml), elems, pos)
