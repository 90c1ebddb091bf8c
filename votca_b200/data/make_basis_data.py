"""Extracts the H, C, N, S entries of the def2 basis sets the BASELINE.json configurations name from the reference's
basis-set library (xtp/share/xtp/basis_sets/*.xml) into votca_b200/data/basis_sets.json.  Run in the build container
(needs /root/reference); the JSON is plain published basis-set data (decays and contraction factors).
Parsing follows BasisSet::Load (xtp/src/libxtp/basisset.cc:150-199): one shell per letter of the shell type."""
import json
import os
import xml.etree.ElementTree as ET

REF = "/root/reference/xtp/share/xtp/basis_sets"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "basis_sets.json")
L_OF = {"S": 0, "P": 1, "D": 2, "F": 3, "G": 4, "H": 5, "I": 6}
SETS = ["def2-svp", "aux-def2-svp", "def2-tzvp", "aux-def2-tzvp"]
ELEMENTS = ["H", "C", "N", "S"]


def load(path):
    out = {}
    for el in ET.parse(path).getroot().findall("element"):
        if el.get("name") not in ELEMENTS:
            continue
        shells = []
        for sh in el.findall("shell"):
            for sub in sh.get("type"):
                prims = []
                for const in sh.findall("constant"):
                    contraction = 0.0
                    for c in const.findall("contractions"):
                        if c.get("type") == sub:
                            contraction = float(c.get("factor"))
                    prims.append([float(const.get("decay")), contraction])
                shells.append([L_OF[sub], prims])
        out[el.get("name")] = shells
    return out


if __name__ == "__main__":
    data = {name: load(os.path.join(REF, name + ".xml")) for name in SETS}
    with open(OUT, "w") as fh:
        json.dump(data, fh, separators=(",", ":"))
    for name, d in data.items():
        print(name, {el: sum(2 * l + 1 for l, _ in sh) for el, sh in d.items()})
