#!/usr/bin/env python
"""Static cost model of the device AO-integral kernel (no GPU needed): counts shell triples, primitive triples and
an estimate of warp-instructions per angular-momentum class for a tier-R system, from the same pair lists the
launcher builds.  A planning aid for profiling, not a measurement.

  python scratch/ao3c_cost_model.py c60-tzvp
"""
import ctypes
import math
import os
import subprocess
import sys
from collections import defaultdict

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from votca_b200 import realsys  # noqa: E402


def nc(l):
    return (l + 1) * (l + 2) // 2


def nh(L):
    return (L + 1) * (L + 2) * (L + 3) // 6


def harness():
    src = os.path.join(ROOT, "tests", "host_harness", "ao3c_host.cc")
    out = os.path.join(ROOT, "tests", "host_harness", "build", "libao3c_host.so")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    subprocess.run(["g++", "-std=c++20", "-O2", "-fPIC", "-shared", "-pthread", "-o", out, src], check=True)
    lib = ctypes.CDLL(out)
    p, i, l = ctypes.c_void_p, ctypes.c_int, ctypes.c_long
    lib.pair_stats_host.argtypes = [i, p, p, p, p, p, l, p, p, p]
    lib.pair_stats_host.restype = l
    return lib


def triple_cost(la, lb, lc, npp, nprc):
    """(warp-instructions per triple, lanes per triple) - rough per-stage instruction estimates."""
    Lab, L = la + lb, la + lb + lc
    nca, ncb, ncc = nc(la), nc(lb), nc(lc)
    width = max(nca * ncb * ncc, nh(Lab) * ncc, nh(L))
    gl = 4 if width <= 4 else 8 if width <= 8 else 16 if width <= 16 else 32
    it = lambda n: math.ceil(n / gl)  # noqa: E731
    esz = (la + 1) * (lb + 1) * (Lab + 1)
    terms_c = ((lc // 3) // 2 + 1) ** 3 if lc else 1          # aux Hermite terms per G entry (typical component)
    terms_ab = (Lab / 3 + 1) ** 3                            # (t,u,v) terms per accumulator (typical component)
    per_pp = it(3 * esz) * 6 + 10
    per_triple_prim = (it(L + 1) * 45 + 10
                       + sum(it(nh(L - n)) * 18 + 10 for n in range(L + 1))
                       + it(nh(Lab) * ncc) * (20 + terms_c * 9) + 10
                       + it(nca * ncb * ncc) * (25 + terms_ab * 10) + 10)
    epilogue = (it(nca * ncb) * (ncc * (2 * lc + 1) * 2 + 2 * ncc) + 10
                + (2 * lc + 1) * (it((2 * la + 1) * ncb) * nca * 3 + it((2 * la + 1) * (2 * lb + 1)) * (ncb * 3 + 12) + 20))
    return (120 + npp * per_pp + npp * nprc * per_triple_prim + epilogue), gl


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "c60-tzvp"
    s = realsys.system(name)
    d, a = s["dft"], s["aux"]
    lib = harness()
    ns = len(d[0])
    cap = ns * (ns + 1) // 2
    la, lb, npp = (np.zeros(cap, dtype=np.int32) for _ in range(3))
    n = lib.pair_stats_host(ns, *[x.ctypes.data for x in d], cap, la.ctypes.data, lb.ctypes.data, npp.ctypes.data)
    la, lb, npp = la[:n], lb[:n], npp[:n]
    aux_by_l = defaultdict(list)
    for l, k in zip(a[0], a[1]):
        aux_by_l[int(l)].append(int(k))
    rows, total = [], 0.0
    for key in sorted(set(zip(la.tolist(), lb.tolist()))):
        sel = (la == key[0]) & (lb == key[1])
        for lc, prims in sorted(aux_by_l.items()):
            instr = 0.0
            for nprc in set(prims):
                cnt_c = prims.count(nprc)
                for q in np.unique(npp[sel]):
                    c, gl = triple_cost(key[0], key[1], lc, int(q), nprc)
                    instr += c * (gl / 32.0) * int((npp[sel] == q).sum()) * cnt_c
            triples = int(sel.sum()) * len(prims)
            rows.append((instr, key, lc, triples, gl))
            total += instr
    rows.sort(reverse=True)
    issue = 148 * 4 * 1.9e9
    print(f"{name}: N={s['nbasis']} Naux={s['naux']} shell pairs kept {n}/{cap}, aux shells {len(a[0])}")
    print(f"estimated warp-instructions {total:.3e} -> {total / issue * 1e3:.0f} ms at perfect issue (4 IPC x 148 SMs), "
          f"{s['naux'] * s['nbasis'] ** 2 / 1e9:.2f} G integrals incl. mirror")
    print("  share  (la lb|lc)  triples      lanes")
    for instr, key, lc, triples, gl in rows[:14]:
        print(f"  {instr / total:5.1%}  ({key[0]} {key[1]}|{lc})  {triples:10d}  {gl:3d}")


if __name__ == "__main__":
    main()
