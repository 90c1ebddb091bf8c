"""Host AO integrals for the oracle (test infrastructure): overlap, 2-centre and
3-centre Coulomb over contracted real-solid-harmonic Gaussian shells.

The reference obtains these from libint2 (third-party, pinned v2.7.1 in
CMakeModules/BuildLibint.cmake:19-25; not vendored under /root/reference):
  * AOOverlap::Fill            xtp/src/libxtp/libint2_calls.cc:163-165
  * AOCoulomb::Fill (xs_xs)    xtp/src/libxtp/libint2_calls.cc:224-271
  * ComputeAO3cBlock (xs_xx)   xtp/src/libxtp/libint2_calls.cc:544-593
This file restates the published McMurchie-Davidson scheme (Hermite expansion
coefficients E, Hermite Coulomb integrals R from the Boys function) and libint's
conventions: pure shells, functions ordered m = -l..l, real solid harmonics in
the Helgaker/Schlegel-Frisch form.  Pinned through tests/test_oracle_golden.py by the
reference's 3-21G (s,p) fixtures and, for higher angular momentum, by its contracted
S/P/D/F overlap, G-shell dipole and I-shell overlap / Coulomb / three-centre (G G | I)
reference matrices (l up to 6).
"""
import math
from functools import lru_cache

import numpy as np
from scipy.special import erf


# ----------------------------------------------------------------------------
# cartesian components and the cartesian -> pure transformation
# ----------------------------------------------------------------------------
@lru_cache(maxsize=None)
def cart_components(l):
    return [(lx, ly, l - lx - ly) for lx in range(l, -1, -1) for ly in range(l - lx, -1, -1)]


def _binom(n, k):
    if k < 0 or k > n or n < 0:
        return 0.0
    return float(math.comb(int(n), int(k)))


@lru_cache(maxsize=None)
def pure_transform(l):
    """(2l+1) x ncart matrix: rows m=-l..l (libint order), columns cart_components(l).

    Real solid harmonics in Racah normalisation, Helgaker/Jorgensen/Olsen eq.
    6.4.47-6.4.50 -- the form libint2's solidharmonics.h generates.
    """
    comps = cart_components(l)
    index = {c: i for i, c in enumerate(comps)}
    T = np.zeros((2 * l + 1, len(comps)))
    for m in range(-l, l + 1):
        am = abs(m)
        norm = (1.0 / (2.0 ** am * math.factorial(l))) * math.sqrt(
            2.0 * math.factorial(l + am) * math.factorial(l - am) / (2.0 if m == 0 else 1.0))
        vm2 = 0 if m >= 0 else 1  # 2*v_m
        for t in range((l - am) // 2 + 1):
            for u in range(t + 1):
                # v runs v_m, v_m+1, ... <= floor(|m|/2 - v_m) + v_m
                nv = int(math.floor(am / 2.0 - vm2 / 2.0))
                for iv in range(nv + 1):
                    twov = 2 * iv + vm2  # 2*v
                    c = ((-1.0) ** (t + iv) * 0.25 ** t * _binom(l, t) * _binom(l - t, am + t)
                         * _binom(t, u) * _binom(am, twov))
                    lx = 2 * t + am - 2 * u - twov
                    ly = 2 * u + twov
                    lz = l - 2 * t - am
                    T[m + l, index[(lx, ly, lz)]] += norm * c
    return T


# ----------------------------------------------------------------------------
# McMurchie-Davidson building blocks (vectorised over primitive combinations)
# ----------------------------------------------------------------------------
def hermite_E(la, lb, a, b, Xab):
    """1-D Hermite expansion coefficients E[i][j][t] (arrays over primitives).

    a, b, Xab broadcastable arrays; b may be 0 (single Gaussian, 'unit' partner).
    """
    p = a + b
    mu = a * b / p
    Xpa = -b / p * Xab
    Xpb = a / p * Xab
    zero = np.zeros(np.broadcast(p, Xab).shape)
    E = [[[zero for _ in range(la + lb + 2)] for _ in range(lb + 1)] for _ in range(la + 1)]
    E[0][0][0] = np.exp(-mu * Xab * Xab) + zero
    inv2p = 0.5 / p
    for i in range(la + 1):
        for j in range(lb + 1):
            if i == 0 and j == 0:
                continue
            for t in range(i + j + 1):
                if i > 0:
                    src = E[i - 1][j]
                    val = Xpa * src[t] + (t + 1) * src[t + 1]
                else:
                    src = E[i][j - 1]
                    val = Xpb * src[t] + (t + 1) * src[t + 1]
                if t > 0:
                    val = val + inv2p * src[t - 1]
                E[i][j][t] = val
    return E


def boys(nmax, x):
    """F_n(x) for n = 0..nmax, x array -> array (nmax+1, ...).

    x < 35: F_nmax from the all-positive series e^-x sum_k (2x)^k / ((2n+1)(2n+3)...(2n+2k+1)), lower orders by the
    stable downward recursion F_{n-1} = (2x F_n + e^-x) / (2n-1).  x >= 35: F_0 = sqrt(pi/x)/2 erf(sqrt x) and the
    upward recursion (stable there).  Accurate to a few ulp (checked against 40-digit arithmetic in
    tests/test_oracle_golden.py); scipy.special.hyp1f1, used here before, loses up to 5e-13 at large x.
    """
    x = np.asarray(x, dtype=np.float64)
    flat = x.ravel()
    out = np.empty((nmax + 1, flat.size))
    small = flat < 35.0
    if np.any(small):
        xs = flat[small]
        ex = np.exp(-xs)
        term = np.full_like(xs, 1.0 / (2 * nmax + 1))
        total = term.copy()
        for k in range(1, 400):
            term = term * (2.0 * xs) / (2 * nmax + 2 * k + 1)
            total += term
            if np.all(term <= 1e-17 * total):
                break
        f = ex * total
        out[nmax, small] = f
        for n in range(nmax, 0, -1):
            f = (2.0 * xs * f + ex) / (2 * n - 1)
            out[n - 1, small] = f
    if np.any(~small):
        xl = flat[~small]
        ex = np.exp(-xl)
        f = 0.5 * np.sqrt(np.pi / xl) * erf(np.sqrt(xl))
        out[0, ~small] = f
        for n in range(nmax):
            f = ((2 * n + 1) * f - ex) / (2.0 * xl)
            out[n + 1, ~small] = f
    return out.reshape((nmax + 1,) + x.shape)


def hermite_R(L, alpha, PC):
    """Hermite Coulomb integrals R_{tuv} for t+u+v <= L.

    alpha: (np,), PC: (np,3).  Returns dict {(t,u,v): array(np)}.
    """
    X, Y, Z = PC[:, 0], PC[:, 1], PC[:, 2]
    r2 = X * X + Y * Y + Z * Z
    F = boys(L, alpha * r2)
    # R[n][(t,u,v)]
    R = [dict() for _ in range(L + 1)]
    for n in range(L + 1):
        R[n][(0, 0, 0)] = (-2.0 * alpha) ** n * F[n]
    for tot in range(1, L + 1):
        for n in range(L - tot + 1):
            for t in range(tot + 1):
                for u in range(tot - t + 1):
                    v = tot - t - u
                    if t > 0:
                        val = X * R[n + 1][(t - 1, u, v)]
                        if t > 1:
                            val = val + (t - 1) * R[n + 1][(t - 2, u, v)]
                    elif u > 0:
                        val = Y * R[n + 1][(t, u - 1, v)]
                        if u > 1:
                            val = val + (u - 1) * R[n + 1][(t, u - 2, v)]
                    else:
                        val = Z * R[n + 1][(t, u, v - 1)]
                        if v > 1:
                            val = val + (v - 1) * R[n + 1][(t, u, v - 2)]
                    R[n][(t, u, v)] = val
    return R[0]


# ----------------------------------------------------------------------------
# overlap
# ----------------------------------------------------------------------------
def _overlap_block(sa, sb):
    a = sa.exps[:, None]
    b = sb.exps[None, :]
    cc = sa.coefs[:, None] * sb.coefs[None, :]
    AB = sa.center - sb.center
    p = a + b
    pref = cc * (math.pi / p) ** 1.5
    Ex = hermite_E(sa.l, sb.l, a, b, AB[0])
    Ey = hermite_E(sa.l, sb.l, a, b, AB[1])
    Ez = hermite_E(sa.l, sb.l, a, b, AB[2])
    ca, cb = cart_components(sa.l), cart_components(sb.l)
    blk = np.zeros((len(ca), len(cb)))
    for i, (ax, ay, az) in enumerate(ca):
        for j, (bx, by, bz) in enumerate(cb):
            blk[i, j] = np.sum(pref * Ex[ax][bx][0] * Ey[ay][by][0] * Ez[az][bz][0])
    return pure_transform(sa.l) @ blk @ pure_transform(sb.l).T


def overlap(basis):
    S = np.zeros((basis.size, basis.size))
    for i, sa in enumerate(basis.shells):
        for sb in basis.shells[:i + 1]:
            blk = _overlap_block(sa, sb)
            S[sa.start:sa.start + sa.nfunc, sb.start:sb.start + sb.nfunc] = blk
            S[sb.start:sb.start + sb.nfunc, sa.start:sa.start + sa.nfunc] = blk.T
    return S


# ----------------------------------------------------------------------------
# Hermite representation of a single shell / a shell pair
# ----------------------------------------------------------------------------
def _hermite_single(sh):
    """Single-centre Hermite expansion of every cartesian component.

    Returns (exps, coefs, list over cart comps of {(t,u,v): array(nprim)}).
    """
    a = sh.exps
    Ex = hermite_E(sh.l, 0, a, 0.0 * a, 0.0)
    comps = []
    for (lx, ly, lz) in cart_components(sh.l):
        d = {}
        for t in range(lx + 1):
            for u in range(ly + 1):
                for v in range(lz + 1):
                    d[(t, u, v)] = Ex[lx][0][t] * Ex[ly][0][u] * Ex[lz][0][v]
        comps.append(d)
    return comps


def _hermite_pair(sa, sb):
    """Hermite expansion of all cartesian products of a shell pair.

    Returns p (npair,), P (npair,3), cc (npair,), and comps[i][j] = {(t,u,v): (npair,)}.
    """
    a = np.repeat(sa.exps, len(sb.exps))
    b = np.tile(sb.exps, len(sa.exps))
    cc = np.repeat(sa.coefs, len(sb.coefs)) * np.tile(sb.coefs, len(sa.coefs))
    p = a + b
    P = (a[:, None] * sa.center[None, :] + b[:, None] * sb.center[None, :]) / p[:, None]
    AB = sa.center - sb.center
    E = [hermite_E(sa.l, sb.l, a, b, AB[k]) for k in range(3)]
    ca, cb = cart_components(sa.l), cart_components(sb.l)
    comps = []
    for (ax, ay, az) in ca:
        row = []
        for (bx, by, bz) in cb:
            d = {}
            for t in range(ax + bx + 1):
                for u in range(ay + by + 1):
                    for v in range(az + bz + 1):
                        d[(t, u, v)] = E[0][ax][bx][t] * E[1][ay][by][u] * E[2][az][bz][v]
            row.append(d)
        comps.append(row)
    return p, P, cc, comps


# ----------------------------------------------------------------------------
# 2-centre Coulomb (P|Q)
# ----------------------------------------------------------------------------
def coulomb2c(basis):
    n = basis.size
    V = np.zeros((n, n))
    herm = [_hermite_single(s) for s in basis.shells]
    for i, sa in enumerate(basis.shells):
        for j, sb in enumerate(basis.shells[:i + 1]):
            na, nb = len(sa.exps), len(sb.exps)
            a = np.repeat(sa.exps, nb)
            b = np.tile(sb.exps, na)
            cc = np.repeat(sa.coefs, nb) * np.tile(sb.coefs, na)
            alpha = a * b / (a + b)
            PC = np.tile((sa.center - sb.center)[None, :], (na * nb, 1))
            R = hermite_R(sa.l + sb.l, alpha, PC)
            pref = cc * 2.0 * math.pi ** 2.5 / (a * b * np.sqrt(a + b))
            ca, cb = cart_components(sa.l), cart_components(sb.l)
            blk = np.zeros((len(ca), len(cb)))
            for ia in range(len(ca)):
                for ib in range(len(cb)):
                    acc = 0.0
                    for (t, u, v), ea in herm[i][ia].items():
                        ea_r = np.repeat(ea, nb)
                        for (t2, u2, v2), eb in herm[j][ib].items():
                            sgn = -1.0 if (t2 + u2 + v2) % 2 else 1.0
                            acc = acc + sgn * ea_r * np.tile(eb, na) * R[(t + t2, u + u2, v + v2)]
                    blk[ia, ib] = np.sum(pref * acc)
            blk = pure_transform(sa.l) @ blk @ pure_transform(sb.l).T
            V[sa.start:sa.start + sa.nfunc, sb.start:sb.start + sb.nfunc] = blk
            V[sb.start:sb.start + sb.nfunc, sa.start:sa.start + sa.nfunc] = blk.T
    return V


# ----------------------------------------------------------------------------
# 3-centre Coulomb (P|mu nu)
# ----------------------------------------------------------------------------
def coulomb3c(auxbasis, dftbasis):
    """Returns array (Naux, N, N): ao3c[P, mu, nu] = (P|mu nu), symmetric in mu,nu.

    Same content as the vector<MatrixXd> ComputeAO3cBlock produces shell by shell
    (libint2_calls.cc:544-593), for all auxiliary shells.
    """
    N, Naux = dftbasis.size, auxbasis.size
    out = np.zeros((Naux, N, N))
    aux_herm = [_hermite_single(s) for s in auxbasis.shells]
    for ia, sa in enumerate(dftbasis.shells):
        for sb in dftbasis.shells[:ia + 1]:
            p, P, cc, pair = _hermite_pair(sa, sb)
            npair = len(p)
            ca, cb = cart_components(sa.l), cart_components(sb.l)
            Ta, Tb = pure_transform(sa.l), pure_transform(sb.l)
            for ic, sc in enumerate(auxbasis.shells):
                nc = len(sc.exps)
                g = np.tile(sc.exps, npair)
                pp = np.repeat(p, nc)
                alpha = pp * g / (pp + g)
                PC = np.repeat(P, nc, axis=0) - sc.center[None, :]
                R = hermite_R(sa.l + sb.l + sc.l, alpha, PC)
                pref = (np.repeat(cc, nc) * np.tile(sc.coefs, npair)
                        * 2.0 * math.pi ** 2.5 / (pp * g * np.sqrt(pp + g)))
                cco = cart_components(sc.l)
                blk = np.zeros((len(cco), len(ca), len(cb)))
                # fold the aux Hermite coefficients (with their sign) into R first
                for k in range(len(cco)):
                    items = [(tuv, (-1.0 if sum(tuv) % 2 else 1.0) * np.tile(e, npair))
                             for tuv, e in aux_herm[ic][k].items()]
                    cache = {}
                    for i in range(len(ca)):
                        for j in range(len(cb)):
                            acc = 0.0
                            for (t, u, v), eab in pair[i][j].items():
                                key = (t, u, v)
                                if key not in cache:
                                    s = 0.0
                                    for (t2, u2, v2), ec in items:
                                        s = s + ec * R[(t + t2, u + u2, v + v2)]
                                    cache[key] = s * pref
                                acc = acc + np.repeat(eab, nc) * cache[key]
                            blk[k, i, j] = np.sum(acc)
                blk = np.einsum("kc,cij->kij", pure_transform(sc.l), blk)
                blk = np.einsum("ai,kij->kaj", Ta, blk)
                blk = np.einsum("bj,kaj->kab", Tb, blk)
                out[sc.start:sc.start + sc.nfunc, sa.start:sa.start + sa.nfunc,
                    sb.start:sb.start + sb.nfunc] = blk
                out[sc.start:sc.start + sc.nfunc, sb.start:sb.start + sb.nfunc,
                    sa.start:sa.start + sa.nfunc] = blk.transpose(0, 2, 1)
    return out


# ----------------------------------------------------------------------------
# dipole integrals <mu| r |nu> about the origin (AODipole, libint2_calls.cc emultipole1)
# ----------------------------------------------------------------------------
def dipole(basis):
    n = basis.size
    D = np.zeros((3, n, n))
    for i, sa in enumerate(basis.shells):
        for sb in basis.shells[:i + 1]:
            a = sa.exps[:, None]
            b = sb.exps[None, :]
            cc = sa.coefs[:, None] * sb.coefs[None, :]
            p = a + b
            AB = sa.center - sb.center
            P = [(a * sa.center[k] + b * sb.center[k]) / p for k in range(3)]
            pref = cc * (math.pi / p) ** 1.5
            E = [hermite_E(sa.l, sb.l, a, b, AB[k]) for k in range(3)]
            ca, cb = cart_components(sa.l), cart_components(sb.l)
            blk = np.zeros((3, len(ca), len(cb)))
            for ia, la in enumerate(ca):
                for ib, lb in enumerate(cb):
                    e0 = [E[k][la[k]][lb[k]][0] for k in range(3)]
                    e1 = [E[k][la[k]][lb[k]][1] for k in range(3)]
                    for k in range(3):
                        f = [e0[0], e0[1], e0[2]]
                        f[k] = e1[k] + P[k] * e0[k]
                        blk[k, ia, ib] = np.sum(pref * f[0] * f[1] * f[2])
            for k in range(3):
                pb = pure_transform(sa.l) @ blk[k] @ pure_transform(sb.l).T
                D[k, sa.start:sa.start + sa.nfunc, sb.start:sb.start + sb.nfunc] = pb
                D[k, sb.start:sb.start + sb.nfunc, sa.start:sa.start + sa.nfunc] = pb.T
    return D


# ----------------------------------------------------------------------------
# kinetic energy <mu| -1/2 nabla^2 |nu>  (AOKinetic, libint2 Operator::kinetic)
# ----------------------------------------------------------------------------
def kinetic(basis):
    n = basis.size
    T = np.zeros((n, n))
    for i, sa in enumerate(basis.shells):
        for sb in basis.shells[:i + 1]:
            a = sa.exps[:, None]
            b = sb.exps[None, :]
            cc = sa.coefs[:, None] * sb.coefs[None, :]
            p = a + b
            AB = sa.center - sb.center
            pref = cc * (math.pi / p) ** 1.5
            E = [hermite_E(sa.l, sb.l + 2, a, b, AB[k]) for k in range(3)]

            def s1(k, ia, jb):  # 1-D overlap without the sqrt(pi/p) factor; zero for negative powers
                return E[k][ia][jb][0] if jb >= 0 else 0.0

            def t1(k, ia, jb):  # 1-D kinetic integral acting on the ket
                return (-2.0 * b * b * s1(k, ia, jb + 2) + b * (2 * jb + 1) * s1(k, ia, jb)
                        - 0.5 * jb * (jb - 1) * s1(k, ia, jb - 2))

            ca, cb = cart_components(sa.l), cart_components(sb.l)
            blk = np.zeros((len(ca), len(cb)))
            for ia, la in enumerate(ca):
                for ib, lb in enumerate(cb):
                    sx, sy, sz = (s1(k, la[k], lb[k]) for k in range(3))
                    val = t1(0, la[0], lb[0]) * sy * sz + sx * t1(1, la[1], lb[1]) * sz + sx * sy * t1(2, la[2], lb[2])
                    blk[ia, ib] = np.sum(pref * val)
            pb = pure_transform(sa.l) @ blk @ pure_transform(sb.l).T
            T[sa.start:sa.start + sa.nfunc, sb.start:sb.start + sb.nfunc] = pb
            T[sb.start:sb.start + sb.nfunc, sa.start:sa.start + sa.nfunc] = pb.T
    return T


# ----------------------------------------------------------------------------
# nuclear attraction sum_C -Z_C <mu| 1/|r - R_C| |nu>  (AOMultipole::FillPotential with the atoms' nuclear charges)
# ----------------------------------------------------------------------------
def nuclear_attraction(basis, charges, positions):
    n = basis.size
    V = np.zeros((n, n))
    charges = np.asarray(charges, dtype=np.float64)
    positions = np.asarray(positions, dtype=np.float64)
    for i, sa in enumerate(basis.shells):
        for sb in basis.shells[:i + 1]:
            p, P, cc, pair = _hermite_pair(sa, sb)
            L = sa.l + sb.l
            ca, cb = cart_components(sa.l), cart_components(sb.l)
            blk = np.zeros((len(ca), len(cb)))
            for Z, C in zip(charges, positions):
                R = hermite_R(L, p, P - C[None, :])
                w = -Z * cc * 2.0 * math.pi / p
                for ia in range(len(ca)):
                    for ib in range(len(cb)):
                        acc = 0.0
                        for tuv, e in pair[ia][ib].items():
                            acc = acc + e * R[tuv]
                        blk[ia, ib] += np.sum(w * acc)
            pb = pure_transform(sa.l) @ blk @ pure_transform(sb.l).T
            V[sa.start:sa.start + sa.nfunc, sb.start:sb.start + sb.nfunc] = pb
            V[sb.start:sb.start + sb.nfunc, sa.start:sa.start + sa.nfunc] = pb.T
    return V
