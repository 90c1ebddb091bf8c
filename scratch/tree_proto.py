"""NumPy prototype of the 1-D treecode Sigma_c evaluator (validates the math of sigma_tree.cu)."""
import numpy as np
P = 24
LEAF = 64
THETA_INV = 4.0
rng = np.random.default_rng(1)
ntotal, npoles, nocc = 150, 400, 40
e = np.sort(np.concatenate([rng.uniform(-1.2, -0.25, nocc), 0.02 + 3 * rng.uniform(0, 1, ntotal - nocc) ** 2]))
om = rng.uniform(0.3, 6, npoles)
fac = rng.uniform(0.1, 1, npoles) * om
M = rng.standard_normal((ntotal, npoles)) * 0.1
eta = 1e-3
a = np.where(np.arange(ntotal)[:, None] < nocc, e[:, None] - om[None, :], e[:, None] + om[None, :])
r = fac[None, :] * M * M
def direct(w):
    t = w - a
    return (r * t / (t * t + eta * eta)).sum(), (r * (eta * eta - t * t) / (t * t + eta * eta) ** 2).sum()
order = np.argsort(a.ravel(), kind="stable")
asort = a.ravel()[order]; rsort = r.ravel()[order]
T = asort.size
NL = -(-T // LEAF)
D = 0
while 4 ** D < NL: D += 1
span = [LEAF * 4 ** (D - d) for d in range(D + 1)]
count = [-(-T // s) for s in span]
def geom(d, j):
    f = j * span[d]; l = min(f + span[d], T) - 1
    c = 0.5 * (asort[f] + asort[l]); rho = 0.5 * (asort[l] - asort[f])
    return f, l, c, max(rho, 1e-300), rho
mom = [np.zeros((count[d], P)) for d in range(D + 1)]
for j in range(count[D]):
    f, l, c, re_, rho = geom(D, j)
    x = (asort[f:l + 1] - c) / re_
    pw = rsort[f:l + 1].copy()
    for o in range(P):
        mom[D][j, o] = pw.sum(); pw = pw * x
for d in range(D - 1, -1, -1):
    for j in range(count[d]):
        f, l, cp, rp, _ = geom(d, j)
        acc = np.zeros(P)
        for ch in range(4 * j, min(4 * j + 4, count[d + 1])):
            _, _, cc, rc, rho_c = geom(d + 1, ch)
            nu = mom[d + 1][ch].copy()
            ratio = rho_c / rp  # true radius ratio (0 if degenerate)
            sc = 1.0
            for m in range(P):
                nu[m] *= sc; sc *= ratio
            s = (cc - cp) / rp
            for t in range(1, P):
                for jj in range(P - 1, t - 1, -1):
                    nu[jj] += s * nu[jj - 1]
            acc += nu
        mom[d][j] = acc
# check M2M against direct moments at depth 0..D-1
for d in range(D):
    for j in range(count[d]):
        f, l, c, re_, rho = geom(d, j)
        x = (asort[f:l + 1] - c) / re_
        ref = np.array([(rsort[f:l + 1] * x ** o).sum() for o in range(P)])
        assert np.abs(ref - mom[d][j]).max() < 1e-12 * np.abs(rsort[f:l + 1]).sum() + 1e-300, (d, j, np.abs(ref - mom[d][j]).max())
def tree_eval(w, d0=1):
    d0 = min(d0, D)
    stack = [(d0, j) for j in range(count[d0])]
    s = ds = 0.0; nfar = nnear = 0
    while stack:
        d, j = stack.pop()
        f, l, c, re_, rho = geom(d, j)
        dist = abs(w - c)
        if dist >= THETA_INV * rho and dist > 0:
            tr = w - c; den = tr * tr + eta * eta; u = complex(tr, eta) / den; q = re_ * u
            A = complex(mom[d][j, P - 1]); B = complex(P * mom[d][j, P - 1])
            for o in range(P - 2, -1, -1):
                A = A * q + mom[d][j, o]; B = B * q + (o + 1) * mom[d][j, o]
            s += (u * A).real; ds += -(u * u * B).real; nfar += 1
        elif d == D:
            t = w - asort[f:l + 1]
            s += (rsort[f:l + 1] * t / (t * t + eta * eta)).sum()
            ds += (rsort[f:l + 1] * (eta * eta - t * t) / (t * t + eta * eta) ** 2).sum(); nnear += 1
        else:
            stack += [(d + 1, ch) for ch in range(4 * j, min(4 * j + 4, count[d + 1]))]
    return s, ds, nfar, nnear
worst = 0
for w in np.concatenate([rng.uniform(-2, 4, 40), [asort[100], asort[5000] + 1e-4]]):
    s0, d0_ = direct(w); s1, d1, nf, nn = tree_eval(w)
    sabs = (np.abs(r) / np.sqrt((w - a) ** 2 + eta ** 2)).sum()
    worst = max(worst, abs(s0 - s1) / sabs)
    print(f"w={w:8.4f} direct={s0: .12e} tree={s1: .12e} rel(abs-sum)={abs(s0-s1)/sabs:.1e} d:{abs(d0_-d1)/max(1,abs(d0_)):.1e} far={nf} nearleaves={nn}")
print("T", T, "D", D, "worst", worst)
