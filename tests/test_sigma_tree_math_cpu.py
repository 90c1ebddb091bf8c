"""CPU check of the mathematics behind votca_b200/csrc/sigma_tree.cu (NumPy restatement of the same steps:
sorted pole positions, 4-ary tree of contiguous ranges, 24 normalised moments per node, moment-to-moment shifts by
the triangular binomial recurrence, opening criterion |w - c| >= 4 rho, complex Horner evaluation and its
derivative) against the term-by-term sum of sigma_ppm.cc:37-91.  The GPU kernels are tested against the same sum
in tests/test_gpu_kernels.py::test_sigma_tree_*; this file guards the constants and formulas without a GPU."""
import numpy as np
import pytest

P, LEAF, OPEN, RHO_MIN = 24, 64, 4.0, 1e-300


class Tree:
    def __init__(self, a, r):
        order = np.argsort(a, kind="stable")
        self.a, self.r = a[order], r[order]
        self.T = self.a.size
        nl = -(-self.T // LEAF)
        self.D = 0
        while 4 ** self.D < nl:
            self.D += 1
        self.span = [LEAF * 4 ** (self.D - d) for d in range(self.D + 1)]
        self.count = [-(-self.T // s) for s in self.span]
        self.mom = [np.zeros((c, P)) for c in self.count]
        for j in range(self.count[self.D]):  # leaf sums
            f, l, c, rho = self.geom(self.D, j)
            x = (self.a[f:l + 1] - c) / max(rho, RHO_MIN)
            pw = self.r[f:l + 1].copy()
            for o in range(P):
                self.mom[self.D][j, o] = pw.sum()
                pw = pw * x
        for d in range(self.D - 1, -1, -1):  # moment-to-moment
            for j in range(self.count[d]):
                _, _, cp, rp = self.geom(d, j)
                inv = 1.0 / max(rp, RHO_MIN)
                acc = np.zeros(P)
                for ch in range(4 * j, min(4 * j + 4, self.count[d + 1])):
                    _, _, cc, rc = self.geom(d + 1, ch)
                    nu = self.mom[d + 1][ch] * (rc * inv) ** np.arange(P)
                    s = (cc - cp) * inv
                    for t in range(1, P):
                        for jj in range(P - 1, t - 1, -1):
                            nu[jj] += s * nu[jj - 1]
                    acc += nu
                self.mom[d][j] = acc

    def geom(self, d, j):
        f = j * self.span[d]
        l = min(f + self.span[d], self.T) - 1
        return f, l, 0.5 * (self.a[f] + self.a[l]), 0.5 * (self.a[l] - self.a[f])

    def evaluate(self, w, eta):
        d0 = min(self.D, 2)
        stack = [(d0, j) for j in range(self.count[d0])]
        s = ds = 0.0
        nfar = nleaf = 0
        while stack:
            d, j = stack.pop()
            f, l, c, rho = self.geom(d, j)
            tr = w - c
            if abs(tr) >= OPEN * rho and abs(tr) > 0.0:
                den = 1.0 / (tr * tr + eta * eta)
                u = complex(tr * den, eta * den)
                q = max(rho, RHO_MIN) * u
                mu = self.mom[d][j]
                A, B = complex(mu[P - 1]), complex(P * mu[P - 1])
                for o in range(P - 2, -1, -1):
                    A = A * q + mu[o]
                    B = B * q + (o + 1) * mu[o]
                s += (u * A).real
                ds -= (u * u * B).real
                nfar += 1
            elif d == self.D:
                t = w - self.a[f:l + 1]
                den = 1.0 / (t * t + eta * eta)
                s += (self.r[f:l + 1] * t * den).sum()
                ds += (self.r[f:l + 1] * den * (2.0 * eta * eta * den - 1.0)).sum()
                nleaf += 1
            else:
                stack += [(d + 1, ch) for ch in range(4 * j, min(4 * j + 4, self.count[d + 1]))]
        return s, ds, nfar, nleaf


def _system(seed, ntotal=120, npoles=260, nocc=35):
    rng = np.random.default_rng(seed)
    e = np.sort(np.concatenate([rng.uniform(-1.2, -0.25, nocc), 0.02 + 3 * rng.uniform(0, 1, ntotal - nocc) ** 2]))
    om = 0.3 + 8.0 * rng.uniform(0, 1, npoles) ** 2
    fac = rng.uniform(0.1, 1, npoles) * om
    M = rng.standard_normal((ntotal, npoles)) * 0.1
    a = np.where(np.arange(ntotal)[:, None] < nocc, e[:, None] - om[None, :], e[:, None] + om[None, :])
    return a.ravel(), (fac[None, :] * M * M).ravel()


@pytest.mark.parametrize("seed", [1, 2])
def test_treecode_matches_term_by_term_sum(seed):
    a, r = _system(seed)
    eta = 1e-3
    tree = Tree(a, r)
    # moment-to-moment shifts reproduce the directly accumulated moments of every inner node
    for d in range(tree.D):
        for j in range(0, tree.count[d], max(1, tree.count[d] // 5)):
            f, l, c, rho = tree.geom(d, j)
            x = (tree.a[f:l + 1] - c) / max(rho, RHO_MIN)
            ref = np.array([(tree.r[f:l + 1] * x ** o).sum() for o in range(P)])
            assert np.abs(ref - tree.mom[d][j]).max() < 1e-12 * np.abs(tree.r[f:l + 1]).sum()
    rng = np.random.default_rng(seed + 10)
    freqs = list(rng.uniform(-3.0, 6.0, 25)) + [a[17], a[4000] + 1e-5, 40.0, -30.0]
    for w in freqs:
        t = w - a
        den = 1.0 / (t * t + eta * eta)
        s_ref, d_ref = (r * t * den).sum(), (r * den * (2 * eta * eta * den - 1.0)).sum()
        scale = (np.abs(r) * np.sqrt(den)).sum()  # sum |r| / |w - a - i eta|: what the truncation bound refers to
        s, ds, nfar, nleaf = tree.evaluate(w, eta)
        assert abs(s - s_ref) < 2e-14 * scale
        assert abs(ds - d_ref) < 1e-9 * max(1.0, abs(d_ref))
        assert nfar + nleaf < 0.2 * tree.count[tree.D]  # far fewer node visits than leaves


def test_degenerate_poles_and_tiny_systems():
    # all poles identical: every node has rho = 0 and only the zeroth moment survives
    a, r = np.full(300, 0.7), np.linspace(0.1, 1.0, 300)
    tree = Tree(a, r)
    for w in (0.7, 0.7 + 1e-6, -2.0):
        s, _, _, _ = tree.evaluate(w, 1e-3)
        t = w - 0.7
        assert abs(s - r.sum() * t / (t * t + 1e-6)) < 1e-12 * max(1.0, abs(s))
    # fewer terms than one leaf
    a, r = np.array([0.1, 0.4, -0.3]), np.array([1.0, 2.0, 3.0])
    s, _, nfar, nleaf = Tree(a, r).evaluate(0.2, 1e-3)
    t = 0.2 - a
    assert abs(s - (r * t / (t * t + 1e-6)).sum()) < 1e-13 and (nfar, nleaf) == (0, 1)
