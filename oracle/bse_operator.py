"""Oracle for BSE_OPERATOR<cqp,cx,cd,cd2> (test infrastructure).

Follows xtp/src/libxtp/gwbse/bse_operator.cc:29-175 with the CPU branches of
xtp/src/libxtp/openmp_cuda.cc:299-478 (PrepareMatrix1/2, Addvec, MultiplyRow,
MultiplyBlocks) and xtp/include/votca/xtp/bseoperator_btda.h:116-149.
Index convention: I = ctotal * v + c (vc2index.h:42-44).

`matmul` is the reference formulation (rebuilds H row by row: the CPU baseline);
`dense` returns the explicit matrix for small cases.
"""
from dataclasses import dataclass

import numpy as np


@dataclass
class BSEOperatorOptions:
    homo: int = 0
    rpamin: int = 0
    qpmin: int = 0
    vmin: int = 0
    cmax: int = 0


class BSEOperator:
    def __init__(self, cqp, cx, cd, cd2, eps_inv, Mmn, Hqp):
        assert not (cd2 != 0 and cd != 0), "Hamiltonian cannot contain Hd and Hd2 at the same time"
        self.cqp, self.cx, self.cd, self.cd2 = cqp, cx, cd, cd2
        self.eps_inv, self.Mmn, self.Hqp = np.asarray(eps_inv), Mmn, np.asarray(Hqp)

    def configure(self, opt):
        self.opt = opt
        self.cmin_abs = opt.homo + 1
        self.vtot = opt.homo - opt.vmin + 1
        self.ctot = opt.cmax - self.cmin_abs + 1
        self.size = self.vtot * self.ctot

    def rows(self):
        return self.size

    def _hqp_row(self, v1, c1):
        vt, ct = self.vtot, self.ctot
        res = np.zeros((ct, vt))
        res[:, v1] += self.cqp * self.Hqp[vt:vt + ct, c1 + vt]
        res[c1, :] -= self.cqp * self.Hqp[:vt, v1]
        return res.reshape(-1, order="F")

    def matmul(self, X):
        """Reference algorithm, bse_operator.cc:40-119 (row-by-row rebuild of H)."""
        X = np.asarray(X, dtype=np.float64)
        if X.ndim == 1:
            X = X[:, None]
        vt, ct = self.vtot, self.ctot
        vmin = self.opt.vmin - self.opt.rpamin
        cmin = self.cmin_abs - self.opt.rpamin
        Y = np.zeros((self.size, X.shape[1]))
        M = self.Mmn
        if self.cd != 0 or self.cd2 != 0 or self.cqp != 0:
            for c1 in range(ct):
                if self.cd != 0:
                    T = (-self.cd * M[c1 + cmin][cmin:cmin + ct, :]) * self.eps_inv[None, :]
                elif self.cd2 != 0:
                    T = (-self.cd2 * M[c1 + cmin][vmin:vmin + vt, :]) * self.eps_inv[None, :]
                for v1 in range(vt):
                    row = np.zeros(self.size)
                    if self.cd != 0:
                        blk = T @ M[v1 + vmin][vmin:vmin + vt, :].T  # (ct, vt)
                        row += blk.reshape(-1, order="F")
                    if self.cd2 != 0:
                        blk = M[v1 + vmin][cmin:cmin + ct, :] @ T.T  # (ct, vt)
                        row += blk.reshape(-1, order="F")
                    if self.cqp != 0:
                        row += self._hqp_row(v1, c1)
                    Y[ct * v1 + c1, :] = row @ X
        if self.cx > 0:
            for v1 in range(vt):
                M1 = self.cx * M[v1 + vmin][cmin:cmin + ct, :]
                for v2 in range(v1, vt):
                    blk = M1 @ M[v2 + vmin][cmin:cmin + ct, :].T
                    Y[v1 * ct:(v1 + 1) * ct] += blk @ X[v2 * ct:(v2 + 1) * ct]
                    if v1 != v2:
                        Y[v2 * ct:(v2 + 1) * ct] += blk.T @ X[v1 * ct:(v1 + 1) * ct]
        return Y

    def matmul_factorised(self, X):
        """Same operator without forming H (the formulation the GPU path uses)."""
        X = np.asarray(X, dtype=np.float64)
        if X.ndim == 1:
            X = X[:, None]
        vt, ct = self.vtot, self.ctot
        k = X.shape[1]
        vmin = self.opt.vmin - self.opt.rpamin
        cmin = self.cmin_abs - self.opt.rpamin
        Mt = self.Mmn.M
        Xr = X.reshape(vt, ct, k)  # [v, c, k]
        Y = np.zeros((vt, ct, k))
        if self.cx != 0:
            Avc = Mt[vmin:vmin + vt, cmin:cmin + ct, :]
            W = np.einsum("vcx,vck->xk", Avc, Xr)
            Y += self.cx * np.einsum("vcx,xk->vck", Avc, W)
        if self.cd != 0:
            Pcc = Mt[cmin:cmin + ct, cmin:cmin + ct, :] * self.eps_inv[None, None, :]
            Pvv = Mt[vmin:vmin + vt, vmin:vmin + vt, :]
            U = np.einsum("adx,wdk->xwak", Pcc, Xr)  # [chi, v2, c1, k]
            Y -= self.cd * np.einsum("vwx,xwak->vak", Pvv, U)
        if self.cd2 != 0:
            Pcv = Mt[cmin:cmin + ct, vmin:vmin + vt, :] * self.eps_inv[None, None, :]  # [c1, v2, chi]
            Pvc = Mt[vmin:vmin + vt, cmin:cmin + ct, :]  # [v1, c2, chi]
            U = np.einsum("vdx,wdk->xvwk", Pvc, Xr)  # [chi, v1, v2, k]
            Y -= self.cd2 * np.einsum("awx,xvwk->vak", Pcv, U)
        if self.cqp != 0:
            Hcc = self.Hqp[vt:vt + ct, vt:vt + ct]
            Hvv = self.Hqp[:vt, :vt]
            Y += self.cqp * (np.einsum("dc,vdk->vck", Hcc, Xr) - np.einsum("wv,wck->vck", Hvv, Xr))
        return Y.reshape(self.size, k)

    def dense(self):
        return self.matmul(np.eye(self.size))

    # bse_operator.cc:134-175
    def diagonal(self):
        vt, ct = self.vtot, self.ctot
        vmin = self.opt.vmin - self.opt.rpamin
        cmin = self.cmin_abs - self.opt.rpamin
        M = self.Mmn
        res = np.zeros(self.size)
        for v in range(vt):
            for c in range(ct):
                entry = 0.0
                if self.cx != 0:
                    entry += self.cx * np.sum(M[v + vmin][cmin + c, :] ** 2)
                if self.cqp != 0:
                    entry += self.cqp * (self.Hqp[c + vt, c + vt] - self.Hqp[v, v])
                if self.cd != 0:
                    entry -= self.cd * np.sum(M[c + cmin][c + cmin, :] * self.eps_inv * M[v + vmin][v + vmin, :])
                if self.cd2 != 0:
                    entry -= self.cd2 * np.sum(M[c + cmin][v + vmin, :] * self.eps_inv * M[v + vmin][c + cmin, :])
                res[ct * v + c] = entry
        return res


def singlet_tda(e, M, H):
    return BSEOperator(1, 2, 1, 0, e, M, H)


def triplet_tda(e, M, H):
    return BSEOperator(1, 0, 1, 0, e, M, H)


def singlet_btda_b(e, M, H):
    return BSEOperator(0, 2, 0, 1, e, M, H)


def hqp_op(e, M, H):
    return BSEOperator(1, 0, 0, 0, e, M, H)


def hx_op(e, M, H):
    return BSEOperator(0, 1, 0, 0, e, M, H)


def hd_op(e, M, H):
    return BSEOperator(0, 0, 1, 0, e, M, H)


def hd2_op(e, M, H):
    return BSEOperator(0, 0, 0, 1, e, M, H)


class HamiltonianOperator:
    """[A B; -B -A], bseoperator_btda.h:59-149."""

    def __init__(self, A, B, factorised=False):
        self.A, self.B = A, B
        self.size = 2 * A.rows()
        self.factorised = factorised
        d = A.diagonal()
        self.diag = np.concatenate([d, -d])

    def rows(self):
        return self.size

    def diagonal(self):
        return self.diag

    def matmul(self, X):
        X = np.asarray(X, dtype=np.float64)
        if X.ndim == 1:
            X = X[:, None]
        half = self.size // 2
        k = X.shape[1]
        stacked = np.concatenate([X[:half], X[half:]], axis=1)  # reshape (half, 2k)
        mm = (lambda op, x: op.matmul_factorised(x)) if self.factorised else (lambda op, x: op.matmul(x))
        tA = mm(self.A, stacked)
        tB = mm(self.B, stacked)
        out = np.zeros_like(X)
        out[:half] = tA[:, :k] + tB[:, k:]
        out[half:] = -tA[:, k:] - tB[:, :k]
        return out
