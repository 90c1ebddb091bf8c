"""Oracle for TCMatrix_gwbse (test infrastructure).

Follows xtp/src/libxtp/threecenter.cc:29-131 and the MO transformation of
xtp/src/libxtp/libint2_calls.cc:595-651.  Storage mirrors the reference:
`M[m]` is an (ntotal x Naux) matrix, m = m_abs - mmin, row = n_abs - nmin.
Held as one array M[m, n, chi].
"""
import numpy as np

from . import integrals


def pseudo_invsqrt_gwbse(S, V, etol=5e-7):
    """AOCoulomb::Pseudo_InvSqrt_GWBSE, xtp/src/libxtp/aomatrices/aomatrix.cc:53-86.

    Returns (((S^-1/2 V S^-1/2)^-1/2 S^-1/2)^T, removedfunctions).
    """
    removed = 0
    ev, evec = np.linalg.eigh(S)
    d = np.zeros_like(ev)
    for i, e in enumerate(ev):
        if e < etol:
            removed += 1
        else:
            d[i] = 1.0 / np.sqrt(e)
    Ssqrt = (evec * d) @ evec.T
    ortho = Ssqrt @ V @ Ssqrt
    ev2, evec2 = np.linalg.eigh(ortho)
    d2 = np.zeros_like(ev2)
    for i, e in enumerate(ev2):
        if e < etol:
            removed += 1
        else:
            d2[i] = 1.0 / np.sqrt(e)
    Vm1 = (evec2 * d2) @ evec2.T
    return (Vm1 @ Ssqrt).T, removed


class TCMatrix:
    """Mmn tensor of the GW-BSE path."""

    def __init__(self, naux, mmin, mmax, nmin, nmax):
        # TCMatrix_gwbse::Initialize, threecenter.cc:29-48
        self.naux = naux
        self.mmin, self.mmax = mmin, mmax
        self.nmin, self.nmax = nmin, nmax
        self.mtotal = mmax - mmin + 1
        self.ntotal = nmax - nmin + 1
        self.M = np.zeros((self.mtotal, self.ntotal, naux))
        self.removed = 0
        self._rebuild = None

    def __getitem__(self, i):
        return self.M[i]

    def msize(self):
        return self.mtotal

    def nsize(self):
        return self.ntotal

    def auxsize(self):
        return self.naux

    # -- Fill3cMO, libint2_calls.cc:595-651 (reference order: Cn^T * ao3c[k] * Cm)
    def fill_3c_mo(self, ao3c, mos):
        Cm = mos[:, self.mmin:self.mmin + self.mtotal]
        Cn = mos[:, self.nmin:self.nmin + self.ntotal]
        for k in range(ao3c.shape[0]):
            t = Cn.T @ ao3c[k] @ Cm  # (ntotal, mtotal)
            self.M[:, :, k] = t.T

    # -- MultiplyRightWithAuxMatrix, threecenter.cc:54-65
    def multiply_right(self, R):
        for m in range(self.mtotal):
            self.M[m] = self.M[m] @ R

    # -- Fill from AO integrals, threecenter.cc:72-90
    def fill_from_integrals(self, ao3c, S_aux, V_aux, mos):
        self._rebuild = (ao3c, S_aux, V_aux, mos)
        self.fill_3c_mo(ao3c, mos)
        L, self.removed = pseudo_invsqrt_gwbse(S_aux, V_aux, 5e-7)
        self.multiply_right(L)
        return L

    def fill(self, auxbasis, dftbasis, mos):
        ao3c = integrals.coulomb3c(auxbasis, dftbasis)
        S = integrals.overlap(auxbasis)
        V = integrals.coulomb2c(auxbasis)
        return self.fill_from_integrals(ao3c, S, V, mos)

    def rebuild(self):
        ao3c, S, V, mos = self._rebuild
        self.fill_from_integrals(ao3c, S, V, mos)

    # -- Rotate (QSGW), threecenter.cc:108-131
    def rotate(self, U, qpmin, qpmax):
        qptotal = qpmax - qpmin + 1
        on = qpmin - self.nmin
        om = qpmin - self.mmin
        for m in range(qptotal):
            self.M[m + om, on:on + qptotal, :] = U.T @ self.M[m + om, on:on + qptotal, :]
