"""Closed-shell Hartree-Fock with density fitting - test infrastructure for real-chemistry ("tier R") inputs.

The reference's DFT (xtp/src/libxtp/dftengine) cannot run here (no libxc), so parity cases on real molecules take
their orbitals from this small RI-RHF instead ("own-integrals"): it only has to produce a physically
sensible, reproducible set of MOs, orbital energies and an exchange matrix for the GW-BSE path to start from.
(mu nu|la si) ~ sum_P B_P[mu,nu] B_P[la,si] with B = V^-1/2 (P|mu nu).
"""
import numpy as np

from oracle import integrals


def _invsqrt(M, etol=1e-10):
    w, U = np.linalg.eigh(M)
    d = np.where(w > etol, 1.0 / np.sqrt(np.where(w > etol, w, 1.0)), 0.0)
    return (U * d) @ U.T


def rhf_ri(dft, aux, charges, positions, nelec, max_iter=100, tol=1e-10):
    """Returns dict(energies, mos, exchange_mo, total_energy, iterations).  exchange_mo = C^T (-1/2 K) C is the
    Vxc matrix of a GW calculation on top of Hartree-Fock (then Hqp = e + Sigma_x + Sigma_c - Vxc = e + Sigma_c)."""
    if nelec % 2:
        raise ValueError("closed shell only")
    nocc = nelec // 2
    S = integrals.overlap(dft)
    H = integrals.kinetic(dft) + integrals.nuclear_attraction(dft, charges, positions)
    B = np.einsum("pq,qmn->pmn", _invsqrt(integrals.coulomb2c(aux)), integrals.coulomb3c(aux, dft))
    X = _invsqrt(S)

    def fock(P):
        J = np.einsum("pmn,p->mn", B, np.einsum("pls,ls->p", B, P))
        K = np.einsum("pml,ls,pns->mn", B, P, B, optimize=True)
        return H + J - 0.5 * K, J, K

    e, Cp = np.linalg.eigh(X.T @ H @ X)
    C = X @ Cp
    P = 2.0 * C[:, :nocc] @ C[:, :nocc].T
    fs, es = [], []
    E_old = 0.0
    for it in range(1, max_iter + 1):
        F, J, K = fock(P)
        E = 0.5 * np.sum(P * (H + F))
        err = X.T @ (F @ P @ S - S @ P @ F) @ X
        fs.append(F)
        es.append(err)
        fs, es = fs[-8:], es[-8:]
        if len(fs) > 1:  # DIIS
            n = len(fs)
            Bm = -np.ones((n + 1, n + 1))
            Bm[n, n] = 0.0
            for i in range(n):
                for j in range(n):
                    Bm[i, j] = np.sum(es[i] * es[j])
            rhs = np.zeros(n + 1)
            rhs[n] = -1.0
            c = np.linalg.lstsq(Bm, rhs, rcond=None)[0][:n]
            F = sum(ci * Fi for ci, Fi in zip(c, fs))
        e, Cp = np.linalg.eigh(X.T @ F @ X)
        C = X @ Cp
        P = 2.0 * C[:, :nocc] @ C[:, :nocc].T
        if abs(E - E_old) < tol and np.abs(err).max() < 1e-7:
            break
        E_old = E
    F, J, K = fock(P)
    e, Cp = np.linalg.eigh(X.T @ F @ X)
    C = X @ Cp
    # fix the sign of every MO (largest coefficient positive) so that the result is reproducible
    for j in range(C.shape[1]):
        if C[np.argmax(np.abs(C[:, j])), j] < 0:
            C[:, j] = -C[:, j]
    enuc = sum(charges[i] * charges[j] / np.linalg.norm(positions[i] - positions[j])
               for i in range(len(charges)) for j in range(i))
    return {"energies": e, "mos": C, "exchange_mo": C.T @ (-0.5 * K) @ C, "overlap": S,
            "total_energy": float(0.5 * np.sum(P * (H + F)) + enuc), "iterations": it}
