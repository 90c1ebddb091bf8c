"""Oracle for BSECoupling (TEST INFRASTRUCTURE - never imported by votca_b200/).

Restates xtp/src/libxtp/bsecoupling.cc: exciton couplings between two monomers A and B from a GW-BSE calculation on
the dimer AB.  Monomer excitons (Frenkel states) and charge-transfer products of monomer orbitals are projected on the
dimer's electron-hole basis, the dimer BSE Hamiltonian (TDA) is formed in that projection, Loewdin-orthogonalised, and
the effective coupling is read off by perturbation theory (:846-916) and by the reduction method (:918-1016).
Pinned on the known answers of xtp/src/tests/test_bsecoupling.cc (j_diag 23.662750 eV, j_pert 9.529579 eV, n_FE 2,
n_CT 18): tests/test_oracle_bsecoupling.py.
"""
from dataclasses import dataclass, field

import numpy as np

from . import bse_operator as bop
from .bse import BSE, BSEOptions

HRT2EV = 27.21138602  # votca::tools::conv::hrt2ev (tools/include/votca/tools/constants.h)


@dataclass
class Fragment:
    """What CalculateCouplings reads from a monomer's Orbitals (bsecoupling.cc:368-400, 516-520)."""
    mos: np.ndarray  # basis x levels
    bse_vmin: int
    bse_vmax: int
    bse_cmin: int
    bse_cmax: int
    singlets: np.ndarray = None  # (vtotal * ctotal) x states, index ctotal * v + c
    triplets: np.ndarray = None
    singlet_energies: np.ndarray = None
    triplet_energies: np.ndarray = None

    @property
    def vtotal(self):
        return self.bse_vmax - self.bse_vmin + 1

    @property
    def ctotal(self):
        return self.bse_cmax - self.bse_cmin + 1


@dataclass
class CouplingOptions:
    """BSECoupling::Initialize, bsecoupling.cc:38-69."""
    spin: str = "singlet"
    use_perturbation: bool = True
    output_tb: bool = False
    levA: int = 1
    levB: int = 1
    occA: int = 3
    unoccA: int = 3
    occB: int = 3
    unoccB: int = 3


@dataclass
class SpinChannel:
    JAB: list = field(default_factory=list)  # [perturbation, reduction], Hartree, (levA + levB)^2
    J_dimer: np.ndarray = None
    S_dimer: np.ndarray = None
    xi: float = 0.0
    pt_rm_discrepancy: float = 0.0
    downfolding_safe: bool = False


def _inv_sqrt(S):
    w, U = np.linalg.eigh(S)
    return (U / np.sqrt(w)) @ U.T


class BSECoupling:
    def __init__(self, opt):
        if opt.spin not in ("singlet", "triplet", "all"):
            raise RuntimeError(f"Choice {opt.spin} for type not known. Available singlet,triplet,all")
        self.opt = opt
        self.do_singlets = opt.spin in ("singlet", "all")
        self.do_triplets = opt.spin in ("triplet", "all")
        self.levA, self.levB = opt.levA, opt.levB
        self.occA, self.unoccA, self.occB, self.unoccB = opt.occA, opt.unoccA, opt.occB, opt.unoccB
        self.channels = {}

    # bsecoupling.cc:321-345
    @staticmethod
    def project_frenkel_excitons(coeffs, X_AB, x_vtotal, x_ctotal, ab_vtotal, ab_ctotal):
        X_occ_occ = X_AB[:, :x_vtotal][:ab_vtotal, :]
        X_unocc_unocc = X_AB[:, X_AB.shape[1] - x_ctotal:][X_AB.shape[0] - ab_ctotal:, :]
        out = np.zeros((ab_vtotal * ab_ctotal, coeffs.shape[1]))
        for i in range(coeffs.shape[1]):
            cm = coeffs[:, i].reshape((x_ctotal, x_vtotal), order="F")
            out[:, i] = (X_unocc_unocc @ cm @ X_occ_occ.T).reshape(-1, order="F")
        return out

    # bsecoupling.cc:268-319
    def setup_ct_states(self, a_vtotal, b_vtotal, ab_vtotal, ab_ctotal, A_AB, B_AB):
        noAB = self.occA * self.unoccB
        noBA = self.unoccA * self.occB
        ct = np.zeros((ab_vtotal * ab_ctotal, noAB + noBA))
        A_occ = A_AB[:, a_vtotal - self.occA:a_vtotal]
        A_unocc = A_AB[:, a_vtotal:a_vtotal + self.unoccA]
        B_occ = B_AB[:, b_vtotal - self.occB:b_vtotal]
        B_unocc = B_AB[:, b_vtotal:b_vtotal + self.unoccB]
        A_occ_occ, B_unocc_unocc = A_occ[:ab_vtotal], B_unocc[B_unocc.shape[0] - ab_ctotal:]
        for a in range(self.occA):
            for b in range(self.unoccB):
                ct[:, a * self.unoccB + b] = np.outer(B_unocc_unocc[:, b], A_occ_occ[:, a]).reshape(-1, order="F")
        A_unocc_unocc, B_occ_occ = A_unocc[A_unocc.shape[0] - ab_ctotal:], B_occ[:ab_vtotal]
        for b in range(self.occB):
            for a in range(self.unoccA):
                ct[:, b * self.unoccA + a + noAB] = np.outer(A_unocc_unocc[:, a], B_occ_occ[:, b]).reshape(-1, order="F")
        return ct

    # bsecoupling.cc:846-916
    def perturbation(self, J_dimer):
        nfe = self.levA + self.levB
        ct = J_dimer.shape[0] - nfe
        Jr = J_dimer
        if ct > 0:
            T = np.eye(J_dimer.shape[0])
            _, T[nfe:, nfe:] = np.linalg.eigh(J_dimer[nfe:, nfe:])
            Jr = T.T @ J_dimer @ T
        out = np.zeros((nfe, nfe))
        for a in range(self.levA):
            Ea = Jr[a, a]
            for b in range(self.levB):
                bd = b + self.levA
                J, Eb = Jr[a, bd], Jr[bd, bd]
                for k in range(nfe, nfe + ct):
                    Eab = Jr[k, k]
                    J += 0.5 * Jr[k, a] * Jr[k, bd] * (1.0 / (Ea - Eab) + 1.0 / (Eb - Eab))
                out[a, bd] = out[bd, a] = J
        return out

    # bsecoupling.cc:918-1016
    def fulldiag(self, J_dimer):
        nfe = self.levA + self.levB
        w, U = np.linalg.eigh(J_dimer)
        out = np.zeros((nfe, nfe))
        for a in range(self.levA):
            for b in range(self.levB):
                bd = b + self.levA
                i0 = int(np.argmax(np.abs(U[a])))
                i1 = int(np.argmax(np.abs(U[bd])))
                if i0 == i1:
                    amp = np.abs(U[bd]).copy()
                    amp[i1] = 0.0
                    i1 = int(np.argmax(amp))
                T = np.zeros((2, 2))
                E = np.zeros((2, 2))
                for i, (k, row) in enumerate(((i0, a), (i1, bd))):
                    sign = np.sign(U[row, k])
                    T[0, i] = sign * U[a, k]
                    T[1, i] = sign * U[bd, k]
                    E[i, i] = w[k]
                T /= np.linalg.norm(T, axis=0)
                if np.linalg.det(T) < 0:
                    T[:, 1] *= -1
                sm1 = _inv_sqrt(T @ T.T)
                E = sm1 @ E @ sm1
                T = T @ sm1
                Js = T @ E @ T.T
                out[a, bd] = Js[0, 1]
                out[bd, a] = Js[1, 0]
        return out

    # bsecoupling.cc:748-787
    def diagnostics(self, ch):
        nfe = self.levA + self.levB
        J = ch.J_dimer
        ct = J.shape[0] - nfe
        for i in range(nfe):
            for k in range(nfe, nfe + ct):
                dE = abs(J[i, i] - J[k, k])
                ch.xi = max(ch.xi, abs(J[i, k]) / dE) if dE > 1e-10 else np.inf
        for i in range(self.levA):
            for j in range(self.levB):
                ch.pt_rm_discrepancy = max(ch.pt_rm_discrepancy, abs(ch.JAB[0][i, j + self.levA] - ch.JAB[1][i, j + self.levA]))
        ch.downfolding_safe = bool(np.isfinite(ch.xi) and ch.xi < 0.3 and ch.pt_rm_discrepancy < 1e-4)

    # bsecoupling.cc:789-835 with OrthogonalizeCTs (:614-669: a plain merge) and CalcJ_dimer (:683-735)
    def project_excitons(self, FE_AB, CT, H):
        P = np.hstack([FE_AB, CT])
        ch = SpinChannel()
        ch.J_dimer = P.T @ H.matmul(P)
        ch.S_dimer = P.T @ P
        Sm1 = _inv_sqrt(ch.S_dimer)
        J_ortho = Sm1 @ ch.J_dimer @ Sm1
        ch.JAB = [self.perturbation(J_ortho), self.fulldiag(J_ortho)]
        self.diagnostics(ch)
        return ch

    def calculate_couplings(self, A, B, mos_AB, overlap_AB, Mmn, Hqp, rpa_input_energies, homo, rpamin, rpamax, qpmin,
                            qpmax, bse_vmin, bse_cmax, use_Hqp_offdiag=True):
        """bsecoupling.cc:356-612.  Mmn: oracle TCMatrix of the dimer, filled (rpamin..qpmax x rpamin..rpamax)."""
        basisA, basisB = A.mos.shape[0], B.mos.shape[0]
        if basisA == 0 or basisB == 0:
            raise RuntimeError("Basis set size is not stored in monomers")
        for spin_on, attr in ((self.do_singlets, "singlets"), (self.do_triplets, "triplets")):
            if spin_on:
                self.levA = min(self.levA, getattr(A, attr).shape[1])
                self.levB = min(self.levB, getattr(B, attr).shape[1])
        if self.unoccA > A.ctotal or self.unoccA < 0:
            self.unoccA = A.ctotal
        if self.unoccB > B.ctotal or self.unoccB < 0:
            self.unoccB = B.ctotal
        if self.occA > A.vtotal or self.occA < 0:
            self.occA = A.vtotal
        if self.occB > B.vtotal or self.occB < 0:
            self.occB = B.vtotal
        ab_vtotal = homo - bse_vmin + 1
        ab_ctotal = bse_cmax - homo
        ab_total = ab_vtotal + ab_ctotal
        MOsA = A.mos[:, A.bse_vmin:A.bse_vmin + A.vtotal + A.ctotal]
        MOsB = B.mos[:, B.bse_vmin:B.bse_vmin + B.vtotal + B.ctotal]
        MOsAB = mos_AB[:, bse_vmin:bse_vmin + ab_total]
        overlap = overlap_AB @ MOsAB
        A_AB = overlap[:basisA].T @ MOsA
        B_AB = overlap[overlap.shape[0] - basisB:].T @ MOsB
        opt = BSEOptions(useTDA=True, homo=homo, rpamin=rpamin, rpamax=rpamax, qpmin=qpmin, qpmax=qpmax, vmin=bse_vmin,
                         cmax=bse_cmax, use_Hqp_offdiag=use_Hqp_offdiag)
        bse = BSE(Mmn)
        bse.configure(opt, rpa_input_energies, Hqp)
        for spin_on, name, attr, make in ((self.do_singlets, "singlet", "singlets", bop.singlet_tda),
                                          (self.do_triplets, "triplet", "triplets", bop.triplet_tda)):
            if not spin_on:
                continue
            FE = np.hstack([
                self.project_frenkel_excitons(getattr(A, attr)[:, :self.levA], A_AB, A.vtotal, A.ctotal, ab_vtotal,
                                              ab_ctotal),
                self.project_frenkel_excitons(getattr(B, attr)[:, :self.levB], B_AB, B.vtotal, B.ctotal, ab_vtotal,
                                              ab_ctotal)])
            CT = self.setup_ct_states(A.vtotal, B.vtotal, ab_vtotal, ab_ctotal, A_AB, B_AB)
            H = bse._configure_op(make(bse.eps_inv, bse.Mmn, bse.Hqp))
            self.channels[name] = self.project_excitons(FE, CT, H)
        return self.channels

    # getSingletCouplingElement / getTripletCouplingElement, bsecoupling.cc:256-266: eV
    def coupling_element(self, spin, levelA, levelB, method):
        return self.channels[spin].JAB[method][levelA, levelB + self.levA] * HRT2EV
