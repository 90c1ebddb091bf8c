"""Oracle for BSE (test infrastructure).  Follows xtp/src/libxtp/gwbse/bse.cc:43-360,
500-716, xtp/include/votca/xtp/bse_initialization.h:47-93 and
xtp/src/libxtp/orbitals.cc:643-674, 742-795 (transition dipoles, oscillator
strengths).
"""
import math
from dataclasses import dataclass

import numpy as np

from . import bse_operator as bop
from .davidson import DavidsonSolver
from .rpa import RPA


@dataclass
class BSEOptions:
    useTDA: bool = True
    homo: int = 0
    rpamin: int = 0
    rpamax: int = 0
    qpmin: int = 0
    qpmax: int = 0
    vmin: int = 0
    cmax: int = 0
    nmax: int = 5
    davidson_correction: str = "DPR"
    davidson_tolerance: str = "normal"
    davidson_update: str = "safe"
    davidson_maxiter: int = 50
    min_print_weight: float = 0.5
    use_Hqp_offdiag: bool = True
    max_dyn_iter: int = 0
    dyn_tolerance: float = 1e-5


def build_full_bse_x_ranked_initial_guess(adiag, bdiag, nroots):
    n = len(adiag)
    nguess = min(n, max(4 * nroots, 8))
    ranked = []
    for i in range(n):
        a, b = adiag[i], bdiag[i]
        disc = max(0.0, (a - b) * (a + b))
        ranked.append((math.sqrt(disc), a, i))
    ranked.sort(key=lambda r: (r[0], r[1]))
    guess = np.zeros((2 * n, nguess))
    for col in range(nguess):
        guess[ranked[col][2], col] = 1.0
    return guess


class BSE:
    def __init__(self, Mmn, factorised=False):
        self.Mmn = Mmn
        self.factorised = factorised  # use the factorised matvec (GPU formulation) instead of row rebuild

    def configure(self, opt, rpa_input_energies, Hqp_in):
        self.opt = opt
        self.vmax = opt.homo
        self.cmin = opt.homo + 1
        self.vtot = self.vmax - opt.vmin + 1
        self.ctot = opt.cmax - self.cmin + 1
        self.size = self.vtot * self.ctot
        H = self.adjust_hqp_size(Hqp_in, rpa_input_energies)
        self.Hqp = H if opt.use_Hqp_offdiag else np.diag(np.diag(H))
        self.setup_direct_interaction_operator(rpa_input_energies, 0.0)

    # bse.cc:147-186
    def adjust_hqp_size(self, Hqp, rpa_e):
        o = self.opt
        hsize = self.vtot + self.ctot
        gwsize = o.qpmax - o.qpmin + 1
        roff = o.vmin - o.rpamin
        H = np.zeros((hsize, hsize))
        if o.vmin >= o.qpmin:
            start = o.vmin - o.qpmin
            if o.cmax <= o.qpmax:
                H = np.array(Hqp[start:start + hsize, start:start + hsize])
            else:
                virtoff = gwsize - start
                H[:virtoff, :virtoff] = Hqp[start:start + virtoff, start:start + virtoff]
                extra = o.cmax - o.qpmax
                idx = np.arange(hsize - extra, hsize)
                H[idx, idx] = rpa_e[roff + virtoff:roff + virtoff + extra]
        if o.vmin < o.qpmin:
            occ_extra = o.qpmin - o.vmin
            idx = np.arange(occ_extra)
            H[idx, idx] = rpa_e[roff:roff + occ_extra]
            H[occ_extra:occ_extra + gwsize, occ_extra:occ_extra + gwsize] = Hqp
            if o.cmax > o.qpmax:
                virtoff = occ_extra + gwsize
                extra = o.cmax - o.qpmax
                idx = np.arange(hsize - extra, hsize)
                H[idx, idx] = rpa_e[roff + virtoff:roff + virtoff + extra]
        return H

    # bse.cc:188-204
    def setup_direct_interaction_operator(self, rpa_e, energy):
        o = self.opt
        rpa = RPA(self.Mmn)
        rpa.configure(o.homo, o.rpamin, o.rpamax)
        rpa.set_rpa_input_energies(rpa_e)
        ev, U = np.linalg.eigh(rpa.calculate_epsilon_r(float(energy)))
        self.Mmn.multiply_right(U)
        self.eps_inv = np.where(ev > 1e-8, 1.0 / np.where(ev > 1e-8, ev, 1.0), 0.0)

    def _configure_op(self, op):
        o = self.opt
        op.configure(bop.BSEOperatorOptions(homo=o.homo, rpamin=o.rpamin, qpmin=o.qpmin, vmin=o.vmin,
                                            cmax=o.cmax))
        if self.factorised:
            op.matmul = op.matmul_factorised
        return op

    def _davidson(self):
        o = self.opt
        ds = DavidsonSolver()
        ds.set_correction(o.davidson_correction)
        ds.set_tolerance(o.davidson_tolerance)
        ds.set_size_update(o.davidson_update)
        ds.set_iter_max(o.davidson_maxiter)
        ds.set_max_search_space(10 * o.nmax)
        return ds

    # bse.cc:266-293
    def _solve_hermitian(self, H):
        ds = self._davidson()
        ds.solve(H, self.opt.nmax)
        self.last_solver = ds
        return {"eigenvalues": ds.eigenvalues, "eigenvectors": ds.eigenvectors}

    # bse.cc:315-360
    def _solve_nonhermitian(self, A, B):
        Hop = bop.HamiltonianOperator(A, B, factorised=False)
        ds = self._davidson()
        ds.set_matrix_type("HAM")
        guess = build_full_bse_x_ranked_initial_guess(A.diagonal(), B.diagonal(), self.opt.nmax)
        ds.solve(Hop, self.opt.nmax, guess)
        self.last_solver = ds
        X = ds.eigenvectors[:A.rows()]
        Y = ds.eigenvectors[A.rows():]
        s = 1.0 / np.sqrt(np.sum(X * X, axis=0) - np.sum(Y * Y, axis=0))
        return {"eigenvalues": ds.eigenvalues, "eigenvectors": X * s[None, :], "eigenvectors2": Y * s[None, :]}

    def solve_singlets(self):
        e, M, H = self.eps_inv, self.Mmn, self.Hqp
        if self.opt.useTDA:
            return self._solve_hermitian(self._configure_op(bop.singlet_tda(e, M, H)))
        return self._solve_nonhermitian(self._configure_op(bop.singlet_tda(e, M, H)),
                                        self._configure_op(bop.singlet_btda_b(e, M, H)))

    def solve_triplets(self):
        e, M, H = self.eps_inv, self.Mmn, self.Hqp
        if self.opt.useTDA:
            return self._solve_hermitian(self._configure_op(bop.triplet_tda(e, M, H)))
        return self._solve_nonhermitian(self._configure_op(bop.triplet_tda(e, M, H)),
                                        self._configure_op(bop.hd2_op(e, M, H)))

    # bse.cc:489-551
    def _expectation(self, es, H, state=None):
        X = es["eigenvectors"] if state is None else es["eigenvectors"][:, state:state + 1]
        temp = H.matmul(X)
        direct = np.sum(X * temp, axis=0)
        cross = np.zeros(0)
        if not self.opt.useTDA:
            Y = es["eigenvectors2"] if state is None else es["eigenvectors2"][:, state:state + 1]
            direct = direct + np.sum(Y * H.matmul(Y), axis=0)
            cross = 2.0 * np.sum(Y * temp, axis=0)
        return direct, cross

    # bse.cc:553-606
    def analyze_eh_interaction(self, es, singlet):
        e, M, H = self.eps_inv, self.Mmn, self.Hqp
        out = {}
        out["qp_contrib"], _ = self._expectation(es, self._configure_op(bop.hqp_op(e, M, H)))
        out["direct_contrib"], _ = self._expectation(es, self._configure_op(bop.hd_op(e, M, H)))
        if not self.opt.useTDA:
            _, cross = self._expectation(es, self._configure_op(bop.hd2_op(e, M, H)))
            out["direct_contrib"] = out["direct_contrib"] + cross
        if singlet:
            d, cross = self._expectation(es, self._configure_op(bop.hx_op(e, M, H)))
            out["exchange_contrib"] = 2.0 * d
            if not self.opt.useTDA:
                out["exchange_contrib"] = out["exchange_contrib"] + 2.0 * cross
        else:
            out["exchange_contrib"] = np.zeros(len(out["direct_contrib"]))
        return out

    # bse.cc:608-716
    def perturbative_dynamical_screening(self, es, rpa_e):
        e, M, H = self.eps_inv, self.Mmn, self.Hqp
        self.setup_direct_interaction_operator(rpa_e, 0.0)
        static, _ = self._expectation(es, self._configure_op(bop.hd_op(self.eps_inv, M, H)))
        if not self.opt.useTDA:
            _, cross = self._expectation(es, self._configure_op(bop.hd2_op(self.eps_inv, M, H)))
            static = static + cross
        E0 = es["eigenvalues"]
        dyn = np.array(E0, dtype=np.float64)
        for i in range(len(E0)):
            for _ in range(self.opt.max_dyn_iter):
                old = dyn[i]
                self.setup_direct_interaction_operator(rpa_e, old)
                d, _ = self._expectation(es, self._configure_op(bop.hd_op(self.eps_inv, M, H)), state=i)
                if not self.opt.useTDA:
                    _, cross = self._expectation(es, self._configure_op(bop.hd2_op(self.eps_inv, M, H)),
                                                 state=i)
                    d = d + cross
                dyn[i] = E0[i] + static[i] - d[0]
                if abs(dyn[i] - old) < self.opt.dyn_tolerance:
                    break
        return dyn


# orbitals.cc:742-795
def free_transition_dipoles(dipole_ao, mos, vmin, vtot, cmin, ctot):
    occ = mos[:, vmin:vmin + vtot]
    emp = mos[:, cmin:cmin + ctot]
    return [emp.T @ dipole_ao[i] @ occ for i in range(3)]


def coupled_transition_dipoles(es, interlevel, ctot, vtot, useTDA):
    out = []
    for s in range(es["eigenvectors"].shape[1]):
        coeffs = es["eigenvectors"][:, s].copy()
        if not useTDA:
            coeffs = coeffs + es["eigenvectors2"][:, s]
        mat = coeffs.reshape((ctot, vtot), order="F")
        out.append(-math.sqrt(2.0) * np.array([np.sum(mat * interlevel[i]) for i in range(3)]))
    return np.array(out)


# orbitals.cc:643-674
def oscillator_strengths(tdip, energies):
    n = min(len(tdip), len(energies))
    return np.array([np.sum(tdip[i] ** 2) * 2.0 / 3.0 * energies[i] for i in range(n)])
