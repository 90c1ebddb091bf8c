"""Oracle for the unrestricted (UKS) twins of the GW-BSE path (TEST INFRASTRUCTURE - never imported by votca_b200/).

Restates, on two oracle TCMatrix objects (alpha, beta):
  RPA_UKS            xtp/src/libxtp/gwbse/rpa_uks.cc:41-71 (energies), :163-201 (shifts), :203-300 (epsilon(i w), epsilon(w)),
                     :306-367 (epsilon(z))
  PPM on RPA_UKS     xtp/src/libxtp/gwbse/ppm.cc (same construction, spin-summed epsilon)
  Sigma_PPM_UKS      xtp/src/libxtp/self_energy_evaluators/sigma_ppm_uks.cc:30-142, sigma_base_uks.cc:25-69
  H2p / screening    xtp/src/libxtp/gwbse/rpa_uks.cc:369-438 (Diagonalize_H2p), :440-556 (AmB, ApB), :91-161 (modes)
  Sigma_Exact_UKS    xtp/src/libxtp/self_energy_evaluators/sigma_exact_uks.cc:37-141
  Sigma_CDA_UKS      xtp/src/libxtp/self_energy_evaluators/sigma_cda_uks.cc:44-159 (the restricted formulas on the
                     spin-summed dielectric matrix and the channel's own energies / Mmn)
  GW_UKS             xtp/src/libxtp/gwbse/gw_uks.cc:38-123 (configure), :188-295 (G0W0 / evGW loop), :314-383 (SolveQP),
                     :758-790 (Hqp)
  BSE_OPERATOR_UKS   xtp/src/libxtp/gwbse/bse_operator_uks.cc:26-264 (matmul, all blocks), :277-352 (diagonal)
  BSE_UKS (TDA)      xtp/src/libxtp/gwbse/bse_uks.cc:51-157, 218-235, 453-479

PARITY UNPINNED against reference data: the reference ships no unit test or fixture for the UKS classes.  What pins this
file instead (tests/test_oracle_uks.py): in the closed-shell limit (alpha = beta) the spin-summed dielectric matrix equals
the restricted one, whose oracle IS pinned on the reference's rpa/ fixtures, and GW_UKS reproduces the restricted GW
oracle (pinned on gw/) level by level; the UKS operator's spin blocks reduce to the pinned restricted blocks.
"""
import math
from dataclasses import dataclass

import numpy as np

from . import gw as ogw
from . import sigma as osigma
from .davidson import DavidsonSolver
from .rpa import RPA


class RPAUKS:
    def __init__(self, Mmn_alpha, Mmn_beta):
        self.spin = (RPA(Mmn_alpha), RPA(Mmn_beta))  # per-spin energies / homo live in restricted containers

    def configure(self, homo_alpha, homo_beta, rpamin, rpamax):
        self.rpamin, self.rpamax = rpamin, rpamax
        self.spin[0].configure(homo_alpha, rpamin, rpamax)
        self.spin[1].configure(homo_beta, rpamin, rpamax)

    def set_rpa_input_energies(self, e_alpha, e_beta):
        self.spin[0].set_rpa_input_energies(e_alpha)
        self.spin[1].set_rpa_input_energies(e_beta)

    def energies(self, s):
        return self.spin[s].get_rpa_input_energies()

    # rpa_uks.cc:41-71 with ShiftUncorrectedEnergies :163-201 (head / tail counted from rpamin, unlike rpa.cc:52-62)
    def update_rpa_input_energies(self, dft_alpha, dft_beta, gw_alpha, gw_beta, qpmin):
        rpatotal = self.rpamax - self.rpamin + 1
        for s, (dft, gwa) in enumerate(((dft_alpha, gw_alpha), (dft_beta, gw_beta))):
            r = self.spin[s]
            e = np.array(dft[self.rpamin:self.rpamin + rpatotal], dtype=np.float64)
            gwsize = len(gwa)
            e[qpmin - self.rpamin:qpmin - self.rpamin + gwsize] = gwa
            lumo, qpmax = r.homo + 1, qpmin + gwsize - 1

            def max_corr(lo, hi):
                if hi < lo:
                    return 0.0
                n = hi - lo + 1
                return float(np.abs(e[lo - self.rpamin:lo - self.rpamin + n] - dft[lo:lo + n]).max())

            occ, virt = max_corr(qpmin, r.homo), max_corr(lumo, qpmax)
            e[:qpmin - self.rpamin] -= occ
            ntail = self.rpamax - qpmax
            if ntail > 0:
                e[len(e) - ntail:] += virt
            r.energies = e

    def _accumulate(self, weights):
        """sum over spins and occupied levels of M^T diag(w) M; weights(dE) -> w"""
        naux = self.spin[0].Mmn.auxsize()
        res = np.zeros((naux, naux))
        for r in self.spin:
            n_occ, n_unocc = r._sizes()
            if n_occ <= 0 or n_unocc <= 0:
                continue
            e = r.energies
            nt = r.Mmn.nsize()
            for m in range(n_occ):
                Mv = r.Mmn[m][nt - n_unocc:, :]
                dE = e[len(e) - n_unocc:] - e[m]
                res += Mv.T @ (weights(dE)[:, None] * Mv)
        return res

    # rpa_uks.cc:203-300
    def calculate_epsilon_i(self, frequency):
        res = self._accumulate(lambda dE: 2.0 * dE / (dE * dE + frequency * frequency))
        res[np.diag_indices(len(res))] += 1.0
        return res

    def calculate_epsilon_r(self, frequency):
        if isinstance(frequency, complex):
            return self._epsilon_r_complex(frequency)
        eta2 = RPA.ETA * RPA.ETA

        def w(dE):
            dm, dp = dE - frequency, dE + frequency
            return dm / (dm * dm + eta2) + dp / (dp * dp + eta2)
        res = self._accumulate(w)
        res[np.diag_indices(len(res))] += 1.0
        return res

    # rpa_uks.cc:306-367
    def _epsilon_r_complex(self, frequency):
        s1, s2 = (frequency.imag + RPA.ETA) ** 2, (frequency.imag - RPA.ETA) ** 2

        def w(dE):
            dEm, dEp = frequency.real - dE, frequency.real + dE
            return dEm / (dEm * dEm + s1) - dEp / (dEp * dEp + s2)
        res = -self._accumulate(w)
        res[np.diag_indices(len(res))] += 1.0
        return res

    # rpa_uks.cc:440-474 (AmB), :475-556 (ApB: same prefactor 2 for the alpha-alpha, beta-beta and mixed blocks)
    def _ph(self, s):
        r = self.spin[s]
        n_occ, n_unocc = r._sizes()
        nt = r.Mmn.nsize()
        # rows (v, c) of the channel's particle-hole space x aux
        return np.concatenate([r.Mmn[v][nt - n_unocc:, :] for v in range(n_occ)], axis=0), n_occ, n_unocc

    def h2p_amb(self):
        out = []
        for r in self.spin:
            n_occ, n_unocc = r._sizes()
            e = r.energies
            out.append((e[n_occ:n_occ + n_unocc][None, :] - e[:n_occ][:, None]).reshape(-1))
        return np.concatenate(out)

    def h2p_apb(self):
        Pa, _, _ = self._ph(0)
        Pb, _, _ = self._ph(1)
        P = np.concatenate([Pa, Pb], axis=0)
        apb = 2.0 * (P @ P.T)
        apb[np.diag_indices(len(apb))] += self.h2p_amb()
        return apb

    # rpa_uks.cc:369-438
    def diagonalize_h2p(self):
        amb = self.h2p_amb()
        apb = self.h2p_apb()
        erpa = -0.25 * (np.trace(apb) + amb.sum())
        sq = np.sqrt(amb)
        ev, evec = np.linalg.eigh(apb * sq[:, None] * sq[None, :])
        if ev.min() <= 0.0:
            raise RuntimeError("Detected non-positive eigenvalue.")
        omega = np.sqrt(ev)
        erpa += 0.5 * omega.sum()
        XpY = (sq[:, None] * evec) / np.sqrt(omega)[None, :]
        return omega, XpY, erpa

    # rpa_uks.cc:91-161: the Coulomb-active modes sum_vc M[v][c,:] (X+Y)[vc, s] of both channels; dark (spin-like)
    # combinations, norm below 1e-10 of the largest, are dropped
    def screening_modes(self):
        omega, XpY, _ = self.diagonalize_h2p()
        Pa, _, _ = self._ph(0)
        Pb, _, _ = self._ph(1)
        modes = Pa.T @ XpY[:len(Pa)] + Pb.T @ XpY[len(Pa):]
        norms = np.linalg.norm(modes, axis=0)
        keep = norms > 1e-10 * max(1.0, norms.max())
        return omega[keep], modes[:, keep]


class SharedPPM:
    """ppm.cc:30-59 on the spin-summed dielectric matrix; shared by both spin channels (gw_uks.cc:107-122, 213-217)."""
    SCREENING_R, SCREENING_I = 0.0, 0.5

    def construct(self, rpa):
        ev, phi = np.linalg.eigh(rpa.calculate_epsilon_r(self.SCREENING_R))
        weight = 1.0 - 1.0 / ev
        eps1inv = np.linalg.inv(phi.T @ rpa.calculate_epsilon_i(self.SCREENING_I) @ phi)
        freq = np.zeros_like(ev)
        for i in range(len(ev)):
            if weight[i] < 1e-5:
                weight[i], freq[i] = 0.0, 0.5
            else:
                nom = eps1inv[i, i] - 1.0
                freq[i] = math.sqrt(abs(-nom / (nom + weight[i]) * self.SCREENING_I ** 2))
        self.phi, self.weight, self.freq = phi, weight, freq


class SigmaPPMUKS(osigma.SigmaPPM):
    """One spin channel: the restricted formulas (sigma_ppm_uks.cc:39-140 are sigma_ppm.cc's with the spin's energies,
    homo and Mmn) with the shared plasmon-pole parameters."""

    def __init__(self, Mmn, rpa_spin, ppm):
        super().__init__(Mmn, rpa_spin)
        self.ppm = ppm

    def prepare_screening(self):  # sigma_ppm_uks.cc:30-37
        self.ppm_phi, self.ppm_weight, self.ppm_freq = self.ppm.phi, self.ppm.weight, self.ppm.freq
        self.Mmn.multiply_right(self.ppm.phi)


class SigmaExactUKS(osigma.SigmaExact):
    """sigma_exact_uks.cc: residues of the channel's Mmn on the shared screening modes; the pole sums carry no
    closed-shell factor 2 (:82, :107, :140 against sigma_exact.cc:57, :76, :106)."""

    def __init__(self, Mmn, rpa_spin, rpa_uks):
        super().__init__(Mmn, rpa_spin)
        self.rpa_uks = rpa_uks

    def prepare_screening(self):  # :37-61
        self.rpa_omegas, modes = self.rpa_uks.screening_modes()
        off = self.opt.qpmin - self.opt.rpamin
        self.residues = [self.Mmn[i + off] @ modes for i in range(self.qptotal)]

    def calc_correlation_diag_element(self, level, frequency):
        return 0.5 * super().calc_correlation_diag_element(level, frequency)

    def calc_correlation_diag_element_derivative(self, level, frequency):
        return 0.5 * super().calc_correlation_diag_element_derivative(level, frequency)

    def calc_correlation_offdiag_element(self, l1, l2, f1, f2):
        return 0.5 * super().calc_correlation_offdiag_element(l1, l2, f1, f2)


class _SpinSummedRPA:
    """What Sigma_CDA_UKS sees of RPA_UKS (sigma_cda_uks.cc:44-159): the dielectric matrix of both channels, the
    energies and homo of its own."""

    def __init__(self, rpa_uks, s):
        self._uks, self._s = rpa_uks, s

    def calculate_epsilon_i(self, frequency):
        return self._uks.calculate_epsilon_i(frequency)

    def calculate_epsilon_r(self, frequency):
        return self._uks.calculate_epsilon_r(frequency)

    def get_rpa_input_energies(self):
        return self._uks.energies(self._s)

    def __getattr__(self, name):  # ETA, homo, rpamin ... of the channel
        return getattr(self._uks.spin[self._s], name)


class SigmaCDAUKS(osigma.SigmaCDA):
    def __init__(self, Mmn, rpa_uks, s):
        super().__init__(Mmn, _SpinSummedRPA(rpa_uks, s))


class _SpinGW(ogw.GW):
    """The per-spin half of GW_UKS: restricted root search (gw_uks.cc:314-747 mirrors gw.cc:323-757) on the spin's
    evaluator; the iteration loop is driven by GWUKS."""

    def __init__(self, Mmn, vxc, dft_energies, rpa_spin, ppm, rpa_uks=None, spin=0):
        super().__init__(Mmn, vxc, dft_energies)
        self.rpa = rpa_spin
        self._ppm = ppm
        self._rpa_uks, self._spin = rpa_uks, spin

    def configure(self, opt):
        self.opt = opt
        ogw.qps.normalize_grid_search_options(opt)
        self.qptotal = opt.qpmax - opt.qpmin + 1
        if opt.sigma_integration == "exact":
            self.sigma = SigmaExactUKS(self.Mmn, self.rpa, self._rpa_uks)
        elif opt.sigma_integration == "cda":
            self.sigma = SigmaCDAUKS(self.Mmn, self._rpa_uks, self._spin)
        else:
            self.sigma = SigmaPPMUKS(self.Mmn, self.rpa, self._ppm)
        self.sigma.configure(osigma.SigmaOptions(
            homo=opt.homo, qpmin=opt.qpmin, qpmax=opt.qpmax, rpamin=opt.rpamin, rpamax=opt.rpamax, eta=opt.eta,
            quadrature_scheme=opt.quadrature_scheme, order=opt.order, alpha=opt.alpha))
        self.Sigma_x = np.zeros((self.qptotal, self.qptotal))
        self.Sigma_c = np.zeros((self.qptotal, self.qptotal))
        self.gw_sc_iteration = 0
        self.sigma_evals = 0


class GWUKS:
    def __init__(self, Mmn_alpha, Mmn_beta, vxc_alpha, vxc_beta, dft_alpha, dft_beta):
        self.rpa = RPAUKS(Mmn_alpha, Mmn_beta)
        self.ppm = SharedPPM()
        self.dft = (np.asarray(dft_alpha, dtype=np.float64), np.asarray(dft_beta, dtype=np.float64))
        self.spin = (_SpinGW(Mmn_alpha, vxc_alpha, dft_alpha, self.rpa.spin[0], self.ppm, self.rpa, 0),
                     _SpinGW(Mmn_beta, vxc_beta, dft_beta, self.rpa.spin[1], self.ppm, self.rpa, 1))

    def configure(self, opt, homo_alpha, homo_beta):
        """opt: oracle GWOptions (homo is overwritten per spin)"""
        if opt.sigma_integration not in ("ppm", "exact", "cda"):
            raise RuntimeError("oracle GW_UKS: sigma_integration is ppm, exact or cda")
        import copy
        self.opt = opt
        self.rpa.configure(homo_alpha, homo_beta, opt.rpamin, opt.rpamax)
        for s, homo in enumerate((homo_alpha, homo_beta)):
            o = copy.copy(opt)
            o.homo = homo
            self.spin[s].configure(o)

    # gw_uks.cc:188-295
    def calculate_gw_perturbation(self):
        o = self.opt
        freqs = []
        shifted = []
        for s in range(2):
            g = self.spin[s]
            g.Sigma_x = (1 - o.ScaHFX) * g.sigma.calc_exchange_matrix()
            sh = self.dft[s].copy()
            sh[g.opt.homo + 1:] += o.shift
            shifted.append(sh)
            freqs.append(sh[o.qpmin:o.qpmin + g.qptotal].copy())
        self.rpa.set_rpa_input_energies(shifted[0][o.rpamin:o.rpamax + 1], shifted[1][o.rpamin:o.rpamax + 1])
        mixing = [ogw.Anderson(o.gw_mixing_order, o.gw_mixing_alpha) for _ in range(2)]
        self.iterations = 0
        for i_gw in range(o.gw_sc_max_iterations):
            self.iterations = i_gw + 1
            for g in self.spin:
                g.gw_sc_iteration = i_gw
            if i_gw % o.reset_3c == 0 and i_gw != 0:
                for g in self.spin:
                    g.Mmn.rebuild()
            if o.sigma_integration == "ppm":  # gw_uks.cc:226-233: the other evaluators build their own screening
                self.ppm.construct(self.rpa)
            for g in self.spin:
                g.sigma.prepare_screening()
            if o.gw_mixing_order > 0 and i_gw > 0:
                for s in range(2):
                    mixing[s].update_input(freqs[s])
            freqs = [self.spin[s].solve_qp(freqs[s]) for s in range(2)]
            if o.gw_sc_max_iterations > 1:
                old = [self.rpa.energies(s).copy() for s in range(2)]
                if o.gw_mixing_order > 0 and i_gw > 0:
                    for s in range(2):
                        mixing[s].update_output(freqs[s])
                        freqs[s] = mixing[s].mix_history()
                self.rpa.update_rpa_input_energies(self.dft[0], self.dft[1], freqs[0], freqs[1], o.qpmin)
                if all(np.abs(self.rpa.energies(s) - old[s]).max() <= o.gw_sc_limit for s in range(2)):
                    break
                if i_gw == o.gw_sc_max_iterations - 1:
                    break
        for s in range(2):
            g = self.spin[s]
            g.Sigma_c[np.diag_indices(g.qptotal)] = g.sigma.calc_correlation_diag(freqs[s])

    def get_gwa_results(self, s):
        return self.spin[s].get_gwa_results()

    # gw_uks.cc:758-790
    def calculate_hqp(self):
        for g in self.spin:
            g.calculate_hqp()

    def get_hqp(self, s):
        return self.spin[s].get_hqp()


@dataclass
class _Block:
    homo: int
    vmin_rpa: int
    cmin_rpa: int
    vtotal: int
    ctotal: int
    size: int
    offset: int


class BSEOperatorUKS:
    """bse_operator_uks.cc: H on the combined excitation space [alpha (v c) | beta (v c)], index ctotal * v + c."""

    def __init__(self, cqp, cx, cd, cd2, eps_inv, Mmn_alpha, Mmn_beta, Hqp_alpha, Hqp_beta):
        assert not (cd != 0 and cd2 != 0)
        self.c = (cqp, cx, cd, cd2)
        self.eps_inv = np.asarray(eps_inv)
        self.M = (Mmn_alpha, Mmn_beta)
        self.Hqp = (np.asarray(Hqp_alpha), np.asarray(Hqp_beta))

    def configure(self, homo_alpha, homo_beta, rpamin, vmin, cmax):
        blocks, off = [], 0
        for homo in (homo_alpha, homo_beta):
            vt, ct = homo - vmin + 1, cmax - homo
            blocks.append(_Block(homo, vmin - rpamin, homo + 1 - rpamin, vt, ct, vt * ct, off))
            off += vt * ct
        self.blk = tuple(blocks)
        self.size = off

    def rows(self):
        return self.size

    def _slab(self, s, m, row0, nrows):
        return self.M[s][m][row0:row0 + nrows, :]

    def matmul(self, X):
        X = np.asarray(X, dtype=np.float64)
        if X.ndim == 1:
            X = X[:, None]
        cqp, cx, cd, cd2 = self.c
        eps = self.eps_inv
        Y = np.zeros((self.size, X.shape[1]))
        for so in range(2):
            ob = self.blk[so]
            y = Y[ob.offset:ob.offset + ob.size]
            x = X[ob.offset:ob.offset + ob.size]
            H = self.Hqp[so]
            for v1 in range(ob.vtotal):
                for c1 in range(ob.ctotal):
                    out = v1 * ob.ctotal + c1
                    if cqp != 0:  # add_qp_block, :47-73
                        row = np.zeros((ob.ctotal, ob.vtotal))
                        row[:, v1] += H[ob.vtotal:ob.vtotal + ob.ctotal, c1 + ob.vtotal]
                        row[c1, :] -= H[:ob.vtotal, v1]
                        y[out] += row.reshape(-1, order="F") @ x
                    if cd != 0:  # add_direct_block, same spin, :102-134
                        left = self._slab(so, c1 + ob.cmin_rpa, ob.cmin_rpa, ob.ctotal)
                        right = self._slab(so, v1 + ob.vmin_rpa, ob.vmin_rpa, ob.vtotal)
                        y[out] += -cd * ((left * eps) @ right.T).reshape(-1, order="F") @ x
                    if cd2 != 0:  # add_direct2_block, same spin, :136-172
                        left = self._slab(so, c1 + ob.cmin_rpa, ob.vmin_rpa, ob.vtotal)
                        right = self._slab(so, v1 + ob.vmin_rpa, ob.cmin_rpa, ob.ctotal)
                        y[out] += -cd2 * ((left * eps) @ right.T).reshape(-1) @ x  # row(v2 * ct + c2) = block(v2, c2)
            if cx != 0:  # add_exchange_block, same spin only, :75-100
                A = np.vstack([self._slab(so, v + ob.vmin_rpa, ob.cmin_rpa, ob.ctotal) for v in range(ob.vtotal)])
                y += cx * (A @ (A.T @ x))
            # cross-spin blocks, :237-255
            si = 1 - so
            ib = self.blk[si]
            xin = X[ib.offset:ib.offset + ib.size]
            if cd != 0:  # add_direct_cross_tda_block: transition densities of both sides, screened, :174-211
                Tin = np.vstack([self._slab(si, v + ib.vmin_rpa, ib.cmin_rpa, ib.ctotal) for v in range(ib.vtotal)])
                W = (Tin.T @ xin) * eps[:, None]
                for v1 in range(ob.vtotal):
                    for c1 in range(ob.ctotal):
                        tout = self.M[so][c1 + ob.cmin_rpa][v1 + ob.vmin_rpa, :]
                        y[v1 * ob.ctotal + c1] += -cd * (tout @ W)
            if cd2 != 0:  # add_direct2_block(out, in = other spin): Mout = M[so], Min = M[si], :252-255
                for v1 in range(ob.vtotal):
                    for c1 in range(ob.ctotal):
                        left = self.M[so][c1 + ob.cmin_rpa][ib.vmin_rpa:ib.vmin_rpa + ib.vtotal, :]
                        right = self.M[si][v1 + ob.vmin_rpa][ib.cmin_rpa:ib.cmin_rpa + ib.ctotal, :]
                        y[v1 * ob.ctotal + c1] += -cd2 * ((left * eps) @ right.T).reshape(-1) @ xin
        return Y

    def dense(self):
        return self.matmul(np.eye(self.size))

    # bse_operator_uks.cc:277-352
    def diagonal(self):
        cqp, cx, cd, cd2 = self.c
        eps = self.eps_inv
        out = np.zeros(self.size)
        for s in range(2):
            b = self.blk[s]
            H = self.Hqp[s]
            for v in range(b.vtotal):
                Mv = self.M[s][v + b.vmin_rpa]
                for c in range(b.ctotal):
                    Mc = self.M[s][c + b.cmin_rpa]
                    e = 0.0
                    if cx != 0:
                        e += cx * np.sum(Mv[b.cmin_rpa + c] ** 2)
                    if cqp != 0:
                        e += H[c + b.vtotal, c + b.vtotal] - H[v, v]
                    if cd != 0:
                        e -= np.sum(Mc[b.cmin_rpa + c] * eps * Mv[b.vmin_rpa + v])
                    if cd2 != 0:
                        e -= np.sum(Mc[b.vmin_rpa + v] * eps * Mv[b.cmin_rpa + c])
                    out[b.offset + v * b.ctotal + c] = e
        return out


def exciton_uks_tda(eps_inv, Ma, Mb, Hqp_a, Hqp_b):  # bse_operator_uks.h: ExcitonUKSOperator_TDA = <1, 1, 1, 0>
    return BSEOperatorUKS(1, 1, 1, 0, eps_inv, Ma, Mb, Hqp_a, Hqp_b)


def exciton_uks_btda_b(eps_inv, Ma, Mb, Hqp_a, Hqp_b):  # ExcitonUKSOperator_BTDA_B = <0, 1, 0, 1>
    return BSEOperatorUKS(0, 1, 0, 1, eps_inv, Ma, Mb, Hqp_a, Hqp_b)


class BSEUKS:
    """bse_uks.cc.  Mmn_alpha / Mmn_beta are rotated in place by the screening eigenvectors."""

    def __init__(self, Mmn_alpha, Mmn_beta):
        self.M = (Mmn_alpha, Mmn_beta)

    def configure(self, opt, homo_alpha, homo_beta, rpa_e_alpha, rpa_e_beta, Hqp_alpha, Hqp_beta):
        """opt: oracle BSEOptions.  Screening as in gwbse.cc:1157-1175 (spin-summed epsilon(0), eigen-decomposition)."""
        from .bse import BSE
        self.opt, self.homo = opt, (homo_alpha, homo_beta)
        rpa = RPAUKS(*self.M)
        rpa.configure(homo_alpha, homo_beta, opt.rpamin, opt.rpamax)
        rpa.set_rpa_input_energies(rpa_e_alpha, rpa_e_beta)
        ev, U = np.linalg.eigh(rpa.calculate_epsilon_r(0.0))
        self.eps_inv = np.where(ev > 1e-8, 1.0 / np.where(ev > 1e-8, ev, 1.0), 0.0)
        self.Hqp = []
        for s, (H, e) in enumerate(((Hqp_alpha, rpa_e_alpha), (Hqp_beta, rpa_e_beta))):
            # AdjustHqpSize, bse_uks.cc:92-133: the restricted routine with the spin's homo
            helper = BSE(self.M[s])
            import copy
            o = copy.copy(opt)
            o.homo = self.homo[s]
            helper.opt = o
            helper.vtot = o.homo - o.vmin + 1
            helper.ctot = o.cmax - o.homo
            Hs = helper.adjust_hqp_size(H, e)
            self.Hqp.append(Hs if opt.use_Hqp_offdiag else np.diag(np.diag(Hs)))
            self.M[s].multiply_right(U)

    # bse_uks.cc:135-157.  The reference rotates a pristine copy of the tensors by the eigenvectors of epsilon(energy);
    # here the tensors are rotated in place again: epsilon computed from rotated tensors has the rotated eigenvectors,
    # so the product of the rotations is the same and the operator, which is all that is used, is identical.
    def setup_direct_interaction_operator(self, rpa_e_alpha, rpa_e_beta, energy):
        rpa = RPAUKS(*self.M)
        rpa.configure(self.homo[0], self.homo[1], self.opt.rpamin, self.opt.rpamax)
        rpa.set_rpa_input_energies(rpa_e_alpha, rpa_e_beta)
        ev, U = np.linalg.eigh(rpa.calculate_epsilon_r(energy))
        self.eps_inv = np.where(ev > 1e-8, 1.0 / np.where(ev > 1e-8, ev, 1.0), 0.0)
        for s in range(2):
            self.M[s].multiply_right(U)

    def _operator(self, cqp, cx, cd, cd2):
        op = BSEOperatorUKS(cqp, cx, cd, cd2, self.eps_inv, self.M[0], self.M[1], self.Hqp[0], self.Hqp[1])
        op.configure(self.homo[0], self.homo[1], self.opt.rpamin, self.opt.vmin, self.opt.cmax)
        return op

    # bse_uks.cc:171-216
    @staticmethod
    def _expectation(es, op, tda, state=None):
        X = es["eigenvectors"] if state is None else es["eigenvectors"][:, [state]]
        HX = op.matmul(X)
        direct = np.sum(X * HX, axis=0)
        cross = None
        if not tda:
            Y = es["eigenvectors2"] if state is None else es["eigenvectors2"][:, [state]]
            direct = direct + np.sum(Y * op.matmul(Y), axis=0)
            cross = 2.0 * np.sum(Y * HX, axis=0)
        return direct, cross

    # bse_uks.cc:640-702
    def perturbative_dynamical_screening(self, es, rpa_e_alpha, rpa_e_beta):
        tda = self.opt.useTDA
        self.setup_direct_interaction_operator(rpa_e_alpha, rpa_e_beta, 0.0)
        static, _ = self._expectation(es, self._operator(0, 0, 1, 0), tda)
        if not tda:
            static = static + self._expectation(es, self._operator(0, 0, 0, 1), tda)[1]
        E0 = np.asarray(es["eigenvalues"], dtype=np.float64)
        dyn = E0.copy()
        for i in range(len(E0)):
            for _ in range(self.opt.max_dyn_iter):
                old = dyn[i]
                self.setup_direct_interaction_operator(rpa_e_alpha, rpa_e_beta, old)
                d, _ = self._expectation(es, self._operator(0, 0, 1, 0), tda, state=i)
                if not tda:
                    d = d + self._expectation(es, self._operator(0, 0, 0, 1), tda, state=i)[1]
                dyn[i] = E0[i] + static[i] - d[0]
                if abs(dyn[i] - old) < self.opt.dyn_tolerance:
                    break
        return dyn

    # Orbitals::CalcCoupledTransition_Dipoles(ExcitonUKS), orbitals.cc:798-877; Oscillatorstrengths :645-674
    def transition_dipoles(self, es, dipole_ao, mos_alpha, mos_beta):
        o = self.opt
        out = []
        inter = []
        for C, homo in ((mos_alpha, self.homo[0]), (mos_beta, self.homo[1])):
            occ, virt = C[:, o.vmin:homo + 1], C[:, homo + 1:o.cmax + 1]
            inter.append([virt.T @ dipole_ao[k] @ occ for k in range(3)])       # ctotal x vtotal
        na = inter[0][0].size
        for s in range(es["eigenvectors"].shape[1]):
            c = es["eigenvectors"][:, s].copy()
            if not o.useTDA:
                c = c + es["eigenvectors2"][:, s]
            ma = c[:na].reshape(inter[0][0].shape, order="F")
            mb = c[na:].reshape(inter[1][0].shape, order="F")
            out.append(-np.array([np.sum(ma * inter[0][k]) + np.sum(mb * inter[1][k]) for k in range(3)]))
        d = np.array(out)
        f = np.sum(d * d, axis=1) * 2.0 / 3.0 * np.asarray(es["eigenvalues"])[:len(d)]
        return d, f

    def operator_tda(self):
        op = exciton_uks_tda(self.eps_inv, self.M[0], self.M[1], self.Hqp[0], self.Hqp[1])
        op.configure(self.homo[0], self.homo[1], self.opt.rpamin, self.opt.vmin, self.opt.cmax)
        return op

    def operator_btda_b(self):
        op = exciton_uks_btda_b(self.eps_inv, self.M[0], self.M[1], self.Hqp[0], self.Hqp[1])
        op.configure(self.homo[0], self.homo[1], self.opt.rpamin, self.opt.vmin, self.opt.cmax)
        return op

    def solve_excitons_uks_btda_dense(self):
        """bse_uks.cc:310-398, the dense branch (dim <= 128): [A B; -B -A], positive real roots ascending, phase with the
        largest |X_k| positive, normalisation X.X - Y.Y = 1."""
        A, B = self.operator_tda().dense(), self.operator_btda_b().dense()
        n = A.shape[0]
        w, V = np.linalg.eig(np.block([[A, B], [-B, -A]]))
        roots = sorted((w[i].real, i) for i in range(2 * n) if abs(w[i].imag) < 1e-8 and w[i].real > 0.0)
        nroots = min(self.opt.nmax, len(roots))
        ev = np.zeros(nroots)
        X, Y = np.zeros((n, nroots)), np.zeros((n, nroots))
        for r in range(nroots):
            vec = V[:, roots[r][1]].real
            x, y = vec[:n], vec[n:]
            if x[np.argmax(np.abs(x))] < 0.0:
                x, y = -x, -y
            f = 1.0 / math.sqrt(abs(x @ x - y @ y))
            ev[r], X[:, r], Y[:, r] = roots[r][0], x * f, y * f
        return {"eigenvalues": ev, "eigenvectors": X, "eigenvectors2": Y}

    def solve_excitons_uks_tda(self):  # bse_uks.cc:218-235, 453-479
        o = self.opt
        H = self.operator_tda()
        ds = DavidsonSolver()
        ds.set_correction(o.davidson_correction)
        ds.set_tolerance(o.davidson_tolerance)
        ds.set_size_update(o.davidson_update)
        ds.set_iter_max(o.davidson_maxiter)
        ds.set_max_search_space(10 * o.nmax)
        ds.solve(H, o.nmax)
        return {"eigenvalues": ds.eigenvalues, "eigenvectors": ds.eigenvectors}
