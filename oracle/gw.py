"""Oracle for GW (test infrastructure).  Follows xtp/src/libxtp/gwbse/gw.cc:35-78,
210-776, xtp/include/votca/xtp/gw.h:214-300 (QPFunc) and
xtp/src/libxtp/anderson_mixing.cc:28-95; QSGW: gw.cc:798-1130 (calculate_qsgw).
"""
import math
from dataclasses import dataclass

import numpy as np

from . import qp_solver as qps
from . import sigma as osigma
from .rpa import RPA


@dataclass
class GWOptions:
    homo: int = 0
    qpmin: int = 0
    qpmax: int = 0
    rpamin: int = 0
    rpamax: int = 0
    eta: float = 1e-3
    g_sc_limit: float = 1e-5
    g_sc_max_iterations: int = 100
    gw_sc_limit: float = 1e-5
    gw_sc_max_iterations: int = 50
    shift: float = 0.0
    ScaHFX: float = 0.0
    sigma_integration: str = "ppm"
    reset_3c: int = 5
    qsgw_max_iterations: int = 20
    qsgw_sc_limit: float = 1e-5
    qsgw_max_virt_correction: float = 0.5
    qp_solver: str = "grid"
    qp_solver_alpha: float = 0.75
    qp_grid_steps: int = 0
    qp_grid_spacing: float = 0.0
    qp_full_window_half_width: float = -1.0
    qp_dense_spacing: float = -1.0
    qp_adaptive_shell_width: float = -1.0
    qp_adaptive_shell_count: int = 0
    gw_mixing_order: int = 20
    gw_mixing_alpha: float = 0.7
    quadrature_scheme: str = "legendre"
    order: int = 12
    alpha: float = 1e-3
    qp_restrict_search: bool = True
    qp_zero_margin: float = 1e-6
    qp_virtual_min_energy: float = -0.1
    qp_root_finder: str = "bisection"
    qp_grid_search_mode: str = "adaptive_with_dense_fallback"


class Anderson:
    def __init__(self, order, alpha):
        self.order = order + 1
        self.alpha = alpha
        self.input, self.output = [], []

    def update_output(self, v):
        if len(self.output) > self.order - 1:
            self.output.pop(0)
        self.output.append(np.array(v))

    def update_input(self, v):
        if len(self.output) > self.order - 1:
            self.input.pop(0)
        self.input.append(np.array(v))

    def mix_history(self):
        iteration = len(self.output)
        used = iteration - 1
        out_mixed = self.output[-1].copy()
        in_mixed = self.input[-1].copy()
        if iteration > 1 and self.order > 1:
            dn = out_mixed - in_mixed
            A = np.zeros((used, used))
            c = np.zeros(used)
            for m in range(1, iteration):
                dm = dn - self.output[used - m] + self.input[used - m]
                c[m - 1] = dm @ dn
                for j in range(1, iteration):
                    A[m - 1, j - 1] = dm @ (dn - self.output[used - j] + self.input[used - j])
            # fullPivHouseholderQr().solve: rank-revealing least squares
            coef = np.linalg.lstsq(A, c, rcond=None)[0]
            for n in range(1, iteration):
                out_mixed += coef[n - 1] * (self.output[used - n] - self.output[used])
                in_mixed += coef[n - 1] * (self.input[used - n] - self.input[used])
        return self.alpha * out_mixed + (1 - self.alpha) * in_mixed


class QPFunc:
    """f(w) = Sigma_c(w) + offset - w  (gw.h:214-249)."""

    def __init__(self, level, sigma, offset):
        self.level, self.sigma_c, self.offset = level, sigma, offset
        self.n_sigma = 0
        self.n_deriv = 0

    def sigma(self, w):
        self.n_sigma += 1
        return self.sigma_c.calc_correlation_diag_element(self.level, w)

    def value(self, w):
        return self.sigma(w) + self.offset - w

    def deriv(self, w):
        self.n_deriv += 1
        return self.sigma_c.calc_correlation_diag_element_derivative(self.level, w) - 1.0


class GW:
    def __init__(self, Mmn, vxc, dft_energies):
        self.Mmn = Mmn
        self.vxc = np.asarray(vxc)
        self.dft_energies = np.asarray(dft_energies, dtype=np.float64)
        self.rpa = RPA(Mmn)

    # gw.cc:35-58
    def configure(self, opt):
        self.opt = opt
        qps.normalize_grid_search_options(opt)
        self.qptotal = opt.qpmax - opt.qpmin + 1
        self.rpa.configure(opt.homo, opt.rpamin, opt.rpamax)
        self.sigma = osigma.create(opt.sigma_integration, self.Mmn, self.rpa)
        self.sigma.configure(osigma.SigmaOptions(
            homo=opt.homo, qpmin=opt.qpmin, qpmax=opt.qpmax, rpamin=opt.rpamin, rpamax=opt.rpamax,
            eta=opt.eta, quadrature_scheme=opt.quadrature_scheme, order=opt.order, alpha=opt.alpha))
        self.Sigma_x = np.zeros((self.qptotal, self.qptotal))
        self.Sigma_c = np.zeros((self.qptotal, self.qptotal))
        self.gw_sc_iteration = 0
        self.sigma_evals = 0

    def _solver_opt(self):
        o = self.opt
        return qps.SolverOptions(
            g_sc_limit=o.g_sc_limit, qp_bisection_max_iter=o.g_sc_max_iterations,
            qp_full_window_half_width=o.qp_full_window_half_width, qp_dense_spacing=o.qp_dense_spacing,
            qp_adaptive_shell_width=o.qp_adaptive_shell_width,
            qp_adaptive_shell_count=o.qp_adaptive_shell_count)

    def get_hqp(self):
        o = self.opt
        return (self.Sigma_x + self.Sigma_c - self.vxc
                + np.diag(self.dft_energies[o.qpmin:o.qpmin + self.qptotal]))

    def get_gwa_results(self):
        o = self.opt
        if getattr(self, "qsgw_final_energies", None) is not None:  # gw.cc:312-318
            return self.qsgw_final_energies
        return (np.diag(self.Sigma_x) + np.diag(self.Sigma_c) - np.diag(self.vxc)
                + self.dft_energies[o.qpmin:o.qpmin + self.qptotal])

    def rpa_input_energies(self):
        return self.rpa.get_rpa_input_energies()

    # gw.cc:1132-1182 (states: IndexParser.cc:36-70); array positions counted from qpmin
    def plot_sigma(self, steps, spacing, states):
        """(levels kept, table of shape steps x 2 * len(levels): omega, E_QP(omega) per level)"""
        o = self.opt
        ids = set()
        for tok in states.replace(",", " ").split():
            if ":" in tok:
                a, b = tok.split(":")
                ids.update(range(int(a), int(b) + 1))
            else:
                ids.add(int(tok))
        levels = [l for l in sorted(ids) if o.qpmin <= l <= o.qpmax]
        e = self.rpa.get_rpa_input_energies()
        table = np.zeros((steps, 2 * len(levels)))
        for i, lvl in enumerate(levels):
            rel = lvl - o.qpmin
            icpt = self.dft_energies[lvl] + self.Sigma_x[rel, rel] - self.vxc[rel, rel]
            for g in range(steps):
                w = e[o.qpmin - o.rpamin + rel] + (g - (steps - 1) / 2.0) * spacing
                table[g, 2 * i] = w
                table[g, 2 * i + 1] = self.sigma.calc_correlation_diag_element(rel, w) + icpt
        return levels, table

    def calc_homo_lumo_shift(self, freqs):
        o = self.opt
        dftgap = self.dft_energies[o.homo + 1] - self.dft_energies[o.homo]
        qpgap = freqs[o.homo + 1 - o.qpmin] - freqs[o.homo - o.qpmin]
        return qpgap - dftgap

    # gw.cc:218-310
    def calculate_gw_perturbation(self):
        o = self.opt
        self.Sigma_x = (1 - o.ScaHFX) * self.sigma.calc_exchange_matrix()
        shifted = self.dft_energies.copy()
        shifted[o.homo + 1:] += o.shift
        self.rpa.set_rpa_input_energies(shifted[o.rpamin:o.rpamax + 1])
        freqs = shifted[o.qpmin:o.qpmin + self.qptotal].copy()
        mixing = Anderson(o.gw_mixing_order, o.gw_mixing_alpha)
        self.iterations = 0
        for i_gw in range(o.gw_sc_max_iterations):
            self.gw_sc_iteration = i_gw
            self.iterations = i_gw + 1
            if i_gw % o.reset_3c == 0 and i_gw != 0:
                self.Mmn.rebuild()
            self.sigma.prepare_screening()
            if o.gw_mixing_order > 0 and i_gw > 0:
                mixing.update_input(freqs)
            freqs = self.solve_qp(freqs)
            if o.gw_sc_max_iterations > 1:
                old = self.rpa.get_rpa_input_energies().copy()
                if o.gw_mixing_order > 0 and i_gw > 0:
                    mixing.update_output(freqs)
                    mixed = mixing.mix_history()
                    self.rpa.update_rpa_input_energies(self.dft_energies, mixed, o.qpmin)
                    freqs = mixed
                else:
                    self.rpa.update_rpa_input_energies(self.dft_energies, freqs, o.qpmin)
                if np.abs(self.rpa.get_rpa_input_energies() - old).max() <= o.gw_sc_limit:
                    break
                elif i_gw == o.gw_sc_max_iterations - 1:
                    break
        self.Sigma_c[np.diag_indices(self.qptotal)] = self.sigma.calc_correlation_diag(freqs)

    # gw.cc:798-1130
    def calculate_qsgw(self):
        """Quasiparticle self-consistent GW.  The caller has restored Mmn to the DFT-MO basis (Mmn.rebuild())."""
        o = self.opt
        e_qp_full = self.get_gwa_results().copy()
        self.qsgw_seed_energies = e_qp_full.copy()
        self.qsgw_final_energies = None
        qsgw_qpmax = o.qpmax
        lumo_local = o.homo - o.qpmin + 1
        for n in range(lumo_local, self.qptotal):
            if abs(e_qp_full[n] - self.dft_energies[o.qpmin + n]) > o.qsgw_max_virt_correction:
                qsgw_qpmax = o.qpmin + n - 1
                break
        nq = qsgw_qpmax - o.qpmin + 1
        trimmed = qsgw_qpmax < o.qpmax

        def sigma_options(qpmax):
            return osigma.SigmaOptions(homo=o.homo, qpmin=o.qpmin, qpmax=qpmax, rpamin=o.rpamin, rpamax=o.rpamax,
                                       eta=o.eta, quadrature_scheme=o.quadrature_scheme, order=o.order, alpha=o.alpha)
        if trimmed:
            self.sigma.configure(sigma_options(qsgw_qpmax))
            self.Sigma_x = np.zeros((nq, nq))
            self.Sigma_c = np.zeros((nq, nq))
        e_qp = e_qp_full[:nq].copy()
        self.qsgw_rotation = np.eye(nq)
        if o.ScaHFX > 0.0:
            raise RuntimeError("GW::CalculateQSGW: QSGW is not compatible with hybrid DFT starting points")
        if o.sigma_integration == "cda":
            raise RuntimeError("GW::CalculateQSGW: QSGW is not supported with the CDA sigma integration method.")
        H0 = -self.vxc[:nq, :nq] + np.diag(self.dft_energies[o.qpmin:o.qpmin + nq])
        M_orig = self.Mmn.M.copy()
        mixer = Anderson(o.gw_mixing_order, o.gw_mixing_alpha)
        U = self.qsgw_rotation
        self.rpa.set_qsgw_rotation(U, o.qpmin, o.homo)
        diff_prev = np.finfo(float).max
        tilde = None
        self.qsgw_iterations = 0
        for it in range(o.qsgw_max_iterations):
            self.qsgw_iterations = it + 1
            self.Mmn.M[...] = M_orig
            if it > 0:
                self.Mmn.rotate(self.qsgw_rotation, o.qpmin, qsgw_qpmax)
            self.rpa.set_qsgw_rotation(self.qsgw_rotation, o.qpmin, o.homo)
            self.sigma.prepare_screening()
            self.Sigma_x = self.sigma.calc_exchange_matrix()
            sc_row = self.sigma.calc_correlation_offdiag(e_qp)
            tilde = self.Sigma_x + 0.5 * (sc_row + sc_row.T)
            tilde[np.diag_indices(nq)] += self.sigma.calc_correlation_diag(e_qp)
            s_flat = tilde.reshape(-1, order="F").copy()
            if it > 0:
                mixer.update_output(s_flat)
                s_flat = mixer.mix_history()
            _, dU = np.linalg.eigh(H0 + s_flat.reshape(nq, nq, order="F"))
            e_new = np.linalg.eigvalsh(H0 + tilde)
            diff = np.abs(e_new - e_qp).max()
            if it > 1 and diff > 2.0 * diff_prev:
                mixer = Anderson(o.gw_mixing_order, o.gw_mixing_alpha)
            diff_prev = diff
            if diff < o.qsgw_sc_limit:
                e_qp = e_new
                self.qsgw_rotation = dU
                self.rpa.update_rpa_input_energies(self.dft_energies, e_qp, o.qpmin)
                break
            e_qp = e_new
            self.qsgw_rotation = dU
            mixer.update_input(tilde.reshape(-1, order="F").copy())
            self.rpa.update_rpa_input_energies(self.dft_energies, e_qp, o.qpmin)
        self.Sigma_c[np.diag_indices(nq)] = self.sigma.calc_correlation_diag(e_qp)
        self.rpa.set_qsgw_rotation(None)
        self.qsgw_energies = e_qp
        if trimmed:
            Ufull = np.eye(self.qptotal)
            Ufull[:nq, :nq] = self.qsgw_rotation
            self.qsgw_rotation = Ufull
            merged = e_qp_full.copy()
            merged[:nq] = e_qp
            self.qsgw_final_energies = merged
            self.rpa.update_rpa_input_energies(self.dft_energies, merged, o.qpmin)
            self.Sigma_x = np.zeros((self.qptotal, self.qptotal))
            self.Sigma_c = np.zeros((self.qptotal, self.qptotal))
            self.sigma.configure(sigma_options(o.qpmax))

    # gw.cc:772-776
    def calculate_hqp(self):
        diag = np.diag(self.Sigma_c).copy()
        self.Sigma_c = self.sigma.calc_correlation_offdiag(self.get_gwa_results())
        self.Sigma_c[np.diag_indices(self.qptotal)] = diag

    # gw.cc:323-410
    def solve_qp(self, freqs):
        o = self.opt
        intercepts = (self.dft_energies[o.qpmin:o.qpmin + self.qptotal] + np.diag(self.Sigma_x)
                      - np.diag(self.vxc))
        new = np.array(freqs, dtype=np.float64)
        self.converged = np.zeros(self.qptotal, dtype=bool)
        for lvl in range(self.qptotal):
            f0, icpt = freqs[lvl], intercepts[lvl]
            newf = None
            if o.qp_solver == "fixedpoint":
                newf = self._solve_fixedpoint(icpt, f0, lvl)
            if newf is not None:
                new[lvl] = newf
                self.converged[lvl] = True
            else:
                newf = self._solve_grid(icpt, f0, lvl)
                if newf is not None:
                    new[lvl] = newf
                    self.converged[lvl] = True
                else:
                    newf = self._solve_linearisation(icpt, f0, lvl)
                    if newf is not None:
                        new[lvl] = newf
        return new

    def _solve_linearisation(self, icpt, f0, lvl):
        fqp = QPFunc(lvl, self.sigma, icpt)
        s = fqp.sigma(f0)
        ds = fqp.deriv(f0)  # dSigma/dw - 1
        # gw.cc:420-425: Z = 1 - dsigma_domega where fqp.deriv already subtracts 1
        Z = 1.0 - ds
        self.sigma_evals += fqp.n_sigma
        if abs(Z) > 1e-9:
            return f0 + (icpt - f0 + s) / Z
        return None

    def _solve_fixedpoint(self, icpt, f0, lvl):
        o = self.opt
        f = QPFunc(lvl, self.sigma, icpt)
        x, ok = qps.newton_raphson(f, f0, o.g_sc_max_iterations, o.g_sc_limit, o.qp_solver_alpha)
        self.sigma_evals += f.n_sigma
        return x if ok else None

    # gw.cc:433-502
    def _windowed_adaptive(self, icpt, f0, lvl, lo, hi, allow_rejected):
        fqp = QPFunc(lvl, self.sigma, icpt)
        res, acc, rej, _ = qps.solve_qp_grid_windowed(
            fqp, f0, lo, hi, self.gw_sc_iteration, self._solver_opt(),
            use_brent=(self.opt.qp_root_finder == "brent"))
        self.sigma_evals += fqp.n_sigma
        if acc:
            return res
        if rej and not allow_rejected:
            return None
        return res

    # gw.cc:504-616
    def _windowed_dense(self, icpt, f0, lvl, lo, hi, allow_rejected):
        o = self.opt
        fqp = QPFunc(lvl, self.sigma, icpt)
        sopt = self._solver_opt()
        use_brent = o.qp_root_finder == "brent"
        acc, rej = [], []
        if lo < hi:
            fprev = lo
            tprev = fqp.value(fprev)
            n_steps = max(2, int(math.ceil((hi - lo) / o.qp_dense_spacing)) + 1)
            for i in range(1, n_steps):
                freq = hi if i == n_steps - 1 else min(hi, lo + float(i) * o.qp_dense_spacing)
                targ = fqp.value(freq)
                if tprev * targ < 0.0:
                    cand = qps.refine_qp_interval(fprev, tprev, freq, targ, fqp, f0, sopt, use_brent)
                    if cand is not None:
                        (acc if cand.accepted else rej).append(cand)
                fprev, tprev = freq, targ
        self.sigma_evals += fqp.n_sigma
        if acc:
            return qps._argmax_first(acc).omega
        if rej:
            if not allow_rejected:
                return None
            return qps._argmax_first(rej).omega
        return None

    # gw.cc:618-675
    def _windowed(self, icpt, f0, lvl, lo, hi, allow_rejected):
        mode = self.opt.qp_grid_search_mode
        if mode == "adaptive":
            return self._windowed_adaptive(icpt, f0, lvl, lo, hi, allow_rejected)
        if mode == "dense":
            return self._windowed_dense(icpt, f0, lvl, lo, hi, allow_rejected)
        if mode == "adaptive_with_dense_fallback":
            r = self._windowed_adaptive(icpt, f0, lvl, lo, hi, allow_rejected)
            if r is not None:
                return r
            return self._windowed_dense(icpt, f0, lvl, lo, hi, allow_rejected)
        raise RuntimeError("Unknown gw.qp_grid_search_mode '" + mode + "'")

    # gw.cc:677-739
    def _solve_grid(self, icpt, f0, lvl):
        o = self.opt
        rng = o.qp_full_window_half_width
        full_lo, full_hi = f0 - rng, f0 + rng
        r_lo, r_hi = full_lo, full_hi
        use_restricted = False
        if o.qp_restrict_search:
            mo_level = lvl + o.qpmin
            if mo_level <= o.homo:
                r_hi = min(full_hi, -o.qp_zero_margin)
            else:
                r_lo = max(full_lo, o.qp_virtual_min_energy)
            tol = 1e-12
            use_restricted = abs(r_lo - full_lo) > tol or abs(r_hi - full_hi) > tol
        if use_restricted and r_lo < r_hi:
            r = self._windowed(icpt, f0, lvl, r_lo, r_hi, False)
            if r is not None:
                return r
            return self._windowed_dense(icpt, f0, lvl, full_lo, full_hi, True)
        return self._windowed(icpt, f0, lvl, full_lo, full_hi, True)
