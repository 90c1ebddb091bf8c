"""Oracle for the self-energy evaluators (test infrastructure).

Follows xtp/src/libxtp/gwbse/sigma_base.cc:36-78,
self_energy_evaluators/sigma_ppm.cc:32-126 + gwbse/ppm.cc:30-59,
self_energy_evaluators/sigma_exact.cc:29-148,
self_energy_evaluators/sigma_cda.cc:30-141 + ImaginaryAxisIntegration.cc:90-176
+ gaussian_quadrature/gauss_legendre_quadrature.h:57-67 (nodes/weights are the
standard Gauss-Legendre values; regenerated with numpy leggauss).
"""
import math
from dataclasses import dataclass

import numpy as np
from scipy.special import erfc


@dataclass
class SigmaOptions:
    homo: int = 0
    qpmin: int = 0
    qpmax: int = 0
    rpamin: int = 0
    rpamax: int = 0
    eta: float = 1e-3
    quadrature_scheme: str = "legendre"
    order: int = 12
    alpha: float = 1e-3


class SigmaBase:
    def __init__(self, Mmn, rpa):
        self.Mmn, self.rpa = Mmn, rpa

    def configure(self, opt):
        self.opt = opt
        self.qptotal = opt.qpmax - opt.qpmin + 1
        self.rpatotal = opt.rpamax - opt.rpamin + 1

    # sigma_base.cc:36-52
    def calc_exchange_matrix(self):
        occ = self.opt.homo - self.opt.rpamin + 1
        off = self.opt.qpmin - self.opt.rpamin
        q = self.qptotal
        res = np.zeros((q, q))
        for i in range(q):
            M1 = self.Mmn[i + off][:occ, :]
            for j in range(i, q):
                M2 = self.Mmn[j + off][:occ, :]
                res[j, i] = -np.sum(M1 * M2)
                res[i, j] = res[j, i]
        return res

    # sigma_base.cc:54-63
    def calc_correlation_diag(self, freqs):
        return np.array([self.calc_correlation_diag_element(i, freqs[i]) for i in range(self.qptotal)])

    # sigma_base.cc:65-78
    def calc_correlation_offdiag(self, freqs):
        q = self.qptotal
        res = np.zeros((q, q))
        for i in range(q):
            for j in range(i + 1, q):
                res[j, i] = self.calc_correlation_offdiag_element(i, j, freqs[i], freqs[j])
                res[i, j] = res[j, i]
        return res


class SigmaPPM(SigmaBase):
    SCREENING_R = 0.0  # ppm.h
    SCREENING_I = 0.5

    def prepare_screening(self):
        rpa = self.rpa
        ev, phi = np.linalg.eigh(rpa.calculate_epsilon_r(self.SCREENING_R))
        weight = 1.0 - 1.0 / ev
        ortho = phi.T @ rpa.calculate_epsilon_i(self.SCREENING_I) @ phi
        eps1inv = np.linalg.inv(ortho)
        freq = np.zeros_like(ev)
        for i in range(len(ev)):
            if weight[i] < 1e-5:
                weight[i] = 0.0
                freq[i] = 0.5
            else:
                nom = eps1inv[i, i] - 1.0
                frac = -1.0 * nom / (nom + weight[i]) * self.SCREENING_I * self.SCREENING_I
                freq[i] = math.sqrt(abs(frac))
        self.ppm_phi, self.ppm_weight, self.ppm_freq = phi, weight, freq
        self.Mmn.multiply_right(phi)

    def _terms(self, frequency):
        """t[n, chi] = w - e_n +- Omega_chi for the active poles."""
        lumo = self.opt.homo + 1
        e = self.rpa.get_rpa_input_energies()
        act = self.ppm_weight >= 1e-9
        t = frequency - e[:, None] + np.zeros((1, act.sum()))
        t[:lumo, :] += self.ppm_freq[act][None, :]
        t[lumo:, :] -= self.ppm_freq[act][None, :]
        return act, t

    def calc_correlation_diag_element(self, level, frequency):
        eta2 = self.opt.eta ** 2
        off = self.opt.qpmin - self.opt.rpamin
        act, t = self._terms(frequency)
        fac = 0.5 * self.ppm_weight[act] * self.ppm_freq[act]
        M2 = self.Mmn[level + off][:, act] ** 2
        return float(np.sum(fac[None, :] * M2 * t / (t * t + eta2)))

    def calc_correlation_diag_element_derivative(self, level, frequency):
        eta2 = self.opt.eta ** 2
        off = self.opt.qpmin - self.opt.rpamin
        act, t = self._terms(frequency)
        fac = 0.5 * self.ppm_weight[act] * self.ppm_freq[act]
        M2 = self.Mmn[level + off][:, act] ** 2
        den = t * t + eta2
        return float(np.sum(fac[None, :] * (eta2 - t * t) * M2 / (den * den)))

    def calc_correlation_offdiag_element(self, l1, l2, f1, f2):
        eta2 = self.opt.eta ** 2
        off = self.opt.qpmin - self.opt.rpamin
        act, t1 = self._terms(f1)
        _, t2 = self._terms(f2)
        fac = 0.25 * self.ppm_weight[act] * self.ppm_freq[act]
        MM = self.Mmn[l1 + off][:, act] * self.Mmn[l2 + off][:, act]
        return float(np.sum(fac[None, :] * (t1 / (t1 * t1 + eta2) + t2 / (t2 * t2 + eta2)) * MM))


class SigmaExact(SigmaBase):
    def prepare_screening(self):
        self.rpa_omegas, XpY, self.erpa = self.rpa.diagonalize_h2p()
        self.residues = [self._calc_residues(i, XpY) for i in range(self.qptotal)]

    # sigma_exact.cc:109-148
    def _calc_residues(self, level, XpY):
        lumo = self.opt.homo + 1
        n_occ = lumo - self.opt.rpamin
        n_unocc = self.opt.rpamax - self.opt.homo
        off = self.opt.qpmin - self.opt.rpamin
        Mi = self.Mmn[level + off]
        res = np.zeros((self.rpatotal, n_occ * n_unocc))
        for v in range(n_occ):
            Mv = self.rpa.hole_slice(v, n_unocc)  # QSGW: rotated inside the QP window (sigma_exact.cc:119-145)
            fc = Mv @ Mi.T
            res += fc.T @ XpY[v * n_unocc:(v + 1) * n_unocc, :]
        return res

    def _terms(self, frequency):
        lumo = self.opt.homo + 1
        n_occ = lumo - self.opt.rpamin
        e = self.rpa.get_rpa_input_energies()
        t = frequency - e[:, None] + np.zeros((1, len(self.rpa_omegas)))
        t[:n_occ, :] += self.rpa_omegas[None, :]
        t[n_occ:, :] -= self.rpa_omegas[None, :]
        return t

    def calc_correlation_diag_element(self, level, frequency):
        eta2 = self.opt.eta ** 2
        t = self._terms(frequency)
        r2 = self.residues[level] ** 2
        return float(2.0 * np.sum(r2 * t / (t * t + eta2)))

    def calc_correlation_diag_element_derivative(self, level, frequency):
        eta2 = self.opt.eta ** 2
        t = self._terms(frequency)
        r2 = self.residues[level] ** 2
        den = t * t + eta2
        return float(2.0 * np.sum((eta2 - t * t) * r2 / (den * den)))

    def calc_correlation_offdiag_element(self, l1, l2, f1, f2):
        eta2 = self.opt.eta ** 2
        t1, t2 = self._terms(f1), self._terms(f2)
        r12 = self.residues[l1] * self.residues[l2]
        return float(2.0 * 0.5 * np.sum(r12 * t1 / (t1 * t1 + eta2) + r12 * t2 / (t2 * t2 + eta2)))


class SigmaCDA(SigmaBase):
    def _quadrature(self):
        x, w = np.polynomial.legendre.leggauss(self.opt.order)
        if self.opt.quadrature_scheme == "legendre":
            pts = np.tan(0.5 * math.pi * x)
            wts = w * 0.5 * math.pi / np.cos(0.5 * math.pi * x) ** 2
            return pts, wts, False
        if self.opt.quadrature_scheme == "modified_legendre":
            pts = 0.5 * (1.0 + x) / (1.0 - x)
            wts = w / (1.0 - x) ** 2
            return pts, wts, True
        raise ValueError("oracle restates the legendre quadratures only")

    # sigma_cda.cc:30-45, ImaginaryAxisIntegration.cc:90-102
    def prepare_screening(self):
        rpa = self.rpa
        k0 = np.linalg.inv(rpa.calculate_epsilon_r(complex(0.0, 0.0)))
        k0[np.diag_indices_from(k0)] -= 1.0
        self.kzero = k0
        self.pts, self.wts, self.symmetry = self._quadrature()
        self.dielinv = []
        for p in self.pts:
            ei = np.linalg.inv(rpa.calculate_epsilon_i(p))
            ei[np.diag_indices_from(ei)] -= 1.0
            self.dielinv.append(-ei + k0 * math.exp(-(self.opt.alpha * p) ** 2))

    # ImaginaryAxisIntegration.cc:104-176
    def _sigma_gq_diag(self, frequency, level, eta):
        lumo = self.opt.homo + 1
        occ = lumo - self.opt.rpamin
        unocc = self.opt.rpamax - self.opt.homo
        off = self.opt.qpmin - self.opt.rpamin
        Imx = self.Mmn[level + off]
        e = self.rpa.get_rpa_input_energies()
        dE = (frequency - e).astype(np.complex128)
        dE[:occ] += 1j * eta
        dE[len(dE) - unocc:] += -1j * eta
        total = 0.0
        for j, (p, w) in enumerate(zip(self.pts, self.wts)):
            cp = 1j * p
            if self.symmetry:
                den = 1.0 / (dE + cp) + 1.0 / (dE - cp)
            else:
                den = 1.0 / (dE + cp)
            val = 0.5 / math.pi * np.sum((Imx @ self.dielinv[j]) * (den[:, None] * Imx)).real
            total += w * val
        return total

    @staticmethod
    def _residue_prefactor(e_f, e_m, frequency):
        tol = 1e-10
        if e_f < e_m and e_m < frequency:
            return 1.0
        if e_f > e_m and e_m > frequency:
            return -1.0
        if abs(e_m - frequency) < tol and e_f > e_m:
            return -0.5
        if abs(e_m - frequency) < tol and e_f < e_m:
            return 0.5
        return 0.0

    def _residue_contribution(self, frequency, level):
        e = self.rpa.get_rpa_input_energies()
        off = self.opt.qpmin - self.opt.rpamin
        homo = self.opt.homo - self.opt.rpamin
        fermi = 0.5 * (e[homo + 1] + e[homo])
        Imx = self.Mmn[level + off]
        sig, tail = 0.0, 0.0
        for i in range(len(e)):
            delta = e[i] - frequency
            ad = abs(delta)
            fac = self._residue_prefactor(fermi, e[i], frequency)
            row = Imx[i, :]
            if abs(fac) > 1e-10:
                eps = self.rpa.calculate_epsilon_r(complex(ad, self.rpa.get_eta()))
                x = np.linalg.solve(eps, row) - row
                sig += fac * float(x @ row)
            if ad > 1e-10:
                ef = 0.5 * math.copysign(1.0, delta) * math.exp((self.opt.alpha * delta) ** 2) * erfc(
                    abs(self.opt.alpha * delta))
                tail += float((row @ self.kzero) @ row) * ef
        return sig + tail

    def calc_correlation_diag_element(self, level, frequency):
        return self._residue_contribution(frequency, level) + self._sigma_gq_diag(
            frequency, level, self.rpa.get_eta())

    def calc_correlation_diag_element_derivative(self, level, frequency):
        h = 1e-3
        return (self.calc_correlation_diag_element(level, frequency + h)
                - self.calc_correlation_diag_element(level, frequency - h)) / (2 * h)

    def calc_correlation_offdiag_element(self, l1, l2, f1, f2):
        return 0.0


def create(name, Mmn, rpa):
    """SigmaFactory, xtp/src/libxtp/factories/sigmafactory.cc:32-36."""
    return {"ppm": SigmaPPM, "exact": SigmaExact, "cda": SigmaCDA}[name](Mmn, rpa)
