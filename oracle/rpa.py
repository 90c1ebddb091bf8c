"""Oracle for RPA (test infrastructure).  Follows xtp/src/libxtp/gwbse/rpa.cc.

The QSGW on-the-fly m-rotation of the hole slices (rpa.cc:95-118, 162-181, 288-310; rpa.h:59-66) is restated in
hole_slice().
"""
import numpy as np


class RPA:
    ETA = 0.0001  # rpa.h:90

    def __init__(self, Mmn):
        self.Mmn = Mmn
        self.energies = None

    def configure(self, homo, rpamin, rpamax):
        self.homo, self.rpamin, self.rpamax = homo, rpamin, rpamax
        self.qsgw_U = None

    # rpa.h:59-66
    def set_qsgw_rotation(self, U, qpmin=0, homo=0):
        self.qsgw_U, self.qsgw_qpmin, self.qsgw_homo = U, qpmin, homo

    def hole_slice(self, m_level, n_unocc):
        """Unoccupied rows of hole slice m_level; inside the QP window they are rotated to the current QP
        wavefunctions: sum_vp U(vp, v_qp) Mmn[vp + qp_offset] (rpa.cc:95-118)."""
        nt = self.Mmn.nsize()
        if self.qsgw_U is not None:
            off = self.qsgw_qpmin - self.rpamin
            qptotal = self.qsgw_U.shape[1]
            end_occ = min(self.qsgw_homo - self.rpamin + 1, off + qptotal)
            if off <= m_level < end_occ:
                out = np.zeros((n_unocc, self.Mmn.auxsize()))
                for vp in range(qptotal):
                    out += self.qsgw_U[vp, m_level - off] * self.Mmn[vp + off][nt - n_unocc:, :]
                return out
        return self.Mmn[m_level][nt - n_unocc:, :]

    def get_eta(self):
        return self.ETA

    def set_rpa_input_energies(self, e):
        self.energies = np.array(e, dtype=np.float64)

    def get_rpa_input_energies(self):
        return self.energies

    # rpa.cc:32-73
    def update_rpa_input_energies(self, dftenergies, gwaenergies, qpmin):
        rpatotal = self.rpamax - self.rpamin + 1
        self.energies = np.array(dftenergies[self.rpamin:self.rpamin + rpatotal], dtype=np.float64)
        gwsize = len(gwaenergies)
        self.energies[qpmin - self.rpamin:qpmin - self.rpamin + gwsize] = gwaenergies
        lumo = self.homo + 1
        qpmax = qpmin + gwsize - 1

        def max_corr(lo, hi):
            rng = hi - lo + 1
            corr = self.energies[lo:lo + rng] - dftenergies[lo - self.rpamin:lo - self.rpamin + rng]
            return np.abs(corr).max()

        occ = max_corr(qpmin, self.homo)
        virt = max_corr(lumo, qpmax)
        self.energies[:qpmin] -= occ
        ntail = self.rpamax - qpmax
        if ntail > 0:
            self.energies[len(self.energies) - ntail:] += virt

    def _sizes(self):
        lumo = self.homo + 1
        n_occ = lumo - self.rpamin
        n_unocc = self.rpamax - lumo + 1
        return n_occ, n_unocc

    # rpa.cc:75-140
    def _epsilon(self, frequency, imag):
        n_occ, n_unocc = self._sizes()
        naux = self.Mmn.auxsize()
        res = np.zeros((naux, naux))
        e = self.energies
        for m in range(n_occ):
            Mv = self.hole_slice(m, n_unocc)
            dE = e[len(e) - n_unocc:] - e[m]
            if imag:
                d = 4.0 * dE / (dE * dE + frequency * frequency)
            else:
                eta2 = self.ETA * self.ETA
                dm = dE - frequency
                s = dm / (dm * dm + eta2)
                dp = dE + frequency
                s = s + dp / (dp * dp + eta2)
                d = 2.0 * s
            res += Mv.T @ (d[:, None] * Mv)
        res[np.diag_indices(naux)] += 1.0
        return res

    def calculate_epsilon_i(self, frequency):
        return self._epsilon(frequency, True)

    def calculate_epsilon_r(self, frequency):
        if isinstance(frequency, complex):
            return self._epsilon_r_complex(frequency)
        return self._epsilon(frequency, False)

    # rpa.cc:145-202
    def _epsilon_r_complex(self, frequency):
        n_occ, n_unocc = self._sizes()
        naux = self.Mmn.auxsize()
        res = np.zeros((naux, naux))
        e = self.energies
        for m in range(n_occ):
            Mv = self.hole_slice(m, n_unocc)
            dE = e[len(e) - n_unocc:] - e[m]
            dEm = frequency.real - dE
            dEp = frequency.real + dE
            s1 = (frequency.imag + self.ETA) ** 2
            s2 = (frequency.imag - self.ETA) ** 2
            chi = dEm / (dEm * dEm + s1) - dEp / (dEp * dEp + s2)
            res += Mv.T @ (chi[:, None] * Mv)
        res = -2.0 * res
        res[np.diag_indices(naux)] += 1.0
        return res

    # rpa.cc:266-326
    def h2p_amb(self):
        n_occ, n_unocc = self._sizes()
        e = self.energies
        amb = np.zeros(n_occ * n_unocc)
        for v in range(n_occ):
            amb[v * n_unocc:(v + 1) * n_unocc] = e[n_occ:n_occ + n_unocc] - e[v]
        return amb

    def h2p_apb(self):
        n_occ, n_unocc = self._sizes()
        S = n_occ * n_unocc
        apb = np.zeros((S, S))
        rot = [self.hole_slice(v, n_unocc) for v in range(n_occ)]  # rpa.cc:288-310
        for v2 in range(n_occ):
            M2 = rot[v2]
            for v1 in range(v2, n_occ):
                M1 = rot[v1]
                apb[v1 * n_unocc:(v1 + 1) * n_unocc, v2 * n_unocc:(v2 + 1) * n_unocc] = 4.0 * M1 @ M2.T
        apb[np.diag_indices(S)] += self.h2p_amb()
        return apb  # lower triangle filled (Eigen solver uses the lower triangle)

    # rpa.cc:204-264
    def diagonalize_h2p(self):
        amb = self.h2p_amb()
        apb = self.h2p_apb()
        erpa = -0.25 * (np.trace(apb) + amb.sum())
        sq = np.sqrt(amb)
        C = apb * sq[:, None] * sq[None, :]
        ev, evec = np.linalg.eigh(C, UPLO="L")
        if ev.min() <= 0.0:
            raise RuntimeError("Detected non-positive eigenvalue.")
        omega = np.sqrt(ev)
        erpa += 0.5 * omega.sum()
        XpY = (sq[:, None] * evec) / np.sqrt(omega)[None, :]
        return omega, XpY, erpa
