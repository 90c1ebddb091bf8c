// RPA - host mirror of xtp/include/votca/xtp/rpa.h:35-115 / xtp/src/libxtp/gwbse/rpa.cc.
// Energies bookkeeping stays on the host (scalar arithmetic, copied statement by statement);
// epsilon assembly and the two-particle matrix are single C-ABI calls.
#pragma once
#include <complex>

#include "threecenter.h"

namespace votca {
namespace xtp {

class RPA {
 public:
  RPA(Logger& log, const TCMatrix_gwbse& Mmn) : log_(log), Mmn_(Mmn) {}

  void configure(Index homo, Index rpamin, Index rpamax) {
    homo_ = homo;
    rpamin_ = rpamin;
    rpamax_ = rpamax;
  }
  double getEta() const { return eta_; }
  Index homo() const { return homo_; }
  Index rpamin() const { return rpamin_; }
  Index rpamax() const { return rpamax_; }

  MatrixXd calculate_epsilon_i(double frequency) const { return epsilon(0, frequency, 0.0, true); }
  MatrixXd calculate_epsilon_r(double frequency) const { return epsilon(1, frequency, 0.0, true); }
  MatrixXd calculate_epsilon_r(std::complex<double> frequency) const {
    return epsilon(2, frequency.real(), frequency.imag(), true);
  }
  // device-resident variants: result stays in the context (gwbse_rpa_epsilon_ptr), nothing crosses PCIe
  double* calculate_epsilon_i_dev(double frequency) const { return epsilon_dev(0, frequency, 0.0); }
  double* calculate_epsilon_r_dev(double frequency) const { return epsilon_dev(1, frequency, 0.0); }
  double* calculate_epsilon_r_dev(std::complex<double> f) const { return epsilon_dev(2, f.real(), f.imag()); }

  const VectorXd& getRPAInputEnergies() const { return energies_; }
  void setRPAInputEnergies(const VectorXd& rpaenergies) { energies_ = rpaenergies; }

  // rpa.cc:32-45
  void UpdateRPAInputEnergies(const VectorXd& dftenergies, const VectorXd& gwaenergies, Index qpmin) {
    Index rpatotal = rpamax_ - rpamin_ + 1;
    energies_ = dftenergies.segment(rpamin_, rpatotal);
    Index gwsize = gwaenergies.size();
    for (Index i = 0; i < gwsize; ++i) energies_(qpmin - rpamin_ + i) = gwaenergies(i);
    ShiftUncorrectedEnergies(dftenergies, qpmin, gwsize);
  }

  struct rpa_eigensolution {
    VectorXd omega;
    MatrixXd XpY;  // host copy only when requested
    double ERPA_correlation;
  };

  // rpa.cc:204-264.  The S x S matrices stay on the device; XpY_dev receives (X+Y).
  rpa_eigensolution Diagonalize_H2p(Device::Buffer* XpY_dev = nullptr, bool fetch_XpY = true) const {
    const Device& dev = Mmn_.device();
    if (dev.world() > 1) throw std::runtime_error("Diagonalize_H2p is single-GPU (S x S matrix not sharded)");
    const Index lumo = homo_ + 1;
    const Index n_occ = lumo - rpamin_;
    const Index n_unocc = rpamax_ - lumo + 1;
    const Index rpasize = n_occ * n_unocc;
    VectorXd AmB = Calculate_H2p_AmB();
    Device::Buffer C = dev.alloc(static_cast<size_t>(rpasize * rpasize));
    dev.check(gwbse_rpa_h2p_apb(dev.ctx(), energies_.data(), (int)homo_, (int)rpamin_, (int)rpamax_, C.get(),
                                (int)rpasize));
    rpa_eigensolution sol;
    // trace(ApB): diagonal = 4 sum_chi M^2 + AmB
    MatrixXd ApBdiag = download_diag(dev, C.get(), rpasize);
    sol.ERPA_correlation = -0.25 * (ApBdiag.col(0).sum() + AmB.sum());
    // C = AmB^1/2 * ApB * AmB^1/2
    VectorXd sq(rpasize);
    for (Index i = 0; i < rpasize; ++i) sq(i) = std::sqrt(AmB(i));
    Device::Buffer dsq = dev.upload(sq);
    dev.check(gwbse_diag_scale_dev(dev.ctx(), 'L', (int)rpasize, (int)rpasize, C.get(), (int)rpasize, dsq.get(),
                                   C.get(), (int)rpasize));
    dev.check(gwbse_diag_scale_dev(dev.ctx(), 'R', (int)rpasize, (int)rpasize, C.get(), (int)rpasize, dsq.get(),
                                   C.get(), (int)rpasize));
    log_(" Diagonalizing two-particle Hamiltonian ");
    VectorXd ev(rpasize);
    dev.check(gwbse_sym_eig_dev(dev.ctx(), (int)rpasize, C.get(), (int)rpasize, ev.data()));
    log_(" Diagonalization done ");
    double minCoeff = ev(0);
    for (Index i = 0; i < rpasize; ++i) minCoeff = std::min(minCoeff, ev(i));
    if (minCoeff <= 0.0) {
      log_(" Detected non-positive eigenvalue: " + std::to_string(minCoeff));
      throw std::runtime_error("Detected non-positive eigenvalue.");
    }
    sol.omega = VectorXd(rpasize);
    VectorXd osi(rpasize);
    for (Index i = 0; i < rpasize; ++i) {
      sol.omega(i) = std::sqrt(ev(i));
      osi(i) = 1.0 / std::sqrt(sol.omega(i));
    }
    sol.ERPA_correlation += 0.5 * sol.omega.sum();
    log_(" RPA correlation energy (Hartree): " + std::to_string(sol.ERPA_correlation));
    // XpY.col(s) = Omega_s^-1/2 * AmB^1/2 .* z_s
    Device::Buffer dos = dev.upload(osi);
    dev.check(gwbse_diag_scale_dev(dev.ctx(), 'L', (int)rpasize, (int)rpasize, C.get(), (int)rpasize, dsq.get(),
                                   C.get(), (int)rpasize));
    dev.check(gwbse_diag_scale_dev(dev.ctx(), 'R', (int)rpasize, (int)rpasize, C.get(), (int)rpasize, dos.get(),
                                   C.get(), (int)rpasize));
    if (fetch_XpY) sol.XpY = dev.download(C.get(), rpasize, rpasize);
    if (XpY_dev) *XpY_dev = std::move(C);
    return sol;
  }

 private:
  static MatrixXd download_diag(const Device& dev, const double* A, Index n) {
    // gather the diagonal with a strided device copy: treat it as a 1 x n block with ld n+1
    Device::Buffer d = dev.alloc(static_cast<size_t>(n));
    dev.check(gwbse_dev_memset_zero(dev.ctx(), d.get(), static_cast<size_t>(n)));
    dev.check(gwbse_axpy_dev(dev.ctx(), 1, (int)n, 1.0, A, (int)(n + 1), d.get(), 1));
    return dev.download(d.get(), n, 1);
  }

  MatrixXd epsilon(int kind, double fre, double fim, bool) const {
    const Device& dev = Mmn_.device();
    MatrixXd out(Mmn_.auxsize(), Mmn_.auxsize());
    dev.check(gwbse_rpa_epsilon(dev.ctx(), kind, fre, fim, eta_, energies_.data(), (int)homo_, (int)rpamin_,
                                (int)rpamax_, out.data(), (int)out.rows()));
    return out;
  }
  double* epsilon_dev(int kind, double fre, double fim) const {
    const Device& dev = Mmn_.device();
    dev.check(gwbse_rpa_epsilon(dev.ctx(), kind, fre, fim, eta_, energies_.data(), (int)homo_, (int)rpamin_,
                                (int)rpamax_, nullptr, 0));
    return gwbse_rpa_epsilon_ptr(dev.ctx());
  }

  // rpa.cc:266-279
  VectorXd Calculate_H2p_AmB() const {
    const Index lumo = homo_ + 1;
    const Index n_occ = lumo - rpamin_;
    const Index n_unocc = rpamax_ - lumo + 1;
    VectorXd AmB(n_occ * n_unocc);
    for (Index v = 0; v < n_occ; v++)
      for (Index c = 0; c < n_unocc; ++c) AmB(v * n_unocc + c) = energies_(n_occ + c) - energies_(v);
    return AmB;
  }

  // rpa.cc:52-73
  void ShiftUncorrectedEnergies(const VectorXd& dftenergies, Index qpmin, Index gwsize) {
    Index lumo = homo_ + 1;
    Index qpmax = qpmin + gwsize - 1;
    double max_correction_occ = getMaxCorrection(dftenergies, qpmin, homo_);
    double max_correction_virt = getMaxCorrection(dftenergies, lumo, qpmax);
    for (Index i = 0; i < qpmin; ++i) energies_(i) -= max_correction_occ;
    const Index ntail = rpamax_ - qpmax;
    for (Index i = energies_.size() - ntail; i < energies_.size(); ++i) energies_(i) += max_correction_virt;
  }
  double getMaxCorrection(const VectorXd& dftenergies, Index min, Index max) const {
    double m = 0.0;
    for (Index i = min; i <= max; ++i) m = std::max(m, std::abs(energies_(i) - dftenergies(i - rpamin_)));
    return m;
  }

  Index homo_ = 0, rpamin_ = 0, rpamax_ = 0;
  const double eta_ = 0.0001;
  VectorXd energies_;
  Logger& log_;
  const TCMatrix_gwbse& Mmn_;
};

}  // namespace xtp
}  // namespace votca
