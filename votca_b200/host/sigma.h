// Sigma_base / Sigma_PPM / Sigma_Exact / Sigma_CDA - host mirror of
// xtp/include/votca/xtp/sigma_base.h:33-106, xtp/src/libxtp/gwbse/sigma_base.cc:36-78,
// self_energy_evaluators/sigma_{ppm,exact,cda}.{h,cc}, gwbse/ppm.cc:30-59,
// ImaginaryAxisIntegration.cc:90-176.  Same virtual interface; in addition every evaluator offers a
// batched EvalBatch() so the QP solver can evaluate many (level, frequency) requests in one kernel pass.
#pragma once
#include <atomic>
#include <cmath>
#include <complex>
#include <memory>
#include <string>
#include <vector>

#include "quadrature.h"
#include "rpa.h"

namespace votca {
namespace xtp {

class Sigma_base {
 public:
  Sigma_base(TCMatrix_gwbse& Mmn, const RPA& rpa) : Mmn_(Mmn), rpa_(rpa) {}
  virtual ~Sigma_base() = default;

  struct options {
    Index homo;
    Index qpmin;
    Index qpmax;
    Index rpamin;
    Index rpamax;
    double eta;
    std::string quadrature_scheme;
    Index order;
    double alpha;
  };

  void configure(options opt) {
    opt_ = opt;
    qptotal_ = opt.qpmax - opt.qpmin + 1;
    rpatotal_ = opt.rpamax - opt.rpamin + 1;
  }

  // sigma_base.cc:36-52
  MatrixXd CalcExchangeMatrix() const {
    MatrixXd result(qptotal_, qptotal_);
    const Device& dev = Mmn_.device();
    dev.check(gwbse_sigma_x(dev.ctx(), (int)opt_.homo, (int)opt_.rpamin, (int)opt_.qpmin, (int)opt_.qpmax,
                            result.data(), (int)qptotal_));
    return result;
  }
  // sigma_base.cc:54-63 (one batched pass instead of an OpenMP loop)
  VectorXd CalcCorrelationDiag(const VectorXd& frequencies) const {
    // every rank evaluates the levels whose Mmn slice it owns; the results are summed over ranks
    const Device& dev = Mmn_.device();
    std::vector<int> lv;
    std::vector<double> fr, s;
    for (Index i = 0; i < qptotal_; ++i) {
      if (!OwnsLevel(i)) continue;
      lv.push_back((int)i);
      fr.push_back(frequencies[i]);
    }
    if (!lv.empty()) EvalBatch(lv, fr, s, nullptr);
    VectorXd out(qptotal_, 0.0);
    for (size_t k = 0; k < lv.size(); ++k) out(lv[k]) = s[k];
    if (dev.world() > 1) dev.allreduce(out.data(), static_cast<size_t>(qptotal_));
    return out;
  }
  // owner of the Mmn slice of a qp level (m-cyclic sharding; always true on one GPU)
  bool OwnsLevel(Index gw_level) const {
    return Mmn_.device().owns_slice(gw_level + opt_.qpmin - opt_.rpamin);
  }
  // sigma_base.cc:65-78
  virtual MatrixXd CalcCorrelationOffDiag(const VectorXd& frequencies) const = 0;

  virtual void PrepareScreening() = 0;
  virtual void EvalBatch(const std::vector<int>& levels, const std::vector<double>& freqs,
                         std::vector<double>& sigma, std::vector<double>* dsigma) const = 0;
  // grouped form: level levels[g] at freqs[gptr[g] .. gptr[g+1]); default = flat batch
  virtual void EvalGroups(const std::vector<int>& levels, const std::vector<int>& gptr,
                          const std::vector<double>& freqs, std::vector<double>& sigma,
                          std::vector<double>* dsigma) const {
    std::vector<int> flat(freqs.size());
    for (size_t g = 0; g < levels.size(); ++g)
      for (int i = gptr[g]; i < gptr[g + 1]; ++i) flat[i] = levels[g];
    EvalBatch(flat, freqs, sigma, dsigma);
  }

  double CalcCorrelationDiagElement(Index gw_level, double frequency) const {
    CountDiagEval();
    std::vector<double> s;
    EvalBatch({(int)gw_level}, {frequency}, s, nullptr);
    return s[0];
  }
  double CalcCorrelationDiagElementDerivative(Index gw_level, double frequency) const {
    std::vector<double> s, ds;
    EvalBatch({(int)gw_level}, {frequency}, s, &ds);
    return ds[0];
  }
  void ResetDiagEvalCounter() const { diag_eval_counter_.store(0); }
  std::size_t GetDiagEvalCounter() const { return diag_eval_counter_.load(); }
  void CountDiagEval(std::size_t n = 1) const { diag_eval_counter_.fetch_add(n); }
  Index qptotal() const { return qptotal_; }

 protected:
  options opt_;
  TCMatrix_gwbse& Mmn_;
  const RPA& rpa_;
  Index qptotal_ = 0;
  Index rpatotal_ = 0;

 private:
  mutable std::atomic<std::size_t> diag_eval_counter_{0};
};

// ---------------------------------------------------------------------------------------------
class PPM {
 public:
  PPM() : screening_r(0.0), screening_i(0.5) {}
  // ppm.cc:30-59; phi stays on the device (phi_dev) and is consumed by MultiplyRight without a PCIe trip
  // RPA_T: RPA, or RPA_UKS (uks.h) whose dielectric matrix is the sum over both spin channels
  template <class RPA_T>
  void PPM_construct_parameters(const RPA_T& rpa, const TCMatrix_gwbse& Mmn) {
    const Device& dev = Mmn.device();
    const Index n = Mmn.auxsize();
    double* eps = rpa.calculate_epsilon_r_dev(screening_r);
    phi_dev_ = dev.alloc(static_cast<size_t>(n * n));
    dev.check(gwbse_d2d(dev.ctx(), phi_dev_.get(), eps, static_cast<size_t>(n * n)));
    VectorXd ev(n);
    dev.check(gwbse_sym_eig_dev(dev.ctx(), (int)n, phi_dev_.get(), (int)n, ev.data()));
    ppm_weight_ = VectorXd(n);
    for (Index i = 0; i < n; ++i) ppm_weight_(i) = 1 - 1.0 / ev(i);
    // ortho = phi^T eps_i phi ; epsilon_1_inv = ortho^-1
    double* eps_i = rpa.calculate_epsilon_i_dev(screening_i);
    Device::Buffer tmp = dev.alloc(static_cast<size_t>(n * n));
    Device::Buffer ortho = dev.alloc(static_cast<size_t>(n * n));
    dev.gemm('T', 'N', n, n, n, 1.0, phi_dev_.get(), n, eps_i, n, 0.0, tmp.get(), n);
    dev.gemm('N', 'N', n, n, n, 1.0, tmp.get(), n, phi_dev_.get(), n, 0.0, ortho.get(), n);
    dev.check(gwbse_inverse_dev(dev.ctx(), (int)n, ortho.get(), (int)n));
    Device::Buffer dg = dev.alloc(static_cast<size_t>(n));
    dev.check(gwbse_dev_memset_zero(dev.ctx(), dg.get(), static_cast<size_t>(n)));
    dev.check(gwbse_axpy_dev(dev.ctx(), 1, (int)n, 1.0, ortho.get(), (int)(n + 1), dg.get(), 1));
    MatrixXd inv_diag = dev.download(dg.get(), n, 1);
    ppm_freq_ = VectorXd(n);
    for (Index i = 0; i < n; i++) {
      if (ppm_weight_(i) < 1.e-5) {
        ppm_weight_(i) = 0.0;
        ppm_freq_(i) = 0.5;
        continue;
      } else {
        double nom = inv_diag(i, 0) - 1.0;
        double frac = -1.0 * nom / (nom + ppm_weight_(i)) * screening_i * screening_i;
        ppm_freq_(i) = std::sqrt(std::abs(frac));
      }
    }
  }
  const VectorXd& getPpm_weight() const { return ppm_weight_; }
  const VectorXd& getPpm_freq() const { return ppm_freq_; }
  const double* getPpm_phi_dev() const { return phi_dev_.get(); }
  void FreeMatrix() { phi_dev_ = Device::Buffer(); }

 private:
  double screening_r, screening_i;
  Device::Buffer phi_dev_;
  VectorXd ppm_freq_, ppm_weight_;
};

class Sigma_PPM : public Sigma_base {
 public:
  Sigma_PPM(TCMatrix_gwbse& Mmn, const RPA& rpa) : Sigma_base(Mmn, rpa) {}
  // sigma_ppm.cc:32-35
  void PrepareScreening() override {
    ppm_.PPM_construct_parameters(rpa_, Mmn_);
    InstallPPM(ppm_);
    ppm_.FreeMatrix();
  }
  void EvalBatch(const std::vector<int>& levels, const std::vector<double>& freqs, std::vector<double>& sigma,
                 std::vector<double>* dsigma) const final {
    const Device& dev = Mmn_.device();
    sigma.resize(levels.size());
    if (dsigma) dsigma->resize(levels.size());
    dev.check(gwbse_sigma_update_energies(dev.ctx(), 0, rpa_.getRPAInputEnergies().data()));
    dev.check(gwbse_sigma_ppm_eval(dev.ctx(), (int)levels.size(), levels.data(), freqs.data(), sigma.data(),
                                   dsigma ? dsigma->data() : nullptr));
  }
  void EvalGroups(const std::vector<int>& levels, const std::vector<int>& gptr, const std::vector<double>& freqs,
                  std::vector<double>& sigma, std::vector<double>* dsigma) const final {
    const Device& dev = Mmn_.device();
    sigma.resize(freqs.size());
    if (dsigma) dsigma->resize(freqs.size());
    dev.check(gwbse_sigma_update_energies(dev.ctx(), 0, rpa_.getRPAInputEnergies().data()));
    dev.check(gwbse_sigma_eval_groups(dev.ctx(), 0, (int)levels.size(), levels.data(), gptr.data(), freqs.data(),
                                      sigma.data(), dsigma ? dsigma->data() : nullptr));
  }
  MatrixXd CalcCorrelationOffDiag(const VectorXd& frequencies) const final {
    MatrixXd out(qptotal_, qptotal_);
    const Device& dev = Mmn_.device();
    dev.check(gwbse_sigma_update_energies(dev.ctx(), 0, rpa_.getRPAInputEnergies().data()));
    dev.check(gwbse_sigma_ppm_offdiag(dev.ctx(), (int)qptotal_, frequencies.data(), out.data(), (int)qptotal_));
    return out;
  }
  const PPM& ppm() const { return ppm_; }

 protected:
  // rotate Mmn into the plasmon-pole basis and hand weights / frequencies to the evaluator kernels
  void InstallPPM(const PPM& ppm) {
    Mmn_.MultiplyRightWithAuxMatrix_dev(ppm.getPpm_phi_dev(), Mmn_.auxsize());
    const Device& dev = Mmn_.device();
    dev.check(gwbse_sigma_ppm_set(dev.ctx(), ppm.getPpm_weight().data(), ppm.getPpm_freq().data(),
                                  rpa_.getRPAInputEnergies().data(), (int)opt_.homo, (int)opt_.rpamin,
                                  (int)opt_.qpmin, opt_.eta));
  }

 private:
  PPM ppm_;
};

class Sigma_Exact : public Sigma_base {
 public:
  Sigma_Exact(TCMatrix_gwbse& Mmn, const RPA& rpa) : Sigma_base(Mmn, rpa) {}
  // sigma_exact.cc:29-38, 109-148
  void PrepareScreening() final {
    if (Mmn_.device().world() > 1)
      throw std::runtime_error("sigma_integrator=exact is single-GPU (the S x S two-particle matrix is not sharded)");
    Device::Buffer XpY;
    RPA::rpa_eigensolution sol = rpa_.Diagonalize_H2p(&XpY, false);
    rpa_omegas_ = sol.omega;
    ERPA_correlation_ = sol.ERPA_correlation;
    const Device& dev = Mmn_.device();
    const Index S = rpa_omegas_.size();
    dev.check(gwbse_sigma_exact_prepare(dev.ctx(), rpa_omegas_.data(), XpY.get(), (int)S,
                                        rpa_.getRPAInputEnergies().data(), (int)opt_.homo, (int)opt_.rpamin,
                                        (int)opt_.rpamax, (int)opt_.qpmin, (int)opt_.qpmax, opt_.eta));
  }
  void EvalBatch(const std::vector<int>& levels, const std::vector<double>& freqs, std::vector<double>& sigma,
                 std::vector<double>* dsigma) const final {
    const Device& dev = Mmn_.device();
    sigma.resize(levels.size());
    if (dsigma) dsigma->resize(levels.size());
    dev.check(gwbse_sigma_update_energies(dev.ctx(), 1, rpa_.getRPAInputEnergies().data()));
    dev.check(gwbse_sigma_exact_eval(dev.ctx(), (int)levels.size(), levels.data(), freqs.data(), sigma.data(),
                                     dsigma ? dsigma->data() : nullptr));
  }
  void EvalGroups(const std::vector<int>& levels, const std::vector<int>& gptr, const std::vector<double>& freqs,
                  std::vector<double>& sigma, std::vector<double>* dsigma) const final {
    const Device& dev = Mmn_.device();
    sigma.resize(freqs.size());
    if (dsigma) dsigma->resize(freqs.size());
    dev.check(gwbse_sigma_update_energies(dev.ctx(), 1, rpa_.getRPAInputEnergies().data()));
    dev.check(gwbse_sigma_eval_groups(dev.ctx(), 1, (int)levels.size(), levels.data(), gptr.data(), freqs.data(),
                                      sigma.data(), dsigma ? dsigma->data() : nullptr));
  }
  MatrixXd CalcCorrelationOffDiag(const VectorXd& frequencies) const final {
    MatrixXd out(qptotal_, qptotal_);
    const Device& dev = Mmn_.device();
    dev.check(gwbse_sigma_update_energies(dev.ctx(), 1, rpa_.getRPAInputEnergies().data()));
    dev.check(gwbse_sigma_exact_offdiag(dev.ctx(), (int)qptotal_, frequencies.data(), out.data(), (int)qptotal_));
    return out;
  }
  const VectorXd& rpa_omegas() const { return rpa_omegas_; }
  double ERPA_correlation() const { return ERPA_correlation_; }

 private:
  VectorXd rpa_omegas_;
  double ERPA_correlation_ = 0.0;
};

class Sigma_CDA : public Sigma_base {
 public:
  Sigma_CDA(TCMatrix_gwbse& Mmn, const RPA& rpa) : Sigma_base(Mmn, rpa) {}

  // sigma_cda.cc:30-45 + ImaginaryAxisIntegration::CalcDielInvVector (ImaginaryAxisIntegration.cc:90-102): the
  // kappa matrices and, new here, the frequency-independent row forms (I kappa_j)[n,:] . I[n,:] of every level are
  // built on the device (gwbse_sigma_cda_prepare); nothing of Mmn crosses PCIe
  void PrepareScreening() final {
    const Device& dev = Mmn_.device();
    if (dev.world() > 1)
      throw std::runtime_error(
          "sigma_integrator=cda is single-GPU: every residue needs a collective eps(z) assembly, and the per-rank QP "
          "searches ask for them at different times");
    BindDielectricSource();
    bool symmetry = false;
    mapped_gauss_legendre(opt_.quadrature_scheme, opt_.order, pts_, wts_, symmetry);
    dev.check(gwbse_sigma_cda_prepare(dev.ctx(), (int)pts_.size(), pts_.data(), wts_.data(), symmetry ? 1 : 0,
                                      opt_.alpha, rpa_.getRPAInputEnergies().data(), (int)opt_.homo, (int)opt_.rpamin,
                                      (int)opt_.rpamax, (int)opt_.qpmin, (int)opt_.qpmax, rpa_.getEta()));
  }

  // sigma_cda.cc:118-124 for a batch of requests; the derivative is the reference's central difference with
  // h = 1e-3 (sigma_cda.h:57-63), its two extra evaluations ride in the same batch
  void EvalBatch(const std::vector<int>& levels, const std::vector<double>& freqs, std::vector<double>& sigma,
                 std::vector<double>* dsigma) const final {
    const Device& dev = Mmn_.device();
    const size_t n = levels.size();
    sigma.resize(n);
    std::vector<int> lv(levels);
    std::vector<double> fr(freqs);
    const double h = 1e-3;
    if (dsigma) {
      dsigma->resize(n);
      for (size_t r = 0; r < n; ++r) {
        lv.push_back(levels[r]);
        fr.push_back(freqs[r] + h);
        lv.push_back(levels[r]);
        fr.push_back(freqs[r] - h);
      }
    }
    std::vector<double> out(lv.size());
    BindDielectricSource();
    dev.check(gwbse_sigma_cda_eval(dev.ctx(), (int)lv.size(), lv.data(), fr.data(), rpa_.getRPAInputEnergies().data(),
                                   out.data()));
    for (size_t r = 0; r < n; ++r) {
      sigma[r] = out[r];
      if (dsigma) (*dsigma)[r] = (out[n + 2 * r] - out[n + 2 * r + 1]) / (2 * h);
    }
  }
  // sigma_cda.h:64-67
  MatrixXd CalcCorrelationOffDiag(const VectorXd&) const final { return MatrixXd::Zero(qptotal_, qptotal_); }

 protected:
  // which dielectric matrix the device assembles: the channel's own (restricted), or - overridden by the unrestricted
  // evaluator - the one of both spin channels
  virtual void BindDielectricSource() const {
    const Device& dev = Mmn_.device();
    dev.check(gwbse_sigma_cda_set_partner(dev.ctx(), nullptr, 0, nullptr));
  }

 private:
  std::vector<double> pts_, wts_;
};

// SigmaFactory, xtp/src/libxtp/factories/sigmafactory.cc:32-36
inline std::unique_ptr<Sigma_base> SigmaFactory_Create(const std::string& name, TCMatrix_gwbse& Mmn, const RPA& rpa) {
  if (name == "ppm") return std::make_unique<Sigma_PPM>(Mmn, rpa);
  if (name == "exact") return std::make_unique<Sigma_Exact>(Mmn, rpa);
  if (name == "cda") return std::make_unique<Sigma_CDA>(Mmn, rpa);
  throw std::runtime_error("SigmaFactory: unknown sigma_integrator '" + name + "'");
}

}  // namespace xtp
}  // namespace votca
