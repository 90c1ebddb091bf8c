// Unrestricted (UKS) twins of the GW-BSE classes - host mirror of xtp/include/votca/xtp/{rpa_uks,sigma_base_uks,
// gw_uks,bse_operator_uks,bse_uks}.h and their sources.
//
// The reference duplicates every restricted class with a spin index.  Here a spin channel IS a restricted pipeline
// on its own kernel-library context: the alpha and the beta Mmn tensor live in two gwbse_ctx objects on the same GPU
// (two streams, one address space), and every per-spin operation - exchange, plasmon-pole rotation, batched
// self-energy evaluation, the lock-step root search, the same-spin blocks of the BSE operator - is the restricted
// code path on that context, unchanged.  What is genuinely unrestricted is small and lives in this file:
//   * the dielectric matrix is the sum over both channels.  The restricted weights carry the closed-shell factor 2
//     (rpa.cc:115-127 against rpa_uks.cc:243-269, :351-363), so eps_uks = (eps_alpha + eps_beta) / 2, combined on the
//     device from the two contexts' own eps buffers;
//   * ONE plasmon-pole model, built from eps_uks, shared by both evaluators (gw_uks.cc:107-122);
//   * the RPA input energies of both channels are updated together (rpa_uks.cc:41-71);
//   * the BSE operator acts on [alpha (v c) | beta (v c)]; its same-spin blocks are gwbse_bse_matmul_dev per context,
//     the coupling between the channels (bse_operator_uks.cc:174-211) is one projection on the input channel's
//     context and one screened expansion on the output channel's (gwbse_bse_vc_project_dev / _expand_dev); the
//     cross-spin block of the full-BSE B operator (:136-172, :252-255) contracts both channels' tensors in one
//     two-leg GEMM chain (gwbse_bse_hd2_cross_dev).
// The two contexts take turns: every call on one is completed (stream synchronised) before the other gets work.  That
// costs nothing here (the channels' work is sequential in the reference too) and it is required: the TMA-staged GEMM
// was found to write wrong tiles when grids of another stream share the SMs with it.
// Scope: sigma_integrator = ppm (the default), exact and cda, BSE in the Tamm-Dancoff approximation and in full, one GPU.
#pragma once
#include "bse.h"
#include "gw.h"

namespace votca {
namespace xtp {

enum class Spin { Alpha = 0, Beta = 1 };

// threecenter.h:185-199 of the reference (TCMatrix_gwbse_spin)
struct TCMatrix_gwbse_spin {
  TCMatrix_gwbse_spin(const Device& dev_alpha, const Device& dev_beta) : alpha(dev_alpha), beta(dev_beta) {}
  TCMatrix_gwbse alpha, beta;
  TCMatrix_gwbse& operator[](Spin s) { return s == Spin::Alpha ? alpha : beta; }
  const TCMatrix_gwbse& operator[](Spin s) const { return s == Spin::Alpha ? alpha : beta; }
};

class RPA_UKS {
 public:
  RPA_UKS(Logger& log, const TCMatrix_gwbse_spin& Mmn) : Mmn_(Mmn), alpha_(log, Mmn.alpha), beta_(log, Mmn.beta) {}

  void configure(Index homo_alpha, Index homo_beta, Index rpamin, Index rpamax) {
    rpamin_ = rpamin;
    rpamax_ = rpamax;
    alpha_.configure(homo_alpha, rpamin, rpamax);
    beta_.configure(homo_beta, rpamin, rpamax);
  }
  void setRPAInputEnergies(const VectorXd& e_alpha, const VectorXd& e_beta) {
    alpha_.setRPAInputEnergies(e_alpha);
    beta_.setRPAInputEnergies(e_beta);
    InvalidateH2pCache();
  }
  // rpa_uks.cc:72-80
  void InvalidateH2pCache() const {
    screening_cached_ = false;
    screening_omegas_ = VectorXd();
    screening_modes_ = Device::Buffer();
  }
  const VectorXd& getRPAInputEnergiesAlpha() const { return alpha_.getRPAInputEnergies(); }
  const VectorXd& getRPAInputEnergiesBeta() const { return beta_.getRPAInputEnergies(); }
  RPA& channel(Spin s) { return s == Spin::Alpha ? alpha_ : beta_; }
  const RPA& channel(Spin s) const { return s == Spin::Alpha ? alpha_ : beta_; }

  // rpa_uks.cc:41-71 with ShiftUncorrectedEnergies / getMaxCorrection (:163-201): GW energies inside the qp window,
  // everything below / above it shifted by the largest correction seen among the occupied / virtual levels
  void UpdateRPAInputEnergies(const VectorXd& dft_alpha, const VectorXd& dft_beta, const VectorXd& gwa_alpha,
                              const VectorXd& gwa_beta, Index qpmin) {
    const Index rpatotal = rpamax_ - rpamin_ + 1;
    for (int s = 0; s < 2; ++s) {
      RPA& ch = s == 0 ? alpha_ : beta_;
      const VectorXd& dft = s == 0 ? dft_alpha : dft_beta;
      const VectorXd& gwa = s == 0 ? gwa_alpha : gwa_beta;
      VectorXd e = dft.segment(rpamin_, rpatotal);
      const Index gwsize = gwa.size(), qpmax = qpmin + gwsize - 1, lumo = ch.homo() + 1;
      for (Index i = 0; i < gwsize; ++i) e(qpmin - rpamin_ + i) = gwa(i);
      auto largest = [&](Index lo, Index hi) {
        double m = 0.0;
        for (Index i = lo; i <= hi; ++i) m = std::max(m, std::abs(e(i - rpamin_) - dft(i)));
        return m;
      };
      const double occ = largest(qpmin, ch.homo()), virt = largest(lumo, qpmax);
      for (Index i = 0; i < qpmin - rpamin_; ++i) e(i) -= occ;
      for (Index i = e.size() - (rpamax_ - qpmax); i < e.size(); ++i) e(i) += virt;
      ch.setRPAInputEnergies(e);
    }
    InvalidateH2pCache();
  }

  struct rpa_eigensolution {
    VectorXd omega;
    double ERPA_correlation = 0.0;
  };
  // rpa_uks.cc:369-438 with Calculate_H2p_AmB / _ApB (:440-556).  The (S_alpha + S_beta)^2 matrix is assembled on
  // the alpha context's device from four GEMMs over the two channels' tensors (gwbse_rpa_h2p_block; the reference
  // computes the lower triangles and the mixed block and mirrors them) and stays there; XpY_dev receives (X+Y).
  rpa_eigensolution Diagonalize_H2p(Device::Buffer* XpY_dev) const {
    const Device& da = Mmn_.alpha.device();
    const Device& db = Mmn_.beta.device();
    if (da.world() > 1) throw std::runtime_error("Diagonalize_H2p is single-GPU (S x S matrix not sharded)");
    const Index homo[2] = {alpha_.homo(), beta_.homo()};
    Index size[2];
    VectorXd AmB;
    {
      std::vector<double> amb;
      for (int s = 0; s < 2; ++s) {
        const VectorXd& e = s == 0 ? alpha_.getRPAInputEnergies() : beta_.getRPAInputEnergies();
        const Index n_occ = homo[s] + 1 - rpamin_, n_unocc = rpamax_ - homo[s];
        size[s] = n_occ * n_unocc;
        for (Index v = 0; v < n_occ; ++v)
          for (Index c = 0; c < n_unocc; ++c) amb.push_back(e(n_occ + c) - e(v));
      }
      AmB = VectorXd(static_cast<Index>(amb.size()));
      for (Index i = 0; i < AmB.size(); ++i) AmB(i) = amb[static_cast<size_t>(i)];
    }
    const Index S = size[0] + size[1];
    Device::Buffer C = da.alloc(static_cast<size_t>(S * S));
    const double kUKSApBPrefactor = 2.0;  // rpa_uks.cc:34
    db.sync();
    for (int sr = 0; sr < 2; ++sr)
      for (int sc = 0; sc < 2; ++sc) {
        const Device& dr = sr == 0 ? da : db;
        const Device& dc = sc == 0 ? da : db;
        const Index r0 = sr == 0 ? 0 : size[0], c0 = sc == 0 ? 0 : size[0];
        dr.check(gwbse_rpa_h2p_block(dr.ctx(), dc.ctx(), (int)homo[sr], (int)homo[sc], (int)rpamin_, (int)rpamax_,
                                     kUKSApBPrefactor, C.get() + r0 + c0 * S, (int)S));
        dr.sync();  // one context at a time (see the note at the top of this file)
      }
    Device::Buffer damb = da.upload(AmB);
    da.check(gwbse_axpy_dev(da.ctx(), 1, (int)S, 1.0, damb.get(), 1, C.get(), (int)(S + 1)));  // + diag(AmB)
    rpa_eigensolution sol;
    {
      Device::Buffer d = da.alloc(static_cast<size_t>(S));
      da.check(gwbse_dev_memset_zero(da.ctx(), d.get(), static_cast<size_t>(S)));
      da.check(gwbse_axpy_dev(da.ctx(), 1, (int)S, 1.0, C.get(), (int)(S + 1), d.get(), 1));
      MatrixXd diag = da.download(d.get(), S, 1);
      sol.ERPA_correlation = -0.25 * (diag.col(0).sum() + AmB.sum());
    }
    VectorXd sq(S);
    for (Index i = 0; i < S; ++i) sq(i) = std::sqrt(AmB(i));
    Device::Buffer dsq = da.upload(sq);
    da.check(gwbse_diag_scale_dev(da.ctx(), 'L', (int)S, (int)S, C.get(), (int)S, dsq.get(), C.get(), (int)S));
    da.check(gwbse_diag_scale_dev(da.ctx(), 'R', (int)S, (int)S, C.get(), (int)S, dsq.get(), C.get(), (int)S));
    VectorXd ev(S);
    da.check(gwbse_sym_eig_dev(da.ctx(), (int)S, C.get(), (int)S, ev.data()));
    double minCoeff = ev(0);
    for (Index i = 0; i < S; ++i) minCoeff = std::min(minCoeff, ev(i));
    if (minCoeff <= 0.0) throw std::runtime_error("Detected non-positive eigenvalue.");
    sol.omega = VectorXd(S);
    VectorXd osi(S);
    for (Index i = 0; i < S; ++i) {
      sol.omega(i) = std::sqrt(ev(i));
      osi(i) = 1.0 / std::sqrt(sol.omega(i));
    }
    sol.ERPA_correlation += 0.5 * sol.omega.sum();
    Device::Buffer dos = da.upload(osi);
    da.check(gwbse_diag_scale_dev(da.ctx(), 'L', (int)S, (int)S, C.get(), (int)S, dsq.get(), C.get(), (int)S));
    da.check(gwbse_diag_scale_dev(da.ctx(), 'R', (int)S, (int)S, C.get(), (int)S, dos.get(), C.get(), (int)S));
    da.sync();
    if (XpY_dev) *XpY_dev = std::move(C);
    return sol;
  }

  // rpa_uks.cc:82-161: the Coulomb-active screening modes sum_vc M[v][c,:] (X+Y)[vc, s] of both channels, cached
  // until the RPA input energies change; dark combinations (norm below 1e-10 of the largest) are dropped.  The
  // modes live on the alpha context's device (naux x nmodes, ld = naux); both channels' evaluators read them.
  void GetCachedScreeningModes(const VectorXd*& omegas, const double*& modes_dev, Index& nmodes) const {
    if (!screening_cached_) BuildCachedScreeningModes();
    omegas = &screening_omegas_;
    modes_dev = screening_modes_.get();
    nmodes = screening_omegas_.size();
  }
  double ERPA_correlation() const { return erpa_; }

 private:
  void BuildCachedScreeningModes() const {
    const Device& da = Mmn_.alpha.device();
    const Device& db = Mmn_.beta.device();
    const Index naux = Mmn_.alpha.auxsize();
    Device::Buffer XpY;
    const rpa_eigensolution sol = Diagonalize_H2p(&XpY);
    erpa_ = sol.ERPA_correlation;
    const Index S = sol.omega.size();
    const Index size_alpha = (alpha_.homo() + 1 - rpamin_) * (rpamax_ - alpha_.homo());
    Device::Buffer Z = da.alloc(static_cast<size_t>(naux * S));
    da.check(gwbse_sigma_exact_project(da.ctx(), XpY.get(), (int)S, (int)S, (int)alpha_.homo(), (int)rpamin_,
                                       (int)rpamax_, 0, Z.get(), (int)naux));
    da.sync();
    db.check(gwbse_sigma_exact_project(db.ctx(), XpY.get() + size_alpha, (int)S, (int)S, (int)beta_.homo(),
                                       (int)rpamin_, (int)rpamax_, 1, Z.get(), (int)naux));
    db.sync();
    MatrixXd modes = da.download(Z.get(), naux, S);
    std::vector<double> norm(static_cast<size_t>(S));
    double max_norm = 0.0;
    for (Index s = 0; s < S; ++s) {
      double n2 = 0.0;
      for (Index x = 0; x < naux; ++x) n2 += modes(x, s) * modes(x, s);
      norm[static_cast<size_t>(s)] = std::sqrt(n2);
      max_norm = std::max(max_norm, norm[static_cast<size_t>(s)]);
    }
    const double tol = 1e-10 * std::max(1.0, max_norm);
    std::vector<Index> keep;
    for (Index s = 0; s < S; ++s)
      if (norm[static_cast<size_t>(s)] > tol) keep.push_back(s);
    screening_omegas_ = VectorXd(static_cast<Index>(keep.size()));
    if (static_cast<Index>(keep.size()) == S) {
      for (Index s = 0; s < S; ++s) screening_omegas_(s) = sol.omega(s);
      screening_modes_ = std::move(Z);
    } else {
      MatrixXd active(naux, static_cast<Index>(keep.size()));
      for (size_t j = 0; j < keep.size(); ++j) {
        screening_omegas_(static_cast<Index>(j)) = sol.omega(keep[j]);
        for (Index x = 0; x < naux; ++x) active(x, static_cast<Index>(j)) = modes(x, keep[j]);
      }
      screening_modes_ = da.upload(active);
    }
    da.sync();
    screening_cached_ = true;
  }

 public:

  // eps_uks on the alpha context's device; valid until the next call
  double* calculate_epsilon_i_dev(double frequency) const {
    return combine(alpha_.calculate_epsilon_i_dev(frequency), [&] { return beta_.calculate_epsilon_i_dev(frequency); });
  }
  double* calculate_epsilon_r_dev(double frequency) const {
    return combine(alpha_.calculate_epsilon_r_dev(frequency), [&] { return beta_.calculate_epsilon_r_dev(frequency); });
  }
  double* calculate_epsilon_r_dev(std::complex<double> f) const {
    return combine(alpha_.calculate_epsilon_r_dev(f), [&] { return beta_.calculate_epsilon_r_dev(f); });
  }
  MatrixXd calculate_epsilon_i(double frequency) const { return fetch(calculate_epsilon_i_dev(frequency)); }
  MatrixXd calculate_epsilon_r(double frequency) const { return fetch(calculate_epsilon_r_dev(frequency)); }
  MatrixXd calculate_epsilon_r(std::complex<double> f) const { return fetch(calculate_epsilon_r_dev(f)); }

 private:
  template <class BetaCall>
  double* combine(double* eps_alpha, BetaCall beta_call) const {
    const Device& da = Mmn_.alpha.device();
    const Device& db = Mmn_.beta.device();
    const Index n = Mmn_.alpha.auxsize();
    double* eps_beta = beta_call();
    db.sync();  // the beta context's stream has produced its matrix before the alpha context reads it
    if (eps_.size() < static_cast<size_t>(n * n)) eps_ = da.alloc(static_cast<size_t>(n * n));
    da.check(gwbse_dev_memset_zero(da.ctx(), eps_.get(), static_cast<size_t>(n * n)));
    da.check(gwbse_axpy_dev(da.ctx(), (int)n, (int)n, 0.5, eps_alpha, (int)n, eps_.get(), (int)n));
    da.check(gwbse_axpy_dev(da.ctx(), (int)n, (int)n, 0.5, eps_beta, (int)n, eps_.get(), (int)n));
    da.sync();
    return eps_.get();
  }
  MatrixXd fetch(const double* p) const {
    const Index n = Mmn_.alpha.auxsize();
    return Mmn_.alpha.device().download(p, n, n);
  }

  const TCMatrix_gwbse_spin& Mmn_;
  RPA alpha_, beta_;
  Index rpamin_ = 0, rpamax_ = 0;
  mutable Device::Buffer eps_;
  mutable bool screening_cached_ = false;
  mutable VectorXd screening_omegas_;
  mutable Device::Buffer screening_modes_;
  mutable double erpa_ = 0.0;
};

// sigma_ppm_uks.h / .cc: the restricted evaluator of one channel with the plasmon-pole model handed in
class Sigma_PPM_UKS final : public Sigma_PPM {
 public:
  Sigma_PPM_UKS(TCMatrix_gwbse& Mmn, const RPA& rpa_channel) : Sigma_PPM(Mmn, rpa_channel) {}
  void SetSharedPPM(const PPM& ppm) { shared_ = &ppm; }
  void PrepareScreening() override {  // sigma_ppm_uks.cc:30-37
    if (!shared_)
      throw std::runtime_error("Sigma_PPM_UKS: shared PPM parameters were not set before PrepareScreening().");
    InstallPPM(*shared_);
  }

 private:
  const PPM* shared_ = nullptr;
};

// sigma_exact_uks.h / .cc: residues of the channel's Mmn on the screening modes both channels share; the pole sums
// carry no closed-shell factor (sigma_exact_uks.cc:82, :107, :140 against sigma_exact.cc:57, :76, :106)
class Sigma_Exact_UKS final : public Sigma_base {
 public:
  Sigma_Exact_UKS(TCMatrix_gwbse& Mmn, const RPA& rpa_channel, const RPA_UKS& rpa_uks)
      : Sigma_base(Mmn, rpa_channel), rpa_uks_(rpa_uks) {}
  void PrepareScreening() override {  // sigma_exact_uks.cc:37-61
    const VectorXd* omegas = nullptr;
    const double* modes = nullptr;
    Index nmodes = 0;
    rpa_uks_.GetCachedScreeningModes(omegas, modes, nmodes);
    if (nmodes == 0) throw std::runtime_error("Sigma_Exact_UKS: no Coulomb-active screening mode");
    const Device& dev = Mmn_.device();
    dev.check(gwbse_sigma_exact_prepare_modes(dev.ctx(), omegas->data(), (int)nmodes, modes, (int)Mmn_.auxsize(),
                                              rpa_.getRPAInputEnergies().data(), (int)opt_.homo, (int)opt_.rpamin,
                                              (int)opt_.rpamax, (int)opt_.qpmin, (int)opt_.qpmax, opt_.eta, 1.0, 0.5));
    dev.sync();
  }
  void EvalBatch(const std::vector<int>& levels, const std::vector<double>& freqs, std::vector<double>& sigma,
                 std::vector<double>* dsigma) const override {
    const Device& dev = Mmn_.device();
    sigma.resize(levels.size());
    if (dsigma) dsigma->resize(levels.size());
    dev.check(gwbse_sigma_update_energies(dev.ctx(), 1, rpa_.getRPAInputEnergies().data()));
    dev.check(gwbse_sigma_exact_eval(dev.ctx(), (int)levels.size(), levels.data(), freqs.data(), sigma.data(),
                                     dsigma ? dsigma->data() : nullptr));
  }
  void EvalGroups(const std::vector<int>& levels, const std::vector<int>& gptr, const std::vector<double>& freqs,
                  std::vector<double>& sigma, std::vector<double>* dsigma) const override {
    const Device& dev = Mmn_.device();
    sigma.resize(freqs.size());
    if (dsigma) dsigma->resize(freqs.size());
    dev.check(gwbse_sigma_update_energies(dev.ctx(), 1, rpa_.getRPAInputEnergies().data()));
    dev.check(gwbse_sigma_eval_groups(dev.ctx(), 1, (int)levels.size(), levels.data(), gptr.data(), freqs.data(),
                                      sigma.data(), dsigma ? dsigma->data() : nullptr));
  }
  MatrixXd CalcCorrelationOffDiag(const VectorXd& frequencies) const override {
    MatrixXd out(qptotal_, qptotal_);
    const Device& dev = Mmn_.device();
    dev.check(gwbse_sigma_update_energies(dev.ctx(), 1, rpa_.getRPAInputEnergies().data()));
    dev.check(gwbse_sigma_exact_offdiag(dev.ctx(), (int)qptotal_, frequencies.data(), out.data(), (int)qptotal_));
    return out;
  }

 private:
  const RPA_UKS& rpa_uks_;
};

// sigma_cda_uks.h / .cc: the restricted contour-deformation formulas with the dielectric matrix of both channels
// (RPA_UKS::calculate_epsilon_*), the channel's own energies, homo and Mmn
class Sigma_CDA_UKS final : public Sigma_CDA {
 public:
  Sigma_CDA_UKS(TCMatrix_gwbse& Mmn, const RPA& rpa_channel, const RPA& rpa_other_channel, const Device& other_device)
      : Sigma_CDA(Mmn, rpa_channel), other_(rpa_other_channel), other_dev_(other_device) {}

 protected:
  void BindDielectricSource() const override {
    const Device& dev = Mmn_.device();
    other_dev_.sync();
    dev.check(gwbse_sigma_cda_set_partner(dev.ctx(), other_dev_.ctx(), (int)other_.homo(),
                                          other_.getRPAInputEnergies().data()));
  }

 private:
  const RPA& other_;
  const Device& other_dev_;
};

class GW_UKS {
 public:
  struct options : GW::options {
    Index homo_alpha = 0, homo_beta = 0;
  };

  GW_UKS(Logger& log, TCMatrix_gwbse_spin& Mmn, const MatrixXd& vxc_alpha, const MatrixXd& vxc_beta,
         const VectorXd& dft_energies_alpha, const VectorXd& dft_energies_beta)
      : log_(log), Mmn_(Mmn), dft_{&dft_energies_alpha, &dft_energies_beta}, rpa_(log, Mmn),
        gw_{std::make_unique<GW>(log, Mmn.alpha, vxc_alpha, dft_energies_alpha),
            std::make_unique<GW>(log, Mmn.beta, vxc_beta, dft_energies_beta)} {}

  // gw_uks.cc:38-123
  void configure(const options& opt) {
    opt_ = opt;
    if (opt_.sigma_integration != "ppm" && opt_.sigma_integration != "exact" && opt_.sigma_integration != "cda")
      throw std::runtime_error("GW_UKS: unknown sigma_integrator '" + opt_.sigma_integration + "'");
    if (opt_.do_qsgw) throw std::runtime_error("GW_UKS: QSGW is defined for restricted references only");
    rpa_.configure(opt_.homo_alpha, opt_.homo_beta, opt_.rpamin, opt_.rpamax);
    for (int s = 0; s < 2; ++s) {
      GW::options o = opt_;
      o.homo = s == 0 ? opt_.homo_alpha : opt_.homo_beta;
      TCMatrix_gwbse& M = s == 0 ? Mmn_.alpha : Mmn_.beta;
      if (opt_.sigma_integration == "exact") {  // sigmafactory_uks.cc
        gw_[s]->configure(o, std::make_unique<Sigma_Exact_UKS>(M, gw_[s]->rpa_, rpa_));
      } else if (opt_.sigma_integration == "cda") {
        TCMatrix_gwbse& Mo = s == 0 ? Mmn_.beta : Mmn_.alpha;
        gw_[s]->configure(o, std::make_unique<Sigma_CDA_UKS>(M, gw_[s]->rpa_, gw_[1 - s]->rpa_, Mo.device()));
      } else {
        auto sigma = std::make_unique<Sigma_PPM_UKS>(M, gw_[s]->rpa_);
        sigma->SetSharedPPM(ppm_);
        gw_[s]->configure(o, std::move(sigma));
      }
      log_(std::string(" UKS ") + (s == 0 ? "alpha" : "beta") + " effective ranges: HOMO=" + std::to_string(o.homo) +
           " LUMO=" + std::to_string(o.homo + 1) + " | RPA [" + std::to_string(o.rpamin) + ":" +
           std::to_string(o.rpamax) + "] | GW [" + std::to_string(o.qpmin) + ":" + std::to_string(o.qpmax) + "]");
    }
    qptotal_ = opt_.qpmax - opt_.qpmin + 1;
  }

  // gw_uks.cc:188-295
  void CalculateGWPerturbation() {
    VectorXd freq[2];
    VectorXd shifted[2];
    for (int s = 0; s < 2; ++s) {
      GW& g = *gw_[s];
      g.Sigma_x_ = (1 - opt_.ScaHFX) * g.sigma_->CalcExchangeMatrix();
      shifted[s] = *dft_[s];
      for (Index i = g.opt_.homo + 1; i < shifted[s].size(); ++i) shifted[s](i) += opt_.shift;
      freq[s] = shifted[s].segment(opt_.qpmin, qptotal_);
    }
    log_(" Calculated spin-resolved Hartree exchange contribution");
    log_(" Scissor shifting alpha/beta DFT energies by: " + std::to_string(opt_.shift) + " Hrt");
    const Index rpatotal = opt_.rpamax - opt_.rpamin + 1;
    rpa_energies(shifted[0].segment(opt_.rpamin, rpatotal), shifted[1].segment(opt_.rpamin, rpatotal));
    Anderson mixing[2];
    for (Anderson& m : mixing) m.Configure(opt_.gw_mixing_order, opt_.gw_mixing_alpha);
    for (Index i_gw = 0; i_gw < opt_.gw_sc_max_iterations; ++i_gw) {
      iterations_ = i_gw + 1;
      for (auto& g : gw_) g->gw_sc_iteration_ = i_gw;
      if (i_gw % opt_.reset_3c == 0 && i_gw != 0) {
        Mmn_.alpha.Rebuild();
        Mmn_.alpha.device().sync();
        Mmn_.beta.Rebuild();
        Mmn_.beta.device().sync();
        log_(" Rebuilding alpha/beta 3c integrals");
      }
      // one plasmon-pole model from the spin-summed dielectric matrix, installed in both channels
      if (opt_.sigma_integration == "ppm") {  // gw_uks.cc:228-235: the other evaluators build their own screening
        ppm_.PPM_construct_parameters(rpa_, Mmn_.alpha);
        Mmn_.alpha.device().sync();
      }
      // one channel after the other, never both contexts' grids on the GPU at once (see the note on streams below)
      for (int s = 0; s < 2; ++s) {
        gw_[s]->sigma_->PrepareScreening();
        (s == 0 ? Mmn_.alpha : Mmn_.beta).device().sync();
      }
      ppm_.FreeMatrix();
      log_(" Calculated unrestricted screening via RPA");
      if (opt_.gw_mixing_order > 0 && i_gw > 0)
        for (int s = 0; s < 2; ++s) mixing[s].UpdateInput(freq[s]);
      for (int s = 0; s < 2; ++s) freq[s] = gw_[s]->SolveQP(freq[s]);
      if (opt_.gw_sc_max_iterations > 1) {
        const VectorXd old[2] = {rpa_.getRPAInputEnergiesAlpha(), rpa_.getRPAInputEnergiesBeta()};
        if (opt_.gw_mixing_order > 0 && i_gw > 0)
          for (int s = 0; s < 2; ++s) {
            mixing[s].UpdateOutput(freq[s]);
            freq[s] = mixing[s].MixHistory();
          }
        rpa_.UpdateRPAInputEnergies(*dft_[0], *dft_[1], freq[0], freq[1], opt_.qpmin);
        sync_channel_energies();
        log_(" GW_Iteration:" + std::to_string(i_gw));
        const bool conv_a = gw_[0]->Converged(rpa_.getRPAInputEnergiesAlpha(), old[0], opt_.gw_sc_limit);
        const bool conv_b = gw_[1]->Converged(rpa_.getRPAInputEnergiesBeta(), old[1], opt_.gw_sc_limit);
        if (conv_a && conv_b) {
          log_(" Converged after " + std::to_string(i_gw + 1) + " unrestricted GW iterations.");
          break;
        } else if (i_gw == opt_.gw_sc_max_iterations - 1) {
          log_(" WARNING! UKS GW-self-consistency cycle not converged after " +
               std::to_string(opt_.gw_sc_max_iterations) + " iterations.");
          break;
        }
      }
    }
    for (int s = 0; s < 2; ++s) {
      const VectorXd diag = gw_[s]->sigma_->CalcCorrelationDiag(freq[s]);
      for (Index i = 0; i < qptotal_; ++i) gw_[s]->Sigma_c_(i, i) = diag(i);
    }
  }

  VectorXd getGWAResultsAlpha() const { return gw_[0]->getGWAResults(); }
  VectorXd getGWAResultsBeta() const { return gw_[1]->getGWAResults(); }
  const VectorXd& RPAInputEnergiesAlpha() const { return rpa_.getRPAInputEnergiesAlpha(); }
  const VectorXd& RPAInputEnergiesBeta() const { return rpa_.getRPAInputEnergiesBeta(); }
  // gw_uks.cc:758-790
  void CalculateHQP() {
    for (auto& g : gw_) g->CalculateHQP();
  }
  MatrixXd getHQPAlpha() const { return gw_[0]->getHQP(); }
  MatrixXd getHQPBeta() const { return gw_[1]->getHQP(); }
  std::pair<VectorXd, MatrixXd> DiagonalizeQPHamiltonianAlpha() const { return gw_[0]->DiagonalizeQPHamiltonian(); }
  std::pair<VectorXd, MatrixXd> DiagonalizeQPHamiltonianBeta() const { return gw_[1]->DiagonalizeQPHamiltonian(); }
  const MatrixXd& Sigma_x(Spin s) const { return gw_[(int)s]->Sigma_x(); }
  const MatrixXd& Sigma_c(Spin s) const { return gw_[(int)s]->Sigma_c(); }
  Index iterations() const { return iterations_; }

 private:
  // the two channels' RPA objects inside the GW halves ARE the channel energies of rpa_: keep both views equal
  void rpa_energies(const VectorXd& ea, const VectorXd& eb) {
    rpa_.setRPAInputEnergies(ea, eb);
    sync_channel_energies();
  }
  void sync_channel_energies() {
    gw_[0]->rpa_.setRPAInputEnergies(rpa_.getRPAInputEnergiesAlpha());
    gw_[1]->rpa_.setRPAInputEnergies(rpa_.getRPAInputEnergiesBeta());
  }

  Logger& log_;
  TCMatrix_gwbse_spin& Mmn_;
  const VectorXd* dft_[2];
  RPA_UKS rpa_;
  PPM ppm_;
  std::unique_ptr<GW> gw_[2];
  options opt_;
  Index qptotal_ = 0, iterations_ = 0;
};

struct BSEOperatorUKS_Options {
  Index homo_alpha, homo_beta, rpamin, qpmin, vmin, cmax;
};

// bse_operator_uks.h / .cc on the combined space [alpha | beta], each channel ordered ctotal * v + c
template <Index cqp, Index cx, Index cd, Index cd2>
class BSE_OPERATOR_UKS final : public MatrixFreeOperator {
 public:
  BSE_OPERATOR_UKS(const VectorXd& Hd_operator, const TCMatrix_gwbse_spin& Mmn, const MatrixXd& Hqp_alpha,
                   const MatrixXd& Hqp_beta)
      : epsilon_0_inv_(Hd_operator), Mmn_(Mmn), Hqp_{&Hqp_alpha, &Hqp_beta} {
    static_assert(!(cd2 != 0 && cd != 0), "Hamiltonian cannot contain Hd and Hd2 at the same time");
  }
  // bse_operator_uks.cc:26-45
  void configure(BSEOperatorUKS_Options opt) {
    opt_ = opt;
    Index offset = 0;
    for (int s = 0; s < 2; ++s) {
      const Index homo = s == 0 ? opt.homo_alpha : opt.homo_beta;
      blk_[s].homo = homo;
      blk_[s].vtotal = homo - opt.vmin + 1;
      blk_[s].ctotal = opt.cmax - homo;
      blk_[s].size = blk_[s].vtotal * blk_[s].ctotal;
      blk_[s].offset = offset;
      offset += blk_[s].size;
    }
    this->set_size(offset);
  }
  Index alpha_size() const { return blk_[0].size; }
  Index beta_size() const { return blk_[1].size; }

  // bse_operator_uks.cc:277-352: the diagonal has no contribution from the coupling between the channels
  VectorXd diagonal() const override {
    VectorXd d(this->size());
    for (int s = 0; s < 2; ++s) {
      bind(s);
      const Device& dev = channel(s).device();
      dev.check(gwbse_bse_diagonal(dev.ctx(), (int)cqp, (int)cx, (int)cd, (int)cd2, d.data() + blk_[s].offset));
    }
    return d;
  }

  void apply_dev(const double* X_dev, Index ldx, Index k, double* Y_dev, Index ldy) const override {
    const Index naux = Mmn_.alpha.auxsize();
    for (int s = 0; s < 2; ++s) {  // same-spin blocks: the restricted operator of the channel (:228-246)
      bind(s);
      const Device& dev = channel(s).device();
      dev.check(gwbse_bse_matmul_dev(dev.ctx(), (int)cqp, (int)cx, (int)cd, (int)cd2, (int)k, X_dev + blk_[s].offset,
                                     (int)ldx, Y_dev + blk_[s].offset, (int)ldy));
      dev.sync();
    }
    if (cd != 0) {  // coupling between the channels, transition-density form, screened (:174-211, :248-252)
      if (W_.size() < static_cast<size_t>(naux * k)) W_ = channel(0).device().alloc(static_cast<size_t>(naux * k));
      for (int in = 0; in < 2; ++in) {
        const int out = 1 - in;
        const Device& din = channel(in).device();
        const Device& dout = channel(out).device();
        din.check(gwbse_bse_vc_project_dev(din.ctx(), (int)k, X_dev + blk_[in].offset, (int)ldx, W_.get()));
        din.sync();
        dout.check(gwbse_bse_vc_expand_dev(dout.ctx(), -double(cd), 1, (int)k, W_.get(), Y_dev + blk_[out].offset,
                                           (int)ldy));
        dout.sync();
      }
    }
    if (cd2 != 0) {  // cross-spin block of the full-BSE B operator: both channels' tensors in one contraction (:252-255)
      for (int in = 0; in < 2; ++in) {
        const int out = 1 - in;
        const Device& din = channel(in).device();
        const Device& dout = channel(out).device();
        dout.check(gwbse_bse_hd2_cross_dev(dout.ctx(), din.ctx(), (int)blk_[in].homo, -double(cd2), (int)k,
                                           X_dev + blk_[in].offset, (int)ldx, Y_dev + blk_[out].offset, (int)ldy));
        dout.sync();
      }
    }
  }

  MatrixXd matmul(const MatrixXd& input) const override {
    if (input.rows() != this->size()) throw std::runtime_error("Shape mismatch in BSE_OPERATOR_UKS::matmul");
    const Device& dev = device();
    Device::Buffer X = dev.upload(input);
    Device::Buffer Y = dev.alloc(static_cast<size_t>(std::max<Index>(input.size(), 1)));
    apply_dev(X.get(), input.rows(), input.cols(), Y.get(), input.rows());
    return dev.download(Y.get(), input.rows(), input.cols());
  }
  const Device& device() const override { return Mmn_.alpha.device(); }

 private:
  struct Block {
    Index homo = 0, vtotal = 0, ctotal = 0, size = 0, offset = 0;
  };
  const TCMatrix_gwbse& channel(int s) const { return s == 0 ? Mmn_.alpha : Mmn_.beta; }
  void bind(int s) const {
    const Device& dev = channel(s).device();
    const MatrixXd& H = *Hqp_[s];
    if (H.rows() < blk_[s].vtotal + blk_[s].ctotal) throw std::runtime_error("Hqp is smaller than the BSE window");
    dev.check(gwbse_bse_configure(dev.ctx(), (int)blk_[s].homo, (int)opt_.rpamin, (int)opt_.vmin, (int)opt_.cmax,
                                  epsilon_0_inv_.data(), H.data(), (int)H.rows()));
  }
  BSEOperatorUKS_Options opt_;
  Block blk_[2];
  const VectorXd& epsilon_0_inv_;
  const TCMatrix_gwbse_spin& Mmn_;
  const MatrixXd* Hqp_[2];
  mutable Device::Buffer W_;
};

typedef BSE_OPERATOR_UKS<1, 1, 1, 0> ExcitonUKSOperator_TDA;
typedef BSE_OPERATOR_UKS<0, 1, 0, 1> ExcitonUKSOperator_BTDA_B;
typedef BSE_OPERATOR_UKS<0, 0, 0, 1> Hd2UKSOperator;
typedef BSE_OPERATOR_UKS<1, 0, 0, 0> HqpUKSOperator;
typedef BSE_OPERATOR_UKS<0, 1, 0, 0> HxUKSOperator;
typedef BSE_OPERATOR_UKS<0, 0, 1, 0> HdUKSOperator;
typedef BSE_OPERATOR_UKS<0, 0, 0, 1> Hd2UKSOperator;

class BSE_UKS {
 public:
  using options = BSE::options;
  BSE_UKS(Logger& log, TCMatrix_gwbse_spin& Mmn) : log_(log), Mmn_(Mmn) {}

  // bse_uks.cc:51-90 with the screening of gwbse.cc:1157-1175 (spin-summed eps(0), eigen-decomposition, both Mmn
  // rotated into its eigenbasis); everything stays on the device
  void configure(const options& opt, Index homo_alpha, Index homo_beta, const VectorXd& RPAInputEnergiesAlpha,
                 const VectorXd& RPAInputEnergiesBeta, const MatrixXd& Hqp_alpha_in, const MatrixXd& Hqp_beta_in) {
    opt_ = opt;
    homo_[0] = homo_alpha;
    homo_[1] = homo_beta;
    for (int s = 0; s < 2; ++s) {
      // AdjustHqpSize (bse_uks.cc:92-133) is the restricted routine with the channel's homo
      BSE::options o = opt_;
      o.homo = homo_[s];
      MatrixXd H = BSE::AdjustHqpSizeFor(o, s == 0 ? Hqp_alpha_in : Hqp_beta_in,
                                         s == 0 ? RPAInputEnergiesAlpha : RPAInputEnergiesBeta);
      Hqp_[s] = opt_.use_Hqp_offdiag ? H : asDiagonal(H.diagonal());
    }
    SetupDirectInteractionOperator(RPAInputEnergiesAlpha, RPAInputEnergiesBeta, 0.0);
  }

  // bse_uks.cc:135-157: spin-summed eps(energy), its eigen-decomposition, both tensors rotated into its eigenbasis.
  // The reference rotates a pristine copy of the tensors each time; here the resident tensors are rotated again -
  // eps computed from rotated tensors has the rotated eigenvectors, so the composed rotation and with it the
  // operator are the same (as in the restricted BSE, bse.h).
  void SetupDirectInteractionOperator(const VectorXd& RPAInputEnergiesAlpha, const VectorXd& RPAInputEnergiesBeta,
                                      double energy) {
    RPA_UKS rpa(log_, Mmn_);
    rpa.configure(homo_[0], homo_[1], opt_.rpamin, opt_.rpamax);
    rpa.setRPAInputEnergies(RPAInputEnergiesAlpha, RPAInputEnergiesBeta);
    const Device& da = Mmn_.alpha.device();
    const Index n = Mmn_.alpha.auxsize();
    double* eps = rpa.calculate_epsilon_r_dev(energy);
    Device::Buffer U = da.alloc(static_cast<size_t>(n * n));
    da.check(gwbse_d2d(da.ctx(), U.get(), eps, static_cast<size_t>(n * n)));
    VectorXd ev(n);
    da.check(gwbse_sym_eig_dev(da.ctx(), (int)n, U.get(), (int)n, ev.data()));
    da.sync();
    Mmn_.alpha.MultiplyRightWithAuxMatrix_dev(U.get(), n);
    Mmn_.alpha.device().sync();
    Mmn_.beta.MultiplyRightWithAuxMatrix_dev(U.get(), n);
    Mmn_.beta.device().sync();
    epsilon_0_inv_ = VectorXd::Zero(n);
    for (Index i = 0; i < n; ++i)
      if (ev(i) > 1e-8) epsilon_0_inv_(i) = 1.0 / ev(i);
  }

  struct ExpectationValues {
    VectorXd direct_term, cross_term;
  };
  // bse_uks.cc:171-216 (state < 0: every state)
  template <typename OPERATOR>
  ExpectationValues ExpectationValue_Operator(const EigenSystem& es, const OPERATOR& H, Index state = -1) const {
    auto cols = [&](const MatrixXd& M) {
      if (state < 0) return M;
      MatrixXd c(M.rows(), 1);
      for (Index i = 0; i < M.rows(); ++i) c(i, 0) = M(i, state);
      return c;
    };
    auto exp_value = [](const MatrixXd& a, const MatrixXd& b) {
      VectorXd r(a.cols(), 0.0);
      for (Index j = 0; j < a.cols(); ++j)
        for (Index i = 0; i < a.rows(); ++i) r(j) += a(i, j) * b(i, j);
      return r;
    };
    ExpectationValues ev;
    const MatrixXd X = cols(es.eigenvectors);
    const MatrixXd HX = H.matmul(X);
    ev.direct_term = exp_value(X, HX);
    if (!opt_.useTDA) {
      const MatrixXd Y = cols(es.eigenvectors2);
      const VectorXd yy = exp_value(Y, H.matmul(Y)), yx = exp_value(Y, HX);
      ev.cross_term = VectorXd(yx.size());
      for (Index j = 0; j < yx.size(); ++j) {
        ev.direct_term(j) += yy(j);
        ev.cross_term(j) = 2.0 * yx(j);
      }
    }
    return ev;
  }

  // bse_uks.cc:640-702: E_dyn = E + <Hd(0)> - <Hd(E_dyn)>, iterated per excitation (full BSE: plus the Hd2 cross terms)
  VectorXd Perturbative_DynamicalScreening(const EigenSystem& es, const VectorXd& RPAInputEnergiesAlpha,
                                           const VectorXd& RPAInputEnergiesBeta) {
    auto contribution = [&](Index state) {
      HdUKSOperator Hd(epsilon_0_inv_, Mmn_, Hqp_[0], Hqp_[1]);
      Hd.configure({homo_[0], homo_[1], opt_.rpamin, opt_.qpmin, opt_.vmin, opt_.cmax});
      VectorXd c = ExpectationValue_Operator(es, Hd, state).direct_term;
      if (!opt_.useTDA) {
        Hd2UKSOperator Hd2(epsilon_0_inv_, Mmn_, Hqp_[0], Hqp_[1]);
        Hd2.configure({homo_[0], homo_[1], opt_.rpamin, opt_.qpmin, opt_.vmin, opt_.cmax});
        const VectorXd x = ExpectationValue_Operator(es, Hd2, state).cross_term;
        for (Index j = 0; j < c.size(); ++j) c(j) += x(j);
      }
      return c;
    };
    SetupDirectInteractionOperator(RPAInputEnergiesAlpha, RPAInputEnergiesBeta, 0.0);
    const VectorXd Hd_static = contribution(-1);
    VectorXd dyn = es.eigenvalues;
    for (Index i = 0; i < dyn.size(); ++i) {
      log_("Dynamical Screening UKS BSE, Excitation " + std::to_string(i) + " static " +
           std::to_string(es.eigenvalues(i)));
      for (Index iter = 0; iter < opt_.max_dyn_iter; ++iter) {
        const double old_energy = dyn(i);
        SetupDirectInteractionOperator(RPAInputEnergiesAlpha, RPAInputEnergiesBeta, old_energy);
        const VectorXd Hd_dyn = contribution(i);
        dyn(i) = es.eigenvalues(i) + Hd_static(i) - Hd_dyn(0);
        log_("Dynamical Screening UKS BSE, excitation " + std::to_string(i) + " iteration " + std::to_string(iter) +
             " dynamic " + std::to_string(dyn(i)));
        if (std::abs(dyn(i) - old_energy) < opt_.dyn_tolerance) break;
      }
    }
    return dyn;
  }

  ExcitonUKSOperator_TDA getExcitonOperator_TDA() const {
    ExcitonUKSOperator_TDA H(epsilon_0_inv_, Mmn_, Hqp_[0], Hqp_[1]);
    H.configure({homo_[0], homo_[1], opt_.rpamin, opt_.qpmin, opt_.vmin, opt_.cmax});
    return H;
  }

  ExcitonUKSOperator_BTDA_B getExcitonOperator_BTDA_B() const {
    ExcitonUKSOperator_BTDA_B H(epsilon_0_inv_, Mmn_, Hqp_[0], Hqp_[1]);
    H.configure({homo_[0], homo_[1], opt_.rpamin, opt_.qpmin, opt_.vmin, opt_.cmax});
    return H;
  }

  // bse_uks.cc:442-451
  EigenSystem Solve_excitons_uks() const { return opt_.useTDA ? Solve_excitons_uks_TDA() : Solve_excitons_uks_BTDA(); }

  // bse_uks.cc:218-235, 453-479
  EigenSystem Solve_excitons_uks_TDA() const {
    ExcitonUKSOperator_TDA H = getExcitonOperator_TDA();
    log_(" Setup combined UKS TDA Hamiltonian ");
    EigenSystem result;
    DavidsonSolver DS(log_);
    configureDavidson(DS);
    DS.solve(H, opt_.nmax);
    result.eigenvalues = DS.eigenvalues();
    result.eigenvectors = DS.eigenvectors();
    result.success = DS.success();
    iterations_ = DS.num_iterations();
    return result;
  }

  // bse_uks.cc:291-440: [A B; -B -A] with A = <1,1,1,0>, B = <0,1,0,1>; dense for up to 128 excitations (the
  // reference's choice for small systems), Davidson on the Hamiltonian operator otherwise
  EigenSystem Solve_excitons_uks_BTDA() const {
    ExcitonUKSOperator_TDA A = getExcitonOperator_TDA();
    ExcitonUKSOperator_BTDA_B B = getExcitonOperator_BTDA_B();
    log_(" Setup combined UKS full BSE exciton hamiltonian ");
    const Index n = A.rows();
    EigenSystem result;
    if (n <= 128) {
      log_(" Using dense full UKS-BSE solve for small system (dim=" + std::to_string(n) + ")");
      const MatrixXd I = MatrixXd::Identity(n, n);
      const MatrixXd Ad = A.matmul(I), Bd = B.matmul(I);
      MatrixXd H(2 * n, 2 * n), One = MatrixXd::Identity(2 * n, 2 * n);
      for (Index j = 0; j < n; ++j)
        for (Index i = 0; i < n; ++i) {
          H(i, j) = Ad(i, j);
          H(i, n + j) = Bd(i, j);
          H(n + i, j) = -Bd(i, j);
          H(n + i, n + j) = -Ad(i, j);
        }
      VectorXd wr(2 * n), wi(2 * n);
      MatrixXd VR(2 * n, 2 * n);
      const Device& dev = Mmn_.alpha.device();
      try {
        dev.check(gwbse_gen_eig_host(dev.ctx(), (int)(2 * n), H.data(), One.data(), wr.data(), wi.data(), VR.data()));
      } catch (const std::exception&) {
        throw std::runtime_error("Dense full UKS-BSE diagonalization failed.");
      }
      std::vector<std::pair<double, Index>> roots;  // positive real roots, ascending
      for (Index i = 0; i < 2 * n; ++i)
        if (std::abs(wi(i)) < 1e-8 && wr(i) > 0.0) roots.emplace_back(wr(i), i);
      std::stable_sort(roots.begin(), roots.end(), [](const auto& a, const auto& b) { return a.first < b.first; });
      if (roots.empty())
        throw std::runtime_error("Dense full UKS-BSE diagonalization produced no positive real roots.");
      const Index nroots = std::min<Index>(opt_.nmax, (Index)roots.size());
      result.eigenvalues = VectorXd(nroots);
      result.eigenvectors = MatrixXd(n, nroots);
      result.eigenvectors2 = MatrixXd(n, nroots);
      for (Index r = 0; r < nroots; ++r) {
        const Index col = roots[r].second;
        Index imax = 0;  // phase convention: the largest |X_k| positive
        for (Index i = 1; i < n; ++i)
          if (std::abs(VR(i, col)) > std::abs(VR(imax, col))) imax = i;
        const double sign = VR(imax, col) < 0.0 ? -1.0 : 1.0;
        double nx = 0.0, ny = 0.0;
        for (Index i = 0; i < n; ++i) {
          nx += VR(i, col) * VR(i, col);
          ny += VR(n + i, col) * VR(n + i, col);
        }
        const double norm = nx - ny;
        if (std::abs(norm) < 1e-12)
          throw std::runtime_error("Dense full UKS-BSE eigenvector has near-zero (X^2-Y^2) norm.");
        const double f = sign / std::sqrt(std::abs(norm));
        result.eigenvalues(r) = roots[r].first;
        for (Index i = 0; i < n; ++i) {
          result.eigenvectors(i, r) = f * VR(i, col);
          result.eigenvectors2(i, r) = f * VR(n + i, col);
        }
      }
      result.success = true;
      iterations_ = 0;
      return result;
    }
    HamiltonianOperator<ExcitonUKSOperator_TDA, ExcitonUKSOperator_BTDA_B> Hop(A, B);
    DavidsonSolver DS(log_);
    configureDavidson(DS);
    DS.set_matrix_type("HAM");
    const MatrixXd initial_guess = BuildFullBSEXRankedInitialGuess(A.diagonal(), B.diagonal(), opt_.nmax);
    DS.solve(Hop, opt_.nmax, initial_guess);
    result.eigenvalues = DS.eigenvalues();
    result.success = DS.success();
    iterations_ = DS.num_iterations();
    const MatrixXd ev = DS.eigenvectors();
    const Index cols = ev.cols();
    result.eigenvectors = MatrixXd(n, cols);
    result.eigenvectors2 = MatrixXd(n, cols);
    for (Index j = 0; j < cols; ++j) {
      double nx = 0.0, ny = 0.0;
      for (Index i = 0; i < n; ++i) {
        nx += ev(i, j) * ev(i, j);
        ny += ev(n + i, j) * ev(n + i, j);
      }
      const double f = std::sqrt(1.0 / (nx - ny));
      for (Index i = 0; i < n; ++i) {
        result.eigenvectors(i, j) = f * ev(i, j);
        result.eigenvectors2(i, j) = f * ev(n + i, j);
      }
    }
    return result;
  }
  const VectorXd& getEpsilonInv() const { return epsilon_0_inv_; }
  const MatrixXd& getHqp(Spin s) const { return Hqp_[(int)s]; }
  Index last_davidson_iterations() const { return iterations_; }

 private:
  void configureDavidson(DavidsonSolver& DS) const {
    DS.set_correction(opt_.davidson_correction);
    DS.set_tolerance(opt_.davidson_tolerance);
    DS.set_size_update(opt_.davidson_update);
    DS.set_iter_max(opt_.davidson_maxiter);
    DS.set_max_search_space(10 * opt_.nmax);
  }
  Logger& log_;
  TCMatrix_gwbse_spin& Mmn_;
  options opt_;
  Index homo_[2] = {0, 0};
  MatrixXd Hqp_[2];
  VectorXd epsilon_0_inv_;
  mutable Index iterations_ = 0;
};

}  // namespace xtp
}  // namespace votca
