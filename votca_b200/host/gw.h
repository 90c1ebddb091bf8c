// GW - host mirror of xtp/include/votca/xtp/gw.h:37-300 and xtp/src/libxtp/gwbse/gw.cc:35-78,210-776
// (G0W0 / evGW; QSGW is outside the BASELINE configs).  Control flow (iteration order, mixing, convergence,
// root selection) is kept; what changes is how Sigma_c is evaluated: the per-level QP searches run on host
// threads exactly as in the reference's `omp parallel for` (gw.cc:344), but their Sigma_c requests are
// collected by a batcher and served by ONE batched kernel pass per round instead of one Eigen loop each.
#pragma once
#include <condition_variable>
#include <cstring>
#include <limits>
#include <mutex>
#include <thread>
#include <unordered_map>
#include <unordered_set>

#include "anderson_mixing.h"
#include "qp_rootsearch.h"
#include "sigma.h"

namespace votca {
namespace xtp {


// Collects Sigma_c requests of the concurrently running per-level searches into batches.  A request is one
// level with a SET of frequencies (the one needed now plus the ones the search is known to need next), so
// the kernel streams the level's data once for all of them.
class SigmaBatcher {
 public:
  explicit SigmaBatcher(const Sigma_base& sigma) : sigma_(sigma) {}

  // called by worker threads: evaluates Sigma_c (and d/dw if want_deriv) at freqs for this level
  void Evaluate(Index level, const std::vector<double>& freqs, bool want_deriv, std::vector<double>& s,
                std::vector<double>& ds) {
    std::unique_lock<std::mutex> lk(mu_);
    Request r{(int)level, &freqs, want_deriv, &s, &ds, false};
    queue_.push_back(&r);
    cv_server_.notify_one();
    cv_workers_.wait(lk, [&] { return r.done; });
    if (!error_.empty()) throw std::runtime_error(error_);
  }
  void WorkerStarted() {
    std::lock_guard<std::mutex> lk(mu_);
    ++running_;
  }
  void WorkerFinished() {
    std::lock_guard<std::mutex> lk(mu_);
    --running_;
    cv_server_.notify_one();
  }
  // run on the calling thread until all workers have finished
  void Serve() {
    std::unique_lock<std::mutex> lk(mu_);
    while (true) {
      cv_server_.wait(lk, [&] { return running_ == 0 || (Index)queue_.size() == running_; });
      if (running_ == 0 && queue_.empty()) return;
      std::vector<Request*> batch;
      batch.swap(queue_);
      std::vector<int> lv(batch.size()), gp(batch.size() + 1, 0);
      std::vector<double> fr, s, ds;
      bool any_deriv = false;
      for (size_t i = 0; i < batch.size(); ++i) {
        lv[i] = batch[i]->level;
        fr.insert(fr.end(), batch[i]->freqs->begin(), batch[i]->freqs->end());
        gp[i + 1] = (int)fr.size();
        any_deriv = any_deriv || batch[i]->want_deriv;
      }
      lk.unlock();
      try {
        sigma_.EvalGroups(lv, gp, fr, s, any_deriv ? &ds : nullptr);
      } catch (const std::exception& e) {
        error_ = e.what();
        s.assign(fr.size(), 0.0);
        ds.assign(fr.size(), 0.0);
      }
      lk.lock();
      ++batches_;
      evaluations_ += fr.size();
      for (size_t i = 0; i < batch.size(); ++i) {
        batch[i]->s->assign(s.begin() + gp[i], s.begin() + gp[i + 1]);
        if (any_deriv)
          batch[i]->ds->assign(ds.begin() + gp[i], ds.begin() + gp[i + 1]);
        else
          batch[i]->ds->clear();
        batch[i]->done = true;
      }
      cv_workers_.notify_all();
    }
  }
  std::size_t batches() const { return batches_; }
  std::size_t evaluations() const { return evaluations_; }

 private:
  struct Request {
    int level;
    const std::vector<double>* freqs;
    bool want_deriv;
    std::vector<double>* s;
    std::vector<double>* ds;
    bool done;
  };
  const Sigma_base& sigma_;
  std::mutex mu_;
  std::condition_variable cv_server_, cv_workers_;
  std::vector<Request*> queue_;
  Index running_ = 0;
  std::size_t batches_ = 0, evaluations_ = 0;
  std::string error_;
};

class GW {
  using EvalStage = qp_solver::EvalStage;
  using QPStats = qp_solver::Stats;
  using QPRootCandidate = qp_solver::RootCandidate;
  using QPWindowDiagnostics = qp_solver::WindowDiagnostics;

 public:
  GW(Logger& log, TCMatrix_gwbse& Mmn, const MatrixXd& vxc, const VectorXd& dft_energies)
      : log_(log), Mmn_(Mmn), vxc_(vxc), dft_energies_(dft_energies), rpa_(log, Mmn) {}

  struct options {
    Index homo = 0, qpmin = 0, qpmax = 0, rpamin = 0, rpamax = 0;
    double eta = 1e-3;
    double g_sc_limit = 1e-5;
    Index g_sc_max_iterations = 100;
    double gw_sc_limit = 1e-5;
    Index gw_sc_max_iterations = 50;
    double shift = 0;
    double ScaHFX = 0.0;
    std::string sigma_integration = "ppm";
    Index reset_3c = 5;
    std::string qp_solver = "grid";
    double qp_solver_alpha = 0.75;
    Index qp_grid_steps = 0;
    double qp_grid_spacing = 0.0;
    double qp_full_window_half_width = -1.0;
    double qp_dense_spacing = -1.0;
    double qp_adaptive_shell_width = -1.0;
    Index qp_adaptive_shell_count = 0;
    Index gw_mixing_order = 20;
    double gw_mixing_alpha = 0.7;
    std::string quadrature_scheme = "legendre";
    Index order = 12;
    double alpha = 1e-3;
    bool qp_restrict_search = true;
    double qp_zero_margin = 1e-6;
    double qp_virtual_min_energy = -0.1;
    std::string qp_root_finder = "bisection";
    std::string qp_grid_search_mode = "adaptive_with_dense_fallback";
    // QSGW (gw.h options of the reference, gwbse.xml:41-44)
    bool do_qsgw = false;
    Index qsgw_max_iterations = 20;
    double qsgw_sc_limit = 1e-5;
    double qsgw_max_virt_correction = 0.5;
  };

  // gw.cc:35-58
  void configure(const options& opt) {
    opt_ = opt;
    qp_solver::NormalizeGridSearchOptions(opt_);
    qptotal_ = opt_.qpmax - opt_.qpmin + 1;
    rpa_.configure(opt_.homo, opt_.rpamin, opt_.rpamax);
    sigma_ = SigmaFactory_Create(opt_.sigma_integration, Mmn_, rpa_);
    Sigma_base::options sigma_opt;
    sigma_opt.homo = opt_.homo;
    sigma_opt.qpmax = opt_.qpmax;
    sigma_opt.qpmin = opt_.qpmin;
    sigma_opt.rpamin = opt_.rpamin;
    sigma_opt.rpamax = opt_.rpamax;
    sigma_opt.eta = opt_.eta;
    sigma_opt.alpha = opt_.alpha;
    sigma_opt.quadrature_scheme = opt_.quadrature_scheme;
    sigma_opt.order = opt_.order;
    sigma_->configure(sigma_opt);
    Sigma_x_ = MatrixXd::Zero(qptotal_, qptotal_);
    Sigma_c_ = MatrixXd::Zero(qptotal_, qptotal_);
  }

  // gw.cc:67-72
  MatrixXd getHQP() const {
    MatrixXd H = Sigma_x_ + Sigma_c_ - vxc_;
    for (Index i = 0; i < qptotal_; ++i) H(i, i) += dft_energies_(opt_.qpmin + i);
    return H;
  }
  // gw.cc:312-321
  VectorXd getGWAResults() const {
    // QSGW with a trimmed virtual window: converged energies of the window, seed energies above it
    if (qsgw_final_energies_.size() > 0) return qsgw_final_energies_;
    VectorXd r(qptotal_);
    for (Index i = 0; i < qptotal_; ++i)
      r(i) = Sigma_x_(i, i) + Sigma_c_(i, i) - vxc_(i, i) + dft_energies_(opt_.qpmin + i);
    return r;
  }
  VectorXd RPAInputEnergies() const { return rpa_.getRPAInputEnergies(); }
  const MatrixXd& Sigma_x() const { return Sigma_x_; }
  const MatrixXd& Sigma_c() const { return Sigma_c_; }
  Index iterations() const { return gw_sc_iteration_ + 1; }
  std::size_t sigma_batches() const { return sigma_batches_; }
  std::size_t sigma_evaluations() const { return sigma_evaluations_; }

  // gw.cc:74-78: eigen-decomposition of Hqp (device symmetric eigensolver)
  std::pair<VectorXd, MatrixXd> DiagonalizeQPHamiltonian() const {
    MatrixXd H = getHQP();
    VectorXd w = Mmn_.device().sym_eig(H);
    return {w, H};
  }

  // gw.cc:218-310
  void CalculateGWPerturbation() {
    Sigma_x_ = (1 - opt_.ScaHFX) * sigma_->CalcExchangeMatrix();
    log_(" Calculated Hartree exchange contribution");
    log_(" Scissor shifting DFT energies by: " + std::to_string(opt_.shift) + " Hrt");
    VectorXd dft_shifted_energies = ScissorShift_DFTlevel(dft_energies_);
    rpa_.setRPAInputEnergies(dft_shifted_energies.segment(opt_.rpamin, opt_.rpamax - opt_.rpamin + 1));
    VectorXd frequencies = dft_shifted_energies.segment(opt_.qpmin, qptotal_);
    Anderson mixing_;
    mixing_.Configure(opt_.gw_mixing_order, opt_.gw_mixing_alpha);
    for (Index i_gw = 0; i_gw < opt_.gw_sc_max_iterations; ++i_gw) {
      gw_sc_iteration_ = i_gw;
      if (i_gw % opt_.reset_3c == 0 && i_gw != 0) {
        Mmn_.Rebuild();
        log_(" Rebuilding 3c integrals");
      }
      sigma_->PrepareScreening();
      log_(" Calculated screening via RPA");
      log_(" Solving QP equations ");
      if (opt_.gw_mixing_order > 0 && i_gw > 0) mixing_.UpdateInput(frequencies);
      frequencies = SolveQP(frequencies);
      if (opt_.gw_sc_max_iterations > 1) {
        VectorXd rpa_energies_old = rpa_.getRPAInputEnergies();
        if (opt_.gw_mixing_order > 0 && i_gw > 0) {
          mixing_.UpdateOutput(frequencies);
          VectorXd mixed_frequencies = mixing_.MixHistory();
          rpa_.UpdateRPAInputEnergies(dft_energies_, mixed_frequencies, opt_.qpmin);
          frequencies = mixed_frequencies;
        } else {
          rpa_.UpdateRPAInputEnergies(dft_energies_, frequencies, opt_.qpmin);
        }
        log_(" GW_Iteration:" + std::to_string(i_gw) + " Shift[Hrt]:" + std::to_string(CalcHomoLumoShift(frequencies)));
        if (Converged(rpa_.getRPAInputEnergies(), rpa_energies_old, opt_.gw_sc_limit)) {
          log_(" Converged after " + std::to_string(i_gw + 1) + " GW iterations.");
          break;
        } else if (i_gw == opt_.gw_sc_max_iterations - 1) {
          log_(" WARNING! GW-self-consistency cycle not converged after " +
               std::to_string(opt_.gw_sc_max_iterations) + " iterations.");
          log_("      Run continues. Inspect results carefully!");
          break;
        }
      }
    }
    VectorXd diag = sigma_->CalcCorrelationDiag(frequencies);
    for (Index i = 0; i < qptotal_; ++i) Sigma_c_(i, i) = diag(i);
  }

  const MatrixXd& getQSGWRotation() const { return qsgw_rotation_; }
  const VectorXd& getQSGWSeedEnergies() const { return qsgw_seed_energies_; }
  Index qsgw_iterations() const { return qsgw_iterations_; }

  // GW::CalculateQSGW, gw.cc:798-1130.  Per iteration: Mmn back to the DFT-MO basis (device snapshot), rotation of
  // the QP-window rows (gwbse_mmn_rotate), screening with the hole slices rotated inside the RPA sums
  // (gwbse_rpa_set_qsgw_rotation), the symmetrised static self-energy from the evaluators' batched kernels, Anderson
  // mixing and two qptotal x qptotal eigenproblems on the host side.  The caller has put Mmn back into the DFT-MO
  // basis (Mmn.Rebuild()) as gwbse.cc does before this call.
  void CalculateQSGW() {
    const Device& dev = Mmn_.device();
    if (dev.world() > 1) throw std::runtime_error("GW::CalculateQSGW: QSGW is single-GPU in this build");
    log_(" Starting QSGW self-consistency loop  ");
    const VectorXd e_qp_full = getGWAResults();
    qsgw_seed_energies_ = e_qp_full;
    qsgw_final_energies_ = VectorXd();
    // virtual-level threshold: trim the window at the first virtual whose perturbative correction is too large
    Index qsgw_qpmax = opt_.qpmax;
    const Index lumo_local = opt_.homo - opt_.qpmin + 1;
    for (Index n = lumo_local; n < qptotal_; ++n) {
      const double corr = std::abs(e_qp_full(n) - dft_energies_(opt_.qpmin + n));
      if (corr > opt_.qsgw_max_virt_correction) {
        qsgw_qpmax = opt_.qpmin + n - 1;
        log_("  QSGW virtual threshold: level " + std::to_string(opt_.qpmin + n) + " exceeds the limit. Trimming QSGW " +
             "window to [" + std::to_string(opt_.qpmin) + "," + std::to_string(qsgw_qpmax) + "].");
        break;
      }
    }
    const Index nq = qsgw_qpmax - opt_.qpmin + 1;
    const bool window_trimmed = qsgw_qpmax < opt_.qpmax;
    auto sigma_options = [&](Index qpmax) {
      Sigma_base::options so;
      so.homo = opt_.homo;
      so.qpmin = opt_.qpmin;
      so.qpmax = qpmax;
      so.rpamin = opt_.rpamin;
      so.rpamax = opt_.rpamax;
      so.eta = opt_.eta;
      so.quadrature_scheme = opt_.quadrature_scheme;
      so.order = opt_.order;
      so.alpha = opt_.alpha;
      return so;
    };
    if (window_trimmed) {
      sigma_->configure(sigma_options(qsgw_qpmax));
      Sigma_x_ = MatrixXd::Zero(nq, nq);
      Sigma_c_ = MatrixXd::Zero(nq, nq);
    }
    VectorXd e_qp = e_qp_full.head(nq);
    qsgw_rotation_ = MatrixXd::Identity(nq, nq);
    if (opt_.ScaHFX > 0.0)
      throw std::runtime_error(
          "GW::CalculateQSGW: QSGW is not compatible with hybrid DFT starting points (ScaHFX = " +
          std::to_string(opt_.ScaHFX) + "). Use a pure GGA or LDA functional as the DFT starting point.");
    if (opt_.sigma_integration == "cda")
      throw std::runtime_error(
          "GW::CalculateQSGW: QSGW is not supported with the CDA sigma integration method. Use sigma_integration=ppm or "
          "sigma_integration=exact instead.");
    MatrixXd H0 = -1.0 * vxc_.block(0, 0, nq, nq);
    for (Index i = 0; i < nq; ++i) H0(i, i) += dft_energies_(opt_.qpmin + i);
    Anderson qsgw_mixer;
    qsgw_mixer.Configure(opt_.gw_mixing_order, opt_.gw_mixing_alpha);
    auto register_rotation = [&](const MatrixXd* U) {
      dev.check(gwbse_rpa_set_qsgw_rotation(dev.ctx(), U ? U->data() : nullptr, U ? (int)U->rows() : 0,
                                            U ? (int)U->cols() : 0, (int)opt_.qpmin, (int)opt_.homo));
    };
    auto flat = [&](const MatrixXd& m) { return VectorXd(m.data(), m.size()); };
    double diff_max_prev = std::numeric_limits<double>::max();
    MatrixXd tilde_Sigma;
    for (Index iter = 0; iter < opt_.qsgw_max_iterations; ++iter) {
      qsgw_iterations_ = iter + 1;
      // Step 1: DFT-MO basis again, then the full rotation in one shot
      Mmn_.Rebuild();
      if (iter > 0) Mmn_.Rotate(qsgw_rotation_, opt_.qpmin, qsgw_qpmax);
      register_rotation(&qsgw_rotation_);
      sigma_->PrepareScreening();
      Sigma_x_ = sigma_->CalcExchangeMatrix();
      const MatrixXd Sc_row = sigma_->CalcCorrelationOffDiag(e_qp);
      tilde_Sigma = Sigma_x_ + 0.5 * (Sc_row + Sc_row.transpose());
      const VectorXd Sc_diag = sigma_->CalcCorrelationDiag(e_qp);
      for (Index i = 0; i < nq; ++i) tilde_Sigma(i, i) += Sc_diag(i);
      // Step 2: Anderson mixing of the flattened self-energy
      VectorXd S_flat = flat(tilde_Sigma);
      if (iter > 0) {
        qsgw_mixer.UpdateOutput(S_flat);
        S_flat = qsgw_mixer.MixHistory();
      }
      // Step 3: rotation from the mixed Hamiltonian, energies (convergence) from the unmixed one
      MatrixXd dU = H0 + MatrixXd(S_flat.data(), nq, nq, nq);
      dev.sym_eig(dU);
      MatrixXd H_new = H0 + tilde_Sigma;
      const VectorXd e_new = dev.sym_eig(H_new);
      const double diff_max = (e_new - e_qp).maxAbs();
      log_("  QSGW iter " + std::to_string(iter) + "  max|dE_QP| = " + std::to_string(diff_max * 27.211386) + " eV");
      if (iter > 1 && diff_max > 2.0 * diff_max_prev) {
        qsgw_mixer = Anderson();
        qsgw_mixer.Configure(opt_.gw_mixing_order, opt_.gw_mixing_alpha);
      }
      diff_max_prev = diff_max;
      if (diff_max < opt_.qsgw_sc_limit) {
        log_("  QSGW converged in " + std::to_string(iter + 1) + " iterations.");
        e_qp = e_new;
        qsgw_rotation_ = dU;
        rpa_.UpdateRPAInputEnergies(dft_energies_, e_qp, opt_.qpmin);
        break;
      }
      if (iter == opt_.qsgw_max_iterations - 1)
        log_("  WARNING: QSGW did not converge in " + std::to_string(opt_.qsgw_max_iterations) +
             " iterations. Inspect results carefully.");
      e_qp = e_new;
      qsgw_rotation_ = dU;
      qsgw_mixer.UpdateInput(flat(tilde_Sigma));
      rpa_.UpdateRPAInputEnergies(dft_energies_, e_qp, opt_.qpmin);
    }
    const VectorXd diag = sigma_->CalcCorrelationDiag(e_qp);
    for (Index i = 0; i < nq; ++i) Sigma_c_(i, i) = diag(i);
    register_rotation(nullptr);
    if (window_trimmed) {
      MatrixXd U_full = MatrixXd::Identity(qptotal_, qptotal_);
      U_full.setBlock(0, 0, qsgw_rotation_);
      qsgw_rotation_ = U_full;
      VectorXd e_merged = e_qp_full;
      for (Index i = 0; i < nq; ++i) e_merged(i) = e_qp(i);
      qsgw_final_energies_ = e_merged;
      rpa_.UpdateRPAInputEnergies(dft_energies_, e_merged, opt_.qpmin);
      Sigma_x_ = MatrixXd::Zero(qptotal_, qptotal_);
      Sigma_c_ = MatrixXd::Zero(qptotal_, qptotal_);
      sigma_->configure(sigma_options(opt_.qpmax));
    }
    log_(" QSGW loop complete.");
  }

  // gw.cc:772-776
  void CalculateHQP() {
    VectorXd diag_backup = Sigma_c_.diagonal();
    Sigma_c_ = sigma_->CalcCorrelationOffDiag(getGWAResults());
    for (Index i = 0; i < qptotal_; ++i) Sigma_c_(i, i) = diag_backup(i);
  }

 private:
  // f(w) = Sigma_c(w) + offset - w, gw.h:214-300.  Sigma_c requests go through the batcher.  Every value the
  // GPU returns is cached under the frequency's bit pattern (the key the reference already uses for its
  // statistics, gw.h:261-267); prefetch() announces frequencies the search will ask for next so that they
  // ride along with the current request.  The cache only removes round trips: each Sigma_c(w) is computed
  // by the same kernel arithmetic whether it was prefetched or not.
  class QPFunc {
   public:
    QPFunc(Index gw_level, SigmaBatcher& batcher, double offset)
        : gw_level_(gw_level), offset_(offset), batcher_(batcher) {}
    std::pair<double, double> operator()(double frequency) const {
      Count(frequency, EvalStage::Other);
      ++stats_.deriv_calls;
      const Entry& e = Fetch(frequency, true);
      return {e.s + offset_ - frequency, e.ds - 1.0};
    }
    double sigma(double frequency, EvalStage stage = EvalStage::Other) const {
      Count(frequency, stage);
      return Fetch(frequency, false).s;
    }
    double value(double frequency, EvalStage stage = EvalStage::Other) const {
      return sigma(frequency, stage) + offset_ - frequency;
    }
    double deriv(double frequency) const {
      ++stats_.deriv_calls;
      return Fetch(frequency, true).ds - 1.0;
    }
    // hint: these frequencies will be requested soon (with_deriv: value and derivative)
    void prefetch(const double* freqs, std::size_t n, bool with_deriv = false) const {
      for (std::size_t i = 0; i < n; ++i) {
        auto it = cache_.find(Key(freqs[i]));
        if (it != cache_.end() && (!with_deriv || it->second.has_ds)) continue;
        pending_.push_back(freqs[i]);
        pending_deriv_ = pending_deriv_ || with_deriv;
      }
    }
    const QPStats& GetStats() const { return stats_; }

   private:
    struct Entry {
      double s = 0.0, ds = 0.0;
      bool has_ds = false;
    };
    static std::uint64_t Key(double x) {
      std::uint64_t key = 0;
      std::memcpy(&key, &x, sizeof(double));
      return key;
    }
    const Entry& Fetch(double frequency, bool need_deriv) const {
      auto it = cache_.find(Key(frequency));
      if (it != cache_.end() && (!need_deriv || it->second.has_ds)) return it->second;
      std::vector<double> fr;
      fr.push_back(frequency);
      std::unordered_set<std::uint64_t> in_req{Key(frequency)};
      for (double f : pending_)
        if (in_req.insert(Key(f)).second && fr.size() < 96) fr.push_back(f);
      const bool want = need_deriv || pending_deriv_;
      pending_.clear();
      pending_deriv_ = false;
      std::vector<double> s, ds;
      batcher_.Evaluate(gw_level_, fr, want, s, ds);
      for (std::size_t i = 0; i < fr.size(); ++i) {
        Entry& e = cache_[Key(fr[i])];
        e.s = s[i];
        if (!ds.empty()) {
          e.ds = ds[i];
          e.has_ds = true;
        }
      }
      return cache_[Key(frequency)];
    }
    void Count(double x, EvalStage stage) const {
      if (!seen_frequencies_.insert(Key(x)).second)
        ++stats_.sigma_repeat_calls;
      else
        ++stats_.sigma_unique_frequencies;
      switch (stage) {
        case EvalStage::Scan: ++stats_.sigma_scan_calls; break;
        case EvalStage::Refine: ++stats_.sigma_refine_calls; break;
        case EvalStage::Derivative: ++stats_.sigma_derivative_calls; break;
        default: ++stats_.sigma_other_calls; break;
      }
    }
    Index gw_level_;
    double offset_;
    SigmaBatcher& batcher_;
    mutable std::unordered_map<std::uint64_t, Entry> cache_;
    mutable std::vector<double> pending_;
    mutable bool pending_deriv_ = false;
    mutable std::unordered_set<std::uint64_t> seen_frequencies_;
    mutable QPStats stats_;
  };

  MatrixXd qsgw_rotation_;
  VectorXd qsgw_seed_energies_, qsgw_final_energies_;
  Index qsgw_iterations_ = 0;

  double CalcHomoLumoShift(const VectorXd& frequencies) const {
    double DFTgap = dft_energies_(opt_.homo + 1) - dft_energies_(opt_.homo);
    double QPgap = frequencies(opt_.homo + 1 - opt_.qpmin) - frequencies(opt_.homo - opt_.qpmin);
    return QPgap - DFTgap;
  }
  VectorXd ScissorShift_DFTlevel(const VectorXd& dft_energies) const {
    VectorXd shifted = dft_energies;
    for (Index i = opt_.homo + 1; i < shifted.size(); ++i) shifted(i) += opt_.shift;
    return shifted;
  }
  bool Converged(const VectorXd& e1, const VectorXd& e2, double epsilon) const {
    Index state = 0;
    double diff_max = 0.0;
    for (Index i = 0; i < e1.size(); ++i)
      if (std::abs(e1(i) - e2(i)) > diff_max) {
        diff_max = std::abs(e1(i) - e2(i));
        state = i;
      }
    log_(" E_diff max=" + std::to_string(diff_max) + " StateNo:" + std::to_string(state));
    return !(diff_max > epsilon);
  }

  // gw.cc:323-410: one host thread per level (the reference's dynamic OpenMP loop), batched Sigma_c
  VectorXd SolveQP(const VectorXd& frequencies) {
    sigma_->ResetDiagEvalCounter();
    VectorXd intercepts(qptotal_);
    for (Index i = 0; i < qptotal_; ++i)
      intercepts(i) = dft_energies_(opt_.qpmin + i) + Sigma_x_(i, i) - vxc_(i, i);
    VectorXd frequencies_new = frequencies;
    std::vector<char> converged(qptotal_, 0);
    std::vector<QPStats> stats(qptotal_);
    std::vector<std::string> errors(qptotal_);
    SigmaBatcher batcher(*sigma_);
    std::vector<std::thread> workers;
    workers.reserve(qptotal_);
    // multi-GPU: every rank searches the roots of the levels whose Mmn slice it owns
    const Device& dev = Mmn_.device();
    std::vector<Index> my_levels;
    for (Index gw_level = 0; gw_level < qptotal_; ++gw_level)
      if (sigma_->OwnsLevel(gw_level)) my_levels.push_back(gw_level);
    for (size_t w = 0; w < my_levels.size(); ++w) batcher.WorkerStarted();
    for (Index gw_level : my_levels) {
      workers.emplace_back([&, gw_level] {
        try {
          double initial_f = frequencies[gw_level];
          double intercept = intercepts[gw_level];
          std::optional<double> newf;
          if (opt_.qp_solver == "fixedpoint")
            newf = SolveQP_FixedPoint(batcher, intercept, initial_f, gw_level, &stats[gw_level]);
          if (newf) {
            frequencies_new[gw_level] = *newf;
            converged[gw_level] = 1;
          } else {
            newf = SolveQP_Grid(batcher, intercept, initial_f, gw_level, &stats[gw_level]);
            if (newf) {
              frequencies_new[gw_level] = *newf;
              converged[gw_level] = 1;
            } else {
              newf = SolveQP_Linearisation(batcher, intercept, initial_f, gw_level, &stats[gw_level]);
              if (newf) frequencies_new[gw_level] = *newf;
            }
          }
        } catch (const std::exception& e) {
          errors[gw_level] = e.what();
        }
        batcher.WorkerFinished();
      });
    }
    batcher.Serve();
    for (auto& t : workers) t.join();
    for (const auto& e : errors)
      if (!e.empty()) throw std::runtime_error(e);
    if (dev.world() > 1) {
      std::vector<double> pack(2 * qptotal_, 0.0);
      for (Index gw_level : my_levels) {
        pack[gw_level] = frequencies_new[gw_level];
        pack[qptotal_ + gw_level] = converged[gw_level] ? 1.0 : 0.0;
      }
      dev.allreduce(pack.data(), pack.size());
      for (Index i = 0; i < qptotal_; ++i) {
        frequencies_new[i] = pack[i];
        converged[i] = pack[qptotal_ + i] > 0.5 ? 1 : 0;
      }
    }
    QPStats total_stats;
    for (const auto& s : stats) total_stats.Add(s);
    sigma_batches_ += batcher.batches();
    sigma_evaluations_ += batcher.evaluations();
    std::string notconv;
    for (Index s = 0; s < qptotal_; ++s)
      if (!converged[s]) notconv += " " + std::to_string(s);
    if (!notconv.empty()) {
      log_(" Not converged PQP states are:" + notconv);
      log_(" Increase the grid search interval");
    }
    log_(" Sigma diagonal evaluations in SolveQP: " + std::to_string(batcher.evaluations()) + " in " +
         std::to_string(batcher.batches()) + " batched kernel passes");
    log_(" QP diagnostics: scan=" + std::to_string(total_stats.sigma_scan_calls) +
         " refine=" + std::to_string(total_stats.sigma_refine_calls) +
         " other=" + std::to_string(total_stats.sigma_other_calls) +
         " total_sigma=" + std::to_string(total_stats.TotalSigmaCalls()) +
         " unique_omega=" + std::to_string(total_stats.sigma_unique_frequencies) +
         " repeat_sigma=" + std::to_string(total_stats.sigma_repeat_calls) +
         " deriv_calls=" + std::to_string(total_stats.deriv_calls));
    return frequencies_new;
  }

  qp_solver::SolverOptions MakeSolverOptions() const {
    qp_solver::SolverOptions s;
    s.g_sc_limit = opt_.g_sc_limit;
    s.qp_bisection_max_iter = opt_.g_sc_max_iterations;
    s.qp_full_window_half_width = opt_.qp_full_window_half_width;
    s.qp_dense_spacing = opt_.qp_dense_spacing;
    s.qp_adaptive_shell_width = opt_.qp_adaptive_shell_width;
    s.qp_adaptive_shell_count = opt_.qp_adaptive_shell_count;
    return s;
  }

  // gw.cc:412-431
  std::optional<double> SolveQP_Linearisation(SigmaBatcher& b, double intercept0, double frequency0, Index gw_level,
                                              QPStats* stats) const {
    std::optional<double> newf;
    QPFunc fqp(gw_level, b, intercept0);
    double sigma = fqp.sigma(frequency0, EvalStage::Other);
    double dsigma_domega = fqp.deriv(frequency0);
    double Z = 1.0 - dsigma_domega;
    if (std::abs(Z) > 1e-9) newf = frequency0 + (intercept0 - frequency0 + sigma) / Z;
    if (stats) *stats = fqp.GetStats();
    return newf;
  }

  // gw.cc:433-502
  std::optional<double> SolveQP_Grid_Windowed_Adaptive(SigmaBatcher& b, double intercept0, double frequency0,
                                                       Index gw_level, double left_limit, double right_limit,
                                                       bool allow_rejected_return, QPStats* stats) const {
    QPFunc fqp(gw_level, b, intercept0);
    qp_solver::SolverOptions solver_opt = MakeSolverOptions();
    QPWindowDiagnostics wdiag;
    std::vector<QPRootCandidate> accepted_roots, rejected_roots;
    const bool use_brent = (opt_.qp_root_finder == "brent");
    auto result = qp_solver::SolveQP_Grid_Windowed(fqp, frequency0, left_limit, right_limit, gw_sc_iteration_,
                                                   solver_opt, &wdiag, &accepted_roots, &rejected_roots, use_brent);
    if (stats) *stats = fqp.GetStats();
    if (!accepted_roots.empty()) return result;
    if (!rejected_roots.empty() && !allow_rejected_return) return std::nullopt;
    return result;
  }

  // gw.cc:504-616
  std::optional<double> SolveQP_Grid_Windowed_Dense(SigmaBatcher& b, double intercept0, double frequency0,
                                                    Index gw_level, double left_limit, double right_limit,
                                                    bool allow_rejected_return, QPStats* stats) const {
    QPFunc fqp(gw_level, b, intercept0);
    qp_solver::SolverOptions solver_opt = MakeSolverOptions();
    const bool use_brent = (opt_.qp_root_finder == "brent");
    std::vector<QPRootCandidate> accepted_roots, rejected_roots;
    if (left_limit < right_limit) {
      double freq_prev = left_limit;
      const Index n_steps =
          std::max<Index>(2, static_cast<Index>(std::ceil((right_limit - left_limit) / opt_.qp_dense_spacing)) + 1);
      auto node = [&](Index i_node) {
        return (i_node == n_steps - 1)
                   ? right_limit
                   : std::min(right_limit, left_limit + static_cast<double>(i_node) * opt_.qp_dense_spacing);
      };
      const Index look = 64;  // the scan nodes are known ahead: announce them in blocks
      auto announce = [&](Index from) {
        std::vector<double> pts;
        for (Index i = from; i < std::min(n_steps, from + look); ++i) pts.push_back(node(i));
        fqp.prefetch(pts.data(), pts.size());
      };
      announce(1);
      double targ_prev = fqp.value(freq_prev, EvalStage::Scan);
      for (Index i_node = 1; i_node < n_steps; ++i_node) {
        if (i_node > 1 && (i_node - 1) % look == 0) announce(i_node);
        const double freq = node(i_node);
        const double targ = fqp.value(freq, EvalStage::Scan);
        if (targ_prev * targ < 0.0) {
          auto cand =
              qp_solver::RefineQPInterval(freq_prev, targ_prev, freq, targ, fqp, frequency0, solver_opt, use_brent);
          if (cand) (cand->accepted ? accepted_roots : rejected_roots).push_back(*cand);
        }
        freq_prev = freq;
        targ_prev = targ;
      }
    }
    if (stats) *stats = fqp.GetStats();
    if (!accepted_roots.empty()) return qp_solver::BestRoot(accepted_roots).omega;
    if (!rejected_roots.empty()) {
      if (!allow_rejected_return) return std::nullopt;
      return qp_solver::BestRoot(rejected_roots).omega;
    }
    return std::nullopt;
  }

  // gw.cc:618-675
  std::optional<double> SolveQP_Grid_Windowed(SigmaBatcher& b, double intercept0, double frequency0, Index gw_level,
                                              double left_limit, double right_limit, bool allow_rejected_return,
                                              QPStats* stats) const {
    if (opt_.qp_grid_search_mode == "adaptive")
      return SolveQP_Grid_Windowed_Adaptive(b, intercept0, frequency0, gw_level, left_limit, right_limit,
                                            allow_rejected_return, stats);
    if (opt_.qp_grid_search_mode == "dense")
      return SolveQP_Grid_Windowed_Dense(b, intercept0, frequency0, gw_level, left_limit, right_limit,
                                         allow_rejected_return, stats);
    if (opt_.qp_grid_search_mode == "adaptive_with_dense_fallback") {
      QPStats total_stats;
      auto adaptive = SolveQP_Grid_Windowed_Adaptive(b, intercept0, frequency0, gw_level, left_limit, right_limit,
                                                     allow_rejected_return, &total_stats);
      if (adaptive) {
        if (stats) *stats = total_stats;
        return adaptive;
      }
      QPStats dense_stats;
      auto dense = SolveQP_Grid_Windowed_Dense(b, intercept0, frequency0, gw_level, left_limit, right_limit,
                                               allow_rejected_return, &dense_stats);
      total_stats.Add(dense_stats);
      if (stats) *stats = total_stats;
      return dense;
    }
    throw std::runtime_error("Unknown gw.qp_grid_search_mode '" + opt_.qp_grid_search_mode + "'");
  }

  // gw.cc:677-739
  std::optional<double> SolveQP_Grid(SigmaBatcher& b, double intercept0, double frequency0, Index gw_level,
                                     QPStats* stats) const {
    const double range = opt_.qp_full_window_half_width;
    const double full_left_limit = frequency0 - range;
    const double full_right_limit = frequency0 + range;
    double restricted_left_limit = full_left_limit;
    double restricted_right_limit = full_right_limit;
    bool use_restricted_window = false;
    if (opt_.qp_restrict_search) {
      const Index mo_level = gw_level + opt_.qpmin;
      const bool is_occupied = (mo_level <= opt_.homo);
      if (is_occupied)
        restricted_right_limit = std::min(full_right_limit, -opt_.qp_zero_margin);
      else
        restricted_left_limit = std::max(full_left_limit, opt_.qp_virtual_min_energy);
      const double tol = 1e-12;
      use_restricted_window = (std::abs(restricted_left_limit - full_left_limit) > tol) ||
                              (std::abs(restricted_right_limit - full_right_limit) > tol);
    }
    if (use_restricted_window && restricted_left_limit < restricted_right_limit) {
      auto restricted = SolveQP_Grid_Windowed(b, intercept0, frequency0, gw_level, restricted_left_limit,
                                              restricted_right_limit, false, stats);
      if (restricted) return restricted;
      return SolveQP_Grid_Windowed_Dense(b, intercept0, frequency0, gw_level, full_left_limit, full_right_limit,
                                         true, stats);
    }
    return SolveQP_Grid_Windowed(b, intercept0, frequency0, gw_level, full_left_limit, full_right_limit, true, stats);
  }

  // gw.cc:741-757
  std::optional<double> SolveQP_FixedPoint(SigmaBatcher& b, double intercept0, double frequency0, Index gw_level,
                                           QPStats* stats) const {
    std::optional<double> newf;
    QPFunc f(gw_level, b, intercept0);
    NewtonRapson<QPFunc> newton(opt_.g_sc_max_iterations, opt_.g_sc_limit, opt_.qp_solver_alpha);
    double freq_new = newton.FindRoot(f, frequency0);
    if (newton.getInfo() == NewtonRapson<QPFunc>::success) newf = freq_new;
    if (stats) *stats = f.GetStats();
    return newf;
  }

  Index qptotal_ = 0;
  MatrixXd Sigma_x_, Sigma_c_;
  options opt_;
  std::unique_ptr<Sigma_base> sigma_;
  Logger& log_;
  TCMatrix_gwbse& Mmn_;
  const MatrixXd& vxc_;
  const VectorXd& dft_energies_;
  Index gw_sc_iteration_ = 0;
  std::size_t sigma_batches_ = 0, sigma_evaluations_ = 0;
  RPA rpa_;
};

}  // namespace xtp
}  // namespace votca
