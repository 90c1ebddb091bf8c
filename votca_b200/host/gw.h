// GW - host mirror of xtp/include/votca/xtp/gw.h:37-300 and xtp/src/libxtp/gwbse/gw.cc:35-78,210-1130
// (G0W0 / evGW / QSGW).  Iteration order, mixing, convergence and root selection give the reference's results; what
// changes is how Sigma_c is evaluated: the per-level QP searches (the reference's `omp parallel for`, gw.cc:344) are
// resumable objects (qp_rootsearch.h) advanced in lock step, so every round of all searches is ONE grouped kernel
// pass over the Mmn slices instead of one Eigen loop per level and frequency.
#pragma once
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <limits>
#include <memory>

#include "anderson_mixing.h"
#include "qp_rootsearch.h"
#include "sigma.h"

namespace votca {
namespace xtp {


class GW {
  friend class GW_UKS;  // uks.h: two GW objects are the spin channels of the unrestricted loop
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

  // gw.cc:35-58.  `evaluator`: a ready-made self-energy evaluator instead of the factory's (GW_UKS installs the
  // spin channel's Sigma_PPM_UKS, gw_uks.cc:46-49)
  void configure(const options& opt, std::unique_ptr<Sigma_base> evaluator = nullptr) {
    opt_ = opt;
    qp_solver::NormalizeGridSearchOptions(opt_);
    qptotal_ = opt_.qpmax - opt_.qpmin + 1;
    rpa_.configure(opt_.homo, opt_.rpamin, opt_.rpamax);
    sigma_ = evaluator ? std::move(evaluator) : SigmaFactory_Create(opt_.sigma_integration, Mmn_, rpa_);
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

  // gw.cc:1132-1182: E_QP(omega) = e_DFT + Sigma_x - Vxc + Sigma_c(omega) on a grid of `steps` points `spacing`
  // apart around the RPA input energy of every chosen state, written as the reference writes it (one column pair
  // per state, "%+1.6f", tab separated).  All (state, omega) pairs go to the device in one grouped call.
  // `states`: IndexParser syntax ("0 3 5:9"); states outside the qp window are skipped.  Array positions are
  // counted from qpmin (the reference indexes them with the absolute level, which is the same thing for qpmin = 0).
  void PlotSigma(const std::string& filename, Index steps, double spacing, const std::string& states) const {
    std::vector<Index> state_inds;
    for (Index gw_level : ParseIndexList(states))
      if (gw_level >= opt_.qpmin && gw_level <= opt_.qpmax) state_inds.push_back(gw_level);
    std::string listed;
    for (Index l : state_inds) listed += (listed.empty() ? "" : " ") + std::to_string(l);
    log_(" PQP(omega) written to '" + filename + "' for states " + listed);
    const Index num_states = static_cast<Index>(state_inds.size());
    const VectorXd& rpa_e = rpa_.getRPAInputEnergies();
    // every rank evaluates the levels whose Mmn slice it owns (all of them on one GPU), the tables are summed
    std::vector<int> levels, gptr{0};
    std::vector<double> freqs, mine_f, mine_s;
    std::vector<size_t> first;  // position of a level's first grid point in the full table, for the owned levels
    for (Index i = 0; i < num_states; ++i) {
      const Index rel = state_inds[i] - opt_.qpmin;
      const bool own = sigma_->OwnsLevel(rel);
      if (own) {
        levels.push_back(static_cast<int>(rel));
        first.push_back(freqs.size());
      }
      for (Index g = 0; g < steps; ++g) {
        const double w = rpa_e(opt_.qpmin - opt_.rpamin + rel) + ((double)g - ((double)(steps - 1) / 2.0)) * spacing;
        freqs.push_back(w);
        if (own) mine_f.push_back(w);
      }
      if (own) gptr.push_back(static_cast<int>(mine_f.size()));
    }
    std::vector<double> sig(freqs.size(), 0.0);
    if (!mine_f.empty()) {
      sigma_->CountDiagEval(mine_f.size());
      sigma_->EvalGroups(levels, gptr, mine_f, mine_s, nullptr);
      for (size_t l = 0; l < levels.size(); ++l)
        for (Index g = 0; g < steps; ++g) sig[first[l] + static_cast<size_t>(g)] = mine_s[static_cast<size_t>(gptr[l]) + g];
    }
    if (Mmn_.device().world() > 1 && !sig.empty()) Mmn_.device().allreduce(sig.data(), sig.size());
    if (Mmn_.device().rank() != 0) return;  // one writer
    std::ofstream out(filename);
    if (!out) throw std::runtime_error("GW::PlotSigma: cannot open " + filename);
    for (Index i = 0; i < num_states; ++i)
      out << "#" << (i == 0 ? "" : "\t") << "omega_" << state_inds[i] << "\tE_QP(omega)_" << state_inds[i];
    out << std::endl;
    char buf[64];
    for (Index g = 0; g < steps; ++g) {
      for (Index i = 0; i < num_states; ++i) {
        const Index rel = state_inds[i] - opt_.qpmin;
        const size_t k = static_cast<size_t>(i * steps + g);
        const double intercept = dft_energies_(opt_.qpmin + rel) + Sigma_x_(rel, rel) - vxc_(rel, rel);
        std::snprintf(buf, sizeof(buf), "%s%+1.6f\t%+1.6f", i == 0 ? "" : "\t", freqs[k], sig[k] + intercept);
        out << buf;
      }
      out << "\n";
    }
    out << std::endl;
  }
  // IndexParser::CreateIndexVector (IndexParser.cc:36-70): tokens separated by blanks / commas, "a:b" ranges,
  // sorted, duplicates removed
  static std::vector<Index> ParseIndexList(const std::string& ids) {
    std::vector<Index> result;
    std::string tok;
    auto flush = [&]() {
      if (tok.empty()) return;
      try {
        const size_t c = tok.find(':');
        size_t used = 0;
        if (c != std::string::npos) {
          const std::string a = tok.substr(0, c), b = tok.substr(c + 1);
          const long start = std::stol(a, &used);
          if (used != a.size()) throw std::invalid_argument(a);
          const long stop = std::stol(b, &used);
          if (used != b.size()) throw std::invalid_argument(b);
          for (long i = start; i <= stop; ++i) result.push_back(static_cast<Index>(i));
        } else {
          const long v = std::stol(tok, &used);
          if (used != tok.size()) throw std::invalid_argument(tok);
          result.push_back(static_cast<Index>(v));
        }
      } catch (const std::exception&) {
        throw std::runtime_error("Could not convert " + tok +
                                 (tok.find(':') != std::string::npos ? " to range of integers." : " to integer."));
      }
      tok.clear();
    };
    for (char ch : ids) {
      if (ch == ' ' || ch == ',' || ch == '\n' || ch == '\t')
        flush();
      else
        tok.push_back(ch);
    }
    flush();
    std::sort(result.begin(), result.end());
    result.erase(std::unique(result.begin(), result.end()), result.end());
    return result;
  }

  // gw.cc:74-78: eigen-decomposition of Hqp (device symmetric eigensolver)
  std::pair<VectorXd, MatrixXd> DiagonalizeQPHamiltonian() const {
    MatrixXd H = getHQP();
    VectorXd w = Mmn_.device().sym_eig(H);
    PrintQP_Energies(w);
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
    PrintGWA_Energies();
  }

  // gw.cc:80-111
  void PrintGWA_Energies() const {
    const VectorXd gwa_energies = getGWAResults();
    log_("  ====== Perturbative quasiparticle energies (Hartree) ====== ");
    char buf[200];
    if (opt_.homo >= opt_.qpmin && opt_.homo + 1 <= opt_.qpmax) {  // the gap needs both frontier levels in the window
      std::snprintf(buf, sizeof(buf), "   DeltaHLGap = %+1.6f Hartree", CalcHomoLumoShift(gwa_energies));
      log_(buf);
    }
    for (Index i = 0; i < qptotal_; ++i) {
      const char* level = (i + opt_.qpmin) == opt_.homo ? "  HOMO " : (i + opt_.qpmin) == opt_.homo + 1 ? "  LUMO " : "  Level";
      std::snprintf(buf, sizeof(buf), "%s = %4ld DFT = %+1.4f VXC = %+1.4f S-X = %+1.4f S-C = %+1.4f GWA = %+1.4f", level,
                    (long)(i + opt_.qpmin), dft_energies_(i + opt_.qpmin), vxc_(i, i), Sigma_x_(i, i), Sigma_c_(i, i),
                    gwa_energies(i));
      log_(buf);
    }
  }
  // gw.cc:183-208
  void PrintQP_Energies(const VectorXd& qp_diag_energies) const {
    const VectorXd gwa_energies = getGWAResults();
    log_(" Full quasiparticle Hamiltonian  ");
    log_("  ====== Diagonalized quasiparticle energies (Hartree) ====== ");
    char buf[160];
    for (Index i = 0; i < qptotal_; ++i) {
      const char* level = (i + opt_.qpmin) == opt_.homo ? "  HOMO " : (i + opt_.qpmin) == opt_.homo + 1 ? "  LUMO " : "  Level";
      std::snprintf(buf, sizeof(buf), "%s = %4ld PQP = %+1.6f DQP = %+1.6f ", level, (long)(i + opt_.qpmin),
                    gwa_energies(i), qp_diag_energies(i));
      log_(buf);
    }
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
  // The search for one level: f(w) = Sigma_c(w) + offset - w (gw.h:214-300 of the reference) on a tape, and the
  // ladder of strategies gw.cc:323-410 + :618-757 climbs until one yields a frequency - Newton iteration when
  // qp_solver=fixedpoint, the shell scan and / or the dense sweep on the window restricted to the physical side of
  // zero, the dense sweep on the full window, and the linearised estimate as the last resort (not "converged").
  class LevelSearch {
   public:
    LevelSearch(Index gw_level, double offset, double frequency0, const options& opt, Index gw_iteration)
        : level_(gw_level), offset_(offset), w0_(frequency0) {
      qp_solver::SolverOptions so;
      so.g_sc_limit = opt.g_sc_limit;
      so.qp_bisection_max_iter = opt.g_sc_max_iterations;
      so.qp_full_window_half_width = opt.qp_full_window_half_width;
      so.qp_dense_spacing = opt.qp_dense_spacing;
      so.qp_adaptive_shell_width = opt.qp_adaptive_shell_width;
      so.qp_adaptive_shell_count = opt.qp_adaptive_shell_count;
      const bool brent = opt.qp_root_finder == "brent";
      const std::string& mode = opt.qp_grid_search_mode;
      const bool shells = mode == "adaptive" || mode == "adaptive_with_dense_fallback";
      const bool sweep = mode == "dense" || mode == "adaptive_with_dense_fallback";
      if (!shells && !sweep) throw std::runtime_error("Unknown gw.qp_grid_search_mode '" + mode + "'");
      if (opt.qp_solver == "fixedpoint")
        ladder_.push_back(std::make_unique<qp_solver::NewtonHunt>(w0_, opt.g_sc_max_iterations, opt.g_sc_limit,
                                                                  opt.qp_solver_alpha));
      // the window, and its part on the physical side of zero (occupied below -margin, virtual above the floor)
      const double lo = w0_ - opt.qp_full_window_half_width, hi = w0_ + opt.qp_full_window_half_width;
      double rlo = lo, rhi = hi;
      if (opt.qp_restrict_search) {
        if (gw_level + opt.qpmin <= opt.homo)
          rhi = std::min(hi, -opt.qp_zero_margin);
        else
          rlo = std::max(lo, opt.qp_virtual_min_energy);
      }
      const bool restricted = (std::abs(rlo - lo) > 1e-12 || std::abs(rhi - hi) > 1e-12) && rlo < rhi;
      auto window = [&](double a, double b, bool allow_rejected) {
        if (shells)
          ladder_.push_back(std::make_unique<qp_solver::ShellHunt>(w0_, a, b, gw_iteration, so, brent, allow_rejected));
        if (sweep) ladder_.push_back(std::make_unique<qp_solver::SweepHunt>(w0_, a, b, so, brent, allow_rejected));
      };
      if (restricted) {
        window(rlo, rhi, false);
        ladder_.push_back(std::make_unique<qp_solver::SweepHunt>(w0_, lo, hi, so, brent, true));
      } else {
        window(lo, hi, true);
      }
    }

    // true when the level is settled; otherwise `ask` holds the frequencies the current rung is blocked on
    bool advance(qp_solver::Ask& ask) {
      while (rung_ < ladder_.size()) {
        if (!ladder_[rung_]->advance(tape_, ask)) return false;
        if (ladder_[rung_]->root()) {
          answer_ = ladder_[rung_]->root();
          converged_ = true;
          return true;
        }
        ++rung_;
      }
      // gw.cc:412-431: w = w0 + (offset - w0 + Sigma_c(w0)) / Z.  The reference forms Z = 1 - fqp.deriv(w0) with
      // fqp.deriv = dSigma_c/dw - 1, i.e. Z = 2 - dSigma_c/dw; kept as it is (results must match the reference's)
      const qp_solver::Sample* s = tape_.find(w0_, true);
      if (!s) {
        ask.want(tape_, w0_, true);
        ask.stage = EvalStage::Other;
        return false;
      }
      const double Z = 1.0 - s->df;
      if (std::abs(Z) > 1e-9) answer_ = w0_ + (offset_ - w0_ + s->sigma) / Z;
      return true;
    }
    // Sigma_c (and dSigma_c/dw when ds != nullptr) at the frequencies of the last ask
    void absorb(const qp_solver::Ask& ask, const double* s, const double* ds) {
      for (std::size_t i = 0; i < ask.w.size(); ++i) {
        qp_solver::Sample smp;
        smp.sigma = s[i];
        smp.f = s[i] + offset_ - ask.w[i];
        if (ds) {
          smp.dsigma = ds[i];
          smp.df = ds[i] - 1.0;
          smp.has_slope = true;
        }
        tape_.record(ask.w[i], smp);
      }
      stats_.Tally(ask.stage, ask.w.size(), ds != nullptr);
    }
    Index level() const { return level_; }
    const std::optional<double>& answer() const { return answer_; }
    bool converged() const { return converged_; }
    const QPStats& stats() const { return stats_; }

   private:
    Index level_;
    double offset_, w0_;
    std::vector<std::unique_ptr<qp_solver::Hunt>> ladder_;
    std::size_t rung_ = 0;
    qp_solver::Tape tape_;
    std::optional<double> answer_;
    bool converged_ = false;
    QPStats stats_;
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

  // gw.cc:323-410.  The reference runs one OpenMP task per level, each calling Sigma_c point by point; here the
  // searches of all levels this rank owns advance together and every round is ONE grouped kernel call
  // (Sigma_base::EvalGroups: a level's Mmn slice is streamed once for all of its frequencies).
  VectorXd SolveQP(const VectorXd& frequencies) {
    sigma_->ResetDiagEvalCounter();
    const Device& dev = Mmn_.device();
    std::vector<LevelSearch> searches;
    for (Index gw_level = 0; gw_level < qptotal_; ++gw_level)
      if (sigma_->OwnsLevel(gw_level))  // multi-GPU: the rank that holds the level's Mmn slice searches its root
        searches.emplace_back(gw_level, dft_energies_(opt_.qpmin + gw_level) + Sigma_x_(gw_level, gw_level) -
                                            vxc_(gw_level, gw_level),
                              frequencies[gw_level], opt_, gw_sc_iteration_);
    std::vector<qp_solver::Ask> asks(searches.size());
    std::vector<char> settled(searches.size(), 0);
    std::size_t rounds = 0, evaluations = 0;
    for (;;) {
      std::vector<int> levels, group_ptr{0};
      std::vector<std::size_t> who;
      std::vector<double> freqs, s, ds;
      bool slope = false;
      for (std::size_t i = 0; i < searches.size(); ++i) {
        if (settled[i]) continue;
        asks[i] = qp_solver::Ask();
        if (searches[i].advance(asks[i])) {
          settled[i] = 1;
          continue;
        }
        who.push_back(i);
        levels.push_back((int)searches[i].level());
        freqs.insert(freqs.end(), asks[i].w.begin(), asks[i].w.end());
        group_ptr.push_back((int)freqs.size());
        slope = slope || asks[i].slope;
      }
      if (who.empty()) break;
      sigma_->EvalGroups(levels, group_ptr, freqs, s, slope ? &ds : nullptr);
      for (std::size_t g = 0; g < who.size(); ++g)
        searches[who[g]].absorb(asks[who[g]], s.data() + group_ptr[g], slope ? ds.data() + group_ptr[g] : nullptr);
      ++rounds;
      evaluations += freqs.size();
    }
    VectorXd frequencies_new = frequencies;
    std::vector<char> converged(qptotal_, 0);
    QPStats total_stats;
    for (const LevelSearch& ls : searches) {
      if (ls.answer()) frequencies_new[ls.level()] = *ls.answer();
      converged[ls.level()] = ls.converged() ? 1 : 0;
      total_stats.Add(ls.stats());
    }
    if (dev.world() > 1) {
      std::vector<double> pack(2 * qptotal_, 0.0);
      for (const LevelSearch& ls : searches) {
        pack[ls.level()] = frequencies_new[ls.level()];
        pack[qptotal_ + ls.level()] = converged[ls.level()] ? 1.0 : 0.0;
      }
      dev.allreduce(pack.data(), pack.size());
      for (Index i = 0; i < qptotal_; ++i) {
        frequencies_new[i] = pack[i];
        converged[i] = pack[qptotal_ + i] > 0.5 ? 1 : 0;
      }
    }
    sigma_batches_ += rounds;
    sigma_evaluations_ += evaluations;
    std::string notconv;
    for (Index lvl = 0; lvl < qptotal_; ++lvl)
      if (!converged[lvl]) notconv += " " + std::to_string(lvl);
    if (!notconv.empty()) {
      log_(" Not converged PQP states are:" + notconv);
      log_(" Increase the grid search interval");
    }
    log_(" Sigma diagonal evaluations in SolveQP: " + std::to_string(evaluations) + " in " + std::to_string(rounds) +
         " batched kernel passes");
    log_(" QP diagnostics: scan=" + std::to_string(total_stats.sigma_scan_calls) +
         " refine=" + std::to_string(total_stats.sigma_refine_calls) +
         " other=" + std::to_string(total_stats.sigma_other_calls) +
         " total_sigma=" + std::to_string(total_stats.TotalSigmaCalls()) +
         " unique_omega=" + std::to_string(total_stats.sigma_unique_frequencies) +
         " deriv_calls=" + std::to_string(total_stats.deriv_calls));
    return frequencies_new;
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
