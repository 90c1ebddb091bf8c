// Quasiparticle root search as an ask / tell protocol.
//
// What it has to reproduce: the decisions of xtp/include/votca/xtp/qp_solver_utils.h (shell scan around a
// linearised centre, 12-fold subdivision of a coarse bracket, bisection / Brent, acceptance by residual and
// renormalisation factor Z, score Z - 0.1 |w - w_ref|) and of xtp/include/votca/xtp/newton_rapson.h.  Root selection
// is discontinuous in Sigma_c, so every frequency that is sampled and every comparison made on the samples is the
// reference's; the form is not.  On a GPU one Sigma_c(w) costs a pass over the level's Mmn slice whether one or a
// hundred frequencies ride along, and a kernel launch whether one or five hundred levels ride along.  So a search
// here never calls a function: it is a resumable object (Hunt) that, given the samples known so far (Tape), either
// finishes or states every frequency it cannot decide without (Ask).  GW::SolveQP advances the hunts of all levels
// in lock step - one batched kernel call per round, no host thread per level.  A hunt may ask for a few samples the
// reference would not take (the look-ahead tree of the interval halving); no decision ever reads one of those.
// The function-object entry points of the reference header (SolveQP_Grid_Windowed, ...) are kept at the bottom as
// drivers of the same objects; tests/test_host_logic_cpu.py runs the reference's known answers
// (test_qp_solver_utils.cc, test_newton_rapson.cc) through them.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <memory>
#include <optional>
#include <stdexcept>
#include <unordered_map>
#include <utility>
#include <vector>

#include "matrix.h"

namespace votca {
namespace xtp {
namespace qp_solver {

// ------------------------------------------------------------------------------------------------------------
// Records of the reference interface (field names are the interface: GW, the options layer and the tests use them)
// ------------------------------------------------------------------------------------------------------------
enum class EvalStage { Scan, Refine, Derivative, Other };

struct SolverOptions {
  double g_sc_limit = 1e-5;
  Index qp_bisection_max_iter = 200;
  double qp_full_window_half_width = 0.75;
  double qp_dense_spacing = 0.002;
  double qp_adaptive_shell_width = 0.025;
  Index qp_adaptive_shell_count = 0;
  double min_accepted_Z = 0.05;
  double max_accepted_Z = 1.5;
};

struct RootCandidate {
  double omega = 0.0, residual = 0.0, deriv = 0.0, Z = 0.0, distance_to_ref = 0.0;
  bool accepted = false;
};

struct WindowDiagnostics {
  Index shells_explored = 0, first_interval_shell = -1, first_accepted_shell = -1, chosen_shell = -1;
  Index intervals_found = 0;
};

// evaluation counters printed by GW::SolveQP (the reference keeps the same tallies, gw.h:261-267)
struct Stats {
  std::size_t sigma_scan_calls = 0, sigma_refine_calls = 0, sigma_derivative_calls = 0, sigma_other_calls = 0;
  std::size_t sigma_repeat_calls = 0, sigma_unique_frequencies = 0, deriv_calls = 0;
  void Tally(EvalStage stage, std::size_t n, bool with_slope) {
    std::size_t& slot = stage == EvalStage::Scan     ? sigma_scan_calls
                        : stage == EvalStage::Refine ? sigma_refine_calls
                        : stage == EvalStage::Derivative ? sigma_derivative_calls
                                                         : sigma_other_calls;
    slot += n;
    sigma_unique_frequencies += n;
    if (with_slope) deriv_calls += n;
  }
  void Add(const Stats& o) {
    const std::size_t* src[] = {&o.sigma_scan_calls,   &o.sigma_refine_calls,       &o.sigma_derivative_calls,
                                &o.sigma_other_calls,  &o.sigma_repeat_calls,       &o.sigma_unique_frequencies,
                                &o.deriv_calls};
    std::size_t* dst[] = {&sigma_scan_calls,  &sigma_refine_calls,  &sigma_derivative_calls,  &sigma_other_calls,
                          &sigma_repeat_calls, &sigma_unique_frequencies, &deriv_calls};
    for (int i = 0; i < 7; ++i) *dst[i] += *src[i];
  }
  std::size_t TotalSigmaCalls() const {
    return sigma_scan_calls + sigma_refine_calls + sigma_derivative_calls + sigma_other_calls;
  }
};

// ------------------------------------------------------------------------------------------------------------
// Option normalisation (qp_solver_utils.h:113-192): the pre-2025 pair (qp_grid_steps, qp_grid_spacing) still
// defines the window when the newer keys are unset
// ------------------------------------------------------------------------------------------------------------
namespace detail {
template <typename Opt>
bool legacy_grid(const Opt& o, double* width) {
  if (o.qp_grid_steps <= 1 || o.qp_grid_spacing <= 0.0) return false;
  *width = o.qp_grid_spacing * double(o.qp_grid_steps - 1);
  return true;
}
}  // namespace detail

template <typename Opt>
inline double LegacyFullWindowHalfWidth(const Opt& opt) {
  double width;
  return detail::legacy_grid(opt, &width) ? 0.5 * width : -1.0;
}

template <typename Opt>
inline double LegacyAdaptiveShellWidth(const Opt& opt) {
  double width;
  if (!detail::legacy_grid(opt, &width)) return -1.0;
  const Index coarse = std::max<Index>(21, opt.qp_grid_steps / 4);  // points of the old coarse pre-scan
  return coarse > 1 ? width / double(coarse - 1) : 4.0 * opt.qp_grid_spacing;
}

template <typename Opt>
inline void NormalizeGridSearchOptions(Opt& opt) {
  double unused;
  const bool legacy = detail::legacy_grid(opt, &unused);
  const bool shells_by_count = opt.qp_adaptive_shell_count > 0;
  if (!(opt.qp_full_window_half_width > 0.0)) opt.qp_full_window_half_width = legacy ? LegacyFullWindowHalfWidth(opt) : 0.75;
  if (!(opt.qp_dense_spacing > 0.0)) opt.qp_dense_spacing = legacy ? opt.qp_grid_spacing : 0.002;
  if (!shells_by_count && !(opt.qp_adaptive_shell_width > 0.0))
    opt.qp_adaptive_shell_width = legacy ? LegacyAdaptiveShellWidth(opt) : 0.025;
  const char* complaint = nullptr;
  if (!(opt.qp_full_window_half_width > 0.0))
    complaint = "qp_full_window_half_width must be > 0";
  else if (!(opt.qp_dense_spacing > 0.0))
    complaint = "qp_dense_spacing must be > 0";
  else if (!shells_by_count && !(opt.qp_adaptive_shell_width > 0.0))
    complaint = "need qp_adaptive_shell_width > 0 or qp_adaptive_shell_count > 0";
  if (complaint) throw std::runtime_error(std::string("Invalid QP search setup: ") + complaint);
}

inline double EffectiveAdaptiveShellWidth(const SolverOptions& opt) {
  return opt.qp_adaptive_shell_count > 0 ? opt.qp_full_window_half_width / double(opt.qp_adaptive_shell_count)
                                         : opt.qp_adaptive_shell_width;
}

// acceptance window of a root (qp_solver_utils.h:309-327) and its score (:329-332)
inline bool AcceptRoot(const RootCandidate& cand, const SolverOptions& opt) {
  const bool finite = std::isfinite(cand.omega) && std::isfinite(cand.Z);
  const bool on_root = std::abs(cand.residual) <= opt.g_sc_limit;
  const bool physical_Z = cand.Z > 0.0 && cand.Z >= opt.min_accepted_Z && cand.Z <= opt.max_accepted_Z;
  return finite && on_root && physical_Z;
}

inline double ScoreRoot(const RootCandidate& cand) { return cand.Z - 0.1 * cand.distance_to_ref; }

// highest score, the earliest found among equals
inline const RootCandidate& BestRoot(const std::vector<RootCandidate>& found) {
  const RootCandidate* best = &found.front();
  for (const RootCandidate& c : found)
    if (ScoreRoot(c) > ScoreRoot(*best)) best = &c;
  return *best;
}

// ------------------------------------------------------------------------------------------------------------
// Tape: every sample of f(w) = Sigma_c(w) + offset - w a level has been told, keyed by the bit pattern of w
// (two hunts of one level - restricted window, then the full one - share it, so nothing is evaluated twice)
// ------------------------------------------------------------------------------------------------------------
struct Sample {
  double f = 0.0, df = 0.0;          // residual and its slope dSigma/dw - 1
  double sigma = 0.0, dsigma = 0.0;  // raw self-energy (linearisation step of GW); unused by the hunts
  bool has_slope = false;
};

class Tape {
 public:
  const Sample* find(double w, bool need_slope = false) const {
    auto it = samples_.find(key(w));
    if (it == samples_.end() || (need_slope && !it->second.has_slope)) return nullptr;
    return &it->second;
  }
  const Sample& at(double w) const {
    const Sample* s = find(w);
    if (!s) throw std::logic_error("QP search: sample missing from the tape");
    return *s;
  }
  void record(double w, const Sample& s) {
    Sample& slot = samples_[key(w)];
    const bool keep_slope = slot.has_slope && !s.has_slope;
    const double df = slot.df, ds = slot.dsigma;
    slot = s;
    if (keep_slope) {
      slot.df = df;
      slot.dsigma = ds;
      slot.has_slope = true;
    }
  }
  std::size_t size() const { return samples_.size(); }

 private:
  static std::uint64_t key(double w) {
    std::uint64_t k;
    std::memcpy(&k, &w, sizeof k);
    return k;
  }
  std::unordered_map<std::uint64_t, Sample> samples_;
};

// What a hunt is blocked on.  One stage label and one slope flag per round: the kernel returns dSigma/dw for a
// whole batch or not at all.
struct Ask {
  std::vector<double> w;
  bool slope = false;
  EvalStage stage = EvalStage::Other;
  void want(const Tape& tape, double x, bool with_slope = false) {
    if (tape.find(x, with_slope)) return;
    if (std::find(w.begin(), w.end(), x) == w.end()) w.push_back(x);
    slope = slope || with_slope;
  }
  bool empty() const { return w.empty(); }
};

// A resumable search.  advance() consumes what the tape holds; false = blocked, `ask` says on what.
class Hunt {
 public:
  virtual ~Hunt() = default;
  virtual bool advance(const Tape& tape, Ask& ask) = 0;
  const std::optional<double>& root() const { return root_; }

 protected:
  std::optional<double> root_;
};

// ------------------------------------------------------------------------------------------------------------
// Bracket polishers
// ------------------------------------------------------------------------------------------------------------
struct Bracket {
  double lo = 0.0, f_lo = 0.0, hi = 0.0, f_hi = 0.0;
  double mid() const { return 0.5 * (lo + hi); }
  bool straddles() const { return !(f_lo * f_hi > 0.0); }
};

// Interval halving (qp_solver_utils.h:194-224).  The midpoints of the next three halvings form a binary tree of
// seven points whatever the signs turn out to be; asking for the tree costs one round instead of three.
class Halving {
 public:
  Halving(const Bracket& b, double tol) : b_(b), tol_(tol) {
    if (!b.straddles()) throw std::runtime_error("Bisection needs a positive and negative function value");
  }
  bool advance(const Tape& tape, Ask& ask) {
    for (;;) {
      const double c = b_.mid();
      if (std::abs(b_.hi - b_.lo) < tol_) return settle(c);
      const Sample* s = tape.find(c);
      if (!s) {
        lookahead(tape, ask, b_.lo, b_.hi, 3);
        return false;
      }
      if (std::abs(s->f) < tol_) return settle(c);
      if (s->f * b_.f_lo > 0.0) {
        b_.lo = c;
        b_.f_lo = s->f;
      } else {
        b_.hi = c;
        b_.f_hi = s->f;
      }
    }
  }
  double root() const { return root_; }

 private:
  bool settle(double c) {
    root_ = c;
    return true;
  }
  static void lookahead(const Tape& tape, Ask& ask, double lo, double hi, int depth) {
    if (depth == 0) return;
    const double c = 0.5 * (lo + hi);
    ask.want(tape, c);
    lookahead(tape, ask, lo, c, depth - 1);
    lookahead(tape, ask, c, hi, depth - 1);
  }
  Bracket b_;
  double tol_, root_ = 0.0;
};

// Brent's method (qp_solver_utils.h:226-307; Numerical Recipes' zbrent with an absolute tolerance): strictly
// sequential inside one bracket, so a round asks for one point - the brackets of all levels still move together.
class BrentSteps {
 public:
  BrentSteps(const Bracket& br, double tol, Index max_iter)
      : a_(br.lo), b_(br.hi), c_(br.lo), fa_(br.f_lo), fb_(br.f_hi), fc_(br.f_lo), tol_(tol), max_iter_(max_iter) {
    if (!br.straddles()) throw std::runtime_error("Brent needs a positive and negative function value");
    d_ = e_ = b_ - a_;
  }
  bool advance(const Tape& tape, Ask& ask) {
    for (;;) {
      if (awaiting_) {
        const Sample* s = tape.find(b_);
        if (!s) {
          ask.want(tape, b_);
          return false;
        }
        fb_ = s->f;
        awaiting_ = false;
        ++done_;
      }
      if (done_ >= max_iter_) throw std::runtime_error("Brent did not converge within qp_bisection_max_iter");
      if ((fb_ > 0.0 && fc_ > 0.0) || (fb_ < 0.0 && fc_ < 0.0)) {  // root no longer between b and c
        c_ = a_;
        fc_ = fa_;
        d_ = e_ = b_ - a_;
      }
      if (std::abs(fc_) < std::abs(fb_)) {  // keep b the better of the two ends
        a_ = b_;
        fa_ = fb_;
        b_ = c_;
        fb_ = fc_;
        c_ = a_;
        fc_ = fa_;
      }
      const double half = 0.5 * (c_ - b_);
      if (std::abs(half) < tol_ || std::abs(fb_) < tol_) return true;
      double move = half;
      bool interpolated = false;
      if (std::abs(e_) >= tol_ && std::abs(fa_) > std::abs(fb_)) {
        const double s = fb_ / fa_;
        double p, q;
        if (a_ == c_) {  // secant
          p = 2.0 * half * s;
          q = 1.0 - s;
        } else {  // inverse quadratic
          const double u = fa_ / fc_, r = fb_ / fc_;
          p = s * (2.0 * half * u * (u - r) - (b_ - a_) * (r - 1.0));
          q = (u - 1.0) * (r - 1.0) * (s - 1.0);
        }
        if (p > 0.0) q = -q;
        p = std::abs(p);
        if (q != 0.0 && 2.0 * p < std::min(3.0 * half * q - std::abs(tol_ * q), std::abs(e_ * q))) {
          e_ = d_;
          d_ = move = p / q;
          interpolated = true;
        }
      }
      if (!interpolated) d_ = e_ = half;
      a_ = b_;
      fa_ = fb_;
      b_ += std::abs(move) > tol_ ? move : (half > 0.0 ? tol_ : -tol_);
      awaiting_ = true;
    }
  }
  double root() const { return b_; }

 private:
  double a_, b_, c_, fa_, fb_, fc_, d_ = 0.0, e_ = 0.0, tol_;
  Index max_iter_, done_ = 0;
  bool awaiting_ = false;
};

// One bracket -> one root candidate (qp_solver_utils.h:334-378): polish, then residual and slope at the root
class Probe {
 public:
  Probe(const Bracket& b, double reference, const SolverOptions& opt, bool brent, Index shell = 0)
      : shell_(shell), reference_(reference), opt_(opt) {
    if (b.f_lo * b.f_hi > 0.0) {
      // no sign change: only an end point that already is a root (within g_sc_limit) yields a candidate
      const bool lo_zero = std::abs(b.f_lo) <= opt.g_sc_limit, hi_zero = std::abs(b.f_hi) <= opt.g_sc_limit;
      if (!lo_zero && !hi_zero) {
        over_ = true;
        return;
      }
      omega_ = std::abs(b.f_lo) <= std::abs(b.f_hi) ? b.lo : b.hi;
      located_ = true;
    } else if (brent) {
      brent_ = std::make_unique<BrentSteps>(b, opt.g_sc_limit, opt.qp_bisection_max_iter);
    } else {
      halving_ = std::make_unique<Halving>(b, opt.g_sc_limit);
    }
  }
  bool advance(const Tape& tape, Ask& ask) {
    if (over_) return true;
    if (!located_) {
      if (brent_ ? !brent_->advance(tape, ask) : !halving_->advance(tape, ask)) return false;
      omega_ = brent_ ? brent_->root() : halving_->root();
      located_ = true;
    }
    const Sample* s = tape.find(omega_, true);
    if (!s) {
      ask.want(tape, omega_, true);
      return false;
    }
    RootCandidate c;
    c.omega = omega_;
    c.residual = s->f;
    c.deriv = s->df;
    c.Z = std::abs(s->df) > 1e-14 ? -1.0 / s->df : std::numeric_limits<double>::infinity();
    c.distance_to_ref = std::abs(omega_ - reference_);
    c.accepted = AcceptRoot(c, opt_);
    found_ = c;
    over_ = true;
    return true;
  }
  const std::optional<RootCandidate>& candidate() const { return found_; }
  Index shell() const { return shell_; }

 private:
  Index shell_;
  double reference_, omega_ = 0.0;
  SolverOptions opt_;
  bool located_ = false, over_ = false;
  std::unique_ptr<Halving> halving_;
  std::unique_ptr<BrentSteps> brent_;
  std::optional<RootCandidate> found_;
};

namespace detail {
// advance every probe that is still open; true when all are closed
inline bool advance_all(std::vector<Probe>& probes, const Tape& tape, Ask& ask) {
  bool all = true;
  for (Probe& p : probes) all = p.advance(tape, ask) && all;
  if (!all) ask.stage = EvalStage::Refine;
  return all;
}
}  // namespace detail

// ------------------------------------------------------------------------------------------------------------
// ShellHunt: qp_solver_utils.h:380-640.  Rounds: (seed) value + slope at the start frequency -> centre;
// (shells) every shell point and both window ends; (split) eleven interior points of every coarse bracket;
// (polish) all brackets together; then the verdict.
// ------------------------------------------------------------------------------------------------------------
class ShellHunt : public Hunt {
 public:
  ShellHunt(double frequency0, double left, double right, Index gw_iteration, const SolverOptions& opt, bool brent,
            bool allow_rejected = true)
      : w0_(frequency0), left_(left), right_(right), first_iteration_(gw_iteration == 0), opt_(opt), brent_(brent),
        allow_rejected_(allow_rejected), width_(EffectiveAdaptiveShellWidth(opt)) {
    if (!(left < right)) phase_ = Phase::Over;
  }

  bool advance(const Tape& tape, Ask& ask) override {
    for (;;) switch (phase_) {
        case Phase::Seed:
          if (first_iteration_ && !tape.find(w0_, true)) {
            ask.want(tape, w0_, true);
            ask.stage = EvalStage::Other;
            return false;
          }
          place_centre(tape);
          phase_ = Phase::Shells;
          break;
        case Phase::Shells:
          for (double x : shell_points_) ask.want(tape, x);
          if (!ask.empty()) {
            ask.stage = EvalStage::Scan;
            return false;
          }
          collect_sign_changes(tape);
          phase_ = Phase::Split;
          break;
        case Phase::Split:
          for (const Coarse& c : coarse_)
            for (Index i = 1; i < kSub; ++i) ask.want(tape, c.lo + double(i) * c.dx());
          if (!ask.empty()) {
            ask.stage = EvalStage::Scan;
            return false;
          }
          for (const Coarse& c : coarse_) probes_.emplace_back(narrow(tape, c), w0_, opt_, brent_, c.shell);
          phase_ = Phase::Polish;
          break;
        case Phase::Polish:
          if (!detail::advance_all(probes_, tape, ask)) return false;
          verdict();
          phase_ = Phase::Over;
          break;
        case Phase::Over:
          return true;
      }
  }
  const WindowDiagnostics& diagnostics() const { return diag_; }
  const std::vector<RootCandidate>& accepted() const { return accepted_; }
  const std::vector<RootCandidate>& rejected() const { return rejected_; }

 private:
  static constexpr Index kSub = 12;
  enum class Phase { Seed, Shells, Split, Polish, Over };
  struct Coarse {
    double lo, f_lo, hi, f_hi;
    Index shell;
    double dx() const { return (hi - lo) / double(kSub); }
  };

  // first GW iteration: one Newton step from the start frequency picks the centre of the shells
  void place_centre(const Tape& tape) {
    centre_ = w0_;
    if (first_iteration_) {
      const Sample& s = tape.at(w0_);
      if (std::isfinite(s.f) && std::isfinite(s.df) && std::abs(s.df) > 1e-6) {
        const double w_lin = w0_ - s.f / s.df;
        if (std::isfinite(w_lin) && w_lin >= left_ && w_lin <= right_) centre_ = w_lin;
      }
    }
    centre_ = std::max(left_, std::min(right_, centre_));
    const double reach = std::max(centre_ - left_, right_ - centre_);
    n_shells_ = static_cast<Index>(std::ceil(reach / width_));
    shell_points_.push_back(centre_);
    for (Index k = 1; k <= n_shells_; ++k) {
      const double delta = double(k) * width_;
      if (centre_ - delta >= left_) shell_points_.push_back(centre_ - delta);
      if (centre_ + delta <= right_) shell_points_.push_back(centre_ + delta);
    }
    shell_points_.push_back(left_);
    shell_points_.push_back(right_);
  }

  // walk outwards shell by shell, left arm before right arm, then the two window ends: every strict sign change
  // between neighbours on an arm is a coarse bracket, in the order the reference meets them
  void collect_sign_changes(const Tape& tape) {
    struct Arm {
      double w, f;
      bool open = true;
    };
    const double fc = tape.at(centre_).f;
    Arm arm[2] = {{centre_, fc}, {centre_, fc}};  // 0: towards lower, 1: towards higher frequencies
    auto bracket = [&](double w_in, double f_in, double w_out, double f_out, Index shell) {
      if (f_in * f_out < 0.0)
        coarse_.push_back(w_in < w_out ? Coarse{w_in, f_in, w_out, f_out, shell} : Coarse{w_out, f_out, w_in, f_in, shell});
    };
    for (Index k = 1; k <= n_shells_; ++k) {
      diag_.shells_explored = k;
      const double delta = double(k) * width_;
      bool moved = false;
      for (int side = 0; side < 2; ++side) {
        if (!arm[side].open) continue;
        const double w = side == 0 ? centre_ - delta : centre_ + delta;
        if (side == 0 ? w < left_ : w > right_) {
          arm[side].open = false;
          continue;
        }
        const double f = tape.at(w).f;
        bracket(arm[side].w, arm[side].f, w, f, k);
        arm[side].w = w;
        arm[side].f = f;
        moved = true;
      }
      if (!moved && !arm[0].open && !arm[1].open) break;
    }
    const Index beyond = diag_.shells_explored + 1;
    if (arm[0].w > left_ + 1e-12) bracket(arm[0].w, arm[0].f, left_, tape.at(left_).f, beyond);
    if (arm[1].w < right_ - 1e-12) bracket(arm[1].w, arm[1].f, right_, tape.at(right_).f, beyond);
  }

  // among the sub-intervals of a coarse bracket that hold a sign change or touch a root, the one closest to the
  // centre (ties: the lower one); the coarse bracket itself if there is none
  Bracket narrow(const Tape& tape, const Coarse& c) const {
    Bracket pick{c.lo, c.f_lo, c.hi, c.f_hi};
    if (!(c.hi > c.lo)) return pick;
    bool have = false;
    double pick_dist = 0.0;
    auto offer = [&](const Bracket& b) {
      const double dist = std::abs(b.mid() - centre_);
      const bool closer = dist < pick_dist - 1e-14;
      const bool level_but_lower = std::abs(dist - pick_dist) <= 1e-14 && b.lo < pick.lo;
      if (!have || closer || level_but_lower) {
        pick = b;
        pick_dist = dist;
        have = true;
      }
    };
    double x0 = c.lo, f0 = c.f_lo;
    for (Index i = 1; i <= kSub; ++i) {
      const bool last = i == kSub;
      const double x1 = last ? c.hi : c.lo + double(i) * c.dx();
      const double f1 = last ? c.f_hi : tape.at(x1).f;
      const bool crosses = (f0 < 0.0 && f1 > 0.0) || (f0 > 0.0 && f1 < 0.0);
      const bool touches = x0 < x1 && (std::abs(f0) <= opt_.g_sc_limit || std::abs(f1) <= opt_.g_sc_limit);
      if (crosses || touches) offer(Bracket{x0, f0, x1, f1});
      x0 = x1;
      f0 = f1;
    }
    return pick;
  }

  void verdict() {
    for (const Probe& p : probes_) {
      if (!p.candidate()) continue;
      if (diag_.first_interval_shell < 0) diag_.first_interval_shell = p.shell();
      ++diag_.intervals_found;
      if (p.candidate()->accepted) {
        if (diag_.first_accepted_shell < 0) diag_.first_accepted_shell = p.shell();
        accepted_.push_back(*p.candidate());
      } else {
        rejected_.push_back(*p.candidate());
      }
    }
    const std::vector<RootCandidate>& pool = !accepted_.empty() ? accepted_ : rejected_;
    if (pool.empty()) return;
    const RootCandidate& best = BestRoot(pool);
    diag_.chosen_shell = static_cast<int>(std::llround(std::abs(best.omega - centre_) / width_));
    if (!accepted_.empty() || allow_rejected_) root_ = best.omega;
  }

  double w0_, left_, right_;
  bool first_iteration_;
  SolverOptions opt_;
  bool brent_, allow_rejected_;
  double width_, centre_ = 0.0;
  Index n_shells_ = 0;
  Phase phase_ = Phase::Seed;
  std::vector<double> shell_points_;
  std::vector<Coarse> coarse_;
  std::vector<Probe> probes_;
  std::vector<RootCandidate> accepted_, rejected_;
  WindowDiagnostics diag_;
};

// ------------------------------------------------------------------------------------------------------------
// SweepHunt: the dense scan of gw.cc:504-616 - equidistant nodes across the window, every strict sign change
// between neighbours polished directly.
// ------------------------------------------------------------------------------------------------------------
class SweepHunt : public Hunt {
 public:
  SweepHunt(double frequency0, double left, double right, const SolverOptions& opt, bool brent, bool allow_rejected)
      : w0_(frequency0), opt_(opt), brent_(brent), allow_rejected_(allow_rejected) {
    if (!(left < right)) {
      scanned_ = polished_ = true;
      return;
    }
    const Index n = std::max<Index>(2, static_cast<Index>(std::ceil((right - left) / opt.qp_dense_spacing)) + 1);
    nodes_.reserve(n);
    for (Index i = 0; i + 1 < n; ++i) nodes_.push_back(std::min(right, left + double(i) * opt.qp_dense_spacing));
    nodes_.push_back(right);
  }
  bool advance(const Tape& tape, Ask& ask) override {
    if (!scanned_) {
      for (double x : nodes_) ask.want(tape, x);
      if (!ask.empty()) {
        ask.stage = EvalStage::Scan;
        return false;
      }
      for (std::size_t i = 1; i < nodes_.size(); ++i) {
        const double f0 = tape.at(nodes_[i - 1]).f, f1 = tape.at(nodes_[i]).f;
        if (f0 * f1 < 0.0) probes_.emplace_back(Bracket{nodes_[i - 1], f0, nodes_[i], f1}, w0_, opt_, brent_);
      }
      scanned_ = true;
    }
    if (!polished_) {
      if (!detail::advance_all(probes_, tape, ask)) return false;
      std::vector<RootCandidate> good, bad;
      for (const Probe& p : probes_)
        if (p.candidate()) (p.candidate()->accepted ? good : bad).push_back(*p.candidate());
      if (!good.empty())
        root_ = BestRoot(good).omega;
      else if (!bad.empty() && allow_rejected_)
        root_ = BestRoot(bad).omega;
      polished_ = true;
    }
    return true;
  }

 private:
  double w0_;
  SolverOptions opt_;
  bool brent_, allow_rejected_, scanned_ = false, polished_ = false;
  std::vector<double> nodes_;
  std::vector<Probe> probes_;
};

// ------------------------------------------------------------------------------------------------------------
// NewtonHunt: damped Newton iteration on f (newton_rapson.h:39-90 as driven by gw.cc:741-757)
// ------------------------------------------------------------------------------------------------------------
class NewtonHunt : public Hunt {
 public:
  enum Outcome { success, smalldenom, notconverged };
  NewtonHunt(double x0, Index max_iterations, double tolerance, double alpha)
      : x_(x0), left_(max_iterations), tol_(tolerance), alpha_(alpha) {}
  bool advance(const Tape& tape, Ask& ask) override {
    while (outcome_ == notconverged && left_ > 0) {
      const Sample* s = tape.find(x_, true);
      if (!s) {
        ask.want(tape, x_, true);
        ask.stage = EvalStage::Other;
        return false;
      }
      --left_;
      if (std::abs(s->df) < 1e-12) {
        outcome_ = smalldenom;
      } else {
        const double step = -alpha_ * s->f / s->df;
        if (std::abs(step) < tol_)
          outcome_ = success;
        else
          x_ += step;
      }
    }
    if (outcome_ == success) root_ = x_;
    return true;
  }
  Outcome outcome() const { return outcome_; }
  double last() const { return x_; }

 private:
  double x_;
  Index left_;
  double tol_, alpha_;
  Outcome outcome_ = notconverged;
};

// ------------------------------------------------------------------------------------------------------------
// Drivers for a plain function object (value(w, stage), deriv(w)): the reference's call surface
// ------------------------------------------------------------------------------------------------------------
template <typename F, typename Machine>
void Drive(Machine& m, Tape& tape, const F& f) {
  Ask ask;
  while (!m.advance(tape, ask)) {
    for (double w : ask.w) {
      Sample s;
      s.f = f.value(w, ask.stage);
      if (ask.slope) {
        s.df = f.deriv(w);
        s.has_slope = true;
      }
      tape.record(w, s);
    }
    ask = Ask();
  }
}

template <typename F>
double SolveQP_Bisection(double lowerbound, double f_lowerbound, double upperbound, double f_upperbound, const F& f,
                         const SolverOptions& opt) {
  Halving h(Bracket{lowerbound, f_lowerbound, upperbound, f_upperbound}, opt.g_sc_limit);
  Tape tape;
  Drive(h, tape, f);
  return h.root();
}

template <typename F>
double SolveQP_Brent(double lowerbound, double f_lowerbound, double upperbound, double f_upperbound, const F& f,
                     const SolverOptions& opt) {
  BrentSteps b(Bracket{lowerbound, f_lowerbound, upperbound, f_upperbound}, opt.g_sc_limit, opt.qp_bisection_max_iter);
  Tape tape;
  Drive(b, tape, f);
  return b.root();
}

template <typename F>
std::optional<RootCandidate> RefineQPInterval(double lowerbound, double f_lowerbound, double upperbound,
                                              double f_upperbound, const F& f, double reference,
                                              const SolverOptions& opt, bool use_brent) {
  Probe p(Bracket{lowerbound, f_lowerbound, upperbound, f_upperbound}, reference, opt, use_brent);
  Tape tape;
  Drive(p, tape, f);
  return p.candidate();
}

template <typename F>
std::optional<double> SolveQP_Grid_Windowed(F& fqp, double frequency0, double left_limit, double right_limit,
                                            Index gw_sc_iteration, const SolverOptions& opt,
                                            WindowDiagnostics* wdiag = nullptr,
                                            std::vector<RootCandidate>* accepted_roots_out = nullptr,
                                            std::vector<RootCandidate>* rejected_roots_out = nullptr,
                                            bool use_brent = false) {
  ShellHunt hunt(frequency0, left_limit, right_limit, gw_sc_iteration, opt, use_brent);
  Tape tape;
  Drive(hunt, tape, fqp);
  if (wdiag) *wdiag = hunt.diagnostics();
  if (accepted_roots_out) *accepted_roots_out = hunt.accepted();
  if (rejected_roots_out) *rejected_roots_out = hunt.rejected();
  return hunt.root();
}

}  // namespace qp_solver

// newton_rapson.h:39-90 for a functor returning (f, f'): the same iteration as NewtonHunt, driven directly
template <class Func>
class NewtonRapson {
 public:
  enum Errors { success, smalldenom, notconverged };
  NewtonRapson(Index max_iterations, double tolerance, double alpha = 1.0)
      : max_iterations_(max_iterations), tolerance_(tolerance), alpha_(alpha) {}
  double FindRoot(const Func& f, double x0) {
    qp_solver::NewtonHunt hunt(x0, max_iterations_, tolerance_, alpha_);
    qp_solver::Tape tape;
    qp_solver::Ask ask;
    while (!hunt.advance(tape, ask)) {
      for (double x : ask.w) {
        const std::pair<double, double> v = f(x);
        qp_solver::Sample s;
        s.f = v.first;
        s.df = v.second;
        s.has_slope = true;
        tape.record(x, s);
      }
      ask = qp_solver::Ask();
    }
    info_ = static_cast<Errors>(hunt.outcome());
    return hunt.last();
  }
  Errors getInfo() const { return info_; }

 private:
  Errors info_ = notconverged;
  Index max_iterations_;
  double tolerance_, alpha_;
};

}  // namespace xtp
}  // namespace votca
