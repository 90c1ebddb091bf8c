// Quasiparticle root search - host control flow of xtp/include/votca/xtp/qp_solver_utils.h (whole file)
// and xtp/include/votca/xtp/newton_rapson.h:39-90.  Root selection is discontinuous in Sigma_c, so the
// decision logic is kept behaviourally identical (same scan points, same bracketing, same acceptance
// and scoring); only the Sigma_c evaluations behind QPFunc run on the GPU.
#pragma once
#include <cmath>
#include <limits>
#include <optional>
#include <stdexcept>
#include <utility>
#include <vector>

#include "matrix.h"

namespace votca {
namespace xtp {
namespace qp_solver {

enum class EvalStage { Scan, Refine, Derivative, Other };

struct Stats {
  std::size_t sigma_scan_calls = 0, sigma_refine_calls = 0, sigma_derivative_calls = 0, sigma_other_calls = 0;
  std::size_t sigma_repeat_calls = 0, sigma_unique_frequencies = 0, deriv_calls = 0;
  void Add(const Stats& o) {
    sigma_scan_calls += o.sigma_scan_calls;
    sigma_refine_calls += o.sigma_refine_calls;
    sigma_derivative_calls += o.sigma_derivative_calls;
    sigma_other_calls += o.sigma_other_calls;
    sigma_repeat_calls += o.sigma_repeat_calls;
    sigma_unique_frequencies += o.sigma_unique_frequencies;
    deriv_calls += o.deriv_calls;
  }
  std::size_t TotalSigmaCalls() const {
    return sigma_scan_calls + sigma_refine_calls + sigma_derivative_calls + sigma_other_calls;
  }
};

struct RootCandidate {
  double omega = 0.0, residual = 0.0, deriv = 0.0, Z = 0.0, distance_to_ref = 0.0;
  bool accepted = false;
};

struct WindowDiagnostics {
  Index shells_explored = 0, first_interval_shell = -1, first_accepted_shell = -1, chosen_shell = -1;
  Index intervals_found = 0;
};

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

template <typename Opt>
inline double LegacyFullWindowHalfWidth(const Opt& opt) {
  if (opt.qp_grid_steps <= 1 || opt.qp_grid_spacing <= 0.0) return -1.0;
  return 0.5 * opt.qp_grid_spacing * double(opt.qp_grid_steps - 1);
}

template <typename Opt>
inline double LegacyAdaptiveShellWidth(const Opt& opt) {
  if (opt.qp_grid_steps <= 1 || opt.qp_grid_spacing <= 0.0) return -1.0;
  const double full_window_width = opt.qp_grid_spacing * double(opt.qp_grid_steps - 1);
  const Index base_coarse_steps = std::max<Index>(21, opt.qp_grid_steps / 4);
  if (base_coarse_steps <= 1) return 4.0 * opt.qp_grid_spacing;
  return full_window_width / double(base_coarse_steps - 1);
}

template <typename Opt>
inline void NormalizeGridSearchOptions(Opt& opt) {
  const bool has_legacy = (opt.qp_grid_steps > 1 && opt.qp_grid_spacing > 0.0);
  if (opt.qp_full_window_half_width <= 0.0)
    opt.qp_full_window_half_width = has_legacy ? LegacyFullWindowHalfWidth(opt) : 0.75;
  if (opt.qp_dense_spacing <= 0.0) opt.qp_dense_spacing = has_legacy ? opt.qp_grid_spacing : 0.002;
  if (opt.qp_adaptive_shell_count <= 0 && opt.qp_adaptive_shell_width <= 0.0)
    opt.qp_adaptive_shell_width = has_legacy ? LegacyAdaptiveShellWidth(opt) : 0.025;
  if (opt.qp_full_window_half_width <= 0.0)
    throw std::runtime_error("Invalid QP search setup: qp_full_window_half_width must be > 0");
  if (opt.qp_dense_spacing <= 0.0) throw std::runtime_error("Invalid QP search setup: qp_dense_spacing must be > 0");
  if (opt.qp_adaptive_shell_count <= 0 && opt.qp_adaptive_shell_width <= 0.0)
    throw std::runtime_error(
        "Invalid QP search setup: need qp_adaptive_shell_width > 0 or qp_adaptive_shell_count > 0");
}

inline double EffectiveAdaptiveShellWidth(const SolverOptions& opt) {
  if (opt.qp_adaptive_shell_count > 0)
    return opt.qp_full_window_half_width / static_cast<double>(opt.qp_adaptive_shell_count);
  return opt.qp_adaptive_shell_width;
}

template <typename QPFunc>
double SolveQP_Bisection(double lowerbound, double f_lowerbound, double upperbound, double f_upperbound,
                         const QPFunc& f, const SolverOptions& opt) {
  if (f_lowerbound * f_upperbound > 0.0)
    throw std::runtime_error("Bisection needs a positive and negative function value");
  Index step = 0;
  while (true) {
    const double c = 0.5 * (lowerbound + upperbound);
    if (std::abs(upperbound - lowerbound) < opt.g_sc_limit) return c;
    if (step++ % 3 == 0) {
      // the next three midpoints lie in this binary tree whatever the signs turn out to be
      double pts[7];
      pts[0] = c;
      pts[1] = 0.5 * (lowerbound + c);
      pts[2] = 0.5 * (c + upperbound);
      pts[3] = 0.5 * (lowerbound + pts[1]);
      pts[4] = 0.5 * (pts[1] + c);
      pts[5] = 0.5 * (c + pts[2]);
      pts[6] = 0.5 * (pts[2] + upperbound);
      f.prefetch(pts, 7);
    }
    const double y_c = f.value(c, EvalStage::Refine);
    if (std::abs(y_c) < opt.g_sc_limit) return c;
    if (y_c * f_lowerbound > 0.0) {
      lowerbound = c;
      f_lowerbound = y_c;
    } else {
      upperbound = c;
      f_upperbound = y_c;
    }
  }
}

template <typename QPFunc>
double SolveQP_Brent(double lowerbound, double f_lowerbound, double upperbound, double f_upperbound, const QPFunc& f,
                     const SolverOptions& opt) {
  if (f_lowerbound * f_upperbound > 0.0) throw std::runtime_error("Brent needs a positive and negative function value");
  double a = lowerbound, b = upperbound, fa = f_lowerbound, fb = f_upperbound;
  double c = a, fc = fa;
  double d = b - a, e = d;
  for (Index iter = 0; iter < opt.qp_bisection_max_iter; ++iter) {
    if ((fb > 0.0 && fc > 0.0) || (fb < 0.0 && fc < 0.0)) {
      c = a;
      fc = fa;
      d = b - a;
      e = d;
    }
    if (std::abs(fc) < std::abs(fb)) {
      a = b;
      b = c;
      c = a;
      fa = fb;
      fb = fc;
      fc = fa;
    }
    const double tol = opt.g_sc_limit;
    const double m = 0.5 * (c - b);
    if (std::abs(m) < tol || std::abs(fb) < opt.g_sc_limit) return b;
    if (std::abs(e) >= tol && std::abs(fa) > std::abs(fb)) {
      double s = fb / fa, p = 0.0, q = 0.0;
      if (a == c) {
        p = 2.0 * m * s;
        q = 1.0 - s;
      } else {
        double q1 = fa / fc, r = fb / fc;
        p = s * (2.0 * m * q1 * (q1 - r) - (b - a) * (r - 1.0));
        q = (q1 - 1.0) * (r - 1.0) * (s - 1.0);
      }
      if (p > 0.0) q = -q;
      p = std::abs(p);
      if (q != 0.0 && 2.0 * p < std::min(3.0 * m * q - std::abs(tol * q), std::abs(e * q))) {
        e = d;
        d = p / q;
      } else {
        d = m;
        e = m;
      }
    } else {
      d = m;
      e = m;
    }
    a = b;
    fa = fb;
    if (std::abs(d) > tol)
      b += d;
    else
      b += (m > 0.0 ? tol : -tol);
    fb = f.value(b, EvalStage::Refine);
  }
  throw std::runtime_error("Brent did not converge within qp_bisection_max_iter");
}

inline bool AcceptRoot(const RootCandidate& cand, const SolverOptions& opt) {
  if (!std::isfinite(cand.omega) || !std::isfinite(cand.Z)) return false;
  if (std::abs(cand.residual) > opt.g_sc_limit) return false;
  if (cand.Z <= 0.0) return false;
  if (cand.Z < opt.min_accepted_Z) return false;
  if (cand.Z > opt.max_accepted_Z) return false;
  return true;
}

inline double ScoreRoot(const RootCandidate& cand) { return cand.Z - 0.1 * cand.distance_to_ref; }

inline const RootCandidate& BestRoot(const std::vector<RootCandidate>& v) {
  // std::max_element semantics: first of the maxima
  size_t best = 0;
  for (size_t i = 1; i < v.size(); ++i)
    if (ScoreRoot(v[best]) < ScoreRoot(v[i])) best = i;
  return v[best];
}

template <typename QPFunc>
std::optional<RootCandidate> RefineQPInterval(double lowerbound, double f_lowerbound, double upperbound,
                                              double f_upperbound, const QPFunc& f, double reference,
                                              const SolverOptions& opt, bool use_brent) {
  RootCandidate cand;
  const bool left_near_zero = std::abs(f_lowerbound) <= opt.g_sc_limit;
  const bool right_near_zero = std::abs(f_upperbound) <= opt.g_sc_limit;
  const bool same_sign = (f_lowerbound * f_upperbound > 0.0);
  if (same_sign) {
    if (left_near_zero || right_near_zero) {
      cand.omega = (std::abs(f_lowerbound) <= std::abs(f_upperbound)) ? lowerbound : upperbound;
    } else {
      return std::nullopt;
    }
  } else {
    cand.omega = use_brent ? SolveQP_Brent(lowerbound, f_lowerbound, upperbound, f_upperbound, f, opt)
                           : SolveQP_Bisection(lowerbound, f_lowerbound, upperbound, f_upperbound, f, opt);
  }
  f.prefetch(&cand.omega, 1, true);  // residual and slope at the same frequency: one evaluation
  cand.residual = f.value(cand.omega, EvalStage::Refine);
  cand.deriv = f.deriv(cand.omega);
  cand.Z = (std::abs(cand.deriv) > 1e-14) ? -1.0 / cand.deriv : std::numeric_limits<double>::infinity();
  cand.distance_to_ref = std::abs(cand.omega - reference);
  cand.accepted = AcceptRoot(cand, opt);
  return cand;
}

template <typename QPFunc>
std::optional<double> SolveQP_Grid_Windowed(QPFunc& fqp, double frequency0, double left_limit, double right_limit,
                                            Index gw_sc_iteration, const SolverOptions& opt,
                                            WindowDiagnostics* wdiag = nullptr,
                                            std::vector<RootCandidate>* accepted_roots_out = nullptr,
                                            std::vector<RootCandidate>* rejected_roots_out = nullptr,
                                            bool use_brent = false) {
  struct Sample {
    double omega = 0.0, fval = 0.0;
  };
  WindowDiagnostics local_diag;
  std::vector<RootCandidate> accepted_roots, rejected_roots;
  auto publish = [&]() {
    if (wdiag) *wdiag = local_diag;
    if (accepted_roots_out) *accepted_roots_out = accepted_roots;
    if (rejected_roots_out) *rejected_roots_out = rejected_roots;
  };
  if (left_limit >= right_limit) {
    publish();
    return std::nullopt;
  }
  const double shell_width = EffectiveAdaptiveShellWidth(opt);
  double center = frequency0;
  if (gw_sc_iteration == 0) {
    fqp.prefetch(&frequency0, 1, true);
    const double f0 = fqp.value(frequency0, EvalStage::Other);
    const double df0 = fqp.deriv(frequency0);
    if (std::isfinite(f0) && std::isfinite(df0) && std::abs(df0) > 1e-6) {
      const double w_lin = frequency0 - f0 / df0;
      if (std::isfinite(w_lin) && w_lin >= left_limit && w_lin <= right_limit) center = w_lin;
    }
  }
  center = std::max(left_limit, std::min(right_limit, center));
  const double max_shell_reach = std::max(center - left_limit, right_limit - center);
  const Index n_shells = static_cast<Index>(std::ceil(max_shell_reach / shell_width));

  auto refine_and_store = [&](double a, double fa, double b, double fb, Index shell_idx) {
    if (b < a) {
      std::swap(a, b);
      std::swap(fa, fb);
    }
    struct LocalBracket {
      double left, f_left, right, f_right;
      double midpoint() const { return 0.5 * (left + right); }
    };
    const Index local_substeps = 12;
    std::vector<LocalBracket> local_brackets;
    if (b > a) {
      const double dx = (b - a) / static_cast<double>(local_substeps);
      {
        std::vector<double> pts;
        for (Index i = 1; i < local_substeps; ++i) pts.push_back(a + static_cast<double>(i) * dx);
        fqp.prefetch(pts.data(), pts.size());
      }
      double x_prev = a, f_prev = fa;
      for (Index i = 1; i <= local_substeps; ++i) {
        const double x_curr = (i == local_substeps) ? b : (a + static_cast<double>(i) * dx);
        const double f_curr = (i == local_substeps) ? fb : fqp.value(x_curr, EvalStage::Scan);
        if ((f_prev < 0.0 && f_curr > 0.0) || (f_prev > 0.0 && f_curr < 0.0))
          local_brackets.push_back({x_prev, f_prev, x_curr, f_curr});
        if (std::abs(f_prev) <= opt.g_sc_limit && x_prev < x_curr)
          local_brackets.push_back({x_prev, f_prev, x_curr, f_curr});
        if (std::abs(f_curr) <= opt.g_sc_limit && x_prev < x_curr)
          local_brackets.push_back({x_prev, f_prev, x_curr, f_curr});
        x_prev = x_curr;
        f_prev = f_curr;
      }
    }
    if (!local_brackets.empty()) {
      size_t best = 0;
      double best_dist = std::abs(local_brackets[0].midpoint() - center);
      for (size_t it = 1; it < local_brackets.size(); ++it) {
        const double dist = std::abs(local_brackets[it].midpoint() - center);
        if (dist < best_dist - 1e-14 ||
            (std::abs(dist - best_dist) <= 1e-14 && local_brackets[it].left < local_brackets[best].left)) {
          best = it;
          best_dist = dist;
        }
      }
      a = local_brackets[best].left;
      fa = local_brackets[best].f_left;
      b = local_brackets[best].right;
      fb = local_brackets[best].f_right;
    }
    auto cand_opt = RefineQPInterval(a, fa, b, fb, fqp, frequency0, opt, use_brent);
    if (!cand_opt) return;
    if (local_diag.first_interval_shell < 0) local_diag.first_interval_shell = shell_idx;
    ++local_diag.intervals_found;
    const RootCandidate& cand = *cand_opt;
    if (cand.accepted) {
      if (local_diag.first_accepted_shell < 0) local_diag.first_accepted_shell = shell_idx;
      accepted_roots.push_back(cand);
    } else {
      rejected_roots.push_back(cand);
    }
  };

  {
    // every shell point is visited below whatever the function values are: announce them all
    std::vector<double> pts;
    for (Index shell = 1; shell <= n_shells; ++shell) {
      const double delta = double(shell) * shell_width;
      if (center - delta >= left_limit) pts.push_back(center - delta);
      if (center + delta <= right_limit) pts.push_back(center + delta);
    }
    pts.push_back(left_limit);
    pts.push_back(right_limit);
    fqp.prefetch(pts.data(), pts.size());
  }
  Sample center_pt{center, fqp.value(center, EvalStage::Scan)};
  bool left_active = true, right_active = true;
  Sample left_prev = center_pt, right_prev = center_pt;
  for (Index shell = 1; shell <= n_shells; ++shell) {
    local_diag.shells_explored = shell;
    bool added_this_shell = false;
    const double delta = double(shell) * shell_width;
    if (left_active) {
      const double omega_left = center - delta;
      if (omega_left >= left_limit) {
        Sample left_curr{omega_left, fqp.value(omega_left, EvalStage::Scan)};
        added_this_shell = true;
        if (left_prev.fval * left_curr.fval < 0.0)
          refine_and_store(left_curr.omega, left_curr.fval, left_prev.omega, left_prev.fval, shell);
        left_prev = left_curr;
      } else {
        left_active = false;
      }
    }
    if (right_active) {
      const double omega_right = center + delta;
      if (omega_right <= right_limit) {
        Sample right_curr{omega_right, fqp.value(omega_right, EvalStage::Scan)};
        added_this_shell = true;
        if (right_prev.fval * right_curr.fval < 0.0)
          refine_and_store(right_prev.omega, right_prev.fval, right_curr.omega, right_curr.fval, shell);
        right_prev = right_curr;
      } else {
        right_active = false;
      }
    }
    if (!added_this_shell && !left_active && !right_active) break;
  }
  if (left_prev.omega > left_limit + 1e-12) {
    Sample left_end{left_limit, fqp.value(left_limit, EvalStage::Scan)};
    if (left_end.fval * left_prev.fval < 0.0)
      refine_and_store(left_end.omega, left_end.fval, left_prev.omega, left_prev.fval, local_diag.shells_explored + 1);
  }
  if (right_prev.omega < right_limit - 1e-12) {
    Sample right_end{right_limit, fqp.value(right_limit, EvalStage::Scan)};
    if (right_prev.fval * right_end.fval < 0.0)
      refine_and_store(right_prev.omega, right_prev.fval, right_end.omega, right_end.fval,
                       local_diag.shells_explored + 1);
  }
  if (!accepted_roots.empty()) {
    const RootCandidate& best = BestRoot(accepted_roots);
    local_diag.chosen_shell = static_cast<int>(std::llround(std::abs(best.omega - center) / shell_width));
    publish();
    return best.omega;
  }
  if (!rejected_roots.empty()) {
    const RootCandidate& least_bad = BestRoot(rejected_roots);
    local_diag.chosen_shell = static_cast<int>(std::llround(std::abs(least_bad.omega - center) / shell_width));
    publish();
    return least_bad.omega;
  }
  publish();
  return std::nullopt;
}

}  // namespace qp_solver

// newton_rapson.h:39-90
template <class Func>
class NewtonRapson {
 public:
  enum Errors { success, smalldenom, notconverged };
  NewtonRapson(Index max_iterations, double tolerance, double alpha = 1.0)
      : max_iterations_(max_iterations), tolerance_(tolerance), alpha_(alpha) {}
  double FindRoot(const Func& f, double x0) {
    info_ = Errors::notconverged;
    double x = x0;
    for (iter_ = 0; iter_ < max_iterations_; iter_++) {
      std::pair<double, double> res = f(x);
      if (std::abs(res.second) < 1e-12) {
        info_ = Errors::smalldenom;
        break;
      }
      double step = -alpha_ * res.first / res.second;
      if (std::abs(step) < tolerance_) {
        info_ = Errors::success;
        break;
      }
      x += step;
    }
    return x;
  }
  Errors getInfo() const { return info_; }

 private:
  Errors info_ = Errors::notconverged;
  Index max_iterations_;
  Index iter_ = 0;
  double tolerance_;
  double alpha_;
};

}  // namespace xtp
}  // namespace votca
