// CPU-only harness around the host layer's QP root search (votca_b200/host/qp_rootsearch.h has no device
// dependency): the known-answer cases of xtp/src/tests/test_qp_solver_utils.cc:87-316 are driven from
// tests/test_host_logic_cpu.py through these C entry points.
#include <cstring>

#include "../../votca_b200/host/anderson_mixing.h"
#include "../../votca_b200/host/qp_rootsearch.h"
#include "../../votca_b200/host/quadrature.h"
#include "../../votca_b200/host/vc2index.h"

using namespace votca;
using namespace votca::xtp;
using namespace votca::xtp::qp_solver;

namespace {
// mock functions of the reference test (:46-84); kind 0: root - w (deriv -1), 1: w - root (deriv +1),
// 2: (w - r1)(w - r2) with a constant deriv of -1
struct MockFunc {
  int kind;
  double r1, r2;
  double value(double w, EvalStage) const {
    if (kind == 0) return r1 - w;
    if (kind == 1) return w - r1;
    return (w - r1) * (w - r2);
  }
  double deriv(double) const { return kind == 1 ? 1.0 : -1.0; }
};
struct LegacyOpt {
  Index qp_grid_steps = 0;
  double qp_grid_spacing = 0.0;
  double qp_full_window_half_width = -1.0, qp_dense_spacing = -1.0, qp_adaptive_shell_width = -1.0;
  Index qp_adaptive_shell_count = 0;
};
}  // namespace

extern "C" {

// out: half width, dense spacing, shell width, shell count, legacy half width, legacy shell width
int qp_normalize(long steps, double spacing, double half_width, double dense, double shell_width, long shell_count,
                 double* out) {
  LegacyOpt o;
  o.qp_grid_steps = steps;
  o.qp_grid_spacing = spacing;
  o.qp_full_window_half_width = half_width;
  o.qp_dense_spacing = dense;
  o.qp_adaptive_shell_width = shell_width;
  o.qp_adaptive_shell_count = shell_count;
  out[4] = LegacyFullWindowHalfWidth(o);
  out[5] = LegacyAdaptiveShellWidth(o);
  try {
    NormalizeGridSearchOptions(o);
  } catch (const std::exception&) {
    return 1;
  }
  out[0] = o.qp_full_window_half_width;
  out[1] = o.qp_dense_spacing;
  out[2] = o.qp_adaptive_shell_width;
  out[3] = (double)o.qp_adaptive_shell_count;
  return 0;
}

double qp_effective_shell_width(double half_width, double shell_width, long shell_count) {
  SolverOptions o;
  o.qp_full_window_half_width = half_width;
  o.qp_adaptive_shell_width = shell_width;
  o.qp_adaptive_shell_count = shell_count;
  return EffectiveAdaptiveShellWidth(o);
}

int qp_accept_root(double residual, double Z, double g_sc_limit, double minZ, double maxZ) {
  SolverOptions o;
  o.g_sc_limit = g_sc_limit;
  o.min_accepted_Z = minZ;
  o.max_accepted_Z = maxZ;
  RootCandidate c;
  c.omega = 0.12;
  c.residual = residual;
  c.deriv = -1.25;
  c.Z = Z;
  c.distance_to_ref = 0.12;
  return AcceptRoot(c, o) ? 1 : 0;
}

// out: found, root, n_accepted, n_rejected, first accepted Z (or first rejected Z), shells_explored,
//      first_interval_shell, first_accepted_shell, chosen_shell, intervals_found
int qp_windowed(int kind, double r1, double r2, double freq0, double left, double right, long iteration,
                double g_sc_limit, double half_width, double dense, double shell_width, int use_brent, double* out) {
  MockFunc f{kind, r1, r2};
  SolverOptions o;
  o.g_sc_limit = g_sc_limit;
  o.qp_bisection_max_iter = 200;
  o.qp_full_window_half_width = half_width;
  o.qp_dense_spacing = dense;
  o.qp_adaptive_shell_width = shell_width;
  o.qp_adaptive_shell_count = 0;
  WindowDiagnostics d;
  std::vector<RootCandidate> acc, rej;
  std::optional<double> root = SolveQP_Grid_Windowed(f, freq0, left, right, iteration, o, &d, &acc, &rej, use_brent != 0);
  out[0] = root ? 1.0 : 0.0;
  out[1] = root ? *root : 0.0;
  out[2] = (double)acc.size();
  out[3] = (double)rej.size();
  out[4] = !acc.empty() ? acc.front().Z : (!rej.empty() ? rej.front().Z : 0.0);
  out[5] = (double)d.shells_explored;
  out[6] = (double)d.first_interval_shell;
  out[7] = (double)d.first_accepted_shell;
  out[8] = (double)d.chosen_shell;
  out[9] = (double)d.intervals_found;
  return 0;
}

// Anderson mixing driven like test_anderson.cc:33-100: nsteps (input, output) pairs of length n; the mixed
// vector of step s is written to mixed[s*n ..] and fed back as the next input when feed_back != 0
int anderson_run(long order, double alpha, long n, long nsteps, const double* first_input, const double* outputs,
                 double* mixed) {
  Anderson mix;
  mix.Configure(order, alpha);
  VectorXd in(first_input, n);
  for (long s = 0; s < nsteps; ++s) {
    mix.UpdateInput(in);
    mix.UpdateOutput(VectorXd(outputs + s * n, n));
    VectorXd m = mix.MixHistory();
    for (long i = 0; i < n; ++i) mixed[s * n + i] = m(i);
    in = m;
  }
  return 0;
}

// test_newton_rapson.cc:31-52: f(x) = x^2 - c
struct SquareMinus {
  double c;
  std::pair<double, double> operator()(double x) const { return {x * x - c, 2 * x}; }
};
int newton_sqrt(double c, double x0, long iterations, double tolerance, double* root) {
  SquareMinus f{c};
  NewtonRapson<SquareMinus> n(iterations, tolerance);
  *root = n.FindRoot(f, x0);
  return (int)n.getInfo();
}

// mapped quadrature points / weights of the CDA integration; returns the symmetry flag (or -1 on error)
int quadrature_points(const char* scheme, long order, double* pts, double* wts) {
  try {
    std::vector<double> p, w;
    bool sym = false;
    mapped_gauss_legendre(scheme, order, p, w, sym);
    for (long i = 0; i < order; ++i) {
      pts[i] = p[i];
      wts[i] = w[i];
    }
    return sym ? 1 : 0;
  } catch (const std::exception&) {
    return -1;
  }
}

// vc2index: what = 0 -> I(a, b), 1 -> v(a), 2 -> c(a)
long vc2index_eval(long vmin, long cmin, long ctotal, int what, long a, long b) {
  votca::xtp::vc2index vc(vmin, cmin, ctotal);
  return what == 0 ? vc.I(a, b) : what == 1 ? vc.v(a) : vc.c(a);
}
}  // extern "C"
