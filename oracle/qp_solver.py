"""Oracle for the quasiparticle root search (test infrastructure).

Follows xtp/include/votca/xtp/qp_solver_utils.h (entire file) and
xtp/include/votca/xtp/newton_rapson.h:39-90.  Control flow kept statement by
statement because root selection is discontinuous in Sigma_c.
"""
import math
from dataclasses import dataclass, field


@dataclass
class SolverOptions:
    g_sc_limit: float = 1e-5
    qp_bisection_max_iter: int = 200
    qp_full_window_half_width: float = 0.75
    qp_dense_spacing: float = 0.002
    qp_adaptive_shell_width: float = 0.025
    qp_adaptive_shell_count: int = 0
    min_accepted_Z: float = 0.05
    max_accepted_Z: float = 1.5


@dataclass
class RootCandidate:
    omega: float = 0.0
    residual: float = 0.0
    deriv: float = 0.0
    Z: float = 0.0
    distance_to_ref: float = 0.0
    accepted: bool = False


@dataclass
class WindowDiagnostics:
    shells_explored: int = 0
    first_interval_shell: int = -1
    first_accepted_shell: int = -1
    chosen_shell: int = -1
    intervals_found: int = 0


def legacy_full_window_half_width(opt):
    if opt.qp_grid_steps <= 1 or opt.qp_grid_spacing <= 0.0:
        return -1.0
    return 0.5 * opt.qp_grid_spacing * float(opt.qp_grid_steps - 1)


def legacy_adaptive_shell_width(opt):
    if opt.qp_grid_steps <= 1 or opt.qp_grid_spacing <= 0.0:
        return -1.0
    full = opt.qp_grid_spacing * float(opt.qp_grid_steps - 1)
    base = max(21, opt.qp_grid_steps // 4)
    if base <= 1:
        return 4.0 * opt.qp_grid_spacing
    return full / float(base - 1)


def normalize_grid_search_options(opt):
    has_legacy = opt.qp_grid_steps > 1 and opt.qp_grid_spacing > 0.0
    if opt.qp_full_window_half_width <= 0.0:
        opt.qp_full_window_half_width = legacy_full_window_half_width(opt) if has_legacy else 0.75
    if opt.qp_dense_spacing <= 0.0:
        opt.qp_dense_spacing = opt.qp_grid_spacing if has_legacy else 0.002
    if opt.qp_adaptive_shell_count <= 0 and opt.qp_adaptive_shell_width <= 0.0:
        opt.qp_adaptive_shell_width = legacy_adaptive_shell_width(opt) if has_legacy else 0.025
    if opt.qp_full_window_half_width <= 0.0:
        raise RuntimeError("Invalid QP search setup: qp_full_window_half_width must be > 0")
    if opt.qp_dense_spacing <= 0.0:
        raise RuntimeError("Invalid QP search setup: qp_dense_spacing must be > 0")
    if opt.qp_adaptive_shell_count <= 0 and opt.qp_adaptive_shell_width <= 0.0:
        raise RuntimeError("Invalid QP search setup: need qp_adaptive_shell_width > 0 or "
                           "qp_adaptive_shell_count > 0")


def effective_adaptive_shell_width(opt):
    if opt.qp_adaptive_shell_count > 0:
        return opt.qp_full_window_half_width / float(opt.qp_adaptive_shell_count)
    return opt.qp_adaptive_shell_width


def solve_bisection(lo, flo, hi, fhi, f, opt):
    if flo * fhi > 0.0:
        raise RuntimeError("Bisection needs a positive and negative function value")
    while True:
        c = 0.5 * (lo + hi)
        if abs(hi - lo) < opt.g_sc_limit:
            return c
        yc = f.value(c)
        if abs(yc) < opt.g_sc_limit:
            return c
        if yc * flo > 0.0:
            lo, flo = c, yc
        else:
            hi, fhi = c, yc


def solve_brent(lo, flo, hi, fhi, f, opt):
    if flo * fhi > 0.0:
        raise RuntimeError("Brent needs a positive and negative function value")
    a, b, fa, fb = lo, hi, flo, fhi
    c, fc = a, fa
    d = b - a
    e = d
    for _ in range(opt.qp_bisection_max_iter):
        if (fb > 0.0 and fc > 0.0) or (fb < 0.0 and fc < 0.0):
            c, fc = a, fa
            d = b - a
            e = d
        if abs(fc) < abs(fb):
            a, b, c = b, c, b
            fa, fb, fc = fb, fc, fb
        tol = opt.g_sc_limit
        m = 0.5 * (c - b)
        if abs(m) < tol or abs(fb) < opt.g_sc_limit:
            return b
        if abs(e) >= tol and abs(fa) > abs(fb):
            s = fb / fa
            if a == c:
                p = 2.0 * m * s
                q = 1.0 - s
            else:
                q1 = fa / fc
                r = fb / fc
                p = s * (2.0 * m * q1 * (q1 - r) - (b - a) * (r - 1.0))
                q = (q1 - 1.0) * (r - 1.0) * (s - 1.0)
            if p > 0.0:
                q = -q
            p = abs(p)
            if q != 0.0 and 2.0 * p < min(3.0 * m * q - abs(tol * q), abs(e * q)):
                e = d
                d = p / q
            else:
                d = m
                e = m
        else:
            d = m
            e = m
        a, fa = b, fb
        if abs(d) > tol:
            b += d
        else:
            b += tol if m > 0.0 else -tol
        fb = f.value(b)
    raise RuntimeError("Brent did not converge within qp_bisection_max_iter")


def accept_root(cand, opt):
    if not math.isfinite(cand.omega) or not math.isfinite(cand.Z):
        return False
    if abs(cand.residual) > opt.g_sc_limit:
        return False
    if cand.Z <= 0.0:
        return False
    if cand.Z < opt.min_accepted_Z:
        return False
    if cand.Z > opt.max_accepted_Z:
        return False
    return True


def score_root(cand):
    return cand.Z - 0.1 * cand.distance_to_ref


def _argmax_first(cands):
    # std::max_element: first of the maximal elements
    best = cands[0]
    for c in cands[1:]:
        if score_root(best) < score_root(c):
            best = c
    return best


def refine_qp_interval(lo, flo, hi, fhi, f, reference, opt, use_brent):
    cand = RootCandidate()
    left_near = abs(flo) <= opt.g_sc_limit
    right_near = abs(fhi) <= opt.g_sc_limit
    same_sign = flo * fhi > 0.0
    if same_sign:
        if left_near or right_near:
            cand.omega = lo if abs(flo) <= abs(fhi) else hi
        else:
            return None
    else:
        cand.omega = (solve_brent if use_brent else solve_bisection)(lo, flo, hi, fhi, f, opt)
    cand.residual = f.value(cand.omega)
    cand.deriv = f.deriv(cand.omega)
    cand.Z = -1.0 / cand.deriv if abs(cand.deriv) > 1e-14 else math.inf
    cand.distance_to_ref = abs(cand.omega - reference)
    cand.accepted = accept_root(cand, opt)
    return cand


def solve_qp_grid_windowed(fqp, frequency0, left_limit, right_limit, gw_sc_iteration, opt,
                           use_brent=False):
    """Adaptive shell scan.  Returns (omega or None, accepted_roots, rejected_roots, diag)."""
    diag = WindowDiagnostics()
    accepted, rejected = [], []
    if left_limit >= right_limit:
        return None, accepted, rejected, diag
    shell_width = effective_adaptive_shell_width(opt)
    center = frequency0
    if gw_sc_iteration == 0:
        f0 = fqp.value(frequency0)
        df0 = fqp.deriv(frequency0)
        if math.isfinite(f0) and math.isfinite(df0) and abs(df0) > 1e-6:
            w_lin = frequency0 - f0 / df0
            if math.isfinite(w_lin) and left_limit <= w_lin <= right_limit:
                center = w_lin
    center = max(left_limit, min(right_limit, center))
    max_reach = max(center - left_limit, right_limit - center)
    n_shells = int(math.ceil(max_reach / shell_width))

    def refine_and_store(a, fa, b, fb, shell_idx):
        if b < a:
            a, b = b, a
            fa, fb = fb, fa
        local_substeps = 12
        brackets = []
        if b > a:
            dx = (b - a) / float(local_substeps)
            x_prev, f_prev = a, fa
            for i in range(1, local_substeps + 1):
                x_curr = b if i == local_substeps else a + float(i) * dx
                f_curr = fb if i == local_substeps else fqp.value(x_curr)
                if (f_prev < 0.0 and f_curr > 0.0) or (f_prev > 0.0 and f_curr < 0.0):
                    brackets.append((x_prev, f_prev, x_curr, f_curr))
                if abs(f_prev) <= opt.g_sc_limit and x_prev < x_curr:
                    brackets.append((x_prev, f_prev, x_curr, f_curr))
                if abs(f_curr) <= opt.g_sc_limit and x_prev < x_curr:
                    brackets.append((x_prev, f_prev, x_curr, f_curr))
                x_prev, f_prev = x_curr, f_curr
        if brackets:
            best = brackets[0]
            best_dist = abs(0.5 * (best[0] + best[2]) - center)
            for br in brackets[1:]:
                dist = abs(0.5 * (br[0] + br[2]) - center)
                if dist < best_dist - 1e-14 or (abs(dist - best_dist) <= 1e-14 and br[0] < best[0]):
                    best, best_dist = br, dist
            a, fa, b, fb = best
        cand = refine_qp_interval(a, fa, b, fb, fqp, frequency0, opt, use_brent)
        if cand is None:
            return
        if diag.first_interval_shell < 0:
            diag.first_interval_shell = shell_idx
        diag.intervals_found += 1
        if cand.accepted:
            if diag.first_accepted_shell < 0:
                diag.first_accepted_shell = shell_idx
            accepted.append(cand)
        else:
            rejected.append(cand)

    center_pt = (center, fqp.value(center))
    left_active = right_active = True
    left_prev = right_prev = center_pt
    for shell in range(1, n_shells + 1):
        diag.shells_explored = shell
        added = False
        delta = float(shell) * shell_width
        if left_active:
            wl = center - delta
            if wl >= left_limit:
                cur = (wl, fqp.value(wl))
                added = True
                if left_prev[1] * cur[1] < 0.0:
                    refine_and_store(cur[0], cur[1], left_prev[0], left_prev[1], shell)
                left_prev = cur
            else:
                left_active = False
        if right_active:
            wr = center + delta
            if wr <= right_limit:
                cur = (wr, fqp.value(wr))
                added = True
                if right_prev[1] * cur[1] < 0.0:
                    refine_and_store(right_prev[0], right_prev[1], cur[0], cur[1], shell)
                right_prev = cur
            else:
                right_active = False
        if not added and not left_active and not right_active:
            break
    if left_prev[0] > left_limit + 1e-12:
        end = (left_limit, fqp.value(left_limit))
        if end[1] * left_prev[1] < 0.0:
            refine_and_store(end[0], end[1], left_prev[0], left_prev[1], diag.shells_explored + 1)
    if right_prev[0] < right_limit - 1e-12:
        end = (right_limit, fqp.value(right_limit))
        if right_prev[1] * end[1] < 0.0:
            refine_and_store(right_prev[0], right_prev[1], end[0], end[1], diag.shells_explored + 1)
    if accepted:
        best = _argmax_first(accepted)
        diag.chosen_shell = int(round(abs(best.omega - center) / shell_width))  # qp_solver_utils.h:660
        return best.omega, accepted, rejected, diag
    if rejected:
        best = _argmax_first(rejected)
        diag.chosen_shell = int(round(abs(best.omega - center) / shell_width))  # qp_solver_utils.h:682
        return best.omega, accepted, rejected, diag
    return None, accepted, rejected, diag


def newton_raphson(f, x0, max_iterations, tolerance, alpha):
    """Returns (x, success).  newton_rapson.h:52-76."""
    x = x0
    for _ in range(max_iterations):
        val, der = f.value(x), f.deriv(x)
        if abs(der) < 1e-12:
            return x, False
        step = -alpha * val / der
        if abs(step) < tolerance:
            return x, True
        x += step
    return x, False
