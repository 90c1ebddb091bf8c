// Host-side tables of the device AO-integral kernels (ao3c_core.cuh): Boys-function grid, cartesian -> pure
// (real solid harmonic) matrices in libint's order m = -l..l, the (t,u,v) enumeration of Hermite indices and the
// single-centre Hermite expansion of a primitive.  Plain C++ (no CUDA) so the CPU test harness
// (tests/host_harness/ao3c_host.cc) builds the very same tables as capi_ao3c.cu.
//
// Conventions restated from the reference's integral provider: libint2 pure shells as built by
// AOShell::LibintShell (xtp/src/libxtp/aoshell.cc:65-79, contr.pure = true), functions ordered m = -l..l
// (aoshell.cc:116-133), cartesian components in libint's standard order (lx descending, then ly descending).
#pragma once
#include <algorithm>
#include <cstdlib>
#include <cmath>
#include <cstdint>
#include <functional>
#include <stdexcept>
#include <thread>
#include <vector>

#include "ao3c_core.cuh"

namespace gwbse {
namespace ao {

constexpr int LMAX_SHELL = 6;                       // i functions (libint is built with --with-max-am=6)
constexpr int LMAX_TOTAL = 16;                      // la + lb + lc of one integral class
constexpr int BOYS_TAYLOR = 8;                      // terms of the Taylor step off the grid
constexpr int BOYS_ORDERS = LMAX_TOTAL + BOYS_TAYLOR;  // F_0 .. F_{ORDERS-1} tabulated
constexpr double BOYS_DX = 0.1;
constexpr int BOYS_POINTS = 361;                    // x = 0 .. 36
constexpr double BOYS_XMAX = 35.95;                 // beyond: F_0 from erf + upward recursion
constexpr int HERM1_STRIDE = (LMAX_SHELL + 1) * (LMAX_SHELL + 1);
// primitive pairs with |c_a c_b| exp(-a b / (a+b) |AB|^2) below this contribute nothing at double precision
constexpr double PRIM_THRESHOLD = 1e-20;

inline int ncart(int l) { return (l + 1) * (l + 2) / 2; }
inline int nherm(int L) { return (L + 1) * (L + 2) * (L + 3) / 6; }

// (t,u,v) of Hermite index h in the degree-major order hidx(t,u,v) = N(N+1)(N+2)/6 + (u+v)(u+v+1)/2 + v,
// N = t+u+v.  The entries of degree l are the cartesian components of a shell in libint order.
inline std::vector<uint32_t> make_tuv_table() {
  std::vector<uint32_t> tab((size_t)nherm(LMAX_TOTAL), 0);
  for (int N = 0; N <= LMAX_TOTAL; ++N)
    for (int t = N; t >= 0; --t)
      for (int u = N - t; u >= 0; --u) {
        const int v = N - t - u;
        const int h = N * (N + 1) * (N + 2) / 6 + (u + v) * (u + v + 1) / 2 + v;
        tab[(size_t)h] = (uint32_t)t | ((uint32_t)u << 8) | ((uint32_t)v << 16);
      }
  return tab;
}

// F_n(x_j), x_j = j * BOYS_DX: highest order from the (all-positive) series
//   F_n(x) = e^-x sum_k (2x)^k / ((2n+1)(2n+3)...(2n+2k+1)),
// the lower ones by the downward recursion F_{n-1} = (2x F_n + e^-x) / (2n-1)  (stable).
inline std::vector<double> make_boys_table() {
  std::vector<double> tab((size_t)BOYS_POINTS * BOYS_ORDERS);
  const int top = BOYS_ORDERS - 1;
  for (int j = 0; j < BOYS_POINTS; ++j) {
    const long double x = (long double)j * (long double)BOYS_DX;
    const long double ex = std::exp(-x);
    long double term = 1.0L / (2 * top + 1), sum = term;
    for (int k = 1; k < 2000; ++k) {
      term *= 2.0L * x / (2 * top + 2 * k + 1);
      sum += term;
      if (term < 1e-22L * sum) break;
    }
    long double f = ex * sum;
    tab[(size_t)j * BOYS_ORDERS + top] = (double)f;
    for (int n = top; n > 0; --n) {
      f = (2.0L * x * f + ex) / (2 * n - 1);
      tab[(size_t)j * BOYS_ORDERS + n - 1] = (double)f;
    }
  }
  return tab;
}

inline double binom(int n, int k) {
  if (n < 0 || k < 0 || k > n) return 0.0;
  double r = 1.0;
  for (int i = 1; i <= k; ++i) r = r * (n - k + i) / i;
  return r;
}
inline double factorial(int n) {
  double r = 1.0;
  for (int i = 2; i <= n; ++i) r *= i;
  return r;
}

// (2l+1) x ncart(l) row-major: real solid harmonics in Racah normalisation (Helgaker, Jorgensen, Olsen,
// "Molecular Electronic-Structure Theory", eq. 6.4.47-6.4.50), the form libint2 generates for pure shells.
inline std::vector<double> make_pure_matrix(int l) {
  const int nc = ncart(l);
  std::vector<double> T((size_t)(2 * l + 1) * nc, 0.0);
  auto cidx = [&](int ly, int lz) { return (ly + lz) * (ly + lz + 1) / 2 + lz; };  // libint order within a shell
  for (int m = -l; m <= l; ++m) {
    const int am = m < 0 ? -m : m;
    const double norm = (1.0 / (std::pow(2.0, am) * factorial(l))) *
                        std::sqrt(2.0 * factorial(l + am) * factorial(l - am) / (m == 0 ? 2.0 : 1.0));
    const int vm2 = m >= 0 ? 0 : 1;  // 2 v_m
    const int nv = (int)std::floor(am / 2.0 - vm2 / 2.0);
    for (int t = 0; t <= (l - am) / 2; ++t)
      for (int u = 0; u <= t; ++u)
        for (int iv = 0; iv <= nv; ++iv) {
          const int twov = 2 * iv + vm2;
          const double c = (((t + iv) & 1) ? -1.0 : 1.0) * std::pow(0.25, t) * binom(l, t) * binom(l - t, am + t) *
                           binom(t, u) * binom(am, twov);
          const int lx = 2 * t + am - 2 * u - twov, ly = 2 * u + twov, lz = l - 2 * t - am;
          if (lx < 0 || ly < 0 || lz < 0) continue;
          T[(size_t)(m + l) * nc + cidx(ly, lz)] += norm * c;
        }
  }
  return T;
}

// x^i exp(-g x^2) = sum_t e[i][t] Lambda_t(x; g): the one-centre case of the McMurchie-Davidson recursion
// (X_PA = 0).  Row-major (LMAX_SHELL+1)^2, zero where i - t is odd or t > i.
inline void fill_herm1(double g, double* e) {
  const int S = LMAX_SHELL + 1;
  for (int k = 0; k < S * S; ++k) e[k] = 0.0;
  e[0] = 1.0;
  const double inv2g = 0.5 / g;
  for (int i = 1; i < S; ++i)
    for (int t = 0; t <= i; ++t) {
      double v = 0.0;
      if (t + 1 <= i - 1) v += (t + 1) * e[(i - 1) * S + t + 1];
      if (t >= 1) v += inv2g * e[(i - 1) * S + t - 1];
      e[i * S + t] = v;
    }
}

// The coefficients a libint2::Shell built by AOShell::LibintShell holds (xtp/src/libxtp/aoshell.cc:65-79) after
// AOShell::normalizeContraction (aoshell.cc:81-89, called from AOBasis::Fill, aobasis.cc:99-100): the raw
// contraction factors of the basis-set file times libint's primitive normalisation of x^l exp(-a r^2), the whole
// contraction divided by the square root of the self-overlap of the shell's first function (libint's own unit
// normalisation of contracted shells is switched off, libint2_calls.cc:168).  Racah-normalised solid harmonics
// share the norm of x^l, so that self-overlap is the one of the x^l component.
inline double double_factorial(int n) {
  double r = 1.0;
  for (int k = n; k > 1; k -= 2) r *= k;
  return r;
}
inline void normalize_contraction(int l, int nprim, const double* exps, const double* raw, double* out) {
  const double pi = 3.14159265358979323846;
  const double df = double_factorial(2 * l - 1);
  for (int p = 0; p < nprim; ++p)
    out[p] = raw[p] * std::sqrt(std::pow(2.0, l) * std::pow(2.0 * exps[p], l + 1.5) / (std::pow(pi, 1.5) * df));
  double s = 0.0;
  for (int p = 0; p < nprim; ++p)
    for (int q = 0; q < nprim; ++q) {
      const double g = exps[p] + exps[q];
      s += out[p] * out[q] * df / std::pow(2.0 * g, l) * std::pow(pi / g, 1.5);
    }
  const double inv = 1.0 / std::sqrt(s);
  for (int p = 0; p < nprim; ++p) out[p] *= inv;
}

// Flat description of a basis as the kernels read it (host copy; capi_ao3c.cu uploads the vectors).
struct HostBasis {
  int nshell = 0, nfunc = 0, nprim = 0, lmax = 0;
  std::vector<int> l, np, prim0, func0;
  std::vector<double> center, exps, coefs, herm1;

  // coefs must already hold the primitive normalisation and VOTCA's shell norm (aoshell.cc:81-89)
  void build(int nshell_, const int* l_, const int* nprim_, const double* center_, const double* exps_,
             const double* coefs_) {
    nshell = nshell_;
    l.assign(l_, l_ + nshell);
    np.assign(nprim_, nprim_ + nshell);
    center.assign(center_, center_ + 3 * (size_t)nshell);
    prim0.resize(nshell);
    func0.resize(nshell);
    nprim = 0;
    nfunc = 0;
    lmax = 0;
    for (int s = 0; s < nshell; ++s) {
      if (l[s] < 0 || l[s] > LMAX_SHELL) throw std::runtime_error("shell angular momentum outside 0..6");
      if (np[s] < 1) throw std::runtime_error("shell without primitives");
      prim0[s] = nprim;
      func0[s] = nfunc;
      nprim += np[s];
      nfunc += 2 * l[s] + 1;
      if (l[s] > lmax) lmax = l[s];
    }
    exps.assign(exps_, exps_ + nprim);
    coefs.assign(coefs_, coefs_ + nprim);
    herm1.resize((size_t)nprim * HERM1_STRIDE);
    for (int p = 0; p < nprim; ++p) {
      if (!(exps[p] > 0.0)) throw std::runtime_error("primitive exponent must be positive");
      fill_herm1(exps[p], &herm1[(size_t)p * HERM1_STRIDE]);
    }
  }

  // Appends the records of shell pair (s, t) to pool - one per primitive pair with
  // |c_a c_b| exp(-a b / (a+b) |AB|^2) >= threshold, layout as PairEntry (ao3c_core.cuh) documents - and returns
  // how many were written.  t < 0: unit partner.  E^{ab}_t is the McMurchie-Davidson recursion in i (shell s)
  // and j (shell t):  E[i+1][j][t] = X_PA E[i][j][t] + (t+1) E[i][j][t+1] + E[i][j][t-1] / 2p, likewise in j.
  int append_pair_records(int s, int t, std::vector<double>& pool, double threshold = PRIM_THRESHOLD) const {
    const bool unit = t < 0;
    const int la = l[s], lb = unit ? 0 : l[t];
    const int T1 = la + lb + 1, ej = (lb + 1) * T1, esz = (la + 1) * ej;
    const double* A = &center[3 * (size_t)s];
    const double* B = unit ? A : &center[3 * (size_t)t];
    const double AB[3] = {A[0] - B[0], A[1] - B[1], A[2] - B[2]};
    const double AB2 = AB[0] * AB[0] + AB[1] * AB[1] + AB[2] * AB[2];
    int written = 0;
    for (int ia = 0; ia < np[s]; ++ia)
      for (int ib = 0; ib < (unit ? 1 : np[t]); ++ib) {
        const double a = exps[prim0[s] + ia], b = unit ? 0.0 : exps[prim0[t] + ib];
        const double cab = coefs[prim0[s] + ia] * (unit ? 1.0 : coefs[prim0[t] + ib]);
        const double p = a + b, mu = a * b / p, inv2p = 0.5 / p;
        if (std::fabs(cab) * std::exp(-mu * AB2) < threshold) continue;
        const size_t base = pool.size();
        pool.resize(base + 5 + 3 * (size_t)esz, 0.0);
        double* rec = &pool[base];
        rec[0] = p;
        for (int d = 0; d < 3; ++d) rec[1 + d] = (a * A[d] + b * B[d]) / p;
        rec[4] = cab;
        for (int d = 0; d < 3; ++d) {
          double* E = rec + 5 + (size_t)d * esz;
          const double Xpa = -b / p * AB[d], Xpb = a / p * AB[d];
          E[0] = std::exp(-mu * AB[d] * AB[d]);
          for (int i = 0; i <= la; ++i)
            for (int j = 0; j <= lb; ++j) {
              if (i == 0 && j == 0) continue;
              const double* src = i > 0 ? E + (i - 1) * ej + j * T1 : E + (j - 1) * T1;
              const double X = i > 0 ? Xpa : Xpb;
              double* dst = E + i * ej + j * T1;
              for (int tt = 0; tt <= i + j; ++tt) {
                double v = 0.0;
                if (tt <= i + j - 1) v += X * src[tt];
                if (tt + 1 <= i + j - 1) v += (tt + 1) * src[tt + 1];
                if (tt >= 1) v += inv2p * src[tt - 1];
                dst[tt] = v;
              }
            }
        }
        ++written;
      }
    return written;
  }
  static int pair_record_doubles(int la, int lb) { return 5 + 3 * (la + 1) * (lb + 1) * (la + lb + 1); }
};

// Launch geometry of one angular-momentum class (l_a >= l_b | l_c): lanes per shell triple = the widest stage of the
// class (accumulators, aux-folded Hermite tensor G, R tensor) rounded up to 4 / 8 / 16 / 32, warps per CTA so that
// about four CTAs share an SM's shared memory where the class is small enough, never more than the opt-in limit.
struct LaunchConfig {
  int group_lanes, groups_per_warp, warps_per_cta, ws_doubles;
  size_t smem_bytes;
  bool fits;
};
inline int lane_div() {
  static const int d = [] {
    const char* e = std::getenv("GWBSE_AO3C_LANEDIV");
    const int v = e ? std::atoi(e) : 4;
    return v < 1 ? 1 : v;
  }();
  return d;
}

inline LaunchConfig launch_config(int la, int lb, int lc, size_t smem_limit) {
  LaunchConfig c{};
  c.ws_doubles = workspace_doubles(la, lb, lc);
  const int Lab = la + lb;
  const int width = std::max({ncart(la) * ncart(lb) * ncart(lc), nherm(Lab) * ncart(lc), nherm(Lab + lc)});
  // lanes per shell triple: the widest stage divided by lane_div() (rounded up to a power of two).  One lane per
  // entry of the widest stage minimises the rounds of a triple but leaves most lanes idle in the narrow stages (the
  // R recursion has nh(L - n) entries at level n); the arithmetic is throughput bound at fixed occupancy, so a
  // group takes a few rounds per stage and a warp carries more triples.
  const int target = (width + lane_div() - 1) / lane_div();
  c.group_lanes = target <= 1 ? 1 : target <= 2 ? 2 : target <= 4 ? 4 : target <= 8 ? 8 : target <= 16 ? 16 : 32;
  c.groups_per_warp = 32 / c.group_lanes;
  const size_t per_warp = sizeof(double) * (size_t)c.ws_doubles * c.groups_per_warp;
  c.warps_per_cta = 8;
  while (c.warps_per_cta > 1 && per_warp * c.warps_per_cta > smem_limit / 4) c.warps_per_cta >>= 1;
  c.smem_bytes = per_warp * c.warps_per_cta;
  c.fits = c.smem_bytes <= smem_limit;
  return c;
}

// Which aux shells a request for the aux FUNCTIONS [f0, f1) touches, per angular momentum: shells [s0, s1) overlap
// the range (shells at the ends may be cut; the kernel writes only functions inside the range), and since by_l[l]
// lists the shells of one l in ascending order they are the contiguous piece [first[l], last[l]) of each list.
struct AuxShellRange {
  int first[LMAX_SHELL + 1], last[LMAX_SHELL + 1];
};
inline AuxShellRange aux_shell_range(const std::vector<int>& func0, const std::vector<int> (&by_l)[LMAX_SHELL + 1], int f0,
                                     int f1) {
  AuxShellRange r{};
  const int s0 = int(std::upper_bound(func0.begin(), func0.end(), f0) - func0.begin()) - 1;
  const int s1 = int(std::lower_bound(func0.begin(), func0.end(), f1) - func0.begin());
  for (int l = 0; l <= LMAX_SHELL; ++l) {
    r.first[l] = int(std::lower_bound(by_l[l].begin(), by_l[l].end(), s0) - by_l[l].begin());
    r.last[l] = int(std::lower_bound(by_l[l].begin(), by_l[l].end(), s1) - by_l[l].begin());
    if (f1 <= f0) r.last[l] = r.first[l];
  }
  return r;
}

// Launch lists of a basis: every unordered shell pair once, oriented so that l_a >= l_b, pairs without a surviving
// primitive pair dropped; or (unit = true) every shell with the unit partner.  Records go to pool.
struct PairLists {
  std::vector<PairEntry> entries;
  std::vector<double> pool;
  long long total = 0;  // pairs before screening
};
// Primitive-pair records of every shell pair (s >= t).  The pairs of a range of s are independent, so the list is
// built by a few host threads on contiguous ranges of s (balanced for the triangular loop) and concatenated in order:
// the result is the serial one bit for bit.  This is on the end-to-end path of every basis-set job (0.57 s single
// threaded for C60 / def2-tzvp).
inline PairLists make_pair_lists(const HostBasis& h, bool unit) {
  PairLists out;
  if (unit) {
    for (int s = 0; s < h.nshell; ++s) {
      ++out.total;
      PairEntry e{s, -1, 0, HostBasis::pair_record_doubles(h.l[s], 0), (long long)out.pool.size()};
      e.npp = h.append_pair_records(s, -1, out.pool);
      if (e.npp) out.entries.push_back(e);
    }
    return out;
  }
  auto build = [&h](int s0, int s1, PairLists& part) {
    for (int s = s0; s < s1; ++s)
      for (int t = 0; t <= s; ++t) {
        ++part.total;
        const bool swap = h.l[t] > h.l[s];
        const int x = swap ? t : s, y = swap ? s : t;
        PairEntry e{x, y, 0, HostBasis::pair_record_doubles(h.l[x], h.l[y]), (long long)part.pool.size()};
        e.npp = h.append_pair_records(x, y, part.pool);
        if (e.npp) part.entries.push_back(e);
      }
  };
  int nthr = (int)std::min<unsigned>(std::max(1u, std::thread::hardware_concurrency()), 16u);
  if (const char* env = std::getenv("GWBSE_HOST_THREADS")) nthr = std::max(1, std::atoi(env));
  if (h.nshell < 64) nthr = 1;
  if (nthr == 1) {
    build(0, h.nshell, out);
    return out;
  }
  // range boundaries at equal shares of the s (s + 1) / 2 pairs
  std::vector<int> cut(nthr + 1, 0);
  for (int i = 1; i < nthr; ++i)
    cut[i] = std::min(h.nshell, (int)std::lround(std::sqrt((double)i / nthr) * h.nshell));
  cut[nthr] = h.nshell;
  std::vector<PairLists> parts(nthr);
  std::vector<std::thread> pool;
  for (int i = 0; i < nthr; ++i) pool.emplace_back(build, cut[i], cut[i + 1], std::ref(parts[i]));
  for (auto& t : pool) t.join();
  size_t ne = 0, nd = 0;
  for (const PairLists& p : parts) {
    ne += p.entries.size();
    nd += p.pool.size();
  }
  out.entries.reserve(ne);
  out.pool.reserve(nd);
  for (const PairLists& p : parts) {
    const long long base = (long long)out.pool.size();
    for (PairEntry e : p.entries) {
      e.off += base;
      out.entries.push_back(e);
    }
    out.pool.insert(out.pool.end(), p.pool.begin(), p.pool.end());
    out.total += p.total;
  }
  return out;
}

}  // namespace ao
}  // namespace gwbse
