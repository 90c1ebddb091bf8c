// CPU harness of the device AO-integral code: runs votca_b200/csrc/ao3c_core.cuh - the same source the sm_100a
// kernel compiles - with one std::thread per lane and a std::barrier as the warp barrier.  Test infrastructure
// only (tests/test_ao3c_core_cpu.py); built with -fsanitize=thread it checks the barrier placement.
#include <algorithm>
#include <barrier>
#include <cstring>
#include <map>
#include <memory>
#include <thread>
#include <vector>

#include "../../votca_b200/csrc/ao3c_tables.h"

using namespace gwbse::ao;

namespace {

struct Tables {
  std::vector<double> boys, pure;
  std::vector<uint32_t> tuv;
  TableView view{};
  Tables() {
    boys = make_boys_table();
    tuv = make_tuv_table();
    int off = 0;
    for (int l = 0; l <= LMAX_SHELL; ++l) {
      view.pure_off[l] = off;
      std::vector<double> T = make_pure_matrix(l);
      pure.insert(pure.end(), T.begin(), T.end());
      off += (int)T.size();
    }
    view.boys = boys.data();
    view.tuv = tuv.data();
    view.pure = pure.data();
    view.boys_orders = BOYS_ORDERS;
    view.boys_taylor = BOYS_TAYLOR;
    view.herm1_stride = HERM1_STRIDE;
    view.herm1_dim = LMAX_SHELL + 1;
    view.boys_dx = BOYS_DX;
    view.boys_xmax = BOYS_XMAX;
  }
};

BasisView view_of(const HostBasis& b) {
  BasisView v{};
  v.nshell = b.nshell;
  v.nfunc = b.nfunc;
  v.l = b.l.data();
  v.np = b.np.data();
  v.prim0 = b.prim0.data();
  v.func0 = b.func0.data();
  v.center = b.center.data();
  v.exps = b.exps.data();
  v.coefs = b.coefs.data();
  v.herm1 = b.herm1.data();
  return v;
}

struct BarrierSync {
  std::barrier<>* b;
  void operator()() { b->arrive_and_wait(); }
};
struct NoSync {
  void operator()() {}
};

}  // namespace

extern "C" {

// out[k][nu][mu] (aux function, N x N column-major) for all aux shells; nl lanes (1 = plain serial run)
int ao3c_host(int nshell, const int* l, const int* nprim, const double* center, const double* exps, const double* coefs,
              int nshell_aux, const int* l_aux, const int* nprim_aux, const double* center_aux, const double* exps_aux,
              const double* coefs_aux, int nl, double* out) {
  try {
    static Tables tb;
    HostBasis dft, aux;
    dft.build(nshell, l, nprim, center, exps, coefs);
    aux.build(nshell_aux, l_aux, nprim_aux, center_aux, exps_aux, coefs_aux);
    const BasisView dv = view_of(dft), av = view_of(aux);
    const long long N = dft.nfunc;
    OutSpec spec{out, N * N, 1, N, 0, aux.nfunc, 1};
    const int wsd = workspace_doubles(dft.lmax, dft.lmax, aux.lmax);
    std::vector<double> ws((size_t)wsd + 8, -7.0e300);  // poisoned: stale reads show up in the result
    // as the launcher does (capi_ao3c.cu): shell pairs without a surviving primitive pair are not visited, their
    // blocks are the zeros of the initial fill
    std::fill(out, out + (size_t)aux.nfunc * N * N, 0.0);
    const PairLists pl = make_pair_lists(dft, false);
    if (nl <= 1) {
      NoSync s;
      for (const PairEntry& pe : pl.entries)
        for (int c = 0; c < aux.nshell; ++c) triple_block(dv, av, tb.view, pe, pl.pool.data(), c, ws.data(), 0, 1, s, spec);
      return 0;
    }
    std::barrier<> bar(nl);
    std::vector<std::thread> th;
    for (int lane = 0; lane < nl; ++lane)
      th.emplace_back([&, lane] {
        BarrierSync s{&bar};
        for (const PairEntry& pe : pl.entries)
          for (int c = 0; c < aux.nshell; ++c)
            triple_block(dv, av, tb.view, pe, pl.pool.data(), c, ws.data(), lane, nl, s, spec);
      });
    for (auto& t : th) t.join();
    return 0;
  } catch (...) {
    return 1;
  }
}

// The launcher's path for a request of aux functions [f0, f1) (capi_ao3c.cu: launch_classes): aux shells picked per
// angular momentum by aux_shell_range, every pair list entry against each of them, output restricted to the range.
// out[k - f0][nu][mu], pitch doubles between columns.
int ao3c_range_host(int nshell, const int* l, const int* nprim, const double* center, const double* exps,
                    const double* coefs, int nshell_aux, const int* l_aux, const int* nprim_aux, const double* center_aux,
                    const double* exps_aux, const double* coefs_aux, int f0, int f1, long pitch, double* out) {
  try {
    static Tables tb;
    HostBasis dft, aux;
    dft.build(nshell, l, nprim, center, exps, coefs);
    aux.build(nshell_aux, l_aux, nprim_aux, center_aux, exps_aux, coefs_aux);
    const BasisView dv = view_of(dft), av = view_of(aux);
    const long long N = dft.nfunc;
    if (pitch == 0) pitch = N;
    std::vector<int> by_l[LMAX_SHELL + 1];
    for (int s = 0; s < aux.nshell; ++s) by_l[aux.l[s]].push_back(s);
    const AuxShellRange r = aux_shell_range(aux.func0, by_l, f0, f1);
    OutSpec spec{out, pitch * N, 1, pitch, f0, f1, 1};
    std::fill(out, out + (size_t)std::max(f1 - f0, 0) * pitch * N, 0.0);
    std::vector<double> ws((size_t)workspace_doubles(dft.lmax, dft.lmax, aux.lmax) + 8, -7.0e300);
    const PairLists pl = make_pair_lists(dft, false);
    NoSync s;
    for (int lc = LMAX_SHELL; lc >= 0; --lc)
      for (const PairEntry& pe : pl.entries)
        for (int i = r.first[lc]; i < r.last[lc]; ++i)
          triple_block(dv, av, tb.view, pe, pl.pool.data(), by_l[lc][i], ws.data(), 0, 1, s, spec);
    return 0;
  } catch (...) {
    return 1;
  }
}

// The whole launch sequence of capi_ao3c.cu emulated on the CPU for a request of aux functions [f0, f1): pair entries
// grouped by class exactly as build_basis does, one "launch" per (l_c, class) with launch_config's geometry, every
// CTA / warp / lane of the grid run through cta_thread - 32 threads play the lanes of a warp and walk over all warps
// of all CTAs, lane groups synchronise on their own std::barrier.  Only the CUDA runtime calls are left out.
int ao3c_grid_host(int nshell, const int* l, const int* nprim, const double* center, const double* exps,
                   const double* coefs, int nshell_aux, const int* l_aux, const int* nprim_aux, const double* center_aux,
                   const double* exps_aux, const double* coefs_aux, int f0, int f1, long pitch, double* out) {
  try {
    static Tables tb;
    HostBasis dft, aux;
    dft.build(nshell, l, nprim, center, exps, coefs);
    aux.build(nshell_aux, l_aux, nprim_aux, center_aux, exps_aux, coefs_aux);
    const BasisView dv = view_of(dft), av = view_of(aux);
    const long long N = dft.nfunc;
    if (pitch == 0) pitch = N;
    std::vector<int> by_l[LMAX_SHELL + 1];
    for (int s = 0; s < aux.nshell; ++s) by_l[aux.l[s]].push_back(s);
    const AuxShellRange r = aux_shell_range(aux.func0, by_l, f0, f1);
    OutSpec spec{out, pitch * N, 1, pitch, f0, f1, 1};
    std::fill(out, out + (size_t)std::max(f1 - f0, 0) * pitch * N, 0.0);
    const PairLists pl = make_pair_lists(dft, false);
    std::map<std::pair<int, int>, std::vector<PairEntry>> classes;
    for (const PairEntry& e : pl.entries) classes[{dft.l[e.a], dft.l[e.b]}].push_back(e);
    const size_t smem_limit = 232448;
    struct HostSyncFactory {
      std::vector<std::unique_ptr<std::barrier<>>>* bars;
      BarrierSync operator()(int sub, int) const { return BarrierSync{(*bars)[sub].get()}; }
    };
    for (int lc = LMAX_SHELL; lc >= 0; --lc) {
      const int naux_shells = r.last[lc] - r.first[lc];
      if (naux_shells <= 0) continue;
      for (auto it = classes.rbegin(); it != classes.rend(); ++it) {
        const LaunchConfig cfg = launch_config(it->first.first, it->first.second, lc, smem_limit);
        if (!cfg.fits) return 2;
        const std::vector<PairEntry>& pairs = it->second;
        const long long groups = (long long)pairs.size() * naux_shells;
        const long long per_cta = (long long)cfg.warps_per_cta * cfg.groups_per_warp;
        const long long blocks = (groups + per_cta - 1) / per_cta;
        std::vector<double> smem(cfg.smem_bytes / sizeof(double), -7.0e300);
        std::vector<std::unique_ptr<std::barrier<>>> bars;
        for (int g = 0; g < cfg.groups_per_warp; ++g) bars.emplace_back(new std::barrier<>(cfg.group_lanes));
        std::vector<std::thread> lanes;
        for (int lane = 0; lane < 32; ++lane)
          lanes.emplace_back([&, lane] {
            HostSyncFactory sf{&bars};
            for (long long b = 0; b < blocks; ++b)
              for (int w = 0; w < cfg.warps_per_cta; ++w)
                cta_thread(b, cfg.warps_per_cta, w * 32 + lane, dv, av, tb.view, pairs.data(), (long long)pairs.size(),
                           pl.pool.data(), by_l[lc].data() + r.first[lc], naux_shells, spec, cfg.ws_doubles,
                           cfg.group_lanes, smem.data(), sf);
          });
        for (auto& t : lanes) t.join();
      }
    }
    return 0;
  } catch (...) {
    return 1;
  }
}

// Every triple with a scratch buffer of exactly workspace_doubles(la, lb, lc) doubles on the heap: built with
// -fsanitize=address this flags any access outside the region the launcher reserves per lane group.  Also runs the
// two-centre, overlap and dipole modes that way.  out as ao3c_host.
int ao3c_exact_scratch_host(int nshell, const int* l, const int* nprim, const double* center, const double* exps,
                            const double* coefs, int nshell_aux, const int* l_aux, const int* nprim_aux,
                            const double* center_aux, const double* exps_aux, const double* coefs_aux, double* out) {
  try {
    static Tables tb;
    HostBasis dft, aux;
    dft.build(nshell, l, nprim, center, exps, coefs);
    aux.build(nshell_aux, l_aux, nprim_aux, center_aux, exps_aux, coefs_aux);
    const BasisView dv = view_of(dft), av = view_of(aux);
    const long long N = dft.nfunc, M = aux.nfunc;
    OutSpec spec{out, N * N, 1, N, 0, aux.nfunc, 1};
    std::fill(out, out + (size_t)aux.nfunc * N * N, 0.0);
    NoSync s;
    const PairLists pl = make_pair_lists(dft, false);
    for (const PairEntry& pe : pl.entries)
      for (int c = 0; c < aux.nshell; ++c) {
        std::vector<double> ws((size_t)workspace_doubles(dft.l[pe.a], dft.l[pe.b], aux.l[c]));
        triple_block(dv, av, tb.view, pe, pl.pool.data(), c, ws.data(), 0, 1, s, spec);
      }
    std::vector<double> small((size_t)std::max(M * M, 3 * N * N));
    const PairLists units = make_pair_lists(aux, true), auxpairs = make_pair_lists(aux, false);
    OutSpec s2{small.data(), M, 1, 0, 0, aux.nfunc, 0};
    for (const PairEntry& pe : units.entries)
      for (int c = 0; c < aux.nshell; ++c) {
        std::vector<double> ws((size_t)workspace_doubles(aux.l[pe.a], 0, aux.l[c]));
        triple_block(av, av, tb.view, pe, units.pool.data(), c, ws.data(), 0, 1, s, s2);
      }
    for (int code = -1; code >= -4; --code) {
      OutSpec s1{small.data(), 0, 1, N, 0, 1, 1};
      for (const PairEntry& pe : pl.entries) {
        std::vector<double> ws((size_t)workspace_doubles(dft.l[pe.a], dft.l[pe.b], 0));
        triple_block(dv, dv, tb.view, pe, pl.pool.data(), code, ws.data(), 0, 1, s, s1);
      }
    }
    (void)auxpairs;
    return 0;
  } catch (...) {
    return 1;
  }
}

// V[q][p] = (p | q) over one basis: unit partner
int coulomb2c_host(int nshell, const int* l, const int* nprim, const double* center, const double* exps,
                   const double* coefs, double* out) {
  try {
    static Tables tb;
    HostBasis bs;
    bs.build(nshell, l, nprim, center, exps, coefs);
    const BasisView v = view_of(bs);
    const long long N = bs.nfunc;
    OutSpec spec{out, N, 1, 0, 0, bs.nfunc, 0};
    std::vector<double> ws((size_t)workspace_doubles(bs.lmax, 0, bs.lmax) + 8, -7.0e300);
    NoSync s;
    const PairLists pl = make_pair_lists(bs, true);
    for (const PairEntry& pe : pl.entries)
      for (int c = 0; c < bs.nshell; ++c) triple_block(v, v, tb.view, pe, pl.pool.data(), c, ws.data(), 0, 1, s, spec);
    return 0;
  } catch (...) {
    return 1;
  }
}

// S[nu][mu] = <mu | nu>
int overlap_host(int nshell, const int* l, const int* nprim, const double* center, const double* exps,
                 const double* coefs, double* out) {
  try {
    static Tables tb;
    HostBasis bs;
    bs.build(nshell, l, nprim, center, exps, coefs);
    const BasisView v = view_of(bs);
    const long long N = bs.nfunc;
    OutSpec spec{out, 0, 1, N, 0, 1, 1};
    std::vector<double> ws((size_t)workspace_doubles(bs.lmax, bs.lmax, 0) + 8, -7.0e300);
    NoSync s;
    std::fill(out, out + (size_t)N * N, 0.0);
    const PairLists pl = make_pair_lists(bs, false);
    for (const PairEntry& pe : pl.entries) triple_block(v, v, tb.view, pe, pl.pool.data(), -1, ws.data(), 0, 1, s, spec);
    return 0;
  } catch (...) {
    return 1;
  }
}

long surviving_pairs_host(int nshell, const int* l, const int* nprim, const double* center, const double* exps,
                          const double* coefs) {
  HostBasis bs;
  bs.build(nshell, l, nprim, center, exps, coefs);
  return (long)make_pair_lists(bs, false).entries.size();
}

int normalize_host(int l, int nprim, const double* exps, const double* raw, double* out) {
  normalize_contraction(l, nprim, exps, raw, out);
  return 0;
}

// per surviving shell pair: l of both shells and the number of primitive-pair records (workload statistics)
long pair_stats_host(int nshell, const int* l, const int* nprim, const double* center, const double* exps,
                     const double* coefs, long cap, int* la, int* lb, int* npp) {
  HostBasis bs;
  bs.build(nshell, l, nprim, center, exps, coefs);
  const PairLists pl = make_pair_lists(bs, false);
  long n = 0;
  for (const PairEntry& e : pl.entries) {
    if (n >= cap) break;
    la[n] = bs.l[e.a];
    lb[n] = bs.l[e.b];
    npp[n] = e.npp;
    ++n;
  }
  return (long)pl.entries.size();
}

// out[6]: group_lanes, groups_per_warp, warps_per_cta, ws_doubles, smem_bytes, fits
int launch_config_host(int la, int lb, int lc, long smem_limit, long* out) {
  const LaunchConfig c = launch_config(la, lb, lc, (size_t)smem_limit);
  out[0] = c.group_lanes;
  out[1] = c.groups_per_warp;
  out[2] = c.warps_per_cta;
  out[3] = c.ws_doubles;
  out[4] = (long)c.smem_bytes;
  out[5] = c.fits ? 1 : 0;
  return 0;
}

// D[k][nu][mu] = <mu | r_k | nu>, k = x, y, z, about the origin
int dipole_host(int nshell, const int* l, const int* nprim, const double* center, const double* exps,
                const double* coefs, double* out) {
  try {
    static Tables tb;
    HostBasis bs;
    bs.build(nshell, l, nprim, center, exps, coefs);
    const BasisView v = view_of(bs);
    const long long N = bs.nfunc;
    std::fill(out, out + (size_t)3 * N * N, 0.0);
    std::vector<double> ws((size_t)workspace_doubles(bs.lmax, bs.lmax, 0) + 8, -7.0e300);
    NoSync s;
    const PairLists pl = make_pair_lists(bs, false);
    for (int k = 0; k < 3; ++k) {
      OutSpec spec{out + (size_t)k * N * N, 0, 1, N, 0, 1, 1};
      for (const PairEntry& pe : pl.entries) triple_block(v, v, tb.view, pe, pl.pool.data(), -2 - k, ws.data(), 0, 1, s, spec);
    }
    return 0;
  } catch (...) {
    return 1;
  }
}

int boys_host(int n, double x, double* out) {
  static Tables tb;
  *out = boys_one(tb.view, n, x);
  return 0;
}

int pure_matrix_host(int l, double* out) {
  std::vector<double> T = make_pure_matrix(l);
  std::memcpy(out, T.data(), sizeof(double) * T.size());
  return 0;
}
}
