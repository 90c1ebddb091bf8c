// TEST INFRASTRUCTURE ONLY - a plain CPU stand-in for libgwbse_b200.so (include/gwbse_b200.h), single rank, naive
// loops, small problems.  SURVEY.md section 7 step 2: "a host-only reference implementation of the ABI *for tests
// only* (the product has no CPU fallback)".  It exists so that the C++ host layer (votca_b200/host: GW loop, QP
// search, BSE driver, Davidson, job facade, checkpoint writer, basis-set path) can run in the CPU test suite against
// the oracle.  Nothing under votca_b200/ builds, links or loads it; tests/test_host_layer_on_mock_cpu.py builds it
// into tests/host_harness/build/mock/ together with a copy of the host library linked against it and runs both in a
// separate process.  Arithmetic follows the same reference formulas the oracle cites (rpa.cc, sigma_base.cc,
// sigma_ppm.cc, sigma_exact.cc, bse_operator.cc, davidsonsolver.cc, threecenter.cc, aomatrix.cc).  Not covered:
// multi-GPU.  The small general eigenproblem of the full-BSE Davidson is delegated to the test
// process's LAPACK through mock_set_gen_eig.
#include <algorithm>
#include <cmath>
#include <complex>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/gwbse_b200.h"
#include "../../votca_b200/csrc/ao3c_tables.h"

using namespace gwbse;

struct gwbse_ctx {
  std::string err;
  long long launches = 0;
  int naux = 0, mmin = 0, mmax = -1, nmin = 0, nmax = -1, mtotal = 0, ntotal = 0;
  std::vector<double> X, Xsnap;  // element (m, n, chi) at X[(chi * mtotal + m) * ntotal + n]
  bool have_snapshot = false;
  std::vector<double> mos;
  int nbasis = 0, nmo = 0;
  std::vector<double> eps, invsqrt;
  // PPM evaluator
  bool ppm_ready = false;
  std::vector<double> fac, pole, energies;
  int lumo = 0, qpoff = 0, q = 0;
  double eta = 0.0;
  // exact Sigma: residues R[level][n][s], poles
  bool exact_ready = false;
  std::vector<double> residues, rpa_omegas, energies_exact;
  int ex_S = 0, ex_q = 0, ex_nocc = 0;
  double ex_eta = 0.0, ex_diag_pref = 2.0, ex_offdiag_pref = 1.0;
  // CDA: kappa matrices [order + 1][naux * naux] (last: kappa_0), quadrature, ranges
  bool cda_ready = false;
  gwbse_ctx* cda_partner = nullptr;
  int cda_partner_homo = 0;
  std::vector<double> cda_partner_e;
  std::vector<double> cda_kappa, cda_pts, cda_wts;
  int cda_order = 0, cda_sym = 0, cda_homo = 0, cda_rpamin = 0, cda_rpamax = 0, cda_qpmin = 0, cda_q = 0;
  double cda_alpha = 0.0, cda_eta = 0.0;
  // BSE
  bool bse_ready = false;
  int homo = 0, vt = 0, ct = 0, voff = 0, coff = 0;
  int bse_vmin = 0, bse_cmax = 0, bse_rpamin = 0;
  std::vector<double> eps_inv, hqp;
  double bse_flops = 0.0;
  long long bse_products = 0, bse_columns = 0;

  // QSGW rotation of the hole slices (rpa.h:59-66)
  std::vector<double> qsgw_U;
  int qsgw_q = 0, qsgw_qpmin = 0, qsgw_homo = 0;

  double& M(int m, int n, int chi) { return X[((size_t)chi * mtotal + m) * ntotal + n]; }
  // element (v, n, chi) of hole slice v as the RPA sums see it (rpa.cc:95-118)
  double hole(int v, int n, int chi) {
    if (!qsgw_U.empty()) {
      const int off = qsgw_qpmin - mmin, end_occ = std::min(qsgw_homo - mmin + 1, off + qsgw_q);
      if (v >= off && v < end_occ) {
        double s = 0.0;
        for (int vp = 0; vp < qsgw_q; ++vp) s += qsgw_U[vp + (size_t)(v - off) * qsgw_q] * M(vp + off, n, chi);
        return s;
      }
    }
    return M(v, n, chi);
  }
};

struct gwbse_basis {
  ao::HostBasis host;
  ao::PairLists pairs, unit_pairs;
};

namespace {

std::string g_create_error;

#define MOCK_BEGIN(ctx) \
  if (!(ctx)) return 1; \
  try {                 \
    (ctx)->launches++;
#define MOCK_END(ctx)               \
  return 0;                         \
  }                                 \
  catch (const std::exception& e) { \
    (ctx)->err = e.what();          \
    return 1;                       \
  }
#define REQUIRE(cond, msg) \
  if (!(cond)) throw std::runtime_error(msg)

void gemm(char ta, char tb, int m, int n, int k, double alpha, const double* A, long lda, const double* B, long ldb,
          double beta, double* C, long ldc) {
  const bool tA = ta == 'T' || ta == 't', tB = tb == 'T' || tb == 't';
  for (int j = 0; j < n; ++j)
    for (int i = 0; i < m; ++i) {
      double s = 0.0;
      for (int l = 0; l < k; ++l) s += (tA ? A[l + i * lda] : A[i + l * lda]) * (tB ? B[j + l * ldb] : B[l + j * ldb]);
      C[i + j * ldc] = alpha * s + (beta == 0.0 ? 0.0 : beta * C[i + j * ldc]);
    }
}

// cyclic Jacobi: A (n x n, lower/upper both read) -> eigenvectors in the columns of A, eigenvalues ascending in w
void sym_eig(int n, double* A, long lda, double* w) {
  std::vector<double> a((size_t)n * n), v((size_t)n * n, 0.0);
  for (int j = 0; j < n; ++j)
    for (int i = 0; i < n; ++i) a[i + (size_t)j * n] = i >= j ? A[i + j * lda] : A[j + i * lda];  // lower triangle
  for (int i = 0; i < n; ++i) v[i + (size_t)i * n] = 1.0;
  for (int sweep = 0; sweep < 100; ++sweep) {
    double off = 0.0, diag = 0.0;
    for (int j = 0; j < n; ++j)
      for (int i = 0; i < n; ++i) (i == j ? diag : off) += a[i + (size_t)j * n] * a[i + (size_t)j * n];
    if (off <= 1e-30 * (diag + 1e-300)) break;
    for (int p = 0; p < n - 1; ++p)
      for (int q = p + 1; q < n; ++q) {
        const double apq = a[p + (size_t)q * n];
        if (std::fabs(apq) < 1e-300) continue;
        const double theta = (a[q + (size_t)q * n] - a[p + (size_t)p * n]) / (2.0 * apq);
        const double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
        const double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
        for (int k = 0; k < n; ++k) {
          const double akp = a[k + (size_t)p * n], akq = a[k + (size_t)q * n];
          a[k + (size_t)p * n] = c * akp - s * akq;
          a[k + (size_t)q * n] = s * akp + c * akq;
        }
        for (int k = 0; k < n; ++k) {
          const double apk = a[p + (size_t)k * n], aqk = a[q + (size_t)k * n];
          a[p + (size_t)k * n] = c * apk - s * aqk;
          a[q + (size_t)k * n] = s * apk + c * aqk;
        }
        for (int k = 0; k < n; ++k) {
          const double vkp = v[k + (size_t)p * n], vkq = v[k + (size_t)q * n];
          v[k + (size_t)p * n] = c * vkp - s * vkq;
          v[k + (size_t)q * n] = s * vkp + c * vkq;
        }
      }
  }
  std::vector<int> order(n);
  for (int i = 0; i < n; ++i) order[i] = i;
  std::sort(order.begin(), order.end(), [&](int x, int y) { return a[x + (size_t)x * n] < a[y + (size_t)y * n]; });
  for (int j = 0; j < n; ++j) {
    w[j] = a[order[j] + (size_t)order[j] * n];
    for (int i = 0; i < n; ++i) A[i + j * lda] = v[i + (size_t)order[j] * n];
  }
}

// Gauss-Jordan with partial pivoting: B <- A^-1 B (A destroyed)
void solve(int n, int nrhs, double* A, long lda, double* B, long ldb) {
  for (int c = 0; c < n; ++c) {
    int piv = c;
    for (int r = c + 1; r < n; ++r)
      if (std::fabs(A[r + c * lda]) > std::fabs(A[piv + c * lda])) piv = r;
    if (A[piv + c * lda] == 0.0) throw std::runtime_error("LU factorisation failed (singular matrix)");
    if (piv != c) {
      for (int j = 0; j < n; ++j) std::swap(A[c + j * lda], A[piv + j * lda]);
      for (int j = 0; j < nrhs; ++j) std::swap(B[c + j * ldb], B[piv + j * ldb]);
    }
    const double inv = 1.0 / A[c + c * lda];
    for (int j = 0; j < n; ++j) A[c + j * lda] *= inv;
    for (int j = 0; j < nrhs; ++j) B[c + j * ldb] *= inv;
    for (int r = 0; r < n; ++r) {
      if (r == c) continue;
      const double f = A[r + c * lda];
      if (f == 0.0) continue;
      for (int j = 0; j < n; ++j) A[r + j * lda] -= f * A[c + j * lda];
      for (int j = 0; j < nrhs; ++j) B[r + j * ldb] -= f * B[c + j * ldb];
    }
  }
}

void require_mmn(gwbse_ctx* ctx) { REQUIRE(!ctx->X.empty(), "Mmn not allocated (gwbse_mmn_alloc)"); }

void mul_right(gwbse_ctx* ctx, const double* R, long ldr) {
  require_mmn(ctx);
  const int naux = ctx->naux;
  const size_t rows = (size_t)ctx->mtotal * ctx->ntotal;
  std::vector<double> out(ctx->X.size(), 0.0);
  for (int c2 = 0; c2 < naux; ++c2)
    for (int c1 = 0; c1 < naux; ++c1) {
      const double r = R[c1 + c2 * ldr];
      if (r == 0.0) continue;
      const double* src = &ctx->X[(size_t)c1 * rows];
      double* dst = &out[(size_t)c2 * rows];
      for (size_t i = 0; i < rows; ++i) dst[i] += src[i] * r;
    }
  ctx->X.swap(out);
}

void fill_block(gwbse_ctx* ctx, int aux_offset, int aux_count, const double* ao3c, long pitch) {
  require_mmn(ctx);
  REQUIRE(!ctx->mos.empty(), "MO coefficients not set (gwbse_mmn_set_mos)");
  REQUIRE(aux_offset >= 0 && aux_offset + aux_count <= ctx->naux, "aux block out of range");
  REQUIRE(ctx->mmax < ctx->nmo && ctx->nmax < ctx->nmo, "level range exceeds number of MOs");
  const int N = ctx->nbasis;
  if (pitch == 0) pitch = N;
  std::vector<double> H((size_t)N * ctx->mtotal);
  for (int k = 0; k < aux_count; ++k) {
    const double* T = ao3c + (size_t)k * pitch * N;
    gemm('N', 'N', N, ctx->mtotal, N, 1.0, T, pitch, &ctx->mos[(size_t)ctx->mmin * N], N, 0.0, H.data(), N);
    for (int m = 0; m < ctx->mtotal; ++m)
      for (int n = 0; n < ctx->ntotal; ++n) {
        double s = 0.0;
        const double* Cn = &ctx->mos[(size_t)(ctx->nmin + n) * N];
        for (int mu = 0; mu < N; ++mu) s += Cn[mu] * H[mu + (size_t)m * N];
        ctx->M(m, n, aux_offset + k) = s;
      }
  }
}

// Sigma_c(level, w) of the plasmon-pole model and its derivative (sigma_ppm.cc:37-91)
void ppm_eval(gwbse_ctx* ctx, int level, double w, double* sigma, double* dsigma) {
  REQUIRE(ctx->ppm_ready, "sigma evaluator not prepared");
  REQUIRE(level >= 0 && level < ctx->q, "gw_level out of range");
  const int m = ctx->qpoff + level;
  const double eta2 = ctx->eta * ctx->eta;
  double s = 0.0, ds = 0.0;
  for (int chi = 0; chi < ctx->naux; ++chi) {
    if (ctx->fac[chi] == 0.0) continue;
    for (int n = 0; n < ctx->ntotal; ++n) {
      const double t = w - ctx->energies[n] + (n < ctx->lumo ? ctx->pole[chi] : -ctx->pole[chi]);
      const double Mv = ctx->M(m, n, chi), den = t * t + eta2;
      s += 0.5 * ctx->fac[chi] * Mv * Mv * t / den;
      ds += 0.5 * ctx->fac[chi] * Mv * Mv * (eta2 - t * t) / (den * den);
    }
  }
  *sigma = s;
  if (dsigma) *dsigma = ds;
}

// Sigma_Exact::CalcCorrelationDiagElement / ...Derivative, sigma_exact.cc:40-83
void exact_eval(gwbse_ctx* ctx, int level, double w, double* sigma, double* dsigma) {
  REQUIRE(ctx->exact_ready, "sigma evaluator not prepared");
  REQUIRE(level >= 0 && level < ctx->ex_q, "gw_level out of range");
  const double eta2 = ctx->ex_eta * ctx->ex_eta;
  const int S = ctx->ex_S, ntot = ctx->ntotal;
  double s = 0.0, ds = 0.0;
  for (int n = 0; n < ntot; ++n)
    for (int p = 0; p < S; ++p) {
      const double t = w - ctx->energies_exact[n] + (n < ctx->ex_nocc ? ctx->rpa_omegas[p] : -ctx->rpa_omegas[p]);
      const double r = ctx->residues[((size_t)level * ntot + n) * S + p], den = t * t + eta2;
      s += ctx->ex_diag_pref * r * r * t / den;
      ds += ctx->ex_diag_pref * r * r * (eta2 - t * t) / (den * den);
    }
  *sigma = s;
  if (dsigma) *dsigma = ds;
}

void bse_matmul(gwbse_ctx* ctx, int cqp, int cx, int cd, int cd2, int k, const double* X, long ldx, double* Y, long ldy) {
  REQUIRE(ctx->bse_ready, "BSE operator not configured (gwbse_bse_configure)");
  REQUIRE(!(cd != 0 && cd2 != 0), "Hamiltonian cannot contain Hd and Hd2 at the same time");
  const int vt = ctx->vt, ct = ctx->ct, B = vt * ct, naux = ctx->naux, vo = ctx->voff, co = ctx->coff, hs = vt + ct;
  REQUIRE(ldx >= B && ldy >= B, "Shape mismatch in BSE matmul");
  // dense H, bse_operator.cc:40-131; index I = ct * v + c
  std::vector<double> H((size_t)B * B, 0.0);
  for (int v1 = 0; v1 < vt; ++v1)
    for (int c1 = 0; c1 < ct; ++c1)
      for (int v2 = 0; v2 < vt; ++v2)
        for (int c2 = 0; c2 < ct; ++c2) {
          double h = 0.0;
          for (int chi = 0; chi < naux; ++chi) {
            if (cx) h += cx * ctx->M(vo + v1, co + c1, chi) * ctx->M(vo + v2, co + c2, chi);
            if (cd) h -= cd * ctx->M(co + c1, co + c2, chi) * ctx->eps_inv[chi] * ctx->M(vo + v1, vo + v2, chi);
            if (cd2) h -= cd2 * ctx->M(co + c1, vo + v2, chi) * ctx->eps_inv[chi] * ctx->M(vo + v1, co + c2, chi);
          }
          if (cqp) {
            if (v1 == v2) h += cqp * ctx->hqp[(vt + c2) + (size_t)(vt + c1) * hs];
            if (c1 == c2) h -= cqp * ctx->hqp[v2 + (size_t)v1 * hs];
          }
          H[(size_t)(ct * v1 + c1) + (size_t)(ct * v2 + c2) * B] = h;
        }
  std::vector<double> out((size_t)B * std::max(k, 1));
  gemm('N', 'N', B, k, B, 1.0, H.data(), B, X, ldx, 0.0, out.data(), B);
  for (int j = 0; j < k; ++j) std::memcpy(Y + (size_t)j * ldy, &out[(size_t)j * B], sizeof(double) * B);
  ctx->bse_products++;
  ctx->bse_columns += k;
}

ao::TableView& tables() {
  static std::vector<double> boys = ao::make_boys_table(), pure;
  static std::vector<uint32_t> tuv = ao::make_tuv_table();
  static ao::TableView view{};
  if (!view.boys) {
    for (int l = 0; l <= ao::LMAX_SHELL; ++l) {
      view.pure_off[l] = (int)pure.size();
      const std::vector<double> T = ao::make_pure_matrix(l);
      pure.insert(pure.end(), T.begin(), T.end());
    }
    view.boys = boys.data();
    view.tuv = tuv.data();
    view.pure = pure.data();
    view.boys_orders = ao::BOYS_ORDERS;
    view.boys_taylor = ao::BOYS_TAYLOR;
    view.herm1_stride = ao::HERM1_STRIDE;
    view.herm1_dim = ao::LMAX_SHELL + 1;
    view.boys_dx = ao::BOYS_DX;
    view.boys_xmax = ao::BOYS_XMAX;
  }
  return view;
}

ao::BasisView view_of(const ao::HostBasis& b) {
  ao::BasisView v{};
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

struct NoSync {
  void operator()() {}
};

// every pair of the list against aux shells sc0..sc1 (or the single code sc0 < 0), serial
void integrals(const gwbse_basis& orb, const ao::PairLists& pl, const gwbse_basis& aux, int sc0, int sc1,
               const ao::OutSpec& out, int lmax_aux) {
  const ao::BasisView ov = view_of(orb.host), av = view_of(aux.host);
  std::vector<double> ws((size_t)ao::workspace_doubles(orb.host.lmax, orb.host.lmax, lmax_aux) + 8);
  NoSync s;
  for (const ao::PairEntry& pe : pl.entries)
    for (int sc = sc0; sc < sc1; ++sc) ao::triple_block(ov, av, tables(), pe, pl.pool.data(), sc, ws.data(), 0, 1, s, out);
}

void ao3c_block(const gwbse_basis* aux, const gwbse_basis* dft, int off, int cnt, double* out, long pitch) {
  REQUIRE(aux && dft && out, "null argument");
  REQUIRE(off >= 0 && cnt >= 0 && off + cnt <= aux->host.nfunc, "aux function range out of bounds");
  const long long N = dft->host.nfunc;
  if (pitch == 0) pitch = N;
  std::fill(out, out + (size_t)cnt * pitch * N, 0.0);
  ao::OutSpec spec{out, pitch * N, 1, pitch, off, off + cnt, 1};
  integrals(*dft, dft->pairs, *aux, 0, aux->host.nshell, spec, aux->host.lmax);
}

}  // namespace

extern "C" {

// ---- context / misc ------------------------------------------------------------------------------------------
int gwbse_ctx_create(int, gwbse_ctx** out) {
  if (!out) return 1;
  *out = new gwbse_ctx;
  return 0;
}
void gwbse_ctx_destroy(gwbse_ctx* ctx) { delete ctx; }
const char* gwbse_last_error(const gwbse_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }
const char* gwbse_create_error(void) { return g_create_error.c_str(); }
int gwbse_sync(gwbse_ctx*) { return 0; }
long long gwbse_launch_count(const gwbse_ctx* ctx) { return ctx ? ctx->launches : 0; }
int gwbse_profile_report(gwbse_ctx*, char* buf, size_t n) {
  if (buf && n) buf[0] = 0;
  return 0;
}
int gwbse_device_count(void) { return 0; }
int gwbse_set_option(gwbse_ctx*, const char*, double) { return 0; }
int gwbse_gemm_profile(gwbse_ctx*, int) { return 0; }
int gwbse_gemm_stats(gwbse_ctx*, double* ms, double* fl, long long* l) {
  if (ms) *ms = 0;
  if (fl) *fl = 0;
  if (l) *l = 0;
  return 0;
}
int gwbse_gemm_shape_report(gwbse_ctx*, char* buf, size_t n) {
  if (buf && n) buf[0] = 0;
  return 0;
}
int gwbse_fp64_peak_probe(gwbse_ctx*, double* t) {
  if (t) *t = 0;
  return 0;
}
int gwbse_timer_start(gwbse_ctx*) { return 0; }
int gwbse_timer_stop_ms(gwbse_ctx*, float* ms) {
  if (ms) *ms = 0;
  return 0;
}
int gwbse_nccl_unique_id(unsigned char* id) {
  if (id) std::memset(id, 0, 128);
  return 0;
}
int gwbse_comm_init(gwbse_ctx* ctx, int, int world, const unsigned char*) {
  if (world != 1 && ctx) ctx->err = "the test mock is single-rank";
  return world == 1 ? 0 : 1;
}
int gwbse_comm_rank(const gwbse_ctx*) { return 0; }
int gwbse_comm_world(const gwbse_ctx*) { return 1; }
int gwbse_comm_allreduce_host(gwbse_ctx*, double*, size_t) { return 0; }
int gwbse_shard_owner(int m, int world) { return world > 0 ? m % world : 0; }
int gwbse_shard_local_index(int m, int world) { return world > 0 ? m / world : m; }
int gwbse_shard_local_count(int total, int rank, int world) { return total <= rank ? 0 : (total - rank + world - 1) / world; }
int gwbse_shard_aux_begin(int naux, int rank, int world) { return (int)((long long)rank * naux / world); }
int gwbse_shard_aux_range(const gwbse_ctx* ctx, int rank, int* b, int* e) {
  if (!ctx || rank != 0) return 1;
  if (b) *b = 0;
  if (e) *e = ctx->naux;
  return 0;
}

// ---- memory ----------------------------------------------------------------------------------------------------
int gwbse_dev_malloc(gwbse_ctx*, size_t bytes, double** out) {
  *out = static_cast<double*>(std::malloc(std::max<size_t>(bytes, 8)));
  return *out ? 0 : 1;
}
int gwbse_dev_free(gwbse_ctx*, double* p) {
  std::free(p);
  return 0;
}
int gwbse_h2d(gwbse_ctx*, double* d, const double* s, size_t n) {
  std::memcpy(d, s, n * sizeof(double));
  return 0;
}
int gwbse_d2h(gwbse_ctx*, double* d, const double* s, size_t n) {
  std::memcpy(d, s, n * sizeof(double));
  return 0;
}
int gwbse_d2d(gwbse_ctx*, double* d, const double* s, size_t n) {
  std::memmove(d, s, n * sizeof(double));
  return 0;
}
int gwbse_dev_memset_zero(gwbse_ctx*, double* d, size_t n) {
  std::memset(d, 0, n * sizeof(double));
  return 0;
}
int gwbse_dev_mem_info(gwbse_ctx*, size_t* f, size_t* t) {
  if (f) *f = (size_t)1 << 34;
  if (t) *t = (size_t)1 << 34;
  return 0;
}
int gwbse_host_malloc(size_t bytes, void** out) {
  *out = std::malloc(std::max<size_t>(bytes, 8));
  return *out ? 0 : 1;
}
int gwbse_host_free(void* p) {
  std::free(p);
  return 0;
}

// ---- dense primitives ----------------------------------------------------------------------------------------------
int gwbse_dgemm_dev(gwbse_ctx* ctx, char ta, char tb, int m, int n, int k, double alpha, const double* A, int lda,
                    const double* B, int ldb, double beta, double* C, int ldc) {
  MOCK_BEGIN(ctx)
  const bool tA = ta == 'T' || ta == 't', tB = tb == 'T' || tb == 't';
  REQUIRE(m >= 0 && n >= 0 && k >= 0 && ldc >= std::max(m, 1) && lda >= std::max(tA ? k : m, 1) &&
              ldb >= std::max(tB ? n : k, 1),
          "Shape mismatch in Cublas gemm");
  gemm(ta, tb, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc);
  MOCK_END(ctx)
}
int gwbse_dgemm_dev_ex(gwbse_ctx* ctx, char ta, char tb, int m, int n, int k, double alpha, const double* A, int lda,
                       const double* B, int ldb, double beta, double* C, int ldc, int, int) {
  return gwbse_dgemm_dev(ctx, ta, tb, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc);
}
int gwbse_diag_scale_dev(gwbse_ctx* ctx, char side, int m, int n, const double* A, int lda, const double* d, double* C,
                         int ldc) {
  MOCK_BEGIN(ctx)
  const bool right = side == 'R' || side == 'r';
  for (int j = 0; j < n; ++j)
    for (int i = 0; i < m; ++i) C[i + (size_t)j * ldc] = A[i + (size_t)j * lda] * (right ? d[j] : d[i]);
  MOCK_END(ctx)
}
int gwbse_axpy_dev(gwbse_ctx* ctx, int m, int n, double alpha, const double* X, int ldx, double* Y, int ldy) {
  MOCK_BEGIN(ctx)
  for (int j = 0; j < n; ++j)
    for (int i = 0; i < m; ++i) Y[i + (size_t)j * ldy] += alpha * X[i + (size_t)j * ldx];
  MOCK_END(ctx)
}
int gwbse_colnorms_dev(gwbse_ctx* ctx, int m, int n, const double* A, int lda, double* norms) {
  MOCK_BEGIN(ctx)
  for (int j = 0; j < n; ++j) {
    double s = 0.0;
    for (int i = 0; i < m; ++i) s += A[i + (size_t)j * lda] * A[i + (size_t)j * lda];
    norms[j] = std::sqrt(s);
  }
  MOCK_END(ctx)
}
int gwbse_coldots_dev(gwbse_ctx* ctx, int m, int n, const double* X, int ldx, const double* Y, int ldy, double* dots) {
  MOCK_BEGIN(ctx)
  for (int j = 0; j < n; ++j) {
    double s = 0.0;
    for (int i = 0; i < m; ++i) s += X[i + (size_t)j * ldx] * Y[i + (size_t)j * ldy];
    dots[j] = s;
  }
  MOCK_END(ctx)
}
int gwbse_scale_cols_dev(gwbse_ctx* ctx, int m, int n, double* A, int lda, const double* s) {
  MOCK_BEGIN(ctx)
  for (int j = 0; j < n; ++j)
    for (int i = 0; i < m; ++i) A[i + (size_t)j * lda] *= s[j];
  MOCK_END(ctx)
}
int gwbse_sym_eig_dev(gwbse_ctx* ctx, int n, double* A, int lda, double* w) {
  MOCK_BEGIN(ctx)
  if (n > 0) sym_eig(n, A, lda, w);
  MOCK_END(ctx)
}
int gwbse_inverse_dev(gwbse_ctx* ctx, int n, double* A, int lda) {
  MOCK_BEGIN(ctx)
  std::vector<double> I((size_t)n * n, 0.0);
  for (int i = 0; i < n; ++i) I[i + (size_t)i * n] = 1.0;
  solve(n, n, A, lda, I.data(), n);
  for (int j = 0; j < n; ++j) std::memcpy(A + (size_t)j * lda, &I[(size_t)j * n], sizeof(double) * n);
  MOCK_END(ctx)
}
int gwbse_lu_solve_dev(gwbse_ctx* ctx, int n, int nrhs, double* A, int lda, double* B, int ldb) {
  MOCK_BEGIN(ctx)
  if (n > 0 && nrhs > 0) solve(n, nrhs, A, lda, B, ldb);
  MOCK_END(ctx)
}
// The real non-symmetric generalized eigenproblem of the harmonic Ritz step is LAPACK work (the product calls
// cuSOLVER Xgeev); the test process lends its own LAPACK (scipy) through this hook, see tests/conftest.py.
typedef int (*mock_gen_eig_fn)(int n, const double* T, const double* B, double* wr, double* wi, double* VR);
static mock_gen_eig_fn g_gen_eig = nullptr;
__attribute__((visibility("default"))) void mock_set_gen_eig(mock_gen_eig_fn fn) { g_gen_eig = fn; }
int gwbse_gen_eig_host(gwbse_ctx* ctx, int n, const double* T, const double* B, double* wr, double* wi, double* VR) {
  MOCK_BEGIN(ctx)
  REQUIRE(g_gen_eig != nullptr, "the generalized eigenproblem needs the LAPACK hook of the test process");
  REQUIRE(g_gen_eig(n, T, B, wr, wi, VR) == 0, "Small generalized eigenvalue problem failed.");
  MOCK_END(ctx)
}

// ---- Mmn ---------------------------------------------------------------------------------------------------------------
int gwbse_mmn_alloc(gwbse_ctx* ctx, int naux, int mmin, int mmax, int nmin, int nmax) {
  MOCK_BEGIN(ctx)
  REQUIRE(naux > 0 && mmax >= mmin && nmax >= nmin && mmin >= 0 && nmin >= 0, "invalid Mmn dimensions");
  ctx->naux = naux;
  ctx->mmin = mmin;
  ctx->mmax = mmax;
  ctx->nmin = nmin;
  ctx->nmax = nmax;
  ctx->mtotal = mmax - mmin + 1;
  ctx->ntotal = nmax - nmin + 1;
  ctx->X.assign((size_t)naux * ctx->mtotal * ctx->ntotal, 0.0);
  ctx->have_snapshot = false;
  ctx->ppm_ready = ctx->bse_ready = false;
  MOCK_END(ctx)
}
int gwbse_mmn_free(gwbse_ctx* ctx) {
  MOCK_BEGIN(ctx)
  ctx->X.clear();
  ctx->Xsnap.clear();
  MOCK_END(ctx)
}
int gwbse_mmn_dims(const gwbse_ctx* ctx, int* naux, int* mtotal, int* ntotal, int* mlocal, int* npad) {
  if (!ctx) return 1;
  if (naux) *naux = ctx->naux;
  if (mtotal) *mtotal = ctx->mtotal;
  if (ntotal) *ntotal = ctx->ntotal;
  if (mlocal) *mlocal = ctx->mtotal;
  if (npad) *npad = ctx->ntotal;
  return 0;
}
int gwbse_mmn_set_mos(gwbse_ctx* ctx, const double* mos, int ldmos, int nbasis, int nmo) {
  MOCK_BEGIN(ctx)
  REQUIRE(mos && ldmos >= nbasis && nbasis > 0 && nmo > 0, "invalid MO matrix");
  ctx->mos.resize((size_t)nbasis * nmo);
  for (int j = 0; j < nmo; ++j) std::memcpy(&ctx->mos[(size_t)j * nbasis], mos + (size_t)j * ldmos, sizeof(double) * nbasis);
  ctx->nbasis = nbasis;
  ctx->nmo = nmo;
  MOCK_END(ctx)
}
int gwbse_mmn_fill_block(gwbse_ctx* ctx, int aux_offset, int aux_count, const double* ao3c) {
  MOCK_BEGIN(ctx)
  fill_block(ctx, aux_offset, aux_count, ao3c, 0);
  MOCK_END(ctx)
}
int gwbse_mmn_fill_block_dev(gwbse_ctx* ctx, int aux_offset, int aux_count, const double* ao3c) {
  return gwbse_mmn_fill_block(ctx, aux_offset, aux_count, ao3c);
}
int gwbse_mmn_fill_begin(gwbse_ctx* ctx, int) { return ctx ? 0 : 1; }
int gwbse_mmn_fill_end(gwbse_ctx* ctx) { return ctx ? 0 : 1; }
int gwbse_mmn_mul_right(gwbse_ctx* ctx, const double* R, int ldr) {
  MOCK_BEGIN(ctx)
  REQUIRE(ldr >= ctx->naux, "Shape mismatch in MultiplyRightWithAuxMatrix");
  mul_right(ctx, R, ldr);
  MOCK_END(ctx)
}
int gwbse_mmn_mul_right_dev(gwbse_ctx* ctx, const double* R, int ldr) { return gwbse_mmn_mul_right(ctx, R, ldr); }
// the device library defers the rows outside the window; the result is the same product
int gwbse_mmn_mul_right_window_dev(gwbse_ctx* ctx, const double* R, int ldr, int, int) {
  return gwbse_mmn_mul_right(ctx, R, ldr);
}
int gwbse_mmn_rotate(gwbse_ctx* ctx, const double* U, int ldu, int qpmin, int qpmax) {
  MOCK_BEGIN(ctx)
  require_mmn(ctx);
  const int q = qpmax - qpmin + 1, on = qpmin - ctx->nmin, om = qpmin - ctx->mmin;
  REQUIRE(q > 0 && on >= 0 && om >= 0 && on + q <= ctx->ntotal && om + q <= ctx->mtotal, "QP window outside Mmn");
  std::vector<double> col(q);
  for (int m = 0; m < q; ++m)
    for (int chi = 0; chi < ctx->naux; ++chi) {
      for (int i = 0; i < q; ++i) {
        double s = 0.0;
        for (int j = 0; j < q; ++j) s += U[j + (size_t)i * ldu] * ctx->M(m + om, on + j, chi);  // U^T
        col[i] = s;
      }
      for (int i = 0; i < q; ++i) ctx->M(m + om, on + i, chi) = col[i];
    }
  MOCK_END(ctx)
}
int gwbse_pseudo_invsqrt(gwbse_ctx* ctx, int n, const double* S, const double* V, double etol, double* L_out, int* removed) {
  MOCK_BEGIN(ctx)
  // AOCoulomb::Pseudo_InvSqrt_GWBSE, aomatrix.cc:53-86
  int rem = 0;
  std::vector<double> U(S, S + (size_t)n * n), w(n), Ssqrt((size_t)n * n), tmp((size_t)n * n), ortho((size_t)n * n);
  sym_eig(n, U.data(), n, w.data());
  std::vector<double> Ud = U;
  for (int j = 0; j < n; ++j) {
    const double d = w[j] < etol ? (++rem, 0.0) : 1.0 / std::sqrt(w[j]);
    for (int i = 0; i < n; ++i) Ud[i + (size_t)j * n] *= d;
  }
  gemm('N', 'T', n, n, n, 1.0, Ud.data(), n, U.data(), n, 0.0, Ssqrt.data(), n);
  gemm('N', 'N', n, n, n, 1.0, Ssqrt.data(), n, V, n, 0.0, tmp.data(), n);
  gemm('N', 'N', n, n, n, 1.0, tmp.data(), n, Ssqrt.data(), n, 0.0, ortho.data(), n);
  sym_eig(n, ortho.data(), n, w.data());
  std::vector<double> Od = ortho, Vm1((size_t)n * n);
  for (int j = 0; j < n; ++j) {
    const double d = w[j] < etol ? (++rem, 0.0) : 1.0 / std::sqrt(w[j]);
    for (int i = 0; i < n; ++i) Od[i + (size_t)j * n] *= d;
  }
  gemm('N', 'T', n, n, n, 1.0, Od.data(), n, ortho.data(), n, 0.0, Vm1.data(), n);
  ctx->invsqrt.assign((size_t)n * n, 0.0);
  gemm('T', 'T', n, n, n, 1.0, Ssqrt.data(), n, Vm1.data(), n, 0.0, ctx->invsqrt.data(), n);  // (Vm1 Ssqrt)^T
  if (L_out) std::memcpy(L_out, ctx->invsqrt.data(), sizeof(double) * n * n);
  if (removed) *removed = rem;
  MOCK_END(ctx)
}
const double* gwbse_pseudo_invsqrt_result_dev(gwbse_ctx* ctx) { return ctx ? ctx->invsqrt.data() : nullptr; }
int gwbse_mmn_get_slice(gwbse_ctx* ctx, int m, double* out, int ld) {
  MOCK_BEGIN(ctx)
  require_mmn(ctx);
  REQUIRE(m >= 0 && m < ctx->mtotal && ld >= ctx->ntotal, "slice out of range");
  for (int chi = 0; chi < ctx->naux; ++chi)
    for (int n = 0; n < ctx->ntotal; ++n) out[n + (size_t)chi * ld] = ctx->M(m, n, chi);
  MOCK_END(ctx)
}
int gwbse_mmn_set_slice(gwbse_ctx* ctx, int m, const double* in, int ld) {
  MOCK_BEGIN(ctx)
  require_mmn(ctx);
  REQUIRE(m >= 0 && m < ctx->mtotal && ld >= ctx->ntotal, "slice out of range");
  for (int chi = 0; chi < ctx->naux; ++chi)
    for (int n = 0; n < ctx->ntotal; ++n) ctx->M(m, n, chi) = in[n + (size_t)chi * ld];
  MOCK_END(ctx)
}
int gwbse_mmn_snapshot(gwbse_ctx* ctx) {
  MOCK_BEGIN(ctx)
  ctx->Xsnap = ctx->X;
  ctx->have_snapshot = true;
  MOCK_END(ctx)
}
int gwbse_mmn_restore(gwbse_ctx* ctx) {
  MOCK_BEGIN(ctx)
  REQUIRE(ctx->have_snapshot, "no Mmn snapshot to restore");
  ctx->X = ctx->Xsnap;
  MOCK_END(ctx)
}

// ---- RPA ---------------------------------------------------------------------------------------------------------------
int gwbse_rpa_epsilon(gwbse_ctx* ctx, int kind, double fre, double fim, double eta, const double* e, int homo, int rpamin,
                      int rpamax, double* eps_out, int ld) {
  MOCK_BEGIN(ctx)
  require_mmn(ctx);
  REQUIRE(rpamin == ctx->mmin && rpamin == ctx->nmin && rpamax - rpamin + 1 == ctx->ntotal, "RPA range must match Mmn");
  const int naux = ctx->naux, n_occ = homo - rpamin + 1, n_unocc = rpamax - homo, ntot = ctx->ntotal;
  ctx->eps.assign((size_t)naux * naux, 0.0);
  std::vector<double> d(n_unocc), Hv((size_t)n_unocc * naux);
  for (int v = 0; v < n_occ; ++v) {
    for (int c = 0; c < n_unocc; ++c) {
      const double dE = e[ntot - n_unocc + c] - e[v];
      if (kind == 0) {
        d[c] = 4.0 * dE / (dE * dE + fre * fre);
      } else if (kind == 1) {
        const double dm = dE - fre, dp = dE + fre;
        d[c] = 2.0 * (dm / (dm * dm + eta * eta) + dp / (dp * dp + eta * eta));
      } else {
        const double dm = fre - dE, dp = fre + dE;
        d[c] = -2.0 * (dm / (dm * dm + (fim + eta) * (fim + eta)) - dp / (dp * dp + (fim - eta) * (fim - eta)));
      }
    }
    for (int c = 0; c < n_unocc; ++c)
      for (int a = 0; a < naux; ++a) Hv[(size_t)c * naux + a] = ctx->hole(v, ntot - n_unocc + c, a);
    for (int c2 = 0; c2 < naux; ++c2)
      for (int c1 = c2; c1 < naux; ++c1) {  // lower triangle, mirrored below
        double s = 0.0;
        for (int c = 0; c < n_unocc; ++c) s += Hv[(size_t)c * naux + c1] * d[c] * Hv[(size_t)c * naux + c2];
        ctx->eps[c1 + (size_t)c2 * naux] += s;
      }
  }
  for (int c2 = 0; c2 < naux; ++c2)
    for (int c1 = c2 + 1; c1 < naux; ++c1) ctx->eps[c2 + (size_t)c1 * naux] = ctx->eps[c1 + (size_t)c2 * naux];
  for (int i = 0; i < naux; ++i) ctx->eps[i + (size_t)i * naux] += 1.0;
  if (eps_out) {
    REQUIRE(ld >= naux, "leading dimension too small");
    for (int j = 0; j < naux; ++j) std::memcpy(eps_out + (size_t)j * ld, &ctx->eps[(size_t)j * naux], sizeof(double) * naux);
  }
  MOCK_END(ctx)
}
double* gwbse_rpa_epsilon_ptr(gwbse_ctx* ctx) { return ctx ? ctx->eps.data() : nullptr; }
int gwbse_rpa_set_qsgw_rotation(gwbse_ctx* ctx, const double* U, int ldu, int qptotal, int qpmin, int homo) {
  MOCK_BEGIN(ctx)
  ctx->qsgw_U.clear();
  if (U) {
    REQUIRE(qptotal > 0 && ldu >= qptotal, "invalid QSGW rotation");
    ctx->qsgw_U.resize((size_t)qptotal * qptotal);
    for (int j = 0; j < qptotal; ++j)
      for (int i = 0; i < qptotal; ++i) ctx->qsgw_U[i + (size_t)j * qptotal] = U[i + (size_t)j * ldu];
    ctx->qsgw_q = qptotal;
    ctx->qsgw_qpmin = qpmin;
    ctx->qsgw_homo = homo;
  }
  MOCK_END(ctx)
}
int gwbse_rpa_h2p_apb(gwbse_ctx* ctx, const double* e, int homo, int rpamin, int rpamax, double* apb, int ld) {
  MOCK_BEGIN(ctx)
  // RPA::Calculate_H2p_ApB, rpa.cc:281-326 (+ the A-B diagonal)
  require_mmn(ctx);
  const int n_occ = homo - rpamin + 1, n_unocc = rpamax - homo, S = n_occ * n_unocc;
  REQUIRE(ld >= S, "leading dimension too small");
  for (int v2 = 0; v2 < n_occ; ++v2)
    for (int c2 = 0; c2 < n_unocc; ++c2)
      for (int v1 = 0; v1 < n_occ; ++v1)
        for (int c1 = 0; c1 < n_unocc; ++c1) {
          double s = 0.0;
          for (int chi = 0; chi < ctx->naux; ++chi) s += ctx->hole(v1, n_occ + c1, chi) * ctx->hole(v2, n_occ + c2, chi);
          double h = 4.0 * s;
          if (v1 == v2 && c1 == c2) h += e[n_occ + c1] - e[v1];
          apb[(v1 * n_unocc + c1) + (size_t)(v2 * n_unocc + c2) * ld] = h;
        }
  MOCK_END(ctx)
}

// ---- Sigma -------------------------------------------------------------------------------------------------------------
int gwbse_sigma_x(gwbse_ctx* ctx, int homo, int rpamin, int qpmin, int qpmax, double* out, int ld) {
  MOCK_BEGIN(ctx)
  require_mmn(ctx);
  const int occ = homo - rpamin + 1, off = qpmin - rpamin, q = qpmax - qpmin + 1;
  REQUIRE(ld >= q && off >= 0 && off + q <= ctx->mtotal, "QP window outside Mmn");
  for (int i = 0; i < q; ++i)
    for (int j = 0; j < q; ++j) {
      double s = 0.0;
      for (int chi = 0; chi < ctx->naux; ++chi)
        for (int n = 0; n < occ; ++n) s += ctx->M(off + i, n, chi) * ctx->M(off + j, n, chi);
      out[i + (size_t)j * ld] = -s;
    }
  MOCK_END(ctx)
}
int gwbse_sigma_ppm_set(gwbse_ctx* ctx, const double* w, const double* f, const double* e, int homo, int rpamin,
                        int qpmin, double eta) {
  MOCK_BEGIN(ctx)
  require_mmn(ctx);
  ctx->fac.resize(ctx->naux);
  for (int i = 0; i < ctx->naux; ++i) ctx->fac[i] = w[i] < 1e-9 ? 0.0 : w[i] * f[i];  // sigma_ppm.cc:47-52
  ctx->pole.assign(f, f + ctx->naux);
  ctx->energies.assign(e, e + ctx->ntotal);
  ctx->lumo = homo + 1;  // sigma_ppm.cc:40,56-57: unshifted
  ctx->qpoff = qpmin - rpamin;
  ctx->q = ctx->mtotal - ctx->qpoff;
  ctx->eta = eta;
  ctx->ppm_ready = true;
  MOCK_END(ctx)
}
int gwbse_sigma_update_energies(gwbse_ctx* ctx, int which, const double* e) {
  MOCK_BEGIN(ctx)
  (which == 0 ? ctx->energies : ctx->energies_exact).assign(e, e + ctx->ntotal);
  MOCK_END(ctx)
}
int gwbse_rpa_h2p_block(gwbse_ctx* ctx, gwbse_ctx* other, int homo, int homo_other, int rpamin, int rpamax, double alpha,
                        double* block, int ld) {
  MOCK_BEGIN(ctx)
  // RPA_UKS::Calculate_H2p_ApB, rpa_uks.cc:475-540 (one spin block, without the A-B diagonal)
  require_mmn(ctx);
  REQUIRE(other && other->naux == ctx->naux, "both channels live on one GPU with one aux basis");
  const int n_occ = homo - rpamin + 1, n_unocc = rpamax - homo, no2 = homo_other - rpamin + 1, nu2 = rpamax - homo_other;
  REQUIRE(n_occ > 0 && n_unocc > 0 && no2 > 0 && nu2 > 0, "empty particle-hole space");
  REQUIRE(ld >= n_occ * n_unocc, "leading dimension too small");
  for (int v2 = 0; v2 < no2; ++v2)
    for (int c2 = 0; c2 < nu2; ++c2)
      for (int v1 = 0; v1 < n_occ; ++v1)
        for (int c1 = 0; c1 < n_unocc; ++c1) {
          double s = 0.0;
          for (int chi = 0; chi < ctx->naux; ++chi) s += ctx->hole(v1, n_occ + c1, chi) * other->hole(v2, no2 + c2, chi);
          block[(v1 * n_unocc + c1) + (size_t)(v2 * nu2 + c2) * ld] = alpha * s;
        }
  MOCK_END(ctx)
}
int gwbse_sigma_ppm_eval(gwbse_ctx* ctx, int nreq, const int* levels, const double* freqs, double* sigma, double* dsigma) {
  MOCK_BEGIN(ctx)
  for (int i = 0; i < nreq; ++i) ppm_eval(ctx, levels[i], freqs[i], sigma + i, dsigma ? dsigma + i : nullptr);
  MOCK_END(ctx)
}
int gwbse_sigma_eval_groups(gwbse_ctx* ctx, int which, int ngroups, const int* levels, const int* gptr, const double* freqs,
                            double* sigma, double* dsigma) {
  MOCK_BEGIN(ctx)
  for (int g = 0; g < ngroups; ++g)
    for (int i = gptr[g]; i < gptr[g + 1]; ++i)
      (which == 0 ? ppm_eval : exact_eval)(ctx, levels[g], freqs[i], sigma + i, dsigma ? dsigma + i : nullptr);
  MOCK_END(ctx)
}
int gwbse_sigma_ppm_offdiag(gwbse_ctx* ctx, int q, const double* freqs, double* out, int ld) {
  MOCK_BEGIN(ctx)
  REQUIRE(ctx->ppm_ready, "sigma evaluator not prepared");
  REQUIRE(q > 0 && ctx->qpoff + q <= ctx->mtotal && ld >= q, "q does not match the prepared evaluator");
  const double eta2 = ctx->eta * ctx->eta;
  for (int i = 0; i < q; ++i)
    for (int j = 0; j < q; ++j) {
      double s = 0.0;
      if (i != j)
        for (int chi = 0; chi < ctx->naux; ++chi) {
          if (ctx->fac[chi] == 0.0) continue;
          for (int n = 0; n < ctx->ntotal; ++n) {
            const double sh = n < ctx->lumo ? ctx->pole[chi] : -ctx->pole[chi];
            const double t1 = freqs[i] - ctx->energies[n] + sh, t2 = freqs[j] - ctx->energies[n] + sh;
            s += 0.25 * ctx->fac[chi] * (t1 / (t1 * t1 + eta2) + t2 / (t2 * t2 + eta2)) * ctx->M(ctx->qpoff + i, n, chi) *
                 ctx->M(ctx->qpoff + j, n, chi);
          }
        }
      out[i + (size_t)j * ld] = s;
    }
  MOCK_END(ctx)
}
int gwbse_sigma_exact_prepare(gwbse_ctx* ctx, const double* omegas, const double* XpY, int ldxpy, const double* e,
                              int homo, int rpamin, int rpamax, int qpmin, int qpmax, double eta) {
  MOCK_BEGIN(ctx)
  // Sigma_Exact::PrepareScreening / CalcResidues, sigma_exact.cc:29-38, 109-148
  require_mmn(ctx);
  REQUIRE(rpamin == ctx->mmin && rpamin == ctx->nmin && rpamax == ctx->nmax, "RPA range must match Mmn");
  const int n_occ = homo + 1 - rpamin, n_unocc = rpamax - homo, S = n_occ * n_unocc;
  const int q = qpmax - qpmin + 1, qpoff = qpmin - rpamin, ntot = ctx->ntotal;
  REQUIRE(ldxpy >= S && q > 0 && qpoff >= 0 && qpoff + q <= ctx->mtotal, "invalid sizes");
  ctx->residues.assign((size_t)q * ntot * S, 0.0);
  std::vector<double> fc((size_t)S);
  for (int i = 0; i < q; ++i)
    for (int n = 0; n < ntot; ++n) {
      for (int v = 0; v < n_occ; ++v)
        for (int c = 0; c < n_unocc; ++c) {
          double t = 0.0;
          for (int chi = 0; chi < ctx->naux; ++chi) t += ctx->hole(v, n_occ + c, chi) * ctx->M(qpoff + i, n, chi);
          fc[(size_t)v * n_unocc + c] = t;
        }
      for (int s2 = 0; s2 < S; ++s2) {
        double r = 0.0;
        for (int vc = 0; vc < S; ++vc) r += fc[vc] * XpY[vc + (size_t)s2 * ldxpy];
        ctx->residues[((size_t)i * ntot + n) * S + s2] = r;
      }
    }
  ctx->rpa_omegas.assign(omegas, omegas + S);
  ctx->energies_exact.assign(e, e + ntot);
  ctx->ex_S = S;
  ctx->ex_q = q;
  ctx->ex_nocc = n_occ;
  ctx->ex_eta = eta;
  ctx->ex_diag_pref = 2.0;
  ctx->ex_offdiag_pref = 1.0;
  ctx->exact_ready = true;
  MOCK_END(ctx)
}
// halves of the call above for the unrestricted evaluator (sigma_exact_uks.cc:37-61, rpa_uks.cc:91-161)
int gwbse_sigma_exact_project(gwbse_ctx* ctx, const double* XpY, int ldxpy, int ncols, int homo, int rpamin, int rpamax,
                              int accumulate, double* Z, int ldz) {
  MOCK_BEGIN(ctx)
  require_mmn(ctx);
  REQUIRE(rpamin == ctx->mmin && rpamin == ctx->nmin && rpamax == ctx->nmax, "RPA range must match Mmn");
  const int n_occ = homo + 1 - rpamin, n_unocc = rpamax - homo, S = n_occ * n_unocc;
  REQUIRE(ldxpy >= S && ldz >= ctx->naux && ncols >= 0, "invalid sizes");
  for (int s2 = 0; s2 < ncols; ++s2)
    for (int chi = 0; chi < ctx->naux; ++chi) {
      double t = 0.0;
      for (int v = 0; v < n_occ; ++v)
        for (int c = 0; c < n_unocc; ++c)
          t += ctx->hole(v, n_occ + c, chi) * XpY[(size_t)v * n_unocc + c + (size_t)s2 * ldxpy];
      double& z = Z[chi + (size_t)s2 * ldz];
      z = accumulate ? z + t : t;
    }
  MOCK_END(ctx)
}
int gwbse_sigma_exact_prepare_modes(gwbse_ctx* ctx, const double* omegas, int nmodes, const double* Z, int ldz,
                                    const double* e, int homo, int rpamin, int rpamax, int qpmin, int qpmax, double eta,
                                    double diag_pref, double offdiag_pref) {
  MOCK_BEGIN(ctx)
  require_mmn(ctx);
  REQUIRE(rpamin == ctx->mmin && rpamin == ctx->nmin && rpamax == ctx->nmax, "RPA range must match Mmn");
  const int q = qpmax - qpmin + 1, qpoff = qpmin - rpamin, ntot = ctx->ntotal;
  REQUIRE(nmodes > 0 && ldz >= ctx->naux && q > 0 && qpoff >= 0 && qpoff + q <= ctx->mtotal, "invalid sizes");
  ctx->residues.assign((size_t)q * ntot * nmodes, 0.0);
  for (int i = 0; i < q; ++i)
    for (int n = 0; n < ntot; ++n)
      for (int s2 = 0; s2 < nmodes; ++s2) {
        double r = 0.0;
        for (int chi = 0; chi < ctx->naux; ++chi) r += ctx->M(qpoff + i, n, chi) * Z[chi + (size_t)s2 * ldz];
        ctx->residues[((size_t)i * ntot + n) * nmodes + s2] = r;
      }
  ctx->rpa_omegas.assign(omegas, omegas + nmodes);
  ctx->energies_exact.assign(e, e + ntot);
  ctx->ex_S = nmodes;
  ctx->ex_q = q;
  ctx->ex_nocc = homo + 1 - rpamin;
  ctx->ex_eta = eta;
  ctx->ex_diag_pref = diag_pref;
  ctx->ex_offdiag_pref = offdiag_pref;
  ctx->exact_ready = true;
  MOCK_END(ctx)
}
int gwbse_sigma_exact_eval(gwbse_ctx* ctx, int nreq, const int* levels, const double* freqs, double* sigma, double* dsigma) {
  MOCK_BEGIN(ctx)
  for (int i = 0; i < nreq; ++i) exact_eval(ctx, levels[i], freqs[i], sigma + i, dsigma ? dsigma + i : nullptr);
  MOCK_END(ctx)
}
int gwbse_sigma_exact_offdiag(gwbse_ctx* ctx, int q, const double* freqs, double* out, int ld) {
  MOCK_BEGIN(ctx)
  REQUIRE(ctx->exact_ready, "sigma evaluator not prepared");
  REQUIRE(q == ctx->ex_q && ld >= q, "q does not match the prepared evaluator");
  const double eta2 = ctx->ex_eta * ctx->ex_eta;
  const int S = ctx->ex_S, ntot = ctx->ntotal;
  for (int i = 0; i < q; ++i)
    for (int j = 0; j < q; ++j) {
      double s = 0.0;
      if (i != j)
        for (int n = 0; n < ntot; ++n)
          for (int p = 0; p < S; ++p) {
            const double sh = n < ctx->ex_nocc ? ctx->rpa_omegas[p] : -ctx->rpa_omegas[p];
            const double t1 = freqs[i] - ctx->energies_exact[n] + sh, t2 = freqs[j] - ctx->energies_exact[n] + sh;
            s += ctx->ex_offdiag_pref * ctx->residues[((size_t)i * ntot + n) * S + p] *
                 ctx->residues[((size_t)j * ntot + n) * S + p] * (t1 / (t1 * t1 + eta2) + t2 / (t2 * t2 + eta2));
          }
      out[i + (size_t)j * ld] = s;
    }
  MOCK_END(ctx)
}

// ---- Sigma_CDA (sigma_cda.cc:30-141, ImaginaryAxisIntegration.cc:90-176), evaluated the reference's way: one
//      I kappa_j product per node and evaluation ------------------------------------------------------------------
// eps of the channel, or the spin-summed one of RPA_UKS when a partner channel is registered (sigma_cda_uks.cc)
static void cda_epsilon(gwbse_ctx* ctx, int kind, double fre, double fim, double eta, const double* e, int homo, int rpamin,
                        int rpamax) {
  if (gwbse_rpa_epsilon(ctx, kind, fre, fim, eta, e, homo, rpamin, rpamax, nullptr, 0)) throw std::runtime_error(ctx->err);
  gwbse_ctx* o = ctx->cda_partner;
  if (!o) return;
  if (gwbse_rpa_epsilon(o, kind, fre, fim, eta, ctx->cda_partner_e.data(), ctx->cda_partner_homo, rpamin, rpamax, nullptr, 0))
    throw std::runtime_error(o->err);
  for (size_t i = 0; i < ctx->eps.size(); ++i) ctx->eps[i] = 0.5 * (ctx->eps[i] + o->eps[i]);
}
int gwbse_sigma_cda_set_partner(gwbse_ctx* ctx, gwbse_ctx* other, int homo_other, const double* energies_other) {
  MOCK_BEGIN(ctx)
  ctx->cda_partner = other;
  ctx->cda_partner_homo = homo_other;
  ctx->cda_partner_e.clear();
  if (other) {
    REQUIRE(other != ctx && energies_other, "the partner channel needs its own filled Mmn");
    ctx->cda_partner_e.assign(energies_other, energies_other + other->ntotal);
  }
  MOCK_END(ctx)
}
int gwbse_sigma_cda_prepare(gwbse_ctx* ctx, int order, const double* points, const double* weights, int symmetry,
                            double alpha, const double* energies, int homo, int rpamin, int rpamax, int qpmin,
                            int qpmax, double eta) {
  MOCK_BEGIN(ctx)
  require_mmn(ctx);
  const int n = ctx->naux;
  const size_t nn = (size_t)n * n;
  ctx->cda_ready = false;
  ctx->cda_kappa.assign(nn * (order + 1), 0.0);
  double* kzero = ctx->cda_kappa.data() + nn * order;
  auto inverse_minus_one = [&](double* dst) {
    std::vector<double> A(ctx->eps), I(nn, 0.0);
    for (int i = 0; i < n; ++i) I[i + (size_t)i * n] = 1.0;
    solve(n, n, A.data(), n, I.data(), n);
    for (int i = 0; i < n; ++i) I[i + (size_t)i * n] -= 1.0;
    std::copy(I.begin(), I.end(), dst);
  };
  cda_epsilon(ctx, 2, 0.0, 0.0, eta, energies, homo, rpamin, rpamax);
  inverse_minus_one(kzero);
  for (int j = 0; j < order; ++j) {
    cda_epsilon(ctx, 0, points[j], 0.0, eta, energies, homo, rpamin, rpamax);
    double* k = ctx->cda_kappa.data() + nn * j;
    inverse_minus_one(k);
    const double sc = std::exp(-std::pow(alpha * points[j], 2));
    for (size_t i = 0; i < nn; ++i) k[i] = -k[i] + sc * kzero[i];
  }
  ctx->cda_pts.assign(points, points + order);
  ctx->cda_wts.assign(weights, weights + order);
  ctx->cda_order = order;
  ctx->cda_sym = symmetry;
  ctx->cda_alpha = alpha;
  ctx->cda_eta = eta;
  ctx->cda_homo = homo;
  ctx->cda_rpamin = rpamin;
  ctx->cda_rpamax = rpamax;
  ctx->cda_qpmin = qpmin;
  ctx->cda_q = qpmax - qpmin + 1;
  ctx->cda_ready = true;
  MOCK_END(ctx)
}

int gwbse_sigma_cda_eval(gwbse_ctx* ctx, int nreq, const int* levels, const double* freqs, const double* energies,
                         double* sigma) {
  MOCK_BEGIN(ctx)
  REQUIRE(ctx->cda_ready, "CDA screening not prepared (gwbse_sigma_cda_prepare)");
  const int n = ctx->naux, nt = ctx->ntotal, order = ctx->cda_order;
  const size_t nn = (size_t)n * n;
  const int occ = ctx->cda_homo + 1 - ctx->cda_rpamin;
  const int homo = ctx->cda_homo - ctx->cda_rpamin;
  const double fermi = 0.5 * (energies[homo + 1] + energies[homo]);
  const double pi = 3.14159265358979323846;
  for (int r = 0; r < nreq; ++r) {
    REQUIRE(levels[r] >= 0 && levels[r] < ctx->cda_q, "level outside the qp window");
    const int m = levels[r] + ctx->cda_qpmin - ctx->cda_rpamin;
    const double w = freqs[r];
    auto rowform = [&](const double* kappa, int i) {  // (I kappa)[i,:] . I[i,:]
      double s = 0.0;
      for (int a = 0; a < n; ++a) {
        double t = 0.0;
        for (int b = 0; b < n; ++b) t += ctx->M(m, i, b) * kappa[b + (size_t)a * n];
        s += t * ctx->M(m, i, a);
      }
      return s;
    };
    double total = 0.0;
    for (int j = 0; j < order; ++j) {
      const double* kappa = ctx->cda_kappa.data() + nn * j;
      double acc = 0.0;
      for (int i = 0; i < nt; ++i) {
        const std::complex<double> dE(w - energies[i], i < occ ? ctx->cda_eta : -ctx->cda_eta), cp(0.0, ctx->cda_pts[j]);
        std::complex<double> den = 1.0 / (dE + cp);
        if (ctx->cda_sym) den += 1.0 / (dE - cp);
        acc += den.real() * rowform(kappa, i);
      }
      total += ctx->cda_wts[j] * 0.5 / pi * acc;
    }
    const double* kzero = ctx->cda_kappa.data() + nn * order;
    for (int i = 0; i < nt; ++i) {
      const double delta = energies[i] - w, ad = std::fabs(delta);
      double factor = 0.0;
      if (fermi < energies[i] && energies[i] < w) factor = 1.0;
      else if (fermi > energies[i] && energies[i] > w) factor = -1.0;
      else if (ad < 1e-10 && fermi > energies[i]) factor = -0.5;
      else if (ad < 1e-10 && fermi < energies[i]) factor = 0.5;
      if (std::fabs(factor) > 1e-10) {
        cda_epsilon(ctx, 2, ad, ctx->cda_eta, ctx->cda_eta, energies, ctx->cda_homo, ctx->cda_rpamin, ctx->cda_rpamax);
        std::vector<double> A(ctx->eps), x(n), row(n);
        for (int a = 0; a < n; ++a) x[a] = row[a] = ctx->M(m, i, a);
        solve(n, 1, A.data(), n, x.data(), n);
        double dot = 0.0;
        for (int a = 0; a < n; ++a) dot += (x[a] - row[a]) * row[a];
        total += factor * dot;
      }
      if (ad > 1e-10)
        total += 0.5 * std::copysign(1.0, delta) * std::exp(std::pow(ctx->cda_alpha * delta, 2)) *
                 std::erfc(std::fabs(ctx->cda_alpha * delta)) * rowform(kzero, i);
    }
    sigma[r] = total;
  }
  MOCK_END(ctx)
}

// ---- BSE ---------------------------------------------------------------------------------------------------------------
int gwbse_bse_configure(gwbse_ctx* ctx, int homo, int rpamin, int vmin, int cmax, const double* eps_inv, const double* Hqp,
                        int ldh) {
  MOCK_BEGIN(ctx)
  require_mmn(ctx);
  REQUIRE(rpamin == ctx->mmin && rpamin == ctx->nmin, "RPA range must match Mmn");
  ctx->homo = homo;
  ctx->vt = homo - vmin + 1;
  ctx->ct = cmax - homo;
  ctx->voff = vmin - rpamin;
  ctx->bse_vmin = vmin;
  ctx->bse_cmax = cmax;
  ctx->bse_rpamin = rpamin;
  ctx->coff = homo + 1 - rpamin;
  REQUIRE(ctx->vt > 0 && ctx->ct > 0 && ctx->voff >= 0, "invalid BSE level ranges");
  REQUIRE(ctx->coff + ctx->ct <= ctx->mtotal && ctx->coff + ctx->ct <= ctx->ntotal, "BSE range exceeds Mmn");
  const int hs = ctx->vt + ctx->ct;
  REQUIRE(ldh >= hs, "Hqp leading dimension too small");
  ctx->eps_inv.assign(eps_inv, eps_inv + ctx->naux);
  ctx->hqp.resize((size_t)hs * hs);
  for (int j = 0; j < hs; ++j) std::memcpy(&ctx->hqp[(size_t)j * hs], Hqp + (size_t)j * ldh, sizeof(double) * hs);
  ctx->bse_ready = true;
  MOCK_END(ctx)
}
int gwbse_bse_matmul(gwbse_ctx* ctx, int cqp, int cx, int cd, int cd2, int k, const double* X, int ldx, double* Y, int ldy) {
  MOCK_BEGIN(ctx)
  bse_matmul(ctx, cqp, cx, cd, cd2, k, X, ldx, Y, ldy);
  MOCK_END(ctx)
}
int gwbse_bse_matmul_dev(gwbse_ctx* ctx, int cqp, int cx, int cd, int cd2, int k, const double* X, int ldx, double* Y,
                         int ldy) {
  return gwbse_bse_matmul(ctx, cqp, cx, cd, cd2, k, X, ldx, Y, ldy);
}
int gwbse_bse_vc_project_dev(gwbse_ctx* ctx, int k, const double* X, int ldx, double* W) {
  MOCK_BEGIN(ctx)
  REQUIRE(ctx->bse_ready, "BSE operator not configured (gwbse_bse_configure)");
  const int vt = ctx->vt, ct = ctx->ct, vo = ctx->voff, co = ctx->coff, naux = ctx->naux;
  REQUIRE(ldx >= vt * ct, "Shape mismatch in BSE projection");
  for (int j = 0; j < k; ++j)
    for (int chi = 0; chi < naux; ++chi) {
      double s = 0.0;
      for (int v = 0; v < vt; ++v)
        for (int c = 0; c < ct; ++c) s += ctx->M(vo + v, co + c, chi) * X[(size_t)j * ldx + ct * v + c];
      W[(size_t)j * naux + chi] = s;
    }
  MOCK_END(ctx)
}
int gwbse_bse_vc_expand_dev(gwbse_ctx* ctx, double alpha, int screened, int k, const double* W, double* Y, int ldy) {
  MOCK_BEGIN(ctx)
  REQUIRE(ctx->bse_ready, "BSE operator not configured (gwbse_bse_configure)");
  const int vt = ctx->vt, ct = ctx->ct, vo = ctx->voff, co = ctx->coff, naux = ctx->naux;
  REQUIRE(ldy >= vt * ct, "Shape mismatch in BSE expansion");
  for (int j = 0; j < k; ++j)
    for (int v = 0; v < vt; ++v)
      for (int c = 0; c < ct; ++c) {
        double s = 0.0;
        for (int chi = 0; chi < naux; ++chi)
          s += ctx->M(vo + v, co + c, chi) * (screened ? ctx->eps_inv[chi] : 1.0) * W[(size_t)j * naux + chi];
        Y[(size_t)j * ldy + ct * v + c] += alpha * s;
      }
  MOCK_END(ctx)
}
int gwbse_bse_hd2_cross_dev(gwbse_ctx* ctx, gwbse_ctx* other, int homo_other, double alpha, int k, const double* X,
                            int ldx, double* Y, int ldy) {
  MOCK_BEGIN(ctx)
  REQUIRE(ctx->bse_ready && other, "BSE operator not configured (gwbse_bse_configure)");
  const int vt = ctx->vt, ct = ctx->ct, vo = ctx->voff, co = ctx->coff, naux = ctx->naux;
  const int vti = homo_other - ctx->bse_vmin + 1, cti = ctx->bse_cmax - homo_other, coi = homo_other + 1 - ctx->bse_rpamin;
  REQUIRE(ldx >= vti * cti && ldy >= vt * ct, "Shape mismatch in the cross-spin BSE block");
  for (int j = 0; j < k; ++j)
    for (int v1 = 0; v1 < vt; ++v1)
      for (int c1 = 0; c1 < ct; ++c1) {
        double s = 0.0;
        for (int v2 = 0; v2 < vti; ++v2)
          for (int c2 = 0; c2 < cti; ++c2) {
            double b = 0.0;
            for (int chi = 0; chi < naux; ++chi)
              b += ctx->M(co + c1, vo + v2, chi) * ctx->eps_inv[chi] * other->M(vo + v1, coi + c2, chi);
            s += b * X[(size_t)j * ldx + cti * v2 + c2];
          }
        Y[(size_t)j * ldy + ct * v1 + c1] += alpha * s;
      }
  MOCK_END(ctx)
}
int gwbse_bse_stats(gwbse_ctx* ctx, double* fl, long long* p, long long* c, int reset) {
  if (!ctx) return 1;
  if (fl) *fl = ctx->bse_flops;
  if (p) *p = ctx->bse_products;
  if (c) *c = ctx->bse_columns;
  if (reset) ctx->bse_products = ctx->bse_columns = 0;
  return 0;
}
int gwbse_bse_dense_stats(gwbse_ctx* ctx, long long* b, long long* c, double* r) {
  if (!ctx) return 1;
  if (b) *b = 0;
  if (c) *c = 0;
  if (r) *r = 0.0;
  return 0;
}
int gwbse_bse_diagonal(gwbse_ctx* ctx, int cqp, int cx, int cd, int cd2, double* diag) {
  MOCK_BEGIN(ctx)
  REQUIRE(ctx->bse_ready, "BSE operator not configured (gwbse_bse_configure)");
  REQUIRE(!(cd != 0 && cd2 != 0), "Hamiltonian cannot contain Hd and Hd2 at the same time");
  const int vt = ctx->vt, ct = ctx->ct, vo = ctx->voff, co = ctx->coff, hs = vt + ct;
  for (int v = 0; v < vt; ++v)
    for (int c = 0; c < ct; ++c) {
      double e = 0.0;
      for (int chi = 0; chi < ctx->naux; ++chi) {
        if (cx) e += cx * ctx->M(vo + v, co + c, chi) * ctx->M(vo + v, co + c, chi);
        if (cd) e -= cd * ctx->M(co + c, co + c, chi) * ctx->eps_inv[chi] * ctx->M(vo + v, vo + v, chi);
        if (cd2) e -= cd2 * ctx->M(co + c, vo + v, chi) * ctx->eps_inv[chi] * ctx->M(vo + v, co + c, chi);
      }
      if (cqp) e += cqp * (ctx->hqp[(vt + c) + (size_t)(vt + c) * hs] - ctx->hqp[v + (size_t)v * hs]);
      diag[ct * v + c] = e;
    }
  MOCK_END(ctx)
}

// ---- Davidson helpers (davidsonsolver.cc:392-478) ---------------------------------------------------------------------------
int gwbse_gramschmidt_dev(gwbse_ctx* ctx, int rows, int ncols, int nstart, double* Q, int ldq) {
  MOCK_BEGIN(ctx)
  REQUIRE(ldq >= rows && nstart >= 0 && nstart <= ncols, "invalid Gram-Schmidt arguments");
  auto col = [&](int j) { return Q + (size_t)j * ldq; };
  auto dot = [&](const double* a, const double* b) {
    double s = 0.0;
    for (int i = 0; i < rows; ++i) s += a[i] * b[i];
    return s;
  };
  std::vector<double> norms0(std::max(ncols - nstart, 0));
  for (int j = nstart; j < ncols; ++j) norms0[j - nstart] = std::sqrt(dot(col(j), col(j)));
  for (int rep = 0; rep < 2; ++rep) {
    if (nstart > 0)
      for (int j = nstart; j < ncols; ++j) {
        std::vector<double> coef(nstart);
        for (int o = 0; o < nstart; ++o) coef[o] = dot(col(o), col(j));
        for (int o = 0; o < nstart; ++o)
          for (int i = 0; i < rows; ++i) col(j)[i] -= coef[o] * col(o)[i];
        const double nr = std::sqrt(dot(col(j), col(j)));
        for (int i = 0; i < rows; ++i) col(j)[i] /= nr;
      }
    for (int j = nstart + 1; j < ncols; ++j) {
      std::vector<double> coef(j - nstart);
      for (int o = nstart; o < j; ++o) coef[o - nstart] = dot(col(o), col(j));
      for (int o = nstart; o < j; ++o)
        for (int i = 0; i < rows; ++i) col(j)[i] -= coef[o - nstart] * col(o)[i];
      const double nj = std::sqrt(dot(col(j), col(j)));
      if (rep == 1 && nj <= 1e-12 * norms0[j - nstart]) throw std::runtime_error("Linear dependencies in Gram-Schmidt.");
      for (int i = 0; i < rows; ++i) col(j)[i] /= nj;
    }
  }
  MOCK_END(ctx)
}
int gwbse_davidson_correction_dev(gwbse_ctx* ctx, int rows, int ncols, int olsen, const double* diag, const double* lambda,
                                  const double* R, int ldr, const double* Q, int ldq, double* W, int ldw) {
  MOCK_BEGIN(ctx)
  auto fin = [](double v) { return std::isfinite(v) ? v : 0.0; };
  for (int j = 0; j < ncols; ++j) {
    double* w = W + (size_t)j * ldw;
    const double* r = R + (size_t)j * ldr;
    for (int i = 0; i < rows; ++i) w[i] = fin(-r[i] / (diag[i] - lambda[j]));
    if (olsen) {
      const double* x = Q + (size_t)j * ldq;
      double num = 0.0, den = 0.0;
      for (int i = 0; i < rows; ++i) {
        num += x[i] * w[i];
        den += x[i] * fin(-x[i] / (diag[i] - lambda[j]));
      }
      for (int i = 0; i < rows; ++i) w[i] = fin(w[i] + num / den * x[i]);
    }
    double nr = 0.0;
    for (int i = 0; i < rows; ++i) nr += w[i] * w[i];
    nr = 1.0 / std::sqrt(nr);
    for (int i = 0; i < rows; ++i) w[i] *= nr;
  }
  MOCK_END(ctx)
}

// ---- AO integrals: the shared host/device source, run serially ---------------------------------------------------------------
int gwbse_basis_normalize(int nshell, const int* l, const int* nprim, const double* exps, const double* contractions,
                          double* out) {
  size_t p0 = 0;
  for (int s = 0; s < nshell; ++s) {
    if (l[s] < 0 || l[s] > ao::LMAX_SHELL || nprim[s] < 1) return 1;
    ao::normalize_contraction(l[s], nprim[s], exps + p0, contractions + p0, out + p0);
    p0 += (size_t)nprim[s];
  }
  return 0;
}
int gwbse_basis_create(gwbse_ctx* ctx, int nshell, const int* l, const int* nprim, const double* centers, const double* exps,
                       const double* coefs, gwbse_basis** out) {
  MOCK_BEGIN(ctx)
  REQUIRE(out && l && nprim && centers && exps && coefs && nshell > 0, "invalid basis description");
  gwbse_basis* b = new gwbse_basis;
  try {
    b->host.build(nshell, l, nprim, centers, exps, coefs);
  } catch (...) {
    delete b;
    throw;
  }
  b->pairs = ao::make_pair_lists(b->host, false);
  b->unit_pairs = ao::make_pair_lists(b->host, true);
  *out = b;
  MOCK_END(ctx)
}
int gwbse_basis_destroy(gwbse_ctx* ctx, gwbse_basis* b) {
  delete b;
  return ctx ? 0 : 1;
}
int gwbse_basis_size(const gwbse_basis* b) { return b ? b->host.nfunc : -1; }
int gwbse_ao3c_block(gwbse_ctx* ctx, const gwbse_basis* aux, const gwbse_basis* dft, int off, int cnt, double* out) {
  MOCK_BEGIN(ctx)
  ao3c_block(aux, dft, off, cnt, out, 0);
  MOCK_END(ctx)
}
int gwbse_ao3c_block_dev(gwbse_ctx* ctx, const gwbse_basis* aux, const gwbse_basis* dft, int off, int cnt, double* out) {
  return gwbse_ao3c_block(ctx, aux, dft, off, cnt, out);
}
int gwbse_ao_coulomb2c(gwbse_ctx* ctx, const gwbse_basis* aux, double* V, int ld) {
  MOCK_BEGIN(ctx)
  REQUIRE(aux && V && ld >= aux->host.nfunc, "invalid argument");
  const int n = aux->host.nfunc;
  std::vector<double> d((size_t)n * n, 0.0);
  ao::OutSpec spec{d.data(), (long long)n, 1, 0, 0, n, 0};
  integrals(*aux, aux->unit_pairs, *aux, 0, aux->host.nshell, spec, aux->host.lmax);
  for (int j = 0; j < n; ++j) std::memcpy(V + (size_t)j * ld, &d[(size_t)j * n], sizeof(double) * n);
  MOCK_END(ctx)
}
static int one_electron(gwbse_ctx* ctx, const gwbse_basis* b, int code, double* out, int ld) {
  MOCK_BEGIN(ctx)
  REQUIRE(b && out && ld >= b->host.nfunc, "invalid argument");
  const int n = b->host.nfunc;
  std::vector<double> d((size_t)n * n, 0.0);
  ao::OutSpec spec{d.data(), 0, 1, (long long)n, 0, 1, 1};
  integrals(*b, b->pairs, *b, code, code + 1, spec, 0);
  for (int j = 0; j < n; ++j) std::memcpy(out + (size_t)j * ld, &d[(size_t)j * n], sizeof(double) * n);
  MOCK_END(ctx)
}
int gwbse_ao_overlap(gwbse_ctx* ctx, const gwbse_basis* b, double* S, int ld) { return one_electron(ctx, b, -1, S, ld); }
int gwbse_ao_dipole(gwbse_ctx* ctx, const gwbse_basis* b, double* D, int ld) {
  for (int k = 0; k < 3; ++k)
    if (one_electron(ctx, b, -2 - k, D + (size_t)k * ld * (b ? b->host.nfunc : 0), ld)) return 1;
  return 0;
}
int gwbse_mmn_fill_from_basis(gwbse_ctx* ctx, const gwbse_basis* aux, const gwbse_basis* dft, int aux_block) {
  MOCK_BEGIN(ctx)
  REQUIRE(aux && dft, "null argument");
  require_mmn(ctx);
  REQUIRE(aux->host.nfunc == ctx->naux, "aux basis does not match the Mmn tensor");
  REQUIRE(dft->host.nfunc == ctx->nbasis, "orbital basis does not match the MO coefficients (gwbse_mmn_set_mos)");
  if (aux_block < 1) aux_block = 64;
  const long pitch = (ctx->nbasis + 1) / 2 * 2;
  std::vector<double> blk((size_t)aux_block * pitch * ctx->nbasis);
  for (int a0 = 0; a0 < ctx->naux; a0 += aux_block) {
    const int cnt = std::min(aux_block, ctx->naux - a0);
    ao3c_block(aux, dft, a0, cnt, blk.data(), pitch);
    fill_block(ctx, a0, cnt, blk.data(), pitch);
  }
  MOCK_END(ctx)
}

}  // extern "C"
