// C ABI, part 3: Sigma_c evaluators (PPM and exact), batched over (level, frequency).
#include <algorithm>
#include <cstring>
#include <vector>

#include "../../include/gwbse_b200.h"
#include "context.cuh"

using namespace gwbse;

namespace {

void upload(gwbse_ctx* ctx, double** dst, const double* src, size_t n) {
  if (*dst) GW_CUDA(cudaFree(*dst));
  *dst = nullptr;
  GW_CUDA(cudaMalloc(dst, sizeof(double) * std::max<size_t>(n, 1)));
  GW_CUDA(cudaMemcpyAsync(*dst, src, sizeof(double) * n, cudaMemcpyHostToDevice, ctx->stream));
}

// groups: levels[g] with frequencies freqs[gptr[g] .. gptr[g+1])
void sigma_eval_groups(gwbse_ctx* ctx, gwbse_ctx::SigmaState& st, int ngroups, const int* levels, const int* gptr,
                       const double* freqs, double* sigma, double* dsigma) {
  GW_REQUIRE(st.ready, "sigma evaluator not prepared");
  if (ngroups <= 0) return;
  const int nfreq = gptr[ngroups];
  if (nfreq <= 0) return;
  // level -> local slice index of the matrix the evaluator reads (sharded Mmn for ppm, residues for exact)
  const bool sharded = (st.mat == ctx->X);
  std::vector<int> slices(ngroups);
  for (int i = 0; i < ngroups; ++i) {
    GW_REQUIRE(levels[i] >= 0 && levels[i] < st.q, "gw_level out of range");
    GW_REQUIRE(gptr[i + 1] >= gptr[i], "group offsets must be non-decreasing");
    const int m = st.qpoff + levels[i];
    if (sharded) {
      GW_REQUIRE(ctx->owns(m), "Sigma_c requested for a level this rank does not own");
      slices[i] = ctx->local_index(m);
    } else {
      slices[i] = m;
    }
  }
  GW_REQUIRE(ngroups <= 65535, "too many sigma request groups in one batch");
  int* lev_d = reinterpret_cast<int*>(ctx->buf("sig_levels", (size_t)ngroups / 2 + 8));
  int* gp_d = reinterpret_cast<int*>(ctx->buf("sig_gptr", (size_t)(ngroups + 1) / 2 + 8));
  double* frq_d = ctx->buf("sig_freqs", nfreq);
  GW_CUDA(cudaMemcpyAsync(lev_d, slices.data(), sizeof(int) * ngroups, cudaMemcpyHostToDevice, ctx->stream));
  GW_CUDA(cudaMemcpyAsync(gp_d, gptr, sizeof(int) * (ngroups + 1), cudaMemcpyHostToDevice, ctx->stream));
  GW_CUDA(cudaMemcpyAsync(frq_d, freqs, sizeof(double) * nfreq, cudaMemcpyHostToDevice, ctx->stream));
  double* out = ctx->buf("sig_out", (size_t)nfreq * 2);
  if ((long long)ctx->ntotal * st.npoles >= ctx->sigma_tree_min_terms) {
    // treecode over the sorted pole positions (sigma_tree.cu)
    const int which = (&st == &ctx->sig_ppm) ? 0 : 1;
    if (sharded && st.seen_mmn_version != ctx->mmn_version) {
      st.seen_mmn_version = ctx->mmn_version;
      st.content_version++;
    }
    const int nslices_total = sharded ? ctx->mlocal : st.qpoff + st.q;
    sigma_tree_eval(ctx, st, which, nslices_total, ngroups, slices.data(), gptr, frq_d, nfreq, dsigma != nullptr,
                    out);
  } else {
    const int nchunks = sigma_multi_chunks(st.npoles);
    double* partial = ctx->buf("sig_partial", (size_t)nfreq * nchunks * 2);
    launch_sigma_multi(st, ctx->ntotal, ngroups, nfreq, lev_d, gp_d, frq_d, partial, out, dsigma != nullptr,
                       ctx->stream);
    ctx->launches += 2;
  }
  GW_CUDA(cudaMemcpyAsync(sigma, out, sizeof(double) * nfreq, cudaMemcpyDeviceToHost, ctx->stream));
  if (dsigma)
    GW_CUDA(cudaMemcpyAsync(dsigma, out + nfreq, sizeof(double) * nfreq, cudaMemcpyDeviceToHost, ctx->stream));
  GW_CUDA(cudaStreamSynchronize(ctx->stream));
}

void sigma_eval(gwbse_ctx* ctx, gwbse_ctx::SigmaState& st, int nreq, const int* levels, const double* freqs,
                double* sigma, double* dsigma) {
  if (nreq <= 0) return;
  // one group per request (consecutive requests of the same level are merged)
  std::vector<int> lv, gp;
  for (int i = 0; i < nreq; ++i) {
    if (i == 0 || levels[i] != levels[i - 1]) {
      lv.push_back(levels[i]);
      gp.push_back(i);
    }
  }
  gp.push_back(nreq);
  sigma_eval_groups(ctx, st, (int)lv.size(), lv.data(), gp.data(), freqs, sigma, dsigma);
}

// Off-diagonal Sigma_c for all level pairs as a weighted GEMM:
//   S[i,j] = sum_{p,n} (pref fac_p g(w_i - e_n +- pole_p) M_i[n,p]) M_j[n,p],  Sigma_c[i,j] = S[i,j] + S[j,i]
// sigma_ppm.cc:93-126 / sigma_exact.cc:85-107 evaluate the same sum pair by pair.
void sigma_offdiag(gwbse_ctx* ctx, gwbse_ctx::SigmaState& st, double pref, int q, const double* freqs, double* out,
                   int ld) {
  GW_REQUIRE(st.ready, "sigma evaluator not prepared");
  GW_REQUIRE(q == st.q && ld >= q, "q does not match the prepared evaluator");
  const bool sharded = (st.mat == ctx->X) && ctx->world > 1;
  double* frq_d = ctx->buf("sig_freqs", q);
  GW_CUDA(cudaMemcpyAsync(frq_d, freqs, sizeof(double) * q, cudaMemcpyHostToDevice, ctx->stream));
  const int npad = (int)st.lstride;
  // levels whose rows this rank computes: all of them on one GPU, the owned ones when Mmn is sharded
  std::vector<int> slice_idx, freq_idx;
  int ifirst = 0, istride = 1;
  if (sharded) {
    ifirst = ctx->first_owned(st.qpoff, ctx->rank) - st.qpoff;
    istride = ctx->world;
    for (int i = ifirst; i < q; i += istride) {
      slice_idx.push_back(ctx->local_index(st.qpoff + i));
      freq_idx.push_back(i);
    }
  } else {
    for (int i = 0; i < q; ++i) {
      slice_idx.push_back((st.mat == ctx->X ? ctx->local_index(st.qpoff + i) : st.qpoff + i));
      freq_idx.push_back(i);
    }
  }
  const int qloc = (int)slice_idx.size();
  int* sl_d = reinterpret_cast<int*>(ctx->buf("sig_off_slices", (size_t)q / 2 + 8));
  int* fi_d = reinterpret_cast<int*>(ctx->buf("sig_off_fidx", (size_t)q / 2 + 8));
  if (qloc) {
    GW_CUDA(cudaMemcpyAsync(sl_d, slice_idx.data(), sizeof(int) * qloc, cudaMemcpyHostToDevice, ctx->stream));
    GW_CUDA(cudaMemcpyAsync(fi_d, freq_idx.data(), sizeof(int) * qloc, cudaMemcpyHostToDevice, ctx->stream));
  }
  const long long ldo = (long long)std::max(qloc, 1) * npad;
  const long long ldg = (long long)q * npad;
  const size_t budget = (size_t)1 << 25;  // doubles (256 MiB)
  const int chunk = (int)std::max<size_t>(1, std::min<size_t>(std::min(st.npoles, 65535), budget / (size_t)ldg));
  double* A = ctx->buf("sig_offdiag_A", (size_t)ldo * chunk);
  double* G = sharded ? ctx->buf("sig_offdiag_G", (size_t)ldg * chunk) : nullptr;
  double* S = ctx->buf("sig_offdiag_S", (size_t)q * q * 2);
  GW_CUDA(cudaMemsetAsync(S, 0, sizeof(double) * (size_t)q * q, ctx->stream));
  for (int p0 = 0; p0 < st.npoles; p0 += chunk) {
    const int np = std::min(chunk, st.npoles - p0);
    if (sharded) gather_slices(ctx, st.qpoff, q, 0, ctx->ntotal, p0, np, G, ldg, npad);
    if (qloc == 0) continue;
    launch_sigma_offdiag_weight(st, ctx->ntotal, npad, qloc, sl_d, fi_d, p0, np, frq_d, pref, A, ldo, ctx->stream);
    ctx->launches++;
    GemmParams p;
    p.M = qloc;
    p.N = q;
    p.Ko = np;
    p.Ki = ctx->ntotal;
    p.A.ptr = A;
    p.A.s_ri = npad;
    p.A.s_ki = 1;
    p.A.s_ko = ldo;
    if (sharded) {
      p.B.ptr = G;
      p.B.s_ri = npad;
      p.B.s_ko = ldg;
    } else {
      p.B.ptr = st.mat + (long long)slice_idx[0] * st.lstride + (long long)p0 * st.ld;
      p.B.s_ri = st.lstride;
      p.B.s_ko = st.ld;
    }
    p.B.s_ki = 1;
    p.C = S + ifirst;  // rows land at their global level index
    p.sC_mi = istride;
    p.sC_ni = q;
    p.beta = 1.0;
    ctx->gemm(p);
  }
  if (sharded) allreduce_dev(ctx, S, (size_t)q * q);
  launch_offdiag_finish(S, q, S + (size_t)q * q, ctx->stream);
  ctx->launches++;
  GW_CUDA(copy2d_async(out, sizeof(double) * ld, S + (size_t)q * q, sizeof(double) * q, sizeof(double) * q, q,
                            cudaMemcpyDeviceToHost, ctx->stream));
  GW_CUDA(cudaStreamSynchronize(ctx->stream));
}

}  // namespace

extern "C" {

int gwbse_sigma_ppm_set(gwbse_ctx* ctx, const double* ppm_weight, const double* ppm_freq, const double* energies,
                        int homo, int rpamin, int qpmin, double eta) {
  GW_API_BEGIN(ctx)
  GW_REQUIRE(ctx->X != nullptr, "Mmn not allocated");
  mmn_complete_rotation(ctx);
  auto& st = ctx->sig_ppm;
  const int naux = ctx->naux;
  std::vector<double> fac(naux);
  for (int i = 0; i < naux; ++i)  // sigma_ppm.cc:47-52: weights below 1e-9 are skipped
    fac[i] = (ppm_weight[i] < 1.e-9) ? 0.0 : ppm_weight[i] * ppm_freq[i];
  upload(ctx, &st.fac, fac.data(), naux);
  upload(ctx, &st.pole, ppm_freq, naux);
  upload(ctx, &st.energies, energies, ctx->ntotal);
  st.npoles = naux;
  st.nocc_boundary = homo + 1;  // sigma_ppm.cc:40,56-57 uses lumo = homo + 1 unshifted
  st.qpoff = qpmin - rpamin;
  st.q = ctx->mtotal - st.qpoff;
  st.eta = eta;
  st.diag_pref = 0.5;
  st.mat = ctx->X;
  st.ld = ctx->ldx;
  st.lstride = ctx->npad;
  st.ready = true;
  st.energies_host.assign(energies, energies + ctx->ntotal);
  st.content_version++;
  sigma_tree_invalidate(st.tree);
  GW_CUDA(cudaStreamSynchronize(ctx->stream));
  GW_API_END(ctx)
}

int gwbse_sigma_update_energies(gwbse_ctx* ctx, int which, const double* energies) {
  GW_API_BEGIN(ctx)
  auto& st = which == 0 ? ctx->sig_ppm : ctx->sig_exact;
  GW_REQUIRE(st.ready && st.energies != nullptr, "sigma evaluator not prepared");
  if ((int)st.energies_host.size() == ctx->ntotal &&
      std::memcmp(st.energies_host.data(), energies, sizeof(double) * ctx->ntotal) == 0)
    return 0;  // unchanged: keep the device copy and the treecode geometry
  st.energies_host.assign(energies, energies + ctx->ntotal);
  sigma_tree_invalidate(st.tree);
  GW_CUDA(cudaMemcpyAsync(st.energies, energies, sizeof(double) * ctx->ntotal, cudaMemcpyHostToDevice, ctx->stream));
  GW_CUDA(cudaStreamSynchronize(ctx->stream));
  GW_API_END(ctx)
}

int gwbse_sigma_ppm_eval(gwbse_ctx* ctx, int nreq, const int* levels, const double* freqs, double* sigma,
                         double* dsigma) {
  GW_API_BEGIN(ctx)
  GW_PROF(ctx, "sigma_ppm_eval");
  mmn_complete_rotation(ctx);
  ctx->sig_ppm.mat = ctx->X;
  sigma_eval(ctx, ctx->sig_ppm, nreq, levels, freqs, sigma, dsigma);
  GW_API_END(ctx)
}

int gwbse_sigma_eval_groups(gwbse_ctx* ctx, int which, int ngroups, const int* levels, const int* group_ptr,
                            const double* freqs, double* sigma, double* dsigma) {
  GW_API_BEGIN(ctx)
  GW_PROF(ctx, "sigma_eval_groups");
  auto& st = which == 0 ? ctx->sig_ppm : ctx->sig_exact;
  if (which == 0) {
    mmn_complete_rotation(ctx);
    st.mat = ctx->X;
  }
  sigma_eval_groups(ctx, st, ngroups, levels, group_ptr, freqs, sigma, dsigma);
  GW_API_END(ctx)
}

int gwbse_sigma_ppm_offdiag(gwbse_ctx* ctx, int q, const double* freqs, double* out, int ld) {
  GW_API_BEGIN(ctx)
  GW_PROF(ctx, "sigma_ppm_offdiag");
  mmn_complete_rotation(ctx);
  ctx->sig_ppm.mat = ctx->X;
  const int qsave = ctx->sig_ppm.q;
  ctx->sig_ppm.q = q;
  try {
    sigma_offdiag(ctx, ctx->sig_ppm, 0.25, q, freqs, out, ld);
  } catch (...) {
    ctx->sig_ppm.q = qsave;
    throw;
  }
  ctx->sig_ppm.q = qsave;
  GW_API_END(ctx)
}

// Z[chi, s] (+)= sum_{v,c} M[v][c,chi] XpY[(v,c), s]: the projection of (X+Y) columns on the auxiliary basis
// (sigma_exact.cc:119-145 first half; rpa_uks.cc:103-121 per spin channel, where both channels add into one Z)
static void exact_project(gwbse_ctx* ctx, const double* XpY_dev, int ldxpy, int ncols, int homo, int rpamin, int rpamax,
                          bool accumulate, double* Z, int ldz) {
  GW_REQUIRE(ctx->X != nullptr, "Mmn not allocated");
  mmn_complete_rotation(ctx);
  GW_REQUIRE(ctx->world == 1, "exact sigma is single-GPU (SURVEY.md 8e)");
  GW_REQUIRE(rpamin == ctx->mmin && rpamin == ctx->nmin && rpamax == ctx->nmax, "RPA range must match Mmn");
  const int n_occ = homo + 1 - rpamin, n_unocc = rpamax - homo;
  const int S = n_occ * n_unocc;
  GW_REQUIRE(ldxpy >= S && ldz >= ctx->naux && ncols >= 0, "invalid sizes");
  if (ncols == 0 || S == 0) return;
  GemmParams p;
  p.M = ctx->naux;
  p.N = ncols;
  p.Ko = n_occ;
  p.Ki = n_unocc;
  const HoleView hv = hole_view(ctx, n_occ);  // QSGW: rotated inside the QP window (sigma_exact.cc:119-145)
  p.A.ptr = hv.ptr;
  p.A.s_ri = hv.s_chi;
  p.A.s_ki = 1;
  p.A.s_ko = hv.s_v;
  p.B.ptr = XpY_dev;
  p.B.s_ri = ldxpy;
  p.B.s_ki = 1;
  p.B.s_ko = n_unocc;
  p.C = Z;
  p.sC_mi = 1;
  p.sC_ni = ldz;
  p.beta = accumulate ? 1.0 : 0.0;
  ctx->gemm(p);
}

// R[(i,n), s] = sum_chi M_i[n,chi] Z[chi,s] for the qp window, and the evaluator state around it
static void exact_install_modes(gwbse_ctx* ctx, const double* omegas, int nmodes, const double* Z, int ldz,
                                const double* energies, int homo, int rpamin, int rpamax, int qpmin, int qpmax,
                                double eta, double diag_pref, double offdiag_pref) {
  GW_REQUIRE(ctx->X != nullptr, "Mmn not allocated");
  mmn_complete_rotation(ctx);
  GW_REQUIRE(ctx->world == 1, "exact sigma is single-GPU (SURVEY.md 8e)");
  GW_REQUIRE(rpamin == ctx->mmin && rpamin == ctx->nmin && rpamax == ctx->nmax, "RPA range must match Mmn");
  const int n_occ = homo + 1 - rpamin;
  const int q = qpmax - qpmin + 1, qpoff = qpmin - rpamin;
  const int naux = ctx->naux, npad = ctx->npad;
  GW_REQUIRE(nmodes > 0 && ldz >= naux && q > 0 && qpoff >= 0 && qpoff + q <= ctx->mtotal, "invalid sizes");
  const long long ldr = (long long)q * npad;
  if (ctx->exact_res) GW_CUDA(cudaFree(ctx->exact_res));
  ctx->exact_res = nullptr;
  GW_CUDA(cudaMalloc(&ctx->exact_res, sizeof(double) * (size_t)ldr * nmodes));
  GemmParams r;
  r.M = (int)ldr;
  r.N = nmodes;
  r.Ki = naux;
  r.A.ptr = ctx->X + (long long)qpoff * npad;
  r.A.s_ri = 1;
  r.A.s_ki = ctx->ldx;
  r.B.ptr = Z;
  r.B.s_ri = ldz;
  r.B.s_ki = 1;
  r.C = ctx->exact_res;
  r.sC_mi = 1;
  r.sC_ni = ldr;
  ctx->gemm(r);
  auto& st = ctx->sig_exact;
  std::vector<double> fac(nmodes, 1.0);
  upload(ctx, &st.fac, fac.data(), nmodes);
  upload(ctx, &st.pole, omegas, nmodes);
  upload(ctx, &st.energies, energies, ctx->ntotal);
  st.npoles = nmodes;
  st.nocc_boundary = n_occ;
  st.qpoff = 0;
  st.q = q;
  st.eta = eta;
  st.diag_pref = diag_pref;
  st.offdiag_pref = offdiag_pref;
  st.mat = ctx->exact_res;
  st.ld = ldr;
  st.lstride = npad;
  st.ready = true;
  st.energies_host.assign(energies, energies + ctx->ntotal);
  st.content_version++;
  sigma_tree_invalidate(st.tree);
  GW_CUDA(cudaStreamSynchronize(ctx->stream));
}

int gwbse_sigma_exact_prepare(gwbse_ctx* ctx, const double* rpa_omegas, const double* XpY_dev, int ldxpy,
                              const double* energies, int homo, int rpamin, int rpamax, int qpmin, int qpmax,
                              double eta) {
  GW_API_BEGIN(ctx)
  GW_PROF(ctx, "sigma_exact_prepare");
  const int S = (homo + 1 - rpamin) * (rpamax - homo);
  GW_REQUIRE(S > 0, "invalid sizes");
  double* Z = ctx->buf("exact_Z", (size_t)ctx->naux * S);
  exact_project(ctx, XpY_dev, ldxpy, S, homo, rpamin, rpamax, false, Z, ctx->naux);
  // closed shell: 2 * sum (sigma_exact.cc:57, :76), off-diagonal 2 * 0.5 * (...) (:103-106)
  exact_install_modes(ctx, rpa_omegas, S, Z, ctx->naux, energies, homo, rpamin, rpamax, qpmin, qpmax, eta, 2.0, 1.0);
  GW_API_END(ctx)
}

int gwbse_sigma_exact_project(gwbse_ctx* ctx, const double* XpY_dev, int ldxpy, int ncols, int homo, int rpamin,
                              int rpamax, int accumulate, double* Z_dev, int ldz) {
  GW_API_BEGIN(ctx)
  GW_PROF(ctx, "sigma_exact_project");
  exact_project(ctx, XpY_dev, ldxpy, ncols, homo, rpamin, rpamax, accumulate != 0, Z_dev, ldz);
  GW_CUDA(cudaStreamSynchronize(ctx->stream));
  GW_API_END(ctx)
}

int gwbse_sigma_exact_prepare_modes(gwbse_ctx* ctx, const double* omegas, int nmodes, const double* Z_dev, int ldz,
                                    const double* energies, int homo, int rpamin, int rpamax, int qpmin, int qpmax,
                                    double eta, double diag_pref, double offdiag_pref) {
  GW_API_BEGIN(ctx)
  GW_PROF(ctx, "sigma_exact_prepare");
  exact_install_modes(ctx, omegas, nmodes, Z_dev, ldz, energies, homo, rpamin, rpamax, qpmin, qpmax, eta, diag_pref,
                      offdiag_pref);
  GW_API_END(ctx)
}

int gwbse_sigma_exact_eval(gwbse_ctx* ctx, int nreq, const int* levels, const double* freqs, double* sigma,
                           double* dsigma) {
  GW_API_BEGIN(ctx)
  GW_PROF(ctx, "sigma_exact_eval");
  sigma_eval(ctx, ctx->sig_exact, nreq, levels, freqs, sigma, dsigma);
  GW_API_END(ctx)
}

int gwbse_sigma_exact_offdiag(gwbse_ctx* ctx, int q, const double* freqs, double* out, int ld) {
  GW_API_BEGIN(ctx)
  GW_PROF(ctx, "sigma_exact_offdiag");
  sigma_offdiag(ctx, ctx->sig_exact, ctx->sig_exact.offdiag_pref, q, freqs, out, ld);
  GW_API_END(ctx)
}

}  // extern "C"
