// C ABI, part 4: the factorised BSE operator and the Davidson device helpers.
//
// BSE_OPERATOR<cqp,cx,cd,cd2>::matmul (bse_operator.cc:40-119) rebuilds every row
// of H for every product (2 B^2 (Naux + k) flops).  Here H is never formed:
//   Hx  : W = Avc^T X            then  Y += cx  Avc W                 (4 B Naux k)
//   Hd  : U = (Mcc^T X) eps^-1   then  Y -= cd  Mvv U                 (2 Naux k B (vt+ct))
//   Hd2 : U = (Mvc^T X) eps^-1   then  Y -= cd2 Mcv U
//   Hqp : two small GEMMs with the cc and vv blocks of Hqp            (2 B (vt+ct) k)
// all on sub-blocks of the device-resident Mmn consumed in place by the DMMA GEMM.
#include <algorithm>
#include <vector>

#include "../../include/gwbse_b200.h"
#include "context.cuh"

using namespace gwbse;

namespace {

inline int round_up(int v, int a) { return (v + a - 1) / a * a; }

void bse_matmul_dev(gwbse_ctx* ctx, int cqp, int cx, int cd, int cd2, int k, const double* Xin, int ldin, double* Y,
                    int ldy) {
  auto& st = ctx->bse;
  GW_REQUIRE(st.ready, "BSE operator not configured (gwbse_bse_configure)");
  GW_REQUIRE(!(cd != 0 && cd2 != 0), "Hamiltonian cannot contain Hd and Hd2 at the same time");
  GW_REQUIRE(ctx->world == 1, "gwbse_bse_matmul: single-GPU build");
  const int vt = st.vt, ct = st.ct, B = st.size, naux = ctx->naux, npad = ctx->npad;
  const int voff = st.voff, coff = st.coff;
  const long long ldx = ctx->ldx;
  const double* X = ctx->X;
  GW_REQUIRE(ldin >= B && ldy >= B, "Shape mismatch in BSE matmul");
  if (k <= 0) return;
  GW_CUDA(cudaMemset2DAsync(Y, sizeof(double) * ldy, 0, sizeof(double) * B, k, ctx->stream));

  if (cqp != 0) {
    const int ldh = vt + ct;
    // Y[(v1,c1)] += cqp sum_c2 Hqp[vt+c2, vt+c1] X[(v1,c2)]
    GemmParams p;
    p.M = ct;
    p.N = vt * k;
    p.Ki = ct;
    p.A.ptr = st.hqp + vt + (long long)vt * ldh;
    p.A.s_ri = ldh;
    p.A.s_ki = 1;
    p.B.ptr = Xin;
    p.B.Lr = vt;
    p.B.s_ri = ct;
    p.B.s_ro = ldin;
    p.B.s_ki = 1;
    p.C = Y;
    p.sC_mi = 1;
    p.Ln = vt;
    p.sC_ni = ct;
    p.sC_no = ldy;
    p.alpha = cqp;
    p.beta = 1.0;
    ctx->gemm(p);
    // Y[(v1,c1)] -= cqp sum_v2 Hqp[v2, v1] X[(v2,c1)]      (batched over the k vectors)
    GemmParams q;
    q.M = ct;
    q.N = vt;
    q.Ki = vt;
    q.Z1 = k;
    q.A.ptr = Xin;
    q.A.s_ri = 1;
    q.A.s_ki = ct;
    q.A.s_z1 = ldin;
    q.B.ptr = st.hqp;
    q.B.s_ri = ldh;
    q.B.s_ki = 1;
    q.C = Y;
    q.sC_mi = 1;
    q.sC_ni = ct;
    q.sC_z1 = ldy;
    q.alpha = -cqp;
    q.beta = 1.0;
    ctx->gemm(q);
  }

  if (cx != 0) {
    double* W = ctx->buf("bse_W", (size_t)naux * k);
    // W[chi, kv] = sum_{v,c} M[v][c,chi] X[(v,c),kv]
    GemmParams p;
    p.M = naux;
    p.N = k;
    p.Ko = vt;
    p.Ki = ct;
    p.A.ptr = X + (long long)voff * npad + coff;
    p.A.s_ri = ldx;
    p.A.s_ki = 1;
    p.A.s_ko = npad;
    p.B.ptr = Xin;
    p.B.s_ri = ldin;
    p.B.s_ki = 1;
    p.B.s_ko = ct;
    p.C = W;
    p.sC_mi = 1;
    p.sC_ni = naux;
    ctx->gemm(p);
    // Y[(v,c), kv] += cx sum_chi M[v][c,chi] W[chi,kv]
    GemmParams q;
    q.M = B;
    q.N = k;
    q.Ki = naux;
    q.A.ptr = X + (long long)voff * npad + coff;
    q.A.Lr = ct;
    q.A.s_ri = 1;
    q.A.s_ro = npad;
    q.A.s_ki = ldx;
    q.B.ptr = W;
    q.B.s_ri = naux;
    q.B.s_ki = 1;
    q.C = Y;
    q.sC_mi = 1;
    q.sC_ni = ldy;
    q.alpha = cx;
    q.beta = 1.0;
    ctx->gemm(q);
  }

  if (cd != 0 || cd2 != 0) {
    const int vtp = round_up(vt, 2);
    const long long ldU = (long long)vtp * naux;
    const int nout = cd != 0 ? ct : vt;  // index that is chunked (c1 for Hd, v1 for Hd2)
    int nc = (int)std::max<long long>(1, (long long)(ctx->bse_chunk_bytes / sizeof(double)) / (ldU * k));
    nc = std::min(nc, nout);
    while ((long long)naux * nc >= (1LL << 31) || (long long)nc * k >= (1LL << 31)) nc = std::max(1, nc / 2);
    double* U = ctx->buf("bse_U", (size_t)ldU * nc * k);
    for (int a = 0; a < nout; a += nc) {
      const int n1 = std::min(nc, nout - a);
      // U[(v2, chi), (l, kv)] = eps_inv[chi] sum_c2 X[c2,(v2,kv)] Mblk[l][c2, chi]
      GemmParams p;
      p.M = vt * k;
      p.N = naux * n1;
      p.Ki = ct;
      p.A.ptr = Xin;
      p.A.Lr = vt;
      p.A.s_ri = ct;
      p.A.s_ro = ldin;
      p.A.s_ki = 1;
      p.B.ptr = X + (long long)((cd != 0 ? coff : voff) + a) * npad + coff;
      p.B.Lr = naux;
      p.B.s_ri = ldx;
      p.B.s_ro = npad;
      p.B.s_ki = 1;
      p.C = U;
      p.Lm = vt;
      p.sC_mi = 1;
      p.sC_mo = ldU * n1;
      p.Ln = naux;
      p.sC_ni = vtp;
      p.sC_no = ldU;
      p.nscale = st.eps_inv;
      p.nscale_mod = naux;
      ctx->gemm(p);
      GemmParams q;
      q.Ko = naux;
      q.Ki = vt;
      q.N = n1 * k;
      q.A.s_ri = npad;
      q.A.s_ki = 1;
      q.A.s_ko = ldx;
      q.B.ptr = U;
      q.B.s_ri = ldU;
      q.B.s_ki = 1;
      q.B.s_ko = vtp;
      q.beta = 1.0;
      q.Ln = n1;
      q.sC_no = ldy;
      if (cd != 0) {
        // Y[(v1, a+l), kv] -= cd sum_{chi,v2} M[v1][v2,chi] U[(v2,chi),(l,kv)]
        q.M = vt;
        q.A.ptr = X + (long long)voff * npad + voff;
        q.C = Y + a;
        q.sC_mi = ct;
        q.sC_ni = 1;
        q.alpha = -cd;
      } else {
        // Y[(a+l, c1), kv] -= cd2 sum_{chi,v2} M[c1][v2,chi] U[(v2,chi),(l,kv)]
        q.M = ct;
        q.A.ptr = X + (long long)coff * npad + voff;
        q.C = Y + (long long)a * ct;
        q.sC_mi = 1;
        q.sC_ni = ct;
        q.alpha = -cd2;
      }
      ctx->gemm(q);
    }
  }
}

}  // namespace

extern "C" {

int gwbse_bse_configure(gwbse_ctx* ctx, int homo, int rpamin, int vmin, int cmax, const double* eps_inv,
                        const double* Hqp, int ldh) {
  GW_API_BEGIN(ctx)
  GW_PROF(ctx, "bse_configure");
  GW_REQUIRE(ctx->X != nullptr, "Mmn not allocated");
  GW_REQUIRE(rpamin == ctx->mmin && rpamin == ctx->nmin, "RPA range must match Mmn");
  auto& st = ctx->bse;
  st.homo = homo;
  st.rpamin = rpamin;
  st.vmin = vmin;
  st.cmax = cmax;
  st.vt = homo - vmin + 1;
  st.ct = cmax - homo;
  st.size = st.vt * st.ct;
  st.voff = vmin - rpamin;
  st.coff = homo + 1 - rpamin;
  GW_REQUIRE(st.vt > 0 && st.ct > 0 && st.voff >= 0, "invalid BSE level ranges");
  GW_REQUIRE(st.coff + st.ct <= ctx->mtotal && st.coff + st.ct <= ctx->ntotal, "BSE range exceeds Mmn");
  const int hs = st.vt + st.ct;
  GW_REQUIRE(ldh >= hs, "Hqp leading dimension too small");
  st.eps_inv = ctx->buf("bse_eps_inv", ctx->naux);
  st.hqp = ctx->buf("bse_hqp", (size_t)hs * hs);
  GW_CUDA(cudaMemcpyAsync(st.eps_inv, eps_inv, sizeof(double) * ctx->naux, cudaMemcpyHostToDevice, ctx->stream));
  GW_CUDA(cudaMemcpy2DAsync(st.hqp, sizeof(double) * hs, Hqp, sizeof(double) * ldh, sizeof(double) * hs, hs,
                            cudaMemcpyHostToDevice, ctx->stream));
  GW_CUDA(cudaStreamSynchronize(ctx->stream));
  st.ready = true;
  GW_API_END(ctx)
}

int gwbse_bse_matmul_dev(gwbse_ctx* ctx, int cqp, int cx, int cd, int cd2, int k, const double* X_dev, int ldx,
                         double* Y_dev, int ldy) {
  GW_API_BEGIN(ctx)
  GW_PROF(ctx, "bse_matmul");
  bse_matmul_dev(ctx, cqp, cx, cd, cd2, k, X_dev, ldx, Y_dev, ldy);
  GW_API_END(ctx)
}

int gwbse_bse_matmul(gwbse_ctx* ctx, int cqp, int cx, int cd, int cd2, int k, const double* X, int ldx, double* Y,
                     int ldy) {
  GW_API_BEGIN(ctx)
  GW_PROF(ctx, "bse_matmul_h2d");
  const int B = ctx->bse.size;
  GW_REQUIRE(ctx->bse.ready, "BSE operator not configured (gwbse_bse_configure)");
  GW_REQUIRE(ldx >= B && ldy >= B, "Shape mismatch in BSE matmul");
  if (k > 0) {
    double* Xd = ctx->buf("bse_Xin", (size_t)B * k);
    double* Yd = ctx->buf("bse_Yout", (size_t)B * k);
    GW_CUDA(cudaMemcpy2DAsync(Xd, sizeof(double) * B, X, sizeof(double) * ldx, sizeof(double) * B, k,
                              cudaMemcpyHostToDevice, ctx->stream));
    bse_matmul_dev(ctx, cqp, cx, cd, cd2, k, Xd, B, Yd, B);
    GW_CUDA(cudaMemcpy2DAsync(Y, sizeof(double) * ldy, Yd, sizeof(double) * B, sizeof(double) * B, k,
                              cudaMemcpyDeviceToHost, ctx->stream));
    GW_CUDA(cudaStreamSynchronize(ctx->stream));
  }
  GW_API_END(ctx)
}

int gwbse_bse_diagonal(gwbse_ctx* ctx, int cqp, int cx, int cd, int cd2, double* diag) {
  GW_API_BEGIN(ctx)
  GW_PROF(ctx, "bse_diagonal");
  auto& st = ctx->bse;
  GW_REQUIRE(st.ready, "BSE operator not configured (gwbse_bse_configure)");
  GW_REQUIRE(!(cd != 0 && cd2 != 0), "Hamiltonian cannot contain Hd and Hd2 at the same time");
  GW_REQUIRE(ctx->world == 1, "gwbse_bse_diagonal: single-GPU build");
  double* d = ctx->buf("bse_diag", st.size);
  launch_bse_diag(ctx->X, ctx->ldx, ctx->npad, ctx->naux, st.vt, st.ct, st.voff, st.coff, st.eps_inv, st.hqp,
                  st.vt + st.ct, cqp, cx, cd, cd2, d, ctx->stream);
  ctx->launches++;
  GW_CUDA(cudaMemcpyAsync(diag, d, sizeof(double) * st.size, cudaMemcpyDeviceToHost, ctx->stream));
  GW_CUDA(cudaStreamSynchronize(ctx->stream));
  GW_API_END(ctx)
}

// ------------------------- Davidson device helpers --------------------------
// DavidsonSolver::gramschmidt, davidsonsolver.cc:442-478, same order of operations.
int gwbse_gramschmidt_dev(gwbse_ctx* ctx, int rows, int ncols, int nstart, double* Q, int ldq) {
  GW_API_BEGIN(ctx)
  GW_PROF(ctx, "gramschmidt");
  GW_REQUIRE(ldq >= rows && nstart >= 0 && nstart <= ncols, "invalid Gram-Schmidt arguments");
  const int nup = ncols - nstart;
  if (nup > 0) {
    std::vector<double> norms0(nup), nrm(nup), inv(nup);
    double* red = ctx->buf("gs_red", ncols);
    double* coef = ctx->buf("gs_coef", (size_t)ncols * std::max(nup, 1));
    double* Qn = Q + (size_t)nstart * ldq;
    auto colnorms = [&](int j0, int n, double* out) {
      launch_colnorms(rows, n, Q + (size_t)j0 * ldq, ldq, red, ctx->stream);
      ctx->launches++;
      GW_CUDA(cudaMemcpyAsync(out, red, sizeof(double) * n, cudaMemcpyDeviceToHost, ctx->stream));
      GW_CUDA(cudaStreamSynchronize(ctx->stream));
    };
    auto normalize = [&](int j0, int n, const double* nr) {
      for (int j = 0; j < n; ++j) inv[j] = 1.0 / nr[j];
      GW_CUDA(cudaMemcpyAsync(red, inv.data(), sizeof(double) * n, cudaMemcpyHostToDevice, ctx->stream));
      launch_scale_cols(rows, n, Q + (size_t)j0 * ldq, ldq, red, ctx->stream);
      ctx->launches++;
      GW_CUDA(cudaStreamSynchronize(ctx->stream));
    };
    colnorms(nstart, nup, norms0.data());
    for (int rep = 0; rep < 2; ++rep) {
      if (nstart > 0) {
        // Qn -= Qold (Qold^T Qn)
        if (gwbse_dgemm_dev(ctx, 'T', 'N', nstart, nup, rows, 1.0, Q, ldq, Qn, ldq, 0.0, coef, nstart))
          throw std::runtime_error(ctx->err);
        if (gwbse_dgemm_dev(ctx, 'N', 'N', rows, nup, nstart, -1.0, Q, ldq, coef, nstart, 1.0, Qn, ldq))
          throw std::runtime_error(ctx->err);
        colnorms(nstart, nup, nrm.data());
        normalize(nstart, nup, nrm.data());
      }
      for (int j = nstart + 1; j < ncols; ++j) {
        const int range = j - nstart;
        double* qj = Q + (size_t)j * ldq;
        if (gwbse_dgemm_dev(ctx, 'T', 'N', range, 1, rows, 1.0, Qn, ldq, qj, ldq, 0.0, coef, range))
          throw std::runtime_error(ctx->err);
        if (gwbse_dgemm_dev(ctx, 'N', 'N', rows, 1, range, -1.0, Qn, ldq, coef, range, 1.0, qj, ldq))
          throw std::runtime_error(ctx->err);
        double nj = 0.0;
        colnorms(j, 1, &nj);
        if (rep == 1 && nj <= 1e-12 * norms0[range]) throw std::runtime_error("Linear dependencies in Gram-Schmidt.");
        normalize(j, 1, &nj);
      }
    }
  }
  GW_API_END(ctx)
}

int gwbse_davidson_correction_dev(gwbse_ctx* ctx, int rows, int ncols, int olsen, const double* diag_dev,
                                  const double* lambda, const double* R, int ldr, const double* Q, int ldq,
                                  double* W, int ldw) {
  GW_API_BEGIN(ctx)
  GW_PROF(ctx, "davidson_correction");
  if (ncols > 0) {
    double* lam = ctx->buf("dav_lambda", ncols);
    GW_CUDA(cudaMemcpyAsync(lam, lambda, sizeof(double) * ncols, cudaMemcpyHostToDevice, ctx->stream));
    launch_dpr(rows, ncols, diag_dev, lam, R, ldr, W, ldw, ctx->stream);
    ctx->launches++;
    if (olsen) {
      double* tmp = ctx->buf("dav_tmp", (size_t)rows * ncols);
      double* num = ctx->buf("dav_num", ncols);
      double* den = ctx->buf("dav_den", ncols);
      launch_dpr(rows, ncols, diag_dev, lam, Q, ldq, tmp, rows, ctx->stream);  // dpr(x)
      launch_coldots(rows, ncols, Q, ldq, W, ldw, num, ctx->stream);          // x . dpr(r)
      launch_coldots(rows, ncols, Q, ldq, tmp, rows, den, ctx->stream);       // x . dpr(x)
      launch_olsen_finish(rows, ncols, diag_dev, lam, Q, ldq, num, den, W, ldw, ctx->stream);
      ctx->launches += 4;
    }
    // w.normalized(), davidsonsolver.cc:372
    std::vector<double> nr(ncols);
    double* red = ctx->buf("gs_red", ncols);
    launch_colnorms(rows, ncols, W, ldw, red, ctx->stream);
    GW_CUDA(cudaMemcpyAsync(nr.data(), red, sizeof(double) * ncols, cudaMemcpyDeviceToHost, ctx->stream));
    GW_CUDA(cudaStreamSynchronize(ctx->stream));
    for (int j = 0; j < ncols; ++j) nr[j] = 1.0 / nr[j];
    GW_CUDA(cudaMemcpyAsync(red, nr.data(), sizeof(double) * ncols, cudaMemcpyHostToDevice, ctx->stream));
    launch_scale_cols(rows, ncols, W, ldw, red, ctx->stream);
    ctx->launches += 2;
    GW_CUDA(cudaStreamSynchronize(ctx->stream));
  }
  GW_API_END(ctx)
}

}  // extern "C"
