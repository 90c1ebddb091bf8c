// C ABI, part 6: contour-deformation self-energy (Sigma_CDA) resident on the device.
//
// Reference: Sigma_CDA::PrepareScreening / CalcCorrelationDiagElement / CalcResidueContribution
// (xtp/src/libxtp/self_energy_evaluators/sigma_cda.cc:30-141) and ImaginaryAxisIntegration::CalcDielInvVector /
// SigmaGQDiag (xtp/src/libxtp/ImaginaryAxisIntegration.cc:90-176).  The reference evaluates, per (level, omega),
//   sum_j w_j/(2 pi) Re sum_n [ (I kappa_j)[n,:] . I[n,:] ] (1/(dE_n + i w_j) [+ 1/(dE_n - i w_j)])      quadrature
// + sum_n tail(e_n - omega) (I kappa_0)[n,:] . I[n,:]                                                    Gaussian tail
// + sum_{n: Theta != 0} Theta [ eps(|e_n - omega| + i eta)^-1 I[n,:] - I[n,:] ] . I[n,:]                  residues
// with I = Mmn[level] (n x Naux): one n x Naux x Naux product per node and evaluation.  The row forms
//   Q_j[level][n] = (I kappa_j)[n,:] . I[n,:]
// do not depend on omega, so they are built once per screening update for every level with ONE MultiplyRight-shaped
// DMMA GEMM per node over the whole device-resident Mmn plus a streaming row-dot pass; an evaluation of the
// quadrature and tail parts is then a sum over (order + 1) * n numbers.  The residue terms keep the reference's
// formulation (a full eps(z) assembly and an LU solve per pole inside the contour) but never leave the device: the
// row is gathered from Mmn by a kernel, the dot product is reduced on the device, one D2H of the results per batch.
#include <algorithm>
#include <cmath>
#include <vector>

#include "../../include/gwbse_b200.h"
#include "context.cuh"

using namespace gwbse;

namespace {

inline int round_up(int v, int a) { return (v + a - 1) / a * a; }

// k <- -(k - 1) + scale * kzero      (ImaginaryAxisIntegration.cc:95-101)
__global__ void cda_kappa_kernel(double* k, const double* kzero, int n, double scale) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int j = blockIdx.y;
  if (i >= n) return;
  const long long o = i + (long long)j * n;
  k[o] = -(k[o] - (i == j ? 1.0 : 0.0)) + scale * kzero[o];
}
__global__ void cda_minus_identity_kernel(double* k, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) k[i + (long long)i * n] -= 1.0;
}

// Q[r] = sum_chi X[chi * ldx + r] * T[chi * ldx + r] for the rows r of the qp-window slices (r contiguous)
__global__ void cda_rowdot_kernel(const double* X, const double* T, long long ldx, int naux, long long rows,
                                  double* Q) {
  const long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (r >= rows) return;
  double s0 = 0.0, s1 = 0.0;
  int chi = 0;
  for (; chi + 1 < naux; chi += 2) {
    s0 += X[chi * ldx + r] * T[chi * ldx + r];
    s1 += X[(chi + 1) * ldx + r] * T[(chi + 1) * ldx + r];
  }
  if (chi < naux) s0 += X[chi * ldx + r] * T[chi * ldx + r];
  Q[r] = s0 + s1;
}

// quadrature + tail part of request r (one CTA): out[r] = sum_n [ sum_j w_j/(2 pi) Re(den_j(n)) Q_j[l][n]
//                                                                 + tail(e_n - w) Q_order[l][n] ]
__global__ void __launch_bounds__(256)
cda_quad_kernel(const double* Q, long long node_stride, int npad, const int* slices, const double* freqs,
                const double* energies, int nt, int occ, int order, const double* pts, const double* wts,
                int symmetry, double alpha, double eta, double* out) {
  const int r = blockIdx.x;
  const double w = freqs[r];
  const double* Ql = Q + (long long)slices[r] * npad;
  double acc = 0.0;
  for (int n = threadIdx.x; n < nt; n += blockDim.x) {
    const double a = w - energies[n];
    const double b = n < occ ? eta : -eta;  // DeltaE.imag(): +eta on the occupied rows, -eta on the rest
    double s = 0.0;
    for (int j = 0; j < order; ++j) {
      const double bp = b + pts[j];
      double re = a / (a * a + bp * bp);  // Re 1/(dE + i w_j)
      if (symmetry) {
        const double bm = b - pts[j];
        re += a / (a * a + bm * bm);
      }
      s += wts[j] * re * Ql[j * node_stride + n];
    }
    s *= 0.15915494309189535;  // 0.5 / pi
    const double delta = -a;
    if (fabs(delta) > 1e-10) {
      const double ad = alpha * delta;
      s += 0.5 * copysign(1.0, delta) * exp(ad * ad) * erfc(fabs(ad)) * Ql[order * node_stride + n];
    }
    acc += s;
  }
  __shared__ double red[256];
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[r] = red[0];
}

// row n of the local slice: b[chi] = X[chi * ldx + slice * npad + n]
__global__ void cda_gather_row_kernel(const double* X, long long ldx, long long row, int naux, double* b, double* b0) {
  const int chi = blockIdx.x * blockDim.x + threadIdx.x;
  if (chi >= naux) return;
  const double v = X[chi * ldx + row];
  b[chi] = v;
  b0[chi] = v;
}
// out[r] += factor * (x - row) . row       (one CTA)
__global__ void __launch_bounds__(256) cda_residue_dot_kernel(const double* x, const double* row, int naux, double factor,
                                                              double* out) {
  double acc = 0.0;
  for (int i = threadIdx.x; i < naux; i += blockDim.x) acc += (x[i] - row[i]) * row[i];
  __shared__ double red[256];
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) *out += factor * red[0];
}

void check(gwbse_ctx* ctx, int rc) {
  if (rc != 0) throw std::runtime_error(ctx->err);
}

__global__ void cda_mean_kernel(double* a, const double* b, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) a[i] = 0.5 * (a[i] + b[i]);
}

// eps into ctx->eps: the channel's own matrix, or - with a partner channel registered - the spin-summed one of
// RPA_UKS (rpa_uks.cc:203-367; the restricted weights carry the closed-shell factor 2, so the sum is the mean).
// The two contexts take turns on the GPU: each stream is drained before the other one gets work.
void cda_epsilon(gwbse_ctx* ctx, int kind, double fre, double fim, double eta, const double* energies, int homo,
                 int rpamin, int rpamax) {
  check(ctx, gwbse_rpa_epsilon(ctx, kind, fre, fim, eta, energies, homo, rpamin, rpamax, nullptr, 0));
  gwbse_ctx* o = ctx->cda.partner;
  if (!o) return;
  GW_CUDA(cudaStreamSynchronize(ctx->stream));
  if (gwbse_rpa_epsilon(o, kind, fre, fim, eta, ctx->cda.partner_energies.data(), ctx->cda.partner_homo, rpamin, rpamax,
                        nullptr, 0))
    throw std::runtime_error(o->err);
  GW_CUDA(cudaSetDevice(ctx->device));
  GW_CUDA(cudaStreamSynchronize(o->stream));
  const size_t nn = (size_t)ctx->naux * ctx->naux;
  cda_mean_kernel<<<(unsigned)((nn + 255) / 256), 256, 0, ctx->stream>>>(ctx->eps, o->eps, nn);
  GW_CUDA(cudaGetLastError());
  ctx->launches++;
}

// sigma_cda.cc:62-77
double residue_prefactor(double e_f, double e_m, double frequency) {
  const double tolerance = 1e-10;
  if (e_f < e_m && e_m < frequency) return 1.0;
  if (e_f > e_m && e_m > frequency) return -1.0;
  if (std::abs(e_m - frequency) < tolerance && e_f > e_m) return -0.5;
  if (std::abs(e_m - frequency) < tolerance && e_f < e_m) return 0.5;
  return 0.0;
}

}  // namespace

extern "C" {

int gwbse_sigma_cda_prepare(gwbse_ctx* ctx, int order, const double* points, const double* weights, int symmetry,
                            double alpha, const double* energies, int homo, int rpamin, int rpamax, int qpmin,
                            int qpmax, double eta) {
  GW_API_BEGIN(ctx)
  GW_PROF(ctx, "sigma_cda_prepare");
  GW_REQUIRE(ctx->X != nullptr, "Mmn not allocated");
  mmn_complete_rotation(ctx);
  GW_REQUIRE(order > 0 && points && weights && energies, "invalid quadrature");
  GW_REQUIRE(rpamin == ctx->mmin && rpamin == ctx->nmin && rpamax == ctx->nmax, "RPA range must match Mmn");
  auto& st = ctx->cda;
  st.ready = false;
  const int n = ctx->naux;
  const size_t nn = (size_t)n * n;
  const int q = qpmax - qpmin + 1, qpoff = qpmin - rpamin;
  GW_REQUIRE(q > 0 && qpoff >= 0 && qpoff + q <= ctx->mtotal, "invalid qp window");
  st.order = order;
  st.symmetry = symmetry;
  st.alpha = alpha;
  st.eta = eta;
  st.homo = homo;
  st.rpamin = rpamin;
  st.rpamax = rpamax;
  st.qpmin = qpmin;
  st.q = q;
  st.pts.assign(points, points + order);
  st.wts.assign(weights, weights + order);
  double* kappa = ctx->buf("cda_kappa", nn * (size_t)(order + 1));
  double* kzero = kappa + nn * (size_t)order;
  // kappa_0 = eps(0)^-1 - 1
  cda_epsilon(ctx, 2, 0.0, 0.0, eta, energies, homo, rpamin, rpamax);
  GW_CUDA(cudaMemcpyAsync(kzero, ctx->eps, sizeof(double) * nn, cudaMemcpyDeviceToDevice, ctx->stream));
  check(ctx, gwbse_inverse_dev(ctx, n, kzero, n));
  cda_minus_identity_kernel<<<(n + 255) / 256, 256, 0, ctx->stream>>>(kzero, n);
  for (int j = 0; j < order; ++j) {
    double* k = kappa + nn * (size_t)j;
    cda_epsilon(ctx, 0, points[j], 0.0, eta, energies, homo, rpamin, rpamax);
    GW_CUDA(cudaMemcpyAsync(k, ctx->eps, sizeof(double) * nn, cudaMemcpyDeviceToDevice, ctx->stream));
    check(ctx, gwbse_inverse_dev(ctx, n, k, n));
    cda_kappa_kernel<<<dim3((n + 255) / 256, n), 256, 0, ctx->stream>>>(k, kzero, n,
                                                                          std::exp(-std::pow(alpha * points[j], 2)));
    GW_CUDA(cudaGetLastError());
  }
  ctx->launches += 1 + order;
  // Q_j for the local slices of the qp window: T = X kappa_j (MultiplyRight shape, into the second Mmn buffer),
  // then the row dots
  st.nq_local = ctx->owned_count(qpoff, q, ctx->rank);
  st.lfirst = st.nq_local ? ctx->local_index(ctx->first_owned(qpoff, ctx->rank)) : 0;
  st.node_stride = (long long)std::max(st.nq_local, 1) * ctx->npad;
  double* Q = ctx->buf("cda_Q", (size_t)st.node_stride * (order + 1));
  if (st.nq_local > 0) {
    double* T = mmn_scratch_x2(ctx);
    const long long rows = (long long)st.nq_local * ctx->npad, r0 = (long long)st.lfirst * ctx->npad;
    const int ldp = round_up(n, 2);
    double* Rp = ctx->buf("mulright_Rp", (size_t)ldp * n);
    for (int j = 0; j <= order; ++j) {
      GW_CUDA(cudaMemcpy2DAsync(Rp, sizeof(double) * ldp, kappa + nn * (size_t)j, sizeof(double) * n, sizeof(double) * n,
                                n, cudaMemcpyDeviceToDevice, ctx->stream));
      GemmParams p;
      p.M = (int)rows;
      p.N = n;
      p.Ki = n;
      p.A.ptr = ctx->X + r0;
      p.A.s_ri = 1;
      p.A.s_ki = ctx->ldx;
      p.B.ptr = Rp;
      p.B.s_ri = ldp;
      p.B.s_ki = 1;
      p.C = T + r0;
      p.sC_mi = 1;
      p.sC_ni = ctx->ldx;
      ctx->gemm(p);
      cda_rowdot_kernel<<<(unsigned)((rows + 255) / 256), 256, 0, ctx->stream>>>(ctx->X + r0, T + r0, ctx->ldx, n, rows,
                                                                               Q + (size_t)j * st.node_stride);
      GW_CUDA(cudaGetLastError());
      ctx->launches++;
    }
  }
  double* pw = ctx->buf("cda_pw", 2 * (size_t)order);
  GW_CUDA(cudaMemcpyAsync(pw, points, sizeof(double) * order, cudaMemcpyHostToDevice, ctx->stream));
  GW_CUDA(cudaMemcpyAsync(pw + order, weights, sizeof(double) * order, cudaMemcpyHostToDevice, ctx->stream));
  GW_CUDA(cudaStreamSynchronize(ctx->stream));
  st.mmn_version = ctx->mmn_version;
  st.ready = true;
  GW_API_END(ctx)
}

int gwbse_sigma_cda_set_partner(gwbse_ctx* ctx, gwbse_ctx* other, int homo_other, const double* energies_other) {
  GW_API_BEGIN(ctx)
  auto& st = ctx->cda;
  if (!other) {
    st.partner = nullptr;
    st.partner_energies.clear();
  } else {
    GW_REQUIRE(other != ctx && other->X != nullptr && energies_other, "the partner channel needs its own filled Mmn");
    GW_REQUIRE(ctx->world == 1 && other->world == 1, "the unrestricted path is single-GPU");
    GW_REQUIRE(other->device == ctx->device && other->naux == ctx->naux && other->ntotal == ctx->ntotal &&
                   other->mmin == ctx->mmin && other->nmin == ctx->nmin,
               "both channels live on one GPU with the same aux basis and level window");
    st.partner = other;
    st.partner_homo = homo_other;
    st.partner_energies.assign(energies_other, energies_other + other->ntotal);
  }
  GW_API_END(ctx)
}

int gwbse_sigma_cda_eval(gwbse_ctx* ctx, int nreq, const int* levels, const double* freqs, const double* energies,
                         double* sigma) {
  GW_API_BEGIN(ctx)
  GW_PROF(ctx, "sigma_cda_eval");
  auto& st = ctx->cda;
  mmn_complete_rotation(ctx);
  GW_REQUIRE(st.ready && st.mmn_version == ctx->mmn_version, "CDA screening not prepared (gwbse_sigma_cda_prepare)");
  if (nreq > 0) {
    const int n = ctx->naux, nt = ctx->ntotal, order = st.order;
    const int occ = st.homo + 1 - st.rpamin;
    const int qpoff = st.qpmin - st.rpamin;
    std::vector<int> slices(nreq);
    for (int r = 0; r < nreq; ++r) {
      GW_REQUIRE(levels[r] >= 0 && levels[r] < st.q, "level outside the qp window");
      const int m = qpoff + levels[r];
      GW_REQUIRE(ctx->owns(m), "CDA request for a level whose Mmn slice lives on another rank");
      slices[r] = ctx->local_index(m) - st.lfirst;
    }
    double* e_dev = ctx->buf("cda_e", nt);
    double* f_dev = ctx->buf("cda_f", nreq);
    double* out = ctx->buf("cda_out", nreq);
    int* s_dev = reinterpret_cast<int*>(ctx->buf("cda_slices", (size_t)nreq / 2 + 1));
    GW_CUDA(cudaMemcpyAsync(e_dev, energies, sizeof(double) * nt, cudaMemcpyHostToDevice, ctx->stream));
    GW_CUDA(cudaMemcpyAsync(f_dev, freqs, sizeof(double) * nreq, cudaMemcpyHostToDevice, ctx->stream));
    GW_CUDA(cudaMemcpyAsync(s_dev, slices.data(), sizeof(int) * nreq, cudaMemcpyHostToDevice, ctx->stream));
    const double* pw = ctx->buf("cda_pw", 2 * (size_t)order);
    cda_quad_kernel<<<nreq, 256, 0, ctx->stream>>>(ctx->buf("cda_Q", 0), st.node_stride, ctx->npad, s_dev, f_dev, e_dev,
                                                   nt, occ, order, pw, pw + order, st.symmetry, st.alpha, st.eta, out);
    GW_CUDA(cudaGetLastError());
    ctx->launches++;
    // residues: poles of G between the Fermi level and omega (sigma_cda.cc:81-110)
    const int homo = st.homo - st.rpamin;
    const double fermi = 0.5 * (energies[homo + 1] + energies[homo]);
    double* b = ctx->buf("cda_b", 2 * (size_t)n);
    for (int r = 0; r < nreq; ++r) {
      for (int i = 0; i < nt; ++i) {
        const double factor = residue_prefactor(fermi, energies[i], freqs[r]);
        if (std::abs(factor) <= 1e-10) continue;
        const double abs_delta = std::abs(energies[i] - freqs[r]);
        cda_epsilon(ctx, 2, abs_delta, st.eta, st.eta, energies, st.homo, st.rpamin, st.rpamax);
        const long long row = (long long)(slices[r] + st.lfirst) * ctx->npad + i;
        cda_gather_row_kernel<<<(n + 255) / 256, 256, 0, ctx->stream>>>(ctx->X, ctx->ldx, row, n, b, b + n);
        check(ctx, gwbse_lu_solve_dev(ctx, n, 1, ctx->eps, n, b, n));
        cda_residue_dot_kernel<<<1, 256, 0, ctx->stream>>>(b, b + n, n, factor, out + r);
        GW_CUDA(cudaGetLastError());
        ctx->launches += 2;
      }
    }
    GW_CUDA(cudaMemcpyAsync(sigma, out, sizeof(double) * nreq, cudaMemcpyDeviceToHost, ctx->stream));
    GW_CUDA(cudaStreamSynchronize(ctx->stream));
  }
  GW_API_END(ctx)
}

}  // extern "C"
