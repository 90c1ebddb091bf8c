// Streaming (HBM-bound) kernels of the GW-BSE path: coalesced, vectorised where
// alignment allows, warp-shuffle reductions.  SURVEY.md section 8a rows a7 (weights),
// a12/a14 (Sigma_c evaluation), a18 (BSE diagonal), a20 (Davidson corrections).
#include <algorithm>
#include <cmath>
#include <cstdint>

#include <cooperative_groups.h>

#include "context.cuh"

namespace gwbse {

namespace {

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// block-wide sum; result valid in thread 0.  blockDim.x multiple of 32, <= 1024.
__device__ __forceinline__ double block_sum(double v, double* sh) {
  v = warp_sum(v);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) sh[w] = v;
  __syncthreads();
  const int nw = blockDim.x >> 5;
  v = (threadIdx.x < nw) ? sh[threadIdx.x] : 0.0;
  if (w == 0) v = warp_sum(v);
  __syncthreads();
  return v;
}

__global__ void symmetrize_lower_kernel(double* A, int n, long long ld) {
  __shared__ double tile[32][33];
  const int tr = blockIdx.y, tc = blockIdx.x;  // destination tile (rows tr, cols tc), upper part: tc >= tr
  if (tc < tr) return;
  // source: lower tile rows tc*32.., cols tr*32..
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int r = tc * 32 + threadIdx.x, c = tr * 32 + j;
    tile[j][threadIdx.x] = (r < n && c < n) ? A[r + (long long)c * ld] : 0.0;
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int r = tr * 32 + threadIdx.x, c = tc * 32 + j;  // dest element (r, c) = source (c, r)
    if (r < n && c < n && c > r) A[r + (long long)c * ld] = tile[threadIdx.x][j];
  }
}

__global__ void add_diagonal_kernel(double* A, int n, long long ld, double v) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) A[i + (long long)i * ld] += v;
}

// RPA transition weights, rpa.cc:115-127 (imag / real) and rpa.cc:178-190 (complex)
__global__ void rpa_weights_kernel(double* w, long long ldw, const double* e, int kind, double fre, double fim,
                                   double eta, int n_occ, int n_unocc, int rank, int world) {
  const int ml = blockIdx.y;
  const int v = rank + ml * world;
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n_unocc || v >= n_occ) return;
  const double dE = e[n_occ + c] - e[v];
  double d;
  if (kind == 0) {
    d = 4.0 * dE / (dE * dE + fre * fre);
  } else if (kind == 1) {
    const double eta2 = eta * eta;
    const double dm = dE - fre, dp = dE + fre;
    d = 2.0 * (dm / (dm * dm + eta2) + dp / (dp * dp + eta2));
  } else {
    const double dm = fre - dE, dp = fre + dE;
    const double s1 = (fim + eta) * (fim + eta), s2 = (fim - eta) * (fim - eta);
    d = dm / (dm * dm + s1) - dp / (dp * dp + s2);
  }
  w[(long long)ml * ldw + c] = d;
}

__global__ void diag_scale_kernel(int side_right, int m, int n, const double* A, long long lda, const double* d,
                                  double* C, long long ldc) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int j = blockIdx.y;
  if (i < m && j < n) C[i + j * ldc] = A[i + j * lda] * (side_right ? d[j] : d[i]);
}

__global__ void axpy_kernel(int m, int n, double alpha, const double* X, long long ldx, double* Y, long long ldy) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int j = blockIdx.y;
  if (i < m && j < n) Y[i + j * ldy] += alpha * X[i + j * ldx];
}

__global__ void coldots_kernel(int m, const double* X, long long ldx, const double* Y, long long ldy, double* out,
                               int take_sqrt) {
  __shared__ double sh[32];
  const int j = blockIdx.x;
  const double* x = X + j * ldx;
  const double* y = Y + j * ldy;
  double s = 0.0;
  for (int i = threadIdx.x; i < m; i += blockDim.x) s += x[i] * y[i];
  s = block_sum(s, sh);
  if (threadIdx.x == 0) out[j] = take_sqrt ? sqrt(s) : s;
}

// Long columns (Davidson vectors: rows = BSE size, 20-30 columns): one CTA per column would leave most SMs idle, so a
// cluster of COLDOT_CL CTAs shares a column; the partial sums meet in rank 0 through distributed shared memory in a
// fixed order (no atomics, no scratch buffer: the result depends on m only).
constexpr int COLDOT_CL = 8;

__global__ void __cluster_dims__(COLDOT_CL, 1, 1) __launch_bounds__(256)
    coldots_cluster_kernel(int m, const double* X, long long ldx, const double* Y, long long ldy, double* out,
                           int take_sqrt) {
  namespace cg = cooperative_groups;
  cg::cluster_group cl = cg::this_cluster();
  __shared__ double sh[32];
  __shared__ double part;
  const int j = blockIdx.x / COLDOT_CL;
  const int r = (int)cl.block_rank();
  const double* x = X + j * ldx;
  const double* y = Y + j * ldy;
  const int chunk = ((m + COLDOT_CL - 1) / COLDOT_CL + 1) & ~1;  // even: chunk starts keep the column's alignment
  const int i0 = min(m, r * chunk), i1 = min(m, i0 + chunk);
  double s0 = 0.0, s1 = 0.0;
  if ((((uintptr_t)(x + i0) | (uintptr_t)(y + i0)) & 15) == 0) {
    const int n2 = (i1 - i0) >> 1;
    const double2* x2 = reinterpret_cast<const double2*>(x + i0);
    const double2* y2 = reinterpret_cast<const double2*>(y + i0);
    for (int i = threadIdx.x; i < n2; i += blockDim.x) {
      const double2 a = x2[i], b = y2[i];
      s0 += a.x * b.x;
      s1 += a.y * b.y;
    }
    if (threadIdx.x == 0 && ((i1 - i0) & 1)) s0 += x[i1 - 1] * y[i1 - 1];
  } else {
    for (int i = i0 + threadIdx.x; i < i1; i += blockDim.x) s0 += x[i] * y[i];
  }
  const double s = block_sum(s0 + s1, sh);
  if (threadIdx.x == 0) part = s;
  cl.sync();
  if (r == 0 && threadIdx.x == 0) {
    double t = 0.0;
    for (int b = 0; b < COLDOT_CL; ++b) t += *cl.map_shared_rank(&part, b);
    out[j] = take_sqrt ? sqrt(t) : t;
  }
  cl.sync();  // no CTA may exit while rank 0 still reads its shared memory
}

__global__ void scale_cols_kernel(int m, int n, double* A, long long lda, const double* s) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int j = blockIdx.y;
  if (i < m && j < n) A[i + j * lda] *= s[j];
}

__global__ void copy_block_kernel(int m, int n, const double* A, long long lda, double* B, long long ldb) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int j = blockIdx.y;
  if (i < m && j < n) B[i + j * ldb] = A[i + j * lda];
}

// Compact, optionally pole-scaled copy of a two-index block of the three-centre tensor (the operands of the
// materialised BSE blocks): out[p * plane + a * L2 + b] = scale[p] * src[p * s_pole + a * s_outer + b].
__global__ void pack_block_kernel(const double* __restrict__ src, long long s_pole, long long s_outer, int L1, int L2,
                                  const double* __restrict__ scale, double* __restrict__ out, long long plane,
                                  int npoles) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= L1 * L2) return;
  const int a = idx / L2, b = idx - a * L2;
  for (int p = blockIdx.y; p < npoles; p += gridDim.y) {
    const double f = scale ? scale[p] : 1.0;
    out[p * plane + idx] = f * src[p * s_pole + a * s_outer + b];
  }
}

__global__ void invsqrt_scale_kernel(double* out, const double* w, int n, double etol, int* removed) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (w[i] < etol) {
    out[i] = 0.0;
    atomicAdd(removed, 1);
  } else {
    out[i] = 1.0 / sqrt(w[i]);
  }
}

// ---------------------------------------------------------------------------
// Sigma_c diagonal element (and derivative) for GROUPS of frequencies per level.
//   sigma(w) = pref * sum_p fac_p sum_n M[n,p]^2 t/(t^2+eta^2),  t = w - e_n +- pole_p
//   dsigma/dw = pref * sum_p fac_p sum_n M[n,p]^2 (eta^2 - t^2)/(t^2+eta^2)^2
// sigma_ppm.cc:37-91 (fac = w*Omega, pref 1/2), sigma_exact.cc:40-83 (fac 1, pref 2).
// One CTA = (group, chunk of SIG_PC poles); the level's M columns are read from HBM once and re-used
// from L1 for every block of 8 frequencies of the group.  The kernel is FP64-arithmetic bound
// (12 FP64 instructions per element and frequency, one of them a reciprocal), not HBM bound, as soon
// as a group holds more than ~2 frequencies.  Partial sums are per (frequency, chunk) in a fixed
// order, so a result does not depend on how requests were batched.
// ---------------------------------------------------------------------------
constexpr int SIG_PC = 8;   // poles per CTA
constexpr int SIG_FB = 8;   // frequencies per register block

__device__ __forceinline__ double rcp_pos(double x) {
  // x = t^2 + eta^2 is positive, finite and normal: MUFU seed + 2 Newton steps (no special-case path)
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  double e = fma(-x, r, 1.0);
  r = fma(r, e, r);
  e = fma(-x, r, 1.0);
  r = fma(r, e, r);
  return r;
}

template <bool DERIV>
__global__ void __launch_bounds__(256) sigma_multi_kernel(const double* __restrict__ mat, long long ld,
                                                          long long lstride, int qpoff, int ntotal, int npoles,
                                                          int boundary, double eta2,
                                                          const double* __restrict__ fac,
                                                          const double* __restrict__ pole,
                                                          const double* __restrict__ energies,
                                                          const int* __restrict__ levels,
                                                          const int* __restrict__ gptr,
                                                          const double* __restrict__ freqs, double* partial,
                                                          int nchunks) {
  __shared__ double sh[2 * SIG_FB][8];
  const int g = blockIdx.y, chunk = blockIdx.x;
  const int level = levels[g];
  const int f_begin = gptr[g], nf = gptr[g + 1] - f_begin;
  const int p0 = chunk * SIG_PC;
  double om[SIG_PC], fc[SIG_PC];
  const double* cols[SIG_PC];
  bool any = false;
#pragma unroll
  for (int p = 0; p < SIG_PC; ++p) {
    const int pp = min(p0 + p, npoles - 1);
    om[p] = pole[pp];
    fc[p] = (p0 + p < npoles) ? fac[pp] : 0.0;
    any = any || fc[p] != 0.0;
    cols[p] = mat + (long long)pp * ld + (long long)(qpoff + level) * lstride;  // level = local slice index
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const double two_eta2 = 2.0 * eta2;
  for (int fb = 0; fb < nf; fb += SIG_FB) {
    double wf[SIG_FB], as[SIG_FB], ad[SIG_FB];
#pragma unroll
    for (int f = 0; f < SIG_FB; ++f) {
      wf[f] = freqs[f_begin + min(fb + f, nf - 1)];
      as[f] = 0.0;
      ad[f] = 0.0;
    }
    if (any) {
      for (int n = threadIdx.x; n < ntotal; n += 256) {
        const double e = energies[n];
        const bool occ = n < boundary;
#pragma unroll
        for (int p = 0; p < SIG_PC; ++p) {
          const double m = cols[p][n];
          const double m2 = fc[p] * m * m;
          const double a = occ ? e - om[p] : e + om[p];  // t = w - a
#pragma unroll
          for (int f = 0; f < SIG_FB; ++f) {
            const double t = wf[f] - a;
            const double r = rcp_pos(fma(t, t, eta2));
            as[f] = fma(m2, t * r, as[f]);
            if (DERIV) ad[f] = fma(m2, r * fma(two_eta2, r, -1.0), ad[f]);
          }
        }
      }
    }
#pragma unroll
    for (int f = 0; f < SIG_FB; ++f) {
      as[f] = warp_sum(as[f]);
      if (DERIV) ad[f] = warp_sum(ad[f]);
    }
    if (lane == 0) {
#pragma unroll
      for (int f = 0; f < SIG_FB; ++f) {
        sh[f][warp] = as[f];
        if (DERIV) sh[SIG_FB + f][warp] = ad[f];
      }
    }
    __syncthreads();
    if (threadIdx.x < SIG_FB * (DERIV ? 2 : 1)) {
      const int f = threadIdx.x % SIG_FB, isd = threadIdx.x / SIG_FB;
      if (fb + f < nf) {
        double v = 0.0;
#pragma unroll
        for (int w = 0; w < 8; ++w) v += sh[threadIdx.x][w];
        partial[((long long)(f_begin + fb + f) * nchunks + chunk) * 2 + isd] = v;
      }
    }
    __syncthreads();
  }
}

__global__ void sigma_multi_reduce_kernel(const double* partial, int nfreq, int nchunks, double pref, int deriv,
                                          double* out) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= nfreq) return;
  double s = 0.0, ds = 0.0;
  for (int k = 0; k < nchunks; ++k) {
    s += partial[((long long)r * nchunks + k) * 2 + 0];
    if (deriv) ds += partial[((long long)r * nchunks + k) * 2 + 1];
  }
  out[r] = pref * s;
  out[nfreq + r] = pref * ds;
}

// A_il[n, p] = pref * fac_p * M[n, p] * t/(t^2+eta^2), t = w - e_n +- pole_p  (off-diagonal Sigma_c as a GEMM);
// il runs over the levels this rank owns: slice_idx[il] = local slice, freq_idx[il] = index into freqs
__global__ void sigma_offdiag_weight_kernel(const double* __restrict__ mat, long long ld, long long lstride,
                                            const int* __restrict__ slice_idx, const int* __restrict__ freq_idx,
                                            int ntotal, int npad, int boundary, double eta2, double pref,
                                            const double* __restrict__ fac, const double* __restrict__ pole,
                                            const double* __restrict__ energies, const double* __restrict__ freqs,
                                            int p0, double* out, long long ldo) {
  const int il = blockIdx.y;
  const int p = p0 + blockIdx.z;
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= npad) return;
  double v = 0.0;
  if (n < ntotal) {
    const double om = pole[p];
    const double t = freqs[freq_idx[il]] - energies[n] + (n < boundary ? om : -om);
    const double m = mat[(long long)p * ld + (long long)slice_idx[il] * lstride + n];
    v = pref * fac[p] * m * t / (t * t + eta2);
  }
  out[(long long)blockIdx.z * ldo + (long long)il * npad + n] = v;
}

__global__ void offdiag_finish_kernel(const double* S, int q, double* out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int j = blockIdx.y;
  if (i >= q || j >= q) return;
  out[i + (long long)j * q] = (i == j) ? 0.0 : S[i + (long long)j * q] + S[j + (long long)i * q];
}

// Diagonal elements M[s][s, chi] of the slices [s0, s0+ns) this rank owns -> D[(s - s0) + ns * chi] (others untouched)
__global__ void slice_diag_kernel(const double* __restrict__ X, long long ldx, int npad, int naux, int s0, int ns,
                                  int rank, int world, double* D) {
  const int chi = blockIdx.x * blockDim.x + threadIdx.x;
  const int sl = blockIdx.y;
  const int s = s0 + sl;
  if (chi >= naux || s % world != rank) return;
  D[sl + (long long)ns * chi] = X[(long long)chi * ldx + (long long)(s / world) * npad + s];
}

// BSE_OPERATOR::diagonal, bse_operator.cc:134-175 for the occupied levels v this rank owns; thread = (c, local v).
// cv(c, v, chi) = M[coff+c][voff+v, chi] comes from X (single GPU) or from the replicated block.
__global__ void bse_diag_kernel(const double* __restrict__ X, long long ldx, int npad, int naux, int vt, int ct,
                                int voff, int coff, int v_rel0, int vstride, int lfirst,
                                const double* __restrict__ cv, long long cv_row, long long cv_pole,
                                const double* __restrict__ Dcc, const double* __restrict__ Dvv,
                                const double* __restrict__ eps_inv, const double* __restrict__ hqp, int ldh, int cqp,
                                int cx, int cd, int cd2, double* out) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  const int l = blockIdx.y;
  const int v = v_rel0 + l * vstride;
  if (c >= ct || v >= vt) return;
  double entry = 0.0;
  const double* Mv = X + (long long)(lfirst + l) * npad;  // local slice of v
  if (cx != 0) {
    double s = 0.0;
    for (int p = 0; p < naux; ++p) {
      const double m = Mv[(long long)p * ldx + coff + c];
      s += m * m;
    }
    entry += cx * s;
  }
  if (cqp != 0) entry += cqp * (hqp[(c + vt) + (long long)(c + vt) * ldh] - hqp[v + (long long)v * ldh]);
  if (cd != 0) {
    double s = 0.0;
    for (int p = 0; p < naux; ++p) s += Dcc[c + (long long)ct * p] * eps_inv[p] * Dvv[v + (long long)vt * p];
    entry -= cd * s;
  }
  if (cd2 != 0) {
    double s = 0.0;
    for (int p = 0; p < naux; ++p)
      s += cv[(long long)p * cv_pole + (long long)c * cv_row + v] * eps_inv[p] * Mv[(long long)p * ldx + coff + c];
    entry -= cd2 * s;
  }
  out[(long long)ct * v + c] = entry;
}

// DavidsonSolver::dpr, davidsonsolver.cc:416-422 (+ the isfinite filter of :411-413)
__global__ void dpr_kernel(int rows, const double* diag, const double* lambda, const double* R, long long ldr,
                           double* W, long long ldw) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int j = blockIdx.y;
  if (i >= rows) return;
  const double v = -R[i + j * ldr] / (diag[i] - lambda[j]);
  W[i + j * ldw] = isfinite(v) ? v : 0.0;
}

// DavidsonSolver::olsen, davidsonsolver.cc:424-440: W holds dpr(r); W += (x.dpr(r) / x.dpr(x)) * x
__global__ void olsen_finish_kernel(int rows, const double* Q, long long ldq, const double* num, const double* den,
                                    double* W, long long ldw) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int j = blockIdx.y;
  if (i >= rows) return;
  const double eps = num[j] / den[j];
  const double v = W[i + j * ldw] + eps * Q[i + j * ldq];
  W[i + j * ldw] = isfinite(v) ? v : 0.0;
}

constexpr int kColdotClusterRows = 8192;  // below this one CTA per column is latency-optimal

inline dim3 grid2(int m, int n, int bx = 256) { return dim3((m + bx - 1) / bx, n); }

}  // namespace

void launch_symmetrize_lower(double* A, int n, long long ld, cudaStream_t s) {
  const int t = (n + 31) / 32;
  symmetrize_lower_kernel<<<dim3(t, t), dim3(32, 8), 0, s>>>(A, n, ld);
  GW_CUDA(cudaGetLastError());
}
void launch_add_diagonal(double* A, int n, long long ld, double v, cudaStream_t s) {
  add_diagonal_kernel<<<(n + 255) / 256, 256, 0, s>>>(A, n, ld, v);
  GW_CUDA(cudaGetLastError());
}
void launch_rpa_weights(double* w, long long ldw, const double* e, int kind, double fre, double fim, double eta,
                        int n_occ, int n_unocc, int rank, int world, int nloc_occ, cudaStream_t s) {
  if (nloc_occ <= 0) return;
  rpa_weights_kernel<<<dim3((n_unocc + 255) / 256, nloc_occ), 256, 0, s>>>(w, ldw, e, kind, fre, fim, eta, n_occ,
                                                                            n_unocc, rank, world);
  GW_CUDA(cudaGetLastError());
}
void launch_diag_scale(char side, int m, int n, const double* A, long long lda, const double* d, double* C,
                       long long ldc, cudaStream_t s) {
  if (m <= 0 || n <= 0) return;
  diag_scale_kernel<<<grid2(m, n), 256, 0, s>>>(side == 'R' || side == 'r', m, n, A, lda, d, C, ldc);
  GW_CUDA(cudaGetLastError());
}
void launch_axpy(int m, int n, double alpha, const double* X, long long ldx, double* Y, long long ldy,
                 cudaStream_t s) {
  if (m <= 0 || n <= 0) return;
  axpy_kernel<<<grid2(m, n), 256, 0, s>>>(m, n, alpha, X, ldx, Y, ldy);
  GW_CUDA(cudaGetLastError());
}
void launch_colnorms(int m, int n, const double* A, long long lda, double* out_dev, cudaStream_t s) {
  if (n <= 0) return;
  if (m >= kColdotClusterRows)
    coldots_cluster_kernel<<<n * COLDOT_CL, 256, 0, s>>>(m, A, lda, A, lda, out_dev, 1);
  else
    coldots_kernel<<<n, 256, 0, s>>>(m, A, lda, A, lda, out_dev, 1);
  GW_CUDA(cudaGetLastError());
}
void launch_coldots(int m, int n, const double* X, long long ldx, const double* Y, long long ldy, double* out_dev,
                    cudaStream_t s) {
  if (n <= 0) return;
  if (m >= kColdotClusterRows)
    coldots_cluster_kernel<<<n * COLDOT_CL, 256, 0, s>>>(m, X, ldx, Y, ldy, out_dev, 0);
  else
    coldots_kernel<<<n, 256, 0, s>>>(m, X, ldx, Y, ldy, out_dev, 0);
  GW_CUDA(cudaGetLastError());
}
void launch_scale_cols(int m, int n, double* A, long long lda, const double* s_dev, cudaStream_t s) {
  if (m <= 0 || n <= 0) return;
  scale_cols_kernel<<<grid2(m, n), 256, 0, s>>>(m, n, A, lda, s_dev);
  GW_CUDA(cudaGetLastError());
}
// X[chi * ldx + (lfirst + il) * npad + row0 + n] = T[n + q * (il + nloc * chi)]
__global__ void rotate_scatter_kernel(const double* __restrict__ T, int q, int nloc, double* __restrict__ X,
                                      long long ldx, int npad, int lfirst, int row0) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  const int il = blockIdx.y, chi = blockIdx.z;
  if (n >= q) return;
  X[(long long)chi * ldx + (long long)(lfirst + il) * npad + row0 + n] = T[n + (long long)q * (il + (long long)nloc * chi)];
}
void launch_rotate_scatter(const double* T, int q, int nloc, int naux, double* X, long long ldx, int npad, int lfirst,
                           int row0, cudaStream_t s) {
  if (q <= 0 || nloc <= 0 || naux <= 0) return;
  dim3 grid((q + 127) / 128, nloc, naux);
  rotate_scatter_kernel<<<grid, 128, 0, s>>>(T, q, nloc, X, ldx, npad, lfirst, row0);
  GW_CUDA(cudaGetLastError());
}
void launch_copy_block(int m, int n, const double* A, long long lda, double* B, long long ldb, cudaStream_t s) {
  if (m <= 0 || n <= 0) return;
  copy_block_kernel<<<grid2(m, n), 256, 0, s>>>(m, n, A, lda, B, ldb);
  GW_CUDA(cudaGetLastError());
}
void launch_pack_block(const double* src, long long s_pole, long long s_outer, int L1, int L2, const double* scale,
                       double* out, long long plane, int npoles, cudaStream_t s) {
  if (L1 <= 0 || L2 <= 0 || npoles <= 0) return;
  GW_REQUIRE((long long)L1 * L2 < (1LL << 31), "block too large to pack");
  dim3 grid((unsigned)(((long long)L1 * L2 + 255) / 256), (unsigned)std::min(npoles, 65535));
  pack_block_kernel<<<grid, 256, 0, s>>>(src, s_pole, s_outer, L1, L2, scale, out, plane, npoles);
  GW_CUDA(cudaGetLastError());
}
void launch_invsqrt_scale(double* out, const double* w, int n, double etol, int* removed_dev, cudaStream_t s) {
  invsqrt_scale_kernel<<<(n + 255) / 256, 256, 0, s>>>(out, w, n, etol, removed_dev);
  GW_CUDA(cudaGetLastError());
}
int sigma_multi_chunks(int npoles) { return (npoles + SIG_PC - 1) / SIG_PC; }

void launch_sigma_multi(const gwbse_ctx::SigmaState& st, int ntotal, int ngroups, int nfreq, const int* levels_dev,
                        const int* gptr_dev, const double* freqs_dev, double* partial_dev, double* out_dev,
                        bool want_deriv, cudaStream_t s) {
  if (ngroups <= 0 || nfreq <= 0) return;
  const int nchunks = sigma_multi_chunks(st.npoles);
  dim3 grid(nchunks, ngroups);
  if (want_deriv)
    sigma_multi_kernel<true><<<grid, 256, 0, s>>>(st.mat, st.ld, st.lstride, 0, ntotal, st.npoles,
                                                  st.nocc_boundary, st.eta * st.eta, st.fac, st.pole, st.energies,
                                                  levels_dev, gptr_dev, freqs_dev, partial_dev, nchunks);
  else
    sigma_multi_kernel<false><<<grid, 256, 0, s>>>(st.mat, st.ld, st.lstride, 0, ntotal, st.npoles,
                                                   st.nocc_boundary, st.eta * st.eta, st.fac, st.pole, st.energies,
                                                   levels_dev, gptr_dev, freqs_dev, partial_dev, nchunks);
  GW_CUDA(cudaGetLastError());
  sigma_multi_reduce_kernel<<<(nfreq + 127) / 128, 128, 0, s>>>(partial_dev, nfreq, nchunks, st.diag_pref,
                                                                want_deriv ? 1 : 0, out_dev);
  GW_CUDA(cudaGetLastError());
}
void launch_sigma_offdiag_weight(const gwbse_ctx::SigmaState& st, int ntotal, int npad, int nlevels,
                                 const int* slice_idx_dev, const int* freq_idx_dev, int p0, int np,
                                 const double* freqs_dev, double pref, double* out, long long ldo, cudaStream_t s) {
  if (np <= 0 || nlevels <= 0) return;
  sigma_offdiag_weight_kernel<<<dim3((npad + 127) / 128, nlevels, np), 128, 0, s>>>(
      st.mat, st.ld, st.lstride, slice_idx_dev, freq_idx_dev, ntotal, npad, st.nocc_boundary, st.eta * st.eta, pref,
      st.fac, st.pole, st.energies, freqs_dev, p0, out, ldo);
  GW_CUDA(cudaGetLastError());
}
void launch_offdiag_finish(const double* S, int q, double* out, cudaStream_t s) {
  offdiag_finish_kernel<<<grid2(q, q, 128), 128, 0, s>>>(S, q, out);
  GW_CUDA(cudaGetLastError());
}
void launch_slice_diag(const double* X, long long ldx, int npad, int naux, int s0, int ns, int rank, int world,
                       double* D, cudaStream_t s) {
  if (ns <= 0) return;
  slice_diag_kernel<<<dim3((naux + 127) / 128, ns), 128, 0, s>>>(X, ldx, npad, naux, s0, ns, rank, world, D);
  GW_CUDA(cudaGetLastError());
}
void launch_bse_diag(const double* X, long long ldx, int npad, int naux, int vt, int ct, int voff, int coff,
                     int v_rel0, int vstride, int lfirst, int nvloc, const double* cv, long long cv_row,
                     long long cv_pole, const double* Dcc, const double* Dvv, const double* eps_inv,
                     const double* hqp, int ldh, int cqp, int cx, int cd, int cd2, double* out, cudaStream_t s) {
  if (nvloc <= 0) return;
  bse_diag_kernel<<<dim3((ct + 127) / 128, nvloc), 128, 0, s>>>(X, ldx, npad, naux, vt, ct, voff, coff, v_rel0,
                                                                  vstride, lfirst, cv, cv_row, cv_pole, Dcc, Dvv,
                                                                  eps_inv, hqp, ldh, cqp, cx, cd, cd2, out);
  GW_CUDA(cudaGetLastError());
}
void launch_dpr(int rows, int ncols, const double* diag, const double* lambda_dev, const double* R, long long ldr,
                double* W, long long ldw, cudaStream_t s) {
  if (ncols <= 0) return;
  dpr_kernel<<<grid2(rows, ncols), 256, 0, s>>>(rows, diag, lambda_dev, R, ldr, W, ldw);
  GW_CUDA(cudaGetLastError());
}
void launch_olsen_finish(int rows, int ncols, const double*, const double*, const double* Q, long long ldq,
                         const double* num_dev, const double* den_dev, double* W, long long ldw, cudaStream_t s) {
  if (ncols <= 0) return;
  olsen_finish_kernel<<<grid2(rows, ncols), 256, 0, s>>>(rows, Q, ldq, num_dev, den_dev, W, ldw);
  GW_CUDA(cudaGetLastError());
}

}  // namespace gwbse
