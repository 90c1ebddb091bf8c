// Sigma_c diagonal element by a 1-D treecode over the pole positions.
//
//   sigma_i(w) = pref * sum_{n,p} r_i[n,p] t/(t^2+eta^2),  t = w - a[n,p],
//   a[n,p] = e_n - pole_p (n below the occupied boundary) or e_n + pole_p,  r_i[n,p] = fac_p M_i[n,p]^2
//   (sigma_ppm.cc:37-91, sigma_exact.cc:40-83 evaluate this sum term by term for every frequency of the
//   QP root search, gw.cc:344 / qp_solver_utils.h).
//
// The pole positions a[n,p] are the same for every level i; only the residues differ.  They are sorted once
// per screening update, cut into leaves of `leaf` consecutive terms, and a 4-ary tree is put on top.  For
// each level the tree nodes carry TREE_P normalised moments  mu_j = sum_k r_k ((a_k - c)/rho)^j  (c, rho =
// midpoint / half width of the node), built from the leaves upwards (leaf sums, then moment-to-moment shifts).
// A frequency then walks the tree: a node with |w - c| >= 4 rho is summed through
//   Re sum_k r_k/(z - d_k) = Re[ u sum_j mu_j (rho u)^j ],  z = w - c - i eta, u = 1/z, |rho u| <= 1/4,
// (truncation 4^-24 ~ 4e-15 of sum|r|/|z|), any other node is opened, and the few leaves next to w are summed
// term by term with the reference formula.  Work per frequency drops from n*npoles terms to O(100) node
// expansions, so the QP search no longer depends on the FP64 pipe.  Summation order is fixed by the tree:
// a result does not depend on how requests are batched.
#include <cub/cub.cuh>

#include <algorithm>
#include <cstring>
#include <vector>

#include "context.cuh"

namespace gwbse {

namespace {

constexpr int TREE_P = 24;
constexpr double TREE_OPEN = 4.0;  // a node is expanded only if |w - c| >= TREE_OPEN * rho
constexpr int TREE_STACK = 2048;   // per-warp traversal stack (entries)
constexpr int TREE_WARPS = 4;
constexpr double TREE_RHO_MIN = 1e-300;

struct TreeGeom {
  long long T;   // number of terms
  long long NT;  // nodes stored per level slot
  int leaf, D, d0;
  long long span[16];   // terms per node at depth d
  long long count[16];  // nodes at depth d
  long long off[16];    // first node of depth d within a slot (depths d0..D)
};

__device__ __forceinline__ void node_geom(const TreeGeom& g, const double* __restrict__ a, int d, long long j,
                                          long long& first, long long& last, double& c, double& rho) {
  first = j * g.span[d];
  last = min(first + g.span[d], g.T) - 1;
  const double lo = a[first], hi = a[last];
  c = 0.5 * (lo + hi);
  rho = 0.5 * (hi - lo);
}

__global__ void tree_keys_kernel(long long T, int ntotal, int boundary, const double* __restrict__ energies,
                                 const double* __restrict__ pole, const double* __restrict__ fac,
                                 double* keys, unsigned int* vals) {
  const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= T) return;
  const int p = (int)(k / ntotal), n = (int)(k - (long long)p * ntotal);
  const double e = energies[n];
  // poles with zero prefactor are skipped by the reference (sigma_ppm.cc:47-52): park them on e_n
  const double om = fac[p] != 0.0 ? pole[p] : 0.0;
  keys[k] = n < boundary ? e - om : e + om;
  vals[k] = (unsigned int)k;
}

__global__ void tree_terms_kernel(long long T, int ntotal, const unsigned int* __restrict__ vals, int2* term) {
  const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= T) return;
  const unsigned int v = vals[k];
  const int p = (int)(v / (unsigned int)ntotal);
  term[k] = make_int2((int)(v - (unsigned int)p * (unsigned int)ntotal), p);
}

// one thread per (leaf, level being built): mu_j = sum_k r_k x_k^j
__global__ void __launch_bounds__(128) tree_leaf_moments_kernel(TreeGeom g, const double* __restrict__ a,
                                                                const int2* __restrict__ term,
                                                                const double* __restrict__ mat, long long ld,
                                                                long long lstride,
                                                                const double* __restrict__ fac,
                                                                const int* __restrict__ build_slice,
                                                                const int* __restrict__ build_slot, double* mom) {
  const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= g.count[g.D]) return;
  const int slice = build_slice[blockIdx.y], slot = build_slot[blockIdx.y];
  long long first, last;
  double c, rho;
  node_geom(g, a, g.D, j, first, last, c, rho);
  const double inv = 1.0 / fmax(rho, TREE_RHO_MIN);
  double mu[TREE_P];
#pragma unroll
  for (int o = 0; o < TREE_P; ++o) mu[o] = 0.0;
  const double* base = mat + (long long)slice * lstride;
  for (long long k = first; k <= last; ++k) {
    const int2 t = term[k];
    const double m = base[(long long)t.y * ld + t.x];
    double pw = fac[t.y] * m * m;
    const double x = (a[k] - c) * inv;
#pragma unroll
    for (int o = 0; o < TREE_P; ++o) {
      mu[o] += pw;
      pw *= x;
    }
  }
  double* out = mom + ((long long)slot * g.NT + g.off[g.D] + j) * TREE_P;
#pragma unroll
  for (int o = 0; o < TREE_P; ++o) out[o] = mu[o];
}

// one thread per (parent node at depth d, level): children rescaled to the parent radius and shifted to its centre
__global__ void __launch_bounds__(128) tree_m2m_kernel(TreeGeom g, int d, const double* __restrict__ a,
                                                       const int* __restrict__ build_slot, double* mom) {
  const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= g.count[d]) return;
  const int slot = build_slot[blockIdx.y];
  long long first, last;
  double cp, rp;
  node_geom(g, a, d, j, first, last, cp, rp);
  const double invp = 1.0 / fmax(rp, TREE_RHO_MIN);
  double acc[TREE_P];
#pragma unroll
  for (int o = 0; o < TREE_P; ++o) acc[o] = 0.0;
  for (long long ch = 4 * j; ch < min(4 * j + 4, g.count[d + 1]); ++ch) {
    double cc, rc;
    node_geom(g, a, d + 1, ch, first, last, cc, rc);
    const double* in = mom + ((long long)slot * g.NT + g.off[d + 1] + ch) * TREE_P;
    double nu[TREE_P];
    const double ratio = rc * invp, s = (cc - cp) * invp;
    double sc = 1.0;
#pragma unroll
    for (int o = 0; o < TREE_P; ++o) {
      nu[o] = in[o] * sc;
      sc *= ratio;
    }
    // binomial shift  nu'_j = sum_m C(j,m) s^(j-m) nu_m  by the triangular recurrence
#pragma unroll
    for (int t = 1; t < TREE_P; ++t) {
#pragma unroll
      for (int jj = TREE_P - 1; jj >= t; --jj) nu[jj] = fma(s, nu[jj - 1], nu[jj]);
    }
#pragma unroll
    for (int o = 0; o < TREE_P; ++o) acc[o] += nu[o];
  }
  double* out = mom + ((long long)slot * g.NT + g.off[d] + j) * TREE_P;
#pragma unroll
  for (int o = 0; o < TREE_P; ++o) out[o] = acc[o];
}

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// one warp per requested (level, frequency)
template <bool DERIV>
__global__ void __launch_bounds__(TREE_WARPS * 32) tree_eval_kernel(
    TreeGeom g, const double* __restrict__ a, const int2* __restrict__ term, const double* __restrict__ mat,
    long long ld, long long lstride, const double* __restrict__ fac, const double* __restrict__ mom, int nfreq,
    const double* __restrict__ freqs, const int* __restrict__ fslice, const int* __restrict__ fslot, double eta,
    double pref, double* out, int* overflow) {
  __shared__ int stack_all[TREE_WARPS][TREE_STACK];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int target = blockIdx.x * TREE_WARPS + warp;
  if (target >= nfreq) return;
  int* stack = stack_all[warp];
  const double w = freqs[target];
  const int slot = fslot[target];
  const double* base = mat + (long long)fslice[target] * lstride;
  const double* mu_base = mom + (long long)slot * g.NT * TREE_P;
  const double eta2 = eta * eta;
  int sp = (int)g.count[g.d0];
  for (int i = lane; i < sp; i += 32) stack[i] = (g.d0 << 27) | i;
  __syncwarp();
  double acc = 0.0, dacc = 0.0;
  while (sp > 0) {
    const int take = min(32, sp);
    sp -= take;
    const int entry = lane < take ? stack[sp + lane] : -1;
    __syncwarp();
    bool open = false, leaf_near = false;
    int d = 0;
    long long j = 0;
    if (entry >= 0) {
      d = entry >> 27;
      j = entry & ((1 << 27) - 1);
      long long first, last;
      double c, rho;
      node_geom(g, a, d, j, first, last, c, rho);
      const double tr = w - c, dist = fabs(tr);
      if (dist >= TREE_OPEN * rho && dist > 0.0) {
        const double den = 1.0 / fma(tr, tr, eta2);
        const double ur = tr * den, ui = eta * den;  // u = 1/(tr - i eta)
        const double re = fmax(rho, TREE_RHO_MIN);
        const double qr = re * ur, qi = re * ui;
        const double* mu = mu_base + (g.off[d] + j) * TREE_P;
        double ar = mu[TREE_P - 1], ai = 0.0;
        double br = TREE_P * ar, bi = 0.0;
#pragma unroll
        for (int o = TREE_P - 2; o >= 0; --o) {
          const double m = mu[o];
          const double nr = fma(ar, qr, fma(-ai, qi, m));
          ai = fma(ar, qi, ai * qr);
          ar = nr;
          if (DERIV) {
            const double mr = fma(br, qr, fma(-bi, qi, (o + 1) * m));
            bi = fma(br, qi, bi * qr);
            br = mr;
          }
        }
        acc += ar * ur - ai * ui;
        if (DERIV) {
          const double u2r = ur * ur - ui * ui, u2i = 2.0 * ur * ui;
          dacc -= br * u2r - bi * u2i;
        }
      } else if (d == g.D) {
        leaf_near = true;
      } else {
        open = true;
      }
    }
    // push the children of opened nodes (missing children as -1)
    const unsigned omask = __ballot_sync(0xffffffffu, open);
    const int npush = 4 * __popc(omask);
    if (sp + npush > TREE_STACK) {
      if (lane == 0) atomicExch(overflow, 1);
      break;
    }
    if (open) {
      const int pos = sp + 4 * __popc(omask & ((1u << lane) - 1u));
#pragma unroll
      for (int ch = 0; ch < 4; ++ch) {
        const long long cj = 4 * j + ch;
        stack[pos + ch] = cj < g.count[d + 1] ? (((d + 1) << 27) | (int)cj) : -1;
      }
    }
    sp += npush;
    __syncwarp();
    // leaves next to w: term by term, the whole warp on one leaf at a time
    unsigned lmask = __ballot_sync(0xffffffffu, leaf_near);
    while (lmask) {
      const int src = __ffs(lmask) - 1;
      lmask &= lmask - 1;
      const long long lj = __shfl_sync(0xffffffffu, j, src);
      const long long first = lj * g.span[g.D], last = min(first + g.span[g.D], g.T) - 1;
      for (long long k = first + lane; k <= last; k += 32) {
        const int2 t = term[k];
        const double m = base[(long long)t.y * ld + t.x];
        const double r = fac[t.y] * m * m;
        const double tt = w - a[k];
        const double den = 1.0 / fma(tt, tt, eta2);
        acc = fma(r, tt * den, acc);
        if (DERIV) dacc = fma(r, den * fma(2.0 * eta2, den, -1.0), dacc);
      }
    }
  }
  acc = warp_sum_d(acc);
  if (DERIV) dacc = warp_sum_d(dacc);
  if (lane == 0) {
    out[target] = pref * acc;
    out[nfreq + target] = DERIV ? pref * dacc : 0.0;
  }
}

TreeGeom make_geom(long long T, int leaf) {
  TreeGeom g{};
  g.T = T;
  g.leaf = leaf;
  const long long NL = ceil_div<long long>(T, leaf);
  int D = 0;
  long long cap = 1;
  while (cap < NL) {
    cap *= 4;
    ++D;
  }
  GW_REQUIRE(D < 14, "sigma tree too deep");
  g.D = D;
  g.d0 = std::min(D, 2);
  long long off = 0;
  for (int d = 0; d <= D; ++d) {
    g.span[d] = (long long)leaf << (2 * (D - d));
    g.count[d] = ceil_div<long long>(T, g.span[d]);
    g.off[d] = 0;
    if (d >= g.d0) {
      g.off[d] = off;
      off += g.count[d];
    }
  }
  g.NT = off;
  GW_REQUIRE(g.count[D] < (1LL << 27), "sigma tree has too many leaves");
  return g;
}

}  // namespace

struct SigmaTree {
  TreeGeom g{};
  bool geom_valid = false;
  std::vector<double> energies_host, pole_host, fac_host;
  int boundary = -1;
  const double* mat = nullptr;
  long long mat_version = -1;
  int nslots = 0;
  std::vector<int> slot_of_slice;  // slice -> slot or -1
  int slots_used = 0;
};

SigmaTree* sigma_tree_create() { return new SigmaTree(); }
void sigma_tree_destroy(SigmaTree* t) { delete t; }
void sigma_tree_invalidate(SigmaTree* t) {
  if (t) t->geom_valid = false;
}

// which: tag for buffer names (0 ppm, 1 exact).  host_* are the arrays last uploaded to st.energies / pole / fac.
void sigma_tree_eval(gwbse_ctx* ctx, gwbse_ctx::SigmaState& st, int which, int nslices_total, int ngroups,
                     const int* slices, const int* gptr, const double* freqs_dev, int nfreq, bool want_deriv,
                     double* out_dev) {
  if (!st.tree) st.tree = sigma_tree_create();
  SigmaTree& tr = *st.tree;
  cudaStream_t s = ctx->stream;
  const std::string tag = which == 0 ? "sigtree0_" : "sigtree1_";
  const long long T = (long long)ctx->ntotal * st.npoles;
  GW_REQUIRE(T < (1LL << 32), "sigma tree: too many poles for 32-bit term ids");
  // ---- geometry: sorted pole positions (shared by all levels) ----
  const bool content_changed = tr.mat != st.mat || tr.mat_version != st.content_version;
  if (!tr.geom_valid || tr.g.T != T) {
    // leaf size: smallest power of two >= 64 whose moment store for all slices fits the budget
    int leaf = 64;
    while (leaf < 8192) {
      TreeGeom gg = make_geom(T, leaf);
      if ((double)gg.NT * TREE_P * 8.0 * nslices_total <= (double)ctx->sigma_tree_bytes) break;
      leaf *= 2;
    }
    tr.g = make_geom(T, leaf);
    double* keys = ctx->buf(tag + "keys", (size_t)T);
    double* keys2 = ctx->buf(tag + "a", (size_t)T);
    unsigned int* vals = reinterpret_cast<unsigned int*>(ctx->buf(tag + "vals", (size_t)T / 2 + 2));
    unsigned int* vals2 = reinterpret_cast<unsigned int*>(ctx->buf(tag + "vals2", (size_t)T / 2 + 2));
    int2* term = reinterpret_cast<int2*>(ctx->buf(tag + "term", (size_t)T));
    const int blocks = (int)ceil_div<long long>(T, 256);
    tree_keys_kernel<<<blocks, 256, 0, s>>>(T, ctx->ntotal, st.nocc_boundary, st.energies, st.pole, st.fac, keys,
                                            vals);
    GW_CUDA(cudaGetLastError());
    size_t tmp_bytes = 0;
    GW_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, keys, keys2, vals, vals2, (int)T, 0, 64, s));
    double* tmp = ctx->buf(tag + "sorttmp", tmp_bytes / 8 + 2);
    GW_CUDA(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, keys, keys2, vals, vals2, (int)T, 0, 64, s));
    tree_terms_kernel<<<blocks, 256, 0, s>>>(T, ctx->ntotal, vals2, term);
    GW_CUDA(cudaGetLastError());
    ctx->launches += 4;
    tr.geom_valid = true;
    tr.slots_used = 0;
    tr.slot_of_slice.assign(nslices_total, -1);
    tr.nslots = 0;
  } else if (content_changed) {
    tr.slots_used = 0;
    tr.slot_of_slice.assign(nslices_total, -1);
  }
  tr.mat = st.mat;
  tr.mat_version = st.content_version;
  const TreeGeom& g = tr.g;
  const double* a = ctx->buf(tag + "a", (size_t)T);
  const int2* term = reinterpret_cast<const int2*>(ctx->buf(tag + "term", (size_t)T));
  // ---- moment store ----
  if (tr.nslots == 0) {
    const double per_slot = (double)g.NT * TREE_P * 8.0;
    tr.nslots = (int)std::max(1.0, std::min<double>(nslices_total, (double)ctx->sigma_tree_bytes / per_slot));
    tr.slot_of_slice.assign(nslices_total, -1);
    tr.slots_used = 0;
  }
  double* mom = ctx->buf(tag + "mom", (size_t)tr.nslots * g.NT * TREE_P);
  int* fslice_d = reinterpret_cast<int*>(ctx->buf(tag + "fslice", (size_t)nfreq / 2 + 2));
  int* fslot_d = reinterpret_cast<int*>(ctx->buf(tag + "fslot", (size_t)nfreq / 2 + 2));
  int* build_d = reinterpret_cast<int*>(ctx->buf(tag + "build", (size_t)nslices_total + 4));
  int* overflow_d = build_d + 2 * (size_t)nslices_total + 2;
  std::vector<int> fslice(nfreq), fslot(nfreq);
  int g0 = 0;
  while (g0 < ngroups) {
    // take groups while their levels fit the slot store; when it is full, recycle it
    std::vector<int> bslice, bslot;
    int g1 = g0;
    for (; g1 < ngroups; ++g1) {
      const int sl = slices[g1];
      GW_REQUIRE(sl >= 0 && sl < nslices_total, "sigma tree: slice out of range");
      if (tr.slot_of_slice[sl] < 0) {
        if (tr.slots_used == tr.nslots) {
          if (g1 > g0) break;  // evaluate what we have, then recycle
          std::fill(tr.slot_of_slice.begin(), tr.slot_of_slice.end(), -1);
          tr.slots_used = 0;
        }
        tr.slot_of_slice[sl] = tr.slots_used++;
        bslice.push_back(sl);
        bslot.push_back(tr.slot_of_slice[sl]);
      }
      for (int f = gptr[g1]; f < gptr[g1 + 1]; ++f) {
        fslice[f] = sl;
        fslot[f] = tr.slot_of_slice[sl];
      }
    }
    const int f0 = gptr[g0], f1 = gptr[g1];
    if (!bslice.empty()) {
      const int nb = (int)bslice.size();
      GW_CUDA(cudaMemcpyAsync(build_d, bslice.data(), sizeof(int) * nb, cudaMemcpyHostToDevice, s));
      GW_CUDA(cudaMemcpyAsync(build_d + nslices_total, bslot.data(), sizeof(int) * nb, cudaMemcpyHostToDevice, s));
      for (int b0 = 0; b0 < nb; b0 += 65535) {
        const int nbb = std::min(65535, nb - b0);
        dim3 grid((unsigned)ceil_div<long long>(g.count[g.D], 128), nbb);
        tree_leaf_moments_kernel<<<grid, 128, 0, s>>>(g, a, term, st.mat, st.ld, st.lstride, st.fac, build_d + b0,
                                                      build_d + nslices_total + b0, mom);
        GW_CUDA(cudaGetLastError());
        ctx->launches++;
        for (int d = g.D - 1; d >= g.d0; --d) {
          dim3 gm((unsigned)ceil_div<long long>(g.count[d], 128), nbb);
          tree_m2m_kernel<<<gm, 128, 0, s>>>(g, d, a, build_d + nslices_total + b0, mom);
          GW_CUDA(cudaGetLastError());
          ctx->launches++;
        }
      }
      // the build lists are re-used by the next batch: wait for the copies to be consumed
      GW_CUDA(cudaStreamSynchronize(s));
    }
    if (f1 > f0) {
      GW_CUDA(cudaMemcpyAsync(fslice_d + f0, fslice.data() + f0, sizeof(int) * (f1 - f0), cudaMemcpyHostToDevice, s));
      GW_CUDA(cudaMemcpyAsync(fslot_d + f0, fslot.data() + f0, sizeof(int) * (f1 - f0), cudaMemcpyHostToDevice, s));
      GW_CUDA(cudaMemsetAsync(overflow_d, 0, sizeof(int), s));
      const int nf = f1 - f0;
      const int blocks = ceil_div(nf, TREE_WARPS);
      double* tmp_out = ctx->buf(tag + "out", (size_t)2 * nf);
      if (want_deriv)
        tree_eval_kernel<true><<<blocks, TREE_WARPS * 32, 0, s>>>(g, a, term, st.mat, st.ld, st.lstride, st.fac, mom,
                                                                 nf, freqs_dev + f0, fslice_d + f0, fslot_d + f0,
                                                                 st.eta, st.diag_pref, tmp_out, overflow_d);
      else
        tree_eval_kernel<false><<<blocks, TREE_WARPS * 32, 0, s>>>(g, a, term, st.mat, st.ld, st.lstride, st.fac,
                                                                  mom, nf, freqs_dev + f0, fslice_d + f0,
                                                                  fslot_d + f0, st.eta, st.diag_pref, tmp_out,
                                                                  overflow_d);
      GW_CUDA(cudaGetLastError());
      ctx->launches++;
      GW_CUDA(cudaMemcpyAsync(out_dev + f0, tmp_out, sizeof(double) * nf, cudaMemcpyDeviceToDevice, s));
      GW_CUDA(cudaMemcpyAsync(out_dev + nfreq + f0, tmp_out + nf, sizeof(double) * nf, cudaMemcpyDeviceToDevice, s));
      int overflow = 0;
      GW_CUDA(cudaMemcpyAsync(&overflow, overflow_d, sizeof(int), cudaMemcpyDeviceToHost, s));
      GW_CUDA(cudaStreamSynchronize(s));
      GW_REQUIRE(overflow == 0, "sigma tree traversal stack overflow");
    }
    g0 = g1;
  }
}

}  // namespace gwbse
