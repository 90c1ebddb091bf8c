// Internal state of a gwbse_ctx (one per process / GPU).
#pragma once
#include <nvtx3/nvToolsExt.h>
#include <cuda_runtime.h>
#include <cusolverDn.h>

#include <chrono>
#include <map>
#include <string>
#include <vector>

#include "common.cuh"
#include "gemm_dmma.cuh"

namespace gwbse {
struct SigmaTree;  // sigma_tree.cu
}

struct DevBuf {
  double* p = nullptr;
  size_t cap = 0;  // doubles
};

struct gwbse_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  // host -> device staging of AO blocks (gwbse_mmn_fill_block): copies run on their own stream into two
  // alternating device buffers so that they overlap the contraction GEMMs of the previous block
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t fill_copied[2] = {nullptr, nullptr}, fill_consumed[2] = {nullptr, nullptr};
  int fill_slot = 0;
  int num_sms = 148;
  std::string err;
  cusolverDnHandle_t solver = nullptr;
  long long launches = 0;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  // region profiler (set_option "profile"): CUDA-event pairs per C-ABI entry point + host wall time
  bool profile = false;
  struct Region {
    double ms = 0.0, wall_ms = 0.0;
    long long calls = 0;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> pending;
  };
  std::map<std::string, Region> regions;
  std::vector<cudaEvent_t> event_pool;
  cudaEvent_t get_event() {
    if (!event_pool.empty()) {
      cudaEvent_t e = event_pool.back();
      event_pool.pop_back();
      return e;
    }
    cudaEvent_t e;
    GW_CUDA(cudaEventCreate(&e));
    return e;
  }
  void collect_regions() {
    GW_CUDA(cudaStreamSynchronize(stream));
    for (auto& kv : regions) {
      for (auto& pr : kv.second.pending) {
        float ms = 0.f;
        GW_CUDA(cudaEventElapsedTime(&ms, pr.first, pr.second));
        kv.second.ms += ms;
        event_pool.push_back(pr.first);
        event_pool.push_back(pr.second);
      }
      kv.second.pending.clear();
    }
  }
  // GEMM accounting (gwbse_gemm_profile)
  bool gemm_profile = false;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> gemm_events;
  size_t gemm_events_used = 0;
  double gemm_ms = 0.0, gemm_flops = 0.0;
  long long gemm_launches = 0;
  // per-shape breakdown (gwbse_gemm_shape_report)
  struct ShapeStat {
    double ms = 0.0, flops = 0.0;
    long long calls = 0;
  };
  std::vector<std::string> gemm_event_keys;
  std::vector<double> gemm_event_flops;
  std::map<std::string, ShapeStat> gemm_shapes;
  void gemm_collect();

  // multi-GPU
  int rank = 0, world = 1;
  void* nccl_comm = nullptr;

  // caching allocator behind gwbse_dev_malloc/free: freed blocks are kept and handed out again for
  // requests of the same size (Davidson / BSE temporaries repeat every iteration)
  std::multimap<size_t, double*> free_blocks;
  std::map<double*, size_t> live_blocks;
  size_t cached_bytes = 0;
  // named grow-only device scratch buffers
  std::map<std::string, DevBuf> bufs;

  // ---- Mmn ----
  int naux = 0, mmin = 0, mmax = -1, nmin = 0, nmax = -1;
  int mtotal = 0, ntotal = 0, npad = 0, mlocal = 0;
  int alloc_world = 0;
  int mlmax = 0;  // slices per rank rounded up (ldx = mlmax * npad on every rank: pad slices are zero)
  // aux-sharded fill (gwbse_mmn_fill_begin/end): this rank contracts aux functions [fill_lo, fill_hi) for ALL m
  // into X2, laid out per destination rank, and the all-to-all of fill_end lands them in the m-sharded X
  bool fill_sharded = false;
  int fill_lo = 0, fill_hi = 0;
  int aux_begin(int r) const { return (int)((long long)r * naux / world); }
  long long ldx = 0;
  double* X = nullptr;      // current Mmn
  double* X2 = nullptr;     // out-of-place target of MultiplyRight
  double* Xsnap = nullptr;  // pristine copy for Rebuild
  double* mos = nullptr;    // MO coefficients on device (nbasis x nmo)
  int nbasis = 0, nmo = 0;

  // ---- RPA ----
  double* eps = nullptr;  // naux x naux

  // ---- Sigma_c evaluators (shared kernel: poles x levels) ----
  struct SigmaState {
    bool ready = false;
    int npoles = 0;        // naux (ppm) or rpasize (exact)
    int nocc_boundary = 0;  // rows [0, boundary) take +pole, the rest -pole
    int qpoff = 0;          // first qp level in Mmn storage index
    int q = 0;
    double eta = 0.0;
    double diag_pref = 1.0;  // multiplies fac[] for the diagonal element
    double offdiag_pref = 1.0;  // exact evaluator: 1 closed shell (2 * 0.5), 0.5 per spin channel
    const double* mat = nullptr;  // matrix holding level blocks: element (level l, n, pole p) at mat[p*ld + l*lstride + n]
    long long ld = 0, lstride = 0;
    double* fac = nullptr;     // per pole prefactor (device)
    double* pole = nullptr;    // per pole frequency (device)
    double* energies = nullptr;  // ntotal (device)
    std::vector<double> energies_host;  // last upload (the treecode geometry is rebuilt only when it changes)
    long long content_version = 0;      // bumped when mat / fac change (treecode moments are rebuilt)
    long long seen_mmn_version = -1;
    gwbse::SigmaTree* tree = nullptr;
  } sig_ppm, sig_exact;
  // treecode evaluator (sigma_tree.cu): used when n * npoles >= sigma_tree_min_terms; moment store budget
  long long sigma_tree_min_terms = 1 << 15;
  size_t sigma_tree_bytes = (size_t)8 << 30;
  double* exact_res = nullptr;  // residues (q*npad) x S

  // ---- QSGW (rpa.h:59-66): rotation of the hole slices inside the QP window, applied in the RPA sums ----
  struct QsgwState {
    bool active = false;
    int qptotal = 0, qpmin = 0, homo = 0;
    double* U = nullptr;              // qptotal x qptotal (device, ld = qptotal)
    long long built_version = -1;     // mmn_version the rotated copy of the occupied slices was built from
    int built_nocc = 0;
  } qsgw;

  // ---- Sigma_CDA (capi_cda.cu) ----
  struct CdaState {
    bool ready = false;
    int order = 0, symmetry = 0;
    double alpha = 0.0, eta = 0.0;
    int homo = 0, rpamin = 0, rpamax = 0, qpmin = 0, q = 0;
    int nq_local = 0, lfirst = 0;  // local slices of the qp window: count and first local index
    long long node_stride = 0;     // doubles between the Q tables of consecutive quadrature nodes
    long long mmn_version = -1;
    std::vector<double> pts, wts;
    // unrestricted reference (sigma_cda_uks.cc): the dielectric matrix is the mean of this channel's and the
    // partner channel's closed-shell-weighted matrices
    gwbse_ctx* partner = nullptr;
    int partner_homo = 0;
    std::vector<double> partner_energies;
  } cda;

  // ---- BSE ----
  struct BseState {
    bool ready = false;
    int homo = 0, rpamin = 0, vmin = 0, cmax = 0;
    int vt = 0, ct = 0, size = 0;
    int voff = 0, coff = 0;  // vmin - rpamin, cmin - rpamin (rows / slices of Mmn)
    double* eps_inv = nullptr;  // naux (device)
    double* hqp = nullptr;      // (vt+ct)^2 device, ld = vt+ct
    long long gathered_version = -1;  // multi-GPU: version of X the replicated vv / cv blocks were built from
    int gathered_voff = -1, gathered_vt = -1, gathered_ct = -1;
    std::vector<double> hqp_host;
    std::vector<double> eps_inv_host;  // last screening uploaded (eps_version counts its changes)
    long long eps_version = 0;
    // Materialised direct-interaction blocks (capi_bse.cu): kind 0 = Hd, 1 = Hd2.  Column j of H holds the
    // coefficients of local output row j against every input row (v2, c2): H[(v2 * ct + c2) + j * ld].
    struct DenseBlock {
      double* H = nullptr;
      long long ld = 0, ncols = 0;
      bool valid = false, in_x2 = false, refused = false;
      long long x2_epoch = -1;
      // key the block (and the payback counter) belongs to
      long long key_mmn = -1, key_eps = -1;
      int vt = 0, ct = 0, voff = 0, coff = 0;
      double spent_flops = 0.0;  // factorised work done under this key so far
    } dense[2];
  } bse;
  // bse_dense: 0 never materialise, 1 when it pays back (default), 2 always (if memory allows)
  int bse_dense_mode = 1;
  double bse_dense_payback = 0.25;  // build once the factorised work under one key reaches this fraction of a build
  long long bse_dense_builds = 0, bse_dense_columns = 0;
  long long x2_epoch = 0;  // bumped whenever the second Mmn buffer is written or handed out as scratch
  // MultiplyRight restricted to a window of n (gwbse_mmn_mul_right_window_dev): rows n in [n_lo, n_hi) of every slice
  // carry the rotation R, the others get it when an entry point that may read them runs (mmn_complete_rotation)
  struct PendingRotation {
    bool active = false;
    int n_lo = 0, n_hi = 0;
    double* R = nullptr;  // naux x naux at pitch ld (named buffer "pending_R")
    int ld = 0;
  } pending_rot;
  double bse_algo_flops = 0.0;  // SURVEY.md 8(d) F_bse summed over the operator products so far
  long long bse_columns = 0, bse_products = 0;
  size_t bse_chunk_bytes = (size_t)8 << 30;  // size of the Hd intermediate per chunk

  double* buf(const std::string& name, size_t n) {
    DevBuf& b = bufs[name];
    if (b.cap < n) {
      if (b.p) GW_CUDA(cudaFree(b.p));
      b.p = nullptr;
      b.cap = 0;
      GW_CUDA(cudaMalloc(&b.p, sizeof(double) * n));
      b.cap = n;
    }
    return b.p;
  }
  void release_buf(const std::string& name) {
    auto it = bufs.find(name);
    if (it == bufs.end()) return;
    if (it->second.p) GW_CUDA(cudaFree(it->second.p));
    bufs.erase(it);
  }
  // algo_flops < 0: 2*M*N*K per batch (half for lower_only)
  void gemm(const gwbse::GemmParams& p, int cfg = -1, int splitk = 0, double algo_flops = -1.0);
  long long mmn_version = 0;  // bumped whenever the content of X changes (gathered blocks are cached per version)
  // first global slice >= s0 owned by rank r, and how many slices of [s0, s0+ns) rank r owns
  int first_owned(int s0, int r) const { return s0 + ((r - s0 % world) % world + world) % world; }
  int owned_count(int s0, int ns, int r) const {
    const int f = first_owned(s0, r);
    return f >= s0 + ns ? 0 : (s0 + ns - f + world - 1) / world;
  }
  bool owns(int m) const { return (m % world) == rank; }
  int local_index(int m) const { return m / world; }
  // number of local levels with global storage index in [0, upto)
  int local_count(int upto) const { return upto <= rank ? 0 : (upto - rank + world - 1) / world; }
};

// Every C-ABI entry point is an NVTX range (nvtx3 is header-only and a no-op unless a tool is attached: nsys / ncu
// --nvtx show the reference's stage names) and, with the option "profile", a CUDA-event region.
struct ProfScope {
  gwbse_ctx* ctx;
  const char* name;
  cudaEvent_t e0 = nullptr;
  std::chrono::steady_clock::time_point t0;
  ProfScope(gwbse_ctx* c, const char* n) : ctx(c), name(n) {
    nvtxRangePushA(n);
    if (!ctx->profile) return;
    e0 = ctx->get_event();
    cudaEventRecord(e0, ctx->stream);
    t0 = std::chrono::steady_clock::now();
  }
  ~ProfScope() {
    nvtxRangePop();
    if (!e0) return;
    cudaEvent_t e1 = ctx->get_event();
    cudaEventRecord(e1, ctx->stream);
    auto& r = ctx->regions[name];
    r.pending.emplace_back(e0, e1);
    r.calls++;
    r.wall_ms += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    if (r.pending.size() > 2048) ctx->collect_regions();
  }
};
#define GW_PROF(ctx, name) ProfScope _prof_scope(ctx, name)

#define GW_API_BEGIN(ctx) \
  if (!(ctx)) return 1;   \
  try {                   \
    GW_CUDA(cudaSetDevice((ctx)->device));
#define GW_API_END(ctx)               \
  return 0;                           \
  }                                   \
  catch (const std::exception& e) {   \
    (ctx)->err = e.what();            \
    return 1;                         \
  }                                   \
  catch (...) {                       \
    (ctx)->err = "unknown exception"; \
    return 1;                         \
  }

namespace gwbse {
// slice gather (capi_shard.cu): rows [row0,row0+nrows) of global slices [s0,s0+ns), poles [p0,p0+np), from all
// ranks into out[(p-p0)*ldo + (s-s0)*rpad + (row-row0)] in natural slice order (ldo >= ns*rpad)
void gather_slices(gwbse_ctx* ctx, int s0, int ns, int row0, int nrows, int p0, int np, double* out, long long ldo,
                   int rpad);
// Hole slices as the RPA sums see them (rpa.cc:92-118): element (v, c, chi) at ptr[chi * s_chi + v * s_v + c], c
// counting the unoccupied rows.  Plain view into X, or - with a QSGW rotation registered - a copy of the occupied
// slices in which those inside the QP window are sum_vp U(vp, v) Mmn[vp] (capi_mmn.cu).
struct HoleView {
  const double* ptr;
  long long s_chi, s_v;
};
HoleView hole_view(gwbse_ctx* ctx, int n_occ);
// applies a rotation left pending by gwbse_mmn_mul_right_window_dev to the rows outside its window (capi_mmn.cu);
// every entry point that may read such rows calls it first, the BSE operator's only if its window is not covered
void mmn_complete_rotation(gwbse_ctx* ctx);
// the second Mmn buffer (out-of-place target of MultiplyRight) as scratch of the same shape as X (capi_mmn.cu)
double* mmn_scratch_x2(gwbse_ctx* ctx);
// Fill3cMO contraction of a device-resident AO block whose columns are pitch doubles apart (capi_mmn.cu)
void mmn_fill_block_pitched(gwbse_ctx* ctx, int aux_offset, int aux_count, const double* ao3c_dev, long long pitch);
// NCCL plumbing (comm.cu)
void allreduce_dev(gwbse_ctx* ctx, double* buf_dev, size_t n);
void allgather_dev(gwbse_ctx* ctx, const double* send_dev, double* recv_dev, size_t n_per_rank);
// send[r] / recv[r]: device buffers exchanged with rank r (counts in doubles), one grouped NCCL call
void alltoallv_dev(gwbse_ctx* ctx, const double* const* send, const size_t* send_count, double* const* recv,
                   const size_t* recv_count);
// streaming kernels (streaming.cu)
void launch_symmetrize_lower(double* A, int n, long long ld, cudaStream_t s);
void launch_add_diagonal(double* A, int n, long long ld, double v, cudaStream_t s);
// w[ml * ldw + c]: rows padded to an even pitch so that the weight tile of a k-step is a 16-byte aligned TMA box
void launch_rpa_weights(double* w, long long ldw, const double* e, int kind, double fre, double fim, double eta,
                        int n_occ, int n_unocc, int rank, int world, int nloc_occ, cudaStream_t s);
void launch_diag_scale(char side, int m, int n, const double* A, long long lda, const double* d, double* C,
                       long long ldc, cudaStream_t s);
void launch_axpy(int m, int n, double alpha, const double* X, long long ldx, double* Y, long long ldy,
                 cudaStream_t s);
void launch_colnorms(int m, int n, const double* A, long long lda, double* out_dev, cudaStream_t s);
void launch_coldots(int m, int n, const double* X, long long ldx, const double* Y, long long ldy, double* out_dev,
                    cudaStream_t s);
void launch_scale_cols(int m, int n, double* A, long long lda, const double* s_dev, cudaStream_t s);
void launch_copy_block(int m, int n, const double* A, long long lda, double* B, long long ldb, cudaStream_t s);
void launch_pack_block(const double* src, long long s_pole, long long s_outer, int L1, int L2, const double* scale,
                       double* out, long long plane, int npoles, cudaStream_t s);
void launch_rotate_scatter(const double* T, int q, int nloc, int naux, double* X, long long ldx, int npad, int lfirst,
                           int row0, cudaStream_t s);
void launch_invsqrt_scale(double* out, const double* w, int n, double etol, int* removed_dev, cudaStream_t s);
// treecode Sigma_c (sigma_tree.cu): slices[g] = local slice of group g, frequencies gptr[g]..gptr[g+1]
void sigma_tree_eval(gwbse_ctx* ctx, gwbse_ctx::SigmaState& st, int which, int nslices_total, int ngroups,
                     const int* slices, const int* gptr, const double* freqs_dev, int nfreq, bool want_deriv,
                     double* out_dev);
void sigma_tree_invalidate(SigmaTree* t);
void sigma_tree_destroy(SigmaTree* t);
int sigma_multi_chunks(int npoles);
void launch_sigma_multi(const gwbse_ctx::SigmaState& st, int ntotal, int ngroups, int nfreq, const int* levels_dev,
                        const int* gptr_dev, const double* freqs_dev, double* partial_dev, double* out_dev,
                        bool want_deriv, cudaStream_t s);
void launch_sigma_offdiag_weight(const gwbse_ctx::SigmaState& st, int ntotal, int npad, int nlevels,
                                 const int* slice_idx_dev, const int* freq_idx_dev, int p0, int np,
                                 const double* freqs_dev, double pref, double* out, long long ldo, cudaStream_t s);
void launch_offdiag_finish(const double* S, int q, double* out, cudaStream_t s);
void launch_slice_diag(const double* X, long long ldx, int npad, int naux, int s0, int ns, int rank, int world,
                       double* D, cudaStream_t s);
void launch_bse_diag(const double* X, long long ldx, int npad, int naux, int vt, int ct, int voff, int coff,
                     int v_rel0, int vstride, int lfirst, int nvloc, const double* cv, long long cv_row,
                     long long cv_pole, const double* Dcc, const double* Dvv, const double* eps_inv,
                     const double* hqp, int ldh, int cqp, int cx, int cd, int cd2, double* out, cudaStream_t s);
void launch_dpr(int rows, int ncols, const double* diag, const double* lambda_dev, const double* R, long long ldr,
                double* W, long long ldw, cudaStream_t s);
void launch_olsen_finish(int rows, int ncols, const double* diag, const double* lambda_dev, const double* Q,
                         long long ldq, const double* num_dev, const double* den_dev, double* W, long long ldw,
                         cudaStream_t s);
}  // namespace gwbse
