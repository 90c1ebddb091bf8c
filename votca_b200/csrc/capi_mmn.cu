// C ABI, part 2: the Mmn tensor (TCMatrix_gwbse), RPA dielectric matrix, Sigma_x.
#include <algorithm>
#include <cstdint>
#include <vector>

#include "../../include/gwbse_b200.h"
#include "context.cuh"

using namespace gwbse;

namespace {

inline int round_up(int v, int a) { return (v + a - 1) / a * a; }

// every entry point of this file may read any row of the tensor: a rotation pending outside its window is applied
void require_mmn(gwbse_ctx* ctx) {
  GW_REQUIRE(ctx->X != nullptr, "Mmn not allocated (gwbse_mmn_alloc)");
  gwbse::mmn_complete_rotation(ctx);
}

// contraction for a block of aux functions whose AO integrals already sit on the device
// second Mmn-sized buffer: out-of-place target of MultiplyRight and staging area of the aux-sharded fill
// (world extra pole rows cover the rounding of the aux partition)
void ensure_x2(gwbse_ctx* ctx) {
  if (ctx->X2) return;
  size_t fr = 0, tot = 0;
  GW_CUDA(cudaMemGetInfo(&fr, &tot));
  const size_t bytes = sizeof(double) * (size_t)ctx->ldx * (ctx->naux + ctx->world);
  if (bytes > fr)
    throw std::runtime_error("There were requested : " + std::to_string((double)bytes / 1048576.0) +
                             " MB but the device has " + std::to_string((double)fr / 1048576.0) + " MB free.");
  GW_CUDA(cudaMalloc(&ctx->X2, bytes));
}

// pitch: doubles between consecutive columns of one N x N block (blocks are pitch * N apart); 0 = N (contiguous)
void fill_block_dev(gwbse_ctx* ctx, int aux_offset, int aux_count, const double* ao3c_dev, long long pitch = 0) {
  require_mmn(ctx);
  GW_REQUIRE(ctx->mos != nullptr, "MO coefficients not set (gwbse_mmn_set_mos)");
  GW_REQUIRE(aux_offset >= 0 && aux_offset + aux_count <= ctx->naux, "aux block out of range");
  const int N = ctx->nbasis;
  if (pitch == 0) pitch = N;
  GW_REQUIRE(pitch >= N, "AO block pitch smaller than the basis size");
  GW_REQUIRE(ctx->mmax < ctx->nmo && ctx->nmax < ctx->nmo, "level range exceeds number of MOs");
  if (aux_count == 0) return;
  const bool sh = ctx->fill_sharded;
  if (sh)
    GW_REQUIRE(aux_offset >= ctx->fill_lo && aux_offset + aux_count <= ctx->fill_hi,
               "aux block outside this rank's share of the aux-sharded fill (gwbse_shard_aux_range)");
  // columns contracted here: the local m slices, or every m when the fill is sharded over aux functions
  const int mcols = sh ? ctx->mtotal : ctx->mlocal;
  if (mcols == 0) return;
  const long long seg = (long long)(ctx->fill_hi - ctx->fill_lo) * ctx->ldx;  // per destination rank
  const int max_batch = 4096;
  for (int b0 = 0; b0 < aux_count; b0 += max_batch) {
    const int nb = std::min(max_batch, aux_count - b0);
    const int Np = round_up(N, 2);  // even pitch of the MO matrix and of the half-transformed blocks
    double* H = ctx->buf("fill_H", (size_t)Np * mcols * nb);
    // H_k[nu, ml] = sum_mu T_k[mu, nu] C[mu, m(ml)]      (2 N^2 mcols flops per aux function)
    GemmParams p;
    p.M = N;
    p.N = mcols;
    p.Ki = N;
    p.Z1 = nb;
    p.A.ptr = ao3c_dev + (size_t)b0 * pitch * N;
    p.A.s_ri = pitch;
    p.A.s_ki = 1;
    p.A.s_z1 = pitch * N;
    p.B.ptr = ctx->mos + (size_t)(ctx->mmin + (sh ? 0 : ctx->rank)) * Np;
    p.B.s_ri = (long long)(sh ? 1 : ctx->world) * Np;
    p.B.s_ki = 1;
    p.C = H;
    p.sC_mi = 1;
    p.sC_ni = Np;
    p.sC_z1 = (long long)Np * mcols;
    ctx->gemm(p);
    // M[m](n, k) = sum_nu C[nu, nmin+n] H_k[nu, ml]       (2 n N mcols flops per aux function)
    GemmParams q;
    q.M = ctx->ntotal;
    q.N = mcols;
    q.Ki = N;
    q.Z1 = nb;
    q.A.ptr = ctx->mos + (size_t)ctx->nmin * Np;
    q.A.s_ri = Np;
    q.A.s_ki = 1;
    q.B.ptr = H;
    q.B.s_ri = Np;
    q.B.s_ki = 1;
    q.B.s_z1 = (long long)Np * mcols;
    q.sC_mi = 1;
    q.sC_z1 = ctx->ldx;
    if (sh) {
      // X2[dest = m % world][chi - fill_lo][m / world][n]: each destination's part is one contiguous segment
      q.C = ctx->X2 + (long long)(aux_offset - ctx->fill_lo + b0) * ctx->ldx;
      q.Ln = ctx->world;
      q.sC_ni = seg;
      q.sC_no = ctx->npad;
    } else {
      q.C = ctx->X + (long long)(aux_offset + b0) * ctx->ldx;
      q.sC_ni = ctx->npad;
    }
    ctx->gemm(q);
  }
  ctx->mmn_version++;
}

}  // namespace

namespace gwbse {
// capi_ao3c.cu: blocks produced on the device carry an even pitch so the contraction takes 16-byte copies
void mmn_fill_block_pitched(gwbse_ctx* ctx, int aux_offset, int aux_count, const double* ao3c_dev, long long pitch) {
  GW_PROF(ctx, "mmn_fill_block");
  fill_block_dev(ctx, aux_offset, aux_count, ao3c_dev, pitch);
}
}  // namespace gwbse

namespace {

void mul_right_dev(gwbse_ctx* ctx, const double* R_dev, int ldr) {
  require_mmn(ctx);
  if (ctx->ldx == 0) return;
  GW_REQUIRE(ldr >= ctx->naux, "Shape mismatch in MultiplyRight");
  ensure_x2(ctx);
  // R with an odd leading dimension (Naux is odd for most basis sets) or an unaligned base would force 8-byte
  // copies for the B operand: re-pitch it to an even ld first (Naux^2 doubles, microseconds)
  if ((ldr & 1) || (reinterpret_cast<uintptr_t>(R_dev) & 15)) {
    const int ldp = round_up(ctx->naux, 2);
    double* Rp = ctx->buf("mulright_Rp", (size_t)ldp * ctx->naux);
    GW_CUDA(cudaMemcpy2DAsync(Rp, sizeof(double) * ldp, R_dev, sizeof(double) * ldr, sizeof(double) * ctx->naux,
                              ctx->naux, cudaMemcpyDeviceToDevice, ctx->stream));
    R_dev = Rp;
    ldr = ldp;
  }
  // X2 = X * R as one flat GEMM over all (m, n) rows (padding rows are zero and stay zero)
  GW_REQUIRE(ctx->ldx < (1LL << 31), "Mmn row count exceeds 2^31");
  GemmParams p;
  p.M = (int)ctx->ldx;
  p.N = ctx->naux;
  p.Ki = ctx->naux;
  p.A.ptr = ctx->X;
  p.A.s_ri = 1;
  p.A.s_ki = ctx->ldx;
  p.B.ptr = R_dev;
  p.B.s_ri = ldr;
  p.B.s_ki = 1;
  p.C = ctx->X2;
  p.sC_mi = 1;
  p.sC_ni = ctx->ldx;
  ctx->gemm(p);
  std::swap(ctx->X, ctx->X2);
  ctx->mmn_version++;
  ctx->x2_epoch++;
  // evaluators hold pointers into X
  ctx->sig_ppm.ready = false;
}

}  // namespace

namespace {
// X[ml][n, :] <- X[ml][n, :] R for n in [n0, n1) of every local slice: the product goes through the second Mmn
// buffer (same addresses), then back in place.  ldx = mlmax * npad, so (chi, ml) is one row index of pitch npad.
void rotate_rows(gwbse_ctx* ctx, const double* Rp, int ldp, int n0, int n1) {
  const int nw = n1 - n0;
  if (nw <= 0 || ctx->ldx == 0) return;
  ensure_x2(ctx);
  GemmParams p;
  p.M = ctx->mlmax * nw;
  p.N = ctx->naux;
  p.Ki = ctx->naux;
  p.A.ptr = ctx->X + n0;
  p.A.Lr = nw;
  p.A.s_ri = 1;
  p.A.s_ro = ctx->npad;
  p.A.s_ki = ctx->ldx;
  p.B.ptr = Rp;
  p.B.s_ri = ldp;
  p.B.s_ki = 1;
  p.C = ctx->X2 + n0;
  p.Lm = nw;
  p.sC_mi = 1;
  p.sC_mo = ctx->npad;
  p.sC_ni = ctx->ldx;
  ctx->gemm(p);
  GW_CUDA(cudaMemcpy2DAsync(ctx->X + n0, sizeof(double) * ctx->npad, ctx->X2 + n0, sizeof(double) * ctx->npad,
                            sizeof(double) * nw, (size_t)ctx->mlmax * ctx->naux, cudaMemcpyDeviceToDevice,
                            ctx->stream));
  ctx->x2_epoch++;
}
}  // namespace

namespace gwbse {
void mmn_complete_rotation(gwbse_ctx* ctx) {
  auto& pr = ctx->pending_rot;
  if (!pr.active) return;
  pr.active = false;  // first: the calls below go through entry-point helpers
  if (ctx->X == nullptr) return;
  rotate_rows(ctx, pr.R, pr.ld, 0, pr.n_lo);
  rotate_rows(ctx, pr.R, pr.ld, pr.n_hi, ctx->ntotal);
  ctx->mmn_version++;
  ctx->sig_ppm.ready = false;
}

HoleView hole_view(gwbse_ctx* ctx, int n_occ) {
  auto& qs = ctx->qsgw;
  if (!qs.active) return {ctx->X + n_occ, ctx->ldx, ctx->npad};
  GW_REQUIRE(ctx->world == 1, "QSGW is single-GPU (the hole rotation mixes slices of every rank)");
  const int npad = ctx->npad, naux = ctx->naux;
  const long long ldo = (long long)n_occ * npad;
  double* Xo = ctx->buf("qsgw_Xocc", (size_t)ldo * naux);
  if (qs.built_version != ctx->mmn_version || qs.built_nocc != n_occ) {
    const int off = qs.qpmin - ctx->mmin;
    const int end_occ = std::min(qs.homo - ctx->mmin + 1, off + qs.qptotal);
    GW_REQUIRE(off >= 0 && off + qs.qptotal <= ctx->mtotal, "QSGW rotation outside the Mmn range");
    // occupied slices outside the QP window: unchanged
    GW_CUDA(cudaMemcpy2DAsync(Xo, sizeof(double) * ldo, ctx->X, sizeof(double) * ctx->ldx, sizeof(double) * ldo, naux,
                              cudaMemcpyDeviceToDevice, ctx->stream));
    if (end_occ > off) {
      // Xo[chi][off + v][n] = sum_vp U(vp, v) X[chi][off + vp][n]: rows (chi, n), k = vp
      GemmParams p;
      GW_REQUIRE((long long)naux * npad < (1LL << 31), "Mmn too large for the hole rotation");
      p.M = naux * npad;
      p.N = end_occ - off;
      p.Ki = qs.qptotal;
      p.A.ptr = ctx->X + (long long)off * npad;
      p.A.Lr = npad;
      p.A.s_ri = 1;
      p.A.s_ro = ctx->ldx;
      p.A.s_ki = npad;
      p.B.ptr = qs.U;
      p.B.s_ri = qs.qptotal;
      p.B.s_ki = 1;
      p.C = Xo + (long long)off * npad;
      p.Lm = npad;
      p.sC_mi = 1;
      p.sC_mo = ldo;
      p.sC_ni = npad;
      ctx->gemm(p);
    }
    qs.built_version = ctx->mmn_version;
    qs.built_nocc = n_occ;
  }
  return {Xo + n_occ, ldo, npad};
}

double* mmn_scratch_x2(gwbse_ctx* ctx) {
  require_mmn(ctx);
  ensure_x2(ctx);
  ctx->x2_epoch++;  // whatever was parked in the buffer (a materialised BSE block) is gone
  return ctx->X2;
}
}  // namespace gwbse

extern "C" {

int gwbse_mmn_alloc(gwbse_ctx* ctx, int naux, int mmin, int mmax, int nmin, int nmax) {
  GW_API_BEGIN(ctx)
  GW_REQUIRE(naux > 0 && mmax >= mmin && nmax >= nmin && mmin >= 0 && nmin >= 0, "invalid Mmn ranges");
  GW_CUDA(cudaStreamSynchronize(ctx->stream));
  // same shape as the tensor already held (a job run again, Rebuild without snapshot): keep the three buffers
  const bool same = ctx->X != nullptr && ctx->naux == naux && ctx->mmin == mmin && ctx->mmax == mmax &&
                    ctx->nmin == nmin && ctx->nmax == nmax && ctx->alloc_world == ctx->world;
  if (!same) {
    for (double** p : {&ctx->X, &ctx->X2, &ctx->Xsnap}) {
      if (*p) GW_CUDA(cudaFree(*p));
      *p = nullptr;
    }
    // materialised BSE blocks of the previous shape: give the memory back before the new tensor is sized
    for (const char* name : {"bse_dense0", "bse_dense1", "bse_packA", "bse_packB"}) ctx->release_buf(name);
  }
  ctx->x2_epoch++;
  ctx->pending_rot.active = false;
  for (auto& blk : ctx->bse.dense) blk.valid = false;
  ctx->alloc_world = ctx->world;
  ctx->naux = naux;
  ctx->mmin = mmin;
  ctx->mmax = mmax;
  ctx->nmin = nmin;
  ctx->nmax = nmax;
  ctx->mtotal = mmax - mmin + 1;
  ctx->ntotal = nmax - nmin + 1;
  ctx->npad = round_up(ctx->ntotal, 16);
  ctx->mlocal = ctx->local_count(ctx->mtotal);
  ctx->mlmax = (ctx->mtotal + ctx->world - 1) / ctx->world;
  ctx->ldx = (long long)ctx->mlmax * ctx->npad;
  ctx->fill_sharded = false;
  // world extra pole rows: X and X2 trade places in MultiplyRight and X2 stages the aux-sharded fill
  const size_t bytes = sizeof(double) * (size_t)std::max<long long>(ctx->ldx, 1) * (naux + ctx->world);
  if (!ctx->X) {
    size_t fr = 0, tot = 0;
    GW_CUDA(cudaMemGetInfo(&fr, &tot));
    if (bytes > fr)
      throw std::runtime_error("There were requested : " + std::to_string((double)bytes / 1048576.0) +
                               " MB but the device has " + std::to_string((double)fr / 1048576.0) + " MB free.");
    GW_CUDA(cudaMalloc(&ctx->X, bytes));
  }
  GW_CUDA(cudaMemsetAsync(ctx->X, 0, bytes, ctx->stream));
  ctx->mmn_version++;
  ctx->sig_ppm.ready = ctx->sig_exact.ready = ctx->bse.ready = false;
  GW_API_END(ctx)
}

int gwbse_mmn_free(gwbse_ctx* ctx) {
  GW_API_BEGIN(ctx)
  GW_CUDA(cudaStreamSynchronize(ctx->stream));
  for (double** p : {&ctx->X, &ctx->X2, &ctx->Xsnap}) {
    if (*p) GW_CUDA(cudaFree(*p));
    *p = nullptr;
  }
  for (const char* name : {"bse_dense0", "bse_dense1", "bse_packA", "bse_packB"}) ctx->release_buf(name);
  ctx->x2_epoch++;
  ctx->pending_rot.active = false;
  for (auto& blk : ctx->bse.dense) blk.valid = false;
  ctx->sig_ppm.ready = ctx->sig_exact.ready = ctx->bse.ready = false;
  GW_API_END(ctx)
}

int gwbse_mmn_dims(const gwbse_ctx* ctx, int* naux, int* mtotal, int* ntotal, int* mlocal, int* npad) {
  if (!ctx) return 1;
  if (naux) *naux = ctx->naux;
  if (mtotal) *mtotal = ctx->mtotal;
  if (ntotal) *ntotal = ctx->ntotal;
  if (mlocal) *mlocal = ctx->mlocal;
  if (npad) *npad = ctx->npad;
  return 0;
}

int gwbse_mmn_set_mos(gwbse_ctx* ctx, const double* mos, int ldmos, int nbasis, int nmo) {
  GW_API_BEGIN(ctx)
  GW_REQUIRE(ldmos >= nbasis && nbasis > 0 && nmo > 0, "invalid MO matrix shape");
  GW_CUDA(cudaStreamSynchronize(ctx->stream));
  if (ctx->mos) GW_CUDA(cudaFree(ctx->mos));
  ctx->mos = nullptr;
  const int Np = round_up(nbasis, 2);  // even pitch: 16-byte aligned columns for the fill GEMMs
  GW_CUDA(cudaMalloc(&ctx->mos, sizeof(double) * (size_t)Np * nmo));
  GW_CUDA(cudaMemsetAsync(ctx->mos, 0, sizeof(double) * (size_t)Np * nmo, ctx->stream));
  GW_CUDA(copy2d_async(ctx->mos, sizeof(double) * Np, mos, sizeof(double) * ldmos, sizeof(double) * nbasis,
                            nmo, cudaMemcpyHostToDevice, ctx->stream));
  GW_CUDA(cudaStreamSynchronize(ctx->stream));
  ctx->nbasis = nbasis;
  ctx->nmo = nmo;
  GW_API_END(ctx)
}

int gwbse_mmn_fill_block_dev(gwbse_ctx* ctx, int aux_offset, int aux_count, const double* ao3c_dev) {
  GW_API_BEGIN(ctx)
  GW_PROF(ctx, "mmn_fill_block");
  fill_block_dev(ctx, aux_offset, aux_count, ao3c_dev);
  GW_API_END(ctx)
}

int gwbse_mmn_fill_block(gwbse_ctx* ctx, int aux_offset, int aux_count, const double* ao3c) {
  GW_API_BEGIN(ctx)
  GW_PROF(ctx, "mmn_fill_block_h2d");
  const size_t per = (size_t)ctx->nbasis * ctx->nbasis;
  if (!ctx->copy_stream) {
    GW_CUDA(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
    for (int i = 0; i < 2; ++i) {
      GW_CUDA(cudaEventCreateWithFlags(&ctx->fill_copied[i], cudaEventDisableTiming));
      GW_CUDA(cudaEventCreateWithFlags(&ctx->fill_consumed[i], cudaEventDisableTiming));
    }
  }
  // sub-blocks of <= 256 MiB through two alternating staging buffers: the copy of block i+1 (copy stream)
  // overlaps the two contraction GEMMs of block i (compute stream)
  const int chunk = (int)std::max<size_t>(1, std::min<size_t>(aux_count, ((size_t)1 << 25) / std::max<size_t>(per, 1)));
  double* stage[2] = {ctx->buf("fill_ao0", per * chunk), ctx->buf("fill_ao1", per * chunk)};
  for (int b0 = 0; b0 < aux_count; b0 += chunk) {
    const int nb = std::min(chunk, aux_count - b0);
    const int s = ctx->fill_slot;
    ctx->fill_slot ^= 1;
    GW_CUDA(cudaStreamWaitEvent(ctx->copy_stream, ctx->fill_consumed[s], 0));
    GW_CUDA(cudaMemcpyAsync(stage[s], ao3c + (size_t)b0 * per, sizeof(double) * per * nb, cudaMemcpyHostToDevice,
                            ctx->copy_stream));
    GW_CUDA(cudaEventRecord(ctx->fill_copied[s], ctx->copy_stream));
    GW_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->fill_copied[s], 0));
    fill_block_dev(ctx, aux_offset + b0, nb, stage[s]);
    GW_CUDA(cudaEventRecord(ctx->fill_consumed[s], ctx->stream));
  }
  // the caller may overwrite ao3c as soon as we return; the GEMMs of the last block keep running
  GW_CUDA(cudaStreamSynchronize(ctx->copy_stream));
  GW_API_END(ctx)
}

int gwbse_host_malloc(size_t bytes, void** out) {
  if (!out) return 1;
  *out = nullptr;
  return cudaHostAlloc(out, std::max<size_t>(bytes, 1), cudaHostAllocDefault) == cudaSuccess ? 0 : 1;
}

int gwbse_host_free(void* p) { return (!p || cudaFreeHost(p) == cudaSuccess) ? 0 : 1; }

int gwbse_shard_aux_range(const gwbse_ctx* ctx, int rank, int* begin, int* end) {
  if (!ctx || rank < 0 || rank >= ctx->world) return 1;
  if (begin) *begin = ctx->aux_begin(rank);
  if (end) *end = ctx->aux_begin(rank + 1);
  return 0;
}

int gwbse_mmn_fill_begin(gwbse_ctx* ctx, int aux_sharded) {
  GW_API_BEGIN(ctx)
  require_mmn(ctx);
  ctx->fill_sharded = aux_sharded != 0 && ctx->world > 1;
  if (ctx->fill_sharded) {
    ctx->fill_lo = ctx->aux_begin(ctx->rank);
    ctx->fill_hi = ctx->aux_begin(ctx->rank + 1);
    ensure_x2(ctx);
    ctx->x2_epoch++;
    const size_t n = (size_t)(ctx->fill_hi - ctx->fill_lo) * ctx->ldx * ctx->world;
    GW_CUDA(cudaMemsetAsync(ctx->X2, 0, sizeof(double) * n, ctx->stream));
  }
  GW_API_END(ctx)
}

int gwbse_mmn_fill_end(gwbse_ctx* ctx) {
  GW_API_BEGIN(ctx)
  GW_PROF(ctx, "mmn_fill_exchange");
  if (ctx->fill_sharded) {
    const int world = ctx->world;
    const long long seg = (long long)(ctx->fill_hi - ctx->fill_lo) * ctx->ldx;
    std::vector<const double*> send(world);
    std::vector<double*> recv(world);
    std::vector<size_t> ns(world), nr(world);
    for (int r = 0; r < world; ++r) {
      send[r] = ctx->X2 + (long long)r * seg;
      ns[r] = (size_t)seg;
      const int lo = ctx->aux_begin(r), hi = ctx->aux_begin(r + 1);
      recv[r] = ctx->X + (long long)lo * ctx->ldx;
      nr[r] = (size_t)(hi - lo) * ctx->ldx;
    }
    alltoallv_dev(ctx, send.data(), ns.data(), recv.data(), nr.data());
    ctx->fill_sharded = false;
    ctx->mmn_version++;
  }
  GW_API_END(ctx)
}

int gwbse_mmn_mul_right_dev(gwbse_ctx* ctx, const double* R_dev, int ldr) {
  GW_API_BEGIN(ctx)
  GW_PROF(ctx, "mmn_mul_right");
  mul_right_dev(ctx, R_dev, ldr);
  GW_API_END(ctx)
}

int gwbse_mmn_mul_right_window_dev(gwbse_ctx* ctx, const double* R_dev, int ldr, int n_lo, int n_hi) {
  GW_API_BEGIN(ctx)
  GW_PROF(ctx, "mmn_mul_right");
  require_mmn(ctx);  // completes an earlier pending rotation: rotations compose in order
  GW_REQUIRE(ldr >= ctx->naux, "Shape mismatch in MultiplyRight");
  GW_REQUIRE(n_lo >= 0 && n_hi <= ctx->ntotal && n_lo < n_hi, "invalid row window");
  if (n_lo == 0 && n_hi == ctx->ntotal) {
    mul_right_dev(ctx, R_dev, ldr);
  } else if (ctx->ldx != 0) {
    auto& pr = ctx->pending_rot;
    pr.ld = round_up(ctx->naux, 2);
    pr.R = ctx->buf("pending_R", (size_t)pr.ld * ctx->naux);
    GW_CUDA(cudaMemcpy2DAsync(pr.R, sizeof(double) * pr.ld, R_dev, sizeof(double) * ldr, sizeof(double) * ctx->naux,
                              ctx->naux, cudaMemcpyDeviceToDevice, ctx->stream));
    // an even first row keeps the window's operand 16-byte aligned (TMA-describable)
    n_lo &= ~1;
    rotate_rows(ctx, pr.R, pr.ld, n_lo, n_hi);
    pr.n_lo = n_lo;
    pr.n_hi = n_hi;
    pr.active = true;
    ctx->mmn_version++;
    ctx->sig_ppm.ready = false;
  }
  GW_API_END(ctx)
}

int gwbse_mmn_mul_right(gwbse_ctx* ctx, const double* R, int ldr) {
  GW_API_BEGIN(ctx)
  GW_PROF(ctx, "mmn_mul_right_h2d");
  require_mmn(ctx);
  GW_REQUIRE(ldr >= ctx->naux, "Shape mismatch in MultiplyRight");
  double* Rd = ctx->buf("mulright_R", (size_t)ctx->naux * ctx->naux);
  GW_CUDA(copy2d_async(Rd, sizeof(double) * ctx->naux, R, sizeof(double) * ldr, sizeof(double) * ctx->naux,
                            ctx->naux, cudaMemcpyHostToDevice, ctx->stream));
  mul_right_dev(ctx, Rd, ctx->naux);
  GW_CUDA(cudaStreamSynchronize(ctx->stream));
  GW_API_END(ctx)
}

int gwbse_mmn_get_slice(gwbse_ctx* ctx, int m, double* out, int ld) {
  GW_API_BEGIN(ctx)
  require_mmn(ctx);
  GW_REQUIRE(m >= 0 && m < ctx->mtotal && ctx->owns(m), "slice index not owned by this rank");
  GW_REQUIRE(ld >= ctx->ntotal, "leading dimension too small");
  const double* src = ctx->X + (long long)ctx->local_index(m) * ctx->npad;
  GW_CUDA(copy2d_async(out, sizeof(double) * ld, src, sizeof(double) * ctx->ldx, sizeof(double) * ctx->ntotal,
                            ctx->naux, cudaMemcpyDeviceToHost, ctx->stream));
  GW_CUDA(cudaStreamSynchronize(ctx->stream));
  GW_API_END(ctx)
}

int gwbse_mmn_set_slice(gwbse_ctx* ctx, int m, const double* in, int ld) {
  GW_API_BEGIN(ctx)
  require_mmn(ctx);
  GW_REQUIRE(m >= 0 && m < ctx->mtotal && ctx->owns(m), "slice index not owned by this rank");
  GW_REQUIRE(ld >= ctx->ntotal, "leading dimension too small");
  double* dst = ctx->X + (long long)ctx->local_index(m) * ctx->npad;
  GW_CUDA(copy2d_async(dst, sizeof(double) * ctx->ldx, in, sizeof(double) * ld, sizeof(double) * ctx->ntotal,
                            ctx->naux, cudaMemcpyHostToDevice, ctx->stream));
  GW_CUDA(cudaStreamSynchronize(ctx->stream));
  ctx->mmn_version++;
  GW_API_END(ctx)
}

int gwbse_mmn_rotate(gwbse_ctx* ctx, const double* U, int ldu, int qpmin, int qpmax) {
  GW_API_BEGIN(ctx)
  GW_PROF(ctx, "mmn_rotate");
  require_mmn(ctx);
  const int q = qpmax - qpmin + 1;
  GW_REQUIRE(q > 0 && ldu >= q, "invalid rotation matrix");
  GW_REQUIRE(qpmin >= ctx->nmin && qpmax <= ctx->nmax && qpmin >= ctx->mmin && qpmax <= ctx->mmax,
             "QP window outside the Mmn ranges");
  const int row0 = qpmin - ctx->nmin, s0 = qpmin - ctx->mmin;
  GW_REQUIRE(ctx->naux <= 65535, "too many aux functions for the rotation scatter");
  double* Ud = ctx->buf("rotate_U", (size_t)q * q);
  GW_CUDA(copy2d_async(Ud, sizeof(double) * q, U, sizeof(double) * ldu, sizeof(double) * q, q, cudaMemcpyHostToDevice,
                       ctx->stream));
  const int nloc = ctx->owned_count(s0, q, ctx->rank);
  if (nloc > 0) {
    const int lfirst = ctx->local_index(ctx->first_owned(s0, ctx->rank));
    // T[n', (il, chi)] = sum_n U[n, n'] X[chi][lfirst + il][row0 + n]
    double* T = ctx->buf("rotate_T", (size_t)q * nloc * ctx->naux);
    GemmParams p;
    p.M = q;
    p.N = nloc * ctx->naux;
    p.Ki = q;
    p.A.ptr = Ud;
    p.A.s_ri = q;
    p.A.s_ki = 1;
    p.B.ptr = ctx->X + (long long)lfirst * ctx->npad + row0;
    p.B.Lr = nloc;
    p.B.s_ri = ctx->npad;
    p.B.s_ro = ctx->ldx;
    p.B.s_ki = 1;
    p.C = T;
    p.sC_mi = 1;
    p.sC_ni = q;
    ctx->gemm(p);
    launch_rotate_scatter(T, q, nloc, ctx->naux, ctx->X, ctx->ldx, ctx->npad, lfirst, row0, ctx->stream);
    ctx->launches++;
  }
  ctx->mmn_version++;
  ctx->sig_ppm.ready = false;
  GW_CUDA(cudaStreamSynchronize(ctx->stream));
  GW_API_END(ctx)
}

// RPA::setQSGWRotation (rpa.h:59-66) / Sigma_base::setQSGWRotation (sigma_base.h:51-58): U == NULL clears it
int gwbse_rpa_set_qsgw_rotation(gwbse_ctx* ctx, const double* U, int ldu, int qptotal, int qpmin, int homo) {
  GW_API_BEGIN(ctx)
  auto& qs = ctx->qsgw;
  qs.built_version = -1;
  if (!U) {
    qs.active = false;
  } else {
    GW_REQUIRE(qptotal > 0 && ldu >= qptotal, "invalid QSGW rotation");
    qs.U = ctx->buf("qsgw_U", (size_t)qptotal * qptotal);
    GW_CUDA(copy2d_async(qs.U, sizeof(double) * qptotal, U, sizeof(double) * ldu, sizeof(double) * qptotal, qptotal,
                         cudaMemcpyHostToDevice, ctx->stream));
    GW_CUDA(cudaStreamSynchronize(ctx->stream));
    qs.qptotal = qptotal;
    qs.qpmin = qpmin;
    qs.homo = homo;
    qs.active = true;
  }
  GW_API_END(ctx)
}

int gwbse_mmn_snapshot(gwbse_ctx* ctx) {
  GW_API_BEGIN(ctx)
  GW_PROF(ctx, "mmn_snapshot");
  require_mmn(ctx);
  const size_t bytes = sizeof(double) * (size_t)std::max<long long>(ctx->ldx, 1) * ctx->naux;
  if (!ctx->Xsnap) GW_CUDA(cudaMalloc(&ctx->Xsnap, bytes));
  GW_CUDA(cudaMemcpyAsync(ctx->Xsnap, ctx->X, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
  GW_API_END(ctx)
}

int gwbse_mmn_restore(gwbse_ctx* ctx) {
  GW_API_BEGIN(ctx)
  GW_PROF(ctx, "mmn_restore");
  ctx->pending_rot.active = false;  // the whole tensor is replaced
  require_mmn(ctx);
  GW_REQUIRE(ctx->Xsnap != nullptr, "no Mmn snapshot to restore");
  const size_t bytes = sizeof(double) * (size_t)std::max<long long>(ctx->ldx, 1) * ctx->naux;
  GW_CUDA(cudaMemcpyAsync(ctx->X, ctx->Xsnap, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
  ctx->mmn_version++;
  ctx->sig_ppm.ready = false;
  GW_API_END(ctx)
}

// AOCoulomb::Pseudo_InvSqrt_GWBSE, aomatrix.cc:53-86: two symmetric eigensolves + GEMMs
int gwbse_pseudo_invsqrt(gwbse_ctx* ctx, int n, const double* S, const double* V, double etol, double* L_out,
                         int* removed) {
  GW_API_BEGIN(ctx)
  GW_PROF(ctx, "pseudo_invsqrt");
  const size_t nn = (size_t)n * n;
  double* dS = ctx->buf("pis_S", nn);
  double* dV = ctx->buf("pis_V", nn);
  double* dT = ctx->buf("pis_T", nn);
  double* dSs = ctx->buf("pis_Ss", nn);
  double* dd = ctx->buf("pis_d", n);
  int* drem = reinterpret_cast<int*>(ctx->buf("pis_rem", 8));
  std::vector<double> w(n);
  GW_CUDA(cudaMemcpyAsync(dS, S, sizeof(double) * nn, cudaMemcpyHostToDevice, ctx->stream));
  GW_CUDA(cudaMemcpyAsync(dV, V, sizeof(double) * nn, cudaMemcpyHostToDevice, ctx->stream));
  GW_CUDA(cudaMemsetAsync(drem, 0, sizeof(int), ctx->stream));
  if (gwbse_sym_eig_dev(ctx, n, dS, n, w.data())) throw std::runtime_error(ctx->err);
  double* wd = ctx->buf("eig_w", n);  // eigenvalues still on the device
  launch_invsqrt_scale(dd, wd, n, etol, drem, ctx->stream);
  // Ssqrt = U diag(d) U^T
  launch_diag_scale('R', n, n, dS, n, dd, dT, n, ctx->stream);
  if (gwbse_dgemm_dev(ctx, 'N', 'T', n, n, n, 1.0, dT, n, dS, n, 0.0, dSs, n)) throw std::runtime_error(ctx->err);
  // ortho = Ssqrt V Ssqrt
  if (gwbse_dgemm_dev(ctx, 'N', 'N', n, n, n, 1.0, dSs, n, dV, n, 0.0, dT, n)) throw std::runtime_error(ctx->err);
  if (gwbse_dgemm_dev(ctx, 'N', 'N', n, n, n, 1.0, dT, n, dSs, n, 0.0, dV, n)) throw std::runtime_error(ctx->err);
  if (gwbse_sym_eig_dev(ctx, n, dV, n, w.data())) throw std::runtime_error(ctx->err);
  launch_invsqrt_scale(dd, wd, n, etol, drem, ctx->stream);
  // Vm1 = W diag(d2) W^T ; result = (Vm1 Ssqrt)^T = Ssqrt Vm1  (both symmetric)
  launch_diag_scale('R', n, n, dV, n, dd, dT, n, ctx->stream);
  if (gwbse_dgemm_dev(ctx, 'N', 'T', n, n, n, 1.0, dT, n, dV, n, 0.0, dS, n)) throw std::runtime_error(ctx->err);
  if (gwbse_dgemm_dev(ctx, 'N', 'N', n, n, n, 1.0, dSs, n, dS, n, 0.0, dT, n)) throw std::runtime_error(ctx->err);
  ctx->launches += 4;
  int hrem = 0;
  if (L_out) GW_CUDA(cudaMemcpyAsync(L_out, dT, sizeof(double) * nn, cudaMemcpyDeviceToHost, ctx->stream));
  GW_CUDA(cudaMemcpyAsync(&hrem, drem, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  GW_CUDA(cudaStreamSynchronize(ctx->stream));
  if (removed) *removed = hrem;
  GW_API_END(ctx)
}

const double* gwbse_pseudo_invsqrt_result_dev(gwbse_ctx* ctx) {
  if (!ctx) return nullptr;
  auto it = ctx->bufs.find("pis_T");
  return it == ctx->bufs.end() ? nullptr : it->second.p;
}

// ------------------------------- RPA ---------------------------------------
int gwbse_rpa_epsilon(gwbse_ctx* ctx, int kind, double fre, double fim, double eta, const double* energies,
                      int homo, int rpamin, int rpamax, double* eps_out, int ld) {
  GW_API_BEGIN(ctx)
  GW_PROF(ctx, "rpa_epsilon");
  require_mmn(ctx);
  GW_REQUIRE(kind >= 0 && kind <= 2, "epsilon kind must be 0 (imag), 1 (real) or 2 (complex)");
  GW_REQUIRE(rpamin == ctx->mmin && rpamin == ctx->nmin && rpamax == ctx->nmax,
             "RPA range must match the Mmn n-range and start at mmin");
  const int n_occ = homo + 1 - rpamin;
  const int n_unocc = rpamax - homo;
  GW_REQUIRE(n_occ > 0 && n_unocc > 0 && n_occ <= ctx->mtotal, "invalid occupied/virtual split");
  const int naux = ctx->naux;
  ctx->eps = ctx->buf("rpa_eps", (size_t)naux * naux);
  const int rpatotal = n_occ + n_unocc;
  double* e_dev = ctx->buf("rpa_e", rpatotal);
  GW_CUDA(cudaMemcpyAsync(e_dev, energies, sizeof(double) * rpatotal, cudaMemcpyHostToDevice, ctx->stream));
  const int nloc_occ = ctx->local_count(n_occ);
  const long long ldw = (n_unocc + 1) & ~1;
  double* w = ctx->buf("rpa_w", (size_t)std::max(nloc_occ, 1) * ldw);
  launch_rpa_weights(w, ldw, e_dev, kind, fre, fim, eta, n_occ, n_unocc, ctx->rank, ctx->world, nloc_occ, ctx->stream);
  ctx->launches++;
  if (nloc_occ > 0) {
    GemmParams p;
    p.M = naux;
    p.N = naux;
    p.Ko = nloc_occ;
    p.Ki = n_unocc;
    const HoleView hv = hole_view(ctx, n_occ);
    p.A.ptr = hv.ptr;
    p.A.s_ri = hv.s_chi;
    p.A.s_ki = 1;
    p.A.s_ko = hv.s_v;
    p.B = p.A;
    p.C = ctx->eps;
    p.sC_mi = 1;
    p.sC_ni = naux;
    p.alpha = (kind == 2) ? -2.0 : 1.0;
    p.w = w;
    p.sW_ko = ldw;
    p.lower_only = 1;
    ctx->gemm(p);
    launch_symmetrize_lower(ctx->eps, naux, naux, ctx->stream);
    ctx->launches++;
  } else {
    GW_CUDA(cudaMemsetAsync(ctx->eps, 0, sizeof(double) * (size_t)naux * naux, ctx->stream));
  }
  if (ctx->world > 1) {
    allreduce_dev(ctx, ctx->eps, (size_t)naux * naux);
  }
  launch_add_diagonal(ctx->eps, naux, naux, 1.0, ctx->stream);
  ctx->launches++;
  if (eps_out) {
    GW_REQUIRE(ld >= naux, "leading dimension too small");
    GW_CUDA(copy2d_async(eps_out, sizeof(double) * ld, ctx->eps, sizeof(double) * naux, sizeof(double) * naux,
                              naux, cudaMemcpyDeviceToHost, ctx->stream));
  }
  GW_CUDA(cudaStreamSynchronize(ctx->stream));
  GW_API_END(ctx)
}

double* gwbse_rpa_epsilon_ptr(gwbse_ctx* ctx) { return ctx ? ctx->eps : nullptr; }

// (A+B)[v1 c1, v2 c2] = 4 sum_chi M[v1][c1,chi] M[v2][c2,chi] + delta (e_c - e_v), rpa.cc:281-326
int gwbse_rpa_h2p_apb(gwbse_ctx* ctx, const double* energies, int homo, int rpamin, int rpamax, double* apb_dev,
                      int ld) {
  GW_API_BEGIN(ctx)
  GW_PROF(ctx, "rpa_h2p_apb");
  require_mmn(ctx);
  GW_REQUIRE(ctx->world == 1, "H2p is single-GPU (exact sigma does not scale, SURVEY.md 8e)");
  GW_REQUIRE(rpamin == ctx->mmin && rpamin == ctx->nmin && rpamax == ctx->nmax, "RPA range must match Mmn");
  const int n_occ = homo + 1 - rpamin, n_unocc = rpamax - homo;
  const int S = n_occ * n_unocc;
  GW_REQUIRE(ld >= S, "leading dimension too small");
  // rows (v, c): compound index with inner length n_unocc
  GemmParams p;
  p.M = S;
  p.N = S;
  p.Ki = ctx->naux;
  const HoleView hv = hole_view(ctx, n_occ);
  p.A.ptr = hv.ptr;
  p.A.Lr = n_unocc;
  p.A.s_ri = 1;
  p.A.s_ro = hv.s_v;
  p.A.s_ki = hv.s_chi;
  p.B = p.A;
  p.C = apb_dev;
  p.sC_mi = 1;
  p.sC_ni = ld;
  p.alpha = 4.0;
  p.lower_only = 1;
  ctx->gemm(p);
  launch_symmetrize_lower(apb_dev, S, ld, ctx->stream);
  // + diag(AmB)
  std::vector<double> amb(S);
  for (int v = 0; v < n_occ; ++v)
    for (int c = 0; c < n_unocc; ++c) amb[(size_t)v * n_unocc + c] = energies[n_occ + c] - energies[v];
  double* d = ctx->buf("h2p_amb", S);
  GW_CUDA(cudaMemcpyAsync(d, amb.data(), sizeof(double) * S, cudaMemcpyHostToDevice, ctx->stream));
  // axpy on the diagonal viewed as a strided vector: 1 x S block with ld+1 stride
  launch_axpy(1, S, 1.0, d, 1, apb_dev, (long long)ld + 1, ctx->stream);
  ctx->launches += 2;
  GW_CUDA(cudaStreamSynchronize(ctx->stream));
  GW_API_END(ctx)
}

// One block of the unrestricted A+B matrix (rpa_uks.cc:475-540: alpha-alpha, beta-beta and the mixed block, all
// with the same prefactor): rows are the particle-hole pairs of `ctx`, columns those of `other` (same GPU; may be
// the same context).  The caller adds diag(AmB) and owns the layout of the combined matrix.
int gwbse_rpa_h2p_block(gwbse_ctx* ctx, gwbse_ctx* other, int homo, int homo_other, int rpamin, int rpamax,
                        double alpha, double* block_dev, int ld) {
  GW_API_BEGIN(ctx)
  GW_PROF(ctx, "rpa_h2p_block");
  require_mmn(ctx);
  GW_REQUIRE(other && other->X, "the other spin channel has no Mmn");
  mmn_complete_rotation(other);
  GW_REQUIRE(ctx->world == 1 && other->world == 1, "H2p is single-GPU (exact sigma does not scale, SURVEY.md 8e)");
  GW_REQUIRE(other->device == ctx->device && other->naux == ctx->naux, "both channels live on one GPU with one aux basis");
  for (const gwbse_ctx* c : {ctx, other})
    GW_REQUIRE(rpamin == c->mmin && rpamin == c->nmin && rpamax == c->nmax, "RPA range must match Mmn");
  const int n_occ = homo + 1 - rpamin, n_unocc = rpamax - homo;
  const int n_occ_o = homo_other + 1 - rpamin, n_unocc_o = rpamax - homo_other;
  GW_REQUIRE(n_occ > 0 && n_unocc > 0 && n_occ_o > 0 && n_unocc_o > 0, "empty particle-hole space");
  GW_REQUIRE(ld >= n_occ * n_unocc, "leading dimension too small");
  GemmParams p;
  p.M = n_occ * n_unocc;
  p.N = n_occ_o * n_unocc_o;
  p.Ki = ctx->naux;
  const HoleView hv = hole_view(ctx, n_occ), ho = hole_view(other, n_occ_o);
  p.A.ptr = hv.ptr;
  p.A.Lr = n_unocc;
  p.A.s_ri = 1;
  p.A.s_ro = hv.s_v;
  p.A.s_ki = hv.s_chi;
  p.B.ptr = ho.ptr;
  p.B.Lr = n_unocc_o;
  p.B.s_ri = 1;
  p.B.s_ro = ho.s_v;
  p.B.s_ki = ho.s_chi;
  p.C = block_dev;
  p.sC_mi = 1;
  p.sC_ni = ld;
  p.alpha = alpha;
  ctx->gemm(p);
  GW_CUDA(cudaStreamSynchronize(ctx->stream));
  GW_API_END(ctx)
}

// ------------------------------ Sigma_x -------------------------------------
int gwbse_sigma_x(gwbse_ctx* ctx, int homo, int rpamin, int qpmin, int qpmax, double* out, int ld) {
  GW_API_BEGIN(ctx)
  GW_PROF(ctx, "sigma_x");
  require_mmn(ctx);
  GW_REQUIRE(rpamin == ctx->mmin && rpamin == ctx->nmin, "RPA range must match Mmn");
  const int q = qpmax - qpmin + 1;
  const int occ = homo - rpamin + 1;
  const int qpoff = qpmin - rpamin;
  GW_REQUIRE(q > 0 && qpoff >= 0 && qpoff + q <= ctx->mtotal && occ > 0 && occ <= ctx->ntotal, "invalid qp window");
  GW_REQUIRE(ld >= q, "leading dimension too small");
  double* sx = ctx->buf("sigma_x", (size_t)q * q);
  GemmParams p;
  p.M = q;
  p.N = q;
  p.Ko = ctx->naux;
  p.Ki = occ;
  if (ctx->world == 1) {
    p.A.ptr = ctx->X + (long long)qpoff * ctx->npad;
    p.A.s_ri = ctx->npad;
    p.A.s_ko = ctx->ldx;
  } else {
    // occupied-row panels of every qp slice, replicated on all ranks (q * n_occ * naux doubles)
    const int rpad = (occ + 1) & ~1;
    const long long ldo = (long long)q * rpad;
    double* panel = ctx->buf("sigma_x_panel", (size_t)ldo * ctx->naux);
    gather_slices(ctx, qpoff, q, 0, occ, 0, ctx->naux, panel, ldo, rpad);
    p.A.ptr = panel;
    p.A.s_ri = rpad;
    p.A.s_ko = ldo;
  }
  p.A.s_ki = 1;
  p.B = p.A;
  p.C = sx;
  p.sC_mi = 1;
  p.sC_ni = q;
  p.alpha = -1.0;
  p.lower_only = 1;
  ctx->gemm(p);
  launch_symmetrize_lower(sx, q, q, ctx->stream);
  ctx->launches++;
  GW_CUDA(copy2d_async(out, sizeof(double) * ld, sx, sizeof(double) * q, sizeof(double) * q, q,
                            cudaMemcpyDeviceToHost, ctx->stream));
  GW_CUDA(cudaStreamSynchronize(ctx->stream));
  GW_API_END(ctx)
}

}  // extern "C"
