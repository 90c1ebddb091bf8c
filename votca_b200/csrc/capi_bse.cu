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
#include <cstdlib>
#include <vector>

#include "../../include/gwbse_b200.h"
#include "context.cuh"

using namespace gwbse;

namespace {

inline int round_up(int v, int a) { return (v + a - 1) / a * a; }

// replicated copies of the occupied-occupied and virtual-occupied blocks when Mmn is sharded:
//   vv[chi][v1][v2] = M[voff+v1][voff+v2, chi],  cv[chi][c1][v2] = M[coff+c1][voff+v2, chi]
struct BlockView {
  const double* ptr;
  long long row, pole;  // strides of the slice index and of chi
};

void ensure_gathered(gwbse_ctx* ctx) {
  auto& st = ctx->bse;
  if (ctx->world == 1) return;
  if (st.gathered_version == ctx->mmn_version && st.gathered_voff == st.voff && st.gathered_vt == st.vt &&
      st.gathered_ct == st.ct)
    return;
  const int vtp = round_up(st.vt, 2);
  double* gvv = ctx->buf("bse_gvv", (size_t)ctx->naux * st.vt * vtp);
  double* gcv = ctx->buf("bse_gcv", (size_t)ctx->naux * st.ct * vtp);
  gather_slices(ctx, st.voff, st.vt, st.voff, st.vt, 0, ctx->naux, gvv, (long long)st.vt * vtp, vtp);
  gather_slices(ctx, st.coff, st.ct, st.voff, st.vt, 0, ctx->naux, gcv, (long long)st.ct * vtp, vtp);
  st.gathered_version = ctx->mmn_version;
  st.gathered_voff = st.voff;
  st.gathered_vt = st.vt;
  st.gathered_ct = st.ct;
}

BlockView vv_view(gwbse_ctx* ctx) {
  auto& st = ctx->bse;
  if (ctx->world == 1) return {ctx->X + (long long)st.voff * ctx->npad + st.voff, ctx->npad, ctx->ldx};
  const int vtp = round_up(st.vt, 2);
  return {ctx->buf("bse_gvv", (size_t)ctx->naux * st.vt * vtp), vtp, (long long)st.vt * vtp};
}
BlockView cv_view(gwbse_ctx* ctx) {
  auto& st = ctx->bse;
  if (ctx->world == 1) return {ctx->X + (long long)st.coff * ctx->npad + st.voff, ctx->npad, ctx->ldx};
  const int vtp = round_up(st.vt, 2);
  return {ctx->buf("bse_gcv", (size_t)ctx->naux * st.ct * vtp), vtp, (long long)st.ct * vtp};
}

// The operator reads rows n in [voff, coff + ct) of every slice only.  A rotation that is still pending outside a
// window (gwbse_mmn_mul_right_window_dev) has to be completed only if that window does not cover them.
void bse_rows_ready(gwbse_ctx* ctx) {
  const auto& pr = ctx->pending_rot;
  if (!pr.active) return;
  const auto& st = ctx->bse;
  if (!st.ready || st.voff < pr.n_lo || st.coff + st.ct > pr.n_hi) mmn_complete_rotation(ctx);
}

// The two halves of the exchange term, also the building blocks of the cross-spin coupling of the unrestricted
// operator (BSE_OPERATOR_UKS::add_direct_cross_tda_block, bse_operator_uks.cc:174-211), where the projection of one
// spin channel is expanded in the other:
//   project: W[chi, kv]     = sum_{v, c} M[v][c, chi] X[(v, c), kv]          (all-reduced over the ranks)
//   expand : Y[(v, c), kv] += alpha sum_chi M[v][c, chi] W[chi, kv]           (rows of the v slices this rank owns)
void vc_project(gwbse_ctx* ctx, int k, const double* Xin, int ldin, double* W) {
  auto& st = ctx->bse;
  const int ct = st.ct, naux = ctx->naux, npad = ctx->npad, world = ctx->world;
  const int nvloc = ctx->owned_count(st.voff, st.vt, ctx->rank);
  const int v_rel0 = ctx->first_owned(st.voff, ctx->rank) - st.voff;
  const int lvfirst = nvloc ? ctx->local_index(st.voff + v_rel0) : 0;
  if (nvloc > 0) {
    GemmParams p;
    p.M = naux;
    p.N = k;
    p.Ko = nvloc;
    p.Ki = ct;
    p.A.ptr = ctx->X + (long long)lvfirst * npad + st.coff;
    p.A.s_ri = ctx->ldx;
    p.A.s_ki = 1;
    p.A.s_ko = npad;
    p.B.ptr = Xin + (long long)v_rel0 * ct;
    p.B.s_ri = ldin;
    p.B.s_ki = 1;
    p.B.s_ko = (long long)world * ct;
    p.C = W;
    p.sC_mi = 1;
    p.sC_ni = naux;
    ctx->gemm(p);
  } else {
    GW_CUDA(cudaMemsetAsync(W, 0, sizeof(double) * (size_t)naux * k, ctx->stream));
  }
  allreduce_dev(ctx, W, (size_t)naux * k);
}

void vc_expand(gwbse_ctx* ctx, double alpha, int k, const double* W, double* Y, int ldy) {
  auto& st = ctx->bse;
  const int ct = st.ct, naux = ctx->naux, npad = ctx->npad, world = ctx->world;
  const int nvloc = ctx->owned_count(st.voff, st.vt, ctx->rank);
  const int v_rel0 = ctx->first_owned(st.voff, ctx->rank) - st.voff;
  const int lvfirst = nvloc ? ctx->local_index(st.voff + v_rel0) : 0;
  if (nvloc <= 0) return;
  GemmParams q;
  q.M = nvloc * ct;
  q.N = k;
  q.Ki = naux;
  q.A.ptr = ctx->X + (long long)lvfirst * npad + st.coff;
  q.A.Lr = ct;
  q.A.s_ri = 1;
  q.A.s_ro = npad;
  q.A.s_ki = ctx->ldx;
  q.B.ptr = W;
  q.B.s_ri = naux;
  q.B.s_ki = 1;
  q.C = Y + (long long)v_rel0 * ct;
  q.Lm = ct;
  q.sC_mi = 1;
  q.sC_mo = (long long)world * ct;
  q.sC_ni = ldy;
  q.alpha = alpha;
  q.beta = 1.0;
  ctx->gemm(q);
}

// ------------------------------------------------------------------------------------------------------------
// Materialised direct-interaction blocks.  The factorised products above cost 2 Naux B (vt + ct) (Hd) and
// 4 Naux vt^2 ct (Hd2) flops per trial vector; forming the block once costs 2 B^2 Naux, i.e. as much as
// B / (vt + ct) resp. ct / 2 vectors - a Davidson solve applies the operator to ten times that many.  With 180 GB of
// HBM the B x B blocks of a C60-sized problem (33 GB each) stay resident, and a product becomes one skinny GEMM that
// streams the block once.  Rows of the result a rank owns (c-slices for Hd, v-slices for Hd2, as in the factorised
// path) are the columns of the block it stores:
//   Hd : H[(v2, c2), (v1, l)] = sum_chi M[v1][v2, chi] eps_inv[chi] M[c1(l)][c2, chi]
//   Hd2: H[(v2, c2), (l, c1)] = sum_chi M[v1(l)][c2, chi] eps_inv[chi] M[c1][v2, chi]
// built by one GEMM each over chi from compact (even-pitch, TMA-describable) copies of the two-index blocks, the
// screening folded into one of the copies.  A block is valid for one (Mmn version, screening, level window); the
// first one is parked in the second Mmn buffer (idle between MultiplyRight calls), the second one is allocated if
// the device has the room, otherwise that term stays factorised.
struct OwnedSlices {
  int nvloc, v_rel0, lvfirst, ncloc, c_rel0, lcfirst;
};
OwnedSlices owned_slices(gwbse_ctx* ctx) {
  auto& st = ctx->bse;
  OwnedSlices o;
  o.nvloc = ctx->owned_count(st.voff, st.vt, ctx->rank);
  o.v_rel0 = ctx->first_owned(st.voff, ctx->rank) - st.voff;
  o.lvfirst = o.nvloc ? ctx->local_index(st.voff + o.v_rel0) : 0;
  o.ncloc = ctx->owned_count(st.coff, st.ct, ctx->rank);
  o.c_rel0 = ctx->first_owned(st.coff, ctx->rank) - st.coff;
  o.lcfirst = o.ncloc ? ctx->local_index(st.coff + o.c_rel0) : 0;
  return o;
}

double factorised_flops_per_column(gwbse_ctx* ctx, int kind) {
  auto& st = ctx->bse;
  return kind == 0 ? 2.0 * ctx->naux * (double)st.vt * st.ct * (st.vt + st.ct)
                   : 4.0 * (double)st.vt * st.vt * st.ct * ctx->naux;
}

// measurement knob (scratch/ncu_bse_dense.py): tile shape of the build GEMM, -1 = the planner's choice
int dense_build_cfg() {
  static const int cfg = [] {
    const char* e = std::getenv("GWBSE_DENSE_BUILD_CFG");
    return e ? std::atoi(e) : -1;
  }();
  return cfg;
}

bool dense_build(gwbse_ctx* ctx, int kind) {
  auto& st = ctx->bse;
  auto& blk = st.dense[kind];
  const int vt = st.vt, ct = st.ct, B = st.size, naux = ctx->naux, npad = ctx->npad;
  const OwnedSlices o = owned_slices(ctx);
  // even (16-byte rows for TMA), and never a multiple of 2 KiB: columns 2^k bytes apart would land on the same
  // L2 slices / HBM channels (a B = 8192 block ran its product 30x slower, profiles/r02_bse_block_build_tiles.txt)
  long long ld = round_up(B, 2);
  if (ld % 256 == 0) ld += 2;
  const long long ncols = kind == 0 ? (long long)vt * o.ncloc : (long long)o.nvloc * ct;
  blk.ld = ld;
  blk.ncols = ncols;
  blk.H = nullptr;
  blk.in_x2 = false;
  if (ncols == 0) {  // this rank owns no row of the result
    blk.valid = true;
    ctx->bse_dense_builds++;
    ctx->bse_algo_flops += 2.0 * (double)B * B * naux;
    return true;
  }
  if (ncols >= (1LL << 31) || (long long)vt * vt >= (1LL << 31) || (long long)ct * ct >= (1LL << 31)) return false;
  // ---- operand copies: the replicated small block whole, the other one in chunks of local slices
  const int nloc = kind == 0 ? o.ncloc : o.nvloc;               // chunked index l
  const long long small_rows = kind == 0 ? (long long)vt * vt : (long long)ct * vt;
  const long long small_plane = round_up((int)small_rows, 2);
  const size_t pack_budget = (size_t)4 << 30;
  int lchunk = (int)std::max<long long>(1, (long long)(pack_budget / sizeof(double)) / ((long long)ct * naux));
  lchunk = std::min(lchunk, nloc);
  const long long big_plane = round_up(lchunk * ct, 2);
  const size_t need_small = (size_t)small_plane * naux, need_big = (size_t)big_plane * naux;
  const size_t need_H = (size_t)ld * (size_t)ncols;
  // ---- where the block lives
  const auto& other = st.dense[1 - kind];
  const bool x2_taken = other.valid && other.in_x2 && other.x2_epoch == ctx->x2_epoch;
  const size_t x2_cap = ctx->X2 ? (size_t)ctx->ldx * (size_t)(naux + ctx->world) : 0;
  const char* own_name = kind == 0 ? "bse_dense0" : "bse_dense1";
  auto held = [&](const char* name) {
    auto it = ctx->bufs.find(name);
    return it == ctx->bufs.end() ? (size_t)0 : it->second.cap;
  };
  size_t extra = 0;  // doubles that would have to be newly allocated
  if (held("bse_packA") < need_small) extra += need_small;
  if (held("bse_packB") < need_big) extra += need_big;
  const bool use_x2 = !x2_taken && x2_cap >= need_H;
  if (!use_x2 && held(own_name) < need_H) extra += need_H;
  {
    size_t free_b = 0, total_b = 0;
    GW_CUDA(cudaMemGetInfo(&free_b, &total_b));
    // buffers that are re-allocated give their old size back first
    size_t back = 0;
    if (held("bse_packA") < need_small) back += held("bse_packA");
    if (held("bse_packB") < need_big) back += held("bse_packB");
    if (!use_x2 && held(own_name) < need_H) back += held(own_name);
    const size_t margin = std::max<size_t>((size_t)4 << 30, total_b / 32);
    if (extra * sizeof(double) + margin > free_b + back * sizeof(double)) return false;
  }
  double* packA = ctx->buf("bse_packA", need_small);
  double* packB = ctx->buf("bse_packB", need_big);
  double* H = use_x2 ? ctx->X2 : ctx->buf(own_name, need_H);
  if (use_x2) blk.x2_epoch = ++ctx->x2_epoch;
  const BlockView vv = vv_view(ctx), cv = cv_view(ctx);
  const double* X = ctx->X;
  if (kind == 0) {
    // packA[chi][(v1, v2)] = eps_inv[chi] M[v1][v2, chi]
    launch_pack_block(vv.ptr, vv.pole, vv.row, vt, vt, st.eps_inv, packA, small_plane, naux, ctx->stream);
  } else {
    // packA[chi][(c1, v2)] = eps_inv[chi] M[c1][v2, chi]
    launch_pack_block(cv.ptr, cv.pole, cv.row, ct, vt, st.eps_inv, packA, small_plane, naux, ctx->stream);
  }
  ctx->launches++;
  for (int a = 0; a < nloc; a += lchunk) {
    const int n1 = std::min(lchunk, nloc - a);
    // packB[chi][(l, c2)] = M[slice(a + l)][coff + c2, chi]: c-slices for Hd, v-slices for Hd2
    const long long lfirst = (kind == 0 ? o.lcfirst : o.lvfirst) + a;
    launch_pack_block(X + lfirst * npad + st.coff, ctx->ldx, npad, n1, ct, nullptr, packB, big_plane, naux,
                      ctx->stream);
    ctx->launches++;
    GemmParams p;
    p.Ki = naux;
    p.A.s_ri = 1;
    p.B.s_ri = 1;
    // either way the GEMM's row index carries c2, the contiguous index of the block: the 8 rows a warp stores per
    // instruction are one 64-byte run
    if (kind == 0) {
      // rows (l, c2), columns (v1, v2) -> H[(v2, c2), v1 * ncloc + a + l]
      p.M = n1 * ct;
      p.N = vt * vt;
      p.A.ptr = packB;
      p.A.s_ki = big_plane;
      p.B.ptr = packA;
      p.B.s_ki = small_plane;
      p.C = H + (long long)a * ld;
      p.Lm = ct;
      p.sC_mo = ld;
      p.sC_mi = 1;
      p.Ln = vt;
      p.sC_no = (long long)o.ncloc * ld;
      p.sC_ni = ct;
    } else {
      // rows (l, c2), columns (c1, v2) -> H[(v2, c2), (a + l) * ct + c1]
      p.M = n1 * ct;
      p.N = ct * vt;
      p.A.ptr = packB;
      p.A.s_ki = big_plane;
      p.B.ptr = packA;
      p.B.s_ki = small_plane;
      p.C = H + (long long)a * ct * ld;
      p.Lm = ct;
      p.sC_mo = (long long)ct * ld;
      p.sC_mi = 1;
      p.Ln = vt;
      p.sC_no = ld;
      p.sC_ni = ct;
    }
    // both operands M-major: the 128 x 128 tile of the TMA kernel runs 8-12 % above the planner's pick
    // (profiles/r02_bse_block_build_tiles.txt); small problems keep the planner
    int cfg = dense_build_cfg();
    if (cfg < 0 && p.M >= 512 && p.N >= 512) cfg = 10;
    ctx->gemm(p, cfg);
  }
  blk.H = H;
  blk.in_x2 = use_x2;
  blk.valid = true;
  ctx->bse_dense_builds++;
  ctx->bse_algo_flops += 2.0 * (double)B * B * naux;
  return true;
}

// true: the block of this kind is resident and current (built here if the policy says so)
bool dense_ready(gwbse_ctx* ctx, int kind, int k) {
  auto& st = ctx->bse;
  auto& blk = st.dense[kind];
  const bool same_key = blk.key_mmn == ctx->mmn_version && blk.key_eps == st.eps_version && blk.vt == st.vt &&
                        blk.ct == st.ct && blk.voff == st.voff && blk.coff == st.coff;
  if (!same_key) {
    blk.valid = false;
    blk.refused = false;
    blk.spent_flops = 0.0;
    blk.key_mmn = ctx->mmn_version;
    blk.key_eps = st.eps_version;
    blk.vt = st.vt;
    blk.ct = st.ct;
    blk.voff = st.voff;
    blk.coff = st.coff;
  }
  if (blk.valid && blk.in_x2 && blk.x2_epoch != ctx->x2_epoch) blk.valid = false;
  if (ctx->bse_dense_mode == 0) return false;
  if (blk.valid) return true;
  if (blk.refused) return false;
  const double per_col = factorised_flops_per_column(ctx, kind);
  if (ctx->bse_dense_mode == 1) {
    const double build = 2.0 * (double)st.size * st.size * ctx->naux;
    if (blk.spent_flops + per_col * k < ctx->bse_dense_payback * build) {
      blk.spent_flops += per_col * k;
      return false;
    }
  }
  if (!dense_build(ctx, kind)) {
    blk.refused = true;  // no room under this key: stay factorised without asking again
    return false;
  }
  return true;
}

// Y[owned rows] += alpha H^T X
void dense_apply(gwbse_ctx* ctx, int kind, double alpha, int k, const double* Xin, int ldin, double* Y, int ldy) {
  auto& st = ctx->bse;
  const auto& blk = st.dense[kind];
  if (blk.ncols == 0) return;
  const OwnedSlices o = owned_slices(ctx);
  GemmParams p;
  p.M = (int)blk.ncols;
  p.N = k;
  p.Ki = st.size;
  p.A.ptr = blk.H;
  p.A.s_ri = blk.ld;
  p.A.s_ki = 1;
  p.B.ptr = Xin;
  p.B.s_ri = ldin;
  p.B.s_ki = 1;
  p.sC_ni = ldy;
  p.alpha = alpha;
  p.beta = 1.0;
  if (kind == 0) {
    // row v1 * ncloc + l -> Y[(v1, c_rel0 + l * world)]
    p.C = Y + o.c_rel0;
    p.Lm = o.ncloc;
    p.sC_mo = st.ct;
    p.sC_mi = ctx->world;
  } else {
    // row l * ct + c1 -> Y[(v_rel0 + l * world, c1)]
    p.C = Y + (long long)o.v_rel0 * st.ct;
    p.Lm = st.ct;
    p.sC_mo = (long long)ctx->world * st.ct;
    p.sC_mi = 1;
  }
  ctx->gemm(p);
  ctx->bse_dense_columns += k;
}

// Y = H X.  With a sharded Mmn every rank computes the part its slices contribute (v-slices for Hx / Hd2,
// c-slices for Hd, rank 0 the Hqp term) into disjoint or additive entries of Y, then Y is all-reduced.
void bse_matmul_dev(gwbse_ctx* ctx, int cqp, int cx, int cd, int cd2, int k, const double* Xin, int ldin, double* Y,
                    int ldy) {
  auto& st = ctx->bse;
  GW_REQUIRE(st.ready, "BSE operator not configured (gwbse_bse_configure)");
  GW_REQUIRE(!(cd != 0 && cd2 != 0), "Hamiltonian cannot contain Hd and Hd2 at the same time");
  const int vt = st.vt, ct = st.ct, B = st.size, naux = ctx->naux, npad = ctx->npad;
  const int voff = st.voff, coff = st.coff, world = ctx->world;
  const long long ldx = ctx->ldx;
  const double* X = ctx->X;
  GW_REQUIRE(ldin >= B && ldy >= B, "Shape mismatch in BSE matmul");
  if (k <= 0) return;
  bse_rows_ready(ctx);
  // work of the formulation that is executed (all ranks together); the direct terms are added where their path
  // is chosen: factorised legs, or the skinny product with the resident block (plus the build, once)
  ctx->bse_algo_flops += (double)k * ((cx != 0 ? 4.0 * B * naux : 0.0) + (cqp != 0 ? 2.0 * B * (vt + ct) : 0.0));
  ctx->bse_columns += k;
  ctx->bse_products++;
  ensure_gathered(ctx);
  // occupied / virtual slices owned by this rank
  const int nvloc = ctx->owned_count(voff, vt, ctx->rank);
  const int v_rel0 = ctx->first_owned(voff, ctx->rank) - voff;
  const int lvfirst = nvloc ? ctx->local_index(voff + v_rel0) : 0;
  const int ncloc = ctx->owned_count(coff, ct, ctx->rank);
  const int c_rel0 = ctx->first_owned(coff, ctx->rank) - coff;
  const int lcfirst = ncloc ? ctx->local_index(coff + c_rel0) : 0;
  GW_CUDA(cudaMemset2DAsync(Y, sizeof(double) * ldy, 0, sizeof(double) * B, k, ctx->stream));

  if (cqp != 0 && ctx->rank == 0) {
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
    vc_project(ctx, k, Xin, ldin, W);
    vc_expand(ctx, (double)cx, k, W, Y, ldy);
  }

  const bool direct_dense = (cd != 0 || cd2 != 0) && dense_ready(ctx, cd != 0 ? 0 : 1, k);
  if (direct_dense) {
    ctx->bse_algo_flops += 2.0 * (double)B * B * k;
    dense_apply(ctx, cd != 0 ? 0 : 1, cd != 0 ? -(double)cd : -(double)cd2, k, Xin, ldin, Y, ldy);
  } else if (cd != 0 || cd2 != 0) {
    ctx->bse_algo_flops += (double)k * factorised_flops_per_column(ctx, cd != 0 ? 0 : 1);
    const int vtp = round_up(vt, 2);
    const long long ldU = (long long)vtp * naux;
    const int nout = cd != 0 ? ncloc : nvloc;  // chunked index: local c1 for Hd, local v1 for Hd2
    if (nout > 0) {
      // chunk budget: the configured size, but never more than the buffer already held plus half the free memory
      size_t budget = ctx->bse_chunk_bytes;
      {
        size_t free_b = 0, total_b = 0;
        GW_CUDA(cudaMemGetInfo(&free_b, &total_b));
        const size_t held = ctx->bufs["bse_U"].cap * sizeof(double);
        budget = std::min(budget, std::max(held, held + free_b / 2));
      }
      int nc = (int)std::max<long long>(1, (long long)(budget / sizeof(double)) / (ldU * k));
      nc = std::min(nc, nout);
      while ((long long)naux * nc >= (1LL << 31) || (long long)nc * k >= (1LL << 31)) nc = std::max(1, nc / 2);
      if (nc < nout) {
        // keep the N extent (nc * k) of the second GEMM a multiple of the 64-wide tile
        int g = 64, b = k;
        while (b) {
          const int t = g % b;
          g = b;
          b = t;
        }
        const int step = 64 / g;
        if (nc > step) nc -= nc % step;
      }
      double* U = ctx->buf("bse_U", (size_t)ldU * nc * k);
      const BlockView vv = vv_view(ctx), cv = cv_view(ctx);
      // trial vectors with an even row stride: X[(v2,c2),kv] -> Xp[(kv,v2)][c2], ld = ctp, so that the operand
      // of the first GEMM is 16-byte aligned row by row (ct is odd for most molecules)
      const int ctp = round_up(ct, 2);
      const double* Xa = Xin;
      long long xa_ri = ct, xa_ro = ldin;
      if (ct != ctp) {
        double* Xp = ctx->buf("bse_Xp", (size_t)ctp * vt * k);
        if (ldin == B) {
          GW_CUDA(cudaMemcpy2DAsync(Xp, sizeof(double) * ctp, Xin, sizeof(double) * ct, sizeof(double) * ct,
                                    (size_t)vt * k, cudaMemcpyDeviceToDevice, ctx->stream));
        } else {
          for (int j = 0; j < k; ++j)
            GW_CUDA(cudaMemcpy2DAsync(Xp + (size_t)j * vt * ctp, sizeof(double) * ctp, Xin + (size_t)j * ldin,
                                      sizeof(double) * ct, sizeof(double) * ct, vt, cudaMemcpyDeviceToDevice,
                                      ctx->stream));
        }
        Xa = Xp;
        xa_ri = ctp;
        xa_ro = (long long)vt * ctp;
      }
      for (int a = 0; a < nout; a += nc) {
        const int n1 = std::min(nc, nout - a);
        // U[(v2, chi), (l, kv)] = eps_inv[chi] sum_c2 X[c2,(v2,kv)] Mblk[l][c2, chi]   (l: local slices of the chunk)
        GemmParams p;
        p.M = vt * k;
        p.N = naux * n1;
        p.Ki = ct;
        p.A.ptr = Xa;
        p.A.Lr = vt;
        p.A.s_ri = xa_ri;
        p.A.s_ro = xa_ro;
        p.A.s_ki = 1;
        p.B.ptr = X + (long long)((cd != 0 ? lcfirst : lvfirst) + a) * npad + coff;
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
        q.A.s_ki = 1;
        q.B.ptr = U;
        q.B.s_ri = ldU;
        q.B.s_ki = 1;
        q.B.s_ko = vtp;
        q.beta = 1.0;
        q.Ln = n1;
        q.sC_no = ldy;
        if (cd != 0) {
          // Y[(v1, c1), kv] -= cd sum_{chi,v2} M[v1][v2,chi] U[(v2,chi),(l,kv)],  c1 = c_rel0 + (a+l) world
          q.M = vt;
          q.A.ptr = vv.ptr;
          q.A.s_ri = vv.row;
          q.A.s_ko = vv.pole;
          q.C = Y + c_rel0 + (long long)a * world;
          q.sC_mi = ct;
          q.sC_ni = world;
          q.alpha = -cd;
        } else {
          // Y[(v1, c1), kv] -= cd2 sum_{chi,v2} M[c1][v2,chi] U[(v2,chi),(l,kv)],  v1 = v_rel0 + (a+l) world
          q.M = ct;
          q.A.ptr = cv.ptr;
          q.A.s_ri = cv.row;
          q.A.s_ko = cv.pole;
          q.C = Y + ((long long)v_rel0 + (long long)a * world) * ct;
          q.sC_mi = 1;
          q.sC_ni = (long long)world * ct;
          q.alpha = -cd2;
        }
        ctx->gemm(q);
      }
    }
  }
  if (world > 1) {
    if (ldy == B) {
      allreduce_dev(ctx, Y, (size_t)B * k);
    } else {
      for (int j = 0; j < k; ++j) allreduce_dev(ctx, Y + (size_t)j * ldy, (size_t)B);
    }
  }
}

// Cross-spin Hd2 block of the unrestricted full-BSE B operator (BSE_OPERATOR_UKS::add_direct2_block with out != in,
// bse_operator_uks.cc:136-172, 252-255): the output channel's tensor (this context) supplies M_out[c1][v2, chi], the
// input channel's tensor (other_X, same layout, the other context's Mmn) supplies M_in[v1][c2, chi]:
//   Y[(v1, c1), kv] += alpha sum_{v2, c2, chi} M_out[c1][v2, chi] eps_inv[chi] M_in[v1][c2, chi] X[(v2, c2), kv]
// v1, c1 run over the output channel's ranges, v2, c2 over the input channel's.  Same two-leg factorisation as Hd2.
void hd2_cross(gwbse_ctx* ctx, const double* other_X, int homo_in, double alpha, int k, const double* Xin, int ldin,
               double* Y, int ldy) {
  auto& st = ctx->bse;
  GW_REQUIRE(st.ready, "BSE operator not configured (gwbse_bse_configure)");
  GW_REQUIRE(ctx->world == 1, "the unrestricted operator is single-GPU");
  mmn_complete_rotation(ctx);
  const int naux = ctx->naux, npad = ctx->npad;
  const long long ldx = ctx->ldx;
  const int vt_o = st.vt, ct_o = st.ct, voff = st.voff, coff_o = st.coff;
  const int vt_i = homo_in - st.vmin + 1, ct_i = st.cmax - homo_in, coff_i = homo_in + 1 - st.rpamin;
  GW_REQUIRE(vt_i > 0 && ct_i > 0 && coff_i + ct_i <= ctx->ntotal, "invalid input-channel ranges");
  GW_REQUIRE(ldin >= vt_i * ct_i && ldy >= vt_o * ct_o, "Shape mismatch in the cross-spin BSE block");
  if (k <= 0) return;
  const int vtp = round_up(vt_i, 2);
  const long long ldU = (long long)vtp * naux;
  size_t budget = ctx->bse_chunk_bytes;
  {
    size_t free_b = 0, total_b = 0;
    GW_CUDA(cudaMemGetInfo(&free_b, &total_b));
    const size_t held = ctx->bufs["bse_U"].cap * sizeof(double);
    budget = std::min(budget, std::max(held, held + free_b / 2));
  }
  int nc = (int)std::max<long long>(1, (long long)(budget / sizeof(double)) / (ldU * k));
  nc = std::min(nc, vt_o);
  while ((long long)naux * nc >= (1LL << 31) || (long long)nc * k >= (1LL << 31)) nc = std::max(1, nc / 2);
  double* U = ctx->buf("bse_U", (size_t)ldU * nc * k);
  // trial vectors at an even row pitch (16-byte copies), as in the same-spin term
  const int ctp = round_up(ct_i, 2);
  const double* Xa = Xin;
  long long xa_ri = ct_i, xa_ro = ldin;
  if (ct_i != ctp) {
    double* Xp = ctx->buf("bse_Xp", (size_t)ctp * vt_i * k);
    for (int j = 0; j < k; ++j)
      GW_CUDA(cudaMemcpy2DAsync(Xp + (size_t)j * vt_i * ctp, sizeof(double) * ctp, Xin + (size_t)j * ldin,
                                sizeof(double) * ct_i, sizeof(double) * ct_i, vt_i, cudaMemcpyDeviceToDevice,
                                ctx->stream));
    Xa = Xp;
    xa_ri = ctp;
    xa_ro = (long long)vt_i * ctp;
  }
  for (int a = 0; a < vt_o; a += nc) {
    const int n1 = std::min(nc, vt_o - a);
    // U[(v2, chi), (l, kv)] = eps_inv[chi] sum_c2 X[c2, (v2, kv)] M_in[voff + a + l][coff_i + c2, chi]
    GemmParams p;
    p.M = vt_i * k;
    p.N = naux * n1;
    p.Ki = ct_i;
    p.A.ptr = Xa;
    p.A.Lr = vt_i;
    p.A.s_ri = xa_ri;
    p.A.s_ro = xa_ro;
    p.A.s_ki = 1;
    p.B.ptr = other_X + (long long)(voff + a) * npad + coff_i;
    p.B.Lr = naux;
    p.B.s_ri = ldx;
    p.B.s_ro = npad;
    p.B.s_ki = 1;
    p.C = U;
    p.Lm = vt_i;
    p.sC_mi = 1;
    p.sC_mo = ldU * n1;
    p.Ln = naux;
    p.sC_ni = vtp;
    p.sC_no = ldU;
    p.nscale = st.eps_inv;
    p.nscale_mod = naux;
    ctx->gemm(p);
    // Y[(a + l, c1), kv] += alpha sum_{chi, v2} M_out[coff_o + c1][voff + v2, chi] U[(v2, chi), (l, kv)]
    GemmParams q;
    q.Ko = naux;
    q.Ki = vt_i;
    q.M = ct_o;
    q.N = n1 * k;
    q.A.ptr = ctx->X + (long long)coff_o * npad + voff;
    q.A.s_ri = npad;
    q.A.s_ki = 1;
    q.A.s_ko = ldx;
    q.B.ptr = U;
    q.B.s_ri = ldU;
    q.B.s_ki = 1;
    q.B.s_ko = vtp;
    q.C = Y + (long long)a * ct_o;
    q.sC_mi = 1;
    q.Ln = n1;
    q.sC_ni = ct_o;
    q.sC_no = ldy;
    q.alpha = alpha;
    q.beta = 1.0;
    ctx->gemm(q);
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
  // a changed screening invalidates the materialised blocks (dense_ready compares eps_version)
  if (st.eps_inv_host.size() != (size_t)ctx->naux ||
      !std::equal(st.eps_inv_host.begin(), st.eps_inv_host.end(), eps_inv)) {
    st.eps_inv_host.assign(eps_inv, eps_inv + ctx->naux);
    st.eps_version++;
  }
  st.eps_inv = ctx->buf("bse_eps_inv", ctx->naux);
  st.hqp = ctx->buf("bse_hqp", (size_t)hs * hs);
  GW_CUDA(cudaMemcpyAsync(st.eps_inv, eps_inv, sizeof(double) * ctx->naux, cudaMemcpyHostToDevice, ctx->stream));
  GW_CUDA(copy2d_async(st.hqp, sizeof(double) * hs, Hqp, sizeof(double) * ldh, sizeof(double) * hs, hs,
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
    GW_CUDA(copy2d_async(Xd, sizeof(double) * B, X, sizeof(double) * ldx, sizeof(double) * B, k,
                              cudaMemcpyHostToDevice, ctx->stream));
    bse_matmul_dev(ctx, cqp, cx, cd, cd2, k, Xd, B, Yd, B);
    GW_CUDA(copy2d_async(Y, sizeof(double) * ldy, Yd, sizeof(double) * B, sizeof(double) * B, k,
                              cudaMemcpyDeviceToHost, ctx->stream));
    GW_CUDA(cudaStreamSynchronize(ctx->stream));
  }
  GW_API_END(ctx)
}

int gwbse_bse_vc_project_dev(gwbse_ctx* ctx, int k, const double* X_dev, int ldx, double* W_dev) {
  GW_API_BEGIN(ctx)
  GW_PROF(ctx, "bse_vc_project");
  GW_REQUIRE(ctx->bse.ready, "BSE operator not configured (gwbse_bse_configure)");
  GW_REQUIRE(ldx >= ctx->bse.size, "Shape mismatch in BSE projection");
  bse_rows_ready(ctx);
  if (k > 0) vc_project(ctx, k, X_dev, ldx, W_dev);
  GW_API_END(ctx)
}

int gwbse_bse_vc_expand_dev(gwbse_ctx* ctx, double alpha, int screened, int k, const double* W_dev, double* Y_dev,
                            int ldy) {
  GW_API_BEGIN(ctx)
  GW_PROF(ctx, "bse_vc_expand");
  GW_REQUIRE(ctx->bse.ready, "BSE operator not configured (gwbse_bse_configure)");
  GW_REQUIRE(ldy >= ctx->bse.size, "Shape mismatch in BSE expansion");
  GW_REQUIRE(ctx->world == 1, "gwbse_bse_vc_expand_dev accumulates into Y in place: single-GPU (the UKS operator)");
  bse_rows_ready(ctx);
  if (k > 0) {
    const double* W = W_dev;
    if (screened) {  // W <- diag(eps^-1) W
      double* Ws = ctx->buf("bse_Ws", (size_t)ctx->naux * k);
      launch_diag_scale('L', ctx->naux, k, W_dev, ctx->naux, ctx->bse.eps_inv, Ws, ctx->naux, ctx->stream);
      ctx->launches++;
      W = Ws;
    }
    vc_expand(ctx, alpha, k, W, Y_dev, ldy);
  }
  GW_API_END(ctx)
}

int gwbse_bse_hd2_cross_dev(gwbse_ctx* ctx, gwbse_ctx* other, int homo_other, double alpha, int k,
                            const double* X_dev, int ldx, double* Y_dev, int ldy) {
  GW_API_BEGIN(ctx)
  GW_PROF(ctx, "bse_hd2_cross");
  GW_REQUIRE(other && other->X && ctx->X, "both contexts need a filled Mmn");
  GW_REQUIRE(other->naux == ctx->naux && other->ldx == ctx->ldx && other->npad == ctx->npad &&
                 other->mmin == ctx->mmin && other->nmin == ctx->nmin && other->device == ctx->device,
             "the two Mmn tensors must have the same shape and live on the same GPU");
  mmn_complete_rotation(other);
  GW_CUDA(cudaStreamSynchronize(other->stream));
  hd2_cross(ctx, other->X, homo_other, alpha, k, X_dev, ldx, Y_dev, ldy);
  GW_API_END(ctx)
}

int gwbse_bse_stats(gwbse_ctx* ctx, double* algo_flops, long long* products, long long* columns, int reset) {
  GW_API_BEGIN(ctx)
  if (algo_flops) *algo_flops = ctx->bse_algo_flops;
  if (products) *products = ctx->bse_products;
  if (columns) *columns = ctx->bse_columns;
  if (reset) {
    ctx->bse_algo_flops = 0.0;
    ctx->bse_products = ctx->bse_columns = 0;
  }
  GW_API_END(ctx)
}

int gwbse_bse_dense_stats(gwbse_ctx* ctx, long long* builds, long long* columns, double* resident_bytes) {
  GW_API_BEGIN(ctx)
  if (builds) *builds = ctx->bse_dense_builds;
  if (columns) *columns = ctx->bse_dense_columns;
  if (resident_bytes) {
    double b = 0.0;
    for (const auto& blk : ctx->bse.dense)
      if (blk.valid && blk.H && (!blk.in_x2 || blk.x2_epoch == ctx->x2_epoch)) b += 8.0 * (double)blk.ld * (double)blk.ncols;
    *resident_bytes = b;
  }
  GW_API_END(ctx)
}

int gwbse_bse_diagonal(gwbse_ctx* ctx, int cqp, int cx, int cd, int cd2, double* diag) {
  GW_API_BEGIN(ctx)
  GW_PROF(ctx, "bse_diagonal");
  auto& st = ctx->bse;
  GW_REQUIRE(st.ready, "BSE operator not configured (gwbse_bse_configure)");
  GW_REQUIRE(!(cd != 0 && cd2 != 0), "Hamiltonian cannot contain Hd and Hd2 at the same time");
  bse_rows_ready(ctx);
  ensure_gathered(ctx);
  const int vt = st.vt, ct = st.ct, naux = ctx->naux;
  double* d = ctx->buf("bse_diag", st.size);
  double* Dcc = ctx->buf("bse_dcc", (size_t)ct * naux);
  double* Dvv = ctx->buf("bse_dvv", (size_t)vt * naux);
  GW_CUDA(cudaMemsetAsync(d, 0, sizeof(double) * st.size, ctx->stream));
  GW_CUDA(cudaMemsetAsync(Dcc, 0, sizeof(double) * (size_t)ct * naux, ctx->stream));
  GW_CUDA(cudaMemsetAsync(Dvv, 0, sizeof(double) * (size_t)vt * naux, ctx->stream));
  launch_slice_diag(ctx->X, ctx->ldx, ctx->npad, naux, st.coff, ct, ctx->rank, ctx->world, Dcc, ctx->stream);
  launch_slice_diag(ctx->X, ctx->ldx, ctx->npad, naux, st.voff, vt, ctx->rank, ctx->world, Dvv, ctx->stream);
  if (ctx->world > 1) {
    allreduce_dev(ctx, Dcc, (size_t)ct * naux);
    allreduce_dev(ctx, Dvv, (size_t)vt * naux);
  }
  const int nvloc = ctx->owned_count(st.voff, vt, ctx->rank);
  const int v_rel0 = ctx->first_owned(st.voff, ctx->rank) - st.voff;
  const int lvfirst = nvloc ? ctx->local_index(st.voff + v_rel0) : 0;
  const BlockView cv = cv_view(ctx);
  launch_bse_diag(ctx->X, ctx->ldx, ctx->npad, naux, vt, ct, st.voff, st.coff, v_rel0, ctx->world, lvfirst, nvloc,
                  cv.ptr, cv.row, cv.pole, Dcc, Dvv, st.eps_inv, st.hqp, vt + ct, cqp, cx, cd, cd2, d, ctx->stream);
  ctx->launches += 3;
  if (ctx->world > 1) allreduce_dev(ctx, d, (size_t)st.size);
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
