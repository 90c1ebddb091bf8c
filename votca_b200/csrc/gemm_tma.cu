// Host side of the TMA-staged DMMA GEMM (gemm_tma.cuh): tensor-map construction, tile / split-K plan, launch.
#include "gemm_tma.cuh"

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>

namespace gwbse {

namespace {

struct TCfg {
  int BM, BN, occ;
  double eff;  // relative throughput of the tile shape (measured on B200, scratch/gemm_tma_sweep.py)
};
// 0: 128x128 (8 consumer warps of 64x32, one CTA per SM)   1: 128x64 (4 warps of 64x32, two CTAs per SM)
// 2: 128x32  (4 warps of 32x32, two CTAs per SM)           3: 64x64  (4 warps of 32x32, three CTAs per SM)
// 4: 128x48  (4 warps of 32x48, two CTAs per SM): the 144- and 287-wide dimensions of the BSE legs
// eff: 4096^3 NN / TN and the long-K shapes on B200 (profiles/r02_gemm_tile_sweep.txt): every 128-row tile runs
// within 2 % of 33 TFLOP/s, the 64 x 64 tile at 29
constexpr TCfg kT[5] = {{128, 128, 1, 1.00}, {128, 64, 2, 0.985}, {128, 32, 2, 0.99}, {64, 64, 3, 0.88}, {128, 48, 2, 0.975}};
constexpr int kNT = 5;

using EncodeFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                              const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                              CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeFn encode_fn() {
  static EncodeFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      p = nullptr;
    return reinterpret_cast<EncodeFn>(p);
  }();
  return fn;
}

bool g_enabled = [] {
  const char* e = std::getenv("GWBSE_NO_TMA");
  return !(e && e[0] == '1');
}();

bool even(long long v) { return (v & 1LL) == 0; }

// Row structure of an operand as the TMA kernel sees it: flat (one row index) or compound ro * Lr + ri with tiles
// that stay inside one ro.
struct RowShape {
  bool kmajor = true, compound = false;
  long long Lr = 1LL << 30, n_ro = 1;
};

bool operand_shape(const GemmOperand& op, int rows, RowShape* rs) {
  if (op.s_ki == 1)
    rs->kmajor = true;
  else if (op.s_ri == 1)
    rs->kmajor = false;
  else
    return false;
  rs->compound = op.Lr < rows && op.s_ro != (long long)op.Lr * op.s_ri;
  if (rs->compound) {
    if (rows % op.Lr != 0) return false;
    rs->Lr = op.Lr;
    rs->n_ro = rows / op.Lr;
  }
  return true;
}

// Tensor map + coordinate recipe of one operand; false if it cannot be expressed.
bool make_operand_map(const GemmOperand& op, int rows, int Ko, int Ki, int Z1, int tile_rows, CUtensorMap* map,
                      TmaOperand* t) {
  RowShape rs;
  if (!operand_shape(op, rows, &rs)) return false;
  const bool kmajor = rs.kmajor;
  // TMA wants the start of every box 16-byte aligned: an 8-byte offset view cannot be expressed by a shifted
  // coordinate either (the innermost coordinate must itself be a multiple of 16 bytes) - cp.async kernel then
  if (reinterpret_cast<uintptr_t>(op.ptr) % 16 != 0) return false;
  const int shift = 0;
  const double* base = op.ptr;
  const long long rows_in = rs.compound ? rs.Lr : rows;  // extent of the inner row index
  const long long s_row = kmajor ? op.s_ri : 1, s_k = kmajor ? 1 : op.s_ki;
  // (ro, ko, z1): broadcast / absent dimensions collapse to extent 1
  long long ext[3] = {rs.compound ? rs.n_ro : 1, Ko, Z1}, str[3] = {op.s_ro, op.s_ko, op.s_z1};
  int use[3];
  for (int d = 0; d < 3; ++d) {
    use[d] = (ext[d] > 1 && str[d] != 0) ? 1 : 0;
    if (!use[d]) ext[d] = 1;
    if (use[d] && (!even(str[d]) || str[d] < 0)) return false;
  }
  if (rs.compound && !use[0]) return false;
  const long long s1 = kmajor ? s_row : s_k;  // stride of tensor dimension 1
  const long long e0 = (kmajor ? Ki : rows_in) + shift, e1 = kmajor ? rows_in : Ki;
  if (e1 > 1 && (!even(s1) || s1 <= 0)) return false;
  cuuint64_t dims[5] = {(cuuint64_t)e0, (cuuint64_t)e1, (cuuint64_t)ext[0], (cuuint64_t)ext[1], (cuuint64_t)ext[2]};
  // strides of dimensions 1..4 in bytes; unused dimensions get a valid dummy (multiple of 16)
  cuuint64_t strides[4];
  strides[0] = (e1 > 1) ? (cuuint64_t)s1 * 8 : 16;
  for (int d = 0; d < 3; ++d) strides[d + 1] = use[d] ? (cuuint64_t)str[d] * 8 : 16;
  for (int d = 0; d < 4; ++d)
    if (strides[d] % 16 != 0 || strides[d] >= (1ULL << 40)) return false;
  for (int d = 0; d < 5; ++d)
    if (dims[d] == 0 || dims[d] > (1ULL << 32)) return false;
  cuuint32_t box[5] = {16, (cuuint32_t)(kmajor ? tile_rows : 16), 1, 1, 1};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  EncodeFn enc = encode_fn();
  if (!enc) return false;
  const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 5, const_cast<double*>(base), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return false;
  t->row0 = kmajor ? 0 : shift;
  t->k0 = kmajor ? shift : 0;
  // a broadcast ro is not supported (compound rows always address memory), ko / z1 may be broadcast
  t->use_ko = use[1];
  t->use_z1 = use[2];
  t->Lr = rs.compound ? (int)rs.Lr : (1 << 30);
  t->tiles_per_ro = rs.compound ? (int)ceil_div<long long>(rs.Lr, tile_rows) : (1 << 30);
  return true;
}

bool make_weight_map(const GemmParams& p, CUtensorMap* map, int* use_ko, int* use_z1) {
  if (reinterpret_cast<uintptr_t>(p.w) % 16 != 0) return false;
  long long ext[2] = {p.Ko, p.Z1}, str[2] = {p.sW_ko, p.sW_z1};
  int use[2];
  for (int d = 0; d < 2; ++d) {
    use[d] = (ext[d] > 1 && str[d] != 0) ? 1 : 0;
    if (!use[d]) ext[d] = 1;
    if (use[d] && (!even(str[d]) || str[d] < 0)) return false;
  }
  cuuint64_t dims[5] = {(cuuint64_t)p.Ki, (cuuint64_t)ext[0], (cuuint64_t)ext[1], 1, 1};
  cuuint64_t strides[4] = {use[0] ? (cuuint64_t)str[0] * 8 : 16, use[1] ? (cuuint64_t)str[1] * 8 : 16, 16, 16};
  cuuint32_t box[5] = {16, 1, 1, 1, 1}, estr[5] = {1, 1, 1, 1, 1};
  EncodeFn enc = encode_fn();
  if (!enc) return false;
  const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 5, const_cast<double*>(p.w), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  *use_ko = use[0];
  *use_z1 = use[1];
  return r == CUDA_SUCCESS;
}

// cheap host-side test (no encoding) of what make_operand_map will accept
bool operand_ok(const GemmOperand& op, int rows, int Ko, int Z1, RowShape* rs) {
  if (!operand_shape(op, rows, rs)) return false;
  if (reinterpret_cast<uintptr_t>(op.ptr) % 16 != 0) return false;
  const long long rows_in = rs->compound ? rs->Lr : rows;
  if (rs->kmajor) {
    if (rows_in > 1 && (!even(op.s_ri) || op.s_ri <= 0)) return false;
  } else {
    if (!even(op.s_ki) || op.s_ki <= 0) return false;
  }
  const long long ext[3] = {rs->compound ? rs->n_ro : 1, Ko, Z1}, str[3] = {op.s_ro, op.s_ko, op.s_z1};
  for (int d = 0; d < 3; ++d)
    if (ext[d] > 1 && str[d] != 0 && (!even(str[d]) || str[d] < 0)) return false;
  if (rs->compound && (op.s_ro == 0 || rs->n_ro < 1)) return false;
  return true;
}

bool eligible(const GemmParams& p, RowShape* ra = nullptr, RowShape* rb = nullptr) {
  if (!g_enabled || !encode_fn()) return false;
  if (p.M <= 0 || p.N <= 0 || p.Ki <= 0 || p.Ko < 1 || p.Z2 != 1) return false;
  RowShape a, b;
  if (!operand_ok(p.A, p.M, p.Ko, p.Z1, &a) || !operand_ok(p.B, p.N, p.Ko, p.Z1, &b)) return false;
  if ((a.compound || b.compound) && p.lower_only) return false;
  if (p.w) {
    if (!a.kmajor) return false;  // weights ride on a K-major A (as in the cp.async kernel)
    if (reinterpret_cast<uintptr_t>(p.w) % 16 != 0) return false;
    const long long ext[2] = {p.Ko, p.Z1}, str[2] = {p.sW_ko, p.sW_z1};
    for (int d = 0; d < 2; ++d)
      if (ext[d] > 1 && str[d] != 0 && (!even(str[d]) || str[d] < 0)) return false;
  }
  if (ra) *ra = a;
  if (rb) *rb = b;
  return true;
}

struct TPlan {
  int cfg = 0, splitk = 1, tiles_m = 0, tiles_n = 0;
  bool swap = false;
  long long ntiles = 0;
};

long long tiles_along(int rows, const RowShape& rs, int tile) {
  return rs.compound ? rs.n_ro * ceil_div<long long>(rs.Lr, tile) : ceil_div(rows, tile);
}

long long count_tiles(int M, int N, const RowShape& ra, const RowShape& rb, int cfg, bool lower) {
  const long long tm = tiles_along(M, ra, kT[cfg].BM), tn = tiles_along(N, rb, kT[cfg].BN);
  if (!lower) return tm * tn;
  long long c = 0;
  for (long long i = 0; i < tm; ++i) c += std::min<long long>(tn, (i * kT[cfg].BM + kT[cfg].BM - 1) / kT[cfg].BN + 1);
  return c;
}

// Run-time estimate of one choice: the persistent grid has num_sms * occ CTAs, items are dealt round robin,
// an item costs its tile area times (k-steps + fixed prologue/epilogue equivalent).
double tplan_cost(const GemmParams& p, const RowShape& ra_in, const RowShape& rb_in, int num_sms, int cfg, bool swap,
                  int splitk, long long T_total) {
  const int M = swap ? p.N : p.M, N = swap ? p.M : p.N;
  const RowShape& ra = swap ? rb_in : ra_in;
  const RowShape& rb = swap ? ra_in : rb_in;
  const long long tiles = count_tiles(M, N, ra, rb, cfg, p.lower_only != 0) * p.Z1 * p.Z2;
  const long long items = tiles * splitk;
  const long long slots = (long long)num_sms * kT[cfg].occ;
  const long long rounds = ceil_div(items, slots);
  // epilogue / pipeline refill, in k-step equivalents: the consumers of a one-CTA-per-SM shape stall the SM while
  // they store, co-resident CTAs overlap it
  const double fixed = kT[cfg].occ == 1 ? 6.0 : 3.0;
  const double ksteps = (double)ceil_div<long long>(T_total, splitk) + fixed;
  const double flops_item = 2.0 * kT[cfg].BM * kT[cfg].BN * ksteps * TMA_BK;
  const double rate_sm = 33.5e12 / 148.0 * kT[cfg].eff;
  double t = (double)rounds * kT[cfg].occ * flops_item / rate_sm;
  if (splitk > 1) {
    const long long tm = tiles_along(M, ra, kT[cfg].BM), tn = tiles_along(N, rb, kT[cfg].BN);
    t += 3.0e-6 + 16.0 * (double)tm * kT[cfg].BM * (double)tn * kT[cfg].BN * p.Z1 * p.Z2 * splitk / 5.0e12;
  }
  if (swap) t *= 1.01;
  return t;
}

TPlan make_tplan(const GemmParams& p, const RowShape& ra, const RowShape& rb, int num_sms, int force_cfg,
                 int force_splitk) {
  TPlan pl;
  if (ra.compound || rb.compound) force_splitk = 1;  // the split-K workspace is indexed by flat rows
  const bool can_swap = !p.w && !p.nscale && !p.lower_only;
  const long long T_total = (long long)p.Ko * ceil_div(p.Ki, TMA_BK);
  static const int kSplits[] = {1, 2, 3, 4, 5, 6, 7, 8, 10, 12, 14, 16, 20, 24, 28, 32, 40, 48, 56, 64};
  double best = 1e300;
  for (int sw = 0; sw <= (can_swap ? 1 : 0); ++sw)
    for (int c = 0; c < kNT; ++c) {
      if (force_cfg >= 0 && c != force_cfg) continue;
      for (int sk : kSplits) {
        if (force_splitk > 0) sk = force_splitk;
        if (sk > 1 && (long long)sk * 4 > T_total && force_splitk <= 0) break;
        const double t = tplan_cost(p, ra, rb, num_sms, c, sw != 0, sk, T_total);
        if (t < best) {
          best = t;
          pl.cfg = c;
          pl.swap = sw != 0;
          pl.splitk = sk;
        }
        if (force_splitk > 0) break;
      }
    }
  const int M = pl.swap ? p.N : p.M, N = pl.swap ? p.M : p.N;
  const RowShape& sa = pl.swap ? rb : ra;
  const RowShape& sb = pl.swap ? ra : rb;
  pl.tiles_m = (int)tiles_along(M, sa, kT[pl.cfg].BM);
  pl.tiles_n = (int)tiles_along(N, sb, kT[pl.cfg].BN);
  pl.splitk = (int)std::min<long long>(pl.splitk, std::max<long long>(1, T_total));
  pl.ntiles = count_tiles(M, N, sa, sb, pl.cfg, p.lower_only != 0);
  return pl;
}

GemmParams swap_ops(GemmParams p) {
  std::swap(p.A, p.B);
  std::swap(p.M, p.N);
  std::swap(p.sC_mi, p.sC_ni);
  std::swap(p.sC_mo, p.sC_no);
  std::swap(p.Lm, p.Ln);
  return p;
}

template <int BM, int BN, int WGM, int WGN, int STAGES, int MINB, bool AK, bool BKM, bool HASW>
bool launch_tma_one(const GemmTmaParams& P, int grid, cudaStream_t stream) {
  using SM = TmaSmem<BM, BN, STAGES, HASW>;
  auto kern = gemm_tma_kernel<BM, BN, WGM, WGN, STAGES, MINB, AK, BKM, HASW>;
  static int state = 0;  // 0 unknown, 1 usable, -1 not usable
  if (state == 0) {
    // setmaxnreg.inc waits for registers the producer warpgroup released: the launch-time allocation must be the
    // one the kernel's constants assume, or the consumers would wait forever
    constexpr int NTHR = (WGM * WGN + 4) * 32;
    constexpr int REG_ALL = (65536 / (NTHR * MINB)) / 8 * 8 > 255 ? 248 : (65536 / (NTHR * MINB)) / 8 * 8;
    cudaFuncAttributes at;
    GW_CUDA(cudaFuncGetAttributes(&at, kern));
    if (at.numRegs != REG_ALL) {
      state = -1;
    } else {
      GW_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SM::bytes));
      state = 1;
    }
  }
  if (state < 0) return false;
  kern<<<grid, (WGM * WGN + 4) * 32, SM::bytes, stream>>>(P);
  GW_CUDA(cudaGetLastError());
  return true;
}

template <int BM, int BN, int WGM, int WGN, int STAGES, int MINB>
bool launch_tma_cfg(const GemmTmaParams& P, int grid, bool ak, bool bk, bool hasw, cudaStream_t stream) {
  if (hasw) {
    if (bk) return launch_tma_one<BM, BN, WGM, WGN, STAGES, MINB, true, true, true>(P, grid, stream);
    return launch_tma_one<BM, BN, WGM, WGN, STAGES, MINB, true, false, true>(P, grid, stream);
  }
  if (ak && bk) return launch_tma_one<BM, BN, WGM, WGN, STAGES, MINB, true, true, false>(P, grid, stream);
  if (ak) return launch_tma_one<BM, BN, WGM, WGN, STAGES, MINB, true, false, false>(P, grid, stream);
  if (bk) return launch_tma_one<BM, BN, WGM, WGN, STAGES, MINB, false, true, false>(P, grid, stream);
  return launch_tma_one<BM, BN, WGM, WGN, STAGES, MINB, false, false, false>(P, grid, stream);
}

}  // namespace

void gemm_tma_set_enabled(bool on) { g_enabled = on; }
bool gemm_tma_enabled() { return g_enabled && encode_fn() != nullptr; }

size_t gemm_tma_ws_bytes(const GemmParams& p, int num_sms, int force_cfg, int force_splitk) {
  RowShape ra, rb;
  if (!eligible(p, &ra, &rb)) return 0;
  const TPlan pl = make_tplan(p, ra, rb, num_sms, force_cfg, force_splitk);
  if (pl.splitk <= 1) return 0;
  return sizeof(double) * (size_t)pl.tiles_m * kT[pl.cfg].BM * pl.tiles_n * kT[pl.cfg].BN * pl.splitk * p.Z1 * p.Z2;
}

// Short K (the K = ct leg of the BSE intermediate, 18 k-steps): the tile's epilogue is a fifth of its run time and
// the persistent kernel has nothing to overlap it with (4 consumer warps per co-resident CTA cannot keep the DMMA
// pipe busy alone) - measured 26.4 against 29.9 TFLOP/s of the cp.async kernel with its 2 x 8 warps per SM
// (scratch/tma_probe.cu, 4320 x 50816 x 288).  Unless a TMA tile shape is forced, those go to the cp.async kernel.
constexpr long long kMinKStepsForTma = 48;

bool gemm_tma_describe(const GemmParams& p, int num_sms, int force_cfg, int force_splitk, int* cfg, int* swap,
                       int* splitk) {
  RowShape ra, rb;
  if (!eligible(p, &ra, &rb)) return false;
  if (force_cfg < 0 && (long long)p.Ko * ceil_div(p.Ki, TMA_BK) < kMinKStepsForTma) return false;
  const TPlan pl = make_tplan(p, ra, rb, num_sms, force_cfg, force_splitk);
  *cfg = pl.cfg;
  *swap = pl.swap ? 1 : 0;
  *splitk = pl.splitk;
  return true;
}

bool gemm_tma_try_launch(const GemmParams& p_in, cudaStream_t stream, double* ws, size_t ws_bytes, int num_sms,
                         int force_cfg, int force_splitk) {
  RowShape ra, rb;
  if (!eligible(p_in, &ra, &rb)) return false;
  TPlan pl = make_tplan(p_in, ra, rb, num_sms, force_cfg, force_splitk);
  GemmParams p = pl.swap ? swap_ops(p_in) : p_in;
  const TCfg& c = kT[pl.cfg];
  if (pl.splitk > 1) {
    const size_t need = sizeof(double) * (size_t)pl.tiles_m * c.BM * pl.tiles_n * c.BN * pl.splitk * p.Z1 * p.Z2;
    if (ws == nullptr || need > ws_bytes) {
      const size_t per = need / pl.splitk;
      const int fit = (ws && per) ? (int)(ws_bytes / per) : 1;
      pl.splitk = std::max(1, std::min(pl.splitk, fit));
    }
  }
  GemmTmaParams P;
  std::memset(&P.mapA, 0, sizeof(CUtensorMap) * 3);
  bool ak = p.A.s_ki == 1, bk = p.B.s_ki == 1;
  if (!make_operand_map(p.A, p.M, p.Ko, p.Ki, p.Z1, c.BM, &P.mapA, &P.A)) return false;
  if (!make_operand_map(p.B, p.N, p.Ko, p.Ki, p.Z1, c.BN, &P.mapB, &P.B)) return false;
  if (p.w && !make_weight_map(p, &P.mapW, &P.w_use_ko, &P.w_use_z1)) return false;
  p.tiles_m = pl.tiles_m;
  p.tiles_n = pl.tiles_n;
  p.splitk = pl.splitk;
  p.ws = pl.splitk > 1 ? ws : nullptr;
  {
    const double slots = (double)num_sms * c.occ;
    double gm = std::sqrt(slots * c.BN / c.BM);
    gm = std::max(gm, slots / pl.tiles_n);
    p.group_m = (int)std::max(1.0, std::min<double>(pl.tiles_m, std::round(gm)));
  }
  P.g = p;
  P.ntiles = (int)pl.ntiles;
  P.items = pl.ntiles * pl.splitk * p.Z1 * p.Z2;
  GW_REQUIRE(pl.ntiles < (1LL << 31), "GEMM grid too large");
  const int grid = (int)std::min<long long>(P.items, (long long)num_sms * c.occ);
  const bool hasw = p.w != nullptr;
  bool ok;
  switch (pl.cfg) {
    case 0:
      ok = launch_tma_cfg<128, 128, 2, 4, 6, 1>(P, grid, ak, bk, hasw, stream);
      break;
    case 1:
      ok = launch_tma_cfg<128, 64, 2, 2, 4, 2>(P, grid, ak, bk, hasw, stream);
      break;
    case 2:
      ok = launch_tma_cfg<128, 32, 4, 1, 4, 2>(P, grid, ak, bk, hasw, stream);
      break;
    case 4:
      ok = launch_tma_cfg<128, 48, 4, 1, 4, 2>(P, grid, ak, bk, hasw, stream);
      break;
    default:
      ok = launch_tma_cfg<64, 64, 2, 2, 4, 3>(P, grid, ak, bk, hasw, stream);
      break;
  }
  if (!ok) return false;
  if (p.splitk > 1) {
    dim3 rblk(128);
    dim3 rg(ceil_div(p.M, 128), std::min(p.N, 65535), p.Z1 * p.Z2);
    GW_REQUIRE(p.Z1 * p.Z2 <= 65535, "split-K reduce grid too large");
    switch (pl.cfg) {
      case 0:
        gemm_splitk_reduce_kernel<128, 128><<<rg, rblk, 0, stream>>>(p);
        break;
      case 1:
        gemm_splitk_reduce_kernel<128, 64><<<rg, rblk, 0, stream>>>(p);
        break;
      case 2:
        gemm_splitk_reduce_kernel<128, 32><<<rg, rblk, 0, stream>>>(p);
        break;
      case 4:
        gemm_splitk_reduce_kernel<128, 48><<<rg, rblk, 0, stream>>>(p);
        break;
      default:
        gemm_splitk_reduce_kernel<64, 64><<<rg, rblk, 0, stream>>>(p);
        break;
    }
    GW_CUDA(cudaGetLastError());
  }
  return true;
}

}  // namespace gwbse
