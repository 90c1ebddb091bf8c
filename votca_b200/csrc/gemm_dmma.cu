// Launcher + instantiations of the FP64 DMMA GEMM (see gemm_dmma.cuh).
#include <algorithm>
#include <cmath>
#include <cstdint>

#include "gemm_dmma.cuh"
#include "gemm_tma.cuh"

namespace gwbse {

namespace {

struct Cfg {
  int BM, BN, occ;
};
constexpr Cfg kCfg[6] = {{128, 128, 1}, {128, 32, 2}, {64, 64, 2}, {128, 128, 1}, {128, 64, 2}, {128, 48, 2}};

template <int BM, int BN, int WGM, int WGN, int STAGES, int MINB, bool AK, bool BKM, bool HASW>
void launch_one(const GemmParams& p, dim3 grid, cudaStream_t stream) {
  using SM = GemmSmem<BM, BN, STAGES, HASW>;
  auto kern = gemm_dmma_kernel<BM, BN, WGM, WGN, STAGES, MINB, AK, BKM, HASW>;
  static bool attr_set = false;
  if (!attr_set) {
    GW_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SM::bytes));
    attr_set = true;
  }
  kern<<<grid, WGM * WGN * 32, SM::bytes, stream>>>(p);
  GW_CUDA(cudaGetLastError());
}

template <int BM, int BN, int WGM, int WGN, int STAGES, int MINB>
void launch_cfg(const GemmParams& p, dim3 grid, bool ak, bool bk, cudaStream_t stream) {
  const bool hasw = p.w != nullptr;
  if (hasw) {
    GW_REQUIRE(ak, "weighted GEMM needs a K-major A operand");
    if (bk)
      launch_one<BM, BN, WGM, WGN, STAGES, MINB, true, true, true>(p, grid, stream);
    else
      launch_one<BM, BN, WGM, WGN, STAGES, MINB, true, false, true>(p, grid, stream);
    return;
  }
  if (ak && bk)
    launch_one<BM, BN, WGM, WGN, STAGES, MINB, true, true, false>(p, grid, stream);
  else if (ak && !bk)
    launch_one<BM, BN, WGM, WGN, STAGES, MINB, true, false, false>(p, grid, stream);
  else if (!ak && bk)
    launch_one<BM, BN, WGM, WGN, STAGES, MINB, false, true, false>(p, grid, stream);
  else
    launch_one<BM, BN, WGM, WGN, STAGES, MINB, false, false, false>(p, grid, stream);
}

bool is_kmajor(const GemmOperand& op, int Ki) {
  if (op.s_ki == 1) return true;
  if (op.s_ri == 1) return false;
  GW_REQUIRE(Ki == 1, "GEMM operand needs a unit stride along k or along its row index");
  return true;
}

int deduce_vec(const GemmOperand& op, bool kmajor, int rows) {
  auto even = [](long long v) { return (v & 1LL) == 0; };
  if (reinterpret_cast<uintptr_t>(op.ptr) % 16 != 0) return 1;
  if (!even(op.s_ko) || !even(op.s_z1) || !even(op.s_z2)) return 1;
  const bool flat = op.Lr >= rows;
  if (!flat && !even(op.s_ro)) return 1;
  if (kmajor) {
    if (!even(op.s_ri)) return 1;
  } else {
    if (!even(op.s_ki)) return 1;
    if (!flat && (op.Lr & 1)) return 1;
  }
  return 2;
}

struct Plan {
  int cfg, splitk, tiles_m, tiles_n;
  bool swap;
};

// Relative throughput of the tile shapes measured on B200 (scratch/gemm_cfgs.py, 4096^3 and the BSE shapes):
// 128x64 with two co-resident CTAs per SM is the fastest; the narrow tiles exist to fit skinny dimensions.
constexpr double kCfgEff[6] = {0.88, 0.55, 0.60, 0.90, 1.00, 0.85};

// Estimated run time (seconds) of one (orientation, tile shape, split-K) choice: CTAs are dealt to the
// num_sms * occ resident slots wave by wave, a CTA's work is its tile area times its K range plus a fixed
// prologue/epilogue equivalent, and a split needs one more pass over the partial sums.
double plan_cost(const GemmParams& p, int num_sms, int cfg, bool swap, int splitk, long long T_total) {
  const int M = swap ? p.N : p.M, N = swap ? p.M : p.N;
  const long long tm = ceil_div(M, kCfg[cfg].BM), tn = ceil_div(N, kCfg[cfg].BN);
  long long tiles = tm * tn;
  if (p.lower_only) tiles = (tiles + tm) / 2;
  tiles *= (long long)p.Z1 * p.Z2;
  const long long ctas = tiles * splitk;
  const long long slots = (long long)num_sms * kCfg[cfg].occ;
  const long long waves = ceil_div(ctas, slots);
  const double ksteps = (double)ceil_div<long long>(T_total, splitk) + 5.0;  // +5: pipeline fill and epilogue
  const double flops_cta = 2.0 * kCfg[cfg].BM * kCfg[cfg].BN * ksteps * GEMM_BK;
  const double rate_sm = 30.0e12 / 148.0 * kCfgEff[cfg];  // measured tile-kernel rate per SM
  double t = (double)waves * kCfg[cfg].occ * flops_cta / rate_sm;
  if (splitk > 1)
    t += 3.0e-6 + 16.0 * (double)tm * kCfg[cfg].BM * (double)tn * kCfg[cfg].BN * p.Z1 * p.Z2 * splitk / 5.0e12;
  if (swap) t *= 1.02;
  return t;
}

Plan make_plan(const GemmParams& p, int num_sms, int force_cfg, int force_splitk) {
  Plan pl{};
  const bool can_swap = force_cfg < 0 && !p.w && !p.nscale && !p.lower_only;
  const long long T_total = (long long)p.Ko * ceil_div(p.Ki, GEMM_BK);
  static const int kSplits[] = {1, 2, 3, 4, 5, 6, 7, 8, 10, 12, 14, 16, 20, 24, 28, 32, 40, 48, 56, 64};
  double best = 1e300;
  for (int sw = 0; sw <= (can_swap ? 1 : 0); ++sw) {
    for (int c : {4, 5, 1, 2}) {
      if (force_cfg >= 0) c = force_cfg;
      for (int sk : kSplits) {
        if (force_splitk > 0) sk = force_splitk;
        if (sk > 1 && (long long)sk * 4 > T_total && force_splitk <= 0) break;
        const double t = plan_cost(p, num_sms, c, sw != 0, sk, T_total);
        if (t < best) {
          best = t;
          pl.cfg = c;
          pl.swap = sw != 0;
          pl.splitk = sk;
        }
        if (force_splitk > 0) break;
      }
      if (force_cfg >= 0) break;
    }
  }
  const int M = pl.swap ? p.N : p.M, N = pl.swap ? p.M : p.N;
  pl.tiles_m = ceil_div(M, kCfg[pl.cfg].BM);
  pl.tiles_n = ceil_div(N, kCfg[pl.cfg].BN);
  pl.splitk = (int)std::min<long long>(pl.splitk, std::max<long long>(1, T_total));
  return pl;
}

GemmParams apply_swap(GemmParams p) {
  std::swap(p.A, p.B);
  std::swap(p.M, p.N);
  std::swap(p.sC_mi, p.sC_ni);
  std::swap(p.sC_mo, p.sC_no);
  std::swap(p.Lm, p.Ln);
  return p;
}

}  // namespace

// force_cfg: -1 auto (TMA kernel when the operands allow it), 0..5 cp.async tile shapes, 10..13 TMA tile shapes
static bool wants_tma(const GemmParams& p, int num_sms, int& force_cfg, int force_splitk) {
  if (force_cfg >= 15 || (force_cfg >= 6 && force_cfg < 10)) force_cfg = -1;  // not a tile shape of either kernel
  if (force_cfg >= 0 && force_cfg < 10) return false;
  int c0, sw0, sk0;
  if (!gemm_tma_describe(p, num_sms, force_cfg >= 10 ? force_cfg - 10 : -1, force_splitk, &c0, &sw0, &sk0)) {
    force_cfg = -1;  // a TMA tile shape was asked for but the operands need the cp.async kernel
    return false;
  }
  return true;
}

size_t gemm_ws_bytes_needed(const GemmParams& p, int num_sms, int force_cfg, int force_splitk) {
  if (wants_tma(p, num_sms, force_cfg, force_splitk)) {
    // the cp.async kernel is the fallback if a tensor map cannot be encoded: reserve for whichever needs more
    const size_t t = gemm_tma_ws_bytes(p, num_sms, force_cfg >= 10 ? force_cfg - 10 : -1, force_splitk);
    Plan pl = make_plan(p, num_sms, -1, force_splitk);
    const size_t o = pl.splitk <= 1 ? 0
                                    : sizeof(double) * (size_t)pl.tiles_m * kCfg[pl.cfg].BM * pl.tiles_n *
                                          kCfg[pl.cfg].BN * pl.splitk * p.Z1 * p.Z2;
    return std::max(t, o);
  }
  Plan pl = make_plan(p, num_sms, force_cfg, force_splitk);
  if (pl.splitk <= 1) return 0;
  return sizeof(double) * (size_t)pl.tiles_m * kCfg[pl.cfg].BM * pl.tiles_n * kCfg[pl.cfg].BN * pl.splitk * p.Z1 *
         p.Z2;
}

void gemm_plan_describe(const GemmParams& p, int num_sms, int force_cfg, int force_splitk, int* cfg, int* swap,
                        int* splitk) {
  if (wants_tma(p, num_sms, force_cfg, force_splitk)) {
    gemm_tma_describe(p, num_sms, force_cfg >= 10 ? force_cfg - 10 : -1, force_splitk, cfg, swap, splitk);
    *cfg += 10;
    return;
  }
  Plan pl = make_plan(p, num_sms, force_cfg, force_splitk);
  *cfg = pl.cfg;
  *swap = pl.swap ? 1 : 0;
  *splitk = pl.splitk;
}

void gemm_launch(GemmParams p, cudaStream_t stream, double* ws, size_t ws_bytes, int num_sms, int force_cfg,
                 int force_splitk) {
  if (p.M <= 0 || p.N <= 0 || p.Z1 <= 0 || p.Z2 <= 0) return;
  GW_REQUIRE(p.Ko >= 1 && p.Ki >= 0, "bad K extents");
  if (wants_tma(p, num_sms, force_cfg, force_splitk)) {
    if (gemm_tma_try_launch(p, stream, ws, ws_bytes, num_sms, force_cfg >= 10 ? force_cfg - 10 : -1, force_splitk)) return;
    force_cfg = -1;  // tensor map not encodable: cp.async kernel
  }
  Plan pl = make_plan(p, num_sms, force_cfg, force_splitk);
  if (pl.swap) p = apply_swap(p);
  const bool ak = is_kmajor(p.A, p.Ki), bk = is_kmajor(p.B, p.Ki);
  p.A.vec = deduce_vec(p.A, ak, p.M);
  p.B.vec = deduce_vec(p.B, bk, p.N);
  p.tiles_m = pl.tiles_m;
  p.tiles_n = pl.tiles_n;
  p.splitk = pl.splitk;
  if (pl.splitk > 1) {
    const size_t need = sizeof(double) * (size_t)pl.tiles_m * kCfg[pl.cfg].BM * pl.tiles_n * kCfg[pl.cfg].BN *
                        pl.splitk * p.Z1 * p.Z2;
    if (ws == nullptr || need > ws_bytes) {
      // fall back to fewer splits that fit (or none)
      const size_t per = need / pl.splitk;
      int fit = per ? (int)(ws_bytes / per) : 0;
      p.splitk = pl.splitk = std::max(1, std::min(pl.splitk, ws ? fit : 1));
    }
    p.ws = ws;
  }
  const long long gz = (long long)p.Z1 * p.Z2 * p.splitk;
  GW_REQUIRE(gz <= 65535 && (long long)pl.tiles_m * pl.tiles_n < (1LL << 31), "GEMM grid too large");
  {
    // group height minimising the bytes one wave of resident CTAs streams: group_m*BM + (slots/group_m)*BN
    const double slots = (double)num_sms * kCfg[pl.cfg].occ;
    double gm = std::sqrt(slots * kCfg[pl.cfg].BN / kCfg[pl.cfg].BM);
    gm = std::max(gm, slots / pl.tiles_n);
    p.group_m = (int)std::max(1.0, std::min<double>(pl.tiles_m, std::round(gm)));
  }
  dim3 grid((unsigned)(pl.tiles_m * pl.tiles_n), 1, (unsigned)gz);
  switch (pl.cfg) {
    case 0:
      launch_cfg<128, 128, 2, 4, 4, 1>(p, grid, ak, bk, stream);
      break;
    case 1:
      launch_cfg<128, 32, 4, 1, 4, 2>(p, grid, ak, bk, stream);
      break;
    case 3:
      launch_cfg<128, 128, 4, 4, 4, 1>(p, grid, ak, bk, stream);
      break;
    case 4:
      launch_cfg<128, 64, 4, 2, 3, 2>(p, grid, ak, bk, stream);
      break;
    case 5:
      launch_cfg<128, 48, 4, 1, 3, 2>(p, grid, ak, bk, stream);
      break;
    default:
      launch_cfg<64, 64, 2, 2, 4, 2>(p, grid, ak, bk, stream);
      break;
  }
  if (p.splitk > 1) {
    dim3 rb(128);
    dim3 rg(ceil_div(p.M, 128), std::min(p.N, 65535), p.Z1 * p.Z2);
    GW_REQUIRE(p.Z1 * p.Z2 <= 65535, "split-K reduce grid too large");
    switch (pl.cfg) {
      case 0:
        gemm_splitk_reduce_kernel<128, 128><<<rg, rb, 0, stream>>>(p);
        break;
      case 1:
        gemm_splitk_reduce_kernel<128, 32><<<rg, rb, 0, stream>>>(p);
        break;
      case 3:
        gemm_splitk_reduce_kernel<128, 128><<<rg, rb, 0, stream>>>(p);
        break;
      case 4:
        gemm_splitk_reduce_kernel<128, 64><<<rg, rb, 0, stream>>>(p);
        break;
      case 5:
        gemm_splitk_reduce_kernel<128, 48><<<rg, rb, 0, stream>>>(p);
        break;
      default:
        gemm_splitk_reduce_kernel<64, 64><<<rg, rb, 0, stream>>>(p);
        break;
    }
    GW_CUDA(cudaGetLastError());
  }
}

void gemm_init_attributes() {}

}  // namespace gwbse
