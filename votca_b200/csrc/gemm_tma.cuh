// TMA-staged, warp-specialised, persistent FP64 DMMA GEMM for sm_100a.
//
// Same contraction as gemm_dmma.cuh (C (+)= alpha * nscale[n] * sum_k w[k] A(m;k) B(n;k)) for operands whose rows
// are a plain (non-compound) index and whose strides are 16-byte multiples - the shapes that carry most of the
// GW-BSE flops: MultiplyRightWithAuxMatrix (2-D), the epsilon SYRK (3-D box: chi x (v, c)), the Fill3cMO
// contractions and the short-K BSE legs.  Differences to the cp.async kernel:
//   * operand tiles are brought in by cp.async.bulk.tensor (SASS UTMALDG) into a ring of 128-byte-swizzled
//     shared-memory stages, completion on mbarriers; one producer warp issues, it never touches the math;
//   * the consumer warps only wait on the stage's "full" barrier and arrive on its "empty" barrier - there is no
//     block-wide barrier in the main loop, warps drift apart and keep the DMMA pipe fed;
//   * the kernel is persistent: grid = resident CTAs, a static tile schedule; the producer runs ahead into the
//     next tile while the consumers are in the epilogue;
//   * TMA zero-fills out-of-range rows / k, so there is no predication in the loop.
// Fragment loads are bank-conflict free for both tile layouts because the 4 k values a DMMA step consumes are a
// permutation chosen against the 128-byte swizzle (see kperm below); A and B use the same permutation, so the
// sum over k is unchanged.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include "gemm_dmma.cuh"

namespace gwbse {

constexpr int TMA_BK = 16;  // one 128-byte swizzle row of doubles

// How the coordinates of a tile load are formed for one operand.  Tensor-map dimensions are always
//   K-major: (k, ri, ro, ko, z1)      box (16, ROWS, 1, 1, 1)   one load per tile
//   M-major: (ri, k, ro, ko, z1)      box (16, 16, 1, 1, 1)     ROWS/16 loads per tile
// The row index may be compound, row = ro * Lr + ri (the (slice, chi) index of the BSE intermediate): tiles then
// never straddle an ro boundary (tiles_per_ro tiles per ro, the last one zero-filled by TMA beyond Lr).
// row0/k0 are coordinate offsets of the view inside the tensor map (multiples of two elements: TMA wants every box
// start 16-byte aligned); dimensions that do not exist (or are broadcast) have extent 1 and their coordinate
// multiplier is 0.
struct TmaOperand {
  int row0 = 0, k0 = 0;
  int use_ko = 0, use_z1 = 0;
  int Lr = 1 << 30, tiles_per_ro = 1 << 30;
};

struct alignas(64) GemmTmaParams {
  CUtensorMap mapA, mapB, mapW;
  TmaOperand A, B;
  int w_use_ko = 0, w_use_z1 = 0;
  GemmParams g;       // shapes, C addressing, epilogue, split-K workspace (operand pointers unused)
  long long items = 0;  // tiles (lower triangle only when lower_only) * splitk * Z1 * Z2
  int ntiles = 0;       // tiles per (z, split)
};

// ---------------------------------------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ unsigned smem_u32(const void* p) { return static_cast<unsigned>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
// bounded spin: a protocol error traps (kernel fails with an error) instead of hanging the device
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  const unsigned a = smem_u32(bar);
  unsigned done = 0;
  for (unsigned spins = 0; !done; ++spins) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(a), "r"(parity)
        : "memory");
    if (!done && spins > (1u << 26)) __trap();
  }
}
__device__ __forceinline__ void tma_load_5d(void* smem, const CUtensorMap* map, unsigned long long* bar, int c0, int c1,
                                            int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];\n" ::
          "r"(smem_u32(smem)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
#ifndef GWBSE_TMA_NO_PREFETCH
  asm volatile("prefetch.tensormap [%0];\n" ::"l"(map) : "memory");
#endif
}

// ---------------------------------------------------------------------------------------------- tile schedule
// item -> (tile_m, tile_n, split, z1, z2); the same enumeration in the producer and in the consumers.
template <int BM, int BN>
__device__ __forceinline__ void tma_decode_item(const GemmTmaParams& P, long long item, int& tile_m, int& tile_n,
                                                int& split, int& z1, int& z2) {
  const GemmParams& p = P.g;
  int t = static_cast<int>(item % P.ntiles);
  long long rest = item / P.ntiles;
  split = static_cast<int>(rest % p.splitk);
  rest /= p.splitk;
  z1 = static_cast<int>(rest % p.Z1);
  z2 = static_cast<int>(rest / p.Z1);
  if (p.lower_only) {
    // row-major over the tiles that touch the lower triangle
    int tm = 0;
    for (;; ++tm) {
      const int cnt = min(p.tiles_n, (tm * BM + BM - 1) / BN + 1);
      if (t < cnt) break;
      t -= cnt;
    }
    tile_m = tm;
    tile_n = t;
  } else {
    const int per_group = p.group_m * p.tiles_n;
    const int group = t / per_group, in_group = t - group * per_group;
    const int first_m = group * p.group_m;
    const int gm = min(p.group_m, p.tiles_m - first_m);
    tile_n = in_group / gm;
    tile_m = first_m + (in_group - tile_n * gm);
  }
}

template <int BM, int BN, int STAGES, bool HASW>
struct TmaSmem {
  static constexpr int A_BYTES = BM * TMA_BK * 8, B_BYTES = BN * TMA_BK * 8;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;  // both multiples of 1024 (BM, BN multiples of 8)
  static constexpr int W_BYTES = HASW ? TMA_BK * 8 : 0;
  static constexpr int TX_BYTES = STAGE_BYTES + W_BYTES;
  static constexpr size_t bytes = 1024 /*alignment slack*/ + (size_t)STAGE_BYTES * STAGES + (size_t)W_BYTES * STAGES +
                                  16 * STAGES;
};

// ---------------------------------------------------------------------------------------------- the kernel
// WGM x WGN consumer warps (warp tile BM/WGM x BN/WGN; one or two warpgroups) + one producer warpgroup of which one
// lane works.  The register file is handed out per warpgroup: the producer warpgroup gives its registers back
// (setmaxnreg.dec) and the consumer warpgroups take them (setmaxnreg.inc) - a 64x32 warp tile needs 128 accumulator
// registers, more than an even split of the register file over all threads leaves.
template <int REGS>
__device__ __forceinline__ void setmaxnreg_inc() {
#ifndef GWBSE_TMA_NO_SETMAXNREG
  asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;\n" ::"n"(REGS));
#endif
}
template <int REGS>
__device__ __forceinline__ void setmaxnreg_dec() {
#ifndef GWBSE_TMA_NO_SETMAXNREG
  asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;\n" ::"n"(REGS));
#endif
}

template <int BM, int BN, int WGM, int WGN, int STAGES, int MINB, bool AK, bool BKM, bool HASW>
__global__ void __launch_bounds__((WGM * WGN + 4) * 32, MINB) gemm_tma_kernel(const __grid_constant__ GemmTmaParams P) {
  constexpr int NCW = WGM * WGN;
  static_assert(NCW % 4 == 0, "consumer warps come in warpgroups");
  // registers per thread the kernel is compiled with (ptxas rounds the resident threads to 128 and the count down
  // to a multiple of 8), what the producer warpgroup keeps and what each consumer thread ends up with
  constexpr int NTHR = (NCW + 4) * 32;
  constexpr int REG_ALL = (65536 / (NTHR * MINB)) / 8 * 8 > 255 ? 248 : (65536 / (NTHR * MINB)) / 8 * 8;
  constexpr int REG_PROD = 40;
  constexpr int REG_CONS_RAW = REG_ALL + (REG_ALL - REG_PROD) * 4 / NCW;
  constexpr int REG_CONS = (REG_CONS_RAW > 232 ? 232 : REG_CONS_RAW) / 8 * 8;
  constexpr int WTM = BM / WGM, WTN = BN / WGN;
  constexpr int MI = WTM / 8, NI = WTN / 8;
  static_assert(WTM % 16 == 0 && WTN % 16 == 0 && BM % 16 == 0 && BN % 16 == 0, "warp tiles must be multiples of 16");
  using SM = TmaSmem<BM, BN, STAGES, HASW>;
  const GemmParams& p = P.g;

  extern __shared__ unsigned char smem_raw_tma[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw_tma) + 1023) & ~uintptr_t(1023));
  unsigned char* sW_all = smem + (size_t)SM::STAGE_BYTES * STAGES;
  unsigned long long* full_bar = reinterpret_cast<unsigned long long*>(sW_all + (size_t)SM::W_BYTES * STAGES);
  unsigned long long* empty_bar = full_bar + STAGES;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
#pragma unroll
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar + s, 1);
      mbar_init(empty_bar + s, NCW);
    }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    tma_prefetch_desc(&P.mapA);
    tma_prefetch_desc(&P.mapB);
    if (HASW) tma_prefetch_desc(&P.mapW);
  }
  __syncthreads();

  const int tiles_per_ko = (p.Ki + TMA_BK - 1) / TMA_BK;
  const int T_total = p.Ko * tiles_per_ko;
  const int per_split = (T_total + p.splitk - 1) / p.splitk;

  if (warp >= NCW) {
    // =============================== producer: one lane issues every TMA load ===============================
    setmaxnreg_dec<REG_PROD>();
    if (warp != NCW || lane != 0) return;
    int stage = 0;
    unsigned phase = 0;
    for (long long item = blockIdx.x; item < P.items; item += gridDim.x) {
      int tile_m, tile_n, split, z1, z2;
      tma_decode_item<BM, BN>(P, item, tile_m, tile_n, split, z1, z2);
      const int t_begin = split * per_split, t_end = min(T_total, t_begin + per_split);
      int ko = t_begin / tiles_per_ko;
      int k0 = (t_begin - ko * tiles_per_ko) * TMA_BK;
      const int roA = tile_m / P.A.tiles_per_ro, roB = tile_n / P.B.tiles_per_ro;
      const int rowA = P.A.row0 + (tile_m - roA * P.A.tiles_per_ro) * BM;
      const int rowB = P.B.row0 + (tile_n - roB * P.B.tiles_per_ro) * BN;
      const int zA = z1 * P.A.use_z1, zB = z1 * P.B.use_z1;
      for (int t = t_begin; t < t_end; ++t) {
        mbar_wait(empty_bar + stage, phase ^ 1u);
        unsigned char* sA = smem + (size_t)stage * SM::STAGE_BYTES;
        unsigned char* sB = sA + SM::A_BYTES;
        mbar_expect_tx(full_bar + stage, SM::TX_BYTES);
        if (AK) {
          tma_load_5d(sA, &P.mapA, full_bar + stage, P.A.k0 + k0, rowA, roA, ko * P.A.use_ko, zA);
        } else {
#pragma unroll
          for (int b = 0; b < BM / 16; ++b)
            tma_load_5d(sA + b * 2048, &P.mapA, full_bar + stage, rowA + 16 * b, P.A.k0 + k0, roA, ko * P.A.use_ko, zA);
        }
        if (BKM) {
          tma_load_5d(sB, &P.mapB, full_bar + stage, P.B.k0 + k0, rowB, roB, ko * P.B.use_ko, zB);
        } else {
#pragma unroll
          for (int b = 0; b < BN / 16; ++b)
            tma_load_5d(sB + b * 2048, &P.mapB, full_bar + stage, rowB + 16 * b, P.B.k0 + k0, roB, ko * P.B.use_ko, zB);
        }
        if (HASW)
          tma_load_5d(sW_all + stage * SM::W_BYTES, &P.mapW, full_bar + stage, k0, ko * P.w_use_ko, z1 * P.w_use_z1, 0, 0);
        k0 += TMA_BK;
        if (k0 >= p.Ki) {
          k0 = 0;
          ++ko;
        }
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1u;
        }
      }
    }
    return;
  }

  // =============================== consumers ===============================
  setmaxnreg_inc<REG_CONS>();
  const int g = lane >> 2, t4 = lane & 3;
  const int wm0 = (warp % WGM) * WTM, wn0 = (warp / WGM) * WTN;
  // k permutation of one DMMA step kk (0..3): lane t4 takes k = 2 * (c0 ^ kk) + p0, c0 = {0,5,2,7}[t4],
  // p0 = t4 >> 1.  The 16 k values of a stage are each used exactly once; against the 128-byte swizzle
  // (16-byte chunk index ^= row % 8) the 16 lanes of a half-warp then hit 16 distinct 8-byte bank pairs, for
  // K-major tiles ([row][k], chunk = k/2 ^ row%8) and for M-major tiles ([row/16][k][row%16], chunk = (row%16)/2 ^ k%8).
  const int c0 = (t4 == 0) ? 0 : (t4 == 1) ? 5 : (t4 == 2) ? 2 : 7;
  const int p0 = t4 >> 1;
  int offA[4], offB[4], kidx[4];
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) {
    const int c = c0 ^ kk, k = 2 * c + p0;
    kidx[kk] = k;
    // byte offsets of the lane's element for row offset 0 of its warp tile (rows advance by immediates below)
    offA[kk] = AK ? ((wm0 + g) * 128 + ((c ^ g) << 4) + (p0 << 3))
                  : ((wm0 >> 4) * 2048 + k * 128 + ((((g >> 1) ^ (k & 7))) << 4) + ((g & 1) << 3));
    offB[kk] = BKM ? ((wn0 + g) * 128 + ((c ^ g) << 4) + (p0 << 3))
                   : ((wn0 >> 4) * 2048 + k * 128 + ((((g >> 1) ^ (k & 7))) << 4) + ((g & 1) << 3));
  }

  int stage = 0;
  unsigned phase = 0;
  for (long long item = blockIdx.x; item < P.items; item += gridDim.x) {
    int tile_m, tile_n, split, z1, z2;
    tma_decode_item<BM, BN>(P, item, tile_m, tile_n, split, z1, z2);
    const int t_begin = split * per_split, t_end = min(T_total, t_begin + per_split);
    const int m_base = tile_m * BM, n_base = tile_n * BN;  // in the padded tile grid (split-K workspace)
    // first global row / column of the tile and how many of its rows / columns exist
    const int roA = tile_m / P.A.tiles_per_ro, roB = tile_n / P.B.tiles_per_ro;
    const int riA = (tile_m - roA * P.A.tiles_per_ro) * BM, riB = (tile_n - roB * P.B.tiles_per_ro) * BN;
    const int m_glob = roA * min(P.A.Lr, p.M) + riA, n_glob = roB * min(P.B.Lr, p.N) + riB;
    const int m_lim = min(min(P.A.Lr, p.M) - riA, p.M - m_glob), n_lim = min(min(P.B.Lr, p.N) - riB, p.N - n_glob);

    double acc[MI][NI][2];
#pragma unroll
    for (int i = 0; i < MI; ++i)
#pragma unroll
      for (int j = 0; j < NI; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    for (int t = t_begin; t < t_end; ++t) {
      mbar_wait(full_bar + stage, phase);
      const unsigned char* sA = smem + (size_t)stage * SM::STAGE_BYTES;
      const unsigned char* sB = sA + SM::A_BYTES;
      const double* sW = reinterpret_cast<const double*>(sW_all + stage * SM::W_BYTES);
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        double a[MI], b[NI];
#if defined(GWBSE_TMA_PROBE_NO_LDS)
#pragma unroll
        for (int i = 0; i < MI; ++i) a[i] = 1.0 + i + kk + 1e-9 * lane;
#pragma unroll
        for (int j = 0; j < NI; ++j) b[j] = 0.5 + j - kk + 1e-9 * lane;
        if (false)
#endif
#pragma unroll
        for (int i = 0; i < MI; ++i) {
          // K-major: rows advance by 8 -> +1024 B.  M-major: row%16 = (i&1)*8 + g flips chunk bit 2 (byte 64),
          // row/16 advances every second i (+2048 B).
          const int o = AK ? offA[kk] + i * 1024 : ((offA[kk] ^ ((i & 1) << 6)) + (i >> 1) * 2048);
          a[i] = *reinterpret_cast<const double*>(sA + o);
        }
        if (HASW) {
          const double wv = sW[kidx[kk]];
#pragma unroll
          for (int i = 0; i < MI; ++i) a[i] *= wv;
        }
#if defined(GWBSE_TMA_PROBE_NO_LDS)
        if (false)
#endif
#pragma unroll
        for (int j = 0; j < NI; ++j) {
          const int o = BKM ? offB[kk] + j * 1024 : ((offB[kk] ^ ((j & 1) << 6)) + (j >> 1) * 2048);
          b[j] = *reinterpret_cast<const double*>(sB + o);
        }
#if defined(GWBSE_TMA_ORDER_JI)
#pragma unroll
        for (int j = 0; j < NI; ++j)
#pragma unroll
          for (int i = 0; i < MI; ++i) dmma884(acc[i][j], a[i], b[j]);
#elif defined(GWBSE_TMA_ORDER_SNAKE)
#pragma unroll
        for (int i = 0; i < MI; ++i)
#pragma unroll
          for (int jj = 0; jj < NI; ++jj) {
            const int j = (i & 1) ? NI - 1 - jj : jj;
            dmma884(acc[i][j], a[i], b[j]);
          }
#else
#pragma unroll
        for (int i = 0; i < MI; ++i)
#pragma unroll
          for (int j = 0; j < NI; ++j) dmma884(acc[i][j], a[i], b[j]);
#endif
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(empty_bar + stage);
      if (++stage == STAGES) {
        stage = 0;
        phase ^= 1u;
      }
    }

    // ------------------------------ epilogue (same addressing as gemm_dmma_kernel) ------------------------------
    if (p.splitk > 1) {
      const long long Mpad = (long long)p.tiles_m * BM, Npad = (long long)p.tiles_n * BN;
      double* wsp = p.ws + ((long long)(z2 * p.Z1 + z1) * p.splitk + split) * Mpad * Npad;
#pragma unroll
      for (int i = 0; i < MI; ++i) {
        const long long row = m_base + wm0 + i * 8 + g;
#pragma unroll
        for (int j = 0; j < NI; ++j) {
          const long long col = n_base + wn0 + j * 8 + 2 * t4;
          wsp[col * Mpad + row] = acc[i][j][0];
          wsp[(col + 1) * Mpad + row] = acc[i][j][1];
        }
      }
      continue;
    }
    double* Cz = p.C + z1 * p.sC_z1 + z2 * p.sC_z2;
#pragma unroll
    for (int j = 0; j < NI; ++j)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int cl = wn0 + j * 8 + 2 * t4 + e;
        if (cl >= n_lim) continue;
        const int col = n_glob + cl;
        const long long coff = (long long)(col / p.Ln) * p.sC_no + (long long)(col % p.Ln) * p.sC_ni;
        const double csc = p.alpha * (p.nscale ? p.nscale[p.nscale_mod ? col % p.nscale_mod : col] : 1.0);
#pragma unroll
        for (int i = 0; i < MI; ++i) {
          const int rl = wm0 + i * 8 + g;
          if (rl >= m_lim) continue;
          const int row = m_glob + rl;
          double* dst = Cz + (long long)(row / p.Lm) * p.sC_mo + (long long)(row % p.Lm) * p.sC_mi + coff;
          double v = csc * acc[i][j][e];
          if (p.beta != 0.0) v += p.beta * (*dst);
          *dst = v;
        }
      }
  }
}

// Host side (gemm_tma.cu): returns false when the operands cannot be described by tensor maps (compound row
// index, odd strides, unaligned base) or the driver entry point is missing - the caller then uses the cp.async kernel.
bool gemm_tma_try_launch(const GemmParams& p, cudaStream_t stream, double* ws, size_t ws_bytes, int num_sms,
                         int force_cfg, int force_splitk);
size_t gemm_tma_ws_bytes(const GemmParams& p, int num_sms, int force_cfg, int force_splitk);
// "tma cfgN skK" or "" when the TMA kernel would not be used
bool gemm_tma_describe(const GemmParams& p, int num_sms, int force_cfg, int force_splitk, int* cfg, int* swap, int* splitk);
void gemm_tma_set_enabled(bool on);
bool gemm_tma_enabled();

}  // namespace gwbse
