// FP64 tensor-core (DMMA) GEMM for sm_100a with generalised operand addressing.
//
// One kernel family serves every dense contraction of the GW-BSE path
// (a3 Fill3cMO, a5 MultiplyRight, a7/a8 epsilon SYRK,
// a9 ApB, a10 Sigma_x, a11 residues, a17 BSE matvec, a20 Davidson projections):
//
//   C[m,n] (+)= alpha * nscale[n] * sum_{ko,ki} w[ko,ki] * A(m; ko,ki) * B(n; ko,ki)
//
// * A and B live in global memory with *separable* addressing
//     addr(r, ko, ki) = ptr + (r / Lr) * s_ro + (r % Lr) * s_ri + ko * s_ko + ki * s_ki
//   so sub-blocks of the Mmn tensor (middleRows views, per-level slices, the
//   compound (v,c) index of the RPA / BSE sums) are consumed in place.
//   Either the k index is contiguous ("K-major", s_ki == 1) or the row index is
//   ("M-major", s_ri == 1); both are staged without transposition.
// * Tiles are staged global -> shared with cp.async (LDGSTS, 16 B when the
//   operand is 16 B aligned, else 8 B), 4-stage ring, one barrier per k-tile.
// * Math is mma.sync.m8n8k4.f64 (SASS DMMA.8x8x4 - the only FP64 tensor shape
//   sm_100a implements; larger PTX shapes lower to it).  Measured issue peak on
//   B200: 37.1 TFLOP/s (profiles/r01_fp64_peak_probe.txt).
// * Shared tiles are padded so that fragment loads (one double per lane,
//   8 rows x 4 k) are bank-conflict free in both layouts.
#pragma once
#include <cuda_runtime.h>

#include "common.cuh"

namespace gwbse {

struct GemmOperand {
  const double* ptr = nullptr;
  long long s_ri = 1, s_ro = 0;  // row index r = ro * Lr + ri
  long long s_ki = 1, s_ko = 0;  // k index (ko, ki), ki in [0, Ki)
  long long s_z1 = 0, s_z2 = 0;  // batch strides
  int Lr = 1 << 30;              // inner row length
  int vec = 1;                   // doubles per cp.async (1 or 2)
};

struct GemmParams {
  GemmOperand A, B;
  int M = 0, N = 0, Ko = 1, Ki = 0, Z1 = 1, Z2 = 1;
  double* C = nullptr;
  long long sC_mi = 1, sC_mo = 0, sC_ni = 0, sC_no = 0, sC_z1 = 0, sC_z2 = 0;
  int Lm = 1 << 30, Ln = 1 << 30;
  double alpha = 1.0, beta = 0.0;
  const double* w = nullptr;  // optional weights on A's k index
  long long sW_ko = 0, sW_z1 = 0, sW_z2 = 0;
  const double* nscale = nullptr;  // optional per-column scale, indexed by col (or col % nscale_mod)
  int nscale_mod = 0;
  int lower_only = 0;              // SYRK: skip tiles strictly above the diagonal
  int splitk = 1;
  double* ws = nullptr;  // split-K partials
  int tiles_m = 0, tiles_n = 0;
  int group_m = 1;  // rasterisation: consecutive CTAs sweep group_m m-tiles x all n-tiles (L2 reuse of both panels)
};

constexpr int GEMM_BK = 16;
constexpr int GEMM_LDK = GEMM_BK + 4;  // padded k extent of a K-major tile row

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, int src_bytes) {
  unsigned s = static_cast<unsigned>(__cvta_generic_to_shared(smem));
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async8(void* smem, const void* gmem, int src_bytes) {
  unsigned s = static_cast<unsigned>(__cvta_generic_to_shared(smem));
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(s), "l"(gmem), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

__device__ __forceinline__ void dmma884(double (&c)[2], double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c[0]), "+d"(c[1])
               : "d"(a), "d"(b));
}

// ---------------------------------------------------------------------------
// global -> shared staging of one operand tile (ROWS x 16 doubles)
// ---------------------------------------------------------------------------
template <int ROWS, bool KMAJOR, int NT>
__device__ __forceinline__ void stage_tile(double* s, const double* base, const GemmOperand& op,
                                           const long long* rowoff, int k0, int Ki, int tid) {
  // generic path: row offsets come from the shared table every time.  Trip counts are compile-time constants and
  // the loops are fully unrolled so that the copies can be scheduled between the DMMAs (operands with odd row
  // strides - AO blocks with odd N, trial vectors with odd ct - can only use 8-byte copies).
  if (KMAJOR) {
    // s[r * LDK + k]
    if (op.vec == 2) {
      constexpr int TOTAL = ROWS * (GEMM_BK / 2);
#pragma unroll
      for (int c = 0; c < (TOTAL + NT - 1) / NT; ++c) {
        const int idx = tid + c * NT;
        if (TOTAL % NT != 0 && idx >= TOTAL) break;
        const int r = idx >> 3, k = (idx & 7) * 2;
        const long long off = rowoff[r];
        const int rem = Ki - (k0 + k);
        const int bytes = (off >= 0 && rem > 0) ? (rem >= 2 ? 16 : 8) : 0;
        const double* src = bytes ? base + off + (k0 + k) : base;
        cp_async16(s + r * GEMM_LDK + k, src, bytes);
      }
    } else {
      constexpr int TOTAL = ROWS * GEMM_BK;
#pragma unroll
      for (int c = 0; c < (TOTAL + NT - 1) / NT; ++c) {
        const int idx = tid + c * NT;
        if (TOTAL % NT != 0 && idx >= TOTAL) break;
        const int r = idx >> 4, k = idx & 15;
        const long long off = rowoff[r];
        const int bytes = (off >= 0 && k0 + k < Ki) ? 8 : 0;
        const double* src = bytes ? base + off + (k0 + k) : base;
        cp_async8(s + r * GEMM_LDK + k, src, bytes);
      }
    }
  } else {
    // s[k * (ROWS + 4) + r]
    constexpr int LD = ROWS + 4;
    if (op.vec == 2) {
      constexpr int VPR = ROWS / 2;
      constexpr int TOTAL = GEMM_BK * VPR;
#pragma unroll
      for (int c = 0; c < (TOTAL + NT - 1) / NT; ++c) {
        const int idx = tid + c * NT;
        if (TOTAL % NT != 0 && idx >= TOTAL) break;
        const int k = idx / VPR, r = (idx % VPR) * 2;
        const long long off = rowoff[r];
        const bool kin = (k0 + k) < Ki;
        const int bytes = (off >= 0 && kin) ? (rowoff[r + 1] >= 0 ? 16 : 8) : 0;
        const double* src = bytes ? base + off + (long long)(k0 + k) * op.s_ki : base;
        cp_async16(s + k * LD + r, src, bytes);
      }
    } else {
      constexpr int TOTAL = GEMM_BK * ROWS;
#pragma unroll
      for (int c = 0; c < (TOTAL + NT - 1) / NT; ++c) {
        const int idx = tid + c * NT;
        if (TOTAL % NT != 0 && idx >= TOTAL) break;
        const int k = idx / ROWS, r = idx % ROWS;
        const long long off = rowoff[r];
        const int bytes = (off >= 0 && (k0 + k) < Ki) ? 8 : 0;
        const double* src = bytes ? base + off + (long long)(k0 + k) * op.s_ki : base;
        cp_async8(s + k * LD + r, src, bytes);
      }
    }
  }
}

// Fast path for 16-byte aligned operands: every thread owns the same (row, k) chunks of each k-tile, so the
// row offsets, validity and shared-memory destinations are resolved once and a k-tile costs one pointer add
// and one cp.async per chunk (straight-line code the compiler can interleave with the DMMAs).
template <int ROWS, bool KMAJOR, int NT>
struct Stager {
  static constexpr int TOTAL = ROWS * (GEMM_BK / 2);
  static constexpr int NCH = (TOTAL + NT - 1) / NT;
  const double* src[NCH];
  int bytes[NCH];  // 0, 8 or 16: what the rows of the chunk allow
  __device__ __forceinline__ void init(const double* base, const GemmOperand& op, const long long* rowoff, int tid) {
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      const int idx = tid + c * NT;
      src[c] = base;
      bytes[c] = 0;
      if (TOTAL % NT == 0 || idx < TOTAL) {
        if (KMAJOR) {
          const int r = idx >> 3, k = (idx & 7) * 2;
          const long long off = rowoff[r];
          if (off >= 0) {
            src[c] = base + off + k;
            bytes[c] = 16;
          }
        } else {
          constexpr int VPR = ROWS / 2;
          const int k = idx / VPR, r = (idx % VPR) * 2;
          const long long off = rowoff[r];
          if (off >= 0) {
            src[c] = base + off + (long long)k * op.s_ki;
            bytes[c] = rowoff[r + 1] >= 0 ? 16 : 8;
          }
        }
      }
    }
  }
  // koff = ko * s_ko + k0 * s_ki (element offset of the k-tile), krem = Ki - k0
  __device__ __forceinline__ void issue(double* s, long long koff, int krem, int tid) const {
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      const int idx = tid + c * NT;
      if (TOTAL % NT != 0 && idx >= TOTAL) break;
      if (KMAJOR) {
        const int r = idx >> 3, k = (idx & 7) * 2;
        const int rem = krem - k;
        const int b = rem >= 2 ? bytes[c] : (rem == 1 ? (bytes[c] ? 8 : 0) : 0);
        cp_async16(s + r * GEMM_LDK + k, b ? src[c] + koff : src[c], b);
      } else {
        constexpr int VPR = ROWS / 2;
        constexpr int LD = ROWS + 4;
        const int k = idx / VPR, r = (idx % VPR) * 2;
        const int b = k < krem ? bytes[c] : 0;
        cp_async16(s + k * LD + r, b ? src[c] + koff : src[c], b);
      }
    }
  }
};

template <int BM, int BN, int STAGES, bool HASW>
struct GemmSmem {
  static constexpr int SA = BM * GEMM_LDK;  // >= GEMM_BK * (BM + 4)
  static constexpr int SB = BN * GEMM_LDK;
  static constexpr int SW = HASW ? GEMM_BK : 0;
  static constexpr int STAGE = SA + SB + SW;
  static constexpr size_t bytes = sizeof(double) * STAGE * STAGES + sizeof(long long) * (BM + BN);
};

// ---------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------
template <int BM, int BN, int WGM, int WGN, int STAGES, int MINB, bool AK, bool BKM, bool HASW>
__global__ void __launch_bounds__(WGM* WGN * 32, MINB) gemm_dmma_kernel(const GemmParams p) {
  constexpr int NT = WGM * WGN * 32;
  constexpr int WTM = BM / WGM, WTN = BN / WGN;
  constexpr int MI = WTM / 8, NI = WTN / 8;
  using SM = GemmSmem<BM, BN, STAGES, HASW>;
  static_assert(GEMM_BK * (BM + 4) <= SM::SA && GEMM_BK * (BN + 4) <= SM::SB, "tile padding");

  // grouped rasterisation of the linear tile index: a wave of CTAs covers group_m m-tiles x (wave / group_m)
  // n-tiles, so the A and the B panels it streams are both shared through L2
  int tile_m, tile_n;
  {
    const int per_group = p.group_m * p.tiles_n;
    const int group = blockIdx.x / per_group, in_group = blockIdx.x - group * per_group;
    const int first_m = group * p.group_m;
    const int gm = min(p.group_m, p.tiles_m - first_m);
    tile_n = in_group / gm;
    tile_m = first_m + (in_group - tile_n * gm);
  }
  const int m_base = tile_m * BM, n_base = tile_n * BN;
  if (p.lower_only && n_base > m_base + BM - 1) return;

  int zz = blockIdx.z;
  const int split = zz % p.splitk;
  zz /= p.splitk;
  const int z1 = zz % p.Z1, z2 = zz / p.Z1;

  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* smem = reinterpret_cast<double*>(smem_raw);
  long long* offA = reinterpret_cast<long long*>(smem + SM::STAGE * STAGES);
  long long* offB = offA + BM;

  const int tid = threadIdx.x;
  for (int r = tid; r < BM; r += NT) {
    const int R = m_base + r;
    offA[r] = (R < p.M) ? (long long)(R / p.A.Lr) * p.A.s_ro + (long long)(R % p.A.Lr) * p.A.s_ri : -1;
  }
  for (int r = tid; r < BN; r += NT) {
    const int R = n_base + r;
    offB[r] = (R < p.N) ? (long long)(R / p.B.Lr) * p.B.s_ro + (long long)(R % p.B.Lr) * p.B.s_ri : -1;
  }
  __syncthreads();

  const double* Abase = p.A.ptr + z1 * p.A.s_z1 + z2 * p.A.s_z2;
  const double* Bbase = p.B.ptr + z1 * p.B.s_z1 + z2 * p.B.s_z2;
  const double* Wbase = HASW ? p.w + z1 * p.sW_z1 + z2 * p.sW_z2 : nullptr;

  // k-tile range of this split
  const int tiles_per_ko = (p.Ki + GEMM_BK - 1) / GEMM_BK;
  const int T_total = p.Ko * tiles_per_ko;
  const int per = (T_total + p.splitk - 1) / p.splitk;
  const int t_begin = split * per;
  const int t_end = min(T_total, t_begin + per);
  const int T = max(0, t_end - t_begin);

  const bool fastA = p.A.vec == 2, fastB = p.B.vec == 2;
  Stager<BM, AK, NT> stA;
  Stager<BN, BKM, NT> stB;
  if (fastA) stA.init(Abase, p.A, offA, tid);
  if (fastB) stB.init(Bbase, p.B, offB, tid);
  // (ko, k0) of the next k-tile to be issued
  int iss_ko = t_begin / tiles_per_ko;
  int iss_k0 = (t_begin - iss_ko * tiles_per_ko) * GEMM_BK;

  auto issue = [&](int slot) {
    const int ko = iss_ko, k0 = iss_k0;
    double* sA = smem + slot * SM::STAGE;
    double* sB = sA + SM::SA;
    if (fastA)
      stA.issue(sA, (long long)ko * p.A.s_ko + (long long)k0 * p.A.s_ki, p.Ki - k0, tid);
    else
      stage_tile<BM, AK, NT>(sA, Abase + (long long)ko * p.A.s_ko, p.A, offA, k0, p.Ki, tid);
    if (fastB)
      stB.issue(sB, (long long)ko * p.B.s_ko + (long long)k0 * p.B.s_ki, p.Ki - k0, tid);
    else
      stage_tile<BN, BKM, NT>(sB, Bbase + (long long)ko * p.B.s_ko, p.B, offB, k0, p.Ki, tid);
    if (HASW) {
      if (tid < GEMM_BK) {
        const bool ok = (k0 + tid) < p.Ki;
        cp_async8(sB + SM::SB + tid, ok ? Wbase + (long long)ko * p.sW_ko + k0 + tid : Wbase, ok ? 8 : 0);
      }
    }
    iss_k0 += GEMM_BK;
    if (iss_k0 >= p.Ki) {
      iss_k0 = 0;
      ++iss_ko;
    }
  };

#pragma unroll
  for (int s = 0; s < STAGES - 1; ++s) {
    if (s < T) issue(s);
    cp_async_commit();
  }

  const int warp = tid >> 5, lane = tid & 31, g = lane >> 2, t4 = lane & 3;
  const int wm0 = (warp % WGM) * WTM, wn0 = (warp / WGM) * WTN;

  double acc[MI][NI][2];
#pragma unroll
  for (int i = 0; i < MI; ++i)
#pragma unroll
    for (int j = 0; j < NI; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

  for (int it = 0; it < T; ++it) {
    cp_async_wait<STAGES - 2>();
    __syncthreads();
    {
      const int nxt = it + STAGES - 1;
      if (nxt < T) issue(nxt % STAGES);
      cp_async_commit();
    }
    const double* sA = smem + (it % STAGES) * SM::STAGE;
    const double* sB = sA + SM::SA;
    const double* sW = sB + SM::SB;
#pragma unroll
    for (int kk = 0; kk < GEMM_BK / 4; ++kk) {
      const int k = kk * 4 + t4;
      double a[MI], b[NI];
#pragma unroll
      for (int i = 0; i < MI; ++i) {
        const int r = wm0 + i * 8 + g;
        a[i] = AK ? sA[r * GEMM_LDK + k] : sA[k * (BM + 4) + r];
      }
      if (HASW) {
        const double wv = sW[k];
#pragma unroll
        for (int i = 0; i < MI; ++i) a[i] *= wv;
      }
#pragma unroll
      for (int j = 0; j < NI; ++j) {
        const int r = wn0 + j * 8 + g;
        b[j] = BKM ? sB[r * GEMM_LDK + k] : sB[k * (BN + 4) + r];
      }
#pragma unroll
      for (int i = 0; i < MI; ++i)
#pragma unroll
        for (int j = 0; j < NI; ++j) dmma884(acc[i][j], a[i], b[j]);
    }
  }
  cp_async_wait<0>();

  // ------------------------------ epilogue ------------------------------
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
    return;
  }

  double* Cz = p.C + z1 * p.sC_z1 + z2 * p.sC_z2;
  long long coff[NI][2];
  double csc[NI][2];
#pragma unroll
  for (int j = 0; j < NI; ++j)
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int col = n_base + wn0 + j * 8 + 2 * t4 + e;
      if (col < p.N) {
        coff[j][e] = (long long)(col / p.Ln) * p.sC_no + (long long)(col % p.Ln) * p.sC_ni;
        csc[j][e] = p.alpha * (p.nscale ? p.nscale[p.nscale_mod ? col % p.nscale_mod : col] : 1.0);
      } else {
        coff[j][e] = -1;
        csc[j][e] = 0.0;
      }
    }
#pragma unroll
  for (int i = 0; i < MI; ++i) {
    const int row = m_base + wm0 + i * 8 + g;
    if (row >= p.M) continue;
    const long long roff = (long long)(row / p.Lm) * p.sC_mo + (long long)(row % p.Lm) * p.sC_mi;
#pragma unroll
    for (int j = 0; j < NI; ++j)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        if (coff[j][e] < 0) continue;
        double* dst = Cz + roff + coff[j][e];
        double v = csc[j][e] * acc[i][j][e];
        if (p.beta != 0.0) v += p.beta * (*dst);
        *dst = v;
      }
  }
}

// Split-K second pass: sums the partials and applies the epilogue.
template <int BM, int BN>
__global__ void gemm_splitk_reduce_kernel(const GemmParams p) {
  const long long Mpad = (long long)p.tiles_m * BM, Npad = (long long)p.tiles_n * BN;
  const int z = blockIdx.z;
  const int z1 = z % p.Z1, z2 = z / p.Z1;
  const int row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= p.M) return;
  // columns beyond the grid's y limit (65535) are taken in further rounds
  for (int col = blockIdx.y; col < p.N; col += gridDim.y) {
    if (p.lower_only && (col / BN) * BN > (row / BM) * BM + BM - 1) continue;
    const double* wsp = p.ws + (long long)z * p.splitk * Mpad * Npad + (long long)col * Mpad + row;
    double s = 0.0;
    for (int k = 0; k < p.splitk; ++k) s += wsp[(long long)k * Mpad * Npad];
    double* dst = p.C + z1 * p.sC_z1 + z2 * p.sC_z2 + (long long)(row / p.Lm) * p.sC_mo +
                  (long long)(row % p.Lm) * p.sC_mi + (long long)(col / p.Ln) * p.sC_no +
                  (long long)(col % p.Ln) * p.sC_ni;
    double v = p.alpha * (p.nscale ? p.nscale[p.nscale_mod ? col % p.nscale_mod : col] : 1.0) * s;
    if (p.beta != 0.0) v += p.beta * (*dst);
    *dst = v;
  }
}

// Host-side launcher (gemm_dmma.cu).  Chooses tile shape, split-K and vector
// width; `ws`/`ws_bytes` is a caller-owned scratch buffer for split-K.
struct GemmPlan {
  int cfg = 0;     // 0: 128x128, 1: 128x32 (skinny N), 2: 64x64
  int splitk = 1;  // 0 = auto
};
void gemm_launch(GemmParams p, cudaStream_t stream, double* ws, size_t ws_bytes, int num_sms,
                 int force_cfg = -1, int force_splitk = 0);
size_t gemm_ws_bytes_needed(const GemmParams& p, int num_sms, int force_cfg = -1, int force_splitk = 0);
// tile configuration / orientation / split the launcher would pick (profiling reports)
void gemm_plan_describe(const GemmParams& p, int num_sms, int force_cfg, int force_splitk, int* cfg, int* swap,
                        int* splitk);
void gemm_init_attributes();

}  // namespace gwbse
