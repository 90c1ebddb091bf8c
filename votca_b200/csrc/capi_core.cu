// C ABI, part 1: context, device memory, dense primitives, dense auxiliaries.
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <vector>

#include "../../include/gwbse_b200.h"
#include "context.cuh"
#include "gemm_tma.cuh"

using namespace gwbse;

static thread_local std::string g_create_error;

void gwbse_ctx::gemm(const GemmParams& p, int cfg, int splitk, double algo_flops) {
  const size_t need = gemm_ws_bytes_needed(p, num_sms, cfg, splitk);
  double* ws = need ? buf("gemm_ws", need / sizeof(double)) : nullptr;
  if (gemm_profile) {
    if (gemm_events_used == gemm_events.size()) {
      if (gemm_events.size() >= 4096) gemm_collect();
      if (gemm_events_used == gemm_events.size()) {
        cudaEvent_t a, b;
        GW_CUDA(cudaEventCreate(&a));
        GW_CUDA(cudaEventCreate(&b));
        gemm_events.emplace_back(a, b);
      }
    }
    GW_CUDA(cudaEventRecord(gemm_events[gemm_events_used].first, stream));
  }
  gemm_launch(p, stream, ws, need, num_sms, cfg, splitk);
  if (gemm_profile) {
    GW_CUDA(cudaEventRecord(gemm_events[gemm_events_used].second, stream));
    ++gemm_events_used;
    if (algo_flops < 0.0) {
      algo_flops = 2.0 * p.M * (double)p.N * (double)p.Ko * (double)p.Ki * p.Z1 * p.Z2;
      if (p.lower_only) algo_flops *= 0.5 * (1.0 + 1.0 / std::max(p.N, 1));
    }
    gemm_flops += algo_flops;
    ++gemm_launches;
    int pc = 0, psw = 0, psk = 1;
    gemm_plan_describe(p, num_sms, cfg, splitk, &pc, &psw, &psk);
    char key[160];
    snprintf(key, sizeof(key), "M=%-7d N=%-7d K=%-8lld Z=%-4d %s%s%s%s cfg%d%s sk%d", p.M, p.N,
             (long long)p.Ko * p.Ki, p.Z1 * p.Z2, p.A.s_ki == 1 ? "Ak" : "Am", p.B.s_ki == 1 ? "Bk" : "Bm",
             p.w ? " w" : "", p.lower_only ? " syrk" : (p.nscale ? " nsc" : ""), pc, psw ? "s" : "", psk);
    gemm_event_keys.emplace_back(key);
    gemm_event_flops.push_back(algo_flops);
  }
  launches += (need ? 2 : 1);
}

void gwbse_ctx::gemm_collect() {
  if (gemm_events_used == 0) return;
  GW_CUDA(cudaStreamSynchronize(stream));
  for (size_t i = 0; i < gemm_events_used; ++i) {
    float ms = 0.f;
    GW_CUDA(cudaEventElapsedTime(&ms, gemm_events[i].first, gemm_events[i].second));
    gemm_ms += ms;
    if (i < gemm_event_keys.size()) {
      auto& st = gemm_shapes[gemm_event_keys[i]];
      st.ms += ms;
      st.flops += gemm_event_flops[i];
      st.calls++;
    }
  }
  gemm_event_keys.clear();
  gemm_event_flops.clear();
  gemm_events_used = 0;
}

namespace {
__global__ void __launch_bounds__(256) dmma_probe_kernel(double* out, int iters, double seed) {
  double a = seed + threadIdx.x * 1e-9, b = seed * 0.5;
  double c[8][2];
#pragma unroll
  for (int j = 0; j < 8; ++j) c[j][0] = c[j][1] = 0.0;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int j = 0; j < 8; ++j) gwbse::dmma884(c[j], a, b);
  }
  double s = 0.0;
#pragma unroll
  for (int j = 0; j < 8; ++j) s += c[j][0] + c[j][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
}  // namespace

static void check_solver(cusolverStatus_t st, const char* what) {
  if (st != CUSOLVER_STATUS_SUCCESS)
    throw std::runtime_error(std::string("cuSOLVER error ") + std::to_string((int)st) + " in " + what);
}

extern "C" {

int gwbse_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
  return n;
}

const char* gwbse_create_error(void) { return g_create_error.c_str(); }

int gwbse_ctx_create(int device, gwbse_ctx** out) {
  if (!out) return 1;
  *out = nullptr;
  gwbse_ctx* ctx = nullptr;
  try {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
      throw std::runtime_error(std::string("gwbse_b200 needs a CUDA device (sm_100a); none found: ") +
                               cudaGetErrorString(e));
    if (device < 0 || device >= n) throw std::runtime_error("gwbse_ctx_create: invalid device index");
    GW_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    GW_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10)
      throw std::runtime_error(std::string("gwbse_b200 is built for sm_100a only; device is ") + prop.name);
    ctx = new gwbse_ctx();
    ctx->device = device;
    ctx->num_sms = prop.multiProcessorCount;
    GW_CUDA(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    GW_CUDA(cudaEventCreate(&ctx->ev0));
    GW_CUDA(cudaEventCreate(&ctx->ev1));
    check_solver(cusolverDnCreate(&ctx->solver), "cusolverDnCreate");
    check_solver(cusolverDnSetStream(ctx->solver, ctx->stream), "cusolverDnSetStream");
    *out = ctx;
    return 0;
  } catch (const std::exception& e) {
    g_create_error = e.what();
    delete ctx;
    return 1;
  }
}

void gwbse_ctx_destroy(gwbse_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  for (auto& kv : ctx->bufs)
    if (kv.second.p) cudaFree(kv.second.p);
  for (auto& kv : ctx->free_blocks) cudaFree(kv.second);
  for (auto& kv : ctx->live_blocks) cudaFree(kv.first);
  for (double* p : {ctx->X, ctx->X2, ctx->Xsnap, ctx->mos, ctx->exact_res})
    if (p) cudaFree(p);
  for (auto* st : {&ctx->sig_ppm, &ctx->sig_exact})
    for (double* p : {st->fac, st->pole, st->energies})
      if (p) cudaFree(p);
  sigma_tree_destroy(ctx->sig_ppm.tree);
  sigma_tree_destroy(ctx->sig_exact.tree);
  if (ctx->solver) cusolverDnDestroy(ctx->solver);
  for (auto& e : ctx->gemm_events) {
    cudaEventDestroy(e.first);
    cudaEventDestroy(e.second);
  }
  if (ctx->ev0) cudaEventDestroy(ctx->ev0);
  if (ctx->ev1) cudaEventDestroy(ctx->ev1);
  for (int i = 0; i < 2; ++i) {
    if (ctx->fill_copied[i]) cudaEventDestroy(ctx->fill_copied[i]);
    if (ctx->fill_consumed[i]) cudaEventDestroy(ctx->fill_consumed[i]);
  }
  if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
  if (ctx->stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
}

const char* gwbse_last_error(const gwbse_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

int gwbse_sync(gwbse_ctx* ctx) {
  GW_API_BEGIN(ctx)
  GW_CUDA(cudaStreamSynchronize(ctx->stream));
  GW_API_END(ctx)
}

long long gwbse_launch_count(const gwbse_ctx* ctx) { return ctx ? ctx->launches : 0; }

int gwbse_set_option(gwbse_ctx* ctx, const char* key, double value) {
  GW_API_BEGIN(ctx)
  const std::string k(key ? key : "");
  if (k == "profile") {
    ctx->collect_regions();
    ctx->regions.clear();
    ctx->profile = value != 0.0;
  } else if (k == "tma") {
    // process-wide switch between the TMA-staged GEMM and the cp.async kernel for operands both can take (A/B timing)
    gwbse::gemm_tma_set_enabled(value != 0.0);
  } else if (k == "bse_chunk_bytes") {
    GW_REQUIRE(value >= 1024, "bse_chunk_bytes too small");
    ctx->bse_chunk_bytes = (size_t)value;
  } else if (k == "bse_dense") {
    // 0: the direct terms of the BSE operator stay factorised; 1: their B x B blocks are materialised once that
    // pays back (default); 2: always, memory permitting
    GW_REQUIRE(value == 0.0 || value == 1.0 || value == 2.0, "bse_dense is 0, 1 or 2");
    ctx->bse_dense_mode = (int)value;
  } else if (k == "bse_dense_payback") {
    GW_REQUIRE(value >= 0.0, "bse_dense_payback must not be negative");
    ctx->bse_dense_payback = value;
  } else if (k == "sigma_tree_min_terms") {
    ctx->sigma_tree_min_terms = (long long)value;
  } else if (k == "sigma_tree_bytes") {
    GW_REQUIRE(value >= 1 << 20, "sigma_tree_bytes too small");
    ctx->sigma_tree_bytes = (size_t)value;
    sigma_tree_invalidate(ctx->sig_ppm.tree);
    sigma_tree_invalidate(ctx->sig_exact.tree);
  } else {
    throw std::runtime_error("unknown option '" + k + "'");
  }
  GW_API_END(ctx)
}

int gwbse_profile_report(gwbse_ctx* ctx, char* buf, size_t buflen) {
  GW_API_BEGIN(ctx)
  ctx->collect_regions();
  std::string out = "region                          calls   device_ms     host_ms\n";
  std::vector<std::pair<double, std::string>> rows;
  for (auto& kv : ctx->regions) {
    char line[160];
    std::snprintf(line, sizeof(line), "%-30s %6lld %11.2f %11.2f\n", kv.first.c_str(), kv.second.calls, kv.second.ms,
                  kv.second.wall_ms);
    rows.emplace_back(-kv.second.ms, line);
  }
  std::sort(rows.begin(), rows.end());
  for (auto& r : rows) out += r.second;
  ctx->regions.clear();
  if (buf && buflen) {
    std::strncpy(buf, out.c_str(), buflen - 1);
    buf[buflen - 1] = 0;
  }
  GW_API_END(ctx)
}

int gwbse_gemm_profile(gwbse_ctx* ctx, int enable) {
  GW_API_BEGIN(ctx)
  ctx->gemm_collect();
  ctx->gemm_profile = enable != 0;
  ctx->gemm_ms = 0.0;
  ctx->gemm_flops = 0.0;
  ctx->gemm_launches = 0;
  ctx->gemm_shapes.clear();
  GW_API_END(ctx)
}

int gwbse_gemm_shape_report(gwbse_ctx* ctx, char* buf, size_t buflen) {
  GW_API_BEGIN(ctx)
  ctx->gemm_collect();
  std::vector<std::pair<double, std::string>> rows;
  for (auto& kv : ctx->gemm_shapes) {
    char line[256];
    snprintf(line, sizeof(line), "%-72s %6lld %10.2f %8.2f\n", kv.first.c_str(), kv.second.calls, kv.second.ms,
             kv.second.ms > 0 ? kv.second.flops / kv.second.ms * 1e-9 : 0.0);
    rows.emplace_back(-kv.second.ms, line);
  }
  std::sort(rows.begin(), rows.end());
  std::string out = "shape (A/B k- or m-major, flags, tile cfg, split-K)                               calls         ms   TFLOP/s\n";
  for (auto& r : rows) out += r.second;
  if (buf && buflen) {
    const size_t n = std::min(buflen - 1, out.size());
    std::memcpy(buf, out.data(), n);
    buf[n] = 0;
  }
  GW_API_END(ctx)
}

int gwbse_gemm_stats(gwbse_ctx* ctx, double* ms, double* flops, long long* launches) {
  GW_API_BEGIN(ctx)
  ctx->gemm_collect();
  if (ms) *ms = ctx->gemm_ms;
  if (flops) *flops = ctx->gemm_flops;
  if (launches) *launches = ctx->gemm_launches;
  GW_API_END(ctx)
}

int gwbse_fp64_peak_probe(gwbse_ctx* ctx, double* tflops) {
  GW_API_BEGIN(ctx)
  const int grid = ctx->num_sms * 4, iters = 20000;
  double* out = ctx->buf("probe_out", (size_t)grid * 256);
  dmma_probe_kernel<<<grid, 256, 0, ctx->stream>>>(out, 1000, 1.0);  // warm-up
  float best = 1e30f;
  for (int r = 0; r < 3; ++r) {
    GW_CUDA(cudaEventRecord(ctx->ev0, ctx->stream));
    dmma_probe_kernel<<<grid, 256, 0, ctx->stream>>>(out, iters, 1.0);
    GW_CUDA(cudaEventRecord(ctx->ev1, ctx->stream));
    GW_CUDA(cudaEventSynchronize(ctx->ev1));
    float ms = 0.f;
    GW_CUDA(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
    best = std::min(best, ms);
  }
  GW_CUDA(cudaGetLastError());
  // warps * iters * 8 mma * (8*8*4*2 flop)
  *tflops = (double)grid * 8 * (double)iters * 8 * 512.0 / (best * 1e-3) / 1e12;
  GW_API_END(ctx)
}

int gwbse_timer_start(gwbse_ctx* ctx) {
  GW_API_BEGIN(ctx)
  GW_CUDA(cudaEventRecord(ctx->ev0, ctx->stream));
  GW_API_END(ctx)
}
int gwbse_timer_stop_ms(gwbse_ctx* ctx, float* ms) {
  GW_API_BEGIN(ctx)
  GW_CUDA(cudaEventRecord(ctx->ev1, ctx->stream));
  GW_CUDA(cudaEventSynchronize(ctx->ev1));
  GW_CUDA(cudaEventElapsedTime(ms, ctx->ev0, ctx->ev1));
  GW_API_END(ctx)
}

// ------------------------------ memory ------------------------------------
int gwbse_dev_malloc(gwbse_ctx* ctx, size_t bytes, double** out_dev) {
  GW_API_BEGIN(ctx)
  if (bytes == 0) bytes = 8;
  auto it = ctx->free_blocks.find(bytes);
  if (it != ctx->free_blocks.end()) {
    *out_dev = it->second;
    ctx->free_blocks.erase(it);
    ctx->cached_bytes -= bytes;
  } else {
    // same pre-check and message style as CudaMatrix::alloc (cudamatrix.cc:97-113)
    size_t fr = 0, tot = 0;
    GW_CUDA(cudaMemGetInfo(&fr, &tot));
    if (bytes > fr && ctx->cached_bytes > 0) {  // give cached blocks back before failing
      GW_CUDA(cudaStreamSynchronize(ctx->stream));
      for (auto& kv : ctx->free_blocks) cudaFree(kv.second);
      ctx->free_blocks.clear();
      ctx->cached_bytes = 0;
      GW_CUDA(cudaMemGetInfo(&fr, &tot));
    }
    if (bytes > fr)
      throw std::runtime_error("There were requested : " + std::to_string((double)bytes / 1048576.0) +
                               " MB but the device has " + std::to_string((double)fr / 1048576.0) + " MB free.");
    GW_CUDA(cudaMalloc(out_dev, bytes));
  }
  ctx->live_blocks[*out_dev] = bytes;
  GW_API_END(ctx)
}
int gwbse_dev_free(gwbse_ctx* ctx, double* p) {
  GW_API_BEGIN(ctx)
  if (p) {
    auto it = ctx->live_blocks.find(p);
    if (it == ctx->live_blocks.end()) throw std::runtime_error("gwbse_dev_free: unknown device pointer");
    const size_t bytes = it->second;
    ctx->live_blocks.erase(it);
    // work queued on the context's stream may still use the block; blocks are only re-used by later work on
    // the same stream, so no synchronisation is needed to cache it
    if (ctx->cached_bytes + bytes <= ((size_t)8 << 30)) {
      ctx->free_blocks.emplace(bytes, p);
      ctx->cached_bytes += bytes;
    } else {
      GW_CUDA(cudaStreamSynchronize(ctx->stream));
      GW_CUDA(cudaFree(p));
    }
  }
  GW_API_END(ctx)
}
int gwbse_h2d(gwbse_ctx* ctx, double* dst, const double* src, size_t n) {
  GW_API_BEGIN(ctx)
  GW_CUDA(cudaMemcpyAsync(dst, src, n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  GW_CUDA(cudaStreamSynchronize(ctx->stream));
  GW_API_END(ctx)
}
int gwbse_d2h(gwbse_ctx* ctx, double* dst, const double* src, size_t n) {
  GW_API_BEGIN(ctx)
  GW_CUDA(cudaMemcpyAsync(dst, src, n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  GW_CUDA(cudaStreamSynchronize(ctx->stream));
  GW_API_END(ctx)
}
int gwbse_d2d(gwbse_ctx* ctx, double* dst, const double* src, size_t n) {
  GW_API_BEGIN(ctx)
  GW_CUDA(cudaMemcpyAsync(dst, src, n * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
  GW_API_END(ctx)
}
int gwbse_dev_memset_zero(gwbse_ctx* ctx, double* dst, size_t n) {
  GW_API_BEGIN(ctx)
  GW_CUDA(cudaMemsetAsync(dst, 0, n * sizeof(double), ctx->stream));
  GW_API_END(ctx)
}
int gwbse_dev_mem_info(gwbse_ctx* ctx, size_t* fr, size_t* tot) {
  GW_API_BEGIN(ctx)
  GW_CUDA(cudaMemGetInfo(fr, tot));
  GW_API_END(ctx)
}

// ------------------------------ GEMM --------------------------------------
static GemmParams blas_params(char ta, char tb, int m, int n, int k, double alpha, const double* A, int lda,
                              const double* B, int ldb, double beta, double* C, int ldc) {
  const bool at = (ta == 'T' || ta == 't'), bt = (tb == 'T' || tb == 't');
  GW_REQUIRE(at || ta == 'N' || ta == 'n', "transa must be N or T");
  GW_REQUIRE(bt || tb == 'N' || tb == 'n', "transb must be N or T");
  // shape checks in the spirit of CudaPipeline::gemm (cudapipeline.h:113-118)
  GW_REQUIRE(lda >= (at ? k : m) && ldb >= (bt ? n : k) && ldc >= m, "Shape mismatch in cuda gemm");
  GemmParams p;
  p.M = m;
  p.N = n;
  p.Ko = 1;
  p.Ki = k;
  p.A.ptr = A;
  if (at) {  // op(A)(i,l) = A(l,i): k contiguous
    p.A.s_ri = lda;
    p.A.s_ki = 1;
  } else {
    p.A.s_ri = 1;
    p.A.s_ki = lda;
  }
  p.B.ptr = B;
  if (bt) {  // op(B)(l,j) = B(j,l): row index j contiguous
    p.B.s_ri = 1;
    p.B.s_ki = ldb;
  } else {
    p.B.s_ri = ldb;
    p.B.s_ki = 1;
  }
  p.C = C;
  p.sC_mi = 1;
  p.sC_ni = ldc;
  p.alpha = alpha;
  p.beta = beta;
  return p;
}

int gwbse_dgemm_dev_ex(gwbse_ctx* ctx, char ta, char tb, int m, int n, int k, double alpha, const double* A,
                       int lda, const double* B, int ldb, double beta, double* C, int ldc, int cfg, int splitk) {
  GW_API_BEGIN(ctx)
  GW_PROF(ctx, "dgemm_dev");
  if (m > 0 && n > 0) {
    if (k <= 0) {
      // C = beta * C
      if (beta == 0.0) {
        GW_CUDA(cudaMemset2DAsync(C, sizeof(double) * ldc, 0, sizeof(double) * m, n, ctx->stream));
      } else if (beta != 1.0) {
        std::vector<double> s(n, beta);
        double* sd = ctx->buf("tmp_scale", n);
        GW_CUDA(cudaMemcpyAsync(sd, s.data(), sizeof(double) * n, cudaMemcpyHostToDevice, ctx->stream));
        GW_CUDA(cudaStreamSynchronize(ctx->stream));
        launch_scale_cols(m, n, C, ldc, sd, ctx->stream);
      }
    } else {
      ctx->gemm(blas_params(ta, tb, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc), cfg, splitk);
    }
  }
  GW_API_END(ctx)
}

int gwbse_dgemm_dev(gwbse_ctx* ctx, char ta, char tb, int m, int n, int k, double alpha, const double* A, int lda,
                    const double* B, int ldb, double beta, double* C, int ldc) {
  return gwbse_dgemm_dev_ex(ctx, ta, tb, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, -1, 0);
}

int gwbse_diag_scale_dev(gwbse_ctx* ctx, char side, int m, int n, const double* A, int lda, const double* d,
                         double* C, int ldc) {
  GW_API_BEGIN(ctx)
  launch_diag_scale(side, m, n, A, lda, d, C, ldc, ctx->stream);
  ctx->launches++;
  GW_API_END(ctx)
}

int gwbse_axpy_dev(gwbse_ctx* ctx, int m, int n, double alpha, const double* X, int ldx, double* Y, int ldy) {
  GW_API_BEGIN(ctx)
  launch_axpy(m, n, alpha, X, ldx, Y, ldy, ctx->stream);
  ctx->launches++;
  GW_API_END(ctx)
}

int gwbse_colnorms_dev(gwbse_ctx* ctx, int m, int n, const double* A, int lda, double* norms) {
  GW_API_BEGIN(ctx)
  if (n > 0) {
    double* d = ctx->buf("colred", n);
    launch_colnorms(m, n, A, lda, d, ctx->stream);
    ctx->launches++;
    GW_CUDA(cudaMemcpyAsync(norms, d, sizeof(double) * n, cudaMemcpyDeviceToHost, ctx->stream));
    GW_CUDA(cudaStreamSynchronize(ctx->stream));
  }
  GW_API_END(ctx)
}

int gwbse_coldots_dev(gwbse_ctx* ctx, int m, int n, const double* X, int ldx, const double* Y, int ldy,
                      double* dots) {
  GW_API_BEGIN(ctx)
  if (n > 0) {
    double* d = ctx->buf("colred", n);
    launch_coldots(m, n, X, ldx, Y, ldy, d, ctx->stream);
    ctx->launches++;
    GW_CUDA(cudaMemcpyAsync(dots, d, sizeof(double) * n, cudaMemcpyDeviceToHost, ctx->stream));
    GW_CUDA(cudaStreamSynchronize(ctx->stream));
  }
  GW_API_END(ctx)
}

int gwbse_scale_cols_dev(gwbse_ctx* ctx, int m, int n, double* A, int lda, const double* s) {
  GW_API_BEGIN(ctx)
  if (n > 0) {
    double* d = ctx->buf("colscale", n);
    GW_CUDA(cudaMemcpyAsync(d, s, sizeof(double) * n, cudaMemcpyHostToDevice, ctx->stream));
    launch_scale_cols(m, n, A, lda, d, ctx->stream);
    ctx->launches++;
    GW_CUDA(cudaStreamSynchronize(ctx->stream));
  }
  GW_API_END(ctx)
}

// --------------------------- dense auxiliaries -----------------------------
int gwbse_sym_eig_dev(gwbse_ctx* ctx, int n, double* A, int lda, double* w) {
  GW_API_BEGIN(ctx)
  GW_PROF(ctx, "sym_eig");
  if (n > 0) {
    int lwork = 0;
    double* wd = ctx->buf("eig_w", n);
    check_solver(cusolverDnDsyevd_bufferSize(ctx->solver, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, n, A,
                                             lda, wd, &lwork),
                 "Dsyevd_bufferSize");
    double* work = ctx->buf("solver_work", (size_t)lwork + 8);
    int* info = reinterpret_cast<int*>(ctx->buf("solver_info", 8));
    check_solver(cusolverDnDsyevd(ctx->solver, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, n, A, lda, wd,
                                  work, lwork, info),
                 "Dsyevd");
    int hinfo = 0;
    GW_CUDA(cudaMemcpyAsync(&hinfo, info, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    GW_CUDA(cudaMemcpyAsync(w, wd, sizeof(double) * n, cudaMemcpyDeviceToHost, ctx->stream));
    GW_CUDA(cudaStreamSynchronize(ctx->stream));
    if (hinfo != 0) throw std::runtime_error("Small hermitian eigenvalue problem failed.");
  }
  GW_API_END(ctx)
}

static void lu_factor(gwbse_ctx* ctx, int n, double* A, int lda, int** ipiv_out) {
  int lwork = 0;
  check_solver(cusolverDnDgetrf_bufferSize(ctx->solver, n, n, A, lda, &lwork), "Dgetrf_bufferSize");
  double* work = ctx->buf("solver_work", (size_t)lwork + 8);
  int* info = reinterpret_cast<int*>(ctx->buf("solver_info", 8));
  int* ipiv = reinterpret_cast<int*>(ctx->buf("solver_ipiv", (size_t)n / 2 + 8));
  check_solver(cusolverDnDgetrf(ctx->solver, n, n, A, lda, work, ipiv, info), "Dgetrf");
  int hinfo = 0;
  GW_CUDA(cudaMemcpyAsync(&hinfo, info, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  GW_CUDA(cudaStreamSynchronize(ctx->stream));
  if (hinfo != 0) throw std::runtime_error("LU factorisation failed (singular matrix), info=" + std::to_string(hinfo));
  *ipiv_out = ipiv;
}

int gwbse_lu_solve_dev(gwbse_ctx* ctx, int n, int nrhs, double* A, int lda, double* B, int ldb) {
  GW_API_BEGIN(ctx)
  if (n > 0 && nrhs > 0) {
    int* ipiv = nullptr;
    lu_factor(ctx, n, A, lda, &ipiv);
    int* info = reinterpret_cast<int*>(ctx->buf("solver_info", 8));
    check_solver(cusolverDnDgetrs(ctx->solver, CUBLAS_OP_N, n, nrhs, A, lda, ipiv, B, ldb, info), "Dgetrs");
  }
  GW_API_END(ctx)
}

__global__ void set_identity_kernel(double* A, int n, long long ld) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int j = blockIdx.y;
  if (i < n && j < n) A[i + j * ld] = (i == j) ? 1.0 : 0.0;
}

int gwbse_inverse_dev(gwbse_ctx* ctx, int n, double* A, int lda) {
  GW_API_BEGIN(ctx)
  GW_PROF(ctx, "inverse");
  if (n > 0) {
    int* ipiv = nullptr;
    lu_factor(ctx, n, A, lda, &ipiv);
    double* I = ctx->buf("inverse_rhs", (size_t)n * n);
    set_identity_kernel<<<dim3((n + 255) / 256, n), 256, 0, ctx->stream>>>(I, n, n);
    GW_CUDA(cudaGetLastError());
    int* info = reinterpret_cast<int*>(ctx->buf("solver_info", 8));
    check_solver(cusolverDnDgetrs(ctx->solver, CUBLAS_OP_N, n, n, A, lda, ipiv, I, n, info), "Dgetrs");
    GW_CUDA(copy2d_async(A, sizeof(double) * lda, I, sizeof(double) * n, sizeof(double) * n, n,
                              cudaMemcpyDeviceToDevice, ctx->stream));
  }
  GW_API_END(ctx)
}

// Generalized non-symmetric eigenproblem T x = lambda B x for the harmonic Ritz step.
// Solved as (B^-1 T) x = lambda x with cusolverDnXgeev (hybrid host/device LAPACK-class routine).
int gwbse_gen_eig_host(gwbse_ctx* ctx, int n, const double* T, const double* B, double* wr, double* wi,
                       double* VR) {
  GW_API_BEGIN(ctx)
  GW_PROF(ctx, "gen_eig_host");
  if (n > 0) {
    const size_t nn = (size_t)n * n;
    double* dT = ctx->buf("geig_T", nn);
    double* dB = ctx->buf("geig_B", nn);
    double* dW = ctx->buf("geig_W", 2 * (size_t)n);  // complex eigenvalues
    double* dVR = ctx->buf("geig_VR", nn);
    GW_CUDA(cudaMemcpyAsync(dT, T, sizeof(double) * nn, cudaMemcpyHostToDevice, ctx->stream));
    GW_CUDA(cudaMemcpyAsync(dB, B, sizeof(double) * nn, cudaMemcpyHostToDevice, ctx->stream));
    // dT <- B^-1 T
    int* ipiv = nullptr;
    lu_factor(ctx, n, dB, n, &ipiv);
    int* info = reinterpret_cast<int*>(ctx->buf("solver_info", 8));
    check_solver(cusolverDnDgetrs(ctx->solver, CUBLAS_OP_N, n, n, dB, n, ipiv, dT, n, info), "Dgetrs");
    cusolverDnParams_t params;
    check_solver(cusolverDnCreateParams(&params), "CreateParams");
    size_t wdev = 0, whost = 0;
    check_solver(cusolverDnXgeev_bufferSize(ctx->solver, params, CUSOLVER_EIG_MODE_NOVECTOR, CUSOLVER_EIG_MODE_VECTOR, n, CUDA_R_64F, dT,
                       n, CUDA_C_64F, dW, CUDA_R_64F, nullptr, n, CUDA_R_64F, dVR, n, CUDA_R_64F, &wdev, &whost),
                 "Xgeev_bufferSize");
    double* dwork = ctx->buf("solver_work", wdev / sizeof(double) + 8);
    std::vector<unsigned char> hwork(whost + 8);
    check_solver(cusolverDnXgeev(ctx->solver, params, CUSOLVER_EIG_MODE_NOVECTOR, CUSOLVER_EIG_MODE_VECTOR, n, CUDA_R_64F, dT,
                        n, CUDA_C_64F, dW, CUDA_R_64F, nullptr, n, CUDA_R_64F, dVR, n, CUDA_R_64F, dwork, wdev,
                        hwork.data(), whost, info),
                 "Xgeev");
    cusolverDnDestroyParams(params);
    int hinfo = 0;
    std::vector<double> W(2 * (size_t)n);
    GW_CUDA(cudaMemcpyAsync(&hinfo, info, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    GW_CUDA(cudaMemcpyAsync(W.data(), dW, sizeof(double) * 2 * n, cudaMemcpyDeviceToHost, ctx->stream));
    GW_CUDA(cudaMemcpyAsync(VR, dVR, sizeof(double) * nn, cudaMemcpyDeviceToHost, ctx->stream));
    GW_CUDA(cudaStreamSynchronize(ctx->stream));
    if (hinfo != 0) throw std::runtime_error("Small generalized eigenvalue problem failed.");
    for (int i = 0; i < n; ++i) {
      wr[i] = W[2 * i];
      wi[i] = W[2 * i + 1];
    }
  }
  GW_API_END(ctx)
}

}  // extern "C"
