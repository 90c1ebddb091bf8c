// C ABI, part 6: AO Coulomb integrals on the device.
//   gwbse_ao3c_block(_dev)   replaces ComputeAO3cBlock          xtp/src/libxtp/libint2_calls.cc:544-593
//   gwbse_ao_coulomb2c       replaces AOCoulomb::Fill           xtp/src/libxtp/libint2_calls.cc:224-271
//   gwbse_mmn_fill_from_basis = TCMatrix_gwbse::Fill3cMO with the producer on the GPU (libint2_calls.cc:595-651):
//                             the (P|mu nu) blocks are written straight into the buffer the contraction GEMMs
//                             read; no AO integral crosses PCIe.
// One warp per (orbital shell pair, auxiliary shell); launches are grouped by angular-momentum class (la >= lb, lc)
// so that each class gets the shared-memory scratch - and therefore the occupancy - of its own size.  The
// arithmetic is ao3c_core.cuh (shared with the CPU harness).
#include <algorithm>
#include <cstdlib>
#include <map>
#include <memory>
#include <vector>

#include "../../include/gwbse_b200.h"
#include "ao3c_core.cuh"
#include "ao3c_tables.h"
#include "context.cuh"

using namespace gwbse;

namespace {

// barrier of one lane group: the whole warp, or an aligned sub-group of 4 / 16 lanes that works on its own shell
// triple (sub-groups may sit in different loop iterations; the mask names only the caller's group)
struct GroupBarrier {
  unsigned mask;
  __device__ __forceinline__ void operator()() const { __syncwarp(mask); }
};
struct GroupBarrierFactory {
  __device__ __forceinline__ GroupBarrier operator()(int sub, int group_lanes) const {
    return GroupBarrier{group_lanes == 32 ? 0xffffffffu : (((1u << group_lanes) - 1u) << (sub * group_lanes))};
  }
};

// group_lanes lanes per (shell pair, aux shell): 32 for the wide classes, fewer for classes whose widest stage has
// only a handful of entries ((ss|s) has one) so that a warp carries several triples instead of idling 31 lanes.
// The indexing lives in ao::cta_thread (shared with the CPU harness).
// MINB: resident CTAs per SM the register allocation is sized for.  The arithmetic is latency bound (dependent
// recursions through shared memory), so classes whose scratch leaves room run four CTAs of 64-register threads
// instead of two of 128 (a few spilled index registers against twice the warps in flight).
template <int MINB>
__global__ void __launch_bounds__(256, MINB)
    ao3c_kernel(ao::BasisView dft, ao::BasisView aux, ao::TableView tb, const ao::PairEntry* __restrict__ pairs,
                long long npairs, const double* __restrict__ pool, const int* __restrict__ aux_shells, int naux_shells,
                ao::OutSpec out, int ws_doubles, int group_lanes) {
  extern __shared__ double ao3c_smem[];
  GroupBarrierFactory sync_of;
  ao::cta_thread((long long)blockIdx.x, (int)(blockDim.x >> 5), (int)threadIdx.x, dft, aux, tb, pairs, npairs, pool,
                 aux_shells, naux_shells, out, ws_doubles, group_lanes, ao3c_smem, sync_of);
}

template <typename T>
T* upload(const std::vector<T>& v) {
  T* d = nullptr;
  GW_CUDA(cudaMalloc(&d, std::max<size_t>(1, v.size()) * sizeof(T)));
  if (!v.empty()) GW_CUDA(cudaMemcpy(d, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
  return d;
}

// Boys grid, (t,u,v) table and cartesian -> pure matrices: built once per device
struct DeviceTables {
  double* boys = nullptr;
  double* pure = nullptr;
  uint32_t* tuv = nullptr;
  ao::TableView view{};
};
std::map<int, DeviceTables>& table_store() {
  static std::map<int, DeviceTables> s;
  return s;
}
const ao::TableView& device_tables(int device) {
  auto& store = table_store();
  auto it = store.find(device);
  if (it != store.end()) return it->second.view;
  DeviceTables t;
  t.boys = upload(ao::make_boys_table());
  t.tuv = upload(ao::make_tuv_table());
  std::vector<double> pure;
  for (int l = 0; l <= ao::LMAX_SHELL; ++l) {
    t.view.pure_off[l] = (int)pure.size();
    const std::vector<double> T = ao::make_pure_matrix(l);
    pure.insert(pure.end(), T.begin(), T.end());
  }
  t.pure = upload(pure);
  t.view.boys = t.boys;
  t.view.tuv = t.tuv;
  t.view.pure = t.pure;
  t.view.boys_orders = ao::BOYS_ORDERS;
  t.view.boys_taylor = ao::BOYS_TAYLOR;
  t.view.herm1_stride = ao::HERM1_STRIDE;
  t.view.herm1_dim = ao::LMAX_SHELL + 1;
  t.view.boys_dx = ao::BOYS_DX;
  t.view.boys_xmax = ao::BOYS_XMAX;
  return store.emplace(device, t).first->second.view;
}

}  // namespace

// A basis set on the device, with the shell lists the launches need
struct gwbse_basis {
  int device = 0;
  ao::HostBasis host;
  ao::BasisView view{};
  std::vector<void*> owned;
  // shells of one angular momentum, ascending (a function range maps to a contiguous piece of each list)
  std::vector<int> by_l[ao::LMAX_SHELL + 1];
  int* by_l_dev[ao::LMAX_SHELL + 1] = {};
  int* no_aux_dev = nullptr;  // {-1, -2, -3, -4}: "aux shell" codes of the overlap and the three dipole launches
  // shell pairs (x, y) with l_x >= l_y, every unordered pair once, grouped by (l_x, l_y); their primitive-pair
  // records (p, P, c_a c_b, E coefficients) in one pool.  Pairs without a surviving primitive pair are absent.
  struct PairClass {
    int la, lb;
    long long count;
    ao::PairEntry* dev;
    const double* pool;
  };
  std::vector<PairClass> pair_classes;
  long long pairs_total = 0, pairs_kept = 0;  // before / after the shell-pair screening
  // (x, unit partner), grouped by l_x (two-centre integrals)
  std::vector<PairClass> unit_classes;

  ~gwbse_basis() {
    for (void* p : owned) cudaFree(p);
  }
  template <typename T>
  T* keep(T* p) {
    owned.push_back(p);
    return p;
  }
};

namespace {

void build_basis(gwbse_basis& b, int device) {
  b.device = device;
  const ao::HostBasis& h = b.host;
  b.view.nshell = h.nshell;
  b.view.nfunc = h.nfunc;
  b.view.l = b.keep(upload(h.l));
  b.view.np = b.keep(upload(h.np));
  b.view.prim0 = b.keep(upload(h.prim0));
  b.view.func0 = b.keep(upload(h.func0));
  b.view.center = b.keep(upload(h.center));
  b.view.exps = b.keep(upload(h.exps));
  b.view.coefs = b.keep(upload(h.coefs));
  b.view.herm1 = b.keep(upload(h.herm1));
  for (int s = 0; s < h.nshell; ++s) b.by_l[h.l[s]].push_back(s);
  for (int l = 0; l <= ao::LMAX_SHELL; ++l) b.by_l_dev[l] = b.keep(upload(b.by_l[l]));
  b.no_aux_dev = b.keep(upload(std::vector<int>{-1, -2, -3, -4}));
  auto classify = [&](const ao::PairLists& pl, std::vector<gwbse_basis::PairClass>& classes) {
    const double* pool = b.keep(upload(pl.pool));
    std::map<std::pair<int, int>, std::vector<ao::PairEntry>> groups;
    for (const ao::PairEntry& e : pl.entries) groups[{h.l[e.a], e.b < 0 ? 0 : h.l[e.b]}].push_back(e);
    // heaviest classes first: the long-running warps start early
    for (auto it = groups.rbegin(); it != groups.rend(); ++it)
      classes.push_back({it->first.first, it->first.second, (long long)it->second.size(), b.keep(upload(it->second)), pool});
  };
  const ao::PairLists regular = ao::make_pair_lists(h, false);
  b.pairs_total = regular.total;
  b.pairs_kept = (long long)regular.entries.size();
  classify(regular, b.pair_classes);
  classify(ao::make_pair_lists(h, true), b.unit_classes);
}

int shared_memory_limit(gwbse_ctx* ctx) {
  int smem_limit = 0;
  GW_CUDA(cudaDeviceGetAttribute(&smem_limit, cudaDevAttrMaxSharedMemoryPerBlockOptin, ctx->device));
  static int smem_set = -1;
  if (smem_set != smem_limit) {
    GW_CUDA(cudaFuncSetAttribute(ao3c_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_limit));
    GW_CUDA(cudaFuncSetAttribute(ao3c_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_limit));
    smem_set = smem_limit;
  }
  return smem_limit;
}

// one launch: every pair of the class against n_aux_shells aux shells of angular momentum lc
void launch_class(gwbse_ctx* ctx, const gwbse_basis& orb, const gwbse_basis::PairClass& pc, const gwbse_basis& aux,
                  const int* aux_shells_dev, int n_aux_shells, int lc, const ao::OutSpec& out, int smem_limit) {
  GW_REQUIRE(pc.la + pc.lb + lc <= ao::LMAX_TOTAL, "angular momentum class beyond the Boys table");
  const ao::LaunchConfig cfg = ao::launch_config(pc.la, pc.lb, lc, (size_t)smem_limit);
  GW_REQUIRE(cfg.fits, "integral class does not fit into shared memory");
  const long long groups = pc.count * (long long)n_aux_shells;
  const long long per_cta = (long long)cfg.warps_per_cta * cfg.groups_per_warp;
  const long long blocks = (groups + per_cta - 1) / per_cta;
  GW_REQUIRE(blocks < (1LL << 31), "too many shell triples for one launch; use smaller aux blocks");
  // region profiler: one line per angular-momentum class (set_option "profile")
  char cls[32];
  std::snprintf(cls, sizeof(cls), "ao3c (%d %d|%d)", pc.la, pc.lb, lc);
  GW_PROF(ctx, cls);
  // four CTAs per SM when threads and scratch allow it (1 KB per CTA is reserved by the driver)
  static const bool force2 = [] { const char* e = std::getenv("GWBSE_AO3C_MINB2"); return e && e[0] == '1'; }();
  const bool four = !force2 && (size_t)(cfg.smem_bytes + 1024) * 4 <= (size_t)smem_limit + 1024;
  if (four)
    ao3c_kernel<4><<<(unsigned)blocks, cfg.warps_per_cta * 32, cfg.smem_bytes, ctx->stream>>>(
        orb.view, aux.view, device_tables(ctx->device), pc.dev, pc.count, pc.pool, aux_shells_dev, n_aux_shells, out,
        cfg.ws_doubles, cfg.group_lanes);
  else
    ao3c_kernel<2><<<(unsigned)blocks, cfg.warps_per_cta * 32, cfg.smem_bytes, ctx->stream>>>(
        orb.view, aux.view, device_tables(ctx->device), pc.dev, pc.count, pc.pool, aux_shells_dev, n_aux_shells, out,
        cfg.ws_doubles, cfg.group_lanes);
  GW_CUDA(cudaGetLastError());
  ctx->launches++;
}

// all (pair class) x (aux l) launches for the aux shells overlapping functions [f0, f1)
void launch_classes(gwbse_ctx* ctx, const gwbse_basis& orb, const std::vector<gwbse_basis::PairClass>& classes,
                    const gwbse_basis& aux, int f0, int f1, const ao::OutSpec& out) {
  GW_REQUIRE(orb.device == ctx->device && aux.device == ctx->device, "basis belongs to another device");
  GW_REQUIRE(f0 >= 0 && f1 <= aux.host.nfunc && f0 <= f1, "aux function range out of bounds");
  if (f0 == f1) return;
  const ao::AuxShellRange r = ao::aux_shell_range(aux.host.func0, aux.by_l, f0, f1);
  const int smem_limit = shared_memory_limit(ctx);
  for (int lc = ao::LMAX_SHELL; lc >= 0; --lc) {
    if (r.last[lc] <= r.first[lc]) continue;
    for (const auto& pc : classes)
      launch_class(ctx, orb, pc, aux, aux.by_l_dev[lc] + r.first[lc], r.last[lc] - r.first[lc], lc, out, smem_limit);
  }
}

// pitch: doubles between consecutive columns of one N x N block (0 = N)
void ao3c_block_dev(gwbse_ctx* ctx, const gwbse_basis* aux, const gwbse_basis* dft, int aux_offset, int aux_count,
                    double* out_dev, long long pitch = 0) {
  GW_REQUIRE(aux && dft && out_dev, "null argument");
  const long long N = dft->host.nfunc;
  if (pitch == 0) pitch = N;
  // blocks of screened-out shell pairs are exact zeros; so is the padding of a pitched block
  if ((dft->pairs_kept < dft->pairs_total || pitch != N) && aux_count > 0)
    GW_CUDA(cudaMemsetAsync(out_dev, 0, sizeof(double) * (size_t)(pitch * N) * (size_t)aux_count, ctx->stream));
  ao::OutSpec out{out_dev, pitch * N, 1, pitch, aux_offset, aux_offset + aux_count, 1};
  launch_classes(ctx, *dft, dft->pair_classes, *aux, aux_offset, aux_offset + aux_count, out);
}

}  // namespace

extern "C" {

int gwbse_basis_normalize(int nshell, const int* l, const int* nprim, const double* exps, const double* contractions,
                          double* coefs_out) {
  if (nshell < 0 || !l || !nprim || !exps || !contractions || !coefs_out) return 1;
  size_t p0 = 0;
  for (int s = 0; s < nshell; ++s) {
    if (l[s] < 0 || l[s] > ao::LMAX_SHELL || nprim[s] < 1) return 1;
    ao::normalize_contraction(l[s], nprim[s], exps + p0, contractions + p0, coefs_out + p0);
    p0 += (size_t)nprim[s];
  }
  return 0;
}

int gwbse_basis_create(gwbse_ctx* ctx, int nshell, const int* l, const int* nprim, const double* centers,
                       const double* exps, const double* coefs, gwbse_basis** out) {
  GW_API_BEGIN(ctx)
  GW_REQUIRE(out && l && nprim && centers && exps && coefs && nshell > 0, "invalid basis description");
  std::unique_ptr<gwbse_basis> b(new gwbse_basis);
  b->host.build(nshell, l, nprim, centers, exps, coefs);
  build_basis(*b, ctx->device);
  *out = b.release();
  GW_API_END(ctx)
}

int gwbse_basis_destroy(gwbse_ctx* ctx, gwbse_basis* basis) {
  GW_API_BEGIN(ctx)
  GW_CUDA(cudaStreamSynchronize(ctx->stream));
  delete basis;
  GW_API_END(ctx)
}

int gwbse_basis_size(const gwbse_basis* basis) { return basis ? basis->host.nfunc : -1; }

int gwbse_ao3c_block_dev(gwbse_ctx* ctx, const gwbse_basis* aux, const gwbse_basis* dft, int aux_offset,
                         int aux_count, double* out_dev) {
  GW_API_BEGIN(ctx)
  GW_PROF(ctx, "ao3c_block");
  ao3c_block_dev(ctx, aux, dft, aux_offset, aux_count, out_dev);
  GW_API_END(ctx)
}

int gwbse_ao3c_block(gwbse_ctx* ctx, const gwbse_basis* aux, const gwbse_basis* dft, int aux_offset, int aux_count,
                     double* out) {
  GW_API_BEGIN(ctx)
  GW_PROF(ctx, "ao3c_block_d2h");
  GW_REQUIRE(aux && dft && out, "null argument");
  const size_t n = (size_t)dft->host.nfunc * dft->host.nfunc * (size_t)std::max(aux_count, 0);
  if (n == 0) return 0;
  double* d = ctx->buf("ao3c_out", n);
  ao3c_block_dev(ctx, aux, dft, aux_offset, aux_count, d);
  GW_CUDA(cudaMemcpyAsync(out, d, sizeof(double) * n, cudaMemcpyDeviceToHost, ctx->stream));
  GW_CUDA(cudaStreamSynchronize(ctx->stream));
  GW_API_END(ctx)
}

int gwbse_ao_coulomb2c(gwbse_ctx* ctx, const gwbse_basis* aux, double* V, int ld) {
  GW_API_BEGIN(ctx)
  GW_PROF(ctx, "ao_coulomb2c");
  GW_REQUIRE(aux && V, "null argument");
  const int n = aux->host.nfunc;
  GW_REQUIRE(ld >= n, "leading dimension too small");
  double* d = ctx->buf("ao2c_out", (size_t)n * n);
  // element (k | mu) at d[k * n + mu]; the unit partner contributes no index
  ao::OutSpec out{d, (long long)n, 1, 0, 0, n, 0};
  launch_classes(ctx, *aux, aux->unit_classes, *aux, 0, n, out);
  GW_CUDA(copy2d_async(V, sizeof(double) * ld, d, sizeof(double) * n, sizeof(double) * n, n, cudaMemcpyDeviceToHost,
                       ctx->stream));
  GW_CUDA(cudaStreamSynchronize(ctx->stream));
  GW_API_END(ctx)
}

int gwbse_ao_overlap(gwbse_ctx* ctx, const gwbse_basis* basis, double* S, int ld) {
  GW_API_BEGIN(ctx)
  GW_PROF(ctx, "ao_overlap");
  GW_REQUIRE(basis && S, "null argument");
  GW_REQUIRE(basis->device == ctx->device, "basis belongs to another device");
  const int n = basis->host.nfunc;
  GW_REQUIRE(ld >= n, "leading dimension too small");
  double* d = ctx->buf("ao2c_out", (size_t)n * n);
  if (basis->pairs_kept < basis->pairs_total) GW_CUDA(cudaMemsetAsync(d, 0, sizeof(double) * (size_t)n * n, ctx->stream));
  ao::OutSpec out{d, 0, 1, (long long)n, 0, 1, 1};
  const int smem_limit = shared_memory_limit(ctx);
  for (const auto& pc : basis->pair_classes) launch_class(ctx, *basis, pc, *basis, basis->no_aux_dev, 1, 0, out, smem_limit);
  GW_CUDA(copy2d_async(S, sizeof(double) * ld, d, sizeof(double) * n, sizeof(double) * n, n, cudaMemcpyDeviceToHost,
                       ctx->stream));
  GW_CUDA(cudaStreamSynchronize(ctx->stream));
  GW_API_END(ctx)
}

int gwbse_ao_dipole(gwbse_ctx* ctx, const gwbse_basis* basis, double* D, int ld) {
  GW_API_BEGIN(ctx)
  GW_PROF(ctx, "ao_dipole");
  GW_REQUIRE(basis && D, "null argument");
  GW_REQUIRE(basis->device == ctx->device, "basis belongs to another device");
  const int n = basis->host.nfunc;
  GW_REQUIRE(ld >= n, "leading dimension too small");
  double* d = ctx->buf("ao2c_out", (size_t)n * n);
  const int smem_limit = shared_memory_limit(ctx);
  for (int k = 0; k < 3; ++k) {
    if (basis->pairs_kept < basis->pairs_total)
      GW_CUDA(cudaMemsetAsync(d, 0, sizeof(double) * (size_t)n * n, ctx->stream));
    ao::OutSpec out{d, 0, 1, (long long)n, 0, 1, 1};
    for (const auto& pc : basis->pair_classes)
      launch_class(ctx, *basis, pc, *basis, basis->no_aux_dev + 1 + k, 1, 0, out, smem_limit);
    GW_CUDA(copy2d_async(D + (size_t)k * ld * n, sizeof(double) * ld, d, sizeof(double) * n, sizeof(double) * n, n,
                         cudaMemcpyDeviceToHost, ctx->stream));
    GW_CUDA(cudaStreamSynchronize(ctx->stream));
  }
  GW_API_END(ctx)
}

int gwbse_mmn_fill_from_basis(gwbse_ctx* ctx, const gwbse_basis* aux, const gwbse_basis* dft, int aux_block) {
  GW_API_BEGIN(ctx)
  GW_REQUIRE(aux && dft, "null argument");
  GW_REQUIRE(ctx->X != nullptr, "Mmn not allocated (gwbse_mmn_alloc)");
  mmn_complete_rotation(ctx);
  GW_REQUIRE(aux->host.nfunc == ctx->naux, "aux basis does not match the Mmn tensor");
  GW_REQUIRE(dft->host.nfunc == ctx->nbasis, "orbital basis does not match the MO coefficients (gwbse_mmn_set_mos)");
  if (aux_block < 1) aux_block = 64;
  // even pitch: the contraction GEMM stages 16-byte vectors (an odd basis size would force 8-byte copies)
  const long long pitch = (ctx->nbasis + 1) / 2 * 2;
  const size_t per = (size_t)pitch * ctx->nbasis;
  // keep one block below 4 GiB
  aux_block = (int)std::max<size_t>(1, std::min<size_t>(aux_block, ((size_t)1 << 29) / std::max<size_t>(per, 1)));
  if (gwbse_mmn_fill_begin(ctx, ctx->world > 1 ? 1 : 0)) return 1;
  int lo = 0, hi = ctx->naux;
  if (ctx->world > 1) {
    lo = ctx->aux_begin(ctx->rank);
    hi = ctx->aux_begin(ctx->rank + 1);
  }
  double* blk = ctx->buf("ao3c_block", per * aux_block);
  for (int a0 = lo; a0 < hi; a0 += aux_block) {
    const int cnt = std::min(aux_block, hi - a0);
    {
      GW_PROF(ctx, "ao3c_block");
      ao3c_block_dev(ctx, aux, dft, a0, cnt, blk, pitch);
    }
    // same stream: the contraction GEMMs of this block run after its integrals, the next block's integrals after them
    mmn_fill_block_pitched(ctx, a0, cnt, blk, pitch);
  }
  if (gwbse_mmn_fill_end(ctx)) return 1;
  GW_API_END(ctx)
}

}  // extern "C"
