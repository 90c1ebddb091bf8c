// NCCL plumbing (one process per GPU).  NCCL is loaded at run time with dlopen so that the
// single-GPU path has no NCCL dependency and a process that already imported torch shares
// torch's NCCL (same SONAME libnccl.so.2).  Replaces the host-side reduction of
// OpenMP_CUDA::getReductionVar (xtp/src/libxtp/openmp_cuda.cc:480-493).
#include <dlfcn.h>

#include <cstring>
#include <mutex>
#include <vector>

#include "../../include/gwbse_b200.h"
#include "context.cuh"

namespace {

typedef struct {
  char internal[128];
} nccl_uid;
typedef void* nccl_comm_t;

struct NcclApi {
  int (*GetUniqueId)(nccl_uid*) = nullptr;
  int (*CommInitRank)(nccl_comm_t*, int, nccl_uid, int) = nullptr;
  int (*CommDestroy)(nccl_comm_t) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, nccl_comm_t, cudaStream_t) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, nccl_comm_t, cudaStream_t) = nullptr;
  int (*Broadcast)(const void*, void*, size_t, int, int, nccl_comm_t, cudaStream_t) = nullptr;
  int (*Send)(const void*, size_t, int, int, nccl_comm_t, cudaStream_t) = nullptr;
  int (*Recv)(void*, size_t, int, int, nccl_comm_t, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  bool ok = false;
  std::string err;
};

NcclApi& nccl() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) {
      api.err = std::string("cannot load NCCL: ") + dlerror();
      return;
    }
    auto sym = [&](const char* n) { return dlsym(h, n); };
    api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(sym("ncclGetUniqueId"));
    api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(sym("ncclCommInitRank"));
    api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(sym("ncclCommDestroy"));
    api.AllReduce = reinterpret_cast<decltype(api.AllReduce)>(sym("ncclAllReduce"));
    api.AllGather = reinterpret_cast<decltype(api.AllGather)>(sym("ncclAllGather"));
    api.Broadcast = reinterpret_cast<decltype(api.Broadcast)>(sym("ncclBroadcast"));
    api.Send = reinterpret_cast<decltype(api.Send)>(sym("ncclSend"));
    api.Recv = reinterpret_cast<decltype(api.Recv)>(sym("ncclRecv"));
    api.GroupStart = reinterpret_cast<decltype(api.GroupStart)>(sym("ncclGroupStart"));
    api.GroupEnd = reinterpret_cast<decltype(api.GroupEnd)>(sym("ncclGroupEnd"));
    api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(sym("ncclGetErrorString"));
    api.ok = api.GetUniqueId && api.CommInitRank && api.AllReduce && api.AllGather && api.Broadcast;
    if (!api.ok) api.err = "NCCL library lacks required symbols";
  });
  return api;
}

constexpr int kNcclFloat64 = 8;  // ncclDouble
constexpr int kNcclSum = 0;

void check_nccl(int rc, const char* what) {
  if (rc != 0) {
    const char* s = nccl().GetErrorString ? nccl().GetErrorString(rc) : "?";
    throw std::runtime_error(std::string("NCCL error in ") + what + ": " + s);
  }
}

}  // namespace

namespace gwbse {

void allreduce_dev(gwbse_ctx* ctx, double* buf, size_t n) {
  if (ctx->world <= 1 || n == 0) return;
  GW_REQUIRE(ctx->nccl_comm != nullptr, "communicator not initialised (gwbse_comm_init)");
  check_nccl(nccl().AllReduce(buf, buf, n, kNcclFloat64, kNcclSum, ctx->nccl_comm, ctx->stream), "allreduce");
}

void allgather_dev(gwbse_ctx* ctx, const double* send, double* recv, size_t n_per_rank) {
  if (ctx->world <= 1) {
    if (send != recv)
      GW_CUDA(cudaMemcpyAsync(recv, send, sizeof(double) * n_per_rank, cudaMemcpyDeviceToDevice, ctx->stream));
    return;
  }
  GW_REQUIRE(ctx->nccl_comm != nullptr, "communicator not initialised (gwbse_comm_init)");
  check_nccl(nccl().AllGather(send, recv, n_per_rank, kNcclFloat64, ctx->nccl_comm, ctx->stream), "allgather");
}

void alltoallv_dev(gwbse_ctx* ctx, const double* const* send, const size_t* send_count, double* const* recv,
                   const size_t* recv_count) {
  if (ctx->world <= 1) {
    if (send_count[0] && send[0] != recv[0])
      GW_CUDA(cudaMemcpyAsync(recv[0], send[0], sizeof(double) * send_count[0], cudaMemcpyDeviceToDevice,
                              ctx->stream));
    return;
  }
  GW_REQUIRE(ctx->nccl_comm != nullptr, "communicator not initialised (gwbse_comm_init)");
  GW_REQUIRE(nccl().Send && nccl().Recv && nccl().GroupStart && nccl().GroupEnd, "NCCL lacks send/recv");
  check_nccl(nccl().GroupStart(), "ncclGroupStart");
  for (int r = 0; r < ctx->world; ++r) {
    if (send_count[r])
      check_nccl(nccl().Send(send[r], send_count[r], kNcclFloat64, r, ctx->nccl_comm, ctx->stream), "ncclSend");
    if (recv_count[r])
      check_nccl(nccl().Recv(recv[r], recv_count[r], kNcclFloat64, r, ctx->nccl_comm, ctx->stream), "ncclRecv");
  }
  check_nccl(nccl().GroupEnd(), "ncclGroupEnd");
}

}  // namespace gwbse

extern "C" {

int gwbse_nccl_unique_id(unsigned char* id128) {
  if (!nccl().ok) return 1;
  nccl_uid id;
  if (nccl().GetUniqueId(&id) != 0) return 1;
  std::memcpy(id128, id.internal, 128);
  return 0;
}

int gwbse_comm_init(gwbse_ctx* ctx, int rank, int world, const unsigned char* id128) {
  GW_API_BEGIN(ctx)
  GW_REQUIRE(world >= 1 && rank >= 0 && rank < world, "invalid rank/world");
  GW_REQUIRE(ctx->X == nullptr, "gwbse_comm_init must precede gwbse_mmn_alloc");
  ctx->rank = rank;
  ctx->world = world;
  if (world > 1) {
    if (!nccl().ok) throw std::runtime_error(nccl().err);
    nccl_uid id;
    std::memcpy(id.internal, id128, 128);
    nccl_comm_t comm = nullptr;
    check_nccl(nccl().CommInitRank(&comm, world, id, rank), "ncclCommInitRank");
    ctx->nccl_comm = comm;
  }
  GW_API_END(ctx)
}

int gwbse_comm_rank(const gwbse_ctx* ctx) { return ctx ? ctx->rank : 0; }
int gwbse_comm_world(const gwbse_ctx* ctx) { return ctx ? ctx->world : 1; }

int gwbse_comm_allreduce_host(gwbse_ctx* ctx, double* buf, size_t n) {
  GW_API_BEGIN(ctx)
  if (ctx->world > 1 && n > 0) {
    double* d = ctx->buf("comm_stage", n);
    GW_CUDA(cudaMemcpyAsync(d, buf, sizeof(double) * n, cudaMemcpyHostToDevice, ctx->stream));
    gwbse::allreduce_dev(ctx, d, n);
    GW_CUDA(cudaMemcpyAsync(buf, d, sizeof(double) * n, cudaMemcpyDeviceToHost, ctx->stream));
    GW_CUDA(cudaStreamSynchronize(ctx->stream));
  }
  GW_API_END(ctx)
}

}  // extern "C"
