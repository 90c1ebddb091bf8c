// m-cyclic sharding helpers and the slice gather used by the multi-GPU paths of Sigma_x, off-diagonal
// Sigma_c and the BSE operator.
#include "../../include/gwbse_b200.h"
#include "context.cuh"

namespace gwbse {

namespace {

// send[(p * cnt_max + il) * rpad + row] = X[(p0+p) * ldx + (lfirst + il) * npad + row0 + row]  (zero padded)
__global__ void pack_slices_kernel(const double* __restrict__ X, long long ldx, int npad, int lfirst, int lcount,
                                   int cnt_max, int row0, int nrows, int rpad, int p0, double* __restrict__ send) {
  const int row = blockIdx.x * blockDim.x + threadIdx.x;
  const int il = blockIdx.y, p = blockIdx.z;
  if (row >= rpad) return;
  double v = 0.0;
  if (il < lcount && row < nrows) v = X[(long long)(p0 + p) * ldx + (long long)(lfirst + il) * npad + row0 + row];
  send[((long long)p * cnt_max + il) * rpad + row] = v;
}

// out[p * ldo + (s - s0) * rpad + row] = recv[r][(p * cnt_max + il) * rpad + row],  s = first_r + il * world
__global__ void unpack_slices_kernel(const double* __restrict__ recv, long long per_rank, int cnt_max, int rpad,
                                     int world, int s0, int ns, double* __restrict__ out, long long ldo) {
  const int row = blockIdx.x * blockDim.x + threadIdx.x;
  const int sl = blockIdx.y, p = blockIdx.z;  // sl = slice index relative to s0
  if (row >= rpad || sl >= ns) return;
  const int s = s0 + sl;
  const int r = s % world;
  const int first = s0 + ((r - s0 % world) % world + world) % world;
  const int il = (s - first) / world;
  out[(long long)p * ldo + (long long)sl * rpad + row] =
      recv[(long long)r * per_rank + ((long long)p * cnt_max + il) * rpad + row];
}

}  // namespace

void gather_slices(gwbse_ctx* ctx, int s0, int ns, int row0, int nrows, int p0, int np, double* out, long long ldo,
                   int rpad) {
  GW_REQUIRE(s0 >= 0 && s0 + ns <= ctx->mtotal && row0 >= 0 && row0 + nrows <= ctx->ntotal, "gather range");
  GW_REQUIRE(rpad >= nrows && ldo >= (long long)ns * rpad, "gather layout");
  if (ns <= 0 || np <= 0) return;
  const int world = ctx->world;
  const int cnt_max = (ns + world - 1) / world;
  const int lcount = ctx->owned_count(s0, ns, ctx->rank);
  const int lfirst = lcount ? ctx->local_index(ctx->first_owned(s0, ctx->rank)) : 0;
  const long long per_rank = (long long)np * cnt_max * rpad;
  GW_REQUIRE(np <= 65535 && cnt_max <= 65535 && ns <= 65535, "gather grid too large");
  double* send = ctx->buf("gather_send", (size_t)per_rank);
  double* recv = ctx->buf("gather_recv", (size_t)per_rank * world);
  dim3 gp((rpad + 127) / 128, cnt_max, np);
  pack_slices_kernel<<<gp, 128, 0, ctx->stream>>>(ctx->X, ctx->ldx, ctx->npad, lfirst, lcount, cnt_max, row0, nrows,
                                                  rpad, p0, send);
  GW_CUDA(cudaGetLastError());
  allgather_dev(ctx, send, recv, (size_t)per_rank);
  dim3 gu((rpad + 127) / 128, ns, np);
  unpack_slices_kernel<<<gu, 128, 0, ctx->stream>>>(recv, per_rank, cnt_max, rpad, world, s0, ns, out, ldo);
  GW_CUDA(cudaGetLastError());
  ctx->launches += 2;
}

}  // namespace gwbse

extern "C" {

int gwbse_shard_owner(int m, int world) { return world > 0 ? m % world : 0; }
int gwbse_shard_local_index(int m, int world) { return world > 0 ? m / world : m; }
int gwbse_shard_aux_begin(int naux, int rank, int world) {
  return world > 0 ? (int)((long long)rank * naux / world) : 0;
}
int gwbse_shard_local_count(int total, int rank, int world) {
  if (world <= 0 || rank < 0 || rank >= world) return 0;
  return total <= rank ? 0 : (total - rank + world - 1) / world;
}

}  // extern "C"
