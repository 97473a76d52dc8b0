// C ABI, multi-GPU part: halo exchange and MIN all-reduce over NCCL (NVLink 5 / NVSwitch).
//
// Replaces MPIHaloExchange (src/zisa/mpi/parallelization/mpi_halo_exchange.cpp:109-201): the rows a
// peer needs are gathered by a pack kernel, sent with ncclSend, and received in place into the
// contiguous halo rows [recv_begin, recv_end) of the state -- one ncclGroup per RK stage on a
// dedicated stream, overlapped with the reconstruction of the tiles that touch no halo cell.
// NCCL is resolved at run time from the library torch already loaded (libnccl.so.2).
#include <dlfcn.h>
#include <nccl.h>

#include <cstdlib>
#include <cstring>
#include <string>

#include "ctx.hpp"

using namespace zfvm;

namespace {

struct NcclApi {
  void *handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
};

NcclApi g_nccl;

int load_nccl() {
  if (g_nccl.handle) return 0;
  const char *env = std::getenv("ZFVM_NCCL_LIB");
  const char *names[] = {env, "libnccl.so.2", "libnccl.so"};
  void *h = nullptr;
  for (const char *nm : names) {
    if (!nm) continue;
    h = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
    if (h) break;
  }
  if (!h) return fail("NCCL library not found (import torch first, or set ZFVM_NCCL_LIB)");
#define ZFVM_SYM(field, name)                                        \
  g_nccl.field = (decltype(g_nccl.field))dlsym(h, name);             \
  if (!g_nccl.field) return fail(std::string("NCCL symbol missing: ") + name)
  ZFVM_SYM(GetUniqueId, "ncclGetUniqueId");
  ZFVM_SYM(CommInitRank, "ncclCommInitRank");
  ZFVM_SYM(CommDestroy, "ncclCommDestroy");
  ZFVM_SYM(GroupStart, "ncclGroupStart");
  ZFVM_SYM(GroupEnd, "ncclGroupEnd");
  ZFVM_SYM(Send, "ncclSend");
  ZFVM_SYM(Recv, "ncclRecv");
  ZFVM_SYM(AllReduce, "ncclAllReduce");
  ZFVM_SYM(GetErrorString, "ncclGetErrorString");
#undef ZFVM_SYM
  g_nccl.handle = h;
  return 0;
}

#define ZFVM_NCCL(call)                                                                      \
  do {                                                                                       \
    ncclResult_t r__ = (call);                                                               \
    if (r__ != ncclSuccess) return fail(std::string(#call) + ": " + g_nccl.GetErrorString(r__)); \
  } while (0)

}  // namespace

int zfvm_halo_post_internal(zfvm_ctx *ctx, double *state_dev, double *avars_dev) {
  // the state must be complete on the compute stream before it is packed
  ZFVM_CUDA(cudaEventRecord(ctx->ev_a, ctx->stream));
  ZFVM_CUDA(cudaStreamWaitEvent(ctx->comm_stream, ctx->ev_a, 0));
  launch_pack_rows(ctx->send_buf, state_dev, ctx->send_index_dev, ctx->n_send, ctx->comm_stream);
  ctx->launches += 1;
  const int na = (avars_dev && ctx->send_buf_a) ? ctx->n_avars : 0;  // avars rows travel in the same group
  if (na > 0) {
    launch_pack_rows_n(ctx->send_buf_a, avars_dev, ctx->send_index_dev, ctx->n_send, na, ctx->comm_stream);
    ctx->launches += 1;
  }
  ncclComm_t comm = (ncclComm_t)ctx->nccl_comm;
  ZFVM_NCCL(g_nccl.GroupStart());
  for (const HaloPeer &p : ctx->peers) {
    if (p.recv_end > p.recv_begin)
      ZFVM_NCCL(g_nccl.Recv(state_dev + p.recv_begin * NVARS, (size_t)((p.recv_end - p.recv_begin) * NVARS), ncclDouble,
                            p.rank, comm, ctx->comm_stream));
    if (p.send_end > p.send_begin)
      ZFVM_NCCL(g_nccl.Send(ctx->send_buf + p.send_begin * NVARS, (size_t)((p.send_end - p.send_begin) * NVARS),
                            ncclDouble, p.rank, comm, ctx->comm_stream));
    if (na > 0 && p.recv_end > p.recv_begin)
      ZFVM_NCCL(g_nccl.Recv(avars_dev + p.recv_begin * na, (size_t)((p.recv_end - p.recv_begin) * na), ncclDouble, p.rank,
                            comm, ctx->comm_stream));
    if (na > 0 && p.send_end > p.send_begin)
      ZFVM_NCCL(g_nccl.Send(ctx->send_buf_a + p.send_begin * na, (size_t)((p.send_end - p.send_begin) * na), ncclDouble,
                            p.rank, comm, ctx->comm_stream));
  }
  ZFVM_NCCL(g_nccl.GroupEnd());
  ZFVM_CUDA(cudaEventRecord(ctx->ev_b, ctx->comm_stream));
  return 0;
}

int zfvm_halo_wait_internal(zfvm_ctx *ctx) {
  ZFVM_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->ev_b, 0));
  return 0;
}

int zfvm_allreduce_min_internal(zfvm_ctx *ctx, double *dev_value) {
  ZFVM_NCCL(g_nccl.AllReduce(dev_value, dev_value, 1, ncclDouble, ncclMin, (ncclComm_t)ctx->nccl_comm, ctx->stream));
  return 0;
}

// min of the CFL quotient and max of the plausibility flag in one group (ReduceOut: double, int)
int zfvm_allreduce_verdict_internal(zfvm_ctx *ctx) {
  ncclComm_t comm = (ncclComm_t)ctx->nccl_comm;
  ZFVM_NCCL(g_nccl.GroupStart());
  ZFVM_NCCL(g_nccl.AllReduce(&ctx->reduce_dev->min_dx_over_ev, &ctx->reduce_dev->min_dx_over_ev, 1, ncclDouble, ncclMin, comm,
                             ctx->stream));
  ZFVM_NCCL(g_nccl.AllReduce(&ctx->reduce_dev->not_plausible, &ctx->reduce_dev->not_plausible, 1, ncclInt, ncclMax, comm,
                             ctx->stream));
  ZFVM_NCCL(g_nccl.GroupEnd());
  return 0;
}

void zfvm_comm_destroy_internal(zfvm_ctx *ctx) {
  if (ctx->nccl_comm && g_nccl.CommDestroy) {
    if (ctx->comm_stream) cudaStreamSynchronize(ctx->comm_stream);
    g_nccl.CommDestroy((ncclComm_t)ctx->nccl_comm);
  }
  ctx->nccl_comm = nullptr;
}

namespace {
template <class T>
void release_registered(zfvm_ctx *ctx, T *&ptr) {
  if (!ptr) return;
  for (size_t a = 0; a < ctx->allocations.size(); ++a)
    if (ctx->allocations[a] == (void *)ptr) {
      ctx->allocations.erase(ctx->allocations.begin() + (std::ptrdiff_t)a);
      break;
    }
  cudaFree((void *)ptr);
  ptr = nullptr;
}
}  // namespace

extern "C" {

int zfvm_nccl_unique_id(char id_out[128]) {
  if (load_nccl()) return 1;
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId size");
  ncclUniqueId id;
  ZFVM_NCCL(g_nccl.GetUniqueId(&id));
  std::memcpy(id_out, &id, sizeof(id));
  return 0;
}

int zfvm_comm_init(zfvm_ctx *ctx, const char id_in[128], int rank, int n_ranks) {
  if (load_nccl()) return 1;
  ZFVM_CUDA(cudaSetDevice(ctx->device));
  ncclUniqueId id;
  std::memcpy(&id, id_in, sizeof(id));
  zfvm_comm_destroy_internal(ctx);  // a second call replaces the communicator
  ncclComm_t comm;
  ZFVM_NCCL(g_nccl.CommInitRank(&comm, n_ranks, id, rank));
  ctx->nccl_comm = comm;
  ctx->rank = rank;
  ctx->n_ranks = n_ranks;
  return 0;
}

int zfvm_set_halo(zfvm_ctx *ctx, int64_t n_owned, int n_peers, const int *peer_rank, const int64_t *recv_begin,
                  const int64_t *recv_end, const int64_t *send_offset, const int32_t *send_index) {
  ZFVM_CUDA(cudaSetDevice(ctx->device));
  if (n_owned < 0 || n_owned > ctx->n_cells) return fail("zfvm_set_halo: n_owned out of range");
  ctx->n_owned = n_owned;
  ctx->peers.clear();
  for (int p = 0; p < n_peers; ++p) {
    if (recv_begin[p] < n_owned || recv_end[p] > ctx->n_cells || recv_end[p] < recv_begin[p])
      return fail("zfvm_set_halo: halo rows must lie behind the owned rows");
    ctx->peers.push_back(HaloPeer{peer_rank[p], recv_begin[p], recv_end[p], send_offset[p], send_offset[p + 1]});
  }
  ctx->n_send = n_peers > 0 ? send_offset[n_peers] : 0;
  for (int64_t r = 0; r < ctx->n_send; ++r)
    if (send_index[r] < 0 || send_index[r] >= n_owned) return fail("zfvm_set_halo: only owned rows can be sent");
  // a second call replaces the plan: the previous buffers are released, not kept until zfvm_destroy
  ZFVM_CUDA(cudaStreamSynchronize(ctx->stream));
  ZFVM_CUDA(cudaStreamSynchronize(ctx->comm_stream));
  release_registered(ctx, ctx->send_index_dev);
  release_registered(ctx, ctx->send_buf);
  release_registered(ctx, ctx->send_buf_a);
  release_registered(ctx, ctx->tiles_interior);
  release_registered(ctx, ctx->tiles_exterior);
  void *p = nullptr;
  ZFVM_CUDA(cudaMalloc(&p, (size_t)std::max<int64_t>(ctx->n_send, 1) * sizeof(int32_t)));
  ctx->allocations.push_back(p);
  ctx->send_index_dev = (int32_t *)p;
  if (ctx->n_send > 0)
    ZFVM_CUDA(cudaMemcpy(p, send_index, (size_t)ctx->n_send * sizeof(int32_t), cudaMemcpyHostToDevice));
  ZFVM_CUDA(cudaMalloc(&p, (size_t)std::max<int64_t>(ctx->n_send, 1) * NVARS * sizeof(double)));
  ctx->allocations.push_back(p);
  ctx->send_buf = (double *)p;
  ctx->device_bytes += ctx->n_send * (int64_t)(NVARS * sizeof(double) + sizeof(int32_t));
  if (ctx->n_avars > 0) {
    ZFVM_CUDA(cudaMalloc(&p, (size_t)std::max<int64_t>(ctx->n_send, 1) * ctx->n_avars * sizeof(double)));
    ctx->allocations.push_back(p);
    ctx->send_buf_a = (double *)p;
    ctx->device_bytes += ctx->n_send * (int64_t)(ctx->n_avars * sizeof(double));
  }

  // tiles whose stencils reference no halo row can be reconstructed before the exchange completes
  std::vector<int32_t> ti, te;
  for (int64_t t = 0; t < ctx->n_tiles; ++t) {
    if (!ctx->tile_needed[(size_t)t]) continue;
    if (ctx->tile_max_ref[(size_t)t] >= n_owned)
      te.push_back((int32_t)t);
    else
      ti.push_back((int32_t)t);
  }
  ctx->n_tiles_interior = (int64_t)ti.size();
  ctx->n_tiles_exterior = (int64_t)te.size();
  ZFVM_CUDA(cudaMalloc(&p, std::max<size_t>(ti.size(), 1) * sizeof(int32_t)));
  ctx->allocations.push_back(p);
  ctx->tiles_interior = (int32_t *)p;
  if (!ti.empty()) ZFVM_CUDA(cudaMemcpy(p, ti.data(), ti.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
  ZFVM_CUDA(cudaMalloc(&p, std::max<size_t>(te.size(), 1) * sizeof(int32_t)));
  ctx->allocations.push_back(p);
  ctx->tiles_exterior = (int32_t *)p;
  if (!te.empty()) ZFVM_CUDA(cudaMemcpy(p, te.data(), te.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
  return 0;
}

// HaloExchange::operator() and HaloExchange::wait() as the reference declares them
// (include/zisa/parallelization/halo_exchange.hpp:11-21): post returns immediately, the transfer runs on the context's
// communication stream; wait makes the compute stream (every later kernel of this context) wait for it.
int zfvm_halo_post(zfvm_ctx *ctx, double *state_dev, double *avars_dev) {
  if (ctx->n_ranks <= 1 || !ctx->nccl_comm) return 0;  // NoHaloExchange, halo_exchange.hpp:23-29
  ZFVM_CUDA(cudaSetDevice(ctx->device));
  if (!state_dev) state_dev = ctx->u_cur;
  if (ctx->n_avars > 0 && !avars_dev) avars_dev = ctx->a_cur;
  ctx->halo_posted = true;
  return zfvm_halo_post_internal(ctx, state_dev, ctx->n_avars > 0 ? avars_dev : nullptr);
}

int zfvm_halo_wait(zfvm_ctx *ctx) {
  if (ctx->n_ranks <= 1 || !ctx->nccl_comm) return 0;
  if (!ctx->halo_posted) return fail("zfvm_halo_wait: no exchange has been posted");
  ZFVM_CUDA(cudaSetDevice(ctx->device));
  ctx->halo_posted = false;
  return zfvm_halo_wait_internal(ctx);
}

int zfvm_halo_exchange(zfvm_ctx *ctx, double *state_dev) {
  if (ctx->n_ranks <= 1 || !ctx->nccl_comm) return 0;  // NoHaloExchange, halo_exchange.hpp:23-29
  ZFVM_CUDA(cudaSetDevice(ctx->device));
  if (!state_dev) state_dev = ctx->u_cur;
  if (zfvm_halo_post_internal(ctx, state_dev, nullptr)) return 1;
  return zfvm_halo_wait_internal(ctx);
}

int zfvm_halo_exchange_av(zfvm_ctx *ctx, double *state_dev, double *avars_dev) {
  if (ctx->n_ranks <= 1 || !ctx->nccl_comm) return 0;
  ZFVM_CUDA(cudaSetDevice(ctx->device));
  if (!state_dev) state_dev = ctx->u_cur;
  if (!avars_dev) avars_dev = ctx->a_cur;
  if (zfvm_halo_post_internal(ctx, state_dev, avars_dev)) return 1;
  return zfvm_halo_wait_internal(ctx);
}

int zfvm_allreduce_min(zfvm_ctx *ctx, double *value) {
  if (ctx->n_ranks <= 1 || !ctx->nccl_comm) return 0;
  ZFVM_CUDA(cudaSetDevice(ctx->device));
  ZFVM_CUDA(cudaMemcpyAsync(&ctx->reduce_dev->min_dx_over_ev, value, sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  if (zfvm_allreduce_min_internal(ctx, &ctx->reduce_dev->min_dx_over_ev)) return 1;
  ZFVM_CUDA(cudaMemcpyAsync(value, &ctx->reduce_dev->min_dx_over_ev, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  ZFVM_CUDA(cudaStreamSynchronize(ctx->stream));
  return 0;
}

}  // extern "C"
