// Internal definition of the opaque handles of include/zfvm.h.
#pragma once
#include <cuda_runtime.h>

#include <string>
#include <vector>

#include "../../include/zfvm.h"
#include "device/layout.hpp"
#include "host/gravity.hpp"
#include "host/handles.hpp"
#include "kernels/kernels.hpp"

struct HaloPeer {
  int rank;
  std::int64_t recv_begin, recv_end;
  std::int64_t send_begin, send_end;  // rows in the packed send buffer
};

struct zfvm_ctx {
  int device = 0;
  cudaStream_t stream = nullptr, comm_stream = nullptr;
  cudaEvent_t ev_a = nullptr, ev_b = nullptr;
  zfvm::SchemeConst sc{};
  zfvm::DevicePlan plan{};
  zfvm_params params{};
  int n_dims = 0, deg_hi = 0, deg_lo = 0;
  bool generic = false;  // stencil family outside the specialised kernels' shape: kernels/recon_generic.cu
  std::int64_t n_cells = 0, n_tiles = 0;
  std::vector<void *> allocations;
  std::int64_t device_bytes = 0;
  double algorithmic_bytes = 0.0;
  std::int64_t launches = 0;

  // resident state and Runge-Kutta buffers
  double *u_cur = nullptr, *u_tmp = nullptr;
  double *k[zfvm::MAX_RK_STAGES] = {nullptr};
  double *tend_work = nullptr;  // device tendency for the host entry point
  double *state_work = nullptr; // device state for the host entry point
  double *frozen = nullptr;      // FrozenBC steady state in use (null: NoBoundaryCondition)
  double *frozen_buf = nullptr, *frozen_a_buf = nullptr;  // its storage, kept and reused across zfvm_set_frozen_bc calls
  // advected scalars [n][n_avars]: resident values, RK buffers, host entry point work arrays, FrozenBC copy
  int n_avars = 0;
  double *a_cur = nullptr, *a_tmp = nullptr;
  double *ka[zfvm::MAX_RK_STAGES] = {nullptr};
  double *tend_work_a = nullptr, *state_work_a = nullptr, *frozen_a = nullptr;
  double *send_buf_a = nullptr;
  zfvm::TracerRecView tracer_view{};
  double *inradius = nullptr;
  std::int32_t *ghost_index = nullptr;
  std::int64_t n_ghost = 0;
  zfvm::ReduceOut *reduce_dev = nullptr;
  zfvm::ReduceOut *reduce_host = nullptr;  // pinned
  int *eq_fail_dev = nullptr;
  int n_stages = 0;
  double tab_a[zfvm::MAX_RK_STAGES][zfvm::MAX_RK_STAGES] = {{0}};
  double tab_b[zfvm::MAX_RK_STAGES] = {0};
  double tab_c[zfvm::MAX_RK_STAGES] = {0};
  double n_k_avg = 2.0;

  // optional per-kernel timing (cudaEvent pairs around K1 / K2 / K3), see zfvm_profile_*
  bool prof_enabled = false;
  std::vector<cudaEvent_t> prof_events[4];  // K1, K2, K3, tracer kernels (T1 + T2 + T3)

  // pinned staging chunks for copies from / to pageable caller memory
  void *stage[2] = {nullptr, nullptr};
  cudaEvent_t stage_ev[2] = {nullptr, nullptr};

  // zfvm_rk_step_host overlapped with its own copies (single-rank contexts): the state travels in chunks of rows;
  // stage 0 reconstructs a tile as soon as every row its stencils read has landed, the last stage finishes the cells
  // chunk by chunk (reconstruction of the chunk's tiles and of its later face neighbours, the chunk's faces, its rows)
  // and each finished chunk goes back while the next one is computed.
  struct HostPipe {
    int n_chunks = 0;
    std::vector<std::int64_t> cell_begin, face_begin;   // [n_chunks + 1]: first cell / first interior face of a chunk
    std::vector<std::int64_t> up_off, dn_off;           // [n_chunks + 1]: offsets into the two device tile lists
    std::int32_t *up_tiles = nullptr;                   // stage 0: tiles that become ready with upload chunk c
    std::int32_t *dn_tiles = nullptr;                   // last stage: tiles to reconstruct before chunk c can be finished
    std::vector<cudaEvent_t> ev_up, ev_dn;
    double *u_out = nullptr;                            // third state buffer: the chunked stage must not overwrite its input
  } pipe;
  cudaStream_t copy_stream = nullptr;

  // zfvm_rk_step replayed as a CUDA graph (single-rank contexts, no profiling): one captured step per (which buffer holds
  // the state, with / without the CFL reduction); the time step travels through a device scalar.  Small grids are bound
  // by launch latency otherwise (49 928 triangles: ~35 us of launches per 10 us of kernels per stage).
  struct StepGraph {
    cudaGraphExec_t exec = nullptr;
    const double *u_cur = nullptr, *u_tmp = nullptr;  // the state buffers of the captured step
    int epoch = -1;
    std::int64_t launches = 0;      // kernels in the graph (zfvm_counters)
  } step_graph[2][2];
  unsigned graph_victim = 0;
  int warm_epoch = -1;              // epoch in which a step has already run without a graph
  int graph_epoch = 0;              // bumped by whatever changes the captured launches (tableau, boundary condition)
  double *dt_dev = nullptr;         // device scalar
  double *dt_host = nullptr;        // pinned
  const double *capturing_dt = nullptr;  // set while rk_step_impl is being captured

  // multi-GPU
  void *nccl_comm = nullptr;
  int rank = 0, n_ranks = 1;
  bool halo_posted = false;  // zfvm_halo_post without its zfvm_halo_wait yet
  std::int64_t n_owned = 0;
  std::vector<HaloPeer> peers;
  std::int32_t *send_index_dev = nullptr;
  double *send_buf = nullptr;
  std::int64_t n_send = 0;
  std::vector<std::int32_t> tile_max_ref;  // per tile: largest cell index any of its stencils reads
  std::vector<std::uint8_t> tile_needed;   // per tile: some cell contributes a trace to the flux loop
  std::int32_t *tiles_needed = nullptr;    // device list of those tiles (null: all tiles)
  std::int64_t n_tiles_needed = 0;
  std::int32_t *tiles_interior = nullptr, *tiles_exterior = nullptr;
  std::int64_t n_tiles_interior = 0, n_tiles_exterior = 0;
};

namespace zfvm {
void set_error(const std::string &msg);
int fail(const std::string &msg);
#define ZFVM_CUDA(call)                                                                     \
  do {                                                                                      \
    cudaError_t err__ = (call);                                                             \
    if (err__ != cudaSuccess)                                                               \
      return ::zfvm::fail(std::string(#call) + ": " + cudaGetErrorString(err__));           \
  } while (0)
}  // namespace zfvm
