// C ABI, device part: context creation (layout + upload), residual evaluation, fused
// Runge-Kutta step, boundary condition, CFL (include/zfvm.h).
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <chrono>
#include <cstring>
#include <exception>
#include <string>

#include "ctx.hpp"

using namespace zfvm;

namespace {

template <class T>
int dev_alloc(zfvm_ctx *ctx, T **ptr, std::int64_t count, bool zero = false) {
  const size_t bytes = (size_t)std::max<std::int64_t>(count, 1) * sizeof(T);
  void *p = nullptr;
  ZFVM_CUDA(cudaMalloc(&p, bytes));
  if (zero) ZFVM_CUDA(cudaMemsetAsync(p, 0, bytes, ctx->stream));
  ctx->allocations.push_back(p);
  ctx->device_bytes += (std::int64_t)bytes;
  *ptr = (T *)p;
  return 0;
}

/// Frees a buffer that dev_alloc registered (setters that are called again replace their buffers).
template <class T>
void dev_release(zfvm_ctx *ctx, T *&ptr, std::int64_t count) {
  if (!ptr) return;
  auto it = std::find(ctx->allocations.begin(), ctx->allocations.end(), (void *)ptr);
  if (it != ctx->allocations.end()) ctx->allocations.erase(it);
  cudaFree((void *)ptr);
  ctx->device_bytes -= (std::int64_t)((size_t)std::max<std::int64_t>(count, 1) * sizeof(T));
  ptr = nullptr;
}

template <class T, class A>
int dev_upload(zfvm_ctx *ctx, const T **ptr, const std::vector<T, A> &host) {
  T *p = nullptr;
  if (dev_alloc(ctx, &p, (std::int64_t)host.size())) return 1;
  if (!host.empty()) ZFVM_CUDA(cudaMemcpy(p, host.data(), host.size() * sizeof(T), cudaMemcpyHostToDevice));
  *ptr = p;
  return 0;
}

// Host <-> device copies of caller-owned buffers.  The reference's RungeKutta hands us plain
// zisa::array storage that alternates between calls (runge_kutta.cpp:109-111).  Buffers the caller has
// pinned go straight to the copy engine; pageable ones are staged through two internal pinned chunks that
// OpenMP threads fill / drain while the other chunk is in flight.  Caller memory is never registered here:
// its lifetime is not ours.
constexpr size_t STAGE_BYTES = 16u << 20;

bool is_pinned(const void *p) {
  cudaPointerAttributes attr;
  if (cudaPointerGetAttributes(&attr, p) == cudaSuccess && attr.type == cudaMemoryTypeHost) return true;
  cudaGetLastError();
  return false;
}

int ensure_stage(zfvm_ctx *ctx) {
  if (ctx->stage[0]) return 0;
  for (int b = 0; b < 2; ++b) {
    ZFVM_CUDA(cudaMallocHost(&ctx->stage[b], STAGE_BYTES));
    ZFVM_CUDA(cudaEventCreateWithFlags(&ctx->stage_ev[b], cudaEventDisableTiming));
  }
  return 0;
}

void parallel_copy(void *dst, const void *src, size_t bytes) {
  const std::int64_t blk = 1 << 20, nb = (std::int64_t)((bytes + blk - 1) / blk);
#pragma omp parallel for schedule(static)
  for (std::int64_t b = 0; b < nb; ++b) {
    const size_t off = (size_t)(b * blk);
    std::memcpy((char *)dst + off, (const char *)src + off, std::min<size_t>((size_t)blk, bytes - off));
  }
}

// asynchronous with respect to the host only for pinned sources
int copy_h2d(zfvm_ctx *ctx, void *dst_dev, const void *src_host, size_t bytes, cudaStream_t stream = nullptr) {
  if (bytes == 0) return 0;
  if (!stream) stream = ctx->stream;
  if (is_pinned(src_host)) {
    ZFVM_CUDA(cudaMemcpyAsync(dst_dev, src_host, bytes, cudaMemcpyHostToDevice, stream));
    return 0;
  }
  if (ensure_stage(ctx)) return 1;
  int b = 0;
  for (size_t off = 0; off < bytes; off += STAGE_BYTES, b ^= 1) {
    const size_t nb = std::min(STAGE_BYTES, bytes - off);
    ZFVM_CUDA(cudaEventSynchronize(ctx->stage_ev[b]));  // the previous transfer out of this chunk is done
    parallel_copy(ctx->stage[b], (const char *)src_host + off, nb);
    ZFVM_CUDA(cudaMemcpyAsync((char *)dst_dev + off, ctx->stage[b], nb, cudaMemcpyHostToDevice, stream));
    ZFVM_CUDA(cudaEventRecord(ctx->stage_ev[b], stream));
  }
  return 0;
}

// pageable destinations: returns after the data is in dst_host; pinned ones: after the copy is queued when `wait` is
// false (the caller synchronises the stream), after it has completed otherwise
int copy_d2h(zfvm_ctx *ctx, void *dst_host, const void *src_dev, size_t bytes, cudaStream_t stream = nullptr,
             bool wait = true) {
  if (bytes == 0) return 0;
  if (!stream) stream = ctx->stream;
  if (is_pinned(dst_host)) {
    ZFVM_CUDA(cudaMemcpyAsync(dst_host, src_dev, bytes, cudaMemcpyDeviceToHost, stream));
    if (wait) ZFVM_CUDA(cudaStreamSynchronize(stream));
    return 0;
  }
  if (ensure_stage(ctx)) return 1;
  int b = 0;
  size_t prev_off = 0, prev_nb = 0;
  for (size_t off = 0; off < bytes; off += STAGE_BYTES, b ^= 1) {
    const size_t nb = std::min(STAGE_BYTES, bytes - off);
    ZFVM_CUDA(cudaEventSynchronize(ctx->stage_ev[b]));
    ZFVM_CUDA(cudaMemcpyAsync(ctx->stage[b], (const char *)src_dev + off, nb, cudaMemcpyDeviceToHost, stream));
    ZFVM_CUDA(cudaEventRecord(ctx->stage_ev[b], stream));
    if (prev_nb) {  // drain the other chunk while this one is in flight
      ZFVM_CUDA(cudaEventSynchronize(ctx->stage_ev[b ^ 1]));
      parallel_copy((char *)dst_host + prev_off, ctx->stage[b ^ 1], prev_nb);
    }
    prev_off = off;
    prev_nb = nb;
  }
  ZFVM_CUDA(cudaEventSynchronize(ctx->stage_ev[b ^ 1]));
  parallel_copy((char *)dst_host + prev_off, ctx->stage[b ^ 1], prev_nb);
  return 0;
}

struct Tableau {
  int n;
  double a[MAX_RK_STAGES][MAX_RK_STAGES];
  double b[MAX_RK_STAGES];
};

// make_tableau, src/zisa/ode/runge_kutta.cpp:145-213
bool make_tableau(const std::string &m, Tableau &t) {
  std::memset(&t, 0, sizeof(t));
  if (m == "forward_euler") {
    t.n = 1;
    t.b[0] = 1.0;
  } else if (m == "ssp2") {
    t.n = 2;
    t.a[1][0] = 1.0;
    t.b[0] = 0.5;
    t.b[1] = 0.5;
  } else if (m == "ssp3") {
    t.n = 3;
    t.a[1][0] = 1.0;
    t.a[2][0] = 0.25;
    t.a[2][1] = 0.25;
    t.b[0] = 1.0 / 6;
    t.b[1] = 1.0 / 6;
    t.b[2] = 2.0 / 3;
  } else if (m == "wicker") {
    t.n = 3;
    t.a[1][0] = 1.0 / 3;
    t.a[2][1] = 0.5;
    t.b[2] = 1.0;
  } else if (m == "rk4") {
    t.n = 4;
    t.a[1][0] = 0.5;
    t.a[2][1] = 0.5;
    t.a[3][2] = 1.0;
    t.b[0] = 1.0 / 6;
    t.b[1] = 1.0 / 3;
    t.b[2] = 1.0 / 3;
    t.b[3] = 1.0 / 6;
  } else if (m == "fehlberg") {
    t.n = 6;
    t.a[1][0] = 0.25;
    t.a[2][0] = 3.0 / 32.0;
    t.a[2][1] = 9.0 / 32.0;
    t.a[3][0] = 1932.0 / 2197.0;
    t.a[3][1] = -7200.0 / 2197.0;
    t.a[3][2] = 7296.0 / 2197.0;
    t.a[4][0] = 439.0 / 216.0;
    t.a[4][1] = -8.0;
    t.a[4][2] = 3680.0 / 513.0;
    t.a[4][3] = -845.0 / 4104;
    t.a[5][0] = -8.0 / 27.0;
    t.a[5][1] = 2.0;
    t.a[5][2] = -3544.0 / 2565.0;
    t.a[5][3] = 1859.0 / 4104.0;
    t.a[5][4] = -11.0 / 40.0;
    t.b[0] = 16.0 / 135.0;
    t.b[2] = 6656.0 / 12825.0;
    t.b[3] = 28561.0 / 56430.0;
    t.b[4] = -9.0 / 50.0;
    t.b[5] = 2.0 / 55.0;
  } else {
    return false;
  }
  return true;
}

// One residual evaluation: K1 (reconstruction + traces + source), K2 (face fluxes), K3 (gather /
// update).  In a multi-rank context the halo exchange is posted first and the tiles whose stencils
// touch no halo cell are reconstructed while it is in flight (flux_loop.hpp:96-104, with a correct
// interior set, see SURVEY.md 5 "Distributed backend").
int residual(zfvm_ctx *ctx, const double *state, const UpdateArgs &upd, const double *avars = nullptr,
             const UpdateArgs *upd_av = nullptr);

}  // namespace

int zfvm_halo_post_internal(zfvm_ctx *ctx, double *state_dev, double *avars_dev);
int zfvm_halo_wait_internal(zfvm_ctx *ctx);
int zfvm_allreduce_min_internal(zfvm_ctx *ctx, double *dev_value);
int zfvm_allreduce_verdict_internal(zfvm_ctx *ctx);
void zfvm_comm_destroy_internal(zfvm_ctx *ctx);

namespace {

int run_recon(zfvm_ctx *ctx, const double *state, const std::int32_t *tiles, std::int64_t n_tiles) {
  if (ctx->generic) return launch_recon_generic(ctx->plan, ctx->sc, state, tiles, n_tiles, ctx->stream);
  return launch_recon(ctx->plan, ctx->sc, ctx->deg_hi, ctx->deg_lo, state, tiles, n_tiles, ctx->stream);
}

// kernels behind one run_recon call: K1 itself plus the equilibrium / source kernels that follow the same tile list
int recon_group_launches(const zfvm_ctx *ctx) {
  int n = 1;
  if (ctx->sc.well_balanced) n += ctx->plan.rec2 ? 4 : 2;          // E1, E2 (+ E3, S1 with tile records)
  else if (ctx->sc.has_gravity && ctx->plan.rec2) n += 1;          // S1
  if (ctx->plan.eq_flag != nullptr) n += 1;                        // E0
  return n;
}

void prof_mark(zfvm_ctx *ctx, int which) {
  if (!ctx->prof_enabled) return;
  if (ctx->prof_events[which].size() >= 2 * 16384) return;  // bounded: profiling left on keeps the first 16 k residuals
  cudaEvent_t ev;
  cudaEventCreate(&ev);
  cudaEventRecord(ev, ctx->stream);
  ctx->prof_events[which].push_back(ev);
}

int residual(zfvm_ctx *ctx, const double *state, const UpdateArgs &upd, const double *avars, const UpdateArgs *upd_av) {
  int rc;
  if (ctx->n_avars > 0 && (!avars || !upd_av))
    return fail("this context carries advected scalars: use the *_av entry points (AllVariables has cvars and avars)");
  prof_mark(ctx, 0);
  if (ctx->n_ranks > 1 && ctx->nccl_comm) {
    if (zfvm_halo_post_internal(ctx, const_cast<double *>(state), const_cast<double *>(avars))) return 1;
    rc = run_recon(ctx, state, ctx->tiles_interior, ctx->n_tiles_interior);
    if (zfvm_halo_wait_internal(ctx)) return 1;  // (also when rc != 0: the posted group must complete)
    if (rc) return fail("no reconstruction kernel is compiled for this scheme");
    rc = run_recon(ctx, state, ctx->tiles_exterior, ctx->n_tiles_exterior);
    if (rc) return fail("no reconstruction kernel is compiled for this scheme");
    ctx->launches += 2 * recon_group_launches(ctx);
  } else {
    rc = run_recon(ctx, state, ctx->tiles_needed, ctx->n_tiles_needed);
    if (rc) return fail("no reconstruction kernel is compiled for this scheme");
    ctx->launches += recon_group_launches(ctx);
  }
  prof_mark(ctx, 0);
  if (ctx->n_avars > 0) {
    // advected scalars, T1: scalar reconstruction + traces (after the halo rows have arrived, before the face kernel,
    // which upwinds them with the wave speeds of its HLLC evaluation)
    prof_mark(ctx, 3);
    const int trc = ctx->generic ? launch_tracer_recon_generic(ctx->plan, ctx->sc, avars, ctx->tiles_needed,
                                                               ctx->n_tiles_needed, ctx->stream)
                                 : launch_tracer_recon(ctx->plan, ctx->sc, ctx->tracer_view, ctx->deg_hi, ctx->deg_lo, avars,
                                                       ctx->tiles_needed, ctx->n_tiles_needed, ctx->stream);
    if (trc) return fail("no tracer reconstruction kernel is compiled for this scheme");
    prof_mark(ctx, 3);
    ctx->launches += 1;
  }
  prof_mark(ctx, 1);
  launch_flux(ctx->plan, ctx->sc, nullptr, ctx->plan.n_interior_edges, ctx->stream);
  prof_mark(ctx, 1);
  prof_mark(ctx, 2);
  UpdateArgs upd_bc = upd;
  upd_bc.flux_bc_state = ctx->params.flux_bc ? state : nullptr;  // FluxBC is part of the rate of change
  upd_bc.flux_bc_kind = ctx->params.flux_bc;
  launch_update(ctx->plan, ctx->sc, upd_bc, ctx->stream);
  if (ctx->n_avars > 0) {  // T3: gather / RK update of the avars rows
    launch_tracer_update(ctx->plan, ctx->n_dims, *upd_av, ctx->stream);
    ctx->launches += 1;
  }
  prof_mark(ctx, 2);
  ctx->launches += 2;
  ZFVM_CUDA(cudaGetLastError());
  return 0;
}

// the avars half of a residual's update arguments: same tableau row, the avars buffers
UpdateArgs avars_update_args(zfvm_ctx *ctx, const UpdateArgs &A) {
  UpdateArgs B = A;
  B.n_avars = ctx->n_avars;
  B.has_source = 0;
  B.reduce_out = nullptr;
  B.flux_bc_state = nullptr;
  B.tendency = nullptr;
  B.u_next = nullptr;
  B.u_base = nullptr;
  B.frozen = nullptr;
  for (int j = 0; j < MAX_RK_STAGES; ++j) B.k_prev[j] = nullptr;
  return B;
}

UpdateArgs base_update_args(zfvm_ctx *ctx) {
  UpdateArgs A;
  std::memset(&A, 0, sizeof(A));
  // halo rows are refreshed by the next exchange: a decomposed run updates (and reduces over) owned rows only
  A.n_cells_update = (ctx->n_ranks > 1) ? ctx->n_owned : ctx->n_cells;
  A.has_source = ctx->sc.has_gravity;
  A.gamma = ctx->sc.gamma;
  A.inradius = ctx->inradius;
  A.dt_dev = ctx->capturing_dt;  // (non-null only while a step is being captured into a graph)
  return A;
}

}  // namespace

extern "C" {

void zfvm_params_default(zfvm_params *p) {
  std::memset(p, 0, sizeof(*p));
  p->recon_mode = RECON_CWENO_AO;
  for (int k = 0; k < 8; ++k) p->linear_weights[k] = 1.0;
  p->linear_weights[0] = 100.0;
  p->epsilon = 1e-10;
  p->exponent = 4.0;
  p->well_balanced = 0;
  p->scaling = SCALING_EULER;
  p->flux = FLUX_HLLC;
  p->gamma = 1.4;
  p->gas_constant = 1.0;
  p->gravity_kind = GRAVITY_NONE;
  p->steps_per_recompute = 1;
  p->recompute_threshold = 0.0;
  p->flux_bc = 0;
}

static int create_impl(zfvm_ctx *ctx, const zfvm_grid *grid, const zfvm_stencils *stencils, const zfvm_params *params);

int zfvm_create(const zfvm_grid *grid, const zfvm_stencils *stencils, const zfvm_params *params, int device,
                zfvm_ctx **out) {
  zfvm_ctx *ctx = nullptr;
  try {
    const HostGrid &g = grid->g;
    const HostStencils &S = stencils->s;
    const int ns = S.n_stencils;
    if (S.n_cells != g.n_cells) return fail("zfvm_create: stencils do not belong to this grid");
    if (params->steps_per_recompute < 1) return fail("zfvm_create: steps_per_recompute must be >= 1");
    if (params->well_balanced && params->gravity_kind == GRAVITY_NONE)
      return fail("zfvm_create: isentropic well-balancing needs a gravity model");
    if (g.q_f > MAX_QF || g.q_c > MAX_QC) return fail("zfvm_create: quadrature rule too large");
    if (params->n_avars < 0 || params->n_avars > MAX_AVARS) return fail("zfvm_create: n_avars must be in [0, 8]");
    if (params->flux_bc < 0 || params->flux_bc > 2) return fail("zfvm_create: unknown flux_bc");
    if (params->flux_bc == 2 && params->gravity_kind == GRAVITY_NONE)
      return fail("zfvm_create: EquilibriumFluxBC needs a gravity model (equilibrium_flux_bc.hpp:18-35)");
    if (ns > MAX_STENCILS) return fail("zfvm_create: too many stencils");
    int n_dev = 0;
    if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0)
      return fail("zfvm_create: no CUDA device available (the B200 path has no CPU fallback)");
    ZFVM_CUDA(cudaSetDevice(device));

    ctx = new zfvm_ctx();
    ctx->device = device;
    // every failure below leaves through this one place: zfvm_destroy releases whatever had been allocated
    if (create_impl(ctx, grid, stencils, params)) {
      zfvm_destroy(ctx);
      return 1;
    }
    *out = ctx;
    return 0;
  } catch (const std::exception &e) {
    if (ctx) zfvm_destroy(ctx);
    return fail(std::string("zfvm_create: ") + e.what());
  }
}

// The body of zfvm_create once the arguments are validated and the device is selected; any non-zero return is followed
// by zfvm_destroy(ctx) in the caller (allocations are registered in ctx->allocations as they are made).
// Plan of the chunked host step (zfvm_ctx::HostPipe).  Chunks are ranges of consecutive cells (multiples of 64: the
// update kernel's block).  Interior faces are numbered in the order of their left (smaller) cell (host/grid.cpp), so
// the faces whose left cell lies in a chunk are a range too, and every face of a cell has its other cell either in
// the same or an earlier chunk (the cell is the right one: the face belongs to that chunk's range) or in a later one
// (the cell is the left one: the right cell's tile is reconstructed early, with this chunk).
static int build_host_pipe(zfvm_ctx *ctx, const HostGrid &g) {
  zfvm_ctx::HostPipe &H = ctx->pipe;
  const std::int64_t n = g.n_cells, T = ctx->n_tiles, EI = g.n_interior_edges;
  int want = 12;
  if (const char *e = std::getenv("ZFVM_HOST_CHUNKS")) want = std::atoi(e);
  // small grids (the extra launches cost more than the copies take), advected scalars: the plain sequence
  std::int64_t min_cells = 1 << 20;
  if (const char *e = std::getenv("ZFVM_HOST_PIPELINE_MIN_CELLS")) min_cells = std::atoll(e);
  if (want < 2 || n < min_cells || (n + 63) / 64 < 2 * want || ctx->n_avars > 0) return 0;
  for (std::int64_t e = 1; e < EI; ++e)
    if (g.left_right[(size_t)(2 * e)] < g.left_right[(size_t)(2 * e - 2)]) return 0;  // (not this library's numbering)
  const std::int64_t blocks = (n + 63) / 64;
  const int C = (int)std::min<std::int64_t>(want, blocks);
  H.cell_begin.assign((size_t)C + 1, n);
  for (int c = 0; c < C; ++c) H.cell_begin[(size_t)c] = std::min(n, (blocks * c / C) * 64);
  auto chunk_of = [&](std::int64_t cell) {
    return (int)(std::upper_bound(H.cell_begin.begin(), H.cell_begin.end(), cell) - H.cell_begin.begin()) - 1;
  };
  H.face_begin.assign((size_t)C + 1, EI);
  {
    std::int64_t e = 0;
    for (int c = 0; c < C; ++c) {
      while (e < EI && g.left_right[(size_t)(2 * e)] < H.cell_begin[(size_t)c]) ++e;
      H.face_begin[(size_t)c] = e;
    }
  }
  // stage 0: a tile is ready when the chunk holding the largest row index its stencils read has landed
  std::vector<std::vector<std::int32_t>> up((size_t)C), dn((size_t)C);
  for (std::int64_t t = 0; t < T; ++t)
    if (ctx->tile_needed[(size_t)t]) up[(size_t)chunk_of(ctx->tile_max_ref[(size_t)t])].push_back((std::int32_t)t);
  // last stage: chunk c needs its own tiles and the tiles of the right cells of its faces
  std::vector<std::uint8_t> done((size_t)T, 0);
  for (int c = 0; c < C; ++c) {
    auto take = [&](std::int64_t t) {
      if (ctx->tile_needed[(size_t)t] && !done[(size_t)t]) {
        done[(size_t)t] = 1;
        dn[(size_t)c].push_back((std::int32_t)t);
      }
    };
    for (std::int64_t t = H.cell_begin[(size_t)c] / TILE; t * TILE < H.cell_begin[(size_t)c + 1]; ++t) take(t);
    for (std::int64_t e = H.face_begin[(size_t)c]; e < H.face_begin[(size_t)c + 1]; ++e)
      take(g.left_right[(size_t)(2 * e + 1)] / TILE);
    std::sort(dn[(size_t)c].begin(), dn[(size_t)c].end());
  }
  std::vector<std::int32_t> up_all, dn_all;
  H.up_off.assign((size_t)C + 1, 0);
  H.dn_off.assign((size_t)C + 1, 0);
  for (int c = 0; c < C; ++c) {
    up_all.insert(up_all.end(), up[(size_t)c].begin(), up[(size_t)c].end());
    dn_all.insert(dn_all.end(), dn[(size_t)c].begin(), dn[(size_t)c].end());
    H.up_off[(size_t)c + 1] = (std::int64_t)up_all.size();
    H.dn_off[(size_t)c + 1] = (std::int64_t)dn_all.size();
  }
  const std::int32_t *p = nullptr;
  if (dev_upload(ctx, &p, up_all)) return 1;
  H.up_tiles = const_cast<std::int32_t *>(p);
  if (dev_upload(ctx, &p, dn_all)) return 1;
  H.dn_tiles = const_cast<std::int32_t *>(p);
  if (dev_alloc(ctx, &H.u_out, n * NVARS, true)) return 1;
  ZFVM_CUDA(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
  H.ev_up.resize((size_t)C);
  H.ev_dn.resize((size_t)C);
  for (int c = 0; c < C; ++c) {
    ZFVM_CUDA(cudaEventCreateWithFlags(&H.ev_up[(size_t)c], cudaEventDisableTiming));
    ZFVM_CUDA(cudaEventCreateWithFlags(&H.ev_dn[(size_t)c], cudaEventDisableTiming));
  }
  H.n_chunks = C;
  return 0;
}

// Record pieces the host builds (header, geometry -- and the weights when ZFVM_PRECOMPUTE=host) travel through two
// pinned staging slots as strided copies into the record array; the device builds the weights meanwhile.
struct RecordStager {
  static constexpr int MAX_PIECES = 3;
  int n_pieces = 0;
  std::int64_t off[MAX_PIECES] = {0, 0, 0}, bytes[MAX_PIECES] = {0, 0, 0};
  std::int64_t chunk = 0, rec_bytes = 0;
  char *dev = nullptr;
  char *host[2][MAX_PIECES] = {{nullptr, nullptr, nullptr}, {nullptr, nullptr, nullptr}};
  cudaEvent_t done[2] = {nullptr, nullptr};
  cudaStream_t stream = nullptr;

  void add_piece(std::int64_t offset, std::int64_t n_bytes) {
    if (n_bytes <= 0) return;
    off[n_pieces] = offset;
    bytes[n_pieces] = n_bytes;
    ++n_pieces;
  }
  cudaError_t init(char *dev_records, std::int64_t record_bytes, std::int64_t tiles_per_chunk, cudaStream_t st) {
    dev = dev_records;
    rec_bytes = record_bytes;
    chunk = tiles_per_chunk;
    stream = st;
    for (int s = 0; s < 2; ++s) {
      cudaError_t e = cudaEventCreateWithFlags(&done[s], cudaEventDisableTiming);
      if (e != cudaSuccess) return e;
      for (int p = 0; p < n_pieces; ++p) {
        e = cudaHostAlloc((void **)&host[s][p], (size_t)(chunk * bytes[p]), cudaHostAllocDefault);
        if (e != cudaSuccess) return e;
      }
    }
    return cudaSuccess;
  }
  /// Waits until slot s may be overwritten and clears it.
  cudaError_t begin(int s) {
    cudaError_t e = cudaEventSynchronize(done[s]);
    if (e != cudaSuccess) return e;
#pragma omp parallel for schedule(static)
    for (std::int64_t t = 0; t < chunk; ++t)
      for (int p = 0; p < n_pieces; ++p) std::memset(host[s][p] + t * bytes[p], 0, (size_t)bytes[p]);
    return cudaSuccess;
  }
  char *piece(int s, int p, std::int64_t tile_in_chunk) const { return host[s][p] + tile_in_chunk * bytes[p]; }
  cudaError_t upload(int s, std::int64_t t0, std::int64_t n_tiles) {
    for (int p = 0; p < n_pieces; ++p) {
      cudaError_t e = cudaMemcpy2DAsync(dev + t0 * rec_bytes + off[p], (size_t)rec_bytes, host[s][p], (size_t)bytes[p],
                                        (size_t)bytes[p], (size_t)n_tiles, cudaMemcpyHostToDevice, stream);
      if (e != cudaSuccess) return e;
    }
    return cudaEventRecord(done[s], stream);
  }
  ~RecordStager() {
    for (int s = 0; s < 2; ++s) {
      if (done[s]) cudaEventDestroy(done[s]);
      for (int p = 0; p < MAX_PIECES; ++p)
        if (host[s][p]) cudaFreeHost(host[s][p]);
    }
  }
};

// The device builds the stencil weights (kernels/precompute.cu) unless ZFVM_PRECOMPUTE=host asks for the host path,
// which is kept as the bit-for-bit cross-check.
static bool weights_on_device() {
  const char *e = std::getenv("ZFVM_PRECOMPUTE");
  return !(e && e[0] == 'h');
}

struct TempDeviceBuffers {
  std::vector<void *> ptrs;
  cudaError_t upload(const double **out, const double *host, size_t count, cudaStream_t st) {
    void *p = nullptr;
    cudaError_t e = cudaMalloc(&p, std::max<size_t>(count, 1) * sizeof(double));
    if (e != cudaSuccess) return e;
    ptrs.push_back(p);
    *out = (const double *)p;
    return cudaMemcpyAsync(p, host, count * sizeof(double), cudaMemcpyHostToDevice, st);
  }
  ~TempDeviceBuffers() {
    for (void *p : ptrs) cudaFree(p);
  }
};

// ZFVM_VERBOSE=1: wall-clock seconds of the phases of zfvm_create on stderr
struct CreateClock {
  bool on = std::getenv("ZFVM_VERBOSE") != nullptr;
  std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
  void lap(const char *what) {
    if (!on) return;
    const auto t1 = std::chrono::steady_clock::now();
    std::fprintf(stderr, "[zfvm create] %-28s %7.2f s\n", what, std::chrono::duration<double>(t1 - t0).count());
    t0 = t1;
  }
};

// zfvm_create in phases; the members are what more than one phase needs.
struct ContextBuilder {
  zfvm_ctx *ctx;
  const zfvm_params *params;
  const HostGrid &g;
  const HostStencils &S;
  const int nd, F, ns;
  const std::int64_t n, T, E, EI;
  SchemeConst &sc;
  DevicePlan &P;
  CreateClock clk;
  double bytes_W = 0.0, bytes_idx = 0.0, bytes_m = 0.0;  // of the counted (non-ghost) cells: SURVEY 8d
  std::int64_t n_counted = 0;
  bool use_tile = false;   // tile records (recon_tile.cuh / recon_coop.cuh) or the generic kernel's plainer records
  bool dev_w = true;       // stencil weights built on the device (precompute.cu)
  int D = 0;               // coefficients of the highest-order polynomial
  TempDeviceBuffers temp;  // inputs of the weight kernel, freed when the builder goes
  LsqWeightArgs lsq{};
  struct ScratchGuard {
    double *p = nullptr;
    ~ScratchGuard() {
      if (p) cudaFree(p);
    }
  } lsq_scratch_guard;
  double *&lsq_scratch = lsq_scratch_guard.p;
  std::int64_t lsq_scratch_bytes = 0;
  int n_sms = 148;

  ContextBuilder(zfvm_ctx *c, const zfvm_grid *grid, const zfvm_stencils *stencils, const zfvm_params *prm)
      : ctx(c), params(prm), g(grid->g), S(stencils->s), nd(grid->g.n_dims), F(grid->g.max_neighbours),
        ns(stencils->s.n_stencils), n(grid->g.n_cells), T((grid->g.n_cells + TILE - 1) / TILE), E(grid->g.n_edges),
        EI(grid->g.n_interior_edges), sc(c->sc), P(c->plan) {}

  int run() {
    if (int rc = streams_and_scheme_constants()) return rc;
    if (int rc = record_layouts()) return rc;
    if (int rc = weight_kernel_inputs()) return rc;
    if (int rc = use_tile ? tile_records() : generic_records()) return rc;
    if (lsq_scratch) {
      cudaFree(lsq_scratch);
      lsq_scratch = nullptr;
    }
    count_algorithmic_bytes();
    clk.lap("records (A, pinv, pack, H2D)");
    if (int rc = cell_geometry()) return rc;
    if (int rc = faces_and_gravity()) return rc;
    if (int rc = work_arrays()) return rc;
    return finish();
  }

  /// streams, SchemeConst (quadrature tables, linear weights, kernel family), empty DevicePlan
  int streams_and_scheme_constants() {
    ctx->params = *params;
    ctx->n_dims = nd;
    ctx->n_cells = g.n_cells;
    ctx->n_owned = g.n_cells;
    ctx->n_tiles = T;
    ZFVM_CUDA(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    ZFVM_CUDA(cudaStreamCreateWithFlags(&ctx->comm_stream, cudaStreamNonBlocking));
    ZFVM_CUDA(cudaEventCreateWithFlags(&ctx->ev_a, cudaEventDisableTiming));
    ZFVM_CUDA(cudaEventCreateWithFlags(&ctx->ev_b, cudaEventDisableTiming));

    // ---- scheme constants ------------------------------------------------------------------
    std::memset(&sc, 0, sizeof(sc));
    sc.n_dims = nd;
    sc.n_stencils = ns;
    sc.q_f = g.q_f;
    sc.q_c = g.q_c;
    sc.recon_mode = params->recon_mode;
    sc.scaling = params->scaling;
    sc.flux = params->flux;
    sc.well_balanced = params->well_balanced;
    // the cell-local source pass (GravitySourceLoop and / or Heating) runs when either term is present; without a
    // gravity model its potential tables stay zero
    sc.has_gravity = params->gravity_kind != GRAVITY_NONE || params->heating_rate != 0.0;
    {
      sc.eos_pow_e = 1.0 / (params->gamma - 1.0);
      const double twice = 2.0 * sc.eos_pow_e, r = std::rint(twice);
      sc.eos_pow_n = (std::fabs(twice - r) <= 1e-12 * twice && r >= 2.0 && r <= 8.0) ? (int)r : 0;
    }
    sc.steps_per_recompute = params->steps_per_recompute;
    sc.recompute_threshold = params->recompute_threshold;
    sc.heating_rate = params->heating_rate;
    sc.heating_r0 = params->heating_r0;
    sc.heating_r1 = params->heating_r1;
    ctx->n_avars = params->n_avars;
    sc.epsilon = params->epsilon;
    sc.exponent = params->exponent;
    sc.gamma = params->gamma;
    // Families the specialised kernels (tile / thread-per-cell with compile-time degrees) are built for: the shape of
    // every parameter set the reference's experiments use -- one leading stencil of the highest order followed by
    // n_dims + 1 stencils of order 2.  Anything else the reference's JSON can describe (first-order families, several
    // central stencils, one-sided stencils of order 3, a lone stencil; test/.../weno_ao.cpp:47-55, cweno_ao.cpp:144-160)
    // runs the generic kernel (kernels/recon_generic.cu), where every stencil keeps its own coefficient count.
    ctx->deg_hi = 0;
    for (int k = 0; k < ns; ++k) ctx->deg_hi = std::max(ctx->deg_hi, S.params.orders[(size_t)k] - 1);
    ctx->deg_lo = 0;
    for (int k = 1; k < ns; ++k) ctx->deg_lo = std::max(ctx->deg_lo, S.params.orders[(size_t)k] - 1);
    bool specialised = (ns == nd + 2) && S.params.orders[0] >= 2 && S.params.orders[0] - 1 == ctx->deg_hi;
    for (int k = 1; k < ns; ++k) specialised = specialised && S.params.orders[(size_t)k] == 2;
    // tests: ZFVM_RECON=generic (or v1, its older name) and, for runs with source terms, ZFVM_SOURCE=v1 put a family of the
    // specialised shape on the generic kernel as a cross-check
    if (const char *e = std::getenv("ZFVM_RECON")) specialised = specialised && e[0] != 'g' && e[0] != 'v';
    if (const char *e = std::getenv("ZFVM_SOURCE"))
      specialised = specialised && !(e[0] == 'v' && (params->gravity_kind != GRAVITY_NONE || params->heating_rate != 0.0));
    ctx->generic = !specialised;
    if ((nd == 2 && ctx->deg_hi >= 5) || (nd == 3 && ctx->deg_hi >= 4))
      return fail("zfvm_create: LSQ matrices exist up to order 5 in 2D and 4 in 3D (lsq_solver.cpp:288,399)");
    if (g.n_moments < poly_dof(ctx->deg_hi, nd))
      return fail("zfvm_create: grid moments_deg is lower than the polynomial degree");
    double wsum = 0.0;
    for (int k = 0; k < ns; ++k) wsum += params->linear_weights[k];
    for (int k = 0; k < ns; ++k) {
      sc.lin_w[k] = params->linear_weights[k] / wsum;  // hybrid_weno.cpp:26-31
      sc.rows_max[k] = S.max_size[(size_t)k] - 1;
      if (sc.rows_max[k] > 255)
        return fail("zfvm_create: a stencil may have at most 256 cells (8-bit row counts in the tile meta word)");
      sc.ncoef[k] = ctx->generic ? poly_dof(S.params.orders[(size_t)k] - 1, nd) - 1
                                 : poly_dof(k == 0 ? ctx->deg_hi : ctx->deg_lo, nd) - 1;
    }
    // probe the dispatch now: a family no kernel is compiled for must fail here, not in the first residual evaluation
    // (in a multi-rank run that would be after the NCCL group has been posted)
    if (ctx->generic && !recon_generic_supported(sc, poly_dof(ctx->deg_hi, nd)))
      return fail("zfvm_create: no reconstruction kernel for this stencil family (at most 6 stencils, order <= 5 in 2D / 4 in 3D)");
    for (int q = 0; q < g.q_f; ++q) {
      sc.face_w[q] = g.face_rule.weights[(size_t)q];
      for (int b = 0; b < g.face_rule.n_bary; ++b) sc.face_bary[q][b] = g.face_rule.bary[(size_t)(q * g.face_rule.n_bary + b)];
    }
    for (int q = 0; q < g.q_c; ++q) {
      sc.cell_w[q] = g.cell_rule.weights[(size_t)q];
      for (int b = 0; b < g.cell_rule.n_bary; ++b) sc.cell_bary[q][b] = g.cell_rule.bary[(size_t)(q * g.cell_rule.n_bary + b)];
    }

    std::memset(&P, 0, sizeof(P));
    P.n_cells = n;
    P.n_tiles = T;
    P.n_edges = E;
    P.n_interior_edges = EI;

    clk.lap("scheme constants");
    return 0;
  }

  /// byte layout of either record kind; row-list capacity of the tile records
  int record_layouts() {
    // ---- tile records (meta | sidx_k | W_k), built and uploaded in chunks of tiles -------------------
    ctx->tile_max_ref.assign((size_t)T, 0);
    for (std::int64_t t = 0; t < T; ++t) ctx->tile_max_ref[(size_t)t] = (std::int32_t)(std::min(n, (t + 1) * TILE) - 1);
    {
      int off = TILE * (int)sizeof(std::uint64_t);
      for (int k = 0; k < ns; ++k) {
        P.off_sidx[k] = off;
        off += sc.rows_max[k] * TILE * (int)sizeof(std::int32_t);
      }
      P.hdr_bytes = off;
      for (int k = 0; k < ns; ++k) {
        P.off_W[k] = off;
        off += sc.rows_max[k] * sc.ncoef[k] * TILE * (int)sizeof(double);
      }
      P.rec_bytes = off;  // every section is a multiple of 128 bytes
    }
    // ---- tile records (kernels/recon_tile.cuh, recon_coop.cuh): header | one-sided W | central W | geometry ----------
    // Built for every family of the specialised shape; the generic kernel reads the plainer records above.
    use_tile = !ctx->generic && ns >= 2 && recon_tile_compiled(sc, ctx->deg_hi, ctx->deg_lo);
    if (use_tile) {
      // distinct cells read by a tile's stencils
      std::vector<std::int32_t> n_union((size_t)T, 0);
  #pragma omp parallel
      {
        std::vector<std::int32_t> seen;
  #pragma omp for schedule(dynamic, 64)
        for (std::int64_t t = 0; t < T; ++t) {
          seen.clear();
          for (int lane = 0; lane < TILE; ++lane) {
            const std::int64_t i = t * TILE + lane;
            seen.push_back((std::int32_t)std::min(i, n - 1));
            if (i >= n) continue;
            for (int k = 0; k < S.n_family[(size_t)i]; ++k) {
              if (S.order[(size_t)(i * ns + k)] <= 1) continue;
              const int size = S.size[(size_t)(i * ns + k)];
              for (int j = 1; j < size; ++j) seen.push_back(S.global(i, k, j));
            }
          }
          std::sort(seen.begin(), seen.end());
          // the own cells occupy 32 list entries even when the last tile repeats the last cell
          const std::int64_t n_own_distinct = std::min<std::int64_t>(TILE, n - t * TILE);
          n_union[(size_t)t] = (std::int32_t)((std::unique(seen.begin(), seen.end()) - seen.begin()) + (TILE - n_own_distinct));
        }
      }
      int cap = TILE;
      for (std::int64_t t = 0; t < T; ++t) cap = std::max(cap, (int)n_union[(size_t)t]);
      cap = (cap + 31) / 32 * 32;
      if (const char *e = std::getenv("ZFVM_TILE_MIN_CAP")) cap = std::max(cap, (std::atoi(e) + 31) / 32 * 32);  // tests: 16-bit indices
      if (cap > 1024) {  // the shared-memory table would not fit: the generic kernel gathers through global indices
        use_tile = false;
        ctx->generic = true;
      }
      P.rec2_cap = cap;
      clk.lap("tile row-list sizes");
    }
    return 0;
  }

  /// what lsq_weights_kernel reads: centres, lengths, moments, the (order -> rows) table of every stencil
  int weight_kernel_inputs() {
    // ---- weights on the device: the members' centres / lengths / moments, the (order -> rows) table of every stencil -------
    dev_w = weights_on_device();
    ZFVM_CUDA(cudaDeviceGetAttribute(&n_sms, cudaDevAttrMultiProcessorCount, ctx->device));
    if (dev_w) {
      lsq.n_dims = nd;
      lsq.n_stencils = ns;
      lsq.n_moments = g.n_moments;
      for (int k = 0; k < ns; ++k) {
        lsq.ncoef[k] = sc.ncoef[k];
        lsq.max_order[k] = S.params.orders[(size_t)k];
        if (lsq.max_order[k] > 7) return fail("zfvm_create: stencil order above 7");
        for (int o = 2; o <= lsq.max_order[k]; ++o)
          lsq.rows_of_order[k][o] = required_stencil_size(o - 1, S.params.overfit_factors[(size_t)k], nd) - 1;
      }
      ZFVM_CUDA(temp.upload(&lsq.centers, g.cell_centers.data(), g.cell_centers.size(), ctx->stream));
      ZFVM_CUDA(temp.upload(&lsq.length, g.characteristic_length.data(), g.characteristic_length.size(), ctx->stream));
      ZFVM_CUDA(temp.upload(&lsq.moments, g.moments.data(), g.moments.size(), ctx->stream));
    }
    return 0;
  }

  /// tile records: header | one-sided W | central W | geometry (recon_tile.cuh, recon_coop.cuh)
  int tile_records() {
    const int D2 = poly_dof(ctx->deg_hi, nd);
    const TileRecLayout L = tile_rec_layout(sc, nd, D2, P.rec2_cap);
    P.rec2_bytes = L.rec_bytes;
    char *d_rec = nullptr;
    if (dev_alloc(ctx, &d_rec, T * L.rec_bytes + 65536)) {  // slack: the kernel may prefetch a little past the last record
              return 1;
    }
    P.rec2 = d_rec;
    P.rec2_off_list = L.off_list;
    P.rec2_off_lidx = L.off_lidx;
    P.rec2_lidx_elem = L.lidx_elem;
    std::vector<int> lo_row0((size_t)ns, 0), lo_w0((size_t)ns, 0);  // row / byte offsets of the one-sided stencils
    {
      int r = 0, b = 0;
      for (int k = 1; k < ns; ++k) {
        lo_row0[(size_t)k] = r;
        lo_w0[(size_t)k] = b;
        r += sc.rows_max[k];
        b += sc.rows_max[k] * sc.ncoef[k] * TILE * 8;
      }
      lo_row0[0] = r;  // the central stencil's rows come last
    }
    {
      TracerRecView &V = ctx->tracer_view;
      V.tile_record = 1;
      V.off_meta = TILE_OFF_META;
      V.off_list = L.off_list;
      V.off_lidx = L.off_lidx;
      V.lidx_elem = L.lidx_elem;
      for (int k = 0; k < ns; ++k) {
        V.row0[k] = lo_row0[(size_t)k];
        V.off_w[k] = (k == 0) ? L.off_whi : L.off_wlo + lo_w0[(size_t)k];
      }
    }
    const int n_mom2 = std::max(D2 - 3, 0);
    const std::int64_t chunk = std::max<std::int64_t>(1, std::min<std::int64_t>(4096, (256ll << 20) / L.rec_bytes));
    RecordStager stage;
    stage.add_piece(0, L.off_wlo);                       // n_list, meta, row list, local index rows
    stage.add_piece(L.off_geo, L.rec_bytes - L.off_geo);  // geometry
    if (!dev_w) stage.add_piece(L.off_wlo, L.off_geo - L.off_wlo);
    ZFVM_CUDA(stage.init(d_rec, L.rec_bytes, chunk, ctx->stream));
    ZFVM_CUDA(cudaMemsetAsync(d_rec, 0, (size_t)(T * L.rec_bytes + 65536), ctx->stream));
    if (dev_w) {
      lsq.rec = d_rec;
      lsq.rec_bytes = L.rec_bytes;
      lsq.view = ctx->tracer_view;
    }
    int slot = 0;
    for (std::int64_t t0 = 0; t0 < T; t0 += chunk, slot ^= 1) {
      const std::int64_t t1 = std::min(T, t0 + chunk);
      ZFVM_CUDA(stage.begin(slot));
  #pragma omp parallel
      {
        std::vector<double> A, W;
        std::vector<std::pair<std::int32_t, std::int32_t>> map;  // (global, local), sorted by global
        std::vector<std::int32_t> refs;
  #pragma omp for schedule(dynamic, 8)
        for (std::int64_t t = t0; t < t1; ++t) {
          char *rec = stage.piece(slot, 0, t - t0);
          char *rec_geo = stage.piece(slot, 1, t - t0);
          char *rec_w = dev_w ? nullptr : stage.piece(slot, 2, t - t0) - L.off_wlo;  // addressed with record offsets
          std::uint64_t *meta = reinterpret_cast<std::uint64_t *>(rec + TILE_OFF_META);
          std::int32_t *list = reinterpret_cast<std::int32_t *>(rec + L.off_list);
          unsigned char *lidx = reinterpret_cast<unsigned char *>(rec + L.off_lidx);
          auto put_lidx = [&](int row, int lane_, int value) {
            if (L.lidx_elem == 1)
              lidx[(size_t)row * TILE + lane_] = (unsigned char)value;
            else
              reinterpret_cast<std::uint16_t *>(lidx)[(size_t)row * TILE + lane_] = (std::uint16_t)value;
          };
          std::int32_t tile_mx = ctx->tile_max_ref[(size_t)t];
          // pass 1: the row list (own cells first, then the other stencil members in ascending order)
          refs.clear();
          for (int lane = 0; lane < TILE; ++lane) {
            const std::int64_t i = t * TILE + lane;
            list[lane] = (std::int32_t)std::min(i, n - 1);
            if (i >= n) continue;
            for (int k = 0; k < S.n_family[(size_t)i]; ++k) {
              if (S.order[(size_t)(i * ns + k)] <= 1) continue;
              const int size = S.size[(size_t)(i * ns + k)];
              for (int j = 1; j < size; ++j) refs.push_back(S.global(i, k, j));
            }
          }
          std::sort(refs.begin(), refs.end());
          refs.erase(std::unique(refs.begin(), refs.end()), refs.end());
          map.clear();
          const std::int32_t own_lo = (std::int32_t)(t * TILE), own_hi = (std::int32_t)std::min<std::int64_t>(n, (t + 1) * TILE);
          int n_list = TILE;
          for (std::int32_t gidx : refs) {
            if (gidx >= own_lo && gidx < own_hi) {
              map.emplace_back(gidx, gidx - own_lo);
            } else {
              list[n_list] = gidx;
              map.emplace_back(gidx, n_list++);
            }
            tile_mx = std::max(tile_mx, gidx);
          }
          *reinterpret_cast<std::int32_t *>(rec) = n_list;
          auto local_of = [&](std::int32_t gidx) {
            auto it = std::lower_bound(map.begin(), map.end(), std::make_pair(gidx, (std::int32_t)-1));
            return (int)it->second;
          };
          // pass 2: per cell meta, local indices, weights, geometry
          double *geo = reinterpret_cast<double *>(rec_geo);
          std::uint32_t *gref = reinterpret_cast<std::uint32_t *>(rec_geo + (size_t)L.geo_doubles * TILE * 8);
          for (int lane = 0; lane < TILE; ++lane) {
            const std::int64_t i = t * TILE + lane;
            std::uint64_t m = 0;
            for (int k = 0; k < ns; ++k) {
              const int RM = sc.rows_max[k], NC = sc.ncoef[k];
              const int row0 = lo_row0[(size_t)k];
              for (int j = 0; j < RM; ++j) put_lidx(row0 + j, lane, lane);  // padded rows: rhs == 0
              if (i >= n || k >= S.n_family[(size_t)i]) continue;
              const int order = S.order[(size_t)(i * ns + k)];
              if (order <= 1) continue;
              int rows = S.size[(size_t)(i * ns + k)] - 1, cols = poly_dof(order - 1, nd) - 1;
              if (cols > NC || rows > RM) continue;  // cannot happen: orders only degrade
              for (int j = 0; j < rows; ++j) put_lidx(row0 + j, lane, local_of(S.global(i, k, j + 1)));
              if (!dev_w) {
                stencil_matrix(A, rows, cols, g, S, i, k);
                W.resize((size_t)(rows * cols));
                pseudo_inverse(A.data(), rows, cols, W.data());
                double *w = reinterpret_cast<double *>(rec_w + (k == 0 ? L.off_whi : L.off_wlo + lo_w0[(size_t)k])) + lane;
                for (int j = 0; j < rows; ++j)
                  for (int c = 0; c < cols; ++c) w[(size_t)(j * NC + c) * TILE] = W[(size_t)(c * rows + j)];
              }
              m |= ((std::uint64_t)rows) << (8 * k);
            }
            if (i < n) {
              m |= ((std::uint64_t)(S.k_high[(size_t)i] & 0xF)) << 56;
              if (S.n_family[(size_t)i] == 1) m |= 1ull << 60;
            }
            meta[lane] = m;
            // geometry: vtx[F][nd] | centre[nd] | 1/len | moments | face_ref u32[F] | face slots (byte k: face k)
            const std::int64_t ic = std::min(i, n - 1);
            for (int k = 0; k < F; ++k) {
              const Vec3 v = g.vertex(ic, k);
              for (int d = 0; d < nd; ++d) geo[(size_t)((k * nd + d) * TILE + lane)] = v[d];
            }
            for (int d = 0; d < nd; ++d) geo[(size_t)((F * nd + d) * TILE + lane)] = g.cell_centers[(size_t)(3 * ic + d)];
            geo[(size_t)((F * nd + nd) * TILE + lane)] = 1.0 / g.characteristic_length[(size_t)ic];
            for (int mm = 0; mm < n_mom2; ++mm)
              geo[(size_t)((F * nd + nd + 1 + mm) * TILE + lane)] = g.moments[(size_t)(ic * g.n_moments + 3 + mm)];
            std::uint32_t slots_all = 0;
            for (int k = 0; k < F; ++k) {
              std::uint32_t r = 0;
              if (i < n) {
                const std::int64_t e = g.edge_indices[(size_t)(i * F + k)];
                const std::int32_t iL = g.left_right[(size_t)(2 * e)], iR = g.left_right[(size_t)(2 * e + 1)];
                r = (std::uint32_t)e & FREF_EDGE_MASK;
                if (iL != (std::int32_t)i) r |= FREF_SIDE;
                if (iR != INVALID) {
                  r |= FREF_INTERIOR;
                  const bool both_ghost = (g.cell_flags[(size_t)iL] & FLAG_GHOST) && (g.cell_flags[(size_t)iR] & FLAG_GHOST);
                  if (!both_ghost) r |= FREF_TRACE;  // flux_loop.hpp:82-87
                }
                slots_all |= ((std::uint32_t)g.face_vertex_slots[(size_t)(i * F + k)] & 0xFFu) << (8 * k);
              }
              gref[(size_t)(k * TILE + lane)] = r;
            }
            gref[(size_t)(F * TILE + lane)] = slots_all;
          }
          ctx->tile_max_ref[(size_t)t] = tile_mx;
        }
      }
      ZFVM_CUDA(stage.upload(slot, t0, t1 - t0));
      if (dev_w) {
        lsq.tile_begin = t0;
        lsq.tile_end = t1;
        if (launch_lsq_weights(lsq, n_sms, &lsq_scratch, &lsq_scratch_bytes, ctx->stream))
          return fail("zfvm_create: the stencil-weight kernel could not be launched");
      }
    }
    ZFVM_CUDA(cudaStreamSynchronize(ctx->stream));
    return 0;
  }

  /// records of the generic kernel: meta | global index rows | W_k
  int generic_records() {
    char *d_rec = nullptr;
    if (dev_alloc(ctx, &d_rec, T * P.rec_bytes)) {
      return 1;
    }
    P.rec = d_rec;
    {
      TracerRecView &V = ctx->tracer_view;
      V.tile_record = 0;
      V.off_meta = 0;
      for (int k = 0; k < ns; ++k) {
        V.off_sidx[k] = P.off_sidx[k];
        V.off_w[k] = P.off_W[k];
      }
    }
    const std::int64_t chunk = std::max<std::int64_t>(1, std::min<std::int64_t>(4096, (256ll << 20) / P.rec_bytes));
    RecordStager stage;
    stage.add_piece(0, P.hdr_bytes);  // meta, global index rows
    if (!dev_w) stage.add_piece(P.hdr_bytes, P.rec_bytes - P.hdr_bytes);
    ZFVM_CUDA(stage.init(d_rec, P.rec_bytes, chunk, ctx->stream));
    ZFVM_CUDA(cudaMemsetAsync(d_rec, 0, (size_t)(T * P.rec_bytes), ctx->stream));
    if (dev_w) {
      lsq.rec = d_rec;
      lsq.rec_bytes = P.rec_bytes;
      lsq.view = ctx->tracer_view;
    }
    int slot = 0;
    for (std::int64_t t0 = 0; t0 < T; t0 += chunk, slot ^= 1) {
      const std::int64_t t1 = std::min(T, t0 + chunk);
      ZFVM_CUDA(stage.begin(slot));
  #pragma omp parallel
      {
        std::vector<double> A, W;
  #pragma omp for schedule(dynamic, 8)
        for (std::int64_t t = t0; t < t1; ++t) {
          char *rec = stage.piece(slot, 0, t - t0);
          char *rec_w = dev_w ? nullptr : stage.piece(slot, 1, t - t0) - P.hdr_bytes;  // addressed with record offsets
          std::uint64_t *meta = reinterpret_cast<std::uint64_t *>(rec);
          std::int32_t tile_mx = ctx->tile_max_ref[(size_t)t];
          for (int lane = 0; lane < TILE; ++lane) {
            const std::int64_t i = t * TILE + lane;
            const std::int64_t ic = std::min(i, n - 1);
            std::uint64_t m = 0;
            for (int k = 0; k < ns; ++k) {
              const int RM = sc.rows_max[k], NC = sc.ncoef[k];
              std::int32_t *si = reinterpret_cast<std::int32_t *>(rec + P.off_sidx[k]) + lane;
              for (int j = 0; j < RM; ++j) si[(size_t)j * TILE] = (std::int32_t)ic;  // padded rows: rhs == 0
              if (i >= n || k >= S.n_family[(size_t)i]) continue;
              const int order = S.order[(size_t)(i * ns + k)];
              if (order <= 1) continue;
              int rows = S.size[(size_t)(i * ns + k)] - 1, cols = poly_dof(order - 1, nd) - 1;
              if (cols > NC || rows > RM) continue;  // cannot happen: orders only degrade
              for (int j = 0; j < rows; ++j) {
                si[(size_t)j * TILE] = S.global(i, k, j + 1);
                tile_mx = std::max(tile_mx, si[(size_t)j * TILE]);
              }
              if (!dev_w) {
                stencil_matrix(A, rows, cols, g, S, i, k);
                W.resize((size_t)(rows * cols));
                pseudo_inverse(A.data(), rows, cols, W.data());
                double *w = reinterpret_cast<double *>(rec_w + P.off_W[k]) + lane;
                for (int j = 0; j < rows; ++j)
                  for (int c = 0; c < cols; ++c) w[(size_t)(j * NC + c) * TILE] = W[(size_t)(c * rows + j)];
              }
              m |= ((std::uint64_t)rows) << (8 * k);
            }
            if (i < n) {
              m |= ((std::uint64_t)(S.k_high[(size_t)i] & 0xF)) << 56;
              if (S.n_family[(size_t)i] == 1) m |= 1ull << 60;
            }
            meta[lane] = m;
          }
          ctx->tile_max_ref[(size_t)t] = tile_mx;
        }
      }
      ZFVM_CUDA(stage.upload(slot, t0, t1 - t0));
      if (dev_w) {
        lsq.tile_begin = t0;
        lsq.tile_end = t1;
        if (launch_lsq_weights(lsq, n_sms, &lsq_scratch, &lsq_scratch_bytes, ctx->stream))
          return fail("zfvm_create: the stencil-weight kernel could not be launched");
      }
    }
    ZFVM_CUDA(cudaStreamSynchronize(ctx->stream));
    return 0;
  }

  void count_algorithmic_bytes() {
    for (std::int64_t i = 0; i < n; ++i) {
      if (!(g.cell_flags[(size_t)i] & FLAG_GHOST)) {
        ++n_counted;
        bytes_m += S.l2g_size[(size_t)i];
        for (int k = 0; k < S.n_family[(size_t)i]; ++k) {
          const int order = S.order[(size_t)(i * ns + k)], size = S.size[(size_t)(i * ns + k)];
          if (order > 1) bytes_W += 8.0 * (size - 1) * (poly_dof(order - 1, nd) - 1);
          bytes_idx += 4.0 * (size - 1);
        }
      }
    }
  }

  /// tile-interleaved cell geometry for K3 and the generic kernel, face references, tiles that need no reconstruction
  int cell_geometry() {
    // ---- geometry -----------------------------------------------------------------------------
    D = poly_dof(ctx->deg_hi, nd);
    P.n_mom = std::max(D - 3, 0);
    {
      // (filled from all threads: 2 GB of tables at 10 M tetrahedra; the padding lanes of the last tile keep the defaults)
      BigVec<double> vtx, center, inv_len, volume, mom;
      BigVec<std::uint32_t> fref;
      BigVec<std::uint8_t> fslots;
      parallel_assign(vtx, (size_t)(T * F * 3 * TILE), 0.0);
      parallel_assign(center, (size_t)(T * 3 * TILE), 0.0);
      parallel_assign(inv_len, (size_t)(T * TILE), 1.0);
      parallel_assign(volume, (size_t)(T * TILE), 1.0);
      parallel_assign(mom, (size_t)(T * std::max(P.n_mom, 1) * TILE), 0.0);
      parallel_assign(fref, (size_t)(T * F * TILE), (std::uint32_t)0);
      parallel_assign(fslots, (size_t)(T * F * TILE), (std::uint8_t)0);
  #pragma omp parallel for schedule(static)
      for (std::int64_t i = 0; i < n; ++i) {
        const std::int64_t t = i / TILE;
        const int lane = (int)(i % TILE);
        for (int k = 0; k < F; ++k) {
          const Vec3 v = g.vertex(i, k);
          for (int d = 0; d < 3; ++d) vtx[(size_t)(((t * F + k) * 3 + d) * TILE + lane)] = v[d];
          const std::int64_t e = g.edge_indices[(size_t)(i * F + k)];
          const std::int32_t iL = g.left_right[(size_t)(2 * e)], iR = g.left_right[(size_t)(2 * e + 1)];
          std::uint32_t r = (std::uint32_t)e & FREF_EDGE_MASK;
          if (iL != (std::int32_t)i) r |= FREF_SIDE;
          if (iR != INVALID) {
            r |= FREF_INTERIOR;
            const bool both_ghost = (g.cell_flags[(size_t)iL] & FLAG_GHOST) && (g.cell_flags[(size_t)iR] & FLAG_GHOST);
            if (!both_ghost) r |= FREF_TRACE;  // flux_loop.hpp:82-87
          }
          fref[(size_t)((t * F + k) * TILE + lane)] = r;
          fslots[(size_t)((t * F + k) * TILE + lane)] = g.face_vertex_slots[(size_t)(i * F + k)];
        }
        for (int d = 0; d < 3; ++d) center[(size_t)((t * 3 + d) * TILE + lane)] = g.cell_centers[(size_t)(3 * i + d)];
        inv_len[(size_t)i] = 1.0 / g.characteristic_length[(size_t)i];
        volume[(size_t)i] = g.volumes[(size_t)i];
        for (int m = 0; m < P.n_mom; ++m)
          mom[(size_t)((t * P.n_mom + m) * TILE + lane)] = g.moments[(size_t)(i * g.n_moments + 3 + m)];
      }
      // tiles none of whose cells contributes a trace to the flux loop (ghost cells deeper than the l1 layer)
      // are not reconstructed at all -- unless the caller wants every cell's polynomial back
      ctx->tile_needed.assign((size_t)T, 1);
      if (!params->keep_polynomials && !sc.has_gravity) {  // (the source loop visits every cell)
        for (std::int64_t t = 0; t < T; ++t) {
          bool any = false;
          for (std::int64_t a = t * F * TILE; a < (t + 1) * F * TILE && !any; ++a) any = (fref[(size_t)a] & FREF_TRACE) != 0;
          ctx->tile_needed[(size_t)t] = any ? 1 : 0;
        }
      }
      if (E > (std::int64_t)FREF_EDGE_MASK) {
        return fail("zfvm_create: too many faces for the packed face reference");
      }
      if (dev_upload(ctx, &P.vtx, vtx) || dev_upload(ctx, &P.center, center) || dev_upload(ctx, &P.inv_len, inv_len) ||
          dev_upload(ctx, &P.volume, volume) || dev_upload(ctx, &P.moments, mom) || dev_upload(ctx, &P.face_ref, fref) ||
          dev_upload(ctx, &P.face_slots, fslots) || dev_upload(ctx, &P.cell_flags, g.cell_flags)) {
        return 1;
      }
      const double *inr = nullptr;
      if (dev_upload(ctx, &inr, g.inradii)) {
        return 1;
      }
      ctx->inradius = const_cast<double *>(inr);
    }
    clk.lap("cell geometry");
    return 0;
  }

  /// face frames, left / right cells, gravity tables at all Gauss points
  int faces_and_gravity() {
    // ---- faces ----------------------------------------------------------------------------------
    {
      BigVec<std::int32_t> lr;  // (every element is written by the loop below: no value-initialisation pass)
      BigVec<double> frame;
      lr.resize((size_t)(2 * E));
      frame.resize((size_t)(10 * E));
  #pragma omp parallel for schedule(static)
      for (std::int64_t e = 0; e < E; ++e) {
        std::int32_t iL = g.left_right[(size_t)(2 * e)], iR = g.left_right[(size_t)(2 * e + 1)];
        bool skip = (iR == INVALID);
        if (!skip) skip = (g.cell_flags[(size_t)iL] & FLAG_GHOST) && (g.cell_flags[(size_t)iR] & FLAG_GHOST);
        lr[(size_t)(2 * e)] = skip ? -1 : iL;
        lr[(size_t)(2 * e + 1)] = iR;
        for (int d = 0; d < 3; ++d) {
          frame[(size_t)(10 * e + d)] = g.face_normal[(size_t)(3 * e + d)];
          frame[(size_t)(10 * e + 3 + d)] = g.face_t1[(size_t)(3 * e + d)];
          frame[(size_t)(10 * e + 6 + d)] = g.face_t2[(size_t)(3 * e + d)];
        }
        frame[(size_t)(10 * e + 9)] = g.face_area[(size_t)e];
      }
      if (dev_upload(ctx, &P.left_right, lr) || dev_upload(ctx, &P.face_frame, frame)) {
        return 1;
      }
    }
    // ---- gravity ----------------------------------------------------------------------------------
    if (sc.has_gravity) {
      double *a = nullptr, *b = nullptr, *c = nullptr;
      if (dev_alloc(ctx, &a, n * g.q_c, true) || dev_alloc(ctx, &b, n * g.q_c * 3, true) ||
          dev_alloc(ctx, &c, E * g.q_f, true)) {
        return 1;
      }
      P.phi_cqp = a;
      P.gradphi_cqp = b;
      P.phi_fqp = c;
      if (params->gravity_kind >= GRAVITY_CONSTANT && params->gravity_kind <= GRAVITY_POLYTROPE) {
        GravityModel gm;
        gm.kind = params->gravity_kind;
        gm.alignment = params->gravity_alignment;
        for (int q = 0; q < 4; ++q) gm.p[q] = params->gravity_p[q];
        if (gm.kind == GRAVITY_POINT_MASS && params->gravity_p[2] != 0.0) {
          // PointMassGravity(G, M, X): GM = G * M
          gm.p[0] = params->gravity_p[0] * params->gravity_p[1];
          gm.p[1] = params->gravity_p[2];
        }
        for (int d = 0; d < 3; ++d) gm.axis[d] = params->gravity_axis[d];
        std::vector<double> h_a, h_b, h_c;
        tabulate_gravity(gm, g, h_a, h_b, h_c);
        ZFVM_CUDA(cudaMemcpy(a, h_a.data(), h_a.size() * sizeof(double), cudaMemcpyHostToDevice));
        ZFVM_CUDA(cudaMemcpy(b, h_b.data(), h_b.size() * sizeof(double), cudaMemcpyHostToDevice));
        ZFVM_CUDA(cudaMemcpy(c, h_c.data(), h_c.size() * sizeof(double), cudaMemcpyHostToDevice));
      }
    }
    clk.lap("faces, gravity tables");
    return 0;
  }

  /// traces, fluxes, sources, equilibrium tables, resident state, scalars
  int work_arrays() {
    // ---- work arrays ------------------------------------------------------------------------------
    // (+ dump blocks for the tile kernel's branch-free trace write-out)
    if (dev_alloc(ctx, &P.trace, (std::max<std::int64_t>(EI, 1) + TRACE_DUMP_BLOCKS / 2) * 2 * g.q_f * NVARS, true) ||
        dev_alloc(ctx, &P.flux, std::max<std::int64_t>(EI, 1) * NVARS, true) || dev_alloc(ctx, &P.source, n * NVARS, true) ||
        dev_alloc(ctx, &ctx->eq_fail_dev, 1, true) || dev_alloc(ctx, &ctx->reduce_dev, 1, true)) {
      return 1;
      }
    P.eq_fail = ctx->eq_fail_dev;
    P.n_poly_coef = D;
    if (P.rec2 != nullptr && sc.has_gravity) {  // the tile kernel hands the polynomial to source_kernel
      if (dev_alloc(ctx, &P.poly_tile, T * (std::int64_t)(D + 1) * NVARS * TILE, true)) {
        return 1;
      }
    }
    if (params->keep_polynomials) {
      if (dev_alloc(ctx, &P.poly, n * D * NVARS, true) || dev_alloc(ctx, &P.poly_scale, n * NVARS, true)) {
        return 1;
      }
    }
    {
      std::vector<std::int32_t> tl;
      for (std::int64_t t = 0; t < T; ++t)
        if (ctx->tile_needed[(size_t)t]) tl.push_back((std::int32_t)t);
      ctx->n_tiles_needed = (std::int64_t)tl.size();
      if (ctx->n_tiles_needed < T) {
        const std::int32_t *p = nullptr;
        if (dev_upload(ctx, &p, tl)) {
          return 1;
      }
        ctx->tiles_needed = const_cast<std::int32_t *>(p);
      }
    }
    if (build_host_pipe(ctx, g)) return 1;
    ZFVM_CUDA(cudaMallocHost((void **)&ctx->reduce_host, sizeof(ReduceOut)));
    // ghost cells (FrozenBC::count_ghost_cells)
    {
      std::vector<std::int32_t> gi;
      for (std::int64_t i = 0; i < n; ++i)
        if (g.cell_flags[(size_t)i] & FLAG_GHOST) gi.push_back((std::int32_t)i);
      ctx->n_ghost = (std::int64_t)gi.size();
      const std::int32_t *p = nullptr;
      if (dev_upload(ctx, &p, gi)) {
        return 1;
      }
      ctx->ghost_index = const_cast<std::int32_t *>(p);
    }
    // resident state / RK buffers
    if (dev_alloc(ctx, &ctx->u_cur, n * NVARS, true) || dev_alloc(ctx, &ctx->u_tmp, n * NVARS, true) ||
        dev_alloc(ctx, &ctx->tend_work, n * NVARS, true) || dev_alloc(ctx, &ctx->state_work, n * NVARS, true)) {
      return 1;
      }

    // well-balanced runs: equilibrium parameters per cell and equilibrium averages per (cell, stencil row)
    if (sc.well_balanced) {
      int r = 0;
      for (int k = 0; k < ns; ++k) {
        P.eq_row0[k] = r;
        r += sc.rows_max[k];
      }
      P.eq_rows = P.rec2 ? r + 1 : r;  // tile records: rows in lidx order plus one row for the cell itself
      if (dev_alloc(ctx, &P.eq_par, n * 4, true) || dev_alloc(ctx, &P.eq_avg, T * (std::int64_t)P.eq_rows * 2 * TILE, true) ||
          (P.rec2 && dev_alloc(ctx, &P.eq_bg, std::max<std::int64_t>(EI, 1) * 2 * g.q_f * 2, true))) {
        return 1;
      }
    }
    // steps_per_recompute != 1: the per-cell history of LocalReconstruction (local_reconstruction.hpp:87-100).  The cached
    // equilibrium, its averages / point values and the scale live in the arrays the equilibrium kernels write anyway.
    if (params->steps_per_recompute != 1) {
      if (!P.rec2)
        return fail("zfvm_create: steps_per_recompute != 1 needs a stencil family of the experiments' shape (tile records)");
      if (dev_alloc(ctx, &P.eq_steps, n, true) || dev_alloc(ctx, &P.scale_state, n * 2, true) ||
          dev_alloc(ctx, &P.eq_flag, n, true)) {
        return 1;
      }
    }
    // advected scalars: traces, face fluxes, resident rows and host-entry work rows
    P.n_avars = ctx->n_avars;
    if (ctx->n_avars > 0) {
      const std::int64_t na = ctx->n_avars;
      if (dev_alloc(ctx, &P.qtrace, std::max<std::int64_t>(EI, 1) * 2 * g.q_f * na, true) ||
          dev_alloc(ctx, &P.qflux, std::max<std::int64_t>(EI, 1) * na, true) || dev_alloc(ctx, &ctx->a_cur, n * na, true) ||
          dev_alloc(ctx, &ctx->a_tmp, n * na, true) || dev_alloc(ctx, &ctx->tend_work_a, n * na, true) ||
          dev_alloc(ctx, &ctx->state_work_a, n * na, true)) {
        return 1;
      }
    }

    clk.lap("work arrays");
    if (clk.on)
      std::fprintf(stderr, "[zfvm create] %lld cells, %lld tiles, %lld reconstructed, row-list capacity %d, %lld B per tile record\n",
                   (long long)n, (long long)T, (long long)ctx->n_tiles_needed, P.rec2_cap, (long long)P.rec2_bytes);
    return 0;
  }

  /// algorithmic bytes per cell and stage, default tableau
  int finish() {
    // ---- algorithmic bytes per cell and stage -----------------------------------------
    {
      const double nc = (double)std::max<std::int64_t>(n_counted, 1);
      const double B_W = bytes_W / nc, B_idx = bytes_idx / nc + 4.0 * bytes_m / nc, m = bytes_m / nc;
      const double B_state = 80.0, B_poly = 2.0 * 40.0 * D;
      const double B_cell = 8.0 * (3 + 1 + 1 + D + 5) + (sc.has_gravity ? 8.0 * 4 * g.q_c : 0.0);
      const double B_face = (F / 2.0) * (8.0 * (9 + 4 * g.q_f) + 8.0);
      const double B_wb = sc.well_balanced ? 8.0 * (2 * m + 4.0 * (g.q_c + F * g.q_f)) : 0.0;
      ctx->algorithmic_bytes = B_W + B_idx + B_state + B_poly + B_cell + B_face + B_wb;  // + B_rk added per tableau
    }
    if (zfvm_set_time_integration(ctx, "ssp3")) {
      return 1;
      }
    ZFVM_CUDA(cudaStreamSynchronize(ctx->stream));
    return 0;
  }
};

static int create_impl(zfvm_ctx *ctx, const zfvm_grid *grid, const zfvm_stencils *stencils, const zfvm_params *params) {
  ContextBuilder b(ctx, grid, stencils, params);
  return b.run();
}

void zfvm_destroy(zfvm_ctx *ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  if (ctx->stream) cudaStreamSynchronize(ctx->stream);
  for (int b = 0; b < 2; ++b) {
    if (ctx->stage[b]) cudaFreeHost(ctx->stage[b]);
    if (ctx->stage_ev[b]) cudaEventDestroy(ctx->stage_ev[b]);
  }
  zfvm_comm_destroy_internal(ctx);
  for (auto &v : ctx->prof_events) {
    for (cudaEvent_t e : v) cudaEventDestroy(e);
    v.clear();
  }
  for (void *p : ctx->allocations) cudaFree(p);
  if (ctx->reduce_host) cudaFreeHost(ctx->reduce_host);
  if (ctx->ev_a) cudaEventDestroy(ctx->ev_a);
  if (ctx->ev_b) cudaEventDestroy(ctx->ev_b);
  if (ctx->stream) cudaStreamDestroy(ctx->stream);
  if (ctx->comm_stream) cudaStreamDestroy(ctx->comm_stream);
  for (auto &row : ctx->step_graph)
    for (auto &g : row)
      if (g.exec) cudaGraphExecDestroy(g.exec);
  if (ctx->dt_host) cudaFreeHost(ctx->dt_host);
  for (cudaEvent_t e : ctx->pipe.ev_up) cudaEventDestroy(e);
  for (cudaEvent_t e : ctx->pipe.ev_dn) cudaEventDestroy(e);
  if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
  delete ctx;
}

int zfvm_set_gravity_values(zfvm_ctx *ctx, const double *phi_cell_qp, const double *grad_phi_cell_qp,
                            const double *phi_face_qp) {
  if (!ctx->sc.has_gravity) return fail("zfvm_set_gravity_values: context was created without gravity");
  ZFVM_CUDA(cudaSetDevice(ctx->device));
  const std::int64_t n = ctx->n_cells, E = ctx->plan.n_edges;
  ZFVM_CUDA(cudaMemcpy(const_cast<double *>(ctx->plan.phi_cqp), phi_cell_qp, (size_t)(n * ctx->sc.q_c) * sizeof(double),
                       cudaMemcpyHostToDevice));
  ZFVM_CUDA(cudaMemcpy(const_cast<double *>(ctx->plan.gradphi_cqp), grad_phi_cell_qp,
                       (size_t)(n * ctx->sc.q_c * 3) * sizeof(double), cudaMemcpyHostToDevice));
  ZFVM_CUDA(cudaMemcpy(const_cast<double *>(ctx->plan.phi_fqp), phi_face_qp, (size_t)(E * ctx->sc.q_f) * sizeof(double),
                       cudaMemcpyHostToDevice));
  return 0;
}

int zfvm_set_gravity_table(zfvm_ctx *ctx, const zfvm_grid *grid, int64_t n, const double *radii, const double *phi) {
  if (!ctx->sc.has_gravity) return fail("zfvm_set_gravity_table: context was created without gravity");
  if (n < 2) return fail("zfvm_set_gravity_table: need at least two table points");
  GravityModel gm;
  gm.kind = GRAVITY_TABLE;
  gm.alignment = ctx->params.gravity_alignment;
  for (int d = 0; d < 3; ++d) gm.axis[d] = ctx->params.gravity_axis[d];
  gm.table_r.assign(radii, radii + n);
  gm.table_phi.assign(phi, phi + n);
  std::vector<double> a, b, c;
  tabulate_gravity(gm, grid->g, a, b, c);
  return zfvm_set_gravity_values(ctx, a.data(), b.data(), c.data());
}

int zfvm_memory_info(const zfvm_ctx *ctx, int64_t *device_bytes, double *algorithmic_bytes_per_cell_stage) {
  if (device_bytes) *device_bytes = ctx->device_bytes;
  if (algorithmic_bytes_per_cell_stage)
    *algorithmic_bytes_per_cell_stage = ctx->algorithmic_bytes + 40.0 * (1.0 + ctx->n_k_avg) + 40.0;
  return 0;
}

void *zfvm_stream(zfvm_ctx *ctx) { return (void *)ctx->stream; }

int zfvm_rate_of_change_device(zfvm_ctx *ctx, double *tendency_dev, const double *state_dev, double /*t*/,
                               int accumulate) {
  ZFVM_CUDA(cudaSetDevice(ctx->device));
  UpdateArgs A = base_update_args(ctx);
  A.tendency = tendency_dev;
  A.accumulate = accumulate;
  return residual(ctx, state_dev, A);
}

int zfvm_rate_of_change_av_device(zfvm_ctx *ctx, double *tendency_dev, double *tendency_avars_dev,
                                  const double *state_dev, const double *state_avars_dev, double /*t*/, int accumulate) {
  if (ctx->n_avars <= 0) return fail("zfvm_rate_of_change_av_device: the context was created with n_avars = 0");
  ZFVM_CUDA(cudaSetDevice(ctx->device));
  UpdateArgs A = base_update_args(ctx);
  A.tendency = tendency_dev;
  A.accumulate = accumulate;
  UpdateArgs B = avars_update_args(ctx, A);
  B.tendency = tendency_avars_dev;
  return residual(ctx, state_dev, A, state_avars_dev, &B);
}

int zfvm_rate_of_change_av(zfvm_ctx *ctx, double *tendency_host, double *tendency_avars_host, const double *state_host,
                           const double *state_avars_host, double t, int accumulate) {
  if (ctx->n_avars <= 0) return fail("zfvm_rate_of_change_av: the context was created with n_avars = 0");
  ZFVM_CUDA(cudaSetDevice(ctx->device));
  const size_t bytes = (size_t)(ctx->n_cells * NVARS) * sizeof(double);
  const size_t bytes_a = (size_t)(ctx->n_cells * ctx->n_avars) * sizeof(double);
  if (copy_h2d(ctx, ctx->state_work, state_host, bytes) || copy_h2d(ctx, ctx->state_work_a, state_avars_host, bytes_a)) return 1;
  if (accumulate && (copy_h2d(ctx, ctx->tend_work, tendency_host, bytes) ||
                     copy_h2d(ctx, ctx->tend_work_a, tendency_avars_host, bytes_a)))
    return 1;
  if (zfvm_rate_of_change_av_device(ctx, ctx->tend_work, ctx->tend_work_a, ctx->state_work, ctx->state_work_a, t, accumulate))
    return 1;
  if (ctx->n_ranks > 1) {
    const size_t na = (size_t)ctx->n_avars;
    for (auto &p : ctx->peers) {
      const size_t off = (size_t)(p.recv_begin * NVARS), cnt = (size_t)((p.recv_end - p.recv_begin) * NVARS);
      if (copy_d2h(ctx, const_cast<double *>(state_host) + off, ctx->state_work + off, cnt * sizeof(double))) return 1;
      const size_t off_a = (size_t)p.recv_begin * na, cnt_a = (size_t)(p.recv_end - p.recv_begin) * na;
      if (copy_d2h(ctx, const_cast<double *>(state_avars_host) + off_a, ctx->state_work_a + off_a, cnt_a * sizeof(double)))
        return 1;
    }
  }
  if (copy_d2h(ctx, tendency_avars_host, ctx->tend_work_a, bytes_a)) return 1;
  return copy_d2h(ctx, tendency_host, ctx->tend_work, bytes);
}

static void flux_and_update(zfvm_ctx *ctx, const double *state, UpdateArgs A, std::int64_t face_begin, std::int64_t face_end);

// RateOfChange::compute with host buffers overlapped with its own copies (zfvm_ctx::HostPipe, like the host time step):
// the state goes up in chunks while the tiles whose rows have landed are reconstructed; then the faces and cells are
// finished chunk by chunk and every finished chunk of the tendency goes down while the next one is computed.
static int rate_of_change_pipelined(zfvm_ctx *ctx, double *tendency_host, const double *state_host, int accumulate) {
  zfvm_ctx::HostPipe &H = ctx->pipe;
  const int C = H.n_chunks;
  // multi-rank contexts: as in rk_step_host_pipelined -- the chunks in front of the first one with halo rows gate interior
  // tiles only, the exchange is posted once every row is up, chunks of halo rows are not computed
  const bool multi = ctx->n_ranks > 1 && ctx->nccl_comm != nullptr;
  const std::int64_t n_upd = (ctx->n_ranks > 1) ? ctx->n_owned : ctx->n_cells;
  int c_halo = C;
  if (multi)
    for (int c = C - 1; c >= 0; --c)
      if (H.cell_begin[(size_t)c + 1] > ctx->n_owned) c_halo = c;
  ZFVM_CUDA(cudaStreamSynchronize(ctx->stream));
  for (int c = 0; c < C; ++c) {
    const std::int64_t r0 = H.cell_begin[(size_t)c] * NVARS, r1 = H.cell_begin[(size_t)c + 1] * NVARS;
    if (copy_h2d(ctx, ctx->state_work + r0, state_host + r0, (size_t)(r1 - r0) * sizeof(double), ctx->copy_stream)) return 1;
    ZFVM_CUDA(cudaEventRecord(H.ev_up[(size_t)c], ctx->copy_stream));
    ZFVM_CUDA(cudaStreamWaitEvent(ctx->stream, H.ev_up[(size_t)c], 0));
    const std::int64_t nt = H.up_off[(size_t)c + 1] - H.up_off[(size_t)c];
    if (nt > 0 && c < c_halo) {
      if (run_recon(ctx, ctx->state_work, H.up_tiles + H.up_off[(size_t)c], nt))
        return fail("no reconstruction kernel is compiled for this scheme");
      ctx->launches += recon_group_launches(ctx);
    }
  }
  if (multi) {
    if (zfvm_halo_post_internal(ctx, ctx->state_work, nullptr) || zfvm_halo_wait_internal(ctx)) return 1;
    const std::int64_t nt = H.up_off[(size_t)C] - H.up_off[(size_t)c_halo];
    if (nt > 0) {
      if (run_recon(ctx, ctx->state_work, H.up_tiles + H.up_off[(size_t)c_halo], nt))
        return fail("no reconstruction kernel is compiled for this scheme");
      ctx->launches += recon_group_launches(ctx);
    }
    // FluxLoop fills the halo rows of the caller's state (flux_loop.hpp:100, const_cast): they go back behind the exchange
    ZFVM_CUDA(cudaEventRecord(H.ev_up[0], ctx->stream));
    ZFVM_CUDA(cudaStreamWaitEvent(ctx->copy_stream, H.ev_up[0], 0));
    for (auto &p : ctx->peers) {
      const size_t off = (size_t)(p.recv_begin * NVARS), cnt = (size_t)((p.recv_end - p.recv_begin) * NVARS);
      if (copy_d2h(ctx, const_cast<double *>(state_host) + off, ctx->state_work + off, cnt * sizeof(double), ctx->copy_stream, false))
        return 1;
    }
  }
  if (accumulate) {  // the caller's tendency rows: behind the state on the copy stream, needed by the update kernel only
    if (copy_h2d(ctx, ctx->tend_work, tendency_host, (size_t)(ctx->n_cells * NVARS) * sizeof(double), ctx->copy_stream)) return 1;
    ZFVM_CUDA(cudaEventRecord(H.ev_up[0], ctx->copy_stream));
    ZFVM_CUDA(cudaStreamWaitEvent(ctx->stream, H.ev_up[0], 0));
  }
  UpdateArgs A = base_update_args(ctx);
  A.tendency = ctx->tend_work;
  A.accumulate = accumulate;
  for (int c = 0; c < C; ++c) {
    if (H.cell_begin[(size_t)c] < n_upd) {  // (a chunk of halo rows only travels)
      UpdateArgs Ac = A;
      Ac.block_begin = H.cell_begin[(size_t)c] / 64;
      Ac.n_cells_update = std::min(H.cell_begin[(size_t)c + 1], n_upd);
      flux_and_update(ctx, ctx->state_work, Ac, H.face_begin[(size_t)c], H.face_begin[(size_t)c + 1]);
    }
    ZFVM_CUDA(cudaEventRecord(H.ev_dn[(size_t)c], ctx->stream));
  }
  for (int c = 0; c < C; ++c) {
    const std::int64_t r0 = H.cell_begin[(size_t)c] * NVARS, r1 = H.cell_begin[(size_t)c + 1] * NVARS;
    ZFVM_CUDA(cudaStreamWaitEvent(ctx->copy_stream, H.ev_dn[(size_t)c], 0));
    if (copy_d2h(ctx, tendency_host + r0, ctx->tend_work + r0, (size_t)(r1 - r0) * sizeof(double), ctx->copy_stream, false)) return 1;
  }
  ZFVM_CUDA(cudaStreamSynchronize(ctx->copy_stream));
  ZFVM_CUDA(cudaStreamSynchronize(ctx->stream));
  ZFVM_CUDA(cudaGetLastError());
  return 0;
}

static bool host_pipeline_enabled(const zfvm_ctx *ctx) {
  static const bool off = [] {
    const char *e = std::getenv("ZFVM_HOST_PIPELINE");
    return e != nullptr && e[0] == '0';
  }();
  // (multi-rank contexts: the chunked routes post the same halo exchanges as the plain sequences)
  return ctx->pipe.n_chunks > 0 && !off;
}

int zfvm_rate_of_change(zfvm_ctx *ctx, double *tendency_host, const double *state_host, double t, int accumulate) {
  ZFVM_CUDA(cudaSetDevice(ctx->device));
  if (host_pipeline_enabled(ctx) && ctx->n_avars == 0) return rate_of_change_pipelined(ctx, tendency_host, state_host, accumulate);
  const size_t bytes = (size_t)(ctx->n_cells * NVARS) * sizeof(double);
  if (copy_h2d(ctx, ctx->state_work, state_host, bytes)) return 1;
  if (accumulate && copy_h2d(ctx, ctx->tend_work, tendency_host, bytes)) return 1;
  if (zfvm_rate_of_change_device(ctx, ctx->tend_work, ctx->state_work, t, accumulate)) return 1;
  if (ctx->n_ranks > 1) {
    // FluxLoop fills the halo rows of the caller's state (flux_loop.hpp:100, const_cast)
    for (auto &p : ctx->peers) {
      const size_t off = (size_t)(p.recv_begin * NVARS), cnt = (size_t)((p.recv_end - p.recv_begin) * NVARS);
      if (copy_d2h(ctx, const_cast<double *>(state_host) + off, ctx->state_work + off, cnt * sizeof(double))) return 1;
    }
  }
  return copy_d2h(ctx, tendency_host, ctx->tend_work, bytes);
}

int zfvm_set_time_integration(zfvm_ctx *ctx, const char *method) {
  ++ctx->graph_epoch;  // captured steps are stale
  Tableau t;
  if (!make_tableau(method, t)) return fail(std::string("Unknown Butcher Tableau. [") + method + "]");
  ZFVM_CUDA(cudaSetDevice(ctx->device));
  ctx->n_stages = t.n;
  std::memcpy(ctx->tab_a, t.a, sizeof(t.a));
  std::memcpy(ctx->tab_b, t.b, sizeof(t.b));
  double nk = 0.0;
  for (int s = 1; s <= t.n; ++s) {
    const double *row = (s < t.n) ? t.a[s] : t.b;
    for (int j = 0; j < t.n; ++j) nk += (row[j] != 0.0);
  }
  ctx->n_k_avg = nk / t.n;
  for (int s = 0; s + 1 < t.n; ++s) {
    if (!ctx->k[s] && dev_alloc(ctx, &ctx->k[s], ctx->n_cells * NVARS, true)) return 1;
    if (ctx->n_avars > 0 && !ctx->ka[s] && dev_alloc(ctx, &ctx->ka[s], ctx->n_cells * ctx->n_avars, true)) return 1;
  }
  return 0;
}

int zfvm_upload_state(zfvm_ctx *ctx, const double *state_host) {
  ZFVM_CUDA(cudaSetDevice(ctx->device));
  if (copy_h2d(ctx, ctx->u_cur, state_host, (size_t)(ctx->n_cells * NVARS) * sizeof(double))) return 1;
  ZFVM_CUDA(cudaStreamSynchronize(ctx->stream));
  return 0;
}

int zfvm_download_state(zfvm_ctx *ctx, double *state_host) {
  ZFVM_CUDA(cudaSetDevice(ctx->device));
  return copy_d2h(ctx, state_host, ctx->u_cur, (size_t)(ctx->n_cells * NVARS) * sizeof(double));
}

double *zfvm_state_device(zfvm_ctx *ctx) { return ctx->u_cur; }

int zfvm_upload_avars(zfvm_ctx *ctx, const double *avars_host) {
  if (ctx->n_avars <= 0) return fail("zfvm_upload_avars: the context was created with n_avars = 0");
  ZFVM_CUDA(cudaSetDevice(ctx->device));
  if (copy_h2d(ctx, ctx->a_cur, avars_host, (size_t)(ctx->n_cells * ctx->n_avars) * sizeof(double))) return 1;
  ZFVM_CUDA(cudaStreamSynchronize(ctx->stream));
  return 0;
}

int zfvm_download_avars(zfvm_ctx *ctx, double *avars_host) {
  if (ctx->n_avars <= 0) return fail("zfvm_download_avars: the context was created with n_avars = 0");
  ZFVM_CUDA(cudaSetDevice(ctx->device));
  return copy_d2h(ctx, avars_host, ctx->a_cur, (size_t)(ctx->n_cells * ctx->n_avars) * sizeof(double));
}

double *zfvm_avars_device(zfvm_ctx *ctx) { return ctx->a_cur; }

int zfvm_set_frozen_bc_av(zfvm_ctx *ctx, const double *steady_state_host, const double *steady_avars_host) {
  ++ctx->graph_epoch;  // captured steps are stale
  if (zfvm_set_frozen_bc(ctx, steady_state_host)) return 1;
  if (!steady_state_host || !steady_avars_host || ctx->n_avars <= 0) {
    ctx->frozen_a = nullptr;
    return 0;
  }
  if (!ctx->frozen_a_buf && dev_alloc(ctx, &ctx->frozen_a_buf, ctx->n_cells * ctx->n_avars)) return 1;  // reused by later calls
  ZFVM_CUDA(cudaMemcpy(ctx->frozen_a_buf, steady_avars_host, (size_t)(ctx->n_cells * ctx->n_avars) * sizeof(double),
                       cudaMemcpyHostToDevice));
  ctx->frozen_a = ctx->frozen_a_buf;
  return 0;
}

int zfvm_set_frozen_bc(zfvm_ctx *ctx, const double *steady_state_host) {
  ++ctx->graph_epoch;  // captured steps are stale
  ZFVM_CUDA(cudaSetDevice(ctx->device));
  if (!steady_state_host) {
    ctx->frozen = nullptr;
    return 0;
  }
  if (!ctx->frozen_buf && dev_alloc(ctx, &ctx->frozen_buf, ctx->n_cells * NVARS)) return 1;  // reused by later calls
  ZFVM_CUDA(cudaStreamSynchronize(ctx->stream));  // a step that still reads the previous steady state
  ZFVM_CUDA(cudaMemcpy(ctx->frozen_buf, steady_state_host, (size_t)(ctx->n_cells * NVARS) * sizeof(double),
                       cudaMemcpyHostToDevice));
  ctx->frozen = ctx->frozen_buf;
  return 0;
}

int zfvm_apply_frozen_bc(zfvm_ctx *ctx, double *state_dev) {
  if (!ctx->frozen) return 0;
  ZFVM_CUDA(cudaSetDevice(ctx->device));
  launch_frozen_bc(state_dev, ctx->frozen, ctx->ghost_index, ctx->n_ghost, ctx->stream);
  ctx->launches += 1;
  ZFVM_CUDA(cudaGetLastError());
  return 0;
}

// RungeKutta::compute_step (runge_kutta.cpp:87-112). Each stage's K3 also forms the next stage
// state u0 + dt * sum_j a[s+1][j] k_j (or the final sum with b) and applies the boundary condition,
// so there is no separate axpy pass over the state.
static int rk_step_impl(zfvm_ctx *ctx, double dt, bool reduce) {
  const int S = ctx->n_stages;
  if (reduce) {
    launch_reset_reduce(ctx->reduce_dev, ctx->stream);
    ctx->launches += 1;
  }
  for (int s = 0; s < S; ++s) {
    const double *in = (s == 0) ? ctx->u_cur : ctx->u_tmp;
    const double *coefs = (s + 1 < S) ? ctx->tab_a[s + 1] : ctx->tab_b;
    UpdateArgs A = base_update_args(ctx);
    // k_s is needed again iff a later sum uses it
    bool needed_later = false;
    for (int r = s + 2; r <= S; ++r) {
      const double *row = (r < S) ? ctx->tab_a[r] : ctx->tab_b;
      if (row[s] != 0.0) needed_later = true;
    }
    A.tendency = needed_later ? ctx->k[s] : nullptr;
    A.accumulate = 0;
    A.u_next = ctx->u_tmp;
    A.u_base = ctx->u_cur;
    A.n_prev = s;
    for (int j = 0; j < s; ++j) {
      A.k_prev[j] = ctx->k[j];
      A.coef_prev[j] = coefs[j];
    }
    A.coef_cur = coefs[s];
    A.dt = dt;
    A.frozen = ctx->frozen;
    A.reduce_out = (reduce && s + 1 == S) ? ctx->reduce_dev : nullptr;
    if (ctx->n_avars > 0) {  // the avars rows take the same Butcher sum (runge_kutta.cpp:122-143 sums AllVariables)
      UpdateArgs B = avars_update_args(ctx, A);
      B.tendency = needed_later ? ctx->ka[s] : nullptr;
      B.u_next = ctx->a_tmp;
      B.u_base = ctx->a_cur;
      for (int j = 0; j < s; ++j) B.k_prev[j] = ctx->ka[j];
      B.frozen = ctx->frozen ? ctx->frozen_a : nullptr;
      if (residual(ctx, in, A, (s == 0) ? ctx->a_cur : ctx->a_tmp, &B)) return 1;
    } else if (residual(ctx, in, A)) {
      return 1;
    }
  }
  std::swap(ctx->u_cur, ctx->u_tmp);
  std::swap(ctx->a_cur, ctx->a_tmp);
  return 0;
}


// One step replayed from a captured graph.  Returns 0 when the step has been queued, 1 on an error, -1 when graphs do
// not apply (the caller then launches the kernels one by one).
static int rk_step_graph(zfvm_ctx *ctx, double dt, bool reduce) {
  static const bool off = [] {
    const char *e = std::getenv("ZFVM_GRAPH");
    return (e != nullptr && e[0] == '0') || std::getenv("ZFVM_TILE_PROF") != nullptr;
  }();
  if (off || ctx->n_ranks > 1 || ctx->nccl_comm || ctx->prof_enabled || ctx->n_stages < 1) return -1;
  if (ctx->warm_epoch != ctx->graph_epoch) {  // the first step after a change runs plainly: every kernel it needs is
    ctx->warm_epoch = ctx->graph_epoch;       // loaded and configured before anything is captured
    return -1;
  }
  if (!ctx->dt_dev) {
    if (dev_alloc(ctx, &ctx->dt_dev, 1, true)) return 1;
    ZFVM_CUDA(cudaMallocHost((void **)&ctx->dt_host, sizeof(double)));
  }
  // the two state buffers alternate: one graph per buffer the step starts from
  zfvm_ctx::StepGraph *g = nullptr;
  for (int k = 0; k < 2; ++k) {
    zfvm_ctx::StepGraph &c = ctx->step_graph[reduce ? 1 : 0][k];
    if (c.exec && c.epoch == ctx->graph_epoch && c.u_cur == ctx->u_cur && c.u_tmp == ctx->u_tmp) g = &c;
  }
  if (!g) {
    zfvm_ctx::StepGraph *row = ctx->step_graph[reduce ? 1 : 0];
    zfvm_ctx::StepGraph *slot = (!row[0].exec || row[0].epoch != ctx->graph_epoch)   ? &row[0]
                                : (!row[1].exec || row[1].epoch != ctx->graph_epoch) ? &row[1]
                                                                                     : &row[ctx->graph_victim++ & 1];
    if (slot->exec) {
      cudaGraphExecDestroy(slot->exec);
      slot->exec = nullptr;
    }
    const std::int64_t launches_before = ctx->launches;
    double *u_cur = ctx->u_cur, *u_tmp = ctx->u_tmp, *a_cur = ctx->a_cur, *a_tmp = ctx->a_tmp;
    cudaGraph_t graph = nullptr;
    ZFVM_CUDA(cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal));
    cudaMemcpyAsync(ctx->dt_dev, ctx->dt_host, sizeof(double), cudaMemcpyHostToDevice, ctx->stream);
    ctx->capturing_dt = ctx->dt_dev;
    const int rc = rk_step_impl(ctx, dt, reduce);
    ctx->capturing_dt = nullptr;
    if (reduce) cudaMemcpyAsync(ctx->reduce_host, ctx->reduce_dev, sizeof(ReduceOut), cudaMemcpyDeviceToHost, ctx->stream);
    const cudaError_t end = cudaStreamEndCapture(ctx->stream, &graph);
    // the capture recorded the step, it did not run it: undo the bookkeeping of rk_step_impl
    ctx->u_cur = u_cur;
    ctx->u_tmp = u_tmp;
    ctx->a_cur = a_cur;
    ctx->a_tmp = a_tmp;
    slot->launches = ctx->launches - launches_before;
    ctx->launches = launches_before;
    if (rc != 0 || end != cudaSuccess || graph == nullptr) {
      if (graph) cudaGraphDestroy(graph);
      cudaGetLastError();
      return rc != 0 ? 1 : -1;
    }
    const cudaError_t inst = cudaGraphInstantiate(&slot->exec, graph, 0);
    cudaGraphDestroy(graph);
    if (inst != cudaSuccess) {
      slot->exec = nullptr;
      cudaGetLastError();
      return -1;
    }
    slot->u_cur = ctx->u_cur;
    slot->u_tmp = ctx->u_tmp;
    slot->epoch = ctx->graph_epoch;
    g = slot;
  }
  *ctx->dt_host = dt;
  ZFVM_CUDA(cudaGraphLaunch(g->exec, ctx->stream));
  ctx->launches += g->launches;
  std::swap(ctx->u_cur, ctx->u_tmp);
  std::swap(ctx->a_cur, ctx->a_tmp);
  return 0;
}

int zfvm_rk_step(zfvm_ctx *ctx, double /*t*/, double dt, double cfl_number, double *dt_next, int *not_plausible) {
  ZFVM_CUDA(cudaSetDevice(ctx->device));
  const bool reduce = (dt_next != nullptr) || (not_plausible != nullptr);
  const int graphed = rk_step_graph(ctx, dt, reduce);
  if (graphed > 0) return 1;
  if (graphed == 0) {
    if (reduce) {  // (the 16-byte copy of the reduction is part of the graph)
      ZFVM_CUDA(cudaStreamSynchronize(ctx->stream));
      if (dt_next) *dt_next = cfl_number * ctx->reduce_host->min_dx_over_ev;
      if (not_plausible) *not_plausible = ctx->reduce_host->not_plausible;
    }
    return 0;
  }
  if (rk_step_impl(ctx, dt, reduce)) return 1;
  if (reduce) {
    // every rank gets the same verdict: min of the CFL quotient, max of the plausibility flag (the reference aborts the
    // whole MPI job through LOG_ERR when one rank sees a bad state)
    if (ctx->n_ranks > 1 && ctx->nccl_comm && zfvm_allreduce_verdict_internal(ctx)) return 1;
    ZFVM_CUDA(cudaMemcpyAsync(ctx->reduce_host, ctx->reduce_dev, sizeof(ReduceOut), cudaMemcpyDeviceToHost, ctx->stream));
    ZFVM_CUDA(cudaStreamSynchronize(ctx->stream));
    if (dt_next) *dt_next = cfl_number * ctx->reduce_host->min_dx_over_ev;
    if (not_plausible) *not_plausible = ctx->reduce_host->not_plausible;
  }
  return 0;
}

// Update arguments of stage s of a step (the Butcher row that follows the stage; runge_kutta.cpp:122-143)
static UpdateArgs stage_update_args(zfvm_ctx *ctx, int s, double dt, bool reduce) {
  const int S = ctx->n_stages;
  const double *coefs = (s + 1 < S) ? ctx->tab_a[s + 1] : ctx->tab_b;
  UpdateArgs A = base_update_args(ctx);
  bool needed_later = false;  // k_s is needed again iff a later sum uses it
  for (int r = s + 2; r <= S; ++r) {
    const double *row = (r < S) ? ctx->tab_a[r] : ctx->tab_b;
    if (row[s] != 0.0) needed_later = true;
  }
  A.tendency = needed_later ? ctx->k[s] : nullptr;
  A.accumulate = 0;
  A.u_next = ctx->u_tmp;
  A.u_base = ctx->u_cur;
  A.n_prev = s;
  for (int j = 0; j < s; ++j) {
    A.k_prev[j] = ctx->k[j];
    A.coef_prev[j] = coefs[j];
  }
  A.coef_cur = coefs[s];
  A.dt = dt;
  A.frozen = ctx->frozen;
  A.reduce_out = (reduce && s + 1 == S) ? ctx->reduce_dev : nullptr;
  return A;
}

// K2 + K3 of a residual on a range of faces / cell blocks (the whole grid: face range [0, n_interior_edges), block 0)
static void flux_and_update(zfvm_ctx *ctx, const double *state, UpdateArgs A, std::int64_t face_begin, std::int64_t face_end) {
  launch_flux(ctx->plan, ctx->sc, nullptr, face_end - face_begin, ctx->stream, face_begin);
  A.flux_bc_state = ctx->params.flux_bc ? state : nullptr;
  A.flux_bc_kind = ctx->params.flux_bc;
  launch_update(ctx->plan, ctx->sc, A, ctx->stream);
  ctx->launches += 2;
}

// TimeIntegration::compute_step with host buffers, overlapped with its own copies (zfvm_ctx::HostPipe): u0 goes up in
// chunks while stage 0 reconstructs the tiles whose rows have landed; the last stage finishes chunk after chunk and every
// finished chunk goes down while the next one is computed.  Same kernels on the same data as the plain sequence: the
// result is bit-identical (tests/test_gpu_parity.py::test_compute_step_host_matches_resident).
//
// Multi-rank contexts (local numbering: owned rows first, then the halo rows): a tile whose stencils read a halo row
// becomes ready with a chunk that holds halo rows (tile_max_ref >= n_owned), so the chunks in front of the first such
// chunk `c_halo` gate interior tiles only.  Stage 0 reconstructs those while the rest goes up, then posts the halo
// exchange (HaloExchange::operator(), the same NCCL group as in residual()) and reconstructs what is left in one launch;
// the last stage exchanges first and then finishes the chunks that hold owned rows.  Every rank posts exactly one
// exchange per stage, as in the plain sequence, whether or not it takes this route.
static int rk_step_host_pipelined(zfvm_ctx *ctx, const double *u0_host, double *u1_host, double dt) {
  zfvm_ctx::HostPipe &H = ctx->pipe;
  const int C = H.n_chunks, S = ctx->n_stages;
  const std::int64_t EI = ctx->plan.n_interior_edges;
  const bool multi = ctx->n_ranks > 1 && ctx->nccl_comm != nullptr;
  const std::int64_t n_upd = (ctx->n_ranks > 1) ? ctx->n_owned : ctx->n_cells;  // rows the update kernel writes (base_update_args)
  int c_halo = C;  // first chunk that holds a halo row
  if (multi)
    for (int c = C - 1; c >= 0; --c)
      if (H.cell_begin[(size_t)c + 1] > ctx->n_owned) c_halo = c;
  auto exchange = [&](const double *state) -> int {
    if (zfvm_halo_post_internal(ctx, const_cast<double *>(state), nullptr)) return 1;
    return zfvm_halo_wait_internal(ctx);
  };
  ZFVM_CUDA(cudaStreamSynchronize(ctx->stream));  // whatever still reads or writes the resident state is done
  // ---- stage 0: upload chunk c, then the tiles chunk c completes ------------------------------------------------------
  for (int c = 0; c < C; ++c) {
    const std::int64_t r0 = H.cell_begin[(size_t)c] * NVARS, r1 = H.cell_begin[(size_t)c + 1] * NVARS;
    if (copy_h2d(ctx, ctx->u_cur + r0, u0_host + r0, (size_t)(r1 - r0) * sizeof(double), ctx->copy_stream)) return 1;
    ZFVM_CUDA(cudaEventRecord(H.ev_up[(size_t)c], ctx->copy_stream));
    ZFVM_CUDA(cudaStreamWaitEvent(ctx->stream, H.ev_up[(size_t)c], 0));
    const std::int64_t nt = H.up_off[(size_t)c + 1] - H.up_off[(size_t)c];
    if (nt > 0 && c < c_halo) {
      if (run_recon(ctx, ctx->u_cur, H.up_tiles + H.up_off[(size_t)c], nt))
        return fail("no reconstruction kernel is compiled for this scheme");
      ctx->launches += recon_group_launches(ctx);
    }
  }
  if (multi) {  // every row is up: exchange, then the tiles that waited for a chunk with halo rows
    if (exchange(ctx->u_cur)) return 1;
    const std::int64_t nt = H.up_off[(size_t)C] - H.up_off[(size_t)c_halo];
    if (nt > 0) {
      if (run_recon(ctx, ctx->u_cur, H.up_tiles + H.up_off[(size_t)c_halo], nt))
        return fail("no reconstruction kernel is compiled for this scheme");
      ctx->launches += recon_group_launches(ctx);
    }
  }
  auto finish_in_chunks = [&](int s, const double *in, double *out) -> int {
    // (K1 of the stage is done for s == 0 and S == 1; otherwise it runs chunk by chunk here)
    UpdateArgs A = stage_update_args(ctx, s, dt, false);
    A.u_next = out;
    if (multi) {
      if (s > 0 && exchange(in)) return 1;
      // the update kernel writes owned rows only: the halo rows of the result buffer are those of the stage's input,
      // as in the plain sequence (where the result is written into the buffer the exchange has just filled)
      const std::int64_t h0 = ctx->n_owned * NVARS, h1 = ctx->n_cells * NVARS;
      if (h1 > h0)
        ZFVM_CUDA(cudaMemcpyAsync(out + h0, in + h0, (size_t)(h1 - h0) * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
    }
    for (int c = 0; c < C; ++c) {
      if (H.cell_begin[(size_t)c] >= n_upd) {  // a chunk of halo rows: nothing to compute, it only travels
        ZFVM_CUDA(cudaEventRecord(H.ev_dn[(size_t)c], ctx->stream));
        continue;
      }
      const std::int64_t nt = H.dn_off[(size_t)c + 1] - H.dn_off[(size_t)c];
      if (s > 0 && nt > 0) {
        if (run_recon(ctx, in, H.dn_tiles + H.dn_off[(size_t)c], nt)) return fail("no reconstruction kernel is compiled for this scheme");
        ctx->launches += recon_group_launches(ctx);
      }
      UpdateArgs Ac = A;
      Ac.block_begin = H.cell_begin[(size_t)c] / 64;
      Ac.n_cells_update = std::min(H.cell_begin[(size_t)c + 1], n_upd);
      flux_and_update(ctx, in, Ac, H.face_begin[(size_t)c], H.face_begin[(size_t)c + 1]);
      ZFVM_CUDA(cudaEventRecord(H.ev_dn[(size_t)c], ctx->stream));
    }
    for (int c = 0; c < C; ++c) {  // (queued after all launches: a pageable destination blocks the host per chunk)
      const std::int64_t r0 = H.cell_begin[(size_t)c] * NVARS, r1 = H.cell_begin[(size_t)c + 1] * NVARS;
      ZFVM_CUDA(cudaStreamWaitEvent(ctx->copy_stream, H.ev_dn[(size_t)c], 0));
      if (copy_d2h(ctx, u1_host + r0, out + r0, (size_t)(r1 - r0) * sizeof(double), ctx->copy_stream, false)) return 1;
    }
    ZFVM_CUDA(cudaStreamSynchronize(ctx->copy_stream));
    return 0;
  };
  if (S == 1) {
    if (finish_in_chunks(0, ctx->u_cur, ctx->u_tmp)) return 1;
    std::swap(ctx->u_cur, ctx->u_tmp);
  } else {
    flux_and_update(ctx, ctx->u_cur, stage_update_args(ctx, 0, dt, false), 0, EI);
    for (int s = 1; s + 1 < S; ++s)
      if (residual(ctx, ctx->u_tmp, stage_update_args(ctx, s, dt, false))) return 1;
    // the last stage reads u_tmp while its finished chunks are written: they go to the third buffer
    if (finish_in_chunks(S - 1, ctx->u_tmp, H.u_out)) return 1;
    std::swap(ctx->u_cur, H.u_out);
  }
  ZFVM_CUDA(cudaStreamSynchronize(ctx->stream));
  ZFVM_CUDA(cudaGetLastError());
  return 0;
}

int zfvm_rk_step_host(zfvm_ctx *ctx, const double *u0_host, double *u1_host, double /*t*/, double dt) {
  ZFVM_CUDA(cudaSetDevice(ctx->device));
  if (host_pipeline_enabled(ctx) && ctx->n_stages >= 1) return rk_step_host_pipelined(ctx, u0_host, u1_host, dt);
  const size_t bytes = (size_t)(ctx->n_cells * NVARS) * sizeof(double);
  if (copy_h2d(ctx, ctx->u_cur, u0_host, bytes)) return 1;
  if (rk_step_impl(ctx, dt, false)) return 1;
  return copy_d2h(ctx, u1_host, ctx->u_cur, bytes);
}

int zfvm_rk_step_host_av(zfvm_ctx *ctx, const double *u0_host, const double *a0_host, double *u1_host, double *a1_host,
                         double /*t*/, double dt) {
  if (ctx->n_avars <= 0) return fail("zfvm_rk_step_host_av: the context was created with n_avars = 0");
  ZFVM_CUDA(cudaSetDevice(ctx->device));
  const size_t bytes = (size_t)(ctx->n_cells * NVARS) * sizeof(double);
  const size_t bytes_a = (size_t)(ctx->n_cells * ctx->n_avars) * sizeof(double);
  if (copy_h2d(ctx, ctx->u_cur, u0_host, bytes) || copy_h2d(ctx, ctx->a_cur, a0_host, bytes_a)) return 1;
  if (rk_step_impl(ctx, dt, false)) return 1;
  if (copy_d2h(ctx, a1_host, ctx->a_cur, bytes_a)) return 1;
  return copy_d2h(ctx, u1_host, ctx->u_cur, bytes);
}

int zfvm_cfl_dt(zfvm_ctx *ctx, const double *state_dev, double cfl_number, double *dt, int *not_plausible) {
  ZFVM_CUDA(cudaSetDevice(ctx->device));
  if (!state_dev) state_dev = ctx->u_cur;
  launch_reset_reduce(ctx->reduce_dev, ctx->stream);
  // owned + physical ghost cells; halo rows are refreshed by the next exchange and are left out
  // (deviation from local_cfl_condition_impl.hpp:25-40, which also scans stale halo rows)
  launch_cfl(state_dev, ctx->inradius, ctx->n_ranks > 1 ? ctx->n_owned : ctx->n_cells, ctx->sc.gamma, ctx->reduce_dev,
             ctx->stream);
  ctx->launches += 2;
  if (ctx->n_ranks > 1 && ctx->nccl_comm && zfvm_allreduce_verdict_internal(ctx)) return 1;
  ZFVM_CUDA(cudaMemcpyAsync(ctx->reduce_host, ctx->reduce_dev, sizeof(ReduceOut), cudaMemcpyDeviceToHost, ctx->stream));
  ZFVM_CUDA(cudaStreamSynchronize(ctx->stream));
  if (dt) *dt = cfl_number * ctx->reduce_host->min_dx_over_ev;
  if (not_plausible) *not_plausible = ctx->reduce_host->not_plausible;
  return 0;
}

int zfvm_profile_enable(zfvm_ctx *ctx, int enable) {
  ZFVM_CUDA(cudaSetDevice(ctx->device));
  ZFVM_CUDA(cudaStreamSynchronize(ctx->stream));
  for (auto &v : ctx->prof_events) {
    for (cudaEvent_t e : v) cudaEventDestroy(e);
    v.clear();
  }
  ctx->prof_enabled = enable != 0;
  return 0;
}

int zfvm_profile_read(zfvm_ctx *ctx, double ms[3], int64_t counts[3]) {
  ZFVM_CUDA(cudaSetDevice(ctx->device));
  ZFVM_CUDA(cudaStreamSynchronize(ctx->stream));
  for (int w = 0; w < 3; ++w) {
    ms[w] = 0.0;
    counts[w] = (int64_t)ctx->prof_events[w].size() / 2;
    for (size_t a = 0; a + 1 < ctx->prof_events[w].size(); a += 2) {
      float t = 0.f;
      ZFVM_CUDA(cudaEventElapsedTime(&t, ctx->prof_events[w][a], ctx->prof_events[w][a + 1]));
      ms[w] += t;
    }
  }
  unsigned long long tp[16];
  if (tile_prof_read(tp) && tp[8] > 0) {  // ZFVM_TILE_PROF=1: phase timers of warp 0 of the tile kernel
    static const char *names[8] = {"table_wait", "one_sided", "central", "table_issue", "hybridise", "geo_wait", "trace", "(ring waits)"};
    std::fprintf(stderr, "[zfvm tile prof] %llu tiles, cycles per tile:", tp[8]);
    for (int k = 0; k < 8; ++k) std::fprintf(stderr, " %s %.0f", names[k], (double)tp[k] / (double)tp[8]);
    std::fprintf(stderr, "\n");
  }
  return 0;
}

int zfvm_profile_read_tracers(zfvm_ctx *ctx, double *ms, int64_t *count) {
  ZFVM_CUDA(cudaSetDevice(ctx->device));
  ZFVM_CUDA(cudaStreamSynchronize(ctx->stream));
  *ms = 0.0;
  *count = (int64_t)ctx->prof_events[3].size() / 2;
  for (size_t a = 0; a + 1 < ctx->prof_events[3].size(); a += 2) {
    float t = 0.f;
    ZFVM_CUDA(cudaEventElapsedTime(&t, ctx->prof_events[3][a], ctx->prof_events[3][a + 1]));
    *ms += t;
  }
  return 0;
}

int zfvm_synchronize(zfvm_ctx *ctx) {
  ZFVM_CUDA(cudaSetDevice(ctx->device));
  ZFVM_CUDA(cudaStreamSynchronize(ctx->stream));
  return 0;
}

int zfvm_counters(zfvm_ctx *ctx, int64_t counters[4]) {
  ZFVM_CUDA(cudaSetDevice(ctx->device));
  int fails = 0;
  ZFVM_CUDA(cudaMemcpyAsync(&fails, ctx->eq_fail_dev, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  ZFVM_CUDA(cudaStreamSynchronize(ctx->stream));
  counters[0] = ctx->launches;
  counters[1] = fails;
  counters[2] = ctx->n_tiles_interior;
  counters[3] = ctx->n_tiles_exterior;
  return 0;
}

int zfvm_download_polynomials(zfvm_ctx *ctx, double *coeffs_host, double *scale_host, int *n_coef) {
  if (!ctx->plan.poly) return fail("zfvm_download_polynomials: create the context with keep_polynomials = 1");
  ZFVM_CUDA(cudaSetDevice(ctx->device));
  ZFVM_CUDA(cudaStreamSynchronize(ctx->stream));
  const std::int64_t n = ctx->n_cells;
  *n_coef = ctx->plan.n_poly_coef;
  ZFVM_CUDA(cudaMemcpy(coeffs_host, ctx->plan.poly, (size_t)(n * ctx->plan.n_poly_coef * NVARS) * sizeof(double),
                       cudaMemcpyDeviceToHost));
  ZFVM_CUDA(cudaMemcpy(scale_host, ctx->plan.poly_scale, (size_t)(n * NVARS) * sizeof(double), cudaMemcpyDeviceToHost));
  return 0;
}

int zfvm_download_work(zfvm_ctx *ctx, const char *name, double *host, int64_t max_count) {
  ZFVM_CUDA(cudaSetDevice(ctx->device));
  ZFVM_CUDA(cudaStreamSynchronize(ctx->stream));
  const std::string s(name);
  const double *src = nullptr;
  std::int64_t count = 0;
  if (s == "trace") {
    src = ctx->plan.trace;
    count = ctx->plan.n_interior_edges * 2 * ctx->sc.q_f * NVARS;
  } else if (s == "flux") {
    src = ctx->plan.flux;
    count = ctx->plan.n_interior_edges * NVARS;
  } else if (s == "source") {
    src = ctx->plan.source;
    count = ctx->n_cells * NVARS;
  } else if (s == "records") {  // the tile records as raw 8-byte words (tests: device-built weights vs host-built ones)
    src = reinterpret_cast<const double *>(ctx->plan.rec2 ? ctx->plan.rec2 : ctx->plan.rec);
    count = ctx->n_tiles * (ctx->plan.rec2 ? ctx->plan.rec2_bytes : ctx->plan.rec_bytes) / 8;
  } else {
    return fail("zfvm_download_work: unknown array");
  }
  if (count > max_count) return fail("zfvm_download_work: buffer too small");
  ZFVM_CUDA(cudaMemcpy(host, src, (size_t)count * sizeof(double), cudaMemcpyDeviceToHost));
  return 0;
}

}  // extern "C"
