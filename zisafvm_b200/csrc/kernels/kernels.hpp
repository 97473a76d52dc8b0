// Host-callable launchers of the residual kernels (implemented in the .cu files of this directory).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "../device/layout.hpp"

namespace zfvm {

constexpr int MAX_RK_STAGES = 6;  // Fehlberg has six (runge_kutta.cpp:190-207)

struct ReduceOut {
  double min_dx_over_ev;  // min over cells of inradius / (|v| + a)
  int not_plausible;      // some cell has rho <= 0, E <= 0 or a non-finite value
  int pad;
};

struct StagePtrs {
  const double *k[MAX_RK_STAGES];
  double coef[MAX_RK_STAGES];
};

struct UpdateArgs {
  std::int64_t n_cells_update;   // cells [64 * block_begin, n_cells_update) are updated
  std::int64_t block_begin;      // first block of 64 cells of the launch (0: from the first cell)
  int has_source;
  // residual output (RateOfChange::compute): may be null when only the fused update is wanted
  double *tendency;
  int accumulate;
  // fused Runge-Kutta stage update u_next = u_base + dt * (sum_s coef_prev[s] k_prev[s] + coef_cur * k_cur)
  double *u_next;
  const double *u_base;
  const double *k_prev[MAX_RK_STAGES];
  double coef_prev[MAX_RK_STAGES];
  int n_prev;
  double coef_cur;
  double dt;
  const double *dt_dev;    // non-null: the time step is read from device memory (steps replayed as CUDA graphs)
  const double *frozen;    // FrozenBC steady state (null: no boundary condition)
  // FluxBC (boundary/flux_bc.hpp:24-42): exterior faces of the cell, evaluated on `state`; null = NoFluxBC
  const double *flux_bc_state;
  int flux_bc_kind;        // 1 FluxBC, 2 EquilibriumFluxBC (boundary/equilibrium_flux_bc.hpp:37-63)
  int n_avars;             // launch_tracer_update only: row length of the avars arrays
  // CFL / plausibility reduction over the updated state
  ReduceOut *reduce_out;
  const double *inradius;
  double gamma;
};

constexpr int TILE_OFF_META = 16;
constexpr int TRACE_DUMP_BLOCKS = 4096;  // per-warp dump blocks behind the trace array (tile kernel write-out)

/// Generic description of the tile record for a scheme (used by the host when it builds the records).
struct TileRecLayout {
  int cap, rows, lidx_elem, off_list, off_lidx, off_wlo, off_whi, off_geo, geo_doubles;
  std::int64_t rec_bytes;
};

inline TileRecLayout tile_rec_layout(const SchemeConst &sc, int n_dims, int dof_hi, int cap) {
  TileRecLayout L{};
  const int ns = sc.n_stencils, F = n_dims + 1;
  L.cap = cap;
  L.rows = 0;
  for (int k = 0; k < ns; ++k) L.rows += sc.rows_max[k];
  L.lidx_elem = cap <= 256 ? 1 : 2;
  L.off_list = TILE_OFF_META + TILE * 8;
  L.off_lidx = (L.off_list + cap * 4 + 127) / 128 * 128;
  L.off_wlo = (L.off_lidx + L.rows * TILE * L.lidx_elem + 127) / 128 * 128;
  int lo = 0;
  for (int k = 1; k < ns; ++k) lo += sc.rows_max[k] * sc.ncoef[k] * TILE * 8;
  L.off_whi = L.off_wlo + lo;
  L.off_geo = L.off_whi + sc.rows_max[0] * sc.ncoef[0] * TILE * 8;
  L.geo_doubles = F * n_dims + n_dims + 1 + (dof_hi > 3 ? dof_hi - 3 : 0);
  const int geo_bytes = L.geo_doubles * TILE * 8 + F * TILE * 4 + TILE * 4;
  L.rec_bytes = L.off_geo + (geo_bytes + 127) / 128 * 128;
  return L;
}

/// Where the tracer reconstruction (tracers.cu) finds a cell's stencil members and weights: either record kind.
struct TracerRecView {
  int tile_record;             // 1: tile-kernel record (DevicePlan::rec2), 0: older record (DevicePlan::rec)
  int off_meta;                // u64[32]
  int off_list, off_lidx, lidx_elem;  // tile record: row list, local index rows [row][32] of u8 / u16
  int row0[MAX_STENCILS];      // tile record: first index row of stencil k
  int off_sidx[MAX_STENCILS];  // older record: i32[rows_max_k][32]
  int off_w[MAX_STENCILS];     // W_k f64[rows_max_k][ncoef_k][32]
};

/// P1 `lsq_weights_kernel` (precompute.cu): W_k = pinv(A_k) of the cells of tiles [tile_begin, tile_end), written into the
/// records whose headers (meta, member indices) are already on the device.
struct LsqWeightArgs {
  char *rec;                 // records of either kind
  std::int64_t rec_bytes;
  TracerRecView view;        // where the members and the weights of stencil k are
  int n_dims, n_stencils, n_moments;
  int ncoef[MAX_STENCILS];             // column count W_k is padded to in the record
  int max_order[MAX_STENCILS];         // the family's order of stencil k
  int rows_of_order[MAX_STENCILS][8];  // rows (members - 1) of stencil k when it achieves order o (stencil.cpp:168-175)
  const double *centers;     // [n_cells][3]   cell barycentres
  const double *length;      // [n_cells]      characteristic lengths
  const double *moments;     // [n_cells][n_moments] normalised moments (grid.cpp:1049-1098)
  double *scratch;           // set by the launcher
  std::int64_t tile_begin, tile_end;
};
/// 0 on success, 1 when no kernel covers the family (more than 34 columns), 2 on a CUDA error.  `*scratch` is a
/// device buffer the launcher grows on demand for stencils too large for shared memory; the caller frees it.
int launch_lsq_weights(const LsqWeightArgs &args, int n_sms, double **scratch, std::int64_t *scratch_bytes,
                       cudaStream_t stream);

/// Phase timers of the tile kernel (ZFVM_TILE_PROF=1): device buffer of 16 counters, or null when profiling is off.
unsigned long long *tile_prof_buffer();
/// Copies the counters to the host and clears them; returns false when profiling is off.
bool tile_prof_read(unsigned long long out[16]);

/// True if the tile kernel (recon_tile.cuh) is compiled for this scheme's dimension, degrees and stencil sizes.
bool recon_tile_compiled(const SchemeConst &sc, int deg_hi, int deg_lo);

/// Returns 0 on success, 1 if no kernel is compiled for this (n_dims, orders, n_stencils) combination.
int launch_recon(const DevicePlan &plan, const SchemeConst &sc, int deg_hi, int deg_lo, const double *state,
                 const std::int32_t *tile_list, std::int64_t n_tiles, cudaStream_t stream);

/// Generic reconstruction (recon_generic.cu): any stencil family up to MAX_STENCILS stencils, per-stencil orders.
bool recon_generic_supported(const SchemeConst &sc, int n_poly_coef);
int launch_recon_generic(const DevicePlan &plan, const SchemeConst &sc, const double *state,
                         const std::int32_t *tile_list, std::int64_t n_tiles, cudaStream_t stream);
int launch_tracer_recon_generic(const DevicePlan &P, const SchemeConst &sc, const double *avars,
                                const std::int32_t *tile_list, std::int64_t n_tiles, cudaStream_t stream);

/// faces face_list[0 .. n_faces) or, without a list, faces [face_begin, face_begin + n_faces)
void launch_flux(const DevicePlan &P, const SchemeConst &sc, const std::int32_t *face_list, std::int64_t n_faces,
                 cudaStream_t stream, std::int64_t face_begin = 0);
void launch_update(const DevicePlan &P, const SchemeConst &sc, const UpdateArgs &A, cudaStream_t stream);
/// Advected scalars: T1 scalar reconstruction + traces, T3 gather / RK update; the tracer face flux
/// is evaluated by launch_flux (K2) when the plan carries scalars (it needs the face's HLLC wave speeds).
int launch_tracer_recon(const DevicePlan &P, const SchemeConst &sc, const TracerRecView &view, int deg_hi, int deg_lo,
                        const double *avars, const std::int32_t *tile_list, std::int64_t n_tiles, cudaStream_t stream);
void launch_tracer_update(const DevicePlan &P, int n_dims, const UpdateArgs &A, cudaStream_t stream);
void launch_pack_rows_n(double *out, const double *state, const std::int32_t *index, std::int64_t n_rows, int row_len,
                        cudaStream_t stream);
void launch_frozen_bc_n(double *u, const double *frozen, const std::int32_t *ghost_index, std::int64_t n_ghost,
                        int row_len, cudaStream_t stream);
void launch_cfl(const double *u, const double *inradius, std::int64_t n, double gamma, ReduceOut *out,
                cudaStream_t stream);
void launch_reset_reduce(ReduceOut *out, cudaStream_t stream);
void launch_frozen_bc(double *u, const double *frozen, const std::int32_t *ghost_index, std::int64_t n_ghost,
                      cudaStream_t stream);
void launch_pack_rows(double *out, const double *state, const std::int32_t *index, std::int64_t n_rows,
                      cudaStream_t stream);
void launch_axpy_stage(double *u_next, const double *u_base, const StagePtrs &K, int n_stages, double dt,
                       std::int64_t n, cudaStream_t stream);

}  // namespace zfvm
