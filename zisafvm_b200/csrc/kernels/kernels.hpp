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
  std::int64_t n_cells_update;
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
  const double *frozen;    // FrozenBC steady state (null: no boundary condition)
  // CFL / plausibility reduction over the updated state
  ReduceOut *reduce_out;
  const double *inradius;
  double gamma;
};

/// Returns 0 on success, 1 if no kernel is compiled for this (n_dims, orders, n_stencils) combination.
int launch_recon(const DevicePlan &plan, const SchemeConst &sc, int deg_hi, int deg_lo, const double *state,
                 const std::int32_t *tile_list, std::int64_t n_tiles, cudaStream_t stream);

void launch_flux(const DevicePlan &P, const SchemeConst &sc, const std::int32_t *face_list, std::int64_t n_faces,
                 cudaStream_t stream);
void launch_update(const DevicePlan &P, int n_dims, const UpdateArgs &A, cudaStream_t stream);
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
