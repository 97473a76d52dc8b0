// Local isentropic equilibrium of a cell (shared by the reconstruction kernel and EquilibriumFluxBC in K3).
#pragma once
#include "common.cuh"

namespace zfvm {

/// Local isentropic equilibrium of one cell (local_equilibrium_impl.hpp:34-94,
/// isentropic_equilibrium.hpp:27-63). All potentials come from precomputed tables.
struct LocalEq {
  double h_ref, K, phi_ref;
  bool found;
  double c1 = 0.0, inv_gm1 = 0.0;  // (gamma-1) / (gamma K), 1 / (gamma-1): set by prepare()
  ZFVM_DEVICE void prepare(double gamma) {
    c1 = (gamma - 1.0) / (gamma * K);
    inv_gm1 = 1.0 / (gamma - 1.0);
  }
  template <int POWN = 0>
  ZFVM_DEVICE void at(double phi, const SchemeConst &sc, double &rho, double &E, double &p) const {
    if (!found) {
      rho = 0.0;
      E = 0.0;
      p = 0.0;
      return;
    }
    isentropic_state_c<POWN>(h_ref + phi_ref - phi, K, c1, sc, inv_gm1, rho, E, p);
  }
};

/// Cell averages of N isentropic equilibria (h_n, K_n, same reference potential) over a cell whose Gauss-point
/// potentials are `phi` (AoS row), accumulated point 0 first (quadrature.hpp:43-48).  The kernel is bound by the
/// latency of the dependent rsqrt / Newton / multiply chain of one state evaluation, so independent evaluations are
/// issued together: the N equilibria of a finite-difference Jacobian and CH Gauss points at a time.  Every value is
/// formed by the same operations as in the one-at-a-time form (bit-identical results).
template <int N, int CH, int POWN = 0>
ZFVM_DEVICE void eq_cell_average_multi(const double h_ref[N], const double K[N], double phi_ref,
                                       const double *__restrict__ phi, const SchemeConst &sc, double rho_bar[N],
                                       double E_bar[N]) {
  const double gamma = sc.gamma, inv_gm1 = 1.0 / (gamma - 1.0);
  double c1[N], hp[N];
#pragma unroll
  for (int n = 0; n < N; ++n) {
    c1[n] = (gamma - 1.0) / (gamma * K[n]);
    hp[n] = h_ref[n] + phi_ref;
    rho_bar[n] = 0.0;
    E_bar[n] = 0.0;
  }
  int q = 0;
  for (; q + CH <= sc.q_c; q += CH) {
    double r[CH][N], E[CH][N], ph[CH];
#pragma unroll
    for (int j = 0; j < CH; ++j) ph[j] = phi[q + j];
#pragma unroll
    for (int j = 0; j < CH; ++j)
#pragma unroll
      for (int n = 0; n < N; ++n) {
        double p;
        isentropic_state_c<POWN>(hp[n] - ph[j], K[n], c1[n], sc, inv_gm1, r[j][n], E[j][n], p);
      }
#pragma unroll
    for (int j = 0; j < CH; ++j)
#pragma unroll
      for (int n = 0; n < N; ++n) {
        rho_bar[n] = fma(sc.cell_w[q + j], r[j][n], rho_bar[n]);
        E_bar[n] = fma(sc.cell_w[q + j], E[j][n], E_bar[n]);
      }
  }
  for (; q < sc.q_c; ++q) {
    const double ph = phi[q];
#pragma unroll
    for (int n = 0; n < N; ++n) {
      double r, E, p;
      isentropic_state_c<POWN>(hp[n] - ph, K[n], c1[n], sc, inv_gm1, r, E, p);
      rho_bar[n] = fma(sc.cell_w[q], r, rho_bar[n]);
      E_bar[n] = fma(sc.cell_w[q], E, E_bar[n]);
    }
  }
}

/// cell average of the equilibrium over a cell whose Gauss-point potentials are `phi` (AoS row)
template <int POWN = 0>
ZFVM_DEVICE void eq_cell_average(const LocalEq &eq, const double *__restrict__ phi, const SchemeConst &sc,
                                 double &rho_bar, double &E_bar) {
  if (!eq.found) {
    rho_bar = 0.0;
    E_bar = 0.0;
    return;
  }
  const double hp = eq.h_ref + eq.phi_ref;
  constexpr int CH = 4;
  rho_bar = 0.0;
  E_bar = 0.0;
  int q = 0;
  const bool vec = (sc.q_c & 1) == 0;  // rows of an even number of doubles are 16-byte aligned: 128-bit loads
  for (; q + CH <= sc.q_c; q += CH) {
    double r[CH], E[CH], ph[CH];
    if (vec) {
#pragma unroll
      for (int j = 0; j < CH; j += 2) {
        const double2 v = *reinterpret_cast<const double2 *>(phi + q + j);
        ph[j] = v.x;
        ph[j + 1] = v.y;
      }
    } else {
#pragma unroll
      for (int j = 0; j < CH; ++j) ph[j] = phi[q + j];
    }
#pragma unroll
    for (int j = 0; j < CH; ++j) {
      double p;
      isentropic_state_c<POWN>(hp - ph[j], eq.K, eq.c1, sc, eq.inv_gm1, r[j], E[j], p);
    }
#pragma unroll
    for (int j = 0; j < CH; ++j) {
      rho_bar = fma(sc.cell_w[q + j], r[j], rho_bar);
      E_bar = fma(sc.cell_w[q + j], E[j], E_bar);
    }
  }
  for (; q < sc.q_c; ++q) {
    double r, E, p;
    isentropic_state_c<POWN>(hp - phi[q], eq.K, eq.c1, sc, eq.inv_gm1, r, E, p);
    rho_bar = fma(sc.cell_w[q], r, rho_bar);
    E_bar = fma(sc.cell_w[q], E, E_bar);
  }
}

/// quasi_newton (quasi_newton.hpp:12-50) on f(theta) = rhoE_bar - avg_cell rhoE_eq(theta) with
/// the central-difference Jacobian of local_equilibrium_impl.hpp:64-86.
template <int POWN = 0>
ZFVM_DEVICE LocalEq solve_local_equilibrium(double rho_bar, double E_bar, const double *__restrict__ phi_own,
                                            const SchemeConst &sc) {
  const double gamma = sc.gamma;
  LocalEq eq;
  eq.phi_ref = phi_own[0];  // x_ref = first cell Gauss point
  eq.found = true;
  const double p0 = E_bar * (gamma - 1.0);
  const double h0 = gamma / (gamma - 1.0) * p0 / rho_bar;
  const double K0 = p0 / ((gamma == 2.0) ? rho_bar * rho_bar : pow(rho_bar, gamma));
  const double atol_h = 1e-13 * h0, atol_K = 1e-13 * K0;

  auto f = [&](double h, double K, double &f0, double &f1) {
    LocalEq t{h, K, eq.phi_ref, true};
    t.prepare(gamma);
    double rb, Eb;
    eq_cell_average<POWN>(t, phi_own, sc, rb, Eb);
    f0 = rho_bar - rb;
    f1 = E_bar - Eb;
  };

  double h = h0, K = K0, f0, f1;
  f(h, K, f0, f1);
  double dx0 = 2.0 * atol_h + 1.0, dx1 = 2.0 * atol_K + 1.0;
  double a0 = 0, a1 = 0, b0 = 0, b1 = 0;  // previous two steps (rolling_convergence_rate.hpp)
  int iter = 0;
  bool ok = true;
  while (!(fabs(dx0) <= atol_h && fabs(dx1) <= atol_K) && iter < 20) {
    double eps_h = 1e-6 * fabs(h), eps_K = 1e-6 * fabs(K);
    // the four states of the central differences, evaluated together (independent dependency chains)
    const double hs[4] = {h + 0.5 * eps_h, h - 0.5 * eps_h, h, h};
    const double Ks[4] = {K, K, K + 0.5 * eps_K, K - 0.5 * eps_K};
    double rb4[4], Eb4[4];
    eq_cell_average_multi<4, 2, POWN>(hs, Ks, eq.phi_ref, phi_own, sc, rb4, Eb4);
    const double d00 = ((rho_bar - rb4[0]) - (rho_bar - rb4[1])) / eps_h;  // df0 = d f / d h
    const double d01 = ((E_bar - Eb4[0]) - (E_bar - Eb4[1])) / eps_h;
    const double d10 = ((rho_bar - rb4[2]) - (rho_bar - rb4[3])) / eps_K;  // df1 = d f / d K
    const double d11 = ((E_bar - Eb4[2]) - (E_bar - Eb4[3])) / eps_K;
    const double inv_det = 1.0 / (d00 * d11 - d01 * d10);
    dx0 = inv_det * (d11 * f0 - d10 * f1);
    dx1 = inv_det * (-d01 * f0 + d00 * f1);
    h -= dx0;
    K -= dx1;
    f(h, K, f0, f1);
    if (iter >= 4) {
      bool conv = (dx0 <= atol_h && dx1 <= atol_K);
      if (!conv) {
        double r0 = log(fabs(dx0) / fabs(b0)) / log(fabs(b0) / fabs(a0));
        double r1 = log(fabs(dx1) / fabs(b1)) / log(fabs(b1) / fabs(a1));
        conv = (r0 >= 0.0 && r1 >= 0.0);
      }
      if (!conv) {
        ok = false;
        break;
      }
    }
    a0 = b0;
    a1 = b1;
    b0 = dx0;
    b1 = dx1;
    ++iter;
  }
  if (ok && iter == 20 && !(fabs(dx0) <= 1000.0 * atol_h && fabs(dx1) <= 1000.0 * atol_K)) ok = false;
  eq.h_ref = ok ? h : h0;
  eq.K = ok ? K : K0;
  eq.found = ok;
  eq.prepare(gamma);
  return eq;
}


/// E0 (steps_per_recompute != 1 only).  One thread per cell: recompute_equilibrium's verdict
/// (local_reconstruction.hpp:87-100) -- every steps_per_recompute-th evaluation, or when (rho, E_int) has moved away from
/// the cached equilibrium average by recompute_threshold in units of the cached scale -- and the bookkeeping of
/// compute_equilibrium / compute (:69-85, :102-120): the scale's (rho, E_int), steps_since_recompute.
template <int UNUSED = 0>  // (a template only so that the header can be included by several translation units)
__global__ void __launch_bounds__(256) eq_decide_kernel(const DevicePlan P, const __grid_constant__ SchemeConst sc,
                                                        const double *__restrict__ state,
                                                        const std::int32_t *__restrict__ tile_list, std::int64_t n_tiles) {
  const std::int64_t t = (std::int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_tiles * TILE) return;
  const std::int64_t tw = t / TILE;
  const int lane = (int)(t - tw * TILE);
  const std::int64_t tile = tile_list ? (std::int64_t)tile_list[tw] : tw;
  const std::int64_t i = tile * TILE + lane;
  if (i >= P.n_cells) return;
  const double *u = state + i * NVARS;
  const double rho = u[0];
  const double eint = u[4] - 0.5 * (u[1] * u[1] + u[2] * u[2] + u[3] * u[3]) / rho;
  int steps = P.eq_steps[i];
  bool recompute = (steps % sc.steps_per_recompute) == 0;
  if (!recompute) {
    double rho_c = 0.0, E_c = 0.0;  // rhoEbar_cache(0): the cell's own equilibrium average (NoEquilibrium: zero)
    if (sc.well_balanced) {
      const double *e0 = P.eq_avg + ((tile * P.eq_rows + (P.eq_rows - 1)) * 2) * TILE + lane;
      rho_c = e0[0];
      E_c = e0[TILE];
    }
    const double s0 = sc.scaling == SCALING_EULER ? P.scale_state[2 * i] : 1.0;
    const double s4 = sc.scaling == SCALING_EULER ? P.scale_state[2 * i + 1] : 1.0;
    const double d0 = (rho - rho_c) / s0, d1 = (eint - E_c) / s4;
    recompute = sqrt(d0 * d0 + d1 * d1) >= sc.recompute_threshold;
  }
  if (recompute) {
    P.scale_state[2 * i] = rho;
    P.scale_state[2 * i + 1] = eint;
    steps = 0;
  }
  P.eq_flag[i] = recompute ? 1 : 0;
  P.eq_steps[i] = steps + 1;
}

/// E1.  One thread per cell: the local equilibrium of the cell's average state (LocalEquilibrium::solve,
/// local_equilibrium_impl.hpp:34-94) -> eq_par[cell] = (h_ref, K, phi_ref, found).  Kept out of the reconstruction
/// kernel: the Newton iteration needs few registers and no stencil data, so it runs at full occupancy here.
template <int POWN>
__global__ void __launch_bounds__(256) eq_solve_kernel(const DevicePlan P, const __grid_constant__ SchemeConst sc,
                                                       const double *__restrict__ state,
                                                       const std::int32_t *__restrict__ tile_list, std::int64_t n_tiles) {
  const std::int64_t t = (std::int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_tiles * TILE) return;
  const std::int64_t tw = t / TILE;
  const std::int64_t i = (tile_list ? (std::int64_t)tile_list[tw] : tw) * TILE + (t - tw * TILE);
  if (i >= P.n_cells) return;
  if (P.eq_flag != nullptr && !P.eq_flag[i]) return;  // cached equilibrium kept (recompute_equilibrium)
  const double *u = state + i * NVARS;
  const double rho = u[0];
  const double eint = u[4] - 0.5 * (u[1] * u[1] + u[2] * u[2] + u[3] * u[3]) / rho;
  const LocalEq eq = solve_local_equilibrium<POWN>(rho, eint, P.phi_cqp + i * sc.q_c, sc);
  if (!eq.found) atomicAdd(P.eq_fail, 1);
  double *out = P.eq_par + i * 4;
  out[0] = eq.h_ref;
  out[1] = eq.K;
  out[2] = eq.phi_ref;
  out[3] = eq.found ? 1.0 : 0.0;
}

/// E2.  One thread per (cell, stencil row): the cell average of cell i's equilibrium over stencil member g
/// (LocalEquilibrium::extrapolate(cell), local_reconstruction.hpp:109-113) -> eq_avg[tile][row][2][lane].  These are
/// the bulk of the equilibrium evaluations (rows x q_c per cell); one per thread, they fill the FP64 pipe instead of
/// serialising inside the register-bound reconstruction kernel.  A warp owns one row of a tile: the member index and
/// the outputs are contiguous across lanes, the member's potentials are a gathered 8 q_c-byte row.
template <int POWN>
__global__ void __launch_bounds__(256) eq_member_kernel(const DevicePlan P, const __grid_constant__ SchemeConst sc,
                                                        const std::int32_t *__restrict__ tile_list, std::int64_t n_tiles) {
  const int lane = threadIdx.x & 31;
  const std::int64_t w = (std::int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (w >= n_tiles * P.eq_rows) return;
  const std::int64_t tw = w / P.eq_rows;
  const int row = (int)(w - tw * P.eq_rows);
  const std::int64_t tile = tile_list ? (std::int64_t)tile_list[tw] : tw;
  int k = 0;
#pragma unroll
  for (int kk = 1; kk < MAX_STENCILS; ++kk)
    if (kk < sc.n_stencils && row >= P.eq_row0[kk]) k = kk;
  const int j = row - P.eq_row0[k];
  const std::int64_t cell = tile * TILE + lane;
  const bool active = cell < P.n_cells;
  const std::uint64_t meta = active ? P.meta_of(tile)[lane] : 0ull;
  const int rows = (int)((meta >> (8 * k)) & 0xFF);
  double rb = 0.0, Eb = 0.0;
  if (j < rows) {
    const std::int64_t g = P.sidx_of(tile, k)[(std::int64_t)j * TILE + lane];
    const double *par = P.eq_par + cell * 4;
    LocalEq eq{par[0], par[1], par[2], par[3] != 0.0};
    eq.prepare(sc.gamma);
    eq_cell_average<POWN>(eq, P.phi_cqp + g * sc.q_c, sc, rb, Eb);
  }
  double *out = P.eq_avg + ((tile * P.eq_rows + row) * 2) * TILE + lane;
  out[0] = rb;
  out[TILE] = Eb;
}


/// E2 for tile records (recon_tile.cuh): rows in lidx order (the one-sided stencils' rows, then the central stencil's),
/// members through the tile's row list, and one more row for the cell itself.  Padded rows of ragged stencils point
/// at the cell itself and get the cell's own average, which makes their rhs vanish like in the plain scheme.
template <int POWN>
__global__ void __launch_bounds__(256) eq_member_tile_kernel(const DevicePlan P, const __grid_constant__ SchemeConst sc,
                                                             const std::int32_t *__restrict__ tile_list,
                                                             std::int64_t n_tiles) {
  const int lane = threadIdx.x & 31;
  const std::int64_t w = (std::int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (w >= n_tiles * P.eq_rows) return;
  const std::int64_t tw = w / P.eq_rows;
  const int row = (int)(w - tw * P.eq_rows);
  const std::int64_t tile = tile_list ? (std::int64_t)tile_list[tw] : tw;
  const std::int64_t cell = tile * TILE + lane;
  const std::int64_t ci = cell < P.n_cells ? cell : P.n_cells - 1;
  const char *rec = P.rec2 + tile * P.rec2_bytes;
  std::int64_t g = ci;
  if (row < P.eq_rows - 1) {
    const char *lrow = rec + P.rec2_off_lidx + (std::size_t)row * TILE * P.rec2_lidx_elem;
    const int li = (P.rec2_lidx_elem == 1) ? (int)reinterpret_cast<const std::uint8_t *>(lrow)[lane]
                                           : (int)reinterpret_cast<const std::uint16_t *>(lrow)[lane];
    g = reinterpret_cast<const std::int32_t *>(rec + P.rec2_off_list)[li];
  }
  if (P.eq_flag != nullptr && !P.eq_flag[ci]) return;  // cached averages kept
  const double *par = P.eq_par + ci * 4;
  LocalEq eq{par[0], par[1], par[2], par[3] != 0.0};
  eq.prepare(sc.gamma);
  double rb, Eb;
  eq_cell_average<POWN>(eq, P.phi_cqp + g * sc.q_c, sc, rb, Eb);
  double *out = P.eq_avg + ((tile * P.eq_rows + row) * 2) * TILE + lane;
  out[0] = rb;
  out[TILE] = Eb;
}

/// E2 for tile records, shared-memory form: one CTA per tile.  The potentials of the tile's distinct stencil members
/// (the record's row list, 125-250 cells) are staged in shared memory once -- contiguous 8 q_c-byte rows -- and every
/// (cell, row) pair then reads its member's row through the record's local index: no gathered global loads (which
/// bound the one-warp-per-row form at 84 % L1 throughput), the FP64 pipe is what is left.
template <int POWN>
__global__ void __launch_bounds__(256) eq_member_tile_smem_kernel(const DevicePlan P, const __grid_constant__ SchemeConst sc,
                                                                  const std::int32_t *__restrict__ tile_list,
                                                                  std::int64_t n_tiles) {
  extern __shared__ __align__(16) double eq_phi_s[];  // [n_list][q_c]
  const std::int64_t tile = tile_list ? (std::int64_t)tile_list[blockIdx.x] : (std::int64_t)blockIdx.x;
  const char *rec = P.rec2 + tile * P.rec2_bytes;
  const int n_list = *reinterpret_cast<const int *>(rec);
  const std::int32_t *list = reinterpret_cast<const std::int32_t *>(rec + P.rec2_off_list);
  const int q_c = sc.q_c;
  if (P.eq_flag != nullptr) {  // nothing to do when every cell of the tile keeps its cached averages
    const std::int64_t c = tile * TILE + (threadIdx.x & 31);
    const int mine = (threadIdx.x < TILE && c < P.n_cells) ? (int)P.eq_flag[c] : 0;
    if (!__syncthreads_or(mine)) return;
  }
  for (int idx = threadIdx.x; idx < n_list * q_c; idx += blockDim.x) {
    const int row = idx / q_c;
    eq_phi_s[idx] = P.phi_cqp[(std::int64_t)list[row] * q_c + (idx - row * q_c)];
  }
  __syncthreads();
  for (int pair = threadIdx.x; pair < P.eq_rows * TILE; pair += blockDim.x) {
    const int row = pair / TILE, lane = pair - row * TILE;
    const std::int64_t cell = tile * TILE + lane;
    const std::int64_t ci = cell < P.n_cells ? cell : P.n_cells - 1;
    int li = lane;  // the last row: the cell itself (list[0..31] are the tile's own cells)
    if (row < P.eq_rows - 1) {
      const char *lrow = rec + P.rec2_off_lidx + (std::size_t)row * TILE * P.rec2_lidx_elem;
      li = (P.rec2_lidx_elem == 1) ? (int)reinterpret_cast<const std::uint8_t *>(lrow)[lane]
                                   : (int)reinterpret_cast<const std::uint16_t *>(lrow)[lane];
    }
    if (P.eq_flag != nullptr && !P.eq_flag[ci]) continue;  // cached averages kept
    const double *par = P.eq_par + ci * 4;
    LocalEq eq{par[0], par[1], par[2], par[3] != 0.0};
    eq.prepare(sc.gamma);
    double rb, Eb;
    eq_cell_average<POWN>(eq, eq_phi_s + li * q_c, sc, rb, Eb);
    double *out = P.eq_avg + ((tile * P.eq_rows + row) * 2) * TILE + lane;
    out[0] = rb;
    out[TILE] = Eb;
  }
}

/// E3 (tile records).  One thread per (cell, face): the cell's equilibrium (rho, E) at the face's Gauss points
/// (LocalReconstruction::background, local_reconstruction.hpp:157-163; the FewPointsCache entries of the face points)
/// -> eq_bg[e][side][q][2], laid out like the trace array; the face-flux kernel adds it to the traces it reads.
template <int POWN>
__global__ void __launch_bounds__(256) eq_face_kernel(const DevicePlan P, const __grid_constant__ SchemeConst sc,
                                                      const std::int32_t *__restrict__ tile_list, std::int64_t n_tiles) {
  const int F = sc.n_dims + 1;
  const int lane = threadIdx.x & 31;
  const std::int64_t w = (std::int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (w >= n_tiles * F) return;
  const std::int64_t tw = w / F;
  const int k = (int)(w - tw * F);
  const std::int64_t tile = tile_list ? (std::int64_t)tile_list[tw] : tw;
  const std::int64_t cell = tile * TILE + lane;
  const std::int64_t ci = cell < P.n_cells ? cell : P.n_cells - 1;
  const std::uint32_t fref = cell < P.n_cells ? P.face_ref[(tile * F + k) * TILE + lane] : 0u;
  if (!(fref & FREF_TRACE)) return;
  if (P.eq_flag != nullptr && !P.eq_flag[ci]) return;  // cached point values kept (FewPointsCache not updated)
  const std::int64_t e = fref & FREF_EDGE_MASK;
  const int side = (fref & FREF_SIDE) ? 1 : 0;
  const double *par = P.eq_par + ci * 4;
  LocalEq eq{par[0], par[1], par[2], par[3] != 0.0};
  eq.prepare(sc.gamma);
  for (int q = 0; q < sc.q_f; ++q) {
    double r, E, p;
    eq.template at<POWN>(P.phi_fqp[e * sc.q_f + q], sc, r, E, p);
    *reinterpret_cast<double2 *>(P.eq_bg + ((e * 2 + side) * sc.q_f + q) * 2) = make_double2(r, E);
  }
}

}  // namespace zfvm
