// K1, generic form: any stencil family the reference can be configured with (numerical_experiment.cpp:150 reads
// orders / biases / overfit factors from JSON): any number of stencils up to MAX_STENCILS, several central stencils,
// every stencil with its own order (first-order families included), any stencil sizes.  Examples from the reference's
// own tests: {{1},{c}}, {{3},{b}}, {{4,2,2,2,2,2},{c,c,b,b,b,b}}, {{4,3,3,3,3,3},..} (test/.../weno_ao.cpp:47-55,
// cweno_ao.cpp:144-160).  The parameter sets the reference's experiments use (one central stencil + n_dims + 1
// one-sided ones of order 2) never come here: they run the tile kernel (recon_tile.cuh).
//
// Same arithmetic and reference lines as recon.cuh; what differs is that stencil count, per-stencil coefficient
// counts and the polynomial degree are run-time values, so the stencil polynomials live in local memory.  One thread
// owns one cell, one warp one tile; records are the `meta | sidx_k | W_k` form (device/layout.hpp).
#include "recon.cuh"
#include "kernels.hpp"

namespace zfvm {

namespace {

template <int ND>
struct GenericLimits {
  static constexpr int DEG = (ND == 2) ? 4 : 3;  // LSQ matrices exist up to order 5 in 2D, 4 in 3D (lsq_solver.cpp:288,399)
  static constexpr int D = dof_of(DEG, ND);
};

/// coefficient count (without the constant) -> polynomial degree
template <int ND>
ZFVM_DEVICE int degree_of_ncoef(int nc) {
  int deg = 0;
  while (dof_of(deg, ND) - 1 < nc) ++deg;
  return deg;
}

template <int ND, int VARIANT, int POWN>
__global__ void __launch_bounds__(128) recon_generic_kernel(const __grid_constant__ ReconArgs args,
                                                            const __grid_constant__ SchemeConst sc) {
  constexpr int F = ND + 1;
  constexpr int DM = GenericLimits<ND>::D;  // storage bound; the scheme's own dof is P.n_poly_coef
  constexpr int CM = DM - 1;
  constexpr bool WB = (VARIANT == RV_WELL_BALANCED);
  constexpr bool GRAV = (VARIANT != RV_PLAIN);
  const DevicePlan &P = args.plan;
  const int NS = sc.n_stencils;
  const int D = P.n_poly_coef;

  const int lane = threadIdx.x & 31;
  const std::int64_t w = (std::int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (w >= args.n_tiles_launch) return;
  const std::int64_t tile = args.tile_list ? (std::int64_t)args.tile_list[w] : w;
  const std::int64_t cell = tile * TILE + lane;
  const bool active = cell < P.n_cells;
  const std::int64_t ci = active ? cell : P.n_cells - 1;

  const std::uint64_t meta = active ? P.meta_of(tile)[lane] : 0ull;
  const int kh = (int)((meta >> 56) & 0xF);
  const bool single = ((meta >> 60) & 1) != 0;

  double u0[NVARS];
#pragma unroll
  for (int v = 0; v < NVARS; ++v) u0[v] = args.state[ci * NVARS + v];
  const double ekin0 = 0.5 * (u0[1] * u0[1] + u0[2] * u0[2] + u0[3] * u0[3]) / u0[0];
  const double eint0 = u0[4] - ekin0;
  double scale[NVARS];
  if (sc.scaling == SCALING_EULER) {  // characteristic_scale.hpp:24-33
    const double p = eint0 * (sc.gamma - 1.0);
    const double cs = sqrt(sc.gamma * p / u0[0]);
    scale[0] = u0[0];
    scale[1] = scale[2] = scale[3] = cs;
    scale[4] = eint0;
  } else {
#pragma unroll
    for (int v = 0; v < NVARS; ++v) scale[v] = 1.0;
  }
  double inv_scale[NVARS];
  inv_scale[0] = 1.0 / scale[0];
  inv_scale[1] = inv_scale[2] = inv_scale[3] = 1.0 / scale[1];
  inv_scale[4] = 1.0 / scale[4];

  LocalEq eq{0.0, 1.0, 0.0, false};
  double eq0_rho = 0.0, eq0_E = 0.0;
  if (WB) {
    const double *par = P.eq_par + ci * 4;
    eq = LocalEq{par[0], par[1], par[2], par[3] != 0.0};
    eq.prepare(sc.gamma);
    eq_cell_average<POWN>(eq, P.phi_cqp + ci * sc.q_c, sc, eq0_rho, eq0_E);
  }
  double q0s[NVARS];
  q0s[0] = (u0[0] - eq0_rho) * inv_scale[0];
  q0s[1] = u0[1] * inv_scale[1];
  q0s[2] = u0[2] * inv_scale[2];
  q0s[3] = u0[3] * inv_scale[3];
  q0s[4] = (u0[4] - eq0_E) * inv_scale[4];

  // ---- stencil polynomials: coef = W_k * rhs (hybrid_weno.cpp:72-92) -----------------------------------
  double pk[MAX_STENCILS][CM][NVARS];  // local memory: run-time indices
  for (int k = 0; k < NS; ++k)
    for (int c = 0; c < CM; ++c)
#pragma unroll
      for (int v = 0; v < NVARS; ++v) pk[k][c][v] = 0.0;
  for (int k = 0; k < NS; ++k) {
    const int NC = sc.ncoef[k];
    const int rows = (int)((meta >> (8 * k)) & 0xFF);
    const int rows_warp = __reduce_max_sync(0xffffffffu, rows);
    const std::int32_t *sidx = P.sidx_of(tile, k) + lane;
    const double *Wk = P.W_of(tile, k) + lane;
    for (int j = 0; j < rows_warp; ++j) {
      if (j < rows) {
        const std::int64_t g = sidx[(std::int64_t)j * TILE];
        double rhs[NVARS];
#pragma unroll
        for (int v = 0; v < NVARS; ++v) rhs[v] = args.state[g * NVARS + v];
        if (WB) {
          const double *av = P.eq_avg + ((tile * P.eq_rows + P.eq_row0[k] + j) * 2) * TILE + lane;
          rhs[0] -= av[0];
          rhs[4] -= av[TILE];
        }
#pragma unroll
        for (int v = 0; v < NVARS; ++v) rhs[v] = rhs[v] * inv_scale[v] - q0s[v];
        const double *wrow = Wk + (std::int64_t)j * NC * TILE;
        for (int c = 0; c < NC; ++c) {
          const double wv = wrow[c * TILE];
#pragma unroll
          for (int v = 0; v < NVARS; ++v) pk[k][c][v] = fma(wv, rhs[v], pk[k][c][v]);
        }
      }
    }
  }

  // ---- CWENO correction of the highest-order polynomial (cweno_ao.cpp:41-50) ------------------------------
  const int n_eff = single ? 1 : NS;
  double a0h[NVARS];
#pragma unroll
  for (int v = 0; v < NVARS; ++v) a0h[v] = q0s[v];
  if (sc.recon_mode == RECON_CWENO_AO) {
    const double gh = single ? 1.0 : sc.lin_w[kh];
    const double inv_gh = 1.0 / gh;
    for (int k = 0; k < n_eff; ++k) {
      if (k == kh) continue;
      const double g = sc.lin_w[k];
#pragma unroll
      for (int v = 0; v < NVARS; ++v) a0h[v] -= g * q0s[v];
      for (int c = 0; c < sc.ncoef[k]; ++c)
#pragma unroll
        for (int v = 0; v < NVARS; ++v) pk[kh][c][v] -= g * pk[k][c][v];
    }
#pragma unroll
    for (int v = 0; v < NVARS; ++v) a0h[v] *= inv_gh;
    for (int c = 0; c < CM; ++c)
#pragma unroll
      for (int v = 0; v < NVARS; ++v) pk[kh][c][v] *= inv_gh;
  }

  // ---- smoothness indicators, non-linear weights, hybridised polynomial (hybrid_weno.cpp:110-128) ---------
  double alpha[MAX_STENCILS];
  double al_tot = 0.0;
  for (int k = 0; k < NS; ++k) {
    double is_max = 0.0;
    const int nck = (k == kh) ? CM : sc.ncoef[k];  // the corrected polynomial has the family's full degree
#pragma unroll
    for (int v = 0; v < NVARS; ++v) {
      double beta = 0.0;
      for (int c = 0; c < nck; ++c) beta += pk[k][c][v] * pk[k][c][v];
      is_max = (v == 0) ? beta : ref_max(is_max, beta);
    }
    double is_pow;
    if (sc.exponent == 4.0) {
      const double s2 = is_max * is_max;
      is_pow = s2 * s2;
    } else if (sc.exponent == 2.0) {
      is_pow = is_max * is_max;
    } else {
      is_pow = pow(is_max, sc.exponent);
    }
    const double g = single ? 1.0 : sc.lin_w[k];
    alpha[k] = (k < n_eff) ? g / (sc.epsilon + is_pow) : 0.0;
    al_tot += alpha[k];
  }
  double coef[DM][NVARS];
  for (int i = 0; i < DM; ++i)
#pragma unroll
    for (int v = 0; v < NVARS; ++v) coef[i][v] = 0.0;
  for (int k = 0; k < NS; ++k) {
    const double wk = alpha[k] / al_tot;
#pragma unroll
    for (int v = 0; v < NVARS; ++v) coef[0][v] += wk * ((k == kh && sc.recon_mode == RECON_CWENO_AO) ? a0h[v] : q0s[v]);
    for (int c = 0; c < CM; ++c)
#pragma unroll
      for (int v = 0; v < NVARS; ++v) coef[1 + c][v] += wk * pk[k][c][v];
  }
  if (P.poly != nullptr && active) {
    for (int i = 0; i < D; ++i)
      for (int v = 0; v < NVARS; ++v) P.poly[(cell * P.n_poly_coef + i) * NVARS + v] = coef[i][v];
    for (int v = 0; v < NVARS; ++v) P.poly_scale[cell * NVARS + v] = scale[v];
  }
  for (int i = 0; i < DM; ++i)
#pragma unroll
    for (int v = 0; v < NVARS; ++v) coef[i][v] *= scale[v];

  // ---- geometry ---------------------------------------------------------------------------------------------
  double vt[F][3];
#pragma unroll
  for (int k = 0; k < F; ++k)
#pragma unroll
    for (int d = 0; d < 3; ++d) vt[k][d] = (ND == 2 && d == 2) ? 0.0 : P.vtx[((tile * F + k) * 3 + d) * TILE + lane];
  double xc[3];
#pragma unroll
  for (int d = 0; d < 3; ++d) xc[d] = (ND == 2 && d == 2) ? 0.0 : P.center[(tile * 3 + d) * TILE + lane];
  const double inv_len = P.inv_len[tile * TILE + lane];
  double cmom[DM];
  for (int i = 0; i < DM; ++i) cmom[i] = (i >= 3 && i < D) ? P.moments[(tile * P.n_mom + (i - 3)) * TILE + lane] : 0.0;

  auto eval_delta = [&](const double x[3], double out[NVARS]) {
    double mono[DM];
    PolyEval<ND, GenericLimits<ND>::DEG>::monomials((x[0] - xc[0]) * inv_len, (x[1] - xc[1]) * inv_len,
                                                    (ND == 3) ? (x[2] - xc[2]) * inv_len : 0.0, cmom, mono);
#pragma unroll
    for (int v = 0; v < NVARS; ++v) {
      double s = coef[0][v];
      for (int i = 1; i < D; ++i) s = fma(coef[i][v], mono[i], s);
      out[v] = s;
    }
  };

  // ---- traces at the face Gauss points; well-balanced face term of the source ---------------------------------
  double src[NVARS] = {0.0, 0.0, 0.0, 0.0, 0.0};
  for (int k = 0; k < F; ++k) {
    const std::uint32_t fref = active ? P.face_ref[(tile * F + k) * TILE + lane] : 0u;
    const std::uint32_t slots = P.face_slots[(tile * F + k) * TILE + lane];
    const std::int64_t e = fref & FREF_EDGE_MASK;
    const int side = (fref & FREF_SIDE) ? 1 : 0;
    const bool want_trace = (fref & FREF_TRACE) != 0;
    if (!want_trace && !WB) continue;
    double fv[3][3];
#pragma unroll
    for (int r = 0; r < ND; ++r) {
      const int s = (slots >> (2 * r)) & 3;
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        double val = vt[0][d];
#pragma unroll
        for (int kk = 1; kk < F; ++kk)
          if (s == kk) val = vt[kk][d];
        fv[r][d] = val;
      }
    }
    double nout[3] = {0.0, 0.0, 0.0}, area = 0.0;
    if (WB) {  // unit_outward_normal (face.cpp:24-27)
      const double *fr = P.face_frame + e * 10;
      const double n0 = fr[0], n1 = fr[1], n2 = fr[2];
      area = fr[9];
      double x0[3];
#pragma unroll
      for (int d = 0; d < 3; ++d)
        x0[d] = (ND == 2) ? sc.face_bary[0][0] * fv[0][d] + sc.face_bary[0][1] * fv[1][d]
                          : fv[0][d] * sc.face_bary[0][0] + fv[1][d] * sc.face_bary[0][1] + fv[2][d] * sc.face_bary[0][2];
      const double dt = n0 * (x0[0] - xc[0]) + n1 * (x0[1] - xc[1]) + n2 * (x0[2] - xc[2]);
      const double sg = (dt > 0.0) ? 1.0 : ((dt < 0.0) ? -1.0 : 0.0);
      nout[0] = sg * n0;
      nout[1] = sg * n1;
      nout[2] = sg * n2;
    }
    double s_face[3] = {0.0, 0.0, 0.0};
    for (int q = 0; q < sc.q_f; ++q) {
      double x[3];
#pragma unroll
      for (int d = 0; d < 3; ++d)
        x[d] = (ND == 2) ? sc.face_bary[q][0] * fv[0][d] + sc.face_bary[q][1] * fv[1][d]
                         : fv[0][d] * sc.face_bary[q][0] + fv[1][d] * sc.face_bary[q][1] + fv[2][d] * sc.face_bary[q][2];
      double bg_rho = 0.0, bg_E = 0.0;
      if (WB) {
        double p_eq;
        eq.template at<POWN>(P.phi_fqp[e * sc.q_f + q], sc, bg_rho, bg_E, p_eq);
        const double wq = area * sc.face_w[q];
#pragma unroll
        for (int d = 0; d < 3; ++d) s_face[d] = (q == 0) ? wq * (p_eq * nout[d]) : s_face[d] + wq * (p_eq * nout[d]);
      }
      if (want_trace) {
        double du[NVARS];
        eval_delta(x, du);
        double *tr = P.trace + ((e * 2 + side) * sc.q_f + q) * NVARS;
        tr[0] = bg_rho + du[0];
        tr[1] = du[1];
        tr[2] = du[2];
        tr[3] = du[3];
        tr[4] = bg_E + du[4];
      }
    }
    if (WB) {
#pragma unroll
      for (int d = 0; d < 3; ++d) src[1 + d] += s_face[d];
    }
  }

  // ---- volume part of the gravity source, heating (gravity_source_loop.hpp:63-81,121-140; heating.hpp:30-44) ----
  if (GRAV) {
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    double heat = 0.0;
    for (int q = 0; q < sc.q_c; ++q) {
      double x[3];
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        if (ND == 2)
          x[d] = vt[0][d] * sc.cell_bary[q][0] + vt[1][d] * sc.cell_bary[q][1] + vt[2][d] * sc.cell_bary[q][2];
        else
          x[d] = vt[0][d] * sc.cell_bary[q][0] + vt[1][d] * sc.cell_bary[q][1] + vt[2][d] * sc.cell_bary[q][2] +
                 vt[F - 1][d] * sc.cell_bary[q][3];
      }
      double du[NVARS];
      eval_delta(x, du);
      const double *gp = P.gradphi_cqp + (ci * sc.q_c + q) * 3;
      const double g0 = gp[0], g1 = gp[1], g2 = gp[2];
      const double rho = du[0];
      const double s1 = -rho * g0, s2 = -rho * g1, s3 = -rho * g2;
      const double s4 = -(du[1] * g0 + du[2] * g1 + du[3] * g2);
      const double wq = sc.cell_w[q];
      if (sc.heating_rate != 0.0) {
        double rho_full = du[0];
        if (WB) {
          double br, bE, bp;
          eq.template at<POWN>(P.phi_cqp[ci * sc.q_c + q], sc, br, bE, bp);
          rho_full += br;
        }
        const double r = sqrt(x[0] * x[0] + x[1] * x[1] + x[2] * x[2]);
        const double hq = rho_full * ((sc.heating_r0 <= r && r <= sc.heating_r1) ? sc.heating_rate : 0.0);
        heat = (q == 0) ? wq * hq : heat + wq * hq;
      }
      if (q == 0) {
        acc[0] = wq * s1;
        acc[1] = wq * s2;
        acc[2] = wq * s3;
        acc[3] = wq * s4;
      } else {
        acc[0] += wq * s1;
        acc[1] += wq * s2;
        acc[2] += wq * s3;
        acc[3] += wq * s4;
      }
    }
    if (WB) {
      const double inv_vol = 1.0 / P.volume[tile * TILE + lane];
      src[1] = src[1] * inv_vol + acc[0];
      src[2] = src[2] * inv_vol + acc[1];
      src[3] = src[3] * inv_vol + acc[2];
      src[4] = acc[3];
    } else {
      src[1] = acc[0];
      src[2] = acc[1];
      src[3] = acc[2];
      src[4] = acc[3];
    }
    src[4] += heat;
    if (active) {
#pragma unroll
      for (int v = 0; v < NVARS; ++v) P.source[cell * NVARS + v] = src[v];
    }
  }
}

struct TracerGenericArgs {
  DevicePlan plan;
  const double *avars;
  const std::int32_t *tile_list;
  std::int64_t n_tiles_launch;
};

/// T1 for generic families: LocalReconstruction::compute_tracer (local_reconstruction.hpp:127-147) per scalar.
template <int ND>
__global__ void __launch_bounds__(128) tracer_generic_kernel(const __grid_constant__ TracerGenericArgs args,
                                                             const __grid_constant__ SchemeConst sc) {
  constexpr int F = ND + 1;
  constexpr int DM = GenericLimits<ND>::D;
  constexpr int CM = DM - 1;
  const DevicePlan &P = args.plan;
  const int NS = sc.n_stencils, D = P.n_poly_coef, NA = P.n_avars;
  const int lane = threadIdx.x & 31;
  const std::int64_t w = (std::int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (w >= args.n_tiles_launch) return;
  const std::int64_t tile = args.tile_list ? (std::int64_t)args.tile_list[w] : w;
  const std::int64_t cell = tile * TILE + lane;
  const bool active = cell < P.n_cells;
  const std::int64_t ci = active ? cell : P.n_cells - 1;
  const std::uint64_t meta = active ? P.meta_of(tile)[lane] : 0ull;
  const int kh = (int)((meta >> 56) & 0xF);
  const bool single = ((meta >> 60) & 1) != 0;
  const int n_eff = single ? 1 : NS;

  double vt[F][3];
#pragma unroll
  for (int k = 0; k < F; ++k)
#pragma unroll
    for (int d = 0; d < 3; ++d) vt[k][d] = (ND == 2 && d == 2) ? 0.0 : P.vtx[((tile * F + k) * 3 + d) * TILE + lane];
  double xc[3];
#pragma unroll
  for (int d = 0; d < 3; ++d) xc[d] = (ND == 2 && d == 2) ? 0.0 : P.center[(tile * 3 + d) * TILE + lane];
  const double inv_len = P.inv_len[tile * TILE + lane];
  double cmom[DM];
  for (int i = 0; i < DM; ++i) cmom[i] = (i >= 3 && i < D) ? P.moments[(tile * P.n_mom + (i - 3)) * TILE + lane] : 0.0;

  for (int a = 0; a < NA; ++a) {
    const double q0 = args.avars[ci * NA + a];
    double pk[MAX_STENCILS][CM];
    for (int k = 0; k < NS; ++k)
      for (int c = 0; c < CM; ++c) pk[k][c] = 0.0;
    for (int k = 0; k < NS; ++k) {
      const int NC = sc.ncoef[k];
      const int rows = (int)((meta >> (8 * k)) & 0xFF);
      const std::int32_t *sidx = P.sidx_of(tile, k) + lane;
      const double *Wk = P.W_of(tile, k) + lane;
      for (int j = 0; j < rows; ++j) {
        const std::int64_t g = sidx[(std::int64_t)j * TILE];
        const double rhs = args.avars[g * NA + a] - q0;  // hybrid_weno.cpp:80-84
        const double *wrow = Wk + (std::int64_t)j * NC * TILE;
        for (int c = 0; c < NC; ++c) pk[k][c] = fma(wrow[c * TILE], rhs, pk[k][c]);
      }
    }
    double a0h = q0;
    const bool cweno = sc.recon_mode == RECON_CWENO_AO;
    if (cweno) {
      const double gh = single ? 1.0 : sc.lin_w[kh];
      const double inv_gh = 1.0 / gh;
      for (int k = 0; k < n_eff; ++k) {
        if (k == kh) continue;
        const double g = sc.lin_w[k];
        a0h -= g * q0;
        for (int c = 0; c < sc.ncoef[k]; ++c) pk[kh][c] -= g * pk[k][c];
      }
      a0h *= inv_gh;
      for (int c = 0; c < CM; ++c) pk[kh][c] *= inv_gh;
    }
    double alpha[MAX_STENCILS], al_tot = 0.0;
    for (int k = 0; k < NS; ++k) {
      double beta = 0.0;
      const int nck = (k == kh) ? CM : sc.ncoef[k];
      for (int c = 0; c < nck; ++c) beta += pk[k][c] * pk[k][c];
      double is_pow;
      if (sc.exponent == 4.0) {
        const double s2 = beta * beta;
        is_pow = s2 * s2;
      } else if (sc.exponent == 2.0) {
        is_pow = beta * beta;
      } else {
        is_pow = pow(beta, sc.exponent);
      }
      const double g = single ? 1.0 : sc.lin_w[k];
      alpha[k] = (k < n_eff) ? g / (sc.epsilon + is_pow) : 0.0;
      al_tot += alpha[k];
    }
    double coef[DM];
    for (int i = 0; i < DM; ++i) coef[i] = 0.0;
    for (int k = 0; k < NS; ++k) {
      const double wk = alpha[k] / al_tot;
      coef[0] += wk * ((k == kh && cweno) ? a0h : q0);
      for (int c = 0; c < CM; ++c) coef[1 + c] += wk * pk[k][c];
    }
    for (int k = 0; k < F; ++k) {
      const std::uint32_t fref = active ? P.face_ref[(tile * F + k) * TILE + lane] : 0u;
      if (!(fref & FREF_TRACE)) continue;
      const std::uint32_t slots = P.face_slots[(tile * F + k) * TILE + lane];
      const std::int64_t e = fref & FREF_EDGE_MASK;
      const int side = (fref & FREF_SIDE) ? 1 : 0;
      double fv[3][3];
#pragma unroll
      for (int r = 0; r < ND; ++r) {
        const int s = (slots >> (2 * r)) & 3;
#pragma unroll
        for (int d = 0; d < 3; ++d) {
          double val = vt[0][d];
#pragma unroll
          for (int kk = 1; kk < F; ++kk)
            if (s == kk) val = vt[kk][d];
          fv[r][d] = val;
        }
      }
      for (int q = 0; q < sc.q_f; ++q) {
        double x[3];
#pragma unroll
        for (int d = 0; d < 3; ++d)
          x[d] = (ND == 2) ? sc.face_bary[q][0] * fv[0][d] + sc.face_bary[q][1] * fv[1][d]
                           : fv[0][d] * sc.face_bary[q][0] + fv[1][d] * sc.face_bary[q][1] + fv[2][d] * sc.face_bary[q][2];
        double mono[DM];
        PolyEval<ND, GenericLimits<ND>::DEG>::monomials((x[0] - xc[0]) * inv_len, (x[1] - xc[1]) * inv_len,
                                                        (ND == 3) ? (x[2] - xc[2]) * inv_len : 0.0, cmom, mono);
        double s = coef[0];
        for (int i = 1; i < D; ++i) s = fma(coef[i], mono[i], s);
        P.qtrace[((e * 2 + side) * sc.q_f + q) * NA + a] = s;
      }
    }
  }
}

}  // namespace

bool recon_generic_supported(const SchemeConst &sc, int n_poly_coef) {
  if (sc.n_stencils < 1 || sc.n_stencils > MAX_STENCILS) return false;
  const int dm = (sc.n_dims == 2) ? GenericLimits<2>::D : GenericLimits<3>::D;
  if (n_poly_coef > dm) return false;
  for (int k = 0; k < sc.n_stencils; ++k)
    if (sc.ncoef[k] > dm - 1 || sc.rows_max[k] > 255) return false;
  return true;
}

int launch_recon_generic(const DevicePlan &plan, const SchemeConst &sc, const double *state,
                         const std::int32_t *tile_list, std::int64_t n_tiles, cudaStream_t stream) {
  if (n_tiles <= 0) return 0;
  if (!recon_generic_supported(sc, plan.n_poly_coef) || plan.rec == nullptr) return 1;
  ReconArgs args;
  args.plan = plan;
  args.state = state;
  args.tile_list = tile_list;
  args.n_tiles_launch = n_tiles;
  const unsigned grid = (unsigned)((n_tiles + 3) / 4);
#define ZFVM_GENERIC(ND)                                                                                    \
  if (sc.well_balanced) {                                                                                   \
    const unsigned g1 = (unsigned)((n_tiles * TILE + 255) / 256);                                           \
    const unsigned g2 = (unsigned)((n_tiles * plan.eq_rows + 7) / 8);                                       \
    switch (sc.eos_pow_n) {                                                                                 \
      case 2:                                                                                               \
        eq_solve_kernel<2><<<g1, 256, 0, stream>>>(plan, sc, state, tile_list, n_tiles);                    \
        if (plan.eq_rows > 0) eq_member_kernel<2><<<g2, 256, 0, stream>>>(plan, sc, tile_list, n_tiles);    \
        recon_generic_kernel<ND, RV_WELL_BALANCED, 2><<<grid, 128, 0, stream>>>(args, sc);                  \
        break;                                                                                              \
      case 3:                                                                                               \
        eq_solve_kernel<3><<<g1, 256, 0, stream>>>(plan, sc, state, tile_list, n_tiles);                    \
        if (plan.eq_rows > 0) eq_member_kernel<3><<<g2, 256, 0, stream>>>(plan, sc, tile_list, n_tiles);    \
        recon_generic_kernel<ND, RV_WELL_BALANCED, 3><<<grid, 128, 0, stream>>>(args, sc);                  \
        break;                                                                                              \
      case 5:                                                                                               \
        eq_solve_kernel<5><<<g1, 256, 0, stream>>>(plan, sc, state, tile_list, n_tiles);                    \
        if (plan.eq_rows > 0) eq_member_kernel<5><<<g2, 256, 0, stream>>>(plan, sc, tile_list, n_tiles);    \
        recon_generic_kernel<ND, RV_WELL_BALANCED, 5><<<grid, 128, 0, stream>>>(args, sc);                  \
        break;                                                                                              \
      default:                                                                                              \
        eq_solve_kernel<0><<<g1, 256, 0, stream>>>(plan, sc, state, tile_list, n_tiles);                    \
        if (plan.eq_rows > 0) eq_member_kernel<0><<<g2, 256, 0, stream>>>(plan, sc, tile_list, n_tiles);    \
        recon_generic_kernel<ND, RV_WELL_BALANCED, 0><<<grid, 128, 0, stream>>>(args, sc);                  \
        break;                                                                                              \
    }                                                                                                       \
  } else if (sc.has_gravity) {                                                                              \
    recon_generic_kernel<ND, RV_GRAVITY, 0><<<grid, 128, 0, stream>>>(args, sc);                            \
  } else {                                                                                                  \
    recon_generic_kernel<ND, RV_PLAIN, 0><<<grid, 128, 0, stream>>>(args, sc);                              \
  }
  if (sc.n_dims == 2) {
    ZFVM_GENERIC(2)
  } else {
    ZFVM_GENERIC(3)
  }
#undef ZFVM_GENERIC
  return 0;
}

int launch_tracer_recon_generic(const DevicePlan &P, const SchemeConst &sc, const double *avars,
                                const std::int32_t *tile_list, std::int64_t n_tiles, cudaStream_t stream) {
  if (n_tiles <= 0 || P.n_avars <= 0) return 0;
  if (!recon_generic_supported(sc, P.n_poly_coef) || P.rec == nullptr) return 1;
  TracerGenericArgs args{P, avars, tile_list, n_tiles};
  const unsigned grid = (unsigned)((n_tiles + 3) / 4);
  if (sc.n_dims == 2)
    tracer_generic_kernel<2><<<grid, 128, 0, stream>>>(args, sc);
  else
    tracer_generic_kernel<3><<<grid, 128, 0, stream>>>(args, sc);
  return 0;
}

}  // namespace zfvm
