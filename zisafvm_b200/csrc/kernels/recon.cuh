// K1: per-cell reconstruction, trace evaluation and cell-local source term.
//
// One thread owns one cell, one warp owns one tile of 32 cells. Per cell this kernel does what
//   EulerGlobalReconstruction::compute   global_reconstruction_impl.hpp:133-174
//   LocalReconstruction::compute         local_reconstruction.hpp:69-120
//   HybridWENO::compute_polys_impl       hybrid_weno.cpp:72-92   (as W_k * rhs, W_k = pinv(A_k))
//   CWENO_AO::reconstruct_impl           cweno_ao.cpp:36-53
//   HybridWENO::eno_hybridize            hybrid_weno.cpp:110-128
//   rc(i)(x) at the face Gauss points    flux_loop.hpp:131-149, local_reconstruction.hpp:149-163
//   GravitySourceLoop (both variants)    gravity_source_loop.hpp:32-87,121-148
// do in the reference, and writes
//   trace[e][side][q][5]   the reconstructed state at every Gauss point of the cell's faces
//   source[i][5]           the cell's gravity source term (already divided by the volume)
// so that the face kernel (K2) is a pure Riemann-solver pass.
//
// The stencil weights W_k stream through once, tile-interleaved ([tile][row][coef][lane]), so every
// load instruction of a warp is one contiguous 256-byte segment.
#pragma once
#include "common.cuh"
#include "equilibrium.cuh"

namespace zfvm {

enum ReconVariant : int { RV_PLAIN = 0, RV_GRAVITY = 1, RV_WELL_BALANCED = 2 };

struct ReconArgs {
  DevicePlan plan;
  const double *state;            // [n][5]
  const std::int32_t *tile_list;  // optional list of tiles to process (null: all)
  std::int64_t n_tiles_launch;
};

template <int ND, int DEG>
struct PolyEval {
  static constexpr int D = dof_of(DEG, ND);
  /// monomials mono[i] = xi^a eta^b zeta^c - c_i for i = 1..D-1 (mono[0] unused: a0 * (1 - 0))
  ZFVM_DEVICE static void monomials(double xi, double eta, double zeta, const double *cmom, double *mono) {
    double px[DEG + 1], py[DEG + 1], pz[DEG + 1];
    px[0] = py[0] = pz[0] = 1.0;
#pragma unroll
    for (int d = 1; d <= DEG; ++d) {
      px[d] = px[d - 1] * xi;
      py[d] = py[d - 1] * eta;
      pz[d] = pz[d - 1] * zeta;
    }
    constexpr ExpoTable<ND, DEG> tab{};
#pragma unroll
    for (int i = 1; i < D; ++i) {
      double m = (ND == 2) ? px[tab.e[i].a] * py[tab.e[i].b] : px[tab.e[i].a] * py[tab.e[i].b] * pz[tab.e[i].c];
      mono[i] = m - cmom[i];
    }
  }
};

template <int ND, int DEG_HI, int DEG_LO, int NS, int VARIANT, int POWN = 0>
__global__ void __launch_bounds__(128) recon_kernel(const __grid_constant__ ReconArgs args,
                                                    const __grid_constant__ SchemeConst sc) {
  constexpr int F = ND + 1;
  constexpr int D = dof_of(DEG_HI, ND);
  constexpr int CHI = dof_of(DEG_HI, ND) - 1;
  constexpr int CLO = dof_of(DEG_LO, ND) - 1;
  constexpr int NHI = CHI - CLO;  // coefficients only the high-order stencil has
  constexpr bool WB = (VARIANT == RV_WELL_BALANCED);
  constexpr bool GRAV = (VARIANT != RV_PLAIN);
  const DevicePlan &P = args.plan;

  const int lane = threadIdx.x & 31;
  const std::int64_t w = (std::int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (w >= args.n_tiles_launch) return;
  const std::int64_t tile = args.tile_list ? (std::int64_t)args.tile_list[w] : w;
  const std::int64_t cell = tile * TILE + lane;
  const bool active = cell < P.n_cells;
  const std::int64_t ci = active ? cell : P.n_cells - 1;  // padded lanes mirror the last cell, write nothing

  const std::uint64_t meta = active ? P.meta_of(tile)[lane] : 0ull;
  const int kh = (int)((meta >> 56) & 0xF);
  const bool single = ((meta >> 60) & 1) != 0;

  // ---- own state, scaling, equilibrium -------------------------------------------------------
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
    // the cell's local equilibrium comes from eq_solve_kernel (E1), its averages over the stencil members from
    // eq_member_kernel (E2); only the cell's own average and point values are evaluated here
    const double *phi_own = P.phi_cqp + ci * sc.q_c;
    const double *par = P.eq_par + ci * 4;
    eq = LocalEq{par[0], par[1], par[2], par[3] != 0.0};
    eq.prepare(sc.gamma);
    eq_cell_average<POWN>(eq, phi_own, sc, eq0_rho, eq0_E);
  }

  // u_local(0) after equilibrium subtraction and scaling (local_reconstruction.hpp:109-116)
  double q0s[NVARS];
  q0s[0] = (u0[0] - eq0_rho) * inv_scale[0];
  q0s[1] = u0[1] * inv_scale[1];
  q0s[2] = u0[2] * inv_scale[2];
  q0s[3] = u0[3] * inv_scale[3];
  q0s[4] = (u0[4] - eq0_E) * inv_scale[4];

  // ---- stencil polynomials: coef = W_k * rhs -------------------------------------------------
  double lo[NS][CLO > 0 ? CLO : 1][NVARS];
  double hi[NHI > 0 ? NHI : 1][NVARS];
#pragma unroll
  for (int k = 0; k < NS; ++k)
#pragma unroll
    for (int c = 0; c < CLO; ++c)
#pragma unroll
      for (int v = 0; v < NVARS; ++v) lo[k][c][v] = 0.0;
#pragma unroll
  for (int c = 0; c < NHI; ++c)
#pragma unroll
    for (int v = 0; v < NVARS; ++v) hi[c][v] = 0.0;

#pragma unroll
  for (int k = 0; k < NS; ++k) {
    const int NC = (k == 0) ? CHI : CLO;  // compile-time after unrolling
    const int rows = (int)((meta >> (8 * k)) & 0xFF);
    const int rows_warp = __reduce_max_sync(0xffffffffu, rows);
    const std::int32_t *sidx = P.sidx_of(tile, k) + lane;
    const double *Wk = P.W_of(tile, k) + lane;
    for (int j = 0; j < rows_warp; ++j) {
      if (j < rows) {
        const std::int64_t g = ld_stream(sidx + (std::int64_t)j * TILE);
        double rhs[NVARS];
#pragma unroll
        for (int v = 0; v < NVARS; ++v) rhs[v] = args.state[g * NVARS + v];
        if (WB) {
          const double *av = P.eq_avg + ((tile * P.eq_rows + P.eq_row0[k] + j) * 2) * TILE + lane;
          rhs[0] -= ld_stream(av);
          rhs[4] -= ld_stream(av + TILE);
        }
#pragma unroll
        for (int v = 0; v < NVARS; ++v) rhs[v] = rhs[v] * inv_scale[v] - q0s[v];
        const double *wrow = Wk + (std::int64_t)j * NC * TILE;
#pragma unroll
        for (int c = 0; c < NC; ++c) {
          const double wv = ld_stream(wrow + c * TILE);
          if (c < CLO) {
#pragma unroll
            for (int v = 0; v < NVARS; ++v) lo[k][c][v] = fma(wv, rhs[v], lo[k][c][v]);
          } else {
#pragma unroll
            for (int v = 0; v < NVARS; ++v) hi[c - CLO][v] = fma(wv, rhs[v], hi[c - CLO][v]);
          }
        }
      }
    }
  }

  // ---- CWENO correction of the highest-order polynomial (cweno_ao.cpp:41-50) ------------------
  const int n_eff = single ? 1 : NS;
  double a0h[NVARS];  // constant coefficient of stencil kh after the correction
#pragma unroll
  for (int v = 0; v < NVARS; ++v) a0h[v] = q0s[v];
  if (sc.recon_mode == RECON_CWENO_AO) {
    double cor[CLO > 0 ? CLO : 1][NVARS];
#pragma unroll
    for (int c = 0; c < CLO; ++c)
#pragma unroll
      for (int v = 0; v < NVARS; ++v) {
        cor[c][v] = 0.0;
#pragma unroll
        for (int k = 0; k < NS; ++k)
          if (k == kh) cor[c][v] = lo[k][c][v];
      }
#pragma unroll
    for (int k = 0; k < NS; ++k) {
      if (k != kh && k < n_eff) {
        const double g = sc.lin_w[k];
#pragma unroll
        for (int v = 0; v < NVARS; ++v) a0h[v] -= g * q0s[v];
#pragma unroll
        for (int c = 0; c < CLO; ++c)
#pragma unroll
          for (int v = 0; v < NVARS; ++v) cor[c][v] -= g * lo[k][c][v];
      }
    }
    double gh = 1.0;
#pragma unroll
    for (int k = 0; k < NS; ++k)
      if (k == kh) gh = single ? 1.0 : sc.lin_w[k];
    const double inv_gh = 1.0 / gh;
#pragma unroll
    for (int v = 0; v < NVARS; ++v) a0h[v] *= inv_gh;
#pragma unroll
    for (int c = 0; c < CLO; ++c)
#pragma unroll
      for (int v = 0; v < NVARS; ++v) {
        const double val = inv_gh * cor[c][v];
#pragma unroll
        for (int k = 0; k < NS; ++k)
          if (k == kh) lo[k][c][v] = val;
      }
    if (kh == 0) {
#pragma unroll
      for (int c = 0; c < NHI; ++c)
#pragma unroll
        for (int v = 0; v < NVARS; ++v) hi[c][v] *= inv_gh;
    }
  }

  // ---- smoothness indicators and non-linear weights (hybrid_weno.cpp:110-128) -----------------
  double alpha[NS];
  double al_tot = 0.0;
#pragma unroll
  for (int k = 0; k < NS; ++k) {
    double is_max = 0.0;
#pragma unroll
    for (int v = 0; v < NVARS; ++v) {
      double beta = 0.0;
#pragma unroll
      for (int c = 0; c < CLO; ++c) beta += lo[k][c][v] * lo[k][c][v];
      if (k == 0) {
#pragma unroll
        for (int c = 0; c < NHI; ++c) beta += hi[c][v] * hi[c][v];
      }
      is_max = (v == 0) ? beta : fmax(is_max, beta);
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

  // ---- hybridised polynomial p = sum_k (alpha_k / al_tot) p_k ----------------------------------
  double coef[D][NVARS];
  {
#pragma unroll
    for (int i = 0; i < D; ++i)
#pragma unroll
      for (int v = 0; v < NVARS; ++v) coef[i][v] = 0.0;
#pragma unroll
    for (int k = 0; k < NS; ++k) {
      const double wk = alpha[k] / al_tot;
#pragma unroll
      for (int v = 0; v < NVARS; ++v) coef[0][v] += wk * ((k == kh) ? a0h[v] : q0s[v]);
#pragma unroll
      for (int c = 0; c < CLO; ++c)
#pragma unroll
        for (int v = 0; v < NVARS; ++v) coef[1 + c][v] += wk * lo[k][c][v];
      if (k == 0) {
#pragma unroll
        for (int c = 0; c < NHI; ++c)
#pragma unroll
          for (int v = 0; v < NVARS; ++v) coef[1 + CLO + c][v] += wk * hi[c][v];
      }
    }
  }
  if (P.poly != nullptr && active) {
    for (int i = 0; i < D; ++i)
      for (int v = 0; v < NVARS; ++v) P.poly[(cell * P.n_poly_coef + i) * NVARS + v] = coef[i][v];
    for (int v = 0; v < NVARS; ++v) P.poly_scale[cell * NVARS + v] = scale[v];
  }
  // fold the characteristic scale into the coefficients: delta(x) = scale * p(x)
#pragma unroll
  for (int i = 0; i < D; ++i)
#pragma unroll
    for (int v = 0; v < NVARS; ++v) coef[i][v] *= scale[v];

  // ---- geometry of the cell ---------------------------------------------------------------------
  double vt[F][3];
#pragma unroll
  for (int k = 0; k < F; ++k)
#pragma unroll
    for (int d = 0; d < 3; ++d)
      vt[k][d] = (ND == 2 && d == 2) ? 0.0 : ld_stream(P.vtx + ((tile * F + k) * 3 + d) * TILE + lane);
  double xc[3];
#pragma unroll
  for (int d = 0; d < 3; ++d) xc[d] = (ND == 2 && d == 2) ? 0.0 : ld_stream(P.center + (tile * 3 + d) * TILE + lane);
  const double inv_len = ld_stream(P.inv_len + tile * TILE + lane);
  double cmom[D];
#pragma unroll
  for (int i = 0; i < D; ++i) cmom[i] = 0.0;
#pragma unroll
  for (int i = 3; i < D; ++i) cmom[i] = ld_stream(P.moments + (tile * P.n_mom + (i - 3)) * TILE + lane);

  auto eval_delta = [&](const double x[3], double out[NVARS]) {
    double mono[D];
    PolyEval<ND, DEG_HI>::monomials((x[0] - xc[0]) * inv_len, (x[1] - xc[1]) * inv_len,
                                    (ND == 3) ? (x[2] - xc[2]) * inv_len : 0.0, cmom, mono);
#pragma unroll
    for (int v = 0; v < NVARS; ++v) {
      double s = coef[0][v];
#pragma unroll
      for (int i = 1; i < D; ++i) s = fma(coef[i][v], mono[i], s);
      out[v] = s;
    }
  };

  // ---- traces at the face Gauss points; well-balanced face term of the source --------------------
  double src[NVARS] = {0.0, 0.0, 0.0, 0.0, 0.0};
#pragma unroll
  for (int k = 0; k < F; ++k) {
    const std::uint32_t fref = active ? ld_stream(P.face_ref + (tile * F + k) * TILE + lane) : 0u;
    const std::uint32_t slots = P.face_slots[(tile * F + k) * TILE + lane];
    const std::int64_t e = fref & FREF_EDGE_MASK;
    const int side = (fref & FREF_SIDE) ? 1 : 0;
    const bool want_trace = (fref & FREF_TRACE) != 0;
    if (!want_trace && !WB) continue;
    // face vertices in the left cell's order
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
    if (WB) {
      // unit_outward_normal (face.cpp:24-27) and the face area; frame as in face_factory.cpp
      const double *fr = P.face_frame + e * 10;
      const double n0 = fr[0], n1 = fr[1], n2 = fr[2];
      area = fr[9];
      double x0[3];
#pragma unroll
      for (int d = 0; d < 3; ++d)
        x0[d] = (ND == 2) ? sc.face_bary[0][0] * fv[0][d] + sc.face_bary[0][1] * fv[1][d]
                          : fv[0][d] * sc.face_bary[0][0] + fv[1][d] * sc.face_bary[0][1] +
                                fv[2][d] * sc.face_bary[0][2];
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
                         : fv[0][d] * sc.face_bary[q][0] + fv[1][d] * sc.face_bary[q][1] +
                               fv[2][d] * sc.face_bary[q][2];
      double bg_rho = 0.0, bg_E = 0.0;
      if (WB) {
        double p_eq;
        eq.template at<POWN>(P.phi_fqp[e * sc.q_f + q], sc, bg_rho, bg_E, p_eq);
        const double wq = area * sc.face_w[q];
        if (q == 0) {
#pragma unroll
          for (int d = 0; d < 3; ++d) s_face[d] = wq * (p_eq * nout[d]);
        } else {
#pragma unroll
          for (int d = 0; d < 3; ++d) s_face[d] = s_face[d] + wq * (p_eq * nout[d]);
        }
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

  // ---- volume part of the gravity source ---------------------------------------------------------
  if (GRAV) {
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    double heat = 0.0;  // Heating::compute (model/heating.hpp:30-44): average(cell, rho(x) * heating_rate(x))
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
      // WB: only the density perturbation feels gravity (gravity_source_loop.hpp:63-81);
      // otherwise the full density (:121-140); delta == full state when there is no background.
      const double rho = du[0];
      const double s1 = -rho * g0, s2 = -rho * g1, s3 = -rho * g2;
      const double s4 = -(du[1] * g0 + du[2] * g1 + du[3] * g2);
      const double wq = sc.cell_w[q];  // reference weight: quadrature / volume
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
      const double inv_vol = 1.0 / ld_stream(P.volume + tile * TILE + lane);
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

}  // namespace zfvm
