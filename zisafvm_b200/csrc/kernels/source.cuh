// S1: cell-local source terms from the stored WENO polynomial -- used when the reconstruction itself ran on the tile
// kernel (recon_tile.cuh), which hands the hybridised polynomial over through DevicePlan::poly.
//
//   GravitySourceLoop<NoEquilibrium>      fvm_loops/gravity_source_loop.hpp:101-160   tendency += avg_cell (0, -rho grad phi, -(rho v).grad phi)
//   GravitySourceLoop (well-balanced)     fvm_loops/gravity_source_loop.hpp:32-99     + sum_faces int p_eq n_out dS, only delta rho feels gravity
//   Heating                               model/heating.hpp:30-44                      dE/dt += avg_cell rho(x) rate(x)
//
// One warp owns a tile, one thread a cell; the same arithmetic as the source part of recon.cuh, with the polynomial
// read back (D x 5 coefficients in the scaled basis + the 5 scales, tile-interleaved: DevicePlan::poly_tile) instead of
// being live in registers.
#pragma once
#include "common.cuh"
#include "equilibrium.cuh"
#include "recon.cuh"

namespace zfvm {

template <int ND, int DEG_HI, bool WB, int POWN>
__global__ void __launch_bounds__(128, 3) source_kernel(const __grid_constant__ ReconArgs args,
                                                     const __grid_constant__ SchemeConst sc) {
  constexpr int F = ND + 1;
  constexpr int D = dof_of(DEG_HI, ND);
  const DevicePlan &P = args.plan;
  const int lane = threadIdx.x & 31;
  const std::int64_t w = (std::int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (w >= args.n_tiles_launch) return;
  const std::int64_t tile = args.tile_list ? (std::int64_t)args.tile_list[w] : w;
  const std::int64_t cell = tile * TILE + lane;
  const bool active = cell < P.n_cells;
  const std::int64_t ci = active ? cell : P.n_cells - 1;

  // hybridised polynomial, scale folded in: delta(x) = scale * p(x).  Up to 10 coefficients per variable they live in
  // registers; above (3D order 4, 2D order 5) they are re-read per Gauss point (L1 hits: the block is 256-byte rows)
  constexpr bool IN_REGS = D <= 10;
  const double *pt = P.poly_tile + tile * ((D + 1) * NVARS * TILE) + lane;  // [tile][D + 1][5][32]
  double scale[NVARS];
#pragma unroll
  for (int v = 0; v < NVARS; ++v) scale[v] = pt[(D * NVARS + v) * TILE];
  double coef[IN_REGS ? D : 1][NVARS];
  if constexpr (IN_REGS) {
#pragma unroll
    for (int i = 0; i < D; ++i)
#pragma unroll
      for (int v = 0; v < NVARS; ++v) coef[i][v] = ld_stream(pt + (i * NVARS + v) * TILE) * scale[v];
  }
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

  LocalEq eq{0.0, 1.0, 0.0, false};
  if (WB) {
    const double *par = P.eq_par + ci * 4;
    eq = LocalEq{par[0], par[1], par[2], par[3] != 0.0};
    eq.prepare(sc.gamma);
  }

  double src[NVARS] = {0.0, 0.0, 0.0, 0.0, 0.0};
  if (WB) {
    // sum over the cell's faces of int p_eq n_out dS (gravity_source_loop.hpp:40-56), n_out = unit_outward_normal
#pragma unroll
    for (int k = 0; k < F; ++k) {
      const std::uint32_t fref = active ? ld_stream(P.face_ref + (tile * F + k) * TILE + lane) : 0u;
      const std::uint32_t slots = P.face_slots[(tile * F + k) * TILE + lane];
      const std::int64_t e = fref & FREF_EDGE_MASK;
      double x0[3];  // first Gauss point of the face (face vertices in the left cell's order)
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        double fv[3];
#pragma unroll
        for (int r = 0; r < ND; ++r) {
          const int s = (slots >> (2 * r)) & 3;
          double val = vt[0][d];
#pragma unroll
          for (int kk = 1; kk < F; ++kk)
            if (s == kk) val = vt[kk][d];
          fv[r] = val;
        }
        x0[d] = (ND == 2) ? sc.face_bary[0][0] * fv[0] + sc.face_bary[0][1] * fv[1]
                          : fv[0] * sc.face_bary[0][0] + fv[1] * sc.face_bary[0][1] + fv[2] * sc.face_bary[0][2];
      }
      const double *fr = P.face_frame + e * 10;
      const double n0 = fr[0], n1 = fr[1], n2 = fr[2], area = fr[9];
      const double dt = n0 * (x0[0] - xc[0]) + n1 * (x0[1] - xc[1]) + n2 * (x0[2] - xc[2]);
      const double sg = (dt > 0.0) ? 1.0 : ((dt < 0.0) ? -1.0 : 0.0);
      const double nout[3] = {sg * n0, sg * n1, sg * n2};
      double s_face[3] = {0.0, 0.0, 0.0};
      for (int q = 0; q < sc.q_f; ++q) {
        double r_, E_, p_eq;
        eq.template at<POWN>(P.phi_fqp[e * sc.q_f + q], sc, r_, E_, p_eq);
        const double wq = area * sc.face_w[q];
#pragma unroll
        for (int d = 0; d < 3; ++d) s_face[d] = (q == 0) ? wq * (p_eq * nout[d]) : s_face[d] + wq * (p_eq * nout[d]);
      }
#pragma unroll
      for (int d = 0; d < 3; ++d) src[1 + d] += s_face[d];
    }
  }

  // volume part at the cell's Gauss points
  double acc[4] = {0.0, 0.0, 0.0, 0.0}, heat = 0.0;
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
    double mono[D], du[NVARS];
    PolyEval<ND, DEG_HI>::monomials((x[0] - xc[0]) * inv_len, (x[1] - xc[1]) * inv_len,
                                    (ND == 3) ? (x[2] - xc[2]) * inv_len : 0.0, cmom, mono);
#pragma unroll
    for (int v = 0; v < NVARS; ++v) {
      if constexpr (IN_REGS) {
        double s = coef[0][v];
#pragma unroll
        for (int i = 1; i < D; ++i) s = fma(coef[i][v], mono[i], s);
        du[v] = s;
      } else {
        double s = pt[v * TILE] * scale[v];
#pragma unroll 5
        for (int i = 1; i < D; ++i) s = fma(pt[(i * NVARS + v) * TILE] * scale[v], mono[i], s);
        du[v] = s;
      }
    }
    const double *gp = P.gradphi_cqp + (ci * sc.q_c + q) * 3;
    const double g0 = gp[0], g1 = gp[1], g2 = gp[2];
    const double rho = du[0];  // WB: only the density perturbation feels gravity; otherwise delta == full state
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

}  // namespace zfvm
