// T1: reconstruction of the advected scalars (AllVariables::avars) and their traces at the face Gauss points.
//
//   EulerGlobalReconstruction::compute, set_tracer_local   global_reconstruction_impl.hpp:110-126,177-190
//   LocalReconstruction::compute_tracer / tracer(x, k)      local_reconstruction.hpp:127-147,153-155
//   HybridWENO::compute_polys_impl for ScalarPoly           hybrid_weno.cpp:72-92
//   CWENO_AO / WENO_AO::reconstruct for ScalarPoly           cweno_ao.cpp:24-53, weno_ao.cpp:26-36
//   HybridWENO::eno_hybridize                                hybrid_weno.cpp:110-128
//
// Every scalar is a lone variable: no equilibrium, no characteristic scaling, its own smoothness indicators and
// non-linear weights; stencils and pseudo-inverse weights W_k are the cell's (the same tile records the Euler
// reconstruction streams, whichever of the two record kinds the context holds).  One warp owns a tile, one thread
// a cell, the scalars are looped over; the W rows of a tile are then served from L1 / L2 for the second scalar on.
// The tracer flux and the update live next to the Euler ones in flux_update.cu.
#include "common.cuh"
#include "kernels.hpp"

namespace zfvm {

namespace {

struct TracerArgs {
  DevicePlan plan;
  TracerRecView view;
  const double *avars;            // [n][n_avars]
  const std::int32_t *tile_list;  // optional list of tiles (null: all)
  std::int64_t n_tiles_launch;
};

template <int ND>
__global__ void __launch_bounds__(128) tracer_recon_kernel(const __grid_constant__ TracerArgs args,
                                                           const __grid_constant__ SchemeConst sc) {
  constexpr int F = ND + 1;
  constexpr int DEGMAX = (ND == 2) ? 4 : 3;  // LSQ matrices exist up to order 5 in 2D, 4 in 3D
  constexpr int DMAX = dof_of(DEGMAX, ND);
  constexpr int CMAX = DMAX - 1;
  const DevicePlan &P = args.plan;
  const TracerRecView &V = args.view;
  const int NA = P.n_avars, NS = sc.n_stencils;

  const int lane = threadIdx.x & 31;
  const std::int64_t w = (std::int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (w >= args.n_tiles_launch) return;
  const std::int64_t tile = args.tile_list ? (std::int64_t)args.tile_list[w] : w;
  const std::int64_t cell = tile * TILE + lane;
  const bool active = cell < P.n_cells;
  const std::int64_t ci = active ? cell : P.n_cells - 1;

  const char *rec = V.tile_record ? P.rec2 + tile * P.rec2_bytes : P.rec + tile * P.rec_bytes;
  const std::uint64_t meta = active ? reinterpret_cast<const std::uint64_t *>(rec + V.off_meta)[lane] : 0ull;
  const int kh = (int)((meta >> 56) & 0xF);
  const bool single = ((meta >> 60) & 1) != 0;
  const int n_eff = single ? 1 : NS;
  const int nc_hi = sc.ncoef[0];  // the first stencil has the highest order: every polynomial fits nc_hi coefficients
  const int D = nc_hi + 1;

  // global index of member j (row j of W_k) of stencil k
  auto member = [&](int k, int j) -> std::int64_t {
    if (V.tile_record) {
      const char *row = rec + V.off_lidx + (std::size_t)(V.row0[k] + j) * TILE * V.lidx_elem;
      const int li = (V.lidx_elem == 1) ? (int)reinterpret_cast<const std::uint8_t *>(row)[lane]
                                        : (int)reinterpret_cast<const std::uint16_t *>(row)[lane];
      return reinterpret_cast<const std::int32_t *>(rec + V.off_list)[li];
    }
    return reinterpret_cast<const std::int32_t *>(rec + V.off_sidx[k])[(std::size_t)j * TILE + lane];
  };

  // geometry of the cell
  double vt[F][3];
#pragma unroll
  for (int k = 0; k < F; ++k)
#pragma unroll
    for (int d = 0; d < 3; ++d) vt[k][d] = (ND == 2 && d == 2) ? 0.0 : P.vtx[((tile * F + k) * 3 + d) * TILE + lane];
  double xc[3];
#pragma unroll
  for (int d = 0; d < 3; ++d) xc[d] = (ND == 2 && d == 2) ? 0.0 : P.center[(tile * 3 + d) * TILE + lane];
  const double inv_len = P.inv_len[tile * TILE + lane];
  double cmom[DMAX];
#pragma unroll
  for (int i = 0; i < DMAX; ++i) cmom[i] = (i >= 3 && i < D) ? P.moments[(tile * P.n_mom + (i - 3)) * TILE + lane] : 0.0;

  for (int a = 0; a < NA; ++a) {
    const double q0 = args.avars[ci * NA + a];
    double pc[MAX_STENCILS][CMAX];
    for (int k = 0; k < NS; ++k) {
      for (int c = 0; c < CMAX; ++c) pc[k][c] = 0.0;
      const int NC = sc.ncoef[k];
      const int rows = (int)((meta >> (8 * k)) & 0xFF);
      const int rows_warp = __reduce_max_sync(0xffffffffu, rows);
      const double *Wk = reinterpret_cast<const double *>(rec + V.off_w[k]) + lane;
      for (int j = 0; j < rows_warp; ++j) {
        if (j < rows) {
          const double rhs = args.avars[member(k, j) * NA + a] - q0;  // hybrid_weno.cpp:80-84
          const double *wrow = Wk + (std::size_t)j * NC * TILE;
          for (int c = 0; c < NC; ++c) pc[k][c] = fma(wrow[c * TILE], rhs, pc[k][c]);
        }
      }
    }
    // CWENO correction of the highest-order polynomial (cweno_ao.cpp:41-50)
    double a0h = q0;
    if (sc.recon_mode == RECON_CWENO_AO) {
      for (int k = 0; k < NS; ++k) {
        if (k != kh && k < n_eff) {
          const double g = sc.lin_w[k];
          a0h -= g * q0;
          for (int c = 0; c < nc_hi; ++c) pc[kh][c] -= g * pc[k][c];
        }
      }
      const double inv_gh = 1.0 / (single ? 1.0 : sc.lin_w[kh]);
      a0h *= inv_gh;
      for (int c = 0; c < nc_hi; ++c) pc[kh][c] *= inv_gh;
    }
    // smoothness indicators, non-linear weights, hybridised polynomial (hybrid_weno.cpp:110-128)
    double alpha[MAX_STENCILS], al_tot = 0.0;
    for (int k = 0; k < NS; ++k) {
      double beta = 0.0;
      for (int c = 0; c < nc_hi; ++c) beta += pc[k][c] * pc[k][c];
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
    double coef[DMAX];
#pragma unroll
    for (int i = 0; i < DMAX; ++i) coef[i] = 0.0;
    for (int k = 0; k < NS; ++k) {
      const double wk = alpha[k] / al_tot;
      coef[0] += wk * ((k == kh) ? a0h : q0);
#pragma unroll
      for (int c = 0; c < CMAX; ++c)
        if (c < nc_hi) coef[1 + c] += wk * pc[k][c];
    }

    // traces at the face Gauss points: scalar_polys[k_var](x), local_reconstruction.hpp:153-155
#pragma unroll
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
        const double xi = (x[0] - xc[0]) * inv_len, eta = (x[1] - xc[1]) * inv_len;
        const double zeta = (ND == 3) ? (x[2] - xc[2]) * inv_len : 0.0;
        double px[DEGMAX + 1], py[DEGMAX + 1], pz[DEGMAX + 1];
        px[0] = py[0] = pz[0] = 1.0;
#pragma unroll
        for (int d = 1; d <= DEGMAX; ++d) {
          px[d] = px[d - 1] * xi;
          py[d] = py[d - 1] * eta;
          pz[d] = pz[d - 1] * zeta;
        }
        constexpr ExpoTable<ND, DEGMAX> tab{};
        double s = coef[0];
#pragma unroll
        for (int i = 1; i < DMAX; ++i) {
          const double m = (ND == 2) ? px[tab.e[i].a] * py[tab.e[i].b] : px[tab.e[i].a] * py[tab.e[i].b] * pz[tab.e[i].c];
          s = fma(coef[i], m - cmom[i], s);  // coefficients beyond the scheme's degree are zero
        }
        P.qtrace[((e * 2 + side) * sc.q_f + q) * NA + a] = s;
      }
    }
  }
}

}  // namespace

int launch_tracer_recon(const DevicePlan &P, const SchemeConst &sc, const TracerRecView &view, const double *avars,
                        const std::int32_t *tile_list, std::int64_t n_tiles, cudaStream_t stream) {
  if (n_tiles <= 0 || P.n_avars <= 0) return 0;
  const int cmax = dof_of(sc.n_dims == 2 ? 4 : 3, sc.n_dims) - 1;
  if (sc.ncoef[0] > cmax) return 1;
  TracerArgs args{P, view, avars, tile_list, n_tiles};
  const int wpc = 4;
  const unsigned grid = (unsigned)((n_tiles + wpc - 1) / wpc);
  if (sc.n_dims == 2)
    tracer_recon_kernel<2><<<grid, 32 * wpc, 0, stream>>>(args, sc);
  else
    tracer_recon_kernel<3><<<grid, 32 * wpc, 0, stream>>>(args, sc);
  return 0;
}

}  // namespace zfvm
