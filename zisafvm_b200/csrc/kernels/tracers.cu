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
// The tracer flux is part of the face-flux kernel (K2, flux_update.cu: it needs the wave speeds of the face's HLLC
// evaluation), the avars update lives next to K3.
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

// Compile-time stencil count and degrees (the reference's parameter sets: one central stencil of degree DEG_HI, ND + 1
// one-sided stencils of degree DEG_LO): the NS * CLO + NHI coefficients of a scalar's stencil polynomials live in
// registers.  Stencil *sizes* stay run-time (ragged stencils next to boundaries, rows_max from the scheme).
template <int ND, int DEG_HI, int DEG_LO, int NS>
__global__ void __launch_bounds__(128, 4) tracer_recon_kernel(const __grid_constant__ TracerArgs args,
                                                           const __grid_constant__ SchemeConst sc) {
  constexpr int F = ND + 1;
  constexpr int D = dof_of(DEG_HI, ND);
  constexpr int CHI = D - 1;
  constexpr int CLO = dof_of(DEG_LO, ND) - 1;
  constexpr int NHI = CHI - CLO;
  const DevicePlan &P = args.plan;
  const TracerRecView &V = args.view;
  const int NA = P.n_avars;

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
  const std::int32_t *list = reinterpret_cast<const std::int32_t *>(rec + V.off_list);

  // global index of member j (row j of W_k) of stencil k
  auto member = [&](int k, int j) -> std::int64_t {
    if (V.tile_record) {
      const char *row = rec + V.off_lidx + (std::size_t)(V.row0[k] + j) * TILE * V.lidx_elem;
      const int li = (V.lidx_elem == 1) ? (int)reinterpret_cast<const std::uint8_t *>(row)[lane]
                                        : (int)reinterpret_cast<const std::uint16_t *>(row)[lane];
      return list[li];
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
  double cmom[D];
#pragma unroll
  for (int i = 0; i < D; ++i) cmom[i] = 0.0;
#pragma unroll
  for (int i = 3; i < D; ++i) cmom[i] = P.moments[(tile * P.n_mom + (i - 3)) * TILE + lane];

  for (int a = 0; a < NA; ++a) {
    const double q0 = args.avars[ci * NA + a];
    double lo[NS][CLO > 0 ? CLO : 1], hi[NHI > 0 ? NHI : 1];
#pragma unroll
    for (int k = 0; k < NS; ++k)
#pragma unroll
      for (int c = 0; c < CLO; ++c) lo[k][c] = 0.0;
#pragma unroll
    for (int c = 0; c < NHI; ++c) hi[c] = 0.0;
#pragma unroll
    for (int k = 0; k < NS; ++k) {
      const int NC = (k == 0) ? CHI : CLO;  // compile-time after unrolling
      const int rows = (int)((meta >> (8 * k)) & 0xFF);
      const int rows_warp = __reduce_max_sync(0xffffffffu, rows);
      const double *Wk = reinterpret_cast<const double *>(rec + V.off_w[k]) + lane;
      // rows in chunks: the index -> list -> scalar gather chain of a row is three dependent loads; a chunk's chains
      // (and its W loads) are independent of each other and are issued together.  Rows beyond a ragged stencil's
      // own count are zero-padded in W and point at the cell itself (rhs == 0): no per-lane predicate is needed.
      constexpr int CHK = 6;
      for (int j0 = 0; j0 < rows_warp; j0 += CHK) {
        double rhs[CHK];
#pragma unroll
        for (int jj = 0; jj < CHK; ++jj) {
          const int j = min(j0 + jj, rows_warp - 1);
          rhs[jj] = args.avars[member(k, j) * NA + a] - q0;  // hybrid_weno.cpp:80-84
        }
#pragma unroll
        for (int jj = 0; jj < CHK; ++jj) {
          if (j0 + jj < rows_warp) {
            const double *wrow = Wk + (std::size_t)(j0 + jj) * NC * TILE;
#pragma unroll
            for (int c = 0; c < NC; ++c) {
              const double wv = ld_stream(wrow + c * TILE);
              if (c < CLO)
                lo[k][c] = fma(wv, rhs[jj], lo[k][c]);
              else
                hi[c - CLO] = fma(wv, rhs[jj], hi[c - CLO]);
            }
          }
        }
      }
    }
    // CWENO correction of the highest-order polynomial (cweno_ao.cpp:41-50)
    double a0h = q0;
    if (sc.recon_mode == RECON_CWENO_AO) {
      double cor[CLO > 0 ? CLO : 1];
#pragma unroll
      for (int c = 0; c < CLO; ++c) {
        cor[c] = 0.0;
#pragma unroll
        for (int k = 0; k < NS; ++k)
          if (k == kh) cor[c] = lo[k][c];
      }
#pragma unroll
      for (int k = 0; k < NS; ++k) {
        if (k != kh && k < n_eff) {
          const double g = sc.lin_w[k];
          a0h -= g * q0;
#pragma unroll
          for (int c = 0; c < CLO; ++c) cor[c] -= g * lo[k][c];
        }
      }
      double gh = 1.0;
#pragma unroll
      for (int k = 0; k < NS; ++k)
        if (k == kh) gh = single ? 1.0 : sc.lin_w[k];
      const double inv_gh = 1.0 / gh;
      a0h *= inv_gh;
#pragma unroll
      for (int c = 0; c < CLO; ++c) {
        const double val = inv_gh * cor[c];
#pragma unroll
        for (int k = 0; k < NS; ++k)
          if (k == kh) lo[k][c] = val;
      }
      if (kh == 0) {
#pragma unroll
        for (int c = 0; c < NHI; ++c) hi[c] *= inv_gh;
      }
    }
    // smoothness indicators, non-linear weights, hybridised polynomial (hybrid_weno.cpp:110-128)
    double alpha[NS], al_tot = 0.0;
#pragma unroll
    for (int k = 0; k < NS; ++k) {
      double beta = 0.0;
#pragma unroll
      for (int c = 0; c < CLO; ++c) beta += lo[k][c] * lo[k][c];
      if (k == 0) {
#pragma unroll
        for (int c = 0; c < NHI; ++c) beta += hi[c] * hi[c];
      }
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
    double coef[D];
#pragma unroll
    for (int i = 0; i < D; ++i) coef[i] = 0.0;
#pragma unroll
    for (int k = 0; k < NS; ++k) {
      const double wk = alpha[k] / al_tot;
      coef[0] += wk * ((k == kh) ? a0h : q0);
#pragma unroll
      for (int c = 0; c < CLO; ++c) coef[1 + c] += wk * lo[k][c];
      if (k == 0) {
#pragma unroll
        for (int c = 0; c < NHI; ++c) coef[1 + CLO + c] += wk * hi[c];
      }
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
        double px[DEG_HI + 1], py[DEG_HI + 1], pz[DEG_HI + 1];
        px[0] = py[0] = pz[0] = 1.0;
#pragma unroll
        for (int d = 1; d <= DEG_HI; ++d) {
          px[d] = px[d - 1] * xi;
          py[d] = py[d - 1] * eta;
          pz[d] = pz[d - 1] * zeta;
        }
        constexpr ExpoTable<ND, DEG_HI> tab{};
        double s = coef[0];
#pragma unroll
        for (int i = 1; i < D; ++i) {
          const double m = (ND == 2) ? px[tab.e[i].a] * py[tab.e[i].b] : px[tab.e[i].a] * py[tab.e[i].b] * pz[tab.e[i].c];
          s = fma(coef[i], m - cmom[i], s);
        }
        P.qtrace[((e * 2 + side) * sc.q_f + q) * NA + a] = s;
      }
    }
  }
}

}  // namespace

int launch_tracer_recon(const DevicePlan &P, const SchemeConst &sc, const TracerRecView &view, int deg_hi, int deg_lo,
                        const double *avars, const std::int32_t *tile_list, std::int64_t n_tiles, cudaStream_t stream) {
  if (n_tiles <= 0 || P.n_avars <= 0) return 0;
  if (sc.n_stencils != sc.n_dims + 2 || !(deg_lo == 1 || (deg_hi == 0 && deg_lo == 0))) return 1;  // as launch_recon
  TracerArgs args{P, view, avars, tile_list, n_tiles};
  const int wpc = 4;
  const unsigned grid = (unsigned)((n_tiles + wpc - 1) / wpc);
#define ZFVM_TRACER_CASE(ND, DEG)                                                                       \
  if (sc.n_dims == ND && deg_hi == DEG) {                                                               \
    tracer_recon_kernel<ND, DEG, (DEG >= 1 ? 1 : 0), ND + 2><<<grid, 32 * wpc, 0, stream>>>(args, sc); \
    return 0;                                                                                           \
  }
  ZFVM_TRACER_CASE(2, 1)
  ZFVM_TRACER_CASE(2, 2)
  ZFVM_TRACER_CASE(2, 3)
  ZFVM_TRACER_CASE(2, 4)
  ZFVM_TRACER_CASE(3, 1)
  ZFVM_TRACER_CASE(3, 2)
  ZFVM_TRACER_CASE(3, 3)
#undef ZFVM_TRACER_CASE
  return 1;
}

}  // namespace zfvm
