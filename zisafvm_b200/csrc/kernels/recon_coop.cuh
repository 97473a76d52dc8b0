// K1, cooperative form: one CTA of four warps owns one tile of 32 cells; lane l of every warp works on cell l.
//
// For the schemes whose central stencil has more coefficients than one thread can keep in registers next to
// everything else (3D order 4: 19 x 5 accumulators, 57 rows; 2D order 5: 14 x 5) and for stencil sizes the tile kernel
// (recon_tile.cuh) is not instantiated for.  Same arithmetic and reference lines as recon_tile.cuh / recon.cuh; the
// work of a cell is split over the four warps instead of being streamed by one:
//
//   * W apply: warp w accumulates the central stencil's coefficients c = w, w + 4, w + 8, .. (<= 5 x 5 accumulators)
//     and the whole one-sided stencil w + 1 (ND x 5 accumulators).  The weights are read straight from the tile record
//     with coalesced 256-byte loads ([row][coef][32] layout: the four warps' loads of a row are adjacent); the latency
//     is hidden by occupancy (3-4 CTAs per SM) and by issuing a row pair's loads together, not by a TMA ring.
//   * the neighbour states come from the same shared-memory table as in the tile kernel (the record's row list and
//     8/16-bit local indices), filled once per tile by all 128 threads.
//   * the raw stencil polynomials meet in shared memory ([coef][var][lane]); every warp then forms the CWENO-AO
//     correction, the smoothness indicators and the non-linear weights for its lane redundantly (identical
//     instruction sequence on identical data: identical bits), the combined coefficients are written back in place,
//     and warp w evaluates the traces of face w from shared memory.
//
// Uses the tile records (DevicePlan::rec2) and, for well-balanced runs, the same equilibrium tables (E1-E3) and the
// same hand-over to source_kernel as the tile kernel.
#pragma once
#include "recon.cuh"

namespace zfvm {

struct CoopCfg {
  int cap, off_list, off_lidx, off_wlo, off_whi, off_geo;
  std::int64_t rec_bytes;
  int rows0, rows_total;     // rows of the central stencil; all stencils' rows (lidx row count)
  int rows_lo[MAX_STENCILS]; // rows of stencil k >= 1 (index k)
  int row0_lo[MAX_STENCILS]; // first lidx row of stencil k >= 1 (the central stencil's rows follow the one-sided ones)
  int wlo_off[MAX_STENCILS]; // byte offset of W_k (k >= 1) from off_wlo
  int s_table, s_coef, s_pk, smem_bytes;
};

constexpr int COOP_WARPS = 4;

template <int ND, int DEG_HI, typename LIDX, bool WB>
__global__ void __launch_bounds__(COOP_WARPS * 32, 3)
    recon_coop_kernel(const __grid_constant__ ReconArgs args, const __grid_constant__ SchemeConst sc,
                      const __grid_constant__ CoopCfg cfg) {
  constexpr int F = ND + 1, NS = ND + 2;
  constexpr int D = dof_of(DEG_HI, ND), CHI = D - 1, CLO = dof_of(1, ND) - 1;
  constexpr int CPW = (CHI + COOP_WARPS - 1) / COOP_WARPS;  // central coefficients per warp
  constexpr int N_MOM = D > 3 ? D - 3 : 0;
  constexpr int GEO_DOUBLES = F * ND + ND + 1 + N_MOM;
  const DevicePlan &P = args.plan;

  extern __shared__ __align__(128) unsigned char smem[];
  double *table = reinterpret_cast<double *>(smem + cfg.s_table);  // [n_list][5]
  double *coef = reinterpret_cast<double *>(smem + cfg.s_coef);    // [D][5][32]: constant | low | high
  double *pk = reinterpret_cast<double *>(smem + cfg.s_pk);        // [NS - 1][CLO][5][32]: one-sided polynomials

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const std::int64_t tile = args.tile_list ? (std::int64_t)args.tile_list[blockIdx.x] : (std::int64_t)blockIdx.x;
  const char *rec = P.rec2 + tile * cfg.rec_bytes;
  const std::int64_t cell = tile * TILE + lane;
  const bool active = cell < P.n_cells;

  // ---- table of the tile's distinct stencil members (flux_loop / global_reconstruction gathers) ------------------
  {
    const int n_list = *reinterpret_cast<const int *>(rec);
    const std::int32_t *list = reinterpret_cast<const std::int32_t *>(rec + cfg.off_list);
    for (int idx = threadIdx.x; idx < n_list * NVARS; idx += blockDim.x) {
      const int row = idx / NVARS;
      table[idx] = args.state[(std::int64_t)__ldg(list + row) * NVARS + (idx - row * NVARS)];
    }
  }
  const std::uint64_t meta = reinterpret_cast<const std::uint64_t *>(rec + TILE_OFF_META)[lane];
  const int kh_m = (int)((meta >> 56) & 0xF);
  const bool single = ((meta >> 60) & 1) != 0;
  __syncthreads();

  // ---- own state and scaling (characteristic_scale.hpp:24-33) --------------------------------------------------
  double q0s[NVARS], inv_scale[NVARS], scale[NVARS];
  {
    double u0[NVARS];
#pragma unroll
    for (int v = 0; v < NVARS; ++v) u0[v] = table[lane * NVARS + v];
    const double ekin0 = 0.5 * (u0[1] * u0[1] + u0[2] * u0[2] + u0[3] * u0[3]) / u0[0];
    double rho_s = u0[0], eint0 = u0[4] - ekin0;
    if (P.scale_state != nullptr) {  // steps_per_recompute != 1: the scale of the last compute_equilibrium
      const double *ss = P.scale_state + 2 * (active ? cell : P.n_cells - 1);
      rho_s = ss[0];
      eint0 = ss[1];
    }
    if (sc.scaling == SCALING_EULER) {
      const double p = eint0 * (sc.gamma - 1.0);
      const double cs = sqrt(sc.gamma * p / rho_s);
      scale[0] = rho_s;
      scale[1] = scale[2] = scale[3] = cs;
      scale[4] = eint0;
    } else {
#pragma unroll
      for (int v = 0; v < NVARS; ++v) scale[v] = 1.0;
    }
    inv_scale[0] = 1.0 / scale[0];
    inv_scale[1] = inv_scale[2] = inv_scale[3] = 1.0 / scale[1];
    inv_scale[4] = 1.0 / scale[4];
    if constexpr (WB) {  // the cell's own equilibrium average: the last row of the tile's eq_avg block
      const double *e0 = P.eq_avg + ((tile * P.eq_rows + cfg.rows_total) * 2) * TILE + lane;
      u0[0] -= e0[0];
      u0[4] -= e0[TILE];
    }
#pragma unroll
    for (int v = 0; v < NVARS; ++v) q0s[v] = u0[v] * inv_scale[v];
  }
  const LIDX *lidx = reinterpret_cast<const LIDX *>(rec + cfg.off_lidx) + lane;
  const double *eq_rows_tile = WB ? P.eq_avg + (tile * P.eq_rows * 2) * TILE + lane : nullptr;
  // rhs of lidx row `row`: u_local(j) - u_local(0) after equilibrium subtraction and scaling, local_reconstruction.hpp:109-116
  auto load_rhs = [&](int row, double rhs[NVARS]) {
    const double *t = table + (int)lidx[row * TILE] * NVARS;
    if constexpr (WB) {
      const double *ea = eq_rows_tile + row * 2 * TILE;
      rhs[0] = fma(t[0] - ea[0], inv_scale[0], -q0s[0]);
      rhs[4] = fma(t[4] - ea[TILE], inv_scale[4], -q0s[4]);
#pragma unroll
      for (int v = 1; v < 4; ++v) rhs[v] = fma(t[v], inv_scale[v], -q0s[v]);
    } else {
#pragma unroll
      for (int v = 0; v < NVARS; ++v) rhs[v] = fma(t[v], inv_scale[v], -q0s[v]);
    }
  };

  // ---- W apply: this warp's share (hybrid_weno.cpp:72-92 with W = pinv(A)) ------------------------------------------
  {
    double acc[CPW][NVARS];
#pragma unroll
    for (int j = 0; j < CPW; ++j)
#pragma unroll
      for (int v = 0; v < NVARS; ++v) acc[j][v] = 0.0;
    const double *w0 = reinterpret_cast<const double *>(rec + cfg.off_whi) + lane + warp * TILE;
    const int row_c0 = cfg.rows_total - cfg.rows0;  // first lidx row of the central stencil
    int r = 0;
    // RB rows at a time: RB * CPW independent weight loads in flight per thread (the kernel has no staging ring; what
    // hides the latency of the weight stream is occupancy times loads in flight)
    constexpr int RB = 4;
    for (; r + RB <= cfg.rows0; r += RB) {
      double wv[RB][CPW], rv[RB][NVARS];
#pragma unroll
      for (int b = 0; b < RB; ++b)
#pragma unroll
        for (int j = 0; j < CPW; ++j)
          wv[b][j] = (warp + COOP_WARPS * j < CHI) ? ld_stream(w0 + ((std::int64_t)(r + b) * CHI + COOP_WARPS * j) * TILE) : 0.0;
#pragma unroll
      for (int b = 0; b < RB; ++b) load_rhs(row_c0 + r + b, rv[b]);
#pragma unroll
      for (int b = 0; b < RB; ++b)
#pragma unroll
        for (int j = 0; j < CPW; ++j)
#pragma unroll
          for (int v = 0; v < NVARS; ++v) acc[j][v] = fma(wv[b][j], rv[b][v], acc[j][v]);
    }
    for (; r < cfg.rows0; ++r) {
      double wa[CPW], ra[NVARS];
#pragma unroll
      for (int j = 0; j < CPW; ++j)
        wa[j] = (warp + COOP_WARPS * j < CHI) ? ld_stream(w0 + ((std::int64_t)r * CHI + COOP_WARPS * j) * TILE) : 0.0;
      load_rhs(row_c0 + r, ra);
#pragma unroll
      for (int j = 0; j < CPW; ++j)
#pragma unroll
        for (int v = 0; v < NVARS; ++v) acc[j][v] = fma(wa[j], ra[v], acc[j][v]);
    }
#pragma unroll
    for (int j = 0; j < CPW; ++j) {
      const int c = warp + COOP_WARPS * j;
      if (c < CHI) {
#pragma unroll
        for (int v = 0; v < NVARS; ++v) coef[((1 + c) * NVARS + v) * TILE + lane] = acc[j][v];
      }
    }
  }
  if (warp + 1 < NS) {  // one-sided stencil k = warp + 1
    const int k = warp + 1;
    double acc[CLO][NVARS];
#pragma unroll
    for (int c = 0; c < CLO; ++c)
#pragma unroll
      for (int v = 0; v < NVARS; ++v) acc[c][v] = 0.0;
    const double *wk = reinterpret_cast<const double *>(rec + cfg.off_wlo + cfg.wlo_off[k]) + lane;
    const int rows = cfg.rows_lo[k], row0 = cfg.row0_lo[k];
    for (int r = 0; r < rows; ++r) {
      double wv[CLO], rhs[NVARS];
#pragma unroll
      for (int c = 0; c < CLO; ++c) wv[c] = ld_stream(wk + ((std::int64_t)r * CLO + c) * TILE);
      load_rhs(row0 + r, rhs);
#pragma unroll
      for (int c = 0; c < CLO; ++c)
#pragma unroll
        for (int v = 0; v < NVARS; ++v) acc[c][v] = fma(wv[c], rhs[v], acc[c][v]);
    }
#pragma unroll
    for (int c = 0; c < CLO; ++c)
#pragma unroll
      for (int v = 0; v < NVARS; ++v) pk[(((k - 1) * CLO + c) * NVARS + v) * TILE + lane] = acc[c][v];
  }
  __syncthreads();

  // ---- hybridisation, per lane, by every warp (cweno_ao.cpp:36-53, hybrid_weno.cpp:110-128) -------------------------
  const bool cweno = sc.recon_mode == RECON_CWENO_AO;
  const int n_eff = single ? 1 : NS;
  const int kh = cweno ? kh_m : -1;
  auto nonlinear_weight = [&](double is_max, double g) {
    double is_pow;
    if (sc.exponent == 4.0) {
      const double s2 = is_max * is_max;
      is_pow = s2 * s2;
    } else if (sc.exponent == 2.0) {
      is_pow = is_max * is_max;
    } else {
      is_pow = pow(is_max, sc.exponent);
    }
    return g * fast_rcp(sc.epsilon + is_pow);
  };
  double corr[CLO][NVARS], wsum[CLO][NVARS];
#pragma unroll
  for (int c = 0; c < CLO; ++c)
#pragma unroll
    for (int v = 0; v < NVARS; ++v) corr[c][v] = wsum[c][v] = 0.0;
  double al_sum = 0.0;
#pragma unroll 1
  for (int k = 1; k < NS; ++k) {
    double a[CLO][NVARS];
#pragma unroll
    for (int c = 0; c < CLO; ++c)
#pragma unroll
      for (int v = 0; v < NVARS; ++v) a[c][v] = pk[(((k - 1) * CLO + c) * NVARS + v) * TILE + lane];
    double is_max = 0.0;
#pragma unroll
    for (int v = 0; v < NVARS; ++v) {
      double beta = 0.0;
#pragma unroll
      for (int c = 0; c < CLO; ++c) beta += a[c][v] * a[c][v];
      is_max = (v == 0) ? beta : ref_max(is_max, beta);
    }
    const double g_k = sc.lin_w[k];
    const double a_k = nonlinear_weight(is_max, single ? 1.0 : g_k);
    const bool exists = k < n_eff, waits = (k == kh);
    const double a_use = (exists && !waits) ? a_k : 0.0;
    const double g_use = (exists && !waits) ? g_k : 0.0;
    al_sum += a_use;
#pragma unroll
    for (int c = 0; c < CLO; ++c)
#pragma unroll
      for (int v = 0; v < NVARS; ++v) {
        corr[c][v] = fma(g_use, a[c][v], corr[c][v]);
        wsum[c][v] = fma(a_use, a[c][v], wsum[c][v]);
      }
  }
  const bool central_high = (kh == 0);
  double gh = 1.0;
#pragma unroll
  for (int k = 0; k < NS; ++k)
    if (k == kh) gh = single ? 1.0 : sc.lin_w[k];
  const double inv_gh = 1.0 / gh;
  const double hi_factor = (central_high && cweno) ? inv_gh : 1.0;
  double lo0[CLO][NVARS];
  double is0[NVARS];
#pragma unroll
  for (int v = 0; v < NVARS; ++v) is0[v] = 0.0;
#pragma unroll
  for (int c = 0; c < CLO; ++c)
#pragma unroll
    for (int v = 0; v < NVARS; ++v) {
      const double raw = coef[((1 + c) * NVARS + v) * TILE + lane];
      lo0[c][v] = (central_high && cweno) ? (raw - corr[c][v]) * inv_gh : raw;
      is0[v] += lo0[c][v] * lo0[c][v];
    }
#pragma unroll 1
  for (int c = CLO; c < CHI; ++c)
#pragma unroll
    for (int v = 0; v < NVARS; ++v) {
      const double h = coef[((1 + c) * NVARS + v) * TILE + lane] * hi_factor;
      is0[v] += h * h;
    }
  double is_max0 = is0[0];
#pragma unroll
  for (int v = 1; v < NVARS; ++v) is_max0 = ref_max(is_max0, is0[v]);
  const double alpha0 = nonlinear_weight(is_max0, single ? 1.0 : sc.lin_w[0]);
  al_sum += alpha0;
  double alpha_h = alpha0;
#pragma unroll
  for (int c = 0; c < CLO; ++c)
#pragma unroll
    for (int v = 0; v < NVARS; ++v) wsum[c][v] = fma(alpha0, lo0[c][v], wsum[c][v]);
  if (kh >= 1) {  // a one-sided stencil is the highest-order one (next to boundaries): it takes the correction
    double keep[CLO][NVARS];  // its raw polynomial, read back from shared memory
    double is_max = 0.0;
#pragma unroll
    for (int v = 0; v < NVARS; ++v) {
      double beta = 0.0;
#pragma unroll
      for (int c = 0; c < CLO; ++c) {
        const double raw = pk[(((kh - 1) * CLO + c) * NVARS + v) * TILE + lane];
        keep[c][v] = inv_gh * (raw - fma(sc.lin_w[0], lo0[c][v], corr[c][v]));
        beta += keep[c][v] * keep[c][v];
      }
      is_max = (v == 0) ? beta : ref_max(is_max, beta);
    }
    alpha_h = nonlinear_weight(is_max, gh);
    al_sum += alpha_h;
#pragma unroll
    for (int c = 0; c < CLO; ++c)
#pragma unroll
      for (int v = 0; v < NVARS; ++v) wsum[c][v] = fma(alpha_h, keep[c][v], wsum[c][v]);
  }
  const double inv_tot = fast_rcp(al_sum);
  double g_others = 0.0;
#pragma unroll
  for (int k = 0; k < NS; ++k)
    if (k != kh && k < n_eff) g_others += sc.lin_w[k];
  const double w_hi = alpha0 * inv_tot * hi_factor;  // factor of the raw high coefficients in the hybridised polynomial
  __syncthreads();  // every warp has read the raw coefficients: they may now be overwritten in place
  if (warp == 0) {
#pragma unroll
    for (int v = 0; v < NVARS; ++v) {
      const double a0h = (kh >= 0) ? inv_gh * (q0s[v] - g_others * q0s[v]) : q0s[v];
      coef[v * TILE + lane] = (alpha_h * a0h + (al_sum - alpha_h) * q0s[v]) * inv_tot;
    }
#pragma unroll
    for (int c = 0; c < CLO; ++c)
#pragma unroll
      for (int v = 0; v < NVARS; ++v) coef[((1 + c) * NVARS + v) * TILE + lane] = wsum[c][v] * inv_tot;
  }
  for (int c = CLO + warp; c < CHI; c += COOP_WARPS)
#pragma unroll
    for (int v = 0; v < NVARS; ++v) coef[((1 + c) * NVARS + v) * TILE + lane] *= w_hi;
  __syncthreads();

  // ---- polynomial outputs (diagnostics, hand-over to source_kernel): coefficients in the scaled basis + scales --------
  if (P.poly != nullptr && active) {
    for (int i = warp; i < D; i += COOP_WARPS)
#pragma unroll
      for (int v = 0; v < NVARS; ++v) P.poly[(cell * P.n_poly_coef + i) * NVARS + v] = coef[(i * NVARS + v) * TILE + lane];
    if (warp == 0) {
#pragma unroll
      for (int v = 0; v < NVARS; ++v) P.poly_scale[cell * NVARS + v] = scale[v];
    }
  }
  if (P.poly_tile != nullptr) {
    double *pt = P.poly_tile + tile * ((D + 1) * NVARS * TILE) + lane;
    for (int i = warp; i < D; i += COOP_WARPS)
#pragma unroll
      for (int v = 0; v < NVARS; ++v) pt[(i * NVARS + v) * TILE] = coef[(i * NVARS + v) * TILE + lane];
    if (warp == 0) {
#pragma unroll
      for (int v = 0; v < NVARS; ++v) pt[(D * NVARS + v) * TILE] = scale[v];
    }
  }

  // ---- traces of face `warp` at its Gauss points (flux_loop.hpp:131-149) ------------------------------------------------
  if (warp < F) {
    const int k = warp;
    const double *geo = reinterpret_cast<const double *>(rec + cfg.off_geo) + lane;
    const std::uint32_t *g32 = reinterpret_cast<const std::uint32_t *>(rec + cfg.off_geo + GEO_DOUBLES * TILE * 8) + lane;
    const std::uint32_t fref = active ? g32[k * TILE] : 0u;
    if (fref & FREF_TRACE) {
      const std::uint32_t slots = (g32[F * TILE] >> (8 * k)) & 0xFFu;
      double fv[ND][ND];
#pragma unroll
      for (int r = 0; r < ND; ++r) {
        const int s = (slots >> (2 * r)) & 3;
#pragma unroll
        for (int d = 0; d < ND; ++d) fv[r][d] = geo[(s * ND + d) * TILE];
      }
      double xc[ND], cmom[D];
#pragma unroll
      for (int d = 0; d < ND; ++d) xc[d] = geo[(F * ND + d) * TILE];
      const double inv_len = geo[(F * ND + ND) * TILE];
#pragma unroll
      for (int i = 0; i < D; ++i) cmom[i] = (i >= 3) ? geo[(F * ND + ND + 1 + (i - 3)) * TILE] : 0.0;
      const std::int64_t blk = (std::int64_t)(fref & FREF_EDGE_MASK) * 2 + ((fref & FREF_SIDE) ? 1 : 0);
      double *tr = P.trace + blk * (sc.q_f * NVARS);
      for (int q = 0; q < sc.q_f; ++q) {
        double xs[3] = {0.0, 0.0, 0.0};
#pragma unroll
        for (int d = 0; d < ND; ++d) {
          const double x = (ND == 2) ? sc.face_bary[q][0] * fv[0][d] + sc.face_bary[q][1] * fv[1][d]
                                     : fv[0][d] * sc.face_bary[q][0] + fv[1][d] * sc.face_bary[q][1] +
                                           fv[ND - 1][d] * sc.face_bary[q][2];
          xs[d] = (x - xc[d]) * inv_len;
        }
        double mono[D];
        PolyEval<ND, DEG_HI>::monomials(xs[0], xs[1], xs[2], cmom, mono);
#pragma unroll
        for (int v = 0; v < NVARS; ++v) {
          double s = coef[v * TILE + lane] * scale[v];
#pragma unroll
          for (int i = 1; i < D; ++i) s = fma(coef[(i * NVARS + v) * TILE + lane] * scale[v], mono[i], s);
          tr[q * NVARS + v] = s;  // WB: the perturbation; K2 adds the equilibrium background (eq_bg)
        }
      }
    }
  }
}

/// Fills the launch configuration; returns false if the context carries no tile records for this scheme shape.
template <int ND, int DEG_HI>
bool coop_config(const DevicePlan &P, const SchemeConst &sc, CoopCfg &c) {
  constexpr int NS = ND + 2, D = dof_of(DEG_HI, ND), CLO = dof_of(1, ND) - 1;
  if (P.rec2 == nullptr || sc.n_stencils != NS || sc.ncoef[0] != D - 1) return false;
  for (int k = 1; k < NS; ++k)
    if (sc.ncoef[k] != CLO) return false;
  const TileRecLayout L = tile_rec_layout(sc, ND, D, P.rec2_cap);
  if (L.rec_bytes != P.rec2_bytes) return false;
  c.cap = L.cap;
  c.off_list = L.off_list;
  c.off_lidx = L.off_lidx;
  c.off_wlo = L.off_wlo;
  c.off_whi = L.off_whi;
  c.off_geo = L.off_geo;
  c.rec_bytes = L.rec_bytes;
  c.rows0 = sc.rows_max[0];
  c.rows_total = L.rows;
  int r = 0, b = 0;
  for (int k = 1; k < NS; ++k) {
    c.rows_lo[k] = sc.rows_max[k];
    c.row0_lo[k] = r;
    c.wlo_off[k] = b;
    r += sc.rows_max[k];
    b += sc.rows_max[k] * sc.ncoef[k] * TILE * 8;
  }
  c.s_table = 0;
  c.s_coef = (L.cap * NVARS * 8 + 127) / 128 * 128;
  c.s_pk = c.s_coef + D * NVARS * TILE * 8;
  c.smem_bytes = c.s_pk + (NS - 1) * CLO * NVARS * TILE * 8;
  return true;
}

template <int ND, int DEG_HI>
int launch_coop(const ReconArgs &args, const SchemeConst &sc, std::int64_t n_tiles, cudaStream_t stream) {
  CoopCfg cfg;
  if (!coop_config<ND, DEG_HI>(args.plan, sc, cfg)) return 1;
  int dev = 0, optin = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  if (cfg.smem_bytes > optin) return 1;
  auto go = [&](auto kern) {
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, cfg.smem_bytes);
    kern<<<(unsigned)n_tiles, COOP_WARPS * 32, (size_t)cfg.smem_bytes, stream>>>(args, sc, cfg);
  };
  const bool wb = sc.well_balanced != 0;
  if (args.plan.rec2_cap <= 256) {
    if (wb)
      go(recon_coop_kernel<ND, DEG_HI, std::uint8_t, true>);
    else
      go(recon_coop_kernel<ND, DEG_HI, std::uint8_t, false>);
  } else {
    if (wb)
      go(recon_coop_kernel<ND, DEG_HI, std::uint16_t, true>);
    else
      go(recon_coop_kernel<ND, DEG_HI, std::uint16_t, false>);
  }
  return 0;
}

}  // namespace zfvm
