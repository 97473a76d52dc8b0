// Instantiation helper: one translation unit per (n_dims, high degree) keeps nvcc compile times parallel.
//
// Families of the shape the reference's experiments use (one leading stencil of degree DEG_HI, n_dims + 1 stencils of
// degree 1) come here and run on tile records (DevicePlan::rec2):
//   * recon_tile_kernel  (recon_tile.cuh)   when it is instantiated for the scheme's stencil sizes and face rule --
//     the reference's parameter sets up to 9 central coefficients (2D orders 2-4, 3D orders 2-3);
//   * recon_coop_kernel  (recon_coop.cuh)   otherwise: 3D order 4, 2D order 5, any other stencil sizes.
// Everything else runs recon_generic.cu.
#pragma once
#include <algorithm>

#include "kernels.hpp"
#include "recon.cuh"
#include "recon_coop.cuh"
#include "recon_tile.cuh"
#include "source.cuh"

#include <cstdlib>

namespace zfvm {

/// Face quadrature sizes the tile kernel is compiled for (edge rules of degree 2-5, triangle rules of degree 2-3).
constexpr bool tile_kernel_qf(int nd, int q_f) { return nd == 2 ? (q_f == 2 || q_f == 3) : (q_f == 3 || q_f == 4); }

template <int ND, int DEG_HI, int DEG_LO, int NS, int RM0, int RLO, int QF>
int launch_tile(const ReconArgs &args, const SchemeConst &sc, std::int64_t n_tiles, cudaStream_t stream, bool wb = false) {
  using T = TileTraits<ND, DEG_HI, DEG_LO, NS, RM0, RLO, QF>;
  int dev = 0, optin = 0, n_sm = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
  // one CTA per SM; as many warps (= tiles in flight) as fit, at most 8
  TileCfg cfg;
  if (!tile_config<ND, DEG_HI, DEG_LO, NS, RM0, RLO, QF>(args.plan, sc, optin, cfg)) return 1;
  int wpc = std::min(tile_max_warps(ND, DEG_HI), optin / cfg.warp_bytes);
  if (const char *e = std::getenv("ZFVM_TILE_WARPS")) wpc = std::max(1, std::min(wpc, std::atoi(e)));
  if (wpc < 1) return 1;
  cfg.prof = tile_prof_buffer();
  if (const char *e = std::getenv("ZFVM_TILE_EVICT")) cfg.evict_normal = (e[0] == 'n') ? 1 : 0;
  if (const char *e = std::getenv("ZFVM_TILE_L2_AHEAD")) cfg.l2_ahead = std::max(0, std::atoi(e));
  if (const char *e = std::getenv("ZFVM_TILE_L2_WHOLE")) cfg.l2_whole = (e[0] == '1') ? 1 : 0;
  const int smem_bytes = cfg.warp_bytes * wpc;
  const char *e_ctas = std::getenv("ZFVM_STREAM_MAX_CTAS");
  const int max_ctas = e_ctas ? std::max(1, std::atoi(e_ctas)) : (1 << 30);
  const unsigned grid = (unsigned)std::min<std::int64_t>((n_tiles + wpc - 1) / wpc, std::min(n_sm, max_ctas));
  auto go = [&](auto kern) {
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
    kern<<<grid, 32 * wpc, (size_t)smem_bytes, stream>>>(args, sc, cfg);
  };
  if (wb) {  // well-balanced: equilibrium tables in, background added to the traces (no phase-timer instantiation)
    if (args.plan.rec2_cap <= 256)
      go(recon_tile_kernel<ND, DEG_HI, DEG_LO, NS, RM0, RLO, QF, std::uint8_t, false, true>);
    else
      go(recon_tile_kernel<ND, DEG_HI, DEG_LO, NS, RM0, RLO, QF, std::uint16_t, false, true>);
  } else if (cfg.prof != nullptr) {  // ZFVM_TILE_PROF=1: instantiation with the phase timers
    if (args.plan.rec2_cap <= 256)
      go(recon_tile_kernel<ND, DEG_HI, DEG_LO, NS, RM0, RLO, QF, std::uint8_t, true>);
    else
      go(recon_tile_kernel<ND, DEG_HI, DEG_LO, NS, RM0, RLO, QF, std::uint16_t, true>);
  } else if (args.plan.rec2_cap <= 256) {
    go(recon_tile_kernel<ND, DEG_HI, DEG_LO, NS, RM0, RLO, QF, std::uint8_t, false>);
  } else {
    go(recon_tile_kernel<ND, DEG_HI, DEG_LO, NS, RM0, RLO, QF, std::uint16_t, false>);
  }
  (void)sizeof(T);
  return 0;
}

/// The tile kernel keeps (dof - 1) x 5 accumulators of the central stencil in registers: compiled up to 9 coefficients.
constexpr bool tile_kernel_enabled(int nd, int deg_hi) { return dof_of(deg_hi, nd) - 1 <= 9; }

/// The equilibrium kernels are independent of dimension and degrees: one definition (recon_dispatch.cu).
template <int POWN>
void launch_eq_solve(const DevicePlan &plan, const SchemeConst &sc, const double *state, const std::int32_t *tile_list,
                     std::int64_t n_tiles, unsigned grid, cudaStream_t stream);

/// E0: recompute_equilibrium's verdict per cell (steps_per_recompute != 1 only; plan.eq_flag is null otherwise).
void launch_eq_decide(const DevicePlan &plan, const SchemeConst &sc, const double *state, const std::int32_t *tile_list,
                      std::int64_t n_tiles, cudaStream_t stream);

/// E2 + E3 for tile records (members through the tile's row list; equilibrium at the face Gauss points).
template <int POWN>
void launch_eq_tile(const DevicePlan &plan, const SchemeConst &sc, const std::int32_t *tile_list, std::int64_t n_tiles,
                    cudaStream_t stream);

template <int ND, int DEG_HI, int NS, int RM0, int RLO>
int launch_recon_variants(const DevicePlan &plan, const SchemeConst &sc, const double *state,
                          const std::int32_t *tile_list, std::int64_t n_tiles, cudaStream_t stream) {
  ReconArgs args;
  args.plan = plan;
  args.state = state;
  args.tile_list = tile_list;
  args.n_tiles_launch = n_tiles;
  if (n_tiles <= 0) return 0;
  if (plan.rec2 == nullptr) return 1;
  const bool wb = sc.well_balanced != 0;
  if (plan.eq_flag != nullptr) launch_eq_decide(plan, sc, state, tile_list, n_tiles, stream);
  if (wb) {  // E1 equilibrium solve, E2 its averages over the stencil members, E3 its values at the face points
    const unsigned g1 = (unsigned)((n_tiles * TILE + 255) / 256);
    switch (sc.eos_pow_n) {
      case 2: launch_eq_solve<2>(plan, sc, state, tile_list, n_tiles, g1, stream); launch_eq_tile<2>(plan, sc, tile_list, n_tiles, stream); break;
      case 3: launch_eq_solve<3>(plan, sc, state, tile_list, n_tiles, g1, stream); launch_eq_tile<3>(plan, sc, tile_list, n_tiles, stream); break;
      case 5: launch_eq_solve<5>(plan, sc, state, tile_list, n_tiles, g1, stream); launch_eq_tile<5>(plan, sc, tile_list, n_tiles, stream); break;
      default: launch_eq_solve<0>(plan, sc, state, tile_list, n_tiles, g1, stream); launch_eq_tile<0>(plan, sc, tile_list, n_tiles, stream); break;
    }
  }
  int rc = 1;
  // tests: ZFVM_RECON=coop runs the cooperative kernel where the tile kernel exists too
  const char *e_recon = std::getenv("ZFVM_RECON");
  const bool force_coop = e_recon != nullptr && e_recon[0] == 'c';
  if constexpr (tile_kernel_enabled(ND, DEG_HI)) {
    if (!force_coop) {
      if constexpr (ND == 2) {
        if (sc.q_f == 2) rc = launch_tile<ND, DEG_HI, 1, NS, RM0, RLO, 2>(args, sc, n_tiles, stream, wb);
        if (sc.q_f == 3) rc = launch_tile<ND, DEG_HI, 1, NS, RM0, RLO, 3>(args, sc, n_tiles, stream, wb);
      } else {
        if (sc.q_f == 3) rc = launch_tile<ND, DEG_HI, 1, NS, RM0, RLO, 3>(args, sc, n_tiles, stream, wb);
        if (sc.q_f == 4) rc = launch_tile<ND, DEG_HI, 1, NS, RM0, RLO, 4>(args, sc, n_tiles, stream, wb);
      }
    }
  }
  if (rc != 0) rc = launch_coop<ND, DEG_HI>(args, sc, n_tiles, stream);
  // cell-local source terms (gravity, heating) from the polynomial K1 stored
  if (rc == 0 && sc.has_gravity) {
    if (plan.poly_tile == nullptr) return 1;
    const unsigned gs = (unsigned)((n_tiles + 3) / 4);
    if (!wb) {
      source_kernel<ND, DEG_HI, false, 0><<<gs, 128, 0, stream>>>(args, sc);
    } else {
      switch (sc.eos_pow_n) {
        case 2: source_kernel<ND, DEG_HI, true, 2><<<gs, 128, 0, stream>>>(args, sc); break;
        case 3: source_kernel<ND, DEG_HI, true, 3><<<gs, 128, 0, stream>>>(args, sc); break;
        case 5: source_kernel<ND, DEG_HI, true, 5><<<gs, 128, 0, stream>>>(args, sc); break;
        default: source_kernel<ND, DEG_HI, true, 0><<<gs, 128, 0, stream>>>(args, sc); break;
      }
    }
  }
  return rc;
}

#define ZFVM_DECLARE_RECON(ND, DEG_HI)                                                                   \
  int launch_recon_##ND##d_deg##DEG_HI(const DevicePlan &plan, const SchemeConst &sc, int deg_lo,       \
                                       const double *state, const std::int32_t *tile_list,              \
                                       std::int64_t n_tiles, cudaStream_t stream)

// RM0 / RLO: rows (stencil size - 1) of the central / one-sided stencils of the reference's parameter set for this
// order: the sizes the tile kernel is instantiated for; other sizes run the cooperative kernel.
#define ZFVM_DEFINE_RECON(ND, DEG_HI, RM0, RLO)                                                          \
  ZFVM_DECLARE_RECON(ND, DEG_HI) {                                                                       \
    if (sc.n_stencils != ND + 2 || deg_lo != 1) return 1;                                                \
    return launch_recon_variants<ND, DEG_HI, ND + 2, RM0, RLO>(plan, sc, state, tile_list, n_tiles, stream); \
  }

ZFVM_DECLARE_RECON(2, 1);
ZFVM_DECLARE_RECON(2, 2);
ZFVM_DECLARE_RECON(2, 3);
ZFVM_DECLARE_RECON(2, 4);
ZFVM_DECLARE_RECON(3, 1);
ZFVM_DECLARE_RECON(3, 2);
ZFVM_DECLARE_RECON(3, 3);

}  // namespace zfvm
