// Instantiation helper: one translation unit per (n_dims, high degree) keeps nvcc compile times parallel.
#pragma once
#include <algorithm>

#include "kernels.hpp"
#include "recon.cuh"
#include "recon_stream.cuh"
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
  // one CTA per SM; as many warps (= tiles in flight) as fit with `slots` ring slots each, at most 8
  int slots = 3;  // measured on B200: 8 warps x 3 slots beat 7 warps x 4 slots (and 8 x 2)
  if (const char *e = std::getenv("ZFVM_TILE_SLOTS")) slots = std::max(2, std::min(13, std::atoi(e)));
  TileCfg cfg;
  if (!tile_config<ND, DEG_HI, DEG_LO, NS, RM0, RLO, QF>(args.plan, sc, 1 << 20, slots, cfg)) return 1;
  int wpc = std::min(TILE_MAX_WARPS, optin / cfg.warp_bytes);
  if (const char *e = std::getenv("ZFVM_TILE_WARPS")) wpc = std::max(1, std::min(wpc, std::atoi(e)));
  if (wpc < 1) return 1;
  // spend what is left on deeper rings
  if (!std::getenv("ZFVM_TILE_SLOTS")) tile_config<ND, DEG_HI, DEG_LO, NS, RM0, RLO, QF>(args.plan, sc, (optin / wpc) / 128 * 128, 13, cfg);
  cfg.prof = tile_prof_buffer();
  cfg.l2_ahead = 0;
  if (const char *e = std::getenv("ZFVM_TILE_L2_AHEAD")) cfg.l2_ahead = std::max(0, std::atoi(e));
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

/// The two equilibrium kernels are independent of dimension and degrees: one definition (recon_dispatch.cu).
template <int POWN>
void launch_eq_solve(const DevicePlan &plan, const SchemeConst &sc, const double *state, const std::int32_t *tile_list,
                     std::int64_t n_tiles, unsigned grid, cudaStream_t stream);
template <int POWN>
void launch_eq_member(const DevicePlan &plan, const SchemeConst &sc, const std::int32_t *tile_list, std::int64_t n_tiles,
                      unsigned grid, cudaStream_t stream);

/// E2 + E3 for tile records (members through the tile's row list; equilibrium at the face Gauss points).
template <int POWN>
void launch_eq_tile(const DevicePlan &plan, const SchemeConst &sc, const std::int32_t *tile_list, std::int64_t n_tiles,
                    cudaStream_t stream);

template <int ND, int DEG_HI, int DEG_LO, int NS, int RM0, int RLO>
int launch_recon_variants(const DevicePlan &plan, const SchemeConst &sc, const double *state,
                          const std::int32_t *tile_list, std::int64_t n_tiles, cudaStream_t stream) {
  ReconArgs args;
  args.plan = plan;
  args.state = state;
  args.tile_list = tile_list;
  args.n_tiles_launch = n_tiles;
  if (n_tiles <= 0) return 0;
  if constexpr (DEG_HI >= 1 && tile_kernel_enabled(ND, DEG_HI)) {
    // tile kernel (recon_tile.cuh): used whenever the context carries tile records (zfvm_create decides)
    if (plan.rec2 != nullptr) {
      int rc = 1;
      const bool wb = sc.well_balanced != 0;
      if (wb) {  // E1 equilibrium solve, E2 its averages over the stencil members, E3 its values at the face points
        const unsigned g1 = (unsigned)((n_tiles * TILE + 255) / 256);
        switch (sc.eos_pow_n) {
          case 2: launch_eq_solve<2>(plan, sc, state, tile_list, n_tiles, g1, stream); launch_eq_tile<2>(plan, sc, tile_list, n_tiles, stream); break;
          case 3: launch_eq_solve<3>(plan, sc, state, tile_list, n_tiles, g1, stream); launch_eq_tile<3>(plan, sc, tile_list, n_tiles, stream); break;
          case 5: launch_eq_solve<5>(plan, sc, state, tile_list, n_tiles, g1, stream); launch_eq_tile<5>(plan, sc, tile_list, n_tiles, stream); break;
          default: launch_eq_solve<0>(plan, sc, state, tile_list, n_tiles, g1, stream); launch_eq_tile<0>(plan, sc, tile_list, n_tiles, stream); break;
        }
      }
      if constexpr (ND == 2) {
        if (sc.q_f == 2) rc = launch_tile<ND, DEG_HI, DEG_LO, NS, RM0, RLO, 2>(args, sc, n_tiles, stream, wb);
        if (sc.q_f == 3) rc = launch_tile<ND, DEG_HI, DEG_LO, NS, RM0, RLO, 3>(args, sc, n_tiles, stream, wb);
      } else {
        if (sc.q_f == 3) rc = launch_tile<ND, DEG_HI, DEG_LO, NS, RM0, RLO, 3>(args, sc, n_tiles, stream, wb);
        if (sc.q_f == 4) rc = launch_tile<ND, DEG_HI, DEG_LO, NS, RM0, RLO, 4>(args, sc, n_tiles, stream, wb);
      }
      // cell-local source terms (gravity, heating) from the polynomial the tile kernel stored
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
  }
  if constexpr (DEG_HI >= 1) {
    // streaming kernel (recon_stream.cuh); ZFVM_RECON=v1 selects the thread-per-cell kernel for comparisons
    const char *e_recon = std::getenv("ZFVM_RECON");
    const bool force_v1 = e_recon && e_recon[0] == 'v' && e_recon[1] == '1';
    if (!sc.well_balanced && !sc.has_gravity && !force_v1) {
      int dev = 0, optin = 0, n_sm = 0;
      cudaGetDevice(&dev);
      cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
      cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
      int budget = optin;
      if (const char *e = std::getenv("ZFVM_STREAM_SMEM_KB")) budget = std::min(optin, std::atoi(e) * 1024);
      StreamCfg cfg;
      if (stream_config<ND, DEG_HI, DEG_LO, NS, RM0, RLO>(plan, sc, budget, cfg)) {
        auto kern = recon_stream_kernel<ND, DEG_HI, DEG_LO, NS, RM0, RLO>;
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, cfg.total_bytes);
        // one persistent CTA per SM; ZFVM_STREAM_MAX_CTAS lowers the count (tests use it to put many tiles on a CTA)
        const char *e_ctas = std::getenv("ZFVM_STREAM_MAX_CTAS");
        const int max_ctas = e_ctas ? std::max(1, std::atoi(e_ctas)) : (1 << 30);
        const unsigned grid = (unsigned)std::min<std::int64_t>(n_tiles, std::min(n_sm, max_ctas));
        const int block = 32 * StreamTraits<ND, DEG_HI, DEG_LO, NS, RM0, RLO>::N_WARPS;
        kern<<<grid, block, (size_t)cfg.total_bytes, stream>>>(args, sc, cfg);
        return 0;
      }
    }
  }
  const int block = 128;  // 4 tiles per CTA
  const unsigned grid = (unsigned)((n_tiles + 3) / 4);
  if (sc.well_balanced) {
    // E1 (equilibrium solve per cell) and E2 (its averages over every stencil member) run ahead of the reconstruction;
    // the isentropic EOS power x^(n/2) is a compile-time constant for gamma = 2, 5/3, 7/5 (n = 2, 3, 5)
    const unsigned g1 = (unsigned)((n_tiles * TILE + 255) / 256);
    const unsigned g2 = (unsigned)((n_tiles * plan.eq_rows + 7) / 8);
#define ZFVM_WB_LAUNCH(POWN)                                                                              \
  {                                                                                                       \
    launch_eq_solve<POWN>(plan, sc, state, tile_list, n_tiles, g1, stream);                                                 \
    launch_eq_member<POWN>(plan, sc, tile_list, n_tiles, g2, stream);                                     \
    recon_kernel<ND, DEG_HI, DEG_LO, NS, RV_WELL_BALANCED, POWN><<<grid, block, 0, stream>>>(args, sc);   \
  }
    switch (sc.eos_pow_n) {
      case 2: ZFVM_WB_LAUNCH(2) break;
      case 3: ZFVM_WB_LAUNCH(3) break;
      case 5: ZFVM_WB_LAUNCH(5) break;
      default: ZFVM_WB_LAUNCH(0) break;
    }
#undef ZFVM_WB_LAUNCH
  } else if (sc.has_gravity)
    recon_kernel<ND, DEG_HI, DEG_LO, NS, RV_GRAVITY><<<grid, block, 0, stream>>>(args, sc);
  else
    recon_kernel<ND, DEG_HI, DEG_LO, NS, RV_PLAIN><<<grid, block, 0, stream>>>(args, sc);
  return 0;
}

#define ZFVM_DECLARE_RECON(ND, DEG_HI)                                                                   \
  int launch_recon_##ND##d_deg##DEG_HI(const DevicePlan &plan, const SchemeConst &sc, int deg_lo,       \
                                       const double *state, const std::int32_t *tile_list,              \
                                       std::int64_t n_tiles, cudaStream_t stream)

// RM0 / RLO: rows (stencil size - 1) of the central / one-sided stencils of the reference's parameter set for this
// order; the streaming kernel is compiled for exactly these, other sizes run the thread-per-cell kernel.
#define ZFVM_DEFINE_RECON(ND, DEG_HI, RM0, RLO)                                                          \
  bool recon_tile_sizes_##ND##d_deg##DEG_HI(const SchemeConst &sc) {                                     \
    if (!tile_kernel_enabled(ND, DEG_HI) || DEG_HI < 1 || sc.n_stencils != ND + 2) return false;          \
    if (!tile_kernel_qf(ND, sc.q_f)) return false;                                                       \
    if (sc.rows_max[0] != RM0) return false;                                                             \
    for (int k = 1; k < sc.n_stencils; ++k)                                                              \
      if (sc.rows_max[k] != RLO) return false;                                                           \
    return true;                                                                                         \
  }                                                                                                      \
  ZFVM_DECLARE_RECON(ND, DEG_HI) {                                                                       \
    if (sc.n_stencils != ND + 2) return 1;                                                               \
    if (deg_lo == 1 || (DEG_HI == 0 && deg_lo == 0))                                                     \
      return launch_recon_variants<ND, DEG_HI, (DEG_HI >= 1 ? 1 : 0), ND + 2, RM0, RLO>(                 \
          plan, sc, state, tile_list, n_tiles, stream);                                                  \
    return 1;                                                                                            \
  }

#define ZFVM_DECLARE_TILE_SIZES(ND, DEG_HI) bool recon_tile_sizes_##ND##d_deg##DEG_HI(const SchemeConst &sc)
ZFVM_DECLARE_TILE_SIZES(2, 1);
ZFVM_DECLARE_TILE_SIZES(2, 2);
ZFVM_DECLARE_TILE_SIZES(2, 3);
ZFVM_DECLARE_TILE_SIZES(2, 4);
ZFVM_DECLARE_TILE_SIZES(3, 1);
ZFVM_DECLARE_TILE_SIZES(3, 2);
ZFVM_DECLARE_TILE_SIZES(3, 3);
ZFVM_DECLARE_RECON(2, 1);
ZFVM_DECLARE_RECON(2, 2);
ZFVM_DECLARE_RECON(2, 3);
ZFVM_DECLARE_RECON(2, 4);
ZFVM_DECLARE_RECON(3, 1);
ZFVM_DECLARE_RECON(3, 2);
ZFVM_DECLARE_RECON(3, 3);

}  // namespace zfvm
