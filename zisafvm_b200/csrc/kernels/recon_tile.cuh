// K1, tile form: one warp owns one tile of 32 cells, one thread owns one cell; warps never talk to each other.
//
// Same arithmetic as recon.cuh (EulerGlobalReconstruction::compute, LocalReconstruction::compute,
// HybridWENO::compute_polys_impl / eno_hybridize, CWENO_AO::reconstruct_impl, rc(i)(x) at the face Gauss
// points; the reference lines are listed there).  What changes is how the bytes move:
//
//   * every table a tile needs is one contiguous *tile record* (layout below) that the warp streams itself
//     with TMA bulk copies (cp.async.bulk + mbarrier complete_tx, L2 evict-first) through a private ring of
//     small shared-memory slots (one one-sided stencil or a few central-stencil rows each).  After the warp
//     has consumed a slot, its lane 0 re-arms the barrier and issues the copy of the segment that is NSLOT
//     positions ahead: no producer warp, no "empty" barriers, and the bytes in flight do not depend on
//     registers.
//   * the neighbour states are not gathered through global indices.  The host lists the distinct cells a
//     tile's stencils read (~125-250 of them for 32 Hilbert-consecutive cells, instead of 32 x 34 gathers) and
//     stores 8/16-bit indices into that list; the warp copies those rows once into a shared-memory table
//     (cp.async, 8 bytes per lane, issued while the previous tile is still being evaluated) and every rhs
//     entry is a shared-memory read.
//   * hybridisation is streamed.  The non-linear weight of a stencil depends on that stencil's polynomial only
//     (alpha_k = gamma_k / (eps + IS_k^p), hybrid_weno.cpp:110-128), so a one-sided polynomial is folded into
//     sum_k alpha_k p_k and into the CWENO correction sum_k gamma_k p_k (cweno_ao.cpp:41-50) as soon as it is
//     complete, and the sums are normalised at the end: only the central stencil's accumulators and two
//     15-double sums live in registers, nothing is parked, and a warp needs ~28 KB of shared memory, so that
//     8 warps (tiles in flight) fit on an SM.  (Cells whose highest-order stencil is not the central one --
//     next to boundaries -- take a slower, per-lane-predicated instantiation of the same code.)
//   * the traces of a face are staged through shared memory and written as contiguous blocks.
//
// Tile record (sections multiples of 128 bytes; CAP = capacity of the row list, a multiple of 32):
//   | n_list u32, pad to 16 B | meta u64[32] | list i32[CAP] |                              list part
//   | lidx u8|u16 [ROWS][32] |                                                              index part
//   | W_1 f64[RLO][CLO][32] | .. | W_{NS-1} |                                               one-sided stencils
//   | W_0 f64[RM0][CHI][32] |                                                               central stencil
//   | vtx f64[F][ND][32] | centre f64[ND][32] | 1/len f64[32] | moments f64[D-3][32] | face_ref u32[F][32] |
//   | face_slots u32[32] (byte k: face k) |                                                 geometry
// lidx rows: the one-sided stencils' rows first, then the central stencil's; list[0..31] are the tile's own
// cells; lidx is 8-bit when CAP <= 256.
#pragma once
#include <type_traits>

#include "ptx.cuh"
#include "recon.cuh"

namespace zfvm {

constexpr int TILE_MAX_WARPS = 8;       // warps per CTA the kernel is compiled for (register budget 65536 / (32 * TILE_MAX_WARPS))
constexpr int TILE_SLOT_TARGET = 4608;  // bytes of a ring slot aimed at (two central rows of the 3D order-3 scheme)

template <int ND, int DEG_HI, int DEG_LO, int NS, int RM0, int RLO, int QF>
struct TileTraits {
  static constexpr int F = ND + 1;
  static constexpr int D = dof_of(DEG_HI, ND);
  static constexpr int CHI = D - 1;
  static constexpr int CLO = dof_of(DEG_LO, ND) - 1;
  static constexpr int NHI = CHI - CLO;
  static constexpr int ROWS = RM0 + (NS - 1) * RLO;
  static constexpr int HI_ROW_BYTES = CHI * TILE * 8;
  static constexpr int LO_ST_BYTES = RLO * CLO * TILE * 8;
  static constexpr int N_MOM = D > 3 ? D - 3 : 0;
  static constexpr int GEO_DOUBLES = F * ND + ND + 1 + N_MOM;
  static constexpr int GEO_BYTES = GEO_DOUBLES * TILE * 8 + F * TILE * 4 + TILE * 4;
  static constexpr int GEO_SECTION = (GEO_BYTES + 127) / 128 * 128;
  static constexpr int R_FIT = TILE_SLOT_TARGET / HI_ROW_BYTES < 1 ? 1 : TILE_SLOT_TARGET / HI_ROW_BYTES;
  static constexpr int R_HI = R_FIT > RM0 ? RM0 : R_FIT;    // central rows per segment
  static constexpr int SLOT_BYTES = LO_ST_BYTES > R_HI * HI_ROW_BYTES ? LO_ST_BYTES : R_HI * HI_ROW_BYTES;  // multiple of 256
  static constexpr int N_HI = (RM0 + R_HI - 1) / R_HI;
  static constexpr int R_TAIL = RM0 - (N_HI - 1) * R_HI;
  static constexpr int N_LO = NS - 1;                       // one one-sided stencil per segment
  static constexpr int GEO_ROWS_PER_SEG = SLOT_BYTES / (TILE * 8);
  static constexpr int N_GEO = (GEO_SECTION + SLOT_BYTES - 1) / SLOT_BYTES;
  static constexpr int GEO_TAIL_BYTES = GEO_SECTION - (N_GEO - 1) * SLOT_BYTES;
  static constexpr int N_SEG = N_LO + N_HI + N_GEO;
  static constexpr int Q_STAGE = 1;                         // Gauss points staged per pass (shared memory is what limits the warps per SM)
  static constexpr int CHUNK = Q_STAGE * NVARS;             // doubles per (cell, face) block written per pass
  static constexpr int STAGE_PITCH = CHUNK | 1;
  static constexpr int STAGE_BYTES = (TILE * STAGE_PITCH * 8 + 127) / 128 * 128;
};

struct TileCfg {
  // record
  int cap, list_bytes, off_list, off_lidx, lidx_bytes, off_wlo;
  std::int64_t rec_bytes;
  const char *rec;
  // shared memory of one warp
  int n_slots, s_list, s_lidx, s_table, s_ring, s_stage, warp_bytes;
  int l2_ahead;              // L2 prefetch distance behind the ring in bytes (0: off, -1 never used)
  int seg_bytes[64];         // bytes of segment s of a record's W | geometry stream
  unsigned long long *prof;  // optional phase timers of warp 0 (clock cycles): see TilePhase; null = off
};

enum TilePhase : int { TP_TABLE_WAIT = 0, TP_LO = 1, TP_HI = 2, TP_TABLE_ISSUE = 3, TP_HYBRID = 4, TP_GEO_WAIT = 5, TP_TRACE = 6, TP_SEG_WAIT = 7, TP_TILES = 8, TP_COUNT = 9 };

namespace ptx {
ZFVM_DEVICE void cp_async8(void *dst_smem, const void *src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(dst_smem)), "l"(src) : "memory");
}
/// expect_tx + TMA bulk copy global -> shared, both predicated on `pred` inside one asm block (no divergent branch).
ZFVM_DEVICE void bulk_g2s_if(bool pred, void *dst, const void *src, std::uint32_t bytes, std::uint64_t *bar,
                             std::uint64_t policy) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.u32 p, %5, 0;\n\t"
      "@p mbarrier.arrive.expect_tx.shared::cta.b64 _, [%3], %2;\n\t"
      "@p cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;\n\t}"
      ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy), "r"((std::uint32_t)pred)
      : "memory");
}
/// L2 prefetch of a global range (no shared-memory destination, no completion tracking), predicated like bulk_g2s_if.
ZFVM_DEVICE void bulk_prefetch_l2_if(bool pred, const void *src, std::uint32_t bytes) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.u32 p, %2, 0;\n\t"
      "@p cp.async.bulk.prefetch.L2.global [%0], %1;\n\t}" ::"l"(src),
      "r"(bytes), "r"((std::uint32_t)pred)
      : "memory");
}
ZFVM_DEVICE void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
ZFVM_DEVICE void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
}  // namespace ptx

// WB (well-balanced runs): the equilibrium kernels (equilibrium.cuh: E1, E2, E3) have written, per tile, the cell
// averages of each cell's local equilibrium over its stencil members in lidx row order plus one row for the cell
// itself (eq_avg): the rhs subtracts them (local_reconstruction.hpp:109-116).  The traces written here are those of
// the perturbation; the equilibrium background at the face Gauss points (local_reconstruction.hpp:149-163) is added by
// the face-flux kernel from E3's table (eq_bg).
template <int ND, int DEG_HI, int DEG_LO, int NS, int RM0, int RLO, int QF, typename LIDX, bool PROF, bool WB = false>
__global__ void __launch_bounds__(TILE_MAX_WARPS * 32, 1)
    recon_tile_kernel(const __grid_constant__ ReconArgs args, const __grid_constant__ SchemeConst sc,
                      const __grid_constant__ TileCfg cfg) {
  using T = TileTraits<ND, DEG_HI, DEG_LO, NS, RM0, RLO, QF>;
  constexpr int F = T::F, D = T::D, CHI = T::CHI, CLO = T::CLO, NHI = T::NHI;
  constexpr int N_HI = T::N_HI, R_HI = T::R_HI, R_TAIL = T::R_TAIL, N_LO = T::N_LO, N_GEO = T::N_GEO, N_SEG = T::N_SEG;
  const DevicePlan &P = args.plan;

  extern __shared__ __align__(128) unsigned char smem_all[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned char *smem = smem_all + (size_t)warp * cfg.warp_bytes;
  std::uint64_t *list_full = reinterpret_cast<std::uint64_t *>(smem);  // [1] (+1 unused)
  std::uint64_t *lidx_full = list_full + 2;                            // [1]
  std::uint64_t *seg_full = list_full + 3;                             // [n_slots]
  unsigned char *list_base = smem + cfg.s_list;
  const LIDX *lidx = reinterpret_cast<const LIDX *>(smem + cfg.s_lidx) + lane;
  double *table = reinterpret_cast<double *>(smem + cfg.s_table);
  unsigned char *ring = smem + cfg.s_ring;
  double *stage = reinterpret_cast<double *>(smem + cfg.s_stage);
  const int NSLOT = cfg.n_slots;

  const std::int64_t n_launch = args.n_tiles_launch;
  const std::int64_t first = (std::int64_t)blockIdx.x * (blockDim.x >> 5) + warp;
  const std::int64_t stride = (std::int64_t)gridDim.x * (blockDim.x >> 5);
  auto has_tile = [&](int m) { return first + (std::int64_t)m * stride < n_launch; };
  // Tile numbers come from a list in global memory when only some tiles are reconstructed (ghost tiles skipped,
  // interior / exterior lists of a decomposed run).  Every use of a tile number sits in front of a copy that the
  // warp is about to wait for, so the list entry of tile m + 2 is fetched at the start of tile m (t_nxt2) and has a
  // whole tile's time to arrive: in the first version these loads were 12 % of all stall samples (long scoreboard).
  auto load_tile_no = [&](int m) -> std::int32_t {
    const std::int64_t idx = first + (std::int64_t)m * stride;
    if (idx >= n_launch) return 0;
    return args.tile_list ? __ldg(args.tile_list + idx) : (std::int32_t)idx;
  };
  if (!has_tile(0)) return;
  std::int32_t t_cur = load_tile_no(0), t_nxt = load_tile_no(1), t_nxt2 = load_tile_no(2);
  int m_cur = 0;  // the tile the warp is working on: tile_of is only ever asked for m_cur and m_cur + 1
  auto tile_of = [&](int m) -> std::int64_t { return (std::int64_t)(m == m_cur ? t_cur : t_nxt); };
  const bool prof = PROF && cfg.prof != nullptr && first == 0;  // PROF = false: the timers compile away
  long long t_mark = 0, t_segwait = 0;
  auto mark = [&](int phase) {
    if (prof) {
      const long long now = clock64();
      if (lane == 0 && phase >= 0) atomicAdd(cfg.prof + phase, (unsigned long long)(now - t_mark));
      t_mark = now;
    }
  };

  if (lane == 0) {
    ptx::mbar_init(&list_full[0], 1);
    ptx::mbar_init(&lidx_full[0], 1);
    for (int s = 0; s < NSLOT; ++s) ptx::mbar_init(&seg_full[s], 1);
    ptx::fence_barrier_init();
  }
  __syncwarp();

  const std::uint64_t pol = ptx::policy_evict_first();

  // ---- issue side (lane 0 only) ---------------------------------------------------------------------
  auto issue_list = [&](int m) {  // n_list | meta | list of tile m (one buffer: see the tile loop for its life time)
    if (lane == 0 && has_tile(m)) {
      ptx::mbar_expect_tx(&list_full[0], (std::uint32_t)cfg.list_bytes);
      ptx::bulk_g2s(list_base, cfg.rec + tile_of(m) * cfg.rec_bytes, (std::uint32_t)cfg.list_bytes, &list_full[0], pol);
    }
  };
  auto issue_lidx = [&](int m) {
    if (lane == 0 && has_tile(m)) {
      ptx::mbar_expect_tx(&lidx_full[0], (std::uint32_t)cfg.lidx_bytes);
      ptx::bulk_g2s(smem + cfg.s_lidx, cfg.rec + tile_of(m) * cfg.rec_bytes + cfg.off_lidx, (std::uint32_t)cfg.lidx_bytes,
                    &lidx_full[0], pol);
    }
  };
  // The W and geometry sections of a record are contiguous: the issue pointer walks through them.  All lanes
  // keep the (warp-uniform) bookkeeping; only the two asynchronous-copy instructions are predicated on lane 0.
  // The ring is shorter than a record's segment list, so while tile m is consumed the issue side moves from
  // tile m's record to tile m+1's exactly once: `nxt_ptr` is set at the start of tile m.
  int iss_s = 0;
  const char *iss_ptr = cfg.rec + tile_of(0) * cfg.rec_bytes + cfg.off_wlo;
  const char *nxt_ptr = nullptr;
  auto issue_seg = [&](int slot) {
    // branch-free: when the tile list is exhausted iss_ptr is null and nothing is issued
    const bool live = iss_ptr != nullptr;
    const int bytes = live ? cfg.seg_bytes[iss_s] : 0;
    ptx::bulk_g2s_if(live && lane == 0, ring + (size_t)slot * T::SLOT_BYTES, iss_ptr, (std::uint32_t)bytes, &seg_full[slot],
                     pol);
    // the ring is short (shared memory): pull the bytes behind it towards L2 so that the next copies are L2 hits
    if (cfg.l2_ahead > 0) {
      const bool in_rec = iss_s + 1 < N_SEG;  // stay inside this record's W | geometry stream
      ptx::bulk_prefetch_l2_if(live && in_rec && lane == 0, iss_ptr + bytes + cfg.l2_ahead, (std::uint32_t)T::SLOT_BYTES);
    }
    const bool last = (iss_s == N_SEG - 1);
    iss_ptr = last ? nxt_ptr : (live ? iss_ptr + bytes : nullptr);
    nxt_ptr = last ? nullptr : nxt_ptr;
    iss_s = last ? 0 : iss_s + 1;
  };
  // ---- consume side ---------------------------------------------------------------------------------
  int cslot = 0, cphase = 0;
  auto wait_seg = [&]() -> const unsigned char * {
    if (prof) {
      const long long a = clock64();
      ptx::mbar_wait(&seg_full[cslot], cphase);
      t_segwait += clock64() - a;
    } else {
      ptx::mbar_wait(&seg_full[cslot], cphase);
    }
    return ring + (size_t)cslot * T::SLOT_BYTES;
  };
  auto release_seg = [&]() {
    __syncwarp();  // every lane has read what it needs from the slot
    issue_seg(cslot);
    if (++cslot == NSLOT) {
      cslot = 0;
      cphase ^= 1;
    }
  };
  // copy the rows of tile m's list into the table (asynchronously)
  auto load_table = [&](int m) {
    if (!has_tile(m)) return;
    const unsigned char *lb = list_base;
    ptx::mbar_wait(&list_full[0], m & 1);
    const int n_list = *reinterpret_cast<const int *>(lb);
    const std::int32_t *list = reinterpret_cast<const std::int32_t *>(lb + cfg.off_list);
#pragma unroll 2
    for (int row = lane; row < n_list; row += 32) {  // a lane copies whole 40-byte rows
      const double *src = args.state + (std::int64_t)list[row] * NVARS;
      double *dst = table + row * NVARS;
#pragma unroll
      for (int v = 0; v < NVARS; ++v) ptx::cp_async8(dst + v, src + v);
    }
    ptx::cp_async_commit();
  };

  issue_list(0);
  issue_lidx(0);
  for (int s = 0; s < NSLOT; ++s) issue_seg(s);
  load_table(0);

  const bool cweno = sc.recon_mode == RECON_CWENO_AO;
  // write-out pattern of the trace staging buffer: element it * 32 + lane of a pass belongs to cell wo_owner, offset wo_j
  int wo_owner[T::CHUNK], wo_j[T::CHUNK];
#pragma unroll
  for (int it = 0; it < T::CHUNK; ++it) {
    wo_owner[it] = (it * TILE + lane) / T::CHUNK;
    wo_j[it] = (it * TILE + lane) - wo_owner[it] * T::CHUNK;
  }
  // one dump block per warp (TRACE_DUMP_BLOCKS of them are allocated): no two warps store to the same lines
  const std::uint32_t dump_blk = (std::uint32_t)(2 * P.n_interior_edges) + (std::uint32_t)(first % TRACE_DUMP_BLOCKS);

#pragma unroll 1
  for (int m = 0; has_tile(m); ++m) {
    if (m > 0) {
      t_cur = t_nxt;
      t_nxt = t_nxt2;
      t_nxt2 = load_tile_no(m + 2);
      m_cur = m;
    }
    mark(-1);
    // The list buffer holds tile m's part (its rows went into the table during tile m-1; only the meta words
    // are still needed): read them, then let tile m+1's part overwrite the buffer while tile m is applied.
    const std::uint64_t meta = reinterpret_cast<const std::uint64_t *>(list_base + TILE_OFF_META)[lane];
    __syncwarp();
    issue_list(m + 1);
    nxt_ptr = has_tile(m + 1) ? cfg.rec + tile_of(m + 1) * cfg.rec_bytes + cfg.off_wlo : nullptr;
    ptx::cp_async_wait_all();
    ptx::mbar_wait(&lidx_full[0], m & 1);
    __syncwarp();  // the table of tile m is complete and visible to the whole warp
    mark(TP_TABLE_WAIT);
    const std::int64_t tile = tile_of(m);
    if constexpr (WB) {
      // the next tile's equilibrium rows (one contiguous block E2 has just written) are pulled towards L2 a whole
      // tile ahead: load_rhs reads them with plain loads one stencil ahead, which hides an L2 hit but not HBM
      const bool nxt = has_tile(m + 1);
      const std::int64_t tn = nxt ? tile_of(m + 1) : tile;
      const std::uint32_t bytes = (std::uint32_t)(P.eq_rows * 2 * TILE * 8);
      ptx::bulk_prefetch_l2_if(nxt && lane == 0, P.eq_avg + tn * P.eq_rows * 2 * TILE, bytes);
    }
    const std::int64_t cell = tile * TILE + lane;
    const bool active = cell < P.n_cells;
    const int kh_m = (int)((meta >> 56) & 0xF);
    const bool single_m = ((meta >> 60) & 1) != 0;
    const bool fast = __all_sync(0xffffffffu, kh_m == 0 && !single_m);

    // ---- own state and scaling (characteristic_scale.hpp:24-33) -------------------------------------
    double q0s[NVARS], inv_scale[NVARS], scale[NVARS];
    {
      double u0[NVARS];
#pragma unroll
      for (int v = 0; v < NVARS; ++v) u0[v] = table[lane * NVARS + v];
      const double ekin0 = 0.5 * (u0[1] * u0[1] + u0[2] * u0[2] + u0[3] * u0[3]) / u0[0];
      const double eint0 = u0[4] - ekin0;
      if (sc.scaling == SCALING_EULER) {
        const double p = eint0 * (sc.gamma - 1.0);
        const double cs = sqrt(sc.gamma * p / u0[0]);
        scale[0] = u0[0];
        scale[1] = scale[2] = scale[3] = cs;
        scale[4] = eint0;
      } else {
#pragma unroll
        for (int v = 0; v < NVARS; ++v) scale[v] = 1.0;
      }
      inv_scale[0] = 1.0 / scale[0];
      inv_scale[1] = inv_scale[2] = inv_scale[3] = 1.0 / scale[1];
      inv_scale[4] = 1.0 / scale[4];
      if constexpr (WB) {  // the cell's own equilibrium average: row ROWS of the tile's eq_avg block
        const double *e0 = P.eq_avg + ((tile * P.eq_rows + T::ROWS) * 2) * TILE + lane;
        u0[0] -= e0[0];
        u0[4] -= e0[TILE];
      }
#pragma unroll
      for (int v = 0; v < NVARS; ++v) q0s[v] = u0[v] * inv_scale[v];
    }
    const double *eq_rows_tile = WB ? P.eq_avg + (tile * P.eq_rows * 2) * TILE + lane : nullptr;
    // rhs of stencil row `row` (rows numbered as in lidx): u_local(j) - u_local(0), local_reconstruction.hpp:109-116
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
    auto nonlinear_weight = [&](double is_max, double g) {  // alpha = g / (eps + IS^p), hybrid_weno.cpp:117-119
      double is_pow;
      if (sc.exponent == 4.0) {
        const double s2 = is_max * is_max;
        is_pow = s2 * s2;
      } else if (sc.exponent == 2.0) {
        is_pow = is_max * is_max;
      } else {
        is_pow = pow(is_max, sc.exponent);
      }
      return g * fast_rcp(sc.epsilon + is_pow);  // within an ulp or two of the quotient
    };

    // coefficients of the hybridised polynomial: constant | low-order part | high-order part
    double c0[NVARS], lo0[CLO][NVARS], hi[NHI > 0 ? NHI : 1][NVARS];

    auto reconstruct = [&](auto fast_tag) {
      constexpr bool FAST = decltype(fast_tag)::value;
      // FAST: every cell of the tile has the full family and the central stencil is the highest-order one.
      // Otherwise per lane: kh = the stencil that takes the CWENO correction (-1: none), n_eff stencils exist.
      const bool single = FAST ? false : single_m;
      const int n_eff = single ? 1 : NS;
      const int kh = FAST ? 0 : (cweno ? kh_m : -1);
      double corr[CLO][NVARS];   // sum over the other stencils of gamma_k p_k            (cweno_ao.cpp:41-50)
      double wsum[CLO][NVARS];   // sum over finished stencils of alpha_k p_k             (hybrid_weno.cpp:121-127)
      double keep[FAST ? 1 : CLO][NVARS];  // one-sided polynomial that waits for its correction (kh >= 1 only)
#pragma unroll
      for (int c = 0; c < CLO; ++c)
#pragma unroll
        for (int v = 0; v < NVARS; ++v) {
          corr[c][v] = 0.0;
          wsum[c][v] = 0.0;
          if constexpr (!FAST) keep[c][v] = 0.0;
        }
      // ---- one-sided stencils: coef = W_k rhs, folded into the sums as soon as it is complete ----------
      // The rhs of a stencil's rows does not depend on the ring: it is fetched one stencil ahead, so that the
      // table reads overlap the previous stencil's arithmetic instead of sitting in front of the barrier wait.
      double rhs_lo[RLO][NVARS];
#pragma unroll
      for (int r = 0; r < RLO; ++r) load_rhs(r, rhs_lo[r]);
      double al_sum = 0.0;  // sum of the non-linear weights folded so far
#pragma unroll 1
      for (int k = 1; k < NS; ++k) {  // rolled: instruction-cache footprint
        const double *wseg = reinterpret_cast<const double *>(wait_seg()) + lane;
        double acc[CLO][NVARS];
#pragma unroll
        for (int c = 0; c < CLO; ++c)
#pragma unroll
          for (int v = 0; v < NVARS; ++v) acc[c][v] = 0.0;
#pragma unroll
        for (int r = 0; r < RLO; ++r)
#pragma unroll
          for (int c = 0; c < CLO; ++c) {
            const double w = wseg[(r * CLO + c) * TILE];
#pragma unroll
            for (int v = 0; v < NVARS; ++v) acc[c][v] = fma(w, rhs_lo[r][v], acc[c][v]);
          }
        release_seg();
        if (k + 1 < NS) {
#pragma unroll
          for (int r = 0; r < RLO; ++r) load_rhs(k * RLO + r, rhs_lo[r]);
        }
        double is_max = 0.0;
#pragma unroll
        for (int v = 0; v < NVARS; ++v) {
          double beta = 0.0;
#pragma unroll
          for (int c = 0; c < CLO; ++c) beta += acc[c][v] * acc[c][v];
          is_max = (v == 0) ? beta : fmax(is_max, beta);
        }
        const double g_k = sc.lin_w[k];
        const double a_k = nonlinear_weight(is_max, single ? 1.0 : g_k);
        if constexpr (FAST) {
          al_sum += a_k;
#pragma unroll
          for (int c = 0; c < CLO; ++c)
#pragma unroll
            for (int v = 0; v < NVARS; ++v) {
              corr[c][v] = fma(g_k, acc[c][v], corr[c][v]);
              wsum[c][v] = fma(a_k, acc[c][v], wsum[c][v]);
            }
        } else {
          const bool exists = k < n_eff, waits = (k == kh);
          const double a_use = (exists && !waits) ? a_k : 0.0;
          const double g_use = (exists && !waits) ? g_k : 0.0;
          al_sum += a_use;
#pragma unroll
          for (int c = 0; c < CLO; ++c)
#pragma unroll
            for (int v = 0; v < NVARS; ++v) {
              corr[c][v] = fma(g_use, acc[c][v], corr[c][v]);
              wsum[c][v] = fma(a_use, acc[c][v], wsum[c][v]);
              if (waits) keep[c][v] = acc[c][v];
            }
        }
      }
      double rhs_hi[R_HI][NVARS];  // first central segment
#pragma unroll
      for (int r = 0; r < R_HI; ++r) load_rhs((NS - 1) * RLO + r, rhs_hi[r]);
      mark(TP_LO);

      // ---- central stencil: accumulators stay in registers -----------------------------------------------
      // When the central stencil takes the CWENO correction its low-order accumulators start at
      // -sum_k gamma_k p_k: the correction sum is dead before the row loop starts (registers).
      const bool central_high = FAST ? true : (kh == 0);
#pragma unroll
      for (int c = 0; c < CLO; ++c)
#pragma unroll
        for (int v = 0; v < NVARS; ++v) lo0[c][v] = (central_high && cweno) ? -corr[c][v] : 0.0;
#pragma unroll
      for (int c = 0; c < NHI; ++c)
#pragma unroll
        for (int v = 0; v < NVARS; ++v) hi[c][v] = 0.0;
      // (a rolled loop: the unrolled form of 18 segments x 2 instantiations does not fit the instruction cache;
      // the rhs of segment hs+1 is fetched while segment hs is accumulated)
      auto central_rows = [&](const double *wseg, const double (*rhs)[NVARS], auto n_rows_tag) {
        constexpr int NR = decltype(n_rows_tag)::value;
#pragma unroll
        for (int r = 0; r < NR; ++r) {
#pragma unroll
          for (int c = 0; c < CHI; ++c) {
            const double w = wseg[(r * CHI + c) * TILE];
            if (c < CLO) {
#pragma unroll
              for (int v = 0; v < NVARS; ++v) lo0[c][v] = fma(w, rhs[r][v], lo0[c][v]);
            } else {
#pragma unroll
              for (int v = 0; v < NVARS; ++v) hi[c - CLO][v] = fma(w, rhs[r][v], hi[c - CLO][v]);
            }
          }
        }
      };
#pragma unroll 2
      for (int hs = 0; hs < N_HI - 1; ++hs) {
        const double *wseg = reinterpret_cast<const double *>(wait_seg()) + lane;
        double w_first = wseg[0];  // start the weight loads before the next segment's table reads are queued
        double rhs_next[R_HI][NVARS];
#pragma unroll
        for (int r = 0; r < R_HI; ++r) {
          // the last segment may be shorter: its missing rows are never used
          const int row = (NS - 1) * RLO + (hs + 1) * R_HI + r;
          if (r < R_TAIL || hs + 1 < N_HI - 1) load_rhs(row < T::ROWS ? row : T::ROWS - 1, rhs_next[r]);
        }
        (void)w_first;
        central_rows(wseg, rhs_hi, std::integral_constant<int, R_HI>{});
        release_seg();
#pragma unroll
        for (int r = 0; r < R_HI; ++r)
#pragma unroll
          for (int v = 0; v < NVARS; ++v) rhs_hi[r][v] = rhs_next[r][v];
      }
      {
        const double *wseg = reinterpret_cast<const double *>(wait_seg()) + lane;
        central_rows(wseg, rhs_hi, std::integral_constant<int, R_TAIL>{});
        release_seg();
      }
      mark(TP_HI);

      // the table and the index rows are dead: fetch the next tile's while this tile is hybridised and evaluated
      __syncwarp();
      issue_lidx(m + 1);
      load_table(m + 1);
      mark(TP_TABLE_ISSUE);

      // ---- CWENO correction of the highest-order polynomial, non-linear weights, normalisation ------------
      double gh = 1.0;  // linear weight of the corrected stencil
#pragma unroll
      for (int k = 0; k < NS; ++k)
        if (k == kh) gh = single ? 1.0 : sc.lin_w[k];
      const double inv_gh = 1.0 / gh;
      if (central_high && cweno) {
#pragma unroll
        for (int c = 0; c < CLO; ++c)
#pragma unroll
          for (int v = 0; v < NVARS; ++v) lo0[c][v] *= inv_gh;
#pragma unroll
        for (int c = 0; c < NHI; ++c)
#pragma unroll
          for (int v = 0; v < NVARS; ++v) hi[c][v] *= inv_gh;
      }
      double alpha0;
      {
        double is_max = 0.0;
#pragma unroll
        for (int v = 0; v < NVARS; ++v) {
          double beta = 0.0;
#pragma unroll
          for (int c = 0; c < CLO; ++c) beta += lo0[c][v] * lo0[c][v];
#pragma unroll
          for (int c = 0; c < NHI; ++c) beta += hi[c][v] * hi[c][v];
          is_max = (v == 0) ? beta : fmax(is_max, beta);
        }
        alpha0 = nonlinear_weight(is_max, single ? 1.0 : sc.lin_w[0]);
      }
      al_sum += alpha0;
      double alpha_h = alpha0;  // non-linear weight of the corrected stencil
#pragma unroll
      for (int c = 0; c < CLO; ++c)
#pragma unroll
        for (int v = 0; v < NVARS; ++v) wsum[c][v] = fma(alpha0, lo0[c][v], wsum[c][v]);
      if constexpr (!FAST) {
        if (kh >= 1) {  // a one-sided stencil is the highest-order one: it takes the correction now
          double is_max = 0.0;
#pragma unroll
          for (int v = 0; v < NVARS; ++v) {
            double beta = 0.0;
#pragma unroll
            for (int c = 0; c < CLO; ++c) {
              keep[c][v] = inv_gh * (keep[c][v] - fma(sc.lin_w[0], lo0[c][v], corr[c][v]));
              beta += keep[c][v] * keep[c][v];
            }
            is_max = (v == 0) ? beta : fmax(is_max, beta);
          }
          alpha_h = nonlinear_weight(is_max, gh);
          al_sum += alpha_h;
#pragma unroll
          for (int c = 0; c < CLO; ++c)
#pragma unroll
            for (int v = 0; v < NVARS; ++v) wsum[c][v] = fma(alpha_h, keep[c][v], wsum[c][v]);
        }
      }
      const double inv_tot = fast_rcp(al_sum);
      // constant coefficient: q0 for every stencil but kh, whose value carries the correction:
      // sum_k w_k a0_k = (alpha_h a0h + (sum alpha - alpha_h) q0) / sum alpha
      double g_others = 0.0;
#pragma unroll
      for (int k = 0; k < NS; ++k)
        if (k != kh && k < n_eff) g_others += sc.lin_w[k];
#pragma unroll
      for (int v = 0; v < NVARS; ++v) {
        const double a0h = (kh >= 0) ? inv_gh * (q0s[v] - g_others * q0s[v]) : q0s[v];
        c0[v] = (alpha_h * a0h + (al_sum - alpha_h) * q0s[v]) * inv_tot;
      }
#pragma unroll
      for (int c = 0; c < CLO; ++c)
#pragma unroll
        for (int v = 0; v < NVARS; ++v) lo0[c][v] = wsum[c][v] * inv_tot;
      const double w0 = alpha0 * inv_tot;
#pragma unroll
      for (int c = 0; c < NHI; ++c)
#pragma unroll
        for (int v = 0; v < NVARS; ++v) hi[c][v] *= w0;
    };
    if (fast)
      reconstruct(std::true_type{});
    else
      reconstruct(std::false_type{});

    // coefficient i of variable v of the hybridised polynomial (compile-time i)
    auto coef_at = [&](int i, int v) -> double & { return i == 0 ? c0[v] : (i <= CLO ? lo0[i - 1][v] : hi[i - 1 - CLO][v]); };
    if (P.poly != nullptr && active) {
#pragma unroll
      for (int i = 0; i < D; ++i)
#pragma unroll
        for (int v = 0; v < NVARS; ++v) P.poly[(cell * P.n_poly_coef + i) * NVARS + v] = coef_at(i, v);
#pragma unroll
      for (int v = 0; v < NVARS; ++v) P.poly_scale[cell * NVARS + v] = scale[v];
    }
    if (P.poly_tile != nullptr) {  // hand-over to source_kernel: [tile][D + 1][5][32], coefficients then scales
      double *pt = P.poly_tile + tile * ((D + 1) * NVARS * TILE) + lane;
#pragma unroll
      for (int i = 0; i < D; ++i)
#pragma unroll
        for (int v = 0; v < NVARS; ++v) pt[(i * NVARS + v) * TILE] = coef_at(i, v);
#pragma unroll
      for (int v = 0; v < NVARS; ++v) pt[(D * NVARS + v) * TILE] = scale[v];
    }
    // fold the characteristic scale into the coefficients: delta(x) = scale * p(x)
#pragma unroll
    for (int i = 0; i < D; ++i)
#pragma unroll
      for (int v = 0; v < NVARS; ++v) coef_at(i, v) *= scale[v];
    mark(TP_HYBRID);

    // ---- geometry: ring -> registers ---------------------------------------------------------------------
    double vt[F][ND], xc[ND], inv_len = 1.0, cmom[D];
    std::uint32_t fref[F], slots_all = 0;
#pragma unroll
    for (int i = 0; i < D; ++i) cmom[i] = 0.0;
#pragma unroll
    for (int gs = 0; gs < N_GEO; ++gs) {
      const unsigned char *gseg = wait_seg();
      // row j of the geometry section lives in segment j / GEO_ROWS_PER_SEG
      const double *geo = reinterpret_cast<const double *>(gseg) + lane - gs * T::GEO_ROWS_PER_SEG * TILE;
#pragma unroll
      for (int j = 0; j < T::GEO_DOUBLES; ++j) {
        if (j / T::GEO_ROWS_PER_SEG == gs) {
          const double val = geo[j * TILE];
          if (j < F * ND)
            vt[j / ND][j % ND] = val;
          else if (j < F * ND + ND)
            xc[j - F * ND] = val;
          else if (j == F * ND + ND)
            inv_len = val;
          else
            cmom[3 + j - (F * ND + ND + 1)] = val;
        }
      }
      // the 32-bit rows follow the doubles: F rows of face_ref, one row of packed face slots (128 bytes each)
      const std::uint32_t *g32 =
          reinterpret_cast<const std::uint32_t *>(gseg - (size_t)gs * T::SLOT_BYTES + T::GEO_DOUBLES * TILE * 8) + lane;
#pragma unroll
      for (int j = 0; j <= F; ++j) {
        if ((T::GEO_DOUBLES * TILE * 8 + j * TILE * 4) / T::SLOT_BYTES == gs) {
          const std::uint32_t val = g32[j * TILE];
          if (j < F)
            fref[j] = active ? val : 0u;
          else
            slots_all = val;
        }
      }
      release_seg();
    }
    mark(TP_GEO_WAIT);

    // ---- traces at the face Gauss points (flux_loop.hpp:131-149) -----------------------------------------
#pragma unroll 1
    for (int k = 0; k < F; ++k) {  // rolled: one copy of the evaluation code
      const std::uint32_t slots = (slots_all >> (8 * k)) & 0xFFu;
      std::uint32_t fref_k = fref[0];
#pragma unroll
      for (int kk = 1; kk < F; ++kk)
        if (k == kk) fref_k = fref[kk];
      double fv[ND][ND];  // face vertices in the left cell's order
#pragma unroll
      for (int r = 0; r < ND; ++r) {
        const int s = (slots >> (2 * r)) & 3;
#pragma unroll
        for (int d = 0; d < ND; ++d) {
          double val = vt[0][d];
#pragma unroll
          for (int kk = 1; kk < F; ++kk)
            if (s == kk) val = vt[kk][d];
          fv[r][d] = val;
        }
      }
      // block of this (cell, face) in the trace array, in units of QF * 5 doubles; cells whose trace nobody reads
      // write to this warp's dump block behind the last face (the write-out loop stays free of branches)
      const std::uint32_t blk = (fref_k & FREF_TRACE)
                                    ? ((fref_k & FREF_EDGE_MASK) * 2u + ((fref_k & FREF_SIDE) ? 1u : 0u))
                                    : dump_blk;
#pragma unroll 1
      for (int q0 = 0; q0 < QF; q0 += T::Q_STAGE) {
#pragma unroll 1
        for (int qq = 0; qq < T::Q_STAGE; ++qq) {  // rolled: instruction-cache footprint
          const int q = q0 + qq;
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
            double s = c0[v];
#pragma unroll
            for (int i = 1; i < D; ++i) s = fma(coef_at(i, v), mono[i], s);
            stage[lane * T::STAGE_PITCH + qq * NVARS + v] = s;  // WB: the perturbation; K2 adds the background
          }
        }
        __syncwarp();
        // coalesced write-out: consecutive lanes write consecutive doubles of a cell's block
#pragma unroll
        for (int it = 0; it < T::CHUNK; ++it) {
          const std::uint32_t b = __shfl_sync(0xffffffffu, blk, wo_owner[it]);
          P.trace[(std::int64_t)b * (QF * NVARS) + (q0 * NVARS + wo_j[it])] = stage[wo_owner[it] * T::STAGE_PITCH + wo_j[it]];
        }
        __syncwarp();  // the staging buffer is reused
      }
    }
    mark(TP_TRACE);
    if (prof && lane == 0) {
      atomicAdd(cfg.prof + TP_SEG_WAIT, (unsigned long long)t_segwait);
      atomicAdd(cfg.prof + TP_TILES, 1ull);
      t_segwait = 0;
    }
  }
}

/// Shared-memory plan; returns false if the tile kernel does not apply.
template <int ND, int DEG_HI, int DEG_LO, int NS, int RM0, int RLO, int QF>
bool tile_config(const DevicePlan &P, const SchemeConst &sc, int smem_per_warp, int want_slots, TileCfg &c) {
  using T = TileTraits<ND, DEG_HI, DEG_LO, NS, RM0, RLO, QF>;
  if (P.rec2 == nullptr) return false;
  if (NS < 2 || T::CLO < 1 || sc.n_stencils != NS || sc.q_f != QF) return false;
  if (sc.rows_max[0] != RM0 || sc.ncoef[0] != T::CHI) return false;
  for (int k = 1; k < NS; ++k)
    if (sc.rows_max[k] != RLO || sc.ncoef[k] != T::CLO) return false;
  const TileRecLayout L = tile_rec_layout(sc, ND, T::D, P.rec2_cap);
  if (L.rec_bytes != P.rec2_bytes || L.geo_doubles != T::GEO_DOUBLES) return false;
  if (L.off_geo - L.off_wlo != (NS - 1) * T::LO_ST_BYTES + RM0 * T::HI_ROW_BYTES) return false;
  c.cap = L.cap;
  c.list_bytes = L.off_lidx;  // n_list | meta | list
  c.off_list = L.off_list;
  c.off_lidx = L.off_lidx;
  c.lidx_bytes = L.off_wlo - L.off_lidx;
  c.off_wlo = L.off_wlo;
  c.rec_bytes = L.rec_bytes;
  c.rec = P.rec2;
  c.s_list = 128;
  c.s_lidx = c.s_list + c.list_bytes;
  c.s_table = c.s_lidx + c.lidx_bytes;
  c.s_stage = c.s_table + (L.cap * NVARS * 8 + 127) / 128 * 128;
  c.s_ring = c.s_stage + T::STAGE_BYTES;
  int ns = (smem_per_warp - c.s_ring) / T::SLOT_BYTES;
  if (ns > 13) ns = 13;
  if (ns > T::N_SEG - 1) ns = T::N_SEG - 1;  // the issue side changes records at most once per consumed tile
  if (want_slots > 0 && ns > want_slots) ns = want_slots;
  if (ns < 2) return false;
  c.n_slots = ns;
  c.warp_bytes = c.s_ring + ns * T::SLOT_BYTES;
  if (T::N_SEG > 64) return false;
  for (int i = 0; i < T::N_SEG; ++i) {
    int bytes = T::LO_ST_BYTES;
    if (i >= T::N_LO) bytes = T::R_HI * T::HI_ROW_BYTES;
    if (i == T::N_LO + T::N_HI - 1) bytes = T::R_TAIL * T::HI_ROW_BYTES;
    if (i >= T::N_LO + T::N_HI) bytes = T::SLOT_BYTES;
    if (i == T::N_SEG - 1) bytes = T::GEO_TAIL_BYTES;
    c.seg_bytes[i] = bytes;
  }
  c.l2_ahead = 0;
  c.prof = nullptr;
  return true;
}

}  // namespace zfvm
