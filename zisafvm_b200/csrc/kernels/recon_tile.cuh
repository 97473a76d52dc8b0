// K1, tile form: one warp owns one tile of 32 cells, one thread owns one cell; warps never talk to each other.
//
// Same arithmetic as recon.cuh (EulerGlobalReconstruction::compute, LocalReconstruction::compute,
// HybridWENO::compute_polys_impl / eno_hybridize, CWENO_AO::reconstruct_impl, rc(i)(x) at the face Gauss
// points; the reference lines are listed there).  What changes is how the bytes move:
//
//   * every table a tile needs is one contiguous *tile record* (layout below) that the warp streams itself with
//     16-byte cp.async copies (LDGSTS, L2 evict-first) through a private ring of N_SLOTS small shared-memory slots
//     (one one-sided stencil or a few central-stencil rows each), one commit group per segment.  After the warp has
//     consumed a slot it issues the copies of the segment that is N_SLOTS positions ahead: no producer warp, no
//     barriers, and the bytes in flight do not depend on registers.  (The first versions moved the segments with TMA
//     bulk copies and mbarriers.  Measured on B200, profiles/r02_tma_ubench.txt: a completed mbarrier.try_wait still
//     costs ~125 cycles and arming + issuing one bulk copy ~275 cycles of the issuing warp, i.e. ~400 cycles for
//     each of a tile's 17 copies -- a fifth of the tile's time at 4.6 KB per copy -- whereas the nine LDGSTS of a
//     segment issue in ~70 cycles and cp.async.wait_group is a scoreboard wait.)
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
/// Sixteen warps per SM (128 registers) for the schemes with few accumulators (2D orders 2-3, 3D order 2) were measured
/// and lost: 0.7-1.7 KB of spills per thread, K1 +25 % (2D order 3) to +120 % (3D order 2) -- profiles/README.md, round 2.
__host__ __device__ constexpr int tile_max_warps(int /*nd*/, int /*deg_hi*/) { return TILE_MAX_WARPS; }
// Ring geometry; the -D overrides exist for the measurements in scratch/build_variants.sh.
#ifndef ZFVM_SLOT_TARGET
#define ZFVM_SLOT_TARGET 4608
#endif
#ifndef ZFVM_NSLOTS
#define ZFVM_NSLOTS 3
#endif
#ifndef ZFVM_Q_STAGE
#define ZFVM_Q_STAGE 1
#endif
// Per-segment bookkeeping (profiles/README.md, round 2, "per-segment bookkeeping of K1"): measured at 9.86 M tets, K1 ms
// on one box -- prefetch as a rolled loop of one line per lane 5.48, one bulk prefetch by lane 0 5.20 (adopted); on a
// second box bulk 5.30, bulk + unpredicated copies of full slots 5.32 (more spills), a line per lane unrolled 5.28.
#ifndef ZFVM_COPY_FAST
#define ZFVM_COPY_FAST 0   // 1: full ring slots (all but a record's last segment) are copied without per-piece predicates
#endif
#ifndef ZFVM_L2_MODE
#define ZFVM_L2_MODE 1     // prefetch behind the ring: 1 = one cp.async.bulk.prefetch.L2 by lane 0, 2 = a line per lane, unrolled
#endif
constexpr int TILE_SLOT_TARGET = ZFVM_SLOT_TARGET;  // bytes of a ring slot aimed at (two central rows of the 3D order-3 scheme)

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
  static constexpr int N_SLOTS = ZFVM_NSLOTS;               // ring slots per warp (8 warps x 3 slots measured best on B200)
  static constexpr int Q_STAGE = (QF % ZFVM_Q_STAGE == 0) ? ZFVM_Q_STAGE : 1;                         // Gauss points staged per pass (shared memory is what limits the warps per SM)
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
  int s_list, s_lidx, s_table, s_ring, s_stage, warp_bytes;
  int evict_normal;          // L2 policy of the record copies: 0 evict-first (default), 1 evict-normal
  int l2_ahead;              // L2 prefetch distance behind the ring in bytes (0: off)
  int l2_whole;              // experiment: one bulk L2 prefetch of the whole next record at the start of a tile
  int stream_bytes;          // bytes of a record's W | geometry stream (rec_bytes - off_wlo)
  int seg_bytes[64];         // bytes of segment s of a record's W | geometry stream
  unsigned long long *prof;  // optional phase timers of warp 0 (clock cycles): see TilePhase; null = off
};

enum TilePhase : int { TP_TABLE_WAIT = 0, TP_LO = 1, TP_HI = 2, TP_TABLE_ISSUE = 3, TP_HYBRID = 4, TP_GEO_WAIT = 5, TP_TRACE = 6, TP_SEG_WAIT = 7, TP_TILES = 8, TP_COUNT = 9 };

namespace ptx {
ZFVM_DEVICE void cp_async8(void *dst_smem, const void *src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(dst_smem)), "l"(src) : "memory");
}
/// 16-byte global -> shared copy past L1 with an L2 cache policy, predicated inside the asm block (no divergent branch)
ZFVM_DEVICE void cp_async16_if(bool pred, void *dst_smem, const void *src, std::uint64_t policy) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.u32 p, %3, 0;\n\t"
      "@p cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2;\n\t}" ::"r"(smem_u32(dst_smem)),
      "l"(src), "l"(policy), "r"((std::uint32_t)pred)
      : "memory");
}
/// L2 prefetch of a global range (no shared-memory destination, no completion tracking), predicated inside the asm block
ZFVM_DEVICE void bulk_prefetch_l2_if(bool pred, const void *src, std::uint32_t bytes) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.u32 p, %2, 0;\n\t"
      "@p cp.async.bulk.prefetch.L2.global [%0], %1;\n\t}" ::"l"(src),
      "r"(bytes), "r"((std::uint32_t)pred)
      : "memory");
}
ZFVM_DEVICE void cp_async16(void *dst_smem, const void *src, std::uint64_t policy) {
  asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2;" ::"r"(smem_u32(dst_smem)), "l"(src), "l"(policy)
               : "memory");
}
ZFVM_DEVICE void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
ZFVM_DEVICE void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
ZFVM_DEVICE void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
/// waits until at most N of this thread's most recent commit groups are still pending
template <int N>
ZFVM_DEVICE void cp_async_wait_pending() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
}  // namespace ptx

// WB (well-balanced runs): the equilibrium kernels (equilibrium.cuh: E1, E2, E3) have written, per tile, the cell
// averages of each cell's local equilibrium over its stencil members in lidx row order plus one row for the cell
// itself (eq_avg): the rhs subtracts them (local_reconstruction.hpp:109-116).  The traces written here are those of
// the perturbation; the equilibrium background at the face Gauss points (local_reconstruction.hpp:149-163) is added by
// the face-flux kernel from E3's table (eq_bg).
template <int ND, int DEG_HI, int DEG_LO, int NS, int RM0, int RLO, int QF, typename LIDX, bool PROF, bool WB = false>
__global__ void __launch_bounds__(tile_max_warps(ND, DEG_HI) * 32, 1)
    recon_tile_kernel(const __grid_constant__ ReconArgs args, const __grid_constant__ SchemeConst sc,
                      const __grid_constant__ TileCfg cfg) {
  using T = TileTraits<ND, DEG_HI, DEG_LO, NS, RM0, RLO, QF>;
  constexpr int F = T::F, D = T::D, CHI = T::CHI, CLO = T::CLO, NHI = T::NHI;
  constexpr int N_HI = T::N_HI, R_HI = T::R_HI, R_TAIL = T::R_TAIL, N_LO = T::N_LO, N_GEO = T::N_GEO, N_SEG = T::N_SEG;
  const DevicePlan &P = args.plan;

  extern __shared__ __align__(128) unsigned char smem_all[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned char *smem = smem_all + (size_t)warp * cfg.warp_bytes;
  unsigned char *list_base = smem + cfg.s_list;
  const LIDX *lidx = reinterpret_cast<const LIDX *>(smem + cfg.s_lidx) + lane;
  double *table = reinterpret_cast<double *>(smem + cfg.s_table);
  unsigned char *ring = smem + cfg.s_ring;
  double *stage = reinterpret_cast<double *>(smem + cfg.s_stage);
  constexpr int NSLOT = T::N_SLOTS;

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

  // (records are read once per stage: evict-first keeps them from flushing the state rows and traces out of L2;
  // cfg.evict_normal is an experiment switch, ZFVM_TILE_EVICT=normal)
  const std::uint64_t pol = cfg.evict_normal ? ptx::policy_evict_normal() : ptx::policy_evict_first();

  // ---- issue side: every lane copies 16 bytes of each 512-byte row of a range ---------------------------
  // Commit-group bookkeeping (cp.async.wait_group counts a thread's most recent groups, so the order of the commits is
  // part of the algorithm): every issue_seg commits one group -- also when it has nothing to copy -- and so does every
  // load_table; the list part and the index rows of the next tile ride in the group of whatever is committed next.
  auto copy_range = [&](unsigned char *dst, const char *src, int bytes, auto max_bytes_tag) {
    constexpr int MAXB = decltype(max_bytes_tag)::value;
#if ZFVM_COPY_FAST
    if (bytes == MAXB && MAXB % 512 == 0) {  // warp-uniform: a full slot (every segment but a record's last one)
#pragma unroll
      for (int j = 0; j < MAXB / 512; ++j) ptx::cp_async16(dst + j * 512 + lane * 16, src + j * 512 + lane * 16, pol);
      return;
    }
#endif
#pragma unroll
    for (int j = 0; j < (MAXB + 511) / 512; ++j)
      ptx::cp_async16_if(j * 512 + lane * 16 < bytes, dst + j * 512 + lane * 16, src + j * 512 + lane * 16, pol);
  };
  auto copy_range_rt = [&](unsigned char *dst, const char *src, int bytes) {  // run-time length (list part, index rows)
    // (rolled on purpose: ptxas 12.9 emits LDGSTS with an odd descriptor register -- an illegal instruction at run
    // time -- in the remainder of the partially unrolled form of this loop)
#pragma unroll 1
    for (int off = lane * 16; off < bytes; off += 512) ptx::cp_async16(dst + off, src + off, pol);
  };
  auto issue_list = [&](int m) {  // n_list | meta | list of tile m (one buffer: see the tile loop for its life time)
    if (has_tile(m)) copy_range_rt(list_base, cfg.rec + tile_of(m) * cfg.rec_bytes, cfg.list_bytes);
  };
  auto issue_lidx = [&](int m) {
    if (has_tile(m)) copy_range_rt(smem + cfg.s_lidx, cfg.rec + tile_of(m) * cfg.rec_bytes + cfg.off_lidx, cfg.lidx_bytes);
  };
  // The W and geometry sections of a record are contiguous: the issue pointer walks through them.  The ring is
  // shorter than a record's segment list, so while tile m is consumed the issue side moves from tile m's record to
  // tile m+1's exactly once: `nxt_ptr` is set at the start of tile m.
  int iss_s = 0, iss_left = cfg.stream_bytes;  // segment index; bytes of the record's W | geometry stream not yet issued
  const char *iss_ptr = cfg.rec + tile_of(0) * cfg.rec_bytes + cfg.off_wlo;
  const char *nxt_ptr = nullptr;
  auto issue_seg = [&](int slot) {
    // branch-free: when the tile list is exhausted iss_ptr is null and nothing is copied (the group is still committed)
    const bool live = iss_ptr != nullptr;
    const int bytes = live ? cfg.seg_bytes[iss_s] : 0;
    copy_range(ring + (size_t)slot * T::SLOT_BYTES, iss_ptr, bytes, std::integral_constant<int, T::SLOT_BYTES>{});
    ptx::cp_async_commit();
    // the ring is short (shared memory): pull the lines `l2_ahead` bytes behind it towards L2 (one line per lane, the
    // range is about the segment that will be issued l2_ahead / SLOT_BYTES positions later, within this record)
    if (cfg.l2_ahead > 0) {
      iss_left -= bytes;
      const int room = iss_left - cfg.l2_ahead;       // what is left of this record's stream behind the prefetch distance
      const int nb = room < bytes ? room : bytes;     // (<= 0: nothing; the next record is reached by the ring itself)
#if ZFVM_L2_MODE == 1
      ptx::bulk_prefetch_l2_if(live && lane == 0 && nb > 0, iss_ptr + bytes + cfg.l2_ahead, (std::uint32_t)(nb > 0 ? nb : 16));
#else
#pragma unroll
      for (int j = 0; j < (T::SLOT_BYTES + 4095) / 4096; ++j) {  // a 128-byte line per lane
        const int rel = (j * 32 + lane) * 128;
        if (live && rel < nb) ptx::prefetch_l2(iss_ptr + bytes + cfg.l2_ahead + rel);
      }
#endif
    }
    const bool last = (iss_s == N_SEG - 1);
    if (last) iss_left = cfg.stream_bytes;
    iss_ptr = last ? nxt_ptr : (live ? iss_ptr + bytes : nullptr);
    nxt_ptr = last ? nullptr : nxt_ptr;
    iss_s = last ? 0 : iss_s + 1;
  };
  // ---- consume side ---------------------------------------------------------------------------------
  // `PENDING` = commit groups that may still be in flight behind the awaited segment: the NSLOT - 1 segments issued
  // after it, plus the table group when that was committed in between (the geometry segments, see the tile loop)
  int cslot = 0;
  auto wait_seg = [&](auto pending_tag) -> const unsigned char * {
    constexpr int PENDING = decltype(pending_tag)::value;
    if (prof) {
      const long long a = clock64();
      ptx::cp_async_wait_pending<PENDING>();
      t_segwait += clock64() - a;
    } else {
      ptx::cp_async_wait_pending<PENDING>();
    }
    __syncwarp();  // every lane's copies of the segment are complete: the slot is visible to the whole warp
    return ring + (size_t)cslot * T::SLOT_BYTES;
  };
  using PendW = std::integral_constant<int, NSLOT - 1>;
  using PendGeo = std::integral_constant<int, NSLOT>;
  auto release_seg = [&]() {
    __syncwarp();  // every lane has read what it needs from the slot
    issue_seg(cslot);
    if (++cslot == NSLOT) cslot = 0;
  };
  // copy the rows of tile m's list into the table (asynchronously)
  // (tile m's list part was copied at least a whole W phase ago and a wait_seg + __syncwarp has covered its group)
  auto load_table = [&](int m) {
    if (!has_tile(m)) {
      ptx::cp_async_commit();
      return;
    }
    const unsigned char *lb = list_base;
    const int n_list = *reinterpret_cast<const int *>(lb);
    const std::int32_t *list = reinterpret_cast<const std::int32_t *>(lb + cfg.off_list);
    // (measured alternatives, same time or worse at the bench size: consecutive lanes copying consecutive doubles of a
    // row; the copies issued in four portions spread over the first trace evaluations -- the ~4 k cycles per tile this
    // loop takes are the load-store unit working through ~200 scattered rows, wherever they are issued)
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
  ptx::cp_async_wait_pending<NSLOT - 1>();  // the first group carries tile 0's list part
  __syncwarp();
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
    if (cfg.l2_whole) ptx::bulk_prefetch_l2_if(nxt_ptr != nullptr && lane == 0, nxt_ptr, (std::uint32_t)cfg.stream_bytes);
    ptx::cp_async_wait_all();  // (the segments in flight were issued before the previous tile's trace phase)
    __syncwarp();  // the table and the index rows of tile m are complete and visible to the whole warp
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
        const double *wseg = reinterpret_cast<const double *>(wait_seg(PendW{})) + lane;
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
          is_max = (v == 0) ? beta : ref_max(is_max, beta);
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
        const double *wseg = reinterpret_cast<const double *>(wait_seg(PendW{})) + lane;
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
        const double *wseg = reinterpret_cast<const double *>(wait_seg(PendW{})) + lane;
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
          is_max = (v == 0) ? beta : ref_max(is_max, beta);
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
            is_max = (v == 0) ? beta : ref_max(is_max, beta);
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
      const unsigned char *gseg = wait_seg(PendGeo{});
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
bool tile_config(const DevicePlan &P, const SchemeConst &sc, int smem_per_warp, TileCfg &c) {
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
  c.s_list = 0;
  c.s_lidx = c.s_list + c.list_bytes;
  c.s_table = c.s_lidx + c.lidx_bytes;
  c.s_stage = c.s_table + (L.cap * NVARS * 8 + 127) / 128 * 128;
  c.s_ring = c.s_stage + T::STAGE_BYTES;
  // (the issue side changes records at most once per consumed tile: the ring must be shorter than a record's segments)
  if (T::N_SLOTS > T::N_SEG - 1) return false;
  c.warp_bytes = c.s_ring + T::N_SLOTS * T::SLOT_BYTES;
  if (c.warp_bytes > smem_per_warp) return false;
  if (T::N_SEG > 64) return false;
  for (int i = 0; i < T::N_SEG; ++i) {
    int bytes = T::LO_ST_BYTES;
    if (i >= T::N_LO) bytes = T::R_HI * T::HI_ROW_BYTES;
    if (i == T::N_LO + T::N_HI - 1) bytes = T::R_TAIL * T::HI_ROW_BYTES;
    if (i >= T::N_LO + T::N_HI) bytes = T::SLOT_BYTES;
    if (i == T::N_SEG - 1) bytes = T::GEO_TAIL_BYTES;
    c.seg_bytes[i] = bytes;
  }
  c.prof = nullptr;
  c.evict_normal = 0;
  c.l2_ahead = T::SLOT_BYTES;  // measured at the bench size: K1 -1 .. -4 %; 2, 4, 8 slots ahead: none or worse
  c.l2_whole = 0;
  c.stream_bytes = (int)(L.rec_bytes - L.off_wlo);
  return true;
}

}  // namespace zfvm
