// K1, streaming form: persistent, warp-specialised reconstruction kernel for sm_100a.
//
// Same arithmetic as recon.cuh (EulerGlobalReconstruction::compute + LocalReconstruction::compute +
// HybridWENO::compute_polys_impl / eno_hybridize + CWENO_AO::reconstruct_impl + rc(i)(x) at the face
// Gauss points; reference lines are listed there), restructured around the memory system as a
// four-stage pipeline inside one persistent CTA per SM (tiles of 32 cells, stride gridDim.x):
//
//   producer (warp 0, one elected lane)
//       streams each tile record -- header (meta + stencil member indices) and pseudo-inverse weights in
//       row segments of <= ~24 KB -- from HBM into shared-memory rings with TMA bulk copies
//       (cp.async.bulk.shared::cluster.global + mbarrier complete_tx, L2 evict-first): the bytes in flight
//       do not depend on occupancy.
//   gather group (5 warps, thread = (cell, variable), the five variables of a cell in adjacent lanes)
//       reads the neighbour states of a segment's stencil rows (a 40-byte state row is read by five
//       adjacent lanes: every L1 wavefront serves 6-7 rows), forms rhs = (u_j / scale - u_0 / scale) and
//       stores it next to the segment's weights in the same ring slot, [row][cell][5].  Loads of segment
//       s+1 are issued before segment s is stored, so L2 latency is hidden without occupancy.
//   apply groups (2 groups x G warps, alternate tiles; thread = (cell, coefficient group g), all 5 variables)
//       coef += W * rhs entirely out of shared memory: per stencil row 5 rhs loads + 1..3 weight loads feed
//       5..15 DFMAs (register tiling; thread g owns low-order coefficient g of every stencil and every
//       G-th high-order coefficient of the central one, so the CWENO-AO combination is thread-local).
//       Smoothness indicators are summed across the G threads of a cell through shared memory; the
//       non-linear weights are computed once per stencil, not once per thread.
//   trace group (one warp per local face, thread = (cell, face))
//       evaluates the hybridised polynomial at the face Gauss points and writes trace[e][side][q][5].
//
// Stencil sizes are compile-time (RM0 rows for the central stencil, RLO for every one-sided one): all
// shared-memory offsets fold into immediates and no row loop carries a predicate.  Other stencil sizes
// use the thread-per-cell kernel of recon.cuh.
#pragma once
#include <type_traits>

#include "recon.cuh"

namespace zfvm {

struct StreamCfg {
  int n_w_slots;
  int n_groups;  // apply groups in use (1 or STREAM_GROUPS)
  int off_hdr, off_info, off_w, off_xchg, off_alpha, off_stage;  // byte offsets into dynamic shared memory (barriers at 0)
  int stage_pitch;  // doubles per lane in the trace staging area (odd: conflict-free 64-bit accesses)
  int total_bytes;
};

constexpr int STREAM_VAR_WARPS = 5;  // gather group: 160 threads = 32 cells x 5 variables
constexpr int STREAM_GROUPS = 2;     // apply groups (alternate tiles)
constexpr int STREAM_HDR_SLOTS = 3;
constexpr int STREAM_BARS_BYTES = 1024;
constexpr int COEF_PAD = 33;         // coef exchange row pitch (doubles)

template <int ND, int DEG_HI, int DEG_LO, int NS, int RM0, int RLO>
struct StreamTraits {
  static constexpr int F = ND + 1;
  static constexpr int D = dof_of(DEG_HI, ND);
  static constexpr int CHI = D - 1;
  static constexpr int CLO = dof_of(DEG_LO, ND) - 1;
  static constexpr int NHI = CHI - CLO;
  static constexpr int G = CLO;                                 // apply warps per group = coefficient groups
  static constexpr int HPT = (NHI + G - 1) / (G > 0 ? G : 1);  // high-order coefficients per apply thread
  static constexpr int R_CAP0 = 16384 / (CHI * TILE * 8);       // weight segments of <= 16 KB
  static constexpr int R_CAP = R_CAP0 < 1 ? 1 : (R_CAP0 > 12 ? 12 : R_CAP0);
  static constexpr int N_HI = (RM0 + R_CAP - 1) / R_CAP;       // central-stencil segments
  static constexpr int R_HI = (RM0 + N_HI - 1) / N_HI;         // rows per segment (balanced)
  static constexpr int R_TAIL = RM0 - (N_HI - 1) * R_HI;       // rows of the last one
  static constexpr int N_LO = NS / 2;                          // two one-sided stencils per segment
  static constexpr int N_SEGS = N_HI + N_LO;
  static constexpr int RAW = R_HI > 2 * RLO ? R_HI : 2 * RLO;  // rows of the largest segment
  static constexpr int HI_BYTES = R_HI * CHI * TILE * 8;
  static constexpr int LO_BYTES = 2 * RLO * CLO * TILE * 8;
  static constexpr int W_BYTES = HI_BYTES > LO_BYTES ? HI_BYTES : LO_BYTES;
  static constexpr int RHS_BYTES = RAW * TILE * NVARS * 8;      // [row][cell][5]
  static constexpr int SLOT_BYTES = W_BYTES + RHS_BYTES;
  // tile record layout (device/layout.hpp) for these stencil sizes
  static constexpr int OFF_SIDX0 = TILE * 8;
  static constexpr __host__ __device__ int off_sidx(int k) { return OFF_SIDX0 + 4 * TILE * (k == 0 ? 0 : RM0 + (k - 1) * RLO); }
  static constexpr int HDR_BYTES = OFF_SIDX0 + 4 * TILE * (RM0 + (NS - 1) * RLO);
  static constexpr __host__ __device__ int off_W(int k) { return HDR_BYTES + 8 * TILE * (k == 0 ? 0 : RM0 * CHI + (k - 1) * RLO * CLO); }
  static constexpr int REC_BYTES = HDR_BYTES + 8 * TILE * (RM0 * CHI + (NS - 1) * RLO * CLO);
  static constexpr int INFO_BYTES = TILE * 2 * NVARS * 8;         // per header slot: q0s[5], scale[5] per cell
  static constexpr int COEF_BYTES = D * NVARS * COEF_PAD * 8;     // polynomial handed to the trace group
  static constexpr int PART_BYTES = G * NS * NVARS * TILE * 8;    // smoothness-indicator partial sums
  static constexpr int XCHG_BYTES = COEF_BYTES > PART_BYTES ? COEF_BYTES : PART_BYTES;  // per apply group (aliased)
  static constexpr int ALPHA_BYTES = NS * TILE * 8;               // per apply group
  static constexpr int N_WARPS = 1 + STREAM_VAR_WARPS + STREAM_GROUPS * G + F;
  static constexpr int VARS_PER_PASS = (20 / D) < 1 ? 1 : ((20 / D) > NVARS ? NVARS : (20 / D));  // trace group
};

namespace ptx {
ZFVM_DEVICE std::uint32_t smem_u32(const void *p) { return (std::uint32_t)__cvta_generic_to_shared(p); }
ZFVM_DEVICE void mbar_init(std::uint64_t *bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
ZFVM_DEVICE void mbar_expect_tx(std::uint64_t *bar, std::uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
ZFVM_DEVICE void mbar_arrive(std::uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
ZFVM_DEVICE void mbar_wait(std::uint64_t *bar, int parity) {
  const std::uint32_t a = smem_u32(bar);
  std::uint32_t done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(a), "r"((std::uint32_t)parity)
        : "memory");
  } while (!done);
}
ZFVM_DEVICE std::uint64_t policy_evict_first() {
  std::uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
/// TMA bulk copy global -> shared, completion counted in bytes on `bar`.
ZFVM_DEVICE void bulk_g2s(void *dst, const void *src, std::uint32_t bytes, std::uint64_t *bar, std::uint64_t policy) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
          smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
      : "memory");
}
ZFVM_DEVICE void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
ZFVM_DEVICE void named_bar_sync(int id, int n_threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n_threads) : "memory");
}
}  // namespace ptx

template <int ND, int DEG_HI, int DEG_LO, int NS, int RM0, int RLO>
__global__ void __launch_bounds__(32 * StreamTraits<ND, DEG_HI, DEG_LO, NS, RM0, RLO>::N_WARPS, 1)
    recon_stream_kernel(const __grid_constant__ ReconArgs args, const __grid_constant__ SchemeConst sc,
                        const __grid_constant__ StreamCfg cfg) {
  using T = StreamTraits<ND, DEG_HI, DEG_LO, NS, RM0, RLO>;
  constexpr int F = T::F, D = T::D, CHI = T::CHI, CLO = T::CLO, NHI = T::NHI, G = T::G, HPT = T::HPT;
  constexpr int N_HI = T::N_HI, R_HI = T::R_HI, R_TAIL = T::R_TAIL, N_LO = T::N_LO, N_SEGS = T::N_SEGS;
  constexpr int RAW = T::RAW, HS = STREAM_HDR_SLOTS;
  constexpr int FIRST_APPLY_WARP = 1 + STREAM_VAR_WARPS, FIRST_TRACE_WARP = FIRST_APPLY_WARP + STREAM_GROUPS * G;
  const int NG = cfg.n_groups;
  const DevicePlan &P = args.plan;

  extern __shared__ __align__(128) unsigned char smem[];
  std::uint64_t *bars = reinterpret_cast<std::uint64_t *>(smem);
  const int WS = cfg.n_w_slots;
  std::uint64_t *hdr_full = bars, *hdr_empty = bars + HS;
  std::uint64_t *coef_full = bars + 2 * HS, *coef_empty = coef_full + STREAM_GROUPS;
  std::uint64_t *w_full = coef_empty + STREAM_GROUPS, *w_empty = w_full + WS;
  unsigned char *hdr_base = smem + cfg.off_hdr;
  unsigned char *w_base = smem + cfg.off_w;
  double *info_base = reinterpret_cast<double *>(smem + cfg.off_info);  // [HS][32][10]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const std::int64_t n_launch = args.n_tiles_launch;
  auto tile_of = [&](int m) -> std::int64_t {
    const std::int64_t idx = blockIdx.x + (std::int64_t)m * gridDim.x;
    return args.tile_list ? (std::int64_t)args.tile_list[idx] : idx;
  };
  auto has_tile = [&](int m) { return blockIdx.x + (std::int64_t)m * gridDim.x < n_launch; };

  if (threadIdx.x == 0) {
    for (int s = 0; s < HS; ++s) {
      ptx::mbar_init(&hdr_full[s], 1);
      ptx::mbar_init(&hdr_empty[s], STREAM_VAR_WARPS + G);  // gather warps + the apply group that owns the tile
    }
    for (int s = 0; s < WS; ++s) {
      ptx::mbar_init(&w_full[s], 1 + STREAM_VAR_WARPS);  // TMA transaction + rhs rows of the gather warps
      ptx::mbar_init(&w_empty[s], G);
    }
    for (int g = 0; g < STREAM_GROUPS; ++g) {
      ptx::mbar_init(&coef_full[g], G);
      ptx::mbar_init(&coef_empty[g], F);
    }
    ptx::fence_barrier_init();
  }
  __syncthreads();

  // =============================== producer ======================================================
  if (warp == 0) {
    if (lane == 0) {
      const std::uint64_t pol = ptx::policy_evict_first();
      int hs = 0, hph = 0, ws = 0, wph = 0;
      for (int m = 0; has_tile(m); ++m) {
        const char *rec = P.rec + tile_of(m) * T::REC_BYTES;
        ptx::mbar_wait(&hdr_empty[hs], hph ^ 1);
        ptx::mbar_expect_tx(&hdr_full[hs], T::HDR_BYTES);
        ptx::bulk_g2s(hdr_base + hs * T::HDR_BYTES, rec, T::HDR_BYTES, &hdr_full[hs], pol);
        if (++hs == HS) hs = 0, hph ^= 1;
#pragma unroll 1
        for (int s = 0; s < N_SEGS; ++s) {
          int off, bytes;
          if (s < N_HI) {
            off = T::off_W(0) + s * R_HI * CHI * TILE * 8;
            bytes = (s == N_HI - 1 ? R_TAIL : R_HI) * CHI * TILE * 8;
          } else {
            const int k0 = 1 + 2 * (s - N_HI);
            off = T::off_W(1) + (k0 - 1) * RLO * CLO * TILE * 8;
            bytes = (k0 + 1 < NS ? 2 : 1) * RLO * CLO * TILE * 8;
          }
          ptx::mbar_wait(&w_empty[ws], wph ^ 1);
          ptx::mbar_expect_tx(&w_full[ws], (std::uint32_t)bytes);
          ptx::bulk_g2s(w_base + ws * T::SLOT_BYTES, rec + off, (std::uint32_t)bytes, &w_full[ws], pol);
          if (++ws == WS) ws = 0, wph ^= 1;
        }
      }
    }
    return;
  }

  // =============================== gather group: thread = (cell, variable) ==========================
  if (warp < FIRST_APPLY_WARP) {
    const int ta = threadIdx.x - 32;
    const int cell = ta / NVARS, var = ta - cell * NVARS;
    constexpr int N_LOW_ROWS = (NS - 1) * RLO;
    // Two register buffers hold the raw neighbour values of a whole tile: the central-stencil rows and the
    // one-sided-stencil rows.  While one half is being scaled and stored into ring slots, the loads of the
    // same half of the NEXT tile are already in flight (17..34 independent loads per thread), which is what
    // hides the L2 / HBM latency of the gather without any occupancy.
    double raw_hi[RM0], raw_lo[N_LOW_ROWS], u0[NVARS];
    double inv_scale_v = 1.0, q0s = 0.0;

    auto load_hi = [&](int m) {  // waits for the tile's header; also fetches the cell's own state
      const int hs = m % HS;
      const unsigned char *hdr = hdr_base + hs * T::HDR_BYTES;
      ptx::mbar_wait(&hdr_full[hs], (m / HS) & 1);
      const std::int64_t cell_idx = min(tile_of(m) * TILE + cell, P.n_cells - 1);
#pragma unroll
      for (int v = 0; v < NVARS; ++v) u0[v] = args.state[cell_idx * NVARS + v];
      const std::int32_t *si = reinterpret_cast<const std::int32_t *>(hdr + T::OFF_SIDX0) + cell;
#pragma unroll
      for (int r = 0; r < RM0; ++r) raw_hi[r] = args.state[(std::int64_t)si[r * TILE] * NVARS + var];
    };
    auto load_lo = [&](int m) {
      const unsigned char *hdr = hdr_base + (m % HS) * T::HDR_BYTES;
      const std::int32_t *si = reinterpret_cast<const std::int32_t *>(hdr + T::off_sidx(1)) + cell;
#pragma unroll
      for (int r = 0; r < N_LOW_ROWS; ++r) raw_lo[r] = args.state[(std::int64_t)si[r * TILE] * NVARS + var];
    };
    // scale and store rows [first, first + count) of `raw` into the rhs part of the ring slot of segment `seg`
    auto store_rows = [&](int m, int seg, const double *raw, auto first_tag, auto count_tag) {
      constexpr int FIRST = decltype(first_tag)::value, COUNT = decltype(count_tag)::value;
      const int gseg = m * N_SEGS + seg;
      const int ws = gseg % WS;
      ptx::mbar_wait(&w_empty[ws], ((gseg / WS) & 1) ^ 1);
      double *rhs = reinterpret_cast<double *>(w_base + ws * T::SLOT_BYTES + T::W_BYTES) + ta;  // [row][cell][5]
#pragma unroll
      for (int r = 0; r < COUNT; ++r) rhs[r * TILE * NVARS] = raw[FIRST + r] * inv_scale_v - q0s;
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&w_full[ws]);
    };
    auto for_each_hi_seg = [&](int m, auto self, auto seg_tag) -> void {
      constexpr int S = decltype(seg_tag)::value;
      if constexpr (S < N_HI) {
        store_rows(m, S, raw_hi, std::integral_constant<int, S * R_HI>{},
                   std::integral_constant<int, (S == N_HI - 1 ? R_TAIL : R_HI)>{});
        self(m, self, std::integral_constant<int, S + 1>{});
      }
    };
    auto for_each_lo_seg = [&](int m, auto self, auto seg_tag) -> void {
      constexpr int S = decltype(seg_tag)::value;
      if constexpr (S < N_LO) {
        constexpr int K0 = 1 + 2 * S;
        store_rows(m, N_HI + S, raw_lo, std::integral_constant<int, (K0 - 1) * RLO>{},
                   std::integral_constant<int, (K0 + 1 < NS ? 2 : 1) * RLO>{});
        self(m, self, std::integral_constant<int, S + 1>{});
      }
    };

    if (has_tile(0)) {
      load_hi(0);
      load_lo(0);
    }
#pragma unroll 1
    for (int m = 0; has_tile(m); ++m) {
      const int hs = m % HS;
      {  // per-tile scalars of this (cell, variable)
        const double ekin0 = 0.5 * (u0[1] * u0[1] + u0[2] * u0[2] + u0[3] * u0[3]) / u0[0];
        const double eint0 = u0[4] - ekin0;
        double scale_v = 1.0;
        if (sc.scaling == SCALING_EULER) {  // characteristic_scale.hpp:24-33
          const double p = eint0 * (sc.gamma - 1.0);
          const double cs = sqrt(sc.gamma * p / u0[0]);
          scale_v = (var == 0) ? u0[0] : ((var == 4) ? eint0 : cs);
        }
        inv_scale_v = 1.0 / scale_v;
        double own = u0[0];
#pragma unroll
        for (int v = 1; v < NVARS; ++v)
          if (var == v) own = u0[v];
        q0s = own * inv_scale_v;
        double *info = info_base + (hs * TILE + cell) * (2 * NVARS);
        info[var] = q0s;
        info[NVARS + var] = scale_v;
      }
      const bool have_nxt = has_tile(m + 1);
      for_each_hi_seg(m, for_each_hi_seg, std::integral_constant<int, 0>{});
      if (have_nxt) load_hi(m + 1);  // raw_hi and u0 are free again
      for_each_lo_seg(m, for_each_lo_seg, std::integral_constant<int, 0>{});
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&hdr_empty[hs]);  // the index rows of this tile are consumed
      if (have_nxt) load_lo(m + 1);
    }
    return;
  }

  // =============================== apply groups: thread = (cell, coefficient group) =================
  if (warp < FIRST_TRACE_WARP) {
    const int grp = (warp - FIRST_APPLY_WARP) / G;
    const int g = (warp - FIRST_APPLY_WARP) - grp * G;
    if (grp >= NG) return;
    const int cell = lane;
    constexpr int N_GROUP_THREADS = 32 * G;
    double *xchg = reinterpret_cast<double *>(smem + cfg.off_xchg + grp * T::XCHG_BYTES);     // partial IS | coefficients
    double *alpha_x = reinterpret_cast<double *>(smem + cfg.off_alpha + grp * T::ALPHA_BYTES);  // [NS][32]

#pragma unroll 1
    for (int m = grp; has_tile(m); m += NG) {
      double lo[NS][NVARS], hi[HPT > 0 ? HPT : 1][NVARS];
#pragma unroll
      for (int k = 0; k < NS; ++k)
#pragma unroll
        for (int v = 0; v < NVARS; ++v) lo[k][v] = 0.0;
#pragma unroll
      for (int h = 0; h < HPT; ++h)
#pragma unroll
        for (int v = 0; v < NVARS; ++v) hi[h][v] = 0.0;

#pragma unroll 1
      for (int seg = 0; seg < N_SEGS; ++seg) {
        const int gseg = m * N_SEGS + seg;
        const int ws = gseg % WS;
        ptx::mbar_wait(&w_full[ws], (gseg / WS) & 1);
        const double *wslot = reinterpret_cast<const double *>(w_base + ws * T::SLOT_BYTES) + g * TILE + cell;
        const double *rslot = reinterpret_cast<const double *>(w_base + ws * T::SLOT_BYTES + T::W_BYTES) + cell * NVARS;
        auto hi_rows = [&](auto n_rows_tag) {
          constexpr int NR = decltype(n_rows_tag)::value;
#pragma unroll
          for (int r = 0; r < NR; ++r) {
            double rhs[NVARS];
#pragma unroll
            for (int v = 0; v < NVARS; ++v) rhs[v] = rslot[r * TILE * NVARS + v];
            const double wl = wslot[(r * CHI) * TILE];  // low-order coefficient g of the central stencil
#pragma unroll
            for (int v = 0; v < NVARS; ++v) lo[0][v] = fma(wl, rhs[v], lo[0][v]);
#pragma unroll
            for (int h = 0; h < HPT; ++h) {
              if (NHI % G == 0 || h * G + g < NHI) {  // high-order coefficient CLO + h*G + g
                const double wh = wslot[(r * CHI + CLO + h * G) * TILE];
#pragma unroll
                for (int v = 0; v < NVARS; ++v) hi[h][v] = fma(wh, rhs[v], hi[h][v]);
              }
            }
          }
        };
        // the segment number is resolved by explicit branches so that lo[k][v] is only ever indexed with
        // compile-time k (the accumulators must stay in registers)
        auto lo_rows = [&](auto seg_tag) {
          constexpr int s = decltype(seg_tag)::value;
          constexpr int k_first = 1 + 2 * s;
          constexpr int ka = k_first < NS ? k_first : 0, kb = k_first + 1 < NS ? k_first + 1 : 0;
#pragma unroll
          for (int r = 0; r < RLO; ++r) {  // the two stencils interleaved: independent accumulator chains
#pragma unroll
            for (int kk = 0; kk < 2; ++kk) {
              if (k_first + kk < NS) {
                double rhs[NVARS];
#pragma unroll
                for (int v = 0; v < NVARS; ++v) rhs[v] = rslot[(kk * RLO + r) * TILE * NVARS + v];
                const double wv = wslot[((kk * RLO + r) * CLO) * TILE];
#pragma unroll
                for (int v = 0; v < NVARS; ++v) {
                  if (kk == 0)
                    lo[ka][v] = fma(wv, rhs[v], lo[ka][v]);
                  else
                    lo[kb][v] = fma(wv, rhs[v], lo[kb][v]);
                }
              }
            }
          }
        };
        if (seg < N_HI) {
          if (R_TAIL != R_HI && seg == N_HI - 1)
            hi_rows(std::integral_constant<int, R_TAIL>{});
          else
            hi_rows(std::integral_constant<int, R_HI>{});
        } else {
          const int ls = seg - N_HI;
          if (ls == 0)
            lo_rows(std::integral_constant<int, 0>{});
          else if (N_LO > 1 && ls == 1)
            lo_rows(std::integral_constant<int, (N_LO > 1 ? 1 : 0)>{});
          else if (N_LO > 2 && ls == 2)
            lo_rows(std::integral_constant<int, (N_LO > 2 ? 2 : 0)>{});
        }
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&w_empty[ws]);
      }

      // ---- hybridise (cweno_ao.cpp:36-53, hybrid_weno.cpp:110-128) -----------------------------------
      const int hs = m % HS;
      const int j_tile = m / NG;  // tiles finished by this group
      const std::uint64_t meta = reinterpret_cast<const std::uint64_t *>(hdr_base + hs * T::HDR_BYTES)[cell];
      const int kh = (int)((meta >> 56) & 0xF);
      const bool single = ((meta >> 60) & 1) != 0;
      const int n_eff = single ? 1 : NS;
      double inv_gh = 1.0;
      // all 32 cells of the warp have the full family with the central stencil as the highest-order one
      // (true away from boundaries): compile-time stencil indices, no select chains
      const bool fast = __all_sync(0xffffffffu, kh == 0 && !single);
      if (sc.recon_mode == RECON_CWENO_AO) {
        if (fast) {
          inv_gh = 1.0 / sc.lin_w[0];
#pragma unroll
          for (int v = 0; v < NVARS; ++v) {
            double cor = lo[0][v];
#pragma unroll
            for (int k = 1; k < NS; ++k) cor -= sc.lin_w[k] * lo[k][v];
            lo[0][v] = inv_gh * cor;
#pragma unroll
            for (int h = 0; h < HPT; ++h) hi[h][v] *= inv_gh;
          }
        } else {
          double gh = 1.0;
#pragma unroll
          for (int k = 0; k < NS; ++k)
            if (k == kh) gh = single ? 1.0 : sc.lin_w[k];
          inv_gh = 1.0 / gh;
#pragma unroll
          for (int v = 0; v < NVARS; ++v) {
            double cor = 0.0;
#pragma unroll
            for (int k = 0; k < NS; ++k)
              if (k == kh) cor = lo[k][v];
#pragma unroll
            for (int k = 0; k < NS; ++k)
              if (k != kh && k < n_eff) cor -= sc.lin_w[k] * lo[k][v];
            const double val = inv_gh * cor;
#pragma unroll
            for (int k = 0; k < NS; ++k)
              if (k == kh) lo[k][v] = val;
          }
          if (kh == 0) {
#pragma unroll
            for (int h = 0; h < HPT; ++h)
#pragma unroll
              for (int v = 0; v < NVARS; ++v) hi[h][v] *= inv_gh;
          }
        }
      }
      // partial smoothness indicators of this thread's coefficients -> shared memory
      ptx::mbar_wait(&coef_empty[grp], (j_tile & 1) ^ 1);  // the exchange area aliases the previous polynomial
#pragma unroll
      for (int k = 0; k < NS; ++k)
#pragma unroll
        for (int v = 0; v < NVARS; ++v) {
          double beta = lo[k][v] * lo[k][v];
          if (k == 0) {
#pragma unroll
            for (int h = 0; h < HPT; ++h) beta += hi[h][v] * hi[h][v];  // unused slots hold exact zeros
          }
          xchg[((g * NS + k) * NVARS + v) * TILE + cell] = beta;
        }
      ptx::named_bar_sync(1 + grp, N_GROUP_THREADS);
      // non-linear weight of stencil k: computed by thread k % G
#pragma unroll
      for (int k = 0; k < NS; ++k) {
        if (k % G == g) {
          double is_max = 0.0;
#pragma unroll
          for (int v = 0; v < NVARS; ++v) {
            // same summation order as a single thread would use: low-order coefficients first (g = 0, 1, ..)
            double beta = xchg[((0 * NS + k) * NVARS + v) * TILE + cell];
#pragma unroll
            for (int gg = 1; gg < G; ++gg) beta += xchg[((gg * NS + k) * NVARS + v) * TILE + cell];
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
          const double gk = single ? 1.0 : sc.lin_w[k];
          alpha_x[k * TILE + cell] = (k < n_eff) ? gk / (sc.epsilon + is_pow) : 0.0;
        }
      }
      ptx::named_bar_sync(1 + grp, N_GROUP_THREADS);
      double wk[NS];
      {
        double al_tot = 0.0;
#pragma unroll
        for (int k = 0; k < NS; ++k) {
          wk[k] = alpha_x[k * TILE + cell];
          al_tot += wk[k];
        }
        const double inv_tot = 1.0 / al_tot;
#pragma unroll
        for (int k = 0; k < NS; ++k) wk[k] *= inv_tot;
      }
      // this thread's coefficients of the hybridised polynomial, times the characteristic scale -> trace group
      const double *info = info_base + (hs * TILE + cell) * (2 * NVARS);
      const std::int64_t ci = tile_of(m) * TILE + cell;
      const bool keep = P.poly != nullptr && ci < P.n_cells;
      double *cx = xchg + cell;
#pragma unroll
      for (int v = 0; v < NVARS; ++v) {
        const double scale_v = info[NVARS + v];
        double c_lo = 0.0;
#pragma unroll
        for (int k = 0; k < NS; ++k) c_lo += wk[k] * lo[k][v];
        cx[((1 + g) * NVARS + v) * COEF_PAD] = c_lo * scale_v;
        if (keep) P.poly[(ci * P.n_poly_coef + 1 + g) * NVARS + v] = c_lo;
#pragma unroll
        for (int h = 0; h < HPT; ++h) {
          if (NHI % G == 0 || h * G + g < NHI) {
            const double c_hi = wk[0] * hi[h][v];
            cx[((1 + CLO + h * G + g) * NVARS + v) * COEF_PAD] = c_hi * scale_v;
            if (keep) P.poly[(ci * P.n_poly_coef + 1 + CLO + h * G + g) * NVARS + v] = c_hi;
          }
        }
        if (g == 0) {  // constant coefficient: q0 for every stencil but kh, whose value carries the CWENO correction
          const double q0 = info[v];
          double a0h = q0;
          if (sc.recon_mode == RECON_CWENO_AO) {
#pragma unroll
            for (int k = 0; k < NS; ++k)
              if (k != kh && k < n_eff) a0h -= sc.lin_w[k] * q0;
            a0h *= inv_gh;
          }
          double c0 = 0.0;
#pragma unroll
          for (int k = 0; k < NS; ++k) c0 += wk[k] * ((k == kh) ? a0h : q0);
          cx[v * COEF_PAD] = c0 * scale_v;
          if (keep) {
            P.poly[(ci * P.n_poly_coef) * NVARS + v] = c0;
            P.poly_scale[ci * NVARS + v] = scale_v;
          }
        }
      }
      __syncwarp();
      if (lane == 0) {
        ptx::mbar_arrive(&coef_full[grp]);
        ptx::mbar_arrive(&hdr_empty[hs]);
      }
    }
    return;
  }

  // =============================== trace group: thread = (cell, face) ===============================
  {
    const int k = warp - FIRST_TRACE_WARP;  // local face
#pragma unroll 1
    for (int m = 0; has_tile(m); ++m) {
      const std::int64_t tile = tile_of(m);
      const std::int64_t cell = tile * TILE + lane;
      const bool active = cell < P.n_cells;
      // geometry of the cell (issued before the wait: the loads overlap the apply groups' work)
      const std::uint32_t fref = active ? ld_stream(P.face_ref + (tile * F + k) * TILE + lane) : 0u;
      const std::uint32_t slots = P.face_slots[(tile * F + k) * TILE + lane];
      double fv[ND][ND];  // face vertices in the left cell's order
#pragma unroll
      for (int r = 0; r < ND; ++r) {
        const int s = (slots >> (2 * r)) & 3;
#pragma unroll
        for (int d = 0; d < ND; ++d) fv[r][d] = P.vtx[((tile * F + s) * 3 + d) * TILE + lane];
      }
      double xc[ND];
#pragma unroll
      for (int d = 0; d < ND; ++d) xc[d] = P.center[(tile * 3 + d) * TILE + lane];
      const double inv_len = P.inv_len[tile * TILE + lane];
      double cmom[D];
#pragma unroll
      for (int i = 0; i < D; ++i) cmom[i] = 0.0;
#pragma unroll
      for (int i = 3; i < D; ++i) cmom[i] = P.moments[(tile * P.n_mom + (i - 3)) * TILE + lane];
      const std::int64_t e = fref & FREF_EDGE_MASK;
      const int side = (fref & FREF_SIDE) ? 1 : 0;
      const bool want_trace = (fref & FREF_TRACE) != 0;

      double *stage = reinterpret_cast<double *>(smem + cfg.off_stage) + k * (TILE * cfg.stage_pitch + TILE);
      long long *stage_base = reinterpret_cast<long long *>(stage + TILE * cfg.stage_pitch);
      const int grp = m % NG;
      ptx::mbar_wait(&coef_full[grp], (m / NG) & 1);
      const double *cx = reinterpret_cast<const double *>(smem + cfg.off_xchg + grp * T::XCHG_BYTES) + lane;
      constexpr int VP = T::VARS_PER_PASS;
#pragma unroll
      for (int v0 = 0; v0 < NVARS; v0 += VP) {
        double coef[D][VP];
#pragma unroll
        for (int i = 0; i < D; ++i)
#pragma unroll
          for (int v = 0; v < VP; ++v)
            if (v0 + v < NVARS) coef[i][v] = cx[(i * NVARS + v0 + v) * COEF_PAD];
        if (v0 + VP >= NVARS) {  // last pass: the exchange buffer can be refilled
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(&coef_empty[grp]);
        }
        {  // every lane evaluates (lanes without a trace hold finite garbage that is never written out)
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
            for (int v = 0; v < VP; ++v) {
              if (v0 + v < NVARS) {
                double s = coef[0][v];
#pragma unroll
                for (int i = 1; i < D; ++i) s = fma(coef[i][v], mono[i], s);
                stage[lane * cfg.stage_pitch + q * NVARS + v0 + v] = s;
              }
            }
          }
        }
      }
      // coalesced write-out: consecutive lanes write consecutive doubles of a cell's 40*q_f-byte trace block
      // (a lane-per-cell store would touch 32 cache lines per instruction)
      stage_base[lane] = want_trace ? (long long)(((e * 2 + side) * sc.q_f) * NVARS) : -1ll;
      __syncwarp();
      {
        const int qv = sc.q_f * NVARS;
        int owner = lane / qv, j = lane - owner * qv;
        const int step_o = TILE / qv, step_j = TILE - step_o * qv;
        for (int it = 0; it < qv; ++it) {
          const long long base = stage_base[owner];
          if (base >= 0) P.trace[base + j] = stage[owner * cfg.stage_pitch + j];
          owner += step_o;
          j += step_j;
          if (j >= qv) {
            j -= qv;
            ++owner;
          }
        }
      }
      __syncwarp();  // the staging area is reused for the next tile
    }
  }
}

/// Shared-memory plan of the streaming kernel; returns false if the kernel does not apply to this plan.
template <int ND, int DEG_HI, int DEG_LO, int NS, int RM0, int RLO>
bool stream_config(const DevicePlan &P, const SchemeConst &sc, int smem_budget, StreamCfg &c) {
  using T = StreamTraits<ND, DEG_HI, DEG_LO, NS, RM0, RLO>;
  if (NS < 2 || T::CLO < 1 || sc.n_stencils != NS) return false;
  if (sc.rows_max[0] != RM0 || sc.ncoef[0] != T::CHI) return false;
  for (int k = 1; k < NS; ++k)
    if (sc.rows_max[k] != RLO || sc.ncoef[k] != T::CLO) return false;
  if (P.rec_bytes != T::REC_BYTES || P.hdr_bytes != T::HDR_BYTES) return false;
  for (int k = 0; k < NS; ++k)
    if (P.off_sidx[k] != T::off_sidx(k) || P.off_W[k] != T::off_W(k)) return false;
  c.stage_pitch = (sc.q_f * NVARS) | 1;
  const int stage_bytes = T::F * (TILE * c.stage_pitch * 8 + TILE * 8);  // values + one block base per lane
  const int fixed = STREAM_BARS_BYTES + STREAM_HDR_SLOTS * (T::HDR_BYTES + T::INFO_BYTES) +
                    STREAM_GROUPS * (T::XCHG_BYTES + T::ALPHA_BYTES) + stage_bytes;
  int ws = (smem_budget - fixed) / T::SLOT_BYTES;
  if (ws > 12) ws = 12;
  // mbarrier waits see one parity bit: an apply group may only wait for use u of a slot once use u-1 has
  // completed.  Its previous segment is N_SEGS + 1 ring positions back when it moves on to its next tile
  // (the other group's tile lies in between), and fills complete in ring order, so WS >= N_SEGS + 1 makes
  // use u-1 of the slot (WS positions back) no younger than a segment the group has already consumed.
  c.n_groups = (ws >= T::N_SEGS + 1) ? STREAM_GROUPS : 1;
  if (ws < 2) return false;
  c.n_w_slots = ws;
  c.off_hdr = STREAM_BARS_BYTES;
  c.off_info = c.off_hdr + STREAM_HDR_SLOTS * T::HDR_BYTES;
  c.off_w = c.off_info + STREAM_HDR_SLOTS * T::INFO_BYTES;
  c.off_xchg = c.off_w + ws * T::SLOT_BYTES;
  c.off_alpha = c.off_xchg + STREAM_GROUPS * T::XCHG_BYTES;
  c.off_stage = c.off_alpha + STREAM_GROUPS * T::ALPHA_BYTES;
  c.total_bytes = c.off_stage + stage_bytes;
  return 2 * STREAM_HDR_SLOTS + 2 * STREAM_GROUPS + 2 * ws <= STREAM_BARS_BYTES / 8;
}

}  // namespace zfvm
