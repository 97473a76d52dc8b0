// K1, streaming form: persistent, warp-specialised reconstruction kernel for sm_100a.
//
// Same arithmetic as recon.cuh (EulerGlobalReconstruction::compute + LocalReconstruction::compute +
// HybridWENO::compute_polys_impl / eno_hybridize + CWENO_AO::reconstruct_impl + rc(i)(x) at the face
// Gauss points; reference lines are listed there), restructured around the memory system:
//
//   * one CTA per SM, looping over tiles (32 cells) with stride gridDim.x;
//   * warp 0 (one elected lane) is the *producer*: it streams each tile record -- header (meta +
//     stencil member indices) and pseudo-inverse weights in row segments of <= ~24 KB -- from HBM into
//     shared-memory rings with TMA bulk copies (cp.async.bulk.shared::cluster.global + mbarrier
//     complete_tx), L2 evict-first, so the bytes in flight do not depend on occupancy;
//   * two *apply* groups of five warps take alternate tiles: thread = (cell, variable), the five
//     variables of a cell sit in adjacent lanes, so a weight is read from shared memory once per
//     warp-row as a broadcast and a gathered neighbour state row (40 B) is read by five adjacent
//     lanes.  They accumulate coef = W_k * rhs, exchange smoothness indicators through shared
//     memory, hybridise, and hand the final polynomial (times the characteristic scale) to the trace
//     group through shared memory;
//   * the *trace* group (one warp per local face, thread = (cell, face)) evaluates the polynomial at
//     the face Gauss points and writes trace[e][side][q][5].
//
// The gather of segment s+1 is issued before the FMAs of segment s, so L2 latency overlaps the math.
// Stencil sizes are compile-time (RM0 rows for the central stencil, RLO for every one-sided one): all
// shared-memory offsets fold into immediates and no row loop carries a predicate.  Other stencil sizes
// use the thread-per-cell kernel of recon.cuh.
#pragma once
#include <type_traits>

#include "recon.cuh"

namespace zfvm {

struct StreamCfg {
  int n_w_slots;
  int off_hdr, off_w, off_coef, off_is;  // byte offsets into dynamic shared memory (barriers at 0)
  int total_bytes;
};

constexpr int STREAM_VAR_WARPS = 5;  // one apply group: 160 threads = 32 cells x 5 variables
constexpr int STREAM_GROUPS = 2;     // apply groups (alternate tiles)
constexpr int STREAM_HDR_SLOTS = 4;
constexpr int STREAM_BARS_BYTES = 1024;
constexpr int COEF_PAD = 33;         // coef exchange row pitch (doubles): spreads (cell, var) writes over banks

template <int ND, int DEG_HI, int DEG_LO, int NS, int RM0, int RLO>
struct StreamTraits {
  static constexpr int F = ND + 1;
  static constexpr int D = dof_of(DEG_HI, ND);
  static constexpr int CHI = D - 1;
  static constexpr int CLO = dof_of(DEG_LO, ND) - 1;
  static constexpr int NHI = CHI - CLO;
  static constexpr int R_CAP0 = 24576 / (CHI * TILE * 8);
  static constexpr int R_CAP = R_CAP0 < 1 ? 1 : (R_CAP0 > 12 ? 12 : R_CAP0);
  static constexpr int N_HI = (RM0 + R_CAP - 1) / R_CAP;       // central-stencil segments
  static constexpr int R_HI = (RM0 + N_HI - 1) / N_HI;         // rows per segment (balanced)
  static constexpr int R_TAIL = RM0 - (N_HI - 1) * R_HI;       // rows of the last one
  static constexpr int N_LO = NS / 2;                          // two one-sided stencils per segment
  static constexpr int N_SEGS = N_HI + N_LO;
  static constexpr int RAW = R_HI > 2 * RLO ? R_HI : 2 * RLO;
  static constexpr int HI_BYTES = R_HI * CHI * TILE * 8;
  static constexpr int LO_BYTES = 2 * RLO * CLO * TILE * 8;
  static constexpr int SLOT_BYTES = HI_BYTES > LO_BYTES ? HI_BYTES : LO_BYTES;
  // tile record layout (device/layout.hpp) for these stencil sizes
  static constexpr int OFF_SIDX0 = TILE * 8;
  static constexpr __host__ __device__ int off_sidx(int k) { return OFF_SIDX0 + 4 * TILE * (k == 0 ? 0 : RM0 + (k - 1) * RLO); }
  static constexpr int HDR_BYTES = OFF_SIDX0 + 4 * TILE * (RM0 + (NS - 1) * RLO);
  static constexpr __host__ __device__ int off_W(int k) { return HDR_BYTES + 8 * TILE * (k == 0 ? 0 : RM0 * CHI + (k - 1) * RLO * CLO); }
  static constexpr int REC_BYTES = HDR_BYTES + 8 * TILE * (RM0 * CHI + (NS - 1) * RLO * CLO);
  static constexpr int COEF_BYTES = D * NVARS * COEF_PAD * 8;     // per apply group
  static constexpr int IS_BYTES = 2 * NS * NVARS * TILE * 8;      // per apply group (two parities)
  static constexpr int N_WARPS = 1 + STREAM_GROUPS * STREAM_VAR_WARPS + F;
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
  constexpr int F = T::F, D = T::D, CHI = T::CHI, CLO = T::CLO, NHI = T::NHI;
  constexpr int N_HI = T::N_HI, R_HI = T::R_HI, R_TAIL = T::R_TAIL, N_LO = T::N_LO, N_SEGS = T::N_SEGS;
  constexpr int RAW = T::RAW, HS = STREAM_HDR_SLOTS, NG = STREAM_GROUPS;
  constexpr int N_APPLY = 32 * STREAM_VAR_WARPS;
  const DevicePlan &P = args.plan;

  extern __shared__ __align__(128) unsigned char smem[];
  std::uint64_t *bars = reinterpret_cast<std::uint64_t *>(smem);
  const int WS = cfg.n_w_slots;
  std::uint64_t *hdr_full = bars, *hdr_empty = bars + HS;
  std::uint64_t *coef_full = bars + 2 * HS, *coef_empty = coef_full + NG;
  std::uint64_t *w_full = coef_empty + NG, *w_empty = w_full + WS;
  unsigned char *hdr_base = smem + cfg.off_hdr;
  unsigned char *w_base = smem + cfg.off_w;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const std::int64_t n_launch = args.n_tiles_launch;

  if (threadIdx.x == 0) {
    for (int s = 0; s < HS; ++s) {
      ptx::mbar_init(&hdr_full[s], 1);
      ptx::mbar_init(&hdr_empty[s], STREAM_VAR_WARPS);
    }
    for (int s = 0; s < WS; ++s) {
      ptx::mbar_init(&w_full[s], 1);
      ptx::mbar_init(&w_empty[s], STREAM_VAR_WARPS);
    }
    for (int g = 0; g < NG; ++g) {
      ptx::mbar_init(&coef_full[g], STREAM_VAR_WARPS);
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
      for (std::int64_t idx = blockIdx.x; idx < n_launch; idx += gridDim.x) {
        const std::int64_t tile = args.tile_list ? (std::int64_t)args.tile_list[idx] : idx;
        const char *rec = P.rec + tile * T::REC_BYTES;
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

  // =============================== apply groups: thread = (cell, variable) ==========================
  if (warp <= NG * STREAM_VAR_WARPS) {
    const int grp = (warp - 1) / STREAM_VAR_WARPS;
    const int ta = threadIdx.x - 32 - grp * N_APPLY;
    const int cell = ta / NVARS, var = ta - cell * NVARS;
    double *coef_x = reinterpret_cast<double *>(smem + cfg.off_coef + grp * T::COEF_BYTES);
    double *is_x = reinterpret_cast<double *>(smem + cfg.off_is + grp * T::IS_BYTES);

    // cursor over (tile, segment): m = position of the tile in this CTA's sequence
    int cur_m = grp, cur_seg = 0;
    // u0: own-cell state of the tile whose first segment was gathered last; it is consumed at that
    // tile's first segment, before the next tile's first gather overwrites it
    double raw_cur[RAW], raw_nxt[RAW], u0c[NVARS];

    auto tile_of = [&](int m) -> std::int64_t {
      const std::int64_t idx = blockIdx.x + (std::int64_t)m * gridDim.x;
      return args.tile_list ? (std::int64_t)args.tile_list[idx] : idx;
    };
    auto has_tile = [&](int m) { return blockIdx.x + (std::int64_t)m * gridDim.x < n_launch; };

    // issue the gather of one segment: raw neighbour values of this thread's variable
    auto issue_gather = [&](int m, int seg, double *raw) {
      const int hs = m % HS;
      const unsigned char *hdr = hdr_base + hs * T::HDR_BYTES;
      if (seg == 0) {
        ptx::mbar_wait(&hdr_full[hs], (m / HS) & 1);
        const std::int64_t cell_idx = min(tile_of(m) * TILE + cell, P.n_cells - 1);
#pragma unroll
        for (int v = 0; v < NVARS; ++v) u0c[v] = args.state[cell_idx * NVARS + v];
      }
      if (seg < N_HI) {
        const std::int32_t *si = reinterpret_cast<const std::int32_t *>(hdr + T::OFF_SIDX0) + seg * R_HI * TILE + cell;
        if (R_TAIL != R_HI && seg == N_HI - 1) {
#pragma unroll
          for (int r = 0; r < R_TAIL; ++r) raw[r] = args.state[(std::int64_t)si[r * TILE] * NVARS + var];
        } else {
#pragma unroll
          for (int r = 0; r < R_HI; ++r) raw[r] = args.state[(std::int64_t)si[r * TILE] * NVARS + var];
        }
      } else {
        // one-sided stencils k0, k0 + 1: their index rows are contiguous in the header
        const int k0 = 1 + 2 * (seg - N_HI);
        const std::int32_t *si = reinterpret_cast<const std::int32_t *>(hdr + T::off_sidx(1)) + (k0 - 1) * RLO * TILE + cell;
        if ((NS - 1) % 2 == 1 && k0 + 1 >= NS) {
#pragma unroll
          for (int r = 0; r < RLO; ++r) raw[r] = args.state[(std::int64_t)si[r * TILE] * NVARS + var];
        } else {
#pragma unroll
          for (int r = 0; r < 2 * RLO; ++r) raw[r] = args.state[(std::int64_t)si[r * TILE] * NVARS + var];
        }
      }
    };

    // per-tile state
    double lo[NS][CLO > 0 ? CLO : 1], hi[NHI > 0 ? NHI : 1];
    double scale_v = 1.0, inv_scale_v = 1.0, q0s = 0.0;
    std::uint64_t meta = 0;

    bool have_cur = has_tile(cur_m);
    if (have_cur) issue_gather(cur_m, 0, raw_cur);
#pragma unroll 1
    while (have_cur) {
      int nxt_m = cur_m, nxt_seg = cur_seg + 1;
      if (nxt_seg == N_SEGS) {
        nxt_seg = 0;
        nxt_m += NG;
      }
      const bool have_nxt = has_tile(nxt_m);

      if (cur_seg == 0) {  // N_SEGS >= 2: the gather issued below never belongs to another tile here
        const unsigned char *hdr = hdr_base + (cur_m % HS) * T::HDR_BYTES;
        meta = reinterpret_cast<const std::uint64_t *>(hdr)[cell];
        const double ekin0 = 0.5 * (u0c[1] * u0c[1] + u0c[2] * u0c[2] + u0c[3] * u0c[3]) / u0c[0];
        const double eint0 = u0c[4] - ekin0;
        if (sc.scaling == SCALING_EULER) {  // characteristic_scale.hpp:24-33
          const double p = eint0 * (sc.gamma - 1.0);
          const double cs = sqrt(sc.gamma * p / u0c[0]);
          scale_v = (var == 0) ? u0c[0] : ((var == 4) ? eint0 : cs);
        } else {
          scale_v = 1.0;
        }
        inv_scale_v = 1.0 / scale_v;
        double own = u0c[0];
#pragma unroll
        for (int v = 1; v < NVARS; ++v)
          if (var == v) own = u0c[v];
        q0s = own * inv_scale_v;
#pragma unroll
        for (int k = 0; k < NS; ++k)
#pragma unroll
          for (int c = 0; c < CLO; ++c) lo[k][c] = 0.0;
#pragma unroll
        for (int c = 0; c < NHI; ++c) hi[c] = 0.0;
      }

      if (have_nxt) issue_gather(nxt_m, nxt_seg, raw_nxt);

      // ---- coef += W_seg * rhs ------------------------------------------------------------------
      const int gseg = cur_m * N_SEGS + cur_seg;  // position in the weight ring
      const int ws = gseg % WS;
      ptx::mbar_wait(&w_full[ws], (gseg / WS) & 1);
      const double *wslot = reinterpret_cast<const double *>(w_base + ws * T::SLOT_BYTES) + cell;
      auto hi_rows = [&](auto n_rows_tag) {
        constexpr int NR = decltype(n_rows_tag)::value;
#pragma unroll
        for (int r = 0; r < NR; ++r) {
          const double rhs = raw_cur[r] * inv_scale_v - q0s;
#pragma unroll
          for (int c = 0; c < CHI; ++c) {
            const double wv = wslot[(r * CHI + c) * TILE];
            if (c < CLO)
              lo[0][c] = fma(wv, rhs, lo[0][c]);
            else
              hi[c - CLO] = fma(wv, rhs, hi[c - CLO]);
          }
        }
      };
      // the segment number is resolved by explicit branches so that lo[k][c] is only ever indexed with
      // compile-time k (the accumulators must stay in registers)
      auto lo_rows = [&](auto seg_tag) {
        constexpr int s = decltype(seg_tag)::value;
#pragma unroll
        for (int kk = 0; kk < 2; ++kk) {
          constexpr int k_first = 1 + 2 * s;
          if (k_first + kk < NS) {
#pragma unroll
            for (int r = 0; r < RLO; ++r) {
              const double rhs = raw_cur[kk * RLO + r] * inv_scale_v - q0s;
#pragma unroll
              for (int c = 0; c < CLO; ++c) {
                const double wv = wslot[((kk * RLO + r) * CLO + c) * TILE];
                constexpr int ka = k_first < NS ? k_first : 0, kb = k_first + 1 < NS ? k_first + 1 : 0;
                if (kk == 0)
                  lo[ka][c] = fma(wv, rhs, lo[ka][c]);
                else
                  lo[kb][c] = fma(wv, rhs, lo[kb][c]);
              }
            }
          }
        }
      };
      if (cur_seg < N_HI) {
        if (R_TAIL != R_HI && cur_seg == N_HI - 1)
          hi_rows(std::integral_constant<int, R_TAIL>{});
        else
          hi_rows(std::integral_constant<int, R_HI>{});
      } else {
        const int ls = cur_seg - N_HI;
        if (ls == 0)
          lo_rows(std::integral_constant<int, 0>{});
        else if (N_LO > 1 && ls == 1)
          lo_rows(std::integral_constant<int, (N_LO > 1 ? 1 : 0)>{});
        else if (N_LO > 2 && ls == 2)
          lo_rows(std::integral_constant<int, (N_LO > 2 ? 2 : 0)>{});
      }
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&w_empty[ws]);

      // ---- end of tile: hybridise and hand over ----------------------------------------------------
      if (cur_seg == N_SEGS - 1) {
        if (lane == 0) ptx::mbar_arrive(&hdr_empty[cur_m % HS]);  // ordered after the __syncwarp above
        const int j_tile = cur_m / NG;  // tiles finished by this group
        const int kh = (int)((meta >> 56) & 0xF);
        const bool single = ((meta >> 60) & 1) != 0;
        const int n_eff = single ? 1 : NS;
        double a0h = q0s;  // constant coefficient of stencil kh after the CWENO correction
        if (sc.recon_mode == RECON_CWENO_AO) {  // cweno_ao.cpp:41-50
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
              a0h -= g * q0s;
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
        // smoothness indicators: this thread's variable -> shared memory -> max over variables
        double *isb = is_x + (j_tile & 1) * (NS * NVARS * TILE);
#pragma unroll
        for (int k = 0; k < NS; ++k) {
          double beta = 0.0;
#pragma unroll
          for (int c = 0; c < CLO; ++c) beta += lo[k][c] * lo[k][c];
          if (k == 0) {
#pragma unroll
            for (int c = 0; c < NHI; ++c) beta += hi[c] * hi[c];
          }
          isb[(k * NVARS + var) * TILE + cell] = beta;
        }
        ptx::named_bar_sync(1 + grp, N_APPLY);
        double alpha[NS];
        double al_tot = 0.0;
#pragma unroll
        for (int k = 0; k < NS; ++k) {
          double is_max = isb[(k * NVARS) * TILE + cell];
#pragma unroll
          for (int v = 1; v < NVARS; ++v) is_max = fmax(is_max, isb[(k * NVARS + v) * TILE + cell]);
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
        double coef[D];
#pragma unroll
        for (int i = 0; i < D; ++i) coef[i] = 0.0;
#pragma unroll
        for (int k = 0; k < NS; ++k) {
          const double wk = alpha[k] / al_tot;
          coef[0] += wk * ((k == kh) ? a0h : q0s);
#pragma unroll
          for (int c = 0; c < CLO; ++c) coef[1 + c] += wk * lo[k][c];
          if (k == 0) {
#pragma unroll
            for (int c = 0; c < NHI; ++c) coef[1 + CLO + c] += wk * hi[c];
          }
        }
        if (P.poly != nullptr) {
          const std::int64_t ci = tile_of(cur_m) * TILE + cell;
          if (ci < P.n_cells) {
            for (int i = 0; i < D; ++i) P.poly[(ci * P.n_poly_coef + i) * NVARS + var] = coef[i];
            P.poly_scale[ci * NVARS + var] = scale_v;
          }
        }
        // hand the polynomial (times the characteristic scale) to the trace group
        ptx::mbar_wait(&coef_empty[grp], (j_tile & 1) ^ 1);
        double *cx = coef_x + var * COEF_PAD + cell;
#pragma unroll
        for (int i = 0; i < D; ++i) cx[i * NVARS * COEF_PAD] = coef[i] * scale_v;
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&coef_full[grp]);
      }

#pragma unroll
      for (int r = 0; r < RAW; ++r) raw_cur[r] = raw_nxt[r];
      cur_m = nxt_m;
      cur_seg = nxt_seg;
      have_cur = have_nxt;
    }
    return;
  }

  // =============================== trace group: thread = (cell, face) ===============================
  {
    const int k = warp - (1 + NG * STREAM_VAR_WARPS);  // local face
    std::int64_t m = 0;
#pragma unroll 1
    for (std::int64_t idx = blockIdx.x; idx < n_launch; idx += gridDim.x, ++m) {
      const std::int64_t tile = args.tile_list ? (std::int64_t)args.tile_list[idx] : idx;
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

      const int grp = (int)(m % NG);
      ptx::mbar_wait(&coef_full[grp], (int)((m / NG) & 1));
      const double *cx = reinterpret_cast<const double *>(smem + cfg.off_coef + grp * T::COEF_BYTES) + lane;
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
        if (want_trace) {
#pragma unroll
          for (int q = 0; q < sc.q_f; ++q) {
            {
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
              double *tr = P.trace + ((e * 2 + side) * sc.q_f + q) * NVARS;
#pragma unroll
              for (int v = 0; v < VP; ++v) {
                if (v0 + v < NVARS) {
                  double s = coef[0][v];
#pragma unroll
                  for (int i = 1; i < D; ++i) s = fma(coef[i][v], mono[i], s);
                  tr[v0 + v] = s;
                }
              }
            }
          }
        }
      }
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
  const int fixed = STREAM_BARS_BYTES + STREAM_HDR_SLOTS * T::HDR_BYTES + STREAM_GROUPS * (T::COEF_BYTES + T::IS_BYTES);
  int ws = (smem_budget - fixed) / T::SLOT_BYTES;
  if (ws > 12) ws = 12;
  if (ws < 3) return false;
  c.n_w_slots = ws;
  c.off_hdr = STREAM_BARS_BYTES;
  c.off_w = c.off_hdr + STREAM_HDR_SLOTS * T::HDR_BYTES;
  c.off_coef = c.off_w + ws * T::SLOT_BYTES;
  c.off_is = c.off_coef + STREAM_GROUPS * T::COEF_BYTES;
  c.total_bytes = c.off_is + STREAM_GROUPS * T::IS_BYTES;
  return 2 * STREAM_HDR_SLOTS + 2 * STREAM_GROUPS + 2 * ws <= STREAM_BARS_BYTES / 8;
}

}  // namespace zfvm
