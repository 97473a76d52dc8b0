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
//   * warps 1..5 are the *apply* group: thread = (cell, variable), the five variables of a cell sit in
//     adjacent lanes, so a weight is read from shared memory once per warp-row as a broadcast and the
//     gathered neighbour state row (40 B) is read by five adjacent lanes.  They accumulate
//     coef = W_k * rhs, exchange smoothness indicators through shared memory, hybridise, and hand the
//     final polynomial (times the characteristic scale) to the trace group through a double buffer;
//   * warps 6..6+F-1 are the *trace* group: thread = (cell, face); they evaluate the polynomial at the
//     face Gauss points and write trace[e][side][q][5].
//
// The gather of segment s+1 is issued before the FMAs of segment s, so L2 latency overlaps the math.
#pragma once
#include <type_traits>

#include "recon.cuh"

namespace zfvm {

struct StreamCfg {
  int n_hdr_slots, n_w_slots;
  int hdr_bytes, slot_bytes;
  int off_bars, off_hdr, off_w, off_coef, off_is;  // byte offsets into dynamic shared memory
  int n_hi_segs, n_lo_segs;
  int total_bytes;
};

constexpr int STREAM_NVAR_WARPS = 5;  // apply group: 160 threads = 32 cells x 5 variables
constexpr int COEF_PAD = 33;          // coef exchange row pitch (doubles): spreads (cell, var) writes over banks

constexpr __host__ __device__ int stream_r_hi(int chi) {
  int r = 24576 / (chi * TILE * 8);
  return r < 1 ? 1 : (r > 12 ? 12 : r);
}
constexpr __host__ __device__ int stream_rlo_max(int nd) { return nd == 2 ? 4 : 6; }

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
ZFVM_DEVICE void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
ZFVM_DEVICE void named_bar_sync(int id, int n_threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n_threads) : "memory");
}
}  // namespace ptx

struct RingPos {
  int slot = 0, phase = 0;
  ZFVM_DEVICE void advance(int n) {
    if (++slot == n) {
      slot = 0;
      phase ^= 1;
    }
  }
};

template <int ND, int DEG_HI, int DEG_LO, int NS>
__global__ void __launch_bounds__(32 * (1 + STREAM_NVAR_WARPS + ND + 1), 1)  // 9-10 warps: 168 registers
    recon_stream_kernel(const __grid_constant__ ReconArgs args, const __grid_constant__ SchemeConst sc,
                        const __grid_constant__ StreamCfg cfg) {
  constexpr int F = ND + 1;
  constexpr int D = dof_of(DEG_HI, ND);
  constexpr int CHI = dof_of(DEG_HI, ND) - 1;
  constexpr int CLO = dof_of(DEG_LO, ND) - 1;
  constexpr int NHI = CHI - CLO;
  constexpr int R_HI = stream_r_hi(CHI);
  constexpr int RLO = stream_rlo_max(ND);
  constexpr int RAW = (R_HI > 2 * RLO) ? R_HI : 2 * RLO;
  constexpr int NLO_SEGS = NS / 2;  // (NS - 1 + 1) / 2: two low-order stencils per segment
  constexpr int N_APPLY = 32 * STREAM_NVAR_WARPS;
  const DevicePlan &P = args.plan;

  extern __shared__ __align__(128) unsigned char smem[];
  std::uint64_t *bars = reinterpret_cast<std::uint64_t *>(smem + cfg.off_bars);
  const int HS = cfg.n_hdr_slots, WS = cfg.n_w_slots;
  std::uint64_t *hdr_full = bars, *hdr_empty = bars + HS;
  std::uint64_t *w_full = bars + 2 * HS, *w_empty = bars + 2 * HS + WS;
  std::uint64_t *coef_full = bars + 2 * HS + 2 * WS, *coef_empty = coef_full + 2;
  unsigned char *hdr_base = smem + cfg.off_hdr;
  unsigned char *w_base = smem + cfg.off_w;
  double *coef_base = reinterpret_cast<double *>(smem + cfg.off_coef);  // [2][D][5][COEF_PAD]
  double *is_base = reinterpret_cast<double *>(smem + cfg.off_is);      // [2][NS][5][32]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const std::int64_t n_launch = args.n_tiles_launch;
  const int n_segs = cfg.n_hi_segs + cfg.n_lo_segs;

  if (threadIdx.x == 0) {
    for (int s = 0; s < HS; ++s) {
      ptx::mbar_init(&hdr_full[s], 1);
      ptx::mbar_init(&hdr_empty[s], STREAM_NVAR_WARPS);
    }
    for (int s = 0; s < WS; ++s) {
      ptx::mbar_init(&w_full[s], 1);
      ptx::mbar_init(&w_empty[s], STREAM_NVAR_WARPS);
    }
    for (int s = 0; s < 2; ++s) {
      ptx::mbar_init(&coef_full[s], STREAM_NVAR_WARPS);
      ptx::mbar_init(&coef_empty[s], F);
    }
    ptx::fence_barrier_init();
  }
  __syncthreads();

  // =============================== producer ======================================================
  if (warp == 0) {
    if (lane == 0) {
      const std::uint64_t pol = ptx::policy_evict_first();
      RingPos h, w;
      for (std::int64_t idx = blockIdx.x; idx < n_launch; idx += gridDim.x) {
        const std::int64_t tile = args.tile_list ? (std::int64_t)args.tile_list[idx] : idx;
        const char *rec = P.rec + tile * P.rec_bytes;
        ptx::mbar_wait(&hdr_empty[h.slot], h.phase ^ 1);
        ptx::mbar_expect_tx(&hdr_full[h.slot], (std::uint32_t)cfg.hdr_bytes);
        ptx::bulk_g2s(hdr_base + h.slot * cfg.hdr_bytes, rec, (std::uint32_t)cfg.hdr_bytes, &hdr_full[h.slot], pol);
        h.advance(HS);
        for (int s = 0; s < n_segs; ++s) {
          int off, bytes;
          if (s < cfg.n_hi_segs) {
            const int j0 = s * R_HI;
            const int nr = min(R_HI, sc.rows_max[0] - j0);
            off = P.off_W[0] + j0 * CHI * TILE * 8;
            bytes = nr * CHI * TILE * 8;
          } else {
            const int k0 = 1 + 2 * (s - cfg.n_hi_segs);
            off = P.off_W[k0];
            bytes = sc.rows_max[k0] * CLO * TILE * 8;
            if (k0 + 1 < NS) bytes += sc.rows_max[k0 + 1] * CLO * TILE * 8;
          }
          ptx::mbar_wait(&w_empty[w.slot], w.phase ^ 1);
          ptx::mbar_expect_tx(&w_full[w.slot], (std::uint32_t)bytes);
          ptx::bulk_g2s(w_base + w.slot * cfg.slot_bytes, rec + off, (std::uint32_t)bytes, &w_full[w.slot], pol);
          w.advance(WS);
        }
      }
    }
    return;
  }

  // =============================== apply group: thread = (cell, variable) ==========================
  if (warp <= STREAM_NVAR_WARPS) {
    const int ta = threadIdx.x - 32;
    const int cell = ta / NVARS, var = ta - cell * NVARS;

    struct Cursor {
      std::int64_t idx;
      int seg;
      RingPos h;  // header slot of the cursor's tile
    };
    Cursor cur{(std::int64_t)blockIdx.x, 0, RingPos()};
    RingPos w;        // weight ring position of the current segment
    int m_tile = 0;   // tiles finished by this CTA (coef / IS double buffers)
    double raw_cur[RAW], raw_nxt[RAW], u0c[NVARS], u0n[NVARS];
    std::int64_t cell_cur = 0, cell_nxt = 0;

    // issue the gather of one segment (raw neighbour values of this thread's variable)
    auto issue_gather = [&](const Cursor &c, double *raw, double *u0, std::int64_t &cell_idx) {
      const std::int64_t tile = args.tile_list ? (std::int64_t)args.tile_list[c.idx] : c.idx;
      const unsigned char *hdr = hdr_base + c.h.slot * cfg.hdr_bytes;
      if (c.seg == 0) {
        ptx::mbar_wait(&hdr_full[c.h.slot], c.h.phase);
        cell_idx = min(tile * TILE + cell, P.n_cells - 1);
#pragma unroll
        for (int v = 0; v < NVARS; ++v) u0[v] = args.state[cell_idx * NVARS + v];
      }
      if (c.seg < cfg.n_hi_segs) {
        const std::int32_t *si = reinterpret_cast<const std::int32_t *>(hdr + P.off_sidx[0]) + cell;
        const int j0 = c.seg * R_HI;
        const int nr = sc.rows_max[0] - j0;
#pragma unroll
        for (int r = 0; r < R_HI; ++r)
          if (r < nr) raw[r] = args.state[(std::int64_t)si[(j0 + r) * TILE] * NVARS + var];
      } else {
        const int k0 = 1 + 2 * (c.seg - cfg.n_hi_segs);
#pragma unroll
        for (int kk = 0; kk < 2; ++kk) {
          const int k = k0 + kk;
          if (k < NS) {
            const std::int32_t *si = reinterpret_cast<const std::int32_t *>(hdr + P.off_sidx[k]) + cell;
            const int nr = sc.rows_max[k];
#pragma unroll
            for (int r = 0; r < RLO; ++r)
              if (r < nr) raw[kk * RLO + r] = args.state[(std::int64_t)si[r * TILE] * NVARS + var];
          }
        }
      }
    };

    // per-tile state
    double lo[NS][CLO > 0 ? CLO : 1], hi[NHI > 0 ? NHI : 1];
    double scale_v = 1.0, inv_scale_v = 1.0, q0s = 0.0;
    std::uint64_t meta = 0;

    bool have_cur = cur.idx < n_launch;
    if (have_cur) issue_gather(cur, raw_cur, u0c, cell_cur);
    while (have_cur) {
      Cursor nxt = cur;
      if (++nxt.seg == n_segs) {
        nxt.seg = 0;
        nxt.idx += gridDim.x;
        nxt.h.advance(HS);
      }
      const bool have_nxt = nxt.idx < n_launch;
      if (have_nxt) issue_gather(nxt, raw_nxt, u0n, cell_nxt);

      if (cur.seg == 0) {
        const unsigned char *hdr = hdr_base + cur.h.slot * cfg.hdr_bytes;
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

      // ---- coef += W_seg * rhs ------------------------------------------------------------------
      ptx::mbar_wait(&w_full[w.slot], w.phase);
      const double *wslot = reinterpret_cast<const double *>(w_base + w.slot * cfg.slot_bytes) + cell;
      if (cur.seg < cfg.n_hi_segs) {
        const int nr = sc.rows_max[0] - cur.seg * R_HI;
#pragma unroll
        for (int r = 0; r < R_HI; ++r) {
          if (r < nr) {
            const double rhs = raw_cur[r] * inv_scale_v - q0s;
            const double *wr = wslot + r * CHI * TILE;
#pragma unroll
            for (int c = 0; c < CHI; ++c) {
              const double wv = wr[c * TILE];
              if (c < CLO)
                lo[0][c] = fma(wv, rhs, lo[0][c]);
              else
                hi[c - CLO] = fma(wv, rhs, hi[c - CLO]);
            }
          }
        }
      } else {
        // two low-order stencils per segment; the segment number is resolved by explicit branches so that
        // the accumulators lo[k][c] are only ever indexed with compile-time k (they must stay in registers)
        auto lo_seg = [&](auto seg_tag) {
          constexpr int s = decltype(seg_tag)::value;
          const double *wk = wslot;
#pragma unroll
          for (int kk = 0; kk < 2; ++kk) {
            constexpr int k_first = 1 + 2 * s;
            if (k_first + kk < NS) {
              const int nr = sc.rows_max[k_first + kk];
#pragma unroll
              for (int r = 0; r < RLO; ++r) {
                if (r < nr) {
                  const double rhs = raw_cur[kk * RLO + r] * inv_scale_v - q0s;
#pragma unroll
                  for (int c = 0; c < CLO; ++c) {
                    const double wv = wk[(r * CLO + c) * TILE];
                    if (kk == 0)
                      lo[k_first < NS ? k_first : 0][c] = fma(wv, rhs, lo[k_first < NS ? k_first : 0][c]);
                    else
                      lo[k_first + 1 < NS ? k_first + 1 : 0][c] = fma(wv, rhs, lo[k_first + 1 < NS ? k_first + 1 : 0][c]);
                  }
                }
              }
              wk += nr * CLO * TILE;
            }
          }
        };
        const int ls = cur.seg - cfg.n_hi_segs;
        if (ls == 0) {
          lo_seg(std::integral_constant<int, 0>{});
        } else if (NLO_SEGS > 1 && ls == 1) {
          lo_seg(std::integral_constant<int, (NLO_SEGS > 1 ? 1 : 0)>{});
        } else if (NLO_SEGS > 2 && ls == 2) {
          lo_seg(std::integral_constant<int, (NLO_SEGS > 2 ? 2 : 0)>{});
        }
      }
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&w_empty[w.slot]);
      w.advance(WS);

      // ---- end of tile: hybridise and hand over ----------------------------------------------------
      if (cur.seg == n_segs - 1) {
        if (lane == 0) ptx::mbar_arrive(&hdr_empty[cur.h.slot]);  // ordered after the __syncwarp above
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
        double *isb = is_base + (m_tile & 1) * (NS * NVARS * TILE);
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
        ptx::named_bar_sync(1, N_APPLY);
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
        const std::int64_t tile = args.tile_list ? (std::int64_t)args.tile_list[cur.idx] : cur.idx;
        const bool active = tile * TILE + cell < P.n_cells;
        if (P.poly != nullptr && active) {
          for (int i = 0; i < D; ++i) P.poly[(cell_cur * P.n_poly_coef + i) * NVARS + var] = coef[i];
          P.poly_scale[cell_cur * NVARS + var] = scale_v;
        }
        // hand the polynomial (times the characteristic scale) to the trace group
        const int cb = m_tile & 1;
        ptx::mbar_wait(&coef_empty[cb], ((m_tile >> 1) & 1) ^ 1);
        double *cx = coef_base + cb * (D * NVARS * COEF_PAD) + var * COEF_PAD + cell;
#pragma unroll
        for (int i = 0; i < D; ++i) cx[i * NVARS * COEF_PAD] = coef[i] * scale_v;
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&coef_full[cb]);
        ++m_tile;
      }

#pragma unroll
      for (int r = 0; r < RAW; ++r) raw_cur[r] = raw_nxt[r];
      if (nxt.seg == 0) {
#pragma unroll
        for (int v = 0; v < NVARS; ++v) u0c[v] = u0n[v];
        cell_cur = cell_nxt;
      }
      cur = nxt;
      have_cur = have_nxt;
    }
    return;
  }

  // =============================== trace group: thread = (cell, face) ===============================
  {
    const int k = warp - (1 + STREAM_NVAR_WARPS);  // local face
    int m_tile = 0;
    for (std::int64_t idx = blockIdx.x; idx < n_launch; idx += gridDim.x, ++m_tile) {
      const std::int64_t tile = args.tile_list ? (std::int64_t)args.tile_list[idx] : idx;
      const std::int64_t cell = tile * TILE + lane;
      const bool active = cell < P.n_cells;
      // geometry of the cell (issued before the wait: the loads overlap the apply group's work)
      const std::uint32_t fref = active ? ld_stream(P.face_ref + (tile * F + k) * TILE + lane) : 0u;
      const std::uint32_t slots = P.face_slots[(tile * F + k) * TILE + lane];
      double fv[3][3];  // face vertices in the left cell's order
#pragma unroll
      for (int r = 0; r < ND; ++r) {
        const int s = (slots >> (2 * r)) & 3;
#pragma unroll
        for (int d = 0; d < 3; ++d)
          fv[r][d] = (ND == 2 && d == 2) ? 0.0 : P.vtx[((tile * F + s) * 3 + d) * TILE + lane];
      }
      double xc[3];
#pragma unroll
      for (int d = 0; d < 3; ++d) xc[d] = (ND == 2 && d == 2) ? 0.0 : P.center[(tile * 3 + d) * TILE + lane];
      const double inv_len = P.inv_len[tile * TILE + lane];
      double cmom[D];
#pragma unroll
      for (int i = 0; i < D; ++i) cmom[i] = 0.0;
#pragma unroll
      for (int i = 3; i < D; ++i) cmom[i] = P.moments[(tile * P.n_mom + (i - 3)) * TILE + lane];

      const int cb = m_tile & 1;
      ptx::mbar_wait(&coef_full[cb], (m_tile >> 1) & 1);
      const double *cx = coef_base + cb * (D * NVARS * COEF_PAD) + lane;
      constexpr bool IN_REGS = (D * NVARS <= 60);
      double coef[IN_REGS ? D : 1][NVARS];
      if (IN_REGS) {
#pragma unroll
        for (int i = 0; i < D; ++i)
#pragma unroll
          for (int v = 0; v < NVARS; ++v) coef[i][v] = cx[(i * NVARS + v) * COEF_PAD];
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&coef_empty[cb]);
      }
      const std::int64_t e = fref & FREF_EDGE_MASK;
      const int side = (fref & FREF_SIDE) ? 1 : 0;
      const bool want_trace = (fref & FREF_TRACE) != 0;
      if (want_trace) {
        for (int q = 0; q < sc.q_f; ++q) {
          double x[3];
#pragma unroll
          for (int d = 0; d < 3; ++d)
            x[d] = (ND == 2) ? sc.face_bary[q][0] * fv[0][d] + sc.face_bary[q][1] * fv[1][d]
                             : fv[0][d] * sc.face_bary[q][0] + fv[1][d] * sc.face_bary[q][1] +
                                   fv[2][d] * sc.face_bary[q][2];
          double mono[D];
          PolyEval<ND, DEG_HI>::monomials((x[0] - xc[0]) * inv_len, (x[1] - xc[1]) * inv_len,
                                          (ND == 3) ? (x[2] - xc[2]) * inv_len : 0.0, cmom, mono);
          double *tr = P.trace + ((e * 2 + side) * sc.q_f + q) * NVARS;
#pragma unroll
          for (int v = 0; v < NVARS; ++v) {
            double s = IN_REGS ? coef[0][v] : cx[v * COEF_PAD];
#pragma unroll
            for (int i = 1; i < D; ++i) s = fma(IN_REGS ? coef[i][v] : cx[(i * NVARS + v) * COEF_PAD], mono[i], s);
            tr[v] = s;
          }
        }
      }
      if (!IN_REGS) {
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&coef_empty[cb]);
      }
    }
  }
}

/// Shared-memory plan of the streaming kernel for one scheme; returns false if it does not apply.
template <int ND, int DEG_HI, int DEG_LO, int NS>
bool stream_config(const DevicePlan &P, const SchemeConst &sc, int smem_budget, StreamCfg &c) {
  constexpr int D = dof_of(DEG_HI, ND);
  constexpr int CHI = D - 1;
  constexpr int CLO = dof_of(DEG_LO, ND) - 1;
  constexpr int R_HI = stream_r_hi(CHI);
  constexpr int RLO = stream_rlo_max(ND);
  if (NS < 2 || CLO < 1) return false;
  for (int k = 1; k < NS; ++k)
    if (sc.rows_max[k] > RLO || sc.ncoef[k] != CLO) return false;
  if (sc.ncoef[0] != CHI) return false;
  c.n_hi_segs = (sc.rows_max[0] + R_HI - 1) / R_HI;
  c.n_lo_segs = NS / 2;
  c.hdr_bytes = P.hdr_bytes;
  int slot = R_HI * CHI * TILE * 8;
  if (c.n_hi_segs == 1) slot = sc.rows_max[0] * CHI * TILE * 8;
  for (int k0 = 1; k0 < NS; k0 += 2) {
    int b = sc.rows_max[k0] * CLO * TILE * 8;
    if (k0 + 1 < NS) b += sc.rows_max[k0 + 1] * CLO * TILE * 8;
    slot = b > slot ? b : slot;
  }
  c.slot_bytes = slot;
  c.n_hdr_slots = 3;
  const int coef_bytes = 2 * D * NVARS * COEF_PAD * 8;
  const int is_bytes = 2 * NS * NVARS * TILE * 8;
  const int bars_bytes = 1024;
  const int fixed = bars_bytes + c.n_hdr_slots * c.hdr_bytes + coef_bytes + is_bytes;
  int ws = (smem_budget - fixed) / slot;
  if (ws > 8) ws = 8;
  if (ws < 2) return false;
  c.n_w_slots = ws;
  c.off_bars = 0;
  c.off_hdr = bars_bytes;
  c.off_w = c.off_hdr + c.n_hdr_slots * c.hdr_bytes;
  c.off_coef = c.off_w + ws * slot;
  c.off_is = c.off_coef + coef_bytes;
  c.total_bytes = c.off_is + is_bytes;
  return 2 * c.n_hdr_slots + 2 * ws + 4 <= bars_bytes / 8;
}

}  // namespace zfvm
