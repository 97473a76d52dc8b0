// K2: face fluxes from the traces;  K3: atomic-free per-cell gather + source + fused RK stage
// update + frozen boundary condition (+ CFL / plausibility reductions on the last stage).
//
//   HLLCBatten::flux, hllc_speeds<IdealGasEOS>      flux/hllc.hpp:36-81,143-176
//   Euler::flux                                     model/euler_impl.hpp:23-36
//   coord_transform / inv_coord_transform           src/zisa/model/euler_variables.cpp:29-68
//   FluxLoop::compute_patch (quadrature + scatter)  fvm_loops/flux_loop.hpp:124-193
//   runge_kutta_sum                                 src/zisa/ode/runge_kutta.cpp:122-143
//   FrozenBC::apply                                 src/zisa/boundary/frozen_boundary_condition.cpp:39-55
//   LocalCFL                                        model/local_cfl_condition_impl.hpp:25-40
//   SanityCheckFor<Euler> / notplausible            model/euler_impl.hpp:44-46
#include <cstdlib>

#include "common.cuh"
#include "equilibrium.cuh"
#include "kernels.hpp"

// measured at 9.86 M tets (scratch/build_variants.sh, profiles/README.md round 2): the face frames read straight into
// registers instead of staged through shared memory: K2 1.886 -> 1.733 ms (22 instead of 17 warps per SM);
// update_kernel with 5 blocks per SM: no change, with 4: 0.60 -> 0.77 ms
#ifndef ZFVM_K2_FRAMES_DIRECT
#define ZFVM_K2_FRAMES_DIRECT 1
#endif
#ifndef ZFVM_K3_MIN_BLOCKS
#define ZFVM_K3_MIN_BLOCKS 6
#endif

namespace zfvm {

namespace {

// HLLCBatten::flux (flux/hllc.hpp:36-81,143-176).  Every quotient is a fast_rcp, every square root x * fast_rsqrt(x):
// 4 rsqrt + 6 rcp per Gauss point instead of 21 divisions and 4 square roots; results differ from the
// operation-by-operation form by rounding only.  The arguments of the roots are the reference's (rho_R / rho_L,
// gamma p / rho, the Roe sound speed squared) and min / max are the comparisons std::min / std::max make, so that a
// reconstruction that undershoots to a negative pressure or density at a Gauss point -- outside the scheme's domain, it
// happens next to strong discontinuities -- takes the same path as in the CPU code: NaN sound speed, comparisons false.
template <bool SPEEDS = false>
ZFVM_DEVICE void hllc_flux(const double uL[NVARS], const double uR[NVARS], double gamma, double nf[NVARS],
                           double *speeds = nullptr) {
  const double iL = fast_rcp(uL[0]), iR = fast_rcp(uR[0]);
  const double pL = (uL[4] - 0.5 * (uL[1] * uL[1] + uL[2] * uL[2] + uL[3] * uL[3]) * iL) * (gamma - 1.0);
  const double pR = (uR[4] - 0.5 * (uR[1] * uR[1] + uR[2] * uR[2] + uR[3] * uR[3]) * iR) * (gamma - 1.0);
  const double a2L = gamma * pL * iL, a2R = gamma * pR * iR;
  const double aL = a2L * fast_rsqrt(a2L), aR = a2R * fast_rsqrt(a2R);  // sqrt(gamma p / rho)

  const double rho_ratio = uR[0] * iL;
  const double roe_ratio = rho_ratio * fast_rsqrt(rho_ratio);  // sqrt(rho_R / rho_L)
  const double inv_den = fast_rcp(1.0 + roe_ratio);
  const double vL = uL[1] * iL, vR = uR[1] * iR;
  const double v_tilda = (vL + vR * roe_ratio) * inv_den;
  const double HL = (uL[4] + pL) * iL, HR = (uR[4] + pR) * iR;
  const double H_tilda = (HL + HR * roe_ratio) * inv_den;
  const double w2 = (uL[2] * iL + uR[2] * iR * roe_ratio) * inv_den;
  const double w3 = (uL[3] * iL + uR[3] * iR * roe_ratio) * inv_den;
  const double vroe_square = v_tilda * v_tilda + w2 * w2 + w3 * w3;
  const double a2 = (gamma - 1.0) * (H_tilda - 0.5 * vroe_square);
  const double a_tilda = a2 * fast_rsqrt(a2);

  // (std::min / std::max written out, see ref_min: a reconstructed pressure below zero makes a sound speed NaN)
  const double sL = ref_min(vL - aL, v_tilda - a_tilda);
  const double sR = ref_max(vR + aR, v_tilda + a_tilda);
  const double s_star =
      (uR[1] * (sR - vR) - uL[1] * (sL - vL) + pL - pR) * fast_rcp(uR[0] * (sR - vR) - uL[0] * (sL - vL));

  if constexpr (SPEEDS) {  // for HLLCBatten::tracer_flux (hllc.hpp:178-197)
    speeds[0] = sL;
    speeds[1] = s_star;
    speeds[2] = sR;
  }
  const bool left = (0.0 <= s_star);
  double uK[NVARS];
#pragma unroll
  for (int v = 0; v < NVARS; ++v) uK[v] = left ? uL[v] : uR[v];
  const double pK = left ? pL : pR;
  const double vK = left ? vL : vR;
  nf[0] = uK[1];  // Euler::flux, euler_impl.hpp:23-36
  nf[1] = vK * uK[1] + pK;
  nf[2] = vK * uK[2];
  nf[3] = vK * uK[3];
  nf[4] = vK * (uK[4] + pK);
  if (sL < 0.0 && 0.0 <= sR) {
    const double sK = left ? sL : sR;
    const double cK = (sK - vK) * fast_rcp(sK - s_star);
    nf[0] += sK * (cK * uK[0] - uK[0]);
    nf[1] += sK * (cK * uK[0] * s_star - uK[1]);
    nf[2] += sK * (cK * uK[2] - uK[2]);
    nf[3] += sK * (cK * uK[3] - uK[3]);
    nf[4] += sK * (cK * (uK[4] + (s_star - vK) * (uK[0] * s_star + pK * fast_rcp(sK - vK))) - uK[4]);
  }
}

// Not in the reference: local Lax-Friedrichs in the face frame.
ZFVM_DEVICE void rusanov_flux(const double uL[NVARS], const double uR[NVARS], double gamma, double nf[NVARS]) {
  const double iL = 1.0 / uL[0], iR = 1.0 / uR[0];
  const double pL = (uL[4] - 0.5 * (uL[1] * uL[1] + uL[2] * uL[2] + uL[3] * uL[3]) * iL) * (gamma - 1.0);
  const double pR = (uR[4] - 0.5 * (uR[1] * uR[1] + uR[2] * uR[2] + uR[3] * uR[3]) * iR) * (gamma - 1.0);
  const double aL = sqrt(gamma * pL * iL), aR = sqrt(gamma * pR * iR);
  const double vL = uL[1] * iL, vR = uR[1] * iR;
  double fL[NVARS], fR[NVARS];
  fL[0] = uL[1];
  fL[1] = vL * uL[1] + pL;
  fL[2] = vL * uL[2];
  fL[3] = vL * uL[3];
  fL[4] = vL * (uL[4] + pL);
  fR[0] = uR[1];
  fR[1] = vR * uR[1] + pR;
  fR[2] = vR * uR[2];
  fR[3] = vR * uR[3];
  fR[4] = vR * (uR[4] + pR);
  const double lam = ref_max(fabs(vL) + aL, fabs(vR) + aR);
#pragma unroll
  for (int v = 0; v < NVARS; ++v) nf[v] = 0.5 * (fL[v] + fR[v]) - 0.5 * lam * (uR[v] - uL[v]);
}

ZFVM_DEVICE void cp_async16(void *dst_smem, const void *src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((std::uint32_t)__cvta_generic_to_shared(dst_smem)), "l"(src)
               : "memory");
}

// K2.  One thread owns one (face, Gauss point) pair: LPF = QF rounded up to a power of two lanes per face,
// 32 / LPF faces per warp.  The Riemann solver is a long chain of dependent FP64 divisions and square roots;
// spreading a face's Gauss points over lanes shortens the chain per thread by QF and keeps the FP64 pipe busy.
// A face's two traces are one contiguous 80 QF-byte block and its frame one 80-byte row: the warp copies its
// faces' blocks into shared memory with 16-byte cp.async (fully coalesced when the faces are consecutive), the
// threads read their points from there, lane q = 0 of a face adds the weighted point fluxes in the reference's
// order (point 0 first, quadrature.hpp:43-48) and the fluxes go back through shared memory as contiguous rows.
template <int FLUX, int QF>
__global__ void __launch_bounds__(128) flux_kernel(const DevicePlan P, const __grid_constant__ SchemeConst sc,
                                                   const std::int32_t *__restrict__ face_list, std::int64_t n_faces,
                                                   std::int64_t face_begin) {
  constexpr int LPF = QF <= 1 ? 1 : (QF <= 2 ? 2 : (QF <= 4 ? 4 : 8));  // lanes per face
  constexpr int FPW = 32 / LPF;                                          // faces per warp
  constexpr int CHUNKS = QF * NVARS;                                     // 16-byte chunks of a trace block [2][QF][5]
  constexpr int PITCH = 2 * ((QF * NVARS + 1) | 1);                      // doubles per staged face row (odd count of 16-byte units)
  __shared__ __align__(16) double flux_smem[4 * FPW * (PITCH + 10)];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double *tr = flux_smem + warp * FPW * (PITCH + 10);
  double *frs = tr + FPW * PITCH;
  const std::int64_t t0 = ((std::int64_t)blockIdx.x * (blockDim.x >> 5) + warp) * FPW;
  if (t0 >= n_faces) return;
  const int f = lane / LPF, q = lane - f * LPF;  // face within the warp, Gauss point
  const std::int64_t t = min(t0 + f, n_faces - 1);
  const bool in_range = t0 + f < n_faces;
  const std::int64_t e = face_list ? (std::int64_t)face_list[t] : t + face_begin;

  // (uniform trip counts: every lane takes part in the shuffles)
#pragma unroll
  for (int c0 = 0; c0 < FPW * CHUNKS; c0 += 32) {
    const int c = min(c0 + lane, FPW * CHUNKS - 1);
    const int fc = c / CHUNKS, part = c - fc * CHUNKS;
    const std::int64_t ef = __shfl_sync(0xffffffffu, e, fc * LPF);
    if (c0 + lane < FPW * CHUNKS) cp_async16(tr + fc * PITCH + part * 2, P.trace + ef * (2 * CHUNKS) + part * 2);
  }
#pragma unroll
  for (int c0 = 0; c0 < FPW * 5; c0 += 32) {
    const int c = min(c0 + lane, FPW * 5 - 1);
    const int fc = c / 5, part = c - fc * 5;
    const std::int64_t ef = __shfl_sync(0xffffffffu, e, fc * LPF);
    if (c0 + lane < FPW * 5) cp_async16(frs + fc * 10 + part * 2, P.face_frame + ef * 10 + part * 2);
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
  const bool skip = !in_range || P.left_right[2 * e] < 0;
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncwarp();

  const double *fr = frs + f * 10;
  double n[3], t1[3], t2[3];
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    n[d] = fr[d];
    t1[d] = fr[3 + d];
    t2[d] = fr[6 + d];
  }
  double wf[NVARS] = {0.0, 0.0, 0.0, 0.0, 0.0};  // weighted flux of this Gauss point in the face frame
  if (q < QF && !skip) {
    double uL[NVARS], uR[NVARS];
#pragma unroll
    for (int v = 0; v < NVARS; ++v) {
      uL[v] = tr[f * PITCH + q * NVARS + v];
      uR[v] = tr[f * PITCH + (QF + q) * NVARS + v];
    }
    auto rot = [&](double u[NVARS]) {
      const double un = u[1] * n[0] + u[2] * n[1] + u[3] * n[2];
      const double ut1 = u[1] * t1[0] + u[2] * t1[1] + u[3] * t1[2];
      const double ut2 = u[1] * t2[0] + u[2] * t2[1] + u[3] * t2[2];
      u[1] = un;
      u[2] = ut1;
      u[3] = ut2;
    };
    rot(uL);
    rot(uR);
    double fl[NVARS];
    if (FLUX == FLUX_HLLC)
      hllc_flux(uL, uR, sc.gamma, fl);
    else
      rusanov_flux(uL, uR, sc.gamma, fl);
    const double wq = fr[9] * sc.face_w[q];
#pragma unroll
    for (int v = 0; v < NVARS; ++v) wf[v] = wq * fl[v];
  }
  // lane q = 0 accumulates point 0, 1, .. in order
  double nf[NVARS];
#pragma unroll
  for (int v = 0; v < NVARS; ++v) {
    nf[v] = wf[v];
#pragma unroll
    for (int qq = 1; qq < QF; ++qq) nf[v] += __shfl_sync(0xffffffffu, wf[v], (lane & ~(LPF - 1)) + qq);
  }
  __syncwarp();  // all lanes are done with the staged traces: reuse the area for the fluxes
  if (q == 0) {
    double *out = tr + f * NVARS;
    out[0] = nf[0];
    out[1] = nf[1] * n[0] + nf[2] * t1[0] + nf[3] * t2[0];
    out[2] = nf[1] * n[1] + nf[2] * t1[1] + nf[3] * t2[1];
    out[3] = nf[1] * n[2] + nf[2] * t1[2] + nf[3] * t2[2];
    out[4] = nf[4];
  }
  __syncwarp();
  for (int j = lane; j < ((FPW * NVARS + 31) / 32) * 32; j += 32) {
    const int fc = min(j / NVARS, FPW - 1);
    const std::int64_t ef = __shfl_sync(0xffffffffu, e, fc * LPF);
    const int skip_f = __shfl_sync(0xffffffffu, (int)skip, fc * LPF);
    if (j < FPW * NVARS && !skip_f) P.flux[ef * NVARS + (j - fc * NVARS)] = tr[j];
  }
}

// K2.  One warp owns 32 faces, one thread one face.  A face's two traces are one contiguous 80 q_f-byte block and
// its frame one 80-byte row, so a thread-per-face load touches 32 cache lines per instruction; instead the warp
// copies its faces' blocks into shared memory with 16-byte cp.async (fully coalesced when the faces are
// consecutive) and the threads read their rows from there (row pitch padded against bank conflicts).  The
// fluxes go back the same way.
// With TRACERS the kernel also upwinds the advected scalars (fvm_loops/flux_loop.hpp:157-161): the wave speeds are
// those of the face's HLLC evaluation on the same rotated traces (HLLCBatten::flux returns them, hllc.hpp:175), every
// scalar goes through HLLCBatten::tracer_flux (hllc.hpp:178-197) -- Rusanov (not in the reference): the same local
// Lax-Friedrichs form as the flux itself -- and the quadrature sums go to qflux[e][n_avars].
// With WBBG (well-balanced runs on the tile kernel) the staged traces are those of the perturbation: the equilibrium
// background (rho, E) of either side at the Gauss point (eq_bg, written by eq_face_kernel) is added first.
template <int FLUX, bool TRACERS, bool WBBG>
__global__ void __launch_bounds__(64) flux_face_kernel(const DevicePlan P, const __grid_constant__ SchemeConst sc,
                                                  const std::int32_t *__restrict__ face_list, std::int64_t n_faces,
                                                  int pitch, std::int64_t face_begin) {
  extern __shared__ __align__(16) double flux_smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpc = blockDim.x >> 5;
#if ZFVM_K2_FRAMES_DIRECT
  double *tr = flux_smem + (size_t)warp * (TILE * pitch);
#else
  double *tr = flux_smem + (size_t)warp * (TILE * pitch + TILE * 10);
  double *frs = tr + TILE * pitch;
#endif
  const std::int64_t t0 = ((std::int64_t)blockIdx.x * wpc + warp) * TILE;
  if (t0 >= n_faces) return;
  const std::int64_t t = min(t0 + lane, n_faces - 1);
  const bool in_range = t0 + lane < n_faces;
  const std::int64_t e = face_list ? (std::int64_t)face_list[t] : t + face_begin;

  const int chunks = sc.q_f * NVARS;  // 16-byte chunks of a face's trace block [2][q_f][5]
  for (int c = lane; c < TILE * chunks; c += TILE) {
    const int f = c / chunks, part = c - f * chunks;
    const std::int64_t ef = __shfl_sync(0xffffffffu, e, f);
    cp_async16(tr + f * pitch + part * 2, P.trace + ef * (2 * chunks) + part * 2);
  }
#if ZFVM_K2_FRAMES_DIRECT
  // the face frame (n, t1, t2, area: 80 bytes) straight into registers while the traces are in flight: 2.5 KB less shared
  // memory per warp, 22 instead of 17 warps per SM
  asm volatile("cp.async.commit_group;" ::: "memory");
  const double2 *frg = reinterpret_cast<const double2 *>(P.face_frame + e * 10);
  const double2 g0 = __ldg(frg), g1 = __ldg(frg + 1), g2 = __ldg(frg + 2), g3 = __ldg(frg + 3), g4 = __ldg(frg + 4);
  const bool skip = !in_range || P.left_right[2 * e] < 0;
  const double n[3] = {g0.x, g0.y, g1.x}, t1[3] = {g1.y, g2.x, g2.y}, t2[3] = {g3.x, g3.y, g4.x};
  const double area = g4.y;
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncwarp();
#else
  for (int c = lane; c < TILE * 5; c += TILE) {
    const int f = c / 5, part = c - f * 5;
    const std::int64_t ef = __shfl_sync(0xffffffffu, e, f);
    cp_async16(frs + f * 10 + part * 2, P.face_frame + ef * 10 + part * 2);
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
  const bool skip = !in_range || P.left_right[2 * e] < 0;
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncwarp();

  const double *fr = frs + lane * 10;
  double n[3], t1[3], t2[3];
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    n[d] = fr[d];
    t1[d] = fr[3 + d];
    t2[d] = fr[6 + d];
  }
  const double area = fr[9];
#endif
  double nf[NVARS] = {0.0, 0.0, 0.0, 0.0, 0.0};
  const double *trL = tr + lane * pitch;
  const double *trR = trL + sc.q_f * NVARS;
  const int NA = TRACERS ? P.n_avars : 0;
  double qnf[TRACERS ? MAX_AVARS : 1];
#pragma unroll
  for (int a = 0; a < (TRACERS ? MAX_AVARS : 1); ++a) qnf[a] = 0.0;
  const double *qL = TRACERS ? P.qtrace + e * (2 * sc.q_f * NA) : nullptr;
  const double *qR = TRACERS ? qL + sc.q_f * NA : nullptr;
  if (!skip) {
    for (int q = 0; q < sc.q_f; ++q) {
      double uL[NVARS], uR[NVARS];
#pragma unroll
      for (int v = 0; v < NVARS; ++v) {
        uL[v] = trL[q * NVARS + v];
        uR[v] = trR[q * NVARS + v];
      }
      if constexpr (WBBG) {
        const double2 bL = *reinterpret_cast<const double2 *>(P.eq_bg + ((e * 2 + 0) * sc.q_f + q) * 2);
        const double2 bR = *reinterpret_cast<const double2 *>(P.eq_bg + ((e * 2 + 1) * sc.q_f + q) * 2);
        uL[0] = bL.x + uL[0];
        uL[4] = bL.y + uL[4];
        uR[0] = bR.x + uR[0];
        uR[4] = bR.y + uR[4];
      }
      auto rot = [&](double u[NVARS]) {
        const double un = u[1] * n[0] + u[2] * n[1] + u[3] * n[2];
        const double ut1 = u[1] * t1[0] + u[2] * t1[1] + u[3] * t1[2];
        const double ut2 = u[1] * t2[0] + u[2] * t2[1] + u[3] * t2[2];
        u[1] = un;
        u[2] = ut1;
        u[3] = ut2;
      };
      rot(uL);
      rot(uR);
      double f[NVARS], sp[3] = {0.0, 0.0, 0.0};
      if (FLUX == FLUX_HLLC)
        hllc_flux<TRACERS>(uL, uR, sc.gamma, f, sp);
      else
        rusanov_flux(uL, uR, sc.gamma, f);
      const double wq = area * sc.face_w[q];
#pragma unroll
      for (int v = 0; v < NVARS; ++v) nf[v] += wq * f[v];
      if constexpr (TRACERS) {
        if (FLUX == FLUX_HLLC) {
          const double sL = sp[0], s_star = sp[1], sR = sp[2];
          const bool left = (0.0 <= s_star);
          const double vK = left ? uL[1] / uL[0] : uR[1] / uR[0];
          const bool fan = (sL < 0.0 && 0.0 < sR);
          const double sK = left ? sL : sR;
          const double cK = (sK - vK) / (sK - s_star);
#pragma unroll
          for (int a = 0; a < MAX_AVARS; ++a) {
            if (a < NA) {
              const double mqK = left ? qL[q * NA + a] : qR[q * NA + a];
              double fq = mqK * vK;
              if (fan) fq = fq + sK * (cK * mqK - mqK);
              qnf[a] += wq * fq;
            }
          }
        } else {
          const double iL = 1.0 / uL[0], iR = 1.0 / uR[0];
          const double pL = (uL[4] - 0.5 * (uL[1] * uL[1] + uL[2] * uL[2] + uL[3] * uL[3]) * iL) * (sc.gamma - 1.0);
          const double pR = (uR[4] - 0.5 * (uR[1] * uR[1] + uR[2] * uR[2] + uR[3] * uR[3]) * iR) * (sc.gamma - 1.0);
          const double aL = sqrt(sc.gamma * pL * iL), aR = sqrt(sc.gamma * pR * iR);
          const double vL = uL[1] * iL, vR = uR[1] * iR;
          const double lam = ref_max(fabs(vL) + aL, fabs(vR) + aR);
#pragma unroll
          for (int a = 0; a < MAX_AVARS; ++a) {
            if (a < NA) {
              const double mL = qL[q * NA + a], mR = qR[q * NA + a];
              qnf[a] += wq * (0.5 * (mL * vL + mR * vR) - 0.5 * lam * (mR - mL));
            }
          }
        }
      }
    }
    if constexpr (TRACERS) {
#pragma unroll
      for (int a = 0; a < MAX_AVARS; ++a)
        if (a < NA) P.qflux[e * NA + a] = qnf[a];
    }
  }
  const double fx = nf[1] * n[0] + nf[2] * t1[0] + nf[3] * t2[0];
  const double fy = nf[1] * n[1] + nf[2] * t1[1] + nf[3] * t2[1];
  const double fz = nf[1] * n[2] + nf[2] * t1[2] + nf[3] * t2[2];
  __syncwarp();  // all lanes are done with the staged traces: reuse the area for the fluxes
  double *out = tr + lane * NVARS;
  out[0] = nf[0];
  out[1] = fx;
  out[2] = fy;
  out[3] = fz;
  out[4] = nf[4];
  __syncwarp();
  for (int j = lane; j < TILE * NVARS; j += TILE) {
    const int f = j / NVARS;
    const std::int64_t ef = __shfl_sync(0xffffffffu, e, f);
    const int skip_f = __shfl_sync(0xffffffffu, (int)skip, f);
    if (!skip_f) P.flux[ef * NVARS + (j - f * NVARS)] = tr[j];
  }
}

// K3.  One thread owns one (cell, variable) pair and a block two tiles of 32 cells: every access to the row-major
// [n][5] arrays (state, tendencies, fluxes) is then contiguous across the warp, and a face's flux row is read by
// five adjacent lanes.  The face fluxes are gathered in the fixed order of the cell's face list (no atomics).
template <int F, bool FLUX_BC>
__global__ void __launch_bounds__(320, FLUX_BC ? 3 : ZFVM_K3_MIN_BLOCKS) update_kernel(const DevicePlan P, const UpdateArgs A,
                                                                      const __grid_constant__ SchemeConst sc) {
  constexpr int CELLS = 2 * TILE;
  __shared__ double s_un[CELLS * NVARS];
  const int cl = threadIdx.x / NVARS, v = threadIdx.x - cl * NVARS;  // cell within the block, variable
  const std::int64_t blk = (std::int64_t)blockIdx.x + A.block_begin;  // (a range of cell blocks: chunked host steps)
  const std::int64_t i = blk * CELLS + cl;
  const std::int64_t iv = i * NVARS + v;
  const bool active = i < A.n_cells_update;
  double un = 1.0, t = 0.0, ub = 0.0, kp[MAX_RK_STAGES - 1];
  if (active) {
    const std::int64_t tile = i / TILE;
    const int lane = (int)(i % TILE);
    // all independent loads first (face references, then the four flux rows, the cell's own rows): the face
    // gather is a two-level dependent chain and the kernel is latency bound otherwise
    std::uint32_t fref[F];
#pragma unroll
    for (int k = 0; k < F; ++k) fref[k] = P.face_ref[(tile * F + k) * TILE + lane];
    const double vol = P.volume[i];
    if (A.u_next) ub = A.u_base[iv];
#pragma unroll
    for (int s = 0; s < MAX_RK_STAGES - 1; ++s)
      kp[s] = (A.u_next && s < A.n_prev && A.coef_prev[s] != 0.0) ? A.k_prev[s][iv] : 0.0;
    double fl[F];
#pragma unroll
    for (int k = 0; k < F; ++k) {
      const bool use = (fref[k] & FREF_TRACE) != 0;
      const std::int64_t e = use ? (std::int64_t)(fref[k] & FREF_EDGE_MASK) : 0;
      fl[k] = P.flux[e * NVARS + v];
      if (!use) fl[k] = 0.0;
    }
    const double inv_vol = 1.0 / vol;
#pragma unroll
    for (int k = 0; k < F; ++k) {
      if (fref[k] & FREF_TRACE) t += ((fref[k] & FREF_SIDE) ? fl[k] : -fl[k]) * inv_vol;
    }
    if constexpr (FLUX_BC) {
      // FluxBC::compute (boundary/flux_bc.hpp:24-42): the cell's exterior faces, in face-list order; the cell is the
      // left cell of an exterior face.  Rare (boundary cells only): every thread of the cell evaluates the 5-vector.
#pragma unroll
      for (int k = 0; k < F; ++k) {
        if (!(fref[k] & FREF_INTERIOR)) {
          const std::int64_t e = fref[k] & FREF_EDGE_MASK;
          const double *fr = P.face_frame + e * 10;
          double u[NVARS];
#pragma unroll
          for (int w = 0; w < NVARS; ++w) u[w] = A.flux_bc_state[i * NVARS + w];
          if (A.flux_bc_kind == 2) {
            // EquilibriumFluxBC::compute (boundary/equilibrium_flux_bc.hpp:37-63): the cell's local equilibrium is solved
            // from its average (rho, E_int); the flux of the resting equilibrium state at the face Gauss points is its
            // pressure along the face normal (Euler::flux of (rho, 0, 0, 0, E), rotated back with inv_coord_transform)
            if (v >= 1 && v <= 3) {
              const double eint = u[4] - 0.5 * (u[1] * u[1] + u[2] * u[2] + u[3] * u[3]) / u[0];
              const LocalEq eq = solve_local_equilibrium(u[0], eint, P.phi_cqp + i * sc.q_c, sc);
              double acc = 0.0;
              for (int q = 0; q < sc.q_f; ++q) {
                double r_, E_, p_;
                eq.at(P.phi_fqp[e * sc.q_f + q], sc, r_, E_, p_);
                const double wq = fr[9] * sc.face_w[q];
                acc = (q == 0) ? wq * (p_ * fr[v - 1]) : acc + wq * (p_ * fr[v - 1]);
              }
              t -= acc / vol;
            }
            continue;
          }
          const double un_ = u[1] * fr[0] + u[2] * fr[1] + u[3] * fr[2];
          const double ut1 = u[1] * fr[3] + u[2] * fr[4] + u[3] * fr[5];
          const double ut2 = u[1] * fr[6] + u[2] * fr[7] + u[3] * fr[8];
          const double p = (u[4] - 0.5 * (u[1] * u[1] + u[2] * u[2] + u[3] * u[3]) / u[0]) * (A.gamma - 1.0);
          const double vn = un_ / u[0];
          const double f1 = vn * un_ + p, f2 = vn * ut1, f3 = vn * ut2;
          double fv = un_;  // mass flux; the momentum components are rotated back (inv_coord_transform)
          if (v == 1) fv = f1 * fr[0] + f2 * fr[3] + f3 * fr[6];
          if (v == 2) fv = f1 * fr[1] + f2 * fr[4] + f3 * fr[7];
          if (v == 3) fv = f1 * fr[2] + f2 * fr[5] + f3 * fr[8];
          if (v == 4) fv = vn * (u[4] + p);
          t -= fr[9] / vol * fv;
        }
      }
    }
    if (A.has_source) t += P.source[iv];
    if (A.tendency) {
      if (A.accumulate)
        A.tendency[iv] += t;
      else
        A.tendency[iv] = t;
    }
  }
  // the fused stage update may overwrite the very rows FluxBC has just read (u_next aliases the stage's input state
  // from the second stage on): all reads of the block's cells are done before any of them is updated
  if constexpr (FLUX_BC) __syncthreads();
  if (active && A.u_next) {
    // runge_kutta_sum: stages in index order, the stage just computed is the last one
    double dudt = 0.0;
#pragma unroll
    for (int s = 0; s < MAX_RK_STAGES - 1; ++s)
      if (s < A.n_prev && A.coef_prev[s] != 0.0) dudt += A.coef_prev[s] * kp[s];
    if (A.coef_cur != 0.0) dudt += A.coef_cur * t;
    un = ub + (A.dt_dev ? *A.dt_dev : A.dt) * dudt;
    if (A.frozen && (P.cell_flags[i] & 2)) un = A.frozen[iv];
    A.u_next[iv] = un;
  }
  if (A.reduce_out) {  // LocalCFL + plausibility over the updated state (uniform branch)
    s_un[threadIdx.x] = un;
    __syncthreads();
    if (threadIdx.x < CELLS) {
      const std::int64_t ic = blk * CELLS + threadIdx.x;
      double dx_over_ev = 1e300;
      int bad = 0;
      if (ic < A.n_cells_update && A.u_next) {
        double u[NVARS];
#pragma unroll
        for (int w = 0; w < NVARS; ++w) u[w] = s_un[threadIdx.x * NVARS + w];
        const double p = pressure_of(u, A.gamma);
        const double a = sqrt(A.gamma * p / u[0]);
        const double v2 = (u[1] * u[1] + u[2] * u[2] + u[3] * u[3]) / (u[0] * u[0]);
        dx_over_ev = A.inradius[ic] / (sqrt(v2) + a);
        bool finite = true;
#pragma unroll
        for (int w = 0; w < NVARS; ++w) finite = finite && isfinite(u[w]);
        bad = (u[0] <= 0.0 || u[4] <= 0.0 || !finite) ? 1 : 0;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        dx_over_ev = fmin(dx_over_ev, __shfl_xor_sync(0xffffffffu, dx_over_ev, o));
        bad |= __shfl_xor_sync(0xffffffffu, bad, o);
      }
      if ((threadIdx.x & 31) == 0) {
        // positive doubles order like their bit patterns; NaN (from a bad state) is caught by `bad`
        if (dx_over_ev == dx_over_ev)
          atomicMin(reinterpret_cast<unsigned long long *>(&A.reduce_out->min_dx_over_ev),
                    (unsigned long long)__double_as_longlong(fmax(dx_over_ev, 0.0)));
        if (bad) atomicOr(&A.reduce_out->not_plausible, 1);
      }
    }
  }
}

__global__ void cfl_kernel(const double *__restrict__ u, const double *__restrict__ inradius, std::int64_t n,
                           double gamma, ReduceOut *out) {
  const std::int64_t i = (std::int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  double r = 1e300;
  int bad = 0;
  if (i < n) {
    double un[NVARS];
#pragma unroll
    for (int v = 0; v < NVARS; ++v) un[v] = u[i * NVARS + v];
    const double p = pressure_of(un, gamma);
    const double a = sqrt(gamma * p / un[0]);
    const double v2 = (un[1] * un[1] + un[2] * un[2] + un[3] * un[3]) / (un[0] * un[0]);
    r = inradius[i] / (sqrt(v2) + a);
    bool finite = true;
#pragma unroll
    for (int v = 0; v < NVARS; ++v) finite = finite && isfinite(un[v]);
    bad = (un[0] <= 0.0 || un[4] <= 0.0 || !finite) ? 1 : 0;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    r = fmin(r, __shfl_xor_sync(0xffffffffu, r, o));
    bad |= __shfl_xor_sync(0xffffffffu, bad, o);
  }
  if ((threadIdx.x & 31) == 0) {
    if (r == r)
      atomicMin(reinterpret_cast<unsigned long long *>(&out->min_dx_over_ev),
                (unsigned long long)__double_as_longlong(fmax(r, 0.0)));
    if (bad) atomicOr(&out->not_plausible, 1);
  }
}

__global__ void reset_reduce_kernel(ReduceOut *out) {
  out->min_dx_over_ev = 1e300;
  out->not_plausible = 0;
}

__global__ void frozen_bc_kernel(double *__restrict__ u, const double *__restrict__ frozen,
                                 const std::int32_t *__restrict__ ghost_index, std::int64_t n_ghost) {
  const std::int64_t t = (std::int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_ghost * NVARS) return;
  const std::int64_t i = ghost_index[t / NVARS];
  u[i * NVARS + t % NVARS] = frozen[i * NVARS + t % NVARS];
}

// pack rows state[index[r]] into a contiguous send buffer (HaloSendPart, mpi_halo_exchange.cpp:132-142)
__global__ void pack_rows_kernel(double *__restrict__ out, const double *__restrict__ state,
                                 const std::int32_t *__restrict__ index, std::int64_t n_rows) {
  const std::int64_t t = (std::int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_rows * NVARS) return;
  out[t] = state[(std::int64_t)index[t / NVARS] * NVARS + t % NVARS];
}

__global__ void axpy_stage_kernel(double *__restrict__ u_next, const double *__restrict__ u_base,
                                  const StagePtrs K, int n_stages, double dt, std::int64_t n) {
  const std::int64_t t = (std::int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  double dudt = 0.0;
  for (int s = 0; s < n_stages; ++s)
    if (K.coef[s] != 0.0) dudt += K.coef[s] * K.k[s][t];
  u_next[t] = u_base[t] + dt * dudt;
}

// T3.  One thread per (cell, scalar): atomic-free gather of the cell's tracer face fluxes (flux_loop.hpp:180-192),
// fused Runge-Kutta sum and FrozenBC on the avars rows -- the avars half of K3 (no sources act on avars).
template <int F>
__global__ void __launch_bounds__(256) tracer_update_kernel(const DevicePlan P, const UpdateArgs A) {
  const int NA = A.n_avars;
  const std::int64_t idx = (std::int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= A.n_cells_update * NA) return;
  const std::int64_t i = idx / NA;
  const int a = (int)(idx - i * NA);
  const std::int64_t tile = i / TILE;
  const int lane = (int)(i % TILE);
  const double inv_vol = 1.0 / P.volume[i];
  double t = 0.0;
#pragma unroll
  for (int k = 0; k < F; ++k) {
    const std::uint32_t fref = P.face_ref[(tile * F + k) * TILE + lane];
    if (fref & FREF_TRACE) {
      const double fl = P.qflux[(std::int64_t)(fref & FREF_EDGE_MASK) * NA + a];
      t += ((fref & FREF_SIDE) ? fl : -fl) * inv_vol;
    }
  }
  if (A.tendency) {
    if (A.accumulate)
      A.tendency[idx] += t;
    else
      A.tendency[idx] = t;
  }
  if (A.u_next) {
    double dudt = 0.0;
    for (int s = 0; s < A.n_prev; ++s)
      if (A.coef_prev[s] != 0.0) dudt += A.coef_prev[s] * A.k_prev[s][idx];
    if (A.coef_cur != 0.0) dudt += A.coef_cur * t;
    double un = A.u_base[idx] + (A.dt_dev ? *A.dt_dev : A.dt) * dudt;
    if (A.frozen && (P.cell_flags[i] & 2)) un = A.frozen[idx];
    A.u_next[idx] = un;
  }
}

__global__ void frozen_bc_n_kernel(double *__restrict__ u, const double *__restrict__ frozen,
                                   const std::int32_t *__restrict__ ghost_index, std::int64_t n_ghost, int row_len) {
  const std::int64_t t = (std::int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_ghost * row_len) return;
  const std::int64_t i = ghost_index[t / row_len];
  u[i * row_len + t % row_len] = frozen[i * row_len + t % row_len];
}

__global__ void pack_rows_n_kernel(double *__restrict__ out, const double *__restrict__ state,
                                   const std::int32_t *__restrict__ index, std::int64_t n_rows, int row_len) {
  const std::int64_t t = (std::int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_rows * row_len) return;
  out[t] = state[(std::int64_t)index[t / row_len] * row_len + t % row_len];
}

}  // namespace

void launch_tracer_update(const DevicePlan &P, int n_dims, const UpdateArgs &A, cudaStream_t stream) {
  if (A.n_cells_update <= 0 || A.n_avars <= 0) return;
  const int block = 256;
  const unsigned grid = (unsigned)((A.n_cells_update * A.n_avars + block - 1) / block);
  if (n_dims == 2)
    tracer_update_kernel<3><<<grid, block, 0, stream>>>(P, A);
  else
    tracer_update_kernel<4><<<grid, block, 0, stream>>>(P, A);
}

void launch_frozen_bc_n(double *u, const double *frozen, const std::int32_t *ghost_index, std::int64_t n_ghost,
                        int row_len, cudaStream_t stream) {
  if (n_ghost <= 0 || row_len <= 0) return;
  const int block = 256;
  frozen_bc_n_kernel<<<(unsigned)((n_ghost * row_len + block - 1) / block), block, 0, stream>>>(u, frozen, ghost_index,
                                                                                                 n_ghost, row_len);
}

void launch_pack_rows_n(double *out, const double *state, const std::int32_t *index, std::int64_t n_rows, int row_len,
                        cudaStream_t stream) {
  if (n_rows <= 0 || row_len <= 0) return;
  const int block = 256;
  pack_rows_n_kernel<<<(unsigned)((n_rows * row_len + block - 1) / block), block, 0, stream>>>(out, state, index, n_rows,
                                                                                               row_len);
}

template <int FLUX>
static void launch_flux_q(const DevicePlan &P, const SchemeConst &sc, const std::int32_t *face_list, std::int64_t n_faces,
                          std::int64_t face_begin, cudaStream_t stream) {
  static const bool per_point = [] {
    const char *e = std::getenv("ZFVM_FLUX");
    return e != nullptr && e[0] == 'p';
  }();
  if (!per_point || P.n_avars > 0 || P.eq_bg != nullptr) {  // one thread per face (the only variant with scalars / WB background)
    const int block = 64, wpc = block / 32;
    // doubles per staged face row: even (16-byte cp.async) with an odd number of 16-byte units, so that the 64-bit
    // reads of 32 consecutive rows spread over all banks (q_f = 3 would otherwise give a pitch of 32 doubles)
    const int pitch = 2 * ((sc.q_f * NVARS + 1) | 1);
#if ZFVM_K2_FRAMES_DIRECT
    const size_t smem = (size_t)wpc * (TILE * pitch) * sizeof(double);
#else
    const size_t smem = (size_t)wpc * (TILE * pitch + TILE * 10) * sizeof(double);
#endif
    const unsigned grid = (unsigned)((n_faces + block - 1) / block);
    const bool bg = P.eq_bg != nullptr;
    if (P.n_avars > 0) {
      if (bg)
        flux_face_kernel<FLUX, true, true><<<grid, block, smem, stream>>>(P, sc, face_list, n_faces, pitch, face_begin);
      else
        flux_face_kernel<FLUX, true, false><<<grid, block, smem, stream>>>(P, sc, face_list, n_faces, pitch, face_begin);
    } else {
      if (bg)
        flux_face_kernel<FLUX, false, true><<<grid, block, smem, stream>>>(P, sc, face_list, n_faces, pitch, face_begin);
      else
        flux_face_kernel<FLUX, false, false><<<grid, block, smem, stream>>>(P, sc, face_list, n_faces, pitch, face_begin);
    }
    return;
  }
  auto go = [&](auto kern, int lanes_per_face) {
    const int faces_per_block = 4 * (32 / lanes_per_face);
    const unsigned grid = (unsigned)((n_faces + faces_per_block - 1) / faces_per_block);
    kern<<<grid, 128, 0, stream>>>(P, sc, face_list, n_faces, face_begin);
  };
  switch (sc.q_f) {  // edge rules: 1-3 points, triangle rules: 1, 3, 4, 6, 7 points
    case 1: go(flux_kernel<FLUX, 1>, 1); break;
    case 2: go(flux_kernel<FLUX, 2>, 2); break;
    case 3: go(flux_kernel<FLUX, 3>, 4); break;
    case 4: go(flux_kernel<FLUX, 4>, 4); break;
    case 5: go(flux_kernel<FLUX, 5>, 8); break;
    case 6: go(flux_kernel<FLUX, 6>, 8); break;
    case 7: go(flux_kernel<FLUX, 7>, 8); break;
    default: go(flux_kernel<FLUX, 8>, 8); break;
  }
}

void launch_flux(const DevicePlan &P, const SchemeConst &sc, const std::int32_t *face_list, std::int64_t n_faces,
                 cudaStream_t stream, std::int64_t face_begin) {
  if (n_faces <= 0) return;
  if (sc.flux == FLUX_HLLC)
    launch_flux_q<FLUX_HLLC>(P, sc, face_list, n_faces, face_begin, stream);
  else
    launch_flux_q<FLUX_RUSANOV>(P, sc, face_list, n_faces, face_begin, stream);
}

void launch_update(const DevicePlan &P, const SchemeConst &sc, const UpdateArgs &A, cudaStream_t stream) {
  const int n_dims = sc.n_dims;
  if (A.n_cells_update <= 0) return;
  const int block = 2 * TILE * NVARS;
  const std::int64_t n_blocks = (A.n_cells_update + 2 * TILE - 1) / (2 * TILE) - A.block_begin;
  if (n_blocks <= 0) return;
  const unsigned grid = (unsigned)n_blocks;
  const bool bc = A.flux_bc_state != nullptr;
  if (n_dims == 2) {
    if (bc)
      update_kernel<3, true><<<grid, block, 0, stream>>>(P, A, sc);
    else
      update_kernel<3, false><<<grid, block, 0, stream>>>(P, A, sc);
  } else {
    if (bc)
      update_kernel<4, true><<<grid, block, 0, stream>>>(P, A, sc);
    else
      update_kernel<4, false><<<grid, block, 0, stream>>>(P, A, sc);
  }
}

void launch_cfl(const double *u, const double *inradius, std::int64_t n, double gamma, ReduceOut *out,
                cudaStream_t stream) {
  const int block = 256;
  cfl_kernel<<<(unsigned)((n + block - 1) / block), block, 0, stream>>>(u, inradius, n, gamma, out);
}

void launch_reset_reduce(ReduceOut *out, cudaStream_t stream) { reset_reduce_kernel<<<1, 1, 0, stream>>>(out); }

void launch_frozen_bc(double *u, const double *frozen, const std::int32_t *ghost_index, std::int64_t n_ghost,
                      cudaStream_t stream) {
  if (n_ghost <= 0) return;
  const int block = 256;
  frozen_bc_kernel<<<(unsigned)((n_ghost * NVARS + block - 1) / block), block, 0, stream>>>(u, frozen, ghost_index,
                                                                                             n_ghost);
}

void launch_pack_rows(double *out, const double *state, const std::int32_t *index, std::int64_t n_rows,
                      cudaStream_t stream) {
  if (n_rows <= 0) return;
  const int block = 256;
  pack_rows_kernel<<<(unsigned)((n_rows * NVARS + block - 1) / block), block, 0, stream>>>(out, state, index,
                                                                                           n_rows);
}

void launch_axpy_stage(double *u_next, const double *u_base, const StagePtrs &K, int n_stages, double dt,
                       std::int64_t n, cudaStream_t stream) {
  if (n <= 0) return;
  const int block = 256;
  axpy_stage_kernel<<<(unsigned)((n + block - 1) / block), block, 0, stream>>>(u_next, u_base, K, n_stages, dt, n);
}

}  // namespace zfvm
