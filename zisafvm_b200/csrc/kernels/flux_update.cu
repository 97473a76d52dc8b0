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
#include "common.cuh"
#include "kernels.hpp"

namespace zfvm {

namespace {

// HLLCBatten::flux (flux/hllc.hpp:36-81,143-176) with the reciprocals 1/rho_L, 1/rho_R and 1/(1 + sqrt(rho_R/rho_L))
// formed once: 6 divisions and 4 square roots per Gauss point instead of 21 and 4 (FP64 divisions are what made
// the first version of this kernel FP64-pipe bound); results differ from the operation-by-operation form by rounding only.
ZFVM_DEVICE void hllc_flux(const double uL[NVARS], const double uR[NVARS], double gamma, double nf[NVARS]) {
  const double iL = 1.0 / uL[0], iR = 1.0 / uR[0];
  const double pL = (uL[4] - 0.5 * (uL[1] * uL[1] + uL[2] * uL[2] + uL[3] * uL[3]) * iL) * (gamma - 1.0);
  const double pR = (uR[4] - 0.5 * (uR[1] * uR[1] + uR[2] * uR[2] + uR[3] * uR[3]) * iR) * (gamma - 1.0);
  const double aL = sqrt(gamma * pL * iL), aR = sqrt(gamma * pR * iR);

  const double roe_ratio = sqrt(uR[0] * iL);
  const double inv_den = 1.0 / (1.0 + roe_ratio);
  const double vL = uL[1] * iL, vR = uR[1] * iR;
  const double v_tilda = (vL + vR * roe_ratio) * inv_den;
  const double HL = (uL[4] + pL) * iL, HR = (uR[4] + pR) * iR;
  const double H_tilda = (HL + HR * roe_ratio) * inv_den;
  const double w2 = (uL[2] * iL + uR[2] * iR * roe_ratio) * inv_den;
  const double w3 = (uL[3] * iL + uR[3] * iR * roe_ratio) * inv_den;
  const double vroe_square = v_tilda * v_tilda + w2 * w2 + w3 * w3;
  const double a_tilda = sqrt((gamma - 1.0) * (H_tilda - 0.5 * vroe_square));

  const double sL = fmin(vL - aL, v_tilda - a_tilda);
  const double sR = fmax(vR + aR, v_tilda + a_tilda);
  const double s_star =
      (uR[1] * (sR - vR) - uL[1] * (sL - vL) + pL - pR) / (uR[0] * (sR - vR) - uL[0] * (sL - vL));

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
    const double cK = (sK - vK) / (sK - s_star);
    nf[0] += sK * (cK * uK[0] - uK[0]);
    nf[1] += sK * (cK * uK[0] * s_star - uK[1]);
    nf[2] += sK * (cK * uK[2] - uK[2]);
    nf[3] += sK * (cK * uK[3] - uK[3]);
    nf[4] += sK * (cK * (uK[4] + (s_star - vK) * (uK[0] * s_star + pK / (sK - vK))) - uK[4]);
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
  const double lam = fmax(fabs(vL) + aL, fabs(vR) + aR);
#pragma unroll
  for (int v = 0; v < NVARS; ++v) nf[v] = 0.5 * (fL[v] + fR[v]) - 0.5 * lam * (uR[v] - uL[v]);
}

ZFVM_DEVICE void cp_async16(void *dst_smem, const void *src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((std::uint32_t)__cvta_generic_to_shared(dst_smem)), "l"(src)
               : "memory");
}

// K2.  One warp owns 32 faces, one thread one face.  A face's two traces are one contiguous 80 q_f-byte block and
// its frame one 80-byte row, so a thread-per-face load touches 32 cache lines per instruction; instead the warp
// copies its faces' blocks into shared memory with 16-byte cp.async (fully coalesced when the faces are
// consecutive) and the threads read their rows from there (row pitch padded against bank conflicts).  The
// fluxes go back the same way.
template <int FLUX>
__global__ void __launch_bounds__(64) flux_kernel(const DevicePlan P, const __grid_constant__ SchemeConst sc,
                                                  const std::int32_t *__restrict__ face_list, std::int64_t n_faces,
                                                  int pitch) {
  extern __shared__ __align__(16) double flux_smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpc = blockDim.x >> 5;
  double *tr = flux_smem + (size_t)warp * (TILE * pitch + TILE * 10);
  double *frs = tr + TILE * pitch;
  const std::int64_t t0 = ((std::int64_t)blockIdx.x * wpc + warp) * TILE;
  if (t0 >= n_faces) return;
  const std::int64_t t = min(t0 + lane, n_faces - 1);
  const bool in_range = t0 + lane < n_faces;
  const std::int64_t e = face_list ? (std::int64_t)face_list[t] : t;

  const int chunks = sc.q_f * NVARS;  // 16-byte chunks of a face's trace block [2][q_f][5]
  for (int c = lane; c < TILE * chunks; c += TILE) {
    const int f = c / chunks, part = c - f * chunks;
    const std::int64_t ef = __shfl_sync(0xffffffffu, e, f);
    cp_async16(tr + f * pitch + part * 2, P.trace + ef * (2 * chunks) + part * 2);
  }
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
  double nf[NVARS] = {0.0, 0.0, 0.0, 0.0, 0.0};
  const double *trL = tr + lane * pitch;
  const double *trR = trL + sc.q_f * NVARS;
  if (!skip) {
    for (int q = 0; q < sc.q_f; ++q) {
      double uL[NVARS], uR[NVARS];
#pragma unroll
      for (int v = 0; v < NVARS; ++v) {
        uL[v] = trL[q * NVARS + v];
        uR[v] = trR[q * NVARS + v];
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
      double f[NVARS];
      if (FLUX == FLUX_HLLC)
        hllc_flux(uL, uR, sc.gamma, f);
      else
        rusanov_flux(uL, uR, sc.gamma, f);
      const double wq = area * sc.face_w[q];
#pragma unroll
      for (int v = 0; v < NVARS; ++v) nf[v] += wq * f[v];
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
template <int F>
__global__ void __launch_bounds__(320) update_kernel(const DevicePlan P, const UpdateArgs A) {
  constexpr int CELLS = 2 * TILE;
  __shared__ double s_un[CELLS * NVARS];
  const int cl = threadIdx.x / NVARS, v = threadIdx.x - cl * NVARS;  // cell within the block, variable
  const std::int64_t i = (std::int64_t)blockIdx.x * CELLS + cl;
  const bool active = i < A.n_cells_update;
  double un = 1.0;
  if (active) {
    const std::int64_t tile = i / TILE;
    const int lane = (int)(i % TILE);
    const double inv_vol = 1.0 / P.volume[i];
    double t = 0.0;
#pragma unroll
    for (int k = 0; k < F; ++k) {
      const std::uint32_t fref = P.face_ref[(tile * F + k) * TILE + lane];
      if (!(fref & FREF_TRACE)) continue;
      const std::int64_t e = fref & FREF_EDGE_MASK;
      const double fl = P.flux[e * NVARS + v];
      t += ((fref & FREF_SIDE) ? fl : -fl) * inv_vol;
    }
    const std::int64_t iv = i * NVARS + v;
    if (A.has_source) t += P.source[iv];
    if (A.tendency) {
      if (A.accumulate)
        A.tendency[iv] += t;
      else
        A.tendency[iv] = t;
    }
    if (A.u_next) {
      // runge_kutta_sum: stages in index order, the stage just computed is the last one
      double dudt = 0.0;
      for (int s = 0; s < A.n_prev; ++s)
        if (A.coef_prev[s] != 0.0) dudt += A.coef_prev[s] * A.k_prev[s][iv];
      if (A.coef_cur != 0.0) dudt += A.coef_cur * t;
      un = A.u_base[iv] + A.dt * dudt;
      if (A.frozen && (P.cell_flags[i] & 2)) un = A.frozen[iv];
      A.u_next[iv] = un;
    }
  }
  if (A.reduce_out) {  // LocalCFL + plausibility over the updated state (uniform branch)
    s_un[threadIdx.x] = un;
    __syncthreads();
    if (threadIdx.x < CELLS) {
      const std::int64_t ic = (std::int64_t)blockIdx.x * CELLS + threadIdx.x;
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

}  // namespace

void launch_flux(const DevicePlan &P, const SchemeConst &sc, const std::int32_t *face_list, std::int64_t n_faces,
                 cudaStream_t stream) {
  if (n_faces <= 0) return;
  const int block = 64, wpc = block / 32;
  const int pitch = 2 * sc.q_f * NVARS + 2;  // doubles per staged face row: 16-byte aligned, 2-way bank conflicts at most
  const size_t smem = (size_t)wpc * (TILE * pitch + TILE * 10) * sizeof(double);
  const unsigned grid = (unsigned)((n_faces + block - 1) / block);
  if (sc.flux == FLUX_HLLC)
    flux_kernel<FLUX_HLLC><<<grid, block, smem, stream>>>(P, sc, face_list, n_faces, pitch);
  else
    flux_kernel<FLUX_RUSANOV><<<grid, block, smem, stream>>>(P, sc, face_list, n_faces, pitch);
}

void launch_update(const DevicePlan &P, int n_dims, const UpdateArgs &A, cudaStream_t stream) {
  if (A.n_cells_update <= 0) return;
  const int block = 2 * TILE * NVARS;
  const unsigned grid = (unsigned)((A.n_cells_update + 2 * TILE - 1) / (2 * TILE));
  if (n_dims == 2)
    update_kernel<3><<<grid, block, 0, stream>>>(P, A);
  else
    update_kernel<4><<<grid, block, 0, stream>>>(P, A);
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
