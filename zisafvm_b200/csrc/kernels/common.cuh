// Device helpers shared by the residual kernels.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "../device/layout.hpp"

namespace zfvm {

#define ZFVM_DEVICE __device__ __forceinline__

constexpr __host__ __device__ int dof_of(int deg, int nd) {
  return nd == 2 ? ((deg + 1) * (deg + 2)) / 2 : ((deg + 1) * (deg + 2) * (deg + 3)) / 6;
}

struct Expo {
  int a, b, c;
};

/// Inverse of poly_index (poly2d_impl.hpp:34-41): linear coefficient index -> exponents.
template <int ND>
constexpr __host__ __device__ Expo expo_of(int i) {
  if (ND == 2) {
    int n = 0;
    while (dof_of(n, 2) <= i) ++n;       // total degree
    int b = i - dof_of(n - 1, 2);
    return Expo{n - b, b, 0};
  } else {
    int n = 0;
    while (dof_of(n, 3) <= i) ++n;
    int j = i - dof_of(n - 1, 3);          // poly_index(b, c) within the degree-n block
    int m = 0;
    while (dof_of(m, 2) <= j) ++m;        // m = b + c
    int c = j - dof_of(m - 1, 2);
    return Expo{n - m, m - c, c};
  }
}

/// Compile-time table of exponents for all coefficient indices of a degree-DEG polynomial.
template <int ND, int DEG>
struct ExpoTable {
  static constexpr int D = dof_of(DEG, ND);
  Expo e[D];
  constexpr ExpoTable() : e{} {
    for (int i = 0; i < D; ++i) e[i] = expo_of<ND>(i);
  }
};

// streaming loads: data touched exactly once per stage must not evict the gathered state from L1/L2
ZFVM_DEVICE double ld_stream(const double *p) { return __ldcs(p); }
ZFVM_DEVICE std::int32_t ld_stream(const std::int32_t *p) { return __ldcs(p); }
ZFVM_DEVICE std::uint32_t ld_stream(const std::uint32_t *p) { return __ldcs(p); }

struct EulerState {
  double v[NVARS];
};

/// p = (gamma-1)(E - |m|^2/(2 rho)),  ideal_gas_eos.hpp:199-202, euler_variables.hpp:50-56
ZFVM_DEVICE double pressure_of(const double u[NVARS], double gamma) {
  double ekin = 0.5 * (u[1] * u[1] + u[2] * u[2] + u[3] * u[3]) / u[0];
  return (u[4] - ekin) * (gamma - 1.0);
}

// Reciprocal and reciprocal square root from the hardware's 2^-23 approximations plus two Newton steps: full
// double precision to within an ulp or two, at about a third of the FP64-pipe cost of the IEEE division /
// square root sequences (which is what bounds the flux kernel).  Arguments here are densities, pressures and
// wave-speed differences: finite, normal numbers; a non-positive argument of rsqrt yields NaN like sqrt would.
ZFVM_DEVICE double fast_rcp(double x) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  r = fma(fma(-x, r, 1.0), r, r);
  r = fma(fma(-x, r, 1.0), r, r);
  return r;
}
ZFVM_DEVICE double fast_rsqrt(double x) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  const double hx = 0.5 * x;
  y = y * fma(-hx * y, y, 1.5);
  y = y * fma(-hx * y, y, 1.5);
  return y;
}

/// std::min / std::max as comparisons (the CPU path's zisa::min / zisa::max, flux/hllc.hpp:73-74): unlike fmin / fmax
/// they do not drop a NaN operand, they return the first argument when the comparison is false.
ZFVM_DEVICE double ref_min(double a, double b) { return (b < a) ? b : a; }
ZFVM_DEVICE double ref_max(double a, double b) { return (a < b) ? b : a; }

/// x^(1/(gamma-1)) for the density of an isentropic state.  For the adiabatic indices in practical use the exponent
/// is a half-integer (gamma = 2, 5/3, 3/2, 7/5, 4/3 -> 1, 3/2, 2, 5/2, 3): square root and multiplications instead of
/// pow(), which is what the well-balanced reconstruction spends its time in (one evaluation per stencil member and
/// Gauss point, each cell with its own (h, K)).  The host classifies gamma once (SchemeConst::eos_pow_n =
/// 2 / (gamma - 1) when that is an integer in [2, 8], else 0): kernel-uniform predicates over straight-line code, no
/// per-point division, rounding test or counted loop.  Multiplication order: ((x rsqrt(x)) x) x .. / (x x) x ..
/// POWN > 0: the exponent n/2 is a compile-time constant (the well-balanced kernels are instantiated for gamma = 2,
/// 5/3, 7/5: n = 2, 3, 5); POWN == 0: decided at run time from sc.eos_pow_n.
template <int POWN = 0>
ZFVM_DEVICE double pow_inv_gamma_minus_one(double x, const SchemeConst &sc) {
  if constexpr (POWN > 0) {
    double pw = x;
    if constexpr ((POWN & 1) != 0) pw = (x * fast_rsqrt(x)) * x;
    if constexpr (POWN >= 4) pw *= x;
    if constexpr (POWN >= 6) pw *= x;
    if constexpr (POWN >= 8) pw *= x;
    return pw;
  } else {
    const int n = sc.eos_pow_n;  // x^(n/2)
    if (n == 2) return x;        // gamma = 2
    if (n == 0) return pow(x, sc.eos_pow_e);
    double pw = x;
    if (n & 1) pw = (x * fast_rsqrt(x)) * x;
    if (n >= 4) pw *= x;
    if (n >= 6) pw *= x;
    if (n >= 8) pw *= x;
    return pw;
  }
}

/// Isentropic ideal-gas state at specific enthalpy h and entropy function K
/// (ideal_gas_eos.hpp:162-168,188-191,208-215): rho = ((gamma-1) h / (gamma K))^(1/(gamma-1)), p = K rho^gamma,
/// E = p / (gamma-1).  rho^(gamma-1) is the base of that power, so p = K rho base needs no second pow().
/// c1 = (gamma-1) / (gamma K) and 1/(gamma-1) are formed once per equilibrium by the caller: no division per point.
template <int POWN = 0>
ZFVM_DEVICE void isentropic_state_c(double h, double K, double c1, const SchemeConst &sc, double inv_gm1, double &rho,
                                    double &E, double &p) {
  const double base = c1 * h;
  rho = pow_inv_gamma_minus_one<POWN>(base, sc);
  p = K * (rho * base);
  E = p * inv_gm1;
}

}  // namespace zfvm
