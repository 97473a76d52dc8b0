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

/// Isentropic ideal-gas state at specific enthalpy h and entropy function K
/// (ideal_gas_eos.hpp:162-168,188-191,208-215).
ZFVM_DEVICE void isentropic_state(double h, double K, double gamma, double &rho, double &E, double &p) {
  double base = 1.0 / K * (gamma - 1.0) / gamma * h;
  double exponent = 1.0 / (gamma - 1.0);
  rho = (gamma == 2.0) ? base : pow(base, exponent);
  p = K * ((gamma == 2.0) ? rho * rho : pow(rho, gamma));
  E = p / (gamma - 1.0);
}

}  // namespace zfvm
