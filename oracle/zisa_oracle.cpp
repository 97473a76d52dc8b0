// ORACLE -- TEST INFRASTRUCTURE ONLY.  Not part of the product: only tests/, __graft_entry__.smoke()
// and bench.py's CPU-baseline / reference arm may load this library.
//
// CPU restatement (C++17 + OpenMP) of ZisaFVM's per-RK-stage residual path, following the reference
// file by file.  The reference itself cannot be compiled here (five un-vendored sibling repos, Eigen,
// Boost, HDF5 -- see DESIGN.md), so this is a "port" oracle.  Pinning: the pieces are checked
// against the reference's own known-answer / property tests (tests/test_oracle_kat.py lists every one
// with its file:line); there is no end-to-end golden vector in the reference, so END-TO-END PARITY IS
// UNPINNED BY THE REFERENCE and pinned piecewise only.
//
// Third-party arithmetic: Eigen 3.3.9 (conanfile.txt) `LDLT` at lsq_solver.cpp:47,82 -- restated
// here from Eigen's published unblocked algorithm (Eigen/src/Cholesky/LDLT.h, ldlt_inplace<Lower>).
//
// Deliberate deviations (all documented in DESIGN.md):
//   * face-flux scatter order is fixed (the reference uses OpenMP atomics, order nondeterministic);
//
// Reference map (all under /root/reference):
//   PolyND, smoothness_indicator            include/zisa/math/poly2d_impl.hpp:34-41,272-337
//   polynomial expressions                  include/zisa/math/polynomial_expr.hpp:18-156
//   LSQSolver::solve_impl                   src/zisa/reconstruction/lsq_solver.cpp:53-85
//   HybridWENO::compute_polys_impl          src/zisa/reconstruction/hybrid_weno.cpp:72-92
//   HybridWENO::eno_hybridize               src/zisa/reconstruction/hybrid_weno.cpp:110-128
//   CWENO_AO::reconstruct_impl              src/zisa/reconstruction/cweno_ao.cpp:36-53
//   WENO_AO::reconstruct                    src/zisa/reconstruction/weno_ao.cpp:14-24
//   LocalReconstruction                     include/zisa/reconstruction/local_reconstruction.hpp:69-163
//   EulerScaling / UnityScaling             include/zisa/model/characteristic_scale.hpp:14-46
//   IdealGasEOS                             include/zisa/model/ideal_gas_eos.hpp:14-287
//   IsentropicEquilibrium                   include/zisa/model/isentropic_equilibrium.hpp:27-63
//   LocalEquilibriumBase::solve_exact       include/zisa/model/local_equilibrium_impl.hpp:41-94
//   quasi_newton, RollingConvergenceRate    include/zisa/math/quasi_newton.hpp:12-50, rolling_convergence_rate.hpp
//   HLLCBatten, hllc_speeds, RoeAverage     include/zisa/flux/hllc.hpp:17-81,124-176
//   Euler::flux                             include/zisa/model/euler_impl.hpp:23-36
//   coord_transform                         src/zisa/model/euler_variables.cpp:29-68
//   FluxLoop::compute_patch                 include/zisa/fvm_loops/flux_loop.hpp:106-195
//   GravitySourceLoop                       include/zisa/fvm_loops/gravity_source_loop.hpp:32-87,121-148
//   runge_kutta_sum, RungeKutta::compute_step, make_tableau   src/zisa/ode/runge_kutta.cpp:87-213
//   FrozenBC::apply                         src/zisa/boundary/frozen_boundary_condition.cpp:39-55
//   LocalCFL                                include/zisa/model/local_cfl_condition_impl.hpp:25-40
#include <algorithm>
#include <array>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <string>
#include <vector>
#if defined(_OPENMP)
#include <omp.h>
#endif

namespace oracle {

using i32 = std::int32_t;
using i64 = std::int64_t;
constexpr int NV = 5;
constexpr int MAX_COEFFS = 35;  // PolyND<35, NV>: WENOPoly (weno_poly.hpp)

// ---------------------------------------------------------------------------------------------
// polynomials
// ---------------------------------------------------------------------------------------------
inline int poly_dof(int deg, int n_dims) {
  return n_dims == 2 ? ((deg + 1) * (deg + 2)) / 2 : ((deg + 1) * (deg + 2) * (deg + 3)) / 6;
}
inline int poly_index(int a, int b) {
  int n = a + b;
  return ((n + 1) * n) / 2 + b;
}
inline int poly_index(int a, int b, int c) { return poly_dof(a + b + c - 1, 3) + poly_index(b, c); }

template <int NVARS>
struct Poly {
  double coeffs[NVARS * MAX_COEFFS];
  double moments[MAX_COEFFS];
  int degree = 0;
  int n_dims = 0;
  double x_center[3] = {0, 0, 0};
  double reference_length = 1.0;

  Poly() {
    std::fill(coeffs, coeffs + NVARS * MAX_COEFFS, 0.0);
    std::fill(moments, moments + MAX_COEFFS, 0.0);
  }
  // PolyND(int degree, moments, x_center, reference_length, n_dims), poly2d_impl.hpp:88-107
  Poly(int degree_, const double *mom, int n_mom, const double *xc, double len, int nd) : Poly() {
    degree = degree_;
    n_dims = nd;
    for (int d = 0; d < 3; ++d) x_center[d] = xc[d];
    reference_length = len;
    moments[0] = moments[1] = moments[2] = 0.0;
    for (int i = 3; i < n_mom && i < MAX_COEFFS; ++i) moments[i] = mom[i];
  }
  int dof() const { return poly_dof(degree, n_dims); }

  // eval_2d / eval_3d, poly2d_impl.hpp:272-323 (same loop nest, same accumulation order)
  void eval(const double *xyz, double *px) const {
    const double x = (xyz[0] - x_center[0]) / reference_length;
    const double y = (xyz[1] - x_center[1]) / reference_length;
    const double z = (xyz[2] - x_center[2]) / reference_length;
    for (int k = 0; k < NVARS; ++k) px[k] = 0.0;
    const int d = degree;
    if (n_dims == 2) {
      double pow_x = 1.0;
      for (int kx = 0; kx <= d; ++kx) {
        double pow_y = 1.0;
        for (int ky = 0; ky <= d - kx; ++ky) {
          int i = poly_index(kx, ky);
          for (int k = 0; k < NVARS; ++k) px[k] += coeffs[i * NVARS + k] * (pow_x * pow_y - moments[i]);
          pow_y *= y;
        }
        pow_x *= x;
      }
    } else {
      double pow_x = 1.0;
      for (int kx = 0; kx <= d; ++kx) {
        double pow_y = 1.0;
        for (int ky = 0; ky <= d - kx; ++ky) {
          double pow_z = 1.0;
          for (int kz = 0; kz <= d - kx - ky; ++kz) {
            int i = poly_index(kx, ky, kz);
            for (int k = 0; k < NVARS; ++k) {
              double ak = coeffs[i * NVARS + k];
              px[k] += ak * (pow_x * pow_y * pow_z - moments[i]);
            }
            pow_z *= z;
          }
          pow_y *= y;
        }
        pow_x *= x;
      }
    }
  }
};

// p = a + b / a - b: coefficients combine pointwise, meta data from the higher-degree operand
// (e1.degree() > e2.degree() ? e1 : e2), polynomial_expr.hpp:33-56.
template <int NVARS>
void poly_meta_from(Poly<NVARS> &dst, const Poly<NVARS> &e1, const Poly<NVARS> &e2) {
  const Poly<NVARS> &m = (e1.degree > e2.degree) ? e1 : e2;
  double mom[MAX_COEFFS], xc[3];
  std::copy(m.moments, m.moments + MAX_COEFFS, mom);
  std::copy(m.x_center, m.x_center + 3, xc);
  const double len = m.reference_length;
  const int deg = std::max(e1.degree, e2.degree), nd = std::max(e1.n_dims, e2.n_dims);
  std::copy(mom, mom + MAX_COEFFS, dst.moments);
  std::copy(xc, xc + 3, dst.x_center);
  dst.reference_length = len;
  dst.degree = deg;
  dst.n_dims = nd;
}
// a -= alpha * b   (PolyND::operator-= with PointwiseScale, poly2d_impl.hpp:218-227)
template <int NVARS>
void poly_sub_scaled(Poly<NVARS> &a, double alpha, const Poly<NVARS> &b) {
  Poly<NVARS> r;
  for (int i = 0; i < NVARS * MAX_COEFFS; ++i) r.coeffs[i] = a.coeffs[i] - alpha * b.coeffs[i];
  poly_meta_from(r, a, b);
  a = r;
}
// a += alpha * b
template <int NVARS>
void poly_add_scaled(Poly<NVARS> &a, double alpha, const Poly<NVARS> &b) {
  Poly<NVARS> r;
  for (int i = 0; i < NVARS * MAX_COEFFS; ++i) r.coeffs[i] = a.coeffs[i] + alpha * b.coeffs[i];
  poly_meta_from(r, a, b);
  a = r;
}
// a /= alpha  ->  a *= 1.0 / alpha   (poly2d_impl.hpp:229-237)
template <int NVARS>
void poly_div(Poly<NVARS> &a, double alpha) {
  const double inv = 1.0 / alpha;
  for (int i = 0; i < NVARS * MAX_COEFFS; ++i) a.coeffs[i] = inv * a.coeffs[i];
}
// smoothness_indicator, poly2d_impl.hpp:326-337
template <int NVARS>
void smoothness_indicator(const Poly<NVARS> &p, double *beta) {
  for (int k = 0; k < NVARS; ++k) beta[k] = 0.0;
  const int n = p.dof();
  for (int i = poly_dof(0, p.n_dims); i < n; ++i)
    for (int k = 0; k < NVARS; ++k) beta[k] += p.coeffs[i * NVARS + k] * p.coeffs[i * NVARS + k];
}

// ---------------------------------------------------------------------------------------------
// Eigen::LDLT restated (Eigen 3.3.9, Cholesky/LDLT.h: ldlt_inplace<Lower>::unblocked, _solve_impl)
// ---------------------------------------------------------------------------------------------
struct LDLT {
  int n = 0;
  std::vector<double> m;   // lower triangle holds L (unit diagonal implied) and D on the diagonal
  std::vector<int> transp;

  void compute(const double *sym, int size) {
    n = size;
    m.assign(sym, sym + (size_t)size * size);
    transp.assign((size_t)size, 0);
    auto M = [&](int r, int c) -> double & { return m[(size_t)r * n + c]; };
    std::vector<double> temp((size_t)size);
    for (int k = 0; k < size; ++k) {
      int biggest = k;
      double big = std::abs(M(k, k));
      for (int i = k + 1; i < size; ++i)
        if (std::abs(M(i, i)) > big) {
          big = std::abs(M(i, i));
          biggest = i;
        }
      transp[(size_t)k] = biggest;
      if (k != biggest) {
        const int s = size - biggest - 1;
        for (int c = 0; c < k; ++c) std::swap(M(k, c), M(biggest, c));
        for (int r = 0; r < s; ++r) std::swap(M(size - s + r, k), M(size - s + r, biggest));
        std::swap(M(k, k), M(biggest, biggest));
        for (int i = k + 1; i < biggest; ++i) {
          double tmp = M(i, k);
          M(i, k) = M(biggest, i);
          M(biggest, i) = tmp;
        }
      }
      const int rs = size - k - 1;
      if (k > 0) {
        for (int c = 0; c < k; ++c) temp[(size_t)c] = M(c, c) * M(k, c);
        double s = 0.0;
        for (int c = 0; c < k; ++c) s += M(k, c) * temp[(size_t)c];
        M(k, k) -= s;
        for (int r = 0; r < rs; ++r) {
          double a = 0.0;
          for (int c = 0; c < k; ++c) a += M(k + 1 + r, c) * temp[(size_t)c];
          M(k + 1 + r, k) -= a;
        }
      }
      const double akk = M(k, k);
      const bool pivot_is_valid = std::abs(akk) > 0.0;
      if (k == 0 && !pivot_is_valid) {
        for (int j = 0; j < size; ++j) transp[(size_t)j] = j;
        return;
      }
      if (rs > 0 && pivot_is_valid)
        for (int r = 0; r < rs; ++r) M(k + 1 + r, k) /= akk;
    }
  }

  // x (n x nrhs, row-major) <- A^{-1} x
  void solve(double *x, int nrhs) const {
    auto M = [&](int r, int c) { return m[(size_t)r * n + c]; };
    for (int k = 0; k < n; ++k)
      if (transp[(size_t)k] != k)
        for (int c = 0; c < nrhs; ++c) std::swap(x[(size_t)k * nrhs + c], x[(size_t)transp[(size_t)k] * nrhs + c]);
    for (int r = 0; r < n; ++r)
      for (int c2 = 0; c2 < r; ++c2)
        for (int c = 0; c < nrhs; ++c) x[(size_t)r * nrhs + c] -= M(r, c2) * x[(size_t)c2 * nrhs + c];
    const double tolerance = 1.0 / std::numeric_limits<double>::max();
    for (int r = 0; r < n; ++r) {
      const double d = M(r, r);
      for (int c = 0; c < nrhs; ++c) {
        if (std::abs(d) > tolerance)
          x[(size_t)r * nrhs + c] /= d;
        else
          x[(size_t)r * nrhs + c] = 0.0;
      }
    }
    for (int r = n - 1; r >= 0; --r)
      for (int c2 = r + 1; c2 < n; ++c2)
        for (int c = 0; c < nrhs; ++c) x[(size_t)r * nrhs + c] -= M(c2, r) * x[(size_t)c2 * nrhs + c];
    for (int k = n - 1; k >= 0; --k)
      if (transp[(size_t)k] != k)
        for (int c = 0; c < nrhs; ++c) std::swap(x[(size_t)k * nrhs + c], x[(size_t)transp[(size_t)k] * nrhs + c]);
  }
};

// ---------------------------------------------------------------------------------------------
// equation of state, gravity-free equilibrium pieces
// ---------------------------------------------------------------------------------------------
struct IdealGasEOS {
  double gamma = 1.4, R = 1.0;
  double kinetic_energy(const double *u) const { return 0.5 * (u[1] * u[1] + u[2] * u[2] + u[3] * u[3]) / u[0]; }
  double internal_energy(const double *u) const { return u[4] - kinetic_energy(u); }
  double pressure_rhoE(double E) const { return E * (gamma - 1.0); }
  double pressure(const double *u) const { return pressure_rhoE(u[4] - kinetic_energy(u)); }
  double sound_speed_rhoP(double rho, double p) const { return std::sqrt(gamma * p / rho); }
  double sound_speed(const double *u) const { return sound_speed_rhoP(u[0], pressure(u)); }
  double enthalpy_rhoP(double rho, double p) const { return gamma / (gamma - 1.0) * p / rho; }
  double K_rhoP(double rho, double p) const { return p / std::pow(rho, gamma); }
  double rho_hK(double h, double K) const {
    double base = 1.0 / K * (gamma - 1.0) / gamma * h;
    double exponent = 1.0 / (gamma - 1.0);
    return std::pow(base, exponent);
  }
  double pressure_rhoK(double rho, double K) const { return K * std::pow(rho, gamma); }
  double ideal_internal_energy(double p) const { return p / (gamma - 1.0); }
  // rhoE(EnthalpyEntropy): ideal_gas_eos.hpp:88-93
  void rhoE_hK(double h, double K, double &rho, double &E) const {
    rho = rho_hK(h, K);
    double p = pressure_rhoK(rho_hK(h, K), K);  // pressure(theta) recomputes rho(theta)
    E = ideal_internal_energy(p);
  }
};

struct Params {
  int n_dims = 2;
  int recon_mode = 0;  // 0 CWENO-AO, 1 WENO-AO
  int n_stencils = 0;
  double linear_weights[8] = {0};
  double epsilon = 1e-10, exponent = 4.0;
  int well_balanced = 0;
  int scaling = 1;  // 0 unity, 1 euler
  int flux = 0;     // 0 HLLC, 1 Rusanov (not in the reference)
  double gamma = 1.4, gas_constant = 1.0;
  int has_gravity = 0;
  int flux_bc = 0;  // 0 NoFluxBC, 1 FluxBC (boundary/flux_bc.hpp:13-52), 2 EquilibriumFluxBC (equilibrium_flux_bc.hpp:18-75)
  int n_avars = 0;  // advected scalars, AllVariables::avars (all_variables.hpp:31-35)
  double heating_rate = 0.0, heating_r0 = 0.0, heating_r1 = 0.0;  // make_heating_source, heating.hpp:54-80
  // LocalRCParams (local_reconstruction.hpp:22-25; read from JSON at euler_experiment_impl.hpp:83-90)
  int steps_per_recompute = 1;
  double recompute_threshold = 0.0;
};

struct Grid {
  int n_dims = 0, F = 0, q_c = 0, q_f = 0, n_moments = 0;
  i64 n_cells = 0, n_edges = 0, n_interior_edges = 0;
  const i32 *left_right = nullptr, *edge_indices = nullptr;
  const double *volumes = nullptr, *cell_centers = nullptr, *char_length = nullptr, *moments = nullptr;
  const double *cell_qp = nullptr, *cell_qw = nullptr, *face_qp = nullptr, *face_qw = nullptr;
  const double *face_normal = nullptr, *face_t1 = nullptr, *face_t2 = nullptr, *inradii = nullptr;
  const std::uint8_t *cell_flags = nullptr;
  // gravity tables (phi, grad phi at every quadrature point; the gravity classes themselves are
  // state independent, gravity_decl.hpp:24-36)
  const double *phi_cqp = nullptr, *gradphi_cqp = nullptr, *phi_fqp = nullptr;
};

struct Stencils {
  int n_stencils = 0, l2g_stride = 0;
  const i32 *l2g = nullptr, *l2g_size = nullptr, *local = nullptr, *local_off = nullptr;
  const i32 *order = nullptr, *size = nullptr, *k_high = nullptr, *n_family = nullptr;
  const double *A = nullptr;  // [n_cells][A_stride], stencil k at A_off[k], row-major (size-1) x cols
  i64 A_stride = 0;
  const i64 *A_off = nullptr;
};

// LSQSolver: A and LDLT(A^T A), lsq_solver.cpp:40-47
struct LSQSolver {
  int order = 1, rows = 0, cols = 0;
  const double *A = nullptr;
  LDLT ldlt;
  void init(const double *A_, int rows_, int cols_, int order_) {
    A = A_;
    rows = rows_;
    cols = cols_;
    order = order_;
    if (order <= 1) return;
    std::vector<double> AtA((size_t)cols * cols);
    for (int i = 0; i < cols; ++i)
      for (int j = 0; j < cols; ++j) {
        double s = 0.0;
        for (int r = 0; r < rows; ++r) s += A[(size_t)r * cols + i] * A[(size_t)r * cols + j];
        AtA[(size_t)i * cols + j] = s;
      }
    ldlt.compute(AtA.data(), cols);
  }
  // solve_impl, lsq_solver.cpp:53-85
  template <int NVARS>
  Poly<NVARS> solve(const double *rhs, const Grid &g, i64 i_cell) const {
    const double *xc = &g.cell_centers[3 * i_cell];
    const double len = g.char_length[i_cell];
    if (order == 1) {
      double zero = 0.0;
      return Poly<NVARS>(0, &zero, 1, xc, len, g.n_dims);
    }
    Poly<NVARS> poly(order - 1, &g.moments[(size_t)i_cell * g.n_moments], g.n_moments, xc, len, g.n_dims);
    std::vector<double> x((size_t)cols * NVARS);
    for (int c = 0; c < cols; ++c)
      for (int v = 0; v < NVARS; ++v) {
        double s = 0.0;
        for (int r = 0; r < rows; ++r) s += A[(size_t)r * cols + c] * rhs[(size_t)r * NVARS + v];
        x[(size_t)c * NVARS + v] = s;
      }
    ldlt.solve(x.data(), NVARS);
    for (int c = 0; c < cols; ++c)
      for (int v = 0; v < NVARS; ++v) poly.coeffs[(c + 1) * NVARS + v] = x[(size_t)c * NVARS + v];
    return poly;
  }
};

// LocalEquilibrium<IsentropicEquilibrium<IdealGasEOS, Gravity>>
struct LocalEquilibrium {
  double h = 0, K = 0, phi_ref = 0;
  bool found = false;
  void extrapolate(const IdealGasEOS &eos, double phi, double &rho, double &E) const {
    if (!found) {
      rho = 0.0;
      E = 0.0;
      return;
    }
    eos.rhoE_hK(h + phi_ref - phi, K, rho, E);
  }
  // average over a cell (quadrature.hpp:33-62 accumulation, then / volume)
  static void cell_average(const IdealGasEOS &eos, double h, double K, double phi_ref, const double *phi,
                           const double *w, int q_c, double vol, double &rho_bar, double &E_bar) {
    double r, E;
    eos.rhoE_hK(h + phi_ref - phi[0], K, r, E);
    double ar = w[0] * r, aE = w[0] * E;
    for (int q = 1; q < q_c; ++q) {
      eos.rhoE_hK(h + phi_ref - phi[q], K, r, E);
      ar = ar + w[q] * r;
      aE = aE + w[q] * E;
    }
    ar = 1.0 * ar;
    aE = 1.0 * aE;
    rho_bar = ar / vol;
    E_bar = aE / vol;
  }
  void extrapolate_cell(const IdealGasEOS &eos, const Grid &g, i64 j, double &rho_bar, double &E_bar) const {
    if (!found) {
      rho_bar = 0.0;
      E_bar = 0.0;
      return;
    }
    cell_average(eos, h, K, phi_ref, &g.phi_cqp[(size_t)j * g.q_c], &g.cell_qw[(size_t)j * g.q_c], g.q_c, g.volumes[j],
                 rho_bar, E_bar);
  }
  // solve_exact, local_equilibrium_impl.hpp:41-94 with quasi_newton.hpp:12-50
  void solve(const IdealGasEOS &eos, const Grid &g, i64 i, double rho_bar, double E_bar) {
    const double *phi = &g.phi_cqp[(size_t)i * g.q_c];
    const double *w = &g.cell_qw[(size_t)i * g.q_c];
    const double vol = g.volumes[i];
    phi_ref = phi[0];
    const double p0 = eos.pressure_rhoE(E_bar);
    const double h0 = eos.enthalpy_rhoP(rho_bar, p0), K0 = eos.K_rhoP(rho_bar, p0);
    auto f = [&](double hh, double KK, double &f0, double &f1) {
      double rb, Eb;
      cell_average(eos, hh, KK, phi_ref, phi, w, g.q_c, vol, rb, Eb);
      f0 = rho_bar - rb;
      f1 = E_bar - Eb;
    };
    const double atol[2] = {1e-13 * h0, 1e-13 * K0};
    double x[2] = {h0, K0}, fx[2];
    f(x[0], x[1], fx[0], fx[1]);
    double dx[2] = {2.0 * atol[0] + 1.0, 2.0 * atol[1] + 1.0};
    // RollingConvergenceRate
    double values[8][2];
    int i_end = 0;
    auto idx = [](int a) { return (a + 8) % 8; };
    auto is_converged = [&](double factor) {
      return std::abs(dx[0]) <= factor * atol[0] && std::abs(dx[1]) <= factor * atol[1];
    };
    int iter = 0;
    bool ok = true;
    while (!is_converged(1.0) && iter < 20) {
      double df0[2], df1[2];
      {
        double eps = 1e-6 * std::abs(x[0]);
        double fp[2], fm[2];
        f(x[0] + 0.5 * eps * 1.0, x[1] + 0.5 * eps * 0.0, fp[0], fp[1]);
        f(x[0] - 0.5 * eps * 1.0, x[1] - 0.5 * eps * 0.0, fm[0], fm[1]);
        df0[0] = (fp[0] - fm[0]) / eps;
        df0[1] = (fp[1] - fm[1]) / eps;
      }
      {
        double eps = 1e-6 * std::abs(x[1]);
        double fp[2], fm[2];
        f(x[0] + 0.5 * eps * 0.0, x[1] + 0.5 * eps * 1.0, fp[0], fp[1]);
        f(x[0] - 0.5 * eps * 0.0, x[1] - 0.5 * eps * 1.0, fm[0], fm[1]);
        df1[0] = (fp[0] - fm[0]) / eps;
        df1[1] = (fp[1] - fm[1]) / eps;
      }
      const double inv_det = 1.0 / (df0[0] * df1[1] - df0[1] * df1[0]);
      dx[0] = inv_det * (df1[1] * fx[0] - df1[0] * fx[1]);
      dx[1] = inv_det * (-df0[1] * fx[0] + df0[0] * fx[1]);
      x[0] -= dx[0];
      x[1] -= dx[1];
      f(x[0], x[1], fx[0], fx[1]);
      values[i_end][0] = dx[0];
      values[i_end][1] = dx[1];
      i_end = idx(i_end + 1);
      if (iter >= 4) {
        bool conv = values[idx(i_end - 1)][0] <= atol[0] && values[idx(i_end - 1)][1] <= atol[1];
        if (!conv) {
          conv = true;
          for (int c = 0; c < 2; ++c) {
            const double a = values[idx(i_end - 3)][c], b = values[idx(i_end - 2)][c], cc = values[idx(i_end - 1)][c];
            const double rate = std::log(std::abs(cc) / std::abs(b)) / std::log(std::abs(b) / std::abs(a));
            conv = conv && (rate >= 0.0);
          }
        }
        if (!conv) {  // reference: LOG_ERR("Is not converging.") terminates; treated as "not found"
          ok = false;
          break;
        }
      }
      ++iter;
    }
    if (ok && iter == 20 && !is_converged(1000.0)) ok = false;
    found = ok;
    h = ok ? x[0] : h0;
    K = ok ? x[1] : K0;
  }
};

struct WorkPolys {
  std::vector<Poly<NV>> polys;
  std::vector<double> rhs, qbar;
};

struct PointValues {  // FewPointsCache payload: (RhoE, xvars) at the cell's own cell+face points
  double rho = 0, E = 0, p = 0, a = 0;
};

struct Oracle {
  Params prm;
  Grid g;
  Stencils st;
  IdealGasEOS eos;
  std::vector<LSQSolver> lsq;         // [n_cells][n_stencils]
  std::vector<double> lin_w_full;     // normalised
  // per-cell reconstruction state (LocalReconstruction members)
  std::vector<Poly<NV>> weno_poly;    // [n_cells]
  std::vector<double> scale;          // [n_cells][5]
  std::vector<LocalEquilibrium> eq;   // [n_cells]
  std::vector<PointValues> pv_cell;   // [n_cells][q_c]
  std::vector<PointValues> pv_face;   // [n_cells][F][q_f]
  std::vector<Poly<1>> scalar_polys;   // [n_cells][n_avars]  (LocalReconstruction::scalar_polys)
  std::vector<i32> steps_since;        // [n_cells]  LocalReconstruction::steps_since_recompute
  std::vector<double> rhoEbar_cache;   // [n_cells][l2g_stride][2]  LocalReconstruction::rhoEbar_cache
  std::vector<double> frozen, frozen_av;
  std::vector<i32> ghost_index;
  bool has_frozen = false;
  int eq_failures = 0;

  void init() {
    eos.gamma = prm.gamma;
    eos.R = prm.gas_constant;
    const i64 n = g.n_cells;
    const int ns = st.n_stencils;
    lsq.resize((size_t)(n * ns));
    double tot = 0.0;
    for (int k = 0; k < ns; ++k) tot += prm.linear_weights[k];
    lin_w_full.resize((size_t)ns);
    for (int k = 0; k < ns; ++k) lin_w_full[(size_t)k] = prm.linear_weights[k] / tot;  // hybrid_weno.cpp:26-31
#pragma omp parallel for schedule(dynamic, 64)
    for (i64 i = 0; i < n; ++i)
      for (int k = 0; k < st.n_family[i]; ++k) {
        const int order = st.order[i * ns + k], size = st.size[i * ns + k];
        const int cols = order > 1 ? poly_dof(order - 1, g.n_dims) - 1 : 0;
        lsq[(size_t)(i * ns + k)].init(st.A + i * st.A_stride + st.A_off[k], size - 1, cols, order);
      }
    weno_poly.resize((size_t)n);
    scale.assign((size_t)(n * NV), 1.0);
    eq.resize((size_t)n);
    pv_cell.assign((size_t)(n * g.q_c), PointValues());
    pv_face.assign((size_t)(n * g.F * g.q_f), PointValues());
    steps_since.assign((size_t)n, 0);
    rhoEbar_cache.assign((size_t)(n * st.l2g_stride * 2), 0.0);
    for (i64 i = 0; i < n; ++i)
      if (g.cell_flags[i] & 2) ghost_index.push_back((i32)i);
  }

  // eno_hybridize, hybrid_weno.cpp:110-128
  template <int NVV>
  Poly<NVV> hybridize(const std::vector<Poly<NVV>> &polys, const double *lw, int n_st) const {
    double nlw[8];
    double al_tot = 0.0;
    for (int k = 0; k < n_st; ++k) {
      double beta[NVV];
      smoothness_indicator(polys[(size_t)k], beta);
      double IS = beta[0];
      for (int v = 1; v < NVV; ++v) IS = std::max(IS, beta[v]);
      double al = lw[k] / (prm.epsilon + std::pow(IS, prm.exponent));
      nlw[k] = al;
      al_tot += al;
    }
    double zero = 0.0, xz[3] = {0, 0, 0};
    Poly<NVV> p(0, &zero, 1, xz, 1.0, 2);
    for (int k = 0; k < n_st; ++k) poly_add_scaled(p, nlw[k] / al_tot, polys[(size_t)k]);
    return p;
  }

  // LocalReconstruction::compute_tracer (local_reconstruction.hpp:127-147) for cell i: every advected scalar is
  // reconstructed on its own (no equilibrium, no scaling; its own smoothness indicators and non-linear weights)
  // with the cell's stencils and LSQ solvers: rc.reconstruct(rhs_view, polys, q_component) ->
  // compute_polys_impl (hybrid_weno.cpp:72-92) + CWENO_AO / WENO_AO for ScalarPoly (cweno_ao.cpp:24-34, weno_ao.cpp:26-36)
  void reconstruct_tracers_cell(i64 i, const double *avars, std::vector<Poly<1>> &polys, std::vector<double> &rhs) {
    const int ns = st.n_stencils, na = prm.n_avars;
    const int n_st = st.n_family[i];
    const i32 *l2g = st.l2g + i * st.l2g_stride;
    for (int a = 0; a < na; ++a) {
      const double q0 = avars[(size_t)l2g[0] * na + a];
      polys.resize((size_t)n_st);
      for (int k = 0; k < n_st; ++k) {
        const int size = st.size[i * ns + k];
        const i32 *loc = st.local + i * st.l2g_stride + st.local_off[k];
        rhs.resize((size_t)std::max(size - 1, 1));
        for (int ig = 0; ig < size - 1; ++ig) rhs[(size_t)ig] = avars[(size_t)l2g[loc[ig + 1]] * na + a] - q0;
        polys[(size_t)k] = lsq[(size_t)(i * ns + k)].solve<1>(rhs.data(), g, i);
        polys[(size_t)k].coeffs[0] = q0;
      }
      double lw[8];
      if (n_st == 1 && st.order[i * ns] == 1) {
        lw[0] = 1.0;
      } else {
        for (int k = 0; k < n_st; ++k) lw[k] = lin_w_full[(size_t)k];
      }
      if (prm.recon_mode == 0) {
        const int k_high = st.k_high[i];
        for (int k = 0; k < n_st; ++k)
          if (k_high != k) poly_sub_scaled(polys[(size_t)k_high], lw[k], polys[(size_t)k]);
        poly_div(polys[(size_t)k_high], lw[k_high]);
      }
      scalar_polys[(size_t)(i * na + a)] = hybridize(polys, lw, n_st);
    }
  }

  void global_tracer_reconstruction(const double *avars) {
    const i64 n = g.n_cells;
    scalar_polys.resize((size_t)(n * prm.n_avars));
#pragma omp parallel
    {
      std::vector<Poly<1>> polys;
      std::vector<double> rhs;
#pragma omp for schedule(static, 8)
      for (i64 i = 0; i < n; ++i) reconstruct_tracers_cell(i, avars, polys, rhs);
    }
  }

  // LocalReconstruction::compute for cell i (local_reconstruction.hpp:69-120)
  void reconstruct_cell(i64 i, const double *state, WorkPolys &wk) {
    const int ns = st.n_stencils;
    const int n_st = st.n_family[i];
    const i32 *l2g = st.l2g + i * st.l2g_stride;
    const int m = st.l2g_size[i];
    wk.qbar.resize((size_t)m * NV);
    for (int il = 0; il < m; ++il)  // set_qbar_local, global_reconstruction_impl.hpp:166-174
      for (int v = 0; v < NV; ++v) wk.qbar[(size_t)il * NV + v] = state[(size_t)l2g[il] * NV + v];

    // recompute_equilibrium, local_reconstruction.hpp:87-100: every steps_per_recompute-th call, or when the cell has
    // drifted from the cached equilibrium average by recompute_threshold (in units of the cached scale)
    const double *u0 = &wk.qbar[0];
    const double rho_self = u0[0], E_self = eos.internal_energy(u0);
    double *sc = &scale[(size_t)i * NV];
    double *cache = &rhoEbar_cache[(size_t)(i * st.l2g_stride) * 2];
    bool recompute = (steps_since[(size_t)i] % prm.steps_per_recompute) == 0;
    if (!recompute) {
      const double d0 = (rho_self - cache[0]) / sc[0], d1 = (E_self - cache[1]) / sc[4];
      recompute = std::sqrt(d0 * d0 + d1 * d1) >= prm.recompute_threshold;
    }
    LocalEquilibrium &le = eq[(size_t)i];
    if (recompute) {  // compute_equilibrium, local_reconstruction.hpp:69-85
      if (prm.scaling == 1) {  // EulerScaling, characteristic_scale.hpp:24-33
        const double p = eos.pressure_rhoE(E_self);
        const double cs = eos.sound_speed_rhoP(rho_self, p);
        sc[0] = rho_self;
        sc[1] = sc[2] = sc[3] = cs;
        sc[4] = E_self;
      } else {
        for (int v = 0; v < NV; ++v) sc[v] = 1.0;
      }
      if (prm.well_balanced) {
        le.solve(eos, g, i, rho_self, E_self);
        if (!le.found) {
#pragma omp atomic
          eq_failures += 1;
        }
        // point_values_cache.update: extrapolate_full at own cell + face points
        for (int q = 0; q < g.q_c; ++q) {
          PointValues &pv = pv_cell[(size_t)(i * g.q_c + q)];
          le.extrapolate(eos, g.phi_cqp[(size_t)(i * g.q_c + q)], pv.rho, pv.E);
          pv.p = le.found ? eos.pressure_rhoE(pv.E) : 0.0;
          pv.a = le.found ? eos.sound_speed_rhoP(pv.rho, pv.p) : 0.0;
        }
        for (int k = 0; k < g.F; ++k) {
          const i64 e = g.edge_indices[i * g.F + k];
          for (int q = 0; q < g.q_f; ++q) {
            PointValues &pv = pv_face[(size_t)((i * g.F + k) * g.q_f + q)];
            le.extrapolate(eos, g.phi_fqp[(size_t)(e * g.q_f + q)], pv.rho, pv.E);
            pv.p = le.found ? eos.pressure_rhoE(pv.E) : 0.0;
            pv.a = le.found ? eos.sound_speed_rhoP(pv.rho, pv.p) : 0.0;
          }
        }
      }
      for (int il = 0; il < m; ++il) {
        double rho_eq_bar = 0.0, E_eq_bar = 0.0;
        if (prm.well_balanced) le.extrapolate_cell(eos, g, l2g[il], rho_eq_bar, E_eq_bar);
        cache[2 * il] = rho_eq_bar;
        cache[2 * il + 1] = E_eq_bar;
      }
      steps_since[(size_t)i] = 0;
    }
    for (int il = 0; il < m; ++il) {
      double *u = &wk.qbar[(size_t)il * NV];
      u[0] -= cache[2 * il];
      u[4] -= cache[2 * il + 1];
      for (int v = 0; v < NV; ++v) u[v] = u[v] / sc[v];
    }
    steps_since[(size_t)i] += 1;

    // compute_polys_impl, hybrid_weno.cpp:72-92
    wk.polys.resize((size_t)n_st);
    for (int k = 0; k < n_st; ++k) {
      const int size = st.size[i * ns + k];
      const i32 *loc = st.local + i * st.l2g_stride + st.local_off[k];
      wk.rhs.resize((size_t)std::max(size - 1, 1) * NV);
      for (int ig = 0; ig < size - 1; ++ig) {
        const int il = loc[ig + 1];
        for (int v = 0; v < NV; ++v) wk.rhs[(size_t)ig * NV + v] = wk.qbar[(size_t)il * NV + v] - wk.qbar[v];
      }
      wk.polys[(size_t)k] = lsq[(size_t)(i * ns + k)].solve<NV>(wk.rhs.data(), g, i);
      for (int v = 0; v < NV; ++v) wk.polys[(size_t)k].coeffs[v] = wk.qbar[v];
    }
    double lw[8];
    if (n_st == 1 && st.order[i * ns] == 1) {
      lw[0] = 1.0;  // o1_params, global_reconstruction_decl.hpp:120-133
    } else {
      for (int k = 0; k < n_st; ++k) lw[k] = lin_w_full[(size_t)k];
    }
    if (prm.recon_mode == 0) {  // CWENO_AO::reconstruct_impl, cweno_ao.cpp:36-53
      const int k_high = st.k_high[i];
      for (int k = 0; k < n_st; ++k)
        if (k_high != k) poly_sub_scaled(wk.polys[(size_t)k_high], lw[k], wk.polys[(size_t)k]);
      poly_div(wk.polys[(size_t)k_high], lw[k_high]);
    }
    weno_poly[(size_t)i] = hybridize(wk.polys, lw, n_st);
  }

  // rc(i)(x) = background(x).first + scale * weno_poly(x), local_reconstruction.hpp:149-163
  void point_value(i64 i, const double *x, const PointValues *bg, double *u) const {
    double px[NV];
    weno_poly[(size_t)i].eval(x, px);
    const double *sc = &scale[(size_t)i * NV];
    const double b[NV] = {bg ? bg->rho : 0.0, 0.0, 0.0, 0.0, bg ? bg->E : 0.0};
    for (int v = 0; v < NV; ++v) u[v] = b[v] + sc[v] * px[v];
  }

  void global_reconstruction(const double *state) {
    const i64 n = g.n_cells;
#pragma omp parallel
    {
      WorkPolys wk;
#pragma omp for schedule(static, 8)
      for (i64 i = 0; i < n; ++i) reconstruct_cell(i, state, wk);
    }
  }

  // ---- numerical fluxes --------------------------------------------------------------------
  void euler_flux(const double *u, double p, double *pf) const {
    double v = u[1] / u[0];
    pf[0] = u[1];
    pf[1] = v * u[1] + p;
    pf[2] = v * u[2];
    pf[3] = v * u[3];
    pf[4] = v * (u[4] + p);
  }

  void hllc(const double *uL, const double *uR, double *nf, double *speeds = nullptr) const {
    const double pL = eos.pressure(uL), aL = eos.sound_speed(uL);
    const double pR = eos.pressure(uR), aR = eos.sound_speed(uR);
    const double roe_ratio = std::sqrt(uR[0] / uL[0]);
    auto roe = [roe_ratio](double qL, double qR) { return (qL + qR * roe_ratio) / (1.0 + roe_ratio); };
    double vL = uL[1] / uL[0], vR = uR[1] / uR[0];
    double v_tilda = roe(vL, vR);
    double HL = (uL[4] + pL) / uL[0], HR = (uR[4] + pR) / uR[0];
    double H_tilda = roe(HL, HR);
    double r1 = roe(uL[1] / uL[0], uR[1] / uR[0]), r2 = roe(uL[2] / uL[0], uR[2] / uR[0]),
           r3 = roe(uL[3] / uL[0], uR[3] / uR[0]);
    double vroe_square = r1 * r1 + r2 * r2 + r3 * r3;
    double a_tilda = std::sqrt((eos.gamma - 1.0) * (H_tilda - 0.5 * vroe_square));
    double sL = std::min(vL - aL, v_tilda - a_tilda);
    double sR = std::max(vR + aR, v_tilda + a_tilda);
    double s_star = (uR[1] * (sR - vR) - uL[1] * (sL - vL) + pL - pR) / (uR[0] * (sR - vR) - uL[0] * (sL - vL));
    if (speeds) {
      speeds[0] = sL;
      speeds[1] = s_star;
      speeds[2] = sR;
    }
    const double *uK = (0.0 <= s_star ? uL : uR);
    const double pK = (0.0 <= s_star ? pL : pR);
    euler_flux(uK, pK, nf);
    if (sL < 0.0 && 0.0 <= sR) {
      double sK = (0.0 <= s_star ? sL : sR);
      double vK = (0.0 <= s_star ? uL[1] / uL[0] : uR[1] / uR[0]);
      double cK = (sK - vK) / (sK - s_star);
      nf[0] += sK * (cK * uK[0] - uK[0]);
      nf[1] += sK * (cK * uK[0] * s_star - uK[1]);
      nf[2] += sK * (cK * uK[2] - uK[2]);
      nf[3] += sK * (cK * uK[3] - uK[3]);
      nf[4] += sK * (cK * (uK[4] + (s_star - vK) * (uK[0] * s_star + pK / (sK - vK))) - uK[4]);
    }
  }

  // HLLCBatten::tracer_flux, flux/hllc.hpp:178-197
  static double hllc_tracer_flux(const double *uL, const double *uR, double mqL, double mqR, const double *speeds) {
    const double sL = speeds[0], s_star = speeds[1], sR = speeds[2];
    double mqK = (0.0 <= s_star ? mqL : mqR);
    double vK = (0.0 <= s_star ? uL[1] / uL[0] : uR[1] / uR[0]);
    double fK = mqK * vK;
    if (sL < 0.0 && 0.0 < sR) {
      double sK = (0.0 <= s_star ? sL : sR);
      double cK = (sK - vK) / (sK - s_star);
      return fK + sK * (cK * mqK - mqK);
    }
    return fK;
  }
  // Rusanov tracer flux: like the Rusanov flux itself not in the reference; same local Lax-Friedrichs form
  double rusanov_tracer_flux(const double *uL, const double *uR, double mqL, double mqR) const {
    const double aL = eos.sound_speed(uL), aR = eos.sound_speed(uR);
    const double vL = uL[1] / uL[0], vR = uR[1] / uR[0];
    const double lam = std::max(std::abs(vL) + aL, std::abs(vR) + aR);
    return 0.5 * (mqL * vL + mqR * vR) - 0.5 * lam * (mqR - mqL);
  }

  // Not in the reference: defined by this project, parity is against this only.
  void rusanov(const double *uL, const double *uR, double *nf) const {
    const double pL = eos.pressure(uL), aL = eos.sound_speed(uL);
    const double pR = eos.pressure(uR), aR = eos.sound_speed(uR);
    double fL[NV], fR[NV];
    euler_flux(uL, pL, fL);
    euler_flux(uR, pR, fR);
    const double lam = std::max(std::abs(uL[1] / uL[0]) + aL, std::abs(uR[1] / uR[0]) + aR);
    for (int v = 0; v < NV; ++v) nf[v] = 0.5 * (fL[v] + fR[v]) - 0.5 * lam * (uR[v] - uL[v]);
  }

  static void coord_transform(double *u, const double *n, const double *t1, const double *t2) {
    double un = u[1] * n[0] + u[2] * n[1] + u[3] * n[2];
    double ut1 = u[1] * t1[0] + u[2] * t1[1] + u[3] * t1[2];
    double ut2 = u[1] * t2[0] + u[2] * t2[1] + u[3] * t2[2];
    u[1] = un;
    u[2] = ut1;
    u[3] = ut2;
  }
  static void inv_coord_transform(double *u, const double *n, const double *t1, const double *t2) {
    double ux = u[1] * n[0] + u[2] * t1[0] + u[3] * t2[0];
    double uy = u[1] * n[1] + u[2] * t1[1] + u[3] * t2[1];
    double uz = u[1] * n[2] + u[2] * t1[2] + u[3] * t2[2];
    u[1] = ux;
    u[2] = uy;
    u[3] = uz;
  }

  int local_face(i64 i, i64 e) const {
    for (int k = 0; k < g.F; ++k)
      if (g.edge_indices[i * g.F + k] == e) return k;
    return -1;
  }

  // FluxLoop::compute_patch, flux_loop.hpp:106-195.  The face fluxes are computed in parallel and
  // scattered serially in edge order (deviation: deterministic order instead of omp atomic).
  void flux_loop(double *tendency, std::vector<double> &face_flux, double *tendency_av = nullptr) {
    const i64 EI = g.n_interior_edges;
    const int na = tendency_av ? prm.n_avars : 0;
    std::vector<double> face_qflux((size_t)(EI * std::max(na, 1)), 0.0);
    face_flux.assign((size_t)(EI * NV), 0.0);
    std::vector<std::uint8_t> active((size_t)EI, 0);
#pragma omp parallel for schedule(static, 8)
    for (i64 e = 0; e < EI; ++e) {
      const i64 iL = g.left_right[2 * e], iR = g.left_right[2 * e + 1];
      if ((g.cell_flags[iL] & 2) && (g.cell_flags[iR] & 2)) continue;  // flux_loop.hpp:82-87
      active[(size_t)e] = 1;
      const double *n = &g.face_normal[3 * e], *t1 = &g.face_t1[3 * e], *t2 = &g.face_t2[3 * e];
      const int kL = local_face(iL, e), kR = local_face(iR, e);
      double nf[NV] = {0, 0, 0, 0, 0};
      double qnf[16] = {0};
      for (int k = 0; k < g.q_f; ++k) {
        const double w = g.face_qw[(size_t)(e * g.q_f + k)];
        const double *x = &g.face_qp[(size_t)((e * g.q_f + k) * 3)];
        double uL[NV], uR[NV];
        point_value(iL, x, prm.well_balanced ? &pv_face[(size_t)((iL * g.F + kL) * g.q_f + k)] : nullptr, uL);
        point_value(iR, x, prm.well_balanced ? &pv_face[(size_t)((iR * g.F + kR) * g.q_f + k)] : nullptr, uR);
        coord_transform(uL, n, t1, t2);
        coord_transform(uR, n, t1, t2);
        double f[NV], speeds[3] = {0, 0, 0};
        if (prm.flux == 0)
          hllc(uL, uR, f, speeds);
        else
          rusanov(uL, uR, f);
        for (int v = 0; v < NV; ++v) nf[v] += w * f[v];
        for (int a = 0; a < na; ++a) {  // flux_loop.hpp:157-161
          double qL, qR;
          scalar_polys[(size_t)(iL * na + a)].eval(x, &qL);
          scalar_polys[(size_t)(iR * na + a)].eval(x, &qR);
          qnf[a] += w * (prm.flux == 0 ? hllc_tracer_flux(uL, uR, qL, qR, speeds) : rusanov_tracer_flux(uL, uR, qL, qR));
        }
      }
      inv_coord_transform(nf, n, t1, t2);
      for (int v = 0; v < NV; ++v) face_flux[(size_t)(e * NV + v)] = nf[v];
      for (int a = 0; a < na; ++a) face_qflux[(size_t)(e * na + a)] = qnf[a];
    }
    for (i64 e = 0; e < EI; ++e) {
      if (!active[(size_t)e]) continue;
      const i64 iL = g.left_right[2 * e], iR = g.left_right[2 * e + 1];
      for (int v = 0; v < NV; ++v) {
        const double nfL = face_flux[(size_t)(e * NV + v)] / g.volumes[iL];
        tendency[iL * NV + v] -= nfL;
        const double nfR = face_flux[(size_t)(e * NV + v)] / g.volumes[iR];
        tendency[iR * NV + v] += nfR;
      }
      for (int a = 0; a < na; ++a) {  // flux_loop.hpp:180-192
        const double qfL = face_qflux[(size_t)(e * na + a)] / g.volumes[iL];
        tendency_av[iL * na + a] -= qfL;
        const double qfR = face_qflux[(size_t)(e * na + a)] / g.volumes[iR];
        tendency_av[iR * na + a] += qfR;
      }
    }
  }

  // GravitySourceLoop, gravity_source_loop.hpp:32-87 (well-balanced) and :121-148 (NoEquilibrium)
  void gravity_source_loop(double *tendency) {
    const i64 n = g.n_cells;
#pragma omp parallel for schedule(static, 8)
    for (i64 i = 0; i < n; ++i) {
      const double vol = g.volumes[i];
      if (prm.well_balanced) {
        const double *xc = &g.cell_centers[3 * i];
        double s[NV] = {0, 0, 0, 0, 0};
        for (int k = 0; k < g.F; ++k) {
          const i64 e = g.edge_indices[i * g.F + k];
          const double *nrm = &g.face_normal[3 * e];
          const double *x0 = &g.face_qp[(size_t)(e * g.q_f * 3)];
          // unit_outward_normal, face.cpp:24-27
          const double dt = nrm[0] * (x0[0] - xc[0]) + nrm[1] * (x0[1] - xc[1]) + nrm[2] * (x0[2] - xc[2]);
          const double sg = (dt > 0.0) ? 1.0 : ((dt < 0.0) ? -1.0 : 0.0);
          const double no[3] = {sg * nrm[0], sg * nrm[1], sg * nrm[2]};
          double acc[NV];
          for (int q = 0; q < g.q_f; ++q) {
            const double p_eq = pv_face[(size_t)((i * g.F + k) * g.q_f + q)].p;
            const double w = g.face_qw[(size_t)(e * g.q_f + q)];
            const double sq[NV] = {0.0, p_eq * no[0], p_eq * no[1], p_eq * no[2], 0.0};
            for (int v = 0; v < NV; ++v) acc[v] = (q == 0) ? w * sq[v] : acc[v] + w * sq[v];
          }
          for (int v = 0; v < NV; ++v) s[v] += 1.0 * acc[v];
        }
        double acc[NV];
        for (int q = 0; q < g.q_c; ++q) {
          const double *x = &g.cell_qp[(size_t)((i * g.q_c + q) * 3)];
          double px[NV], du[NV];
          weno_poly[(size_t)i].eval(x, px);
          for (int v = 0; v < NV; ++v) du[v] = scale[(size_t)i * NV + v] * px[v];
          const PointValues &pv = pv_cell[(size_t)(i * g.q_c + q)];
          const double u[NV] = {pv.rho + du[0], 0.0 + du[1], 0.0 + du[2], 0.0 + du[3], pv.E + du[4]};
          const double drho = du[0];
          const double *gp = &g.gradphi_cqp[(size_t)((i * g.q_c + q) * 3)];
          const double sq[NV] = {0.0, -drho * gp[0], -drho * gp[1], -drho * gp[2],
                                 -(u[1] * gp[0] + u[2] * gp[1] + u[3] * gp[2])};
          const double w = g.cell_qw[(size_t)(i * g.q_c + q)];
          for (int v = 0; v < NV; ++v) acc[v] = (q == 0) ? w * sq[v] : acc[v] + w * sq[v];
        }
        for (int v = 0; v < NV; ++v) s[v] += 1.0 * acc[v];
        for (int v = 0; v < NV; ++v) tendency[i * NV + v] += s[v] / vol;
      } else {
        double acc[NV];
        for (int q = 0; q < g.q_c; ++q) {
          const double *x = &g.cell_qp[(size_t)((i * g.q_c + q) * 3)];
          double u[NV];
          point_value(i, x, nullptr, u);
          const double *gp = &g.gradphi_cqp[(size_t)((i * g.q_c + q) * 3)];
          const double sq[NV] = {0.0, -u[0] * gp[0], -u[0] * gp[1], -u[0] * gp[2],
                                 -(u[1] * gp[0] + u[2] * gp[1] + u[3] * gp[2])};
          const double w = g.cell_qw[(size_t)(i * g.q_c + q)];
          for (int v = 0; v < NV; ++v) acc[v] = (q == 0) ? w * sq[v] : acc[v] + w * sq[v];
        }
        for (int v = 0; v < NV; ++v) tendency[i * NV + v] += (1.0 * acc[v]) / vol;  // average(cell, s)
      }
    }
  }

  // Sum[FluxLoop, GravitySourceLoop]::compute (accumulating; ZeroRateOfChange is the caller's)
  void rate_of_change(double *tendency, const double *state, double *tendency_av = nullptr,
                      const double *state_av = nullptr) {
    global_reconstruction(state);
    if (prm.n_avars > 0 && state_av && tendency_av) global_tracer_reconstruction(state_av);
    std::vector<double> ff;
    flux_loop(tendency, ff, (prm.n_avars > 0 && state_av) ? tendency_av : nullptr);
    if (prm.has_gravity) gravity_source_loop(tendency);
    if (prm.heating_rate != 0.0) heating_loop(tendency);
    if (prm.flux_bc == 1) flux_bc_loop(tendency, state);
    if (prm.flux_bc == 2) equilibrium_flux_bc_loop(tendency, state);
  }

  // Heating::compute, model/heating.hpp:30-44 with the rate of make_heating_source (:54-80):
  // dudt(i, 4) += average(cell, rho(x) * epsilon * [r0 <= |x| <= r1]), rho = first component of rc(i)(x)
  void heating_loop(double *tendency) {
    const i64 n = g.n_cells;
#pragma omp parallel for schedule(static, 8)
    for (i64 i = 0; i < n; ++i) {
      double acc = 0.0;
      for (int q = 0; q < g.q_c; ++q) {
        const double *x = &g.cell_qp[(size_t)((i * g.q_c + q) * 3)];
        double u[NV];
        point_value(i, x, prm.well_balanced ? &pv_cell[(size_t)(i * g.q_c + q)] : nullptr, u);
        const double r = std::sqrt(x[0] * x[0] + x[1] * x[1] + x[2] * x[2]);
        const double rate = prm.heating_rate * ((prm.heating_r0 <= r) && (r <= prm.heating_r1) ? 1.0 : 0.0);
        const double f = u[0] * rate;
        const double w = g.cell_qw[(size_t)(i * g.q_c + q)];
        acc = (q == 0) ? w * f : acc + w * f;
      }
      tendency[i * NV + 4] += (1.0 * acc) / g.volumes[i];
    }
  }

  // EquilibriumFluxBC::compute, boundary/equilibrium_flux_bc.hpp:37-63: per exterior face the local equilibrium of
  // the adjacent cell is solved from its average (rho, E_int) and the flux of the (resting) equilibrium state at the
  // face Gauss points, i.e. its pressure along the face normal, leaves the cell.
  void equilibrium_flux_bc_loop(double *tendency, const double *state) const {
    for (i64 e = g.n_interior_edges; e < g.n_edges; ++e) {
      const i64 i = g.left_right[2 * e];
      const double *n = &g.face_normal[3 * e], *t1 = &g.face_t1[3 * e], *t2 = &g.face_t2[3 * e];
      LocalEquilibrium le;
      le.solve(eos, g, i, state[i * NV], eos.internal_energy(&state[i * NV]));
      double acc[NV] = {0, 0, 0, 0, 0};
      for (int q = 0; q < g.q_f; ++q) {
        double rho, E;
        le.extrapolate(eos, g.phi_fqp[(size_t)(e * g.q_f + q)], rho, E);
        double u[NV] = {rho, 0.0, 0.0, 0.0, E}, f[NV];
        euler_flux(u, le.found ? eos.pressure(u) : 0.0, f);  // (reference: 0/0 when the solve failed)
        if (!le.found) f[0] = f[2] = f[3] = f[4] = 0.0;
        inv_coord_transform(f, n, t1, t2);
        const double w = g.face_qw[(size_t)(e * g.q_f + q)];
        for (int v = 0; v < NV; ++v) acc[v] = (q == 0) ? w * f[v] : acc[v] + w * f[v];
      }
      for (int v = 0; v < NV; ++v) tendency[i * NV + v] -= (1.0 * acc[v]) / g.volumes[i];
    }
  }

  // FluxBC::compute, boundary/flux_bc.hpp:24-42: on every exterior face the physical flux of the adjacent cell's
  // *average* state leaves the cell: tendency(i) -= |face| / |cell| * F(u_i) . n
  void flux_bc_loop(double *tendency, const double *state) const {
    for (i64 e = g.n_interior_edges; e < g.n_edges; ++e) {
      const i64 i = g.left_right[2 * e];
      const double *n = &g.face_normal[3 * e], *t1 = &g.face_t1[3 * e], *t2 = &g.face_t2[3 * e];
      double u[NV], f[NV];
      for (int v = 0; v < NV; ++v) u[v] = state[i * NV + v];
      coord_transform(u, n, t1, t2);
      euler_flux(u, eos.pressure(u), f);
      inv_coord_transform(f, n, t1, t2);
      double area = 0.0;
      for (int q = 0; q < g.q_f; ++q) area += g.face_qw[(size_t)(e * g.q_f + q)];
      const double c = area / g.volumes[i];
      for (int v = 0; v < NV; ++v) tendency[i * NV + v] -= c * f[v];
    }
  }

  void apply_bc(double *u) const {
    if (!has_frozen) return;
    for (i32 i : ghost_index)
      for (int v = 0; v < NV; ++v) u[(size_t)i * NV + v] = frozen[(size_t)i * NV + v];
  }

  void apply_bc_av(double *a) const {  // FrozenBC::apply also resets avars (frozen_boundary_condition.cpp:39-55)
    if (!has_frozen || frozen_av.empty()) return;
    const int na = prm.n_avars;
    for (i32 i : ghost_index)
      for (int v = 0; v < na; ++v) a[(size_t)i * na + v] = frozen_av[(size_t)i * na + v];
  }

  double cfl_dt(const double *u, double cfl_number) const {
    double m = std::numeric_limits<double>::max();
    for (i64 i = 0; i < g.n_cells; ++i) {
      const double *ui = &u[i * NV];
      const double a = eos.sound_speed(ui);
      const double v2 = (ui[1] * ui[1] + ui[2] * ui[2] + ui[3] * ui[3]) / (ui[0] * ui[0]);
      m = std::min(m, g.inradii[i] / (std::sqrt(v2) + a));
    }
    return cfl_number * m;
  }
};

struct Tableau {
  int n = 0;
  double a[6][6] = {{0}};
  double b[6] = {0};
};

bool make_tableau(const std::string &m, Tableau &t) {
  t = Tableau();
  if (m == "forward_euler") {
    t.n = 1;
    t.b[0] = 1.0;
  } else if (m == "ssp2") {
    t.n = 2;
    t.a[1][0] = 1.0;
    t.b[0] = 0.5;
    t.b[1] = 0.5;
  } else if (m == "ssp3") {
    t.n = 3;
    t.a[1][0] = 1.0;
    t.a[2][0] = 0.25;
    t.a[2][1] = 0.25;
    t.b[0] = 1.0 / 6;
    t.b[1] = 1.0 / 6;
    t.b[2] = 2.0 / 3;
  } else if (m == "wicker") {
    t.n = 3;
    t.a[1][0] = 1.0 / 3;
    t.a[2][1] = 0.5;
    t.b[2] = 1.0;
  } else if (m == "rk4") {
    t.n = 4;
    t.a[1][0] = 0.5;
    t.a[2][1] = 0.5;
    t.a[3][2] = 1.0;
    t.b[0] = 1.0 / 6;
    t.b[1] = 1.0 / 3;
    t.b[2] = 1.0 / 3;
    t.b[3] = 1.0 / 6;
  } else if (m == "fehlberg") {
    t.n = 6;
    const double a[6][6] = {{0.0, 0.0, 0.0, 0.0, 0.0, 0.0},
                            {0.25, 0.0, 0.0, 0.0, 0.0, 0.0},
                            {3.0 / 32.0, 9.0 / 32.0, 0.0, 0.0, 0.0, 0.0},
                            {1932.0 / 2197.0, -7200.0 / 2197.0, 7296.0 / 2197.0, 0.0, 0.0, 0.0},
                            {439.0 / 216.0, -8.0, 3680.0 / 513.0, -845.0 / 4104, 0.0, 0.0},
                            {-8.0 / 27.0, 2.0, -3544.0 / 2565.0, 1859.0 / 4104.0, -11.0 / 40.0, 0.0}};
    const double b[6] = {16.0 / 135.0, 0.0, 6656.0 / 12825.0, 28561.0 / 56430.0, -9.0 / 50.0, 2.0 / 55.0};
    for (int r = 0; r < 6; ++r) {
      t.b[r] = b[r];
      for (int c = 0; c < 6; ++c) t.a[r][c] = a[r][c];
    }
  } else {
    return false;
  }
  return true;
}

// runge_kutta_sum, runge_kutta.cpp:122-143
void runge_kutta_sum(double *u1, const double *u0, const std::vector<std::vector<double>> &k, const double *coeffs,
                     int n_stages, double dt, i64 n) {
#pragma omp parallel for schedule(static)
  for (i64 i = 0; i < n; ++i) {
    double dudt = 0.0;
    for (int s = 0; s < n_stages; ++s)
      if (coeffs[s] != 0.0) dudt += coeffs[s] * k[(size_t)s][(size_t)i];
    u1[i] = u0[i] + dt * dudt;
  }
}

}  // namespace oracle

using namespace oracle;

extern "C" {

struct oracle_grid_desc {
  int n_dims, q_c, q_f, n_moments;
  int64_t n_cells, n_edges, n_interior_edges;
  const int32_t *left_right, *edge_indices;
  const double *volumes, *cell_centers, *char_length, *moments, *cell_qp, *cell_qw, *face_qp, *face_qw;
  const double *face_normal, *face_t1, *face_t2, *inradii;
  const uint8_t *cell_flags;
  const double *phi_cqp, *gradphi_cqp, *phi_fqp;
};

struct oracle_stencil_desc {
  int n_stencils, l2g_stride;
  const int32_t *l2g, *l2g_size, *local, *local_off, *order, *size, *k_high, *n_family;
  const double *A;
  int64_t A_stride;
  const int64_t *A_off;
};

struct oracle_params {
  int recon_mode;
  double linear_weights[8];
  double epsilon, exponent;
  int well_balanced, scaling, flux;
  double gamma, gas_constant;
  int has_gravity;
  int flux_bc;
};

void *oracle_create(const oracle_grid_desc *gd, const oracle_stencil_desc *sd, const oracle_params *pp) {
  Oracle *o = new Oracle();
  Grid &g = o->g;
  g.n_dims = gd->n_dims;
  g.F = gd->n_dims + 1;
  g.q_c = gd->q_c;
  g.q_f = gd->q_f;
  g.n_moments = gd->n_moments;
  g.n_cells = gd->n_cells;
  g.n_edges = gd->n_edges;
  g.n_interior_edges = gd->n_interior_edges;
  g.left_right = gd->left_right;
  g.edge_indices = gd->edge_indices;
  g.volumes = gd->volumes;
  g.cell_centers = gd->cell_centers;
  g.char_length = gd->char_length;
  g.moments = gd->moments;
  g.cell_qp = gd->cell_qp;
  g.cell_qw = gd->cell_qw;
  g.face_qp = gd->face_qp;
  g.face_qw = gd->face_qw;
  g.face_normal = gd->face_normal;
  g.face_t1 = gd->face_t1;
  g.face_t2 = gd->face_t2;
  g.inradii = gd->inradii;
  g.cell_flags = gd->cell_flags;
  g.phi_cqp = gd->phi_cqp;
  g.gradphi_cqp = gd->gradphi_cqp;
  g.phi_fqp = gd->phi_fqp;
  Stencils &s = o->st;
  s.n_stencils = sd->n_stencils;
  s.l2g_stride = sd->l2g_stride;
  s.l2g = sd->l2g;
  s.l2g_size = sd->l2g_size;
  s.local = sd->local;
  s.local_off = sd->local_off;
  s.order = sd->order;
  s.size = sd->size;
  s.k_high = sd->k_high;
  s.n_family = sd->n_family;
  s.A = sd->A;
  s.A_stride = sd->A_stride;
  s.A_off = sd->A_off;
  Params &p = o->prm;
  p.n_dims = gd->n_dims;
  p.recon_mode = pp->recon_mode;
  p.n_stencils = sd->n_stencils;
  for (int k = 0; k < 8; ++k) p.linear_weights[k] = pp->linear_weights[k];
  p.epsilon = pp->epsilon;
  p.exponent = pp->exponent;
  p.well_balanced = pp->well_balanced;
  p.scaling = pp->scaling;
  p.flux = pp->flux;
  p.gamma = pp->gamma;
  p.gas_constant = pp->gas_constant;
  p.has_gravity = pp->has_gravity;
  p.flux_bc = pp->flux_bc;
  o->init();
  return o;
}

void oracle_destroy(void *h) { delete (Oracle *)h; }

void oracle_set_frozen_bc(void *h, const double *steady) {
  Oracle *o = (Oracle *)h;
  if (!steady) {
    o->has_frozen = false;
    return;
  }
  o->frozen.assign(steady, steady + o->g.n_cells * NV);
  o->has_frozen = true;
}

/* Sum[FluxLoop, GravitySourceLoop]::compute; tendency is accumulated into. */
void oracle_rate_of_change(void *h, double *tendency, const double *state) {
  ((Oracle *)h)->rate_of_change(tendency, state);
}

/* EulerGlobalReconstruction::compute only; coefficients [n][n_coef][5] in the scaled basis, scales [n][5]. */
void oracle_reconstruct(void *h, const double *state, double *coeffs, int n_coef, double *scale) {
  Oracle *o = (Oracle *)h;
  o->global_reconstruction(state);
  for (int64_t i = 0; i < o->g.n_cells; ++i) {
    for (int c = 0; c < n_coef; ++c)
      for (int v = 0; v < NV; ++v) coeffs[(i * n_coef + c) * NV + v] = o->weno_poly[(size_t)i].coeffs[c * NV + v];
    for (int v = 0; v < NV; ++v) scale[i * NV + v] = o->scale[(size_t)i * NV + v];
  }
}

/* value of the reconstruction of cell i at x (rc(i)(x)); for points of the cell's own rules only
 * when well-balanced (the FewPointsCache holds exactly those): kind 0 cell point q, 1 face (k, q) */
void oracle_point_value(void *h, int64_t i, int kind, int k, int q, double *u) {
  Oracle *o = (Oracle *)h;
  const Grid &g = o->g;
  if (kind == 0) {
    const double *x = &g.cell_qp[(size_t)((i * g.q_c + q) * 3)];
    o->point_value(i, x, o->prm.well_balanced ? &o->pv_cell[(size_t)(i * g.q_c + q)] : nullptr, u);
  } else {
    const int64_t e = g.edge_indices[i * g.F + k];
    const double *x = &g.face_qp[(size_t)((e * g.q_f + q) * 3)];
    o->point_value(i, x, o->prm.well_balanced ? &o->pv_face[(size_t)((i * g.F + k) * g.q_f + q)] : nullptr, u);
  }
}

/* rc(i)(x) at every cell Gauss point of every cell: out[n_cells][q_c][5] (after oracle_reconstruct) */
void oracle_cell_point_values(void *h, double *out) {
  Oracle *o = (Oracle *)h;
  const Grid &g = o->g;
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < g.n_cells; ++i)
    for (int q = 0; q < g.q_c; ++q)
      o->point_value(i, &g.cell_qp[(size_t)((i * g.q_c + q) * 3)],
                     o->prm.well_balanced ? &o->pv_cell[(size_t)(i * g.q_c + q)] : nullptr, &out[(size_t)((i * g.q_c + q) * NV)]);
}

void oracle_eval_at(void *h, int64_t i, const double *x, double *u) { ((Oracle *)h)->point_value(i, x, nullptr, u); }

/* RungeKutta::compute_step, runge_kutta.cpp:87-112 with Sum[Zero, Sum[FluxLoop, GravitySourceLoop]] */
int oracle_rk_step(void *h, const char *method, const double *u0, double *u1, double dt) {
  Oracle *o = (Oracle *)h;
  Tableau t;
  if (!make_tableau(method, t)) return 1;
  const int64_t n = o->g.n_cells * NV;
  std::vector<std::vector<double>> k((size_t)t.n, std::vector<double>((size_t)n, 0.0));
  std::vector<double> ux((size_t)n);
  o->rate_of_change(k[0].data(), u0);
  for (int stage = 1; stage < t.n; ++stage) {
    runge_kutta_sum(ux.data(), u0, k, t.a[stage], t.n, dt, n);
    o->apply_bc(ux.data());
    std::fill(k[(size_t)stage].begin(), k[(size_t)stage].end(), 0.0);  // ZeroRateOfChange
    o->rate_of_change(k[(size_t)stage].data(), ux.data());
  }
  runge_kutta_sum(u1, u0, k, t.b, t.n, dt, n);
  o->apply_bc(u1);
  return 0;
}

/* ---- extensions: advected scalars (a27), Heating, EquilibriumFluxBC -------------------------------- */
void oracle_set_tracers(void *h, int n_avars) { ((Oracle *)h)->prm.n_avars = n_avars; }
void oracle_set_heating(void *h, double rate, double r0, double r1) {
  Oracle *o = (Oracle *)h;
  o->prm.heating_rate = rate;
  o->prm.heating_r0 = r0;
  o->prm.heating_r1 = r1;
}
void oracle_set_flux_bc(void *h, int kind) { ((Oracle *)h)->prm.flux_bc = kind; }
/* LocalRCParams{steps_per_recompute, recompute_threshold}; resets the per-cell counters */
void oracle_set_local_rc_params(void *h, int steps_per_recompute, double recompute_threshold) {
  Oracle *o = (Oracle *)h;
  o->prm.steps_per_recompute = steps_per_recompute < 1 ? 1 : steps_per_recompute;
  o->prm.recompute_threshold = recompute_threshold;
  std::fill(o->steps_since.begin(), o->steps_since.end(), 0);
}
void oracle_set_frozen_bc_av(void *h, const double *steady, const double *steady_av) {
  Oracle *o = (Oracle *)h;
  oracle_set_frozen_bc(h, steady);
  if (steady_av && o->prm.n_avars > 0)
    o->frozen_av.assign(steady_av, steady_av + o->g.n_cells * o->prm.n_avars);
  else
    o->frozen_av.clear();
}
void oracle_rate_of_change_av(void *h, double *tendency, double *tendency_av, const double *state, const double *state_av) {
  ((Oracle *)h)->rate_of_change(tendency, state, tendency_av, state_av);
}
/* scalar polynomial coefficients [n][n_avars][n_coef] after a rate_of_change_av */
void oracle_tracer_polys(void *h, double *coeffs, int n_coef) {
  Oracle *o = (Oracle *)h;
  const int na = o->prm.n_avars;
  for (int64_t i = 0; i < o->g.n_cells * na; ++i)
    for (int c = 0; c < n_coef; ++c) coeffs[i * n_coef + c] = o->scalar_polys[(size_t)i].coeffs[c];
}
double oracle_hllc_tracer_flux(double gamma, const double *uL, const double *uR, double mqL, double mqR) {
  Oracle o;
  o.eos.gamma = gamma;
  double nf[NV], speeds[3];
  o.hllc(uL, uR, nf, speeds);
  return Oracle::hllc_tracer_flux(uL, uR, mqL, mqR, speeds);
}
/* RungeKutta::compute_step on AllVariables{cvars, avars} */
int oracle_rk_step_av(void *h, const char *method, const double *u0, const double *a0, double *u1, double *a1, double dt) {
  Oracle *o = (Oracle *)h;
  Tableau t;
  if (!make_tableau(method, t)) return 1;
  const int na = o->prm.n_avars;
  const int64_t n = o->g.n_cells * NV, m = o->g.n_cells * na;
  std::vector<std::vector<double>> k((size_t)t.n, std::vector<double>((size_t)n, 0.0));
  std::vector<std::vector<double>> ka((size_t)t.n, std::vector<double>((size_t)m, 0.0));
  std::vector<double> ux((size_t)n), ax((size_t)m);
  o->rate_of_change(k[0].data(), u0, ka[0].data(), a0);
  for (int stage = 1; stage < t.n; ++stage) {
    runge_kutta_sum(ux.data(), u0, k, t.a[stage], t.n, dt, n);
    runge_kutta_sum(ax.data(), a0, ka, t.a[stage], t.n, dt, m);
    o->apply_bc(ux.data());
    o->apply_bc_av(ax.data());
    std::fill(k[(size_t)stage].begin(), k[(size_t)stage].end(), 0.0);
    std::fill(ka[(size_t)stage].begin(), ka[(size_t)stage].end(), 0.0);
    o->rate_of_change(k[(size_t)stage].data(), ux.data(), ka[(size_t)stage].data(), ax.data());
  }
  runge_kutta_sum(u1, u0, k, t.b, t.n, dt, n);
  runge_kutta_sum(a1, a0, ka, t.b, t.n, dt, m);
  o->apply_bc(u1);
  o->apply_bc_av(a1);
  return 0;
}

double oracle_cfl_dt(void *h, const double *u, double cfl_number) { return ((Oracle *)h)->cfl_dt(u, cfl_number); }
int oracle_eq_failures(void *h) { return ((Oracle *)h)->eq_failures; }

/* ---- stand-alone pieces for the known-answer tests ------------------------------------------------ */
void oracle_hllc(double gamma, const double *uL, const double *uR, double *nf) {
  Oracle o;
  o.eos.gamma = gamma;
  o.hllc(uL, uR, nf);
}
void oracle_rusanov(double gamma, const double *uL, const double *uR, double *nf) {
  Oracle o;
  o.eos.gamma = gamma;
  o.rusanov(uL, uR, nf);
}
void oracle_euler_flux(double gamma, const double *u, double *pf) {
  Oracle o;
  o.eos.gamma = gamma;
  o.euler_flux(u, o.eos.pressure(u), pf);
}
int oracle_poly_dof(int deg, int n_dims) { return poly_dof(deg, n_dims); }
int oracle_poly_index2(int a, int b) { return poly_index(a, b); }
int oracle_poly_index3(int a, int b, int c) { return poly_index(a, b, c); }
/* PolyND<35, n_vars> evaluation: coeffs [dof][n_vars], moments [dof] */
void oracle_poly_eval(int n_dims, int degree, int n_vars, const double *coeffs, const double *moments, int n_mom,
                      const double *x_center, double length, const double *x, double *out) {
  if (n_vars == 1) {
    Poly<1> p(degree, moments, n_mom, x_center, length, n_dims);
    for (int i = 0; i < poly_dof(degree, n_dims); ++i) p.coeffs[i] = coeffs[i];
    p.eval(x, out);
  } else if (n_vars == 2) {
    Poly<2> p(degree, moments, n_mom, x_center, length, n_dims);
    for (int i = 0; i < 2 * poly_dof(degree, n_dims); ++i) p.coeffs[i] = coeffs[i];
    p.eval(x, out);
  } else {
    Poly<NV> p(degree, moments, n_mom, x_center, length, n_dims);
    for (int i = 0; i < NV * poly_dof(degree, n_dims); ++i) p.coeffs[i] = coeffs[i];
    p.eval(x, out);
  }
}
/* 0.2 p + q - 0.4 p on single-variable polynomials (poly2d.cpp:131-141 "saxpy-like") */
void oracle_poly_saxpy(int n_dims, int deg_p, const double *cp, const double *mp, int deg_q, const double *cq,
                       const double *mq, const double *x, double *out) {
  double xc[3] = {0, 0, 0};
  Poly<1> p(deg_p, mp, poly_dof(deg_p, n_dims), xc, 1.0, n_dims), q(deg_q, mq, poly_dof(deg_q, n_dims), xc, 1.0, n_dims);
  for (int i = 0; i < poly_dof(deg_p, n_dims); ++i) p.coeffs[i] = cp[i];
  for (int i = 0; i < poly_dof(deg_q, n_dims); ++i) q.coeffs[i] = cq[i];
  double zero = 0.0;
  Poly<1> r(0, &zero, 1, xc, 1.0, n_dims);
  poly_add_scaled(r, 0.2, p);
  poly_add_scaled(r, 1.0, q);
  poly_sub_scaled(r, 0.4, p);
  r.eval(x, out);
}
/* LDLT(A^T A).solve(A^T rhs): coefficients [cols][nrhs] */
void oracle_lsq_solve(const double *A, int rows, int cols, const double *rhs, int nrhs, double *x) {
  LSQSolver s;
  s.init(A, rows, cols, 2);
  for (int c = 0; c < cols; ++c)
    for (int v = 0; v < nrhs; ++v) {
      double acc = 0.0;
      for (int r = 0; r < rows; ++r) acc += A[(size_t)r * cols + c] * rhs[(size_t)r * nrhs + v];
      x[(size_t)c * nrhs + v] = acc;
    }
  s.ldlt.solve(x, nrhs);
}
/* IdealGasEOS conversions for model/eos.cpp */
void oracle_eos_rhoE_to_hK(double gamma, double rho, double E, double *h, double *K) {
  IdealGasEOS e;
  e.gamma = gamma;
  const double p = e.pressure_rhoE(E);
  *h = e.enthalpy_rhoP(rho, p);
  *K = e.K_rhoP(rho, p);
}
void oracle_eos_hK_to_rhoE(double gamma, double h, double K, double *rho, double *E) {
  IdealGasEOS e;
  e.gamma = gamma;
  e.rhoE_hK(h, K, *rho, *E);
}
/* LocalEquilibrium::solve + extrapolate on one cell given potentials at its Gauss points
 * (model/local_equilibrium.cpp:15-60). */
int oracle_local_equilibrium(double gamma, int q_c, const double *phi_cell, const double *w_cell, double vol,
                             double rho_bar, double E_bar, double *h, double *K, double *phi_ref) {
  Oracle o;
  o.eos.gamma = gamma;
  Grid &g = o.g;
  g.q_c = q_c;
  g.n_cells = 1;
  g.phi_cqp = phi_cell;
  g.cell_qw = w_cell;
  g.volumes = &vol;
  LocalEquilibrium le;
  le.solve(o.eos, g, 0, rho_bar, E_bar);
  *h = le.h;
  *K = le.K;
  *phi_ref = le.phi_ref;
  return le.found ? 1 : 0;
}
/* generic RungeKutta::compute_step on a user rate of change du/dt = f(t, u) (ode/runge_kutta.cpp test) */
typedef void (*oracle_rhs_fn)(double t, const double *u, double *dudt, int64_t n);
int oracle_rk_generic(const char *method, oracle_rhs_fn f, double *u, int64_t n, double t, double dt) {
  Tableau tb;
  if (!make_tableau(method, tb)) return 1;
  double c[6] = {0};
  for (int s = 0; s < tb.n; ++s)
    for (int j = 0; j < tb.n; ++j) c[s] += tb.a[s][j];
  std::vector<std::vector<double>> k((size_t)tb.n, std::vector<double>((size_t)n, 0.0));
  std::vector<double> u0(u, u + n), ux((size_t)n);
  f(t, u0.data(), k[0].data(), n);
  for (int stage = 1; stage < tb.n; ++stage) {
    runge_kutta_sum(ux.data(), u0.data(), k, tb.a[stage], tb.n, dt, n);
    f(t + c[stage] * dt, ux.data(), k[(size_t)stage].data(), n);
  }
  runge_kutta_sum(u, u0.data(), k, tb.b, tb.n, dt, n);
  return 0;
}

int oracle_num_threads(void) {
#if defined(_OPENMP)
  return omp_get_max_threads();
#else
  return 1;
#endif
}

}  // extern "C"
