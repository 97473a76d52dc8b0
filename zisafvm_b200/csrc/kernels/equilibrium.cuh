// Local isentropic equilibrium of a cell (shared by the reconstruction kernel and EquilibriumFluxBC in K3).
#pragma once
#include "common.cuh"

namespace zfvm {

/// Local isentropic equilibrium of one cell (local_equilibrium_impl.hpp:34-94,
/// isentropic_equilibrium.hpp:27-63). All potentials come from precomputed tables.
struct LocalEq {
  double h_ref, K, phi_ref;
  bool found;
  double c1 = 0.0, inv_gm1 = 0.0;  // (gamma-1) / (gamma K), 1 / (gamma-1): set by prepare()
  ZFVM_DEVICE void prepare(double gamma) {
    c1 = (gamma - 1.0) / (gamma * K);
    inv_gm1 = 1.0 / (gamma - 1.0);
  }
  ZFVM_DEVICE void at(double phi, double gamma, double &rho, double &E, double &p) const {
    if (!found) {
      rho = 0.0;
      E = 0.0;
      p = 0.0;
      return;
    }
    isentropic_state_c(h_ref + phi_ref - phi, K, c1, gamma, inv_gm1, rho, E, p);
  }
};

/// cell average of the equilibrium over a cell whose Gauss-point potentials are `phi` (AoS row)
ZFVM_DEVICE void eq_cell_average(const LocalEq &eq, const double *__restrict__ phi, const SchemeConst &sc,
                                 double &rho_bar, double &E_bar) {
  if (!eq.found) {
    rho_bar = 0.0;
    E_bar = 0.0;
    return;
  }
  double r, E, p;
  eq.at(phi[0], sc.gamma, r, E, p);
  rho_bar = sc.cell_w[0] * r;
  E_bar = sc.cell_w[0] * E;
  for (int q = 1; q < sc.q_c; ++q) {
    eq.at(phi[q], sc.gamma, r, E, p);
    rho_bar += sc.cell_w[q] * r;
    E_bar += sc.cell_w[q] * E;
  }
}

/// quasi_newton (quasi_newton.hpp:12-50) on f(theta) = rhoE_bar - avg_cell rhoE_eq(theta) with
/// the central-difference Jacobian of local_equilibrium_impl.hpp:64-86.
ZFVM_DEVICE LocalEq solve_local_equilibrium(double rho_bar, double E_bar, const double *__restrict__ phi_own,
                                            const SchemeConst &sc) {
  const double gamma = sc.gamma;
  LocalEq eq;
  eq.phi_ref = phi_own[0];  // x_ref = first cell Gauss point
  eq.found = true;
  const double p0 = E_bar * (gamma - 1.0);
  const double h0 = gamma / (gamma - 1.0) * p0 / rho_bar;
  const double K0 = p0 / ((gamma == 2.0) ? rho_bar * rho_bar : pow(rho_bar, gamma));
  const double atol_h = 1e-13 * h0, atol_K = 1e-13 * K0;

  auto f = [&](double h, double K, double &f0, double &f1) {
    LocalEq t{h, K, eq.phi_ref, true};
    t.prepare(gamma);
    double rb, Eb;
    eq_cell_average(t, phi_own, sc, rb, Eb);
    f0 = rho_bar - rb;
    f1 = E_bar - Eb;
  };

  double h = h0, K = K0, f0, f1;
  f(h, K, f0, f1);
  double dx0 = 2.0 * atol_h + 1.0, dx1 = 2.0 * atol_K + 1.0;
  double a0 = 0, a1 = 0, b0 = 0, b1 = 0;  // previous two steps (rolling_convergence_rate.hpp)
  int iter = 0;
  bool ok = true;
  while (!(fabs(dx0) <= atol_h && fabs(dx1) <= atol_K) && iter < 20) {
    double eps_h = 1e-6 * fabs(h), eps_K = 1e-6 * fabs(K);
    double fp0, fp1, fm0, fm1;
    f(h + 0.5 * eps_h, K, fp0, fp1);
    f(h - 0.5 * eps_h, K, fm0, fm1);
    const double d00 = (fp0 - fm0) / eps_h, d01 = (fp1 - fm1) / eps_h;  // df0 = d f / d h
    f(h, K + 0.5 * eps_K, fp0, fp1);
    f(h, K - 0.5 * eps_K, fm0, fm1);
    const double d10 = (fp0 - fm0) / eps_K, d11 = (fp1 - fm1) / eps_K;  // df1 = d f / d K
    const double inv_det = 1.0 / (d00 * d11 - d01 * d10);
    dx0 = inv_det * (d11 * f0 - d10 * f1);
    dx1 = inv_det * (-d01 * f0 + d00 * f1);
    h -= dx0;
    K -= dx1;
    f(h, K, f0, f1);
    if (iter >= 4) {
      bool conv = (dx0 <= atol_h && dx1 <= atol_K);
      if (!conv) {
        double r0 = log(fabs(dx0) / fabs(b0)) / log(fabs(b0) / fabs(a0));
        double r1 = log(fabs(dx1) / fabs(b1)) / log(fabs(b1) / fabs(a1));
        conv = (r0 >= 0.0 && r1 >= 0.0);
      }
      if (!conv) {
        ok = false;
        break;
      }
    }
    a0 = b0;
    a1 = b1;
    b0 = dx0;
    b1 = dx1;
    ++iter;
  }
  if (ok && iter == 20 && !(fabs(dx0) <= 1000.0 * atol_h && fabs(dx1) <= 1000.0 * atol_K)) ok = false;
  eq.h_ref = ok ? h : h0;
  eq.K = ok ? K : K0;
  eq.found = ok;
  eq.prepare(gamma);
  return eq;
}

}  // namespace zfvm
