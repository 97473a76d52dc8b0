// Gravity potentials on the host: phi and grad phi are state independent, so they are tabulated
// once at every quadrature point (model/gravity_decl.hpp:17-353, model/gravity_impl.hpp:13-99,
// math/linear_interpolation.hpp:14-45).
#include <algorithm>
#include <cmath>
#include <limits>

#include "gravity.hpp"

namespace zfvm {

namespace {
const double PI = 3.14159265358979323846;
}

double GravityModel::coordinate(const double x[3]) const {
  if (alignment == 0) return std::sqrt(x[0] * x[0] + x[1] * x[1] + x[2] * x[2]);
  return x[0] * axis[0] + x[1] * axis[1] + x[2] * axis[2];
}

double GravityModel::dx(const double x[3], int dir) const {
  if (alignment == 0) {
    double r = std::sqrt(x[0] * x[0] + x[1] * x[1] + x[2] * x[2]);
    return x[dir] / (r + 1e-50);  // RadialAlignment::epsilon
  }
  return axis[dir];
}

size_t GravityModel::table_index(double r) const {
  size_t i = (size_t)(std::lower_bound(table_r.begin(), table_r.end(), r) - table_r.begin());
  if (i == 0) return 0;
  return std::min(i - 1, table_r.size() - 2);
}

double GravityModel::phi_chi(double chi) const {
  switch (kind) {
    case 1: return p[0] * chi;
    case 2: return p[0] / (p[1] + chi);  // GM / (X + chi)
    case 3: {
      const double rhoC = p[0], K = p[1], G = p[2];
      const double alpha = std::sqrt(2.0 * PI * G / K);
      const double chi_eff = alpha * (chi + std::numeric_limits<double>::min());
      return -2.0 * K * rhoC * std::sin(chi_eff) / chi_eff;
    }
    case 4: {
      // NonUniformLinearInterpolation::operator(), linear_interpolation.hpp:20-41
      const size_t i = table_index(chi);
      const double a = (chi - table_r[i]) / (table_r[i + 1] - table_r[i]);
      return (1 - a) * table_phi[i] + a * table_phi[i + 1];
    }
    default: return 0.0;
  }
}

double GravityModel::dphi_chi(double chi) const {
  switch (kind) {
    case 1: return p[0];
    case 2: return -p[0] / ((p[1] + chi) * (p[1] + chi));
    case 3: {
      const double rhoC = p[0], K = p[1], G = p[2];
      const double alpha = std::sqrt(2.0 * PI * G / K);
      const double chi_eff = alpha * (chi + std::numeric_limits<double>::min());
      const double dphi = (std::cos(chi_eff) - std::sin(chi_eff) / chi_eff) / chi_eff;
      return -2.0 * K * rhoC * dphi * alpha;
    }
    case 4: {
      const size_t i = table_index(chi);
      return (table_phi[i + 1] - table_phi[i]) / (table_r[i + 1] - table_r[i]);
    }
    default: return 0.0;
  }
}

double GravityModel::phi(const double x[3]) const { return phi_chi(coordinate(x)); }

void GravityModel::grad_phi(const double x[3], double g[3]) const {
  const double d = dphi_chi(coordinate(x));
  for (int dir = 0; dir < 3; ++dir) g[dir] = d * dx(x, dir);
}

void tabulate_gravity(const GravityModel &gm, const HostGrid &g, std::vector<double> &phi_cqp,
                      std::vector<double> &gradphi_cqp, std::vector<double> &phi_fqp) {
  const i64 n = g.n_cells, E = g.n_edges;
  phi_cqp.resize((size_t)(n * g.q_c));
  gradphi_cqp.resize((size_t)(n * g.q_c * 3));
  phi_fqp.resize((size_t)(E * g.q_f));
#pragma omp parallel for schedule(static)
  for (i64 a = 0; a < n * g.q_c; ++a) {
    phi_cqp[(size_t)a] = gm.phi(&g.cell_qp[(size_t)(3 * a)]);
    gm.grad_phi(&g.cell_qp[(size_t)(3 * a)], &gradphi_cqp[(size_t)(3 * a)]);
  }
#pragma omp parallel for schedule(static)
  for (i64 a = 0; a < E * g.q_f; ++a) phi_fqp[(size_t)a] = gm.phi(&g.face_qp[(size_t)(3 * a)]);
}

}  // namespace zfvm
