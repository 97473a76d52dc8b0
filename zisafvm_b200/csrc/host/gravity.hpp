// Host gravity models (see gravity.cpp).
#pragma once
#include <vector>

#include "zfvm_host.hpp"

namespace zfvm {

struct GravityModel {
  int kind = 0;       // layout.hpp GravityKind
  int alignment = 0;  // 0 radial, 1 axial
  double p[4] = {0, 0, 0, 0};
  double axis[3] = {0, 0, 0};
  std::vector<double> table_r, table_phi;

  double coordinate(const double x[3]) const;
  double dx(const double x[3], int dir) const;
  size_t table_index(double r) const;
  double phi_chi(double chi) const;
  double dphi_chi(double chi) const;
  double phi(const double x[3]) const;
  void grad_phi(const double x[3], double g[3]) const;
};

void tabulate_gravity(const GravityModel &gm, const HostGrid &g, std::vector<double> &phi_cqp,
                      std::vector<double> &gradphi_cqp, std::vector<double> &phi_fqp);

}  // namespace zfvm
