// Reference quadrature rules on the unit edge / triangle / tetrahedron.
#include <cmath>
#include <stdexcept>

#include "zfvm_host.hpp"

namespace zfvm {

int poly_dof(int deg, int n_dims) {
  if (n_dims == 2) return ((deg + 1) * (deg + 2)) / 2;
  return ((deg + 1) * (deg + 2) * (deg + 3)) / 6;
}
int poly_index2(int a, int b) {
  int n = a + b;
  return ((n + 1) * n) / 2 + b;
}
int poly_index3(int a, int b, int c) { return poly_dof(a + b + c - 1, 3) + poly_index2(b, c); }

int required_stencil_size(int deg, double factor, int n_dims) {
  if (deg == 0) return 1;
  return (int)(double(poly_dof(deg, n_dims) - 1) * factor + 1);
}
int deduce_max_order(int stencil_size, double factor, int n_dims) {
  int deg = 0;
  while (required_stencil_size(deg + 1, factor, n_dims) <= stencil_size) deg += 1;
  return deg + 1;
}

// Gauss-Legendre nodes as the roots of the Fourier form of P_n, found by Newton's method
// from nearly equi-spaced guesses in theta (gauss_legendre.hpp:39-187, newton.hpp:15-33).
void gauss_legendre(int n, double *points, double *weights) {
  const double pi = 3.14159265358979323846;
  std::vector<double> a(n + 1, 0.0);
  {
    double ann = std::sqrt(2.0);
    for (int i = 1; i <= n; ++i) ann = std::sqrt(1.0 - 1.0 / (4.0 * double(i) * double(i))) * ann;
    a[n] = ann;
    for (int l = 2; l <= n; l += 2) {
      double factor = double((l - 1) * (2 * n - l + 2)) / double(l * (2 * n - l + 1));
      a[n - l] = factor * a[n - l + 2];
    }
    if (n % 2 == 0) a[0] *= 0.5;
  }
  auto p = [&](double th) {
    double fx = a[0];
    for (int i = 1; i <= n; ++i) fx += a[i] * std::cos(i * th);
    return fx;
  };
  auto dp = [&](double th) {
    double fx = 0.0;
    for (int i = 1; i <= n; ++i) fx += -a[i] * i * std::sin(i * th);
    return fx;
  };
  auto newton = [&](double x) {
    double fx = p(x), dfx = dp(x);
    int iter = 0;
    while (std::abs(fx / dfx) >= 1e-12 && iter < 100) {
      x = x - fx / dfx;
      fx = p(x);
      dfx = dp(x);
      ++iter;
    }
    return x;
  };
  const bool even = (n % 2 == 0);
  std::vector<double> th(n, 0.0);
  if (!even) {
    th[n / 2] = pi / 2;
    for (int k = 1; k <= n / 2; ++k) {
      double guess = (k < 2 ? pi / 2 - k * pi / n : 2.0 * th[n / 2 + k - 1] - th[n / 2 + k - 2]);
      th[n / 2 + k] = newton(guess);
    }
    weights[n / 2] = (2 * n + 1) / (dp(th[n / 2]) * dp(th[n / 2]));
  } else {
    for (int k = 0; k < n / 2; ++k) {
      double guess = (k < 2 ? pi / 2 - (k + 0.5) * pi / n : 2.0 * th[n / 2 + k - 1] - th[n / 2 + k - 2]);
      th[n / 2 + k] = newton(guess);
    }
  }
  for (int k = 0; k < n / 2; ++k) {
    int iu = even ? n / 2 + k : n / 2 + k + 1;
    int il = n / 2 - 1 - k;
    double d = dp(th[iu]);
    weights[iu] = (2 * n + 1) / (d * d);
    weights[il] = weights[iu];
  }
  if (!even) points[n / 2] = std::cos(th[n / 2]);
  for (int k = 0; k < n / 2; ++k) {
    int iu = even ? n / 2 + k : n / 2 + k + 1;
    int il = n / 2 - 1 - k;
    points[iu] = std::cos(th[iu]);
    points[il] = -points[iu];
  }
}

RefRule make_edge_rule(int deg) {
  int n = deg / 2 + 1;
  if (n > 4) throw std::runtime_error("edge rule: implement case");
  RefRule r;
  r.n_points = n;
  r.n_bary = 2;
  r.xi.resize(n);
  r.weights.resize(n);
  std::vector<double> w(n);
  gauss_legendre(n, r.xi.data(), w.data());
  r.bary.resize(2 * n);
  for (int k = 0; k < n; ++k) {
    r.weights[k] = 0.5 * w[k];
    r.bary[2 * k + 0] = 0.5 - 0.5 * r.xi[k];  // edge.hpp:26-30
    r.bary[2 * k + 1] = 0.5 + 0.5 * r.xi[k];
  }
  return r;
}

static void permutate(RefRule &r, double w, std::initializer_list<double> lam_) {
  std::vector<double> lam(lam_);
  auto push = [&](double a, double b, double c) {
    r.weights.push_back(w);
    r.bary.push_back(a);
    r.bary.push_back(b);
    r.bary.push_back(c);
    r.n_points += 1;
  };
  if (lam.size() == 1) {
    push(lam[0], lam[0], lam[0]);
  } else if (lam.size() == 2) {
    push(lam[0], lam[1], lam[1]);
    push(lam[1], lam[0], lam[1]);
    push(lam[1], lam[1], lam[0]);
  } else {
    throw std::runtime_error("permutate: implement case");
  }
}

RefRule make_triangular_rule(int deg) {
  RefRule r;
  r.n_bary = 3;
  if (deg <= 0) deg = 1;
  if (deg == 1) {
    permutate(r, 1.0, {1.0 / 3.0});
  } else if (deg == 2) {
    permutate(r, 1.0 / 3.0, {2.0 / 3.0, 1.0 / 6.0});
  } else if (deg == 3) {
    permutate(r, -0.5625, {1.0 / 3.0});
    permutate(r, 1.5625 / 3.0, {0.6, 0.2});
  } else if (deg == 4) {
    permutate(r, 0.109951743655322, {0.816847572980459, 0.091576213509771});
    permutate(r, 0.223381589678011, {0.108103018168070, 0.445948490915965});
  } else if (deg == 5) {
    permutate(r, 0.225, {1.0 / 3.0});
    permutate(r, 0.132394152788506, {0.059715871789770, 0.470142064105115});
    permutate(r, 0.125939180544827, {0.797426985353087, 0.101286507323456});
  } else {
    throw std::runtime_error("triangular rule: implement the missing case");
  }
  return r;
}

RefRule make_tetrahedral_rule(int deg) {
  RefRule r;
  r.n_bary = 4;
  auto push = [&](double w, double a, double b, double c, double d) {
    r.weights.push_back(w);
    r.bary.push_back(a);
    r.bary.push_back(b);
    r.bary.push_back(c);
    r.bary.push_back(d);
    r.n_points += 1;
  };
  if (deg <= 1) {
    push(1.0, 0.25, 0.25, 0.25, 0.25);
  } else if (deg == 2) {
    const double a = 0.5854101966249680, b = 0.1381966011250110;
    push(0.25, a, b, b, b);
    push(0.25, b, a, b, b);
    push(0.25, b, b, a, b);
    push(0.25, b, b, b, a);
  } else if (deg == 3) {
    const double w1 = 0.0476331348432089, a = 0.7784952948213300, b = 0.0738349017262234;
    push(w1, a, b, b, b);
    push(w1, b, a, b, b);
    push(w1, b, b, a, b);
    push(w1, b, b, b, a);
    const double w2 = 0.1349112434378610, c = 0.4062443438840510, d = 0.0937556561159491;
    push(w2, c, c, d, d);
    push(w2, c, d, c, d);
    push(w2, c, d, d, c);
    push(w2, d, c, c, d);
    push(w2, d, c, d, c);
    push(w2, d, d, c, c);
  } else {
    throw std::runtime_error("tetrahedral rules of degree 4 and higher are not implemented");
  }
  return r;
}

}  // namespace zfvm
