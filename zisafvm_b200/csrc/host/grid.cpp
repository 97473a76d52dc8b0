// Flattened grid construction. Numbering and orientation conventions follow
// src/zisa/grid/grid.cpp (see zfvm_host.hpp for the line map).
#include <algorithm>
#include <array>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <stdexcept>
#if defined(_OPENMP)
#include <omp.h>
#include <parallel/algorithm>
#endif

#include "vec3.hpp"
#include "zfvm_host.hpp"

namespace zfvm {

int relative_vertex_index(int n_dims, int k, int rel) {
  if (n_dims == 2) return (k + rel) % 3;
  static const int table[4][3] = {{0, 1, 3}, {0, 2, 1}, {0, 3, 2}, {1, 2, 3}};
  return table[k][rel];
}

int relative_off_vertex_index(int n_dims, int k) {
  if (n_dims == 2) return (k + 2) % 3;
  static const int table[4] = {2, 3, 1, 0};
  return table[k];
}

namespace {

// grid.cpp:152-203
void enforce_standard_vertex_order(int n_dims, const std::vector<double> &vertices, std::vector<i32> &vi) {
  const int F = n_dims + 1;
  const i64 n_cells = (i64)vi.size() / F;
  auto V = [&](i32 v) { return Vec3{vertices[3 * (i64)v], vertices[3 * (i64)v + 1], vertices[3 * (i64)v + 2]}; };
#pragma omp parallel for schedule(static)
  for (i64 i = 0; i < n_cells; ++i) {
    i32 *c = &vi[i * F];
    if (n_dims == 3) {
      Vec3 cc = 0.25 * (V(c[0]) + V(c[1]) + V(c[2]) + V(c[3]));
      for (int k = 0; k < 4; ++k) {
        Vec3 a = V(c[relative_vertex_index(3, k, 0)]);
        Vec3 b = V(c[relative_vertex_index(3, k, 1)]);
        Vec3 d = V(c[relative_vertex_index(3, k, 2)]);
        Vec3 n = cross(b - a, d - a);
        Vec3 fc = (1.0 / 3.0) * (a + b + d);
        if (dot(n, fc - cc) < 0.0) {
          std::swap(c[2], c[3]);
          break;
        }
      }
    } else {
      Vec3 n = cross(V(c[1]) - V(c[0]), V(c[2]) - V(c[0]));
      if (n.z < 0.0) std::swap(c[1], c[2]);
    }
  }
}

struct FaceKey {
  i32 v[3];
  i32 payload;  // cell * F + k
  bool operator<(const FaceKey &o) const {
    if (v[0] != o.v[0]) return v[0] < o.v[0];
    if (v[1] != o.v[1]) return v[1] < o.v[1];
    if (v[2] != o.v[2]) return v[2] < o.v[2];
    return payload < o.payload;
  }
};

// Same result as grid.cpp:520-540 (common_face over vertex neighbours): two cells are
// neighbours across local faces (ki, kj) iff those faces have identical vertex sets.
void compute_neighbours(int n_dims, const std::vector<i32> &vi, std::vector<i32> &nb) {
  const int F = n_dims + 1;
  const i64 n_cells = (i64)vi.size() / F;
  if (n_cells * F > (i64)2147483647) throw std::runtime_error("grid too large for 32-bit face ids");
  std::vector<FaceKey> keys((size_t)(n_cells * F));
#pragma omp parallel for schedule(static)
  for (i64 i = 0; i < n_cells; ++i) {
    for (int k = 0; k < F; ++k) {
      FaceKey fk;
      fk.v[2] = 0;
      for (int r = 0; r < n_dims; ++r) fk.v[r] = vi[i * F + relative_vertex_index(n_dims, k, r)];
      if (n_dims == 2) {
        if (fk.v[0] > fk.v[1]) std::swap(fk.v[0], fk.v[1]);
      } else {
        std::sort(fk.v, fk.v + 3);
      }
      fk.payload = (i32)(i * F + k);
      keys[(size_t)(i * F + k)] = fk;
    }
  }
#if defined(_OPENMP)
  __gnu_parallel::sort(keys.begin(), keys.end());
#else
  std::sort(keys.begin(), keys.end());
#endif
  nb.assign((size_t)(n_cells * F), INVALID);
  const i64 m = (i64)keys.size();
#pragma omp parallel for schedule(static)
  for (i64 a = 0; a < m - 1; ++a) {
    const FaceKey &x = keys[(size_t)a], &y = keys[(size_t)a + 1];
    if (x.v[0] == y.v[0] && x.v[1] == y.v[1] && x.v[2] == y.v[2]) {
      nb[(size_t)x.payload] = y.payload / F;
      nb[(size_t)y.payload] = x.payload / F;
    }
  }
}

double herons_formula(double a, double b, double c) {
  double s = 0.5 * (a + b + c);
  return std::sqrt(s * (s - a) * (s - b) * (s - c));
}

struct CellGeom {
  double volume, inradius, circum, length;
};

// triangle.cpp:11-42, tetrahedron.cpp:18-66
CellGeom cell_geometry(int n_dims, const Vec3 *v) {
  CellGeom g;
  if (n_dims == 2) {
    double a = norm(v[1] - v[2]), b = norm(v[0] - v[2]), c = norm(v[0] - v[1]);
    g.volume = herons_formula(a, b, c);
    Vec3 bc = (v[0] + v[1] + v[2]) / 3.0;
    g.circum = std::max(std::max(norm(v[0] - bc), norm(v[1] - bc)), norm(v[2] - bc));
    g.inradius = 0.5 * std::sqrt((b + c - a) * (c + a - b) * (a + b - c) / (a + b + c));
    g.length = g.circum;
  } else {
    Vec3 d1 = v[1] - v[0], d2 = v[2] - v[0], d3 = v[3] - v[0];
    double vol = 1.0 / 6.0 *
                 (d1.x * d2.y * d3.z + d2.x * d3.y * d1.z + d3.x * d1.y * d2.z - d1.x * d3.y * d2.z -
                  d2.x * d1.y * d3.z - d3.x * d2.y * d1.z);
    g.volume = std::abs(vol);
    Vec3 bc = 0.25 * (v[0] + v[1] + v[2] + v[3]);
    double l = 0.0, rmax = 0.0, rmin = 0.0;
    for (int k = 0; k < 4; ++k) {
      l += 2.0 * norm(bc - v[k]);
      rmax = (k == 0) ? norm(v[k] - bc) : std::max(rmax, norm(v[k] - bc));
      Vec3 fb = (v[relative_vertex_index(3, k, 0)] + v[relative_vertex_index(3, k, 1)] +
                 v[relative_vertex_index(3, k, 2)]) /
                3.0;
      double r = norm(fb - bc);
      rmin = (k == 0) ? r : std::min(rmin, r);
    }
    g.length = 0.25 * l;
    g.circum = rmax;
    g.inradius = rmin;
  }
  return g;
}

// denormalize (denormalized_rule.hpp:36-51) + barycentric coord (barycentric.cpp:23-25,57-60).
void denormalize(const RefRule &r, const Vec3 *v, double vol, double *points, double *weights) {
  for (int q = 0; q < r.n_points; ++q) {
    const double *lam = &r.bary[(size_t)q * r.n_bary];
    Vec3 x;
    if (r.n_bary == 2) {
      x = lam[0] * v[0] + lam[1] * v[1];
    } else if (r.n_bary == 3) {
      x = v[0] * lam[0] + v[1] * lam[1] + v[2] * lam[2];
    } else {
      x = v[0] * lam[0] + v[1] * lam[1] + v[2] * lam[2] + v[3] * lam[3];
    }
    points[3 * q] = x.x;
    points[3 * q + 1] = x.y;
    points[3 * q + 2] = x.z;
    weights[q] = vol * r.weights[q];
  }
}

// average(qr, x -> x): quadrature.hpp:33-62 accumulation order.
Vec3 rule_barycenter(int n, const double *points, const double *weights, double vol) {
  Vec3 ret = weights[0] * Vec3{points[0], points[1], points[2]};
  for (int q = 1; q < n; ++q) ret = ret + weights[q] * Vec3{points[3 * q], points[3 * q + 1], points[3 * q + 2]};
  ret = 1.0 * ret;
  return ret / vol;
}

}  // namespace

void build_grid(HostGrid &g, int n_dims, std::vector<double> vertices, std::vector<i32> vertex_indices,
                const QRDegrees &deg) {
  if (n_dims != 2 && n_dims != 3) throw std::runtime_error("only works in 2D or 3D");
  const int F = n_dims + 1;
  g = HostGrid();
  g.n_dims = n_dims;
  g.max_neighbours = F;
  g.deg = deg;
  g.n_vertices = (i64)vertices.size() / 3;
  g.n_cells = (i64)vertex_indices.size() / F;
  g.vertices = std::move(vertices);
  g.vertex_indices = std::move(vertex_indices);
  const i64 n = g.n_cells;

  // ZFVM_VERBOSE=1: wall-clock seconds of the phases on stderr
  const bool verbose = std::getenv("ZFVM_VERBOSE") != nullptr;
  auto t_last = std::chrono::steady_clock::now();
  auto lap = [&](const char *what) {
    if (!verbose) return;
    const auto t1 = std::chrono::steady_clock::now();
    std::fprintf(stderr, "[zfvm grid]   %-28s %7.2f s\n", what, std::chrono::duration<double>(t1 - t_last).count());
    t_last = t1;
  };
  enforce_standard_vertex_order(n_dims, g.vertices, g.vertex_indices);
  lap("vertex order");
  compute_neighbours(n_dims, g.vertex_indices, g.neighbours);
  lap("neighbours");

  // edge numbering: interior edges first, in (cell, local face) order of the smaller cell;
  // then boundary edges in (cell, local face) order.  grid.cpp:385-413
  i64 n_int = 0, n_ext = 0;
  for (i64 i = 0; i < n; ++i)
    for (int k = 0; k < F; ++k) {
      i32 j = g.neighbours[i * F + k];
      if (j == INVALID)
        ++n_ext;
      else if (i < j)
        ++n_int;
    }
  g.n_interior_edges = n_int;
  g.n_edges = n_int + n_ext;
  g.edge_indices.assign((size_t)(n * F), INVALID);
  g.left_right.assign((size_t)(2 * g.n_edges), INVALID);
  {
    i64 ci = 0, ce = n_int;
    for (i64 i = 0; i < n; ++i)
      for (int k = 0; k < F; ++k) {
        i32 j = g.neighbours[i * F + k];
        if (j == INVALID) {
          g.edge_indices[i * F + k] = (i32)ce;
          g.left_right[2 * ce] = (i32)i;
          ++ce;
        } else if (i < j) {
          g.edge_indices[i * F + k] = (i32)ci;
          g.left_right[2 * ci] = (i32)i;
          g.left_right[2 * ci + 1] = j;
          ++ci;
        } else {
          int kj = -1;
          for (int l = 0; l < F; ++l)
            if (g.neighbours[(i64)j * F + l] == (i32)i) kj = l;
          if (kj < 0) throw std::runtime_error("failed to find myself");
          g.edge_indices[i * F + k] = g.edge_indices[(i64)j * F + kj];
        }
      }
  }

  lap("edge numbering");
  // per-cell geometry and quadrature
  g.cell_rule = (n_dims == 2) ? make_triangular_rule(deg.volume_deg) : make_tetrahedral_rule(deg.volume_deg);
  g.face_rule = (n_dims == 2) ? make_edge_rule(deg.face_deg) : make_triangular_rule(deg.face_deg);
  RefRule mom_rule = (n_dims == 2) ? make_triangular_rule(deg.moments_deg) : make_tetrahedral_rule(deg.moments_deg);
  g.q_c = g.cell_rule.n_points;
  g.q_f = g.face_rule.n_points;
  g.n_moments = poly_dof(deg.moments_deg, n_dims);
  if (deg.moments_deg > 8 || mom_rule.n_points > 16) throw std::runtime_error("moment rule beyond the compiled bounds (degree 8, 16 points)");

  g.volumes.resize((size_t)n);
  g.inradii.resize((size_t)n);
  g.circum_radii.resize((size_t)n);
  g.characteristic_length.resize((size_t)n);
  g.cell_centers.resize((size_t)(3 * n));
  g.cell_qp.resize((size_t)(n * g.q_c * 3));
  g.cell_qw.resize((size_t)(n * g.q_c));
  parallel_assign(g.moments, (size_t)(n * g.n_moments), 0.0);
  g.cell_flags.assign((size_t)n, FLAG_INTERIOR);  // CellFlags() default: interior (cell_flags.hpp)

#pragma omp parallel for schedule(static)
  for (i64 i = 0; i < n; ++i) {
    Vec3 v[4];
    for (int k = 0; k < F; ++k) v[k] = g.vertex(i, k);
    CellGeom cg = cell_geometry(n_dims, v);
    g.volumes[i] = cg.volume;
    g.inradii[i] = cg.inradius;
    g.circum_radii[i] = cg.circum;
    g.characteristic_length[i] = cg.length;
    double *qp = &g.cell_qp[(size_t)(i * g.q_c * 3)];
    double *qw = &g.cell_qw[(size_t)(i * g.q_c)];
    denormalize(g.cell_rule, v, cg.volume, qp, qw);
    Vec3 c = rule_barycenter(g.q_c, qp, qw, cg.volume);
    g.cell_centers[3 * i] = c.x;
    g.cell_centers[3 * i + 1] = c.y;
    g.cell_centers[3 * i + 2] = c.z;

    // normalized moments, grid.cpp:1049-1098 with cell.cpp:13-26
    double mp[3 * 16], mw[16];
    denormalize(mom_rule, v, cg.volume, mp, mw);
    Vec3 mc = rule_barycenter(mom_rule.n_points, mp, mw, cg.volume);
    // powers of the centred coordinates, formed once per point (the same std::pow calls cell.cpp:17-27 makes for every
    // monomial: 3 (deg + 1) of them per point instead of 3 per point and monomial)
    constexpr int MAX_MOM_DEG = 8;
    double pw[16][3][MAX_MOM_DEG + 1];
    for (int q = 0; q < mom_rule.n_points; ++q) {
      const double xyz[3] = {mp[3 * q] - mc.x, mp[3 * q + 1] - mc.y, mp[3 * q + 2] - mc.z};
      for (int d = 0; d < 3; ++d)
        for (int e = 0; e <= deg.moments_deg; ++e) pw[q][d][e] = std::pow(xyz[d], (double)e);
    }
    auto avg_moment = [&](int a, int b, int c2) {
      auto f = [&](int q) { return pw[q][0][a] * pw[q][1][b] * pw[q][2][c2]; };
      double ret = mw[0] * f(0);
      for (int q = 1; q < mom_rule.n_points; ++q) ret = ret + mw[q] * f(q);
      ret = 1.0 * ret;
      return ret / cg.volume;
    };
    double *m = &g.moments[(size_t)(i * g.n_moments)];
    double length_d = 1.0;
    for (int d = 0; d <= deg.moments_deg; ++d) {
      for (int a = 0; a <= d; ++a) {
        if (n_dims == 2) {
          int b = d - a;
          m[poly_index2(a, b)] = avg_moment(a, b, 0) / length_d;
        } else {
          for (int b = 0; b <= d - a; ++b) {
            int c2 = d - a - b;
            m[poly_index3(a, b, c2)] = avg_moment(a, b, c2) / length_d;
          }
        }
      }
      length_d *= cg.length;
    }
  }

  lap("cells, quadrature, moments");
  // faces: geometry defined by the left cell's local face (grid.cpp:583-649, face_factory.cpp)
  const i64 E = g.n_edges;
  g.face_qp.resize((size_t)(E * g.q_f * 3));
  g.face_qw.resize((size_t)(E * g.q_f));
  g.face_area.resize((size_t)E);
  g.face_normal.resize((size_t)(3 * E));
  g.face_t1.resize((size_t)(3 * E));
  g.face_t2.resize((size_t)(3 * E));
  g.face_centers.resize((size_t)(3 * E));
  g.face_vertex_slots.assign((size_t)(n * F), 0);

#pragma omp parallel for schedule(static)
  for (i64 i = 0; i < n; ++i) {
    for (int k = 0; k < F; ++k) {
      i32 j = g.neighbours[i * F + k];
      i64 e = g.edge_indices[i * F + k];
      if (j == INVALID || i < j) {
        Vec3 fv[3];
        std::uint8_t slots = 0;
        for (int r = 0; r < n_dims; ++r) {
          int s = relative_vertex_index(n_dims, k, r);
          fv[r] = g.vertex(i, s);
          slots |= (std::uint8_t)(s << (2 * r));
        }
        g.face_vertex_slots[i * F + k] = slots;
        Vec3 nrm, t1, t2;
        double area;
        if (n_dims == 2) {
          Vec3 t = normalize(fv[1] - fv[0]);
          nrm = Vec3{t.y, -t.x, 0.0};  // rotate_right
          t1 = t;
          area = norm(fv[1] - fv[0]);
        } else {
          nrm = normalize(cross(fv[1] - fv[0], fv[2] - fv[0]));
          t1 = normalize(fv[1] - fv[0]);
          double a = norm(fv[1] - fv[2]), b = norm(fv[0] - fv[2]), c = norm(fv[0] - fv[1]);
          area = herons_formula(a, b, c);
        }
        t2 = cross(nrm, t1);
        double *qp = &g.face_qp[(size_t)(e * g.q_f * 3)];
        double *qw = &g.face_qw[(size_t)(e * g.q_f)];
        denormalize(g.face_rule, fv, area, qp, qw);
        Vec3 fc = rule_barycenter(g.q_f, qp, qw, area);
        g.face_area[e] = area;
        for (int d = 0; d < 3; ++d) {
          g.face_normal[3 * e + d] = nrm[d];
          g.face_t1[3 * e + d] = t1[d];
          g.face_t2[3 * e + d] = t2[d];
          g.face_centers[3 * e + d] = fc[d];
        }
      } else {
        // right cell: find the positions of the left cell's face vertices in my vertex list
        int kj = -1;
        for (int l = 0; l < F; ++l)
          if (g.neighbours[(i64)j * F + l] == (i32)i) kj = l;
        std::uint8_t slots = 0;
        for (int r = 0; r < n_dims; ++r) {
          i32 gv = g.vertex_indices[(i64)j * F + relative_vertex_index(n_dims, kj, r)];
          int s = -1;
          for (int l = 0; l < F; ++l)
            if (g.vertex_indices[i * F + l] == gv) s = l;
          slots |= (std::uint8_t)(s << (2 * r));
        }
        g.face_vertex_slots[i * F + k] = slots;
      }
    }
  }
  lap("faces");
}

void mask_ghost_cells(HostGrid &g, const std::uint8_t *mask) {
  const int F = g.max_neighbours;
  for (i64 i = 0; i < g.n_cells; ++i) {
    if (!mask[i]) continue;
    std::uint8_t f = g.cell_flags[i];
    f &= (std::uint8_t)~FLAG_INTERIOR;
    f |= FLAG_GHOST;
    for (int k = 0; k < F; ++k) {
      i32 j = g.neighbours[i * F + k];
      if (j != INVALID && !mask[j]) f |= FLAG_GHOST_L1;
    }
    g.cell_flags[i] = f;
  }
}

}  // namespace zfvm
