// Pieces of the stencil selection (src/zisa/reconstruction/stencil.cpp:178-393, src/zisa/math/cone.cpp:7-30) that the
// host search (stencil.cpp) and the device search (kernels/stencil_search.cu) evaluate from one source.  Everything a
// decision depends on -- cone membership of a point, distances, singular values -- is plain IEEE double arithmetic
// without fused multiply-adds on both sides (g++ has no FMA target here; the .cu file is compiled with -fmad=false),
// so both searches take the same branches.
#pragma once

#include <cmath>

#include "lsq_shared.hpp"

namespace zfvm {
namespace sel {

struct V3 {
  double x, y, z;
};
ZFVM_HD V3 sub(V3 a, V3 b) { return V3{a.x - b.x, a.y - b.y, a.z - b.z}; }
ZFVM_HD V3 cross(V3 a, V3 b) { return V3{a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
ZFVM_HD double dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
ZFVM_HD double det3(V3 a, V3 b, V3 c) { return dot(a, cross(b, c)); }
ZFVM_HD double norm(V3 a) { return sqrt(dot(a, a)); }

/// Raw view of the grid arrays the search reads (HostGrid on the host, device copies in the kernel).
struct GridView {
  int nd, F;
  long long n_cells;
  const int *nb;          // [n][F] neighbours, -1 on the boundary
  const int *vi;          // [n][F] vertex indices
  const double *vtx;      // [nv][3]
  const double *cc;       // [n][3] cell centres
  const double *len;      // [n] characteristic lengths
  const double *mom;      // [n][n_mom] normalised moments
  int n_mom;
  const double *face_c;   // [n_edges][3] face centres
  const int *edge;        // [n][F] edge indices
  int nq;                 // query points: the degree-5 triangle / degree-3 tetrahedron rule (stencil.cpp:178-190)
  double qbary[10][4];
};
ZFVM_HD V3 vertex(const GridView &g, long long i, int k) {
  const double *p = g.vtx + 3 * (long long)g.vi[i * g.F + k];
  return V3{p[0], p[1], p[2]};
}
ZFVM_HD V3 center(const GridView &g, long long i) { return V3{g.cc[3 * i], g.cc[3 * i + 1], g.cc[3 * i + 2]}; }

/// gmsh_reader.cpp:22-82
ZFVM_HD int rel_vertex(int nd, int k, int rel) {
  if (nd == 2) return (k + rel) % 3;
  const int t[4][3] = {{0, 1, 3}, {0, 2, 1}, {0, 3, 2}, {1, 2, 3}};
  return t[k][rel];
}
ZFVM_HD int rel_off_vertex(int nd, int k) {
  if (nd == 2) return (k + 2) % 3;
  const int t[4] = {2, 3, 1, 0};
  return t[k];
}

/// TriangularCone / TetrahedralCone (cone.cpp) with the inward normals of its faces for the margin tests below.
struct Cone {
  int kind;  // 0 full sphere, 2 triangular cone, 3 tetrahedral cone
  V3 A, dB, dC, dD;
  V3 nrm[3];
  int n_planes;
};
ZFVM_HD Cone full_sphere() {
  Cone c{};
  c.kind = 0;
  c.n_planes = 0;
  return c;
}
/// make_cone, stencil.cpp:258-278
ZFVM_HD Cone make_cone(const GridView &g, long long i, V3 apex, int k) {
  Cone r{};
  r.A = apex;
  r.dB = sub(vertex(g, i, rel_vertex(g.nd, k, 0)), apex);
  r.dC = sub(vertex(g, i, rel_vertex(g.nd, k, 1)), apex);
  if (g.nd == 2) {
    r.kind = 2;
    r.nrm[0] = V3{-r.dB.y, r.dB.x, 0.0};
    r.nrm[1] = V3{r.dC.y, -r.dC.x, 0.0};
    r.n_planes = 2;
  } else {
    r.kind = 3;
    r.dD = sub(vertex(g, i, rel_vertex(g.nd, k, 2)), apex);
    r.nrm[0] = cross(r.dB, r.dC);
    r.nrm[1] = cross(r.dC, r.dD);
    r.nrm[2] = cross(r.dD, r.dB);
    r.n_planes = 3;
  }
  return r;
}
ZFVM_HD bool is_inside(const Cone &c, V3 x) {
  if (c.kind == 0) return true;
  const V3 dx = sub(x, c.A);
  if (c.kind == 2) return cross(c.dB, dx).z >= 0.0 && cross(dx, c.dC).z >= 0.0;  // cone.cpp:11-14
  return det3(c.dB, c.dC, dx) >= 0.0 && det3(c.dC, c.dD, dx) >= 0.0 && det3(c.dD, c.dB, dx) >= 0.0;  // :23-32
}

/// Any query point or the centre of cell `cand` inside the region (stencil.cpp:200-209; the tests are independent, the
/// cheap one first).  The cone's faces are planes through its apex, n_p . (x - A) >= 0, and a query point is a convex
/// combination of the cell's vertices (positive barycentric coordinates): its plane values are the same combination of
/// the vertices' values.  Deciding the points from those 3 x F numbers is exact whenever a value clears zero by a margin
/// nine orders of magnitude above the round-off of either evaluation; otherwise the cell takes the reference's
/// point-by-point test.  (Most tested cells touch the cone's boundary: this is where the search spends its time.)
ZFVM_HD bool cell_inside(const GridView &g, const Cone &region, long long cand) {
  if (region.kind == 0) return true;
  if (is_inside(region, center(g, cand))) return true;
  const int F = g.F;
  V3 v[4];
  for (int k = 0; k < F; ++k) v[k] = vertex(g, cand, k);
  double h[3][4], mag[3][4];
  const int np = region.n_planes;
  for (int p = 0; p < np; ++p) {
    bool all_behind = true;
    for (int k = 0; k < F; ++k) {
      const V3 dx = sub(v[k], region.A);
      const double tx = region.nrm[p].x * dx.x, ty = region.nrm[p].y * dx.y, tz = region.nrm[p].z * dx.z;
      h[p][k] = tx + ty + tz;
      mag[p][k] = fabs(tx) + fabs(ty) + fabs(tz);
      all_behind = all_behind && h[p][k] < -1e-7 * mag[p][k];
    }
    if (all_behind) return false;  // the whole cell lies behind one face
  }
  bool ambiguous = false;
  for (int q = 0; q < g.nq && !ambiguous; ++q) {
    const double *lam = g.qbary[q];
    bool inside = true;
    for (int p = 0; p < np; ++p) {
      double val = 0.0, m = 0.0;
      for (int k = 0; k < F; ++k) {
        val += lam[k] * h[p][k];
        m += lam[k] * mag[p][k];
      }
      if (val < -1e-7 * m) {
        inside = false;
        break;
      }
      if (!(val > 1e-7 * m)) {
        ambiguous = true;
        break;
      }
    }
    if (ambiguous) break;
    if (inside) return true;
  }
  if (!ambiguous) return false;
  for (int q = 0; q < g.nq; ++q) {
    const double *lam = g.qbary[q];
    V3 x{v[0].x * lam[0] + v[1].x * lam[1] + v[2].x * lam[2], v[0].y * lam[0] + v[1].y * lam[1] + v[2].y * lam[2],
         v[0].z * lam[0] + v[1].z * lam[1] + v[2].z * lam[2]};
    if (F == 4) x = V3{x.x + v[3].x * lam[3], x.y + v[3].y * lam[3], x.z + v[3].z * lam[3]};
    if (is_inside(region, x)) return true;
  }
  return false;
}

/// Rank of the rows x cols matrix in U (row-major, destroyed): singular values by one-sided (Hestenes) Jacobi rotations,
/// Eigen's JacobiSVD::rank() threshold  sigma_i > sigma_max * min(rows, cols) * eps  (stencil.cpp:352).
ZFVM_HD int matrix_rank_inplace(double *U, int rows, int cols) {
  if (rows < cols) return rows < 0 ? 0 : (rows < cols - 1 ? rows : cols - 1);  // under-determined: never full column rank
  for (int sweep = 0; sweep < 60; ++sweep) {
    bool rotated = false;
    for (int p = 0; p < cols - 1; ++p)
      for (int q = p + 1; q < cols; ++q) {
        double app = 0.0, aqq = 0.0, apq = 0.0;
        for (int r = 0; r < rows; ++r) app += U[r * cols + p] * U[r * cols + p];
        for (int r = 0; r < rows; ++r) aqq += U[r * cols + q] * U[r * cols + q];
        for (int r = 0; r < rows; ++r) apq += U[r * cols + p] * U[r * cols + q];
        if (fabs(apq) <= 1e-15 * sqrt(app * aqq) || apq == 0.0) continue;
        rotated = true;
        const double zeta = (aqq - app) / (2.0 * apq);
        const double t = (zeta >= 0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
        const double c = 1.0 / sqrt(1.0 + t * t), s = c * t;
        for (int r = 0; r < rows; ++r) {
          const double up = U[r * cols + p], uq = U[r * cols + q];
          U[r * cols + p] = c * up - s * uq;
          U[r * cols + q] = s * up + c * uq;
        }
      }
    if (!rotated) break;
  }
  double smax = 0.0;
  for (int p = 0; p < cols; ++p) {
    double d = 0.0;
    for (int r = 0; r < rows; ++r) d += U[r * cols + p] * U[r * cols + p];
    const double sv = sqrt(d);
    if (sv > smax) smax = sv;
  }
  if (smax == 0.0) return 0;
  const double thresh = smax * (rows < cols ? rows : cols) * 2.220446049250313e-16;
  int rank = 0;
  for (int p = 0; p < cols; ++p) {
    double d = 0.0;
    for (int r = 0; r < rows; ++r) d += U[r * cols + p] * U[r * cols + p];
    if (sqrt(d) > thresh) ++rank;
  }
  return rank;
}

/// assemble_weno_ao_matrix (lsq_solver.cpp:168-403) of the stencil s[0 .. n) into A (row-major (n - 1) x cols).
ZFVM_HD void assemble_matrix(const GridView &g, const int *s, int n, int order, double *A, int cols) {
  const long long i0 = s[0];
  const V3 x0 = center(g, i0);
  const double l0 = g.len[i0];
  for (int ii = 0; ii < n - 1; ++ii) {
    const long long j = s[ii + 1];
    const V3 c = center(g, j);
    double *row = A + ii * cols;
    for (int a = 0; a < cols; ++a) row[a] = 0.0;
    lsq::lsq_row(row, g.nd, order, (c.x - x0.x) / l0, (c.y - x0.y) / l0, (c.z - x0.z) / l0, g.len[j] / l0,
                 g.mom + i0 * g.n_mom, g.mom + j * g.n_mom);
  }
}

}  // namespace sel
}  // namespace zfvm
