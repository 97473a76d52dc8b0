// Stencil selection and least-squares matrices.
//   region-grown candidates, distance sort, truncation   stencil.cpp:192-258
//   biased stencils: cone at off-vertex, cone at centre, random retry  stencil.cpp:260-393
//   local/global index bookkeeping                        stencil.cpp:82-104
//   family: k_biased counter, k_high, achieved order      stencil_family.cpp:15-45,120-136
//   o1 families for cells that are neither interior nor ghost_l1  stencil_family.cpp:99-117
//   A assembly                                            lsq_solver.cpp:168-403
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <stdexcept>

#include "lsq_shared.hpp"
#include "stencil_shared.hpp"
#include "vec3.hpp"
#include "zfvm_host.hpp"

namespace zfvm {

namespace {

using Region = sel::Cone;  // full sphere | triangular cone | tetrahedral cone (stencil_shared.hpp)

}  // namespace

sel::GridView make_grid_view(const HostGrid &g) {
  sel::GridView v{};
  v.nd = g.n_dims;
  v.F = g.max_neighbours;
  v.n_cells = g.n_cells;
  v.nb = g.neighbours.data();
  v.vi = g.vertex_indices.data();
  v.vtx = g.vertices.data();
  v.cc = g.cell_centers.data();
  v.len = g.characteristic_length.data();
  v.mom = g.moments.data();
  v.n_mom = g.n_moments;
  v.face_c = g.face_centers.data();
  v.edge = g.edge_indices.data();
  // MAX_*_RULE_DEGREE rule, stencil.cpp:178-190
  const RefRule q = g.n_dims == 2 ? make_triangular_rule(MAX_TRIANGULAR_RULE_DEGREE) : make_tetrahedral_rule(MAX_TETRAHEDRAL_RULE_DEGREE);
  if (q.n_points > 10) throw std::runtime_error("query rule too large");
  v.nq = q.n_points;
  for (int a = 0; a < q.n_points; ++a)
    for (int b = 0; b < v.F; ++b) v.qbary[a][b] = q.bary[(size_t)a * v.F + b];
  return v;
}

namespace {

struct Selector {
  const HostGrid &g;
  sel::GridView gv;
  explicit Selector(const HostGrid &g_) : g(g_), gv(make_grid_view(g_)) {}

  bool cell_inside(const Region &region, i32 cand) const { return sel::cell_inside(gv, region, cand); }

  void candidates(std::vector<i32> &cands, i32 i_center, int n_points, const Region &region) const {
    const int F = g.max_neighbours;
    const size_t max_points = (size_t)5 * n_points;
    cands.clear();
    cands.push_back(i_center);
    // "seen" = already a candidate, or already found outside the region (the membership test is pure: it is not repeated
    // for a cell met again as the neighbour of a later candidate).  Per-thread stamp array instead of linear searches.
    static thread_local std::vector<std::uint32_t> stamp;
    static thread_local std::uint32_t epoch = 0;
    if (stamp.size() != (size_t)g.n_cells) {
      stamp.assign((size_t)g.n_cells, 0u);
      epoch = 0;
    }
    if (++epoch == 0) {  // wrapped around
      std::fill(stamp.begin(), stamp.end(), 0u);
      epoch = 1;
    }
    stamp[(size_t)i_center] = epoch;
    for (size_t p = 0; p < max_points; ++p) {
      if (p >= cands.size()) break;
      i32 j = cands[p];
      for (int k = 0; k < F; ++k) {
        i32 cand = g.neighbours[(i64)j * F + k];
        if (cand == INVALID) continue;
        if (stamp[(size_t)cand] == epoch) continue;
        stamp[(size_t)cand] = epoch;
        if (cell_inside(region, cand)) cands.push_back(cand);
      }
    }
  }

  void region_stencil(std::vector<i32> &cands, i32 i_center, int n_points, const Region &region) const {
    candidates(cands, i_center, n_points, region);
    Vec3 xc = g.center(i_center);
    // sorted by distance of the centres (stencil.cpp:240-250); the distances are formed once per candidate, the
    // comparisons (and so the permutation std::sort produces) are the ones of the comparator on the cell indices
    struct Keyed {
      double d;
      i32 c;
    };
    Keyed keyed[512];
    const size_t m = cands.size();
    if (m > 512) {
      auto closer = [&](i32 a, i32 b) { return norm(g.center(a) - xc) < norm(g.center(b) - xc); };
      std::sort(cands.begin(), cands.end(), closer);
    } else {
      for (size_t a = 0; a < m; ++a) keyed[a] = Keyed{norm(g.center(cands[a]) - xc), cands[a]};
      auto less = [](const Keyed &a, const Keyed &b) { return a.d < b.d; };
      const size_t np = (size_t)n_points;
      bool done = false;
      if (m > 2 * np) {
        // Only the n_points closest cells are kept.  When their distances are pairwise distinct and smaller than all the
        // others, every sort yields the same first n_points entries: select them, sort them, check; ties fall back to
        // the full std::sort on the untouched array, whose permutation is the reference's.
        Keyed part[512];
        std::copy(keyed, keyed + m, part);
        std::nth_element(part, part + np, part + m, less);
        std::sort(part, part + np, less);
        double rest_min = part[np].d;
        for (size_t a = np + 1; a < m; ++a) rest_min = std::min(rest_min, part[a].d);
        bool distinct = part[np - 1].d < rest_min;
        for (size_t a = 1; a < np && distinct; ++a) distinct = part[a - 1].d < part[a].d;
        if (distinct) {
          for (size_t a = 0; a < np; ++a) cands[a] = part[a].c;
          done = true;
        }
      }
      if (!done) {
        std::sort(keyed, keyed + m, less);
        for (size_t a = 0; a < m; ++a) cands[a] = keyed[a].c;
      }
    }
    if ((int)cands.size() > n_points) cands.resize((size_t)n_points);
  }

  Region make_cone(i32 i, Vec3 apex, int k) const { return sel::make_cone(gv, i, sel::V3{apex.x, apex.y, apex.z}, k); }

  bool is_good(const i32 *s, int n, int order) const {
    std::vector<double> A;
    int rows, cols;
    assemble_weno_ao_matrix(A, rows, cols, g, s, n, order);
    return matrix_rank(A.data(), rows, cols) == cols;
  }

  // stencil.cpp:355-393; returns false on the reference's LOG_ERR paths.
  bool biased(std::vector<i32> &s, i32 i, int k, int n_points, int order, std::mt19937 &rng,
              std::string &err) const {
    Region r1 = make_cone(i, g.vertex(i, relative_off_vertex_index(g.n_dims, k)), k);
    region_stencil(s, i, n_points, r1);
    if ((int)s.size() == n_points && is_good(s.data(), n_points, order)) return true;

    Region r2 = make_cone(i, g.center(i), k);
    region_stencil(s, i, n_points, r2);
    if ((int)s.size() == n_points && is_good(s.data(), n_points, order)) return true;

    i64 e = g.edge_indices[(i64)i * g.max_neighbours + k];
    Vec3 fc{g.face_centers[3 * e], g.face_centers[3 * e + 1], g.face_centers[3 * e + 2]};
    Region r3 = make_cone(i, fc, k);
    candidates(s, i, n_points, r3);
    if ((int)s.size() < n_points) {
      s.assign(1, i);
      return true;
    }
    if (!is_good(s.data(), (int)s.size(), order)) {
      err = "It is impossible to find a suitable stencil. i_center = " + std::to_string(i);
      return false;
    }
    // Deviation: the reference re-seeds from std::random_device each iteration
    // (stencil.cpp:332-334); we use a deterministic per-cell generator.
    for (int iter = 0; iter < 100; ++iter) {
      std::shuffle(s.begin() + 1, s.end(), rng);
      if (is_good(s.data(), n_points, order)) {
        s.resize((size_t)n_points);
        return true;
      }
    }
    err = "Did not find a suitable stencil given the candidates. i_center = " + std::to_string(i);
    return false;
  }
};

}  // namespace

// ---- LSQ matrix (lsq_solver.cpp:168-403) ------------------------------------------------------
void assemble_weno_ao_matrix(std::vector<double> &A, int &n_rows, int &n_cols, const HostGrid &g,
                             const i32 *stencil, int n, int order) {
  const int nd = g.n_dims;
  if (order <= 0) throw std::runtime_error("a non-positive convergence order?");
  if (order == 1) {
    n_rows = 1;
    n_cols = 1;
    A.assign(1, 1.0);
    return;
  }
  if ((nd == 2 && order >= 6) || (nd == 3 && order >= 5)) throw std::runtime_error("LSQ order not implemented");
  n_rows = n - 1;
  n_cols = poly_dof(order - 1, nd) - 1;
  A.assign((size_t)n_rows * n_cols, 0.0);

  const i32 i0 = stencil[0];
  const Vec3 x0 = g.center(i0);
  const double l0 = g.characteristic_length[i0];
  const double *C0 = &g.moments[(size_t)i0 * g.n_moments];
  for (int ii = 0; ii < n_rows; ++ii) {
    const i32 j = stencil[ii + 1];
    const Vec3 d = (g.center(j) - x0) / l0;
    lsq::lsq_row(&A[(size_t)ii * n_cols], nd, order, d.x, d.y, d.z, g.characteristic_length[j] / l0, C0,
                 &g.moments[(size_t)j * g.n_moments]);
  }
}

// ---- tiny dense linear algebra -------------------------------------------------------------------
// Singular values by one-sided (Hestenes) Jacobi rotations; rank with Eigen's default
// threshold  sigma_i > sigma_max * min(rows, cols) * eps  (JacobiSVD::rank()).
int matrix_rank(const double *A, int rows, int cols) {
  std::vector<double> U(A, A + (size_t)std::max(rows, 0) * std::max(cols, 0));
  return sel::matrix_rank_inplace(U.data(), rows, cols);
}

// W = R^{-1} Q^T by Householder QR (lsq_shared.hpp: the same code builds the weights on the device).
void pseudo_inverse(const double *A, int rows, int cols, double *W) {
  std::vector<double> work((size_t)rows * cols + 2 * (size_t)cols + (size_t)rows);
  double *R = work.data(), *vk = R + (size_t)rows * cols, *vn = vk + cols, *y = vn + cols;
  for (size_t a = 0; a < (size_t)rows * cols; ++a) R[a] = A[a];
  lsq::pinv_householder(lsq::Strided{R, 1}, lsq::Strided{vk, 1}, lsq::Strided{vn, 1}, lsq::Strided{y, 1}, rows, cols,
                        [&](int i, int c, double v) { W[(size_t)i * rows + c] = v; });
}

// ---- families --------------------------------------------------------------------------------------
void compute_stencils(HostStencils &S, const HostGrid &g, const StencilFamilyParams &params, std::uint64_t seed) {
  const i64 n = g.n_cells;
  const int ns = params.n_stencils();
  const int nd = g.n_dims;
  S = HostStencils();
  S.n_cells = n;
  S.n_dims = nd;
  S.n_stencils = ns;
  S.params = params;
  S.max_size.resize((size_t)ns);
  S.local_off.assign((size_t)ns + 1, 0);
  for (int k = 0; k < ns; ++k) {
    S.max_size[(size_t)k] = required_stencil_size(params.orders[(size_t)k] - 1, params.overfit_factors[(size_t)k], nd);
    S.local_off[(size_t)k + 1] = S.local_off[(size_t)k] + S.max_size[(size_t)k];
  }
  const int L = S.local_off[(size_t)ns];
  S.l2g_stride = L;
  S.l2g_size.assign((size_t)n, 1);
  parallel_assign(S.l2g, (size_t)(n * L), INVALID);
  parallel_assign(S.local, (size_t)(n * L), 0);
  S.order.assign((size_t)(n * ns), 1);
  S.size.assign((size_t)(n * ns), 0);
  S.k_high.assign((size_t)n, 0);
  S.family_order.assign((size_t)n, 1);
  S.n_family.assign((size_t)n, 1);
  Selector sel(g);
  int error = 0;
  std::string error_msg;

  // members from the device search where it decided them (same families: stencil_shared.hpp)
  DeviceStencilSearch dev;
  bool use_dev = false;
  {
    const char *e = std::getenv("ZFVM_STENCILS");
    const bool force_host = e && e[0] == 'h', force_dev = e && e[0] == 'd';
    if (!force_host && (force_dev || n >= 50000)) {
      std::string why;
      use_dev = device_stencil_search(g, params, dev, why);
      if (!use_dev && force_dev) throw std::runtime_error("ZFVM_STENCILS=device: " + why);
      if (std::getenv("ZFVM_VERBOSE")) {
        i64 n_redo = 0;
        if (use_dev)
          for (i64 i = 0; i < n; ++i) n_redo += dev.redo[(size_t)i] != 0;
        std::fprintf(stderr, "[zfvm stencils] %s%s; %lld of %lld cells left to the host search\n",
                     use_dev ? "device search" : "host search: ", use_dev ? "" : why.c_str(), (long long)(use_dev ? n_redo : n),
                     (long long)n);
      }
    }
  }

#pragma omp parallel
  {
    std::vector<i32> s;
#pragma omp for schedule(dynamic, 256)
    for (i64 i = 0; i < n; ++i) {
      i32 *l2g = &S.l2g[(size_t)(i * L)];
      i32 *local = &S.local[(size_t)(i * L)];
      i32 *order = &S.order[(size_t)(i * ns)];
      i32 *size = &S.size[(size_t)(i * ns)];
      int n_l2g = 0;
      const bool full = (g.cell_flags[i] & FLAG_INTERIOR) || (g.cell_flags[i] & FLAG_GHOST_L1);
      if (!full) {  // StencilFamilyParams{{1}, {"c"}, {1.0}}
        l2g[0] = (i32)i;
        S.l2g_size[(size_t)i] = 1;
        order[0] = 1;
        size[0] = 1;
        continue;
      }
      const bool from_dev = use_dev && !dev.redo[(size_t)i];
      std::mt19937 rng((std::uint32_t)(seed * 2654435761u + (std::uint64_t)i));
      int k_biased = 0;
      for (int k = 0; k < ns; ++k) {
        const int max_order = params.orders[(size_t)k];
        const double factor = params.overfit_factors[(size_t)k];
        const int max_size = S.max_size[(size_t)k];
        if (from_dev) {
          const i32 *m = &dev.members[(size_t)(i * dev.L + S.local_off[(size_t)k])];
          s.assign(m, m + dev.count[(size_t)(i * ns + k)]);
        } else if (params.biases[(size_t)k] == 1) {
          std::string err;
          if (!sel.biased(s, (i32)i, k_biased, max_size, max_order, rng, err)) {
#pragma omp critical
            {
              error = 1;
              error_msg = err;
            }
            s.assign(1, (i32)i);
          }
          ++k_biased;
        } else {
          sel.region_stencil(s, (i32)i, max_size, zfvm::sel::full_sphere());
        }
        // assign_local_indices, stencil.cpp:82-104 (every found cell enters l2g)
        i32 *loc = local + S.local_off[(size_t)k];
        for (size_t a = 0; a < s.size(); ++a) {
          i32 *it = std::find(l2g, l2g + n_l2g, s[a]);
          loc[a] = (i32)(it - l2g);
          if (it == l2g + n_l2g) l2g[n_l2g++] = s[a];
        }
        order[k] = deduce_max_order((int)s.size(), factor, nd);
        size[k] = required_stencil_size(order[k] - 1, factor, nd);
      }
      S.l2g_size[(size_t)i] = n_l2g;
      S.n_family[(size_t)i] = ns;
      int fo = 1;
      for (int k = 0; k < ns; ++k) fo = std::max(fo, (int)order[k]);
      S.family_order[(size_t)i] = fo;
      int kh = 0;  // highest_order_central_stencil, stencil_family.cpp:120-136
      for (int k = 1; k < ns; ++k) {
        if (order[k] > order[kh])
          kh = k;
        else if (order[k] == order[kh] && params.biases[(size_t)k] == 0)
          kh = k;
      }
      S.k_high[(size_t)i] = kh;
    }
  }
  S.error = error;
  S.error_msg = error_msg;
}

void stencil_matrix(std::vector<double> &A, int &rows, int &cols, const HostGrid &g, const HostStencils &s,
                    i64 i, int k) {
  const int size = s.size[(size_t)(i * s.n_stencils + k)];
  const int order = s.order[(size_t)(i * s.n_stencils + k)];
  i32 glob[256];  // zfvm_stencils_compute rejects families whose stencils need more than 256 cells
  for (int j = 0; j < size && j < 256; ++j) glob[j] = s.global(i, k, j);
  assemble_weno_ao_matrix(A, rows, cols, g, glob, std::min(size, 256), order);
}

bool extract_stencils(HostStencils &out, const HostStencils &src, i64 n_local, const i32 *local_to_src,
                      std::string &err) {
  const int ns = src.n_stencils, L = src.l2g_stride;
  out = HostStencils();
  out.n_cells = n_local;
  out.n_dims = src.n_dims;
  out.n_stencils = ns;
  out.params = src.params;
  out.max_size = src.max_size;
  out.local_off = src.local_off;
  out.l2g_stride = L;
  out.l2g_size.assign((size_t)n_local, 1);
  parallel_assign(out.l2g, (size_t)(n_local * L), INVALID);
  parallel_assign(out.local, (size_t)(n_local * L), 0);
  out.order.assign((size_t)(n_local * ns), 1);
  out.size.assign((size_t)(n_local * ns), 0);
  out.k_high.assign((size_t)n_local, 0);
  out.family_order.assign((size_t)n_local, 1);
  out.n_family.assign((size_t)n_local, 1);
  std::vector<i32> src_to_local((size_t)src.n_cells, INVALID);
  for (i64 a = 0; a < n_local; ++a) {
    const i32 i = local_to_src[a];
    if (i < 0 || i >= src.n_cells) {
      err = "extract_stencils: source cell index out of range";
      return false;
    }
    src_to_local[(size_t)i] = (i32)a;
  }
  i64 bad = -1;
#pragma omp parallel for schedule(static)
  for (i64 a = 0; a < n_local; ++a) {
    const i64 i = local_to_src[a];
    i32 *l2g = &out.l2g[(size_t)(a * L)];
    i32 *local = &out.local[(size_t)(a * L)];
    int n_l2g = 0;
    l2g[n_l2g++] = (i32)a;
    const int nf = src.n_family[(size_t)i];
    for (int k = 0; k < nf; ++k) {
      const int size = src.size[(size_t)(i * ns + k)];
      out.order[(size_t)(a * ns + k)] = src.order[(size_t)(i * ns + k)];
      out.size[(size_t)(a * ns + k)] = size;
      i32 *loc = local + src.local_off[(size_t)k];
      for (int j = 0; j < size; ++j) {
        const i32 m = src_to_local[(size_t)src.global(i, k, j)];
        if (m == INVALID) {
#pragma omp critical
          bad = i;
          continue;
        }
        i32 *it = std::find(l2g, l2g + n_l2g, m);
        loc[j] = (i32)(it - l2g);
        if (it == l2g + n_l2g) l2g[n_l2g++] = m;
      }
    }
    out.l2g_size[(size_t)a] = n_l2g;
    out.n_family[(size_t)a] = nf;
    out.family_order[(size_t)a] = src.family_order[(size_t)i];
    out.k_high[(size_t)a] = src.k_high[(size_t)i];
  }
  if (bad >= 0) {
    err = "extract_stencils: a stencil member of source cell " + std::to_string(bad) + " is not part of the sub-grid";
    return false;
  }
  return true;
}

bool import_stencils(HostStencils &S, const HostGrid &g, const StencilFamilyParams &params, const i32 *n_family,
                     const i32 *order, const i32 *size, const i64 *global_offset, const i32 *global, std::string &err) {
  const i64 n = g.n_cells;
  const int ns = params.n_stencils(), nd = g.n_dims;
  S = HostStencils();
  S.n_cells = n;
  S.n_dims = nd;
  S.n_stencils = ns;
  S.params = params;
  S.max_size.resize((size_t)ns);
  S.local_off.assign((size_t)ns + 1, 0);
  for (int k = 0; k < ns; ++k) {
    S.max_size[(size_t)k] = required_stencil_size(params.orders[(size_t)k] - 1, params.overfit_factors[(size_t)k], nd);
    S.local_off[(size_t)k + 1] = S.local_off[(size_t)k] + S.max_size[(size_t)k];
  }
  const int L = S.local_off[(size_t)ns];
  S.l2g_stride = L;
  S.l2g_size.assign((size_t)n, 1);
  parallel_assign(S.l2g, (size_t)(n * L), INVALID);
  parallel_assign(S.local, (size_t)(n * L), 0);
  S.order.assign((size_t)(n * ns), 1);
  S.size.assign((size_t)(n * ns), 0);
  S.k_high.assign((size_t)n, 0);
  S.family_order.assign((size_t)n, 1);
  S.n_family.assign((size_t)n, 1);
  i64 bad = -1;
  int why = 0;
#pragma omp parallel for schedule(static)
  for (i64 i = 0; i < n; ++i) {
    i32 *l2g = &S.l2g[(size_t)(i * L)];
    i32 *local = &S.local[(size_t)(i * L)];
    const int nf = n_family[i];
    int reason = 0;
    if (nf != 1 && nf != ns) reason = 1;  // a family is the full parameter set or truncate_to_first_order
    int n_l2g = 0;
    for (int k = 0; k < nf && !reason; ++k) {
      const int sz = size[i * ns + k], ord = order[i * ns + k];
      const i32 *glob = global + global_offset[i * ns + k];
      if (sz < 1 || sz > S.max_size[(size_t)k] || ord < 1 || ord > params.orders[(size_t)k]) reason = 2;
      // Stencil::Stencil truncates to the size the achieved order needs (stencil.cpp:58-59,78-79)
      else if (nf > 1 && sz != required_stencil_size(ord - 1, params.overfit_factors[(size_t)k], nd)) reason = 3;
      else if (glob[0] != (i32)i) reason = 4;  // member 0 is the cell itself (stencil.cpp:200, LSQ rows start at 1)
      if (reason) break;
      i32 *loc = local + S.local_off[(size_t)k];
      for (int j = 0; j < sz; ++j) {  // assign_local_indices, stencil.cpp:82-104
        if (glob[j] < 0 || glob[j] >= n) {
          reason = 5;
          break;
        }
        i32 *it = std::find(l2g, l2g + n_l2g, glob[j]);
        loc[j] = (i32)(it - l2g);
        if (it == l2g + n_l2g) l2g[n_l2g++] = glob[j];
      }
      S.order[(size_t)(i * ns + k)] = ord;
      S.size[(size_t)(i * ns + k)] = sz;
    }
    if (reason) {
#pragma omp critical
      {
        bad = i;
        why = reason;
      }
      continue;
    }
    S.l2g_size[(size_t)i] = n_l2g;
    S.n_family[(size_t)i] = nf;
    const i32 *ordi = &S.order[(size_t)(i * ns)];
    int fo = 1, kh = 0;
    for (int k = 0; k < nf; ++k) fo = std::max(fo, (int)ordi[k]);
    for (int k = 1; k < nf; ++k) {  // highest_order_central_stencil, stencil_family.cpp:120-136
      if (ordi[k] > ordi[kh])
        kh = k;
      else if (ordi[k] == ordi[kh] && params.biases[(size_t)k] == 0)
        kh = k;
    }
    S.family_order[(size_t)i] = fo;
    S.k_high[(size_t)i] = kh;
  }
  if (bad >= 0) {
    static const char *const text[] = {"", "a family must hold one stencil or the full parameter set",
                                       "stencil size or order outside the parameter set",
                                       "stencil size does not match required_stencil_size(order - 1, overfit factor)",
                                       "member 0 of a stencil must be the cell itself", "stencil member out of range"};
    err = std::string("import_stencils: cell ") + std::to_string(bad) + ": " + text[why];
    return false;
  }
  return true;
}

}  // namespace zfvm
