// C ABI, host part: grid flattening, stencil families, synthetic meshes (include/zfvm.h).
#include <cstdlib>
#include <cstring>
#include <exception>
#include <string>

#include "../../include/zfvm.h"
#include "host/handles.hpp"

namespace zfvm {
static thread_local std::string g_last_error;
void set_error(const std::string &msg) { g_last_error = msg; }
int fail(const std::string &msg) {
  g_last_error = msg;
  return 1;
}
}  // namespace zfvm

using namespace zfvm;

extern "C" {

const char *zfvm_last_error(void) { return g_last_error.c_str(); }
int zfvm_version(void) { return 100; }

int zfvm_grid_from_mesh(int n_dims, int64_t n_vertices, const double *vertices, int64_t n_cells,
                        const int32_t *vertex_indices, int face_deg, int volume_deg, int moments_deg,
                        zfvm_grid **out) {
  try {
    if (n_dims != 2 && n_dims != 3) return fail("zfvm_grid_from_mesh: n_dims must be 2 or 3");
    if (n_cells <= 0 || n_vertices <= 0) return fail("zfvm_grid_from_mesh: empty mesh");
    const int F = n_dims + 1;
    for (int64_t a = 0; a < n_cells * F; ++a)
      if (vertex_indices[a] < 0 || vertex_indices[a] >= n_vertices)
        return fail("zfvm_grid_from_mesh: vertex index out of range");
    auto *h = new zfvm_grid();
    QRDegrees deg;
    deg.face_deg = face_deg;
    deg.volume_deg = volume_deg;
    deg.moments_deg = moments_deg;
    build_grid(h->g, n_dims, std::vector<double>(vertices, vertices + 3 * n_vertices),
               std::vector<i32>(vertex_indices, vertex_indices + F * n_cells), deg);
    *out = h;
    return 0;
  } catch (const std::exception &e) {
    return fail(std::string("zfvm_grid_from_mesh: ") + e.what());
  }
}

int zfvm_grid_mask_ghost(zfvm_grid *grid, const uint8_t *mask) {
  mask_ghost_cells(grid->g, mask);
  return 0;
}

int zfvm_grid_set_flags(zfvm_grid *grid, const uint8_t *flags) {
  std::memcpy(grid->g.cell_flags.data(), flags, (size_t)grid->g.n_cells);
  return 0;
}

#define ZFVM_RETURN_ARRAY(vec, DT, ...)                       \
  do {                                                         \
    const int64_t shp__[] = {__VA_ARGS__};                     \
    *data = (vec).data();                                      \
    *dtype = (DT);                                             \
    *ndim = (int)(sizeof(shp__) / sizeof(shp__[0]));           \
    for (int d__ = 0; d__ < *ndim; ++d__) shape[d__] = shp__[d__]; \
    return 0;                                                  \
  } while (0)

int zfvm_grid_get(const zfvm_grid *grid, const char *name, const void **data, int *dtype, int *ndim,
                  int64_t shape[4]) {
  const HostGrid &g = grid->g;
  const std::string s(name);
  const int64_t n = g.n_cells, E = g.n_edges, F = g.max_neighbours;
  if (s == "vertices") ZFVM_RETURN_ARRAY(g.vertices, ZFVM_F64, g.n_vertices, 3);
  if (s == "vertex_indices") ZFVM_RETURN_ARRAY(g.vertex_indices, ZFVM_I32, n, F);
  if (s == "neighbours") ZFVM_RETURN_ARRAY(g.neighbours, ZFVM_I32, n, F);
  if (s == "edge_indices") ZFVM_RETURN_ARRAY(g.edge_indices, ZFVM_I32, n, F);
  if (s == "left_right") ZFVM_RETURN_ARRAY(g.left_right, ZFVM_I32, E, 2);
  if (s == "volumes") ZFVM_RETURN_ARRAY(g.volumes, ZFVM_F64, n);
  if (s == "inradii") ZFVM_RETURN_ARRAY(g.inradii, ZFVM_F64, n);
  if (s == "circum_radii") ZFVM_RETURN_ARRAY(g.circum_radii, ZFVM_F64, n);
  if (s == "characteristic_length") ZFVM_RETURN_ARRAY(g.characteristic_length, ZFVM_F64, n);
  if (s == "cell_centers") ZFVM_RETURN_ARRAY(g.cell_centers, ZFVM_F64, n, 3);
  if (s == "cell_qp") ZFVM_RETURN_ARRAY(g.cell_qp, ZFVM_F64, n, g.q_c, 3);
  if (s == "cell_qw") ZFVM_RETURN_ARRAY(g.cell_qw, ZFVM_F64, n, g.q_c);
  if (s == "face_qp") ZFVM_RETURN_ARRAY(g.face_qp, ZFVM_F64, E, g.q_f, 3);
  if (s == "face_qw") ZFVM_RETURN_ARRAY(g.face_qw, ZFVM_F64, E, g.q_f);
  if (s == "face_area") ZFVM_RETURN_ARRAY(g.face_area, ZFVM_F64, E);
  if (s == "face_normal") ZFVM_RETURN_ARRAY(g.face_normal, ZFVM_F64, E, 3);
  if (s == "face_t1") ZFVM_RETURN_ARRAY(g.face_t1, ZFVM_F64, E, 3);
  if (s == "face_t2") ZFVM_RETURN_ARRAY(g.face_t2, ZFVM_F64, E, 3);
  if (s == "face_centers") ZFVM_RETURN_ARRAY(g.face_centers, ZFVM_F64, E, 3);
  if (s == "face_vertex_slots") ZFVM_RETURN_ARRAY(g.face_vertex_slots, ZFVM_U8, n, F);
  if (s == "moments") ZFVM_RETURN_ARRAY(g.moments, ZFVM_F64, n, g.n_moments);
  if (s == "cell_flags") ZFVM_RETURN_ARRAY(g.cell_flags, ZFVM_U8, n);
  if (s == "cell_rule_weights") ZFVM_RETURN_ARRAY(g.cell_rule.weights, ZFVM_F64, g.q_c);
  if (s == "cell_rule_bary") ZFVM_RETURN_ARRAY(g.cell_rule.bary, ZFVM_F64, g.q_c, g.cell_rule.n_bary);
  if (s == "face_rule_weights") ZFVM_RETURN_ARRAY(g.face_rule.weights, ZFVM_F64, g.q_f);
  if (s == "face_rule_bary") ZFVM_RETURN_ARRAY(g.face_rule.bary, ZFVM_F64, g.q_f, g.face_rule.n_bary);
  return fail(std::string("zfvm_grid_get: unknown array '") + name + "'");
}

int zfvm_grid_info(const zfvm_grid *grid, int64_t info[8]) {
  const HostGrid &g = grid->g;
  info[0] = g.n_dims;
  info[1] = g.n_cells;
  info[2] = g.n_vertices;
  info[3] = g.n_edges;
  info[4] = g.n_interior_edges;
  info[5] = g.q_c;
  info[6] = g.q_f;
  info[7] = g.n_moments;
  return 0;
}

void zfvm_grid_free(zfvm_grid *grid) { delete grid; }

static int export_mesh(RawMesh &m, int hilbert, int64_t *n_vertices, double **vertices, int64_t *n_cells,
                       int32_t **vertex_indices) {
  const int F = m.n_dims + 1;
  const int64_t nc = (int64_t)m.vertex_indices.size() / F;
  if (hilbert) {
    // renumber_grid.cpp:68-82: cell centre = vertex average
    std::vector<double> c((size_t)(3 * nc), 0.0);
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < nc; ++i)
      for (int d = 0; d < 3; ++d) {
        double s = 0.0;
        for (int k = 0; k < F; ++k) s += m.vertices[3 * (size_t)m.vertex_indices[(size_t)(i * F + k)] + d];
        c[(size_t)(3 * i + d)] = s / F;
      }
    auto perm = hilbert_permutation(m.n_dims, nc, c.data());
    renumber_mesh_cells(m, perm);
  }
  *n_vertices = (int64_t)m.vertices.size() / 3;
  *n_cells = nc;
  *vertices = (double *)std::malloc(m.vertices.size() * sizeof(double));
  *vertex_indices = (int32_t *)std::malloc(m.vertex_indices.size() * sizeof(int32_t));
  if (!*vertices || !*vertex_indices) return fail("mesh export: out of memory");
  std::memcpy(*vertices, m.vertices.data(), m.vertices.size() * sizeof(double));
  std::memcpy(*vertex_indices, m.vertex_indices.data(), m.vertex_indices.size() * sizeof(int32_t));
  return 0;
}

int zfvm_mesh_square(int nx, int ny, double x0, double x1, double y0, double y1, double jitter, uint64_t seed,
                     int hilbert, int64_t *n_vertices, double **vertices, int64_t *n_cells,
                     int32_t **vertex_indices) {
  try {
    RawMesh m = make_square_mesh(nx, ny, x0, x1, y0, y1, jitter, seed);
    return export_mesh(m, hilbert, n_vertices, vertices, n_cells, vertex_indices);
  } catch (const std::exception &e) {
    return fail(std::string("zfvm_mesh_square: ") + e.what());
  }
}

int zfvm_mesh_cube(int nx, int ny, int nz, double h, double x0, double y0, double z0, double jitter,
                   uint64_t seed, int hilbert, const int offset[3], const int global[3], int64_t *n_vertices,
                   double **vertices, int64_t *n_cells, int32_t **vertex_indices) {
  try {
    int o[3] = {0, 0, 0}, gl[3] = {-1, -1, -1};
    if (offset) std::memcpy(o, offset, sizeof(o));
    if (global) std::memcpy(gl, global, sizeof(gl));
    RawMesh m = make_cube_mesh(nx, ny, nz, h, x0, y0, z0, jitter, seed, o[0], o[1], o[2], gl[0], gl[1], gl[2]);
    return export_mesh(m, hilbert, n_vertices, vertices, n_cells, vertex_indices);
  } catch (const std::exception &e) {
    return fail(std::string("zfvm_mesh_cube: ") + e.what());
  }
}

int zfvm_mesh_read_msh_h5(const char *path, int *n_dims, int64_t *n_vertices, double **vertices, int64_t *n_cells,
                          int32_t **vertex_indices) {
  try {
    RawMesh m;
    read_msh_h5(path, m);
    *n_dims = m.n_dims;
    return export_mesh(m, 0, n_vertices, vertices, n_cells, vertex_indices);
  } catch (const std::exception &e) {
    return fail(std::string("zfvm_mesh_read_msh_h5: ") + e.what());
  }
}

int zfvm_mesh_write_msh_h5(const char *path, int n_dims, int64_t n_vertices, const double *vertices, int64_t n_cells,
                           const int32_t *vertex_indices) {
  try {
    if (n_dims != 2 && n_dims != 3) return fail("zfvm_mesh_write_msh_h5: n_dims must be 2 or 3");
    RawMesh m;
    m.n_dims = n_dims;
    m.vertices.assign(vertices, vertices + 3 * n_vertices);
    m.vertex_indices.assign(vertex_indices, vertex_indices + (n_dims + 1) * n_cells);
    write_msh_h5(path, m);
    return 0;
  } catch (const std::exception &e) {
    return fail(std::string("zfvm_mesh_write_msh_h5: ") + e.what());
  }
}

int zfvm_mesh_read_subgrid_h5(const char *path, int *n_dims, int64_t *n_vertices, double **vertices, int64_t *n_cells,
                              int32_t **vertex_indices, int64_t **partition, int64_t **global_cell_indices) {
  try {
    RawMesh m;
    read_msh_h5(path, m);
    if (m.partition.empty())
      return fail(std::string("zfvm_mesh_read_subgrid_h5: ") + path + " has no 'partition' / 'global_cell_indices' (a plain grid file?)");
    *n_dims = m.n_dims;
    const size_t nc = m.partition.size();
    int64_t *pt = static_cast<int64_t *>(std::malloc(std::max<size_t>(nc, 1) * sizeof(int64_t)));
    int64_t *gc = static_cast<int64_t *>(std::malloc(std::max<size_t>(nc, 1) * sizeof(int64_t)));
    if (!pt || !gc) {
      std::free(pt);
      std::free(gc);
      return fail("zfvm_mesh_read_subgrid_h5: out of memory");
    }
    std::memcpy(pt, m.partition.data(), nc * sizeof(int64_t));
    std::memcpy(gc, m.global_cell_indices.data(), nc * sizeof(int64_t));
    if (export_mesh(m, 0, n_vertices, vertices, n_cells, vertex_indices)) {
      std::free(pt);
      std::free(gc);
      return 1;
    }
    *partition = pt;
    *global_cell_indices = gc;
    return 0;
  } catch (const std::exception &e) {
    return fail(std::string("zfvm_mesh_read_subgrid_h5: ") + e.what());
  }
}

int zfvm_mesh_write_subgrid_h5(const char *path, int n_dims, int64_t n_vertices, const double *vertices, int64_t n_cells,
                               const int32_t *vertex_indices, const int64_t *partition, const int64_t *global_cell_indices) {
  try {
    if (n_dims != 2 && n_dims != 3) return fail("zfvm_mesh_write_subgrid_h5: n_dims must be 2 or 3");
    if (n_cells < 1) return fail("zfvm_mesh_write_subgrid_h5: an empty sub-grid");
    for (int64_t i = 0; i < n_cells; ++i)
      if (partition[i] < 0 || global_cell_indices[i] < 0)
        return fail("zfvm_mesh_write_subgrid_h5: partition and global cell indices are unsigned in the reference (int_t)");
    RawMesh m;
    m.n_dims = n_dims;
    m.vertices.assign(vertices, vertices + 3 * n_vertices);
    m.vertex_indices.assign(vertex_indices, vertex_indices + (n_dims + 1) * n_cells);
    m.partition.assign(partition, partition + n_cells);
    m.global_cell_indices.assign(global_cell_indices, global_cell_indices + n_cells);
    write_msh_h5(path, m);
    return 0;
  } catch (const std::exception &e) {
    return fail(std::string("zfvm_mesh_write_subgrid_h5: ") + e.what());
  }
}

void zfvm_free(void *p) { std::free(p); }

int zfvm_stencils_compute(const zfvm_grid *grid, int n_stencils, const int *orders, const char *biases,
                          const double *overfit_factors, uint64_t seed, zfvm_stencils **out) {
  try {
    if (n_stencils <= 0 || n_stencils > 6) return fail("zfvm_stencils_compute: 1..6 stencils supported");
    StencilFamilyParams p;
    for (int k = 0; k < n_stencils; ++k) {
      p.orders.push_back(orders[k]);
      if (biases[k] != 'c' && biases[k] != 'b') return fail("zfvm_stencils_compute: bias must be 'c' or 'b'");
      p.biases.push_back(biases[k] == 'b' ? 1 : 0);
      p.overfit_factors.push_back(overfit_factors[k]);
      if (orders[k] < 1) return fail("zfvm_stencils_compute: a non-positive convergence order?");
      if (!(overfit_factors[k] >= 1.0)) return fail("zfvm_stencils_compute: overfit factors must be >= 1");
      if (required_stencil_size(orders[k] - 1, overfit_factors[k], grid->g.n_dims) > 256)
        return fail("zfvm_stencils_compute: a stencil of more than 256 cells (order / overfit factor too large)");
      if (poly_dof(orders[k] - 1, grid->g.n_dims) > grid->g.n_moments && orders[k] > 2)
        return fail("zfvm_stencils_compute: moments_deg of the grid is lower than the polynomial degree");
    }
    auto *h = new zfvm_stencils();
    compute_stencils(h->s, grid->g, p, seed);
    if (h->s.error) {
      std::string msg = h->s.error_msg;
      delete h;
      return fail("zfvm_stencils_compute: " + msg);
    }
    *out = h;
    return 0;
  } catch (const std::exception &e) {
    return fail(std::string("zfvm_stencils_compute: ") + e.what());
  }
}

int zfvm_stencils_get(const zfvm_stencils *st, const char *name, const void **data, int *dtype, int *ndim,
                      int64_t shape[4]) {
  const HostStencils &S = st->s;
  const std::string s(name);
  const int64_t n = S.n_cells, ns = S.n_stencils, L = S.l2g_stride;
  if (s == "l2g") ZFVM_RETURN_ARRAY(S.l2g, ZFVM_I32, n, L);
  if (s == "l2g_size") ZFVM_RETURN_ARRAY(S.l2g_size, ZFVM_I32, n);
  if (s == "local") ZFVM_RETURN_ARRAY(S.local, ZFVM_I32, n, L);
  if (s == "local_off") ZFVM_RETURN_ARRAY(S.local_off, ZFVM_I32, ns + 1);
  if (s == "max_size") ZFVM_RETURN_ARRAY(S.max_size, ZFVM_I32, ns);
  if (s == "order") ZFVM_RETURN_ARRAY(S.order, ZFVM_I32, n, ns);
  if (s == "size") ZFVM_RETURN_ARRAY(S.size, ZFVM_I32, n, ns);
  if (s == "k_high") ZFVM_RETURN_ARRAY(S.k_high, ZFVM_I32, n);
  if (s == "n_family") ZFVM_RETURN_ARRAY(S.n_family, ZFVM_I32, n);
  if (s == "family_order") ZFVM_RETURN_ARRAY(S.family_order, ZFVM_I32, n);
  return fail(std::string("zfvm_stencils_get: unknown array '") + name + "'");
}

/* LSQ matrix A of stencil k of cell i, row-major; returns rows/cols (lsq_solver.cpp:40-47). */
int zfvm_stencil_matrix(const zfvm_grid *grid, const zfvm_stencils *st, int64_t i, int k, double *A, int max_count,
                        int *rows, int *cols) {
  try {
    std::vector<double> tmp;
    stencil_matrix(tmp, *rows, *cols, grid->g, st->s, i, k);
    if ((int)tmp.size() > max_count) return fail("zfvm_stencil_matrix: buffer too small");
    std::memcpy(A, tmp.data(), tmp.size() * sizeof(double));
    return 0;
  } catch (const std::exception &e) {
    return fail(std::string("zfvm_stencil_matrix: ") + e.what());
  }
}

/* Batch variant: all matrices of all cells, padded to rows_max*cols_max per stencil
 * (A_off[k] offsets into a per-cell record of A_stride doubles). */
int zfvm_stencil_matrices(const zfvm_grid *grid, const zfvm_stencils *st, double *A, int64_t A_stride,
                          const int64_t *A_off) {
  try {
    const HostStencils &S = st->s;
    const int ns = S.n_stencils;
#pragma omp parallel
    {
      std::vector<double> tmp;
#pragma omp for schedule(dynamic, 256)
      for (int64_t i = 0; i < S.n_cells; ++i)
        for (int k = 0; k < S.n_family[(size_t)i]; ++k) {
          if (S.order[(size_t)(i * ns + k)] <= 1) continue;
          int rows, cols;
          stencil_matrix(tmp, rows, cols, grid->g, S, i, k);
          std::memcpy(A + i * A_stride + A_off[k], tmp.data(), tmp.size() * sizeof(double));
        }
    }
    return 0;
  } catch (const std::exception &e) {
    return fail(std::string("zfvm_stencil_matrices: ") + e.what());
  }
}

int zfvm_pseudo_inverse(const double *A, int rows, int cols, double *W) {
  pseudo_inverse(A, rows, cols, W);
  return 0;
}

int zfvm_quadrature_rule(int kind, int deg, int *n_points, int *n_bary, double *weights, double *bary, int max_points) {
  try {
    RefRule r = kind == 1 ? make_edge_rule(deg) : (kind == 2 ? make_triangular_rule(deg) : make_tetrahedral_rule(deg));
    if (r.n_points > max_points) return fail("zfvm_quadrature_rule: buffer too small");
    *n_points = r.n_points;
    *n_bary = r.n_bary;
    std::memcpy(weights, r.weights.data(), r.weights.size() * sizeof(double));
    std::memcpy(bary, r.bary.data(), r.bary.size() * sizeof(double));
    return 0;
  } catch (const std::exception &e) {
    return fail(std::string("zfvm_quadrature_rule: ") + e.what());
  }
}

int zfvm_gauss_legendre(int n, double *points, double *weights) {
  gauss_legendre(n, points, weights);
  return 0;
}

int zfvm_deduce_max_order(int stencil_size, double factor, int n_dims) {
  return deduce_max_order(stencil_size, factor, n_dims);
}

void zfvm_stencils_free(zfvm_stencils *st) { delete st; }

int zfvm_stencils_extract(const zfvm_stencils *src, int64_t n_local, const int32_t *local_to_src, zfvm_stencils **out) {
  try {
    auto *h = new zfvm_stencils();
    std::string err;
    if (!extract_stencils(h->s, src->s, n_local, local_to_src, err)) {
      delete h;
      return fail("zfvm_stencils_extract: " + err);
    }
    *out = h;
    return 0;
  } catch (const std::exception &e) {
    return fail(std::string("zfvm_stencils_extract: ") + e.what());
  }
}

int zfvm_stencils_from_arrays(const zfvm_grid *grid, int n_stencils, const int *orders, const char *biases,
                              const double *overfit_factors, const int32_t *n_family, const int32_t *order,
                              const int32_t *size, const int64_t *global_offset, const int32_t *global,
                              zfvm_stencils **out) {
  try {
    if (n_stencils <= 0 || n_stencils > 6) return fail("zfvm_stencils_from_arrays: 1..6 stencils supported");
    StencilFamilyParams p;
    for (int k = 0; k < n_stencils; ++k) {
      if (biases[k] != 'c' && biases[k] != 'b') return fail("zfvm_stencils_from_arrays: bias must be 'c' or 'b'");
      if (orders[k] < 1) return fail("zfvm_stencils_from_arrays: a non-positive convergence order?");
      if (!(overfit_factors[k] >= 1.0)) return fail("zfvm_stencils_from_arrays: overfit factors must be >= 1");
      if (required_stencil_size(orders[k] - 1, overfit_factors[k], grid->g.n_dims) > 256)
        return fail("zfvm_stencils_from_arrays: a stencil of more than 256 cells (order / overfit factor too large)");
      if (poly_dof(orders[k] - 1, grid->g.n_dims) > grid->g.n_moments && orders[k] > 2)
        return fail("zfvm_stencils_from_arrays: moments_deg of the grid is lower than the polynomial degree");
      p.orders.push_back(orders[k]);
      p.biases.push_back(biases[k] == 'b' ? 1 : 0);
      p.overfit_factors.push_back(overfit_factors[k]);
    }
    auto *h = new zfvm_stencils();
    std::string err;
    if (!import_stencils(h->s, grid->g, p, n_family, order, size, global_offset, global, err)) {
      delete h;
      return fail("zfvm_stencils_from_arrays: " + err);
    }
    *out = h;
    return 0;
  } catch (const std::exception &e) {
    return fail(std::string("zfvm_stencils_from_arrays: ") + e.what());
  }
}

int zfvm_partition_kway(const zfvm_grid *grid, const zfvm_stencils *stencils, int n_parts, int32_t *partition) {
  try {
    std::vector<i32> part;
    std::string err;
    if (stencils != nullptr && stencils->s.n_cells != grid->g.n_cells)
      return fail("zfvm_partition_kway: stencils do not belong to this grid");
    if (!partition_kway(part, grid->g, stencils ? &stencils->s : nullptr, n_parts, err)) return fail("zfvm_" + err);
    std::memcpy(partition, part.data(), part.size() * sizeof(int32_t));
    return 0;
  } catch (const std::exception &e) {
    return fail(std::string("zfvm_partition_kway: ") + e.what());
  }
}

int zfvm_has_metis(void) { return metis_available() ? 1 : 0; }

int zfvm_hilbert_permutation(int n_dims, int64_t n, const double *centers, int32_t *perm) {
  try {
    if (n_dims != 2 && n_dims != 3) return fail("zfvm_hilbert_permutation: n_dims must be 2 or 3");
    const std::vector<i32> p = hilbert_permutation(n_dims, n, centers);
    std::memcpy(perm, p.data(), (size_t)n * sizeof(int32_t));
    return 0;
  } catch (const std::exception &e) {
    return fail(std::string("zfvm_hilbert_permutation: ") + e.what());
  }
}

}  // extern "C"
