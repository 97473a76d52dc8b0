// Host-side precompute for the B200 residual path: flattened (SoA) grid, quadrature,
// stencil families, least-squares matrices and pseudo-inverse weights.
//
// Everything here is *input preparation* for the device kernels; it is executed once.
// The algorithms follow the reference so that numbering, orientation and stencil
// membership are the ones a ZisaFVM driver would hand us:
//   grid connectivity / numbering     src/zisa/grid/grid.cpp:385-442,520-540,651-723
//   face-vertex tables                src/zisa/grid/gmsh_reader.cpp:22-107
//   quadrature rules                  src/zisa/math/{edge_rule,triangular_rule,tetrahedral_rule}.cpp
//   normalized moments                src/zisa/grid/grid.cpp:1049-1098
//   stencil selection                 src/zisa/reconstruction/stencil.cpp:168-399
//   LSQ matrix assembly               src/zisa/reconstruction/lsq_solver.cpp:168-403
#pragma once

#include <cstdint>
#include <memory>
#include <new>
#include <string>
#include <utility>
#include <vector>

namespace zfvm {

using i32 = std::int32_t;
using i64 = std::int64_t;

constexpr i32 INVALID = -1;  // reference: magic_index_value (grid_impl.hpp:19)

enum CellFlagBits : std::uint8_t {
  FLAG_INTERIOR = 1,   // cell_flags.hpp:9-15
  FLAG_GHOST = 2,
  FLAG_GHOST_L1 = 4,
};

struct Vec3 {
  double x, y, z;
  double &operator[](int i) { return (&x)[i]; }
  double operator[](int i) const { return (&x)[i]; }
};

/// Reference quadrature rule on the unit element (barycentric points).
struct RefRule {
  int n_points = 0;
  int n_bary = 0;               // 2 (edge: stored as (0.5-0.5 xi, 0.5+0.5 xi)), 3, 4
  std::vector<double> weights;  // sum to 1
  std::vector<double> bary;     // [n_points][n_bary]
  std::vector<double> xi;       // edge rule only: Gauss-Legendre nodes on [-1,1]
};

RefRule make_edge_rule(int deg);          // edge_rule.cpp:10-32 (+ gauss_legendre.hpp)
RefRule make_triangular_rule(int deg);    // triangular_rule.cpp:33-91
RefRule make_tetrahedral_rule(int deg);   // tetrahedral_rule.cpp:10-122
void gauss_legendre(int n, double *points, double *weights);  // Fourier-Newton

constexpr int MAX_TRIANGULAR_RULE_DEGREE = 5;
constexpr int MAX_TETRAHEDRAL_RULE_DEGREE = 3;

int poly_dof(int deg, int n_dims);             // poly2d_impl.hpp:12-18
int poly_index2(int a, int b);                 // poly2d_impl.hpp:34-37
int poly_index3(int a, int b, int c);          // poly2d_impl.hpp:39-41
int required_stencil_size(int deg, double factor, int n_dims);  // stencil.cpp:168-175
int deduce_max_order(int stencil_size, double factor, int n_dims);  // stencil.cpp:158-165

struct QRDegrees {
  int face_deg = 1, volume_deg = 1, moments_deg = 1;
};

/// Flattened grid. Mirrors zisa::Grid (grid_decl.hpp:34-107) as plain arrays.
/// std::vector whose resize() leaves trivially constructible elements uninitialised, and a fill that touches the pages from
/// all threads: the per-cell / per-face tables are 0.1-1.6 GB each at 10 M tetrahedra (every element is written by the
/// parallel loop that follows the resize), and value-initialising them from one thread does not scale with the cores.
template <class T>
struct NoInitAlloc : std::allocator<T> {
  template <class U>
  struct rebind {
    using other = NoInitAlloc<U>;
  };
  template <class U>
  void construct(U *p) {
    ::new ((void *)p) U;
  }
  template <class U, class... A>
  void construct(U *p, A &&...a) {
    ::new ((void *)p) U(std::forward<A>(a)...);
  }
};
template <class T>
using BigVec = std::vector<T, NoInitAlloc<T>>;
using BigVecI32 = BigVec<i32>;
template <class T>
inline void parallel_assign(BigVec<T> &v, size_t n, T value) {
  v.clear();
  v.resize(n);
  T *p = v.data();
#pragma omp parallel for schedule(static)
  for (long long i = 0; i < (long long)n; ++i) p[i] = value;
}

struct HostGrid {
  int n_dims = 0;
  int max_neighbours = 0;  // faces (= vertices) per cell: 3 or 4
  i64 n_cells = 0, n_vertices = 0, n_edges = 0, n_interior_edges = 0;
  QRDegrees deg;

  std::vector<double> vertices;       // [n_vertices][3]
  std::vector<i32> vertex_indices;    // [n_cells][F], standard (outward) order
  std::vector<i32> neighbours;        // [n_cells][F], INVALID on the boundary
  std::vector<i32> edge_indices;      // [n_cells][F]
  std::vector<i32> left_right;        // [n_edges][2], right = INVALID on the boundary

  BigVec<double> volumes, inradii, circum_radii, characteristic_length;  // [n_cells]
  BigVec<double> cell_centers;   // [n_cells][3] quadrature barycentre (cell.cpp:9-11)

  RefRule cell_rule, face_rule;
  int q_c = 0, q_f = 0;
  BigVec<double> cell_qp;        // [n_cells][q_c][3]
  BigVec<double> cell_qw;        // [n_cells][q_c]     (physical weights)
  BigVec<double> face_qp;        // [n_edges][q_f][3]
  BigVec<double> face_qw;        // [n_edges][q_f]
  BigVec<double> face_area;      // [n_edges]
  BigVec<double> face_normal, face_t1, face_t2;  // [n_edges][3]
  BigVec<double> face_centers;   // [n_edges][3]
  /// For cell i, local face k: positions (within cell i's vertex list) of the face's
  /// vertices *in the order used by the left cell* (2 bits each, v0 | v1<<2 | v2<<4).
  std::vector<std::uint8_t> face_vertex_slots;  // [n_cells][F]

  int n_moments = 0;                  // poly_dof(moments_deg)
  BigVec<double> moments;        // [n_cells][n_moments]

  std::vector<std::uint8_t> cell_flags;  // [n_cells]

  Vec3 vertex(i64 i, int k) const {
    const double *p = &vertices[3 * (i64)vertex_indices[i * max_neighbours + k]];
    return {p[0], p[1], p[2]};
  }
  Vec3 center(i64 i) const { return {cell_centers[3 * i], cell_centers[3 * i + 1], cell_centers[3 * i + 2]}; }
  bool is_valid(i64 i, int k) const { return neighbours[i * max_neighbours + k] != INVALID; }
};

/// tet face k -> local vertex; triangle: (k + rel) % 3.  gmsh_reader.cpp:22-56
int relative_vertex_index(int n_dims, int k, int rel);
/// vertex opposite to face k.  gmsh_reader.cpp:58-82
int relative_off_vertex_index(int n_dims, int k);

/// Build the flattened grid from raw vertices / connectivity (Grid::Grid, grid.cpp:651-723).
/// `vertex_indices` is [n_cells][n_dims+1]; it is re-ordered to the standard orientation.
void build_grid(HostGrid &g, int n_dims, std::vector<double> vertices, std::vector<i32> vertex_indices,
                const QRDegrees &deg);

/// mask_ghost_cells (grid.cpp:1122-1136). mask[i] != 0 -> ghost.
void mask_ghost_cells(HostGrid &g, const std::uint8_t *mask);

/// Hilbert-curve renumbering of cells (src/renumber_grid.cpp:46-135).
/// Returns perm with perm[new] = old.
std::vector<i32> hilbert_permutation(int n_dims, i64 n, const double *centers);

// ---- synthetic grids ---------------------------------------------------------
struct RawMesh {
  int n_dims = 0;
  std::vector<double> vertices;     // [nv][3]
  std::vector<i32> vertex_indices;  // [nc][n_dims+1]
  // sub-grid files only (src/domain_decomposition.cpp:84-88, load_distributed_grid): owner rank and global index of every
  // local cell; both empty for a plain grid file
  std::vector<std::int64_t> partition, global_cell_indices;  // [nc]
};
/// [x0,x1]x[y0,y1], nx*ny squares split on alternating diagonals, interior vertices jittered.
RawMesh make_square_mesh(int nx, int ny, double x0, double x1, double y0, double y1, double jitter,
                         std::uint64_t seed);
/// box of nx*ny*nz cubes, 6 Kuhn tetrahedra each (conforming), interior vertices jittered.
/// (ox,oy,oz) and (gx,gy,gz) place the box inside a global lattice so that the jitter of a
/// vertex depends only on its global lattice position (used for domain-decomposed generation).
RawMesh make_cube_mesh(int nx, int ny, int nz, double h, double x0, double y0, double z0, double jitter,
                       std::uint64_t seed, int ox = 0, int oy = 0, int oz = 0, int gx = -1, int gy = -1,
                       int gz = -1);
void renumber_mesh_cells(RawMesh &m, const std::vector<i32> &perm);
/// The reference's grid file `*.msh.h5` (load_grid_gmsh_h5, grid.cpp:889-901): datasets n_dims, vertex_indices, vertices.
/// Dependency-free subset of the HDF5 file format (msh_h5.cpp); both throw std::runtime_error with the reason.
void read_msh_h5(const std::string &path, RawMesh &mesh);
void write_msh_h5(const std::string &path, const RawMesh &mesh);

// ---- stencils -------------------------------------------------------------------------------
struct StencilFamilyParams {
  std::vector<int> orders;
  std::vector<int> biases;  // 0 = central ("c"), 1 = one-sided ("b")
  std::vector<double> overfit_factors;
  int n_stencils() const { return (int)orders.size(); }
};

/// All stencil families of a grid, fixed-stride storage. Mirrors StencilFamily / Stencil
/// (stencil_family.cpp:15-45, stencil.cpp:42-104).
struct HostStencils {
  i64 n_cells = 0;
  int n_dims = 0;
  int n_stencils = 0;            // per full family; order-1 families use slot 0 only
  StencilFamilyParams params;
  std::vector<int> max_size;     // [n_stencils] required_stencil_size(order-1, factor)
  std::vector<int> local_off;    // [n_stencils+1] prefix sums of max_size
  int l2g_stride = 0;            // = local_off[n_stencils]
  std::vector<i32> l2g_size;     // [n_cells]
  BigVecI32 l2g;                 // [n_cells][l2g_stride]; l2g[i][0] == i
  std::vector<i32> order;        // [n_cells][n_stencils] achieved order (1 for unused slots)
  std::vector<i32> size;         // [n_cells][n_stencils] cells actually used (0 for unused slots)
  BigVecI32 local;               // [n_cells][l2g_stride]; stencil k at local_off[k], `size` entries
  std::vector<i32> k_high;       // [n_cells]
  std::vector<i32> family_order; // [n_cells]
  std::vector<i32> n_family;     // [n_cells] number of stencils in the family (1 for o1 cells)
  int error = 0;
  std::string error_msg;

  i32 global(i64 i, int k, int j) const {  // j-th member of stencil k of cell i
    return l2g[(size_t)(i * l2g_stride + local[(size_t)(i * l2g_stride + local_off[(size_t)k] + j)])];
  }
};

/// compute_stencil_families (stencil_family.cpp:99-117).  The members are selected on the current CUDA device when there
/// is one (kernels/stencil_search.cu; cells it leaves undecided -- equal distances, random retries -- and everything
/// else on the host); ZFVM_STENCILS=host keeps the whole search on the host, ZFVM_STENCILS=device insists on the device
/// also for small grids.  Both give the same families (tests/test_gpu_parity.py).
void compute_stencils(HostStencils &s, const HostGrid &g, const StencilFamilyParams &params,
                      std::uint64_t seed = 0);

/// Result of the device search: stencil k of cell i has count[i][k] members at members[i * L + local_off[k] ..];
/// redo[i] != 0: the host search decides the cell.
struct DeviceStencilSearch {
  int L = 0;
  BigVecI32 members;             // [n_cells][L] (1.6 GB at 10 M tetrahedra: pages touched from all threads)
  std::vector<i32> count;
  std::vector<std::uint8_t> redo;
};
/// false (with the reason) when there is no device or the family is outside the kernel's limits.
bool device_stencil_search(const HostGrid &g, const StencilFamilyParams &params, DeviceStencilSearch &out, std::string &why);

/// LSQ matrix of stencil k of cell i (LSQSolver ctor, lsq_solver.cpp:40-47); row-major rows x cols.
void stencil_matrix(std::vector<double> &A, int &rows, int &cols, const HostGrid &g, const HostStencils &s,
                    i64 i, int k);

/// assemble_weno_ao_matrix (lsq_solver.cpp:168-403). `stencil` = global indices, A is
/// (n-1) x (dof(order-1)-1) row-major.
void assemble_weno_ao_matrix(std::vector<double> &A, int &n_rows, int &n_cols, const HostGrid &g,
                             const i32 *stencil, int n, int order);

namespace sel { struct GridView; }
/// Raw-pointer view of the arrays the stencil search reads (stencil_shared.hpp); valid while `g` lives.
sel::GridView make_grid_view(const HostGrid &g);

/// singular values by one-sided Jacobi; used for the rank test of stencil.cpp:352.
int matrix_rank(const double *A, int rows, int cols);

/// W = pinv(A) (cols x rows, row-major) by Householder QR.
void pseudo_inverse(const double *A, int rows, int cols, double *W);

// ---- domain decomposition -------------------------------------------------------------------
/// Stencils of a sub-grid cut out of the grid `src` was computed on (the reference extracts the stencils
/// of a partition from the global ones, domain_decomposition.cpp:412-447): local cell a is cell
/// local_to_src[a] of the source grid.  Families keep their members (renumbered); every used member
/// must be part of the sub-grid.  Returns false and sets `err` otherwise.
bool extract_stencils(HostStencils &out, const HostStencils &src, i64 n_local, const i32 *local_to_src,
                      std::string &err);

/// METIS k-way partition of the cells on the stencil graph (stencils != nullptr) or the face-neighbour graph, with the
/// reference's options (compute_partition_full_stencil, src/zisa/parallelization/domain_decomposition.cpp:27-113).
bool metis_available();
bool partition_kway(std::vector<i32> &part, const HostGrid &g, const HostStencils *stencils, int n_parts, std::string &err);

/// Stencil families the caller already holds (the reference's array<StencilFamily, 1>,
/// include/zisa/reconstruction/global_reconstruction_decl.hpp:107-147): stencil k of cell i has size[i][k] members whose
/// global indices start at global[global_offset[i * n_stencils + k]] (Stencil::global(), member 0 = the cell), achieved
/// order order[i][k]; n_family[i] is 1 for families truncated to first order.  l2g / local are rebuilt by
/// assign_local_indices (stencil.cpp:82-104), k_high by stencil_family.cpp:120-136.
bool import_stencils(HostStencils &out, const HostGrid &g, const StencilFamilyParams &params, const i32 *n_family,
                     const i32 *order, const i32 *size, const i64 *global_offset, const i32 *global, std::string &err);

}  // namespace zfvm
