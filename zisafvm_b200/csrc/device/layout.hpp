// Device data layout for the per-stage residual path (see DESIGN.md "Data layout in HBM").
//
// Cells are grouped in *tiles* of 32 consecutive cells. Every per-cell array that is only read by
// the owning cell is stored tile-interleaved, [tile][field][32], so that a warp's access to one
// field is one fully coalesced 256-byte (FP64) or 128-byte (int32) transaction.
// The reconstruction tables of a tile (stencil meta data, stencil member indices, pseudo-inverse
// weights) form one contiguous *tile record*, so that the reconstruction kernel can stream them
// with a handful of TMA bulk copies (cp.async.bulk) per tile:
//   record = | meta u64[32] | sidx_0 i32[RM_0][32] | .. | sidx_{NS-1} | W_0 f64[RM_0][NC_0][32] | .. | W_{NS-1} |
// Arrays that are gathered through indices (the state, equilibrium potentials, face fluxes) stay
// row-major AoS, the layout of zisa::GridVariables (grid_variables_decl.hpp:17).
#pragma once
#include <cstdint>
#ifndef __CUDACC__
#define __host__
#define __device__
#endif

namespace zfvm {

constexpr int TILE = 32;
constexpr int MAX_STENCILS = 6;
constexpr int NVARS = 5;        // rho, m1, m2, m3, E  (euler_variables.hpp:30-36)
constexpr int MAX_QF = 8;
constexpr int MAX_QC = 16;
constexpr int MAX_AVARS = 8;    // advected scalars per cell (AllVariables::avars)

// face_ref bit layout
constexpr std::uint32_t FREF_EDGE_MASK = 0x0FFFFFFFu;  // edge index
constexpr std::uint32_t FREF_SIDE = 1u << 28;           // 0: this cell is the left cell, 1: right
constexpr std::uint32_t FREF_TRACE = 1u << 29;          // the flux loop needs this cell's trace here
constexpr std::uint32_t FREF_INTERIOR = 1u << 30;       // edge has two cells

enum GravityKind : int {
  GRAVITY_NONE = 0,
  GRAVITY_CONSTANT = 1,     // ConstantGravity       gravity_impl.hpp:13-20
  GRAVITY_POINT_MASS = 2,   // PointMassGravity      gravity_impl.hpp:22-29
  GRAVITY_POLYTROPE = 3,    // PolytropeGravity      gravity_impl.hpp:31-58
  GRAVITY_TABLE = 4,        // SphericalGravity (piecewise-linear table)
  GRAVITY_USER = 5,         // potentials supplied by the caller at all quadrature points
};
enum AlignmentKind : int { ALIGN_RADIAL = 0, ALIGN_AXIAL = 1 };
enum ScalingKind : int { SCALING_UNITY = 0, SCALING_EULER = 1 };  // characteristic_scale.hpp:14-46
enum FluxKind : int { FLUX_HLLC = 0, FLUX_RUSANOV = 1 };
enum ReconMode : int { RECON_CWENO_AO = 0, RECON_WENO_AO = 1 };

/// Quadrature tables and scheme constants, passed by value to kernels (fits the 4 KB param space).
struct SchemeConst {
  int n_dims, n_stencils, q_f, q_c;
  int recon_mode, scaling, flux, well_balanced, has_gravity;
  int rows_max[MAX_STENCILS];    // padded rows of W_k (= max stencil size - 1)
  int ncoef[MAX_STENCILS];       // padded coefficient count of stencil k (without the constant)
  double lin_w[MAX_STENCILS];    // normalised linear weights (hybrid_weno.cpp:26-31)
  double epsilon, exponent;
  double gamma;
  double face_w[MAX_QF];         // reference weights (sum to 1)
  double face_bary[MAX_QF][3];   // barycentric coordinates w.r.t. the left cell's face vertices
  double cell_w[MAX_QC];
  double cell_bary[MAX_QC][4];
  double heating_rate, heating_r0, heating_r1;  // Heating (model/heating.hpp:54-80); rate 0: off
  // isentropic EOS power x^(1/(gamma-1)): eos_pow_n = 2/(gamma-1) when that is an integer in [2, 8] (square-root and
  // multiplication forms), else 0 -> pow(x, eos_pow_e)
  int eos_pow_n;
  double eos_pow_e;
  // LocalRCParams (local_reconstruction.hpp:22-25): the equilibrium, its averages and the characteristic scale of a cell
  // are refreshed every steps_per_recompute-th evaluation or when the cell has drifted by recompute_threshold
  int steps_per_recompute;
  double recompute_threshold;
};

/// Raw device pointers of one context. Sizes in comments use n = n_cells, T = n_tiles, E = n_edges.
struct DevicePlan {
  std::int64_t n_cells, n_tiles, n_edges, n_interior_edges;
  // reconstruction: tile records (layout above). All offsets in bytes from the start of a record.
  const char *rec;                          // [T][rec_bytes]
  std::int64_t rec_bytes;                   // multiple of 128
  int hdr_bytes;                            // meta + all sidx_k (= offset of W_0)
  int off_sidx[MAX_STENCILS];               // sidx_k: [rows_max_k][32] global index of stencil member
  int off_W[MAX_STENCILS];                  // W_k:    [rows_max_k][ncoef_k][32]
  // meta (offset 0): [32] u64, byte k: rows of stencil k; byte 7: k_high | single<<4
  // reconstruction, tile kernel (kernels/recon_tile.cuh): self-contained tile records; null if not built
  const char *rec2;                         // [T][rec2_bytes]
  std::int64_t rec2_bytes;
  int rec2_cap;                             // capacity of a record's row list
  // geometry (tile-interleaved)
  const double *vtx;                        // [T][F][3][32]
  const double *center;                     // [T][3][32]
  const double *inv_len;                    // [T][32]   1 / characteristic_length
  const double *volume;                     // [T][32]
  const double *moments;                    // [T][n_mom][32]  c(3 .. n_mom+2)
  int n_mom;
  const std::uint32_t *face_ref;            // [T][F][32]
  const std::uint8_t *face_slots;           // [T][F][32]
  const std::uint8_t *cell_flags;           // [n]
  // faces
  const std::int32_t *left_right;           // [E][2]; left = -1 for faces the flux loop skips
  const double *face_frame;                 // [E][10]: n(3) t1(3) t2(3) area
  // gravity tables
  const double *phi_cqp;                    // [n][q_c]
  const double *gradphi_cqp;                // [n][q_c][3]
  const double *phi_fqp;                    // [E][q_f]
  // work arrays
  double *trace;                            // [E][2][q_f][5]
  double *flux;                             // [E][5]
  double *source;                           // [n][5]
  // helpers
  __host__ __device__ const std::uint64_t *meta_of(std::int64_t tile) const {
    return reinterpret_cast<const std::uint64_t *>(rec + tile * rec_bytes);
  }
  __host__ __device__ const std::int32_t *sidx_of(std::int64_t tile, int k) const {
    return reinterpret_cast<const std::int32_t *>(rec + tile * rec_bytes + off_sidx[k]);
  }
  __host__ __device__ const double *W_of(std::int64_t tile, int k) const {
    return reinterpret_cast<const double *>(rec + tile * rec_bytes + off_W[k]);
  }
  double *poly;                             // optional diagnostics [n][n_poly_coef][5]; may be null
  double *poly_scale;                       // optional [n][5]
  double *poly_tile;                        // tile kernel -> source_kernel hand-over [T][D + 1][5][32]; may be null
  int n_poly_coef;
  int *eq_fail;                             // counter of cells whose equilibrium solve failed
  // well-balanced runs: per-cell local equilibrium (h_ref, K, phi_ref, found) and its cell averages over every
  // stencil member, written by the two equilibrium kernels (equilibrium.cuh) ahead of the reconstruction
  double *eq_par;                           // [n][4]
  double *eq_avg;                           // [T][eq_rows][2][32]: (rho_bar, E_bar) of stencil row r = row0_k + j
  int eq_rows;                              // sum_k rows_max_k (+ 1 with tile records: the last row is the cell itself)
  int eq_row0[MAX_STENCILS];                // older records: first row of stencil k (tile records: lidx row order)
  double *eq_bg;                            // tile records: [E_int][2][q_f][2] equilibrium (rho, E) at the face Gauss points
  // steps_per_recompute != 1 (null otherwise): LocalReconstruction::steps_since_recompute, the (rho, E_int) the cached
  // scale was formed from, and this evaluation's verdict of recompute_equilibrium (local_reconstruction.hpp:87-100)
  std::int32_t *eq_steps;                   // [n]
  double *scale_state;                      // [n][2]
  std::uint8_t *eq_flag;                    // [n] 1: the cell's equilibrium data are recomputed in this evaluation
  int rec2_off_list, rec2_off_lidx, rec2_lidx_elem;  // where a tile record keeps its row list and local indices
  // advected scalars (tracers.cu): traces and face fluxes of the n_avars scalars; null when n_avars == 0
  int n_avars;
  double *qtrace;                           // [E_int][2][q_f][n_avars]
  double *qflux;                            // [E_int][n_avars]
};

}  // namespace zfvm
