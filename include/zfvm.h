/* zfvm.h -- C ABI of the B200 residual path for ZisaFVM-style solvers.
 *
 * The reference (1uc/ZisaFVM) has no FFI layer: the operator API its hot path sits behind is the
 * C++ virtual interface zisa::RateOfChange (include/zisa/ode/rate_of_change.hpp:23-43) plus the
 * secondary interfaces TimeIntegration::compute_step (include/zisa/ode/time_integration.hpp:42-44),
 * CFLCondition::operator() (include/zisa/model/cfl_condition.hpp:12-18), BoundaryCondition::apply
 * (include/zisa/boundary/boundary_condition.hpp:17) and HaloExchange
 * (include/zisa/parallelization/halo_exchange.hpp:11-21).  Each entry point below names the
 * reference member function it stands in for; INTEGRATION.md shows the C++ adapter classes a
 * maintainer adds on the ZisaFVM side.
 *
 * Conventions: plain pointers and sizes only; every function returns 0 on success and a non-zero
 * code on failure with a message available from zfvm_last_error() (the reference reports errors
 * with LOG_ERR, which prints and terminates -- the adapter turns non-zero into LOG_ERR).
 * One context per GPU / rank; a context is driven from a single host thread, like
 * RungeKutta::compute_step drives RateOfChange::compute.  State arrays are row-major
 * [n_cells][5] doubles in the order (rho, rho v1, rho v2, rho v3, E), the layout of
 * zisa::AllVariables::cvars (include/zisa/model/all_variables.hpp:31-35).
 */
#ifndef ZFVM_H_
#define ZFVM_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct zfvm_grid zfvm_grid;         /* flattened host grid (mirrors zisa::Grid) */
typedef struct zfvm_stencils zfvm_stencils; /* stencil families + least-squares matrices */
typedef struct zfvm_ctx zfvm_ctx;           /* device context: one GPU, one (sub-)grid */

enum zfvm_dtype { ZFVM_F64 = 0, ZFVM_I32 = 1, ZFVM_I64 = 2, ZFVM_U8 = 3 };

/* ---- errors ---------------------------------------------------------------------------------- */
const char *zfvm_last_error(void);
int zfvm_version(void);

/* ---- host precompute: grid ---------------------------------------------------------------------
 * zfvm_grid_from_mesh  <->  zisa::Grid::Grid(element_type, vertices, vertex_indices, QRDegrees)
 *                           (src/zisa/grid/grid.cpp:651-723); n_dims = 2 triangles, 3 tetrahedra.
 * zfvm_grid_mask_ghost <->  zisa::mask_ghost_cells (src/zisa/grid/grid.cpp:1122-1136)
 * zfvm_grid_set_flags       copies zisa::Grid::cell_flags verbatim (bit0 interior, bit1 ghost_cell,
 *                           bit2 ghost_cell_l1; include/zisa/grid/cell_flags.hpp:9-15)
 */
int zfvm_grid_from_mesh(int n_dims, int64_t n_vertices, const double *vertices, int64_t n_cells,
                        const int32_t *vertex_indices, int face_deg, int volume_deg, int moments_deg,
                        zfvm_grid **out);
int zfvm_grid_mask_ghost(zfvm_grid *grid, const uint8_t *mask);
int zfvm_grid_set_flags(zfvm_grid *grid, const uint8_t *flags);
/* Named array access, e.g. "volumes", "cell_centers", "left_right", "neighbours", "edge_indices",
 * "cell_qp", "cell_qw", "face_qp", "face_qw", "face_normal", "moments", "cell_flags", "inradii", ...
 * Pointers stay valid until zfvm_grid_free. */
int zfvm_grid_get(const zfvm_grid *grid, const char *name, const void **data, int *dtype, int *ndim,
                  int64_t shape[4]);
int zfvm_grid_info(const zfvm_grid *grid, int64_t info[8]); /* n_dims, n_cells, n_vertices, n_edges,
                                                               n_interior_edges, q_c, q_f, n_moments */
void zfvm_grid_free(zfvm_grid *grid);

/* Synthetic meshes (the reference ships no grids, rsync.exclude: grids/ *) and the Hilbert
 * renumbering of src/renumber_grid.cpp:46-135.  Outputs are malloc'ed; free with zfvm_free. */
int zfvm_mesh_square(int nx, int ny, double x0, double x1, double y0, double y1, double jitter, uint64_t seed,
                     int hilbert, int64_t *n_vertices, double **vertices, int64_t *n_cells,
                     int32_t **vertex_indices);
int zfvm_mesh_cube(int nx, int ny, int nz, double h, double x0, double y0, double z0, double jitter,
                   uint64_t seed, int hilbert, const int offset[3], const int global[3], int64_t *n_vertices,
                   double **vertices, int64_t *n_cells, int32_t **vertex_indices);
/* The reference's grid file *.msh.h5 (load_grid_gmsh_h5, src/zisa/grid/grid.cpp:889-901: datasets n_dims, vertex_indices,
 * vertices; written by src/renumber_grid.cpp:129-132).  A dependency-free subset of the HDF5 file format
 * (csrc/host/msh_h5.cpp: superblock 0-3, symbol-table or compact-link root group, contiguous / compact datasets of
 * little-endian integers and IEEE floats); anything else is rejected with the reason in zfvm_last_error.  The writer
 * emits what libhdf5 1.8 / h5py write by default, with the reference's 64-bit unsigned indices.  Outputs are
 * malloc'ed; free with zfvm_free.  The result goes to zfvm_grid_from_mesh like a generated mesh. */
int zfvm_mesh_read_msh_h5(const char *path, int *n_dims, int64_t *n_vertices, double **vertices, int64_t *n_cells,
                          int32_t **vertex_indices);
int zfvm_mesh_write_msh_h5(const char *path, int n_dims, int64_t n_vertices, const double *vertices, int64_t n_cells,
                           const int32_t *vertex_indices);
/* Sub-grid files subgrid-%04d.msh.h5 of the reference's partition tool (src/domain_decomposition.cpp:70-88: the datasets of
 * a grid file plus `partition` and `global_cell_indices`, one int_t per local cell; read back by load_distributed_grid,
 * src/zisa/parallelization/distributed_grid.cpp:10-17): a rank's cells in the reference's local numbering -- owned cells
 * first, then the halo grouped per owner (make_halo, src/zisa/mpi/parallelization/mpi_halo_exchange.cpp:203-238) -- with
 * the owner rank and the global index of every cell.  zfvm_set_halo takes exactly the ranges make_halo derives from them. */
int zfvm_mesh_read_subgrid_h5(const char *path, int *n_dims, int64_t *n_vertices, double **vertices, int64_t *n_cells,
                              int32_t **vertex_indices, int64_t **partition, int64_t **global_cell_indices);
int zfvm_mesh_write_subgrid_h5(const char *path, int n_dims, int64_t n_vertices, const double *vertices, int64_t n_cells,
                               const int32_t *vertex_indices, const int64_t *partition, const int64_t *global_cell_indices);
void zfvm_free(void *p);

/* ---- host precompute: stencils -------------------------------------------------------------------
 * zfvm_stencils_compute <-> zisa::compute_stencil_families (src/zisa/reconstruction/stencil_family.cpp:99-117)
 *                           + LSQSolver ctor (src/zisa/reconstruction/lsq_solver.cpp:40-47) per stencil.
 * biases: one char per stencil, 'c' central or 'b' one-sided (stencil_bias.cpp). */
int zfvm_stencils_compute(const zfvm_grid *grid, int n_stencils, const int *orders, const char *biases,
                          const double *overfit_factors, uint64_t seed, zfvm_stencils **out);
/* names: "l2g_offset", "l2g", "order", "size", "local_offset", "local", "k_high", "n_family",
 * "A_offset", "A" */
int zfvm_stencils_get(const zfvm_stencils *st, const char *name, const void **data, int *dtype, int *ndim,
                      int64_t shape[4]);
void zfvm_stencils_free(zfvm_stencils *st);
/* Stencils of a sub-grid cut out of the grid `src` belongs to: local cell a is cell local_to_src[a] of the
 * source grid (extraction of a partition's stencils, src/domain_decomposition.cpp /
 * src/zisa/grid/domain_decomposition.cpp:412-447).  Every used stencil member must be part of the sub-grid. */
int zfvm_stencils_extract(const zfvm_stencils *src, int64_t n_local, const int32_t *local_to_src, zfvm_stencils **out);
/* Import of stencil families the caller already holds -- the reference's array<StencilFamily, 1> as
 * EulerGlobalReconstruction keeps it (include/zisa/reconstruction/global_reconstruction_decl.hpp:107-147) -- instead of
 * re-running the selection: wherever the reference's choice is not deterministic (unordered ties of std::sort,
 * std::random_device retries, src/zisa/reconstruction/stencil.cpp:240-250,332-334) this is how the drop-in works "on the
 * same inputs".  orders / biases / overfit_factors: the StencilFamilyParams; n_family[i]: StencilFamily::size() (1 after
 * truncate_to_first_order); order[i][k], size[i][k] ([n_cells][n_stencils]): Stencil::order() / size();
 * global[global_offset[i * n_stencils + k] ..]: Stencil::global() (member 0 is cell i).  local2global() and
 * Stencil::local() are rebuilt in the reference's order (assign_local_indices, stencil.cpp:82-104). */
int zfvm_stencils_from_arrays(const zfvm_grid *grid, int n_stencils, const int *orders, const char *biases,
                              const double *overfit_factors, const int32_t *n_family, const int32_t *order,
                              const int32_t *size, const int64_t *global_offset, const int32_t *global,
                              zfvm_stencils **out);
/* Owner part of every cell from METIS_PartGraphKway on the stencil graph (stencils != NULL: an edge between a cell and
 * every member of its combined stencil) or on the face-neighbour graph (stencils == NULL), with the reference's options
 * OBJTYPE_VOL, NCUTS 10, NITER 20, UFACTOR 100 (compute_partition_full_stencil / compute_partitioned_grid,
 * src/zisa/parallelization/domain_decomposition.cpp:27-113,250-270).  The space-filling-curve partition
 * (compute_partitioned_grid_by_sfc, :577-609) needs no entry point: contiguous chunks of the Hilbert order.
 * zfvm_has_metis() == 0: built without the METIS archive, zfvm_partition_kway fails like ZISA_HAS_METIS == 0 does. */
int zfvm_partition_kway(const zfvm_grid *grid, const zfvm_stencils *stencils, int n_parts, int32_t *partition);
int zfvm_has_metis(void);
/* Hilbert ordering of cell centres [n][3] (src/renumber_grid.cpp:60-126): perm[new] = old. */
int zfvm_hilbert_permutation(int n_dims, int64_t n, const double *centers, int32_t *perm);
/* LSQSolver::A of stencil k of cell i, row-major rows x cols (lsq_solver.cpp:40-47,168-403) */
int zfvm_stencil_matrix(const zfvm_grid *grid, const zfvm_stencils *st, int64_t i, int k, double *A, int max_count,
                        int *rows, int *cols);
/* all matrices: cell i, stencil k at A[i * A_stride + A_off[k]] */
int zfvm_stencil_matrices(const zfvm_grid *grid, const zfvm_stencils *st, double *A, int64_t A_stride,
                          const int64_t *A_off);
/* W = pinv(A), cols x rows row-major: what the device applies instead of LDLT(A^T A).solve(A^T rhs)
 * (lsq_solver.cpp:82) */
int zfvm_pseudo_inverse(const double *A, int rows, int cols, double *W);
/* reference rules: kind 1 EdgeRule (edge_rule.cpp), 2 TriangularRule (triangular_rule.cpp),
 * 3 TetrahedralRule (tetrahedral_rule.cpp); Gauss-Legendre nodes by Fourier-Newton (gauss_legendre.hpp) */
int zfvm_quadrature_rule(int kind, int deg, int *n_points, int *n_bary, double *weights, double *bary, int max_points);
int zfvm_gauss_legendre(int n, double *points, double *weights);
int zfvm_deduce_max_order(int stencil_size, double factor, int n_dims); /* stencil.cpp:158-165 */

/* ---- scheme parameters ---------------------------------------------------------------------------- */
typedef struct zfvm_params {
  /* HybridWENOParams (include/zisa/reconstruction/hybrid_weno_params.hpp) */
  int recon_mode;            /* 0 CWENO-AO (cweno_ao.cpp), 1 WENO-AO (weno_ao.cpp) */
  double linear_weights[8];  /* un-normalised, one per stencil */
  double epsilon, exponent;
  /* "well-balancing.mode": 0 constant (NoEquilibrium), 1 isentropic (euler_experiment_impl.hpp:385-397) */
  int well_balanced;
  int scaling;               /* 0 UnityScaling, 1 EulerScaling (characteristic_scale.hpp) */
  int flux;                  /* 0 HLLC (flux/hllc.hpp), 1 Rusanov (not in the reference) */
  /* IdealGasEOS(gamma, specific_gas_constant) */
  double gamma, gas_constant;
  /* gravity (model/gravity_decl.hpp): kind 0 none, 1 constant g, 2 point mass (GM, X),
   * 3 polytrope (rhoC, K, G), 4 radial table (set with zfvm_set_gravity_table),
   * 5 user potentials at all quadrature points (zfvm_set_gravity_values);
   * alignment 0 radial, 1 axial (axis) */
  int gravity_kind, gravity_alignment;
  double gravity_p[4];
  double gravity_axis[3];
  /* LocalRCParams{steps_per_recompute, recompute_threshold} (local_reconstruction.hpp:22-25, 87-100; JSON keys
   * "reconstruction.steps_per_recompute" / ".recompute_threshold", euler_experiment_impl.hpp:83-90): a cell's local
   * equilibrium, its averages over the stencil, its values at the cell's Gauss points and the characteristic scale are
   * refreshed every steps_per_recompute-th evaluation of the rate of change, or earlier when (rho, E_int) has moved
   * away from the cached equilibrium average by recompute_threshold in units of the cached scale (the threshold is the
   * last member of this struct).  1 = every evaluation.  Other values need a family of the experiments' shape. */
  int steps_per_recompute;
  int keep_polynomials;      /* diagnostics: store every cell's WENO polynomial */
  /* "flux-bc" (numerical_experiment.cpp:238-256 adds it to the FVM rate of change): 0 NoFluxBC, 1 FluxBC
   * (include/zisa/boundary/flux_bc.hpp:13-52): on exterior faces the physical flux of the cell average leaves the cell,
   * 2 EquilibriumFluxBC (include/zisa/boundary/equilibrium_flux_bc.hpp:18-75): the pressure of the cell's local
   * isentropic equilibrium at the exterior face's Gauss points (needs a gravity model) */
  int flux_bc;
  /* advected scalars: AllVariables::avars[n_cells][n_avars] (include/zisa/model/all_variables.hpp:31-35), each
   * reconstructed on its own (LocalReconstruction::compute_tracer, local_reconstruction.hpp:127-147) and upwinded
   * with the HLLC wave speeds (HLLCBatten::tracer_flux, flux/hllc.hpp:178-197).  At most 8. */
  int n_avars;
  /* Heating (include/zisa/model/heating.hpp:18-80): dE/dt += average(rho * heating_rate * [r0 <= |x| <= r1]);
   * heating_rate == 0: no heating term */
  double heating_rate, heating_r0, heating_r1;
  double recompute_threshold; /* LocalRCParams::recompute_threshold, see steps_per_recompute */
} zfvm_params;

void zfvm_params_default(zfvm_params *p);

/* ---- device context --------------------------------------------------------------------------------
 * zfvm_create builds what EulerExperiment::choose_physical_rate_of_change builds
 * (include/zisa/experiments/euler_experiment_impl.hpp:303-313): reconstruction array, flux loop,
 * gravity source loop -- as device-resident weight / index / geometry tables.  The host arrays of
 * `grid` and `stencils` are only borrowed during the call. */
int zfvm_create(const zfvm_grid *grid, const zfvm_stencils *stencils, const zfvm_params *params, int device,
                zfvm_ctx **out);
void zfvm_destroy(zfvm_ctx *ctx);
int zfvm_set_gravity_table(zfvm_ctx *ctx, const zfvm_grid *grid, int64_t n, const double *radii, const double *phi);
int zfvm_set_gravity_values(zfvm_ctx *ctx, const double *phi_cell_qp, const double *grad_phi_cell_qp,
                            const double *phi_face_qp);
/* bytes of device memory held; algorithmic bytes per cell and RK stage (SURVEY.md 8d formula,
 * evaluated on the actual stencils) */
int zfvm_memory_info(const zfvm_ctx *ctx, int64_t *device_bytes, double *algorithmic_bytes_per_cell_stage);
void *zfvm_stream(zfvm_ctx *ctx); /* cudaStream_t used by all kernels of this context */

/* RateOfChange::compute(tendency, current_state, t) for Sum[FluxLoop, GravitySourceLoop]
 * (fvm_loops/flux_loop.hpp:96-104, fvm_loops/gravity_source_loop.hpp:32-87,121-148).
 * accumulate != 0: tendency += rate (the reference contract after ZeroRateOfChange);
 * accumulate == 0: tendency  = rate (ZeroRateOfChange folded in).
 * Host version: copies state in, tendency out (and in, when accumulating). */
int zfvm_rate_of_change(zfvm_ctx *ctx, double *tendency_host, const double *state_host, double t, int accumulate);
int zfvm_rate_of_change_device(zfvm_ctx *ctx, double *tendency_dev, const double *state_dev, double t,
                               int accumulate);

/* The same for AllVariables{cvars, avars} (n_avars > 0): the avars tendency is the tracer part of
 * FluxLoop::compute_patch (fvm_loops/flux_loop.hpp:157-161,180-192); sources do not touch avars. */
int zfvm_rate_of_change_av(zfvm_ctx *ctx, double *tendency_host, double *tendency_avars_host, const double *state_host,
                           const double *state_avars_host, double t, int accumulate);
int zfvm_rate_of_change_av_device(zfvm_ctx *ctx, double *tendency_dev, double *tendency_avars_dev,
                                  const double *state_dev, const double *state_avars_dev, double t, int accumulate);

/* TimeIntegration::compute_step for RungeKutta (src/zisa/ode/runge_kutta.cpp:87-143).
 * method: "forward_euler", "ssp2", "ssp3", "wicker", "rk4", "fehlberg" (make_tableau :145-213).
 * The state lives on the device between calls. */
int zfvm_set_time_integration(zfvm_ctx *ctx, const char *method);
int zfvm_upload_state(zfvm_ctx *ctx, const double *state_host);
int zfvm_download_state(zfvm_ctx *ctx, double *state_host);
double *zfvm_state_device(zfvm_ctx *ctx);
/* resident advected scalars [n_cells][n_avars]; zfvm_rk_step advances them together with the state */
int zfvm_upload_avars(zfvm_ctx *ctx, const double *avars_host);
int zfvm_download_avars(zfvm_ctx *ctx, double *avars_host);
double *zfvm_avars_device(zfvm_ctx *ctx);
/* FrozenBC (src/zisa/boundary/frozen_boundary_condition.cpp:11-55): ghost rows are reset to
 * `steady_state_host` after every stage.  Pass NULL for NoBoundaryCondition. */
int zfvm_set_frozen_bc(zfvm_ctx *ctx, const double *steady_state_host);
int zfvm_apply_frozen_bc(zfvm_ctx *ctx, double *state_dev);
/* FrozenBC on AllVariables: ghost rows of cvars and avars (frozen_boundary_condition.cpp:39-55 copies both) */
int zfvm_set_frozen_bc_av(zfvm_ctx *ctx, const double *steady_state_host, const double *steady_avars_host);
/* One RK step on the resident state.  If dt_next / not_plausible are non-NULL the CFL time step
 * cfl_number * min inradius/(|v|+a) (LocalCFL, model/local_cfl_condition_impl.hpp:25-40) and the
 * SanityCheckFor<Euler> flag of the new state come back with it (one 16-byte D2H copy). */
int zfvm_rk_step(zfvm_ctx *ctx, double t, double dt, double cfl_number, double *dt_next, int *not_plausible);
/* Same with host buffers: u0_host -> u1_host (one H2D + one D2H copy of the state).  Grids of a million cells and more
 * take a chunked route -- the rows go up in chunks that gate the stage-0 reconstruction of the tiles they complete, the
 * last stage is finished and sent back chunk by chunk -- with bit-identical results; in a multi-rank context the route
 * posts the same halo exchanges (one per stage) as the plain sequence, after the last upload of stage 0 and ahead of the
 * last stage's chunks, so ranks may mix the two routes. */
int zfvm_rk_step_host(zfvm_ctx *ctx, const double *u0_host, double *u1_host, double t, double dt);
int zfvm_rk_step_host_av(zfvm_ctx *ctx, const double *u0_host, const double *a0_host, double *u1_host, double *a1_host,
                         double t, double dt);
/* LocalCFL on a device / the resident state */
int zfvm_cfl_dt(zfvm_ctx *ctx, const double *state_dev, double cfl_number, double *dt, int *not_plausible);
int zfvm_synchronize(zfvm_ctx *ctx);
/* per-kernel device timing: CUDA events around K1 (reconstruction), K2 (face flux), K3 (update) of every
 * residual evaluation while enabled; read returns the summed milliseconds and the launch counts */
int zfvm_profile_enable(zfvm_ctx *ctx, int enable);
int zfvm_profile_read(zfvm_ctx *ctx, double ms[3], int64_t counts[3]);
/* the same for the advected-scalar kernels (T1 reconstruction + T2 flux + T3 update, one event pair per residual) */
int zfvm_profile_read_tracers(zfvm_ctx *ctx, double *ms, int64_t *count);
/* counters: [0] kernels launched since creation, [1] cells whose equilibrium solve failed */
int zfvm_counters(zfvm_ctx *ctx, int64_t counters[4]);
/* diagnostics (keep_polynomials): [n_cells][n_coef][5] coefficients in the scaled basis, [n_cells][5] scales */
int zfvm_download_polynomials(zfvm_ctx *ctx, double *coeffs_host, double *scale_host, int *n_coef);
int zfvm_download_work(zfvm_ctx *ctx, const char *name, double *host, int64_t max_count);

/* ---- multi-GPU: HaloExchange (src/zisa/mpi/parallelization/mpi_halo_exchange.cpp:109-251) -----------
 * Local cells are [0, n_owned) owned, then halo cells grouped contiguously per owner rank.
 * recv_begin/recv_end: rows of the local state that peer p fills; send_index: local rows packed for
 * peer p (concatenated, send_offset[n_peers+1]).  The exchange itself is NCCL send/recv in one group. */
int zfvm_nccl_unique_id(char id_out[128]);
int zfvm_comm_init(zfvm_ctx *ctx, const char id[128], int rank, int n_ranks);
int zfvm_set_halo(zfvm_ctx *ctx, int64_t n_owned, int n_peers, const int *peer_rank, const int64_t *recv_begin,
                  const int64_t *recv_end, const int64_t *send_offset, const int32_t *send_index);
/* HaloExchange::operator()(AllVariables&) and HaloExchange::wait() (include/zisa/parallelization/halo_exchange.hpp:11-21;
 * MPIHaloExchange::exchange / wait, mpi_halo_exchange.cpp:178-201): post packs and starts the NCCL group on the context's
 * communication stream and returns; wait orders every later kernel of the context behind the transfer.  state_dev NULL:
 * the resident state; avars_dev NULL with n_avars > 0: the resident scalars. */
int zfvm_halo_post(zfvm_ctx *ctx, double *state_dev, double *avars_dev);
int zfvm_halo_wait(zfvm_ctx *ctx);
int zfvm_halo_exchange(zfvm_ctx *ctx, double *state_dev); /* post + wait in one call */
/* the same exchange for cvars and avars rows in one NCCL group (mpi_halo_exchange.cpp:178-201 exchanges both) */
int zfvm_halo_exchange_av(zfvm_ctx *ctx, double *state_dev, double *avars_dev);
int zfvm_allreduce_min(zfvm_ctx *ctx, double *value);     /* MPIAllReduce MIN (mpi_all_reduce.cpp:16-24) */

#ifdef __cplusplus
}
#endif
#endif /* ZFVM_H_ */
