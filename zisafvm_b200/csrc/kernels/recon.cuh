// K1: shared definitions of the reconstruction kernels (arguments, polynomial evaluation).
//
// Per cell, K1 (recon_tile.cuh, recon_coop.cuh, recon_generic.cu) does what
//   EulerGlobalReconstruction::compute   global_reconstruction_impl.hpp:133-174
//   LocalReconstruction::compute         local_reconstruction.hpp:69-120
//   HybridWENO::compute_polys_impl       hybrid_weno.cpp:72-92   (as W_k * rhs, W_k = pinv(A_k))
//   CWENO_AO::reconstruct_impl           cweno_ao.cpp:36-53
//   HybridWENO::eno_hybridize            hybrid_weno.cpp:110-128
//   rc(i)(x) at the face Gauss points    flux_loop.hpp:131-149, local_reconstruction.hpp:149-163
//   GravitySourceLoop (both variants)    gravity_source_loop.hpp:32-87,121-148
// do in the reference, and writes
//   trace[e][side][q][5]   the reconstructed state at every Gauss point of the cell's faces
//   source[i][5]           the cell's gravity source term (already divided by the volume; source.cuh or the generic kernel)
// so that the face kernel (K2) is a pure Riemann-solver pass.
#pragma once
#include "common.cuh"
#include "equilibrium.cuh"

namespace zfvm {

enum ReconVariant : int { RV_PLAIN = 0, RV_GRAVITY = 1, RV_WELL_BALANCED = 2 };

struct ReconArgs {
  DevicePlan plan;
  const double *state;            // [n][5]
  const std::int32_t *tile_list;  // optional list of tiles to process (null: all)
  std::int64_t n_tiles_launch;
};

template <int ND, int DEG>
struct PolyEval {
  static constexpr int D = dof_of(DEG, ND);
  /// monomials mono[i] = xi^a eta^b zeta^c - c_i for i = 1..D-1 (mono[0] unused: a0 * (1 - 0))
  ZFVM_DEVICE static void monomials(double xi, double eta, double zeta, const double *cmom, double *mono) {
    double px[DEG + 1], py[DEG + 1], pz[DEG + 1];
    px[0] = py[0] = pz[0] = 1.0;
#pragma unroll
    for (int d = 1; d <= DEG; ++d) {
      px[d] = px[d - 1] * xi;
      py[d] = py[d - 1] * eta;
      pz[d] = pz[d - 1] * zeta;
    }
    constexpr ExpoTable<ND, DEG> tab{};
#pragma unroll
    for (int i = 1; i < D; ++i) {
      double m = (ND == 2) ? px[tab.e[i].a] * py[tab.e[i].b] : px[tab.e[i].a] * py[tab.e[i].b] * pz[tab.e[i].c];
      mono[i] = m - cmom[i];
    }
  }
};

}  // namespace zfvm
