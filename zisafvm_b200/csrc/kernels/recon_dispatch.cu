// Chooses the compiled reconstruction kernel for a scheme.
#include "recon_inst.cuh"

namespace zfvm {

int launch_recon(const DevicePlan &plan, const SchemeConst &sc, int deg_hi, int deg_lo, const double *state,
                 const std::int32_t *tile_list, std::int64_t n_tiles, cudaStream_t stream) {
  if (sc.n_dims == 2) {
    switch (deg_hi) {
      case 1: return launch_recon_2d_deg1(plan, sc, deg_lo, state, tile_list, n_tiles, stream);
      case 2: return launch_recon_2d_deg2(plan, sc, deg_lo, state, tile_list, n_tiles, stream);
      case 3: return launch_recon_2d_deg3(plan, sc, deg_lo, state, tile_list, n_tiles, stream);
      case 4: return launch_recon_2d_deg4(plan, sc, deg_lo, state, tile_list, n_tiles, stream);
    }
  } else if (sc.n_dims == 3) {
    switch (deg_hi) {
      case 1: return launch_recon_3d_deg1(plan, sc, deg_lo, state, tile_list, n_tiles, stream);
      case 2: return launch_recon_3d_deg2(plan, sc, deg_lo, state, tile_list, n_tiles, stream);
      case 3: return launch_recon_3d_deg3(plan, sc, deg_lo, state, tile_list, n_tiles, stream);
    }
  }
  return 1;
}

}  // namespace zfvm
