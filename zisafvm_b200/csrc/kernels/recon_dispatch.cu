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

void launch_eq_decide(const DevicePlan &plan, const SchemeConst &sc, const double *state, const std::int32_t *tile_list,
                      std::int64_t n_tiles, cudaStream_t stream) {
  const unsigned grid = (unsigned)((n_tiles * TILE + 255) / 256);
  eq_decide_kernel<0><<<grid, 256, 0, stream>>>(plan, sc, state, tile_list, n_tiles);
}
template <int POWN>
void launch_eq_solve(const DevicePlan &plan, const SchemeConst &sc, const double *state, const std::int32_t *tile_list,
                     std::int64_t n_tiles, unsigned grid, cudaStream_t stream) {
  eq_solve_kernel<POWN><<<grid, 256, 0, stream>>>(plan, sc, state, tile_list, n_tiles);
}
template <int POWN>
void launch_eq_tile(const DevicePlan &plan, const SchemeConst &sc, const std::int32_t *tile_list, std::int64_t n_tiles,
                    cudaStream_t stream) {
  const unsigned g2 = (unsigned)((n_tiles * plan.eq_rows + 7) / 8);
  const unsigned g3 = (unsigned)((n_tiles * (sc.n_dims + 1) + 7) / 8);
  const size_t smem = (size_t)plan.rec2_cap * sc.q_c * sizeof(double);  // potentials of the tile's row list
  static const bool e2_v1 = [] {
    const char *e = std::getenv("ZFVM_EQ_MEMBER");
    return e != nullptr && e[0] == 'v';
  }();
  if (smem <= 48 * 1024 && !e2_v1)
    eq_member_tile_smem_kernel<POWN><<<(unsigned)n_tiles, 256, smem, stream>>>(plan, sc, tile_list, n_tiles);
  else
    eq_member_tile_kernel<POWN><<<g2, 256, 0, stream>>>(plan, sc, tile_list, n_tiles);
  eq_face_kernel<POWN><<<g3, 256, 0, stream>>>(plan, sc, tile_list, n_tiles);
}
#define ZFVM_EQ_INST(POWN)                                                                                          \
  template void launch_eq_solve<POWN>(const DevicePlan &, const SchemeConst &, const double *, const std::int32_t *,    \
                                      std::int64_t, unsigned, cudaStream_t);                                            \
  template void launch_eq_tile<POWN>(const DevicePlan &, const SchemeConst &, const std::int32_t *, std::int64_t, cudaStream_t);
ZFVM_EQ_INST(0)
ZFVM_EQ_INST(2)
ZFVM_EQ_INST(3)
ZFVM_EQ_INST(5)
#undef ZFVM_EQ_INST

namespace {
unsigned long long *g_tile_prof = nullptr;
bool g_tile_prof_init = false;
}  // namespace

unsigned long long *tile_prof_buffer() {
  if (!g_tile_prof_init) {
    g_tile_prof_init = true;
    const char *e = std::getenv("ZFVM_TILE_PROF");
    if (e && e[0] == '1' && cudaMalloc((void **)&g_tile_prof, 16 * sizeof(unsigned long long)) == cudaSuccess)
      cudaMemset(g_tile_prof, 0, 16 * sizeof(unsigned long long));
    else
      g_tile_prof = nullptr;
  }
  return g_tile_prof;
}

bool tile_prof_read(unsigned long long out[16]) {
  if (!g_tile_prof) return false;
  cudaDeviceSynchronize();
  cudaMemcpy(out, g_tile_prof, 16 * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
  cudaMemset(g_tile_prof, 0, 16 * sizeof(unsigned long long));
  return true;
}

bool recon_tile_compiled(const SchemeConst &sc, int deg_hi, int deg_lo) {
  // tile records serve the tile kernel and the cooperative kernel: every family of the specialised shape
  if (deg_lo != 1 || deg_hi < 1) return false;
  return (sc.n_dims == 2 && deg_hi <= 4) || (sc.n_dims == 3 && deg_hi <= 3);
}

}  // namespace zfvm
