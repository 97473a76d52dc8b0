// P1 `lsq_weights_kernel`: the stencil weights W_k = pinv(A_k) of every cell, built on the device and written straight
// into the tile records (reference: LSQSolver ctor, src/zisa/reconstruction/lsq_solver.cpp:40-47 with
// assemble_weno_ao_matrix :168-403, run once per stencil by StencilFamily / GlobalReconstruction construction).
//
// A warp owns a tile, a lane a cell: the lane reads its stencil members from the record header the host has already
// uploaded (row list + local index rows, or the global index rows of the plainer record), assembles A from the members'
// centres, lengths and normalised moments, factorises it by Householder reflections and writes the rows of
// W = R^{-1} Q^T lane-interleaved where K1 streams them from.  The arithmetic is host/lsq_shared.hpp, compiled here
// with -fmad=false: the result is bit-identical to the host path (ZFVM_PRECOMPUTE=host), which stays as the cross-check.
// A cell's working set (R with the reflectors in its lower part, |v|^2, the unit column) lives in shared memory,
// [element][thread], when 32 cells' worth fits; in an L2-resident global scratch otherwise.
#include <algorithm>
#include <cstdio>

#include "../host/lsq_shared.hpp"
#include "kernels.hpp"

namespace zfvm {

namespace {

constexpr int LSQ_MAX_COLS = 34;  // 3D order 5 would have 34 columns; the matrices exist up to 19 (3D order 4)

__global__ void lsq_weights_kernel(LsqWeightArgs a) {
  extern __shared__ double lsq_smem[];
  const int lane = threadIdx.x & 31;
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
  double *base;
  long long stride;
  if (a.scratch == nullptr) {
    base = lsq_smem + threadIdx.x;
    stride = blockDim.x;
  } else {
    stride = (long long)gridDim.x * blockDim.x;
    base = a.scratch + (long long)blockIdx.x * blockDim.x + threadIdx.x;
  }
  const int nd = a.n_dims;
  for (long long t = a.tile_begin + warp; t < a.tile_end; t += n_warps) {
    char *rec = a.rec + t * a.rec_bytes;
    const unsigned long long meta = reinterpret_cast<const unsigned long long *>(rec + a.view.off_meta)[lane];
    const long long i0 = t * TILE + lane;
    for (int k = 0; k < a.n_stencils; ++k) {
      const int rows = (int)((meta >> (8 * k)) & 0xFFull);
      if (rows == 0) continue;  // unused slot, first-order stencil or padding lane
      int order = 0;
      for (int o = 2; o <= a.max_order[k]; ++o)
        if (a.rows_of_order[k][o] == rows) order = o;
      if (order == 0) continue;
      const int cols = lsq::dof(order - 1, nd) - 1;
      const lsq::Strided R{base, stride}, vk{base + (long long)rows * cols * stride, stride},
          vn{base + ((long long)rows * cols + cols) * stride, stride},
          y{base + ((long long)rows * cols + 2 * cols) * stride, stride};
      const double x0 = a.centers[3 * i0], y0 = a.centers[3 * i0 + 1], z0 = a.centers[3 * i0 + 2];
      const double l0 = a.length[i0];
      const double *C0 = a.moments + i0 * a.n_moments;
      for (int j = 0; j < rows; ++j) {
        long long g;
        if (a.view.tile_record) {
          const int row = a.view.row0[k] + j;
          const int li = a.view.lidx_elem == 1
                             ? (int)reinterpret_cast<const unsigned char *>(rec + a.view.off_lidx)[row * TILE + lane]
                             : (int)reinterpret_cast<const unsigned short *>(rec + a.view.off_lidx)[row * TILE + lane];
          g = reinterpret_cast<const int *>(rec + a.view.off_list)[li];
        } else {
          g = reinterpret_cast<const int *>(rec + a.view.off_sidx[k])[j * TILE + lane];
        }
        double row[LSQ_MAX_COLS];
        for (int c = 0; c < cols; ++c) row[c] = 0.0;
        lsq::lsq_row(row, nd, order, (a.centers[3 * g] - x0) / l0, (a.centers[3 * g + 1] - y0) / l0,
                     (a.centers[3 * g + 2] - z0) / l0, a.length[g] / l0, C0, a.moments + g * a.n_moments);
        for (int c = 0; c < cols; ++c) R[j * cols + c] = row[c];
      }
      double *w = reinterpret_cast<double *>(rec + a.view.off_w[k]) + lane;
      const int NC = a.ncoef[k];
      lsq::pinv_householder(R, vk, vn, y, rows, cols,
                            [&](int i, int c, double v) { w[(long long)(c * NC + i) * TILE] = v; });
    }
  }
}

}  // namespace

std::int64_t lsq_scratch_doubles_per_cell(const LsqWeightArgs &a) {
  std::int64_t m = 0;
  for (int k = 0; k < a.n_stencils; ++k) {
    const int rows = a.rows_of_order[k][a.max_order[k]];
    const int cols = lsq::dof(a.max_order[k] - 1, a.n_dims) - 1;
    m = std::max<std::int64_t>(m, (std::int64_t)rows * cols + 2 * cols + rows);
  }
  return m;
}

int launch_lsq_weights(const LsqWeightArgs &args, int n_sms, double **scratch, std::int64_t *scratch_bytes,
                       cudaStream_t stream) {
  LsqWeightArgs a = args;
  for (int k = 0; k < a.n_stencils; ++k)
    if (lsq::dof(a.max_order[k] - 1, a.n_dims) - 1 > LSQ_MAX_COLS) return 1;
  const std::int64_t per_cell = lsq_scratch_doubles_per_cell(a) * (std::int64_t)sizeof(double);
  static bool attr_set = false;
  constexpr int SMEM_LIMIT = 200 * 1024;
  if (!attr_set) {
    if (cudaFuncSetAttribute(lsq_weights_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT) != cudaSuccess)
      return 2;
    attr_set = true;
  }
  int threads = (int)std::min<std::int64_t>(256, SMEM_LIMIT / std::max<std::int64_t>(per_cell, 1) / 32 * 32);
  const std::int64_t n_tiles = a.tile_end - a.tile_begin;
  if (n_tiles <= 0) return 0;
  if (threads >= 32) {  // shared-memory working set, one block per SM
    a.scratch = nullptr;
    const int blocks = (int)std::min<std::int64_t>(n_sms, (n_tiles + threads / 32 - 1) / (threads / 32));
    lsq_weights_kernel<<<blocks, threads, (size_t)(per_cell * threads), stream>>>(a);
  } else {  // large stencils: global scratch, kept small enough to stay in L2
    threads = 128;
    const int blocks = (int)std::min<std::int64_t>(2 * n_sms, (n_tiles + 3) / 4);
    const std::int64_t need = per_cell * threads * blocks;
    if (*scratch_bytes < need) {
      if (*scratch) cudaFree(*scratch);
      *scratch = nullptr;
      if (cudaMalloc(scratch, (size_t)need) != cudaSuccess) return 2;
      *scratch_bytes = need;
    }
    a.scratch = *scratch;
    lsq_weights_kernel<<<blocks, threads, 0, stream>>>(a);
  }
  return cudaGetLastError() == cudaSuccess ? 0 : 2;
}

}  // namespace zfvm
