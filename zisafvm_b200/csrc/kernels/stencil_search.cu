// P2 `stencil_search_kernel`: the stencil families of all cells, selected on the device (reference:
// compute_stencil_families, src/zisa/reconstruction/stencil_family.cpp:99-117, Stencil ctors stencil.cpp:42-80 with
// central_stencil / biased_stencil :347-399).  One thread per cell walks the reference's algorithm for every stencil of
// the family: region-grown candidates (breadth first over face neighbours, at most 5 n_points expansions, a cell enters
// when its centre or one of its query points lies in the region), the n_points closest by centre distance, and for the
// one-sided stencils the rank test of the least-squares matrix, first with the cone at the off vertex, then at the centre.
// The decisions are taken by the code the host search uses (host/stencil_shared.hpp, no fused multiply-adds here), so
// the members are the host's.  Whatever the reference leaves to its implementation or to chance is not decided here:
// equal distances among the kept candidates (std::sort's permutation), the random retries of tryhard_stencil
// (stencil.cpp:303-345) and the rare overflow of the per-thread buffers flag the cell, and the host search redoes it.
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "../host/stencil_shared.hpp"
#include "../host/zfvm_host.hpp"

namespace zfvm {

namespace {

constexpr int SEARCH_CMAX = 640;   // candidates of one region search (central 3D order 4: up to ~590)
constexpr int SEARCH_HASH = 1024;  // open-addressing set of the cells met in one region search
constexpr int SEARCH_TOP = 66;     // n_points + 1 closest candidates kept sorted
constexpr int SEARCH_AMAX = 256;   // entries of the least-squares matrix of the rank test

struct SearchParams {
  int ns;
  int orders[6], biased[6], max_size[6], local_off[7];
  int L;
};

struct Scratch {
  int cands[SEARCH_CMAX];
  int hash[SEARCH_HASH];
  double top_d[SEARCH_TOP];
  int top_i[SEARCH_TOP];
  double U[SEARCH_AMAX];
};

/// true if `c` was met before in this search; records it otherwise
__device__ inline bool seen_before(int *hash, int c) {
  unsigned h = ((unsigned)c * 2654435761u) >> 22;  // 10 bits
  for (int probe = 0; probe < SEARCH_HASH; ++probe) {
    const int v = hash[h];
    if (v == c) return true;
    if (v < 0) {
      hash[h] = c;
      return false;
    }
    h = (h + 1) & (SEARCH_HASH - 1);
  }
  return false;  // (cannot fill up: at most SEARCH_CMAX + tested cells entries, checked by the caller)
}

/// region_based_stencil (stencil.cpp:192-256): returns the number of members written to `out` (<= n_points), or -1 when
/// the cell has to be redone on the host (buffer overflow, equal distances among the kept candidates).
__device__ int region_stencil(const sel::GridView &g, const sel::Cone &region, int i_center, int n_points, Scratch &w,
                              int *out) {
  for (int a = 0; a < SEARCH_HASH; ++a) w.hash[a] = -1;
  const int max_points = 5 * n_points;
  int n = 0, met = 1;
  w.cands[n++] = i_center;
  seen_before(w.hash, i_center);
  for (int p = 0; p < max_points; ++p) {
    if (p >= n) break;
    const long long j = w.cands[p];
    for (int k = 0; k < g.F; ++k) {
      const int c = g.nb[j * g.F + k];
      if (c < 0) continue;
      if (seen_before(w.hash, c)) continue;
      if (++met > SEARCH_HASH - 64) return -1;
      if (sel::cell_inside(g, region, c)) {
        if (n >= SEARCH_CMAX) return -1;
        w.cands[n++] = c;
      }
    }
  }
  // the n_points + 1 closest, sorted; equal distances among them: the order is std::sort's business
  const int keep = n < n_points + 1 ? n : n_points + 1;
  const sel::V3 xc = sel::center(g, i_center);
  int kept = 0;
  for (int a = 0; a < n; ++a) {
    const double d = sel::norm(sel::sub(sel::center(g, w.cands[a]), xc));
    if (kept == keep && !(d < w.top_d[kept - 1])) {
      if (d == w.top_d[kept - 1]) return -1;
      continue;
    }
    int pos = kept < keep ? kept : keep - 1;  // slot that becomes free (the current last entry drops out when full)
    while (pos > 0 && d < w.top_d[pos - 1]) {
      w.top_d[pos] = w.top_d[pos - 1];
      w.top_i[pos] = w.top_i[pos - 1];
      --pos;
    }
    if (pos > 0 && d == w.top_d[pos - 1]) return -1;
    w.top_d[pos] = d;
    w.top_i[pos] = w.cands[a];
    if (kept < keep) ++kept;
  }
  const int m = kept < n_points ? kept : n_points;
  for (int a = 0; a < m; ++a) out[a] = w.top_i[a];
  return m;
}

__device__ bool is_good(const sel::GridView &g, const int *s, int n, int order, Scratch &w, bool &unsupported) {
  const int rows = n - 1, cols = lsq::dof(order - 1, g.nd) - 1;
  if (order <= 1) return true;  // a 1 x 1 matrix of ones
  if (rows * cols > SEARCH_AMAX || cols > 34) {
    unsupported = true;
    return false;
  }
  sel::assemble_matrix(g, s, n, order, w.U, cols);
  return sel::matrix_rank_inplace(w.U, rows, cols) == cols;
}

__global__ void __launch_bounds__(128) stencil_search_kernel(const sel::GridView g, const SearchParams p,
                                                             const unsigned char *__restrict__ flags, int *members,
                                                             int *count, unsigned char *redo, Scratch *scratch) {
  Scratch &w = scratch[(long long)blockIdx.x * blockDim.x + threadIdx.x];
  const long long n_threads = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < g.n_cells; i += n_threads) {
    int *mem = members + i * p.L;
    int *cnt = count + i * p.ns;
    redo[i] = 0;
    for (int k = 0; k < p.ns; ++k) cnt[k] = 0;
    const bool full = (flags[i] & FLAG_INTERIOR) || (flags[i] & FLAG_GHOST_L1);
    if (!full) {  // StencilFamilyParams{{1}, {"c"}, {1.0}}, stencil_family.cpp:99-117
      mem[0] = (int)i;
      cnt[0] = 1;
      continue;
    }
    int k_biased = 0;
    bool bad = false;
    for (int k = 0; k < p.ns && !bad; ++k) {
      int *out = mem + p.local_off[k];
      const int n_points = p.max_size[k];
      int m;
      if (p.biased[k]) {  // biased_stencil, stencil.cpp:347-393
        const int kf = k_biased++;
        bool unsupported = false;
        const sel::Cone r1 = sel::make_cone(g, i, sel::vertex(g, i, sel::rel_off_vertex(g.nd, kf)), kf);
        m = region_stencil(g, r1, (int)i, n_points, w, out);
        if (m < 0) {
          bad = true;
          break;
        }
        if (!(m == n_points && is_good(g, out, n_points, p.orders[k], w, unsupported))) {
          if (unsupported) {
            bad = true;
            break;
          }
          const sel::Cone r2 = sel::make_cone(g, i, sel::center(g, i), kf);
          m = region_stencil(g, r2, (int)i, n_points, w, out);
          if (m < 0 || !(m == n_points && is_good(g, out, n_points, p.orders[k], w, unsupported))) {
            bad = true;  // tryhard_stencil: the host's business
            break;
          }
        }
      } else {
        m = region_stencil(g, sel::full_sphere(), (int)i, n_points, w, out);
        if (m < 0) {
          bad = true;
          break;
        }
      }
      cnt[k] = m;
    }
    if (bad) redo[i] = 1;
  }
}

template <class T>
struct DevBuf {
  T *p = nullptr;
  cudaError_t alloc(size_t count) { return cudaMalloc((void **)&p, std::max<size_t>(count, 1) * sizeof(T)); }
  cudaError_t upload(const T *host, size_t count) {
    cudaError_t e = alloc(count);
    if (e != cudaSuccess) return e;
    return cudaMemcpy(p, host, count * sizeof(T), cudaMemcpyHostToDevice);
  }
  ~DevBuf() {
    if (p) cudaFree(p);
  }
};

}  // namespace

bool device_stencil_search(const HostGrid &g, const StencilFamilyParams &params, DeviceStencilSearch &out, std::string &why) {
  int n_dev = 0;
  if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0) {
    cudaGetLastError();
    why = "no CUDA device";
    return false;
  }
  const int ns = params.n_stencils();
  if (ns > 6) {
    why = "more than six stencils";
    return false;
  }
  SearchParams p{};
  p.ns = ns;
  p.local_off[0] = 0;
  for (int k = 0; k < ns; ++k) {
    p.orders[k] = params.orders[(size_t)k];
    p.biased[k] = params.biases[(size_t)k];
    p.max_size[k] = required_stencil_size(params.orders[(size_t)k] - 1, params.overfit_factors[(size_t)k], g.n_dims);
    p.local_off[k + 1] = p.local_off[k] + p.max_size[k];
    if (p.max_size[k] + 1 > SEARCH_TOP) {
      why = "stencil too large for the device search";
      return false;
    }
  }
  p.L = p.local_off[ns];
  const std::int64_t n = g.n_cells;
  const bool verbose = std::getenv("ZFVM_VERBOSE") != nullptr;
  auto t_last = std::chrono::steady_clock::now();
  auto lap = [&](const char *what) {
    if (!verbose) return;
    const auto t1 = std::chrono::steady_clock::now();
    std::fprintf(stderr, "[zfvm stencils]   %-24s %7.2f s\n", what, std::chrono::duration<double>(t1 - t_last).count());
    t_last = t1;
  };
  sel::GridView v = make_grid_view(g);
  DevBuf<int> nb, vi, members, count;
  DevBuf<double> vtx, cc, len, mom;
  DevBuf<unsigned char> flags, redo;
  DevBuf<Scratch> scratch;
  int dev = 0, n_sm = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
  const int threads = 128;
  const int blocks = (int)std::min<std::int64_t>((n + threads - 1) / threads, (std::int64_t)n_sm * 8);
  auto check = [&](cudaError_t e, const char *what) {
    if (e == cudaSuccess) return true;
    why = std::string(what) + ": " + cudaGetErrorString(e);
    cudaGetLastError();
    return false;
  };
  if (!check(nb.upload(g.neighbours.data(), g.neighbours.size()), "neighbours") ||
      !check(vi.upload(g.vertex_indices.data(), g.vertex_indices.size()), "vertex_indices") ||
      !check(vtx.upload(g.vertices.data(), g.vertices.size()), "vertices") ||
      !check(cc.upload(g.cell_centers.data(), g.cell_centers.size()), "cell_centers") ||
      !check(len.upload(g.characteristic_length.data(), g.characteristic_length.size()), "characteristic_length") ||
      !check(mom.upload(g.moments.data(), g.moments.size()), "moments") ||
      !check(flags.upload(g.cell_flags.data(), g.cell_flags.size()), "cell_flags") ||
      !check(members.alloc((size_t)(n * p.L)), "members") || !check(count.alloc((size_t)(n * ns)), "count") ||
      !check(redo.alloc((size_t)n), "redo") || !check(scratch.alloc((size_t)blocks * threads), "scratch"))
    return false;
  v.nb = nb.p;
  v.vi = vi.p;
  v.vtx = vtx.p;
  v.cc = cc.p;
  v.len = len.p;
  v.mom = mom.p;
  v.face_c = nullptr;
  v.edge = nullptr;
  lap("upload");
  if (!check(cudaMemset(members.p, 0xFF, (size_t)(n * p.L) * sizeof(int)), "memset")) return false;
  stencil_search_kernel<<<blocks, threads>>>(v, p, flags.p, members.p, count.p, redo.p, scratch.p);
  if (!check(cudaGetLastError(), "stencil_search_kernel launch") || !check(cudaDeviceSynchronize(), "stencil_search_kernel"))
    return false;
  lap("search kernel");
  out.L = p.L;
  parallel_assign(out.members, (size_t)(n * p.L), (i32)INVALID);
  out.count.resize((size_t)(n * ns));
  out.redo.resize((size_t)n);
  if (!check(cudaMemcpy(out.members.data(), members.p, out.members.size() * sizeof(int), cudaMemcpyDeviceToHost), "members") ||
      !check(cudaMemcpy(out.count.data(), count.p, out.count.size() * sizeof(int), cudaMemcpyDeviceToHost), "count") ||
      !check(cudaMemcpy(out.redo.data(), redo.p, out.redo.size(), cudaMemcpyDeviceToHost), "redo"))
    return false;
  lap("download");
  return true;
}

}  // namespace zfvm
