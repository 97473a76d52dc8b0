// Synthetic meshes (the reference ships no grids) and Hilbert renumbering
// (src/renumber_grid.cpp:46-135; the curve itself lives in the external ZisaSFC, so the
// standard Skilling transform is used -- any Hilbert orientation gives the same locality).
#include <algorithm>
#include <cmath>
#include <limits>
#include <numeric>
#if defined(_OPENMP)
#include <parallel/algorithm>
#endif

#include "zfvm_host.hpp"

namespace zfvm {

namespace {
inline std::uint64_t splitmix64(std::uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}
// uniform in [-1, 1)
inline double urand(std::uint64_t key) { return (double)(splitmix64(key) >> 11) * (2.0 / 9007199254740992.0) - 1.0; }
}  // namespace

RawMesh make_square_mesh(int nx, int ny, double x0, double x1, double y0, double y1, double jitter,
                         std::uint64_t seed) {
  RawMesh m;
  m.n_dims = 2;
  const double hx = (x1 - x0) / nx, hy = (y1 - y0) / ny;
  const i64 nvx = nx + 1, nvy = ny + 1;
  m.vertices.resize((size_t)(3 * nvx * nvy));
  for (i64 iy = 0; iy < nvy; ++iy)
    for (i64 ix = 0; ix < nvx; ++ix) {
      i64 v = iy * nvx + ix;
      double x = x0 + hx * ix, y = y0 + hy * iy;
      if (ix > 0 && ix < nx && iy > 0 && iy < ny) {
        x += jitter * hx * urand(seed * 1000003ull + 2 * (std::uint64_t)v);
        y += jitter * hy * urand(seed * 1000003ull + 2 * (std::uint64_t)v + 1);
      }
      m.vertices[3 * v] = x;
      m.vertices[3 * v + 1] = y;
      m.vertices[3 * v + 2] = 0.0;
    }
  m.vertex_indices.reserve((size_t)(6 * (i64)nx * ny));
  for (i64 iy = 0; iy < ny; ++iy)
    for (i64 ix = 0; ix < nx; ++ix) {
      i32 a = (i32)(iy * nvx + ix), b = a + 1, c = (i32)((iy + 1) * nvx + ix), d = c + 1;
      if ((ix + iy) % 2 == 0) {
        m.vertex_indices.insert(m.vertex_indices.end(), {a, b, d});
        m.vertex_indices.insert(m.vertex_indices.end(), {a, d, c});
      } else {
        m.vertex_indices.insert(m.vertex_indices.end(), {a, b, c});
        m.vertex_indices.insert(m.vertex_indices.end(), {b, d, c});
      }
    }
  return m;
}

RawMesh make_cube_mesh(int nx, int ny, int nz, double h, double x0, double y0, double z0, double jitter,
                       std::uint64_t seed, int ox, int oy, int oz, int gx, int gy, int gz) {
  if (gx < 0) gx = nx;
  if (gy < 0) gy = ny;
  if (gz < 0) gz = nz;
  RawMesh m;
  m.n_dims = 3;
  const i64 nvx = nx + 1, nvy = ny + 1, nvz = nz + 1;
  m.vertices.resize((size_t)(3 * nvx * nvy * nvz));
#pragma omp parallel for schedule(static)
  for (i64 iz = 0; iz < nvz; ++iz)
    for (i64 iy = 0; iy < nvy; ++iy)
      for (i64 ix = 0; ix < nvx; ++ix) {
        const i64 v = (iz * nvy + iy) * nvx + ix;
        const i64 Gx = ox + ix, Gy = oy + iy, Gz = oz + iz;  // global lattice position
        double x = x0 + h * Gx, y = y0 + h * Gy, z = z0 + h * Gz;
        if (Gx > 0 && Gx < gx && Gy > 0 && Gy < gy && Gz > 0 && Gz < gz) {
          const std::uint64_t gid = (std::uint64_t)((Gz * (gy + 1) + Gy) * (gx + 1) + Gx);
          x += jitter * h * urand(seed * 1000003ull + 3 * gid);
          y += jitter * h * urand(seed * 1000003ull + 3 * gid + 1);
          z += jitter * h * urand(seed * 1000003ull + 3 * gid + 2);
        }
        m.vertices[3 * v] = x;
        m.vertices[3 * v + 1] = y;
        m.vertices[3 * v + 2] = z;
      }
  static const int perms[6][3] = {{0, 1, 2}, {0, 2, 1}, {1, 0, 2}, {1, 2, 0}, {2, 0, 1}, {2, 1, 0}};
  const i64 nc = (i64)nx * ny * nz;
  m.vertex_indices.resize((size_t)(24 * nc));
#pragma omp parallel for schedule(static)
  for (i64 c = 0; c < nc; ++c) {
    const i64 ix = c % nx, iy = (c / nx) % ny, iz = c / ((i64)nx * ny);
    for (int p = 0; p < 6; ++p) {
      i64 pos[3] = {ix, iy, iz};
      i32 *t = &m.vertex_indices[(size_t)(24 * c + 4 * p)];
      t[0] = (i32)((pos[2] * nvy + pos[1]) * nvx + pos[0]);
      for (int s = 0; s < 3; ++s) {
        pos[perms[p][s]] += 1;
        t[s + 1] = (i32)((pos[2] * nvy + pos[1]) * nvx + pos[0]);
      }
    }
  }
  return m;
}

void renumber_mesh_cells(RawMesh &m, const std::vector<i32> &perm) {
  const int F = m.n_dims + 1;
  std::vector<i32> vi(m.vertex_indices.size());
  const i64 n = (i64)perm.size();
#pragma omp parallel for schedule(static)
  for (i64 i = 0; i < n; ++i)
    for (int k = 0; k < F; ++k) vi[(size_t)(i * F + k)] = m.vertex_indices[(size_t)((i64)perm[(size_t)i] * F + k)];
  m.vertex_indices.swap(vi);
}

namespace {
// Skilling, "Programming the Hilbert curve" (2004): axes -> transposed Hilbert index.
std::uint64_t hilbert_key(std::uint32_t *X, int n, int bits) {
  std::uint32_t M = 1u << (bits - 1), P, Q, t;
  for (Q = M; Q > 1; Q >>= 1) {
    P = Q - 1;
    for (int i = 0; i < n; ++i) {
      if (X[i] & Q)
        X[0] ^= P;
      else {
        t = (X[0] ^ X[i]) & P;
        X[0] ^= t;
        X[i] ^= t;
      }
    }
  }
  for (int i = 1; i < n; ++i) X[i] ^= X[i - 1];
  t = 0;
  for (Q = M; Q > 1; Q >>= 1)
    if (X[n - 1] & Q) t ^= Q - 1;
  for (int i = 0; i < n; ++i) X[i] ^= t;
  std::uint64_t key = 0;
  for (int b = bits - 1; b >= 0; --b)
    for (int i = 0; i < n; ++i) key = (key << 1) | ((X[i] >> b) & 1u);
  return key;
}
}  // namespace

std::vector<i32> hilbert_permutation(int n_dims, i64 n, const double *centers) {
  double lo[3] = {std::numeric_limits<double>::max(), std::numeric_limits<double>::max(),
                  std::numeric_limits<double>::max()};
  double hi[3] = {-lo[0], -lo[0], -lo[0]};
  for (i64 i = 0; i < n; ++i)
    for (int d = 0; d < n_dims; ++d) {
      lo[d] = std::min(lo[d], centers[3 * i + d]);
      hi[d] = std::max(hi[d], centers[3 * i + d]);
    }
  const int bits = n_dims == 2 ? 31 : 21;
  const double scale = (double)((1u << bits) - 1);
  std::vector<std::uint64_t> keys((size_t)n);
#pragma omp parallel for schedule(static)
  for (i64 i = 0; i < n; ++i) {
    std::uint32_t X[3] = {0, 0, 0};
    for (int d = 0; d < n_dims; ++d) {
      double u = (centers[3 * i + d] - lo[d]) / (hi[d] - lo[d] + 1e-10 * std::abs(hi[d]) + 1e-300);
      X[d] = (std::uint32_t)(u * scale);
    }
    keys[(size_t)i] = hilbert_key(X, n_dims, bits);
  }
  std::vector<i32> perm((size_t)n);
  std::iota(perm.begin(), perm.end(), 0);
  auto cmp = [&](i32 a, i32 b) {
    return keys[(size_t)a] != keys[(size_t)b] ? keys[(size_t)a] < keys[(size_t)b] : a < b;
  };
#if defined(_OPENMP)
  __gnu_parallel::sort(perm.begin(), perm.end(), cmp);
#else
  std::sort(perm.begin(), perm.end(), cmp);
#endif
  return perm;
}

}  // namespace zfvm
