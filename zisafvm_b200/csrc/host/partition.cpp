// Graph partition of the cells with METIS (k-way, communication-volume objective), as the reference computes it:
//   compute_partition_full_stencil     src/zisa/parallelization/domain_decomposition.cpp:27-113
//   compute_cell_permutation           :160-176   (cells sorted by part)
//   compute_partition_boundaries       :178-198
//
// METIS itself is third-party library code: the image carries METIS 5.1.0 as a static archive next to the CUDA
// libraries (libmetis_static.a, 64-bit idx_t) but not metis.h, so the one entry point and the option slots used here are
// declared by hand from the published 5.1.0 API (the slots were checked against the library's own debug print-out of
// its run-time parameters).  Built without the archive (ZFVM_HAS_METIS == 0) the entry point reports an error, like the
// reference built with ZISA_HAS_METIS == 0 (domain_decomposition.cpp:108-110).
#include <algorithm>
#include <cstdint>
#include <string>
#include <vector>

#include "zfvm_host.hpp"

#if ZFVM_HAS_METIS
extern "C" {
int METIS_SetDefaultOptions(std::int64_t *options);
int METIS_PartGraphKway(std::int64_t *nvtxs, std::int64_t *ncon, std::int64_t *xadj, std::int64_t *adjncy, std::int64_t *vwgt,
                        std::int64_t *vsize, std::int64_t *adjwgt, std::int64_t *nparts, float *tpwgts, float *ubvec,
                        std::int64_t *options, std::int64_t *objval, std::int64_t *part);
}
#endif

namespace zfvm {

bool metis_available() { return ZFVM_HAS_METIS != 0; }

bool partition_kway(std::vector<i32> &part, const HostGrid &g, const HostStencils *stencils, int n_parts,
                    std::string &err) {
  const i64 n = g.n_cells;
  part.assign((size_t)n, 0);
  if (n_parts < 1) {
    err = "partition_kway: n_parts must be positive";
    return false;
  }
  if (n_parts == 1) return true;
#if ZFVM_HAS_METIS
  // the graph: an edge between i and every member of its (combined) stencil, or between face neighbours when no
  // stencils are given; symmetric, every edge once per direction, neighbours sorted (:52-71)
  std::vector<std::vector<i32>> graph((size_t)n);
  const int F = g.max_neighbours;
  for (i64 i = 0; i < n; ++i) {
    auto link = [&](i32 j) {
      if (j == (i32)i) return;
      auto &gi = graph[(size_t)i];
      if (std::find(gi.begin(), gi.end(), j) == gi.end()) {
        gi.push_back(j);
        graph[(size_t)j].push_back((i32)i);
      }
    };
    if (stencils == nullptr) {
      for (int k = 0; k < F; ++k)
        if (g.is_valid(i, k)) link(g.neighbours[(size_t)(i * F + k)]);
    } else {
      const i32 *l2g = &stencils->l2g[(size_t)(i * stencils->l2g_stride)];
      for (int a = 0; a < stencils->l2g_size[(size_t)i]; ++a) link(l2g[a]);
    }
  }
  std::vector<std::int64_t> xadj((size_t)n + 1, 0), adjncy;
  for (i64 i = 0; i < n; ++i) {
    auto &gi = graph[(size_t)i];
    std::sort(gi.begin(), gi.end());
    gi.erase(std::unique(gi.begin(), gi.end()), gi.end());
    xadj[(size_t)i + 1] = xadj[(size_t)i] + (std::int64_t)gi.size();
  }
  adjncy.reserve((size_t)xadj[(size_t)n]);
  for (i64 i = 0; i < n; ++i)
    for (i32 j : graph[(size_t)i]) adjncy.push_back(j);
  std::vector<std::vector<i32>>().swap(graph);

  std::int64_t nvtxs = n, ncon = 1, nparts = n_parts, objval = -1;
  std::vector<std::int64_t> p64((size_t)n, 0);
  std::int64_t options[40];  // METIS_NOPTIONS
  METIS_SetDefaultOptions(options);
  options[1] = 1;     // METIS_OPTION_OBJTYPE = METIS_OBJTYPE_VOL     (:94)
  options[7] = 10;    // METIS_OPTION_NCUTS                            (:95)
  options[6] = 20;    // METIS_OPTION_NITER                            (:96)
  options[16] = 100;  // METIS_OPTION_UFACTOR                          (:97)
  const int rc = METIS_PartGraphKway(&nvtxs, &ncon, xadj.data(), adjncy.data(), nullptr, nullptr, nullptr, &nparts, nullptr,
                                     nullptr, options, &objval, p64.data());
  if (rc != 1) {  // METIS_OK
    err = "partition_kway: METIS_PartGraphKway failed with code " + std::to_string(rc);
    return false;
  }
  for (i64 i = 0; i < n; ++i) part[(size_t)i] = (i32)p64[(size_t)i];
  return true;
#else
  (void)stencils;
  err = "partition_kway: this build has no METIS (libmetis_static.a was not found next to the CUDA libraries); use the "
        "space-filling-curve partition";
  return false;
#endif
}

}  // namespace zfvm
