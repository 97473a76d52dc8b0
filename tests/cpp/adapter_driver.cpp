// Plain C++ program that drives libzfvm_b200.so through include/zfvm_zisa_adapter.hpp the way ZisaFVM's time loop would
// (src/zisa/ode/time_loop.cpp:115-175: dt = cfl(u); u = time_integration->compute_step(u, t, dt); sanity check) and
// through the one-object swap (Sum[Zero, CudaEulerRateOfChange]).  TEST INFRASTRUCTURE: no Python, no ctypes on this side;
// tests/test_cpp_adapter.py writes the inputs, runs the binary and compares its outputs with the oracle.
//
//   adapter_driver run <in.bin> <out.bin>     needs a GPU
//   adapter_driver nodevice                   on a box without a GPU: the failure must arrive through LOG_ERR
//
// in.bin : i64 n_dims, n_vertices, n_cells, n_steps, order | f64 gamma, cfl | f64 vertices[nv][3] | i32 vertex_indices[nc][F]
//          | u8 cell_flags[nc] (bit0 interior, bit1 ghost_cell, bit2 ghost_cell_l1) | f64 u0[nc][5]
// out.bin: f64 tendency[nc][5] | f64 tendency_imported_stencils[nc][5] | f64 u_final[nc][5] | f64 dt[n_steps]
//          | f64 u_final_resident[nc][5]
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iostream>

#define ZFVM_ADAPTER_ZISA_MOCK "mock_zisa.hpp"
#include "zfvm_zisa_adapter.hpp"

using namespace zisa;

namespace {

template <class T>
void read_into(std::ifstream &in, T *dst, std::size_t count) {
  in.read(reinterpret_cast<char *>(dst), (std::streamsize)(count * sizeof(T)));
  if (!in) throw std::runtime_error("adapter_driver: short read");
}

HybridWENOParams weno_params(int n_dims, int order) {
  // the reference's parameter sets (test/zisa/unit_test/reconstruction/cweno_ao.cpp:48-55, lsq_solver.cpp:21-22)
  HybridWENOParams w;
  const int ns = n_dims + 2;
  w.stencil_family_params.orders.assign((std::size_t)ns, 2);
  w.stencil_family_params.orders[0] = order;
  w.stencil_family_params.biases.assign((std::size_t)ns, "b");
  w.stencil_family_params.biases[0] = "c";
  w.stencil_family_params.overfit_factors.assign((std::size_t)ns, 1.5);
  w.stencil_family_params.overfit_factors[0] = 2.0;
  w.linear_weights.assign((std::size_t)ns, 1.0);
  w.linear_weights[0] = 100.0;
  w.epsilon = 1e-10;
  w.exponent = 4.0;
  return w;
}

/// array<StencilFamily, 1> as the reference would hold it, rebuilt from the library's own selection
array<StencilFamily, 1> families_of(const b200::Context &c, int_t n_cells, int ns) {
  auto get = [&](const char *name) {
    const void *data;
    int dtype, ndim;
    std::int64_t shape[4];
    b200::check(zfvm_stencils_get(c.stencils(), name, &data, &dtype, &ndim, shape));
    return std::make_pair(static_cast<const std::int32_t *>(data), shape[ndim - 1]);
  };
  const auto [l2g, L] = get("l2g");
  const auto local = get("local").first;
  const auto local_off = get("local_off").first;
  const auto order = get("order").first;
  const auto size = get("size").first;
  const auto n_family = get("n_family").first;
  array<StencilFamily, 1> out(n_cells);
  for (int_t i = 0; i < n_cells; ++i) {
    std::vector<Stencil> st;
    for (int k = 0; k < n_family[i]; ++k) {
      std::vector<int_t> glob;
      for (int j = 0; j < size[i * ns + k]; ++j) glob.push_back((int_t)l2g[i * L + local[i * L + local_off[k] + j]]);
      st.emplace_back(glob, order[i * ns + k]);
    }
    out[i] = StencilFamily(st);
  }
  return out;
}

int run(const char *in_path, const char *out_path) {
  std::ifstream in(in_path, std::ios::binary);
  if (!in) throw std::runtime_error("adapter_driver: cannot open the input file");
  std::int64_t head[5];
  double par[2];
  read_into(in, head, 5);
  read_into(in, par, 2);
  const int n_dims = (int)head[0], order = (int)head[4];
  const int_t nv = (int_t)head[1], nc = (int_t)head[2], n_steps = (int_t)head[3], F = (int_t)n_dims + 1;

  Grid grid;
  grid.n_cells = nc;
  grid.n_vertices = nv;
  grid.max_neighbours = F;
  grid.vertices = array<XYZ, 1>(nv);
  grid.vertex_indices = array<int_t, 2>(nc, F);
  grid.cell_flags = array<CellFlags, 1>(nc);
  read_into(in, reinterpret_cast<double *>(grid.vertices.raw()), nv * 3);
  {
    std::vector<std::int32_t> vi(nc * F);
    read_into(in, vi.data(), vi.size());
    for (int_t a = 0; a < nc * F; ++a) grid.vertex_indices[a] = (int_t)vi[a];
    std::vector<std::uint8_t> flags(nc);
    read_into(in, flags.data(), nc);
    for (int_t i = 0; i < nc; ++i) {  // what mask_ghost_cells left behind, src/zisa/grid/grid.cpp:1122-1136
      grid.cell_flags[i].interior = (flags[i] & 1) != 0;
      grid.cell_flags[i].ghost_cell = (flags[i] & 2) != 0;
      grid.cell_flags[i].ghost_cell_l1 = (flags[i] & 4) != 0;
    }
  }
  auto u0 = std::make_shared<AllVariables>(AllVariablesDimensions{nc, 5, 0});
  read_into(in, u0->cvars.raw(), nc * 5);

  const HybridWENOParams weno = weno_params(n_dims, order);
  zfvm_params p = b200::make_params(weno);
  p.gamma = par[0];
  const QRDegrees qr{3, 3, 4};
  auto context = std::make_shared<b200::Context>(grid, qr, weno, p, /*device=*/0);

  std::ofstream out(out_path, std::ios::binary);
  auto write = [&](const double *src, std::size_t count) {
    out.write(reinterpret_cast<const char *>(src), (std::streamsize)(count * sizeof(double)));
  };

  // ---- one-object swap: aggregate_rates_of_change = Sum[Zero, fvm] (numerical_experiment.cpp:238-256) ---------------
  AllVariables tendency(u0->dims());
  {
    SumRatesOfChange sum;
    sum.add_term(std::make_shared<ZeroRateOfChange>());
    sum.add_term(std::make_shared<b200::CudaEulerRateOfChange>(context));
    for (double &x : tendency.cvars) x = 123.0;  // Zero must wipe it, the device term accumulates
    sum.compute(tendency, *u0, 0.0);
    write(tendency.cvars.raw(), nc * 5);
  }
  // ---- the same on stencil families handed over as the reference holds them -------------------------------------------
  {
    const array<StencilFamily, 1> fam = families_of(*context, nc, n_dims + 2);
    auto imported = std::make_shared<b200::Context>(grid, qr, weno, p, 0, &fam);
    AllVariables t2(u0->dims());
    ZeroRateOfChange().compute(t2, *u0, 0.0);
    b200::CudaEulerRateOfChange(imported).compute(t2, *u0, 0.0);
    write(t2.cvars.raw(), nc * 5);
  }
  // ---- full swap: the reference's time loop over the adapter classes ---------------------------------------------------
  std::vector<double> dts;
  for (int pass = 0; pass < 2; ++pass) {  // pass 0: host state refreshed every step; pass 1: resident, one download
    auto rk = std::make_shared<b200::CudaRungeKutta>(context, "ssp3", par[1], /*download_every_step=*/pass == 0);
    std::shared_ptr<BoundaryCondition> bc = std::make_shared<b200::CudaFrozenBC>(context, *u0);
    std::shared_ptr<CFLCondition> cfl = std::make_shared<b200::CudaCFL>(rk);
    std::shared_ptr<SanityCheck> sane = std::make_shared<b200::CudaSanityCheck>(rk);
    std::shared_ptr<TimeIntegration> ti = rk;
    std::shared_ptr<AllVariables> u = u0;
    double t = 0.0;
    for (int_t s = 0; s < n_steps; ++s) {
      const double dt = (*cfl)(*u);
      u = ti->compute_step(u, t, dt);
      bc->apply(*u, t + dt);
      LOG_ERR_IF(!(*sane)(*u), "adapter_driver: implausible state");
      t += dt;
      if (pass == 0) dts.push_back(dt);
    }
    if (pass == 1) rk->download(*u);
    write(u->cvars.raw(), nc * 5);
    if (pass == 0) write(dts.data(), dts.size());
  }
  std::printf("adapter_driver: %s; %zu cells, %zu steps\n", b200::CudaEulerRateOfChange(context).str().c_str(), nc, n_steps);
  return out ? 0 : 1;
}

int nodevice() {
  // two triangles; creation must fail with the library's message, delivered through LOG_ERR (the mock throws)
  Grid grid;
  grid.n_cells = 2;
  grid.n_vertices = 4;
  grid.max_neighbours = 3;
  grid.vertices = array<XYZ, 1>(4);
  const double xy[4][2] = {{0, 0}, {1, 0}, {1, 1}, {0, 1}};
  for (int_t i = 0; i < 4; ++i) grid.vertices[i] = XYZ{{xy[i][0], xy[i][1], 0.0}};
  grid.vertex_indices = array<int_t, 2>(2, 3);
  const int_t vi[6] = {0, 1, 3, 1, 2, 3};
  for (int_t a = 0; a < 6; ++a) grid.vertex_indices[a] = vi[a];
  grid.cell_flags = array<CellFlags, 1>(2);
  HybridWENOParams weno = weno_params(2, 2);
  zfvm_params p = b200::make_params(weno);
  try {
    b200::Context c(grid, QRDegrees{2, 2, 2}, weno, p, 0);
  } catch (const std::runtime_error &e) {
    std::printf("LOG_ERR: %s\n", e.what());
    return std::strstr(e.what(), "CUDA") != nullptr ? 0 : 2;
  }
  std::printf("a context was created: this box has a GPU\n");
  return 0;
}

}  // namespace

int main(int argc, char **argv) {
  try {
    if (argc == 2 && std::strcmp(argv[1], "nodevice") == 0) return nodevice();
    if (argc == 4 && std::strcmp(argv[1], "run") == 0) return run(argv[2], argv[3]);
  } catch (const std::exception &e) {
    std::fprintf(stderr, "adapter_driver failed: %s\n", e.what());
    return 1;
  }
  std::fprintf(stderr, "usage: adapter_driver run <in.bin> <out.bin> | adapter_driver nodevice\n");
  return 64;
}
