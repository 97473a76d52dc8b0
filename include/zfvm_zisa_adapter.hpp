// zfvm_zisa_adapter.hpp -- the ZisaFVM-side adapter of libzfvm_b200.so: header-only, C++17.
//
// A ZisaFVM maintainer adds this one header to the tree (e.g. as zisa/cuda/zfvm_adapter.hpp), links the
// executable with -lzfvm_b200 -lcudart, and returns the classes below from the experiment's virtual factory
// functions.  Nothing else of the driver changes: the time loop keeps calling the reference's own interfaces.
//
//   zisa::b200::Context               what EulerExperiment::choose_physical_rate_of_change builds
//                                     (include/zisa/experiments/euler_experiment_impl.hpp:303-313), as one device context
//   zisa::b200::CudaEulerRateOfChange zisa::RateOfChange            include/zisa/ode/rate_of_change.hpp:23-43
//   zisa::b200::CudaRungeKutta        zisa::TimeIntegration         include/zisa/ode/time_integration.hpp:42-44
//   zisa::b200::CudaCFL               zisa::CFLCondition            include/zisa/model/cfl_condition.hpp:12-18
//   zisa::b200::CudaSanityCheck       zisa::SanityCheck             include/zisa/model/sanity_check.hpp:12-16
//   zisa::b200::CudaFrozenBC          zisa::BoundaryCondition       include/zisa/boundary/boundary_condition.hpp:10-19
//   zisa::b200::NcclHaloExchange      zisa::HaloExchange            include/zisa/parallelization/halo_exchange.hpp:11-21
//
// Errors: every C-ABI call returns int; non-zero becomes LOG_ERR(zfvm_last_error()), which is how the reference reports
// failures (src/zisa/reconstruction/lsq_solver.cpp:134, src/zisa/ode/time_loop.cpp:173).
//
// tests/cpp/ compiles this header against a stand-in of the interfaces (-DZFVM_ADAPTER_ZISA_MOCK=...) and drives the C ABI
// through it from a plain C++ program.
#pragma once

#include <cstdint>
#include <memory>
#include <string>
#include <vector>

#include "zfvm.h"

#ifdef ZFVM_ADAPTER_ZISA_MOCK
#include ZFVM_ADAPTER_ZISA_MOCK
#else
#include <zisa/boundary/boundary_condition.hpp>
#include <zisa/config.hpp>
#include <zisa/grid/grid.hpp>
#include <zisa/model/all_variables.hpp>
#include <zisa/model/cfl_condition.hpp>
#include <zisa/model/sanity_check.hpp>
#include <zisa/ode/rate_of_change.hpp>
#include <zisa/ode/time_integration.hpp>
#include <zisa/parallelization/halo_exchange.hpp>
#include <zisa/reconstruction/hybrid_weno_params.hpp>
#include <zisa/reconstruction/stencil_family.hpp>
#endif

namespace zisa {
namespace b200 {

inline void check(int rc) { LOG_ERR_IF(rc != 0, zfvm_last_error()); }

/// zfvm_params with the HybridWENOParams part filled in (hybrid_weno_params.hpp); the caller sets the model part
/// (gamma, gravity, well_balanced, flux_bc, n_avars, heating) from its own configuration.
inline zfvm_params make_params(const HybridWENOParams &weno, bool cweno = true) {
  zfvm_params p;
  zfvm_params_default(&p);
  p.recon_mode = cweno ? 0 : 1;
  LOG_ERR_IF(weno.linear_weights.size() > 8, "zfvm: at most 8 stencils per family");
  for (std::size_t k = 0; k < weno.linear_weights.size(); ++k) p.linear_weights[k] = weno.linear_weights[k];
  p.epsilon = weno.epsilon;
  p.exponent = weno.exponent;
  return p;
}

/// Owner of the flattened grid, the stencil tables and the device context of one rank.
class Context {
public:
  /// `families`: the array<StencilFamily, 1> the experiment already computed (global_reconstruction_decl.hpp:107-147),
  /// handed over as it is so that the device path reconstructs on exactly the reference's stencils; nullptr: the library
  /// selects them itself (compute_stencil_families, stencil_family.cpp:99-117).
  Context(const Grid &grid, const QRDegrees &qr, const HybridWENOParams &weno, const zfvm_params &params, int device,
          const array<StencilFamily, 1> *families = nullptr)
      : n_cells_(grid.n_cells), n_avars_(params.n_avars) {
    // zisa::array is contiguous row-major (all_variables.hpp:31-35, mpi_halo_exchange.cpp:148-150): raw pointers go over;
    // nothing is kept after the calls.  int_t (64-bit unsigned) -> int32.
    const int_t nv = grid.max_neighbours;
    std::vector<std::int32_t> vi(grid.n_cells * nv);
    for (int_t i = 0; i < grid.n_cells; ++i)
      for (int_t k = 0; k < nv; ++k) vi[i * nv + k] = (std::int32_t)grid.vertex_indices(i, k);
    static_assert(sizeof(XYZ) == 3 * sizeof(double), "XYZ is three doubles");
    check(zfvm_grid_from_mesh(grid.n_dims(), (std::int64_t)grid.n_vertices,
                              reinterpret_cast<const double *>(grid.vertices.raw()), (std::int64_t)grid.n_cells, vi.data(),
                              (int)qr.face_deg, (int)qr.volume_deg, (int)qr.moments_deg, &grid_));
    std::vector<std::uint8_t> flags(grid.n_cells);  // cell_flags.hpp:9-15: bit0 interior, bit1 ghost_cell, bit2 ghost_cell_l1
    for (int_t i = 0; i < grid.n_cells; ++i) {
      const CellFlags f = grid.cell_flags[i];
      flags[i] = (std::uint8_t)((f.interior ? 1 : 0) | (f.ghost_cell ? 2 : 0) | (f.ghost_cell_l1 ? 4 : 0));
    }
    check(zfvm_grid_set_flags(grid_, flags.data()));

    const StencilFamilyParams &sp = weno.stencil_family_params;
    const int ns = (int)sp.n_stencils();
    std::string biases;
    for (const std::string &b : sp.biases) biases += (b == "c" || b == "central") ? 'c' : 'b';  // stencil_bias.cpp
    if (families == nullptr) {
      check(zfvm_stencils_compute(grid_, ns, sp.orders.data(), biases.c_str(), sp.overfit_factors.data(), 0, &stencils_));
    } else {
      LOG_ERR_IF(families->size() != grid.n_cells, "zfvm: one stencil family per cell expected");
      std::vector<std::int32_t> n_family(grid.n_cells), order(grid.n_cells * ns, 1), size(grid.n_cells * ns, 0), global;
      std::vector<std::int64_t> offset(grid.n_cells * ns + 1, 0);
      for (int_t i = 0; i < grid.n_cells; ++i) {
        const StencilFamily &fam = (*families)[i];
        n_family[i] = (std::int32_t)fam.size();
        for (int k = 0; k < ns; ++k) {
          offset[i * ns + k] = (std::int64_t)global.size();
          if ((int_t)k >= fam.size()) continue;
          const Stencil &s = fam[k];
          order[i * ns + k] = s.order();
          size[i * ns + k] = (std::int32_t)s.size();
          for (int_t j = 0; j < s.size(); ++j) global.push_back((std::int32_t)s.global(j));
        }
      }
      offset[grid.n_cells * ns] = (std::int64_t)global.size();
      check(zfvm_stencils_from_arrays(grid_, ns, sp.orders.data(), biases.c_str(), sp.overfit_factors.data(),
                                      n_family.data(), order.data(), size.data(), offset.data(), global.data(), &stencils_));
    }
    check(zfvm_create(grid_, stencils_, &params, device, &ctx_));
  }
  Context(const Context &) = delete;
  Context &operator=(const Context &) = delete;
  ~Context() {
    if (ctx_) zfvm_destroy(ctx_);
    if (stencils_) zfvm_stencils_free(stencils_);
    if (grid_) zfvm_grid_free(grid_);
  }

  zfvm_ctx *ctx() const { return ctx_; }
  const zfvm_grid *grid() const { return grid_; }
  const zfvm_stencils *stencils() const { return stencils_; }
  int_t n_cells() const { return n_cells_; }
  int n_avars() const { return n_avars_; }

private:
  zfvm_grid *grid_ = nullptr;
  zfvm_stencils *stencils_ = nullptr;
  zfvm_ctx *ctx_ = nullptr;
  int_t n_cells_;
  int n_avars_;
};

/// The minimal swap: ONE rate-of-change object replaces Sum[FluxLoop, GravitySourceLoop, (FluxBC, Heating)] inside
/// aggregate_rates_of_change (src/zisa/experiments/numerical_experiment.cpp:238-256); host AllVariables stay where they are.
/// `tendency` arrives zeroed by ZeroRateOfChange and the contract is "+=" (src/zisa/ode/rate_of_change.cpp:49-54).
/// RungeKutta swaps its buffers between calls (runge_kutta.cpp:109-111): no host pointer is cached.
class CudaEulerRateOfChange : public RateOfChange {
public:
  explicit CudaEulerRateOfChange(std::shared_ptr<Context> context) : context_(std::move(context)) {}

  void compute(AllVariables &tendency, const AllVariables &current_state, double t) const override {
    LOG_ERR_IF(current_state.cvars.shape(0) != context_->n_cells(), "zfvm: state has a different number of cells");
    if (context_->n_avars() > 0)
      check(zfvm_rate_of_change_av(context_->ctx(), tendency.cvars.raw(), tendency.avars.raw(), current_state.cvars.raw(),
                                   current_state.avars.raw(), t, /*accumulate=*/1));
    else
      check(zfvm_rate_of_change(context_->ctx(), tendency.cvars.raw(), current_state.cvars.raw(), t, /*accumulate=*/1));
  }
  std::string str() const override { return "B200 flux loop + gravity source loop (libzfvm_b200)"; }

private:
  std::shared_ptr<Context> context_;
};

/// The full swap: the state stays on the device between steps (built where make_time_integration is called,
/// src/zisa/ode/time_integration_factory.cpp:10-29).  compute_step returns a host AllVariables like the reference; with
/// `download_every_step = false` its contents are refreshed only by download() (the driver calls it when
/// clock.is_plotting_step(), src/zisa/ode/time_loop.cpp:163-167) and CudaCFL / CudaSanityCheck answer from the device.
class CudaRungeKutta : public TimeIntegration {
public:
  CudaRungeKutta(std::shared_ptr<Context> context, const std::string &method, double cfl_number,
                 bool download_every_step = true)
      : context_(std::move(context)), method_(method), cfl_number_(cfl_number), download_every_step_(download_every_step) {
    check(zfvm_set_time_integration(context_->ctx(), method.c_str()));
  }

  std::shared_ptr<AllVariables> compute_step(const std::shared_ptr<AllVariables> &u0, double t, double dt) override {
    // "the returned smart pointer does not point to the same object as u0" and "RungeKutta ... can modify u0 in a later
    // call" (time_integration.hpp:20-31): two host buffers, the step writes into the one that is not u0
    const int w = (buf_[0].get() == u0.get()) ? 1 : 0;
    if (!buf_[w]) buf_[w] = std::make_shared<AllVariables>(u0->dims());
    if (!resident_ || u0.get() != last_returned_) upload(*u0);  // somebody else's state: take it
    check(zfvm_rk_step(context_->ctx(), t, dt, cfl_number_, &dt_next_, &not_plausible_));
    have_verdict_ = true;
    if (download_every_step_) download(*buf_[w]);
    last_returned_ = buf_[w].get();
    return buf_[w];
  }
  std::string str() const override { return "RungeKutta '" + method_ + "' on the device (libzfvm_b200)"; }

  void upload(const AllVariables &u) {
    check(zfvm_upload_state(context_->ctx(), u.cvars.raw()));
    if (context_->n_avars() > 0) check(zfvm_upload_avars(context_->ctx(), u.avars.raw()));
    resident_ = true;
    have_verdict_ = false;
  }
  void download(AllVariables &u) const {
    check(zfvm_download_state(context_->ctx(), u.cvars.raw()));
    if (context_->n_avars() > 0) check(zfvm_download_avars(context_->ctx(), u.avars.raw()));
  }
  /// LocalCFL of the resident state (model/local_cfl_condition_impl.hpp:25-40): the value that came back with the last
  /// step, or one reduction if there was none.
  double cfl_dt(const AllVariables &u) {
    if (!resident_) upload(u);
    if (!have_verdict_) {
      check(zfvm_cfl_dt(context_->ctx(), nullptr, cfl_number_, &dt_next_, &not_plausible_));
      have_verdict_ = true;
    }
    return dt_next_;
  }
  bool plausible(const AllVariables &u) {
    (void)cfl_dt(u);
    return not_plausible_ == 0;
  }
  const std::shared_ptr<Context> &context() const { return context_; }

private:
  std::shared_ptr<Context> context_;
  std::string method_;
  double cfl_number_;
  bool download_every_step_;
  bool resident_ = false, have_verdict_ = false;
  double dt_next_ = 0.0;
  int not_plausible_ = 0;
  std::shared_ptr<AllVariables> buf_[2];
  const AllVariables *last_returned_ = nullptr;
};

/// choose_cfl_condition() (euler_experiment_impl.hpp:231-242): dt = cfl_number * min inradius / (|v| + a), reduced on the
/// device inside the last stage of the previous step.
class CudaCFL : public CFLCondition {
public:
  explicit CudaCFL(std::shared_ptr<CudaRungeKutta> rk) : rk_(std::move(rk)) {}
  double operator()(const AllVariables &u) override { return rk_->cfl_dt(u); }

private:
  std::shared_ptr<CudaRungeKutta> rk_;
};

/// SanityCheckFor<Euler> (model/sanity_check_for.hpp:24-44): the plausibility flag travels with the CFL reduction.
class CudaSanityCheck : public SanityCheck {
public:
  explicit CudaSanityCheck(std::shared_ptr<CudaRungeKutta> rk) : rk_(std::move(rk)) {}
  bool operator()(const AllVariables &u) const override { return rk_->plausible(u); }

private:
  std::shared_ptr<CudaRungeKutta> rk_;
};

/// FrozenBC (src/zisa/boundary/frozen_boundary_condition.cpp:11-55): the ghost rows are reset inside the device step after
/// every stage; the host-side apply() the time loop may still call has nothing left to do.
class CudaFrozenBC : public BoundaryCondition {
public:
  CudaFrozenBC(std::shared_ptr<Context> context, const AllVariables &steady_state) : context_(std::move(context)) {
    if (context_->n_avars() > 0)
      check(zfvm_set_frozen_bc_av(context_->ctx(), steady_state.cvars.raw(), steady_state.avars.raw()));
    else
      check(zfvm_set_frozen_bc(context_->ctx(), steady_state.cvars.raw()));
  }
  void apply(AllVariables &, double) override {}
  std::string str() const override { return "FrozenBC inside the device step (libzfvm_b200)"; }

private:
  std::shared_ptr<Context> context_;
};

/// choose_halo_exchange() (euler_experiment_impl.hpp:315-319).  Inside zfvm_rk_step / zfvm_rate_of_change the exchange is
/// posted by the library and overlapped with the interior reconstruction; the two members are for callers that drive the
/// resident state themselves (mpi_halo_exchange.cpp:178-201).
class NcclHaloExchange : public HaloExchange {
public:
  explicit NcclHaloExchange(std::shared_ptr<Context> context, bool explicit_exchange = false)
      : context_(std::move(context)), explicit_(explicit_exchange) {}
  void operator()(AllVariables &) override {
    if (explicit_) check(zfvm_halo_post(context_->ctx(), nullptr, nullptr));
  }
  void wait() override {
    if (explicit_) check(zfvm_halo_wait(context_->ctx()));
  }

private:
  std::shared_ptr<Context> context_;
  bool explicit_;
};

}  // namespace b200
}  // namespace zisa
