// Stand-in for the ZisaFVM / ZisaCore / ZisaMemory declarations include/zfvm_zisa_adapter.hpp uses -- TEST INFRASTRUCTURE.
// The reference cannot be compiled here (its sibling libraries are not vendored), so the adapter is compiled
// against these few classes instead.  Only names, signatures and member layouts the adapter touches are reproduced, each
// with the reference header it stands for; bodies are the shortest thing that works.
#pragma once

#include <cstddef>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace zisa {

// zisa/config.hpp (ZisaCore): int_t is a 64-bit unsigned integer; LOG_ERR prints and terminates -- the tests want to
// see the message, so the mock throws.
using int_t = std::size_t;
#define LOG_ERR(msg)                                 \
  do {                                               \
    std::ostringstream zfvm_mock_os;                 \
    zfvm_mock_os << msg;                             \
    throw std::runtime_error(zfvm_mock_os.str());    \
  } while (0)
#define LOG_ERR_IF(cond, msg) \
  do {                        \
    if (cond) LOG_ERR(msg);   \
  } while (0)

// zisa/memory/array.hpp (ZisaMemory): contiguous row-major storage with raw(), shape(k), size(), (i[, j]) access
template <class T, int N>
class array {
public:
  array() = default;
  explicit array(int_t n0, int_t n1 = 1) : shape_{n0, n1}, data_(n0 * n1) {}
  T *raw() { return data_.data(); }
  const T *raw() const { return data_.data(); }
  int_t shape(int_t k) const { return shape_[k]; }
  int_t size() const { return data_.size(); }
  T &operator[](int_t i) { return data_[i]; }
  const T &operator[](int_t i) const { return data_[i]; }
  T &operator()(int_t i) { return data_[i]; }
  const T &operator()(int_t i) const { return data_[i]; }
  T &operator()(int_t i, int_t j) { return data_[i * shape_[1] + j]; }
  const T &operator()(int_t i, int_t j) const { return data_[i * shape_[1] + j]; }
  auto begin() { return data_.begin(); }
  auto end() { return data_.end(); }

private:
  int_t shape_[2] = {0, 1};
  std::vector<T> data_;
};

// zisa/math/cartesian.hpp
struct XYZ {
  double x[3];
  double &operator[](int_t k) { return x[k]; }
  double operator[](int_t k) const { return x[k]; }
};

// zisa/model/grid_variables.hpp, zisa/model/all_variables.hpp:15-60
using GridVariables = array<double, 2>;
struct AllVariablesDimensions {
  int_t n_cells, n_cvars, n_avars;
};
class AllVariables {
public:
  GridVariables cvars, avars;
  AllVariables() = default;
  explicit AllVariables(const AllVariablesDimensions &d) : cvars(d.n_cells, d.n_cvars), avars(d.n_cells, d.n_avars) {}
  AllVariablesDimensions dims() const { return {cvars.shape(0), cvars.shape(1), avars.shape(1)}; }
};

// zisa/ode/rate_of_change.hpp:23-43,45-88
class RateOfChange {
public:
  virtual ~RateOfChange() = default;
  virtual void compute(AllVariables &tendency, const AllVariables &current_state, double t) const = 0;
  virtual std::string str() const = 0;
};
class ZeroRateOfChange : public RateOfChange {
public:
  void compute(AllVariables &tendency, const AllVariables &, double) const override {
    for (double &x : tendency.cvars) x = 0.0;
    for (double &x : tendency.avars) x = 0.0;
  }
  std::string str() const override { return "zero"; }
};
class SumRatesOfChange : public RateOfChange {
public:
  void add_term(const std::shared_ptr<RateOfChange> &rate) { terms_.push_back(rate); }
  void compute(AllVariables &tendency, const AllVariables &current_state, double t) const override {
    for (const auto &r : terms_) r->compute(tendency, current_state, t);
  }
  std::string str() const override { return "sum"; }

private:
  std::vector<std::shared_ptr<RateOfChange>> terms_;
};

// zisa/ode/time_integration.hpp:10-50
class TimeIntegration {
public:
  virtual ~TimeIntegration() = default;
  virtual std::shared_ptr<AllVariables> compute_step(const std::shared_ptr<AllVariables> &u0, double t, double dt) = 0;
  virtual std::string str() const = 0;
};

// zisa/model/cfl_condition.hpp:8-18, zisa/model/sanity_check.hpp:10-22
class CFLCondition {
public:
  virtual ~CFLCondition() = default;
  virtual double operator()(const AllVariables &u) = 0;
};
class SanityCheck {
public:
  virtual ~SanityCheck() = default;
  virtual bool operator()(const AllVariables &all_variables) const = 0;
};

// zisa/boundary/boundary_condition.hpp:8-20
class BoundaryCondition {
public:
  virtual ~BoundaryCondition() = default;
  virtual void apply(AllVariables &u, double t) = 0;
  virtual std::string str() const = 0;
};

// zisa/parallelization/halo_exchange.hpp:9-22
class HaloExchange {
public:
  virtual ~HaloExchange() = default;
  virtual void operator()(AllVariables &all_vars) = 0;
  virtual void wait() = 0;
};

// zisa/grid/cell_flags.hpp:6-14, zisa/grid/grid_decl.hpp:28-107 (the members the adapter reads)
struct CellFlags {
  bool interior : 1;
  bool ghost_cell : 1;
  bool ghost_cell_l1 : 1;
  CellFlags() : interior(true), ghost_cell(false), ghost_cell_l1(false) {}
};
struct QRDegrees {
  int_t face_deg, volume_deg, moments_deg;
};
struct Grid {
  int_t n_cells = 0, n_vertices = 0, max_neighbours = 0;
  array<int_t, 2> vertex_indices;
  array<XYZ, 1> vertices;
  array<CellFlags, 1> cell_flags;
  int n_dims() const { return (int)max_neighbours - 1; }
};

// zisa/reconstruction/stencil_family_params.hpp:14-33, hybrid_weno_params.hpp:11-32
struct StencilFamilyParams {
  std::vector<int> orders;
  std::vector<std::string> biases;
  std::vector<double> overfit_factors;
  int_t n_stencils() const { return orders.size(); }
};
struct HybridWENOParams {
  StencilFamilyParams stencil_family_params;
  std::vector<double> linear_weights;
  double epsilon, exponent;
};

// zisa/reconstruction/stencil.hpp:15-85, stencil_family.hpp:10-75 (read access only)
class Stencil {
public:
  Stencil() = default;
  Stencil(std::vector<int_t> global, int order) : order_(order), global_(std::move(global)) {}
  int_t global(int_t k) const { return global_[k]; }
  int order() const { return order_; }
  int_t size() const { return global_.size(); }

private:
  int order_ = 1;
  std::vector<int_t> global_;
};
class StencilFamily {
public:
  StencilFamily() = default;
  explicit StencilFamily(std::vector<Stencil> stencils) : stencils_(std::move(stencils)) {}
  const Stencil &operator[](int_t k) const { return stencils_[k]; }
  int_t size() const { return stencils_.size(); }

private:
  std::vector<Stencil> stencils_;
};

}  // namespace zisa
