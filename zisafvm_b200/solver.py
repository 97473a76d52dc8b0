"""Host-side mirror of the reference's operator interfaces for the residual path.

Names, argument meaning and error behaviour follow the reference classes they replace:

* ``CudaEulerRateOfChange.compute(tendency, current_state, t)``  <->  ``zisa::RateOfChange::compute``
  (include/zisa/ode/rate_of_change.hpp:37-39) for ``Sum[FluxLoop, GravitySourceLoop]``
  (include/zisa/fvm_loops/flux_loop.hpp:96-104, gravity_source_loop.hpp:32-87,121-148);
* ``CudaRungeKutta.compute_step(u0, t, dt)``  <->  ``RungeKutta::compute_step``
  (src/zisa/ode/runge_kutta.cpp:87-112);
* ``FrozenBC``  <->  src/zisa/boundary/frozen_boundary_condition.cpp;
* ``LocalCFL``  <->  include/zisa/model/local_cfl_condition_impl.hpp:25-40.

Everything executes in ``libzfvm_b200.so`` on the GPU; this file is argument marshalling only.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Optional, Sequence

import numpy as np

from . import _capi
from ._capi import ZfvmParams, check, lib
from .grid import Grid, HybridWENOParams, StencilFamilies

GRAVITY_KINDS = {"none": 0, "constant": 1, "point_mass": 2, "polytrope": 3, "table": 4, "user": 5}


@dataclass
class Gravity:
    """Gravity model selector (include/zisa/model/gravity_decl.hpp)."""

    kind: str = "none"
    params: Sequence[float] = ()
    alignment: str = "radial"  # RadialAlignment | AxialAlignment
    axis: Sequence[float] = (0.0, 1.0, 0.0)
    table: Optional[tuple] = None  # (radii, phi) for SphericalGravity


@dataclass
class EulerParams:
    """Scheme parameters: what ``EulerExperiment`` reads from the JSON config."""

    weno: HybridWENOParams
    reconstruction: str = "CWENO-AO"  # "reconstruction.mode"
    well_balancing: str = "constant"  # "well-balancing.mode": constant | isentropic
    scaling: str = "euler"  # EulerScaling | UnityScaling
    flux: str = "hllc"  # hllc | rusanov
    gamma: float = 1.4
    gas_constant: float = 1.0
    gravity: Gravity = field(default_factory=Gravity)
    keep_polynomials: bool = False
    flux_bc: str = "none"  # "flux-bc": none | flux (FluxBC, boundary/flux_bc.hpp) | equilibrium (EquilibriumFluxBC)
    n_avars: int = 0  # advected scalars, AllVariables::avars (all_variables.hpp:31-35)
    heating: Optional[tuple] = None  # (rate, lower_boundary, upper_boundary) of "heating" (model/heating.hpp:54-80)
    # LocalRCParams, "reconstruction.steps_per_recompute" / ".recompute_threshold" (euler_experiment_impl.hpp:83-90)
    steps_per_recompute: int = 1
    recompute_threshold: float = 0.0

    def to_c(self) -> ZfvmParams:
        p = ZfvmParams()
        lib.zfvm_params_default(C.byref(p))
        p.recon_mode = {"CWENO-AO": 0, "WENO-AO": 1}[self.reconstruction]
        for k in range(8):
            p.linear_weights[k] = 0.0
        for k, w in enumerate(self.weno.linear_weights):
            p.linear_weights[k] = float(w)
        p.epsilon = float(self.weno.epsilon)
        p.exponent = float(self.weno.exponent)
        p.well_balanced = {"constant": 0, "isentropic": 1}[self.well_balancing]
        p.scaling = {"unity": 0, "euler": 1}[self.scaling]
        p.flux = {"hllc": 0, "rusanov": 1}[self.flux]
        p.gamma = float(self.gamma)
        p.gas_constant = float(self.gas_constant)
        p.gravity_kind = GRAVITY_KINDS[self.gravity.kind]
        p.gravity_alignment = {"radial": 0, "axial": 1}[self.gravity.alignment]
        for k, v in enumerate(self.gravity.params):
            p.gravity_p[k] = float(v)
        for k, v in enumerate(self.gravity.axis):
            p.gravity_axis[k] = float(v)
        p.steps_per_recompute = int(self.steps_per_recompute)
        p.recompute_threshold = float(self.recompute_threshold)
        p.keep_polynomials = int(self.keep_polynomials)
        p.flux_bc = {"none": 0, "flux": 1, "equilibrium": 2}[self.flux_bc]
        p.n_avars = int(self.n_avars)
        if self.heating is not None:
            p.heating_rate, p.heating_r0, p.heating_r1 = (float(x) for x in self.heating)
        return p


class AllVariables:
    """``zisa::AllVariables`` (all_variables.hpp:31-35): ``cvars[n_cells][5]`` and ``avars[n_cells][n_avars]`` (host)."""

    def __init__(self, n_cells: int, cvars: Optional[np.ndarray] = None, avars: Optional[np.ndarray] = None, n_avars: int = 0):
        self.cvars = np.zeros((n_cells, 5)) if cvars is None else np.ascontiguousarray(cvars, dtype=np.float64)
        assert self.cvars.shape == (n_cells, 5)
        if avars is None:
            self.avars = np.zeros((n_cells, n_avars))
        else:
            self.avars = np.ascontiguousarray(avars, dtype=np.float64).reshape(n_cells, -1)


class CudaContext:
    """One GPU, one (sub-)grid: owns the device-resident tables of the residual path."""

    def __init__(self, grid: Grid, stencils: StencilFamilies, params: EulerParams, device: int = 0):
        self.grid = grid
        self.stencils = stencils
        self.params = params
        cp = params.to_c()
        h = C.c_void_p()
        check(lib.zfvm_create(grid._h, stencils._h, C.byref(cp), device, C.byref(h)))
        self._h = h
        self.n_cells = grid.n_cells
        self.n_avars = int(params.n_avars)
        if params.gravity.kind == "table":
            r, phi = params.gravity.table
            r = _capi.as_f64(r)
            phi = _capi.as_f64(phi)
            check(lib.zfvm_set_gravity_table(h, grid._h, r.size, _capi.ptr_f64(r), _capi.ptr_f64(phi)))

    def close(self):
        h = getattr(self, "_h", None)
        if h:
            lib.zfvm_destroy(h)
            self._h = None

    def __del__(self):
        self.close()

    # -- information ------------------------------------------------------------------------------
    def memory_info(self):
        b = C.c_int64()
        a = C.c_double()
        check(lib.zfvm_memory_info(self._h, C.byref(b), C.byref(a)))
        return int(b.value), float(a.value)

    def counters(self):
        c = (C.c_int64 * 4)()
        check(lib.zfvm_counters(self._h, c))
        return {"launches": int(c[0]), "eq_failures": int(c[1]), "tiles_interior": int(c[2]), "tiles_exterior": int(c[3])}

    def profile(self, enable: bool):
        check(lib.zfvm_profile_enable(self._h, int(enable)))

    def profile_read(self):
        """Summed device milliseconds and launch counts of (reconstruction, face flux, update) kernels."""
        ms = (C.c_double * 3)()
        cnt = (C.c_int64 * 3)()
        check(lib.zfvm_profile_read(self._h, ms, cnt))
        out_ms, out_cnt = [float(x) for x in ms], [int(x) for x in cnt]
        if self.n_avars > 0:  # fourth entry: the advected-scalar kernels (T1 + T2 + T3)
            tms, tcnt = C.c_double(), C.c_int64()
            check(lib.zfvm_profile_read_tracers(self._h, C.byref(tms), C.byref(tcnt)))
            out_ms.append(float(tms.value))
            out_cnt.append(int(tcnt.value))
        return out_ms, out_cnt

    def stream(self) -> int:
        return int(lib.zfvm_stream(self._h) or 0)

    def synchronize(self):
        check(lib.zfvm_synchronize(self._h))

    def set_gravity_values(self, phi_cqp, gradphi_cqp, phi_fqp):
        a, b, c = _capi.as_f64(phi_cqp), _capi.as_f64(gradphi_cqp), _capi.as_f64(phi_fqp)
        check(lib.zfvm_set_gravity_values(self._h, _capi.ptr_f64(a), _capi.ptr_f64(b), _capi.ptr_f64(c)))

    def polynomials(self):
        n_coef = C.c_int()
        coeffs = np.zeros((self.n_cells, 35, 5))
        scale = np.zeros((self.n_cells, 5))
        # the library writes [n][n_coef][5] compactly
        flat = np.zeros(self.n_cells * 35 * 5)
        check(lib.zfvm_download_polynomials(self._h, _capi.ptr_f64(flat), _capi.ptr_f64(scale), C.byref(n_coef)))
        d = n_coef.value
        return flat[: self.n_cells * d * 5].reshape(self.n_cells, d, 5).copy(), scale

    def work_array(self, name: str) -> np.ndarray:
        g = self.grid
        count = {"trace": g.n_interior_edges * 2 * g.q_f * 5, "flux": g.n_interior_edges * 5, "source": g.n_cells * 5}[name]
        out = np.zeros(count)
        check(lib.zfvm_download_work(self._h, name.encode(), _capi.ptr_f64(out), count))
        return out


    def records(self) -> np.ndarray:
        """The tile records (headers, stencil weights, geometry) as raw bytes -- for tests of the record builders."""
        dev_bytes, _ = self.memory_info()
        out = np.zeros(dev_bytes // 8 + 1)
        check(lib.zfvm_download_work(self._h, b"records", _capi.ptr_f64(out), out.size))
        return out.view(np.uint8)


class CudaEulerRateOfChange:
    """Drop-in for the reference's ``Sum[FluxLoop, GravitySourceLoop]`` rate of change."""

    def __init__(self, ctx: CudaContext):
        self.ctx = ctx

    def compute(self, tendency: AllVariables, current_state: AllVariables, t: float = 0.0, accumulate: bool = True):
        """``RateOfChange::compute``: ``tendency += rate(current_state)`` (host buffers)."""
        assert tendency.cvars.flags.c_contiguous and current_state.cvars.flags.c_contiguous
        if self.ctx.n_avars > 0:
            assert tendency.avars.shape == current_state.avars.shape == (self.ctx.n_cells, self.ctx.n_avars)
            check(lib.zfvm_rate_of_change_av(self.ctx._h, _capi.ptr_f64(tendency.cvars), _capi.ptr_f64(tendency.avars),
                                             _capi.ptr_f64(current_state.cvars), _capi.ptr_f64(current_state.avars),
                                             float(t), int(accumulate)))
            return
        check(lib.zfvm_rate_of_change(self.ctx._h, _capi.ptr_f64(tendency.cvars), _capi.ptr_f64(current_state.cvars),
                                      float(t), int(accumulate)))

    def compute_device(self, tendency_ptr: int, state_ptr: int, t: float = 0.0, accumulate: bool = False):
        check(lib.zfvm_rate_of_change_device(self.ctx._h, tendency_ptr, state_ptr, float(t), int(accumulate)))

    def str(self) -> str:
        return "B200 flux loop + gravity source loop (libzfvm_b200)"


class FrozenBC:
    """``zisa::FrozenBC``: ghost rows are reset to the saved steady state after every stage."""

    def __init__(self, ctx: CudaContext, steady_state: AllVariables):
        self.ctx = ctx
        s = _capi.as_f64(steady_state.cvars)
        if ctx.n_avars > 0:
            a = _capi.as_f64(steady_state.avars)
            check(lib.zfvm_set_frozen_bc_av(ctx._h, _capi.ptr_f64(s), _capi.ptr_f64(a)))
        else:
            check(lib.zfvm_set_frozen_bc(ctx._h, _capi.ptr_f64(s)))

    def apply_device(self, state_ptr: int):
        check(lib.zfvm_apply_frozen_bc(self.ctx._h, state_ptr))


class LocalCFL:
    """``zisa::LocalCFL``: ``cfl_number * min_i inradius_i / (|v_i| + a_i)``, evaluated on the device."""

    def __init__(self, ctx: CudaContext, cfl_number: float):
        self.ctx = ctx
        self.cfl_number = cfl_number

    def __call__(self, state_ptr: Optional[int] = None):
        dt = C.c_double()
        bad = C.c_int()
        check(lib.zfvm_cfl_dt(self.ctx._h, state_ptr, self.cfl_number, C.byref(dt), C.byref(bad)))
        return float(dt.value), bool(bad.value)


class CudaRungeKutta:
    """``zisa::RungeKutta`` with device-resident stages (make_tableau names: ssp3, ssp2, rk4, ...)."""

    def __init__(self, ctx: CudaContext, method: str = "ssp3"):
        self.ctx = ctx
        self.method = method
        check(lib.zfvm_set_time_integration(ctx._h, method.encode()))

    def upload(self, u0: AllVariables):
        check(lib.zfvm_upload_state(self.ctx._h, _capi.ptr_f64(u0.cvars)))
        if self.ctx.n_avars > 0:
            check(lib.zfvm_upload_avars(self.ctx._h, _capi.ptr_f64(u0.avars)))

    def download(self, out: Optional[AllVariables] = None) -> AllVariables:
        out = out or AllVariables(self.ctx.n_cells, n_avars=self.ctx.n_avars)
        check(lib.zfvm_download_state(self.ctx._h, _capi.ptr_f64(out.cvars)))
        if self.ctx.n_avars > 0:
            check(lib.zfvm_download_avars(self.ctx._h, _capi.ptr_f64(out.avars)))
        return out

    def step(self, t: float, dt: float, cfl_number: Optional[float] = None):
        """One step on the resident state; returns (dt_next, not_plausible) when ``cfl_number`` is given."""
        if cfl_number is None:
            check(lib.zfvm_rk_step(self.ctx._h, t, dt, 0.0, None, None))
            return None
        dtn = C.c_double()
        bad = C.c_int()
        check(lib.zfvm_rk_step(self.ctx._h, t, dt, cfl_number, C.byref(dtn), C.byref(bad)))
        return float(dtn.value), bool(bad.value)

    def compute_step(self, u0: AllVariables, t: float, dt: float, out: Optional[AllVariables] = None) -> AllVariables:
        """``TimeIntegration::compute_step(u0, t, dt) -> u1`` with host buffers (H2D + D2H inside).  ``out`` lets the
        caller hand in the result buffer (the reference's RungeKutta owns and swaps its buffers, runge_kutta.cpp:109-111);
        pinned buffers are copied without staging."""
        u1 = AllVariables(self.ctx.n_cells, n_avars=self.ctx.n_avars) if out is None else out
        if self.ctx.n_avars > 0:
            check(lib.zfvm_rk_step_host_av(self.ctx._h, _capi.ptr_f64(u0.cvars), _capi.ptr_f64(u0.avars),
                                           _capi.ptr_f64(u1.cvars), _capi.ptr_f64(u1.avars), t, dt))
            return u1
        check(lib.zfvm_rk_step_host(self.ctx._h, _capi.ptr_f64(u0.cvars), _capi.ptr_f64(u1.cvars), t, dt))
        return u1
