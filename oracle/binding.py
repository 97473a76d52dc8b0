"""ctypes binding of the CPU oracle (oracle/zisa_oracle.cpp).

TEST INFRASTRUCTURE: imported only by tests/, __graft_entry__.smoke() and bench.py's CPU-baseline /
reference arm.  The product package (zisafvm_b200/) never imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libzisa_oracle.so")

dp = C.POINTER(C.c_double)
ip = C.POINTER(C.c_int32)
lp = C.POINTER(C.c_int64)
bp = C.POINTER(C.c_uint8)


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "zisa_oracle.cpp")
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return LIB_PATH


class GridDesc(C.Structure):
    _fields_ = [("n_dims", C.c_int), ("q_c", C.c_int), ("q_f", C.c_int), ("n_moments", C.c_int),
                ("n_cells", C.c_int64), ("n_edges", C.c_int64), ("n_interior_edges", C.c_int64),
                ("left_right", ip), ("edge_indices", ip),
                ("volumes", dp), ("cell_centers", dp), ("char_length", dp), ("moments", dp), ("cell_qp", dp),
                ("cell_qw", dp), ("face_qp", dp), ("face_qw", dp), ("face_normal", dp), ("face_t1", dp),
                ("face_t2", dp), ("inradii", dp), ("cell_flags", bp), ("phi_cqp", dp), ("gradphi_cqp", dp),
                ("phi_fqp", dp)]


class StencilDesc(C.Structure):
    _fields_ = [("n_stencils", C.c_int), ("l2g_stride", C.c_int), ("l2g", ip), ("l2g_size", ip), ("local", ip),
                ("local_off", ip), ("order", ip), ("size", ip), ("k_high", ip), ("n_family", ip), ("A", dp),
                ("A_stride", C.c_int64), ("A_off", lp)]


class OracleParams(C.Structure):
    _fields_ = [("recon_mode", C.c_int), ("linear_weights", C.c_double * 8), ("epsilon", C.c_double),
                ("exponent", C.c_double), ("well_balanced", C.c_int), ("scaling", C.c_int), ("flux", C.c_int),
                ("gamma", C.c_double), ("gas_constant", C.c_double), ("has_gravity", C.c_int), ("flux_bc", C.c_int)]


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(LIB_PATH)
        _lib.oracle_create.restype = C.c_void_p
        _lib.oracle_create.argtypes = [C.POINTER(GridDesc), C.POINTER(StencilDesc), C.POINTER(OracleParams)]
        _lib.oracle_destroy.argtypes = [C.c_void_p]
        _lib.oracle_set_frozen_bc.argtypes = [C.c_void_p, dp]
        _lib.oracle_rate_of_change.argtypes = [C.c_void_p, dp, dp]
        _lib.oracle_reconstruct.argtypes = [C.c_void_p, dp, dp, C.c_int, dp]
        _lib.oracle_point_value.argtypes = [C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int, dp]
        _lib.oracle_eval_at.argtypes = [C.c_void_p, C.c_int64, dp, dp]
        _lib.oracle_cell_point_values.argtypes = [C.c_void_p, dp]
        _lib.oracle_rk_step.argtypes = [C.c_void_p, C.c_char_p, dp, dp, C.c_double]
        _lib.oracle_rk_step.restype = C.c_int
        _lib.oracle_cfl_dt.argtypes = [C.c_void_p, dp, C.c_double]
        _lib.oracle_cfl_dt.restype = C.c_double
        _lib.oracle_eq_failures.argtypes = [C.c_void_p]
        _lib.oracle_hllc.argtypes = [C.c_double, dp, dp, dp]
        _lib.oracle_rusanov.argtypes = [C.c_double, dp, dp, dp]
        _lib.oracle_euler_flux.argtypes = [C.c_double, dp, dp]
        _lib.oracle_poly_eval.argtypes = [C.c_int, C.c_int, C.c_int, dp, dp, C.c_int, dp, C.c_double, dp, dp]
        _lib.oracle_poly_saxpy.argtypes = [C.c_int, C.c_int, dp, dp, C.c_int, dp, dp, dp, dp]
        _lib.oracle_lsq_solve.argtypes = [dp, C.c_int, C.c_int, dp, C.c_int, dp]
        _lib.oracle_eos_rhoE_to_hK.argtypes = [C.c_double, C.c_double, C.c_double, dp, dp]
        _lib.oracle_eos_hK_to_rhoE.argtypes = [C.c_double, C.c_double, C.c_double, dp, dp]
        _lib.oracle_local_equilibrium.argtypes = [C.c_double, C.c_int, dp, dp, C.c_double, C.c_double, C.c_double, dp, dp, dp]
        _lib.oracle_local_equilibrium.restype = C.c_int
        _lib.oracle_set_tracers.argtypes = [C.c_void_p, C.c_int]
        _lib.oracle_set_heating.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_double]
        _lib.oracle_set_local_rc_params.argtypes = [C.c_void_p, C.c_int, C.c_double]
        _lib.oracle_set_flux_bc.argtypes = [C.c_void_p, C.c_int]
        _lib.oracle_set_frozen_bc_av.argtypes = [C.c_void_p, dp, dp]
        _lib.oracle_rate_of_change_av.argtypes = [C.c_void_p, dp, dp, dp, dp]
        _lib.oracle_tracer_polys.argtypes = [C.c_void_p, dp, C.c_int]
        _lib.oracle_hllc_tracer_flux.argtypes = [C.c_double, dp, dp, C.c_double, C.c_double]
        _lib.oracle_hllc_tracer_flux.restype = C.c_double
        _lib.oracle_rk_step_av.argtypes = [C.c_void_p, C.c_char_p, dp, dp, dp, dp, C.c_double]
        _lib.oracle_rk_step_av.restype = C.c_int
    return _lib


def _p(a, t=dp):
    return a.ctypes.data_as(t) if a is not None else None


def f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


class Oracle:
    """CPU oracle for one grid + stencil set + parameter set (arrays come from the product's host library,
    the arithmetic of the path is the oracle's own)."""

    def __init__(self, grid, stencils, params, gravity_tables=None):
        """grid/stencils: zisafvm_b200 Grid / StencilFamilies; params: zisafvm_b200 EulerParams;
        gravity_tables: (phi_cqp, gradphi_cqp, phi_fqp) numpy arrays when gravity is on."""
        L = lib()
        self._keep = []

        def keep(a):
            self._keep.append(a)
            return a

        g = GridDesc()
        g.n_dims, g.q_c, g.q_f, g.n_moments = grid.n_dims, grid.q_c, grid.q_f, grid.n_moments
        g.n_cells, g.n_edges, g.n_interior_edges = grid.n_cells, grid.n_edges, grid.n_interior_edges
        g.left_right = _p(keep(grid.array("left_right")), ip)
        g.edge_indices = _p(keep(grid.array("edge_indices")), ip)
        for name, attr in [("volumes", "volumes"), ("cell_centers", "cell_centers"),
                           ("characteristic_length", "char_length"), ("moments", "moments"), ("cell_qp", "cell_qp"),
                           ("cell_qw", "cell_qw"), ("face_qp", "face_qp"), ("face_qw", "face_qw"),
                           ("face_normal", "face_normal"), ("face_t1", "face_t1"), ("face_t2", "face_t2"),
                           ("inradii", "inradii")]:
            setattr(g, attr, _p(keep(grid.array(name))))
        g.cell_flags = _p(keep(grid.array("cell_flags")), bp)
        if gravity_tables is not None:
            a, b, c = (keep(f64(x)) for x in gravity_tables)
            g.phi_cqp, g.gradphi_cqp, g.phi_fqp = _p(a), _p(b), _p(c)
        s = StencilDesc()
        s.n_stencils = stencils.n_stencils
        l2g = keep(stencils.array("l2g"))
        s.l2g_stride = l2g.shape[1]
        s.l2g = _p(l2g, ip)
        for name in ["l2g_size", "local", "local_off", "order", "size", "k_high", "n_family"]:
            setattr(s, name, _p(keep(np.ascontiguousarray(stencils.array(name), dtype=np.int32)), ip))
        A, a_off, stride = stencils.all_matrices()
        keep(A), keep(a_off)
        s.A, s.A_stride, s.A_off = _p(A), stride, _p(a_off, lp)
        p = OracleParams()
        p.recon_mode = {"CWENO-AO": 0, "WENO-AO": 1}[params.reconstruction]
        for k, w in enumerate(params.weno.linear_weights):
            p.linear_weights[k] = float(w)
        p.epsilon, p.exponent = params.weno.epsilon, params.weno.exponent
        p.well_balanced = int(params.well_balancing == "isentropic")
        p.scaling = {"unity": 0, "euler": 1}[params.scaling]
        p.flux = {"hllc": 0, "rusanov": 1}[params.flux]
        p.gamma, p.gas_constant = params.gamma, params.gas_constant
        p.has_gravity = int(params.gravity.kind != "none")
        p.flux_bc = {"none": 0, "flux": 1, "equilibrium": 2}[getattr(params, "flux_bc", "none")]
        self._descs = (g, s, p)
        self.n_cells = grid.n_cells
        self._h = L.oracle_create(C.byref(g), C.byref(s), C.byref(p))
        self.n_avars = int(getattr(params, "n_avars", 0))
        L.oracle_set_tracers(self._h, self.n_avars)
        spr = int(getattr(params, "steps_per_recompute", 1))
        if spr != 1:
            L.oracle_set_local_rc_params(self._h, spr, float(getattr(params, "recompute_threshold", 0.0)))
        heating = getattr(params, "heating", None)
        if heating is not None:
            L.oracle_set_heating(self._h, float(heating[0]), float(heating[1]), float(heating[2]))

    def __del__(self):
        if getattr(self, "_h", None):
            lib().oracle_destroy(self._h)
            self._h = None

    def set_frozen_bc(self, steady):
        lib().oracle_set_frozen_bc(self._h, _p(f64(steady)) if steady is not None else None)

    def rate_of_change(self, state, tendency=None):
        state = f64(state)
        t = np.zeros_like(state) if tendency is None else tendency
        lib().oracle_rate_of_change(self._h, _p(t), _p(state))
        return t

    def reconstruct(self, state, n_coef):
        state = f64(state)
        coeffs = np.zeros((self.n_cells, n_coef, 5))
        scale = np.zeros((self.n_cells, 5))
        lib().oracle_reconstruct(self._h, _p(state), _p(coeffs), n_coef, _p(scale))
        return coeffs, scale

    def point_value(self, i, kind, k, q):
        u = np.zeros(5)
        lib().oracle_point_value(self._h, i, kind, k, q, _p(u))
        return u

    def cell_point_values(self, q_c):
        """rc(i)(x) at all cell Gauss points, [n][q_c][5]; call after reconstruct() / rate_of_change()."""
        out = np.zeros((self.n_cells, q_c, 5))
        lib().oracle_cell_point_values(self._h, _p(out))
        return out

    def eval_at(self, i, x):
        u = np.zeros(5)
        lib().oracle_eval_at(self._h, i, _p(f64(x)), _p(u))
        return u

    def rk_step(self, method, u0, dt):
        u0 = f64(u0)
        u1 = np.zeros_like(u0)
        rc = lib().oracle_rk_step(self._h, method.encode(), _p(u0), _p(u1), dt)
        if rc:
            raise ValueError(f"Unknown Butcher Tableau. [{method}]")
        return u1

    # -- AllVariables{cvars, avars} (advected scalars) ----------------------------------
    def set_frozen_bc_av(self, steady, steady_av):
        lib().oracle_set_frozen_bc_av(self._h, _p(f64(steady)), _p(f64(steady_av)))

    def rate_of_change_av(self, state, avars):
        state, avars = f64(state), f64(avars)
        t, ta = np.zeros_like(state), np.zeros_like(avars)
        lib().oracle_rate_of_change_av(self._h, _p(t), _p(ta), _p(state), _p(avars))
        return t, ta

    def tracer_polys(self, n_coef):
        out = np.zeros((self.n_cells, self.n_avars, n_coef))
        lib().oracle_tracer_polys(self._h, _p(out), n_coef)
        return out

    def rk_step_av(self, method, u0, a0, dt):
        u0, a0 = f64(u0), f64(a0)
        u1, a1 = np.zeros_like(u0), np.zeros_like(a0)
        rc = lib().oracle_rk_step_av(self._h, method.encode(), _p(u0), _p(a0), _p(u1), _p(a1), dt)
        if rc:
            raise ValueError(f"Unknown Butcher Tableau. [{method}]")
        return u1, a1

    def cfl_dt(self, u, cfl_number):
        return lib().oracle_cfl_dt(self._h, _p(f64(u)), cfl_number)

    def eq_failures(self):
        return lib().oracle_eq_failures(self._h)


def hllc_tracer_flux(gamma, uL, uR, mqL, mqR) -> float:
    """HLLCBatten::tracer_flux (flux/hllc.hpp:178-197) with the speeds of HLLCBatten::flux for (uL, uR)."""
    return lib().oracle_hllc_tracer_flux(gamma, _p(f64(uL)), _p(f64(uR)), float(mqL), float(mqR))


def num_threads() -> int:
    return lib().oracle_num_threads()
