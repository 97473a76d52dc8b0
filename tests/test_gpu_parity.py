"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle on the same seeded inputs.

Tolerances: north_star asks for relative L1 / L-inf <= 1e-11 per variable after a fixed number of steps;
single residual evaluations and polynomial coefficients are held to tighter bounds (stated per test).
"""
import numpy as np
import pytest

import zisafvm_b200 as z
from zisafvm_b200 import cases

from util import active_vars, rel_err, rel_l1

pytestmark = pytest.mark.gpu


def _oracle(case, st):
    from oracle.binding import Oracle

    tables = cases.gravity_tables(case.grid, case.params.gravity) if case.params.gravity.kind != "none" else None
    return Oracle(case.grid, st, case.params, tables)


CASES = {
    "vortex_o3_hllc": lambda: cases.isentropic_vortex(n=40, order=3, flux="hllc"),
    "vortex_o3_rusanov": lambda: cases.isentropic_vortex(n=40, order=3, flux="rusanov"),
    "vortex_o2": lambda: cases.isentropic_vortex(n=33, order=2),
    "vortex_o4": lambda: cases.isentropic_vortex(n=33, order=4),
    "vortex_o5": lambda: cases.isentropic_vortex(n=30, order=5),
    "blast_o2": lambda: cases.blast_3d(n=7, order=2, kind="blast"),
    "blast_o3": lambda: cases.blast_3d(n=7, order=3, kind="blast"),
    "sod_o3": lambda: cases.blast_3d(n=6, order=3, kind="sod"),
    "smooth3d_o3": lambda: cases.blast_3d(n=7, order=3, kind="smooth"),
    "smooth3d_o4": lambda: cases.blast_3d(n=8, order=4, kind="smooth"),
    "polytrope_wb": lambda: cases.polytrope_2d(n=36, order=3, well_balanced=True),
    "polytrope_wb_perturbed": lambda: cases.polytrope_2d(n=36, order=3, well_balanced=True, amplitude=1e-3),
    "polytrope_nowb": lambda: cases.polytrope_2d(n=36, order=3, well_balanced=False),
    "atmosphere_wb": lambda: cases.stellar_atmosphere_3d(n=7, order=3, well_balanced=True),
    "atmosphere_nowb": lambda: cases.stellar_atmosphere_3d(n=7, order=2, well_balanced=False),
}


@pytest.fixture(scope="module", params=sorted(CASES))
def setup(request):
    case = CASES[request.param]()
    if case.name.endswith("o4") and case.grid.n_dims == 3:
        pytest.skip("3D order 4 kernel is not compiled yet")
    st = case.ensure_stencils()
    case.params.keep_polynomials = True
    ctx = z.CudaContext(case.grid, st, case.params)
    ora = _oracle(case, st)
    yield request.param, case, st, ctx, ora
    ctx.close()


def test_polynomial_coefficients(setup):
    """WENO polynomial of every cell, scaled basis: <= 1e-12 relative to the largest coefficient of the variable."""
    name, case, st, ctx, ora = setup
    roc = z.CudaEulerRateOfChange(ctx)
    tend = z.AllVariables(case.grid.n_cells)
    roc.compute(tend, z.AllVariables(case.grid.n_cells, case.u0), accumulate=False)
    coef, scale = ctx.polynomials()
    ref_coef, ref_scale = ora.reconstruct(case.u0, coef.shape[1])
    assert np.allclose(scale, ref_scale, rtol=1e-14, atol=0)
    for v in active_vars(case.grid.n_dims):
        den = np.abs(ref_coef[:, :, v]).max()
        err = np.abs(coef[:, :, v] - ref_coef[:, :, v]).max() / den
        assert err < 1e-12, (name, v, err)


def test_rate_of_change(setup):
    """RateOfChange::compute, overwrite and accumulate semantics: <= 1e-12 of the largest tendency."""
    name, case, st, ctx, ora = setup
    n = case.grid.n_cells
    roc = z.CudaEulerRateOfChange(ctx)
    tend = z.AllVariables(n)
    roc.compute(tend, z.AllVariables(n, case.u0), accumulate=False)
    ref = ora.rate_of_change(case.u0)
    err = rel_err(tend.cvars, ref)
    assert err.max() < 1e-12, (name, err)
    # accumulate: tendency += rate (the contract after ZeroRateOfChange, rate_of_change.cpp:21-27)
    base = np.random.default_rng(0).normal(size=(n, 5))
    tend2 = z.AllVariables(n, base.copy())
    roc.compute(tend2, z.AllVariables(n, case.u0), accumulate=True)
    assert np.allclose(tend2.cvars - base, tend.cvars, rtol=0, atol=1e-12 * np.abs(ref).max())
    assert ctx.counters()["eq_failures"] == 0 and ora.eq_failures() == 0


def test_runge_kutta_steps(setup):
    """N steps of the fused device RK against the oracle's Butcher-form RK: rel. L1 and L-inf <= 1e-11."""
    name, case, st, ctx, ora = setup
    n = case.grid.n_cells
    n_steps = 10 if case.grid.n_dims == 2 else 5
    rk = z.CudaRungeKutta(ctx, case.method)
    z.FrozenBC(ctx, z.AllVariables(n, case.u0))
    ora.set_frozen_bc(case.u0)
    rk.upload(z.AllVariables(n, case.u0))
    u_ref = case.u0.copy()
    dt = ora.cfl_dt(u_ref, case.cfl)
    t = 0.0
    for _ in range(n_steps):
        dt_next, bad = rk.step(t, dt, case.cfl)
        assert not bad
        u_ref = ora.rk_step(case.method, u_ref, dt)
        dt_ref = ora.cfl_dt(u_ref, case.cfl)
        assert abs(dt_next - dt_ref) <= 1e-11 * dt_ref
        t += dt
        dt = dt_ref  # both integrators use the same step sequence
    u = rk.download().cvars
    vol = case.grid.array("volumes")
    vs = active_vars(case.grid.n_dims)
    assert rel_err(u, u_ref)[vs].max() < 1e-11, (name, rel_err(u, u_ref))
    assert rel_l1(u, u_ref, vol)[vs].max() < 1e-11, (name, rel_l1(u, u_ref, vol))
    # ghost rows stay frozen
    gh = case.grid.is_ghost
    assert np.array_equal(u[gh], case.u0[gh])


def test_compute_step_host_matches_resident(setup):
    """TimeIntegration::compute_step with host buffers == the resident-state step, bit for bit."""
    name, case, st, ctx, ora = setup
    n = case.grid.n_cells
    rk = z.CudaRungeKutta(ctx, case.method)
    z.FrozenBC(ctx, z.AllVariables(n, case.u0))
    dt = ora.cfl_dt(case.u0, case.cfl)
    rk.upload(z.AllVariables(n, case.u0))
    rk.step(0.0, dt)
    a = rk.download().cvars
    b = rk.compute_step(z.AllVariables(n, case.u0), 0.0, dt).cvars
    assert np.array_equal(a, b)
