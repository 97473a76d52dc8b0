"""GPU parity of the rows that widen the path: advected scalars, `Heating`,
`EquilibriumFluxBC` -- the CUDA path through the C ABI against the CPU oracle on the same seeded inputs.

Tolerances as in test_gpu_parity.py: one residual <= 1e-12 of the flux scale, N Runge-Kutta steps relative
L1 / L-inf <= 1e-11 per variable (north_star), frozen ghost rows bit-identical.
"""
import numpy as np
import pytest

import zisafvm_b200 as z
from zisafvm_b200 import cases

from util import active_vars, rel_err, rel_l1, state_scales, tendency_scales

pytestmark = pytest.mark.gpu


def _oracle(case, st):
    from oracle.binding import Oracle

    tables = cases.gravity_tables(case.grid, case.params.gravity) if case.params.gravity.kind != "none" else None
    return Oracle(case.grid, st, case.params, tables)


# every record kind / kernel family the scalars have to ride on: tile records (2D / 3D order 3, 2D order 2), the
# older records (2D order 5; gravity / well-balanced runs), Rusanov, degraded stencils at an open boundary
TRACER_CASES = {
    "vortex_o3_hllc_2q": lambda: cases.with_tracers(cases.isentropic_vortex(n=36, order=3, flux="hllc"), 2),
    "vortex_o3_rusanov_1q": lambda: cases.with_tracers(cases.isentropic_vortex(n=30, order=3, flux="rusanov"), 1),
    "vortex_o2_1q": lambda: cases.with_tracers(cases.isentropic_vortex(n=30, order=2), 1),
    "vortex_o5_1q": lambda: cases.with_tracers(cases.isentropic_vortex(n=28, order=5), 1),
    "blast_o3_2q": lambda: cases.with_tracers(cases.blast_3d(n=7, order=3, kind="blast"), 2),
    "smooth3d_o2_3q": lambda: cases.with_tracers(cases.blast_3d(n=6, order=2, kind="smooth"), 3),
    "polytrope_wb_1q": lambda: cases.with_tracers(cases.polytrope_2d(n=30, order=3, well_balanced=True, amplitude=1e-3), 1),
    "vortex_fluxbc_1q": lambda: cases.with_tracers(cases.isentropic_vortex(n=22, order=3, ghost_ring_cells=0, flux_bc="flux"), 1),
}


@pytest.fixture(scope="module", params=sorted(TRACER_CASES))
def tracer_setup(request):
    case = TRACER_CASES[request.param]()
    st = case.ensure_stencils()
    ctx = z.CudaContext(case.grid, st, case.params)
    ora = _oracle(case, st)
    yield request.param, case, st, ctx, ora
    ctx.close()


def _avars_scale(case):
    """flux scale of the scalars: m q * a / inradius."""
    gamma = case.params.gamma
    u = case.u0
    p = (gamma - 1.0) * (u[:, 4] - 0.5 * (u[:, 1:4] ** 2).sum(axis=1) / u[:, 0])
    a = np.sqrt(gamma * np.abs(p) / u[:, 0])
    return np.abs(case.a0).max(axis=0) * (a / case.grid.array("inradii")).max()


def test_tracer_rate_of_change(tracer_setup):
    name, case, st, ctx, ora = tracer_setup
    n, na = case.grid.n_cells, case.params.n_avars
    roc = z.CudaEulerRateOfChange(ctx)
    tend = z.AllVariables(n, n_avars=na)
    state = z.AllVariables(n, case.u0, case.a0)
    roc.compute(tend, state, accumulate=False)
    ref, ref_a = ora.rate_of_change_av(case.u0, case.a0)
    scale = tendency_scales(case.u0, case.params.gamma, case.grid.array("inradii"))
    assert (np.abs(tend.cvars - ref).max(axis=0) / scale).max() < 1e-12, name
    err_a = np.abs(tend.avars - ref_a).max(axis=0) / _avars_scale(case)
    assert np.abs(ref_a).max() > 0.0
    assert err_a.max() < 1e-12, (name, err_a)
    # accumulate semantics on both halves of AllVariables
    rng = np.random.default_rng(1)
    base, base_a = rng.normal(size=(n, 5)), rng.normal(size=(n, na))
    tend2 = z.AllVariables(n, base.copy(), base_a.copy())
    roc.compute(tend2, state, accumulate=True)
    assert np.allclose(tend2.avars - base_a, tend.avars, rtol=0, atol=1e-12 * max(np.abs(ref_a).max(), 1.0))
    assert np.allclose(tend2.cvars - base, tend.cvars, rtol=0, atol=1e-12 * max(np.abs(ref).max(), 1.0))
    # the scatter is conservative on the device too
    vol = case.grid.array("volumes")
    tot = (vol[:, None] * tend.avars).sum(axis=0)
    gross = (vol[:, None] * np.abs(tend.avars)).sum(axis=0)
    assert np.all(np.abs(tot) <= 1e-11 * gross), (name, tot, gross)


def test_tracer_runge_kutta_steps(tracer_setup):
    name, case, st, ctx, ora = tracer_setup
    n, na = case.grid.n_cells, case.params.n_avars
    n_steps = 8 if case.grid.n_dims == 2 else 4
    rk = z.CudaRungeKutta(ctx, case.method)
    if case.frozen_bc:
        z.FrozenBC(ctx, z.AllVariables(n, case.u0, case.a0))
        ora.set_frozen_bc_av(case.u0, case.a0)
    rk.upload(z.AllVariables(n, case.u0, case.a0))
    u_ref, a_ref = case.u0.copy(), case.a0.copy()
    dt = ora.cfl_dt(u_ref, case.cfl)
    for _ in range(n_steps):
        dt_next, bad = rk.step(0.0, dt, case.cfl)
        assert not bad
        u_ref, a_ref = ora.rk_step_av(case.method, u_ref, a_ref, dt)
        dt_ref = ora.cfl_dt(u_ref, case.cfl)
        assert abs(dt_next - dt_ref) <= 1e-11 * dt_ref
        dt = dt_ref
    out = rk.download()
    vol = case.grid.array("volumes")
    assert rel_err(out.avars, a_ref).max() < 1e-11, (name, rel_err(out.avars, a_ref))
    assert rel_l1(out.avars, a_ref, vol).max() < 1e-11, (name, rel_l1(out.avars, a_ref, vol))
    sc = state_scales(case.u0, case.params.gamma)
    assert (np.abs(out.cvars - u_ref).max(axis=0) / sc).max() < 1e-11, name
    assert np.abs(out.avars - case.a0).max() > 0.0
    if case.frozen_bc:
        gh = case.grid.is_ghost
        assert np.array_equal(out.avars[gh], case.a0[gh]) and np.array_equal(out.cvars[gh], case.u0[gh])
    # TimeIntegration::compute_step with host buffers == the resident-state step, bit for bit
    rk.upload(z.AllVariables(n, case.u0, case.a0))
    rk.step(0.0, dt)
    a = rk.download()
    b = rk.compute_step(z.AllVariables(n, case.u0, case.a0), 0.0, dt)
    assert np.array_equal(a.avars, b.avars) and np.array_equal(a.cvars, b.cvars)


def test_tracers_leave_the_euler_path_untouched():
    """The conserved variables' tendency is bit-identical with and without scalars on board (the scalar kernels only
    read the traces), and a context with scalars refuses the cvars-only entry point instead of dropping them."""
    case = cases.isentropic_vortex(n=30, order=3)
    st = case.ensure_stencils()
    n = case.grid.n_cells
    ctx0 = z.CudaContext(case.grid, st, case.params)
    t0 = z.AllVariables(n)
    z.CudaEulerRateOfChange(ctx0).compute(t0, z.AllVariables(n, case.u0), accumulate=False)
    ctx0.close()
    cases.with_tracers(case, 2)
    ctx = z.CudaContext(case.grid, st, case.params)
    t1 = z.AllVariables(n, n_avars=2)
    z.CudaEulerRateOfChange(ctx).compute(t1, z.AllVariables(n, case.u0, case.a0), accumulate=False)
    assert np.array_equal(t0.cvars, t1.cvars)
    with pytest.raises(z.ZfvmError):
        z._capi.check(z._capi.lib.zfvm_rate_of_change(ctx._h, z._capi.ptr_f64(t1.cvars), z._capi.ptr_f64(case.u0), 0.0, 0))
    ctx.close()


HEATING_CASES = {
    # Heating alone (no gravity model): the cell-local source pass runs with zero potentials
    "smooth3d_o3_heating": lambda: cases.blast_3d(n=6, order=3, kind="smooth"),
    "vortex_o3_heating": lambda: cases.isentropic_vortex(n=24, order=3),
    # next to the gravity source, with and without the equilibrium background in rho(x)
    "atmosphere_wb_heating": lambda: cases.stellar_atmosphere_3d(n=6, order=3, well_balanced=True),
    "polytrope_nowb_heating": lambda: cases.polytrope_2d(n=24, order=3, well_balanced=False),
}


@pytest.mark.parametrize("name", sorted(HEATING_CASES))
def test_heating(name):
    case = HEATING_CASES[name]()
    c = case.grid.array("cell_centers")
    r = np.linalg.norm(c, axis=1)
    case.params.heating = (0.3, float(np.quantile(r, 0.3)), float(np.quantile(r, 0.7)))  # a shell cutting through cells
    st = case.ensure_stencils()
    n = case.grid.n_cells
    ctx = z.CudaContext(case.grid, st, case.params)
    ora = _oracle(case, st)
    tend = z.AllVariables(n)
    z.CudaEulerRateOfChange(ctx).compute(tend, z.AllVariables(n, case.u0), accumulate=False)
    ref = ora.rate_of_change(case.u0)
    scale = tendency_scales(case.u0, case.params.gamma, case.grid.array("inradii"))
    assert (np.abs(tend.cvars - ref).max(axis=0) / scale).max() < 1e-12, name
    # the heating term itself, isolated: rate with - rate without
    case.params.heating = None
    ctx0 = z.CudaContext(case.grid, st, case.params)
    t0 = z.AllVariables(n)
    z.CudaEulerRateOfChange(ctx0).compute(t0, z.AllVariables(n, case.u0), accumulate=False)
    d = tend.cvars - t0.cvars
    d_ref = ref - _oracle(case, st).rate_of_change(case.u0)
    assert np.abs(d_ref[:, 4]).max() > 0.0
    assert np.abs(d[:, 4] - d_ref[:, 4]).max() <= 1e-13 * scale[4] + 1e-12 * np.abs(d_ref[:, 4]).max()
    assert np.abs(d[:, :4]).max() <= 1e-12 * scale[:4].max()
    # RK steps with the source on
    case.params.heating = (0.3, float(np.quantile(r, 0.3)), float(np.quantile(r, 0.7)))
    rk = z.CudaRungeKutta(ctx, case.method)
    z.FrozenBC(ctx, z.AllVariables(n, case.u0))
    ora.set_frozen_bc(case.u0)
    rk.upload(z.AllVariables(n, case.u0))
    u_ref = case.u0.copy()
    dt = ora.cfl_dt(u_ref, case.cfl)
    for _ in range(3):
        rk.step(0.0, dt)
        u_ref = ora.rk_step(case.method, u_ref, dt)
    sc = state_scales(case.u0, case.params.gamma)
    assert (np.abs(rk.download().cvars - u_ref).max(axis=0) / sc).max() < 1e-11, name
    ctx.close()
    ctx0.close()


@pytest.mark.parametrize("wb", [True, False])
def test_equilibrium_flux_bc(wb):
    """EquilibriumFluxBC on a polytrope without a ghost ring: parity with the oracle, and (well-balanced) the
    hydrostatic state is closed to round-off -- boundary cells included -- and stays put under RK steps."""
    case = cases.polytrope_2d(n=28, order=3, well_balanced=wb, ghost=False, flux_bc="equilibrium")
    st = case.ensure_stencils()
    n = case.grid.n_cells
    ctx = z.CudaContext(case.grid, st, case.params)
    ora = _oracle(case, st)
    tend = z.AllVariables(n)
    z.CudaEulerRateOfChange(ctx).compute(tend, z.AllVariables(n, case.u0), accumulate=False)
    ref = ora.rate_of_change(case.u0)
    scale = tendency_scales(case.u0, case.params.gamma, case.grid.array("inradii"))
    assert (np.abs(tend.cvars - ref).max(axis=0) / scale).max() < 1e-12
    lr = case.grid.array("left_right")
    boundary = np.zeros(n, dtype=bool)
    boundary[lr[case.grid.n_interior_edges:, 0]] = True
    if wb:
        assert (np.abs(tend.cvars).max(axis=0) / scale).max() < 1e-11
    else:
        assert np.abs(tend.cvars[boundary, 1:3]).max() > 1e-6 * scale[1]
    rk = z.CudaRungeKutta(ctx, "ssp3")
    rk.upload(z.AllVariables(n, case.u0))
    u_ref = case.u0.copy()
    dt = ora.cfl_dt(u_ref, 0.4)
    for _ in range(5):
        rk.step(0.0, dt)
        u_ref = ora.rk_step("ssp3", u_ref, dt)
    u = rk.download().cvars
    sc = state_scales(case.u0, case.params.gamma)
    assert (np.abs(u - u_ref).max(axis=0) / sc).max() < 1e-11
    if wb:
        assert (np.abs(u - case.u0).max(axis=0) / sc).max() < 1e-12
    assert ctx.counters()["eq_failures"] == 0
    ctx.close()


def test_equilibrium_flux_bc_needs_gravity():
    case = cases.isentropic_vortex(n=12, order=3, ghost_ring_cells=0)
    case.params.flux_bc = "equilibrium"
    with pytest.raises(z.ZfvmError, match="EquilibriumFluxBC needs a gravity model"):
        z.CudaContext(case.grid, case.ensure_stencils(), case.params)


# ---- LocalRCParams: history-dependent refresh of the local equilibrium (local_reconstruction.hpp:87-100) ----------------
RC_CASES = {
    "polytrope_wb": lambda: cases.polytrope_2d(n=30, order=3, well_balanced=True, amplitude=1e-3),
    "atmosphere_wb": lambda: cases.stellar_atmosphere_3d(n=6, order=3, well_balanced=True),
    "atmosphere_wb_o4": lambda: cases.stellar_atmosphere_3d(n=7, order=4, well_balanced=True),   # cooperative kernel
    "polytrope_nowb": lambda: cases.polytrope_2d(n=30, order=3, well_balanced=False),            # only the scale is cached
    "vortex": lambda: cases.isentropic_vortex(n=30, order=3),
}


@pytest.mark.parametrize("name", sorted(RC_CASES))
@pytest.mark.parametrize("spr,threshold", [(3, 1e300), (4, 2e-4), (2, 0.0)])
def test_steps_per_recompute(name, spr, threshold):
    """`steps_per_recompute` / `recompute_threshold` (JSON "reconstruction", euler_experiment_impl.hpp:83-90): the cell's
    equilibrium, its stencil averages, its point values and the characteristic scale are kept between evaluations of the
    rate of change.  Four SSP3 steps (twelve evaluations) against the oracle, which restates recompute_equilibrium:
    refresh every third evaluation only; every fourth or when a cell has drifted by 2e-4 scaled units (some do, most do
    not); threshold 0 = refresh always, which must reproduce the steps_per_recompute = 1 run bit for bit."""
    case = RC_CASES[name]()
    case.params.steps_per_recompute, case.params.recompute_threshold = spr, threshold
    st = case.ensure_stencils()
    n = case.grid.n_cells
    ctx = z.CudaContext(case.grid, st, case.params)
    ora = _oracle(case, st)
    rk = z.CudaRungeKutta(ctx, case.method)
    z.FrozenBC(ctx, z.AllVariables(n, case.u0))
    ora.set_frozen_bc(case.u0)
    rk.upload(z.AllVariables(n, case.u0))
    u_ref = case.u0.copy()
    dt = 0.8 * ora.cfl_dt(u_ref, case.cfl)
    for _ in range(4):
        rk.step(0.0, dt)
        u_ref = ora.rk_step(case.method, u_ref, dt)
    u = rk.download().cvars
    assert ctx.counters()["eq_failures"] == ora.eq_failures()
    ctx.close()
    sc = state_scales(case.u0, case.params.gamma)
    assert (np.abs(u - u_ref).max(axis=0) / sc).max() < 1e-11, (name, np.abs(u - u_ref).max(axis=0) / sc)
    if threshold == 0.0:
        case.params.steps_per_recompute, case.params.recompute_threshold = 1, 0.0
        ctx = z.CudaContext(case.grid, st, case.params)
        rk = z.CudaRungeKutta(ctx, case.method)
        z.FrozenBC(ctx, z.AllVariables(n, case.u0))
        rk.upload(z.AllVariables(n, case.u0))
        for _ in range(4):
            rk.step(0.0, dt)
        assert np.array_equal(rk.download().cvars, u)
        ctx.close()
    elif name.endswith("_wb") or name.endswith("_o4"):
        # the cached equilibrium is history: the run differs from the refresh-always run (else the test tests nothing)
        case.params.steps_per_recompute, case.params.recompute_threshold = 1, 0.0
        ctx = z.CudaContext(case.grid, st, case.params)
        rk = z.CudaRungeKutta(ctx, case.method)
        z.FrozenBC(ctx, z.AllVariables(n, case.u0))
        rk.upload(z.AllVariables(n, case.u0))
        for _ in range(4):
            rk.step(0.0, dt)
        assert not np.array_equal(rk.download().cvars, u)
        ctx.close()
