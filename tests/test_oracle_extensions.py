"""CPU checks of the oracle's restatement of the SURVEY.md 8f rows that widen the path: advected scalars (a27),
`Heating`, `EquilibriumFluxBC`.  The reference holds no unit test for any of the three (test/zisa/unit_test has
no tracer, heating or flux-bc file), so they are pinned by the properties their definitions imply:

  flux/hllc.hpp:178-197           tracer flux of m q = c rho is c times the HLLC mass flux, in every wave pattern
  local_reconstruction.hpp:127-147  a scalar is reconstructed like a lone conserved variable with unit scaling
  fvm_loops/flux_loop.hpp:157-192   the scatter is conservative; a uniform state has zero tendency
  model/heating.hpp:30-44,54-80     dE/dt = average(rho * epsilon * [r0 <= |x| <= r1])
  boundary/equilibrium_flux_bc.hpp:37-63  closes a hydrostatic domain without a ghost ring to round-off
"""
import numpy as np
import pytest

import zisafvm_b200 as z
from zisafvm_b200 import cases
from oracle import binding
from oracle.binding import Oracle


def _hllc_mass_flux(gamma, uL, uR):
    nf = np.zeros(5)
    binding.lib().oracle_hllc(gamma, binding._p(binding.f64(uL)), binding._p(binding.f64(uR)), binding._p(nf))
    return nf[0]


@pytest.mark.parametrize("uL,uR", [
    ((1.0, 0.1, 0.0, 0.0, 2.5), (0.125, -0.05, 0.0, 0.0, 0.25)),      # subsonic, s* > 0
    ((0.125, 0.02, 0.0, 0.0, 0.25), (1.0, -0.3, 0.1, 0.0, 2.5)),      # subsonic, s* < 0
    ((1.0, 3.0, 0.2, 0.0, 7.0), (0.9, 2.8, 0.0, 0.1, 6.5)),           # supersonic to the right
    ((1.0, -3.0, 0.2, 0.0, 7.0), (0.9, -2.8, 0.0, 0.1, 6.5)),         # supersonic to the left
])
def test_tracer_flux_is_concentration_times_mass_flux(uL, uR):
    gamma = 1.4
    for c in (0.0, 0.37, 2.0):
        fq = binding.hllc_tracer_flux(gamma, uL, uR, c * uL[0], c * uR[0])
        assert fq == pytest.approx(c * _hllc_mass_flux(gamma, uL, uR), rel=1e-13, abs=1e-15)


def test_tracer_flux_upwinds_contact():
    # pure contact moving right: the left concentration is carried
    gamma, u = 1.4, 0.5
    uL = (1.0, u, 0.0, 0.0, 2.5 + 0.5 * u * u)
    uR = (0.5, 0.5 * u, 0.0, 0.0, 2.5 + 0.25 * u * u)
    assert binding.hllc_tracer_flux(gamma, uL, uR, 0.8, 0.1) == pytest.approx(0.8 * u, rel=1e-12)


@pytest.fixture(scope="module")
def vortex_tracers():
    case = cases.with_tracers(cases.isentropic_vortex(n=16, order=3), n_avars=2)
    st = case.ensure_stencils()
    return case, st, Oracle(case.grid, st, case.params)


def test_uniform_state_has_zero_tracer_tendency(vortex_tracers):
    case, st, ora = vortex_tracers
    n = case.grid.n_cells
    u = np.tile(np.array([1.3, 0.4, -0.2, 0.0, 3.0]), (n, 1))
    a = np.tile(np.array([0.7 * 1.3, 0.2 * 1.3]), (n, 1))
    t, ta = ora.rate_of_change_av(u, a)
    interior = ~case.grid.is_ghost
    scale = 1.3 * np.sqrt(1.4 * 1.0) / case.grid.array("inradii").min()
    assert np.abs(ta[interior]).max() <= 1e-12 * scale
    assert np.abs(t[interior]).max() <= 1e-12 * 3.0 * scale


def test_tracer_scatter_is_conservative(vortex_tracers):
    case, st, ora = vortex_tracers
    t, ta = ora.rate_of_change_av(case.u0, case.a0)
    vol = case.grid.array("volumes")
    total = (vol[:, None] * ta).sum(axis=0)
    gross = (vol[:, None] * np.abs(ta)).sum(axis=0)
    assert np.all(gross > 0.0)
    assert np.all(np.abs(total) <= 1e-12 * gross)
    # the conserved variables are untouched by the presence of tracers
    t_plain = ora.rate_of_change(case.u0)
    assert np.array_equal(t, t_plain)


def test_scalar_reconstruction_is_single_variable_reconstruction():
    # same stencils / weights: with unit scaling and all other variables constant the density polynomial of the
    # Euler reconstruction (IS = max over variables = IS of rho) equals the scalar polynomial of a tracer = rho
    case = cases.isentropic_vortex(n=14, order=3)
    case.params.scaling = "unity"
    case.params.n_avars = 1
    st = case.ensure_stencils()
    ora = Oracle(case.grid, st, case.params)
    u = np.zeros_like(case.u0)
    u[:, 0] = case.u0[:, 0]
    u[:, 4] = 2.5
    a = u[:, :1].copy()
    ora.rate_of_change_av(u, a)
    tp = ora.tracer_polys(6)[:, 0, :]
    coeffs, _ = ora.reconstruct(u, 6)
    assert np.abs(tp - coeffs[:, :, 0]).max() <= 4e-16   # (the compiler contracts the two instantiations differently)


def test_rk_step_carries_tracers_and_frozen_bc(vortex_tracers):
    case, st, ora = vortex_tracers
    ora.set_frozen_bc_av(case.u0, case.a0)
    dt = ora.cfl_dt(case.u0, 0.4)
    u1, a1 = ora.rk_step_av("ssp3", case.u0, case.a0, dt)
    ghost = case.grid.is_ghost
    assert np.array_equal(a1[ghost], case.a0[ghost]) and np.array_equal(u1[ghost], case.u0[ghost])
    assert np.abs(a1[~ghost] - case.a0[~ghost]).max() > 0.0
    # forward Euler == u0 + dt * rate
    t, ta = ora.rate_of_change_av(case.u0, case.a0)
    _, a_fe = ora.rk_step_av("forward_euler", case.u0, case.a0, dt)
    assert np.allclose(a_fe[~ghost], (case.a0 + dt * ta)[~ghost], rtol=0, atol=1e-15)
    ora.set_frozen_bc(None)


def test_heating_of_uniform_gas():
    case = cases.blast_3d(n=4, order=2, kind="smooth")
    st = case.ensure_stencils()
    n = case.grid.n_cells
    u = np.tile(np.array([1.7, 0.0, 0.0, 0.0, 2.5]), (n, 1))
    base = Oracle(case.grid, st, case.params).rate_of_change(u)
    case.params.heating = (0.25, 0.0, 10.0)           # the whole box lies inside the heated shell
    heated = Oracle(case.grid, st, case.params).rate_of_change(u)
    d = heated - base
    assert np.abs(d[:, :4]).max() == 0.0
    assert np.allclose(d[:, 4], 1.7 * 0.25, rtol=1e-13)
    case.params.heating = (0.25, 0.4, 0.8)            # partial shell: 0 <= dE/dt <= rho eps, both attained
    shell = Oracle(case.grid, st, case.params).rate_of_change(u) - base
    assert shell[:, 4].min() >= 0.0 and shell[:, 4].max() <= 1.7 * 0.25 * (1 + 1e-13)
    assert (shell[:, 4] == 0.0).any() and (shell[:, 4] > 0.0).any()


def test_equilibrium_flux_bc_closes_a_hydrostatic_domain():
    # polytrope without a ghost ring: FluxLoop + well-balanced source leave the boundary cells' exterior faces open;
    # EquilibriumFluxBC supplies exactly the equilibrium pressure there
    case = cases.polytrope_2d(n=20, order=3, well_balanced=True, ghost=False)
    grid = case.grid
    assert not grid.is_ghost.any()
    st = case.ensure_stencils()
    tables = cases.gravity_tables(grid, case.params.gravity)
    gamma = case.params.gamma
    p = (gamma - 1.0) * case.u0[:, 4]
    a = np.sqrt(gamma * p / case.u0[:, 0])
    scale = (p / grid.array("inradii")).max()         # magnitude of the pressure terms that cancel
    open_res = Oracle(grid, st, case.params, tables).rate_of_change(case.u0)
    case.params.flux_bc = "equilibrium"
    closed = Oracle(grid, st, case.params, tables).rate_of_change(case.u0)
    boundary = np.zeros(grid.n_cells, dtype=bool)
    lr = grid.array("left_right")
    boundary[lr[grid.n_interior_edges:, 0]] = True
    assert np.abs(open_res[boundary, 1:3]).max() > 1e-3 * scale          # open: O(p / h) imbalance
    assert np.abs(closed[:, 1:3]).max() <= 1e-11 * scale                 # closed: round-off
    assert np.array_equal(closed[~boundary], open_res[~boundary])
    assert a.min() > 0


def test_steps_per_recompute_in_the_oracle():
    """recompute_equilibrium (local_reconstruction.hpp:87-100): threshold 0 refreshes at every evaluation and reproduces
    steps_per_recompute = 1 bit for bit; an unreachable threshold keeps the cached equilibrium between refreshes, which
    changes a perturbed run but leaves an unperturbed hydrostatic state in equilibrium to round-off."""
    from oracle.binding import Oracle

    def run(case, spr, thr, steps=3):
        case.params.steps_per_recompute, case.params.recompute_threshold = spr, thr
        ora = Oracle(case.grid, case.ensure_stencils(), case.params, cases.gravity_tables(case.grid, case.params.gravity))
        ora.set_frozen_bc(case.u0)
        u = case.u0.copy()
        dt = 0.8 * ora.cfl_dt(u, case.cfl)
        for _ in range(steps):
            u = ora.rk_step(case.method, u, dt)
        return u

    pert = cases.polytrope_2d(n=16, order=3, well_balanced=True, amplitude=1e-3)
    base = run(pert, 1, 0.0)
    assert np.array_equal(run(pert, 5, 0.0), base)
    cached = run(pert, 3, 1e300)
    assert not np.array_equal(cached, base)
    assert np.abs(cached - base).max() < 1e-5          # second order in the perturbation over a few steps
    rest = cases.polytrope_2d(n=16, order=3, well_balanced=True, amplitude=0.0)
    u = run(rest, 3, 1e300)
    assert np.abs(u - rest.u0).max() < 1e-12
