"""GPU properties that do not need the oracle, on grids larger than the oracle can follow in seconds
(BASELINE configs 3/5 shape: 3D blast, CWENO-AO order 3 and 2, HLLC): conservation of the atomic-free
face-flux gather, free-stream preservation, bit-reproducibility, agreement of the stage-by-stage host path with
the fused device path."""
import numpy as np
import pytest

import zisafvm_b200 as z
from zisafvm_b200 import cases

from util import state_scales, tendency_scales

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", params=[(3, 40), (2, 32)], ids=["o3_n40", "o2_n32"])
def big(request):
    order, n = request.param
    case = cases.blast_3d(n=n, order=order, kind="blast")
    st = case.ensure_stencils()
    ctx = z.CudaContext(case.grid, st, case.params)
    yield case, ctx
    ctx.close()


def test_conservation(big):
    """Every interior face flux enters its two cells with opposite signs (flux_loop.hpp:169-190), ghost-ghost faces
    are skipped for both: sum_i vol_i dU_i/dt vanishes up to round-off of the individual contributions."""
    case, ctx = big
    n = case.grid.n_cells
    tend = z.AllVariables(n)
    z.CudaEulerRateOfChange(ctx).compute(tend, z.AllVariables(n, case.u0), accumulate=False)
    vol = case.grid.array("volumes")
    total = (tend.cvars * vol[:, None]).sum(axis=0)
    gross = (np.abs(tend.cvars) * vol[:, None]).sum(axis=0)
    scale = tendency_scales(case.u0, case.params.gamma, case.grid.array("inradii")) * vol.sum()
    assert np.all(np.abs(total) <= 1e-12 * np.maximum(gross, 1e-6 * scale)), (total, gross)


def test_free_stream(big):
    """A constant state has a zero residual (all stencil differences vanish, the closed-surface flux sum cancels):
    <= 1e-12 of the flux scale in every cell that is not a ghost cell."""
    case, ctx = big
    n = case.grid.n_cells
    u = np.tile(np.array([1.3, 0.4, -0.2, 0.3, 2.9]), (n, 1))
    tend = z.AllVariables(n)
    z.CudaEulerRateOfChange(ctx).compute(tend, z.AllVariables(n, u), accumulate=False)
    scale = tendency_scales(u, case.params.gamma, case.grid.array("inradii"))
    interior = ~case.grid.is_ghost
    assert (np.abs(tend.cvars[interior]).max(axis=0) / scale).max() < 1e-12


def test_bit_reproducible_and_host_path(big):
    """Two evaluations give identical bits (no atomics anywhere), accumulate adds exactly, and three SSP3 stages driven
    through RateOfChange::compute + a host-side Butcher sum reproduce the fused device step to round-off."""
    case, ctx = big
    n = case.grid.n_cells
    roc = z.CudaEulerRateOfChange(ctx)
    u0 = z.AllVariables(n, case.u0)
    a, b = z.AllVariables(n), z.AllVariables(n)
    roc.compute(a, u0, accumulate=False)
    roc.compute(b, u0, accumulate=False)
    assert np.array_equal(a.cvars, b.cvars)
    roc.compute(b, u0, accumulate=True)
    assert np.array_equal(b.cvars, a.cvars + a.cvars)

    rk = z.CudaRungeKutta(ctx, case.method)
    z.FrozenBC(ctx, u0)
    rk.upload(u0)
    dt, bad = z.LocalCFL(ctx, case.cfl)()
    assert not bad and dt > 0
    rk.step(0.0, dt)
    u1 = rk.download().cvars
    gh = case.grid.is_ghost
    if case.method == "ssp3":   # runge_kutta.cpp:154-165
        A, bw = [[], [1.0], [0.25, 0.25]], [1 / 6, 1 / 6, 2 / 3]
    else:                       # ssp2, runge_kutta.cpp:145-153
        A, bw = [[], [1.0]], [0.5, 0.5]
    ks, us = [], case.u0.copy()
    for s in range(len(bw)):
        us = case.u0 + dt * sum(c * k for c, k in zip(A[s], ks)) if s > 0 else case.u0.copy()
        us[gh] = case.u0[gh]
        k = z.AllVariables(n)
        roc.compute(k, z.AllVariables(n, us), accumulate=False)
        ks.append(k.cvars.copy())
    ref = case.u0 + dt * sum(c * k for c, k in zip(bw, ks))
    ref[gh] = case.u0[gh]
    sc = state_scales(case.u0, case.params.gamma)
    assert (np.abs(u1 - ref).max(axis=0) / sc).max() < 1e-13
    assert np.array_equal(u1[gh], case.u0[gh])
