"""Golden vectors (tests/golden/*.npz, written by tests/golden/make_golden.py with the CPU oracle): the oracle still
reproduces them (CPU), the CUDA path matches them through the C ABI without the oracle (GPU).
Tolerances: residual <= 1e-12 of the flux scale, three RK steps <= 1e-11 relative per variable (north_star)."""
import importlib.util
import os

import numpy as np
import pytest

from util import active_vars, state_scales, tendency_scales

HERE = os.path.dirname(os.path.abspath(__file__))
_spec = importlib.util.spec_from_file_location("make_golden", os.path.join(HERE, "golden", "make_golden.py"))
make_golden = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(make_golden)
NAMES = sorted(make_golden.CASES)


def _load(name):
    return np.load(os.path.join(HERE, "golden", name + ".npz"))


@pytest.mark.parametrize("name", NAMES)
def test_oracle_reproduces_golden(name):
    g = _load(name)
    case, now = make_golden.compute(name)
    assert int(g["n_cells"]) == case.grid.n_cells
    assert np.array_equal(g["u0"], now["u0"])          # seeded generators and initial data are deterministic
    assert abs(float(g["dt"]) - float(now["dt"])) <= 1e-15 * float(g["dt"])
    sc_t = tendency_scales(case.u0, case.params.gamma, case.grid.array("inradii"))
    assert (np.abs(g["tendency"] - now["tendency"]).max(axis=0) / sc_t).max() < 1e-13
    sc_u = state_scales(case.u0, case.params.gamma)
    assert (np.abs(g["u3"] - now["u3"]).max(axis=0) / sc_u).max() < 1e-13
    if "a0" in g:
        assert np.array_equal(g["a0"], now["a0"])
        sa = np.abs(g["a0"]).max(axis=0)
        assert (np.abs(g["a3"] - now["a3"]).max(axis=0) / sa).max() < 1e-13
        assert (np.abs(g["tendency_a"] - now["tendency_a"]).max(axis=0) / (sa * sc_t[0] / sc_u[0])).max() < 1e-13


@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
def test_cuda_matches_golden(name):
    import zisafvm_b200 as z

    g = _load(name)
    case = make_golden.CASES[name]()
    st = case.ensure_stencils()
    n = case.grid.n_cells
    ctx = z.CudaContext(case.grid, st, case.params)
    a0 = g["a0"] if "a0" in g else None
    na = 0 if a0 is None else a0.shape[1]
    tend = z.AllVariables(n, n_avars=na)
    z.CudaEulerRateOfChange(ctx).compute(tend, z.AllVariables(n, g["u0"], a0), accumulate=False)
    sc_t = tendency_scales(case.u0, case.params.gamma, case.grid.array("inradii"))
    assert (np.abs(tend.cvars - g["tendency"]).max(axis=0) / sc_t).max() < 1e-12
    rk = z.CudaRungeKutta(ctx, case.method)
    if case.frozen_bc:
        z.FrozenBC(ctx, z.AllVariables(n, g["u0"], a0))
    rk.upload(z.AllVariables(n, g["u0"], a0))
    for _ in range(3):
        rk.step(0.0, float(g["dt"]))
    out = rk.download()
    sc_u = state_scales(case.u0, case.params.gamma)
    vs = active_vars(case.grid.n_dims)
    assert (np.abs(out.cvars - g["u3"]).max(axis=0) / sc_u)[vs].max() < 1e-11
    if a0 is not None:
        sa = np.abs(a0).max(axis=0)
        assert (np.abs(tend.avars - g["tendency_a"]).max(axis=0) / (sa * sc_t[0] / sc_u[0])).max() < 1e-12
        assert (np.abs(out.avars - g["a3"]).max(axis=0) / sa).max() < 1e-11
    ctx.close()
