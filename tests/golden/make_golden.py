"""Generates the golden vectors of tests/golden/*.npz with the CPU oracle (oracle/zisa_oracle.cpp).

The reference itself cannot be built or run in this image, so these vectors pin the *oracle's*
output at the time of writing: `tests/test_golden.py` checks that the oracle still reproduces them (CPU) and that the
CUDA path matches them (GPU) without needing the oracle at test time.  Seeds and sizes are fixed; run from the
repository root:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle.binding import Oracle  # noqa: E402
from zisafvm_b200 import cases  # noqa: E402

CASES = {
    "vortex_o3_hllc_n12": lambda: cases.isentropic_vortex(n=12, order=3, flux="hllc"),
    "vortex_o3_rusanov_n12": lambda: cases.isentropic_vortex(n=12, order=3, flux="rusanov"),
    "blast_o3_n7": lambda: cases.blast_3d(n=7, order=3, kind="blast"),
    "blast_o2_n7": lambda: cases.blast_3d(n=7, order=2, kind="blast"),
    "polytrope_wb_n12": lambda: cases.polytrope_2d(n=12, order=3, well_balanced=True, amplitude=1e-3),
    # rows that widen the path: advected scalars, Heating, EquilibriumFluxBC
    "vortex_o3_tracers_n12": lambda: cases.with_tracers(cases.isentropic_vortex(n=12, order=3, flux="hllc"), 2),
    "blast_o3_tracer_n6": lambda: cases.with_tracers(cases.blast_3d(n=6, order=3, kind="blast"), 1),
    "smooth3d_o3_heating_n6": lambda: _heated(cases.blast_3d(n=6, order=3, kind="smooth"), (0.3, 0.45, 0.75)),
    "polytrope_wb_eqfluxbc_n12": lambda: cases.polytrope_2d(n=12, order=3, well_balanced=True, amplitude=1e-3,
                                                            ghost=False, flux_bc="equilibrium"),
}


def _heated(case, heating):
    case.params.heating = heating
    return case


def compute(name):
    case = CASES[name]()
    st = case.ensure_stencils()
    tables = cases.gravity_tables(case.grid, case.params.gravity) if case.params.gravity.kind != "none" else None
    ora = Oracle(case.grid, st, case.params, tables)
    dt = ora.cfl_dt(case.u0, case.cfl)
    if case.a0 is not None:  # AllVariables{cvars, avars}
        if case.frozen_bc:
            ora.set_frozen_bc_av(case.u0, case.a0)
        tend, tend_a = ora.rate_of_change_av(case.u0, case.a0)
        u, a = case.u0.copy(), case.a0.copy()
        for _ in range(3):
            u, a = ora.rk_step_av(case.method, u, a, dt)
        return case, dict(u0=case.u0, tendency=tend, dt=np.float64(dt), u3=u, n_cells=np.int64(case.grid.n_cells),
                          a0=case.a0, tendency_a=tend_a, a3=a)
    if case.frozen_bc:
        ora.set_frozen_bc(case.u0)
    tend = ora.rate_of_change(case.u0)
    u = case.u0.copy()
    for _ in range(3):
        u = ora.rk_step(case.method, u, dt)
    return case, dict(u0=case.u0, tendency=tend, dt=np.float64(dt), u3=u, n_cells=np.int64(case.grid.n_cells))


if __name__ == "__main__":
    here = os.path.dirname(os.path.abspath(__file__))
    only = sys.argv[1:]  # optional: names to (re)generate; existing vectors are otherwise left alone
    for name in CASES:
        if only and name not in only:
            continue
        _, data = compute(name)
        np.savez_compressed(os.path.join(here, name + ".npz"), **data)
        print(name, data["u0"].shape, "dt", float(data["dt"]))
