"""Diagnostics for the BASELINE-size parity failures: per-step errors against the oracle, generic-kernel cross-check."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np

import zisafvm_b200 as z
from zisafvm_b200 import cases
from oracle.binding import Oracle
from util import rel_err, tendency_scales


def residual_check(case, label):
    st = case.ensure_stencils()
    n = case.grid.n_cells
    ora = Oracle(case.grid, st, case.params)
    ref = ora.rate_of_change(case.u0)
    scale = tendency_scales(case.u0, case.params.gamma, case.grid.array("inradii"))
    out = {}
    for mode in ("tile", "generic"):
        if mode == "generic":
            os.environ["ZFVM_RECON"] = "generic"
        else:
            os.environ.pop("ZFVM_RECON", None)
        ctx = z.CudaContext(case.grid, st, case.params)
        roc = z.CudaEulerRateOfChange(ctx)
        t = z.AllVariables(n)
        roc.compute(t, z.AllVariables(n, case.u0), accumulate=False)
        out[mode] = t.cvars.copy()
        err = np.abs(t.cvars - ref) / scale
        bad = np.flatnonzero(err.max(axis=1) > 1e-11)
        fl = case.grid.array("cell_flags")
        print(f"[{label}] {mode}: max err {err.max():.3e}, cells > 1e-11: {bad.size}, counters {ctx.counters()}", flush=True)
        if bad.size:
            print("   first bad cells", bad[:10], "flags", fl[bad[:10]], "tile", bad[:10] // 32, "err", err[bad[:10]].max(axis=1))
            print("   bad cells ghost?", ((fl[bad] & 2) != 0).mean(), " n_family", st.array("n_family")[bad[:10]], "order", st.array("order")[bad[:5]])
        ctx.close()
    os.environ.pop("ZFVM_RECON", None)
    d = np.abs(out["tile"] - out["generic"]) / scale
    print(f"[{label}] tile vs generic max {d.max():.3e}", flush=True)
    return ora, st


def steps_check(case, ora, st, n_steps, label):
    n = case.grid.n_cells
    ctx = z.CudaContext(case.grid, st, case.params)
    rk = z.CudaRungeKutta(ctx, case.method)
    z.FrozenBC(ctx, z.AllVariables(n, case.u0))
    ora.set_frozen_bc(case.u0)
    rk.upload(z.AllVariables(n, case.u0))
    u_ref = case.u0.copy()
    dt = ora.cfl_dt(u_ref, case.cfl)
    for s in range(n_steps):
        dt_next, bad = rk.step(0.0, dt, case.cfl)
        u_ref = ora.rk_step(case.method, u_ref, dt)
        dt_ref = ora.cfl_dt(u_ref, case.cfl)
        u = rk.download().cvars
        e = rel_err(u, u_ref)
        rho, p = u_ref[:, 0], (case.params.gamma - 1) * (u_ref[:, 4] - 0.5 * (u_ref[:, 1:4] ** 2).sum(axis=1) / u_ref[:, 0])
        print(f"[{label}] step {s}: bad {bad} dt rel diff {abs(dt_next - dt_ref) / dt_ref:.2e} state err {e.max():.2e} "
              f"oracle min rho {rho.min():.3e} min p {p.min():.3e}", flush=True)
        if bad or e.max() > 1e-6:
            w = np.abs(u - u_ref).max(axis=1)
            worst = np.argsort(-w)[:5]
            print("   worst cells", worst, w[worst], "flags", case.grid.array("cell_flags")[worst])
            break
        dt = dt_ref
    ctx.close()


which = sys.argv[1] if len(sys.argv) > 1 else "all"
if which in ("all", "o3"):
    case = cases.blast_3d(n=64, order=3, kind="blast")
    ora, st = residual_check(case, "o3_n64_blast")
    steps_check(case, ora, st, 20, "o3_n64_blast")
if which in ("all", "o2"):
    case = cases.blast_3d(n=48, order=2, kind="sod")
    ora, st = residual_check(case, "o2_n48_sod")
    steps_check(case, ora, st, 20, "o2_n48_sod")
if which in ("all", "gl"):
    case = cases.blast_3d(n=48, order=3, kind="blast", ghosts_last=True)
    ora, st = residual_check(case, "o3_n48_ghosts_last")

if which in ("o2s",):
    for n in (24, 48):
        case = cases.blast_3d(n=n, order=2, kind="smooth")
        ora, st = residual_check(case, f"o2_n{n}_smooth")
        steps_check(case, ora, st, 20, f"o2_n{n}_smooth")
