"""Timing of the well-balanced configurations (BASELINE configs 2 and 4 shape) on one GPU: ms per kernel and stage."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import zisafvm_b200 as z
from zisafvm_b200 import cases

def run(name, case, steps=5):
    st = case.ensure_stencils()
    n = case.grid.n_cells
    ctx = z.CudaContext(case.grid, st, case.params)
    rk = z.CudaRungeKutta(ctx, case.method)
    z.FrozenBC(ctx, z.AllVariables(n, case.u0))
    rk.upload(z.AllVariables(n, case.u0))
    dt, bad = z.LocalCFL(ctx, case.cfl)()
    for _ in range(2):
        rk.step(0.0, 0.5 * dt)
    ctx.profile(True)
    for _ in range(steps):
        rk.step(0.0, 0.5 * dt)
    kms, kcnt = ctx.profile_read()
    ctx.profile(False)
    per = [kms[i] / max(kcnt[i], 1) for i in range(3)]
    n_int = int((~case.grid.is_ghost).sum())
    print(f"{name}: {n} cells, K1 {per[0]:.3f} K2 {per[1]:.3f} K3 {per[2]:.3f} ms per stage -> {n_int / (sum(per) * 1e-3):.4g} cell-updates/s", flush=True)
    ctx.close()

run("atmosphere 3D o3 WB (gamma 5/3)", cases.stellar_atmosphere_3d(n=40, order=3, well_balanced=True))
run("atmosphere 3D o3 no WB, gravity", cases.stellar_atmosphere_3d(n=40, order=3, well_balanced=False))
run("polytrope 2D o3 WB (gamma 2)", cases.polytrope_2d(n=600, order=3, well_balanced=True))
if os.environ.get("WB_O4", "1") == "1":
    # BASELINE config 4 itself (order 4, well-balanced) and its plain counterpart
    run("atmosphere 3D o4 WB (C4)", cases.stellar_atmosphere_3d(n=32, order=4, well_balanced=True), steps=3)
    run("smooth 3D o4, no gravity", cases.blast_3d(n=32, order=4, kind="smooth"), steps=3)
