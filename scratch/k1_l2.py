"""K1 with HBM out of the picture: a grid whose tile records fit in L2, few CTAs so that every warp walks several tiles.
Phase timers (ZFVM_TILE_PROF=1) give cycles per tile per phase; compare with the bench-size numbers."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["ZFVM_TILE_PROF"] = "1"
os.environ["ZFVM_STREAM_MAX_CTAS"] = sys.argv[2] if len(sys.argv) > 2 else "12"
import zisafvm_b200 as z
from zisafvm_b200 import cases

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
case = cases.blast_3d(n=n, order=3, kind="smooth")
st = case.ensure_stencils()
ctx = z.CudaContext(case.grid, st, case.params)
nc = case.grid.n_cells
rk = z.CudaRungeKutta(ctx, case.method)
z.FrozenBC(ctx, z.AllVariables(nc, case.u0))
rk.upload(z.AllVariables(nc, case.u0))
dt, bad = z.LocalCFL(ctx, case.cfl)()
for rep in range(3):
    for _ in range(3):
        rk.step(0.0, 0.5 * dt)
    ctx.synchronize()
    ctx.profile(True)
    for _ in range(4):
        rk.step(0.0, 0.5 * dt)
    ms, cnt = ctx.profile_read()
    ctx.profile(False)
    print(f"n={n} cells {nc} K1 {ms[0] / cnt[0]:.4f} ms", flush=True)
