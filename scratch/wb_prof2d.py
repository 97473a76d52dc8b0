"""One well-balanced 2D polytrope step for ncu."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import zisafvm_b200 as z
from zisafvm_b200 import cases

case = cases.polytrope_2d(n=int(os.environ.get("WB_N", "600")), order=3, well_balanced=True)
st = case.ensure_stencils()
n = case.grid.n_cells
ctx = z.CudaContext(case.grid, st, case.params)
rk = z.CudaRungeKutta(ctx, case.method)
z.FrozenBC(ctx, z.AllVariables(n, case.u0))
rk.upload(z.AllVariables(n, case.u0))
dt, bad = z.LocalCFL(ctx, case.cfl)()
for _ in range(2):
    rk.step(0.0, 0.5 * dt)
ctx.synchronize()
ctx.close()
