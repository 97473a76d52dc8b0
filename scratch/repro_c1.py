import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import zisafvm_b200 as z
from zisafvm_b200 import cases
n = int(sys.argv[1]) if len(sys.argv) > 1 else 158
case = cases.isentropic_vortex(n=n, order=3)
st = case.ensure_stencils()
nc = case.grid.n_cells
ctx = z.CudaContext(case.grid, st, case.params)
rk = z.CudaRungeKutta(ctx, case.method)
z.FrozenBC(ctx, z.AllVariables(nc, case.u0))
rk.upload(z.AllVariables(nc, case.u0))
dt, bad = z.LocalCFL(ctx, case.cfl)()
for s in range(int(sys.argv[2]) if len(sys.argv) > 2 else 3):
    dt, bad = rk.step(0.0, dt, case.cfl)
    print("step", s, dt, bad, flush=True)
print("ok", np.abs(rk.download().cvars).max())
