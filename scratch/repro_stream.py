import sys, numpy as np
sys.path.insert(0, '.')
import zisafvm_b200 as z
from zisafvm_b200 import cases
case = cases.blast_3d(n=int(sys.argv[1]) if len(sys.argv) > 1 else 7, order=3)
st = case.ensure_stencils()
ctx = z.CudaContext(case.grid, st, case.params)
roc = z.CudaEulerRateOfChange(ctx)
t = z.AllVariables(case.grid.n_cells)
roc.compute(t, z.AllVariables(case.grid.n_cells, case.u0), accumulate=False)
print("ok", np.abs(t.cvars).sum())
