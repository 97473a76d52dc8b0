import sys, numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import zisafvm_b200 as z
from zisafvm_b200 import cases
from oracle.binding import Oracle
from util import rel_err
for name, mk in [("polytrope_nowb", lambda: cases.polytrope_2d(n=36, order=3, well_balanced=False)),
                 ("polytrope_wb", lambda: cases.polytrope_2d(n=36, order=3, well_balanced=True)),
                 ("polytrope_wb_pert", lambda: cases.polytrope_2d(n=36, order=3, well_balanced=True, amplitude=1e-3)),
                 ("atmosphere_wb", lambda: cases.stellar_atmosphere_3d(n=7, order=3, well_balanced=True))]:
    case = mk(); st = case.ensure_stencils(); case.params.keep_polynomials = True
    ctx = z.CudaContext(case.grid, st, case.params)
    ora = Oracle(case.grid, st, case.params, cases.gravity_tables(case.grid, case.params.gravity))
    n = case.grid.n_cells
    tend = z.AllVariables(n)
    z.CudaEulerRateOfChange(ctx).compute(tend, z.AllVariables(n, case.u0), accumulate=False)
    ref = ora.rate_of_change(case.u0)
    coef, scale = ctx.polynomials(); rc, rs = ora.reconstruct(case.u0, coef.shape[1])
    print(name, "n", n, "eqfail", ctx.counters()["eq_failures"], ora.eq_failures())
    print("  max|tend ref|", np.abs(ref).max(axis=0), " max|diff|", np.abs(tend.cvars-ref).max(axis=0))
    gh = case.grid.is_ghost
    print("  nonghost max|tend ref|", np.abs(ref[~gh]).max(axis=0), " max|diff|", np.abs(tend.cvars-ref)[~gh].max(axis=0))
    print("  coef max|ref|", np.abs(rc).max(axis=(0,1)), " max|diff|", np.abs(coef-rc).max(axis=(0,1)))
    print("  scale diff", np.abs(scale-rs).max())
    src = ctx.work_array("source").reshape(n,5)
    print("  src max", np.abs(src).max(axis=0))
    ctx.close()
