"""K1 experiments on one built context: the run-time knobs of recon_tile_kernel (read by the library at every launch)
and the phase timers of warp 0.  Usage: python scratch/k1_knobs.py [n] [order]; prints ms per K1 / K2 / K3 launch."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["ZFVM_GRAPH"] = "0"  # the knobs are read at launch: no replayed graphs here
import numpy as np

import zisafvm_b200 as z
from zisafvm_b200 import cases

n = int(sys.argv[1]) if len(sys.argv) > 1 else 118
order = int(sys.argv[2]) if len(sys.argv) > 2 else 3
t0 = time.perf_counter()
case = cases.blast_3d(n=n, order=order, kind="blast", ghosts_last=bool(os.environ.get("ZFVM_KNOB_GHOSTS_LAST")))
st = case.ensure_stencils()
ctx = z.CudaContext(case.grid, st, case.params)
nc = case.grid.n_cells
counted = int((~case.grid.is_ghost).sum())
rk = z.CudaRungeKutta(ctx, case.method)
z.FrozenBC(ctx, z.AllVariables(nc, case.u0))
rk.upload(z.AllVariables(nc, case.u0))
dt_next, bad = z.LocalCFL(ctx, case.cfl)()
dt = 0.5 * dt_next
print(f"setup {time.perf_counter() - t0:.1f} s, {nc} cells, {counted} counted", flush=True)
_, alg = ctx.memory_info()


def measure(label, env):
    for k in list(os.environ):
        if k.startswith("ZFVM_TILE") or k in ("ZFVM_RECON",):
            del os.environ[k]
    os.environ.update(env)
    for _ in range(2):
        rk.step(0.0, dt)
    ctx.synchronize()
    ctx.profile(True)
    for _ in range(4):
        rk.step(0.0, dt)
    ms, cnt = ctx.profile_read()
    ctx.profile(False)
    per = [m / max(c, 1) for m, c in zip(ms, cnt)]
    print(f"{label:40s} K1 {per[0]:.3f} ms  K2 {per[1]:.3f}  K3 {per[2]:.3f}  stage {sum(per):.3f}", flush=True)
    rk.upload(z.AllVariables(nc, case.u0))


CONFIGS = [
    ("default", {}),
    ("l2_ahead 4608", {"ZFVM_TILE_L2_AHEAD": "4608"}),
    ("l2_ahead 4608 sector", {"ZFVM_TILE_L2_AHEAD": "4608", "ZFVM_TILE_L2_GRAN": "32"}),
    ("l2_ahead 9216 sector", {"ZFVM_TILE_L2_AHEAD": "9216", "ZFVM_TILE_L2_GRAN": "32"}),
    ("l2_ahead 18432 sector", {"ZFVM_TILE_L2_AHEAD": "18432", "ZFVM_TILE_L2_GRAN": "32"}),
    ("default again", {}),
    ("l2_ahead 4608 again", {"ZFVM_TILE_L2_AHEAD": "4608"}),
]
if os.environ.get("ZFVM_KNOB_L2_WHOLE"):
    CONFIGS = [("default", {}), ("whole record prefetch", {"ZFVM_TILE_L2_WHOLE": "1"}),
               ("whole record prefetch, no per-segment prefetch", {"ZFVM_TILE_L2_WHOLE": "1", "ZFVM_TILE_L2_AHEAD": "0"}),
               ("default again", {}), ("whole again", {"ZFVM_TILE_L2_WHOLE": "1", "ZFVM_TILE_L2_AHEAD": "0"})]
if os.environ.get("ZFVM_KNOB_DEFAULT_ONLY"):
    CONFIGS = [(os.environ["ZFVM_KNOB_DEFAULT_ONLY"], {})] * 2
if os.environ.get("ZFVM_TILE_PROF"):
    CONFIGS = [("phase timers", {"ZFVM_TILE_PROF": "1"})] * 2
if os.environ.get("ZFVM_KNOB_GHOSTS_LAST"):
    CONFIGS = CONFIGS[:2]
for label, env in CONFIGS:
    measure(label, env)
