#!/usr/bin/env python
"""Benchmark of the per-RK-stage residual path (BASELINE.json metric: cell-updates/s per RK stage).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...   # the reference algorithm on the host cores (oracle port)

One "step" is one SSP-RK time step = `stages` passes of the hot path (reconstruction, face fluxes, gather +
stage update + boundary condition) over the whole grid.  `value` = owned non-ghost cells x stages x K / time with
the state resident in HBM; `e2e` = the same metric through the reference-facing call RateOfChange::compute
(zfvm_rate_of_change) with HOST buffers, i.e. one H2D copy of the state and one D2H copy of the tendency per stage.

Workload at N=1: BASELINE config[2] "3D Sod/blast on synthetic ~10M-tetrahedra grid, order 3, single B200"
(118^3 cubes x 6 = 9 858 192 tets, CWENO-AO {3,2,2,2,2}, HLLC, SSP3) -- the configuration the north-star's
">= 60 % of HBM roofline on one B200" target is quoted on.  N>1: weak scaling, every rank owns an equal box of a
global lattice (NCCL halo exchange per stage).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

# torchrun pins OMP_NUM_THREADS=1; the host precompute (stencils, pseudo-inverses) is OpenMP code, so give every
# rank its share of the host cores before any OpenMP runtime is loaded
# (the reference arm runs on rank 0 alone and takes all host cores at every N)
if int(os.environ.get("WORLD_SIZE", "1")) > 1:
    _lws = int(os.environ.get("LOCAL_WORLD_SIZE", os.environ.get("WORLD_SIZE", "1")))
    _ref = "reference" in sys.argv[1:] or any(a.startswith("--impl=reference") for a in sys.argv[1:])
    os.environ["OMP_NUM_THREADS"] = str(max(1, (os.cpu_count() or 8) // (1 if _ref else max(_lws, 1))))

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "cell_updates_per_sec_per_rk_stage"
UNIT = "cell-updates/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--n", type=int, default=int(os.environ.get("ZFVM_BENCH_N", "118")), help="cubes per direction per GPU")
    ap.add_argument("--order", type=int, default=3)
    ap.add_argument("--kind", default="blast", help="blast | sod | smooth (3D, BASELINE configs 3/5); vortex2d (BASELINE "
                    "config 1 scaled up: --n squares per direction x 2 triangles; single GPU, extra measurement); "
                    "atmosphere (BASELINE config 4: 3D well-balanced stellar atmosphere, --order 3 or 4; single GPU); "
                    "polytrope2d (BASELINE config 2 scaled up; single GPU)")
    ap.add_argument("--avars", type=int, default=0, help="advected scalars carried along (extra measurement; BASELINE configs have 0)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="N > 1: weak = every rank owns an n^3 box of a lattice (default, the driver's scaling run); strong = "
                         "ONE global n^3 mesh cut into N chunks of the Hilbert curve (the reference's SFC partition path)")
    ap.add_argument("--cpu-n", type=int, default=0, help="cubes (squares) per direction of the bounded CPU sample; 0 = "
                    "chosen from the requested steps so that the CPU run stays within about a minute")
    ap.add_argument("--partition", default="sfc", choices=["sfc", "metis", "metis_stencils"],
                    help="--scaling strong: contiguous chunks of the Hilbert curve (the reference's shipped path) or METIS "
                         "k-way on the face-neighbour / stencil graph (domain_decomposition.cpp:27-113)")
    ap.add_argument("--ghosts-last", action="store_true",
                    help="number the first-order ghost cells of the FrozenBC shell behind the reconstructed cells (like a "
                         "partition's halo) instead of leaving them interleaved along the Hilbert curve: 8 % fewer tiles, "
                         "but the boundary tiles' row lists outgrow 256 entries (16-bit indices): measured no gain")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device: int):
        self.device = device
        self.rows = []
        self.proc = None

    def start(self):
        """Started before the warm-up (nvidia-smi needs ~0.1 s to come up); rows are time-stamped on arrival and
        ``stop`` keeps those that fall inside the timed region."""
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.device}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def stop(self, t_begin: float = 0.0, t_end: float = float("inf")):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        time.sleep(0.05)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        rows = [r for (t, r) in self.rows if t_begin <= t <= t_end]
        window = "timed region"
        if not rows:  # region shorter than the sampling period: take the samples closest to it
            rows = [r for (t, r) in self.rows if t_begin - 0.25 <= t <= t_end + 0.25]
            window = "timed region +- 0.25 s"
        sm, smax, reasons = [], [], set()
        for r in rows:
            try:
                sm.append(float(r[0]))
                smax.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for name, val in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], r[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm), "window": window}


def measured_peak_gbs():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except (OSError, KeyError, ValueError):
        return 6650.0, "fallback (B200_PROFILING.md, 6.65 TB/s)"


def workload_name(args, method: str) -> str:
    """The same string in both arms (the driver pairs the lines by metric and config)."""
    extra = f", {args.avars} advected scalar(s)" if args.avars else ""
    if args.kind == "vortex2d":
        return (f"2D isentropic vortex on [0,10]^2, {args.n}^2 squares x 2 triangles, CWENO-AO order {args.order}, HLLC, "
                f"{method}, FrozenBC ghost ring" + extra)
    if args.kind == "atmosphere":
        return (f"3D well-balanced stellar atmosphere (gamma 5/3, point-mass gravity, isentropic equilibrium) on [-1,1]^3, "
                f"{args.n}^3 cubes x 6 Kuhn tets, CWENO-AO order {args.order}, HLLC, {method}, FrozenBC ghost shell" + extra)
    if args.kind == "polytrope2d":
        return (f"2D well-balanced polytrope (gamma 2) on [-0.6,0.6]^2, {args.n}^2 squares x 2 triangles, CWENO-AO order "
                f"{args.order}, HLLC, {method}, FrozenBC for r > 0.5" + extra)
    per = "in total" if (args.scaling == "strong" and args.gpus > 1) else "per GPU"
    return (f"3D {args.kind} on [0,1]^3, {args.n}^3 cubes x 6 Kuhn tets {per}, CWENO-AO order {args.order} "
            f"{{{args.order},2,2,2,2}}, HLLC, {method}, FrozenBC ghost shell"
            + (" numbered behind the reconstructed cells" if (args.ghosts_last and args.gpus == 1) else "") + extra)


def make_case(args, n: int):
    from zisafvm_b200 import cases

    if args.kind == "vortex2d":
        case = cases.isentropic_vortex(n=n, order=args.order)
    elif args.kind == "atmosphere":
        case = cases.stellar_atmosphere_3d(n=n, order=args.order, well_balanced=True)
    elif args.kind == "polytrope2d":
        case = cases.polytrope_2d(n=n, order=args.order, well_balanced=True)
    else:
        case = cases.blast_3d(n=n, order=args.order, kind=args.kind, ghosts_last=args.ghosts_last)
    if args.avars > 0:
        cases.with_tracers(case, args.avars)
    return case


def measured_traffic(args, n_cells: int, world: int):
    """dram__bytes_read.sum + dram__bytes_write.sum of one K1 launch from the committed ``ncu --set full`` capture
    of this very configuration (profiles/r02_k1_traffic.json), else null."""
    try:
        with open(os.path.join(ROOT, "profiles", "r02_k1_traffic.json")) as f:
            t = json.load(f)
        if world == 1 and t["n"] == args.n and t["order"] == args.order and t["cells"] == n_cells:
            return {"dram_bytes_per_launch": t["dram_bytes_per_launch"], "bytes_per_cell": t["dram_bytes_per_launch"] / n_cells,
                    "source": t["source"]}
    except (OSError, KeyError, ValueError):
        pass
    return None


def rank_box(rank: int, n_ranks: int):
    """Weak scaling: ranks tile a (px, py, pz) arrangement of equal boxes."""
    shapes = {1: (1, 1, 1), 2: (2, 1, 1), 4: (2, 2, 1), 8: (2, 2, 2)}
    if n_ranks not in shapes:
        raise SystemExit(f"--gpus must be one of {sorted(shapes)}")
    px, py, pz = shapes[n_ranks]
    return (px, py, pz), (rank % px, (rank // px) % py, rank // (px * py))


def cpu_sample_n(args, total_steps: int) -> int:
    """Cubes (squares) per direction of the bounded CPU sample: as large as ~60 s of oracle time allow for
    ``total_steps`` steps at a conservative 0.1 M cell-updates/s per host core (3D order 3; 2D is ~2x faster)."""
    if args.cpu_n > 0:
        return args.cpu_n if args.kind not in ("vortex2d", "polytrope2d") else min(args.n, args.cpu_n)
    cores = os.cpu_count() or 8
    budget_updates = 60.0 * 0.1e6 * cores * (0.3 if args.order >= 4 else 1.0) * (0.4 if args.kind == "atmosphere" else 1.0)
    cells = budget_updates / (3.0 * max(total_steps, 1))
    if args.kind in ("vortex2d", "polytrope2d"):
        return int(max(32, min(args.n, (cells / 2.0) ** 0.5)))
    return int(max(16, min(args.n, 64, (cells / 6.0) ** (1.0 / 3.0))))


def cpu_reference_run(args, steps: int, warmup: int):
    """The reference algorithm (CPU oracle port, OpenMP on all host cores) on a bounded sample of the workload."""
    from oracle import binding as ob
    from zisafvm_b200 import cases

    cpu_n = cpu_sample_n(args, steps + warmup)
    case = make_case(args, cpu_n)
    st = case.ensure_stencils()
    tables = cases.gravity_tables(case.grid, case.params.gravity) if case.params.gravity.kind != "none" else None
    ora = ob.Oracle(case.grid, st, case.params, tables)
    n_int = int((~case.grid.is_ghost).sum())
    stages = {"ssp3": 3, "ssp2": 2}[case.method]
    dt = ora.cfl_dt(case.u0, case.cfl)
    if case.a0 is not None:
        ora.set_frozen_bc_av(case.u0, case.a0)
        state = (case.u0, case.a0)
        step = lambda s: ora.rk_step_av(case.method, s[0], s[1], dt)  # noqa: E731
    else:
        ora.set_frozen_bc(case.u0)
        state = case.u0
        step = lambda s: ora.rk_step(case.method, s, dt)  # noqa: E731
    for _ in range(warmup):
        state = step(state)
    t0 = time.perf_counter()
    for _ in range(steps):
        state = step(state)
    el = time.perf_counter() - t0
    value = n_int * stages * steps / el
    shape = f"{cpu_n}^2 squares x 2 triangles" if case.grid.n_dims == 2 else f"{cpu_n}^3 cubes x 6 Kuhn tets"
    sample = (f"{shape} = {case.grid.n_cells} cells of the same {args.kind} order-{args.order} workload, "
              f"{steps} {case.method} steps after {warmup} warm-up steps")
    return value, el / steps * 1e3, ob.num_threads(), sample, case


def run_reference(args):
    """The reference's own algorithm for the path on the host cores (oracle port: the reference cannot be compiled in this
    image).  Rank 0 alone runs it, with all host cores, for exactly --steps / --warmup steps."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warmup = max(1, args.steps), max(0, args.warmup)
    value, ms, cores, sample, case = cpu_reference_run(args, steps, warmup)
    out = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args, case.method) + f" [CPU arm: bounded sample of it, {sample}]",
                   "sample": sample + "; reference algorithm restated on the CPU with OpenMP on all host cores (oracle "
                             "port: the reference binary cannot be built in this image, DESIGN.md section 4); the metric "
                             "is per cell and stage, so the sample size does not enter it beyond cache effects (the sample's "
                             "working set is far larger than the host caches)",
                   "cells": int(case.grid.n_cells), "l2_flush": "state + weights larger than any cache"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(out))


def multi_gpu_parity_check(args, rank: int, world: int, local_rank: int):
    """N > 1: before anything is timed, a small run on the reference's own partition (ONE Hilbert-ordered mesh cut into
    `world` contiguous chunks, halo rows by NCCL send / recv, interior tiles overlapped with the exchange, ncclMin for dt)
    is compared on rank 0 with the single-domain CPU oracle on the same mesh: three steps, CFL-driven dt."""
    import torch.distributed as dist

    import zisafvm_b200 as z
    from zisafvm_b200 import cases
    from zisafvm_b200 import distributed as zd

    n_small, kind, steps, cfl = 10, "smooth", 3, 0.4
    order = args.order if args.order in (2, 3) else 3
    run = zd.make_strong_scaling_case(rank, world, n=n_small, order=order, kind=kind, device=local_rank)
    sub, case, ctx = run.sub, run.case, run.ctx
    n = sub.n_local
    rk = z.CudaRungeKutta(ctx, case.method)
    z.FrozenBC(ctx, z.AllVariables(n, case.u0))
    u0 = case.u0.copy()
    u0[sub.n_owned:] = 1e300  # halo rows must come from the exchange
    rk.upload(z.AllVariables(n, u0))
    dt, bad = z.LocalCFL(ctx, cfl)()
    dts = [dt]
    for _ in range(steps):
        dt, bad = rk.step(0.0, dt, cfl)
        dts.append(dt)
    u = rk.download().cvars[: sub.n_owned]
    gid = sub.global_index[: sub.n_owned]
    ctx.close()
    parts = [None] * world
    dist.all_gather_object(parts, (gid, u, dts))
    if rank != 0:
        return None
    from oracle.binding import Oracle

    g_case = cases.blast_3d(n=n_small, order=order, kind=kind)
    ora = Oracle(g_case.grid, g_case.ensure_stencils(), g_case.params)
    ora.set_frozen_bc(g_case.u0)
    u_ref = g_case.u0.copy()
    d = ora.cfl_dt(u_ref, cfl)
    dts_ref = [d]
    for _ in range(steps):
        u_ref = ora.rk_step(g_case.method, u_ref, d)
        d = ora.cfl_dt(u_ref, cfl)
        dts_ref.append(d)
    scale = np.abs(u_ref).max(axis=0)
    max_rel, max_dt, seen = 0.0, 0.0, 0
    for gid_r, u_r, dts_r in parts:
        max_rel = max(max_rel, float((np.abs(u_r - u_ref[gid_r]).max(axis=0) / scale).max()))
        max_dt = max(max_dt, float(np.abs(np.array(dts_r) / np.array(dts_ref) - 1.0).max()))
        seen += gid_r.size
    return {"max_rel": max_rel, "max_rel_dt": max_dt, "tolerance": 1e-11, "ok": bool(max_rel <= 1e-11 and max_dt <= 1e-11),
            "case": f"3D {kind}, {n_small}^3 cubes x 6 tets, order {order}, SFC partition into {world} chunks, {steps} "
                    f"{g_case.method} steps with LocalCFL dt (ncclMin), vs the single-domain CPU oracle",
            "cells_compared": int(seen), "cells_total": int(g_case.grid.n_cells)}


def run_b200(args):
    import torch
    import torch.distributed as dist

    import zisafvm_b200 as z
    from zisafvm_b200 import cases

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the B200 path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    distributed = world > 1
    if distributed:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        if world != args.gpus:
            raise SystemExit("--gpus must equal WORLD_SIZE under torchrun")

    parity_check = multi_gpu_parity_check(args, rank, world, local_rank) if distributed else None

    t_setup = time.perf_counter()
    setup_parts = None
    if distributed:
        from zisafvm_b200 import distributed as zd

        maker = zd.make_strong_scaling_case if args.scaling == "strong" else zd.make_weak_scaling_case
        extra = {"partition": args.partition} if args.scaling == "strong" else {}
        sub = maker(rank, world, n=args.n, order=args.order, kind=args.kind, device=local_rank, n_avars=args.avars, **extra)
        case, ctx = sub.case, sub.ctx
        n_counted = sub.n_counted
    else:
        case = make_case(args, args.n)
        t_case = time.perf_counter()
        st = case.ensure_stencils()
        t_st = time.perf_counter()
        ctx = z.CudaContext(case.grid, st, case.params, device=local_rank)
        t_ctx = time.perf_counter()
        setup_parts = {"mesh_geometry_initial_data": round(t_case - t_setup, 1), "stencil_search_device_plus_host_indexing": round(t_st - t_case, 1),
                       "records_weights_on_device_upload": round(t_ctx - t_st, 1)}
        n_counted = int((~case.grid.is_ghost).sum())
    n = case.grid.n_cells
    rk = z.CudaRungeKutta(ctx, case.method)
    z.FrozenBC(ctx, z.AllVariables(n, case.u0, case.a0))
    stages = {"ssp3": 3, "ssp2": 2}[case.method]
    rk.upload(z.AllVariables(n, case.u0, case.a0))
    dt_next, bad = z.LocalCFL(ctx, case.cfl)()
    dt = 0.5 * dt_next
    setup_s = time.perf_counter() - t_setup
    dev_bytes, alg_bytes = ctx.memory_info()

    stream = torch.cuda.ExternalStream(ctx.stream(), device=torch.device("cuda", local_rank))

    def barrier():
        ctx.synchronize()
        torch.cuda.synchronize()
        if distributed:
            dist.barrier()

    # ---- device-resident timing --------------------------------------------------------------------------
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for _ in range(args.warmup):
        rk.step(0.0, dt)
    barrier()
    t_begin = time.time()
    launches0 = ctx.counters()["launches"]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        rk.step(0.0, dt)
    e1.record(stream)
    barrier()
    ms_total = e0.elapsed_time(e1)
    clocks = sampler.stop(t_begin, time.time()) if rank == 0 else None
    launches = ctx.counters()["launches"] - launches0
    if distributed:
        tmax = torch.tensor([ms_total], device="cuda", dtype=torch.float64)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        ms_total = float(tmax.item())
        cnt = torch.tensor([n_counted], device="cuda", dtype=torch.int64)
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
        total_counted = int(cnt.item())
        lt = torch.tensor([launches], device="cuda", dtype=torch.int64)
        dist.all_reduce(lt, op=dist.ReduceOp.SUM)
        launches = int(lt.item())
    else:
        total_counted = n_counted
    value = total_counted * stages * args.steps / (ms_total * 1e-3)
    state_ok = not rk.step(0.0, dt, case.cfl)[1]

    # The metric is per RK stage with a fixed dt.  A driver also asks for the next time step once per step (LocalCFL +
    # SanityCheck fused into the last stage's K3, with N > 1 an ncclAllReduce(min) and a 16-byte read-back: the one
    # latency-bound collective of the design): the same steps again with it, reported next to `value`.
    barrier()
    c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    c0.record(stream)
    for _ in range(args.steps):
        rk.step(0.0, dt, case.cfl)
    c1.record(stream)
    barrier()
    ms_cfl = c0.elapsed_time(c1)
    if distributed:
        tmax = torch.tensor([ms_cfl], device="cuda", dtype=torch.float64)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        ms_cfl = float(tmax.item())
    with_cfl = {"value": total_counted * stages * args.steps / (ms_cfl * 1e-3), "ms_per_step": ms_cfl / args.steps,
                "what": "the same steps with dt_next = LocalCFL (+ ncclAllReduce(min) with N > 1) read back every step"}

    # ---- per-kernel timing for the roofline (separate pass; event pairs around every launch) ----------------
    ctx.profile(True)
    prof_steps = max(2, min(args.steps, 5))
    for _ in range(prof_steps):
        rk.step(0.0, dt)
    kms, kcnt = ctx.profile_read()
    ctx.profile(False)
    peak, peak_src = measured_peak_gbs()
    nd = case.grid.n_dims
    D = {2: 4, 3: 10, 4: 20}[args.order] if nd == 3 else {2: 3, 3: 6, 4: 10, 5: 15}[args.order]
    b_k2 = 40.0 * D + ((nd + 1) / 2.0) * (8.0 * (9 + 4 * case.grid.q_f) + 8.0)   # polynomial read + face data (F/2 faces)
    b_k3 = 40.0 + 40.0 * (1.0 + {3: 2.0, 2: 1.5}[stages]) + 40.0        # tendency write + RK sum
    b_k1 = alg_bytes - b_k2 - b_k3
    t_k1 = kms[0] / max(kcnt[0], 1) * 1e-3
    ach_k1 = n_counted * b_k1 / t_k1 / 1e9
    t_stage = sum(kms) / max(kcnt[0], 1) * 1e-3
    tile_capable = (nd == 2 and args.order <= 4) or (nd == 3 and args.order <= 3)
    recon_name = "recon_tile_kernel" if tile_capable else "recon_coop_kernel (four warps per tile)"
    wb = case.params.well_balancing == "isentropic"
    k1_name = (recon_name if case.params.gravity.kind == "none" and case.params.heating is None
               else ("equilibrium kernels E1-E3 + " if wb else "") + recon_name + " + source_kernel")
    traffic = measured_traffic(args, int(n), world)
    roofline = {
        # well-balanced runs: the equilibrium kernels (Newton solve per cell, rows x q_c equilibrium evaluations per cell)
        # are FP64-pipe bound (committed ncu captures: E1 73 %, E2 62.5 % pipe utilisation), so the HBM fraction below is
        # reported for comparison, not as the bound of the stage
        "bound": "fp64" if wb else "hbm", "kernel": k1_name + " (K1: stencil-weight apply + CWENO-AO + traces)",
        "achieved": ach_k1, "peak": peak, "unit": "GB/s", "frac": ach_k1 / peak,
        # DRAM bytes of one K1 launch (ncu dram__bytes_read.sum + dram__bytes_write.sum), a number like `achieved`'s
        # numerator; where it comes from: traffic_detail
        "traffic": (traffic or {}).get("dram_bytes_per_launch"), "traffic_unit": "bytes per K1 launch", "traffic_detail": traffic,
        "peak_source": peak_src, "algorithmic_bytes_per_cell": {"K1": b_k1, "K2": b_k2, "K3": b_k3, "stage": alg_bytes},
        "kernel_ms": {"K1_recon": kms[0] / max(kcnt[0], 1), "K2_flux": kms[1] / max(kcnt[1], 1),
                      "K3_update": kms[2] / max(kcnt[2], 1),
                      "T_tracers": (kms[3] / max(kcnt[3], 1)) if len(kms) > 3 and kcnt[3] else None},
        "stage": {"achieved": n_counted * alg_bytes / t_stage / 1e9, "frac": n_counted * alg_bytes / t_stage / 1e9 / peak},
    }
    if wb:
        roofline["fp64_pipe_utilisation_ncu"] = {
            "eq_solve_kernel": 0.73, "eq_member_tile_smem_kernel": 0.625, "recon_tile_kernel_wb": 0.23,
            "source": "profiles/r01_wb_split_ncu_details_n40.csv, profiles/r01_wb_tile_ncu_details_n40.csv (sm__pipe_fp64_cycles_active)"}

    # ---- end to end through RateOfChange::compute with host buffers ---------------------------------------------
    e2e = None
    if not args.no_e2e:
        roc = z.CudaEulerRateOfChange(ctx)
        h_state = torch.from_numpy(case.u0.copy()).pin_memory()
        h_tend = torch.zeros_like(h_state).pin_memory()
        na = args.avars
        h_a = torch.from_numpy(case.a0.copy()).pin_memory() if na else None
        h_ta = torch.zeros_like(h_a).pin_memory() if na else None
        s_av = z.AllVariables(n, h_state.numpy(), h_a.numpy() if na else None)
        t_av = z.AllVariables(n, h_tend.numpy(), h_ta.numpy() if na else None)
        calls = max(3, min(3 * args.steps, 9))
        for _ in range(2):
            roc.compute(t_av, s_av, 0.0, accumulate=False)
        barrier()
        t0 = time.perf_counter()
        for _ in range(calls):
            roc.compute(t_av, s_av, 0.0, accumulate=False)
        barrier()
        el = time.perf_counter() - t0
        if distributed:
            tm = torch.tensor([el], device="cuda", dtype=torch.float64)
            dist.all_reduce(tm, op=dist.ReduceOp.MAX)
            el = float(tm.item())
        checksum = float(np.abs(t_av.cvars).sum())
        roc_value = total_counted * calls / el
        # TimeIntegration::compute_step with host buffers (one H2D of u0 + one D2H of u1 per time step): the boundary
        # SURVEY.md 8b recommends so that the stages of a step do not cross PCIe
        u_av = z.AllVariables(n, h_state.numpy(), h_a.numpy() if na else None)
        o_av = z.AllVariables(n, h_tend.numpy(), h_ta.numpy() if na else None)
        rk.compute_step(u_av, 0.0, dt, out=o_av)
        barrier()
        t0 = time.perf_counter()
        reps = max(1, min(args.steps, 5))
        for _ in range(reps):
            rk.compute_step(u_av, 0.0, dt, out=o_av)
        barrier()
        el2 = time.perf_counter() - t0
        if distributed:
            tm = torch.tensor([el2], device="cuda", dtype=torch.float64)
            dist.all_reduce(tm, op=dist.ReduceOp.MAX)
            el2 = float(tm.item())
        e2e = {"value": total_counted * stages * reps / el2, "unit": UNIT, "h2d_bytes_per_step": int(n * (40 + 8 * na)),
               "d2h_bytes_per_step": int(n * (40 + 8 * na)),
               "call": "zfvm_rk_step_host (TimeIntegration::compute_step), pinned host buffers, per time step",
               "steps": reps, "state_l1": float(np.abs(o_av.cvars).sum()),
               "rate_of_change": {"value": roc_value, "call": "zfvm_rate_of_change (RateOfChange::compute), pinned host "
                                  "buffers, one H2D + one D2H per stage", "calls": calls,
                                  "h2d_bytes_per_call": int(n * (40 + 8 * na)), "d2h_bytes_per_call": int(n * (40 + 8 * na)),
                                  "tendency_l1": checksum}}

    cpu_baseline = None
    if rank == 0 and not args.no_cpu_baseline and not distributed:
        v, ms, cores, sample, _ = cpu_reference_run(args, steps=2, warmup=1)
        cpu_baseline = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample}

    if rank == 0:
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_total / args.steps, "higher_is_better": True,
            "scaling": args.scaling if distributed else "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {
                "workload": workload_name(args, case.method) + (
                    f"; ONE global mesh of {args.n}^3 cubes cut into {world} chunks of the Hilbert curve (SFC partition)"
                    if distributed and args.scaling == "strong" else ""),
                "cells_per_gpu": int(n), "counted_cells": int(total_counted), "stages_per_step": stages,
                "device_bytes": int(dev_bytes), "setup_seconds": round(setup_s, 1), "setup_parts": setup_parts,
                "time_step": "fixed dt inside the timed region; with the per-step CFL reduction: see with_cfl",
                "l2_flush": "inputs larger than L2 (weights + state >> 126 MB per stage)",
                "parallelism": f"domain decomposition x{world}" if distributed else "single GPU",
            },
            "roofline": roofline, "cpu_baseline": cpu_baseline, "e2e": e2e, "with_cfl": with_cfl, "gpu_launches": int(launches),
            "clocks": clocks, "state_plausible": bool(state_ok),
        }
        if parity_check is not None:
            out["parity_check"] = parity_check
        print(json.dumps(out))
    ctx.close()
    if distributed:
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
