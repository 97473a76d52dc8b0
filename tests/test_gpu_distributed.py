"""Two-GPU parity: a domain-decomposed run (NCCL halo exchange per stage, interior tiles overlapped with the
exchange, fused RK update, ncclMin for dt) against the single-domain CPU oracle on the global mesh.
Needs two visible GPUs (``gpurun --gpus 2``); skipped otherwise."""
import os
import socket
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

N_PER_RANK, ORDER, KIND, STEPS, CFL = 6, 3, "smooth", 3, 0.4
N_SFC = 8    # cubes per direction of the global mesh of the SFC-partition test
N_AVARS = 2  # advected scalars ride along: their halo rows travel in the same NCCL group


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out_dir, partition="lattice"):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch
    import torch.distributed as dist

    import zisafvm_b200 as z
    from zisafvm_b200 import distributed as zd

    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        if partition == "sfc_wb":  # the same partition for a well-balanced run with gravity (equilibrium kernels per tile list)
            run = zd.make_strong_scaling_case(rank, world, n=N_SFC, order=ORDER, kind="atmosphere", device=rank, n_avars=N_AVARS)
        elif partition == "metis":  # METIS k-way on the stencil graph (domain_decomposition.cpp:27-113): ragged parts
            run = zd.make_strong_scaling_case(rank, world, n=N_SFC, order=ORDER, kind=KIND, device=rank, n_avars=N_AVARS,
                                              partition="metis_stencils")
        elif partition == "sfc":  # one global mesh, contiguous chunks of the Hilbert curve (the reference's shipped path)
            run = zd.make_strong_scaling_case(rank, world, n=N_SFC, order=ORDER, kind=KIND, device=rank, n_avars=N_AVARS)
        else:
            run = zd.make_weak_scaling_case(rank, world, n=N_PER_RANK, order=ORDER, kind=KIND, device=rank, n_avars=N_AVARS)
        sub, case, ctx = run.sub, run.case, run.ctx
        n = sub.n_local
        rk = z.CudaRungeKutta(ctx, case.method)
        z.FrozenBC(ctx, z.AllVariables(n, case.u0, case.a0))
        u0, a0 = case.u0.copy(), case.a0.copy()
        u0[sub.n_owned:] = 1e300  # halo rows must come from the exchange, not from the upload
        a0[sub.n_owned:] = 1e300
        rk.upload(z.AllVariables(n, u0, a0))
        dt, bad = z.LocalCFL(ctx, CFL)()
        dts = [dt]
        for _ in range(STEPS):
            dt_next, bad = rk.step(0.0, dt, CFL)
            assert not bad
            dt = dt_next
            dts.append(dt)
        out = rk.download()
        u = out.cvars
        np.save(os.path.join(out_dir, f"a_{rank}.npy"), out.avars[: sub.n_owned])
        cnt = ctx.counters()
        # (a METIS part of this small mesh may consist of ghost cells only, or be so ragged that every tile reads a halo
        # row: the overlap split is only asserted for the contiguous partitions)
        assert partition == "metis" or (cnt["tiles_exterior"] > 0 and cnt["tiles_interior"] > 0), cnt
        np.save(os.path.join(out_dir, f"u_{rank}.npy"), u[: sub.n_owned])
        np.save(os.path.join(out_dir, f"gid_{rank}.npy"), sub.global_index[: sub.n_owned])
        np.save(os.path.join(out_dir, f"dt_{rank}.npy"), np.array(dts))
        ctx.close()
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_two_gpu_run_matches_single_domain_oracle(tmp_path):
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp

    import zisafvm_b200 as z
    from oracle.binding import Oracle
    from zisafvm_b200 import cases
    from zisafvm_b200 import distributed as zd
    from zisafvm_b200.grid import cube_mesh

    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)

    # the same global mesh in one piece, natural (generator) order = global index
    px, py, pz = zd.rank_lattice(world)
    G = (px * N_PER_RANK, py * N_PER_RANK, pz * N_PER_RANK)
    h = 1.0 / max(G)
    verts, vi = cube_mesh(G[0], G[1], G[2], h, jitter=0.1, seed=0, hilbert=False, offset=(0, 0, 0), global_shape=G)
    grid = z.Grid(3, verts, vi, cases.blast_qr(ORDER))
    c = np.arange(vi.shape[0]) // 6
    cx, cy, cz = c % G[0], (c // G[0]) % G[1], c // (G[0] * G[1])
    g2 = 2
    grid.mask_ghost_cells((cx < g2) | (cx >= G[0] - g2) | (cy < g2) | (cy >= G[1] - g2) | (cz < g2) | (cz >= G[2] - g2))
    case = cases.blast_3d_on_grid(grid, order=ORDER, kind=KIND)
    cases.with_tracers(case, N_AVARS, box=((0.0, 0.0, 0.0), tuple(float(g) * h for g in G)))
    st = case.ensure_stencils()
    ora = Oracle(grid, st, case.params)
    ora.set_frozen_bc_av(case.u0, case.a0)
    u_ref, a_ref = case.u0.copy(), case.a0.copy()
    dt = ora.cfl_dt(u_ref, CFL)
    dts = [dt]
    for _ in range(STEPS):
        u_ref, a_ref = ora.rk_step_av(case.method, u_ref, a_ref, dt)
        dt = ora.cfl_dt(u_ref, CFL)
        dts.append(dt)

    scale = np.abs(u_ref).max(axis=0)
    for r in range(world):
        u = np.load(tmp_path / f"u_{r}.npy")
        gid = np.load(tmp_path / f"gid_{r}.npy")
        err = np.abs(u - u_ref[gid]).max(axis=0) / scale
        assert err.max() < 1e-11, (r, err)
        a = np.load(tmp_path / f"a_{r}.npy")
        err_a = np.abs(a - a_ref[gid]).max(axis=0) / np.abs(a_ref).max(axis=0)
        assert err_a.max() < 1e-11, (r, err_a)
        # ncclMin of the local CFL steps == the global CFL step
        assert np.allclose(np.load(tmp_path / f"dt_{r}.npy"), dts, rtol=1e-11, atol=0)


@pytest.mark.parametrize("partition", ["sfc", "metis"])
def test_two_gpu_sfc_partition_matches_single_domain_oracle(tmp_path, partition):
    """The same check for the reference's partitions of one global mesh: "sfc" -- the Hilbert-ordered mesh cut into two
    contiguous chunks of the curve, the shipped path -- and "metis" -- METIS k-way on the stencil graph
    (domain_decomposition.cpp:27-113); ragged partition boundaries, halo rows grouped per owner."""
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp

    from oracle.binding import Oracle
    from zisafvm_b200 import cases
    from zisafvm_b200 import distributed as zd

    if partition == "metis" and not zd.has_metis():
        pytest.skip("libzfvm_b200.so was built without METIS")
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path), partition), nprocs=world, join=True)

    case = cases.with_tracers(cases.blast_3d(n=N_SFC, order=ORDER, kind=KIND), N_AVARS)
    st = case.ensure_stencils()
    ora = Oracle(case.grid, st, case.params)
    ora.set_frozen_bc_av(case.u0, case.a0)
    u_ref, a_ref = case.u0.copy(), case.a0.copy()
    dt = ora.cfl_dt(u_ref, CFL)
    dts = [dt]
    for _ in range(STEPS):
        u_ref, a_ref = ora.rk_step_av(case.method, u_ref, a_ref, dt)
        dt = ora.cfl_dt(u_ref, CFL)
        dts.append(dt)
    scale = np.abs(u_ref).max(axis=0)
    seen = np.zeros(case.grid.n_cells, dtype=bool)
    for r in range(world):
        u = np.load(tmp_path / f"u_{r}.npy")
        a = np.load(tmp_path / f"a_{r}.npy")
        gid = np.load(tmp_path / f"gid_{r}.npy")
        seen[gid] = True
        assert (np.abs(u - u_ref[gid]).max(axis=0) / scale).max() < 1e-11, r
        assert (np.abs(a - a_ref[gid]).max(axis=0) / np.abs(a_ref).max(axis=0)).max() < 1e-11, r
        assert np.allclose(np.load(tmp_path / f"dt_{r}.npy"), dts, rtol=1e-11, atol=0)
    assert seen.all()


def test_two_gpu_well_balanced_matches_single_domain_oracle(tmp_path):
    """Well-balanced atmosphere with gravity on two GPUs (SFC partition): the equilibrium kernels, the tile kernel and the
    source kernel run per tile list (interior tiles while the halo is in flight, the rest after it)."""
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp

    from oracle.binding import Oracle
    from zisafvm_b200 import cases

    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path), "sfc_wb"), nprocs=world, join=True)

    case = cases.with_tracers(cases.stellar_atmosphere_3d(n=N_SFC, order=ORDER, well_balanced=True), N_AVARS)
    st = case.ensure_stencils()
    ora = Oracle(case.grid, st, case.params, cases.gravity_tables(case.grid, case.params.gravity))
    ora.set_frozen_bc_av(case.u0, case.a0)
    u_ref, a_ref = case.u0.copy(), case.a0.copy()
    dt = ora.cfl_dt(u_ref, CFL)
    dts = [dt]
    for _ in range(STEPS):
        u_ref, a_ref = ora.rk_step_av(case.method, u_ref, a_ref, dt)
        dt = ora.cfl_dt(u_ref, CFL)
        dts.append(dt)
    gamma = case.params.gamma
    p = (gamma - 1.0) * case.u0[:, 4]
    ra = (case.u0[:, 0] * np.sqrt(gamma * p / case.u0[:, 0])).max()   # acoustic momentum scale (the momenta are ~ 0)
    scale = np.array([case.u0[:, 0].max(), ra, ra, ra, case.u0[:, 4].max()])
    for r in range(world):
        u = np.load(tmp_path / f"u_{r}.npy")
        a = np.load(tmp_path / f"a_{r}.npy")
        gid = np.load(tmp_path / f"gid_{r}.npy")
        assert (np.abs(u - u_ref[gid]).max(axis=0) / scale).max() < 1e-11, r
        assert (np.abs(a - a_ref[gid]).max(axis=0) / np.abs(a_ref).max(axis=0)).max() < 1e-11, r
        assert np.allclose(np.load(tmp_path / f"dt_{r}.npy"), dts, rtol=1e-11, atol=0)


def _worker_host_step(rank, world, port, partition):
    """compute_step with host buffers in a decomposed run: the chunked route (uploads gating the interior tiles, the halo
    exchange posted once every row is up, last stage finished and downloaded chunk by chunk) against the resident step."""
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      ZFVM_HOST_PIPELINE_MIN_CELLS="0", ZFVM_HOST_CHUNKS="5")
    import torch
    import torch.distributed as dist

    import zisafvm_b200 as z
    from zisafvm_b200 import distributed as zd

    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        if partition == "sfc_wb":
            run = zd.make_strong_scaling_case(rank, world, n=N_SFC, order=ORDER, kind="atmosphere", device=rank)
        elif partition == "sfc":
            run = zd.make_strong_scaling_case(rank, world, n=N_SFC, order=ORDER, kind=KIND, device=rank)
        else:
            run = zd.make_weak_scaling_case(rank, world, n=N_PER_RANK, order=ORDER, kind=KIND, device=rank)
        sub, case, ctx = run.sub, run.case, run.ctx
        n, no = sub.n_local, sub.n_owned
        z.FrozenBC(ctx, z.AllVariables(n, case.u0))
        u0 = case.u0.copy()
        u0[no:] = 1e300  # halo rows must come from the exchange, not from the upload
        for method in (case.method, "forward_euler", "rk4"):
            rk = z.CudaRungeKutta(ctx, method)
            rk.upload(z.AllVariables(n, u0))
            dt, bad = z.LocalCFL(ctx, CFL)()
            l0 = ctx.counters()["launches"]
            rk.step(0.0, dt)
            rk.step(dt, dt)
            l1 = ctx.counters()["launches"]
            a2 = rk.download().cvars.copy()
            h0 = torch.from_numpy(u0.copy()).pin_memory().numpy()
            h1 = torch.full((n, 5), float("nan"), dtype=torch.float64).pin_memory().numpy()
            h2 = np.full((n, 5), np.nan)   # pageable destination
            rk.compute_step(z.AllVariables(n, h0), 0.0, dt, out=z.AllVariables(n, h1))
            rk.compute_step(z.AllVariables(n, h1), dt, dt, out=z.AllVariables(n, h2))
            l2 = ctx.counters()["launches"]
            assert np.array_equal(h2[:no], a2[:no]), (method, np.abs(h2[:no] - a2[:no]).max())
            assert np.isfinite(h2).all() and np.abs(h2).max() < 1e200, method   # halo rows of the result: exchanged values
            assert np.array_equal(rk.download().cvars[:no], a2[:no]), method     # the resident state is the step's result
            assert l2 - l1 > l1 - l0, (method, l0, l1, l2)                       # the chunked route really ran
        # RateOfChange::compute with host buffers takes the same chunked route: against the residual on device buffers
        # (overwrite and accumulate contracts; the halo rows of the caller's state come back filled, flux_loop.hpp:100)
        roc = z.CudaEulerRateOfChange(ctx)
        base = np.random.default_rng(1).normal(size=(n, 5))
        dev = torch.device("cuda", rank)
        for accumulate in (False, True):
            s_dev = torch.from_numpy(u0).to(dev)
            t_dev = torch.from_numpy(base).to(dev)
            torch.cuda.synchronize()
            roc.compute_device(t_dev.data_ptr(), s_dev.data_ptr(), 0.0, accumulate=accumulate)
            ctx.synchronize()
            t_ref, s_ref = t_dev.cpu().numpy(), s_dev.cpu().numpy()
            for pinned in (True, False):
                s_host = torch.from_numpy(u0.copy()).pin_memory().numpy() if pinned else u0.copy()
                t_host = torch.from_numpy(base.copy()).pin_memory().numpy() if pinned else base.copy()
                l3 = ctx.counters()["launches"]
                roc.compute(z.AllVariables(n, t_host), z.AllVariables(n, s_host), 0.0, accumulate=accumulate)
                l4 = ctx.counters()["launches"]
                assert np.array_equal(t_host[:no], t_ref[:no]), (accumulate, pinned, np.abs(t_host[:no] - t_ref[:no]).max())
                assert np.array_equal(s_host, s_ref), (accumulate, pinned)   # owned rows untouched, halo rows exchanged
                assert np.abs(s_host).max() < 1e200
                assert l4 > l3
        ctx.close()
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("partition", ["lattice", "sfc", "sfc_wb"])
def test_two_gpu_host_step_matches_resident_step(partition):
    """zfvm_rk_step_host and zfvm_rate_of_change of a multi-rank context take the chunked, copy-overlapped route too:
    bit-identical owned rows to upload + zfvm_rk_step + download for three-, one- and four-stage tableaux (lattice boxes,
    ragged SFC chunks, the well-balanced kernels per tile list) and to the residual on device buffers, the same number
    of NCCL groups per call on every rank."""
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp

    mp.spawn(_worker_host_step, args=(2, _free_port(), partition), nprocs=2, join=True)
