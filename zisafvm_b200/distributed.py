"""Domain decomposition and halo plans for the multi-GPU residual path (one process per GPU).

Host-side mirror of the reference's distributed set-up:

* ``partition_by_sfc``        <->  ``compute_partitioned_grid_by_sfc`` (contiguous balanced chunks of the
  Hilbert-ordered cells, src/zisa/parallelization/domain_decomposition.cpp:577-609);
* ``partition_by_metis``      <->  ``compute_partitioned_grid`` (METIS k-way on the stencil graph, :27-113,250-270);
* ``extract_subdomain``       <->  ``extract_subgrid`` / ``extract_stencils`` / ``StencilBasedIndicator``
  (:300-326, :412-447): owned cells first, then the halo -- every cell a stencil of an owned cell or of a
  face-neighbour of an owned cell reads -- grouped contiguously per owner;
* ``save_partitioned_grid`` / ``load_local_grid``  <->  the partition tool and its reader (src/domain_decomposition.cpp:36-113,
  src/zisa/parallelization/local_grid.cpp:11-58): one ``subgrid-%04d.msh.h5`` per part, an oversized chunk with ``partition``
  and ``global_cell_indices``;
* ``HaloPlan`` / ``connect``  <->  ``make_mpi_halo_exchange`` (src/zisa/mpi/parallelization/
  mpi_halo_exchange.cpp:203-251): ranks tell each other which of their cells they need, by global index.

The exchange itself (pack kernel + ncclSend/ncclRecv, overlapped with the reconstruction of the tiles that read
no halo row) lives in ``libzfvm_b200.so`` (csrc/capi_comm.cu); this file only builds index lists.  The lists are
plain numpy, so the same plan drives the ``gloo`` CPU tests (tests/test_distributed_cpu.py).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Callable, List, Optional, Sequence

import numpy as np

from . import _capi
from ._capi import check, lib
from .grid import Grid, QRDegrees, StencilFamilies, compute_stencil_families, cube_mesh, hilbert_permutation

FLAG_INTERIOR, FLAG_GHOST, FLAG_GHOST_L1 = 1, 2, 4


def partition_by_sfc(n_cells: int, n_parts: int) -> np.ndarray:
    """Owner rank of every cell: balanced contiguous chunks (the cells are assumed Hilbert ordered)."""
    part = np.zeros(n_cells, dtype=np.int32)
    chunk, n_large = divmod(n_cells, n_parts)
    for k in range(n_parts):
        lo = k * chunk + min(k, n_large)
        hi = (k + 1) * chunk + min(k + 1, n_large)
        part[lo:hi] = k
    return part


def partition_by_metis(grid: Grid, stencils: Optional[StencilFamilies], n_parts: int) -> np.ndarray:
    """Owner rank of every cell from ``METIS_PartGraphKway`` with the reference's options (OBJTYPE_VOL, NCUTS 10,
    NITER 20, UFACTOR 100) on the stencil graph, or on the face-neighbour graph when ``stencils`` is None
    (``compute_partition_full_stencil``, src/zisa/parallelization/domain_decomposition.cpp:27-113).  Cells of a part
    keep their relative (Hilbert) order, like ``compute_cell_permutation`` (:160-176)."""
    part = np.zeros(grid.n_cells, dtype=np.int32)
    check(lib.zfvm_partition_kway(grid._h, stencils._h if stencils is not None else None, int(n_parts),
                                  part.ctypes.data_as(_capi.c_int32_p)))
    return part


def has_metis() -> bool:
    return bool(lib.zfvm_has_metis())


@dataclass
class HaloPlan:
    """Index lists of one rank's halo exchange (``HaloReceivePart`` / ``HaloSendPart`` of the reference)."""

    peers: np.ndarray          # [n_peers] ranks this rank exchanges with, ascending
    recv_begin: np.ndarray     # [n_peers] first local row filled by peer p
    recv_end: np.ndarray       # [n_peers]
    need_global: List[np.ndarray]  # per peer: global indices of rows [recv_begin, recv_end), in row order
    send_offset: Optional[np.ndarray] = None  # [n_peers + 1]
    send_index: Optional[np.ndarray] = None   # local (owned) rows packed for peer p, concatenated

    @property
    def n_send(self) -> int:
        return 0 if self.send_offset is None else int(self.send_offset[-1])

    @property
    def n_recv(self) -> int:
        return int((self.recv_end - self.recv_begin).sum())


@dataclass
class SubDomain:
    rank: int
    n_ranks: int
    grid: Grid
    stencils: StencilFamilies
    n_owned: int
    global_index: np.ndarray   # [n_local] global cell index of every local cell
    owner: np.ndarray          # [n_local] owning rank
    halo: HaloPlan
    local_to_src: np.ndarray   # [n_local] index into the mesh the sub-domain was cut from

    @property
    def n_local(self) -> int:
        return int(self.global_index.size)

    @property
    def counted(self) -> np.ndarray:
        """Owned, non-ghost cells: the cells this rank's updates count for."""
        m = np.zeros(self.n_local, dtype=bool)
        m[: self.n_owned] = True
        return m & ((self.grid.array("cell_flags") & FLAG_GHOST) == 0)


def extract_subdomain(n_dims: int, vertices: np.ndarray, vertex_indices: np.ndarray, owner: np.ndarray,
                      global_index: np.ndarray, rank: int, n_ranks: int, qr: QRDegrees, stencil_params,
                      physical_ghost: Optional[np.ndarray] = None, seed: int = 0) -> SubDomain:
    """Cut rank ``rank``'s sub-domain out of a mesh that contains its owned cells and enough of their surroundings.

    ``owner[i]`` is the owning rank of mesh cell ``i`` (``-1``: unknown / beyond the region any stencil may
    reach -- needing such a cell is an error), ``global_index[i]`` its global number, ``physical_ghost[i]`` marks
    the ghost cells of physical boundaries (FrozenBC rings), which are owned like any other cell.
    """
    vi = np.ascontiguousarray(vertex_indices, dtype=np.int32).reshape(-1, n_dims + 1)
    owner = np.asarray(owner, dtype=np.int32)
    gidx = np.asarray(global_index, dtype=np.int64)
    n_src = vi.shape[0]
    phys = np.zeros(n_src, dtype=bool) if physical_ghost is None else np.asarray(physical_ghost, dtype=bool)
    owned = owner == rank

    # stencils on the source mesh: full families only where the reference keeps them (interior || ghost_cell_l1,
    # stencil_family.cpp:108-114); with everything not owned masked as ghost that is owned cells plus the face
    # neighbours of owned interior cells -- the "l1" layer whose reconstruction is recomputed redundantly.
    src_grid = Grid(n_dims, vertices, vi, qr)
    src_grid.mask_ghost_cells(~owned | phys)
    src_st = compute_stencil_families(src_grid, stencil_params, seed)
    flags = src_grid.array("cell_flags").copy()
    full = ((flags & FLAG_INTERIOR) != 0) | ((flags & FLAG_GHOST_L1) != 0)

    # cells some full stencil reads (used members only) + the l1 layer itself
    ns = src_st.n_stencils
    l2g, local = src_st.array("l2g"), src_st.array("local")
    size, off = src_st.array("size"), src_st.array("local_off")
    needed = np.zeros(n_src, dtype=bool)
    needed[owned] = True
    needed[full] = True
    rows = np.nonzero(full)[0]
    for k in range(ns):
        sk = size[rows, k]
        for j in range(int(sk.max()) if rows.size else 0):
            sel = rows[sk > j]
            needed[l2g[sel, local[sel, off[k] + j]]] = True
    halo = np.nonzero(needed & ~owned)[0]
    if np.any(owner[halo] < 0):
        raise ValueError("extract_subdomain: a stencil reaches beyond the part of the mesh with known owners; "
                         "enlarge the overlap region")
    # halo rows grouped per owner, by global index within a group (any order both sides agree on would do)
    halo = halo[np.lexsort((gidx[halo], owner[halo]))]
    own_idx = np.nonzero(owned)[0]
    sel = np.concatenate([own_idx, halo]).astype(np.int32)
    n_owned = int(own_idx.size)

    # local mesh: compress vertices; the source grid's vertex order per cell is already the standard orientation
    src_vi = src_grid.array("vertex_indices")[sel]
    used, inv = np.unique(src_vi.reshape(-1), return_inverse=True)
    loc_vertices = np.ascontiguousarray(src_grid.array("vertices")[used])
    loc_vi = inv.reshape(-1, n_dims + 1).astype(np.int32)
    grid = Grid(n_dims, loc_vertices, loc_vi, qr)
    if not np.array_equal(used[grid.array("vertex_indices")], src_vi):
        raise RuntimeError("extract_subdomain: the vertex order of a cell changed during extraction")
    grid.set_flags(flags[sel])
    stencils = StencilFamilies.extract(src_st, grid, sel)

    own_of_halo = owner[halo]
    peers = np.unique(own_of_halo).astype(np.int32)
    recv_begin = np.array([n_owned + np.searchsorted(own_of_halo, p, "left") for p in peers], dtype=np.int64)
    recv_end = np.array([n_owned + np.searchsorted(own_of_halo, p, "right") for p in peers], dtype=np.int64)
    need = [gidx[halo[b - n_owned: e - n_owned]].copy() for b, e in zip(recv_begin, recv_end)]
    plan = HaloPlan(peers, recv_begin, recv_end, need)
    return SubDomain(rank, n_ranks, grid, stencils, n_owned, gidx[sel].copy(), owner[sel].copy(), plan, sel)


def subgrid_file(dirname: str, n_parts: int, rank: int) -> str:
    """``<dirname>/partitioned/<n_parts>/subgrid-%04d.msh.h5`` (mpi_numerical_experiment.hpp:309-315)."""
    import os

    return os.path.join(dirname, "partitioned", str(int(n_parts)), "subgrid-%04d.msh.h5" % int(rank))


def save_partitioned_grid(dirname: str, grid: Grid, owner: np.ndarray, n_parts: int, layers: int = 12) -> List[str]:
    """The reference's partition tool (``save_partitioned_grid``, src/domain_decomposition.cpp:36-113): one sub-grid
    file per part holding the part's cells followed by an *oversized* surrounding -- here every cell within ``layers``
    face-neighbour layers of the part, grouped per owner -- with the owner (``partition``) and the global index
    (``global_cell_indices``) of every cell.  The loading rank recomputes its stencils on that chunk and keeps what they
    need (``load_local_grid``); the chunk must therefore contain the search region of the part's outermost stencils
    (the reference oversizes with a large central stencil per owned cell, :466-505).  Returns the file names."""
    import os

    from .grid import write_subgrid_h5

    owner = np.asarray(owner, dtype=np.int64)
    nb = grid.array("neighbours")
    verts, vi = grid.array("vertices"), grid.array("vertex_indices")
    safe = np.maximum(nb, 0)
    names = []
    for p in range(int(n_parts)):
        owned = owner == p
        if not owned.any():
            raise ValueError(f"save_partitioned_grid: part {p} owns no cell")
        mask = owned.copy()
        for _ in range(int(layers)):
            grown = (mask[safe] & (nb >= 0)).any(axis=1)
            if not (grown & ~mask).any():
                break
            mask |= grown
        halo = np.nonzero(mask & ~owned)[0]
        halo = halo[np.lexsort((halo, owner[halo]))]   # grouped per owner (make_halo reads runs of equal owners)
        sel = np.concatenate([np.nonzero(owned)[0], halo])
        used, inv = np.unique(vi[sel].reshape(-1), return_inverse=True)
        name = subgrid_file(dirname, n_parts, p)
        os.makedirs(os.path.dirname(name), exist_ok=True)
        write_subgrid_h5(name, grid.n_dims, verts[used], inv.reshape(-1, grid.n_dims + 1), owner[sel], sel)
        names.append(name)
    return names


def load_local_grid(path: str, rank: int, n_ranks: int, qr: QRDegrees, stencil_params, boundary_mask=None,
                    seed: int = 0) -> SubDomain:
    """``zisa::load_local_grid`` (src/zisa/parallelization/local_grid.cpp:11-58): read rank ``rank``'s oversized chunk,
    mask everything it does not own (and the physical boundary cells ``boundary_mask`` marks) as ghost, compute the
    stencil families on the chunk, keep the cells they need (``StencilBasedIndicator``), and extract grid, stencils and
    the distributed-grid arrays for them.  ``boundary_mask``: boolean array over the chunk's cells, or a callable
    ``(vertices, vertex_indices, global_cell_indices) -> mask``."""
    from .grid import read_subgrid_h5

    nd, verts, vi, part, gci = read_subgrid_h5(path)
    n_mine = int((part == rank).sum())
    if n_mine == 0 or not np.all(part[:n_mine] == rank):
        raise ValueError(f"{path}: the cells of part {rank} must come first (is this rank {rank}'s file?)")
    if int(part.max()) >= n_ranks:
        raise ValueError(f"{path}: written for more than {n_ranks} parts")
    phys = None
    if boundary_mask is not None:
        phys = boundary_mask(verts, vi, gci) if callable(boundary_mask) else np.asarray(boundary_mask, dtype=bool)
    return extract_subdomain(nd, verts, vi, part.astype(np.int32), gci, rank, n_ranks, qr, stencil_params,
                             physical_ghost=phys, seed=seed)


def complete_halo_plan(sub: SubDomain, requests: Sequence[dict]) -> None:
    """Fill the send side of ``sub.halo`` from every rank's request table.

    ``requests[r]`` is rank r's ``{peer: global indices it needs from peer}`` (``request_table``); rows are packed
    for a peer in exactly the order the peer asked for them, which is the order of its receive rows.
    """
    order = np.argsort(sub.global_index[: sub.n_owned], kind="stable")
    sorted_gid = sub.global_index[: sub.n_owned][order]
    senders = sorted(r for r in range(sub.n_ranks) if r != sub.rank and sub.rank in requests[r])
    if set(senders) != set(int(p) for p in sub.halo.peers):
        # the reference assumes symmetric neighbourhoods too (mpi_halo_exchange.cpp:203-251 posts one send and
        # one receive per neighbour); stencils are not symmetric, so allow one-sided peers by merging the sets
        all_peers = sorted(set(senders) | set(int(p) for p in sub.halo.peers))
        rb, re_, need = [], [], []
        for p in all_peers:
            hit = np.nonzero(sub.halo.peers == p)[0]
            if hit.size:
                rb.append(sub.halo.recv_begin[hit[0]])
                re_.append(sub.halo.recv_end[hit[0]])
                need.append(sub.halo.need_global[hit[0]])
            else:
                rb.append(sub.n_owned)
                re_.append(sub.n_owned)
                need.append(np.zeros(0, dtype=np.int64))
        sub.halo.peers = np.array(all_peers, dtype=np.int32)
        sub.halo.recv_begin = np.array(rb, dtype=np.int64)
        sub.halo.recv_end = np.array(re_, dtype=np.int64)
        sub.halo.need_global = need
    send_offset = [0]
    chunks = []
    for p in sub.halo.peers:
        want = np.asarray(requests[int(p)].get(sub.rank, np.zeros(0, dtype=np.int64)), dtype=np.int64)
        pos = np.searchsorted(sorted_gid, want)
        if want.size and (np.any(pos >= sorted_gid.size) or np.any(sorted_gid[np.minimum(pos, sorted_gid.size - 1)] != want)):
            raise ValueError(f"rank {sub.rank}: rank {int(p)} asks for cells this rank does not own")
        chunks.append(order[pos].astype(np.int32))
        send_offset.append(send_offset[-1] + want.size)
    sub.halo.send_offset = np.array(send_offset, dtype=np.int64)
    sub.halo.send_index = np.concatenate(chunks).astype(np.int32) if chunks else np.zeros(0, dtype=np.int32)


def request_table(sub: SubDomain) -> dict:
    return {int(p): g for p, g in zip(sub.halo.peers, sub.halo.need_global)}


def exchange_requests(sub: SubDomain, group=None) -> None:
    """All ranks publish their request tables (torch.distributed, any backend) and complete their plans."""
    import torch.distributed as dist

    tables = [None] * sub.n_ranks
    dist.all_gather_object(tables, request_table(sub), group=group)
    complete_halo_plan(sub, tables)


def connect(sub: SubDomain, ctx, group=None) -> None:
    """Give the device context its NCCL communicator and halo plan (``zfvm_comm_init`` + ``zfvm_set_halo``)."""
    import torch.distributed as dist

    if sub.halo.send_offset is None:
        exchange_requests(sub, group)
    uid = C.create_string_buffer(128)
    if sub.rank == 0:
        check(lib.zfvm_nccl_unique_id(uid))
    box = [uid.raw]
    dist.broadcast_object_list(box, src=0, group=group)
    check(lib.zfvm_comm_init(ctx._h, box[0], sub.rank, sub.n_ranks))
    h = sub.halo
    n_peers = int(h.peers.size)
    peers = np.ascontiguousarray(h.peers, dtype=np.int32)
    rb = np.ascontiguousarray(h.recv_begin, dtype=np.int64)
    re_ = np.ascontiguousarray(h.recv_end, dtype=np.int64)
    so = np.ascontiguousarray(h.send_offset, dtype=np.int64)
    si = np.ascontiguousarray(h.send_index, dtype=np.int32)
    if si.size == 0:
        si = np.zeros(1, dtype=np.int32)
    check(lib.zfvm_set_halo(ctx._h, sub.n_owned, n_peers, peers.ctypes.data_as(C.POINTER(C.c_int)),
                            rb.ctypes.data_as(_capi.c_int64_p), re_.ctypes.data_as(_capi.c_int64_p),
                            so.ctypes.data_as(_capi.c_int64_p), si.ctypes.data_as(_capi.c_int32_p)))


def halo_exchange_host(sub: SubDomain, state: np.ndarray, group=None) -> None:
    """The exchange on host arrays over torch.distributed point-to-point (``gloo`` in the CPU tests): what
    ``zfvm_halo_exchange`` does on the device with a pack kernel and ncclSend / ncclRecv."""
    import torch
    import torch.distributed as dist

    h = sub.halo
    ops, keep = [], []
    for k, p in enumerate(h.peers):
        b, e = int(h.recv_begin[k]), int(h.recv_end[k])
        if e > b:
            t = torch.from_numpy(state[b:e])  # contiguous rows: received in place
            ops.append(dist.P2POp(dist.irecv, t, int(p), group=group))
        s0, s1 = int(h.send_offset[k]), int(h.send_offset[k + 1])
        if s1 > s0:
            buf = torch.from_numpy(np.ascontiguousarray(state[h.send_index[s0:s1]]))
            keep.append(buf)
            ops.append(dist.P2POp(dist.isend, buf, int(p), group=group))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()


# ---- weak-scaling lattice of boxes (bench.py, BASELINE config 5) ----------------------------------------------------
def rank_lattice(n_ranks: int):
    shapes = {1: (1, 1, 1), 2: (2, 1, 1), 4: (2, 2, 1), 8: (2, 2, 2)}
    if n_ranks not in shapes:
        raise ValueError(f"n_ranks must be one of {sorted(shapes)}")
    return shapes[n_ranks]


def box_subdomain(rank: int, n_ranks: int, n: int, stencil_params, qr: QRDegrees, overlap: int = 4,
                  ghost_cubes: int = 2, jitter: float = 0.1, seed: int = 0, lattice=None):
    """Rank ``rank``'s box of a global lattice of ``P = px*py*pz`` boxes of ``n^3`` cubes (6 Kuhn tetrahedra each).

    Every rank generates only its own box plus ``overlap`` layers of cubes towards its neighbours (vertex jitter is
    a function of the global lattice position, so all ranks see the same global mesh), Hilbert-orders it, and cuts
    its sub-domain out of it.  Returns ``(SubDomain, h, (gx, gy, gz))``.
    """
    px, py, pz = lattice or rank_lattice(n_ranks)
    P = (px, py, pz)
    pos = (rank % px, (rank // px) % py, rank // (px * py))
    G = (px * n, py * n, pz * n)
    h = 1.0 / max(G)
    lo = [max(pos[d] * n - overlap, 0) for d in range(3)]
    hi = [min((pos[d] + 1) * n + overlap, G[d]) for d in range(3)]
    shape = [hi[d] - lo[d] for d in range(3)]
    verts, vi = cube_mesh(shape[0], shape[1], shape[2], h, jitter=jitter, seed=seed, hilbert=False,
                          offset=tuple(lo), global_shape=G)
    nc = vi.shape[0]
    # natural order of the generator: cell = 6 * cube + tet, cube = (iz * ny + iy) * nx + ix
    cube = np.arange(nc, dtype=np.int64) // 6
    tet = np.arange(nc, dtype=np.int64) % 6
    cx = cube % shape[0] + lo[0]
    cy = (cube // shape[0]) % shape[1] + lo[1]
    cz = cube // (shape[0] * shape[1]) + lo[2]
    gid = ((cz * G[1] + cy) * G[0] + cx) * 6 + tet
    owner = ((cz // n) * py + (cy // n)) * px + (cx // n)
    phys = ((cx < ghost_cubes) | (cx >= G[0] - ghost_cubes) | (cy < ghost_cubes) | (cy >= G[1] - ghost_cubes) |
            (cz < ghost_cubes) | (cz >= G[2] - ghost_cubes)) if ghost_cubes > 0 else np.zeros(nc, dtype=bool)
    # Hilbert order (src/renumber_grid.cpp) of the vertex-average centres
    centers = verts[vi].mean(axis=1)
    perm = hilbert_permutation(3, centers)
    sub = extract_subdomain(3, verts, vi[perm], owner[perm].astype(np.int32), gid[perm], rank, n_ranks, qr,
                            stencil_params, physical_ghost=phys[perm], seed=seed)
    return sub, h, G


@dataclass
class RankRun:
    """What bench.py drives on one rank: the sub-domain, its set-up and its device context."""

    sub: SubDomain
    case: object
    ctx: object
    n_counted: int


def make_weak_scaling_case(rank: int, n_ranks: int, n: int, order: int = 3, kind: str = "blast", device: int = 0,
                           group=None, n_avars: int = 0) -> RankRun:
    """BASELINE config 5, weak scaling: every rank owns an ``n^3``-cube box of the lattice, NCCL halo exchange."""
    from . import cases
    from .grid import WENO_PARAMS
    from .solver import CudaContext

    weno = WENO_PARAMS[f"3d_o{order}"]
    sub, h, G = box_subdomain(rank, n_ranks, n, weno.stencil_family_params, cases.blast_qr(order))
    case = cases.blast_3d_on_grid(sub.grid, order=order, kind=kind, stencils=sub.stencils)
    if n_avars > 0:  # advected scalars as functions of the global position
        cases.with_tracers(case, n_avars, box=((0.0, 0.0, 0.0), tuple(float(g) * h for g in G)))
    ctx = CudaContext(sub.grid, sub.stencils, case.params, device=device)
    connect(sub, ctx, group)
    return RankRun(sub, case, ctx, int(sub.counted.sum()))


def make_strong_scaling_case(rank: int, n_ranks: int, n: int, order: int = 3, kind: str = "blast", device: int = 0,
                             group=None, n_avars: int = 0, partition: str = "sfc") -> RankRun:
    """BASELINE config 5, strong scaling: ONE global ``n^3``-cube mesh (the single-GPU workload), Hilbert-ordered and cut
    into ``n_ranks`` contiguous chunks of the space-filling curve -- the reference's shipped partition path
    (``compute_partitioned_grid_by_sfc``, src/zisa/grid/domain_decomposition.cpp:577-609) -- each rank extracting its
    sub-grid, halo and stencils from the global mesh (``extract_subgrid`` / ``extract_stencils``, :300-326,412-447).
    ``partition = "metis" | "metis_stencils"``: the reference's other partitioner, METIS k-way (:27-113)."""
    from . import cases
    from .solver import CudaContext

    if kind == "atmosphere":  # well-balanced stellar atmosphere (BASELINE config 4 shape): gravity tables per sub-grid
        g_case = cases.stellar_atmosphere_3d(n=n, order=order, well_balanced=True)
    else:
        g_case = cases.blast_3d(n=n, order=order, kind=kind)
    if n_avars > 0:
        cases.with_tracers(g_case, n_avars)
    g = g_case.grid
    n_cells = g.n_cells
    if partition == "sfc":
        part = partition_by_sfc(n_cells, n_ranks)
    elif partition == "metis":        # METIS k-way on the face-neighbour graph (compute_partitioned_grid(grid, n_parts), :266-270)
        part = partition_by_metis(g, None, n_ranks)
    elif partition == "metis_stencils":  # ... on the stencil graph (:250-264): needs the global stencils on every rank
        part = partition_by_metis(g, g_case.ensure_stencils(), n_ranks)
    else:
        raise ValueError(f"unknown partition '{partition}'")
    sub = extract_subdomain(g.n_dims, g.array("vertices").copy(), g.array("vertex_indices").copy(), part, np.arange(n_cells),
                            rank, n_ranks, g.qr, g_case.params.weno.stencil_family_params, physical_ghost=g.is_ghost.copy())
    exchange_requests(sub, group)
    case = cases.Case(g_case.name, sub.grid, g_case.params, g_case.u0[sub.global_index].copy(), g_case.method, g_case.cfl,
                      stencils=sub.stencils, a0=None if g_case.a0 is None else g_case.a0[sub.global_index].copy())
    ctx = CudaContext(sub.grid, sub.stencils, case.params, device=device)
    connect(sub, ctx, group)
    return RankRun(sub, case, ctx, int(sub.counted.sum()))
