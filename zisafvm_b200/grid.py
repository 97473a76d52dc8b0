"""Host-side mirror of the reference's grid / stencil objects, backed by ``libzfvm_b200.so``.

``Grid`` mirrors ``zisa::Grid`` (include/zisa/grid/grid_decl.hpp:34-107) and
``compute_stencil_families`` mirrors src/zisa/reconstruction/stencil_family.cpp:99-117.  All arrays are
zero-copy numpy views of the flattened storage the device layout is built from.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Sequence

import numpy as np

from . import _capi
from ._capi import check, lib, named_array


@dataclass
class QRDegrees:
    """``zisa::QRDegrees{face_deg, volume_deg, moments_deg}``."""

    face_deg: int = 1
    volume_deg: int = 1
    moments_deg: int = 1


class Grid:
    """Flattened unstructured grid of triangles (``n_dims=2``) or tetrahedra (``n_dims=3``)."""

    def __init__(self, n_dims: int, vertices: np.ndarray, vertex_indices: np.ndarray, qr: QRDegrees):
        v = np.ascontiguousarray(vertices, dtype=np.float64).reshape(-1, 3)
        vi = np.ascontiguousarray(vertex_indices, dtype=np.int32).reshape(-1, n_dims + 1)
        h = C.c_void_p()
        check(
            lib.zfvm_grid_from_mesh(
                n_dims, v.shape[0], v.ctypes.data_as(_capi.c_double_p), vi.shape[0], vi.ctypes.data_as(_capi.c_int32_p),
                qr.face_deg, qr.volume_deg, qr.moments_deg, C.byref(h),
            )
        )
        self._h = h
        self.qr = qr
        info = (C.c_int64 * 8)()
        check(lib.zfvm_grid_info(h, info))
        (self.n_dims, self.n_cells, self.n_vertices, self.n_edges, self.n_interior_edges, self.q_c, self.q_f,
         self.n_moments) = (int(x) for x in info)
        self.max_neighbours = self.n_dims + 1

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            lib.zfvm_grid_free(h)
            self._h = None

    def array(self, name: str) -> np.ndarray:
        return named_array(lib.zfvm_grid_get, self._h, name, owner=self)

    def __getattr__(self, name: str):
        # grid.volumes, grid.cell_centers, grid.left_right, ... (names of zisa::Grid members)
        if name.startswith("_"):
            raise AttributeError(name)
        try:
            return self.array(name)
        except _capi.ZfvmError as e:
            raise AttributeError(name) from e

    def mask_ghost_cells(self, mask: np.ndarray) -> None:
        """``zisa::mask_ghost_cells`` (src/zisa/grid/grid.cpp:1122-1136)."""
        m = np.ascontiguousarray(mask, dtype=np.uint8)
        assert m.shape == (self.n_cells,)
        check(lib.zfvm_grid_mask_ghost(self._h, m.ctypes.data_as(_capi.c_uint8_p)))

    def set_flags(self, flags: np.ndarray) -> None:
        f = np.ascontiguousarray(flags, dtype=np.uint8)
        assert f.shape == (self.n_cells,)
        check(lib.zfvm_grid_set_flags(self._h, f.ctypes.data_as(_capi.c_uint8_p)))

    @property
    def is_ghost(self) -> np.ndarray:
        return (self.array("cell_flags") & 2) != 0


def _take_mesh(nv, verts, nc, vi, n_dims):
    v = np.ctypeslib.as_array(verts, shape=(nv.value, 3)).copy()
    c = np.ctypeslib.as_array(vi, shape=(nc.value, n_dims + 1)).copy()
    lib.zfvm_free(C.cast(verts, C.c_void_p))
    lib.zfvm_free(C.cast(vi, C.c_void_p))
    return v, c


def square_mesh(nx: int, ny: int, x0=0.0, x1=1.0, y0=0.0, y1=1.0, jitter=0.15, seed=0, hilbert=True):
    """Jittered triangulation of a rectangle; cells in Hilbert order."""
    nv, nc = C.c_int64(), C.c_int64()
    verts, vi = _capi.c_double_p(), _capi.c_int32_p()
    check(lib.zfvm_mesh_square(nx, ny, x0, x1, y0, y1, jitter, seed, int(hilbert), C.byref(nv), C.byref(verts),
                               C.byref(nc), C.byref(vi)))
    return _take_mesh(nv, verts, nc, vi, 2)


def cube_mesh(nx: int, ny: int, nz: int, h: float, origin=(0.0, 0.0, 0.0), jitter=0.1, seed=0, hilbert=True,
              offset=None, global_shape=None):
    """Jittered Kuhn triangulation of a box of cubes."""
    nv, nc = C.c_int64(), C.c_int64()
    verts, vi = _capi.c_double_p(), _capi.c_int32_p()
    off = (C.c_int * 3)(*offset) if offset is not None else None
    glb = (C.c_int * 3)(*global_shape) if global_shape is not None else None
    check(lib.zfvm_mesh_cube(nx, ny, nz, h, origin[0], origin[1], origin[2], jitter, seed, int(hilbert), off, glb,
                             C.byref(nv), C.byref(verts), C.byref(nc), C.byref(vi)))
    return _take_mesh(nv, verts, nc, vi, 3)


def read_msh_h5(path: str):
    """``zisa::load_grid_gmsh_h5`` (src/zisa/grid/grid.cpp:889-901): (n_dims, vertices [nv][3], vertex_indices [nc][n_dims+1])
    of a ``*.msh.h5`` grid file, read by the library's own HDF5-subset reader; feed them to ``Grid``."""
    nd, nv, nc = C.c_int(), C.c_int64(), C.c_int64()
    verts, vi = _capi.c_double_p(), _capi.c_int32_p()
    check(lib.zfvm_mesh_read_msh_h5(str(path).encode(), C.byref(nd), C.byref(nv), C.byref(verts), C.byref(nc), C.byref(vi)))
    v, c = _take_mesh(nv, verts, nc, vi, nd.value)
    return nd.value, v, c


def write_msh_h5(path: str, n_dims: int, vertices: np.ndarray, vertex_indices: np.ndarray) -> None:
    """Writes a mesh the way src/renumber_grid.cpp:129-132 does (datasets n_dims, vertex_indices, vertices), so that the
    reference can load the synthetic grids of this package."""
    v = np.ascontiguousarray(vertices, dtype=np.float64)
    c = np.ascontiguousarray(vertex_indices, dtype=np.int32)
    check(lib.zfvm_mesh_write_msh_h5(str(path).encode(), int(n_dims), v.shape[0], _capi.ptr_f64(v), c.shape[0],
                                     c.ctypes.data_as(_capi.c_int32_p)))


def read_subgrid_h5(path: str):
    """A sub-grid file ``subgrid-%04d.msh.h5`` of the reference's partition tool (src/domain_decomposition.cpp:70-88;
    ``load_grid`` + ``load_distributed_grid``, src/zisa/parallelization/distributed_grid.cpp:10-17): ``(n_dims, vertices,
    vertex_indices, partition, global_cell_indices)`` -- owner rank and global index of every local cell."""
    nd, nv, nc = C.c_int(), C.c_int64(), C.c_int64()
    verts, vi = _capi.c_double_p(), _capi.c_int32_p()
    part, gci = _capi.c_int64_p(), _capi.c_int64_p()
    check(lib.zfvm_mesh_read_subgrid_h5(str(path).encode(), C.byref(nd), C.byref(nv), C.byref(verts), C.byref(nc),
                                        C.byref(vi), C.byref(part), C.byref(gci)))
    p = np.ctypeslib.as_array(part, shape=(nc.value,)).copy()
    g = np.ctypeslib.as_array(gci, shape=(nc.value,)).copy()
    lib.zfvm_free(C.cast(part, C.c_void_p))
    lib.zfvm_free(C.cast(gci, C.c_void_p))
    v, c = _take_mesh(nv, verts, nc, vi, nd.value)
    return nd.value, v, c, p, g


def write_subgrid_h5(path: str, n_dims: int, vertices: np.ndarray, vertex_indices: np.ndarray, partition: np.ndarray,
                     global_cell_indices: np.ndarray) -> None:
    """Writes what src/domain_decomposition.cpp:80-88 writes: a grid file plus ``partition`` and ``global_cell_indices``
    (64-bit unsigned, the reference's ``int_t``)."""
    v = np.ascontiguousarray(vertices, dtype=np.float64)
    c = np.ascontiguousarray(vertex_indices, dtype=np.int32)
    p = np.ascontiguousarray(partition, dtype=np.int64)
    g = np.ascontiguousarray(global_cell_indices, dtype=np.int64)
    if p.shape != (c.shape[0],) or g.shape != (c.shape[0],):
        raise ValueError("write_subgrid_h5: partition / global_cell_indices need one entry per cell")
    check(lib.zfvm_mesh_write_subgrid_h5(str(path).encode(), int(n_dims), v.shape[0], _capi.ptr_f64(v), c.shape[0],
                                         c.ctypes.data_as(_capi.c_int32_p), p.ctypes.data_as(_capi.c_int64_p),
                                         g.ctypes.data_as(_capi.c_int64_p)))


@dataclass
class StencilFamilyParams:
    """``zisa::StencilFamilyParams{orders, biases, overfit_factors}``."""

    orders: Sequence[int]
    biases: Sequence[str]
    overfit_factors: Sequence[float]


@dataclass
class HybridWENOParams:
    """``zisa::HybridWENOParams`` (include/zisa/reconstruction/hybrid_weno_params.hpp)."""

    stencil_family_params: StencilFamilyParams
    linear_weights: Sequence[float]
    epsilon: float = 1e-10
    exponent: float = 4.0


#: parameter sets used by the reference
WENO_PARAMS = {
    "2d_o2": HybridWENOParams(StencilFamilyParams([2, 2, 2, 2], "cbbb", [3.0, 2.0, 2.0, 2.0]), [100.0, 1.0, 1.0, 1.0]),
    "2d_o3": HybridWENOParams(StencilFamilyParams([3, 2, 2, 2], "cbbb", [2.0, 1.5, 1.5, 1.5]), [100.0, 1.0, 1.0, 1.0]),
    "2d_o4": HybridWENOParams(StencilFamilyParams([4, 2, 2, 2], "cbbb", [2.0, 1.5, 1.5, 1.5]), [100.0, 1.0, 1.0, 1.0]),
    "2d_o5": HybridWENOParams(StencilFamilyParams([5, 2, 2, 2], "cbbb", [2.0, 1.5, 1.5, 1.5]), [100.0, 1.0, 1.0, 1.0]),
    "3d_o2": HybridWENOParams(StencilFamilyParams([2, 2, 2, 2, 2], "cbbbb", [3.0, 2.0, 2.0, 2.0, 2.0]),
                              [100.0, 1.0, 1.0, 1.0, 1.0]),
    "3d_o3": HybridWENOParams(StencilFamilyParams([3, 2, 2, 2, 2], "cbbbb", [2.0, 1.5, 1.5, 1.5, 1.5]),
                              [100.0, 1.0, 1.0, 1.0, 1.0]),
    "3d_o4": HybridWENOParams(StencilFamilyParams([4, 2, 2, 2, 2], "cbbbb", [3.0, 2.0, 2.0, 2.0, 2.0]),
                              [100.0, 1.0, 1.0, 1.0, 1.0]),
    # parameter sets of the reference's own reconstruction tests
    # test/zisa/unit_test/reconstruction/cweno_ao.cpp:144-160: six stencils, two of them central
    "3d_o4_six": HybridWENOParams(StencilFamilyParams([4, 2, 2, 2, 2, 2], "ccbbbb", [4.0, 4.0, 2.5, 2.5, 2.5, 2.5]),
                                  [100.0, 10.0, 1.0, 1.0, 1.0, 1.0]),
    "3d_o4_six_o3": HybridWENOParams(StencilFamilyParams([4, 3, 3, 3, 3, 3], "ccbbbb", [4.0, 4.0, 2.5, 2.5, 2.5, 2.5]),
                                     [100.0, 10.0, 1.0, 1.0, 1.0, 1.0]),
    # test/zisa/unit_test/reconstruction/weno_ao.cpp:47-62: lone stencils and a wider central stencil
    "2d_o1_c": HybridWENOParams(StencilFamilyParams([1], "c", [2.0]), [1.0]),
    "2d_o2_b": HybridWENOParams(StencilFamilyParams([2], "b", [2.0]), [1.0]),
    "2d_o3_c": HybridWENOParams(StencilFamilyParams([3], "c", [2.0]), [1.0]),
    "2d_o4_c": HybridWENOParams(StencilFamilyParams([4], "c", [2.0]), [1.0]),
    "2d_o3_wide": HybridWENOParams(StencilFamilyParams([3, 2, 2, 2], "cbbb", [3.0, 1.5, 1.5, 1.5]), [100.0, 1.0, 1.0, 1.0]),
    "2d_o4_w10": HybridWENOParams(StencilFamilyParams([4, 2, 2, 2], "cbbb", [2.0, 1.5, 1.5, 1.5]), [10.0, 1.0, 1.0, 1.0]),
}


class StencilFamilies:
    """Result of ``compute_stencil_families``: every cell's stencil family, fixed-stride arrays."""

    def __init__(self, grid: Grid, params: StencilFamilyParams, seed: int = 0):
        ns = len(params.orders)
        orders = (C.c_int * ns)(*[int(o) for o in params.orders])
        biases = "".join(params.biases).encode()
        factors = (C.c_double * ns)(*[float(f) for f in params.overfit_factors])
        h = C.c_void_p()
        check(lib.zfvm_stencils_compute(grid._h, ns, orders, biases, factors, seed, C.byref(h)))
        self._h = h
        self.grid = grid
        self.params = params
        self.n_stencils = ns

    @classmethod
    def extract(cls, src: "StencilFamilies", local_grid: Grid, local_to_src: np.ndarray) -> "StencilFamilies":
        """Stencils of a sub-grid cut out of ``src.grid``: local cell ``a`` is source cell ``local_to_src[a]``
        (what the reference's partitioner does with the global stencils, domain_decomposition.cpp:412-447)."""
        idx = np.ascontiguousarray(local_to_src, dtype=np.int32)
        assert idx.shape == (local_grid.n_cells,)
        h = C.c_void_p()
        check(lib.zfvm_stencils_extract(src._h, idx.size, idx.ctypes.data_as(_capi.c_int32_p), C.byref(h)))
        self = cls.__new__(cls)
        self._h = h
        self.grid = local_grid
        self.params = src.params
        self.n_stencils = src.n_stencils
        return self

    @classmethod
    def from_arrays(cls, grid: Grid, params: StencilFamilyParams, n_family, order, size, global_offset,
                    global_indices) -> "StencilFamilies":
        """Families the caller already holds (the reference's ``array<StencilFamily, 1>``,
        global_reconstruction_decl.hpp:107-147) instead of a new selection: ``order[i][k]`` / ``size[i][k]`` are
        ``Stencil::order()`` / ``size()``, ``global_indices[global_offset[i * ns + k]:][:size]`` is ``Stencil::global()``."""
        ns = len(params.orders)
        orders = (C.c_int * ns)(*[int(o) for o in params.orders])
        biases = "".join(params.biases).encode()
        factors = (C.c_double * ns)(*[float(f) for f in params.overfit_factors])
        nf = np.ascontiguousarray(n_family, dtype=np.int32)
        od = np.ascontiguousarray(order, dtype=np.int32).reshape(grid.n_cells, ns)
        sz = np.ascontiguousarray(size, dtype=np.int32).reshape(grid.n_cells, ns)
        go = np.ascontiguousarray(global_offset, dtype=np.int64).reshape(-1)
        gi = np.ascontiguousarray(global_indices, dtype=np.int32).reshape(-1)
        assert nf.shape == (grid.n_cells,) and go.size >= grid.n_cells * ns
        h = C.c_void_p()
        i32p, i64p = _capi.c_int32_p, _capi.c_int64_p
        check(lib.zfvm_stencils_from_arrays(grid._h, ns, orders, biases, factors, nf.ctypes.data_as(i32p),
                                            od.ctypes.data_as(i32p), sz.ctypes.data_as(i32p), go.ctypes.data_as(i64p),
                                            gi.ctypes.data_as(i32p), C.byref(h)))
        self = cls.__new__(cls)
        self._h = h
        self.grid = grid
        self.params = params
        self.n_stencils = ns
        return self

    def export_arrays(self):
        """(n_family, order, size, global_offset, global_indices): the inverse of :meth:`from_arrays`."""
        ns, n = self.n_stencils, self.grid.n_cells
        nf, order, size = self.array("n_family"), self.array("order"), self.array("size")
        off, l2g, local = self.array("local_off"), self.array("l2g"), self.array("local")
        used = np.where(np.arange(ns)[None, :] < nf[:, None], size, 0).astype(np.int64)
        go = np.zeros(n * ns + 1, dtype=np.int64)
        np.cumsum(used.reshape(-1), out=go[1:])
        gi = np.zeros(int(go[-1]), dtype=np.int32)
        for k in range(ns):
            for j in range(int(used[:, k].max(initial=0))):
                rows = np.nonzero(used[:, k] > j)[0]
                gi[go[rows * ns + k] + j] = l2g[rows, local[rows, off[k] + j]]
        return np.array(nf), np.array(order), np.array(size), go, gi

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            lib.zfvm_stencils_free(h)
            self._h = None

    def array(self, name: str) -> np.ndarray:
        return named_array(lib.zfvm_stencils_get, self._h, name, owner=self)

    def stencil(self, i: int, k: int) -> np.ndarray:
        """Global indices of stencil k of cell i (``Stencil::global()``, truncated to ``size()``)."""
        off = self.array("local_off")
        size = int(self.array("size")[i, k])
        loc = self.array("local")[i, off[k]: off[k] + size]
        return self.array("l2g")[i][loc]

    def matrix(self, i: int, k: int) -> np.ndarray:
        """``LSQSolver::A`` of stencil k of cell i."""
        buf = np.zeros(4096, dtype=np.float64)
        rows, cols = C.c_int(), C.c_int()
        check(lib.zfvm_stencil_matrix(self.grid._h, self._h, i, k, _capi.ptr_f64(buf), buf.size, C.byref(rows),
                                      C.byref(cols)))
        return buf[: rows.value * cols.value].reshape(rows.value, cols.value).copy()

    def matrix_layout(self):
        """(A_off[k], A_stride) of the padded all-cell matrix record."""
        ms = self.array("max_size")
        nd = self.grid.n_dims
        offs = [0]
        for k in range(self.n_stencils):
            deg = self.params.orders[k] - 1
            cols = ((deg + 1) * (deg + 2)) // 2 - 1 if nd == 2 else ((deg + 1) * (deg + 2) * (deg + 3)) // 6 - 1
            offs.append(offs[-1] + max(int(ms[k]) - 1, 1) * max(cols, 1))
        return np.asarray(offs[:-1], dtype=np.int64), int(offs[-1])

    def all_matrices(self) -> tuple[np.ndarray, np.ndarray, int]:
        a_off, stride = self.matrix_layout()
        A = np.zeros((self.grid.n_cells, stride), dtype=np.float64)
        check(lib.zfvm_stencil_matrices(self.grid._h, self._h, _capi.ptr_f64(A), stride,
                                        a_off.ctypes.data_as(_capi.c_int64_p)))
        return A, a_off, stride


def compute_stencil_families(grid: Grid, params: StencilFamilyParams, seed: int = 0) -> StencilFamilies:
    return StencilFamilies(grid, params, seed)


def hilbert_permutation(n_dims: int, centers: np.ndarray) -> np.ndarray:
    """Hilbert-curve order of cell centres ``[n][3]`` (src/renumber_grid.cpp:60-126): ``perm[new] = old``."""
    c = np.ascontiguousarray(centers, dtype=np.float64).reshape(-1, 3)
    perm = np.zeros(c.shape[0], dtype=np.int32)
    check(lib.zfvm_hilbert_permutation(n_dims, c.shape[0], c.ctypes.data_as(_capi.c_double_p),
                                       perm.ctypes.data_as(_capi.c_int32_p)))
    return perm
