"""The reference's grid file ``*.msh.h5`` (load_grid_gmsh_h5, src/zisa/grid/grid.cpp:889-901) through the library's own
HDF5-subset reader / writer (csrc/host/msh_h5.cpp).  The image has no HDF5 library: the reader is pinned against the
library's writer and against files assembled byte by byte from the format specification (tests/h5_files.py), which take
the other branches of the format (latest-format headers and links, user block, continuation block, narrow integers)."""
import struct

import numpy as np
import pytest

import zisafvm_b200 as z
from zisafvm_b200 import _capi, grid as G

import h5_files


def _meshes():
    v3, c3 = G.cube_mesh(3, 2, 4, 0.25, jitter=0.1, seed=3)
    v2, c2 = G.square_mesh(5, 4, jitter=0.15, seed=1)
    return [(3, v3, c3), (2, v2, c2)]


@pytest.mark.parametrize("case", [0, 1])
def test_write_read_round_trip(tmp_path, case):
    nd, v, c = _meshes()[case]
    path = tmp_path / "grid.msh.h5"
    G.write_msh_h5(path, nd, v, c)
    nd2, v2, c2 = G.read_msh_h5(path)
    assert nd2 == nd and np.array_equal(v, v2) and np.array_equal(c, c2)
    # the grid built from the file is the grid built from the arrays
    qr = G.QRDegrees(2, 2, 2)
    g0, g1 = G.Grid(nd, v, c, qr), G.Grid(nd2, v2, c2, qr)
    for name in ("neighbours", "edge_indices", "volumes", "cell_centers", "moments"):
        assert np.array_equal(g0.array(name), g1.array(name))


def test_written_file_has_the_layout_libhdf5_writes_by_default(tmp_path):
    nd, v, c = _meshes()[0]
    path = tmp_path / "grid.msh.h5"
    G.write_msh_h5(path, nd, v, c)
    b = path.read_bytes()
    assert b[:8] == b"\x89HDF\r\n\x1a\n" and b[8] == 0            # superblock version 0
    assert b[13] == 8 and b[14] == 8                                # 8-byte offsets and lengths
    assert struct.unpack_from("<HH", b, 16) == (4, 16)              # group leaf / internal node K
    assert struct.unpack_from("<Q", b, 40)[0] == len(b)             # end-of-file address
    root_header, cache_type = struct.unpack_from("<Q", b, 64)[0], struct.unpack_from("<I", b, 72)[0]
    btree, heap = struct.unpack_from("<QQ", b, 80)
    assert cache_type == 1 and b[btree:btree + 4] == b"TREE" and b[heap:heap + 4] == b"HEAP"
    assert b[root_header] == 1 and struct.unpack_from("<H", b, root_header + 16)[0] == 0x11   # symbol table message
    snod = struct.unpack_from("<Q", b, btree + 24 + 8)[0]
    assert b[snod:snod + 4] == b"SNOD" and struct.unpack_from("<H", b, snod + 6)[0] == 3
    heap_data = struct.unpack_from("<Q", b, heap + 24)[0]
    names = []
    for e in range(3):
        off, hdr = struct.unpack_from("<QQ", b, snod + 8 + 40 * e)
        end = b.index(b"\0", heap_data + off)
        names.append(b[heap_data + off:end].decode())
        assert b[hdr] == 1 and hdr % 8 == 0                         # version-1 object header, aligned
    assert names == sorted(names) == ["n_dims", "vertex_indices", "vertices"]
    # vertex_indices are the reference's int_t: 64-bit unsigned, contiguous, bit for bit in the file
    assert np.asarray(c, dtype=np.uint64).tobytes() in b and np.asarray(v, dtype=np.float64).tobytes() in b


def test_reads_latest_format_files(tmp_path):
    nd, v, c = _meshes()[1]
    path = tmp_path / "latest.msh.h5"
    h5_files.write_latest_format(path, [("vertices", v.astype(np.float32), False), ("n_dims", np.array(nd, dtype=np.int8), True),
                                        ("vertex_indices", c.astype(np.uint16), False)])
    nd2, v2, c2 = G.read_msh_h5(path)
    assert nd2 == nd and np.array_equal(c, c2) and np.array_equal(v.astype(np.float32).astype(np.float64), v2)


def test_reads_old_format_behind_a_user_block(tmp_path):
    nd, v, c = _meshes()[0]
    path = tmp_path / "user_block.msh.h5"
    h5_files.write_old_format_with_user_block(path, nd, v, c)
    nd2, v2, c2 = G.read_msh_h5(path)
    assert nd2 == nd and np.array_equal(v, v2) and np.array_equal(c, c2)


def test_rejects_what_it_does_not_implement(tmp_path):
    nd, v, c = _meshes()[1]
    lib = _capi.lib
    out = (G.C.c_int(), G.C.c_int64(), _capi.c_double_p(), G.C.c_int64(), _capi.c_int32_p())

    def read(path):
        return lib.zfvm_mesh_read_msh_h5(str(path).encode(), G.C.byref(out[0]), G.C.byref(out[1]), G.C.byref(out[2]),
                                         G.C.byref(out[3]), G.C.byref(out[4]))

    p = tmp_path / "not_hdf5.msh.h5"
    p.write_bytes(b"$MeshFormat\n" * 100)
    assert read(p) != 0 and "not an HDF5 file" in _capi.lib.zfvm_last_error().decode()
    assert read(tmp_path / "missing.msh.h5") != 0 and "cannot open" in _capi.lib.zfvm_last_error().decode()
    p = tmp_path / "no_vertices.msh.h5"
    h5_files.write_latest_format(p, [("n_dims", np.array(nd, dtype=np.int32), True), ("vertex_indices", c, False)])
    assert read(p) != 0 and "no dataset 'vertices'" in _capi.lib.zfvm_last_error().decode()
    p = tmp_path / "bad_index.msh.h5"
    bad = c.copy()
    bad[0, 0] = v.shape[0]
    h5_files.write_latest_format(p, [("n_dims", np.array(nd, dtype=np.int32), True), ("vertex_indices", bad, False),
                                     ("vertices", v, False)])
    assert read(p) != 0 and "out of range" in _capi.lib.zfvm_last_error().decode()
    # a chunked dataset: data layout class 2
    p = tmp_path / "chunked.msh.h5"
    G.write_msh_h5(p, nd, v, c)
    b = bytearray(p.read_bytes())
    at = b.index(struct.pack("<BBQ", 3, 1, b.index(np.asarray(v).tobytes())))
    b[at + 1] = 2
    p.write_bytes(bytes(b))
    assert read(p) != 0 and "chunked" in _capi.lib.zfvm_last_error().decode()
    # truncated file
    p = tmp_path / "truncated.msh.h5"
    G.write_msh_h5(p, nd, v, c)
    p.write_bytes(p.read_bytes()[:-64])
    assert read(p) != 0 and "beyond the end" in _capi.lib.zfvm_last_error().decode()


def test_subgrid_file_round_trip(tmp_path):
    """subgrid-%04d.msh.h5 (src/domain_decomposition.cpp:80-88): the three datasets of a grid file plus `partition` and
    `global_cell_indices` as 64-bit unsigned integers; written by the library, read back by the library and -- dataset by
    dataset -- by the independent pure-Python walker of tests/h5_files.py when it is available."""
    from zisafvm_b200._capi import ZfvmError
    from zisafvm_b200.grid import cube_mesh, read_msh_h5, read_subgrid_h5, write_msh_h5, write_subgrid_h5

    verts, vi = cube_mesh(3, 2, 2, 0.5, jitter=0.1, seed=3)
    nc = vi.shape[0]
    rng = np.random.default_rng(0)
    part = np.sort(rng.integers(0, 3, size=nc)).astype(np.int64)
    gci = rng.permutation(10 * nc)[:nc].astype(np.int64)
    path = tmp_path / "subgrid-0001.msh.h5"
    write_subgrid_h5(str(path), 3, verts, vi, part, gci)
    nd, v, c, p, g = read_subgrid_h5(str(path))
    assert nd == 3 and np.array_equal(v, verts) and np.array_equal(c, vi)
    assert np.array_equal(p, part) and np.array_equal(g, gci)
    # a sub-grid file is a grid file too (load_grid reads the same file, local_grid.cpp:17-18)
    nd2, v2, c2 = read_msh_h5(str(path))
    assert nd2 == 3 and np.array_equal(v2, verts) and np.array_equal(c2, vi)
    # a plain grid file is not a sub-grid file
    plain = tmp_path / "grid.msh.h5"
    write_msh_h5(str(plain), 3, verts, vi)
    with pytest.raises(ZfvmError, match="partition"):
        read_subgrid_h5(str(plain))
    with pytest.raises(ZfvmError, match="unsigned"):
        write_subgrid_h5(str(tmp_path / "bad.h5"), 3, verts, vi, part - 1, gci)
    with pytest.raises(ValueError):
        write_subgrid_h5(str(tmp_path / "bad.h5"), 3, verts, vi, part[:-1], gci)
    # the same five datasets assembled byte by byte in the other format branch (superblock 2, version-2 headers, link
    # messages; 32-bit owner ranks, 64-bit unsigned global indices): an independent writer for the reader
    other = tmp_path / "subgrid-latest.msh.h5"
    h5_files.write_latest_format(other, [("n_dims", np.array(3, dtype=np.int32), True), ("vertex_indices", vi.astype(np.uint64), False),
                                         ("vertices", verts, False), ("partition", part.astype(np.uint32), False),
                                         ("global_cell_indices", gci.astype(np.uint64), False)])
    nd3, v3, c3, p3, g3 = read_subgrid_h5(str(other))
    assert nd3 == 3 and np.array_equal(v3, verts) and np.array_equal(c3, vi) and np.array_equal(p3, part) and np.array_equal(g3, gci)
    # one of the two datasets alone is not a sub-grid file
    half = tmp_path / "half.msh.h5"
    h5_files.write_latest_format(half, [("n_dims", np.array(3, dtype=np.int32), True), ("vertex_indices", vi.astype(np.uint64), False),
                                        ("vertices", verts, False), ("partition", part.astype(np.uint64), False)])
    with pytest.raises(ZfvmError, match="come together"):
        read_subgrid_h5(str(half))
