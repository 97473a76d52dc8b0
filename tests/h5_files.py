"""Hand-assembled HDF5 files for the tests of the ``*.msh.h5`` reader (zisafvm_b200/csrc/host/msh_h5.cpp).

No HDF5 library exists in this image; these writers follow the HDF5 File Format Specification (version 3.0) byte by byte
and deliberately take the branches the library's own writer does not: the "latest" format (superblock 2, version-2 object
headers, compact Link messages, a compact scalar) and an old-format file behind a 512-byte user block with 4-byte signed
indices, a rank-1 vertex array and an object-header continuation block.  Checksums are written as zero (the reader does
not verify them).
"""
import struct

import numpy as np

SIG = b"\x89HDF\r\n\x1a\n"
UNDEF = 0xFFFFFFFFFFFFFFFF


def _dtype_msg(arr: np.ndarray) -> bytes:
    size = arr.dtype.itemsize
    if arr.dtype.kind in "iu":
        bits0 = 0x08 if arr.dtype.kind == "i" else 0x00
        return struct.pack("<BBBBIHH", 0x10, bits0, 0, 0, size, 0, 8 * size)
    if arr.dtype == np.float64:
        return struct.pack("<BBBBIHHBBBBI", 0x11, 0x20, 0x3F, 0, 8, 0, 64, 52, 11, 0, 52, 1023)
    if arr.dtype == np.float32:
        return struct.pack("<BBBBIHHBBBBI", 0x11, 0x20, 0x1F, 0, 4, 0, 32, 23, 8, 0, 23, 127)
    raise ValueError(arr.dtype)


def write_latest_format(path, datasets):
    """datasets: list of (name, ndarray, compact: bool).  Superblock 2, OHDR version 2, Link messages."""
    out = bytearray(48)  # superblock, patched at the end
    headers = {}

    def msg(t, body):
        return struct.pack("<BHB", t, len(body), 0) + body

    pending = []
    for name, arr, compact in datasets:
        arr = np.ascontiguousarray(arr)
        if arr.ndim == 0:
            space = struct.pack("<BBBB", 2, 0, 0, 0)
        else:
            space = struct.pack("<BBBB", 2, arr.ndim, 0, 1) + b"".join(struct.pack("<Q", d) for d in arr.shape)
        raw = arr.tobytes()
        if compact:
            layout = struct.pack("<BBH", 3, 0, len(raw)) + raw
        else:
            layout = struct.pack("<BBQQ", 3, 1, 0, len(raw))  # address patched below
        body = msg(0x01, space) + msg(0x03, _dtype_msg(arr)) + msg(0x08, layout)
        hdr_at = len(out)
        out += b"OHDR" + struct.pack("<BBB", 2, 0, len(body)) + body + b"\0\0\0\0"
        headers[name] = hdr_at
        if not compact:
            addr_at = hdr_at + 7 + len(body) - 16
            pending.append((addr_at, raw))
    for addr_at, raw in pending:
        while len(out) % 8:
            out += b"\0"
        out[addr_at:addr_at + 8] = struct.pack("<Q", len(out))
        out += raw
    links = b""
    for name in headers:
        nm = name.encode()
        links += msg(0x06, struct.pack("<BBB", 1, 0, len(nm)) + nm + struct.pack("<Q", headers[name]))
    root = len(out)
    out += b"OHDR" + struct.pack("<BBB", 2, 0, len(links)) + links + b"\0\0\0\0"
    out[0:48] = SIG + struct.pack("<BBBBQQQQI", 2, 8, 8, 0, 0, UNDEF, len(out), root, 0)
    with open(path, "wb") as f:
        f.write(bytes(out))


def write_old_format_with_user_block(path, n_dims, vertices, vertex_indices):
    """Superblock 1 at offset 512 (user block in front, addresses relative to the base address 512), 4-byte signed indices,
    vertices as one flat array, the data-layout message of `vertices` in a continuation block."""
    base = 512
    body = bytearray()

    def here():
        return len(body)  # relative address

    def pad8():
        while len(body) % 8:
            body.append(0)

    def msg1(t, data):
        padded = (len(data) + 7) // 8 * 8
        return struct.pack("<HHB3x", t, padded, 0) + data + b"\0" * (padded - len(data))

    body += b"\0" * 100  # superblock version 1 (100 bytes), patched at the end
    pad8()
    sets = [("n_dims", np.array(n_dims, dtype=np.int64)), ("vertex_indices", np.ascontiguousarray(vertex_indices, dtype=np.int32)),
            ("vertices", np.ascontiguousarray(vertices, dtype=np.float64).reshape(-1))]
    headers, data_patches = [], []
    for k, (name, arr) in enumerate(sets):
        pad8()
        space = struct.pack("<BBBB4x", 1, arr.ndim, 0, 0) + b"".join(struct.pack("<Q", d) for d in arr.shape)
        raw = arr.tobytes()
        layout = struct.pack("<BBQQ", 3, 1, 0, len(raw))
        if name == "vertices":  # layout message lives in a continuation block
            cont_block = msg1(0x08, layout)
            first = msg1(0x01, space) + msg1(0x03, _dtype_msg(arr)) + msg1(0x10, struct.pack("<QQ", 0, len(cont_block)))
            hdr = here()
            body += struct.pack("<BBHII4x", 1, 0, 4, 1, len(first)) + first
            cont_at = here()
            body[cont_at - 16:cont_at - 8] = struct.pack("<Q", cont_at)  # continuation offset
            body += cont_block
            data_patches.append((cont_at + 8 + 2, raw))
        else:
            first = msg1(0x01, space) + msg1(0x03, _dtype_msg(arr)) + msg1(0x08, layout)
            hdr = here()
            body += struct.pack("<BBHII4x", 1, 0, 3, 1, len(first)) + first
            data_patches.append((here() - 24 + 2, raw))
        headers.append(hdr)
    for at, raw in data_patches:
        pad8()
        body[at:at + 8] = struct.pack("<Q", here())
        body += raw
    # local heap, symbol table node, B-tree, root header
    pad8()
    heap_data = bytearray(8)
    offs = []
    for name, _ in sets:
        offs.append(len(heap_data))
        heap_data += name.encode() + b"\0"
        while len(heap_data) % 8:
            heap_data.append(0)
    free = len(heap_data)
    heap_data += struct.pack("<QQ", 1, 16)
    heap = here()
    body += b"HEAP" + struct.pack("<B3xQQQ", 0, len(heap_data), free, heap + 32) + heap_data
    pad8()
    snod = here()
    body += b"SNOD" + struct.pack("<BBH", 1, 0, 3)
    for off, hdr in zip(offs, headers):
        body += struct.pack("<QQII16x", off, hdr, 0, 0)
    body += b"\0" * (8 + 8 * 40 - (here() - snod))
    btree = here()
    body += b"TREE" + struct.pack("<BBHQQ", 0, 0, 1, UNDEF, UNDEF) + struct.pack("<QQQ", 0, snod, offs[-1])
    body += b"\0" * (24 + 33 * 8 + 32 * 8 - (here() - btree))
    root = here()
    stab = msg1(0x11, struct.pack("<QQ", btree, heap))
    body += struct.pack("<BBHII4x", 1, 0, 1, 1, len(stab)) + stab
    sb = SIG + struct.pack("<BBBBBBBBHHI", 1, 0, 0, 0, 0, 8, 8, 0, 4, 16, 0) + struct.pack("<HH", 32, 0)
    sb += struct.pack("<QQQQ", base, UNDEF, base + len(body), UNDEF)
    sb += struct.pack("<QQII", 0, root, 1, 0) + struct.pack("<QQ", btree, heap)
    assert len(sb) == 100
    body[0:100] = sb
    with open(path, "wb") as f:
        f.write(b"user block".ljust(base, b"\0") + bytes(body))
