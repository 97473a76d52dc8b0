"""Host precompute of the product library (grid flattening, quadrature, stencils, LSQ matrices) against the
reference's unit tests that need no grid files, and against its property tests re-run on synthetic grids.
No GPU is needed: these entry points of libzfvm_b200.so are pure host code.

Reference tests restated (under /root/reference/test/zisa/unit_test):
  grid/grid.cpp:21-108            two-triangle grid: neighbours, edge_indices, normal, volume
  grid/grid.cpp:195-213           normals point from the left to the right cell
  reconstruction/stencil.cpp:126-145   deduce_max_order table
  math/gauss_legendre.cpp:10-58   4- and 5-point closed forms (2e-14)
  math/edge_rule.cpp:12-37        degree -> number of points, symmetry, unit total weight
  math/face.cpp:12-42             face barycentre from the face rule
  reconstruction/hybrid_weno.hpp:361-412, lsq_solver.cpp:20-52   every LSQ matrix has full column rank
  reconstruction/hybrid_weno.hpp:53-102   interior cells reach the requested stencil orders
"""
import ctypes as C
import math

import numpy as np
import pytest

import zisafvm_b200 as z
from zisafvm_b200 import _capi, cases
from zisafvm_b200._capi import lib

INVALID = -1


def two_triangles():
    v = np.array([[0, 0, 0], [1, 0, 0], [1, 1, 0], [0, 1, 0]], dtype=float)
    vi = np.array([[0, 1, 3], [1, 2, 3]], dtype=np.int32)
    return z.Grid(2, v, vi, z.QRDegrees(1, 1, 1))


def test_two_triangles_connectivity():
    g = two_triangles()
    nb = g.array("neighbours")
    assert nb.tolist() == [[INVALID, 1, INVALID], [INVALID, INVALID, 0]]
    ei = g.array("edge_indices")
    assert ei[0, 1] == 0 and ei[1, 2] == 0                       # the only interior (diagonal) edge comes first
    assert (ei[0, 0], ei[0, 2], ei[1, 0], ei[1, 1]) == (1, 2, 3, 4)
    assert g.n_edges == 5 and g.n_interior_edges == 1
    n0 = g.array("face_normal")[0]
    assert np.abs(n0 - [0.5 / math.sqrt(0.5), 0.5 / math.sqrt(0.5), 0.0]).max() < 1e-12
    assert abs(g.array("volumes")[0] - 0.5) < 1e-12
    assert g.array("left_right")[0].tolist() == [0, 1]           # left = smaller index (grid.cpp:415-442)


def test_deduce_max_order_table():
    table = {1: 1, 2: 1, 4: 1, 5: 2, 10: 2, 11: 3, 18: 3, 19: 4, 28: 4, 29: 5}
    for n, order in table.items():
        assert lib.zfvm_deduce_max_order(n, 2.0, 2) == order


def _gl(n):
    x, w = np.zeros(n), np.zeros(n)
    lib.zfvm_gauss_legendre(n, _capi.ptr_f64(x), _capi.ptr_f64(w))
    return x, w


def test_gauss_legendre_closed_forms():
    s = math.sqrt
    x5 = [-s(5 + 2 * s(10 / 7)) / 3, -s(5 - 2 * s(10 / 7)) / 3, 0.0, s(5 - 2 * s(10 / 7)) / 3, s(5 + 2 * s(10 / 7)) / 3]
    w5 = [(322 - 13 * s(70)) / 900, (322 + 13 * s(70)) / 900, 128 / 225, (322 + 13 * s(70)) / 900, (322 - 13 * s(70)) / 900]
    x, w = _gl(5)
    assert np.abs(x - x5).max() < 2e-14 and np.abs(w - w5).max() < 2e-14
    x4 = [-s(3 / 7 + 2 / 7 * s(6 / 5)), -s(3 / 7 - 2 / 7 * s(6 / 5)), s(3 / 7 - 2 / 7 * s(6 / 5)), s(3 / 7 + 2 / 7 * s(6 / 5))]
    w4 = [(18 - s(30)) / 36, (18 + s(30)) / 36, (18 + s(30)) / 36, (18 - s(30)) / 36]
    x, w = _gl(4)
    assert np.abs(x - x4).max() < 2e-14 and np.abs(w - w4).max() < 2e-14


def _rule(kind, deg):
    n, nb = C.c_int(), C.c_int()
    w, b = np.zeros(64), np.zeros(256)
    _capi.check(lib.zfvm_quadrature_rule(kind, deg, C.byref(n), C.byref(nb), _capi.ptr_f64(w), _capi.ptr_f64(b), 64))
    return w[: n.value].copy(), b[: n.value * nb.value].reshape(n.value, nb.value).copy()


def test_edge_rule_basics():
    for deg, n_points in [(0, 1), (1, 1), (2, 2), (3, 2), (4, 3), (5, 3), (6, 4), (7, 4)]:
        w, b = _rule(1, deg)
        assert w.size == n_points
        assert abs(w.sum() - 1.0) < 1e-12
        xi = b[:, 1] - b[:, 0]                                    # bary = (0.5 - 0.5 xi, 0.5 + 0.5 xi)
        assert np.abs(xi + xi[::-1]).max() < 1e-12 and np.abs(w - w[::-1]).max() < 1e-12


@pytest.mark.parametrize("kind,max_deg,nv", [(2, 5, 3), (3, 3, 4)])
def test_simplex_rules_integrate_monomials_exactly(kind, max_deg, nv):
    """TriangularRule deg 1-5 (triangular_rule.cpp:33-91) and TetrahedralRule deg 1-3 (tetrahedral_rule.cpp:10-122):
    barycentric monomials int l0^a l1^b l2^c (l3^d) = a! b! c! (d!) dim! / (a+b+c(+d)+dim)!  (weights sum to 1)."""
    from itertools import product

    dim = nv - 1
    for deg in range(1, max_deg + 1):
        w, b = _rule(kind, deg)
        assert b.shape[1] == nv and abs(w.sum() - 1.0) < 1e-14 and np.abs(b.sum(axis=1) - 1.0).max() < 1e-14
        for e in product(range(deg + 1), repeat=nv):
            if sum(e) > deg:
                continue
            exact = math.factorial(dim) * math.prod(math.factorial(a) for a in e) / math.factorial(sum(e) + dim)
            approx = (w * np.prod(b ** np.array(e), axis=1)).sum()
            assert abs(approx - exact) < 1e-14, (kind, deg, e)


def test_rule_point_counts():
    assert [_rule(2, d)[0].size for d in (1, 2, 3, 4, 5)] == [1, 3, 4, 6, 7]     # Dunavant
    assert [_rule(3, d)[0].size for d in (1, 2, 3)] == [1, 4, 10]                 # Shunn-Ham


def test_face_barycentre():
    """A one-cell grid: the face rule's weighted mean of the face points is the vertex average (face.cpp:12-42)."""
    v = np.array([[1.2, 3.0, 0.0], [4.2, 3.2, 4.0], [2.2, 3.2, 4.0], [1.0, 0.0, 1.0]])
    g = z.Grid(3, v, np.array([[0, 1, 2, 3]], dtype=np.int32), z.QRDegrees(3, 2, 2))
    qp, qw, area = g.array("face_qp"), g.array("face_qw"), g.array("face_area")
    for e in range(4):
        bary = (qw[e, :, None] * qp[e]).sum(axis=0) / area[e]
        # the face barycentre must coincide with the mean of three of the four vertices
        cands = [v[[a, b_, c]].mean(axis=0) for a in range(4) for b_ in range(a + 1, 4) for c in range(b_ + 1, 4)]
        assert min(np.abs(bary - c).max() for c in cands) < 1e-10
        assert abs(qw[e].sum() - area[e]) < 1e-13


@pytest.fixture(scope="module")
def square():
    verts, vi = z.square_mesh(12, 12, 0.0, 1.0, 0.0, 1.0, jitter=0.15, seed=1)
    return z.Grid(2, verts, vi, z.QRDegrees(3, 3, 4))


@pytest.fixture(scope="module")
def cube():
    verts, vi = z.cube_mesh(5, 5, 5, 0.2, jitter=0.1, seed=1)
    return z.Grid(3, verts, vi, z.QRDegrees(3, 2, 3))


@pytest.mark.parametrize("which", ["square", "cube"])
def test_grid_geometry_properties(which, request):
    g = request.getfixturevalue(which)
    F = g.max_neighbours
    vol, cc = g.array("volumes"), g.array("cell_centers")
    assert abs(vol.sum() - 1.0) < 1e-12                            # check_volume, grid.cpp:215-225
    lr, n = g.array("left_right"), g.array("face_normal")
    EI = g.n_interior_edges
    assert (lr[:EI, 1] != INVALID).all() and (lr[EI:, 1] == INVALID).all()  # interior faces first (grid.cpp:385-413)
    assert (lr[:EI, 0] < lr[:EI, 1]).all()
    d = cc[lr[:EI, 1]] - cc[lr[:EI, 0]]
    assert ((n[:EI] * d).sum(axis=1) > 0).all()                    # "orientation normals"
    t1, t2 = g.array("face_t1"), g.array("face_t2")
    assert np.abs((n * t1).sum(axis=1)).max() < 1e-14 and np.abs(np.cross(n, t1) - t2).max() < 1e-14
    assert np.abs(np.linalg.norm(n, axis=1) - 1).max() < 1e-14
    # divergence theorem per cell: sum of outward area-weighted normals vanishes
    ei, area = g.array("edge_indices"), g.array("face_area")
    acc = np.zeros((g.n_cells, 3))
    for k in range(F):
        e = ei[:, k]
        sign = np.where(lr[e, 0] == np.arange(g.n_cells), 1.0, -1.0)
        acc += sign[:, None] * area[e, None] * n[e]
    assert np.abs(acc).max() < 1e-13
    # quadrature points lie inside their cell / weights sum to the volume
    assert np.abs(g.array("cell_qw").sum(axis=1) - vol).max() < 1e-15
    # normalized moments (grid.cpp:1049-1098): degree-1 moments vanish, all bounded by 2 (grid.cpp:183-193)
    m = g.array("moments")
    assert np.abs(m[:, 0] - 1.0).max() < 1e-13 and np.abs(m[:, 1:F]).max() < 1e-13 and np.abs(m).max() < 2.0
    # moments are the cell averages of ((x - x_c)/l)^alpha: check xi^2 with the cell rule
    L = g.array("characteristic_length")
    qp, qw = g.array("cell_qp"), g.array("cell_qw")
    xi = (qp[:, :, 0] - cc[:, None, 0]) / L[:, None]
    m_xx = (qw * xi * xi).sum(axis=1) / vol
    idx = 3 if g.n_dims == 2 else 4                                 # poly_index(2,0) / poly_index(2,0,0)
    assert np.abs(m[:, idx] - m_xx).max() < 1e-13


def test_mask_ghost_cells_and_frozen_rows(square):
    """mask_ghost_cells (grid.cpp:1122-1136): ghost bit set, interior bit cleared; l1 = ghost with an interior neighbour."""
    g = square
    cc = g.array("cell_centers")
    mask = (cc[:, 0] < 0.2) | (cc[:, 0] > 0.8)
    g.mask_ghost_cells(mask)
    fl = g.array("cell_flags")
    assert (((fl & 2) != 0) == mask).all() and (((fl & 1) != 0) == ~mask).all()
    nb = g.array("neighbours")
    has_int_nb = np.zeros(g.n_cells, dtype=bool)
    for k in range(3):
        ok = nb[:, k] != INVALID
        has_int_nb[ok] |= ~mask[nb[ok, k]]
    assert (((fl & 4) != 0) == (mask & has_int_nb)).all()
    g.mask_ghost_cells(np.zeros(g.n_cells, dtype=bool))


def test_hilbert_order_is_a_local_permutation():
    a, ia = z.square_mesh(16, 16, hilbert=False, seed=2)
    b, ib = z.square_mesh(16, 16, hilbert=True, seed=2)
    assert np.array_equal(a, b)
    key = lambda t: sorted(map(tuple, np.sort(t, axis=1).tolist()))
    assert key(ia) == key(ib)                                      # same cells, renumbered
    cb = b[ib].mean(axis=1)
    ca = a[ia].mean(axis=1)
    # consecutive cells are close along the curve: mean jump well below the row-major ordering's
    jump = lambda c: np.linalg.norm(np.diff(c, axis=0), axis=1)
    assert jump(cb).max() < 0.2 and jump(cb).mean() < 0.07


PARAM_SETS = [("square", "2d_o3"), ("square", "2d_o4"), ("square", "2d_o5"), ("cube", "3d_o2"), ("cube", "3d_o3"), ("cube", "3d_o4"),
              # the reference's own test families: six stencils with two central ones (cweno_ao.cpp:144-160,
              # test_hybrid_weno_matrices :172-180), lone stencils / first order / wide central (weno_ao.cpp:47-62)
              ("cube", "3d_o4_six"), ("cube", "3d_o4_six_o3"), ("square", "2d_o1_c"), ("square", "2d_o2_b"),
              ("square", "2d_o3_c"), ("square", "2d_o4_c"), ("square", "2d_o3_wide")]


@pytest.mark.parametrize("which,key", PARAM_SETS)
def test_stencil_families(which, key, request):
    g = request.getfixturevalue(which)
    prm = z.WENO_PARAMS[key].stencil_family_params
    if key.endswith("o5") and g.n_moments < 15:
        pytest.skip("grid built with moments_deg 4")
    if key.startswith("3d_o4"):
        verts, vi = z.cube_mesh(7, 7, 7, 1.0 / 7, jitter=0.1, seed=1)
        g = z.Grid(3, verts, vi, z.QRDegrees(3, 3, 3))
    st = z.compute_stencil_families(g, prm)
    nd, ns = g.n_dims, len(prm.orders)
    l2g, size, order = st.array("l2g"), st.array("size"), st.array("order")
    assert (l2g[:, 0] == np.arange(g.n_cells)).all()               # stencil.cpp:82-104
    dof = lambda deg: (deg + 1) * (deg + 2) // 2 if nd == 2 else (deg + 1) * (deg + 2) * (deg + 3) // 6
    req = [int((dof(o - 1) - 1) * f + 1) for o, f in zip(prm.orders, prm.overfit_factors)]  # stencil.cpp:168-175
    assert st.array("max_size").tolist() == req
    n_family = st.array("n_family")
    full = 0
    for i in range(g.n_cells):
        for k in range(n_family[i]):
            s = st.stencil(i, k)
            assert s[0] == i and len(set(s.tolist())) == len(s)   # the cell itself first, no duplicates
            o = order[i, k]
            assert lib.zfvm_deduce_max_order(len(s), prm.overfit_factors[k], nd) >= o
            if o > 1:
                A = st.matrix(i, k)
                assert A.shape == (len(s) - 1, dof(o - 1) - 1)
                assert np.linalg.matrix_rank(A) == A.shape[1]       # hybrid_weno.hpp:361-412
        full += int(n_family[i] == ns and (order[i] == np.array(prm.orders)).all())
    # cells away from the boundary reach the requested orders (test_hybrid_weno_valid_stencil)
    assert full > (0.2 if not key.startswith("3d_o4_six") else 0.02) * g.n_cells
    # biased stencil k lies in the half space behind face k-1... at least it must differ from the central one
    k_high = st.array("k_high")
    assert ((k_high >= 0) & (k_high < ns)).all()


@pytest.mark.parametrize("which,key", [("square", "2d_o3"), ("cube", "3d_o3"), ("square", "2d_o1_c"), ("cube", "3d_o4_six_o3")])
def test_stencil_import_round_trip(which, key, request):
    """zfvm_stencils_from_arrays: families handed over as the reference holds them (Stencil::global(), order(), size()
    per stencil; global_reconstruction_decl.hpp:107-147) give the same tables as the selection that produced them, and
    the LSQ matrices built from them are bit-identical."""
    g = request.getfixturevalue(which)
    prm = z.WENO_PARAMS[key].stencil_family_params
    st = z.compute_stencil_families(g, prm)
    nf, order, size, go, gi = st.export_arrays()
    imp = z.StencilFamilies.from_arrays(g, prm, nf, order, size, go, gi)
    for name in ("l2g", "l2g_size", "local", "order", "size", "k_high", "n_family", "family_order", "max_size", "local_off"):
        a, b = st.array(name), imp.array(name)
        if name == "local":     # slots behind a stencil's size are never read; compare the used part
            off = st.array("local_off")
            for k in range(len(prm.orders)):
                used = np.arange(off[k + 1] - off[k])[None, :] < np.where(np.arange(len(prm.orders))[None, :] < nf[:, None], size, 0)[:, k:k + 1]
                assert (np.where(used, a[:, off[k]:off[k + 1]], 0) == np.where(used, b[:, off[k]:off[k + 1]], 0)).all(), name
        else:
            assert (a == b).all(), name
    for i in (0, g.n_cells // 3, g.n_cells - 1):
        for k in range(int(nf[i])):
            if order[i, k] > 1:
                assert (st.matrix(i, k) == imp.matrix(i, k)).all()
    # what the import refuses: a stencil that does not start with its own cell, a size the order does not need
    i_full = int(np.nonzero(nf == len(prm.orders))[0][0]) if (nf == len(prm.orders)).any() and len(prm.orders) > 1 else None
    if i_full is not None:
        bad = gi.copy()
        bad[go[i_full * len(prm.orders)]] = (i_full + 1) % g.n_cells
        with pytest.raises(_capi.ZfvmError, match="member 0"):
            z.StencilFamilies.from_arrays(g, prm, nf, order, size, go, bad)
        bad_size = size.copy()
        bad_size[i_full, 0] -= 1
        with pytest.raises(_capi.ZfvmError, match="required_stencil_size"):
            z.StencilFamilies.from_arrays(g, prm, nf, order, bad_size, go, gi)


def _monomial_exponents(nd, deg):
    """poly_index order (poly2d_impl.hpp:34-41): by total degree, then 2D (a, b) -> b ascending; 3D (a, b, c) ->
    idx2(b, c) ascending."""
    out = []
    for n in range(deg + 1):
        if nd == 2:
            out += [(n - b, b, 0) for b in range(n + 1)]
        else:
            for m in range(n + 1):            # m = b + c
                out += [(n - m, m - c, c) for c in range(m + 1)]
    return out


@pytest.mark.parametrize("nd,key,qdeg", [(2, "2d_o3", 3), (2, "2d_o4", 4), (2, "2d_o5", 5), (3, "3d_o2", 2), (3, "3d_o3", 3),
                                         (3, "3d_o4", 3), (3, "3d_o4_six_o3", 3)])
def test_lsq_matrix_reproduces_polynomial_averages(nd, key, qdeg):
    """Independent pin of the precompute the oracle shares with the product (moments and LSQ matrices): row j of A
    holds the cell averages over stencil cell j of the centre cell's zero-mean scaled monomials
    (lsq_solver.cpp:168-403, exact binomial moment formulas), and the normalised moments are the centre cell's own
    monomial averages (grid.cpp:1049-1098).  Both are re-derived here in numpy by direct quadrature with a rule that
    is exact for the degree -- nothing of the host library's assembly code is on this side."""
    if nd == 2:
        verts, vi = z.square_mesh(9, 9, jitter=0.15, seed=2)
        g = z.Grid(2, verts, vi, z.QRDegrees(3, qdeg, qdeg))
    else:
        verts, vi = z.cube_mesh(6, 6, 6, 1.0 / 6, jitter=0.1, seed=2)
        g = z.Grid(3, verts, vi, z.QRDegrees(3, qdeg, qdeg))
    prm = z.WENO_PARAMS[key].stencil_family_params
    st = z.compute_stencil_families(g, prm)
    cc, L, m = g.array("cell_centers"), g.array("characteristic_length"), g.array("moments")
    qp, qw, vol = g.array("cell_qp"), g.array("cell_qw"), g.array("volumes")
    order, n_family = st.array("order"), st.array("n_family")
    expo = _monomial_exponents(nd, max(prm.orders) - 1)

    def averages(i, j, n_mono):  # average over cell j of the monomials of cell i's scaled basis
        xi = (qp[j] - cc[i]) / L[i]
        return np.array([(qw[j] * xi[:, 0] ** a * xi[:, 1] ** b * xi[:, 2] ** c).sum() / vol[j] for a, b, c in expo[:n_mono]])

    checked = 0
    for i in range(0, g.n_cells, 13):
        mom_i = averages(i, i, g.n_moments if g.n_moments <= len(expo) else len(expo))
        ref_m = m[i, : mom_i.size].copy()
        assert abs(ref_m[0] - 1.0) < 1e-13 or ref_m[0] == 0.0
        # c_0 .. c_nd are forced to zero (the constant, and the centre is the quadrature barycentre), poly2d_impl.hpp:95
        assert np.abs(mom_i[1 + nd:] - ref_m[1 + nd:]).max(initial=0.0) < 1e-13
        assert np.abs(mom_i[1: 1 + nd]).max() < 1e-13
        for k in range(n_family[i]):
            o = order[i, k]
            if o <= 1 or o - 1 > qdeg:
                continue
            s, A = st.stencil(i, k), st.matrix(i, k)
            n_mono = A.shape[1] + 1
            c_i = np.zeros(n_mono)
            c_i[1 + nd:] = m[i, 1 + nd: n_mono]
            for r, j in enumerate(s[1:]):
                row = averages(i, j, n_mono) - c_i
                assert np.abs(A[r] - row[1:]).max() < 1e-12 * max(1.0, np.abs(row).max()), (i, k, r)
            checked += 1
    assert checked > 10


def test_pseudo_inverse_matches_numpy():
    rng = np.random.default_rng(0)
    for rows, cols in [(3, 2), (10, 5), (18, 9), (57, 19)]:
        A = rng.normal(size=(rows, cols))
        W = np.zeros((cols, rows))
        _capi.check(lib.zfvm_pseudo_inverse(_capi.ptr_f64(np.ascontiguousarray(A)), rows, cols, _capi.ptr_f64(W)))
        assert np.abs(W - np.linalg.pinv(A)).max() < 1e-12 * np.abs(W).max() * np.linalg.cond(A)


def test_error_reporting():
    """Failures come back as a non-zero status plus zfvm_last_error() (the reference would LOG_ERR)."""
    with pytest.raises(_capi.ZfvmError, match="n_dims"):
        z.Grid(4, np.zeros((5, 3)), np.zeros((1, 5), dtype=np.int32), z.QRDegrees())
    with pytest.raises(_capi.ZfvmError, match="out of range"):
        z.Grid(2, np.zeros((3, 3)), np.array([[0, 1, 7]], dtype=np.int32), z.QRDegrees())
    g = two_triangles()
    with pytest.raises(_capi.ZfvmError, match="bias"):
        z.compute_stencil_families(g, z.StencilFamilyParams([1], "x", [1.0]))


def test_gravity_tables_match_numpy_restatement():
    """Host tabulation (csrc/host/gravity.cpp) vs the independent numpy restatement in cases.gravity_tables
    (gravity_impl.hpp:13-58, RadialAlignment gravity_decl.hpp:76-95)."""
    case = cases.polytrope_2d(n=10)
    a, b, c = cases.gravity_tables(case.grid, case.params.gravity)
    assert np.isfinite(a).all() and np.isfinite(b).all() and np.isfinite(c).all()
    r = np.linalg.norm(case.grid.array("cell_qp"), axis=2)
    alpha = math.sqrt(2 * math.pi)
    assert np.abs(a + 2.0 * np.sin(alpha * r) / (alpha * r)).max() < 1e-13
