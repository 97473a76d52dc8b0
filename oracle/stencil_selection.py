"""TEST INFRASTRUCTURE ONLY -- pure-Python restatement of the reference's stencil selection, small grids only.

The C++ oracle (zisa_oracle.cpp) receives its stencils from the product's host library; this file restates the selection
itself so that the host library's stencils are pinned by something that shares no code with it
(tests/test_stencil_selection.py).  It follows

    required_stencil_size / deduce_max_order   src/zisa/reconstruction/stencil.cpp:158-175
    query points                               stencil.cpp:178-190 (degree-5 triangle rule / degree-3 tetrahedron rule)
    region_based_candidates                    stencil.cpp:192-238
    region_based_stencil                       stencil.cpp:240-256
    make_cone, cones' membership tests         stencil.cpp:258-278, src/zisa/math/cone.cpp:7-30
    conservative / less conservative stencil   stencil.cpp:280-301
    tryhard_stencil (up to its random retries) stencil.cpp:303-345
    biased_stencil, central_stencil            stencil.cpp:347-399
    StencilFamily (k_biased counter, families of ghost cells)   stencil_family.cpp:14-45, 99-117
    assign_local_indices                       stencil.cpp:82-104
    relative_vertex_index / off vertex         src/zisa/grid/gmsh_reader.cpp:22-82

Inputs are plain numpy arrays of a grid (vertices, vertex_indices, neighbours, cell centres, cell quadrature); the rank
test of `is_good` builds the least-squares matrix from its definition (cell averages of the centre cell's zero-mean
scaled monomials) by quadrature, not from the reference's expanded formulas, and takes numpy's SVD with Eigen's
JacobiSVD::rank() threshold.  The random retries of tryhard_stencil (std::random_device) cannot be restated: reaching
them raises NeedsRandomRetry and the test skips that cell.
"""
from __future__ import annotations

import numpy as np

INVALID = -1

# degree-5 triangle rule and degree-3 tetrahedron rule in barycentric coordinates (triangular_rule.cpp:62-91,
# tetrahedral_rule.cpp:60-122); only the points matter here
_TRI5 = None
_TET3 = None


def _tri5():
    a1, b1 = 0.059715871789770, 0.470142064105115
    a2, b2 = 0.797426985353087, 0.101286507323456
    pts = [(1 / 3, 1 / 3, 1 / 3)]
    for a, b in ((a1, b1), (a2, b2)):
        pts += [(a, b, b), (b, a, b), (b, b, a)]
    return np.array(pts)


def _tet3():
    a, b = 0.7784952948213300, 0.0738349017262234
    c, d = 0.4062443438840510, 0.0937556561159491
    pts = [(a, b, b, b), (b, a, b, b), (b, b, a, b), (b, b, b, a)]
    pts += [(c, c, d, d), (c, d, c, d), (c, d, d, c), (d, c, c, d), (d, c, d, c), (d, d, c, c)]
    return np.array(pts)


class NeedsRandomRetry(Exception):
    pass


def poly_dof(deg: int, n_dims: int) -> int:
    return (deg + 1) * (deg + 2) // 2 if n_dims == 2 else (deg + 1) * (deg + 2) * (deg + 3) // 6


def required_stencil_size(deg: int, factor: float, n_dims: int) -> int:
    if deg == 0:
        return 1
    return int(float(poly_dof(deg, n_dims) - 1) * factor + 1)


def deduce_max_order(stencil_size: int, factor: float, n_dims: int) -> int:
    deg = 0
    while required_stencil_size(deg + 1, factor, n_dims) <= stencil_size:
        deg += 1
    return deg + 1


def relative_vertex_index(n_dims: int, k: int, rel: int) -> int:
    if n_dims == 2:
        return (k + rel) % 3
    return {0: (0, 1, 3), 1: (0, 2, 1), 2: (0, 3, 2), 3: (1, 2, 3)}[k][rel]


def relative_off_vertex_index(n_dims: int, k: int) -> int:
    if n_dims == 2:
        return (k + 2) % 3
    return {0: 2, 1: 3, 2: 1, 3: 0}[k]


def _det3(a, b, c):
    return float(np.dot(a, np.cross(b, c)))


class Cone:
    def __init__(self, n_dims, apex, pts):
        self.n_dims = n_dims
        self.A = np.asarray(apex, dtype=float)
        self.d = [np.asarray(p, dtype=float) - self.A for p in pts]

    def is_inside(self, x) -> bool:
        dx = np.asarray(x, dtype=float) - self.A
        if self.n_dims == 2:
            dB, dC = self.d
            return np.cross(dB, dx)[2] >= 0.0 and np.cross(dx, dC)[2] >= 0.0
        dB, dC, dD = self.d
        return _det3(dB, dC, dx) >= 0.0 and _det3(dC, dD, dx) >= 0.0 and _det3(dD, dB, dx) >= 0.0


class FullSphere:
    def is_inside(self, x) -> bool:
        return True


class Selection:
    def __init__(self, n_dims, vertices, vertex_indices, neighbours, cell_centers, cell_qp, cell_qw, volumes, char_length,
                 cell_flags):
        self.nd = n_dims
        self.F = n_dims + 1
        self.v = np.asarray(vertices, dtype=float)
        self.vi = np.asarray(vertex_indices)
        self.nb = np.asarray(neighbours)
        self.cc = np.asarray(cell_centers, dtype=float)
        self.qp, self.qw, self.vol, self.len = cell_qp, cell_qw, volumes, char_length
        self.flags = cell_flags
        self.query = _tri5() if n_dims == 2 else _tet3()

    # stencil.cpp:178-190
    def query_points(self, i):
        return self.query @ self.v[self.vi[i]]

    # stencil.cpp:192-238
    def candidates(self, i_center, n_points, region):
        max_points = 5 * n_points
        cands = [i_center]

        def is_inside(c):
            for x in self.query_points(c):
                if region.is_inside(x):
                    return True
            return region.is_inside(self.cc[c])

        p = 0
        while p < max_points and p < len(cands):
            j = cands[p]
            for k in range(self.F):
                c = int(self.nb[j, k])
                if c == INVALID:
                    continue
                if c not in cands and is_inside(c):
                    cands.append(c)
            p += 1
        return cands

    # stencil.cpp:240-256
    def region_stencil(self, i_center, n_points, region):
        cands = self.candidates(i_center, n_points, region)
        xc = self.cc[i_center]
        dist = {c: float(np.sqrt(((self.cc[c] - xc) ** 2).sum())) for c in cands}
        if len(set(dist.values())) != len(dist):
            raise NeedsRandomRetry("equidistant candidates: std::sort's permutation is unspecified")
        cands.sort(key=lambda c: dist[c])
        return cands[: min(n_points, len(cands))]

    # stencil.cpp:258-278
    def make_cone(self, i_center, apex, k):
        pts = [self.v[self.vi[i_center, relative_vertex_index(self.nd, k, r)]] for r in range(self.nd)]
        return Cone(self.nd, apex, pts)

    def face_center(self, i, k):
        pts = [self.v[self.vi[i, relative_vertex_index(self.nd, k, r)]] for r in range(self.nd)]
        return np.mean(pts, axis=0)

    # least-squares matrix from its definition (what lsq_solver.cpp:168-403 expands in closed form)
    def lsq_matrix(self, s, order):
        i0 = s[0]
        expo = []
        for n in range(1, order):
            if self.nd == 2:
                expo += [(n - b, b, 0) for b in range(n + 1)]
            else:
                for m in range(n + 1):
                    expo += [(n - m, m - c, c) for c in range(m + 1)]

        def averages(j):
            xi = (self.qp[j] - self.cc[i0]) / self.len[i0]
            return np.array([(self.qw[j] * xi[:, 0] ** a * xi[:, 1] ** b * xi[:, 2] ** c).sum() / self.vol[j] for a, b, c in expo])

        own = averages(i0)
        own[: self.nd] = 0.0   # the linear moments vanish (the centre is the barycentre)
        return np.array([averages(j) - own for j in s[1:]])

    def is_good(self, s, order):
        A = self.lsq_matrix(list(s), order)
        sv = np.linalg.svd(A, compute_uv=False)
        thresh = sv.max() * min(A.shape) * np.finfo(float).eps   # Eigen JacobiSVD::rank()
        return int((sv > thresh).sum()) == A.shape[1]

    # stencil.cpp:347-393
    def biased_stencil(self, i_center, k, n_points, order):
        off = self.v[self.vi[i_center, relative_off_vertex_index(self.nd, k)]]
        s = self.region_stencil(i_center, n_points, self.make_cone(i_center, off, k))
        if len(s) == n_points and self.is_good(s, order):
            return s
        s = self.region_stencil(i_center, n_points, self.make_cone(i_center, self.cc[i_center], k))
        if len(s) == n_points and self.is_good(s, order):
            return s
        cands = self.candidates(i_center, n_points, self.make_cone(i_center, self.face_center(i_center, k), k))
        if len(cands) < n_points:
            return [i_center]
        raise NeedsRandomRetry("tryhard_stencil would draw random permutations")

    def central_stencil(self, i_center, n_points):
        return self.region_stencil(i_center, n_points, FullSphere())

    # stencil_family.cpp:14-45 with the family choice of :99-117
    def family(self, i, orders, biases, factors):
        interior = bool(self.flags[i] & 1) or bool(self.flags[i] & 4)   # interior || ghost_cell_l1
        if not interior:
            orders, biases, factors = [1], ["c"], [1.0]
        l2g, out = [], []
        k_biased = 0
        for o, b, f in zip(orders, biases, factors):
            max_size = required_stencil_size(o - 1, f, self.nd)
            if b == "b":
                s = self.biased_stencil(i, k_biased, max_size, o)
                k_biased += 1
            else:
                s = self.central_stencil(i, max_size)
            local = []
            for c in s:   # assign_local_indices
                if c in l2g:
                    local.append(l2g.index(c))
                else:
                    local.append(len(l2g))
                    l2g.append(c)
            order = deduce_max_order(len(s), f, self.nd)
            size = required_stencil_size(order - 1, f, self.nd)
            out.append({"global": s, "local": local, "order": order, "size": size})
        return out, l2g
