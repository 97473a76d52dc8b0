"""Synthetic set-ups of the BASELINE.json configurations.

The reference ships neither grids nor these experiments; grids come from the seeded
generators of the host library, initial data are cell averages of analytic fields taken with the cell
quadrature rule (like ``average(cell, ic)`` in src/zisa/experiments/polytrope.cpp:45-49).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, Optional

import numpy as np

from .grid import (Grid, HybridWENOParams, QRDegrees, WENO_PARAMS, compute_stencil_families, cube_mesh,
                   square_mesh)
from .solver import EulerParams, Gravity


@dataclass
class Case:
    name: str
    grid: Grid
    params: EulerParams
    u0: np.ndarray          # [n_cells][5]
    method: str             # Butcher tableau name
    cfl: float
    frozen_bc: bool = True
    stencils: object = None
    a0: Optional[np.ndarray] = None   # [n_cells][n_avars] advected scalars (conserved form rho * q)

    def ensure_stencils(self):
        if self.stencils is None:
            self.stencils = compute_stencil_families(self.grid, self.params.weno.stencil_family_params)
        return self.stencils


def cell_average(grid: Grid, f: Callable[[np.ndarray], np.ndarray]) -> np.ndarray:
    """average(cell, f) with the grid's cell rule; f maps points [m][3] -> values [m][k].  Large grids are evaluated in
    blocks of cells on a thread pool (numpy releases the GIL inside its loops); the per-cell arithmetic is unchanged."""
    qp = grid.array("cell_qp")          # [n][q][3]
    qw = grid.array("cell_qw")          # [n][q]
    vol = grid.array("volumes")
    n, q, _ = qp.shape

    def block(lo, hi):
        m = hi - lo
        vals = f(qp[lo:hi].reshape(-1, 3)).reshape(m, q, -1)
        acc = qw[lo:hi, 0, None] * vals[:, 0]
        for k in range(1, q):
            acc = acc + qw[lo:hi, k, None] * vals[:, k]
        return acc / vol[lo:hi, None]

    if n < 400_000:
        return block(0, n)
    import os
    from concurrent.futures import ThreadPoolExecutor

    threads = max(1, min(32, len(os.sched_getaffinity(0))))
    step = 131_072
    bounds = [(lo, min(n, lo + step)) for lo in range(0, n, step)]
    with ThreadPoolExecutor(threads) as pool:
        parts = list(pool.map(lambda b: block(*b), bounds))
    return np.concatenate(parts, axis=0)


def cvars_from_primitive(rho, v, p, gamma):
    u = np.zeros((rho.shape[0], 5))
    u[:, 0] = rho
    u[:, 1:4] = rho[:, None] * v
    u[:, 4] = p / (gamma - 1.0) + 0.5 * rho * np.sum(v * v, axis=1)
    return u


def with_tracers(case: "Case", n_avars: int = 1, box=None) -> "Case":
    """Adds advected scalars rho * q_a with smooth concentrations q_a (the reference's Rayleigh-Taylor set-up carries
    one, src/zisa/experiments/rayleigh_taylor.cpp:26): cell averages of rho(x) q_a(x) with rho taken as the cell's
    average density (a second-order accurate initial field, which is all the parity tests need)."""
    c = case.grid.array("cell_centers")
    if box is None:  # concentrations in coordinates relative to the grid's bounding box ...
        lo, span = c.min(axis=0), np.maximum(c.max(axis=0) - c.min(axis=0), 1e-300)
    else:            # ... or to a given box (lo, hi): sub-domains of one global mesh then carry the same field
        lo, span = np.asarray(box[0], dtype=float), np.asarray(box[1], dtype=float) - np.asarray(box[0], dtype=float)
    xi = (c - lo) / span
    a0 = np.zeros((case.grid.n_cells, n_avars))
    for a in range(n_avars):
        q = 0.6 + 0.3 * np.sin((2.0 + a) * np.pi * xi[:, 0]) * np.cos((1.0 + a) * np.pi * xi[:, 1]) + 0.1 * a * xi[:, 2]
        if a % 2 == 1:  # a discontinuous one as well: the non-linear weights must switch
            q = np.where(xi[:, 0] + 0.5 * xi[:, 1] < 0.7, 1.0, 0.1) + 0.05 * xi[:, 1]
        a0[:, a] = case.u0[:, 0] * q
    case.a0 = a0
    case.params.n_avars = n_avars
    return case


def ghost_ring(grid: Grid, lo, hi, width):
    """Cells whose centre is within `width` of the box boundary are ghost cells."""
    c = grid.array("cell_centers")
    nd = grid.n_dims
    mask = np.zeros(grid.n_cells, dtype=bool)
    for d in range(nd):
        mask |= (c[:, d] < lo[d] + width) | (c[:, d] > hi[d] - width)
    return mask


def renumber_ghosts_last(grid: Grid, verts: np.ndarray, qr: QRDegrees) -> Grid:
    """The same mesh with the ghost cells nobody reconstructs (ghost_cell without ghost_cell_l1: first-order families,
    stencil_family.cpp:108-114) numbered behind all other cells; both groups keep their relative order."""
    flags = np.array(grid.array("cell_flags"))
    deep = ((flags & 2) != 0) & ((flags & 4) == 0)
    perm = np.concatenate([np.nonzero(~deep)[0], np.nonzero(deep)[0]])
    vi = np.array(grid.array("vertex_indices"))[perm]
    out = Grid(grid.n_dims, verts, vi, qr)
    out.set_flags(flags[perm])
    return out


# ---- C1: 2D isentropic vortex ------------------------------------------------------------------------
def isentropic_vortex(n: int = 158, order: int = 3, flux: str = "hllc", seed: int = 0, jitter: float = 0.15,
                      ghost_ring_cells: int = 3, flux_bc: str = "none", reconstruction: str = "CWENO-AO",
                      scaling: str = "euler", method: str = "ssp3", weno: Optional[HybridWENOParams] = None) -> Case:
    """``ghost_ring_cells = 0, flux_bc = "flux"``: no ghost ring, the domain boundary is closed with ``FluxBC``
    (boundary/flux_bc.hpp) -- the set-up of the reference's domains without a halo of frozen cells."""
    gamma = 1.4
    verts, vi = square_mesh(n, n, 0.0, 10.0, 0.0, 10.0, jitter=jitter, seed=seed)
    grid = Grid(2, verts, vi, QRDegrees(face_deg=3, volume_deg=3, moments_deg=4))
    h = 10.0 / n
    if ghost_ring_cells > 0:
        grid.mask_ghost_cells(ghost_ring(grid, (0.0, 0.0), (10.0, 10.0), ghost_ring_cells * h))
    beta = 5.0

    def ic(x):
        dx, dy = x[:, 0] - 5.0, x[:, 1] - 5.0
        r2 = dx * dx + dy * dy
        e = np.exp(0.5 * (1.0 - r2))
        vel = np.zeros((x.shape[0], 3))
        vel[:, 0] = -beta / (2 * np.pi) * e * dy
        vel[:, 1] = beta / (2 * np.pi) * e * dx
        T = 1.0 - (gamma - 1.0) * beta * beta / (8 * gamma * np.pi ** 2) * e * e
        rho = T ** (1.0 / (gamma - 1.0))
        return cvars_from_primitive(rho, vel, rho * T, gamma)

    params = EulerParams(weno=weno or WENO_PARAMS[f"2d_o{order}"], flux=flux, gamma=gamma, flux_bc=flux_bc,
                         reconstruction=reconstruction, scaling=scaling)
    return Case("isentropic_vortex", grid, params, cell_average(grid, ic), method, 0.4, frozen_bc=ghost_ring_cells > 0)


# ---- C2: 2D well-balanced polytrope ------------------------------------------------------------------
def polytrope_alpha(G=1.0, K=1.0):
    return np.sqrt(2.0 * np.pi * G / K)


def polytrope_2d(n: int = 158, order: int = 3, well_balanced: bool = True, amplitude: float = 0.0,
                 width: float = 0.05, seed: int = 0, ghost: bool = True, flux_bc: str = "none") -> Case:
    """gamma = 2 polytrope in hydrostatic equilibrium (src/zisa/experiments/polytrope.cpp:12-63)."""
    gamma = 2.0
    verts, vi = square_mesh(n, n, -0.6, 0.6, -0.6, 0.6, jitter=0.15, seed=seed)
    grid = Grid(2, verts, vi, QRDegrees(face_deg=3, volume_deg=3, moments_deg=4))
    c = grid.array("cell_centers")
    if ghost:
        grid.mask_ghost_cells(np.linalg.norm(c, axis=1) > 0.5)  # boundary_mask, polytrope.cpp:55-63
    alpha = polytrope_alpha()

    def ic(x):
        r = np.linalg.norm(x, axis=1)
        r_eff = alpha * (r + np.finfo(float).tiny)
        rho = np.sin(r_eff) / r_eff
        p = rho * rho * (1.0 + amplitude * np.exp(-((r / width) ** 2)))
        return cvars_from_primitive(rho, np.zeros((x.shape[0], 3)), p, gamma)

    params = EulerParams(
        weno=WENO_PARAMS[f"2d_o{order}"], gamma=gamma,
        well_balancing="isentropic" if well_balanced else "constant",
        gravity=Gravity(kind="polytrope", params=(1.0, 1.0, 1.0), alignment="radial"), flux_bc=flux_bc,
    )
    return Case("polytrope_2d", grid, params, cell_average(grid, ic), "ssp3", 0.4, frozen_bc=ghost)


def blast_ic(kind: str, gamma: float = 1.4):
    """Initial data of the 3D set-ups on [0,1]^3: ``kind`` in {"blast", "sod", "smooth"}."""

    def ic(x):
        m = x.shape[0]
        vel = np.zeros((m, 3))
        if kind == "sod":
            left = x[:, 0] < 0.5
            rho = np.where(left, 1.0, 0.125)
            p = np.where(left, 1.0, 0.1)
        elif kind == "blast":
            r = np.linalg.norm(x - 0.5, axis=1)
            rho = np.ones(m)
            p = np.where(r < 0.1, 10.0, 0.1)
        else:  # smooth: moving density wave, used for accuracy / parity at high order
            rho = 1.0 + 0.2 * np.sin(2 * np.pi * x[:, 0]) * np.cos(2 * np.pi * x[:, 1]) * np.cos(2 * np.pi * x[:, 2])
            vel[:, 0], vel[:, 1], vel[:, 2] = 0.3, -0.2, 0.1
            p = 1.0 + 0.1 * np.cos(2 * np.pi * x[:, 0])
        return cvars_from_primitive(rho, vel, p, gamma)

    return ic


def blast_qr(order: int) -> QRDegrees:
    return QRDegrees(face_deg=2 if order == 2 else 3, volume_deg=2, moments_deg=max(order - 1, 2))


def blast_3d_on_grid(grid: Grid, order: int = 3, kind: str = "blast", stencils=None) -> Case:
    """The C3 set-up on a grid whose ghost flags are already set (sub-domains of a decomposed run)."""
    gamma = 1.4
    params = EulerParams(weno=WENO_PARAMS[f"3d_o{order}"], gamma=gamma)
    method = "ssp2" if order == 2 else "ssp3"
    return Case(f"{kind}_3d_o{order}", grid, params, cell_average(grid, blast_ic(kind, gamma)), method, 0.4,
                stencils=stencils)


# ---- C3: 3D Sod / blast --------------------------------------------------------------------------------
def blast_3d(n: int = 16, order: int = 3, kind: str = "blast", seed: int = 0, ghost_cubes: int = 2,
             hilbert: bool = True, offset=None, global_n: Optional[int] = None, shape=None, flux_bc: str = "none",
             reconstruction: str = "CWENO-AO", scaling: str = "euler", method: Optional[str] = None,
             weno: Optional[HybridWENOParams] = None, ghosts_last: bool = False) -> Case:
    """[0,1]^3 (n^3 cubes x 6 Kuhn tetrahedra), gamma = 1.4; `kind` in {"blast", "sod", "smooth"}.

    ``ghosts_last``: cells are numbered like the reference numbers a partition -- the cells whose reconstruction
    somebody reads first (interior and ghost_cell_l1, in Hilbert order), the remaining ghost cells behind them
    (domain_decomposition.cpp:300-326 puts the halo behind the owned cells) -- so that tiles of 32 consecutive cells are
    either reconstructed or skipped as a whole."""
    gamma = 1.4
    gn = global_n or n
    h = 1.0 / gn
    nx, ny, nz = shape if shape is not None else (n, n, n)
    verts, vi = cube_mesh(nx, ny, nz, h, jitter=0.1, seed=seed, hilbert=hilbert, offset=offset,
                          global_shape=(gn, gn, gn) if offset is not None else None)
    grid = Grid(3, verts, vi, blast_qr(order))
    if ghost_cubes > 0:
        grid.mask_ghost_cells(ghost_ring(grid, (0.0, 0.0, 0.0), (1.0, 1.0, 1.0), ghost_cubes * h))
        if ghosts_last:
            grid = renumber_ghosts_last(grid, verts, blast_qr(order))

    ic = blast_ic(kind, gamma)

    params = EulerParams(weno=weno or WENO_PARAMS[f"3d_o{order}"], gamma=gamma, flux_bc=flux_bc,
                         reconstruction=reconstruction, scaling=scaling)
    method = method or ("ssp2" if order == 2 else "ssp3")
    return Case(f"{kind}_3d_o{order}", grid, params, cell_average(grid, ic), method, 0.4, frozen_bc=ghost_cubes > 0)


# ---- C4: 3D stellar atmosphere -----------------------------------------------------------------------------
def stellar_atmosphere_3d(n: int = 12, order: int = 3, well_balanced: bool = True, amplitude: float = 1e-3,
                          seed: int = 0, gravity: str = "point_mass", scaling: str = "euler") -> Case:
    """Isentropic hydrostatic atmosphere in a softened point-mass potential, gamma = 5/3 ideal gas,
    plus a pressure perturbation.  Orders 2 and 3 have compiled 3D kernels."""
    gamma = 5.0 / 3.0
    h = 2.0 / n
    verts, vi = cube_mesh(n, n, n, h, origin=(-1.0, -1.0, -1.0), jitter=0.1, seed=seed)
    vdeg = 3 if order >= 3 else 2
    grid = Grid(3, verts, vi, QRDegrees(face_deg=min(order, 4), volume_deg=vdeg, moments_deg=max(order - 1, 2)))
    grid.mask_ghost_cells(ghost_ring(grid, (-1.0,) * 3, (1.0,) * 3, 2 * h))
    GM, X = -1.0, 1.0  # PointMassGravity: phi = GM / (X + r)
    h_c, K = 4.0, 1.0

    def ic(x):
        r = np.linalg.norm(x, axis=1)
        phi = GM / (X + r)
        hh = h_c + GM / X - phi
        rho = ((gamma - 1.0) / (gamma * K) * hh) ** (1.0 / (gamma - 1.0))
        p = K * rho ** gamma * (1.0 + amplitude * np.exp(-((r / 0.3) ** 2)))
        return cvars_from_primitive(rho, np.zeros((x.shape[0], 3)), p, gamma)

    if gravity == "table":
        # the same potential as a RadialGravity table (gravity_decl.hpp:312-338, piecewise linear in r): fine enough that
        # the hydrostatic state above stays a near-equilibrium of the tabulated potential
        radii = np.linspace(0.0, 2.0, 4001)
        grav = Gravity(kind="table", alignment="radial", table=(radii, GM / (X + radii)))
    else:
        grav = Gravity(kind="point_mass", params=(GM, X), alignment="radial")
    params = EulerParams(
        weno=WENO_PARAMS[f"3d_o{order}"], gamma=gamma, scaling=scaling,
        well_balancing="isentropic" if well_balanced else "constant", gravity=grav,
    )
    return Case("stellar_atmosphere_3d", grid, params, cell_average(grid, ic), "ssp3", 0.4)


def constant_gravity_2d(n: int = 32, order: int = 3, well_balanced: bool = True, amplitude: float = 1e-3,
                        seed: int = 0, axis=(0.0, 1.0, 0.0)) -> Case:
    """Isentropic hydrostatic layer in a constant gravitational field along ``axis`` (``ConstantGravityAxial``,
    gravity_impl.hpp:13-20 with ``AxialAlignment``, gravity_decl.hpp:97-120): phi = g (x . axis), gamma = 1.4,
    plus a Gaussian pressure perturbation.  The set-up of the reference's Rayleigh-Taylor / gaussian-bump experiments
    without the tracer."""
    gamma, g_acc = 1.4, 1.0
    verts, vi = square_mesh(n, n, 0.0, 1.0, 0.0, 1.0, jitter=0.15, seed=seed)
    grid = Grid(2, verts, vi, QRDegrees(face_deg=3, volume_deg=3, moments_deg=4))
    grid.mask_ghost_cells(ghost_ring(grid, (0.0, 0.0), (1.0, 1.0), 3.0 / n))
    ax = np.asarray(axis, dtype=float)
    h0, K = 3.0, 1.0

    def ic(x):
        phi = g_acc * (x @ ax)
        hh = h0 - phi
        rho = ((gamma - 1.0) / (gamma * K) * hh) ** (1.0 / (gamma - 1.0))
        r2 = (x[:, 0] - 0.5) ** 2 + (x[:, 1] - 0.5) ** 2
        p = K * rho ** gamma * (1.0 + amplitude * np.exp(-r2 / 0.05 ** 2))
        return cvars_from_primitive(rho, np.zeros((x.shape[0], 3)), p, gamma)

    params = EulerParams(
        weno=WENO_PARAMS[f"2d_o{order}"], gamma=gamma,
        well_balancing="isentropic" if well_balanced else "constant",
        gravity=Gravity(kind="constant", params=(g_acc,), alignment="axial", axis=tuple(axis)),
    )
    return Case("constant_gravity_2d", grid, params, cell_average(grid, ic), "ssp3", 0.4)


def gravity_tables(grid: Grid, gravity: Gravity):
    """phi / grad phi at all quadrature points, in numpy (independent of the host library's tables).
    Follows include/zisa/model/gravity_impl.hpp:13-58 and RadialAlignment (gravity_decl.hpp:76-95)."""
    def phi_dphi(chi):
        kind, p = gravity.kind, gravity.params
        if kind == "constant":
            return p[0] * chi, np.full_like(chi, p[0])
        if kind == "point_mass":
            return p[0] / (p[1] + chi), -p[0] / (p[1] + chi) ** 2
        if kind == "polytrope":
            rhoC, K, G = p[:3]
            alpha = np.sqrt(2.0 * np.pi * G / K)
            ce = alpha * (chi + np.finfo(float).tiny)
            return -2.0 * K * rhoC * np.sin(ce) / ce, -2.0 * K * rhoC * ((np.cos(ce) - np.sin(ce) / ce) / ce) * alpha
        if kind == "table":  # NonUniformLinearInterpolation (math/linear_interpolation.hpp:14-45)
            pts, val = (np.asarray(a, dtype=float) for a in gravity.table)
            i = np.searchsorted(pts, chi, side="left")      # std::lower_bound
            i = np.minimum(np.where(i == 0, 0, i - 1), pts.size - 2)
            a = (chi - pts[i]) / (pts[i + 1] - pts[i])
            return (1 - a) * val[i] + a * val[i + 1], (val[i + 1] - val[i]) / (pts[i + 1] - pts[i])
        raise ValueError(kind)

    def at(x):
        if gravity.alignment == "radial":
            r = np.sqrt(x[:, 0] ** 2 + x[:, 1] ** 2 + x[:, 2] ** 2)
            ph, dph = phi_dphi(r)
            return ph, dph[:, None] * (x / (r + 1e-50)[:, None])
        ax = np.asarray(gravity.axis, dtype=float)
        chi = x @ ax
        ph, dph = phi_dphi(chi)
        return ph, dph[:, None] * ax[None, :]

    cq = grid.array("cell_qp").reshape(-1, 3)
    fq = grid.array("face_qp").reshape(-1, 3)
    phi_c, g_c = at(cq)
    phi_f, _ = at(fq)
    return (phi_c.reshape(grid.n_cells, grid.q_c), g_c.reshape(grid.n_cells, grid.q_c, 3),
            phi_f.reshape(grid.n_edges, grid.q_f))
