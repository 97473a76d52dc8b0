"""Shared helpers of the parity tests (test infrastructure)."""
import numpy as np


def rel_err(a, b):
    """max |a - b| / max |b| per conserved variable (columns)."""
    den = np.maximum(np.abs(b).max(axis=0), 1e-300)
    return np.abs(a - b).max(axis=0) / den


def rel_l1(a, b, w=None):
    w = np.ones(a.shape[0]) if w is None else w
    den = np.maximum((np.abs(b) * w[:, None]).sum(axis=0), 1e-300)
    return (np.abs(a - b) * w[:, None]).sum(axis=0) / den


def active_vars(n_dims):
    return [0, 1, 2, 4] if n_dims == 2 else [0, 1, 2, 3, 4]


def state_scales(u, gamma):
    """Acoustic scales of the conserved variables: (rho, rho a, rho a, rho a, E).  Round-off in a residual is
    proportional to the magnitude of the fluxes that cancel in it, not to the (possibly vanishing) residual."""
    rho = u[:, 0]
    p = (gamma - 1.0) * (u[:, 4] - 0.5 * (u[:, 1:4] ** 2).sum(axis=1) / rho)
    a = np.sqrt(gamma * np.abs(p) / rho)
    ra = (rho * a).max()
    return np.array([rho.max(), ra, ra, ra, u[:, 4].max()])


def tendency_scales(u, gamma, inradius):
    """Magnitude of the individual flux contributions to d/dt of each variable: state scale * a / inradius."""
    rho = u[:, 0]
    p = (gamma - 1.0) * (u[:, 4] - 0.5 * (u[:, 1:4] ** 2).sum(axis=1) / rho)
    a = np.sqrt(gamma * np.abs(p) / rho)
    return state_scales(u, gamma) * (a / inradius).max()
