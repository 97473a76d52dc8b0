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
