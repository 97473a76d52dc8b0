"""ctypes view of ``include/zfvm.h`` (the C ABI of ``libzfvm_b200.so``).

The library is built in-tree by ``__graft_entry__.build()`` (``zisafvm_b200/csrc/Makefile``); there is
no fallback of any kind when it is missing: importing this module raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libzfvm_b200.so")

c_double_p = C.POINTER(C.c_double)
c_int32_p = C.POINTER(C.c_int32)
c_int64_p = C.POINTER(C.c_int64)
c_uint8_p = C.POINTER(C.c_uint8)


class ZfvmParams(C.Structure):
    """``zfvm_params`` of include/zfvm.h."""

    _fields_ = [
        ("recon_mode", C.c_int),
        ("linear_weights", C.c_double * 8),
        ("epsilon", C.c_double),
        ("exponent", C.c_double),
        ("well_balanced", C.c_int),
        ("scaling", C.c_int),
        ("flux", C.c_int),
        ("gamma", C.c_double),
        ("gas_constant", C.c_double),
        ("gravity_kind", C.c_int),
        ("gravity_alignment", C.c_int),
        ("gravity_p", C.c_double * 4),
        ("gravity_axis", C.c_double * 3),
        ("steps_per_recompute", C.c_int),
        ("keep_polynomials", C.c_int),
        ("flux_bc", C.c_int),
        ("n_avars", C.c_int),
        ("heating_rate", C.c_double),
        ("heating_r0", C.c_double),
        ("heating_r1", C.c_double),
        ("recompute_threshold", C.c_double),
    ]


class ZfvmError(RuntimeError):
    pass


def _load() -> C.CDLL:
    global LIB_PATH
    LIB_PATH = os.environ.get("ZFVM_LIB_PATH", LIB_PATH)  # kernel experiments: a variant build of the same library
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(make -C zisafvm_b200/csrc). The B200 path has no CPU or PyTorch fallback."
        )
    return C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)


lib = _load()

_vp = C.c_void_p
_SIGNATURES = {
    "zfvm_last_error": (C.c_char_p, []),
    "zfvm_version": (C.c_int, []),
    "zfvm_grid_from_mesh": (C.c_int, [C.c_int, C.c_int64, c_double_p, C.c_int64, c_int32_p, C.c_int, C.c_int, C.c_int, C.POINTER(_vp)]),
    "zfvm_grid_mask_ghost": (C.c_int, [_vp, c_uint8_p]),
    "zfvm_grid_set_flags": (C.c_int, [_vp, c_uint8_p]),
    "zfvm_grid_get": (C.c_int, [_vp, C.c_char_p, C.POINTER(_vp), C.POINTER(C.c_int), C.POINTER(C.c_int), c_int64_p]),
    "zfvm_grid_info": (C.c_int, [_vp, c_int64_p]),
    "zfvm_grid_free": (None, [_vp]),
    "zfvm_mesh_square": (C.c_int, [C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double, C.c_uint64, C.c_int,
                                   c_int64_p, C.POINTER(c_double_p), c_int64_p, C.POINTER(c_int32_p)]),
    "zfvm_mesh_cube": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double, C.c_uint64,
                                 C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), c_int64_p, C.POINTER(c_double_p), c_int64_p,
                                 C.POINTER(c_int32_p)]),
    "zfvm_mesh_read_msh_h5": (C.c_int, [C.c_char_p, C.POINTER(C.c_int), c_int64_p, C.POINTER(c_double_p), c_int64_p,
                                        C.POINTER(c_int32_p)]),
    "zfvm_mesh_write_msh_h5": (C.c_int, [C.c_char_p, C.c_int, C.c_int64, c_double_p, C.c_int64, c_int32_p]),
    "zfvm_mesh_read_subgrid_h5": (C.c_int, [C.c_char_p, C.POINTER(C.c_int), c_int64_p, C.POINTER(c_double_p), c_int64_p,
                                            C.POINTER(c_int32_p), C.POINTER(c_int64_p), C.POINTER(c_int64_p)]),
    "zfvm_mesh_write_subgrid_h5": (C.c_int, [C.c_char_p, C.c_int, C.c_int64, c_double_p, C.c_int64, c_int32_p, c_int64_p,
                                             c_int64_p]),
    "zfvm_free": (None, [_vp]),
    "zfvm_stencils_compute": (C.c_int, [_vp, C.c_int, C.POINTER(C.c_int), C.c_char_p, c_double_p, C.c_uint64, C.POINTER(_vp)]),
    "zfvm_stencils_get": (C.c_int, [_vp, C.c_char_p, C.POINTER(_vp), C.POINTER(C.c_int), C.POINTER(C.c_int), c_int64_p]),
    "zfvm_stencils_free": (None, [_vp]),
    "zfvm_stencils_extract": (C.c_int, [_vp, C.c_int64, c_int32_p, C.POINTER(_vp)]),
    "zfvm_stencils_from_arrays": (C.c_int, [_vp, C.c_int, C.POINTER(C.c_int), C.c_char_p, c_double_p, c_int32_p, c_int32_p, c_int32_p,
                                            c_int64_p, c_int32_p, C.POINTER(_vp)]),
    "zfvm_partition_kway": (C.c_int, [_vp, _vp, C.c_int, c_int32_p]),
    "zfvm_has_metis": (C.c_int, []),
    "zfvm_hilbert_permutation": (C.c_int, [C.c_int, C.c_int64, c_double_p, c_int32_p]),
    "zfvm_stencil_matrix": (C.c_int, [_vp, _vp, C.c_int64, C.c_int, c_double_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "zfvm_stencil_matrices": (C.c_int, [_vp, _vp, c_double_p, C.c_int64, c_int64_p]),
    "zfvm_pseudo_inverse": (C.c_int, [c_double_p, C.c_int, C.c_int, c_double_p]),
    "zfvm_quadrature_rule": (C.c_int, [C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), c_double_p, c_double_p, C.c_int]),
    "zfvm_gauss_legendre": (C.c_int, [C.c_int, c_double_p, c_double_p]),
    "zfvm_deduce_max_order": (C.c_int, [C.c_int, C.c_double, C.c_int]),
    "zfvm_params_default": (None, [C.POINTER(ZfvmParams)]),
    "zfvm_create": (C.c_int, [_vp, _vp, C.POINTER(ZfvmParams), C.c_int, C.POINTER(_vp)]),
    "zfvm_destroy": (None, [_vp]),
    "zfvm_set_gravity_table": (C.c_int, [_vp, _vp, C.c_int64, c_double_p, c_double_p]),
    "zfvm_set_gravity_values": (C.c_int, [_vp, c_double_p, c_double_p, c_double_p]),
    "zfvm_memory_info": (C.c_int, [_vp, c_int64_p, c_double_p]),
    "zfvm_stream": (_vp, [_vp]),
    "zfvm_rate_of_change": (C.c_int, [_vp, c_double_p, c_double_p, C.c_double, C.c_int]),
    "zfvm_rate_of_change_device": (C.c_int, [_vp, _vp, _vp, C.c_double, C.c_int]),
    "zfvm_set_time_integration": (C.c_int, [_vp, C.c_char_p]),
    "zfvm_upload_state": (C.c_int, [_vp, c_double_p]),
    "zfvm_download_state": (C.c_int, [_vp, c_double_p]),
    "zfvm_state_device": (_vp, [_vp]),
    "zfvm_set_frozen_bc": (C.c_int, [_vp, c_double_p]),
    "zfvm_apply_frozen_bc": (C.c_int, [_vp, _vp]),
    "zfvm_rk_step": (C.c_int, [_vp, C.c_double, C.c_double, C.c_double, c_double_p, C.POINTER(C.c_int)]),
    "zfvm_rk_step_host": (C.c_int, [_vp, c_double_p, c_double_p, C.c_double, C.c_double]),
    "zfvm_cfl_dt": (C.c_int, [_vp, _vp, C.c_double, c_double_p, C.POINTER(C.c_int)]),
    "zfvm_synchronize": (C.c_int, [_vp]),
    "zfvm_counters": (C.c_int, [_vp, c_int64_p]),
    "zfvm_profile_enable": (C.c_int, [_vp, C.c_int]),
    "zfvm_profile_read": (C.c_int, [_vp, c_double_p, c_int64_p]),
    "zfvm_profile_read_tracers": (C.c_int, [_vp, c_double_p, c_int64_p]),
    "zfvm_download_polynomials": (C.c_int, [_vp, c_double_p, c_double_p, C.POINTER(C.c_int)]),
    "zfvm_download_work": (C.c_int, [_vp, C.c_char_p, c_double_p, C.c_int64]),
    "zfvm_rate_of_change_av": (C.c_int, [_vp, c_double_p, c_double_p, c_double_p, c_double_p, C.c_double, C.c_int]),
    "zfvm_rate_of_change_av_device": (C.c_int, [_vp, _vp, _vp, _vp, _vp, C.c_double, C.c_int]),
    "zfvm_upload_avars": (C.c_int, [_vp, c_double_p]),
    "zfvm_download_avars": (C.c_int, [_vp, c_double_p]),
    "zfvm_avars_device": (_vp, [_vp]),
    "zfvm_set_frozen_bc_av": (C.c_int, [_vp, c_double_p, c_double_p]),
    "zfvm_rk_step_host_av": (C.c_int, [_vp, c_double_p, c_double_p, c_double_p, c_double_p, C.c_double, C.c_double]),
    "zfvm_halo_exchange_av": (C.c_int, [_vp, _vp, _vp]),
    "zfvm_nccl_unique_id": (C.c_int, [C.c_char_p]),
    "zfvm_comm_init": (C.c_int, [_vp, C.c_char_p, C.c_int, C.c_int]),
    "zfvm_set_halo": (C.c_int, [_vp, C.c_int64, C.c_int, C.POINTER(C.c_int), c_int64_p, c_int64_p, c_int64_p, c_int32_p]),
    "zfvm_halo_post": (C.c_int, [_vp, _vp, _vp]),
    "zfvm_halo_wait": (C.c_int, [_vp]),
    "zfvm_halo_exchange": (C.c_int, [_vp, _vp]),
    "zfvm_allreduce_min": (C.c_int, [_vp, c_double_p]),
}

for _name, (_res, _args) in _SIGNATURES.items():
    _fn = getattr(lib, _name)
    _fn.restype = _res
    _fn.argtypes = _args

DECLARED_SYMBOLS = tuple(_SIGNATURES)


def check(rc: int) -> None:
    """Non-zero status -> exception carrying ``zfvm_last_error()`` (the reference would LOG_ERR)."""
    if rc != 0:
        raise ZfvmError(lib.zfvm_last_error().decode("utf-8", "replace"))


_DTYPES = {0: np.float64, 1: np.int32, 2: np.int64, 3: np.uint8}


class _HandleView(np.ndarray):
    """ndarray view of library-owned storage that keeps the owning Python object (and so the handle) alive."""

    _owner = None


def named_array(getter, handle, name: str, owner=None) -> np.ndarray:
    """Zero-copy numpy view of a named array owned by a grid / stencil handle."""
    data = _vp()
    dtype = C.c_int()
    ndim = C.c_int()
    shape = (C.c_int64 * 4)()
    check(getter(handle, name.encode(), C.byref(data), C.byref(dtype), C.byref(ndim), shape))
    shp = tuple(int(shape[d]) for d in range(ndim.value))
    count = int(np.prod(shp)) if shp else 1
    np_dtype = np.dtype(_DTYPES[dtype.value])
    if count == 0 or not data.value:
        return np.zeros(shp, dtype=np_dtype)
    buf = (C.c_char * (count * np_dtype.itemsize)).from_address(data.value)
    out = np.frombuffer(buf, dtype=np_dtype).reshape(shp).view(_HandleView)
    out._owner = owner
    return out


def as_f64(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float64)


def ptr_f64(a: np.ndarray):
    return a.ctypes.data_as(c_double_p)
