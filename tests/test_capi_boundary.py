"""The C-ABI boundary: libzfvm_b200.so loads, exports every symbol include/zfvm.h declares, and fails loudly
(no CPU fallback) when no CUDA device is present.  No compute calls are made; no GPU is needed."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import zisafvm_b200 as z
from zisafvm_b200 import _capi, cases

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "zfvm.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(zfvm_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported():
    lib = C.CDLL(_capi.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 45
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_python_binding_covers_the_header():
    assert sorted(_capi.DECLARED_SYMBOLS) == declared_symbols()


def test_params_struct_layout_matches_header():
    """zfvm_params_default writes through the ctypes mirror of the struct: field offsets must agree."""
    p = _capi.ZfvmParams()
    _capi.lib.zfvm_params_default(C.byref(p))
    assert p.recon_mode == 0 and list(p.linear_weights)[:2] == [100.0, 1.0]
    assert p.epsilon == 1e-10 and p.exponent == 4.0 and p.gamma == 1.4 and p.steps_per_recompute == 1
    assert p.keep_polynomials == 0 and p.flux == 0 and p.scaling == 1
    # the fields behind flux_bc (appended later): defaults are "off"
    assert p.flux_bc == 0 and p.n_avars == 0 and p.heating_rate == 0.0 and p.heating_r0 == 0.0 and p.heating_r1 == 0.0
    assert p.recompute_threshold == 0.0
    p.n_avars, p.recompute_threshold = 3, 0.75  # and it is the last field: sizeof agrees with the header's layout
    assert C.sizeof(_capi.ZfvmParams) == _capi.ZfvmParams.recompute_threshold.offset + 8


def test_no_cpu_fallback_without_a_device():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    case = cases.isentropic_vortex(n=8)
    st = case.ensure_stencils()
    with pytest.raises(_capi.ZfvmError, match="no CUDA device|no CPU fallback|cuda"):
        z.CudaContext(case.grid, st, case.params)


def test_product_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing under zisafvm_b200/ may reference it."""
    pkg = os.path.join(ROOT, "zisafvm_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h", "Makefile")):
                text = open(os.path.join(dirpath, f), errors="replace").read()
                assert "oracle" not in text.lower(), os.path.join(dirpath, f)


def test_unknown_tableau_is_rejected_like_the_reference():
    """make_tableau LOG_ERRs 'Unknown Butcher Tableau. [name]' (runge_kutta.cpp:212); checked on the host-only path
    of the oracle here and on the device context in the GPU tests."""
    from oracle import binding as ob

    case = cases.isentropic_vortex(n=8)
    ora = ob.Oracle(case.grid, case.ensure_stencils(), case.params)
    with pytest.raises(ValueError, match="Unknown Butcher Tableau"):
        ora.rk_step("heun17", case.u0, 1e-3)


def test_scheme_parameters_reach_the_c_struct():
    """EulerParams.to_c(): the host mirror fills the zfvm_params fields of the rows that widen the path."""
    from zisafvm_b200.grid import WENO_PARAMS

    p = z.EulerParams(weno=WENO_PARAMS["3d_o3"], flux_bc="equilibrium", n_avars=2, heating=(0.3, 0.4, 0.8),
                      well_balancing="isentropic", gravity=z.Gravity(kind="point_mass", params=(-1.0, 1.0))).to_c()
    assert p.flux_bc == 2 and p.n_avars == 2 and p.well_balanced == 1 and p.gravity_kind == 2
    assert (p.heating_rate, p.heating_r0, p.heating_r1) == (0.3, 0.4, 0.8)
    q = z.EulerParams(weno=WENO_PARAMS["2d_o3"]).to_c()
    assert q.flux_bc == 0 and q.n_avars == 0 and q.heating_rate == 0.0
    assert q.steps_per_recompute == 1 and q.recompute_threshold == 0.0
    r = z.EulerParams(weno=WENO_PARAMS["2d_o3"], steps_per_recompute=4, recompute_threshold=1e-3).to_c()
    assert r.steps_per_recompute == 4 and r.recompute_threshold == 1e-3


def test_all_variables_carries_avars():
    a = z.AllVariables(4, n_avars=2)
    assert a.cvars.shape == (4, 5) and a.avars.shape == (4, 2)
    b = z.AllVariables(3, np.ones((3, 5)), np.arange(3.0))
    assert b.avars.shape == (3, 1) and b.avars.flags.c_contiguous
    assert z.AllVariables(3).avars.shape == (3, 0)


def test_no_malformed_ldgsts_in_the_tile_kernels():
    """ptxas 12.9 was seen to emit ``LDGSTS ... desc[UR1]`` (an odd uniform-register descriptor: 'illegal instruction'
    at run time, only on the path that executes it) in the remainder of a partially unrolled cp.async loop of
    recon_tile.cuh.  The loop is rolled now; this scans the built kernels so that the pattern cannot come back unseen."""
    import glob
    import re
    import shutil
    import subprocess

    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not available")
    objs = sorted(glob.glob(os.path.join(ROOT, "zisafvm_b200", "csrc", "build", "kernels", "recon_*d_deg*.o")))
    if not objs:
        pytest.skip("object files not present (library built elsewhere)")
    bad = re.compile(r"LDGSTS.*desc\[UR\d*[13579]\]")
    n_ldgsts = 0
    for o in objs:
        sass = subprocess.run(["cuobjdump", "-sass", o], capture_output=True, text=True).stdout
        n_ldgsts += sass.count("LDGSTS")
        assert not bad.search(sass), o
    assert n_ldgsts > 0   # the tile kernels stream their records with cp.async
