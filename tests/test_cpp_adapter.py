"""The ZisaFVM-side adapter (include/zfvm_zisa_adapter.hpp) compiled against a stand-in of the reference's interfaces
(tests/cpp/mock_zisa.hpp) and driven from a plain C++ program (tests/cpp/adapter_driver.cpp): the C ABI exercised
without Python in between, through the classes a maintainer would register (RateOfChange, TimeIntegration, CFLCondition,
SanityCheck, BoundaryCondition; rate_of_change.hpp:23-43, time_integration.hpp:42-44).

CPU: the adapter and the header compile (C++17 / C99), unresolved symbols would fail the link, and a failing C-ABI call
arrives as LOG_ERR.  GPU: the C++ program's results against the oracle.
"""
import os
import struct
import subprocess

import numpy as np
import pytest

import zisafvm_b200 as z
from zisafvm_b200 import _capi, cases

from util import active_vars, rel_err, tendency_scales

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CPP = os.path.join(ROOT, "tests", "cpp")
BIN = os.path.join(CPP, "build", "adapter_driver")


@pytest.fixture(scope="module")
def driver():
    os.makedirs(os.path.dirname(BIN), exist_ok=True)
    libdir = os.path.dirname(_capi.LIB_PATH)
    cmd = ["g++", "-std=c++17", "-O2", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"), "-I", CPP,
           os.path.join(CPP, "adapter_driver.cpp"), "-o", BIN, "-L", libdir, "-lzfvm_b200", f"-Wl,-rpath,{libdir}"]
    subprocess.run(cmd, check=True, capture_output=True, text=True)
    return BIN


def test_header_is_plain_c():
    """include/zfvm.h is a C header: plain pointers and sizes, no C++ in the signatures."""
    r = subprocess.run(["gcc", "-std=c99", "-pedantic", "-Wall", "-Werror", "-fsyntax-only", "-x", "c",
                        os.path.join(ROOT, "include", "zfvm.h")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_adapter_compiles_links_and_reports_errors(driver):
    """Every C-ABI symbol the adapter uses resolves against the built library, and a failing call (no CUDA device on the
    CPU box) reaches the caller through the reference's LOG_ERR path with the library's message."""
    r = subprocess.run([driver, "nodevice"], capture_output=True, text=True)
    assert r.returncode == 0, (r.stdout, r.stderr)
    from conftest import _cuda_device_count

    if _cuda_device_count() == 0:
        assert "LOG_ERR: zfvm_create: no CUDA device" in r.stdout


@pytest.mark.gpu
def test_cpp_time_loop_matches_oracle(driver, tmp_path):
    """Sum[Zero, CudaEulerRateOfChange] and the reference-shaped time loop over CudaRungeKutta / CudaCFL / CudaSanityCheck /
    CudaFrozenBC, run by the C++ program, against the oracle: residual <= 1e-12 of the flux scale, state after 8 SSP3
    steps <= 1e-11 per variable, the CFL time steps <= 1e-11; the families imported with zfvm_stencils_from_arrays give
    the bit-identical residual; resident and host-refreshed stepping agree bit for bit."""
    from oracle.binding import Oracle

    case = cases.isentropic_vortex(n=24, order=3)
    g = case.grid
    n, n_steps = g.n_cells, 8
    inp, outp = tmp_path / "in.bin", tmp_path / "out.bin"
    with open(inp, "wb") as f:
        f.write(struct.pack("<5q", 2, g.n_vertices, n, n_steps, 3))
        f.write(struct.pack("<2d", case.params.gamma, case.cfl))
        f.write(np.ascontiguousarray(g.array("vertices"), dtype=np.float64).tobytes())
        f.write(np.ascontiguousarray(g.array("vertex_indices"), dtype=np.int32).tobytes())
        f.write(np.ascontiguousarray(g.array("cell_flags"), dtype=np.uint8).tobytes())
        f.write(np.ascontiguousarray(case.u0, dtype=np.float64).tobytes())
    r = subprocess.run([driver, "run", str(inp), str(outp)], capture_output=True, text=True)
    assert r.returncode == 0, (r.stdout, r.stderr)
    raw = np.fromfile(outp, dtype=np.float64)
    assert raw.size == 4 * n * 5 + n_steps
    tend, tend_imp, u_host = (raw[k * n * 5:(k + 1) * n * 5].reshape(n, 5) for k in range(3))
    dts = raw[3 * n * 5:3 * n * 5 + n_steps]
    u_res = raw[3 * n * 5 + n_steps:].reshape(n, 5)

    st = case.ensure_stencils()
    ora = Oracle(g, st, case.params, None)
    ref = ora.rate_of_change(case.u0)
    scale = tendency_scales(case.u0, case.params.gamma, g.array("inradii"))
    assert (np.abs(tend - ref).max(axis=0) / scale).max() < 1e-12
    assert np.array_equal(tend, tend_imp)
    ora.set_frozen_bc(case.u0)
    u_ref = case.u0.copy()
    for s in range(n_steps):
        dt_ref = ora.cfl_dt(u_ref, case.cfl)
        assert abs(dts[s] - dt_ref) <= 1e-11 * dt_ref
        u_ref = ora.rk_step("ssp3", u_ref, dt_ref)
    vs = active_vars(2)
    assert rel_err(u_host, u_ref)[vs].max() < 1e-11, rel_err(u_host, u_ref)
    assert np.array_equal(u_host, u_res)
    assert np.array_equal(u_host[g.is_ghost], case.u0[g.is_ghost])
