"""Pins the CPU oracle (oracle/zisa_oracle.cpp) against the reference's own known-answer tests.

The reference cannot be compiled here (DESIGN.md "Oracle"), and it ships no end-to-end golden vector for the
residual path: its own unit tests are the only fixtures that exist.  Every test below restates
one of them (file:line under /root/reference/test/zisa/unit_test) with the same numbers and the same tolerance,
evaluated by the oracle's restatement of the function under test.  No GPU is needed.
"""
import ctypes as C
import math

import numpy as np
import pytest

from oracle import binding as ob

dp = ob.dp
RHS = C.CFUNCTYPE(None, C.c_double, dp, dp, C.c_int64)


def _arr(x):
    return np.ascontiguousarray(x, dtype=np.float64)


def _p(a):
    return a.ctypes.data_as(dp)


@pytest.fixture(scope="module")
def L():
    lib = ob.lib()
    lib.oracle_poly_dof.restype = C.c_int
    lib.oracle_poly_index2.restype = C.c_int
    lib.oracle_poly_index3.restype = C.c_int
    lib.oracle_rk_generic.restype = C.c_int
    lib.oracle_rk_generic.argtypes = [C.c_char_p, RHS, dp, C.c_int64, C.c_double, C.c_double]
    return lib


# ---- flux/hllc.cpp:10-26 -------------------------------------------------------------------------------------
def test_hllc_consistency(L):
    """HLLC(u, u) == F(u) to 1e-12 for gamma = 1.6, u = (1, -0.2, 0.3, 0.8, 12)."""
    u = _arr([1.0, -0.2, 0.3, 0.8, 12.0])
    nf, pf = np.zeros(5), np.zeros(5)
    L.oracle_hllc(1.6, _p(u), _p(u), _p(nf))
    L.oracle_euler_flux(1.6, _p(u), _p(pf))
    assert np.abs(nf - pf).max() < 1e-12
    # Euler::flux itself (model/euler_impl.hpp:23-36), from the definition
    p = (1.6 - 1.0) * (12.0 - 0.5 * (0.04 + 0.09 + 0.64) / 1.0)
    v = -0.2
    assert np.allclose(pf, [-0.2, v * -0.2 + p, v * 0.3, v * 0.8, v * (12.0 + p)], rtol=0, atol=1e-15)


def test_hllc_supersonic_upwinding(L):
    """Not a reference test: in supersonic flow HLLC must return the upwind physical flux (s_L >= 0 / s_R < 0)."""
    gamma = 1.4
    uL = _arr([1.0, 3.0, 0.1, 0.0, 1.0 / 0.4 + 0.5 * 9.01])
    uR = _arr([0.9, 2.8, 0.0, 0.1, 0.9 / 0.4 + 0.5 * (2.8 ** 2 + 0.01) / 0.9])
    nf, pf = np.zeros(5), np.zeros(5)
    L.oracle_hllc(gamma, _p(uL), _p(uR), _p(nf))
    L.oracle_euler_flux(gamma, _p(uL), _p(pf))
    assert np.abs(nf - pf).max() < 1e-13
    # mirrored state: flow to the left, the right state is upwind
    mL, mR = uR.copy(), uL.copy()
    mL[1] *= -1
    mR[1] *= -1
    L.oracle_hllc(gamma, _p(mL), _p(mR), _p(nf))
    L.oracle_euler_flux(gamma, _p(mR), _p(pf))
    assert np.abs(nf - pf).max() < 1e-13


def test_rusanov_consistency(L):
    """Rusanov is not in the reference; pinned only by consistency F(u, u) = F(u)."""
    u = _arr([1.0, -0.2, 0.3, 0.8, 12.0])
    nf, pf = np.zeros(5), np.zeros(5)
    L.oracle_rusanov(1.6, _p(u), _p(u), _p(nf))
    L.oracle_euler_flux(1.6, _p(u), _p(pf))
    assert np.abs(nf - pf).max() < 1e-14


# ---- math/poly2d.cpp:10-64 -------------------------------------------------------------------------------------
def test_poly_dof_tables(L):
    assert [L.oracle_poly_dof(d, 2) for d in range(5)] == [1, 3, 6, 10, 15]
    assert [L.oracle_poly_dof(d, 3) for d in range(5)] == [1, 4, 10, 20, 35]


def test_poly_index_enumerates_in_iterator_order(L):
    """PolyIndexRange<2>/<3>(3): degree-major order; within a degree, a descending (poly2d_decl.hpp iterator)."""
    count = 0
    for deg in range(4):
        for b in range(deg + 1):
            assert L.oracle_poly_index2(deg - b, b) == count
            count += 1
    assert count == 10
    count = 0
    for deg in range(4):
        for a in range(deg, -1, -1):
            for b in range(deg - a, -1, -1):
                assert L.oracle_poly_index3(a, b, deg - a - b) == count
                count += 1
    assert count == 20


# ---- math/poly2d.cpp:94-197 ------------------------------------------------------------------------------------
def _eval(L, nd, deg, nv, coeffs, moments, x):
    out = np.zeros(nv)
    c, m = _arr(coeffs), _arr(moments)
    xc = np.zeros(3)
    L.oracle_poly_eval(nd, deg, nv, _p(c), _p(m), m.size, _p(xc), 1.0, _p(_arr(x)), _p(out))
    return out


def test_poly2d_examples(L):
    pc, pm = [1.0, 2.0, 3.0, 4.0, 5.0, 6.0], [0.0, 0.0, 0.0, 1.0, 2.0, 3.0]
    assert abs(_eval(L, 2, 2, 1, pc, pm, [0, 0, 0])[0] - (1.0 - 4.0 - 10.0 - 18.0)) < 1e-14
    x, y = -3.0, 2.0
    exact = 1.0 + 2.0 * x + 3.0 * y + 4.0 * (x * x - 1.0) + 5.0 * (x * y - 2.0) + 6.0 * (y * y - 3.0)
    assert abs(_eval(L, 2, 2, 1, pc, pm, [x, y, 0])[0] - exact) < 1e-14
    # saxpy-like: Poly(0.2 p + q - 0.4 p)(x) == 0.2 p(x) + q(x) - 0.4 p(x)
    qc, qm = [1.0, 2.0, 3.0], [0.0, 0.0, 0.0]
    xx = [-3.4, 2.138, 0.0]
    out = np.zeros(1)
    L.oracle_poly_saxpy(2, 2, _p(_arr(pc)), _p(_arr(pm)), 1, _p(_arr(qc)), _p(_arr(qm)), _p(_arr(xx)), _p(out))
    px, qx = _eval(L, 2, 2, 1, pc, pm, xx)[0], _eval(L, 2, 1, 1, qc, qm, xx)[0]
    assert abs(out[0] - (0.2 * px + qx - 0.4 * px)) < 1e-14


# ---- math/poly2d.cpp:199-283 -----------------------------------------------------------------------------------
def test_poly3d_two_variables(L):
    mom = [0.0, 0.0, 0.0, 0.0, 1.0, 2.0, 3.0, 4.0, 5.0, 6.0]
    P = np.array([[1, -1], [2, -1], [-1, -1], [-3, -1], [2, -3], [4, -2], [2, -5], [1, -3], [1, -3], [3, -1]], dtype=float)
    x, y, z = -3.0, 2.0, 0.2
    exact = (P[0] + P[1] * x + P[2] * y + P[3] * z + P[4] * (x * x - 1.0) + P[5] * (x * y - 2.0) + P[6] * (x * z - 3.0)
             + P[7] * (y * y - 4.0) + P[8] * (y * z - 5.0) + P[9] * (z * z - 6.0))
    approx = _eval(L, 3, 2, 2, P.ravel(), mom, [x, y, z])
    assert np.abs(approx - exact).max() < 1e-14


# ---- model/eos.cpp:7-23, model/generic_eos_test.hpp:58-100 ----------------------------------------------------------
def test_ideal_gas_round_trips(L):
    gamma = 1.2
    h, K = C.c_double(), C.c_double()
    L.oracle_eos_rhoE_to_hK(gamma, 1.0, 2.0, C.byref(h), C.byref(K))
    rho, E = C.c_double(), C.c_double()
    L.oracle_eos_hK_to_rhoE(gamma, h.value, K.value, C.byref(rho), C.byref(E))
    assert abs(rho.value - 1.0) < 1e-10 and abs(E.value - 2.0) < 1e-10
    # closed forms (ideal_gas_eos.hpp:162-215): p = (gamma-1) E, h = gamma/(gamma-1) p/rho, K = p / rho^gamma
    assert abs(h.value - gamma / (gamma - 1) * (gamma - 1) * 2.0) < 1e-14
    assert abs(K.value - (gamma - 1) * 2.0) < 1e-14


# ---- model/local_equilibrium.cpp:15-60 -------------------------------------------------------------------------------
def _triangle_rule_deg2():
    """Cell Gauss points/weights of one triangle through the host library (TriangularRule deg 2, make_cell)."""
    import zisafvm_b200 as z

    def cell(tri):
        g = z.Grid(2, np.array(tri, dtype=float), np.array([[0, 1, 2]], dtype=np.int32), z.QRDegrees(1, 2, 2))
        return g.array("cell_qp")[0].copy(), g.array("cell_qw")[0].copy(), float(g.array("volumes")[0])

    return cell


def test_local_equilibrium_solve_and_extrapolate(L):
    gamma, grav = 1.2, 0.9  # IdealGasEOS(1.2, 0.9), ConstantGravityRadial(0.9): phi = g * |x|
    phi = lambda x: grav * np.linalg.norm(x, axis=-1)
    x_ref, h_ref, K_ref = np.array([0.5, 0.6, 0.0]), 10.0, 3.0

    def rhoE_eq(x):
        h = h_ref + phi(x_ref) - phi(x)
        rho = ((gamma - 1.0) / (gamma * K_ref) * h) ** (1.0 / (gamma - 1.0))
        return np.stack([rho, K_ref * rho ** gamma / (gamma - 1.0)], axis=-1)

    cell = _triangle_rule_deg2()
    qp, qw, vol = cell([[1.0, 1.0, 0.0], [1.01, 1.0, 0.0], [1.0, 1.01, 0.0]])
    bar = (qw[:, None] * rhoE_eq(qp)).sum(axis=0) / vol
    h, K, pref = C.c_double(), C.c_double(), C.c_double()
    phis = _arr(phi(qp))
    found = L.oracle_local_equilibrium(gamma, qp.shape[0], _p(phis), _p(_arr(qw)), vol, bar[0], bar[1], C.byref(h),
                                       C.byref(K), C.byref(pref))
    assert found == 1
    assert pref.value == phis[0]  # x_ref of the local equilibrium is the first cell Gauss point

    def extrap(x):
        hh = h.value + pref.value - phi(x)
        rho = ((gamma - 1.0) / (gamma * K.value) * hh) ** (1.0 / (gamma - 1.0))
        return np.stack([rho, K.value * rho ** gamma / (gamma - 1.0)], axis=-1)

    xy = np.array([1.1, 2.1, 0.0])
    assert np.abs(extrap(xy) - rhoE_eq(xy)).max() < 1e-9              # "extrapolate to point"
    qp2, qw2, vol2 = cell([[1.2, 1.1, 0.0], [1.21, 1.1, 0.0], [1.2, 1.11, 0.0]])
    a = (qw2[:, None] * extrap(qp2)).sum(axis=0) / vol2
    b = (qw2[:, None] * rhoE_eq(qp2)).sum(axis=0) / vol2
    assert np.abs(a - b).max() < 1e-9                                   # "extrapolate to triangle"


# ---- ode/runge_kutta.cpp:172-199 -----------------------------------------------------------------------------------------


def _rk_error(L, ode, method, dt):
    fns = {
        "constant": (lambda t, u: 0.1 + 0 * u, lambda t, u0: u0 + 0.1 * t),
        "exp": (lambda t, u: u, lambda t, u0: u0 * math.exp(t)),
        "t_square": (lambda t, u: t * t + 0 * u, lambda t, u0: u0 + t ** 3 / 3.0),
    }
    f, sol = fns[ode]

    @RHS
    def rhs(t, u, dudt, n):
        uu = np.ctypeslib.as_array(u, shape=(n,))
        np.ctypeslib.as_array(dudt, shape=(n,))[:] = f(t, uu)

    u = np.ones(30 * 5)  # AllVariablesDimensions{30, 2, 3}
    t, t_final = 0.0, 3.0
    while t < t_final - 0.5 * dt:
        assert L.oracle_rk_generic(method.encode(), rhs, _p(u), u.size, t, dt) == 0
        t += dt
    return np.abs(u - sol(t_final, 1.0)).max()


EXPERIMENTS = [("forward_euler", (0.9, 1.1), (1e-2, 1e-3)), ("ssp2", (1.9, 2.1), (1e-2, 1e-3)),
               ("ssp3", (2.9, 3.1), (1e-1, 1e-2)), ("wicker", (1.9, 3.1), (1e-1, 1e-2)), ("rk4", (3.9, 4.1), (1e-1, 1e-2)),
               ("fehlberg", (4.9, 5.1), (1e-1, 1e-2))]


@pytest.mark.parametrize("method,rates,dts", EXPERIMENTS)
def test_runge_kutta_exact_for_constant_rhs(L, method, rates, dts):
    assert _rk_error(L, "constant", method, 0.01) < 1e-12


@pytest.mark.parametrize("ode", ["exp", "t_square"])
@pytest.mark.parametrize("method,rates,dts", EXPERIMENTS)
def test_runge_kutta_convergence_rate(L, ode, method, rates, dts):
    """Observed order inside the reference's interval (the reference uses dt = 1e-3/1e-4 for the two low-order
    schemes; 1e-2/1e-3 here keeps the Python callback count small, the asymptotic rate is the same)."""
    coarse, fine = _rk_error(L, ode, method, dts[0]), _rk_error(L, ode, method, dts[1])
    if fine > 1e-12:
        rate = (math.log(fine) - math.log(coarse)) / (math.log(dts[1]) - math.log(dts[0]))
        assert rates[0] <= rate <= rates[1], (method, ode, rate)


def test_unknown_tableau_is_an_error(L):
    u = np.ones(4)

    @RHS
    def rhs(t, u_, d, n):
        pass

    assert L.oracle_rk_generic(b"heun17", rhs, _p(u), 4, 0.0, 0.1) == 1


# ---- Eigen LDLT boundary (lsq_solver.cpp:47,82): restated solver against numpy's normal-equation solution -----------
def test_ldlt_normal_equations_match_lstsq(L):
    rng = np.random.default_rng(3)
    for rows, cols in [(3, 2), (10, 5), (18, 9), (57, 19)]:
        A = rng.normal(size=(rows, cols))
        rhs = rng.normal(size=(rows, 5))
        x = np.zeros((cols, 5))
        L.oracle_lsq_solve(_p(_arr(A)), rows, cols, _p(_arr(rhs)), 5, _p(x))
        ref = np.linalg.lstsq(A, rhs, rcond=None)[0]
        assert np.abs(x - ref).max() < 1e-11 * max(1.0, np.abs(ref).max())
