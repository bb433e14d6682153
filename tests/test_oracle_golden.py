"""Pins the CPU oracle (oracle/cts_ref.py) against the reference's own golden values, analytic
solutions and expected convergence orders.  CPU only."""
import math

import numpy as np
import pytest

import ref_problems as P
from ref_problems import R

EPS = np.finfo(np.float64).eps


# test/integration/single_column_ars.jl:308-315: literal ||u||_2 after one step (dt = 100 s, 1 Newton iter)
GOLDEN_NORMS = {
    "ARS111": 860.2745315698107,
    "ARS121": 860.2745315698107,
    "ARS122": 860.4393569534262,
    "ARS232": 860.452530117785,
    "ARS222": 860.452530117785,
}


def _single_column_norm(name):
    sc = P.SingleColumn()
    f = sc.ode_function()
    alg = R.imex_algorithm(name, R.NewtonsMethod(max_iters=1))
    u = sc.u0.copy()
    _, step = R.make_stepper(f, alg, u)
    R.solve(u, None, (0.0, 100.0), 100.0, step)
    return float(np.linalg.norm(u))


@pytest.mark.parametrize("name", sorted(GOLDEN_NORMS))
def test_single_column_golden_norms(name):
    # tolerance of the reference test: atol = 2e3 * eps()  (single_column_ars.jl:330)
    assert abs(_single_column_norm(name) - GOLDEN_NORMS[name]) <= 2e3 * EPS


def test_single_column_ars343_vs_in_test_comparator():
    # single_column_ars.jl:234-297,314: ARS343 is checked against an independent in-test implementation
    ref = P.SingleColumn().in_test_ars343_reference()
    assert abs(_single_column_norm("ARS343") - ref) <= 2e3 * EPS


# ---------------------------------------------------------------- convergence orders -----------
def _final_error(case, alg, dt, split=True):
    f = case.split_f if split else case.f
    u = case.Y0.copy()
    _, step = R.make_stepper(f, alg, u)
    t, u, _ = R.solve(u, None, (0.0, case.t_end), dt, step)
    return float(np.linalg.norm(u - case.analytic_sol(case.t_end)))


def _order(case, name, dts, **kw):
    errs = [_final_error(case, R.imex_algorithm(name, R.NewtonsMethod(max_iters=kw.pop("max_iters", 2)), **kw), dt)
            for dt in dts]
    return np.polyfit(np.log(dts), np.log(errs), 1)[0]


# expected orders: test/utils/convergence_utils.jl:20-49 (ARS343 -> 3, SSP333 -> 3, ARS222/232 -> 2, ARS111 -> 1)
@pytest.mark.parametrize("name,order", [("ARS111", 1), ("ARS122", 2), ("ARS222", 2), ("ARS232", 2),
                                        ("ARS233", 3), ("ARS343", 3), ("ARS443", 3),
                                        ("SSP222", 2), ("SSP332", 2), ("SSP333", 3), ("SSP433", 3)])
def test_expected_order_on_ark_analytic_sys(name, order):
    # BASELINE config 1 problem (perf/benchmark.jl:9-11): ark_analytic_sys, linear, all-implicit
    case = P.ark_analytic_sys()
    dts = [case.t_end / n for n in (10, 20, 40, 80)]
    got = _order(case, name, dts, max_iters=1)
    assert abs(got - order) < 0.35, (name, got)


@pytest.mark.parametrize("name,order", [("ARS343", 3), ("SSP333", 3), ("ARS232", 2)])
def test_expected_order_on_split_ark_analytic(name, order):
    case = P.ark_analytic()
    # λ = -100: stay in the asymptotic regime (λ dt <~ 0.3), where the reference's own test measures the order
    dts = [1 / n for n in (320, 640, 1280, 2560)]
    errs = []
    for dt in dts:
        alg = R.imex_algorithm(name, R.NewtonsMethod(max_iters=1))
        u = case.Y0.copy()
        _, step = R.make_stepper(case.split_f, alg, u)
        R.solve(u, None, (0.0, 2.0), dt, step)
        errs.append(abs(u[0] - math.atan(2.0)))
    got = np.polyfit(np.log(dts), np.log(errs), 1)[0]
    assert abs(got - order) < 0.4, (name, got, errs)


def test_config1_ars343_newton2_matches_analytic():
    # BASELINE config 1: IMEXAlgorithm(ARS343(), NewtonsMethod(; max_iters = 2)), dt = 0.01, 5 steps
    case = P.ark_analytic_sys()
    alg = R.imex_algorithm("ARS343", R.NewtonsMethod(max_iters=2))
    u = case.Y0.copy()
    _, step = R.make_stepper(case.split_f, alg, u)
    t, u, _ = R.solve(u, None, (0.0, case.t_end), 0.01, step)
    assert t == case.t_end
    err = np.abs(u - case.analytic_sol(case.t_end))
    assert np.all(err < 2e-3)
    # values recorded by the survey probe -- an independent restatement
    np.testing.assert_allclose(u, [0.23860154917387005, 0.26815540003824573, -0.4882555299804378], rtol=1e-12)


@pytest.mark.parametrize("name", R.SSP_NAMES)
def test_unconstrained_equals_ssp(name):
    # test/utils/convergence_utils.jl:228-238: both code paths agree to 100 eps (no limiter)
    case = P.ark_analytic()
    res = []
    for constraint in ("Unconstrained", "SSP"):
        alg = R.imex_algorithm(name, R.NewtonsMethod(max_iters=1), constraint)
        u = case.Y0.copy()
        _, step = R.make_stepper(case.split_f, alg, u)
        R.solve(u, None, (0.0, 1.0), 1 / 200, step)
        res.append(u.copy())
    assert np.linalg.norm(res[0] - res[1]) <= 100 * EPS * max(1.0, np.linalg.norm(res[0]))


def test_ssp333_is_not_bitwise_sdirk():
    # a22 != a33 in Float64 => cache γ === nothing => NewTimeStep update is rejected
    tab = R.tableau("SSP333")
    assert tab.a_imp[1, 1] != tab.a_imp[2, 2]
    case = P.ark_analytic()
    alg = R.imex_algorithm("SSP333", R.NewtonsMethod(update_j=R.UpdateEvery("NewTimeStep")))
    with pytest.raises(ValueError):
        R.make_stepper(case.split_f, alg, case.Y0.copy())


# ---------------------------------------------------------------- LSRK ------------------------
@pytest.mark.parametrize("prob", ["linear", "sincos"])
def test_lsrk54_order4(prob):
    # test/solvers/lsrk.jl:11-25: LSRK54CarpenterKennedy order 4 +- 0.05 (dts = 0.5 .^ (4:7))
    f, u0, tspan, p, sol = P.linear_prob() if prob == "linear" else P.sincos_prob()
    tab = R.lsrk54_tableau(np.float64)
    dts = [0.5 ** k for k in range(4, 8)]
    errs = []
    for dt in dts:
        u = u0.copy()
        du = np.zeros_like(u)
        R.solve(u, p, tspan, dt, lambda u_, p_, t, h: R.lsrk_step(u_, du, p_, t, h, f, tab))
        errs.append(np.linalg.norm(u - sol(u0, p, tspan[1])))
    got = np.polyfit(np.log(dts), np.log(errs), 1)[0]
    assert abs(got - 4) < 0.05, got


def test_lsrk54_tableau_consistency():
    A, B, C = R.lsrk54_tableau(np.float64)
    # c_{k+1} = c_k + B_k * (1 + A_{k+1} * (...)) -- the 2N consistency: C[1] = B[0]
    assert C[1] == B[0]
    assert A[0] == 0 and C[0] == 0
    # SURVEY Appendix B decimal values
    np.testing.assert_allclose(A, [0, -0.41789047449985195, -1.192151694642677, -1.6977846924715279,
                                   -1.5141834442571558], rtol=1e-15)
    np.testing.assert_allclose(B, [0.14965902199922912, 0.37921031299962726, 0.8229550293869817,
                                   0.6994504559491221, 0.15305724796815198], rtol=1e-15)


# ---------------------------------------------------------------- tableaux (test/unit/tableaus.jl) ---
@pytest.mark.parametrize("name", ["ARS111", "ARS121", "ARS122", "ARS233", "ARS232", "ARS222", "ARS343", "ARS443",
                                  "SSP222", "SSP322", "SSP332", "SSP333", "SSP433"])
def test_tableau_properties(name):
    tab = R.tableau(name)
    # test/unit/tableaus.jl:48-91: c = rowsum(a) (1e-13), sum(b) = 1
    np.testing.assert_allclose(tab.c_exp, tab.a_exp.sum(axis=1), atol=1e-13)
    np.testing.assert_allclose(tab.c_imp, tab.a_imp.sum(axis=1), atol=1e-13)
    assert abs(tab.b_exp.sum() - 1) < 1e-13 and abs(tab.b_imp.sum() - 1) < 1e-13


def test_ars343_float_values():
    # SURVEY Appendix B hex values (IEEE double, expression order of imex_tableaus.jl:293-323)
    tab = R.tableau("ARS343")
    assert float(tab.a_exp[1, 0]).hex() == "0x1.be53cb1d33509p-2"
    assert float(tab.b_exp[1]).hex() == "0x1.35600951895b4p+0"
    assert float(tab.b_exp[2]).hex() == "-0x1.49e9f831ac5eep-1"
    assert float(tab.a_exp[2, 0]).hex() == "0x1.48fd55118d956p-2"
    assert float(tab.a_exp[2, 1]).hex() == "0x1.962c907d0c12ap-2"
    assert float(tab.a_exp[3, 0]).hex() == "-0x1.b19877fa39610p-4"
    assert float(tab.a_imp[2, 1]).hex() == "0x1.20d61a716657cp-2"
    assert not tab.is_fsal()


# ---------------------------------------------------------------- Newton / GMRES / JFNK ----------
# test/solvers/newtons_method.jl is restated as written (functional forms, step adjustments, max_iters, checker, all four
# algorithms, EisenstatWalkerForcing, step-size variants, chord method) in tests/test_newton_reference_cpu.py.
