"""Rosenbrock: the numpy oracle against the reference's own test expectations
(test/solvers/rosenbrock.jl, test/unit/tableaus.jl:142-160, test/unit/cache_sizes.jl:46-54) and the host tableau."""
import math

import numpy as np
import pytest

from ref_problems import R, ark_analytic, ark_analytic_nonlin, ark_analytic_sys, stiff_linear


def test_sspknoth_tableau_structure():
    tab = R.ssp_knoth()
    up = lambda M: np.triu(M)  # UpperTriangular includes the diagonal
    assert np.all(up(tab.A) == 0) and np.all(up(tab.alpha) == 0)
    assert np.all(np.triu(tab.Gamma, 1) == 0) and np.all(np.diag(tab.Gamma) != 0)
    assert np.all(np.abs(up(tab.C)) < 1e-14)
    # transformed coefficients of the 3-stage tableau are dyadic except m
    np.testing.assert_array_equal(tab.A, [[0, 0, 0], [1, 0, 0], [0.25, 0.25, 0]])
    np.testing.assert_array_equal(tab.C, [[0, 0, 0], [0, 0, 0], [-0.75, -0.75, 0]])
    np.testing.assert_allclose(tab.m, [2 / 3, 2 / 3, 2 / 3], rtol=2e-16)


def test_host_tableau_equals_oracle_tableau():
    import cts_b200 as cts  # the library loads without a GPU; no compute call is made here
    t, o = cts.tableau(cts.SSPKnoth()), R.ssp_knoth()
    assert isinstance(cts.RosenbrockAlgorithm(t).tableau, cts.RosenbrockTableau)
    c = t.to_c()
    assert c.s == 3 and c.Gamma[6] == -0.75 and c.m[2] == o.m[2]
    for name in ("A", "alpha", "C", "Gamma", "m"):
        np.testing.assert_array_equal(getattr(t, name), getattr(o, name))


@pytest.mark.parametrize("case,n0,split", [(ark_analytic_sys, 200, True), (ark_analytic_nonlin, 450, False),
                                           (stiff_linear, 1000, True), (ark_analytic, 6000, True)])
def test_sspknoth_converges_at_second_order(case, n0, split):
    # test/solvers/rosenbrock.jl:15-52 with imex_convergence_orders(::SSPKnoth) = (2, 3, 2) (convergence_utils.jl:49):
    # order 2 on the split problems
    c = case()
    errs = []
    for n in (n0, 2 * n0, 4 * n0):
        u = c.Y0.copy()
        _, step = R.make_stepper(c.split_f if split else c.f, R.RosenbrockAlgorithm(R.ssp_knoth()), u)
        _, u, _ = R.solve(u, None, (0.0, c.t_end), c.t_end / n, step)
        errs.append(float(np.linalg.norm(u - c.analytic_sol(c.t_end))))
    orders = [math.log2(errs[i] / errs[i + 1]) for i in range(2)]
    # (ark_analytic is still pre-asymptotic at these step counts -- the reference uses 25000 steps -- hence the margin)
    assert all(abs(o - 2.0) < 0.3 for o in orders), (errs, orders)


def test_rosenbrock_cache_sizes_and_stage_count():
    c = ark_analytic_sys()
    cache = R.init_cache_rosenbrock(c.Y0.copy(), c.split_f, R.RosenbrockAlgorithm(R.ssp_knoth()))
    assert len(cache["k"]) == 3
    for name in ("U", "fU", "fU_imp", "fU_exp", "fU_lim", "dYdt"):
        assert np.all(cache[name] == 0) and cache[name].shape == c.Y0.shape
    assert cache["W"] is c.split_f.T_imp.jac_prototype  # the prototype itself, not a copy (rosenbrock.jl:117)


def test_explicit_only_rosenbrock_is_an_explicit_rk():
    # without T_imp! the stages are k_i = dtγ f(U_i) (rosenbrock.jl:225-227): for SSPKnoth that is SSP33ShuOsher
    lam = -0.7
    f = R.ClimaODEFunction(T_exp=lambda du, u, p, t: du.__setitem__(Ellipsis, lam * u + np.sin(t)))
    u = np.array([1.0, -0.5])
    _, step = R.make_stepper(f, R.RosenbrockAlgorithm(R.ssp_knoth()), u)
    v = u.copy()
    for k in range(5):
        step(u, None, 0.1 * k, 0.1)
    # reference SSPRK3 by hand
    t = 0.0
    g = lambda y, s: lam * y + np.sin(s)
    for k in range(5):
        y1 = v + 0.1 * g(v, t)
        y2 = 0.75 * v + 0.25 * (y1 + 0.1 * g(y1, t + 0.1))
        v = v / 3 + 2 / 3 * (y2 + 0.1 * g(y2, t + 0.05))
        t += 0.1
    np.testing.assert_allclose(u, v, rtol=1e-13)
