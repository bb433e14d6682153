"""ConvergenceChecker / ConvergenceCondition / LineSearch: the oracle restatement against the
reference's own known-answer tests (test/utils/convergence_checker.jl, test/unit/line_search.jl,
test/solvers/newtons_method.jl:58-91), and the host mirror's flattening of condition trees for the C ABI."""
import numpy as np
import pytest

from ref_problems import R

EPS = float(np.finfo(np.float64).eps)
val_func = lambda it: np.array([60.0, -80.0])
err_func1 = lambda it: np.array([6.0, -8.0]) * (-1 / 10) ** min(it, 11)
err_func2 = lambda it: np.array([60.0, -80.0]) * (-1 / 10) ** min(it, 12)
CONDITIONS = lambda m: (m.MaximumError(1e-10 + EPS), m.MaximumRelativeError(1e-10 + EPS),
                        m.MaximumErrorReduction(1e-10 + EPS), m.MinimumRateOfConvergence(0.1 + EPS, 1))
LAST1, LAST2 = (11, 9, 10, 12), (12, 10, 10, 13)   # test/utils/convergence_checker.jl:25-26


def _passes(checker, last1, last2):
    cache = checker.allocate_cache(val_func(0))
    for ef, last in ((err_func1, last1), (err_func2, last2)):
        for it in range(last):
            if R.is_converged(checker, cache, val_func(it), ef(it), it):
                return False
        if not R.is_converged(checker, cache, val_func(last), ef(last), last):
            return False
    return True


def test_single_conditions_hit_the_reference_iteration_numbers():
    for cond, a, b in zip(CONDITIONS(R), LAST1, LAST2):
        assert _passes(R.ConvergenceChecker(norm_condition=cond), a, b)
        assert _passes(R.ConvergenceChecker(component_condition=cond), a, b)


@pytest.mark.parametrize("combiner,pick", [(all, max), (any, min)])
def test_multiple_conditions(combiner, pick):
    cond = R.MultipleConditions(combiner, *CONDITIONS(R))
    assert _passes(R.ConvergenceChecker(norm_condition=cond), pick(LAST1), pick(LAST2))
    assert _passes(R.ConvergenceChecker(component_condition=cond), pick(LAST1), pick(LAST2))


@pytest.mark.parametrize("combiner,pick", [("&", max), ("|", min)])
def test_norm_and_component_conditions_combined(combiner, pick):
    allc, anyc = R.MultipleConditions(all, *CONDITIONS(R)), R.MultipleConditions(any, *CONDITIONS(R))
    for nc, cc in ((allc, anyc), (anyc, allc)):
        ch = R.ConvergenceChecker(norm_condition=nc, component_condition=cc, condition_combiner=combiner)
        assert _passes(ch, pick(LAST1), pick(LAST2))


def test_checker_without_conditions_is_an_error():
    with pytest.raises(ValueError):
        R.ConvergenceChecker().allocate_cache(np.zeros(2))


# ---- line search (test/unit/line_search.jl) ----------------------------------------------------
def test_line_search_accepts_a_full_step_that_reduces_the_residual():
    A, b = np.array([[2.0, 0.0], [0.0, 3.0]]), np.array([4.0, 9.0])
    f_ = lambda f, x: f.__setitem__(Ellipsis, A @ x - b)
    x_old = np.array([1.0, 1.0])
    f = np.zeros(2)
    f_(f, x_old)
    dx = np.linalg.solve(A, f)
    x = x_old - dx
    R.line_search(R.LineSearch(), x, dx, f, f_, None)
    np.testing.assert_allclose(x, np.linalg.solve(A, b), atol=1e-12)


def test_line_search_backtracking_and_nan():
    f_ = lambda f, x: f.__setitem__(Ellipsis, x ** 3 - 1)
    x_old = np.array([5.0])
    f = np.zeros(1)
    f_(f, x_old)
    n0 = np.linalg.norm(f)
    dx = f / (3 * x_old ** 2)
    x = x_old - dx
    R.line_search(R.LineSearch(), x, dx, f, f_, None)
    g = np.zeros(1)
    f_(g, x)
    assert np.linalg.norm(g) <= n0 + EPS
    # NaN for negative x: five halvings walk back into the valid region
    f2_ = lambda f, x: f.__setitem__(Ellipsis, x[0] ** 2 - 1 if x[0] >= 0 else np.nan)
    x = np.array([2.0]) - np.array([5.0])
    f = np.array([3.0])
    R.line_search(R.LineSearch(), x, np.array([5.0]), f, f2_, None)
    assert x[0] >= 0
    g = np.zeros(1)
    f2_(g, x)
    assert np.linalg.norm(g) <= 3.0
    # no-op when line_search is nothing
    x = np.array([1.0, 2.0])
    R.line_search(None, x, np.array([0.1, 0.2]), np.zeros(2), lambda f, x: None, None)
    np.testing.assert_array_equal(x, [1.0, 2.0])


# ---- Newton with a checker / line search (test/solvers/newtons_method.jl:58-91; tolerance, not data) --------------
def _equation(linear, FT, n, seed=1):
    rng = np.random.default_rng(seed)
    A = rng.random((n, n)).astype(FT)
    b = rng.random(n).astype(FT)
    if linear:
        return (lambda f, x: f.__setitem__(Ellipsis, A @ x - b)), (lambda j, x: j.M.__setitem__(Ellipsis, A)), np.linalg.solve(A, b)
    f_ = lambda f, x: f.__setitem__(Ellipsis, A @ np.sinh(A @ x - b) - b)
    j_ = lambda j, x: j.M.__setitem__(Ellipsis, A @ np.diag(np.cosh(A @ x - b)) @ A)
    return f_, j_, np.linalg.solve(A, np.arcsinh(np.linalg.solve(A, b)) + b)


@pytest.mark.parametrize("linear,n", [(True, 1), (True, 10), (False, 1), (False, 10)])
@pytest.mark.parametrize("use_line_search", [False, True])
def test_newton_with_convergence_checker(linear, n, use_line_search):
    FT = np.float64
    rtol = 100 * EPS
    f_, j_, x_exact = _equation(linear, FT, n)
    alg = R.NewtonsMethod(max_iters=1 if linear else 10,
                          convergence_checker=R.ConvergenceChecker(norm_condition=R.MaximumRelativeError(rtol)),
                          line_search=R.LineSearch() if use_line_search else None)
    x = np.zeros(n, FT)
    cache = R.allocate_newton_cache(alg, x, R.DenseJacobian(n, FT))
    R.solve_newton(alg, cache, x, f_, j_)
    assert np.linalg.norm(x - x_exact) / np.linalg.norm(x_exact) < (rtol if linear else 1e3 * rtol)
    if not linear:
        assert cache["newton_iters"] < 10   # the checker stopped the iteration early


def test_host_mirror_flattens_condition_trees():
    import cts_b200 as cts
    from cts_b200 import _lib as L
    ch = cts.ConvergenceChecker(norm_condition=cts.MultipleConditions(any, cts.MaximumError(1e-3), cts.MinimumRateOfConvergence(0.5, 2)),
                                component_condition=cts.MaximumRelativeError(1e-8), condition_combiner="|")
    c = ch.to_c()
    assert c.n_norm == 3 and c.n_comp == 1 and c.combiner_or == 1
    assert c.norm_cond[2].kind == L.COND_ANY and c.norm_cond[2].nchildren == 2
    kinds = sorted(c.norm_cond[k].kind for k in range(2))
    assert kinds == [L.COND_MAX_ERROR, L.COND_MIN_RATE]
    with pytest.raises(ValueError):
        cts.ConvergenceChecker().to_c()
    nm = cts.NewtonsMethod(max_iters=3, convergence_checker=ch, line_search=cts.LineSearch())
    assert nm.to_c([]).max_iters == 3   # no longer NotImplemented
