"""test/solvers/newtons_method.jl restated on the numpy oracle (see newton_reference_cases.py for what is kept as
written and what cannot be: the Julia RNG stream).  Every assertion uses the reference's own tolerance."""
import numpy as np
import pytest

from newton_reference_cases import ROWS, SEEDS, Equation, main_algs
from ref_problems import R


def _solve(alg, use_j, eq):
    x = eq.x_init.copy()
    cache = R.allocate_newton_cache(alg, x, R.DenseJacobian(eq.n, eq.FT) if use_j else None)

    def f_(f, xx):
        f[...] = eq.f(xx)

    def j_(j, xx):
        j.M[...] = eq.j(xx)

    with np.errstate(all="ignore"):
        R.solve_newton(alg, cache, x, f_, j_ if use_j else None)
    return x, cache


@pytest.mark.parametrize("seed", SEEDS)
@pytest.mark.parametrize("row", ROWS, ids=lambda r: f"{'lin' if r[0] else 'nonlin'}-{np.dtype(r[1]).name}-n{r[2]}")
def test_newtons_method_all_four_algorithms(row, seed):
    # :51-91 -- direct, GMRES with mul!, JFNK with the row's step_adjustment, direct + LineSearch
    is_linear, FT, n, adj = row
    eq = Equation("linear" if is_linear else "nonlinear", FT, n, seed)
    rtol, algs = main_algs(R, FT, is_linear, adj)
    for k, (alg, use_j) in enumerate(algs):
        x, _ = _solve(alg, use_j, eq)
        assert eq.rel_err(x) < rtol, (k + 1, eq.rel_err(x) / rtol)


@pytest.mark.parametrize("seed", SEEDS)
@pytest.mark.parametrize("n", [1, 3, 10])
def test_eisenstat_walker_forcing(n, seed):
    # :93-115 -- max_iters = 30, GMRES memory 30, default EisenstatWalkerForcing()
    FT = np.float64
    eq = Equation("nonlinear", FT, n, seed)
    rtol = 100 * float(np.finfo(FT).eps)
    alg = R.NewtonsMethod(max_iters=30, krylov_method=R.KrylovMethod(forcing_term=R.EisenstatWalkerForcing(), memory=30),
                          convergence_checker=R.ConvergenceChecker(norm_condition=R.MaximumRelativeError(rtol)))
    x, cache = _solve(alg, True, eq)
    assert eq.rel_err(x) < rtol
    # the forcing term really varied: first rtol is initial_rtol = 0.5, later ones are tighter
    assert cache["forcing"]["prev_rtol"] <= 0.9 and cache["newton_iters"] >= 2


@pytest.mark.parametrize("seed", SEEDS)
@pytest.mark.parametrize("step", [1, 2, 3])
def test_forward_diff_step_size_variants(step, seed):
    # :117-141 (StepSize2 and StepSize3 upstream; StepSize1, newtons_method.jl:107-109, added with the same criterion)
    FT, n = np.float64, 5
    eq = Equation("gentle", FT, n, seed)
    rtol = 100 * float(np.finfo(FT).eps)
    alg = R.NewtonsMethod(max_iters=30, krylov_method=R.KrylovMethod(
        jacobian_free_jvp=R.ForwardDiffJVP(default_step=step, step_adjustment=1.0), forcing_term=R.ConstantForcing(rtol)),
        convergence_checker=R.ConvergenceChecker(norm_condition=R.MaximumRelativeError(rtol)))
    x, _ = _solve(alg, False, eq)
    assert eq.rel_err(x) < 0.01


@pytest.mark.parametrize("seed", SEEDS)
def test_chord_method_jacobian_reuse(seed):
    # :143-175 -- UpdateEvery(NewNewtonSolve): exact in one iteration on the linear problem, converging on the gentle one
    FT, n = np.float64, 5
    rtol = 100 * float(np.finfo(FT).eps)
    cc = lambda: R.ConvergenceChecker(norm_condition=R.MaximumRelativeError(rtol))
    eq = Equation("linear", FT, n, seed)
    x, _ = _solve(R.NewtonsMethod(max_iters=1, update_j=R.UpdateEvery("NewNewtonSolve"), convergence_checker=cc()), True, eq)
    assert eq.rel_err(x) < rtol
    eq = Equation("gentle", FT, n, seed)
    x, cache = _solve(R.NewtonsMethod(max_iters=30, update_j=R.UpdateEvery("NewNewtonSolve"), convergence_checker=cc()), True, eq)
    assert eq.rel_err(x) < 0.01
    assert 2 <= cache["newton_iters"] < 30   # a frozen Jacobian needs several iterations, the checker ends them
