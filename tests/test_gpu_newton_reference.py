"""test/solvers/newtons_method.jl through the DEVICE solver: `allocate_cache` + `solve_newton!` of the host mirror call
`cts_newton_cache_create` / `cts_newton_solve`; the Newton loop, GMRES, the forward-difference JVP with its step sizes, the
forcing terms, the line search and the convergence checker all run in libcts_sm100a.so.  f!/j! and the dense Jacobian type
are the caller's (here torch on the same device), exactly the split the reference has.  Same cases, seeds and tolerances as
tests/test_newton_reference_cpu.py; in Float64 the iterates also agree with the oracle's to a few ulp-scale digits."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from newton_reference_cases import ROWS, SEEDS, Equation, main_algs  # noqa: E402
from ref_problems import R  # noqa: E402


@pytest.fixture(scope="module")
def cts():
    import cts_b200
    return cts_b200


class DenseJacobianGPU:
    """`j isa DenseMatrix` (newtons_method.jl:630-631): `ldiv!(Δx, lu(j), f)`; `mul!` for GMRES without a JVP (:481)."""

    def __init__(self, n, dtype):
        import torch
        self.n, self.dtype = n, np.dtype(dtype)
        self.M = torch.zeros((n, n), dtype=torch.float64 if self.dtype == np.float64 else torch.float32, device="cuda")

    def zero(self):
        return DenseJacobianGPU(self.n, self.dtype)

    def ldiv(self, x, b):
        import torch
        x.tensor.copy_(torch.linalg.solve(self.M, b.tensor))

    def mul(self, y, x):
        y.tensor.copy_(self.M @ x.tensor)


def _solve(cts, alg, use_j, eq):
    import torch
    x = cts.B200Vector.from_host(eq.x_init.copy())
    cache = cts.allocate_cache(alg, x, DenseJacobianGPU(eq.n, eq.FT) if use_j else None)

    def f_(f, xx):   # the user's residual: evaluated on the host in FT here (the test's closures are not the subject)
        f.tensor.copy_(torch.from_numpy(eq.f(xx.to_host())).cuda())

    def j_(j, xx):
        j.M.copy_(torch.from_numpy(np.ascontiguousarray(eq.j(xx.to_host()))).cuda())

    cts.solve_newton_(alg, cache, x, f_, j_ if use_j else None)
    return x.to_host(), cache


@pytest.mark.parametrize("seed", SEEDS)
@pytest.mark.parametrize("row", ROWS, ids=lambda r: f"{'lin' if r[0] else 'nonlin'}-{np.dtype(r[1]).name}-n{r[2]}")
def test_newtons_method_all_four_algorithms(cts, row, seed):
    is_linear, FT, n, adj = row
    eq = Equation("linear" if is_linear else "nonlinear", FT, n, seed)
    rtol, algs = main_algs(cts, FT, is_linear, adj)
    for k, (alg, use_j) in enumerate(algs):
        x, cache = _solve(cts, alg, use_j, eq)
        assert eq.rel_err(x) < rtol, (k + 1, eq.rel_err(x) / rtol)
        st = cache.stats()
        assert 1 <= st["newton_iters"] <= (1 if is_linear else 10)
        if k in (1, 2):
            assert st["krylov_iters"] >= 1


@pytest.mark.parametrize("seed", SEEDS)
@pytest.mark.parametrize("n", [1, 3, 10])
def test_eisenstat_walker_forcing(cts, n, seed):
    FT = np.float64
    eq = Equation("nonlinear", FT, n, seed)
    rtol = 100 * float(np.finfo(FT).eps)
    mk = lambda m: m.NewtonsMethod(max_iters=30, krylov_method=m.KrylovMethod(forcing_term=m.EisenstatWalkerForcing(), memory=30),
                                   convergence_checker=m.ConvergenceChecker(norm_condition=m.MaximumRelativeError(rtol)))
    x, cache = _solve(cts, mk(cts), True, eq)
    assert eq.rel_err(x) < rtol
    # same control flow as the oracle: number of Newton and Krylov iterations
    import test_newton_reference_cpu as T
    _, ocache = T._solve(mk(R), True, eq)
    st = cache.stats()
    # (the caller's dense closures round differently on the two sides -- LAPACK vs cuSOLVER/cuBLAS -- so GMRES may stop one
    # iteration earlier or later in a solve; the Newton iteration count is robust)
    assert st["newton_iters"] == ocache["newton_iters"]
    assert abs(st["krylov_iters"] - sum(s["niter"] for s in ocache["stats"])) <= max(2, st["newton_iters"])


@pytest.mark.parametrize("seed", SEEDS)
@pytest.mark.parametrize("step", [1, 2, 3])
def test_forward_diff_step_size_variants(cts, step, seed):
    FT, n = np.float64, 5
    eq = Equation("gentle", FT, n, seed)
    rtol = 100 * float(np.finfo(FT).eps)
    kind = {1: cts.ForwardDiffStepSize1, 2: cts.ForwardDiffStepSize2, 3: cts.ForwardDiffStepSize3}[step]()
    alg = cts.NewtonsMethod(max_iters=30, krylov_method=cts.KrylovMethod(
        jacobian_free_jvp=cts.ForwardDiffJVP(default_step=kind, step_adjustment=1.0), forcing_term=cts.ConstantForcing(rtol)),
        convergence_checker=cts.ConvergenceChecker(norm_condition=cts.MaximumRelativeError(rtol)))
    x, cache = _solve(cts, alg, False, eq)
    assert eq.rel_err(x) < 0.01
    assert cache.stats()["jvp_evals"] >= 1


@pytest.mark.parametrize("seed", SEEDS)
def test_chord_method_jacobian_reuse(cts, seed):
    FT, n = np.float64, 5
    rtol = 100 * float(np.finfo(FT).eps)
    cc = lambda: cts.ConvergenceChecker(norm_condition=cts.MaximumRelativeError(rtol))
    eq = Equation("linear", FT, n, seed)
    x, cache = _solve(cts, cts.NewtonsMethod(max_iters=1, update_j=cts.UpdateEvery(cts.NewNewtonSolve), convergence_checker=cc()), True, eq)
    assert eq.rel_err(x) < rtol and cache.stats()["wfact_evals"] == 1
    eq = Equation("gentle", FT, n, seed)
    x, cache = _solve(cts, cts.NewtonsMethod(max_iters=30, update_j=cts.UpdateEvery(cts.NewNewtonSolve), convergence_checker=cc()), True, eq)
    assert eq.rel_err(x) < 0.01
    st = cache.stats()
    assert st["wfact_evals"] == 1 and 2 <= st["newton_iters"] < 30   # one Jacobian, several iterations
