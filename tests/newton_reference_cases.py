"""The reference's Newton / Krylov unit tests (test/solvers/newtons_method.jl:8-175) restated once, for two back ends:
the numpy oracle (tests/test_newton_reference_cpu.py) and the device solver through `cts_newton_solve`
(tests/test_gpu_newton_reference.py).

Kept as written upstream: uniform `rand` matrices and right-hand sides, the three equation families
(`linear_equation` :8-17, `nonlinear_equation` with the sinh nonlinearity :19-28, `nonlinear_equation_gentle` :30-49), the
eight `(is_linear, FT, n, step_adjustment)` rows (:52-61), `max_iters = is_linear ? 1 : 10`, the
`ConvergenceChecker(norm_condition = MaximumRelativeError(100 eps(FT)))`, alg1..alg4 (:68-80) and the pass criterion
`norm(x - x_exact) / norm(x_exact) < rtol` (:88); the EisenstatWalkerForcing set (:93-115), the ForwardDiffStepSize2/3 set
(:117-141) and the chord tests (:143-175).

Not reproducible here: the stream of Julia's `MersenneTwister(1)`.  numpy's MT19937 `RandomState(seed)` supplies uniform
[0,1) samples instead, and every case is run for several seeds -- the tolerance is the pin, the data are not.
"""
import numpy as np

SEEDS = (1, 2, 3)

# (is_linear, FT, n, step_adjustment)   test/solvers/newtons_method.jl:52-61
ROWS = (
    (True, np.float32, 1, 100000),
    (True, np.float32, 10, 1000000),
    (True, np.float64, 1, 1000000000),
    (True, np.float64, 10, 10000000000),
    (False, np.float32, 1, 1),
    (False, np.float32, 10, 1),
    (False, np.float64, 1, 0.1),
    (False, np.float64, 10, 10),
)


def _rand(FT, n, seed):
    rng = np.random.RandomState(seed)
    A = rng.random_sample((n, n)).astype(FT)   # rand(rng, FT, n, n)
    b = rng.random_sample(n).astype(FT)        # rand(rng, FT, n)
    return A, b


class Equation:
    """f(x) and j(x) as dense numpy expressions in FT (the back ends wrap them as f!/j! on their own array types)."""

    def __init__(self, kind, FT, n, seed):
        self.kind, self.FT, self.n = kind, FT, n
        A, b = _rand(FT, n, seed)
        if kind == "gentle":
            A = (A + np.eye(n, dtype=FT) * n).astype(FT)   # :33-34 diagonally dominant
        self.A, self.b = A, b
        if kind == "linear":
            self.x_exact = np.linalg.solve(A, b)
        elif kind == "nonlinear":
            self.x_exact = np.linalg.solve(A, np.arcsinh(np.linalg.solve(A, b)) + b)
        else:
            x = np.linalg.solve(A, b)
            for _ in range(10):   # :39-45 full Newton on the host
                fv = A @ x + 0.1 * np.sin(x) - b
                jv = A + np.diag(0.1 * np.cos(x))
                x = x - np.linalg.solve(jv, fv)
            self.x_exact = x
        self.x_init = np.zeros(n, FT)

    def f(self, x):
        A, b = self.A, self.b
        x = np.asarray(x, dtype=self.FT)
        if self.kind == "linear":
            return (A @ x - b).astype(self.FT)
        if self.kind == "nonlinear":
            return (A @ np.sinh(A @ x - b) - b).astype(self.FT)
        return (A @ x + self.FT(0.1) * np.sin(x) - b).astype(self.FT)

    def j(self, x):
        A, b = self.A, self.b
        x = np.asarray(x, dtype=self.FT)
        if self.kind == "linear":
            return A
        if self.kind == "nonlinear":
            return (A @ np.diag(np.cosh(A @ x - b)) @ A).astype(self.FT)
        return (A + np.diag(self.FT(0.1) * np.cos(x))).astype(self.FT)

    def rel_err(self, x):
        x = np.asarray(x, dtype=np.float64)
        xe = self.x_exact.astype(np.float64)
        return float(np.linalg.norm(x - xe) / np.linalg.norm(xe))


def main_algs(m, FT, is_linear, step_adjustment):
    """alg1..alg4 of :64-80 built from module `m` (the oracle `cts_ref` or the product `cts_b200`); -> [(alg, use_j)]."""
    max_iters = 1 if is_linear else 10
    rtol = 100 * float(np.finfo(FT).eps)
    cc = lambda: m.ConvergenceChecker(norm_condition=m.MaximumRelativeError(rtol))
    return rtol, [
        (m.NewtonsMethod(max_iters=max_iters, convergence_checker=cc()), True),
        (m.NewtonsMethod(max_iters=max_iters, krylov_method=m.KrylovMethod(forcing_term=m.ConstantForcing(rtol)),
                         convergence_checker=cc()), True),
        (m.NewtonsMethod(max_iters=max_iters, krylov_method=m.KrylovMethod(
            jacobian_free_jvp=m.ForwardDiffJVP(step_adjustment=step_adjustment), forcing_term=m.ConstantForcing(rtol)),
            convergence_checker=cc()), False),
        (m.NewtonsMethod(max_iters=max_iters, convergence_checker=cc(), line_search=m.LineSearch()), True),
    ]
