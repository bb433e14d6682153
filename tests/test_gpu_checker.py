"""ConvergenceChecker and LineSearch on the device path: the reference's known-answer sequences
(test/utils/convergence_checker.jl) through cts_newton_cache_is_converged, and whole IMEX steps whose Newton iteration
is controlled by a checker / line search against the oracle."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from ref_problems import R  # noqa: E402
from test_checker_cpu import CONDITIONS, LAST1, LAST2, err_func1, err_func2, val_func  # noqa: E402
from test_gpu_parity import _run_imex_pair  # noqa: E402


@pytest.fixture(scope="module")
def cts():
    import cts_b200
    return cts_b200


class _Checker:
    """A stand-alone Newton cache (cts_newton_cache_create) that only serves is_converged!."""

    def __init__(self, cts, checker, dtype=np.float64, n=2):
        from cts_b200 import _lib as L
        self.cts, self.ctx, self.dtype = cts, cts.Context.default(), dtype
        cfg = L.NewtonConfig()
        cfg.present, cfg.max_iters, cfg.memory = 1, 1, 20
        self.h = C.c_void_p()
        dummy_W = C.c_void_p(1)
        self.ctx.check(self.ctx.lib.cts_newton_cache_create(self.ctx.handle, 1 if dtype == np.float64 else 0, n, C.byref(cfg), dummy_W,
                                                            C.cast(None, L.LDIV_FN), C.cast(None, L.LDIV_FN), None, C.byref(self.h)))
        cc = checker.to_c()
        self.ctx.check(self.ctx.lib.cts_newton_cache_set_convergence_checker(self.h, C.byref(cc)))

    def is_converged(self, val, err, it):
        v, e = self.cts.B200Vector.from_host(val.astype(self.dtype)), self.cts.B200Vector.from_host(err.astype(self.dtype))
        out = C.c_int(-1)
        self.ctx.check(self.ctx.lib.cts_newton_cache_is_converged(self.h, v.ptr, e.ptr, it, C.byref(out)))
        return bool(out.value)

    def __del__(self):
        self.ctx.lib.cts_newton_cache_destroy(self.h)


def _passes(ch, last1, last2):
    for ef, last in ((err_func1, last1), (err_func2, last2)):
        for it in range(last):
            if ch.is_converged(val_func(it), ef(it), it):
                return False
        if not ch.is_converged(val_func(last), ef(last), last):
            return False
    return True


def test_single_conditions_reference_iteration_numbers(cts):
    for cond, a, b in zip(CONDITIONS(cts), LAST1, LAST2):
        assert _passes(_Checker(cts, cts.ConvergenceChecker(norm_condition=cond)), a, b), cond
        assert _passes(_Checker(cts, cts.ConvergenceChecker(component_condition=cond)), a, b), cond


@pytest.mark.parametrize("combiner,pick", [(all, max), (any, min)])
def test_multiple_conditions(cts, combiner, pick):
    cond = cts.MultipleConditions(combiner, *CONDITIONS(cts))
    assert _passes(_Checker(cts, cts.ConvergenceChecker(norm_condition=cond)), pick(LAST1), pick(LAST2))
    assert _passes(_Checker(cts, cts.ConvergenceChecker(component_condition=cond)), pick(LAST1), pick(LAST2))


@pytest.mark.parametrize("combiner,pick", [("&", max), ("|", min)])
def test_norm_and_component_combined(cts, combiner, pick):
    allc, anyc = cts.MultipleConditions(all, *CONDITIONS(cts)), cts.MultipleConditions(any, *CONDITIONS(cts))
    for nc, cc in ((allc, anyc), (anyc, allc)):
        ch = _Checker(cts, cts.ConvergenceChecker(norm_condition=nc, component_condition=cc, condition_combiner=combiner))
        assert _passes(ch, pick(LAST1), pick(LAST2))


@pytest.mark.parametrize("line_search", [False, True])
@pytest.mark.parametrize("kind", ["norm", "component", "both"])
def test_imex_steps_with_checker_and_line_search_match_oracle(cts, line_search, kind):
    # nonlinear column diffusion, Newton up to 8 iterations, stopped by the checker; the oracle's norm is the same fixed tree
    rtol = 1e-9
    mk = lambda m: m.ConvergenceChecker(
        norm_condition=m.MaximumRelativeError(rtol) if kind in ("norm", "both") else None,
        component_condition=m.MultipleConditions(any, m.MaximumRelativeError(1e-7), m.MaximumError(1e-12)) if kind in ("component", "both") else None)

    class KW(dict):
        pass

    import test_gpu_parity as P
    model, op, T0 = P._column_pair(cts, nlev=16, nx=4, ny=4, halo=1, nonlinear=True)
    alg_o = R.imex_algorithm("ARS343", R.NewtonsMethod(max_iters=8, convergence_checker=mk(R),
                                                         line_search=R.LineSearch() if line_search else None))
    alg_c = cts.IMEXAlgorithm(cts.ARS343(), cts.NewtonsMethod(max_iters=8, convergence_checker=mk(cts),
                                                               line_search=cts.LineSearch() if line_search else None))
    uo = T0.copy()
    model.dss(uo, None, 0.0)
    cache_o, step_o = R.make_stepper(model.ode_function(), alg_o, uo)
    uc = cts.B200Vector.from_host(T0)
    op.dss(uc, None, 0.0)
    integ = cts.init(cts.ODEProblem(op.ode_function(), uc, (0.0, 1e9), None), alg_c, dt=0.01)
    t = 0.0
    for _ in range(2):
        step_o(uo, None, t, 0.01)
        cts.step_(integ)
        t += 0.01
    np.testing.assert_array_equal(uc.to_host(), uo)
    st = integ.cache.stats()
    assert st["newton_iters"] == cache_o["nm_cache"]["newton_iters"]
    assert st["newton_iters"] < 2 * 3 * 8    # the checker ended the iterations early
    assert st["fused_stage_launches"] == 0
