"""Algorithm objects of the reference API (host side).

  IMEXAlgorithm / ExplicitAlgorithm       src/solvers/imex_tableaus.jl:112-160, explicit_tableaus.jl:50-68
  NewtonsMethod / KrylovMethod / ForwardDiffJVP / forcing terms      src/nl_solvers/newtons_method.jl
  LSRK54CarpenterKennedy / LSRK144NiegemannDiehlBusch / LSRKEulerMethod   src/solvers/lsrk.jl
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Any, Optional

import numpy as np

from . import _lib as L
from . import tableaus as T
from .functions import NewNewtonIteration, UpdateEvery, UpdateSignalHandler


class TimeSteppingAlgorithm:
    pass


Unconstrained, SSP = "Unconstrained", "SSP"


# ---- Newton / Krylov -----------------------------------------------------------------------------
@dataclass
class ForwardDiffStepSize1:
    kind: int = 1


@dataclass
class ForwardDiffStepSize2:
    kind: int = 2


@dataclass
class ForwardDiffStepSize3:
    kind: int = 3


@dataclass
class ForwardDiffJVP:
    """newtons_method.jl:165-182."""
    default_step: Any = field(default_factory=ForwardDiffStepSize3)
    step_adjustment: float = 1


@dataclass
class ConstantForcing:
    """newtons_method.jl:211-220."""
    rtol: float = 0.0


@dataclass
class EisenstatWalkerForcing:
    """newtons_method.jl:248-280."""
    initial_rtol: float = 0.5
    gamma: float = 1
    alpha: float = 2
    min_rtol_threshold: float = 0.1
    max_rtol: float = 0.9


@dataclass
class KrylovMethod:
    """newtons_method.jl:419-438 -- GMRES only (`type = Val(GmresWorkspace)`, the default)."""
    jacobian_free_jvp: Optional[ForwardDiffJVP] = None
    forcing_term: Any = field(default_factory=ConstantForcing)
    memory: int = 20                    # kwargs = (; memory = 20)
    itmax: int = 0                      # solve_kwargs itmax (0 -> 2n)
    disable_preconditioner: bool = False
    batch_gram_schmidt: bool = False    # B200 option: one fused sweep of dots per Arnoldi step (classical GS)


# ---- ConvergenceChecker / LineSearch ---------------------------------------------------------------
class ConvergenceCondition:
    """convergence_condition.jl:31-33; ``nodes()`` flattens the condition tree into the postfix list the C ABI takes."""

    def nodes(self):
        raise NotImplementedError


@dataclass
class MaximumError(ConvergenceCondition):
    max_err: float

    def nodes(self):
        return [(L.COND_MAX_ERROR, 0, float(self.max_err), 0.0)]


@dataclass
class MaximumRelativeError(ConvergenceCondition):
    max_rel_err: float

    def nodes(self):
        return [(L.COND_MAX_REL_ERROR, 0, float(self.max_rel_err), 0.0)]


@dataclass
class MaximumErrorReduction(ConvergenceCondition):
    max_reduction: float

    def nodes(self):
        return [(L.COND_MAX_ERROR_REDUCTION, 0, float(self.max_reduction), 0.0)]


@dataclass
class MinimumRateOfConvergence(ConvergenceCondition):
    rate: float
    order: float = 1

    def nodes(self):
        return [(L.COND_MIN_RATE, 0, float(self.rate), float(self.order))]


class MultipleConditions(ConvergenceCondition):
    """``MultipleConditions(condition_combiner = all, conditions...)`` (convergence_condition.jl:116-143)."""

    def __init__(self, *args):
        if args and args[0] in (all, any):
            self.condition_combiner, self.conditions = args[0], tuple(args[1:])
        else:
            self.condition_combiner, self.conditions = all, tuple(args)
        if not self.conditions or not all(isinstance(c, ConvergenceCondition) for c in self.conditions):
            raise TypeError("MultipleConditions takes ConvergenceConditions")

    def nodes(self):
        out = []
        for c in reversed(self.conditions):  # the evaluator pops children in reverse order; all/any are symmetric anyway
            out += c.nodes()
        out.append((L.COND_ALL if self.condition_combiner is all else L.COND_ANY, len(self.conditions), 0.0, 0.0))
        return out


@dataclass
class ConvergenceChecker:
    """``ConvergenceChecker(; norm_condition, component_condition, condition_combiner = &)`` with ``norm =
    LinearAlgebra.norm`` (convergence_checker.jl:32-43)."""
    norm_condition: Optional[ConvergenceCondition] = None
    component_condition: Optional[ConvergenceCondition] = None
    condition_combiner: str = "&"

    def to_c(self) -> "L.ConvergenceCheckerC":
        if self.norm_condition is None and self.component_condition is None:
            raise ValueError("ConvergenceChecker must have at least one ConvergenceCondition")  # :47-49
        c = L.ConvergenceCheckerC()
        for cond, count, arr in ((self.norm_condition, "n_norm", c.norm_cond), (self.component_condition, "n_comp", c.comp_cond)):
            nodes = [] if cond is None else cond.nodes()
            if len(nodes) > L.MAX_COND_NODES:
                raise ValueError("condition tree too large")
            setattr(c, count, len(nodes))
            for k, (kind, nch, p0, p1) in enumerate(nodes):
                arr[k].kind, arr[k].nchildren, arr[k].p0, arr[k].p1 = kind, nch, p0, p1
        assert self.condition_combiner in ("&", "|")
        c.combiner_or = int(self.condition_combiner == "|")
        return c


@dataclass
class LineSearch:
    """``LineSearch()`` (line_search.jl:23-25): backtracking by halving, at most 5 times, on ``norm(f)``."""


@dataclass
class NewtonsMethod:
    """newtons_method.jl:568-581."""
    max_iters: int = 1
    update_j: UpdateSignalHandler = field(default_factory=lambda: UpdateEvery(NewNewtonIteration))
    krylov_method: Optional[KrylovMethod] = None
    convergence_checker: Any = None
    line_search: Any = None

    def to_c(self, keep: list) -> L.NewtonConfig:
        c = L.NewtonConfig()  # (convergence_checker / line_search are attached to the Newton cache, see configure_cache)
        c.present = 1
        c.max_iters = int(self.max_iters)
        c.update_j = self.update_j.to_c(keep)
        km = self.krylov_method
        c.memory = 20
        c.jvp_step_kind, c.jvp_step_adjustment = 3, 1.0
        if km is not None:
            c.use_krylov = 1
            c.memory, c.itmax = int(km.memory), int(km.itmax)
            c.disable_preconditioner = int(km.disable_preconditioner)
            c.batch_gram_schmidt = int(km.batch_gram_schmidt)
            if km.jacobian_free_jvp is not None:
                c.jacobian_free_jvp = 1
                c.jvp_step_kind = int(km.jacobian_free_jvp.default_step.kind)
                c.jvp_step_adjustment = float(km.jacobian_free_jvp.step_adjustment)
            ft = km.forcing_term
            if isinstance(ft, ConstantForcing):
                c.forcing_kind, c.forcing_rtol = 0, float(ft.rtol)
            else:
                c.forcing_kind = 1
                c.ew_initial_rtol, c.ew_gamma, c.ew_alpha = float(ft.initial_rtol), float(ft.gamma), float(ft.alpha)
                c.ew_min_rtol_threshold, c.ew_max_rtol = float(ft.min_rtol_threshold), float(ft.max_rtol)
        return c

    def configure_cache(self, ctx, newton_handle):
        """`allocate_cache(::NewtonsMethod, ...)` extras (newtons_method.jl:583-598): checker cache, line search."""
        import ctypes as C
        if self.convergence_checker is not None:
            cc = self.convergence_checker.to_c()
            ctx.check(ctx.lib.cts_newton_cache_set_convergence_checker(newton_handle, C.byref(cc)), "convergence_checker")
        if self.line_search is not None:
            ctx.check(ctx.lib.cts_newton_cache_set_line_search(newton_handle, 1), "line_search")


# ---- IMEX / explicit ------------------------------------------------------------------------------
class IMEXAlgorithm(TimeSteppingAlgorithm):
    """``IMEXAlgorithm(name_or_tableau, newtons_method, [constraint]; cast_tableau_to_state_eltype)``."""

    def __init__(self, name_or_tableau, newtons_method: Optional[NewtonsMethod], constraint: Optional[str] = None, *,
                 cast_tableau_to_state_eltype: bool = False):
        if isinstance(name_or_tableau, T.IMEXTableau):
            self.name, self.tableau = None, name_or_tableau
            default = Unconstrained
        else:
            self.name, self.tableau = name_or_tableau, name_or_tableau.tableau()
            default = name_or_tableau.default_constraint
        self.constraint = default if constraint is None else constraint
        assert self.constraint in (Unconstrained, SSP)
        self.newtons_method = newtons_method
        self.options = {"cast_tableau_to_state_eltype": bool(cast_tableau_to_state_eltype)}


def ExplicitAlgorithm(name_or_tableau, constraint: Optional[str] = None) -> IMEXAlgorithm:
    """explicit_tableaus.jl:50-68: an IMEXAlgorithm with a_imp ≡ a_exp and ``newtons_method = nothing``."""
    return IMEXAlgorithm(name_or_tableau, None, constraint)


class RosenbrockAlgorithm(TimeSteppingAlgorithm):
    """``RosenbrockAlgorithm(tableau)`` (rosenbrock.jl:48-70): one linear solve per stage, no Newton iteration."""

    def __init__(self, tableau: T.RosenbrockTableau):
        if not isinstance(tableau, T.RosenbrockTableau):
            raise TypeError("RosenbrockAlgorithm takes a RosenbrockTableau, e.g. tableau(SSPKnoth())")
        self.tableau = tableau


# ---- low-storage RK --------------------------------------------------------------------------------
class LowStorageRungeKutta2N(TimeSteppingAlgorithm):
    def tableau(self, dtype):
        raise NotImplementedError


class LSRK54CarpenterKennedy(LowStorageRungeKutta2N):
    """lsrk.jl:115-143."""

    def tableau(self, dtype=np.float64):
        return T.lsrk54_carpenter_kennedy(dtype)


class LSRK144NiegemannDiehlBusch(LowStorageRungeKutta2N):
    """lsrk.jl:151-207."""

    def tableau(self, dtype=np.float64):
        return T.lsrk144_niegemann_diehl_busch(dtype)


class LSRKEulerMethod(LowStorageRungeKutta2N):
    """lsrk.jl:216-217."""

    def tableau(self, dtype=np.float64):
        return T.lsrk_euler(dtype)
