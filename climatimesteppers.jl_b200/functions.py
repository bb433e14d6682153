"""Host mirror of the reference's function / problem containers and update-signal handlers.

  ClimaODEFunction         src/functions.jl:81-150
  ODEFunction, IncrementingODEFunction, ODEProblem, ODESolution   src/problems.jl:17-147
  UpdateEvery / UpdateEveryN / UpdateEveryDt + signals             src/utilities/update_signal_handler.jl

Hooks receive :class:`~.device.B200Vector` arguments exactly where the reference passes the user's
array type.  A hook may also be a *library-native* operator method (``operators.py``); those are
handed to the engine as raw function addresses so it can recognise them and fuse.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Any, Callable, Optional

from . import _lib as L

# ---- update signals ----------------------------------------------------------------------------
NewTimeStep, NewNewtonSolve, NewNewtonIteration = "NewTimeStep", "NewNewtonSolve", "NewNewtonIteration"
WithDSS, EndOfStage, EndOfStep = "WithDSS", "EndOfStage", "EndOfStep"
SIGNAL_ID = {NewTimeStep: L.SIG_NEW_TIMESTEP, NewNewtonSolve: L.SIG_NEW_NEWTON_SOLVE,
             NewNewtonIteration: L.SIG_NEW_NEWTON_ITERATION, WithDSS: L.SIG_WITH_DSS,
             EndOfStage: L.SIG_END_OF_STAGE, EndOfStep: L.SIG_END_OF_STEP}
_ID_SIGNAL = {v: k for k, v in SIGNAL_ID.items()}
_SUPERS = {NewTimeStep: (NewTimeStep,), NewNewtonSolve: (NewNewtonSolve,), NewNewtonIteration: (NewNewtonIteration,),
           WithDSS: (WithDSS,), EndOfStage: (EndOfStage, WithDSS), EndOfStep: (EndOfStep, EndOfStage, WithDSS)}


class UpdateSignalHandler:
    signal: str = ""

    def needs_update(self, signal: str, t=None) -> bool:  # needs_update!(::UpdateSignalHandler, ::UpdateSignal) = false
        return False

    def to_c(self, keep: list) -> L.UpdateHandler:
        """Built-in handlers travel as data; stateful ones as a host callback."""
        h = L.UpdateHandler()
        h.every_signal = SIGNAL_ID[self.signal]
        if type(self) is not UpdateEvery:
            def _cb(ud, sig, t, _self=self):
                return 1 if _self.needs_update(_ID_SIGNAL[sig], t) else 0
            cb = L.UPDATE_FN(_cb)
            keep.append(cb)
            h.fn = cb
        return h


class UpdateEvery(UpdateSignalHandler):
    """update_signal_handler.jl:96-99: fires on the signal type or any subtype of it."""

    def __init__(self, signal: str):
        assert signal in SIGNAL_ID
        self.signal = signal

    def needs_update(self, signal, t=None):
        return self.signal in _SUPERS[signal]


class UpdateEveryN(UpdateSignalHandler):
    """update_signal_handler.jl:110-135."""

    def __init__(self, n: int, signal: str, reset_signal: Optional[str] = None):
        if reset_signal == signal:
            raise ValueError("Reset and update signal types cannot be the same.")
        self.n, self.signal, self.reset_signal, self.counter = n, signal, reset_signal, 0

    def needs_update(self, signal, t=None):
        if signal == self.signal:
            result = self.counter == 0
            self.counter += 1
            if self.counter == self.n:
                self.counter = 0
            return result
        if self.reset_signal is not None and signal == self.reset_signal:
            self.counter = 0
        return False


class UpdateEveryDt(UpdateSignalHandler):
    """update_signal_handler.jl:147-165."""
    signal = NewTimeStep

    def __init__(self, dt):
        self.dt, self.is_first_t, self.prev_update_t = dt, True, None

    def needs_update(self, signal, t=None):
        if signal != NewTimeStep:
            return False
        if self.is_first_t or abs(t - self.prev_update_t) >= self.dt:
            self.is_first_t, self.prev_update_t = False, t
            return True
        return False


# ---- functions ---------------------------------------------------------------------------------
def Returns_nothing(*a, **k):
    """``Returns(nothing)``: the default no-op hook (functions.jl:97-102)."""
    return None


@dataclass
class ODEFunction:
    """problems.jl:101-110: ``f(du,u,p,t)`` + ``jac_prototype`` / ``Wfact(W,u,p,dtγ,t)`` / ``tgrad``."""
    f: Callable
    jac_prototype: Any = None
    Wfact: Optional[Callable] = None
    tgrad: Optional[Callable] = None

    def __call__(self, *args):
        return self.f(*args)


@dataclass
class IncrementingODEFunction:
    """problems.jl:77-85: ``f(du,u,p,t,α=true,β=false)`` computes ``du .= α .* f(u,p,t) .+ β .* du``."""
    f: Callable
    iip: bool = True

    def __call__(self, *args):
        return self.f(*args)


class ClimaODEFunction:
    """functions.jl:81-144; the constructor normalises T_exp!/T_lim! into T_exp_T_lim! exactly like the reference."""

    def __init__(self, *, T_exp_T_lim=None, T_lim=None, T_exp=None, T_imp=None, lim=Returns_nothing,
                 dss=Returns_nothing, constrain_state=Returns_nothing, initialize_imp=Returns_nothing,
                 cache=Returns_nothing, cache_imp=None, update_cache: Optional[UpdateSignalHandler] = None,
                 update_constrain_state: Optional[UpdateSignalHandler] = None):
        if T_exp_T_lim is not None:
            assert T_exp is None, "`T_exp_T_lim!` was passed, `T_exp!` must be `nothing`"
            assert T_lim is None, "`T_exp_T_lim!` was passed, `T_lim!` must be `nothing`"
            has_lim = True
        elif T_exp is not None and T_lim is not None:
            def T_exp_T_lim(du_exp, du_lim, u, p, t, _e=T_exp, _l=T_lim):
                _e(du_exp, u, p, t)
                _l(du_lim, u, p, t)
            has_lim = True
        elif T_exp is not None:
            if getattr(T_exp, "_cts_native_exp_lim", None) is not None:
                T_exp_T_lim = T_exp  # a library-native explicit operator already has the fused form
            else:
                def T_exp_T_lim(du_exp, du_lim, u, p, t, _e=T_exp):
                    _e(du_exp, u, p, t)
            has_lim = False
        elif T_lim is not None:
            def T_exp_T_lim(du_exp, du_lim, u, p, t, _l=T_lim):
                _l(du_lim, u, p, t)
            has_lim = True
        else:
            has_lim = False
        self.T_exp_T_lim, self.T_imp = T_exp_T_lim, T_imp
        self.lim, self.dss, self.constrain_state = lim, dss, constrain_state
        self.initialize_imp, self.cache = initialize_imp, cache
        self.cache_imp = cache if cache_imp is None else cache_imp
        self.update_cache = UpdateEvery(EndOfStage) if update_cache is None else update_cache
        self.update_constrain_state = UpdateEvery(EndOfStep) if update_constrain_state is None else update_constrain_state
        self._has_lim = has_lim


def has_T_exp(f: ClimaODEFunction) -> bool:  # functions.jl:147
    return f.T_exp_T_lim is not None


def has_T_lim(f: ClimaODEFunction) -> bool:  # functions.jl:150
    return f._has_lim


@dataclass
class ODEProblem:
    """problems.jl:17-22."""
    f: Any
    u0: Any
    tspan: tuple
    p: Any = None


@dataclass
class ODESolution:
    """problems.jl:129-147; ``sol(t)`` is a nearest-neighbour lookup, not an interpolation."""
    t: list
    u: list
    prob: Any
    alg: Any

    def __call__(self, t):
        idx = min(range(len(self.t)), key=lambda i: abs(self.t[i] - t))
        return self.u[idx]
