"""Integrator driver: ``init`` / ``solve!`` / ``step!`` / ``reinit!`` / tstops / saving callback.

Host-only, O(1) work per step; mirrors ``src/integrators.jl`` (cited per function) and the callback
types of ``src/ClimaTimeSteppers.jl:139-167`` and ``src/Callbacks.jl``.  Julia's ``f!`` names are spelled
with a trailing underscore (``step_`` = ``step!``).
"""
from __future__ import annotations

import bisect
import time
from dataclasses import dataclass, field
from typing import Any, Callable, List, Optional, Sequence

import numpy as np

from .functions import ODEProblem, ODESolution
from .solvers import init_cache, step_u


class SortedQueue:
    """integrators.jl:12-46: next tstop first; ascending times forward, descending in reverse."""

    def __init__(self, vals, forward: bool):
        self.forward = forward
        self.data = sorted(vals, reverse=not forward)

    def __len__(self):
        return len(self.data)

    def isempty(self):
        return not self.data

    def first(self):
        return self.data[0]

    def pop(self):
        return self.data.pop(0)

    def push(self, val):
        if self.forward:
            bisect.insort(self.data, val)
        else:
            keys = [-x for x in self.data]
            self.data.insert(bisect.bisect_left(keys, -val), val)
        return self

    def empty(self):
        self.data.clear()
        return self


@dataclass
class DiscreteCallback:
    """ClimaTimeSteppers.jl:139-151."""
    condition: Callable
    affect: Callable
    initialize: Callable = lambda cb, u, t, integrator: None
    finalize: Callable = lambda cb, u, t, integrator: None


class CallbackSet:
    """ClimaTimeSteppers.jl:159-167 (flattens nested sets, drops ``nothing``)."""

    def __init__(self, *callbacks):
        out: List[DiscreteCallback] = []
        for cb in callbacks:
            if cb is None:
                continue
            if isinstance(cb, CallbackSet):
                out.extend(cb.discrete_callbacks)
            else:
                out.append(cb)
        self.discrete_callbacks = tuple(out)


def _tstops_and_saveat_queues(t0, tf, tstops, saveat=()):
    """integrators.jl:131-144."""
    forward = tf > t0
    ts = [t for t in tstops if (t0 < t < tf) or (tf < t < t0)] + [tf]
    if saveat is None:
        saveat = (t0, tf)
    return SortedQueue(ts, forward), SortedQueue(list(saveat), forward)


class TimeStepperIntegrator:
    """integrators.jl:78-108."""

    def __init__(self, **kw):
        self.__dict__.update(kw)


def _saving_callback(save_func, sol: ODESolution, save_everystep: bool) -> DiscreteCallback:
    """NonInterpolatingSavingCallback (integrators.jl:495-526)."""
    if save_everystep:
        condition = lambda u, t, integrator: True
    else:
        def condition(u, t, integrator):
            cond = False
            while (not integrator.saveat.isempty()) and (integrator.saveat.first() == integrator.t
                                                          or is_past_t(integrator, integrator.saveat.first())):
                cond = True
                integrator.saveat.pop()
            return cond

    def affect(integrator):
        sol.t.append(integrator.t)
        sol.u.append(save_func(integrator.u, integrator.t))

    def initialize(cb, u, t, integrator):
        if condition(u, t, integrator):
            affect(integrator)

    def finalize(cb, u, t, integrator):
        if (not save_everystep) and (not integrator.saveat.isempty()):
            affect(integrator)

    cb = DiscreteCallback(condition, affect, initialize, finalize)
    cb.is_saving = True
    return cb


def init(prob: ODEProblem, alg, *, dt, tstops=(), saveat=None, save_everystep=False, callback=None,
         advance_to_tstop=False, save_func=None, save=True, dtchangeable=True, stepstop=-1, tdir=None, **kwargs):
    """integrators.jl:187-257.  ``integrator.u`` aliases ``prob.u0`` (it is not copied, :236)."""
    u0, p = prob.u0, prob.p
    t0, tf = prob.tspan
    if not dt > 0:
        raise ValueError("dt must be positive")
    if isinstance(t0, np.float32) and isinstance(tf, np.float32):
        dt = np.float32(dt)  # promote(t0, tf, dt)
    _dt = dt
    dt = dt if tf > t0 else -dt
    if tdir is None:
        tdir = -1 if t0 > tf else 1
    q_tstops, q_saveat = _tstops_and_saveat_queues(t0, tf, tstops, saveat)
    if save_func is None:
        save_func = lambda u, t: u.copy()
    sol = ODESolution([], [], prob, alg)
    callback = CallbackSet(callback, _saving_callback(save_func, sol, save_everystep)) if save else CallbackSet(callback)
    integrator = TimeStepperIntegrator(
        alg=alg, u=u0, p=p, t=t0, dt=dt, _dt=_dt, dtchangeable=dtchangeable, tstops=q_tstops, _tstops=tstops,
        saveat=q_saveat, _saveat=saveat, step=0, stepstop=stepstop, callback=callback,
        advance_to_tstop=advance_to_tstop, cache=init_cache(prob, alg, dt=dt, **kwargs), sol=sol, tdir=tdir)
    f = prob.f
    if hasattr(f, "cache") and callable(getattr(f, "cache", None)):  # initialize_function! functions.jl:159-161
        f.cache(u0, p, t0)
    for cb in callback.discrete_callbacks:
        cb.initialize(cb, u0, t0, integrator)
    return integrator


_DEFAULT = object()   # "use what init was given": compared by identity, so arrays are fine as `saveat`


def reinit_(integrator, u0=None, *, t0=None, tf=None, erase_sol=True, tstops=None, saveat=_DEFAULT,
            reinit_callbacks=True):
    """integrators.jl:273-303."""
    prob = integrator.sol.prob
    u0 = prob.u0 if u0 is None else u0
    t0 = prob.tspan[0] if t0 is None else t0
    tf = prob.tspan[1] if tf is None else tf
    tstops = integrator._tstops if tstops is None else tstops
    saveat = integrator._saveat if saveat is _DEFAULT else saveat
    if u0 is not integrator.u:
        integrator.u.copyto(u0)
    integrator.t = t0
    integrator.tstops, integrator.saveat = _tstops_and_saveat_queues(t0, tf, tstops, saveat)
    integrator.step = 0
    if erase_sol:
        del integrator.sol.t[:]
        del integrator.sol.u[:]
    if reinit_callbacks:
        for cb in integrator.callback.discrete_callbacks:
            cb.initialize(cb, u0, t0, integrator)
    elif integrator.callback.discrete_callbacks:
        cb = integrator.callback.discrete_callbacks[-1]
        if getattr(cb, "is_saving", False):
            cb.initialize(cb, u0, t0, integrator)


def is_past_t(integrator, t) -> bool:
    return integrator.tdir * (t - integrator.t) < 0


def reached_tstop(integrator, tstop, stop_at_tstop=None) -> bool:
    if stop_at_tstop is None:
        stop_at_tstop = integrator.dtchangeable
    return integrator.t == tstop or ((not stop_at_tstop) and is_past_t(integrator, tstop))


def _eps(t) -> float:
    return float(np.spacing(np.abs(t if isinstance(t, np.float32) else np.float64(t))))


def __step_(integrator):
    """integrators.jl:410-439."""
    tstops = integrator.tstops
    integrator.step += 1
    if (not tstops.isempty()) and integrator.dtchangeable:
        integrator.dt = integrator.tdir * min(integrator._dt, abs(tstops.first() - integrator.t))
    else:
        integrator.dt = integrator.tdir * integrator._dt
    step_u(integrator)
    t_plus_dt = integrator.t + integrator.dt
    max_t_error = 100 * _eps(integrator.t)
    if (not tstops.isempty()) and abs(float(tstops.first()) - float(t_plus_dt)) < max_t_error:
        integrator.t = tstops.first()
    else:
        integrator.t = t_plus_dt
    for cb in integrator.callback.discrete_callbacks:
        if cb.condition(integrator.u, integrator.t, integrator):
            cb.affect(integrator)
    while (not tstops.isempty()) and reached_tstop(integrator, tstops.first()):
        tstops.pop()


def solve_(integrator):
    """``solve!`` (integrators.jl:345-351; `NVTX.@annotate`d there, so the range is opened here as well)."""
    lib = _nvtx_lib(integrator)
    if lib is not None:
        lib.cts_nvtx_range_push(b"solve!")
    try:
        while (not integrator.tstops.isempty()) and integrator.step != integrator.stepstop:
            __step_(integrator)
        for cb in integrator.callback.discrete_callbacks:
            cb.finalize(cb, integrator.u, integrator.t, integrator)
    finally:
        if lib is not None:
            lib.cts_nvtx_range_pop()
    return integrator.sol


def _nvtx_lib(integrator):
    """The C-ABI library behind a device state (None for the host-only integrators of the CPU tests)."""
    ctx = getattr(integrator.u, "ctx", None)
    return getattr(ctx, "lib", None)


def solve(prob, alg, **kwargs):
    """integrators.jl:329-337."""
    return solve_(init(prob, alg, **kwargs))


def step_(integrator, dt=None, stop_at_tdt=False):
    """``step!`` (integrators.jl:367-388)."""
    if dt is None:
        if integrator.advance_to_tstop:
            tstop = integrator.tstops.first()
            while not reached_tstop(integrator, tstop):
                __step_(integrator)
        else:
            __step_(integrator)
        return
    if dt <= 0:
        raise ValueError("dt must be positive")
    if stop_at_tdt and not integrator.dtchangeable:
        raise ValueError("Cannot stop at t + dt if dtchangeable is false")
    t_plus_dt = integrator.t + integrator.tdir * dt
    if stop_at_tdt:
        add_tstop_(integrator, t_plus_dt)
    while not reached_tstop(integrator, t_plus_dt, stop_at_tdt):
        __step_(integrator)


def get_dt(integrator):
    return integrator._dt


def set_dt_(integrator, dt):
    if dt <= 0:
        raise ValueError("dt must be positive")
    integrator._dt = dt


def add_tstop_(integrator, t):
    if is_past_t(integrator, t):
        raise ValueError(f"Cannot add a tstop at {t} because that is behind the current integrator time {integrator.t}")
    integrator.tstops.push(t)


def add_saveat_(integrator, t):
    if is_past_t(integrator, t):
        raise ValueError(f"Cannot add a saveat point at {t} because that is behind the current integrator time {integrator.t}")
    integrator.saveat.push(t)


# ---- Callbacks.jl ----------------------------------------------------------------------------------
def _ext_initialize(f, integrator):
    """`Callbacks.initialize!(f!, integrator)` (Callbacks.jl:17-21): an extension point of the affect; default no-op."""
    h = getattr(f, "initialize", None)
    if callable(h):
        h(integrator)


def _ext_finalize(f, integrator):
    """`Callbacks.finalize!(f!, integrator)` (Callbacks.jl:23-28)."""
    h = getattr(f, "finalize", None)
    if callable(h):
        h(integrator)


def EveryXSimulationSteps(f, dn: int, atinit=False) -> DiscreteCallback:
    """Callbacks.jl:144-174: fires when the step counter reaches `steps_next`, which then advances by Δsteps."""
    state = {"steps": 0, "steps_next": 0}

    def condition(u, t, integrator):
        state["steps"] += 1
        if state["steps"] >= state["steps_next"]:
            state["steps_next"] += dn
            return True
        return False

    def initialize(cb, u, t, integrator):
        state["steps"], state["steps_next"] = 0, dn
        _ext_initialize(f, integrator)
        if atinit:
            f(integrator)

    def finalize(cb, u, t, integrator):
        _ext_finalize(f, integrator)

    return DiscreteCallback(condition, f, initialize, finalize)


def EveryXSimulationTime(f, dt_cb, atinit=False) -> DiscreteCallback:
    """Callbacks.jl:101-133: `t_next` starts at Δt (an ABSOLUTE time, independent of t0) and advances past t on firing."""
    state = {"t_next": 0.0}

    def condition(u, t, integrator):
        if t >= state["t_next"]:
            while t >= state["t_next"]:
                state["t_next"] += dt_cb
            return True
        return False

    def initialize(cb, u, t, integrator):
        state["t_next"] = dt_cb
        _ext_initialize(f, integrator)
        if atinit:
            f(integrator)

    def finalize(cb, u, t, integrator):
        _ext_finalize(f, integrator)

    return DiscreteCallback(condition, f, initialize, finalize)


def EveryXWallTimeSeconds(f, dt_wall, ctx=None, atinit=False, clock=None) -> DiscreteCallback:
    """Callbacks.jl:40-90: the wall clock is all-reduced (max) so every rank fires on the same step (:59,:72):
    `ClimaComms.allreduce(comm_ctx, time(), max)` -> `cts_allreduce_max_f64_host` on the context's communicator.
    `clock` replaces `time.time` (tests)."""
    import ctypes as C
    state = {"next": 0.0}
    now_fn = clock if clock is not None else time.time

    def _now():
        t = now_fn()
        if ctx is not None and ctx.nranks > 1:
            v = C.c_double(t)
            ctx.check(ctx.lib.cts_allreduce_max_f64_host(ctx.handle, C.byref(v)), "allreduce(max)")
            t = v.value
        return t

    def condition(u, t, integrator):
        now = _now()
        if now >= state["next"]:
            while now >= state["next"]:
                state["next"] += dt_wall
            return True
        return False

    def initialize(cb, u, t, integrator):
        state["next"] = _now() + dt_wall
        _ext_initialize(f, integrator)
        if atinit:
            f(integrator)

    def finalize(cb, u, t, integrator):
        _ext_finalize(f, integrator)

    return DiscreteCallback(condition, f, initialize, finalize)


class HostSaver:
    """A `save_func` for `init(...; save_func)` (integrators.jl:197,495-526) that saves to the HOST without stalling the
    step loop: every save is a stream-ordered device-to-host copy (`cts_memcpy_d2h`) into its own pinned buffer; the
    values are valid after `wait()` (or any later synchronisation of the context's stream)."""

    def __init__(self, ctx):
        self.ctx = ctx
        self._bufs = []

    def __call__(self, u, t):
        import torch
        host = torch.empty(len(u), dtype=u.tensor.dtype, pin_memory=True)
        self.ctx.check(self.ctx.lib.cts_memcpy_d2h(self.ctx.handle, host.data_ptr(), u.ptr, host.numel() * host.element_size()),
                       "cts_memcpy_d2h")
        self._bufs.append(host)
        return host

    def wait(self):
        self.ctx.sync()
        return [b.numpy() for b in self._bufs]
