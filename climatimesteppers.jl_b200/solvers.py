"""``init_cache`` / ``step_u!`` for the B200 state type -- the solver-dispatch contract of the reference
(``src/integrators.jl:250,441-444``: "solvers need to define this interface").

In Julia these are methods ``init_cache(prob, alg; dt)`` and ``step_u!(integrator, cache::IMEXARKCache{<:B200Vector})``
that ``ccall`` into the library; here the same calls are made through ctypes.  The
stage loop itself runs inside the library (``csrc/solver.cu``), which calls back into the user's hooks
through C function pointers at exactly the places the reference calls them.
"""
from __future__ import annotations

import ctypes as C
from typing import Any, Dict, List, Optional

import numpy as np

from . import _lib as L
from .algorithms import SSP, IMEXAlgorithm, LowStorageRungeKutta2N, RosenbrockAlgorithm
from .device import B200Vector, Context
from .functions import ClimaODEFunction, IncrementingODEFunction, Returns_nothing
from .operators import ColumnTridiag, NativeHook


class _Views:
    """device pointer -> B200Vector (hooks receive the arrays the engine works on)."""

    def __init__(self, ctx: Context, n: int, dtype):
        self.ctx, self.n, self.dtype = ctx, n, np.dtype(dtype)
        self.map: Dict[int, B200Vector] = {}

    def add(self, v: B200Vector):
        self.map[v.ptr] = v

    def get(self, ptr) -> Optional[B200Vector]:
        if not ptr:
            return None
        v = self.map.get(ptr)
        if v is None:
            v = B200Vector.from_ptr(ptr, self.n, self.dtype, self.ctx)
            self.map[ptr] = v
        return v


def _guard(fn):
    """Exceptions must not cross the C ABI: report them as a non-zero hook status and re-raise afterwards."""
    def wrapped(self, *a):
        try:
            fn(self, *a)
            return 0
        except BaseException as e:  # noqa: BLE001
            self.pending_exception = e
            return 1
    return wrapped


class _HookTable:
    """Builds the ``cts_ode_function`` for a ClimaODEFunction: native hooks by address, Python hooks as callbacks."""

    def __init__(self, f: ClimaODEFunction, p: Any, views: _Views, jac):
        self.f, self.p, self.views, self.jac = f, p, views, jac
        self.keep: List[Any] = []
        self.pending_exception: Optional[BaseException] = None
        self.c = L.OdeFunction()
        self._build()

    # Python trampolines ---------------------------------------------------------------------
    @_guard
    def _exp_lim(self, ud, du_exp, du_lim, u, t):
        g = self.views.get
        self.f.T_exp_T_lim(g(du_exp), g(du_lim), g(u), self.p, t)

    @_guard
    def _timp(self, ud, du, u, t):
        g = self.views.get
        self.f.T_imp(g(du), g(u), self.p, t)

    @_guard
    def _wfact(self, ud, W, u, dtg, t):
        self.f.T_imp.Wfact(self.jac, self.views.get(u), self.p, dtg, t)

    @_guard
    def _ldiv(self, ud, x, W, b):
        self.jac.ldiv(self.views.get(x), self.views.get(b))

    @_guard
    def _jmul(self, ud, y, W, x):
        self.jac.mul(self.views.get(y), self.views.get(x))

    def _state(self, fn):
        def cb(ud, u, t):
            try:
                fn(self.views.get(u), self.p, t)
                return 0
            except BaseException as e:  # noqa: BLE001
                self.pending_exception = e
                return 1
        return cb

    @_guard
    def _init_imp(self, ud, u, dtg):
        self.f.initialize_imp(self.views.get(u), self.p, dtg)

    @_guard
    def _lim(self, ud, u, t, u_ref):
        self.f.lim(self.views.get(u), self.p, t, self.views.get(u_ref))

    def _build(self):
        f, c, keep = self.f, self.c, self.keep
        natives = [h for h in (f.T_exp_T_lim, f.T_imp.f if f.T_imp is not None and hasattr(f.T_imp, "f") else None,
                               getattr(f.T_imp, "Wfact", None) if f.T_imp is not None else None, f.dss)
                   if isinstance(h, NativeHook)]
        ops = {id(h.op) for h in natives}
        primary = natives[0].op if len(ops) == 1 else None  # one `ud` per function table
        c.ud = primary.handle if primary is not None else None

        def native(h):
            return isinstance(h, NativeHook) and primary is not None and h.op is primary

        def put(field, hook, fntype, tramp):
            if hook is None:
                return
            if native(hook):
                setattr(c, field, C.cast(hook.address, fntype))
            else:
                cb = fntype(tramp)
                keep.append(cb)
                setattr(c, field, cb)

        put("T_exp_T_lim", f.T_exp_T_lim, L.EXP_LIM_FN, lambda *a: self._exp_lim(*a))
        c.has_lim = int(f._has_lim)
        if f.T_imp is not None:
            timp_fn = f.T_imp.f if hasattr(f.T_imp, "f") else f.T_imp
            put("T_imp", timp_fn, L.TENDENCY_FN, lambda *a: self._timp(*a))
            wf = getattr(f.T_imp, "Wfact", None)
            if wf is not None and self.jac is not None:
                put("Wfact", wf, L.WFACT_FN, lambda *a: self._wfact(*a))
                if isinstance(self.jac, ColumnTridiag) and native(wf):
                    c.ldiv = C.cast(L.fn_addr("cts_op_column_ldiv"), L.LDIV_FN)
                    c.jac_mul = C.cast(L.fn_addr("cts_op_column_jac_mul"), L.LDIV_FN)
                    c.W = self.jac.handle
                else:
                    cb = L.LDIV_FN(lambda *a: self._ldiv(*a))
                    keep.append(cb)
                    c.ldiv = cb
                    if hasattr(self.jac, "mul"):
                        cb2 = L.LDIV_FN(lambda *a: self._jmul(*a))
                        keep.append(cb2)
                        c.jac_mul = cb2
                    c.W = C.c_void_p(id(self.jac))  # opaque token; the trampolines use the Python object
        for name in ("dss", "constrain_state", "cache", "cache_imp"):
            hook = getattr(f, name)
            if hook is None or hook is Returns_nothing:
                continue  # NULL == Returns(nothing)
            put(name, hook, L.STATE_FN, self._state(hook))
        if f.initialize_imp is None:
            c.initialize_imp_is_nothing = 1
        elif f.initialize_imp is not Returns_nothing:
            cb = L.INIT_IMP_FN(lambda *a: self._init_imp(*a))
            keep.append(cb)
            c.initialize_imp = cb
        if f.lim is not None and f.lim is not Returns_nothing:
            cb = L.LIM_FN(lambda *a: self._lim(*a))
            keep.append(cb)
            c.lim = cb
        c.update_cache = f.update_cache.to_c(keep)
        c.update_constrain_state = f.update_constrain_state.to_c(keep)


class IMEXCache:
    """IMEXARKCache / IMEXSSPRKCache (imex_ark.jl:24-33, imex_ssprk.jl:19-31) living in the library."""

    def __init__(self, prob, alg: IMEXAlgorithm):
        u0: B200Vector = prob.u0
        f: ClimaODEFunction = prob.f
        self.ctx, self.alg, self.n, self.dtype = u0.ctx, alg, len(u0), u0.dtype
        self.views = _Views(self.ctx, self.n, self.dtype)
        self.views.add(u0)
        T_imp = f.T_imp
        has_jac = (T_imp is not None and getattr(T_imp, "Wfact", None) is not None
                   and getattr(T_imp, "jac_prototype", None) is not None)  # has_jac, imex_ark.jl:3-7
        self.jac = T_imp.jac_prototype.zero() if (has_jac and alg.newtons_method is not None) else None
        self.hooks = _HookTable(f, prob.p, self.views, self.jac)
        keep = self.hooks.keep
        tab_c = alg.tableau.to_c(alg.options["cast_tableau_to_state_eltype"])
        ncfg = alg.newtons_method.to_c(keep) if alg.newtons_method is not None else L.NewtonConfig()
        h = C.c_void_p()
        lib = self.ctx.lib
        self.ctx.check(lib.cts_imex_cache_create(self.ctx.handle, u0.ft, self.n, C.byref(tab_c),
                                                 1 if alg.constraint == SSP else 0, C.byref(self.hooks.c),
                                                 C.byref(ncfg), C.byref(h)), "init_cache")
        self.handle = h
        self._keep = (tab_c, ncfg)
        if alg.newtons_method is not None:
            nm = C.c_void_p()
            self.ctx.check(lib.cts_imex_cache_newton(h, C.byref(nm)))
            self.newton_handle = nm
            if nm.value:
                alg.newtons_method.configure_cache(self.ctx, nm)

    def buffer(self, name: str) -> Optional[B200Vector]:
        p = C.c_void_p()
        self.ctx.check(self.ctx.lib.cts_imex_cache_buffer(self.handle, name.encode(), C.byref(p)), "cache buffer")
        return self.views.get(p.value)

    def nbuffers(self) -> int:
        n = C.c_int()
        self.ctx.check(self.ctx.lib.cts_imex_cache_nbuffers(self.handle, C.byref(n)))
        return n.value

    def stats(self) -> dict:
        s = L.SolverStats()
        self.ctx.check(self.ctx.lib.cts_imex_cache_stats(self.handle, C.byref(s)))
        return s.as_dict()

    def set_fusion(self, enabled: bool):
        self.ctx.check(self.ctx.lib.cts_imex_cache_set_fusion(self.handle, int(enabled)))

    def set_graph(self, enabled: bool):
        """Step through a CUDA graph (captured every step, applied with cudaGraphExecUpdate) when the step is all library
        kernels and the context's stream is capturable."""
        self.ctx.check(self.ctx.lib.cts_imex_cache_set_graph(self.handle, int(enabled)))

    def graph_stats(self) -> dict:
        a, b = C.c_uint64(), C.c_uint64()
        self.ctx.check(self.ctx.lib.cts_imex_cache_graph_stats(self.handle, C.byref(a), C.byref(b)))
        return {"graph_launches": int(a.value), "instantiations": int(b.value)}

    def step_u(self, integrator):
        u: B200Vector = integrator.u
        self.views.add(u)
        dt = integrator.dt
        r = self.ctx.lib.cts_imex_step(self.handle, u.ptr, float(integrator.t), float(dt), int(isinstance(dt, np.float32)))
        exc, self.hooks.pending_exception = self.hooks.pending_exception, None
        if exc is not None:
            raise exc
        self.ctx.check(r, "step_u!")

    def __del__(self):
        try:
            if self.handle:
                self.ctx.lib.cts_imex_cache_destroy(self.handle)
        except Exception:
            pass


class RosenbrockCache:
    """RosenbrockCache (rosenbrock.jl:72-124) living in the library: U, fU, fU_imp, fU_exp, fU_lim, ∂Y∂t, k[1..N]."""

    def __init__(self, prob, alg: RosenbrockAlgorithm):
        u0: B200Vector = prob.u0
        f: ClimaODEFunction = prob.f
        self.ctx, self.alg, self.n, self.dtype = u0.ctx, alg, len(u0), u0.dtype
        self.views = _Views(self.ctx, self.n, self.dtype)
        self.views.add(u0)
        T_imp = f.T_imp
        self.jac = None
        if T_imp is not None:
            if getattr(T_imp, "Wfact", None) is None or getattr(T_imp, "jac_prototype", None) is None:
                raise ValueError("RosenbrockAlgorithm needs T_imp!.Wfact and T_imp!.jac_prototype")
            # `W = prob.f.T_imp!.jac_prototype` (:117): the prototype itself; the library's lazy column prototype is
            # materialised first
            self.jac = T_imp.jac_prototype
            if isinstance(self.jac, ColumnTridiag) and self.jac.handle is None:
                self.jac = self.jac.zero()
        self.hooks = _HookTable(f, prob.p, self.views, self.jac)
        keep = self.hooks.keep
        tgrad = getattr(T_imp, "tgrad", None) if T_imp is not None else None
        if tgrad is not None:
            hooks = self.hooks

            @_guard
            def tg(hooks_self, ud, du, u, t):
                tgrad(hooks.views.get(du), hooks.views.get(u), hooks.p, t)

            cb = L.TENDENCY_FN(lambda *a: tg(hooks, *a))
            keep.append(cb)
        else:
            cb = C.cast(None, L.TENDENCY_FN)
        tab_c = alg.tableau.to_c()
        h = C.c_void_p()
        self.ctx.check(self.ctx.lib.cts_rosenbrock_cache_create(self.ctx.handle, u0.ft, self.n, C.byref(tab_c),
                                                                C.byref(self.hooks.c), cb, C.byref(h)), "init_cache")
        self.handle = h
        self._keep = (tab_c, cb)

    def buffer(self, name: str) -> Optional[B200Vector]:
        p = C.c_void_p()
        self.ctx.check(self.ctx.lib.cts_rosenbrock_cache_buffer(self.handle, name.encode(), C.byref(p)), "cache buffer")
        return self.views.get(p.value)

    @property
    def k(self):
        return [self.buffer(f"k{i + 1}") for i in range(self.alg.tableau.s)]

    def stats(self) -> dict:
        s = L.SolverStats()
        self.ctx.check(self.ctx.lib.cts_rosenbrock_cache_stats(self.handle, C.byref(s)))
        return s.as_dict()

    def set_fusion(self, enabled: bool):
        self.ctx.check(self.ctx.lib.cts_rosenbrock_cache_set_fusion(self.handle, 1 if enabled else 0))

    def step_u(self, integrator):
        u: B200Vector = integrator.u
        self.views.add(u)
        r = self.ctx.lib.cts_rosenbrock_step(self.handle, u.ptr, float(integrator.t), float(integrator.dt))
        exc, self.hooks.pending_exception = self.hooks.pending_exception, None
        if exc is not None:
            raise exc
        self.ctx.check(r, "step_u!")

    def __del__(self):
        try:
            if self.handle:
                self.ctx.lib.cts_rosenbrock_cache_destroy(self.handle)
        except Exception:
            pass


class LSRKCache:
    """LowStorageRungeKutta2NIncCache (lsrk.jl:44-52)."""

    def __init__(self, prob, alg: LowStorageRungeKutta2N):
        u0: B200Vector = prob.u0
        self.ctx, self.n, self.dtype, self.p = u0.ctx, len(u0), u0.dtype, prob.p
        self.views = _Views(self.ctx, self.n, self.dtype)
        self.views.add(u0)
        A, B, Cc = alg.tableau(self.dtype)  # tableau(alg, eltype(du))  lsrk.jl:51
        self.tableau = (A, B, Cc)
        ns = len(A)
        arr = lambda v: (C.c_double * ns)(*[float(x) for x in v])
        fobj = prob.f.f if isinstance(prob.f, IncrementingODEFunction) else prob.f
        self.pending_exception = None
        if isinstance(fobj, NativeHook):
            fptr, ud = C.c_void_p(fobj.address), fobj.handle
            self._cb = None
        else:
            def cb(ud, du, u, t, alpha, beta):
                try:
                    fobj(self.views.get(du), self.views.get(u), self.p, t, alpha, beta)
                    return 0
                except BaseException as e:  # noqa: BLE001
                    self.pending_exception = e
                    return 1
            self._cb = L.INCR_FN(cb)
            fptr, ud = C.cast(self._cb, C.c_void_p), None
        h = C.c_void_p()
        self.ctx.check(self.ctx.lib.cts_lsrk_cache_create(self.ctx.handle, u0.ft, self.n, ns, arr(A), arr(B), arr(Cc),
                                                          fptr, ud, C.byref(h)), "init_cache (LSRK)")
        self.handle = h

    def buffer(self, name="du") -> B200Vector:
        p = C.c_void_p()
        self.ctx.check(self.ctx.lib.cts_lsrk_cache_buffer(self.handle, name.encode(), C.byref(p)))
        return self.views.get(p.value)

    def set_fusion(self, enabled: bool):
        self.ctx.check(self.ctx.lib.cts_lsrk_cache_set_fusion(self.handle, int(enabled)))

    def stats(self) -> dict:
        s = L.SolverStats()
        self.ctx.check(self.ctx.lib.cts_lsrk_cache_stats(self.handle, C.byref(s)))
        return s.as_dict()

    def nbuffers(self) -> int:
        return 1

    def step_u(self, integrator):
        u: B200Vector = integrator.u
        self.views.add(u)
        dt = integrator.dt
        r = self.ctx.lib.cts_lsrk_step(self.handle, u.ptr, float(integrator.t), float(dt), int(isinstance(dt, np.float32)))
        exc, self.pending_exception = self.pending_exception, None
        if exc is not None:
            raise exc
        self.ctx.check(r, "step_u! (LSRK)")

    def __del__(self):
        try:
            if self.handle:
                self.ctx.lib.cts_lsrk_cache_destroy(self.handle)
        except Exception:
            pass


def init_cache(prob, alg, **kwargs):
    """``init_cache(prob, alg; dt, kwargs...)`` (integrators.jl:250) dispatching on the algorithm type."""
    if not isinstance(prob.u0, B200Vector):
        raise TypeError("the B200 path steps B200Vector states only (there is no CPU fallback)")
    if isinstance(alg, IMEXAlgorithm):
        return IMEXCache(prob, alg)
    if isinstance(alg, LowStorageRungeKutta2N):
        return LSRKCache(prob, alg)
    if isinstance(alg, RosenbrockAlgorithm):
        return RosenbrockCache(prob, alg)
    raise TypeError(f"no B200 step_u! for {type(alg).__name__} (out of the hot-path scope, SURVEY §8)")


def step_u(integrator):
    """``step_u!(integrator)`` (integrators.jl:442-444)."""
    integrator.cache.step_u(integrator)


# ---- NewtonsMethod on its own (newtons_method.jl:583-665) ------------------------------------------------------
class NewtonCache:
    """``allocate_cache(alg::NewtonsMethod, x_prototype, j_prototype)`` (newtons_method.jl:583-598) for ``B200Vector``
    states: Δx, f, the Krylov workspace and the checker cache live in the library (``cts_newton_cache_create``);
    ``j_prototype`` is any object with ``zero()``, ``ldiv(x, b)`` and, for GMRES without a JVP, ``mul(y, x)`` -- the
    Jacobian type stays the user's, as in the reference."""

    def __init__(self, alg, x_prototype: B200Vector, j_prototype=None):
        self.ctx, self.alg = x_prototype.ctx, alg
        self.n, self.dtype = len(x_prototype), x_prototype.dtype
        self.views = _Views(self.ctx, self.n, self.dtype)
        self.jac = j_prototype.zero() if j_prototype is not None else None
        self.pending_exception: Optional[BaseException] = None
        self.keep: List[Any] = []
        cfg = alg.to_c(self.keep)

        def ldiv(ud, x, W, b):
            try:
                self.jac.ldiv(self.views.get(x), self.views.get(b))
                return 0
            except BaseException as e:  # noqa: BLE001
                self.pending_exception = e
                return 1

        def jmul(ud, y, W, x):
            try:
                self.jac.mul(self.views.get(y), self.views.get(x))
                return 0
            except BaseException as e:  # noqa: BLE001
                self.pending_exception = e
                return 1

        null = C.cast(None, L.LDIV_FN)
        self._ldiv = L.LDIV_FN(ldiv) if self.jac is not None else null
        self._jmul = L.LDIV_FN(jmul) if (self.jac is not None and hasattr(self.jac, "mul")) else null
        h = C.c_void_p()
        token = C.c_void_p(id(self.jac)) if self.jac is not None else None
        self.ctx.check(self.ctx.lib.cts_newton_cache_create(self.ctx.handle, x_prototype.ft, self.n, C.byref(cfg), token,
                                                            self._ldiv, self._jmul, None, C.byref(h)), "allocate_cache")
        self.handle, self._cfg = h, cfg
        alg.configure_cache(self.ctx, h)

    def stats(self) -> dict:
        s = L.SolverStats()
        self.ctx.check(self.ctx.lib.cts_newton_cache_stats(self.handle, C.byref(s)))
        return s.as_dict()

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                self.ctx.lib.cts_newton_cache_destroy(self.handle)
                self.handle = None
        except Exception:
            pass


def allocate_cache(alg, x_prototype: B200Vector, j_prototype=None) -> NewtonCache:
    return NewtonCache(alg, x_prototype, j_prototype)


def solve_newton_(alg, cache: NewtonCache, x: B200Vector, f_, j_=None, prepare_for_f_=None):
    """``solve_newton!(alg, cache, x, f!, j!, prepare_for_f!)`` (newtons_method.jl:609-665): the iteration, the linear
    solves (direct / GMRES / JFNK), forcing terms, line search and convergence checks run in the library
    (``cts_newton_solve``); ``f!(f, x)``, ``j!(j, x)`` and ``prepare_for_f!(x)`` are the caller's."""
    views = cache.views
    views.add(x)

    def guard(fn):
        def cb(*a):
            try:
                fn(*a)
                return 0
            except BaseException as e:  # noqa: BLE001
                cache.pending_exception = e
                return 1
        return cb

    fcb = L.RESIDUAL_FN(guard(lambda ud, f, xx: f_(views.get(f), views.get(xx))))
    jcb = L.JAC_FN(guard(lambda ud, W, xx: j_(cache.jac, views.get(xx)))) if (j_ is not None and cache.jac is not None) \
        else C.cast(None, L.JAC_FN)
    pcb = L.PREPARE_FN(guard(lambda ud, xx: prepare_for_f_(views.get(xx)))) if prepare_for_f_ is not None \
        else C.cast(None, L.PREPARE_FN)
    r = cache.ctx.lib.cts_newton_solve(cache.handle, x.ptr, fcb, jcb, pcb, None)
    exc, cache.pending_exception = cache.pending_exception, None
    if exc is not None:
        raise exc
    cache.ctx.check(r, "solve_newton!")
    return x
