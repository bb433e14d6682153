"""Library-native model operators: the "user model" (layer L5 of the reference) of the synthetic
benchmark problems, implemented as device code inside ``libcts_sm100a.so`` (``csrc/ops.cu``).

Each operator exposes hooks with the reference's signatures (``T_imp!(du,u,p,t)``, ``Wfact(W,u,p,dtγ,t)``,
``dss!(u,p,t)``, ``f(du,u,p,t,α,β)``); they can be called from Python like any user function, and when
they are plugged into a :class:`~.functions.ClimaODEFunction` the engine receives their C addresses and
may substitute fused kernels (bit-identical results).
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np

from . import _lib as L
from .device import B200Vector, Context
from .functions import ClimaODEFunction, IncrementingODEFunction, ODEFunction


class NativeHook:
    """A hook implemented by an exported C function of the library, bound to an operator handle."""

    def __init__(self, op, symbol: str, kind: str):
        self.op, self.symbol, self.kind = op, symbol, kind
        self._fn = getattr(op.ctx.lib, symbol)
        if kind == "exp_lim":
            self._cts_native_exp_lim = True

    @property
    def address(self) -> int:
        return L.fn_addr(self.symbol)

    @property
    def handle(self):
        return self.op.handle

    def __call__(self, *args):
        op, k = self.op, self.kind
        p = lambda v: None if v is None else (v.ptr if isinstance(v, B200Vector) else v)
        if k == "tendency":      # (du, u, p, t)
            du, u, _, t = args
            r = self._fn(op.handle, p(du), p(u), float(t))
        elif k == "exp_lim":     # (du_exp, du_lim, u, p, t)  or the plain T_exp!(du, u, p, t) form
            if len(args) == 4:
                du, u, _, t = args
                r = self._fn(op.handle, p(du), None, p(u), float(t))
            else:
                du, dl, u, _, t = args
                r = self._fn(op.handle, p(du), p(dl), p(u), float(t))
        elif k == "wfact":       # (W, u, p, dtγ, t)
            W, u, _, dtg, t = args
            r = self._fn(op.handle, W.handle, p(u), float(dtg), float(t))
        elif k == "state":       # (u, p, t)
            u, _, t = args
            r = self._fn(op.handle, p(u), float(t))
        elif k == "incr":        # (du, u, p, t, α=true, β=false)
            du, u, _, t = args[:4]
            alpha = args[4] if len(args) > 4 else True
            beta = args[5] if len(args) > 5 else False
            r = self._fn(op.handle, p(du), p(u), float(t), float(alpha), float(beta))
        else:
            raise TypeError(k)
        op.ctx.check(r, self.symbol)


class AdvectionOperator:
    """C2: 1-D periodic first-order upwind advection ``u_t = -c u_x`` as an IncrementingODEFunction."""

    def __init__(self, n: int, c: float = 1.0, dx: Optional[float] = None, dtype=np.float64, ctx: Optional[Context] = None):
        self.ctx = ctx or Context.default()
        self.n, self.c, self.dx, self.dtype = n, float(c), float(1.0 / n if dx is None else dx), np.dtype(dtype)
        h = C.c_void_p()
        ft = L.CTS_F64 if self.dtype == np.float64 else L.CTS_F32
        self.ctx.check(self.ctx.lib.cts_op_advection_create(self.ctx.handle, ft, n, self.c, self.dx, C.byref(h)),
                       "cts_op_advection_create")
        self.handle = h
        self.incr = NativeHook(self, "cts_op_advection_incr", "incr")

    def ode_function(self) -> IncrementingODEFunction:
        return IncrementingODEFunction(self.incr)

    def __del__(self):
        try:
            if self.handle:
                self.ctx.lib.cts_op_advection_destroy(self.handle)
        except Exception:
            pass


class ColumnTridiag:
    """The ``jac_prototype`` type of the column model: batched tridiagonal ``W`` with ``zero`` and ``ldiv!``
    (the plug-in point of newtons_method.jl:596,633; cf. MockDiagonalJacobian in test/performance/allocations.jl:26-30)."""

    def __init__(self, op: "ColumnOperator", allocate: bool = False):
        self.op, self.ctx = op, op.ctx
        self._w = None
        if allocate:
            w = C.POINTER(L.Tridiag)()
            self.ctx.check(self.ctx.lib.cts_op_column_make_W(op.handle, C.byref(w)), "cts_op_column_make_W")
            self._w = w

    def zero(self) -> "ColumnTridiag":  # Base.zero(j_prototype)
        return ColumnTridiag(self.op, allocate=True)

    @property
    def handle(self):
        return C.cast(self._w, C.c_void_p) if self._w else None

    def diagonals(self):
        """(dl, d, du) as B200Vector views (tests)."""
        w = self._w.contents
        n = w.ncols * w.nlev
        return tuple(B200Vector.from_ptr(p, n, self.op.dtype, self.ctx, keep=self) for p in (w.dl, w.d, w.du))

    def ldiv(self, x: B200Vector, b: B200Vector):  # LinearAlgebra.ldiv!(x, W, b)
        self.ctx.check(self.ctx.lib.cts_op_column_ldiv(self.op.handle, x.ptr, self.handle, b.ptr), "ldiv!")

    def mul(self, y: B200Vector, x: B200Vector):  # LinearAlgebra.mul!(y, W, x)
        self.ctx.check(self.ctx.lib.cts_op_column_jac_mul(self.op.handle, y.ptr, self.handle, x.ptr), "mul!")

    def __del__(self):
        try:
            if self._w:
                self.ctx.lib.cts_op_column_free_W(self.op.handle, self._w)
        except Exception:
            pass


class ColumnOperator:
    """C3/C4/C5 column model on a ``[rows][nx][nlev]`` state (level fastest, ``rows = ny_local + 2*halo``).

    ``T_imp``: vertical diffusion ``K_col d2/dz2`` (linear) or ``d/dz(K_col(1+U^2) dU/dz)`` (nonlinear), Dirichlet 0;
    ``Wfact``: ``W = dtγ J - I`` as a batched tridiagonal; ``T_exp``: source ``S[lev]*(1+sin(2πt)/2)`` + ``nu`` x 5-point
    horizontal Laplacian on owned rows; ``dss``: ghost-row refresh (local periodic copy or NCCL halo exchange).
    """

    def __init__(self, nlev: int, nx: int, ny_local: int, K_col: B200Vector, *, dz: Optional[float] = None,
                 S_lev: Optional[B200Vector] = None, nu: float = 0.0, halo: int = 0, nonlinear: bool = False,
                 ctx: Optional[Context] = None):
        self.ctx = ctx or K_col.ctx
        self.nlev, self.nx, self.ny_local, self.halo = nlev, nx, ny_local, halo
        self.dz = float(1.0 / (nlev + 1) if dz is None else dz)
        self.nu, self.nonlinear, self.dtype = float(nu), bool(nonlinear), K_col.dtype
        self.K_col, self.S_lev = K_col, S_lev
        rows = ny_local + 2 * halo
        assert len(K_col) == rows * nx, "K_col needs one value per local column (ghost rows included)"
        assert S_lev is None or (len(S_lev) == nlev and S_lev.dtype == K_col.dtype)
        d = L.ColumnDesc(nlev, nx, ny_local, halo, self.dz, self.nu, int(nonlinear), K_col.ptr,
                         S_lev.ptr if S_lev is not None else None)
        h = C.c_void_p()
        self.ctx.check(self.ctx.lib.cts_op_column_create(self.ctx.handle, K_col.ft, C.byref(d), C.byref(h)),
                       "cts_op_column_create")
        self.handle = h
        self.T_imp = NativeHook(self, "cts_op_column_T_imp", "tendency")
        self.T_exp = NativeHook(self, "cts_op_column_T_exp", "exp_lim")
        self.Wfact = NativeHook(self, "cts_op_column_Wfact", "wfact")
        self.dss = NativeHook(self, "cts_op_column_dss", "state")

    @property
    def ndof(self) -> int:
        return int(self.ctx.lib.cts_op_column_ndof(self.handle))

    @property
    def ndof_owned(self) -> int:
        return self.ny_local * self.nx * self.nlev

    def set_matrix_free(self, enabled: bool = True):
        self.ctx.check(self.ctx.lib.cts_op_column_set_matrix_free(self.handle, int(enabled)), "set_matrix_free")
        return self

    def jac_prototype(self) -> ColumnTridiag:
        return ColumnTridiag(self)

    def ode_function(self, *, explicit: bool = True, dss: Optional[bool] = None, **kw) -> ClimaODEFunction:
        """``ClimaODEFunction(; T_exp!, T_imp! = ODEFunction(T_imp!; jac_prototype, Wfact), dss!)``."""
        T_imp = ODEFunction(self.T_imp, jac_prototype=self.jac_prototype(), Wfact=self.Wfact)
        use_dss = self.halo > 0 if dss is None else dss
        args = dict(T_imp=T_imp)
        if explicit:
            args["T_exp"] = self.T_exp
        if use_dss:
            args["dss"] = self.dss
        args.update(kw)
        return ClimaODEFunction(**args)

    def foreign_ode_function(self, *, explicit: bool = True, dss: Optional[bool] = None, **kw) -> ClimaODEFunction:
        """The same model handed to the solver as OPAQUE user closures (Python callables around the library kernels), the
        way a ClimaAtmos-style model supplies `T_imp!`/`Wfact`/`ldiv!`: the engine cannot recognise the operator, so every
        step runs the generic pass-by-pass path through host callbacks (used to measure and test that path)."""
        op = self

        class _ForeignTridiag(ColumnTridiag):   # a user Jacobian type: `zero`, `ldiv!`, `mul!`
            def zero(self_inner):
                z = _ForeignTridiag(op, allocate=True)
                return z

        T_imp = ODEFunction(lambda du, u, p, t: op.T_imp(du, u, p, t), jac_prototype=_ForeignTridiag(self),
                            Wfact=lambda W, u, p, dtg, t: op.Wfact(W, u, p, dtg, t))
        use_dss = self.halo > 0 if dss is None else dss
        args = dict(T_imp=T_imp)
        if explicit:
            args["T_exp"] = lambda du, u, p, t: op.T_exp(du, u, p, t)
        if use_dss:
            args["dss"] = lambda u, p, t: op.dss(u, p, t)
        args.update(kw)
        return ClimaODEFunction(**args)

    def __del__(self):
        try:
            if self.handle:
                self.ctx.lib.cts_op_column_destroy(self.handle)
        except Exception:
            pass
