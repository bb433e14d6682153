"""numpy restatement of the synthetic benchmark models of SURVEY §8d (configs C2-C5): hashed inputs,
upwind advection, the column diffusion model (T_imp, Wfact, T_exp, ghost-row dss!) and the batched
two-sided product-form Thomas solve that plays ``ldiv!`` for the column-stacked tridiagonal ``W``.

*** TEST INFRASTRUCTURE ONLY *** (see oracle/cts_ref.py).  These models are OURS, not the reference's
(the reference ships no model at this scale); they are defined once here, operation by operation, and
``csrc/ops.cu`` / ``csrc/tile_device.cuh`` must reproduce them bit for bit (numpy never contracts to FMA).
"""
from __future__ import annotations

import math

import numpy as np

import cts_ref as R

MASK64 = (1 << 64) - 1


def hash_uniform(seed: int, idx: np.ndarray) -> np.ndarray:
    """splitmix64(seed ^ idx) -> uniform [0,1) from the top 53 bits."""
    x = (np.uint64(seed) ^ idx.astype(np.uint64)) + np.uint64(0x9E3779B97F4A7C15)
    x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    x = x ^ (x >> np.uint64(31))
    return (x >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


def fill_hash(n, seed, offset=0, scale=1.0, shift=0.0, dtype=np.float64):
    """out[i] = scale * (hash(seed, offset+i) + shift), rounded to dtype (cts_fill_hash)."""
    with np.errstate(over="ignore"):
        u = hash_uniform(seed, np.arange(offset, offset + n, dtype=np.uint64))
    return (scale * (u + shift)).astype(dtype)


# ------------------------------------------------------------------------------ advection (C2) ---
class Advection:
    """rhs_i = (-c (u_i - u_{i-1})) / dx ; du = α rhs + β du   -- all in the state precision."""

    def __init__(self, n, c=1.0, dx=None, dtype=np.float64):
        self.FT = np.dtype(dtype).type
        self.n, self.negc, self.dx = n, self.FT(-c), self.FT(1.0 / n if dx is None else dx)

    def incr(self, du, u, p, t, alpha=True, beta=False):
        FT = self.FT
        rhs = (self.negc * (u - np.roll(u, 1))) / self.dx
        du[...] = FT(alpha) * rhs + FT(beta) * du


# ------------------------------------------------------------------------------ tridiagonal ldiv! ---
def _pow2_rescale(P):
    """Exact power of two 2^-e, e = unbiased exponent of P (1 for zero/subnormal/huge/non-finite P): the bit
    logic of csrc/tile_device.cuh::pow2_rescale."""
    if P.dtype == np.float64:
        be = ((P.view(np.uint64) >> np.uint64(52)) & np.uint64(0x7FF)).astype(np.int64)
        ok = (be != 0) & (be < 2046)
        sc = np.ldexp(np.float64(1.0), np.where(ok, 1023 - be, 0).astype(np.int32))
    else:
        be = ((P.view(np.uint32) >> np.uint32(23)) & np.uint32(0xFF)).astype(np.int64)
        ok = (be != 0) & (be < 254)
        sc = np.ldexp(np.float32(1.0), np.where(ok, 127 - be, 0).astype(np.int32)).astype(np.float32)
    return sc


def _rcp_checked(P):
    """csrc/tile_device.cuh::rcp_checked: correctly rounded 1/P for pivot products inside the safe exponent
    range, NaN otherwise (zero / subnormal / overflowing / non-finite product = singular or badly scaled W)."""
    if P.dtype == np.float64:
        be = ((np.ascontiguousarray(P).view(np.uint64) >> np.uint64(52)) & np.uint64(0x7FF)).astype(np.int64)
        bad = (be < 25) | (be > 2021)
    else:
        be = ((np.ascontiguousarray(P).view(np.uint32) >> np.uint32(23)) & np.uint32(0xFF)).astype(np.int64)
        bad = (be == 0) | (be == 255)
    return np.where(bad, P.dtype.type(np.nan), P.dtype.type(1) / P)


class _ProductSweep:
    """Product-form elimination (csrc/tile_device.cuh::ProductSweep): P_i = d_i P_{i-1} - (l_i u_{i-1}) P_{i-2},
    S_i = r_i P_{i-1} - l_i S_{i-1}, cp_i = (u_i P_{i-1})/P_i, rp_i = S_i/P_i, power-of-two rescale every 8 rows (4 in Float32)."""

    def __init__(self, nc, FT):
        self.P1, self.P2 = np.ones(nc, FT), np.zeros(nc, FT)
        self.S1, self.aprev = np.zeros(nc, FT), np.zeros(nc, FT)
        self.one = FT(1)
        self.every = 8 if np.dtype(FT).itemsize == 8 else 4  # RescaleEvery<T>

    def row(self, toward, dd, away, rr, s):
        lu = toward * self.aprev
        P = dd * self.P1 - lu * self.P2
        S = rr * self.P1 - toward * self.S1
        inv = _rcp_checked(P)
        fac = (away * self.P1) * inv
        rhs = S * inv
        self.P2, self.P1, self.S1, self.aprev = self.P1, P, S, away
        if (s + 1) % self.every == 0:
            sc = _pow2_rescale(np.ascontiguousarray(self.P1))
            self.P1, self.P2, self.S1 = self.P1 * sc, self.P2 * sc, self.S1 * sc
        return fac, rhs


def babe_solve(dl, d, du, rhs, nlev):
    """Two-sided, product-form Thomas elimination: identical operation sequence to csrc/tile_device.cuh
    (babe_solve_lanepair) and csrc/tridiag.cu (generic kernel).  Arrays are column-stacked (level fastest);
    returns x.  dl[0] / du[nlev-1] of each column are ignored."""
    FT = rhs.dtype.type
    sh = (-1, nlev)
    dl, d, du, r = (np.ascontiguousarray(a.reshape(sh)) for a in (dl, d, du, rhs))
    nc = r.shape[0]
    x = np.empty_like(r)
    one = FT(1)
    if nlev == 1:
        return (r[:, 0] * (one / d[:, 0])).reshape(rhs.shape)
    m = nlev // 2
    fac = np.empty_like(r)
    zero = np.zeros(nc, FT)
    with np.errstate(all="ignore"):
        sw = _ProductSweep(nc, FT)
        for i in range(m):  # down sweep
            fp, rp = sw.row(zero if i == 0 else dl[:, i], d[:, i], du[:, i], r[:, i], i)
            fac[:, i], x[:, i] = fp, rp
        cpm, rpm = fp, rp
        sw = _ProductSweep(nc, FT)
        for s, j in enumerate(range(nlev - 1, m - 1, -1)):  # up sweep
            fp, rp = sw.row(zero if j == nlev - 1 else du[:, j], d[:, j], dl[:, j], r[:, j], s)
            fac[:, j], x[:, j] = fp, rp
        lpm, spm = fp, rp
        det = one - cpm * lpm
        xm1 = (rpm - cpm * spm) / det
        xm = spm - lpm * xm1
        x[:, m - 1], x[:, m] = xm1, xm
        for i in range(m - 2, -1, -1):
            x[:, i] = x[:, i] - fac[:, i] * x[:, i + 1]
        for j in range(m + 1, nlev):
            x[:, j] = x[:, j] - fac[:, j] * x[:, j - 1]
    return x.reshape(rhs.shape)


def thomas_reference(dl, d, du, rhs, nlev):
    """Textbook Thomas in extended precision (accuracy yardstick for the product form; not used for parity)."""
    L = lambda a: a.reshape(-1, nlev).astype(np.longdouble)
    dl, d, du, r = L(dl), L(d), L(du), L(rhs)
    cp, rp = np.zeros_like(d), np.zeros_like(d)
    fp, q = np.zeros(d.shape[0], np.longdouble), np.zeros(d.shape[0], np.longdouble)
    for i in range(nlev):
        l = dl[:, i] if i > 0 else 0 * dl[:, i]
        piv = d[:, i] - l * fp
        fp = du[:, i] / piv
        q = (r[:, i] - l * q) / piv
        cp[:, i], rp[:, i] = fp, q
    x = np.zeros_like(d)
    x[:, -1] = rp[:, -1]
    for i in range(nlev - 2, -1, -1):
        x[:, i] = rp[:, i] - cp[:, i] * x[:, i + 1]
    return x.reshape(rhs.shape)


class TridiagW:
    """jac_prototype of the column model: (dl, d, du) + ldiv!/mul!."""

    def __init__(self, n, nlev, dtype):
        self.nlev = nlev
        self.dl, self.d, self.du = (np.zeros(n, dtype) for _ in range(3))

    def zero(self):
        return TridiagW(self.d.size, self.nlev, self.d.dtype)

    def ldiv(self, x, b):
        x[...] = babe_solve(self.dl, self.d, self.du, b, self.nlev)

    def mul(self, y, x):
        n = self.nlev
        X = x.reshape(-1, n)
        dl, d, du = (a.reshape(-1, n) for a in (self.dl, self.d, self.du))
        acc = d * X
        acc[:, 1:] = dl[:, 1:] * X[:, :-1] + acc[:, 1:]
        acc[:, :-1] = acc[:, :-1] + du[:, :-1] * X[:, 1:]
        y[...] = acc.reshape(y.shape)


# ------------------------------------------------------------------------------ column model ---
class ColumnModel:
    """State layout [rows][nx][nlev], rows = ny_local + 2*halo.  See include/cts.h (cts_column_desc)."""

    def __init__(self, nlev, nx, ny_local, K_col, dz=None, S_lev=None, nu=0.0, halo=0, nonlinear=False,
                 exchange=None):
        self.nlev, self.nx, self.ny, self.halo = nlev, nx, ny_local, halo
        self.dtype = K_col.dtype
        FT = self.FT = K_col.dtype.type
        self.dz = FT(1.0 / (nlev + 1) if dz is None else dz)
        self.K, self.S, self.nu, self.nonlinear = K_col, S_lev, FT(nu), nonlinear
        self.rows = ny_local + 2 * halo
        self.n = self.rows * nx * nlev
        self.a = (K_col / (self.dz * self.dz))[:, None]  # per-column diffusion number, state precision
        self.exchange = exchange  # (first_row, last_row) -> (ghost_lo, ghost_hi); None = local periodic wrap

    def _nb(self, u):
        U = u.reshape(-1, self.nlev)
        um = np.zeros_like(U)
        up = np.zeros_like(U)
        um[:, 1:] = U[:, :-1]
        up[:, :-1] = U[:, 1:]
        return U, um, up

    def _kface(self, ua, ub):
        FT = self.FT
        ubar = FT(0.5) * (ua + ub)
        return self.a * (FT(1) + ubar * ubar)

    def T_imp(self, du, u, p, t):
        FT = self.FT
        U, um, up = self._nb(u)
        if not self.nonlinear:
            out = self.a * ((um - FT(2) * U) + up)
        else:
            out = self._kface(U, up) * (up - U) - self._kface(um, U) * (U - um)
        du[...] = out.reshape(du.shape)

    def Wfact(self, W, u, p, dtg, t):
        FT = self.FT
        shp = (-1, self.nlev)
        if not self.nonlinear:
            o = float(dtg) * self.a.astype(np.float64) * np.ones(shp[1])
            W.dl[...] = o.astype(FT).ravel()
            W.du[...] = o.astype(FT).ravel()
            W.d[...] = ((-2.0 * o) - 1.0).astype(FT).ravel()
        else:
            U, um, up = self._nb(u)
            lo = float(dtg) * self._kface(um, U).astype(np.float64)
            hi = float(dtg) * self._kface(U, up).astype(np.float64)
            W.dl[...] = lo.astype(FT).ravel()
            W.du[...] = hi.astype(FT).ravel()
            W.d[...] = ((-(lo + hi)) - 1.0).astype(FT).ravel()

    def jac_prototype(self):
        return TridiagW(self.n, self.nlev, self.dtype)

    def T_exp(self, du, u, p, t):
        FT = self.FT
        h, ny, nx, nl = self.halo, self.ny, self.nx, self.nlev
        U = u.reshape(self.rows, nx, nl)
        out = du.reshape(self.rows, nx, nl)
        g = FT(1.0 + 0.5 * math.sin(2.0 * math.pi * t))
        val = np.zeros((ny, nx, nl), self.dtype)
        if self.S is not None:
            val = val + (self.S * g)[None, None, :]
        if self.nu != 0:
            own = slice(h, h + ny)
            uc = U[own]
            uw, ue = np.roll(U, 1, axis=1)[own], np.roll(U, -1, axis=1)[own]
            if h:
                us, un = U[h - 1:h - 1 + ny], U[h + 1:h + 1 + ny]
            else:
                us, un = np.roll(U, 1, axis=0), np.roll(U, -1, axis=0)
            lap = (((uw + ue) + us) + un) - FT(4) * uc
            val = val + self.nu * lap
        out[h:h + ny] = val

    def dss(self, u, p, t):
        if not self.halo:
            return
        U = u.reshape(self.rows, self.nx * self.nlev)
        if self.exchange is None:
            U[0], U[-1] = U[self.ny].copy(), U[1].copy()
        else:
            U[0], U[-1] = self.exchange(U[1].copy(), U[self.ny].copy())

    def ode_function(self, explicit=True, dss=None, **kw):
        T_imp = R.ODEFunction(self.T_imp, jac_prototype=self.jac_prototype(), Wfact=self.Wfact)
        args = dict(T_imp=T_imp)
        if explicit:
            args["T_exp"] = self.T_exp
        if (self.halo > 0) if dss is None else dss:
            args["dss"] = self.dss
        args.update(kw)
        return R.ClimaODEFunction(**args)


def column_inputs(nlev, nx, ny_global, row0, ny_local, halo, dtype=np.float64, seed_K=2, seed_T=3):
    """Hashed inputs of C3: K_col = 0.5(1 + hash(2,col)/2), T0 = hash(3, idx), S[lev] from the
    tutorial source 5 exp(-(z-0.3)^2/0.01); ghost rows hold the periodic neighbours' values so that any
    sharding sees the same global field.  Returns (K_col, T0, S_lev) for global rows [row0-halo, row0+ny_local+halo)."""
    rows = np.arange(row0 - halo, row0 + ny_local + halo) % ny_global
    cols = (rows[:, None] * nx + np.arange(nx)[None, :]).ravel()
    with np.errstate(over="ignore"):
        K = (0.5 * (1.0 + 0.5 * hash_uniform(seed_K, cols.astype(np.uint64)))).astype(dtype)
        idx = (cols[:, None] * nlev + np.arange(nlev)[None, :]).ravel()
        T0 = hash_uniform(seed_T, idx.astype(np.uint64)).astype(dtype)
    dz = 1.0 / (nlev + 1)
    z = (np.arange(nlev) + 1) * dz
    S = (5.0 * np.exp(-((z - 0.3) ** 2) / 0.01)).astype(dtype)
    return K, T0, S
