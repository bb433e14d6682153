"""CPU restatement (numpy) of the ClimaTimeSteppers.jl per-step state-update path.

*** TEST INFRASTRUCTURE ONLY ***  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import this module.  The
product (``climatimesteppers.jl_b200``) never imports anything under ``oracle/``.

The reference is pure Julia and Julia is not installed in this image, so the reference
cannot be executed; this file restates its algorithm line by line and is *pinned* against the
reference's own golden values (``tests/test_oracle_golden.py``):
  * ``test/integration/single_column_ars.jl:308-315,330`` literal norms (ARS111/121/122/232/222)
    and the in-test independent ARS343 comparator (``:234-297``),
  * analytic solutions + expected orders of ``test/problems/*.jl`` / ``test/utils/convergence_utils.jl``,
  * ``Unconstrained == SSP`` to 100 eps (``test/utils/convergence_utils.jl:228-238``).
GMRES lives in the un-vendored Krylov.jl (v0.10.6 per ``docs/Manifest.toml:931-935``); its
published algorithm is restated in :func:`gmres` -- integrator-level JFNK parity is unpinned upstream.

Arithmetic conventions: every element-wise expression is evaluated in the
order Julia's broadcast evaluates it, with no FMA contraction (numpy never fuses), coefficients
``dt*a_ij`` rounded once on the host, Float64 tableau coefficients promote Float32 states to
Float64 inside each expression and round on store.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Any, Callable, Optional, Sequence

import numpy as np

# --------------------------------------------------------------------------------------
# update-signal handlers  (src/utilities/update_signal_handler.jl:29-165)
# --------------------------------------------------------------------------------------
# signal hierarchy: EndOfStep <: EndOfStage <: WithDSS (update_signal_handler.jl:65-86)
_SIGNAL_SUPERS = {
    "NewTimeStep": ("NewTimeStep",),
    "NewNewtonSolve": ("NewNewtonSolve",),
    "NewNewtonIteration": ("NewNewtonIteration",),
    "WithDSS": ("WithDSS",),
    "EndOfStage": ("EndOfStage", "WithDSS"),
    "EndOfStep": ("EndOfStep", "EndOfStage", "WithDSS"),
}


class UpdateEvery:
    """update_signal_handler.jl:96-99 -- fires on the target signal type or any subtype."""

    def __init__(self, signal: str):
        self.signal = signal

    def needs_update(self, signal: str, t=None) -> bool:
        return self.signal in _SIGNAL_SUPERS[signal]


class UpdateEveryN:
    """update_signal_handler.jl:110-135 (exact-type match, counter semantics)."""

    def __init__(self, n: int, signal: str, reset_signal: Optional[str] = None):
        self.n, self.signal, self.reset_signal, self.counter = n, signal, reset_signal, 0

    def needs_update(self, signal: str, t=None) -> bool:
        if signal == self.signal:
            result = self.counter == 0
            self.counter += 1
            if self.counter == self.n:
                self.counter = 0
            return result
        if self.reset_signal is not None and signal == self.reset_signal:
            self.counter = 0
        return False


def _fires_on_new_timestep(h) -> bool:  # async_utils.jl:85-88
    return isinstance(h, (UpdateEvery, UpdateEveryN)) and h.signal == "NewTimeStep"


# --------------------------------------------------------------------------------------
# function containers  (src/functions.jl:81-150, src/problems.jl:77-110)
# --------------------------------------------------------------------------------------
def _noop(*a, **k):
    return None


@dataclass
class ODEFunction:
    """problems.jl:101-110: callable ``f(du,u,p,t)`` + optional ``jac_prototype``/``Wfact``."""

    f: Callable
    jac_prototype: Any = None
    Wfact: Optional[Callable] = None
    tgrad: Optional[Callable] = None

    def __call__(self, *args):
        return self.f(*args)


class ClimaODEFunction:
    """functions.jl:81-144 -- normalises T_exp!/T_lim! into T_exp_T_lim! exactly as the ctor does."""

    def __init__(self, *, T_exp_T_lim=None, T_lim=None, T_exp=None, T_imp=None, lim=_noop, dss=_noop,
                 constrain_state=_noop, initialize_imp=_noop, cache=_noop, cache_imp=None,
                 update_cache=None, update_constrain_state=None):
        if T_exp_T_lim is not None:
            assert T_exp is None and T_lim is None
            has_lim = True
        elif T_exp is not None and T_lim is not None:
            def T_exp_T_lim(du_exp, du_lim, u, p, t, _e=T_exp, _l=T_lim):
                _e(du_exp, u, p, t)
                _l(du_lim, u, p, t)
            has_lim = True
        elif T_exp is not None:
            def T_exp_T_lim(du_exp, du_lim, u, p, t, _e=T_exp):
                _e(du_exp, u, p, t)
            has_lim = False
        elif T_lim is not None:
            def T_exp_T_lim(du_exp, du_lim, u, p, t, _l=T_lim):
                _l(du_lim, u, p, t)
            has_lim = True
        else:
            has_lim = False
        self.T_exp_T_lim, self.T_imp, self.lim, self.dss = T_exp_T_lim, T_imp, lim, dss
        self.constrain_state, self.initialize_imp, self.cache = constrain_state, initialize_imp, cache
        self.cache_imp = cache if cache_imp is None else cache_imp
        self.update_cache = UpdateEvery("EndOfStage") if update_cache is None else update_cache
        self.update_constrain_state = (UpdateEvery("EndOfStep") if update_constrain_state is None
                                       else update_constrain_state)
        self._has_lim = has_lim

    @property
    def has_T_exp(self):  # functions.jl:147
        return self.T_exp_T_lim is not None

    @property
    def has_T_lim(self):  # functions.jl:150
        return self._has_lim


# --------------------------------------------------------------------------------------
# tableaux  (src/solvers/imex_tableaus.jl, src/solvers/lsrk.jl:117-143)
# --------------------------------------------------------------------------------------
@dataclass
class IMEXTableau:
    """imex_tableaus.jl:35-79: defaults b = last row of a, c = row sums of a."""

    a_exp: np.ndarray
    a_imp: np.ndarray
    b_exp: Optional[np.ndarray] = None
    b_imp: Optional[np.ndarray] = None
    c_exp: Optional[np.ndarray] = None
    c_imp: Optional[np.ndarray] = None

    def __post_init__(self):
        self.a_exp = np.array(self.a_exp, dtype=np.float64)
        self.a_imp = np.array(self.a_imp, dtype=np.float64)
        if self.b_exp is None:
            self.b_exp = self.a_exp[-1, :].copy()
        if self.b_imp is None:
            self.b_imp = self.a_imp[-1, :].copy()
        self.b_exp = np.array(self.b_exp, dtype=np.float64)
        self.b_imp = np.array(self.b_imp, dtype=np.float64)
        if self.c_exp is None:
            self.c_exp = _rowsum(self.a_exp)
        if self.c_imp is None:
            self.c_imp = _rowsum(self.a_imp)
        self.c_exp = np.array(self.c_exp, dtype=np.float64)
        self.c_imp = np.array(self.c_imp, dtype=np.float64)
        assert np.all(np.triu(self.a_exp) == 0)
        assert np.all(np.triu(self.a_imp, 1) == 0)

    @property
    def s(self):
        return len(self.b_exp)

    def is_fsal(self):  # imex_tableaus.jl:59-62
        return (np.array_equal(self.b_exp, self.a_exp[-1, :])
                and np.array_equal(self.b_imp, self.a_imp[-1, :]))


def _rowsum(a):
    # Julia `sum(a; dims=2)` on an SMatrix: sequential left-to-right accumulation per row.
    out = np.zeros(a.shape[0])
    for i in range(a.shape[0]):
        acc = 0.0
        for j in range(a.shape[1]):
            acc = acc + float(a[i, j])
        out[i] = acc
    return out


def tableau(name: str, **kw) -> IMEXTableau:
    """Named tableaux restated from imex_tableaus.jl (file:line in each branch)."""
    s2, s3 = math.sqrt(2.0), math.sqrt(3.0)
    if name == "ARS111":  # :176-179
        return IMEXTableau(a_exp=[[0, 0], [1, 0]], a_imp=[[0, 0], [0, 1]])
    if name == "ARS121":  # :188-195
        return IMEXTableau(a_exp=[[0, 0], [1, 0]], b_exp=[0, 1], a_imp=[[0, 0], [0, 1]])
    if name == "ARS122":  # :203-211
        return IMEXTableau(a_exp=[[0, 0], [0.5, 0]], b_exp=[0, 1], a_imp=[[0, 0], [0, 0.5]], b_imp=[0, 1])
    if name == "ARS233":  # :219-237
        g = 1 / 2 + s3 / 6
        return IMEXTableau(a_exp=[[0, 0, 0], [g, 0, 0], [g - 1, 2 - 2 * g, 0]], b_exp=[0, 0.5, 0.5],
                           a_imp=[[0, 0, 0], [0, g, 0], [0, 1 - 2 * g, g]], b_imp=[0, 0.5, 0.5])
    if name == "ARS232":  # :245-262
        g = 1 - s2 / 2
        d = -2 * s2 / 3
        return IMEXTableau(a_exp=[[0, 0, 0], [g, 0, 0], [d, 1 - d, 0]], b_exp=[0, 1 - g, g],
                           a_imp=[[0, 0, 0], [0, g, 0], [0, 1 - g, g]])
    if name == "ARS222":  # :270-286   (`1 / 2γ` is 1/(2γ) in Julia)
        g = 1 - s2 / 2
        d = 1 - 1 / (2 * g)
        return IMEXTableau(a_exp=[[0, 0, 0], [g, 0, 0], [d, 1 - d, 0]],
                           a_imp=[[0, 0, 0], [0, g, 0], [0, 1 - g, g]])
    if name == "ARS343":  # :293-323 (expression order kept)
        g = 0.4358665215084590
        a42 = 0.5529291480359398
        a43 = a42
        b1 = -3 / 2 * g ** 2 + 4 * g - 1 / 4
        b2 = 3 / 2 * g ** 2 - 5 * g + 5 / 4
        a31 = ((1 - 9 / 2 * g + 3 / 2 * g ** 2) * a42 + (11 / 4 - 21 / 2 * g + 15 / 4 * g ** 2) * a43
               - 7 / 2 + 13 * g - 9 / 2 * g ** 2)
        a32 = ((-1 + 9 / 2 * g - 3 / 2 * g ** 2) * a42 + (-11 / 4 + 21 / 2 * g - 15 / 4 * g ** 2) * a43
               + 4 - 25 / 2 * g + 9 / 2 * g ** 2)
        a41 = 1 - a42 - a43
        return IMEXTableau(
            a_exp=[[0, 0, 0, 0], [g, 0, 0, 0], [a31, a32, 0, 0], [a41, a42, a43, 0]],
            b_exp=[0, b1, b2, g],
            a_imp=[[0, 0, 0, 0], [0, g, 0, 0], [0, (1 - g) / 2, g, 0], [0, b1, b2, g]])
    if name == "ARS443":  # :331-353
        return IMEXTableau(
            a_exp=[[0, 0, 0, 0, 0], [1 / 2, 0, 0, 0, 0], [11 / 18, 1 / 18, 0, 0, 0],
                   [5 / 6, -5 / 6, 1 / 2, 0, 0], [1 / 4, 7 / 4, 3 / 4, -7 / 4, 0]],
            a_imp=[[0, 0, 0, 0, 0], [0, 1 / 2, 0, 0, 0], [0, 1 / 6, 1 / 2, 0, 0],
                   [0, -1 / 2, 1 / 2, 1 / 2, 0], [0, 3 / 2, -3 / 2, 1 / 2, 1 / 2]])
    if name == "SSP222":  # :654-668
        g = 1 - s2 / 2
        return IMEXTableau(a_exp=[[0, 0], [1, 0]], b_exp=[0.5, 0.5],
                           a_imp=[[g, 0], [1 - 2 * g, g]], b_imp=[0.5, 0.5])
    if name == "SSP322":  # :676-692
        return IMEXTableau(a_exp=[[0, 0, 0], [0, 0, 0], [0, 1, 0]], b_exp=[0, 0.5, 0.5],
                           a_imp=[[0.5, 0, 0], [-0.5, 0.5, 0], [0, 0.5, 0.5]], b_imp=[0, 0.5, 0.5])
    if name == "SSP332":  # :700-717
        g = 1 - s2 / 2
        return IMEXTableau(a_exp=[[0, 0, 0], [1, 0, 0], [1 / 4, 1 / 4, 0]], b_exp=[1 / 6, 1 / 6, 2 / 3],
                           a_imp=[[g, 0, 0], [1 - 2 * g, g, 0], [1 / 2 - g, 0, g]], b_imp=[1 / 6, 1 / 6, 2 / 3])
    if name == "SSP333":  # :728-748
        beta = kw.get("beta", 1 / 2 + s3 / 6)
        assert beta > 1 / 2
        g = (2 * beta ** 2 - 3 * beta / 2 + 1 / 3) / (2 - 4 * beta)
        return IMEXTableau(a_exp=[[0, 0, 0], [1, 0, 0], [1 / 4, 1 / 4, 0]], b_exp=[1 / 6, 1 / 6, 2 / 3],
                           a_imp=[[0, 0, 0], [4 * g + 2 * beta, 1 - 4 * g - 2 * beta, 0],
                                  [1 / 2 - beta - g, g, beta]], b_imp=[1 / 6, 1 / 6, 2 / 3])
    if name == "SSP433":  # :756-777
        al, be, et = 0.24169426078821, 0.06042356519705, 0.12915286960590
        return IMEXTableau(
            a_exp=[[0, 0, 0, 0], [0, 0, 0, 0], [0, 1, 0, 0], [0, 1 / 4, 1 / 4, 0]], b_exp=[0, 1 / 6, 1 / 6, 2 / 3],
            a_imp=[[al, 0, 0, 0], [-al, al, 0, 0], [0, 1 - al, al, 0], [be, et, 1 / 2 - al - be - et, al]],
            b_imp=[0, 1 / 6, 1 / 6, 2 / 3])
    if name in _IMKG:  # :399-607 (IMKGTableau :378-389)
        return _imkg_tableau(*_IMKG[name]())
    if name == "DBM453":  # :789-822
        g = 0.32591194130117247
        bb = [0.87795339639076675, -0.72692641526151547, 0.7520413715737272, -0.22898029400415088, g]
        return IMEXTableau(
            a_exp=[[0, 0, 0, 0, 0],
                   [0.10306208811591838, 0, 0, 0, 0],
                   [-0.94124866143519894, 1.6626399742527356, 0, 0, 0],
                   [-1.3670975201437765, 1.3815852911016873, 1.2673234025619065, 0, 0],
                   [-0.81287582068772448, 0.81223739060505738, 0.90644429603699305, 0.094194134045674111, 0]],
            b_exp=bb,
            a_imp=[[0, 0, 0, 0, 0],
                   [-0.2228498531852541, g, 0, 0, 0],
                   [-0.46801347074080545, 0.86349284225716961, g, 0, 0],
                   [-0.46509906651927421, 0.81063103116959553, 0.61036726756832357, g, 0],
                   bb])
    if name == "HOMMEM1":  # :831-856
        ae = np.zeros((6, 6))
        ai = np.zeros((6, 6))
        for i, v in enumerate((1 / 5, 1 / 5, 1 / 3, 1 / 2, 1)):
            ae[i + 1, i] = v
        for i, v in enumerate((1 / 5, 1 / 5, 1 / 3, 1 / 2)):
            ai[i + 1, i + 1] = v
        ai[5, :] = [5 / 18, 5 / 18, 0, 0, 0, 8 / 18]
        return IMEXTableau(a_exp=ae, a_imp=ai)
    if name == "ARK2GKC":  # :866-881
        a32 = (1 / 2 + s2 / 3) if kw.get("paper_version", False) else 1 / 2
        return IMEXTableau(a_exp=[[0, 0, 0], [2 - s2, 0, 0], [1 - a32, a32, 0]], b_exp=[s2 / 4, s2 / 4, 1 - s2 / 2],
                           a_imp=[[0, 0, 0], [1 - s2 / 2, 1 - s2 / 2, 0], [s2 / 4, s2 / 4, 1 - s2 / 2]])
    if name in ("ARK437L2SA1", "ARK548L2SA2"):  # :890-1066: Rational{Int64} tables (oracle/kc2019_tables.json)
        t = _kc2019()[name]
        q = lambda p: p[0] / p[1]  # Float64(num // den): both < 2^53, so this is the correctly rounded quotient
        b = [q(p) for p in t["b"]]
        c = [q(p) for p in t["c"]]
        return IMEXTableau(a_exp=[[q(p) for p in r] for r in t["a_exp"]], a_imp=[[q(p) for p in r] for r in t["a_imp"]],
                           b_exp=b, b_imp=b, c_exp=c, c_imp=c)
    if name in _EXPLICIT:  # explicit_tableaus.jl:82-124; IMEXTableau((; a, b, c)) = IMEXTableau(a, b, c, a, b, c) (:75)
        a, b = _EXPLICIT[name]
        return IMEXTableau(a_exp=a, a_imp=a, b_exp=b, b_imp=b)
    raise KeyError(name)


def _imkg_tableau(al, alh, dh, be=None):
    """IMKGTableau(α, α̂, δ̂, β) (imex_tableaus.jl:375-389): s = length(α̂) + 1;
    a_exp[i,j] = α[j] on the first sub-diagonal, β[i-2] in the first column of rows i > 2;
    a_imp likewise with α̂, plus δ̂[i-1] on the diagonal of rows 1 < i <= length(α̂)  (1-based indices)."""
    s = len(alh) + 1
    be = (0.0,) * len(dh) if be is None else be
    ae, ai = np.zeros((s, s)), np.zeros((s, s))
    for i in range(1, s + 1):
        for j in range(1, s + 1):
            if i == j + 1:
                ae[i - 1, j - 1] = al[j - 1]
                ai[i - 1, j - 1] = alh[j - 1]
            elif i > 2 and j == 1:
                ae[i - 1, j - 1] = be[i - 3]
                ai[i - 1, j - 1] = be[i - 3]
            elif 1 < i <= len(alh) and i == j:
                ai[i - 1, j - 1] = dh[i - 2]
    return IMEXTableau(a_exp=ae, a_imp=ai)


_R2, _R3 = math.sqrt(2.0), math.sqrt(3.0)
_A5 = (1 / 4, 1 / 6, 3 / 8, 1 / 2, 1)  # α of the five-stage IMKG25x family
_IMKG = {  # name -> (α, α̂, δ̂[, β]) as written at imex_tableaus.jl:399-607
    "IMKG232a": lambda: ((1 / 2, 1 / 2, 1), (0, -1 / 2 + _R2 / 2, 1), (1 - _R2 / 2, 1 - _R2 / 2)),
    "IMKG232b": lambda: ((1 / 2, 1 / 2, 1), (0, -1 / 2 - _R2 / 2, 1), (1 + _R2 / 2, 1 + _R2 / 2)),
    "IMKG242a": lambda: ((1 / 4, 1 / 3, 1 / 2, 1), (0, 0, -1 / 2 + _R2 / 2, 1), (0, 1 - _R2 / 2, 1 - _R2 / 2)),
    "IMKG242b": lambda: ((1 / 4, 1 / 3, 1 / 2, 1), (0, 0, -1 / 2 - _R2 / 2, 1), (0, 1 + _R2 / 2, 1 + _R2 / 2)),
    "IMKG243a": lambda: ((1 / 4, 1 / 3, 1 / 2, 1), (0, 1 / 6, -_R3 / 6, 1), (1 / 2 + _R3 / 6,) * 3),
    "IMKG252a": lambda: (_A5, (0, 0, 0, -1 / 2 + _R2 / 2, 1), (0, 0, 1 - _R2 / 2, 1 - _R2 / 2)),
    "IMKG252b": lambda: (_A5, (0, 0, 0, -1 / 2 - _R2 / 2, 1), (0, 0, 1 + _R2 / 2, 1 + _R2 / 2)),
    "IMKG253a": lambda: (_A5, (0, 0, _R3 / 4 * (1 - _R3 / 3) * ((1 + _R3 / 3) ** 2 - 2), _R3 / 6, 1),
                         (0, 1 / 2 - _R3 / 6, 1 / 2 - _R3 / 6, 1 / 2 - _R3 / 6)),
    "IMKG253b": lambda: (_A5, (0, 0, _R3 / 4 * (1 + _R3 / 3) * ((1 - _R3 / 3) ** 2 - 2), -_R3 / 6, 1),
                         (0, 1 / 2 + _R3 / 6, 1 / 2 + _R3 / 6, 1 / 2 + _R3 / 6)),
    "IMKG254a": lambda: (_A5, (0, -3 / 10, 5 / 6, -3 / 2, 1), (-1 / 2, 1, 1, 2)),
    "IMKG254b": lambda: (_A5, (0, -1 / 20, 5 / 4, -1 / 2, 1), (-1 / 2, 1, 1, 1)),
    "IMKG254c": lambda: (_A5, (0, 1 / 20, 5 / 36, 1 / 3, 1), (1 / 6, 1 / 6, 1 / 6, 1 / 6)),
    "IMKG342a": lambda: ((1 / 4, 2 / 3, 1 / 3, 3 / 4), (0, 1 / 6 - _R3 / 6, -1 / 6 - _R3 / 6, 3 / 4),
                         (0, 1 / 2 + _R3 / 6, 1 / 2 + _R3 / 6), (0, 1 / 3, 1 / 4)),
    "IMKG343a": lambda: ((1 / 4, 2 / 3, 1 / 3, 3 / 4), (0, -1 / 3, -2 / 3, 3 / 4), (-1 / 3, 1, 1), (0, 1 / 3, 1 / 4)),
}
_EXPLICIT = {  # name -> (a, b)
    "SSP22Heuns": ([[0, 0], [1, 0]], [1 / 2, 1 / 2]),
    "SSP33ShuOsher": ([[0, 0, 0], [1, 0, 0], [1 / 4, 1 / 4, 0]], [1 / 6, 1 / 6, 2 / 3]),
    "RK4": ([[0, 0, 0, 0], [1 / 2, 0, 0, 0], [0, 1 / 2, 0, 0], [0, 0, 1, 0]], [1 / 6, 1 / 3, 1 / 3, 1 / 6]),
}
_KC2019 = None


def _kc2019():
    global _KC2019
    if _KC2019 is None:
        import json
        import os
        with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "kc2019_tables.json")) as fh:
            _KC2019 = json.load(fh)
    return _KC2019


IMEX_NAMES = ("ARS111", "ARS121", "ARS122", "ARS232", "ARS222", "IMKG232a", "IMKG232b", "IMKG242a", "IMKG242b", "IMKG243a",
              "IMKG252a", "IMKG252b", "IMKG253a", "IMKG253b", "IMKG254a", "IMKG254b", "IMKG254c", "HOMMEM1", "ARS233",
              "ARS343", "ARS443", "IMKG342a", "IMKG343a", "DBM453", "ARK2GKC", "ARK437L2SA1", "ARK548L2SA2", "SSP222",
              "SSP322", "SSP332", "SSP333", "SSP433")  # test/solvers/imex_ark.jl:11-44
EXPLICIT_NAMES = ("SSP22Heuns", "SSP33ShuOsher", "RK4")  # test/solvers/explicit_rk.jl:51-55
# imex_convergence_orders (test/utils/convergence_utils.jl:20-49): (implicit only, explicit only, both)
CONVERGENCE_ORDERS = {
    "ARS111": (1, 1, 1), "ARS121": (1, 1, 1), "ARS122": (2, 2, 2), "ARS222": (2, 2, 2), "ARS232": (2, 3, 2),
    "ARS233": (3, 3, 3), "ARS343": (3, 4, 3), "ARS443": (3, 3, 3), "IMKG232a": (2, 2, 2), "IMKG232b": (2, 2, 2),
    "IMKG242a": (2, 4, 2), "IMKG242b": (2, 4, 2), "IMKG243a": (2, 4, 2), "IMKG252a": (2, 2, 2), "IMKG252b": (2, 2, 2),
    "IMKG253a": (2, 2, 2), "IMKG253b": (2, 2, 2), "IMKG254a": (2, 2, 2), "IMKG254b": (2, 2, 2), "IMKG254c": (2, 2, 2),
    "IMKG342a": (3, 4, 3), "IMKG343a": (3, 4, 3), "SSP222": (2, 2, 2), "SSP322": (2, 2, 2), "SSP332": (2, 3, 2),
    "SSP333": (3, 3, 3), "SSP433": (3, 3, 3), "DBM453": (3, 3, 3), "HOMMEM1": (2, 3, 2), "ARK2GKC": (2, 2, 2),
    "ARK437L2SA1": (4, 4, 4), "ARK548L2SA2": (5, 5, 5), "SSP22Heuns": (2, 2, 2), "SSP33ShuOsher": (3, 3, 3),
    "RK4": (4, 4, 4), "SSPKnoth": (2, 3, 2),
}
SSP_NAMES = ("SSP222", "SSP322", "SSP332", "SSP333", "SSP433")  # default_constraint = SSP (:646)


def lsrk54_tableau(dtype=np.float64):
    """lsrk.jl:117-143 -- ``RT(num // den)`` = RT(num)/RT(den) (Base rational.jl)."""
    RT = np.dtype(dtype).type
    q = lambda n, d: RT(RT(n) / RT(d))
    A = (RT(0), q(-567301805773, 1357537059087), q(-2404267990393, 2016746695238),
         q(-3550918686646, 2091501179385), q(-1275806237668, 842570457699))
    B = (q(1432997174477, 9575080441755), q(5161836677717, 13612068292357), q(1720146321549, 2090206949498),
         q(3134564353537, 4481467310338), q(2277821191437, 14882151754819))
    C = (RT(0), q(1432997174477, 9575080441755), q(2526269341429, 6820363962896),
         q(2006345519317, 3224310063776), q(2802321613138, 2924317926251))
    return A, B, C


# --------------------------------------------------------------------------------------
# fused increments  (src/utilities/fused_increment.jl)
# --------------------------------------------------------------------------------------
def _wide(x):
    """Promote a state array to Float64 for mixed expressions (Float64 coefficient * Float32 array)."""
    return x.astype(np.float64, copy=False)


def _jl(x):
    """A scalar with its Julia type: a Float32 stays Float32, anything else is Float64 (numpy >= 2 promotes typed
    scalars the way Julia does: Float64*Float32 -> Float64, Float32*Float32 -> Float32)."""
    return x if isinstance(x, np.float32) else np.float64(x)


def _scale(coef, T):
    """``coef .* T`` with Julia promotion: Float32 only when both are Float32."""
    if isinstance(coef, np.float32) and T.dtype == np.float32:
        return coef * T
    return np.float64(coef) * _wide(T)


def downcast_tableau(FT, tab: "IMEXTableau"):
    """imex_tableaus.jl:127-133: every coefficient array mapped through FT (used when cast_tableau_to_state_eltype)."""
    import copy
    t = copy.copy(tab)
    for k in ("a_exp", "b_exp", "c_exp", "a_imp", "b_imp", "c_imp"):
        setattr(t, k, getattr(tab, k).astype(FT))
    return t


def fused_raw_increment(dt, coeffs: Sequence[float], tends: Sequence[Optional[np.ndarray]], cdtype=None):
    """fused_increment.jl:115-172.  ``coeffs[j]`` pairs with ``tends[j]``; zero coefficients are dropped
    (sparse_coeffs.jl:29-35).  Each term is ``(dt*c_j) .* T_j`` with the scalar product rounded first (:58,:79) in
    ``promote(typeof(dt), eltype(tableau))``; terms are summed by a left fold (:90-91).  Returns ``None`` for
    NullBroadcasted (:127-128,:155-156).  The tableau eltype is the dtype of ``coeffs`` (Float64 unless
    cast_tableau_to_state_eltype downcast it, imex_ark.jl:59-61)."""
    acc = None
    dt = _jl(dt)
    for c, T in zip(coeffs, tends):
        if c == 0:
            continue
        term = _scale(dt * _jl(c), T)
        acc = term if acc is None else acc + term
    return acc


def _axpy_dt(x, dt, T):
    """``x + dt*T`` with Julia promotion: a Float32 ``dt`` keeps Float32 arithmetic, anything else is Float64."""
    if isinstance(dt, np.float32) and x.dtype == np.float32:
        return x + dt * T
    return _wide(x) + float(dt) * _wide(T)


def assign_with_increments(U, base, inc_exp, inc_imp):
    """fused_increment.jl:220-236: ``U .= (base + inc_exp) + inc_imp`` with Null handling."""
    if inc_exp is None and inc_imp is None:
        U[...] = base
    elif inc_exp is None:
        U[...] = _wide(base) + inc_imp if inc_imp.dtype == np.float64 else base + inc_imp
    elif inc_imp is None:
        U[...] = _wide(base) + inc_exp if inc_exp.dtype == np.float64 else base + inc_exp
    else:
        b = _wide(base) if inc_exp.dtype == np.float64 else base
        U[...] = (b + inc_exp) + inc_imp


def assign_fused_increment(U, u, dt, coeffs, tends, cdtype=np.float64):
    """fused_increment.jl:201-205: ``@. U = u + inc`` (``u`` alone when inc is Null)."""
    inc = fused_raw_increment(dt, coeffs, tends, cdtype)
    if inc is None:
        U[...] = u
    else:
        U[...] = (_wide(u) if inc.dtype == np.float64 else u) + inc


# --------------------------------------------------------------------------------------
# Newton / Krylov  (src/nl_solvers/newtons_method.jl)
# --------------------------------------------------------------------------------------
_RED_THREADS = 256      # csrc/cts_internal.h: kReduceThreads
REDUCTION_CHUNK = 8192  # csrc/cts_internal.h: kReduceChunk (cts_set_reduction_layout may choose another power of two)


def _warp_tree(a):
    """xor-butterfly over the last axis (32 lanes): every level adds lane l and lane l^o (commutative, so all
    lanes agree); equivalent to repeated halving."""
    for o in (16, 8, 4, 2, 1):
        a = a[..., :o] + a[..., o:2 * o]
    return a[..., 0]


def _block_tree(acc):
    """[..., 256] per-thread values -> one value per block: warp shuffles, then warp 0 folds the 8 warp sums."""
    w = _warp_tree(acc.reshape(acc.shape[:-1] + (_RED_THREADS // 32, 32)))  # [..., 8]
    w = w[..., :4] + w[..., 4:8]
    w = w[..., :2] + w[..., 2:4]
    return w[..., 0] + w[..., 1]


def tree_sum(terms: np.ndarray, vn: int = 2, chunk: Optional[int] = None) -> float:
    """The library's deterministic reduction order (csrc/reduce.cu), restated so that norms and dots --
    and everything downstream of them (JVP step sizes, GMRES) -- are bit-identical on both sides.
    LinearAlgebra.norm's own order is implementation defined, so this fixed tree IS the
    specification.  It is defined on the GLOBAL element index, hence independent of any sharding: the domain is
    cut into chunks of `chunk` elements; within a chunk thread t of 256 accumulates (from 0) the 16-byte vectors
    t, t+256, ... (elements of a vector in order), then warp/block trees give the chunk partial; the chunk
    partials are folded in order by one block (thread t takes partials t, t+256, ...) with the same tree.
    `terms` are the Float64 summands; `vn` = elements per 16-byte vector (2 for Float64, 4 for Float32)."""
    chunk = int(chunk or REDUCTION_CHUNK)
    v = np.ascontiguousarray(terms, dtype=np.float64).ravel()
    n = v.size
    nc = max(1, -(-n // chunk))
    rounds = max(1, -(-(chunk // vn) // _RED_THREADS))
    per = rounds * _RED_THREADS * vn            # padded chunk (absent elements add +0.0, which is exact)
    buf = np.zeros((nc, per))
    full = n // chunk
    if full:
        buf[:full, :chunk] = v[:full * chunk].reshape(full, chunk)
    if n - full * chunk:
        buf[full, :n - full * chunk] = v[full * chunk:]
    buf = buf.reshape(nc, rounds, _RED_THREADS, vn)
    acc = np.zeros((nc, _RED_THREADS))
    for r in range(rounds):
        for e in range(vn):
            acc = acc + buf[:, r, :, e]
    partials = _block_tree(acc)  # [nc]
    fin = np.zeros(_RED_THREADS)
    for b0 in range(0, nc, _RED_THREADS):
        c = partials[b0:b0 + _RED_THREADS]
        fin[:c.size] = fin[:c.size] + c
    return float(_block_tree(fin))


REDUCTION_WINDOW = None  # (offset, count): restrict norms/dots to owned elements (cts_set_reduction_window)


def _vn(x):
    return 16 // x.dtype.itemsize


def _win(x):
    x = x.ravel()
    if REDUCTION_WINDOW is None:
        return x
    off, cnt = REDUCTION_WINDOW
    return x[off:off + cnt]


def norm2(x):
    xd = _wide(_win(x))
    return math.sqrt(tree_sum(xd * xd, _vn(x)))


def dot(x, y):
    return tree_sum(_wide(_win(x)) * _wide(_win(y)), _vn(x))


def sum1pabs(x):
    return tree_sum(1.0 + np.abs(_wide(_win(x))), _vn(x))


@dataclass
class ForwardDiffJVP:
    """newtons_method.jl:165-182; step sizes :107-135 (1, 2 or 3; default 3)."""

    default_step: int = 3
    step_adjustment: float = 1

    def step(self, dx, x):
        FT = x.dtype.type
        eps = float(np.finfo(x.dtype).eps)
        if self.default_step == 1:
            e = math.sqrt(eps) / norm2(dx)
        elif self.default_step == 2:
            e = math.sqrt(eps * (1 + norm2(x))) / norm2(dx)
        else:
            e = math.sqrt(eps) * sum1pabs(x) / (_win(x).size * norm2(dx))
        return FT(FT(self.step_adjustment) * FT(e))


@dataclass
class ConstantForcing:  # :211-220
    rtol: float = 0.0

    def get_rtol(self, cache, f, n):
        return float(f.dtype.type(self.rtol))


@dataclass
class EisenstatWalkerForcing:  # :248-280
    initial_rtol: float = 0.5
    gamma: float = 1
    alpha: float = 2
    min_rtol_threshold: float = 0.1
    max_rtol: float = 0.9

    def get_rtol(self, cache, f, n):
        norm_f = norm2(f)
        if n == 0:
            rtol = self.initial_rtol
        else:
            rtol = self.gamma * (norm_f / cache["prev_norm_f"]) ** self.alpha
            min_rtol = self.gamma * cache["prev_rtol"] ** self.alpha
            if min_rtol > self.min_rtol_threshold:
                rtol = max(rtol, min_rtol)
        rtol = min(rtol, self.max_rtol)
        cache["prev_norm_f"], cache["prev_rtol"] = norm_f, rtol
        return rtol


@dataclass
class KrylovMethod:  # :419-438
    jacobian_free_jvp: Optional[ForwardDiffJVP] = None
    forcing_term: Any = field(default_factory=ConstantForcing)
    memory: int = 20
    disable_preconditioner: bool = False
    itmax: int = 0


@dataclass
class NewtonsMethod:  # :568-581
    max_iters: int = 1
    update_j: Any = field(default_factory=lambda: UpdateEvery("NewNewtonIteration"))
    krylov_method: Optional[KrylovMethod] = None
    convergence_checker: Any = None
    line_search: Any = None


# --------------------------------------------------------------------------------------
# ConvergenceChecker / ConvergenceCondition / LineSearch
# (src/utilities/convergence_condition.jl:41-143, convergence_checker.jl:32-108, line_search.jl:23-47)
# --------------------------------------------------------------------------------------
class ConvergenceCondition:
    has_cache = False

    def needs_cache_update(self, it):
        return False


@dataclass
class MaximumError(ConvergenceCondition):  # :41-44
    max_err: float

    def has_converged(self, cache, val, err, it):
        return err <= self.max_err


@dataclass
class MaximumRelativeError(ConvergenceCondition):  # :52-56
    max_rel_err: float

    def has_converged(self, cache, val, err, it):
        return err <= self.max_rel_err * val


@dataclass
class MaximumErrorReduction(ConvergenceCondition):  # :66-75
    max_reduction: float
    has_cache = True

    def has_converged(self, cache, val, err, it):
        return (err <= cache) if it >= 1 else (err != err) & False  # `iter >= 1 && err <= cache.max_err`

    def needs_cache_update(self, it):
        return it == 0

    def updated_cache(self, cache, val, err, it):
        return self.max_reduction * err


@dataclass
class MinimumRateOfConvergence(ConvergenceCondition):  # :96-108
    rate: float
    order: float = 1
    has_cache = True

    def has_converged(self, cache, val, err, it):
        return (err >= cache) if it >= 1 else (err != err) & False

    def needs_cache_update(self, it):
        return True

    def updated_cache(self, cache, val, err, it):
        return self.rate * err ** self.order


class MultipleConditions(ConvergenceCondition):  # :116-143
    has_cache = True

    def __init__(self, *args):
        if args and args[0] in (all, any):
            self.combiner, self.conditions = args[0], tuple(args[1:])
        else:
            self.combiner, self.conditions = all, tuple(args)

    def has_converged(self, caches, val, err, it):
        res = [c.has_converged(k, val, err, it) for c, k in zip(self.conditions, caches)]
        if isinstance(res[0], np.ndarray) or any(isinstance(r, np.ndarray) for r in res):
            res = [np.broadcast_to(np.asarray(r, dtype=bool), np.shape(err)) for r in res]
            return np.logical_and.reduce(res) if self.combiner is all else np.logical_or.reduce(res)
        return self.combiner(bool(r) for r in res)

    def needs_cache_update(self, it):
        return any(c.needs_cache_update(it) for c in self.conditions)

    def updated_cache(self, caches, val, err, it):
        return tuple(c.updated_cache(k, val, err, it) if c.needs_cache_update(it) else k
                     for c, k in zip(self.conditions, caches))


def _empty_cond_cache(cond, proto):
    """`Ref{cache_type(...)}()` / `similar(val_prototype, cache_type(...))`: uninitialised in the reference; NaN here, so a
    condition that reads its cache before the first update is false (documented deviation from undefined behaviour)."""
    if isinstance(cond, MultipleConditions):
        return tuple(_empty_cond_cache(c, proto) for c in cond.conditions)
    if not cond.has_cache:
        return None
    return np.nan if proto is None else np.full_like(proto, np.nan)


@dataclass
class ConvergenceChecker:  # convergence_checker.jl:32-43
    norm_condition: Any = None
    component_condition: Any = None
    condition_combiner: str = "&"
    norm: Callable = None

    def allocate_cache(self, val_prototype):  # :45-60
        if self.norm_condition is None and self.component_condition is None:
            raise ValueError("ConvergenceChecker must have at least one ConvergenceCondition")
        return {"norm_cache": None if self.norm_condition is None else _empty_cond_cache(self.norm_condition, None),
                "component_cache": None if self.component_condition is None
                else _empty_cond_cache(self.component_condition, val_prototype)}


def is_converged(alg: Optional[ConvergenceChecker], cache, val, err, it) -> bool:
    """is_converged!, convergence_checker.jl:75-108."""
    if alg is None:
        return False
    nc, cc = alg.norm_condition, alg.component_condition
    nrm = alg.norm if alg.norm is not None else norm2

    def comp():
        r = cc.has_converged(cache["component_cache"], np.abs(val), np.abs(err), it)
        return bool(np.all(np.broadcast_to(np.asarray(r, dtype=bool), np.shape(val))))

    if nc is None:
        converged = comp()
    else:
        norm_val, norm_err = nrm(val), nrm(err)
        converged = bool(nc.has_converged(cache["norm_cache"], norm_val, norm_err, it))
        if cc is not None:
            converged = (converged and comp()) if alg.condition_combiner == "&" else (converged or comp())
    if not converged:
        if nc is not None and nc.needs_cache_update(it):
            cache["norm_cache"] = nc.updated_cache(cache["norm_cache"], norm_val, norm_err, it)
        if cc is not None and cc.needs_cache_update(it):
            cache["component_cache"] = cc.updated_cache(cache["component_cache"], np.abs(val), np.abs(err), it)
    return converged


@dataclass
class LineSearch:  # line_search.jl:23-25
    norm: Callable = None


def line_search(alg: Optional[LineSearch], x, dx, f, f_, prepare_for_f):
    """line_search!, line_search.jl:27-47: x already holds x_old - Δx, f holds f(x_old)."""
    if alg is None:
        return
    nrm = alg.norm if alg.norm is not None else norm2
    normf0 = nrm(f)
    if prepare_for_f is not None:
        prepare_for_f(x)
    f_(f, x)
    normf = nrm(f)
    i, alpha = 1, 0.5
    while ((normf > normf0) or not math.isfinite(normf)) and i <= 5:
        x[...] = (x.astype(np.float64) + alpha * dx.astype(np.float64)).astype(x.dtype)  # `x .+= α * Δx`
        if prepare_for_f is not None:
            prepare_for_f(x)
        f_(f, x)
        normf = nrm(f)
        alpha /= 2
        i += 1


def sym_givens(a: float, b: float):
    """Krylov.jl ``sym_givens`` (krylov_utils.jl, v0.10.x; real case): returns (c, s, rho) with
    [c s; s -c] [a; b] = [rho; 0].  Restated from the published algorithm (Choi's SymGivens2)."""
    if b == 0:
        c = 1.0 if a == 0 else math.copysign(1.0, a)
        return c, 0.0, abs(a)
    if a == 0:
        return 0.0, math.copysign(1.0, b), abs(b)
    if abs(b) > abs(a):
        t = a / b
        s = math.copysign(1.0, b) / math.sqrt(1 + t * t)
        c = s * t
        return c, s, b / s
    t = b / a
    c = math.copysign(1.0, a) / math.sqrt(1 + t * t)
    s = c * t
    return c, s, a / c


def gmres(matvec, b, *, precond=None, rtol=0.0, atol=0.0, memory=20, itmax=0, stats=None):
    """Krylov.jl v0.10 ``gmres!`` as called at newtons_method.jl:490 (left preconditioner M via
    ``ldiv!``, N = I, restart = false, reorthogonalization = false, itmax = 0 -> 2n).  Restated from
    the published algorithm: MGS Arnoldi, Givens on the Hessenberg column,
    stop on preconditioned residual ``rNorm <= atol + rtol*||M^-1 b||`` or ``rNorm + 1 <= 1``,
    breakdown when ``Hbis <= eps^(3/4)``; back substitution skips pivots <= btol (inconsistent).
    ``matvec(w, v)`` writes A*v into w; ``precond(q, w)`` writes M^-1 w into q.  Vectors beyond
    ``memory`` are appended (no restart)."""
    FT = b.dtype.type
    n = b.size
    x = np.zeros_like(b)
    w = b.copy()
    if precond is not None:
        r0 = np.zeros_like(b)
        precond(r0, w)
    else:
        r0 = w
    beta = norm2(r0)
    st = {"niter": 0, "solved": False, "inconsistent": False, "status": "unknown"}
    if stats is not None:
        stats.update(st)
    if beta == 0:
        st.update(solved=True, status="x = 0 is a zero-residual solution")
        if stats is not None:
            stats.update(st)
        return x
    rNorm = beta
    eps_tol = atol + rtol * beta
    btol = float(np.finfo(b.dtype).eps) ** 0.75
    itmax = 2 * n if itmax == 0 else itmax
    V = [(_wide(r0) / rNorm).astype(b.dtype)]
    z = [beta]
    cs, ss = [], []
    R = []  # packed upper-triangular columns
    solved = rNorm <= eps_tol
    breakdown = False
    k = 0
    q = np.zeros_like(b)
    while not (solved or k >= itmax or breakdown):
        k += 1
        matvec(w, V[k - 1])
        if precond is not None:
            precond(q, w)
        else:
            q[...] = w
        col = []
        for i in range(k):
            h = dot(V[i], q)
            col.append(h)
            q[...] = _wide(q) - h * _wide(V[i])
        Hbis = norm2(q)
        for i in range(k - 1):  # apply stored reflections
            rtmp = cs[i] * col[i] + ss[i] * col[i + 1]
            col[i + 1] = ss[i] * col[i] - cs[i] * col[i + 1]
            col[i] = rtmp
        c, s, rho = sym_givens(col[k - 1], Hbis)
        col[k - 1] = rho
        cs.append(c)
        ss.append(s)
        R.append(col)
        zeta = s * z[k - 1]
        z[k - 1] = c * z[k - 1]
        rNorm = abs(zeta)
        solved = rNorm <= eps_tol or rNorm + 1 <= 1
        breakdown = Hbis <= btol
        if not (solved or k >= itmax or breakdown):
            V.append((_wide(q) / Hbis).astype(b.dtype))
            z.append(zeta)
    # back substitution R y = z
    y = [0.0] * k
    inconsistent = False
    for i in range(k - 1, -1, -1):
        acc = z[i]
        for j in range(i + 1, k):
            acc -= R[j][i] * y[j]
        if abs(R[i][i]) <= btol:
            y[i] = 0.0
            inconsistent = True
        else:
            y[i] = acc / R[i][i]
    xd = np.zeros(n, dtype=np.float64).reshape(b.shape)
    for i in range(k):
        xd = xd + y[i] * _wide(V[i])
    x[...] = xd
    st.update(niter=k, solved=solved and not inconsistent, inconsistent=inconsistent,
              status="solved" if solved else ("breakdown" if breakdown else "maxiter"))
    if stats is not None:
        stats.update(st)
    return x


class DenseJacobian:
    """``j isa DenseMatrix`` path: ``ldiv!(dx, lu(j), f)`` (newtons_method.jl:630-631)."""

    def __init__(self, n, dtype=np.float64):
        self.M = np.zeros((n, n), dtype=dtype)

    def zero(self):
        return DenseJacobian(self.M.shape[0], self.M.dtype)

    def ldiv(self, dx, f):
        dx[...] = np.linalg.solve(self.M, f)

    def mul(self, out, v):
        out[...] = self.M @ v


def allocate_newton_cache(alg: NewtonsMethod, x_prototype, j_prototype=None):
    """newtons_method.jl:583-598."""
    km = alg.krylov_method
    assert not (j_prototype is None and (km is None or km.jacobian_free_jvp is None))
    cache = {"dx": np.zeros_like(x_prototype), "f": np.zeros_like(x_prototype),
             "j": None if j_prototype is None else j_prototype.zero(), "stats": [],
             "checker": None if alg.convergence_checker is None else alg.convergence_checker.allocate_cache(x_prototype)}
    if km is not None:
        cache["forcing"] = {}
        if km.jacobian_free_jvp is not None:
            cache["x2"] = np.zeros_like(x_prototype)
            cache["f2"] = np.zeros_like(x_prototype)
    return cache


def solve_newton(alg: NewtonsMethod, cache, x, f_, j_=None, prepare_for_f=None):
    """newtons_method.jl:609-665."""
    if cache is None:
        return
    dx, f, j = cache["dx"], cache["f"], cache["j"]
    km = alg.krylov_method
    if j is not None and alg.update_j.needs_update("NewNewtonSolve"):
        j_(j, x)
    f_(f, x)
    for n in range(1, alg.max_iters + 1):
        if j is not None and alg.update_j.needs_update("NewNewtonIteration"):
            j_(j, x)
        if km is None:
            j.ldiv(dx, f)
        else:
            _solve_krylov(km, cache, dx, x, f_, f, n - 1, prepare_for_f, j)
        x[...] = x - dx  # :651
        line_search(alg.line_search, x, dx, f, f_, prepare_for_f)  # :652
        if is_converged(alg.convergence_checker, cache.get("checker"), x, dx, n):  # :656
            cache["newton_iters"] = cache.get("newton_iters", 0) + n
            break
        elif n < alg.max_iters and alg.line_search is None:  # :658-660
            if prepare_for_f is not None:
                prepare_for_f(x)
            f_(f, x)
    else:
        cache["newton_iters"] = cache.get("newton_iters", 0) + alg.max_iters


def _solve_krylov(km: KrylovMethod, cache, dx, x, f_, f, n, prepare_for_f, j):
    """newtons_method.jl:465-509."""
    jvp = km.jacobian_free_jvp

    def matvec(out, v):
        if jvp is None:
            j.mul(out, v)
        else:  # jvp! :173-182
            x2, f2 = cache["x2"], cache["f2"]
            e = jvp.step(v, x)
            x2[...] = x + e * v
            if prepare_for_f is not None:
                prepare_for_f(x2)
            f_(f2, x2)
            out[...] = (f2 - f) / e

    use_M = not (km.disable_preconditioner or j is None or jvp is None)
    precond = (lambda q, w: j.ldiv(q, w)) if use_M else None
    rtol = km.forcing_term.get_rtol(cache["forcing"], f, n)
    st = {}
    sol = gmres(matvec, f, precond=precond, rtol=rtol, atol=0.0, memory=km.memory, itmax=km.itmax, stats=st)
    cache["stats"].append(st)
    dx[...] = sol


# --------------------------------------------------------------------------------------
# algorithms + caches
# --------------------------------------------------------------------------------------
@dataclass
class IMEXAlgorithm:
    """imex_tableaus.jl:112-160.  constraint: 'Unconstrained' | 'SSP'."""

    tableau: IMEXTableau
    newtons_method: Optional[NewtonsMethod]
    constraint: str = "Unconstrained"
    name: Optional[str] = None
    cast_tableau_to_state_eltype: bool = False


def imex_algorithm(name, newtons_method, constraint=None, **kw):
    tab = tableau(name, **{k: v for k, v in kw.items() if k in ("beta", "paper_version")})
    if constraint is None:  # default_constraint: SSP for IMEXSSPRK and SSPRK names (imex_tableaus.jl:646, explicit_tableaus.jl:70)
        constraint = "SSP" if name in SSP_NAMES or name in ("SSP22Heuns", "SSP33ShuOsher") else "Unconstrained"
    return IMEXAlgorithm(tab, newtons_method, constraint, name,
                         kw.get("cast_tableau_to_state_eltype", False))


def _sdirk_gamma(a_imp):
    d = [float(v) for v in np.diag(a_imp) if v != 0]
    uniq = []
    for v in d:
        if v not in uniq:
            uniq.append(v)
    return uniq[0] if len(uniq) == 1 else None


def init_cache_ark(u0, f: ClimaODEFunction, alg: IMEXAlgorithm):
    """imex_ark.jl:35-66."""
    tab = alg.tableau
    s = tab.s
    T_imp = f.T_imp
    idx_exp = [i for i in range(s) if np.any(tab.a_exp[:, i] != 0) or tab.b_exp[i] != 0]
    idx_imp = [i for i in range(s) if np.any(tab.a_imp[:, i] != 0) or tab.b_imp[i] != 0]
    z = lambda: np.zeros_like(u0)
    gamma = _sdirk_gamma(tab.a_imp)
    nm = alg.newtons_method
    if gamma is None and nm is not None and _fires_on_new_timestep(nm.update_j):  # async_utils.jl:93-98
        raise ValueError("tableau is not SDIRK: do not update on the NewTimeStep signal")
    has_jac = T_imp is not None and getattr(T_imp, "Wfact", None) is not None and T_imp.jac_prototype is not None
    nmc = None
    if T_imp is not None and nm is not None:
        nmc = allocate_newton_cache(nm, u0, T_imp.jac_prototype if has_jac else None)
    opt_tb = downcast_tableau(u0.dtype, tab) if alg.cast_tableau_to_state_eltype else tab  # imex_ark.jl:59-61
    return {"U": z(), "T_lim": {i: z() for i in idx_exp}, "T_exp": {i: z() for i in idx_exp},
            "T_imp": {i: z() for i in idx_imp}, "temp": z(), "gamma": gamma, "nm_cache": nmc, "tableau": opt_tb}


def _maybe_update_jacobian(f, alg, cache, u, p, t, dt):
    """async_utils.jl:11-31."""
    T_imp, nm, nmc = f.T_imp, alg.newtons_method, cache["nm_cache"]
    if T_imp is None or nm is None or nmc is None or nmc["j"] is None:
        return
    if not nm.update_j.needs_update("NewTimeStep", t):
        return
    T_imp.Wfact(nmc["j"], u, p, float(dt) * cache["gamma"], t)  # γ comes from the tableau as given (imex_ark.jl:50-52)


def _solve_implicit_equation(U, U0, p, t_imp, dtg, T, nm, nmc, cache_imp):
    """imex_ark.jl:283-308: residual = (U0 + dtγ*T(U')) - U'."""

    def residual(res, Up):
        T(res, Up, p, float(t_imp))
        if isinstance(dtg, np.float32) and res.dtype == np.float32:
            res[...] = (U0 + dtg * res) - Up  # dtγ in Float32 (Float32 dt and a downcast tableau): Float32 broadcast
        else:
            res[...] = (_wide(U0) + float(dtg) * _wide(res)) - _wide(Up)

    solve_newton(nm, nmc, U, residual,
                 (lambda jac, Up: T.Wfact(jac, Up, p, float(dtg), float(t_imp))) if getattr(T, "Wfact", None) else None,
                 lambda Up: cache_imp(Up, p, float(t_imp)))


def _timp_from_stage(U, temp, dtg):
    """``@. T_imp[i] = (U - temp) / dtγ`` (imex_ark.jl:269): the subtraction in the state precision, the division in
    promote(eltype(U), typeof(dtγ))."""
    if isinstance(dtg, np.float32) and U.dtype == np.float32:
        return (U - temp) / dtg
    return _wide(U - temp) / float(dtg)


def imex_ark_step(u, p, t, dt, f: ClimaODEFunction, alg: IMEXAlgorithm, cache):
    """imex_ark.jl:69-104 (step_u!) with the stage loop of :107-272 inlined."""
    tab = cache["tableau"]
    s = tab.s
    U, T_lim, T_exp, T_imp, temp = cache["U"], cache["T_lim"], cache["T_exp"], cache["T_imp"], cache["temp"]
    cd = None
    Timp_f = f.T_imp
    nm, nmc = alg.newtons_method, cache["nm_cache"]
    _maybe_update_jacobian(f, alg, cache, u, p, t, dt)
    tend = lambda d, n: [d.get(j) for j in range(n)]
    for i in range(s):
        t_exp = float(_jl(t) + _jl(dt) * tab.c_exp[i])
        t_imp = float(_jl(t) + _jl(dt) * tab.c_imp[i])
        dtg = _jl(dt) * tab.a_imp[i, i]  # `float(dt) * a_imp[i, i]` :122 in promote(typeof(dt), eltype(tableau))
        # compute_stage_value! :140-165
        inc_exp = fused_raw_increment(dt, tab.a_exp[i, :i], tend(T_exp, i), cd) if f.has_T_exp else None
        inc_imp = None if Timp_f is None else fused_raw_increment(dt, tab.a_imp[i, :i], tend(T_imp, i), cd)
        if f.has_T_lim:
            assign_fused_increment(U, u, dt, tab.a_exp[i, :i], tend(T_lim, i), cd)
            if i != 0:
                f.lim(U, p, t_exp, u)
            assign_with_increments(U, U, inc_exp, inc_imp)
        else:
            assign_with_increments(U, u, inc_exp, inc_imp)
        # solve_stage_implicit! :174-242
        no_imp = Timp_f is None or dtg == 0
        top_sig = "EndOfStage" if no_imp else "WithDSS"
        if i != 0:
            f.dss(U, p, t_exp)
            if f.update_constrain_state.needs_update(top_sig):
                f.constrain_state(U, p, t_exp)
            if no_imp and f.update_cache.needs_update("EndOfStage"):
                f.cache(U, p, t_exp)
        if not no_imp:
            assert nm is not None
            temp[...] = U
            if f.initialize_imp is not None:
                f.initialize_imp(U, p, dtg)
                f.dss(U, p, t_exp)
                if f.update_constrain_state.needs_update("WithDSS"):
                    f.constrain_state(U, p, t_exp)
                f.cache_imp(U, p, t_imp)
            elif i != 0:
                f.cache_imp(U, p, t_imp)
            _solve_implicit_equation(U, temp, p, t_imp, dtg, Timp_f, nm, nmc, f.cache_imp)
            f.dss(U, p, t_imp)
            if not (i == s - 1 and tab.is_fsal()):
                if f.update_constrain_state.needs_update("EndOfStage"):
                    f.constrain_state(U, p, t_imp)
                if f.update_cache.needs_update("EndOfStage"):
                    f.cache(U, p, t_imp)
        # evaluate_stage_tendencies! :252-272
        if np.any(tab.a_exp[:, i] != 0) or tab.b_exp[i] != 0:
            if f.T_exp_T_lim is not None:
                f.T_exp_T_lim(T_exp[i], T_lim[i], U, p, t_exp)
        if np.any(tab.a_imp[:, i] != 0) or tab.b_imp[i] != 0:
            if dtg == 0:
                if Timp_f is not None:
                    Timp_f(T_imp[i], U, p, t_imp)
            elif Timp_f is not None:
                T_imp[i][...] = _timp_from_stage(U, temp, dtg)
    # final update :79-91
    t_final = t + dt
    inc_exp = fused_raw_increment(dt, tab.b_exp, tend(T_exp, s), cd) if f.has_T_exp else None
    inc_imp = None if Timp_f is None else fused_raw_increment(dt, tab.b_imp, tend(T_imp, s), cd)
    if f.has_T_lim:
        assign_fused_increment(temp, u, dt, tab.b_exp, tend(T_lim, s), cd)
        f.lim(temp, p, t_final, u)
        assign_with_increments(u, temp, inc_exp, inc_imp)
    else:
        assign_with_increments(u, u, inc_exp, inc_imp)
    f.dss(u, p, t_final)
    if f.update_constrain_state.needs_update("EndOfStep"):
        f.constrain_state(u, p, t_final)
    if f.update_cache.needs_update("EndOfStep"):
        f.cache(u, p, t_final)
    return u


def init_cache_ssprk(u0, f: ClimaODEFunction, alg: IMEXAlgorithm):
    """imex_ssprk.jl:33-93."""
    tab = alg.tableau
    s = tab.s
    idx_imp = [i for i in range(s) if np.any(tab.a_imp[:, i] != 0) or tab.b_imp[i] != 0]
    z = lambda: np.zeros_like(u0)
    a_hat = np.vstack([tab.a_exp, tab.b_exp[None, :]])
    beta = np.array([a_hat[i + 1, i] for i in range(s)])
    for i in range(s):
        if not np.array_equal(a_hat[i + 1:, i], np.cumprod(beta[i:])):
            raise ValueError("not a low-storage IMEX SSPRK tableau (imex_ssprk.jl:51-64)")
    gamma = _sdirk_gamma(tab.a_imp)
    nm = alg.newtons_method
    if gamma is None and nm is not None and _fires_on_new_timestep(nm.update_j):
        raise ValueError("tableau is not SDIRK: do not update on the NewTimeStep signal")
    T_imp = f.T_imp
    has_jac = T_imp is not None and getattr(T_imp, "Wfact", None) is not None and T_imp.jac_prototype is not None
    nmc = None
    if T_imp is not None and nm is not None:
        nmc = allocate_newton_cache(nm, u0, T_imp.jac_prototype if has_jac else None)
    opt_tb = downcast_tableau(u0.dtype, tab) if alg.cast_tableau_to_state_eltype else tab  # imex_ssprk.jl:75-78
    return {"U": z(), "U_exp": z(), "U_lim": z(), "T_lim": z(), "T_exp": z(),
            "T_imp": {i: z() for i in idx_imp}, "temp": z(), "beta": beta, "gamma": gamma,
            "nm_cache": nmc, "tableau": opt_tb}  # β is formed from the tableau as given (:48-49): Float64 for named methods


def imex_ssprk_step(u, p, t, dt, f: ClimaODEFunction, alg: IMEXAlgorithm, cache):
    """imex_ssprk.jl:95-220."""
    tab = cache["tableau"]
    s = tab.s
    U, U_exp, U_lim, T_lim, T_exp = cache["U"], cache["U_exp"], cache["U_lim"], cache["T_lim"], cache["T_exp"]
    T_imp, temp, beta, cd = cache["T_imp"], cache["temp"], cache["beta"], None
    Timp_f, nm, nmc = f.T_imp, alg.newtons_method, cache["nm_cache"]
    _maybe_update_jacobian(f, alg, cache, u, p, t, dt)
    tend = lambda d, n: [d.get(j) for j in range(n)]
    w = _wide            # β is Float64, so every Shu-Osher blend is a Float64 broadcast (rounded on store)
    cdt = float

    def plus_inc(x, inc):  # `x + inc_imp`: Float32 only when the increment is (Float32 dt and a downcast tableau)
        return (x + inc) if (inc.dtype == np.float32 and x.dtype == np.float32) else (_wide(x) + _wide(inc))

    for i in range(s):
        t_exp = float(_jl(t) + _jl(dt) * tab.c_exp[i])
        t_imp = float(_jl(t) + _jl(dt) * tab.c_imp[i])
        dtg = _jl(dt) * tab.a_imp[i, i]
        if i == 0:
            U_exp[...] = u
        elif beta[i - 1] != 0:
            if f.has_T_lim:
                U_lim[...] = _axpy_dt(U_exp, dt, T_lim)  # :158
                f.lim(U_lim, p, t_exp, U_exp)
                U_exp[...] = U_lim
            if f.has_T_exp:
                U_exp[...] = _axpy_dt(U_exp, dt, T_exp)  # :163
            b = cdt(beta[i - 1])
            U_exp[...] = (1 - b) * w(u) + b * w(U_exp)  # :165
        inc_imp = None if Timp_f is None else fused_raw_increment(dt, tab.a_imp[i, :i], tend(T_imp, i), cd)
        if inc_imp is None:
            # `inc_imp = 0` (:168) when T_imp! === nothing; a NullBroadcasted adds nothing
            U[...] = U_exp
        else:
            U[...] = plus_inc(U_exp, inc_imp)
        no_imp = Timp_f is None or tab.a_imp[i, i] == 0
        top_sig = "EndOfStage" if no_imp else "WithDSS"
        if i != 0:
            f.dss(U, p, t_exp)
            if f.update_constrain_state.needs_update(top_sig):
                f.constrain_state(U, p, t_exp)
            if no_imp and f.update_cache.needs_update("EndOfStage"):
                f.cache(U, p, t_exp)
        if not no_imp:
            assert nm is not None
            if i != 0:
                f.cache_imp(U, p, t_imp)
            temp[...] = U
            _solve_implicit_equation(U, temp, p, t_imp, dtg, Timp_f, nm, nmc, f.cache_imp)
            f.dss(U, p, t_imp)
            if f.update_constrain_state.needs_update("EndOfStage"):
                f.constrain_state(U, p, t_imp)
            if f.update_cache.needs_update("EndOfStage"):
                f.cache(U, p, t_imp)
        if beta[i] != 0 and f.T_exp_T_lim is not None:
            f.T_exp_T_lim(T_exp, T_lim, U, p, t_exp)
        if np.any(tab.a_imp[:, i] != 0) or tab.b_imp[i] != 0:
            if tab.a_imp[i, i] == 0:
                if Timp_f is not None:
                    Timp_f(T_imp[i], U, p, t_imp)
            elif Timp_f is not None:
                T_imp[i][...] = _timp_from_stage(U, temp, dtg)
    t_final = t + dt
    inc_imp = None if Timp_f is None else fused_raw_increment(dt, tab.b_imp, tend(T_imp, s), cd)
    if beta[s - 1] != 0:
        if f.has_T_lim:
            U_lim[...] = _axpy_dt(U_exp, dt, T_lim)
            f.lim(U_lim, p, t_final, U_exp)
            U_exp[...] = U_lim
        if f.has_T_exp:
            U_exp[...] = _axpy_dt(U_exp, dt, T_exp)
        b = cdt(beta[s - 1])
        if inc_imp is not None:
            u[...] = ((1 - b) * w(u) + b * w(U_exp)) + _wide(inc_imp)  # :120
        else:
            u[...] = (1 - b) * w(u) + b * w(U_exp)
    elif inc_imp is not None:
        u[...] = plus_inc(u, inc_imp)
    f.dss(u, p, t_final)
    if f.update_constrain_state.needs_update("EndOfStep"):
        f.constrain_state(u, p, t_final)
    if f.update_cache.needs_update("EndOfStep"):
        f.cache(u, p, t_final)
    return u


def lsrk_step(u, du, p, t, dt, f_incr, tab):
    """lsrk.jl:56-67: ``f(du,u,p,t+C*dt,1,A)`` then ``u .+= (dt*B) .* du``.
    The tableau is in ``eltype(u)`` (:51); ``dt*B`` is evaluated in ``promote(typeof(dt), FT)``."""
    A, B, C = tab
    f32_dt = isinstance(dt, np.float32)
    for k in range(len(A)):
        if f32_dt and isinstance(C[k], np.float32):
            stage_time = t + C[k] * dt
        else:
            stage_time = float(t) + float(C[k]) * float(dt)
        f_incr(du, u, p, stage_time, 1, A[k])
        if u.dtype == np.float32 and f32_dt:
            u[...] = u + (dt * B[k]) * du
        else:
            u[...] = _wide(u) + (float(dt) * float(B[k])) * _wide(du)
    return u


# --------------------------------------------------------------------------------------
# Rosenbrock  (src/solvers/rosenbrock.jl)
# --------------------------------------------------------------------------------------
class RosenbrockTableau:
    """rosenbrock.jl:25-46: A = α Γ⁻¹, C = diag(Γ⁻¹) − Γ⁻¹, m = b Γ⁻¹ (Float64).  For SSPKnoth Γ⁻¹ is exact in binary
    floating point, so the only rounding is in m = b Γ⁻¹, evaluated here as the row-times-matrix product in increasing
    column order (StaticArrays' `b / Γ` cannot be run in this container: last-bit unpinned for m)."""

    def __init__(self, alpha, Gamma, b):
        alpha, Gamma, b = (np.array(x, dtype=np.float64) for x in (alpha, Gamma, b))
        n = len(b)
        inv = np.linalg.inv(Gamma)
        # a unit-lower-triangular Γ with dyadic entries inverts exactly; make that explicit for the 3-stage tableau
        if np.allclose(np.diag(Gamma), 1.0) and np.allclose(np.triu(Gamma, 1), 0.0) and n <= 3:
            inv = np.eye(n)
            for i in range(n):
                for j in range(i):
                    inv[i, j] = -sum(Gamma[i, k] * inv[k, j] for k in range(j, i))
        self.alpha, self.Gamma = alpha, Gamma
        self.A = np.array([[sum(alpha[i, k] * inv[k, j] for k in range(n)) for j in range(n)] for i in range(n)])
        self.C = np.diag(np.diag(inv)) - inv
        self.m = np.array([sum(b[k] * inv[k, j] for k in range(n)) for j in range(n)])
        self.s = n


def ssp_knoth() -> RosenbrockTableau:
    """tableau(::SSPKnoth), rosenbrock.jl:252-267."""
    return RosenbrockTableau([[0, 0, 0], [1, 0, 0], [1 / 4, 1 / 4, 0]], [[1, 0, 0], [0, 1, 0], [-3 / 4, -3 / 4, 1]],
                             [1 / 6, 1 / 6, 2 / 3])


@dataclass
class RosenbrockAlgorithm:
    tableau: RosenbrockTableau


def init_cache_rosenbrock(u0, f: ClimaODEFunction, alg: RosenbrockAlgorithm):
    """rosenbrock.jl:100-124: U, fU, fU_imp, fU_exp, fU_lim, ∂Y∂t, k[1..N]; W is the jac_prototype ITSELF (not a copy)."""
    z = lambda: np.zeros_like(u0)
    return dict(U=z(), fU=z(), fU_imp=z(), fU_exp=z(), fU_lim=z(), dYdt=z(), k=[z() for _ in range(alg.tableau.s)],
                W=f.T_imp.jac_prototype if f.T_imp is not None else None)


def rosenbrock_step(u, p, t, dt, f: ClimaODEFunction, alg: RosenbrockAlgorithm, cache):
    """step_u!(int, cache::RosenbrockCache), rosenbrock.jl:136-239 — every `.=`/`.+=` below is its own pass, evaluated in
    Float64 (the tableau is Float64) and rounded to the state type on store."""
    tab = alg.tableau
    U, fU, fU_imp, fU_exp, fU_lim, dYdt, k, W = (cache[x] for x in ("U", "fU", "fU_imp", "fU_exp", "fU_lim", "dYdt", "k", "W"))
    T_imp, T_exp_T_lim = f.T_imp, f.T_exp_T_lim
    tgrad = None if T_imp is None else T_imp.tgrad
    st = u.dtype
    w = lambda x: x.astype(np.float64)
    fdt = float(dt)
    dtg = fdt * tab.Gamma[0, 0]
    if T_imp is not None:
        T_imp.Wfact(W, u, p, dtg, t)
    if tgrad is not None:
        tgrad(dYdt, u, p, t)
    for i in range(tab.s):
        fU[...] = 0
        ai = 0.0
        gi = 0.0
        for j in range(i):
            ai += tab.alpha[i, j]
        for j in range(i + 1):
            gi += tab.Gamma[i, j]
        U[...] = u
        for j in range(i):
            U[...] = (w(U) + tab.A[i, j] * w(k[j])).astype(st)
        ti = t + ai * dt
        if i != 0:
            f.dss(U, p, ti)
            if f.update_constrain_state.needs_update("EndOfStage"):
                f.constrain_state(U, p, ti)
            if f.update_cache.needs_update("EndOfStage"):
                f.cache(U, p, ti)
        if T_imp is not None:
            T_imp(fU_imp, U, p, ti)
            fU[...] = fU + fU_imp
        if T_exp_T_lim is not None:
            T_exp_T_lim(fU_exp, fU_lim, U, p, ti)
            fU[...] = fU + fU_exp
            if f.has_T_lim:
                fU[...] = fU + fU_lim
        if tgrad is not None:
            fU[...] = (w(fU) + (gi * fdt) * w(dYdt)).astype(st)
        for j in range(i):
            fU[...] = (w(fU) + (tab.C[i, j] / fdt) * w(k[j])).astype(st)
        fU[...] = (w(fU) * (-float(dtg))).astype(st)
        if T_imp is not None:
            W.ldiv(k[i], fU)
        else:
            k[i][...] = -fU
    for i in range(tab.s):
        u[...] = (w(u) + tab.m[i] * w(k[i])).astype(st)
    f.dss(u, p, t + dt)
    if f.update_constrain_state.needs_update("EndOfStep"):
        f.constrain_state(u, p, t + dt)
    if f.update_cache.needs_update("EndOfStep"):
        f.cache(u, p, t + dt)


# --------------------------------------------------------------------------------------
# integrator driver  (src/integrators.jl:187-257, 345-351, 410-439)
# --------------------------------------------------------------------------------------
def solve(u0, p, tspan, dt, step_fn, *, save_everystep=False):
    """``init`` + ``solve!``: dt clipping to the next tstop (:415-419), ``t`` snapped to the tstop
    within 100 eps (:424-429).  ``u0`` is advanced in place (``u = u0`` is not copied, :236).
    ``step_fn(u, p, t, dt)`` performs ``step_u!``.  Returns (t_final, u, saved)."""
    t0, tf = tspan
    tdir = -1 if t0 > tf else 1
    _dt = dt
    tstops = [tf]
    t = t0
    u = u0
    saved = [(t, u.copy())]
    nstep = 0
    while tstops:
        nstep += 1
        dt_step = tdir * min(_dt, abs(tstops[0] - t))
        step_fn(u, p, t, dt_step)
        t_plus = t + dt_step
        max_t_err = 100 * float(np.spacing(np.abs(t if isinstance(t, np.float32) else np.float64(t))))
        t = tstops[0] if abs(float(tstops[0]) - float(t_plus)) < max_t_err else t_plus
        if save_everystep:
            saved.append((t, u.copy()))
        while tstops and (t == tstops[0] or tdir * (tstops[0] - t) < 0):
            tstops.pop(0)
    if not save_everystep:
        saved.append((t, u.copy()))
    return t, u, saved


def make_stepper(f, alg, u0):
    """init_cache dispatch (imex_ark.jl:35 for Unconstrained, imex_ssprk.jl:33 for SSP, rosenbrock.jl:100)."""
    if isinstance(alg, RosenbrockAlgorithm):
        cache = init_cache_rosenbrock(u0, f, alg)
        return cache, (lambda u, p, t, dt: rosenbrock_step(u, p, t, dt, f, alg, cache))
    if alg.constraint == "SSP":
        cache = init_cache_ssprk(u0, f, alg)
        return cache, (lambda u, p, t, dt: imex_ssprk_step(u, p, t, dt, f, alg, cache))
    cache = init_cache_ark(u0, f, alg)
    return cache, (lambda u, p, t, dt: imex_ark_step(u, p, t, dt, f, alg, cache))
