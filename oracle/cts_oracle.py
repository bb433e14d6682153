"""ctypes wrapper of oracle/libcts_oracle.so (C restatement, reference pass structure; see cts_oracle.c).
TEST INFRASTRUCTURE ONLY."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_PATH = os.path.join(_HERE, "libcts_oracle.so")
_lib = None
_dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")


class ColModel(C.Structure):
    _fields_ = [("nlev", C.c_int), ("nx", C.c_size_t), ("ny_local", C.c_size_t), ("halo", C.c_int), ("dz", C.c_double),
                ("nu", C.c_double), ("nonlinear", C.c_int), ("K", C.c_void_p), ("S", C.c_void_p)]


class Tableau(C.Structure):
    _fields_ = [("s", C.c_int), ("a_exp", C.c_double * 64), ("b_exp", C.c_double * 8), ("c_exp", C.c_double * 8),
                ("a_imp", C.c_double * 64), ("b_imp", C.c_double * 8), ("c_imp", C.c_double * 8)]


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(_PATH):
            subprocess.run(["make", "-C", _HERE], check=True)
        lib = C.CDLL(_PATH)
        lib.oracle_num_threads.restype = C.c_int
        lib.oracle_set_num_threads.argtypes = [C.c_int]
        lib.oracle_fill_hash.argtypes = [_dp, C.c_size_t, C.c_uint64, C.c_uint64, C.c_double, C.c_double]
        lib.oracle_col_T_imp.argtypes = [C.POINTER(ColModel), _dp, _dp]
        lib.oracle_col_T_exp.argtypes = [C.POINTER(ColModel), _dp, _dp, C.c_double]
        lib.oracle_col_Wfact.argtypes = [C.POINTER(ColModel), _dp, _dp, _dp, _dp, C.c_double]
        lib.oracle_col_dss.argtypes = [C.POINTER(ColModel), _dp]
        lib.oracle_tridiag_solve.argtypes = [C.c_int, C.c_size_t, _dp, _dp, _dp, _dp, _dp]
        lib.oracle_ark_create.restype = C.c_void_p
        lib.oracle_ark_create.argtypes = [C.POINTER(ColModel), C.POINTER(Tableau), C.c_int, C.c_int]
        lib.oracle_ark_destroy.argtypes = [C.c_void_p]
        lib.oracle_ark_step.argtypes = [C.c_void_p, C.POINTER(ColModel), _dp, C.c_double, C.c_double]
        lib.oracle_ark_buffer.restype = C.POINTER(C.c_double)
        lib.oracle_ark_buffer.argtypes = [C.c_void_p, C.c_char_p, C.c_int]
        lib.oracle_lsrk_advection_step.argtypes = [_dp, _dp, C.c_size_t, C.c_double, C.c_double, C.c_double, C.c_int,
                                                   _dp, _dp]
        _lib = lib
    return _lib


def fill_hash(n, seed, offset=0, scale=1.0, shift=0.0):
    out = np.empty(n)
    load().oracle_fill_hash(out, n, seed, offset, scale, shift)
    return out


def tableau_c(tab) -> Tableau:
    """`tab`: an object with numpy a_exp/b_exp/c_exp/a_imp/b_imp/c_imp (oracle or product IMEXTableau: data only)."""
    t = Tableau()
    s = len(tab.b_exp)
    t.s = s
    for i in range(s):
        for j in range(s):
            t.a_exp[i * s + j] = tab.a_exp[i, j]
            t.a_imp[i * s + j] = tab.a_imp[i, j]
        t.b_exp[i], t.b_imp[i], t.c_exp[i], t.c_imp[i] = tab.b_exp[i], tab.b_imp[i], tab.c_exp[i], tab.c_imp[i]
    return t


UPDATE_J = {"NewTimeStep": 0, "NewNewtonSolve": 1, "NewNewtonIteration": 2}


class ColumnARK:
    """ARK integrator on the column model in the reference's pass structure (one loop per broadcast)."""

    def __init__(self, nlev, nx, ny_local, K, S=None, nu=0.0, halo=0, nonlinear=False, dz=None, tab=None, max_iters=1,
                 update_j="NewNewtonIteration"):
        self.lib = load()
        self.K = np.ascontiguousarray(K, dtype=np.float64)
        self.S = None if S is None else np.ascontiguousarray(S, dtype=np.float64)
        self.m = ColModel(nlev, nx, ny_local, halo, 1.0 / (nlev + 1) if dz is None else dz, nu, int(nonlinear),
                          self.K.ctypes.data, None if self.S is None else self.S.ctypes.data)
        self.n = (ny_local + 2 * halo) * nx * nlev
        self.tab = tableau_c(tab)
        self.h = self.lib.oracle_ark_create(C.byref(self.m), C.byref(self.tab), max_iters, UPDATE_J[update_j])

    def dss(self, u):
        self.lib.oracle_col_dss(C.byref(self.m), u)

    def step(self, u, t, dt):
        self.lib.oracle_ark_step(self.h, C.byref(self.m), u, t, dt)

    def buffer(self, name, idx=0):
        p = self.lib.oracle_ark_buffer(self.h, name.encode(), idx)
        return np.ctypeslib.as_array(p, shape=(self.n,)) if p else None

    def __del__(self):
        try:
            self.lib.oracle_ark_destroy(self.h)
        except Exception:
            pass


def lsrk_advection_step(u, du, c, dx, dt, A, B):
    A = np.ascontiguousarray(A, dtype=np.float64)
    B = np.ascontiguousarray(B, dtype=np.float64)
    load().oracle_lsrk_advection_step(u, du, u.size, c, dx, dt, len(A), A, B)
