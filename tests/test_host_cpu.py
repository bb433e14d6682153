"""CPU-only tests of the host side: the C-ABI library loads and exports every symbol declared in include/cts.h,
the ctypes mirrors match the header, the tableaux are the reference's, update-signal handlers / tstop queues behave
like the reference, and the product fails loudly without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

import cts_b200 as cts
from cts_b200 import _lib as L
from cts_b200 import integrators as I

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "cts.h")


def _declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(cts_[A-Za-z0-9_]+)\s*\(", src)) - {"cts_update_handler"})


def test_library_exports_every_declared_symbol():
    lib = L.load()
    names = _declared_symbols()
    assert len(names) >= 60
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/cts.h but not exported by libcts_sm100a.so"
    assert {n for n, _, _ in L.SIGNATURES} == set(names)
    assert lib.cts_abi_version() == 1 and b"sm100a" in lib.cts_version()


def test_ctypes_structs_match_header_layout():
    """Compile a tiny C program against include/cts.h and compare sizeof/offsetof with the ctypes mirrors."""
    prog = r"""
#include <stdio.h>
#include <stddef.h>
#include "cts.h"
int main(void){
 printf("%zu %zu %zu %zu %zu %zu %zu\n", sizeof(cts_ode_function), sizeof(cts_imex_tableau), sizeof(cts_newton_config),
        sizeof(cts_solver_stats), sizeof(cts_tridiag), sizeof(cts_column_desc), sizeof(cts_update_handler));
 printf("%zu %zu %zu %zu\n", offsetof(cts_ode_function, W), offsetof(cts_ode_function, update_constrain_state),
        offsetof(cts_newton_config, forcing_rtol), offsetof(cts_imex_tableau, cast_to_state_eltype));
 printf("%zu %zu %zu %zu %zu %zu\n", sizeof(cts_rosenbrock_tableau), sizeof(cts_cond_node), sizeof(cts_convergence_checker),
        offsetof(cts_rosenbrock_tableau, m), offsetof(cts_convergence_checker, comp_cond), offsetof(cts_convergence_checker, combiner_or));
 return 0; }"""
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        src = os.path.join(d, "t.c")
        open(src, "w").write(prog)
        exe = os.path.join(d, "t")
        subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), src, "-o", exe], check=True)
        out = subprocess.run([exe], capture_output=True, text=True, check=True).stdout.split()
    sizes = [int(x) for x in out]
    assert sizes[:7] == [C.sizeof(L.OdeFunction), C.sizeof(L.ImexTableau), C.sizeof(L.NewtonConfig), C.sizeof(L.SolverStats),
                         C.sizeof(L.Tridiag), C.sizeof(L.ColumnDesc), C.sizeof(L.UpdateHandler)]
    assert sizes[7:11] == [L.OdeFunction.W.offset, L.OdeFunction.update_constrain_state.offset,
                           L.NewtonConfig.forcing_rtol.offset, L.ImexTableau.cast_to_state_eltype.offset]
    assert sizes[11:] == [C.sizeof(L.RosenbrockTableau), C.sizeof(L.CondNode), C.sizeof(L.ConvergenceCheckerC),
                          L.RosenbrockTableau.m.offset, L.ConvergenceCheckerC.comp_cond.offset,
                          L.ConvergenceCheckerC.combiner_or.offset]


def test_no_gpu_means_loud_failure_not_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    h = C.c_void_p()
    assert L.load().cts_ctx_create(C.byref(h), 0, None) == -4  # CTS_ENODEVICE
    with pytest.raises(L.CTSError):
        cts.Context()
    with pytest.raises(TypeError):  # only B200Vector states are stepped: no numpy/CPU path in the product
        cts.init(cts.ODEProblem(cts.ClimaODEFunction(T_exp=lambda *a: None), np.zeros(3), (0.0, 1.0), None),
                 cts.ExplicitAlgorithm(cts.RK4()), dt=0.1)


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "climatimesteppers.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".jl")):
                txt = open(os.path.join(dirpath, f), errors="replace").read()
                # comments may cite the oracle files; importing, including, linking or loading them is forbidden
                assert not re.search(r"^\s*(import|from)\s+(cts_ref|cts_oracle|cts_models)\b", txt, flags=re.M), f
                assert not re.search(r"#include\s+[\"<].*oracle", txt), f
                assert "libcts_oracle" not in txt and "sys.path" not in txt, f


# ---------------------------------------------------------------- tableaux -----------------------
def test_all_named_tableaux_satisfy_reference_properties():
    # test/unit/tableaus.jl:48-130
    assert len(cts.IMEX_NAMES) == 32
    for name in cts.IMEX_NAMES + cts.EXPLICIT_NAMES:
        tab = getattr(cts, name)().tableau()
        s = tab.s
        assert np.all(np.triu(tab.a_exp) == 0) and np.all(np.triu(tab.a_imp, 1) == 0)
        np.testing.assert_allclose(tab.c_exp, tab.a_exp.sum(axis=1), atol=1e-13, err_msg=name)
        np.testing.assert_allclose(tab.c_imp, tab.a_imp.sum(axis=1), atol=1e-13, err_msg=name)
        assert abs(tab.b_exp.sum() - 1) < 1e-13 and abs(tab.b_imp.sum() - 1) < 1e-13, name
        assert s <= L.CTS_MAX_STAGES
    # SDIRK diagonals (tableaus.jl:96-111): one distinct non-zero diagonal value
    for name in ("ARS222", "ARS232", "ARS233", "ARS343", "ARS443", "SSP222", "SSP332", "ARK437L2SA1", "ARK548L2SA2", "DBM453"):
        d = np.diag(getattr(cts, name)().tableau().a_imp)
        assert len(set(d[d != 0])) == 1, name
    with pytest.raises(AssertionError):  # tableaus.jl:135-139
        cts.SSP333(beta=0.5).tableau()


def test_product_tableaux_equal_oracle_tableaux_bitwise():
    import cts_ref as R
    for name in ("ARS111", "ARS121", "ARS122", "ARS233", "ARS232", "ARS222", "ARS343", "ARS443", "SSP222", "SSP322", "SSP332",
                 "SSP333", "SSP433"):
        a, b = getattr(cts, name)().tableau(), R.tableau(name)
        for f in ("a_exp", "b_exp", "c_exp", "a_imp", "b_imp", "c_imp"):
            np.testing.assert_array_equal(getattr(a, f), getattr(b, f), err_msg=f"{name}.{f}")
    for dt in (np.float64, np.float32):
        for x, y in zip(cts.LSRK54CarpenterKennedy().tableau(dt), R.lsrk54_tableau(dt)):
            np.testing.assert_array_equal(np.array(x), np.array(y))
    assert cts.IMEXAlgorithm(cts.SSP333(), cts.NewtonsMethod()).constraint == cts.SSP       # imex_tableaus.jl:646
    assert cts.IMEXAlgorithm(cts.ARS343(), cts.NewtonsMethod()).constraint == cts.Unconstrained
    assert cts.ExplicitAlgorithm(cts.SSP33ShuOsher()).constraint == cts.SSP and cts.ExplicitAlgorithm(cts.RK4()).newtons_method is None
    assert not cts.is_fsal(cts.ARS343().tableau()) and cts.is_fsal(cts.ARS222().tableau())


# ---------------------------------------------------------------- update signal handlers ------------
def test_update_signal_handlers():
    # test/unit/update_signal_handler.jl
    every = cts.UpdateEvery(cts.WithDSS)
    assert every.needs_update(cts.EndOfStep) and every.needs_update(cts.EndOfStage) and every.needs_update(cts.WithDSS)
    assert not every.needs_update(cts.NewTimeStep)
    assert cts.UpdateEvery(cts.EndOfStep).needs_update(cts.EndOfStep) and not cts.UpdateEvery(cts.EndOfStep).needs_update(cts.EndOfStage)
    n3 = cts.UpdateEveryN(3, cts.NewNewtonIteration, cts.NewNewtonSolve)
    assert [n3.needs_update(cts.NewNewtonIteration) for _ in range(7)] == [True, False, False, True, False, False, True]
    n3.needs_update(cts.NewNewtonIteration)
    assert n3.needs_update(cts.NewNewtonSolve) is False          # reset signal
    assert n3.needs_update(cts.NewNewtonIteration) is True
    with pytest.raises(ValueError):
        cts.UpdateEveryN(2, cts.NewTimeStep, cts.NewTimeStep)
    d = cts.UpdateEveryDt(1.0)
    assert [d.needs_update(cts.NewTimeStep, t) for t in (0.0, 0.5, 1.0, 1.2, 2.5)] == [True, False, True, False, True]
    keep = []
    h = cts.UpdateEveryN(2, cts.NewTimeStep).to_c(keep)
    assert h.every_signal == L.SIG_NEW_TIMESTEP and bool(h.fn) and [h.fn(None, L.SIG_NEW_TIMESTEP, 0.0) for _ in range(3)] == [1, 0, 1]
    assert not bool(cts.UpdateEvery(cts.NewTimeStep).to_c(keep).fn)


def test_clima_ode_function_normalisation():
    # functions.jl:103-131
    f = cts.ClimaODEFunction(T_exp=lambda *a: None)
    assert cts.has_T_exp(f) and not cts.has_T_lim(f)
    f = cts.ClimaODEFunction(T_lim=lambda *a: None)
    assert cts.has_T_exp(f) and cts.has_T_lim(f)
    f = cts.ClimaODEFunction(T_imp=cts.ODEFunction(lambda *a: None))
    assert not cts.has_T_exp(f) and f.cache_imp is f.cache
    with pytest.raises(AssertionError):
        cts.ClimaODEFunction(T_exp_T_lim=lambda *a: None, T_exp=lambda *a: None)


# ---------------------------------------------------------------- integrator queues -----------------
def test_sorted_queue_and_tstops():
    # integrators.jl:12-46,131-144
    q = I.SortedQueue([3.0, 1.0, 2.0], True)
    assert q.first() == 1.0 and q.pop() == 1.0 and q.push(0.5).first() == 0.5 and len(q) == 3
    r = I.SortedQueue([1.0, 3.0, 2.0], False)
    assert r.first() == 3.0 and r.push(2.5).data == [3.0, 2.5, 2.0, 1.0]
    ts, sv = I._tstops_and_saveat_queues(0.0, 1.0, (0.5, 2.0, -1.0, 0.25), None)
    assert ts.data == [0.25, 0.5, 1.0] and sv.data == [0.0, 1.0]
    ts, sv = I._tstops_and_saveat_queues(1.0, 0.0, (0.5,), ())
    assert ts.data == [0.5, 0.0] and sv.isempty()


class _FakeCache:
    def __init__(self):
        self.calls = []

    def step_u(self, integrator):
        self.calls.append((integrator.t, integrator.dt))


def _fake_integrator(t0, tf, dt, tstops=(), saveat=None, **kw):
    """The integrator driver without a device: a recording step_u! stands in for the library."""
    prob = cts.ODEProblem(None, np.zeros(1), (t0, tf), None)
    ts, sv = I._tstops_and_saveat_queues(t0, tf, tstops, saveat)
    sol = cts.ODESolution([], [], prob, None)
    cb = I.CallbackSet(kw.pop("callback", None), I._saving_callback(lambda u, t: u.copy(), sol, kw.pop("save_everystep", False)))
    integ = I.TimeStepperIntegrator(alg=None, u=prob.u0, p=None, t=t0, dt=dt if tf > t0 else -dt, _dt=dt, dtchangeable=True,
                                    tstops=ts, _tstops=tstops, saveat=sv, _saveat=saveat, step=0, stepstop=-1, callback=cb,
                                    advance_to_tstop=False, cache=_FakeCache(), sol=sol, tdir=1 if tf > t0 else -1)
    for c in cb.discrete_callbacks:
        c.initialize(c, prob.u0, t0, integ)
    return integ


def test_step_clips_dt_to_tstops_and_snaps_t():
    # integrators.jl:410-439 and test/integration/integrator.jl
    integ = _fake_integrator(0.0, 1.0, 0.3, tstops=(0.5,))
    I.solve_(integ)
    ts = [c[0] for c in integ.cache.calls]
    dts = [c[1] for c in integ.cache.calls]
    assert integ.t == 1.0 and ts[0] == 0.0
    assert any(abs(d - 0.2) < 1e-15 for d in dts) and max(dts) <= 0.3 + 1e-15     # clipped to hit 0.5 exactly
    assert 0.5 in ts
    assert integ.sol.t == [0.0, 1.0]                                               # default saveat = (t0, tf)
    # reverse time
    integ = _fake_integrator(1.0, 0.0, 0.25)
    I.solve_(integ)
    assert integ.t == 0.0 and all(d < 0 for _, d in integ.cache.calls) and len(integ.cache.calls) == 4


def test_save_everystep_saveat_and_callbacks():
    integ = _fake_integrator(0.0, 1.0, 0.25, save_everystep=True)
    I.solve_(integ)
    assert integ.sol.t == [0.0, 0.25, 0.5, 0.75, 1.0]
    integ = _fake_integrator(0.0, 1.0, 0.25, saveat=(0.3, 0.75))
    I.solve_(integ)
    assert integ.sol.t == [0.5, 0.75]   # recorded at the first step on or after each save time
    hits = []
    integ = _fake_integrator(0.0, 1.0, 0.1, callback=I.EveryXSimulationSteps(lambda i: hits.append(i.step), 3))
    I.solve_(integ)
    assert hits == [3, 6, 9]
    with pytest.raises(ValueError):
        I.add_tstop_(integ, 0.5)        # behind the current time (integrators.jl:469-473)
    I.set_dt_(integ, 0.05)
    assert I.get_dt(integ) == 0.05
    with pytest.raises(ValueError):
        I.set_dt_(integ, -1.0)


def test_step_with_dt_stops_exactly():
    integ = _fake_integrator(0.0, 10.0, 0.3)
    I.step_(integ, 1.0, True)           # step!(integrator, dt, stop_at_tdt = true)
    assert integ.t == 1.0
    I.step_(integ)
    assert abs(integ.t - 1.3) < 1e-15
