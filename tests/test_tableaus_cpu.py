"""Every named tableau of the reference (32 IMEX + 3 explicit; SSPKnoth lives in test_rosenbrock_cpu.py) against the oracle's
OWN transcription (oracle/cts_ref.py `tableau`, oracle/kc2019_tables.json) and against the orders the reference expects:
`test_convergence!` (test/utils/convergence_utils.jl:174-226) restated -- same sampling of step counts, average RMS error
over all saved steps, log-log regression with a 99 % Student-t confidence interval and the same pass criterion --
with `imex_convergence_orders` (:20-49).  A transcription error in any coefficient changes the measured order."""
import math

import numpy as np
import pytest
from scipy import stats

import ref_problems as P
from ref_problems import R

ALL = R.IMEX_NAMES + R.EXPLICIT_NAMES


# ---------------------------------------------------------------- product table == oracle table ------------------
@pytest.mark.parametrize("name", ALL)
def test_product_tableau_equals_oracle_transcription(name):
    import cts_b200
    from cts_b200 import tableaus as T
    o, p = R.tableau(name), T.tableau(getattr(cts_b200, name)())
    for k in ("a_exp", "a_imp", "b_exp", "b_imp", "c_exp", "c_imp"):
        a, b = np.asarray(getattr(o, k), dtype=np.float64), np.asarray(getattr(p, k), dtype=np.float64)
        assert a.shape == b.shape and np.array_equal(a, b), (name, k)


@pytest.mark.parametrize("name", ALL)
def test_tableau_identities(name):
    # test/unit/tableaus.jl:48-91: c = rowsum(a) and sum(b) = 1 to 1e-13; strictly lower / lower triangular
    tab = R.tableau(name)
    np.testing.assert_allclose(tab.c_exp, tab.a_exp.sum(axis=1), atol=1e-13)
    np.testing.assert_allclose(tab.c_imp, tab.a_imp.sum(axis=1), atol=1e-13)
    assert abs(tab.b_exp.sum() - 1) < 1e-13 and abs(tab.b_imp.sum() - 1) < 1e-13


def test_ark2gkc_paper_version_differs():
    a, b = R.tableau("ARK2GKC"), R.tableau("ARK2GKC", paper_version=True)
    assert a.a_exp[2, 1] == 0.5 and b.a_exp[2, 1] == 1 / 2 + math.sqrt(2.0) / 3


# ---------------------------------------------------------------- test_convergence! ------------------------------
def convergence_order(dts, errs, confidence=0.99):
    """convergence_utils.jl:53-82."""
    x, y = np.log10(dts), np.log10(errs)
    n = len(dts)
    order, icpt = np.linalg.lstsq(np.stack([x, np.ones(n)], 1), y, rcond=None)[0]
    ndof = n - 2
    tval = -stats.t.ppf((1 - confidence) / 2, ndof)
    sd = math.sqrt(np.sum((y - (x * order + icpt)) ** 2) / ndof)
    return order, tval * sd / math.sqrt(np.sum((x - x.mean()) ** 2))


def check_convergence(name, f, Y0, t_end, analytic, linear_implicit, default_num_steps, predicted, shifts=1, explicit=False):
    """test_convergence! (:174-226) with the analytic solution as the reference."""
    powers = np.arange(-1, 1.01, 0.5) - shifts * max(0, predicted - 3) / 2
    nsteps = np.round(default_num_steps * 4.0 ** powers).astype(int)
    dts = t_end / nsteps
    errs = []
    for dt in dts:
        nm = None if explicit else R.NewtonsMethod(max_iters=1 if linear_implicit else 2)
        alg = R.imex_algorithm(name, nm)
        u = np.array(Y0, dtype=np.float64)
        _, step = R.make_stepper(f, alg, u)
        _, _, saved = R.solve(u, None, (0.0, t_end), dt, step, save_everystep=True)
        rms = lambda a: float(np.linalg.norm(a) / math.sqrt(np.size(a)))
        errs.append(rms(np.array([rms(np.abs(uu - analytic(tt))) for tt, uu in saved])))
    order, unc = convergence_order(dts, np.array(errs))
    assert not math.isnan(order)
    assert abs(order - predicted) <= unc or unc > predicted / 10, (name, order, unc, predicted, errs)
    return order


@pytest.mark.parametrize("name", R.IMEX_NAMES)
def test_imex_convergence_ark_analytic_nonlin(name):
    # test/solvers/imex_ark.jl:50: implicit-only nonlinear problem, 450 steps, two Newton iterations
    case = P.ark_analytic_nonlin()
    check_convergence(name, case.split_f, case.Y0, case.t_end, case.analytic_sol, False, 450, R.CONVERGENCE_ORDERS[name][0])


@pytest.mark.parametrize("name", R.IMEX_NAMES)
def test_imex_convergence_ark_analytic_sys(name):
    # :51 (BASELINE config 1 problem): implicit-only linear system, 200 steps
    case = P.ark_analytic_sys()
    check_convergence(name, case.split_f, case.Y0, case.t_end, case.analytic_sol, True, 200, R.CONVERGENCE_ORDERS[name][0])


@pytest.mark.parametrize("name", ["ARS343", "SSP333", "ARS232", "IMKG343a", "IMKG242a", "HOMMEM1", "DBM453", "ARK2GKC"])
def test_imex_convergence_split_stiff_linear(name):
    # :77-81 stiff_linear (both tendencies -> the combined order); the reference uses ARK548L2SA2 with 64x more steps as the
    # numerical reference, the oracle can use the analytic matrix exponential.  A subset of names keeps the CPU suite short.
    case = P.stiff_linear()
    check_convergence(name, case.split_f, case.Y0, case.t_end, case.analytic_sol, True, 1000 // 4, R.CONVERGENCE_ORDERS[name][2])


@pytest.mark.parametrize("name", R.EXPLICIT_NAMES + ("SSP333", "SSP222"))
def test_explicit_convergence(name):
    # test/solvers/explicit_rk.jl:11-49,61-62: du = p u (100 steps) and the sin/cos rotation (800 steps), ExplicitAlgorithm
    order = R.CONVERGENCE_ORDERS[name][1]
    p = 1.01
    f = R.ClimaODEFunction(T_exp=lambda du, u, _p, t: du.__setitem__(Ellipsis, p * u))
    check_convergence(name, f, [0.5], 1.0, lambda t: np.array([0.5 * math.exp(p * t)]), False, 100, order, explicit=True)
    w = 2.0

    def rot(du, u, _p, t):
        du[0], du[1] = w * u[1], -w * u[0]

    sol = lambda t: np.array([math.sin(w * t), math.cos(w * t)])
    check_convergence(name, R.ClimaODEFunction(T_exp=rot), [0.0, 1.0], 1.0, sol, False, 800, order, explicit=True)
