"""The C restatement (oracle/cts_oracle.c, used at full size and as the CPU baseline) must agree bit for bit
with the numpy oracle, which is the one pinned to the reference's golden values.  CPU only."""
import numpy as np
import pytest

import cts_models as M
import cts_oracle as CO
import cts_ref as R


def test_hash_matches():
    np.testing.assert_array_equal(CO.fill_hash(5000, 3, 77, 0.5, 1.0), M.fill_hash(5000, 3, 77, 0.5, 1.0))


@pytest.mark.parametrize("nlev,ncols", [(64, 40), (10, 7), (1, 5), (3, 4), (128, 3)])
def test_babe_matches(nlev, ncols):
    rng = np.random.default_rng(0)
    n = nlev * ncols
    dl, du = rng.uniform(.1, 1, n), rng.uniform(.1, 1, n)
    d = -(dl + du) - 1.5
    rhs = rng.standard_normal(n)
    x = np.empty(n)
    CO.load().oracle_tridiag_solve(nlev, ncols, dl, d, du, rhs, x)
    np.testing.assert_array_equal(x, M.babe_solve(dl, d, du, rhs, nlev))


@pytest.mark.parametrize("name", ["ARS343", "ARS222", "ARS443", "SSP333"])
@pytest.mark.parametrize("update_j", ["NewTimeStep", "NewNewtonSolve", "NewNewtonIteration"])
@pytest.mark.parametrize("nonlinear,max_iters", [(False, 1), (True, 2)])
def test_column_ark_step_matches_numpy_oracle(name, update_j, nonlinear, max_iters):
    if name == "SSP333" and update_j == "NewTimeStep":
        pytest.skip("SSP333 is not bitwise SDIRK (SURVEY Appendix B)")
    nlev, nx, ny, halo = 16, 6, 5, 1
    K, T0, S = M.column_inputs(nlev, nx, ny, 0, ny, halo)
    model = M.ColumnModel(nlev, nx, ny, K, S_lev=S, nu=0.8, halo=halo, nonlinear=nonlinear)
    tab = R.tableau(name)
    alg = R.IMEXAlgorithm(tab, R.NewtonsMethod(max_iters=max_iters, update_j=R.UpdateEvery(update_j)), "Unconstrained", name)
    u1 = T0.copy()
    model.dss(u1, None, 0.0)
    _, step = R.make_stepper(model.ode_function(), alg, u1)
    co = CO.ColumnARK(nlev, nx, ny, K, S, 0.8, halo, nonlinear, tab=tab, max_iters=max_iters, update_j=update_j)
    u2 = T0.copy()
    co.dss(u2)
    t = 0.0
    for _ in range(3):
        step(u1, None, t, 0.02)
        co.step(u2, t, 0.02)
        t += 0.02
    np.testing.assert_array_equal(u2, u1)


def test_thread_count_does_not_change_results():
    nlev, nx, ny = 64, 16, 8
    K, T0, S = M.column_inputs(nlev, nx, ny, 0, ny, 1)
    outs = []
    lib = CO.load()
    nthr = lib.oracle_num_threads()
    for th in (1, max(2, nthr)):
        lib.oracle_set_num_threads(th)
        co = CO.ColumnARK(nlev, nx, ny, K, S, 1.0, 1, tab=R.tableau("ARS343"))
        u = T0.copy()
        co.dss(u)
        co.step(u, 0.0, 0.02)
        outs.append(u)
    lib.oracle_set_num_threads(nthr)
    np.testing.assert_array_equal(outs[0], outs[1])


def test_lsrk_advection_matches():
    n = 4099
    u0 = M.fill_hash(n, 1)
    adv = M.Advection(n)
    A, B, C = R.lsrk54_tableau()
    u1, du1 = u0.copy(), np.zeros(n)
    u2, du2 = u0.copy(), np.zeros(n)
    for k in range(3):
        R.lsrk_step(u1, du1, None, 0.0, 0.5 / n, adv.incr, (A, B, C))
        CO.lsrk_advection_step(u2, du2, 1.0, 1.0 / n, 0.5 / n, A, B)
    np.testing.assert_array_equal(u2, u1)
    np.testing.assert_array_equal(du2, du1)


@pytest.mark.parametrize("dtype,scale", [(np.float64, 0.1), (np.float64, 22.0), (np.float64, 1e4), (np.float32, 22.0)])
def test_product_form_accuracy(dtype, scale):
    """The product-form elimination (reciprocal off the dependent chain) is as accurate as textbook Thomas on
    diagonally dominant W = dtγJ - I, and never leaves the exponent range thanks to the power-of-two rescaling."""
    rng = np.random.default_rng(0)
    nc, n = 500, 64
    dl = (scale * rng.uniform(.5, 1, nc * n)).astype(dtype)
    du = (scale * rng.uniform(.5, 1, nc * n)).astype(dtype)
    d = (-(dl + du) - 1).astype(dtype)
    r = rng.standard_normal(nc * n).astype(dtype)
    x = M.babe_solve(dl, d, du, r, n)
    xe = M.thomas_reference(dl, d, du, r, n)
    err = np.abs(x.astype(np.longdouble) - xe).reshape(nc, n).max(axis=1) / np.abs(xe).reshape(nc, n).max(axis=1)
    eps = np.finfo(dtype).eps
    assert np.isfinite(x).all()
    assert float(err.max()) < 40 * eps * max(1.0, scale ** 0.5)


def test_product_form_large_pivots_do_not_overflow():
    # pivots ~ 1e30 per row over 128 rows would overflow any un-scaled running product
    n, nc = 128, 4
    dl = np.full(nc * n, 1e29)
    du = np.full(nc * n, 1e29)
    d = np.full(nc * n, -1e30)
    r = np.linspace(1, 2, nc * n)
    x = M.babe_solve(dl, d, du, r, n)
    xe = M.thomas_reference(dl, d, du, r, n)
    assert np.isfinite(x).all()
    assert float(np.max(np.abs(x - xe)) / np.max(np.abs(xe))) < 1e-14
