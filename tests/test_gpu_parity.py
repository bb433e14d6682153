"""GPU parity tests: the CUDA path, driven through the C ABI (ctypes mirror), against the CPU oracle on the
same seeded inputs.  Float64/Float32 element-wise kernels, the tridiagonal solve and whole steps with
library-native operators are required to be BIT-EXACT (the kernels use un-contracted IEEE ops in the
reference's evaluation order); paths downstream of a reduction (GMRES/JFNK) are held to the north-star
tolerance: max relative error <= 1e-12 per step in Float64, <= 1e-5 in Float32."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import cts_models as M  # noqa: E402
import ref_problems as P  # noqa: E402
from ref_problems import R  # noqa: E402


@pytest.fixture(scope="module")
def cts():
    import cts_b200
    return cts_b200


@pytest.fixture(scope="module")
def ctx(cts):
    return cts.Context.default()


def dev(cts, a):
    return cts.B200Vector.from_host(a)


def rel_err(a, b):
    return float(np.max(np.abs(a.astype(np.float64) - b.astype(np.float64))) / max(1e-300, np.max(np.abs(b.astype(np.float64)))))


# ------------------------------------------------------------------ inputs / element-wise --------
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_fill_hash_bit_exact(cts, dtype):
    n = 10007
    v = cts.B200Vector.zeros(n, dtype).fill_hash(3, global_offset=12345, scale=0.5, shift=1.0)
    np.testing.assert_array_equal(v.to_host(), M.fill_hash(n, 3, 12345, 0.5, 1.0, dtype))


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("n", [1, 7, 4096, 100003])
@pytest.mark.parametrize("n1,n2", [(0, 0), (1, 0), (0, 1), (2, 3), (4, 3), (8, 6)])
def test_axpy_n_matches_fused_increment(cts, dtype, n, n1, n2):
    # K1 semantics of test/unit/fused_increment.jl:83-98 (exact ==) against oracle fused_raw_increment / assign_with_increments
    rng = np.random.default_rng(n1 * 17 + n2)
    base = rng.standard_normal(n).astype(dtype)
    T = [rng.standard_normal(n).astype(dtype) for _ in range(n1 + n2)]
    a = rng.standard_normal(n1 + n2)
    dt = 0.0371
    inc1 = R.fused_raw_increment(dt, a[:n1], T[:n1])
    inc2 = R.fused_raw_increment(dt, a[n1:], T[n1:])
    want = np.zeros(n, dtype)
    R.assign_with_increments(want, base, inc1, inc2)
    out = cts.B200Vector.zeros(n, dtype)
    out2 = cts.B200Vector.zeros(n, dtype)
    dT = [dev(cts, t) for t in T]
    d_base = dev(cts, base)
    cts.axpy_n(out, d_base, [(dt * a[j], dT[j]) for j in range(n1)],
               [(dt * a[n1 + j], dT[n1 + j]) for j in range(n2)], out2=out2)
    np.testing.assert_array_equal(out.to_host(), want)
    np.testing.assert_array_equal(out2.to_host(), want)


def test_axpy_n_in_place_unaligned_and_scaled_base(cts):
    # aliasing out == base (assign_with_increments!(u, u, ...), imex_ark.jl:90) on a view that is not 16-byte aligned
    import torch
    n = 1001
    rng = np.random.default_rng(5)
    big = torch.from_numpy(rng.standard_normal(n + 1)).cuda()
    u = cts.B200Vector(big[1:])  # 8-byte offset
    T1 = dev(cts, rng.standard_normal(n))
    u_h, T_h = u.to_host().copy(), T1.to_host()
    beta = 0.25
    cts.axpy_n(u, u, [(beta, T1)], c0=1 - beta)  # (1-β)u + β T   (imex_ssprk.jl:165)
    np.testing.assert_array_equal(u.to_host(), (1 - beta) * u_h + beta * T_h)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_small_elementwise_kernels(cts, ctx, dtype):
    n = 5003
    rng = np.random.default_rng(11)
    FT = np.dtype(dtype).type
    r, U0, Up, dx = (rng.standard_normal(n).astype(dtype) for _ in range(4))
    lib, h = ctx.lib, ctx.handle
    ft = 1 if dtype == np.float64 else 0
    dtg = 0.0123
    w = lambda x: x.astype(np.float64)
    d_r = dev(cts, r)
    g_U0, g_Up, g_dx = dev(cts, U0), dev(cts, Up), dev(cts, dx)  # keep the device buffers alive
    ctx.check(lib.cts_newton_residual(h, ft, n, d_r.ptr, g_U0.ptr, dtg, g_Up.ptr))  # K3
    np.testing.assert_array_equal(d_r.to_host(), ((w(U0) + dtg * w(r)) - w(Up)).astype(dtype))
    d_x = dev(cts, U0)
    ctx.check(lib.cts_sub_inplace(h, ft, n, d_x.ptr, g_dx.ptr))  # K5
    np.testing.assert_array_equal(d_x.to_host(), U0 - dx)
    d_T = cts.B200Vector.zeros(n, dtype)
    ctx.check(lib.cts_timp_from_stage(h, ft, n, d_T.ptr, g_Up.ptr, g_U0.ptr, dtg))  # K6
    np.testing.assert_array_equal(d_T.to_host(), (w(Up - U0) / dtg).astype(dtype))
    eps = float(FT(1.7e-4))
    d_x2 = cts.B200Vector.zeros(n, dtype)
    ctx.check(lib.cts_jvp_perturb(h, ft, n, d_x2.ptr, g_U0.ptr, eps, g_dx.ptr))  # K7
    np.testing.assert_array_equal(d_x2.to_host(), U0 + FT(eps) * dx)
    d_q = cts.B200Vector.zeros(n, dtype)
    ctx.check(lib.cts_jvp_quotient(h, ft, n, d_q.ptr, g_Up.ptr, g_U0.ptr, eps))  # K8
    np.testing.assert_array_equal(d_q.to_host(), (Up - U0) / FT(eps))
    d_u = dev(cts, U0)
    ctx.check(lib.cts_lsrk_update(h, ft, n, d_u.ptr, 0.37, g_dx.ptr, 0))  # K10 (Float64 dt*B)
    np.testing.assert_array_equal(d_u.to_host(), (w(U0) + 0.37 * w(dx)).astype(dtype))


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("n", [1, 1000, 2 ** 20 + 3])
def test_reductions(cts, dtype, n):
    rng = np.random.default_rng(2)
    x, y = rng.standard_normal(n).astype(dtype), rng.standard_normal(n).astype(dtype)
    dx_, dy_ = dev(cts, x), dev(cts, y)
    # the oracle restates the library's fixed reduction tree: bit-identical, and close to numpy's own sum
    assert dx_.norm() == R.norm2(x)
    assert dx_.dot(dy_) == R.dot(x, y)
    assert dx_.sum1pabs() == R.sum1pabs(x)
    assert abs(R.norm2(x) - np.linalg.norm(x.astype(np.float64))) <= 1e-13 * R.norm2(x)
    assert dx_.norm() == dx_.norm()  # deterministic


# ------------------------------------------------------------------ tridiagonal solve (K4) -------
def _random_W(ncols, nlev, dtype, seed=0):
    rng = np.random.default_rng(seed)
    n = ncols * nlev
    dl = rng.uniform(0.1, 1.0, n).astype(dtype)
    du = rng.uniform(0.1, 1.0, n).astype(dtype)
    d = (-(dl + du) - 1 - rng.uniform(0, 1, n)).astype(dtype)
    rhs = rng.standard_normal(n).astype(dtype)
    return dl, d, du, rhs


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("ncols,nlev", [(32, 64), (37, 64), (1, 64), (1000, 64), (64, 16), (5, 10), (33, 1), (7, 3), (40, 128)])
def test_tridiag_solve_bit_exact(cts, ctx, dtype, ncols, nlev):
    from cts_b200 import _lib as L
    dl, d, du, rhs = _random_W(ncols, nlev, dtype)
    want = M.babe_solve(dl, d, du, rhs, nlev)
    vs = [dev(cts, a) for a in (dl, d, du, rhs)]
    x = cts.B200Vector.zeros(ncols * nlev, dtype)
    W = L.Tridiag(nlev, ncols, vs[0].ptr, vs[1].ptr, vs[2].ptr)
    ctx.check(ctx.lib.cts_tridiag_solve_batched(ctx.handle, x.ft, C.byref(W), vs[3].ptr, x.ptr), "tridiag")
    got = x.to_host()
    np.testing.assert_array_equal(got, want)
    # and it really solves the system (LAPACK-style check of the oracle itself)
    for c in range(min(ncols, 3)):
        s = slice(c * nlev, (c + 1) * nlev)
        A = np.diag(d[s].astype(np.float64)) + np.diag(dl[s][1:].astype(np.float64), -1) + np.diag(du[s][:-1].astype(np.float64), 1)
        ref = np.linalg.solve(A, rhs[s].astype(np.float64))
        assert rel_err(got[s], ref) < (1e-13 if dtype == np.float64 else 2e-5)
    # mul! is the inverse operation
    y = cts.B200Vector.zeros(ncols * nlev, dtype)
    ctx.check(ctx.lib.cts_tridiag_matvec(ctx.handle, x.ft, C.byref(W), x.ptr, y.ptr), "matvec")
    assert rel_err(y.to_host(), rhs) < (1e-12 if dtype == np.float64 else 1e-4)


def test_tridiag_solve_without_tma_is_identical():
    # the cooperative-load variant of the tile kernel (CTS_TILE_LOADER selects cooperative / bulk-TMA / cp.async loaders) must give the same bits
    code = r"""
import sys, numpy as np, ctypes as C
sys.path[:0] = [%r, %r, %r]
import cts_b200 as cts, cts_models as M
from cts_b200 import _lib as L
ctx = cts.Context.default()
rng = np.random.default_rng(0); nc, nl = 70, 64; n = nc*nl
dl = rng.uniform(.1,1,n); du = rng.uniform(.1,1,n); d = -(dl+du)-1.5; rhs = rng.standard_normal(n)
vs = [cts.B200Vector.from_host(a) for a in (dl,d,du,rhs)]
x = cts.B200Vector.zeros(n)
W = L.Tridiag(nl, nc, vs[0].ptr, vs[1].ptr, vs[2].ptr)
ctx.check(ctx.lib.cts_tridiag_solve_batched(ctx.handle, 1, C.byref(W), vs[3].ptr, x.ptr))
assert np.array_equal(x.to_host(), M.babe_solve(dl,d,du,rhs,nl))
print("OK")
""" % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))), os.path.dirname(os.path.abspath(__file__)),
       os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
    env = dict(os.environ, CTS_TILE_LOADER="1")
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "OK" in out.stdout, out.stderr[-2000:]


# ------------------------------------------------------------------ column model hooks -----------
def _column_pair(cts, nlev=64, nx=8, ny=6, halo=1, nu=0.7, nonlinear=False, dtype=np.float64, source=True):
    K, T0, S = M.column_inputs(nlev, nx, ny, 0, ny, halo, dtype)
    model = M.ColumnModel(nlev, nx, ny, K, S_lev=S if source else None, nu=nu, halo=halo, nonlinear=nonlinear)
    op = cts.ColumnOperator(nlev, nx, ny, dev(cts, K), S_lev=dev(cts, S) if source else None, nu=nu, halo=halo,
                            nonlinear=nonlinear)
    return model, op, T0


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("nonlinear", [False, True])
@pytest.mark.parametrize("halo", [0, 1])
def test_column_hooks_bit_exact(cts, dtype, nonlinear, halo):
    model, op, T0 = _column_pair(cts, halo=halo, nonlinear=nonlinear, dtype=dtype)
    n = model.n
    assert op.ndof == n
    u = dev(cts, T0)
    want, got = np.zeros(n, dtype), cts.B200Vector.zeros(n, dtype)
    model.T_imp(want, T0, None, 0.3)
    op.T_imp(got, u, None, 0.3)
    np.testing.assert_array_equal(got.to_host(), want)
    want[:] = 0
    got = cts.B200Vector.zeros(n, dtype)
    model.T_exp(want, T0, None, 0.3)
    op.T_exp(got, u, None, 0.3)
    np.testing.assert_array_equal(got.to_host(), want)
    Wm, Wd = model.jac_prototype().zero(), op.jac_prototype().zero()
    model.Wfact(Wm, T0, None, 0.0087, 0.3)
    op.Wfact(Wd, u, None, 0.0087, 0.3)
    for a, b in zip(Wd.diagonals(), (Wm.dl, Wm.d, Wm.du)):
        np.testing.assert_array_equal(a.to_host(), b)
    rhs = M.fill_hash(n, 9, dtype=dtype)
    xm = np.zeros(n, dtype)
    Wm.ldiv(xm, rhs)
    xd = cts.B200Vector.zeros(n, dtype)
    d_rhs = dev(cts, rhs)
    Wd.ldiv(xd, d_rhs)
    np.testing.assert_array_equal(xd.to_host(), xm)
    ud, uh = dev(cts, T0), T0.copy()
    op.dss(ud, None, 0.0)
    model.dss(uh, None, 0.0)
    np.testing.assert_array_equal(ud.to_host(), uh)


# ------------------------------------------------------------------ whole IMEX steps -------------
def _run_imex_pair(cts, name, newton_kw, nsteps, dt, *, dtype=np.float64, fusion=True, matrix_free=False, halo=1,
                   nonlinear=False, constraint=None, krylov=None, nlev=64, nx=8, ny=6):
    model, op, T0 = _column_pair(cts, nlev=nlev, nx=nx, ny=ny, halo=halo, nonlinear=nonlinear, dtype=dtype)
    if matrix_free:
        op.set_matrix_free(True)
    # oracle
    o_kw = dict(newton_kw)
    c_kw = dict(newton_kw)
    if "update_j" in newton_kw:
        o_kw["update_j"] = R.UpdateEvery(newton_kw["update_j"])
        c_kw["update_j"] = cts.UpdateEvery(newton_kw["update_j"])
    if krylov is not None:
        o_kw["krylov_method"] = R.KrylovMethod(jacobian_free_jvp=R.ForwardDiffJVP() if krylov["jfnk"] else None,
                                               forcing_term=R.ConstantForcing(krylov["rtol"]))
        c_kw["krylov_method"] = cts.KrylovMethod(jacobian_free_jvp=cts.ForwardDiffJVP() if krylov["jfnk"] else None,
                                                 forcing_term=cts.ConstantForcing(krylov["rtol"]))
    try:
        alg_o = R.imex_algorithm(name, R.NewtonsMethod(**o_kw), constraint)
    except KeyError:  # tableaux the oracle does not restate: take the coefficient data from the product (data only)
        tb = getattr(cts, name)().tableau()
        alg_o = R.IMEXAlgorithm(R.IMEXTableau(tb.a_exp, tb.a_imp, tb.b_exp, tb.b_imp, tb.c_exp, tb.c_imp),
                                R.NewtonsMethod(**o_kw), constraint or "Unconstrained", name)
    uo = T0.copy()
    fo = model.ode_function()
    model.dss(uo, None, 0.0)
    cache_o, step_o = R.make_stepper(fo, alg_o, uo)
    # CUDA
    alg_c = cts.IMEXAlgorithm(getattr(cts, name)(), cts.NewtonsMethod(**c_kw), constraint)
    uc = dev(cts, T0)
    op.dss(uc, None, 0.0)
    prob = cts.ODEProblem(op.ode_function(), uc, (0.0, 1000 * dt), None)  # far tstop: no dt clipping (integrators.jl:415-419)
    integ = cts.init(prob, alg_c, dt=dt)
    integ.cache.set_fusion(fusion)
    t = 0.0
    for _ in range(nsteps):
        step_o(uo, None, t, dt)
        cts.step_(integ)
        t += dt
    return uo, integ.u.to_host(), integ, cache_o, model


@pytest.mark.parametrize("name", ["ARS343", "ARS222", "ARS233", "ARS443", "IMKG343a", "DBM453", "SSP333"])
@pytest.mark.parametrize("fusion,matrix_free", [(False, False), (True, False), (True, True)])
def test_imex_ark_column_steps_bit_exact(cts, name, fusion, matrix_free):
    # config C3 in miniature: ARS343 & co, NewtonsMethod(max_iters = 1), per-iteration Wfact, tridiagonal ldiv!
    constraint = "Unconstrained" if name == "SSP333" else None
    uo, uc, integ, _, model = _run_imex_pair(cts, name, dict(max_iters=1), 3, 0.02, fusion=fusion,
                                             matrix_free=matrix_free, constraint=constraint)
    own = slice(model.nx * model.nlev, -model.nx * model.nlev)
    np.testing.assert_array_equal(uc[own], uo[own])
    st = integ.cache.stats()
    assert (st["fused_stage_launches"] > 0) == fusion


@pytest.mark.parametrize("update_j", ["NewTimeStep", "NewNewtonSolve", "NewNewtonIteration"])
def test_imex_ark_jacobian_update_policies(cts, update_j):
    # tutorial configuration (docs/src/tutorials/imex_diffusion.md:161-167): update_j = UpdateEvery(NewTimeStep)
    uo, uc, integ, _, model = _run_imex_pair(cts, "ARS343", dict(max_iters=1, update_j=update_j), 2, 0.02)
    np.testing.assert_array_equal(uc, uo)
    uo2, uc2, *_ = _run_imex_pair(cts, "ARS343", dict(max_iters=1, update_j=update_j), 2, 0.02, fusion=False)
    np.testing.assert_array_equal(uc2, uo2)


@pytest.mark.parametrize("dtype,tol", [(np.float64, 0.0), (np.float32, 0.0)])
@pytest.mark.parametrize("name", ["ARS343", "SSP333", "SSP222", "SSP433"])
def test_imex_default_constraint_paths(cts, name, dtype, tol):
    # SSP names take the IMEXSSPRKCache path by default (imex_tableaus.jl:646); Float32 states with Float64 tableau
    uo, uc, integ, _, _ = _run_imex_pair(cts, name, dict(max_iters=2), 2, 0.02, dtype=dtype, fusion=False)
    np.testing.assert_array_equal(uc, uo)
    uo, uc, *_ = _run_imex_pair(cts, name, dict(max_iters=1), 2, 0.02, dtype=dtype, fusion=True)
    np.testing.assert_array_equal(uc, uo)


def test_nonlinear_column_direct_newton(cts):
    uo, uc, *_ = _run_imex_pair(cts, "ARS343", dict(max_iters=2), 2, 0.01, nonlinear=True)
    np.testing.assert_array_equal(uc, uo)


@pytest.mark.parametrize("dtype,tol", [(np.float64, 1e-12), (np.float32, 1e-5)])
@pytest.mark.parametrize("jfnk", [True, False])
def test_ssp333_gmres_newton_krylov(cts, dtype, tol, jfnk):
    # config C4 in miniature: SSP333 (SSPRK path) + KrylovMethod(GMRES, ForwardDiffJVP, ConstantForcing), W as left preconditioner
    uo, uc, integ, cache_o, _ = _run_imex_pair(cts, "SSP333", dict(max_iters=2), 2, 0.01, dtype=dtype, nonlinear=True,
                                                krylov=dict(jfnk=jfnk, rtol=1e-6 if dtype == np.float64 else 1e-3),
                                                nx=4, ny=4, nlev=16)
    assert rel_err(uc, uo) <= 2 * tol  # two steps
    if dtype == np.float64:
        np.testing.assert_array_equal(uc, uo)  # reductions share one fixed tree => even JFNK is bit-exact
    st = integ.cache.stats()
    assert st["krylov_iters"] > 0
    # same number of Krylov iterations as the oracle's GMRES
    assert st["krylov_iters"] == sum(s["niter"] for s in cache_o["nm_cache"]["stats"])


def test_rejects_non_sdirk_with_newtimestep_and_bad_ssp_tableau(cts):
    model, op, T0 = _column_pair(cts)
    prob = cts.ODEProblem(op.ode_function(), dev(cts, T0), (0.0, 1.0), None)
    # async_utils.jl:93-98 (SSP333's implicit diagonal is not bitwise uniform)
    with pytest.raises(cts.CTSError):
        cts.init(prob, cts.IMEXAlgorithm(cts.SSP333(), cts.NewtonsMethod(update_j=cts.UpdateEvery(cts.NewTimeStep))), dt=0.01)
    # imex_ssprk.jl:51-64: ARS343 is not a low-storage SSPRK tableau
    with pytest.raises(cts.CTSError):
        cts.init(prob, cts.IMEXAlgorithm(cts.ARS343(), cts.NewtonsMethod(), cts.SSP), dt=0.01)


def test_cache_buffer_counts(cts):
    # test/unit/cache_sizes.jl:24-30: ARS343 keeps 4 T_exp and 3 T_imp (T_lim only when a limiter exists here)
    model, op, T0 = _column_pair(cts)
    prob = cts.ODEProblem(op.ode_function(), dev(cts, T0), (0.0, 1.0), None)
    integ = cts.init(prob, cts.IMEXAlgorithm(cts.ARS343(), cts.NewtonsMethod()), dt=0.01)
    c = integ.cache
    assert all(c.buffer(f"T_exp{i}") is not None for i in (1, 2, 3, 4))
    assert c.buffer("T_imp1") is None and all(c.buffer(f"T_imp{i}") is not None for i in (2, 3, 4))
    assert c.nbuffers() == 2 + 4 + 3 + 2  # U, temp, T_exp x4, T_imp x3, Newton Δx and f


# ------------------------------------------------------------------ generic Python hooks ---------
class DenseJacobianGPU:
    """`j isa DenseMatrix` test path (newtons_method.jl:630-631): dense W solved with LU on the device."""

    def __init__(self, n, dtype):
        import torch
        self.M = torch.zeros((n, n), dtype=torch.float64 if dtype == np.float64 else torch.float32, device="cuda")

    def zero(self):
        return DenseJacobianGPU(self.M.shape[0], np.float64 if self.M.dtype.itemsize == 8 else np.float32)

    def ldiv(self, x, b):
        import torch
        x.tensor.copy_(torch.linalg.solve(self.M, b.tensor))


@pytest.mark.parametrize("name,newton_iters", [("ARS343", 2), ("ARS232", 1), ("SSP333", 1)])
def test_config1_ark_analytic_sys_python_hooks(cts, name, newton_iters):
    # BASELINE config 1 (perf/benchmark.jl:9-11) through user-supplied (Python) hooks and a dense Jacobian type
    import torch
    case = P.ark_analytic_sys()
    A = torch.from_numpy(np.array([[1, -1, 1], [-1, 2, 1], [0, -1, 2.0]]) @ np.diag([-0.5, -0.1, -100.0])
                         @ (np.array([[5, 1, -3], [2, 2, -2], [1, 1, 1.0]]) / 4)).cuda()

    def T_imp(du, u, p, t):
        du.tensor.copy_(A @ u.tensor)

    def Wfact(W, u, p, dtg, t):
        W.M.copy_(dtg * A - torch.eye(3, dtype=torch.float64, device="cuda"))

    f = cts.ClimaODEFunction(T_imp=cts.ODEFunction(T_imp, jac_prototype=DenseJacobianGPU(3, np.float64), Wfact=Wfact))
    u = dev(cts, case.Y0)
    sol = cts.solve(cts.ODEProblem(f, u, (0.0, case.t_end), None),
                    cts.IMEXAlgorithm(getattr(cts, name)(), cts.NewtonsMethod(max_iters=newton_iters)), dt=0.01)
    alg_o = R.imex_algorithm(name, R.NewtonsMethod(max_iters=newton_iters))
    uo = case.Y0.copy()
    _, step = R.make_stepper(case.f, alg_o, uo)
    R.solve(uo, None, (0.0, case.t_end), 0.01, step)
    assert sol.t[-1] == case.t_end and len(sol.u) == 2
    assert rel_err(u.to_host(), uo) <= 1e-12
    assert np.all(np.abs(u.to_host() - case.analytic_sol(case.t_end)) < 5e-3)


def test_hook_call_sites_match_reference_order(cts):
    # the engine must call user hooks exactly where imex_ark.jl does: record the trace on both sides
    def make(trace, wrap):
        def T_exp(du, u, p, t):
            trace.append(("T_exp", round(t, 12)))
            wrap(du)[...] = 1.0
        def T_imp(du, u, p, t):
            trace.append(("T_imp", round(t, 12)))
            wrap(du)[...] = -wrap(u)
        def dss(u, p, t):
            trace.append(("dss", round(t, 12)))
        def cache(u, p, t):
            trace.append(("cache", round(t, 12)))
        def cache_imp(u, p, t):
            trace.append(("cache_imp", round(t, 12)))
        def constrain(u, p, t):
            trace.append(("constrain", round(t, 12)))
        def init_imp(u, p, dtg):
            trace.append(("init_imp", round(dtg, 12)))
        return T_exp, T_imp, dss, cache, cache_imp, constrain, init_imp

    class DiagJ:  # MockDiagonalJacobian (test/performance/allocations.jl:26-30)
        def __init__(self, n, gpu):
            self.n, self.gpu, self.dtg = n, gpu, 0.0
        def zero(self):
            return DiagJ(self.n, self.gpu)
        def ldiv(self, x, b):
            if self.gpu:
                x.tensor.copy_(b.tensor / (-self.dtg - 1))
            else:
                x[...] = b / (-self.dtg - 1)

    def Wfact_factory(trace):
        def Wfact(W, u, p, dtg, t):
            trace.append(("Wfact", round(dtg, 12), round(t, 12)))
            W.dtg = dtg
        return Wfact

    n = 16
    u0 = np.linspace(1, 2, n)
    for name in ("ARS343", "ARS222", "SSP333", "SSP222"):
        tr_o, tr_c = [], []
        ho = make(tr_o, lambda a: a)
        fo = R.ClimaODEFunction(T_exp=ho[0], T_imp=R.ODEFunction(ho[1], jac_prototype=DiagJ(n, False), Wfact=Wfact_factory(tr_o)),
                                dss=ho[2], cache=ho[3], cache_imp=ho[4], constrain_state=ho[5], initialize_imp=ho[6],
                                update_constrain_state=R.UpdateEvery("WithDSS"))
        hc = make(tr_c, lambda a: a.tensor)
        fc = cts.ClimaODEFunction(T_exp=hc[0], T_imp=cts.ODEFunction(hc[1], jac_prototype=DiagJ(n, True), Wfact=Wfact_factory(tr_c)),
                                  dss=hc[2], cache=hc[3], cache_imp=hc[4], constrain_state=hc[5], initialize_imp=hc[6],
                                  update_constrain_state=cts.UpdateEvery(cts.WithDSS))
        uo = u0.copy()
        _, step = R.make_stepper(fo, R.imex_algorithm(name, R.NewtonsMethod(max_iters=2)), uo)
        step(uo, None, 0.0, 0.1)
        uc = dev(cts, u0)
        integ = cts.init(cts.ODEProblem(fc, uc, (0.0, 1.0), None),
                         cts.IMEXAlgorithm(getattr(cts, name)(), cts.NewtonsMethod(max_iters=2)), dt=0.1)
        del tr_c[:]  # `init` fires cache!(u0, p, t0) (functions.jl:159-161)
        cts.step_(integ)
        assert tr_c == tr_o, name
        assert rel_err(uc.to_host(), uo) <= 1e-14


# ------------------------------------------------------------------ LSRK --------------------------
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("fusion", [False, True])
@pytest.mark.parametrize("n", [1000, 4099])
def test_lsrk54_advection_bit_exact(cts, dtype, fusion, n):
    # config C2 in miniature
    u0 = M.fill_hash(n, 1, dtype=dtype)
    adv = M.Advection(n, dtype=dtype)
    dt = 0.5 / n
    uo, du = u0.copy(), np.zeros(n, dtype)
    tab = R.lsrk54_tableau(dtype)
    op = cts.AdvectionOperator(n, dtype=dtype)
    uc = dev(cts, u0)
    integ = cts.init(cts.ODEProblem(op.ode_function(), uc, (0.0, 1.0), None), cts.LSRK54CarpenterKennedy(), dt=dt)
    integ.cache.set_fusion(fusion)
    t = 0.0
    for _ in range(4):
        R.lsrk_step(uo, du, None, t, dt, adv.incr, tab)
        cts.step_(integ)
        t += dt
    np.testing.assert_array_equal(integ.u.to_host(), uo)
    np.testing.assert_array_equal(integ.cache.buffer("du").to_host(), du)
    assert (integ.cache.stats()["fused_stage_launches"] > 0) == fusion


def test_lsrk54_python_incrementing_function_order4(cts):
    # test/solvers/lsrk.jl:11-25 with a user (Python) IncrementingODEFunction on the device
    f_o, u0, tspan, p, sol = P.linear_prob()

    def f(du, u, p, t, alpha=True, beta=False):
        du.tensor.copy_(alpha * p * u.tensor + beta * du.tensor)

    errs, dts = [], [0.5 ** k for k in range(4, 8)]
    for dt in dts:
        u = dev(cts, u0)
        cts.solve(cts.ODEProblem(cts.IncrementingODEFunction(f), u, tspan, p), cts.LSRK54CarpenterKennedy(), dt=dt)
        errs.append(abs(u.to_host()[0] - sol(u0, p, tspan[1])[0]))
    order = np.polyfit(np.log(dts), np.log(errs), 1)[0]
    assert abs(order - 4) < 0.05


# ------------------------------------------------------------------ full-size properties ---------
def test_full_size_fused_equals_unfused_checksum(cts):
    # BASELINE size for one GPU is 2^21 columns x 64 levels; here a 2^17-column slice keeps the test short while
    # exercising > 1 wave of tiles: fused and un-fused paths must agree bit for bit (checksum of the state bits).
    nlev, nx, ny = 64, 2 ** 9, 2 ** 8
    K = cts.B200Vector.zeros((ny + 2) * nx).fill_hash(2, scale=0.25, shift=2.0)
    S = dev(cts, M.column_inputs(nlev, 1, 1, 0, 1, 0)[2])
    sums = []
    for fusion, mf in ((False, False), (True, False), (True, True)):
        op = cts.ColumnOperator(nlev, nx, ny, K, S_lev=S, nu=1.0, halo=1).set_matrix_free(mf)
        u = cts.B200Vector.zeros(op.ndof).fill_hash(3)
        op.dss(u, None, 0.0)
        integ = cts.init(cts.ODEProblem(op.ode_function(), u, (0.0, 1.0), None),
                         cts.IMEXAlgorithm(cts.ARS343(), cts.NewtonsMethod(max_iters=1, update_j=cts.UpdateEvery(cts.NewTimeStep))), dt=0.02)
        integ.cache.set_fusion(fusion)
        for _ in range(2):
            cts.step_(integ)
        import torch
        bits = u.tensor.view(torch.int64)
        sums.append((int(bits.sum().item()), int((bits ^ (bits >> 7)).sum().item())))
        assert bool(torch.isfinite(u.tensor).all())
    assert sums[0] == sums[1] == sums[2]
