"""GPU parity of the Rosenbrock step (cts_rosenbrock_step, cts_axpy_chain) against the numpy oracle restatement of
src/solvers/rosenbrock.jl:136-239."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import cts_models as M  # noqa: E402
from ref_problems import R  # noqa: E402


@pytest.fixture(scope="module")
def cts():
    import cts_b200
    return cts_b200


def dev(cts, a):
    return cts.B200Vector.from_host(a)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("nterms,base,scale", [(0, True, None), (1, True, None), (3, True, None), (2, False, -0.37), (6, False, 1.7), (12, True, -2.0)])
def test_axpy_chain_matches_sequential_passes(cts, dtype, nterms, base, scale):
    ctx = cts.Context.default()
    n = 1031
    ft = 1 if dtype == np.float64 else 0
    xs = [(M.fill_hash(n, 20 + k) - 0.5).astype(dtype) for k in range(nterms)]
    cs = [0.3 * (k + 1) * (-1) ** k for k in range(nterms)]
    b = (M.fill_hash(n, 7) - 0.5).astype(dtype)
    want = b.copy() if base else np.zeros(n, dtype)
    for c, x in zip(cs, xs):
        want = (want.astype(np.float64) + c * x.astype(np.float64)).astype(dtype)   # `y .+= c .* x`
    if scale is not None:
        want = (want.astype(np.float64) * scale).astype(dtype)                       # `y .*= s`
    dx = [dev(cts, x) for x in xs]
    db = dev(cts, b)
    out = cts.B200Vector.zeros(n, dtype)
    coef = (C.c_double * max(1, nterms))(*cs)
    terms = (C.c_void_p * max(1, nterms))(*[d.ptr for d in dx])
    ctx.check(ctx.lib.cts_axpy_chain(ctx.handle, ft, n, out.ptr, db.ptr if base else None, nterms, coef, terms,
                                     0.0 if scale is None else scale, int(scale is not None)))
    np.testing.assert_array_equal(out.to_host(), want)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("fusion", [True, False])
@pytest.mark.parametrize("halo,nu,nonlinear,explicit,nlev,nx,ny", [(1, 0.7, False, True, 32, 6, 5), (0, 0.0, False, False, 32, 6, 5),
                                                                   (0, 0.5, True, True, 32, 6, 5), (1, 0.7, False, True, 64, 40, 9),
                                                                   (0, 0.5, True, False, 64, 33, 7)])
def test_rosenbrock_column_steps_bit_exact(cts, dtype, fusion, halo, nu, nonlinear, explicit, nlev, nx, ny):
    # fusion: T_imp! + the `.+=` chain + ldiv! of a stage as one tile kernel (cts_column_rosenbrock_stage) vs the three passes
    K, T0, S = M.column_inputs(nlev, nx, ny, 0, ny, halo, dtype)
    model = M.ColumnModel(nlev, nx, ny, K, S_lev=S, nu=nu, halo=halo, nonlinear=nonlinear)
    op = cts.ColumnOperator(nlev, nx, ny, dev(cts, K), S_lev=dev(cts, S), nu=nu, halo=halo, nonlinear=nonlinear)
    fo = R.ClimaODEFunction(T_exp=model.T_exp if explicit else None,
                            T_imp=R.ODEFunction(model.T_imp, jac_prototype=model.jac_prototype().zero(), Wfact=model.Wfact),
                            dss=model.dss if halo else R._noop)
    uo = T0.copy()
    _, step = R.make_stepper(fo, R.RosenbrockAlgorithm(R.ssp_knoth()), uo)
    u = dev(cts, T0)
    integ = cts.init(cts.ODEProblem(op.ode_function(explicit=explicit), u, (0.0, 1e9), None),
                     cts.RosenbrockAlgorithm(cts.tableau(cts.SSPKnoth())), dt=dtype(0.01) if dtype == np.float32 else 0.01, save=False)
    integ.cache.set_fusion(fusion)
    t = 0.0
    for _ in range(4):
        step(uo, None, t, 0.01 if dtype == np.float64 else float(np.float32(0.01)))
        cts.step_(integ)
        t = integ.t
    assert integ.cache.stats()["ldiv_calls"] == 12 and integ.cache.stats()["wfact_evals"] == 4
    assert integ.cache.stats()["fused_stage_launches"] == (12 if fusion else 0)
    np.testing.assert_array_equal(u.to_host(), uo)
    assert len(integ.cache.k) == 3 and integ.cache.buffer("fU_lim") is not None   # test/unit/cache_sizes.jl:46-54


def test_rosenbrock_python_hooks_tgrad_and_limiter_terms(cts):
    # generic hooks through the C ABI: T_imp! with tgrad, T_exp_T_lim! with a limited tendency, cache!/dss! call order
    n = 513
    lam = -3.0
    u0 = np.linspace(0.5, 1.5, n)

    class DiagW:
        def __init__(self, gpu):
            self.gpu, self.dtg = gpu, 0.0
        def zero(self):
            return DiagW(self.gpu)
        def ldiv(self, x, b):
            if self.gpu:
                x.tensor.copy_(b.tensor / (lam * self.dtg - 1))
            else:
                x[...] = b / (lam * self.dtg - 1)

    def mk(wrap, trace):
        def T_imp(du, u, p, t):
            wrap(du)[...] = lam * wrap(u) + np.float64(np.cos(t))
        def tgrad(dT, u, p, t):
            wrap(dT)[...] = 0 * wrap(u) - np.float64(np.sin(t))
        def T_exp(du, u, p, t):
            wrap(du)[...] = 0.2 * wrap(u)
        def T_lim(du, u, p, t):
            wrap(du)[...] = -0.1 * wrap(u)
        def dss(u, p, t):
            trace.append(("dss", round(t, 12)))
        def cache(u, p, t):
            trace.append(("cache", round(t, 12)))
        return T_imp, tgrad, T_exp, T_lim, dss, cache

    def Wfact(W, u, p, dtg, t):
        W.dtg = dtg

    tr_o, tr_c = [], []
    ho = mk(lambda a: a, tr_o)
    fo = R.ClimaODEFunction(T_exp=ho[2], T_lim=ho[3], T_imp=R.ODEFunction(ho[0], jac_prototype=DiagW(False), Wfact=Wfact, tgrad=ho[1]),
                            dss=ho[4], cache=ho[5])
    hc = mk(lambda a: a.tensor, tr_c)
    fc = cts.ClimaODEFunction(T_exp=hc[2], T_lim=hc[3], T_imp=cts.ODEFunction(hc[0], jac_prototype=DiagW(True), Wfact=Wfact, tgrad=hc[1]),
                              dss=hc[4], cache=hc[5])
    uo = u0.copy()
    _, step = R.make_stepper(fo, R.RosenbrockAlgorithm(R.ssp_knoth()), uo)
    uc = dev(cts, u0)
    integ = cts.init(cts.ODEProblem(fc, uc, (0.0, 100.0), None), cts.RosenbrockAlgorithm(cts.tableau(cts.SSPKnoth())), dt=0.05)
    t = 0.0
    for _ in range(3):
        step(uo, None, t, 0.05)
        cts.step_(integ)
        t += 0.05
    assert tr_c[0] == ("cache", 0.0)   # `init` fills the cache once (integrators.jl:252)
    assert tr_c[1:] == tr_o and len(tr_o) == 3 * (2 * 2 + 2)   # 2 later stages + end of step, dss! and cache! each
    assert float(np.max(np.abs(uc.to_host() - uo)) / np.max(np.abs(uo))) <= 1e-14
