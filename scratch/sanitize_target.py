"""Small invocations of every kernel family on cuda:0, meant to run under compute-sanitizer
(`scratch/sanitize.sh`: memcheck, racecheck, synccheck, initcheck).  Sizes are CI-sized: a few tiles per kernel, ragged
last tiles, both state types.  Prints one line per case; any sanitizer finding is in the tool's own log."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import cts_b200 as cts


def column_case(dtype, mf, fusion, nlev=64, nx=40, ny=9, nonlinear=False, alg="ARS343", krylov=False, graph=True):
    if krylov:  # BASELINE configs[3] in small: nonlinear column diffusion only, no halo
        K = cts.B200Vector.zeros(ny * nx, dtype).fill_hash(2, 0, 0.25, 2.0)
        op = cts.ColumnOperator(nlev, nx, ny, K, nu=0.0, halo=0, nonlinear=True)
        u = cts.B200Vector.zeros(op.ndof, dtype).fill_hash(3, 0, 0.02, 25.0)
        fo = op.ode_function(explicit=False)
    else:
        K = cts.B200Vector.zeros((ny + 2) * nx, dtype).fill_hash(2, 0, 0.25, 2.0)
        z = (np.arange(nlev) + 1) / (nlev + 1)
        S = cts.B200Vector.from_host((5.0 * np.exp(-((z - 0.3) ** 2) / 0.01)).astype(dtype))
        op = cts.ColumnOperator(nlev, nx, ny, K, S_lev=S, nu=1.0, halo=1, nonlinear=nonlinear).set_matrix_free(mf)
        u = cts.B200Vector.zeros(op.ndof, dtype).fill_hash(3)
        op.dss(u, None, 0.0)
        fo = op.ode_function()
    if krylov:
        km = cts.KrylovMethod(jacobian_free_jvp=cts.ForwardDiffJVP(), forcing_term=cts.ConstantForcing(1e-2))
        nm = cts.NewtonsMethod(max_iters=2, krylov_method=km)
    else:
        nm = cts.NewtonsMethod(max_iters=1, update_j=cts.UpdateEvery(cts.NewTimeStep))
    name = getattr(cts, alg)()
    integ = cts.init(cts.ODEProblem(fo, u, (0.0, 1e9), None), cts.IMEXAlgorithm(name, nm), dt=2e-3, save=False)
    integ.cache.set_fusion(fusion)
    integ.cache.set_graph(graph)
    for _ in range(2):
        cts.step_(integ)
    h = u.to_host()
    assert np.isfinite(h).all()
    return float(np.abs(h).max())


def rosenbrock_case(dtype, fusion):
    nlev, nx, ny = 64, 33, 7
    K = cts.B200Vector.zeros(ny * nx, dtype).fill_hash(2, 0, 0.25, 2.0)
    op = cts.ColumnOperator(nlev, nx, ny, K, nu=0.0, halo=0)
    u = cts.B200Vector.zeros(op.ndof, dtype).fill_hash(3)
    integ = cts.init(cts.ODEProblem(op.ode_function(explicit=False), u, (0.0, 1e9), None),
                     cts.RosenbrockAlgorithm(cts.tableau(cts.SSPKnoth())), dt=dtype(0.01) if dtype == np.float32 else 0.01, save=False)
    integ.cache.set_fusion(fusion)
    for _ in range(2):
        cts.step_(integ)
    h = u.to_host()
    assert np.isfinite(h).all()
    return float(np.abs(h).max())


def lsrk_case(dtype, n):
    u = cts.B200Vector.zeros(n, dtype).fill_hash(1)
    integ = cts.init(cts.ODEProblem(cts.AdvectionOperator(n, dtype=dtype).ode_function(), u, (0.0, 1.0), None),
                     cts.LSRK54CarpenterKennedy(), dt=0.5 / n, save=False)
    for _ in range(2):
        cts.step_(integ)
    h = u.to_host()
    assert np.isfinite(h).all()
    return float(np.abs(h).max())


def vector_case(dtype, n):
    a = cts.B200Vector.zeros(n, dtype).fill_hash(5)
    b = cts.B200Vector.zeros(n, dtype).fill_hash(6)
    return a.dot(b) + a.norm()


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    cases = []
    for dt in (np.float64, np.float32):
        cases += [(f"ars343 fused stored-W {dt.__name__}", lambda dt=dt: column_case(dt, False, True)),
                  (f"ars343 fused matrix-free {dt.__name__}", lambda dt=dt: column_case(dt, True, True)),
                  (f"ars343 generic passes {dt.__name__}", lambda dt=dt: column_case(dt, False, False, graph=False)),
                  (f"ars343 fused 32 levels {dt.__name__}", lambda dt=dt: column_case(dt, False, True, nlev=32, nx=6, ny=5)),
                  (f"ars343 fused 48 levels (5-tile kernel) {dt.__name__}", lambda dt=dt: column_case(dt, False, True, nlev=48, nx=7, ny=5)),
                  (f"ssp333 jfnk {dt.__name__}", lambda dt=dt: column_case(dt, False, True, nx=16, ny=6, nonlinear=True, alg="SSP333", krylov=True)),
                  (f"rosenbrock fused {dt.__name__}", lambda dt=dt: rosenbrock_case(dt, True)),
                  (f"rosenbrock passes {dt.__name__}", lambda dt=dt: rosenbrock_case(dt, False)),
                  (f"lsrk54 {dt.__name__}", lambda dt=dt: lsrk_case(dt, (1 << 14) + 37)),
                  (f"reductions {dt.__name__}", lambda dt=dt: vector_case(dt, (1 << 15) + 5))]
    for name, fn in cases:
        if which != "all" and which not in name:
            continue
        print(f"{name}: ok {fn():.6g}", flush=True)
    print("kernels launched", cts.Context.default().launch_count())
