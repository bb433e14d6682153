"""Developer probe: per-step error of the CUDA-graph step against the oracle (graph on / off, private stream)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
import torch
import cts_b200 as cts
import cts_models as M
import cts_ref as R

def run(graph, ny, dts):
    stream = torch.cuda.Stream()
    with torch.cuda.stream(stream):
        ctx = cts.Context(0)
        nlev, nx = 16, 8
        K, T0, S = M.column_inputs(nlev, nx, ny, 0, ny, 1)
        dev = lambda a: cts.B200Vector.from_host(a, ctx)
        op = cts.ColumnOperator(nlev, nx, ny, dev(K), S_lev=dev(S), nu=1.0, halo=1, ctx=ctx)
        u = dev(T0); op.dss(u, None, 0.0)
        model = M.ColumnModel(nlev, nx, ny, K, S_lev=S, nu=1.0, halo=1)
        uo = T0.copy(); model.dss(uo, None, 0.0)
        alg = cts.IMEXAlgorithm(cts.ARS343(), cts.NewtonsMethod(max_iters=1, update_j=cts.UpdateEvery(cts.NewTimeStep)))
        integ = cts.init(cts.ODEProblem(op.ode_function(), u, (0.0, 1e9), None), alg, dt=0.01, save=False)
        integ.cache.set_graph(graph)
        _, step = R.make_stepper(model.ode_function(), R.imex_algorithm("ARS343", R.NewtonsMethod(max_iters=1, update_j=R.UpdateEvery("NewTimeStep"))), uo)
        t = 0.0
        for k, dt in enumerate(dts):
            integ.dt = dt
            cts.step_(integ)
            step(uo, None, t, dt); t += dt
            got = u.to_host()
            print(f"graph={graph} ny={ny} step {k} dt={dt}: max abs err {np.max(np.abs(got-uo)):.3e} exact={np.array_equal(got,uo)} integ.t={integ.t} {integ.cache.graph_stats()}", flush=True)
        del integ, op
        ctx.sync()

for graph in (False, True):
    for ny in (8, 40):
        run(graph, ny, (0.01, 0.01, 0.004, 0.01))
