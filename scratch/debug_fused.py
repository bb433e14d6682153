import sys, os
sys.path[:0] = [os.getcwd(), os.path.join(os.getcwd(), "oracle")]
import numpy as np, torch
import cts_b200 as cts, cts_models as M
nlev = 64
def run(nx, ny, fusion, mf, nsteps=1, Kmode="hash"):
    K = cts.B200Vector.zeros((ny + 2) * nx).fill_hash(2, scale=0.25, shift=2.0)
    S = cts.B200Vector.from_host(M.column_inputs(nlev, 1, 1, 0, 1, 0)[2])
    op = cts.ColumnOperator(nlev, nx, ny, K, S_lev=S, nu=1.0, halo=1).set_matrix_free(mf)
    u = cts.B200Vector.zeros(op.ndof).fill_hash(3)
    op.dss(u, None, 0.0)
    integ = cts.init(cts.ODEProblem(op.ode_function(), u, (0.0, 1.0), None),
                     cts.IMEXAlgorithm(cts.ARS343(), cts.NewtonsMethod(max_iters=1, update_j=cts.UpdateEvery(cts.NewTimeStep))), dt=0.02)
    integ.cache.set_fusion(fusion)
    for _ in range(nsteps):
        cts.step_(integ)
    out = {"u": u.to_host()}
    for nm in ("U", "T_exp1", "T_exp2", "T_imp2", "T_imp3", "T_imp4"):
        out[nm] = integ.cache.buffer(nm).to_host()
    return out, (nx, ny)
for nx, ny, ns in ((8, 6, 1),):
    a, _ = run(nx, ny, False, False, ns)
    b, _ = run(nx, ny, True, False, ns)
    c, _ = run(nx, ny, True, True, ns)
    a2, _ = run(nx, ny, False, False, ns)
    print('unfused repeatable:', all(np.array_equal(a[k], a2[k]) for k in a), 'finite', np.isfinite(a['u']).all(), np.isfinite(b['u']).all())
    own = slice(nx * nlev, -nx * nlev)
    for nm in a:
        for lab, o in (("fused", b), ("mfree", c)):
            d = np.nonzero(a[nm][own] != o[nm][own])[0]
            if d.size:
                i = d[0]
                print(f"nx={nx} ny={ny} {nm} unfused vs {lab}: {d.size} mismatches, first at owned idx {i} (col {i//nlev}, lev {i%nlev}) "
                      f"{a[nm][own][i]!r} vs {o[nm][own][i]!r}; cols hit: {np.unique(d//nlev)[:10]}")
    print(f"nx={nx} ny={ny} done")
