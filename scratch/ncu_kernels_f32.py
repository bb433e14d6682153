import sys, os
sys.path[:0] = [os.getcwd()]
import numpy as np, torch
import cts_b200 as cts
ctx = cts.Context.default()
nlev, nx, ny = 64, 2048, 1024
K = cts.B200Vector.zeros((ny + 2) * nx, np.float32).fill_hash(2, scale=0.25, shift=2.0)
op = cts.ColumnOperator(nlev, nx, ny, K, nu=1.0, halo=1)
u = cts.B200Vector.zeros(op.ndof, np.float32).fill_hash(3)
integ = cts.init(cts.ODEProblem(op.ode_function(), u, (0.0, 1e9), None),
                 cts.IMEXAlgorithm(cts.ARS343(), cts.NewtonsMethod(max_iters=1, update_j=cts.UpdateEvery(cts.NewTimeStep))), dt=0.02, save=False)
for _ in range(2):
    cts.step_(integ)
ctx.sync()
print("done")
