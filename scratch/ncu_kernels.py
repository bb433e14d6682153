import sys, os, ctypes as C
sys.path[:0] = [os.getcwd()]
import numpy as np, torch
import cts_b200 as cts
from cts_b200 import _lib as L
ctx = cts.Context.default()
nlev, nx, ny = 64, 2048, 1024
n = nlev * nx * (ny + 2)
vs = [cts.B200Vector.zeros(n).fill_hash(10 + k) for k in range(5)]
vs[1].tensor.add_(-4.0)
W = L.Tridiag(nlev, n // nlev, vs[0].ptr, vs[1].ptr, vs[2].ptr)
for _ in range(2):
    ctx.check(ctx.lib.cts_tridiag_solve_batched(ctx.handle, 1, C.byref(W), vs[3].ptr, vs[4].ptr))
del vs
K = cts.B200Vector.zeros((ny + 2) * nx).fill_hash(2, scale=0.25, shift=2.0)
mf = len(sys.argv) > 1 and sys.argv[1] == "mf"
op = cts.ColumnOperator(nlev, nx, ny, K, nu=1.0, halo=1).set_matrix_free(mf)
u = cts.B200Vector.zeros(op.ndof).fill_hash(3)
integ = cts.init(cts.ODEProblem(op.ode_function(), u, (0.0, 1e9), None),
                 cts.IMEXAlgorithm(cts.ARS343(), cts.NewtonsMethod(max_iters=1, update_j=cts.UpdateEvery(cts.NewTimeStep))), dt=0.02, save=False)
for _ in range(2):
    cts.step_(integ)
ctx.sync()
print("done")
