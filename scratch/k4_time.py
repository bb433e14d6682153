import sys, os, ctypes as C
sys.path[:0] = [os.getcwd()]
import numpy as np, torch
import cts_b200 as cts
from cts_b200 import _lib as L
ctx = cts.Context.default()
n = 1 << 27; nlev = 64
vs = [cts.B200Vector.zeros(n).fill_hash(10 + k) for k in range(5)]
vs[1].tensor.add_(-4.0)
W = L.Tridiag(nlev, n // nlev, vs[0].ptr, vs[1].ptr, vs[2].ptr)
def run(): ctx.check(ctx.lib.cts_tridiag_solve_batched(ctx.handle, 1, C.byref(W), vs[3].ptr, vs[4].ptr))
for _ in range(3): run()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); a.record()
for _ in range(10): run()
b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b) / 10
print(f"loader={os.environ.get('CTS_TILE_LOADER','default')} tridiag 2^21 cols: {ms:.3f} ms  {5*8*n/ms/1e6:.0f} GB/s")
