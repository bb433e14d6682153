"""A/B timing of the T_exp! kernel variants (CTS_TEXP_KERNEL=0 flat, 1 row-tiled, 2 marching)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
import cts_b200 as cts
nlev, nx, ny = 64, 1024, 2048
K = cts.B200Vector.zeros((ny + 2) * nx).fill_hash(2, 0, 0.25, 2.0)
z = (np.arange(nlev) + 1) / (nlev + 1)
S = cts.B200Vector.from_host(5.0 * np.exp(-((z - 0.3) ** 2) / 0.01))
op = cts.ColumnOperator(nlev, nx, ny, K, S_lev=S, nu=1.0, halo=1)
u = cts.B200Vector.zeros(op.ndof).fill_hash(3)
du = cts.B200Vector.zeros(op.ndof)
for _ in range(3):
    op.T_exp(du, u, None, 0.1)
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize()
a.record()
for _ in range(20):
    op.T_exp(du, u, None, 0.1)
b.record()
torch.cuda.synchronize()
ms = a.elapsed_time(b) / 20
print(f"variant={os.environ.get('CTS_TEXP_KERNEL','default')} T_exp {ms:.4f} ms  {2*8*nlev*nx*ny/ms/1e6:.0f} GB/s  checksum {float(du.tensor.double().sum()):.17g}")
