import os, sys
import numpy as np, torch, torch.distributed as dist
ROOT = os.getcwd(); sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
import cts_b200 as cts, cts_models as M
rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1: dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
ctx = cts.Context(local)
if world > 1: cts.attach_communicator(ctx, dist)
nlev, nx, ny_g = 64, 16, 8 * world
part = cts.SlabPartition(ny_g, world, rank)
K, T0, S = M.column_inputs(nlev, nx, ny_g, part.row0, part.ny_local, 1)
dev = lambda a: cts.B200Vector.from_host(a, ctx)
op = cts.ColumnOperator(nlev, nx, part.ny_local, dev(K), S_lev=dev(S), nu=1.0, halo=1, nonlinear=True, ctx=ctx)
u = dev(T0); op.dss(u, None, 0.0)
if os.environ.get("WINDOW", "1") == "1":
    off, cnt = part.owned_window(nx, nlev); ctx.set_reduction_window(off, cnt)
km = cts.KrylovMethod(jacobian_free_jvp=cts.ForwardDiffJVP(), forcing_term=cts.ConstantForcing(1e-6), itmax=int(os.environ.get("ITMAX", "60")))
integ = cts.init(cts.ODEProblem(op.ode_function(), u, (0.0, 1e9), None), cts.IMEXAlgorithm(cts.SSP333(), cts.NewtonsMethod(max_iters=2, krylov_method=km)), dt=0.02, save=False)
for k in range(2):
    cts.step_(integ)
    print(f"[rank {rank}/{world}] step {k} stats {integ.cache.stats()} norm {u.norm():.15g}", flush=True)
if world > 1: dist.destroy_process_group()
