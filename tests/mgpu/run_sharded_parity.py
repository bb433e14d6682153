"""Launched by torchrun (one rank per GPU): config C5 in miniature.  Every rank steps its slab of the column
problem on its GPU (library dss! = NVLink peer-memory halo kernel, or NCCL send/recv with CTS_HALO_NCCL=1); rank 0 gathers
the owned rows and compares them with the single-process numpy oracle of the GLOBAL problem -- bit for bit, including the
sharded JFNK step: reductions use the shard-invariant chunk layout (cts_set_reduction_layout), so dots, norms, the JVP step
size and GMRES see the same bits for any rank count."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
import cts_b200 as cts  # noqa: E402
import cts_models as M  # noqa: E402
import cts_ref as R  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)  # a non-default stream shared by torch and the library (as in bench.py)
    ctx = cts.Context(local)
    cts.attach_communicator(ctx, dist)
    log = lambda *a: print(f"[rank {rank}]", *a, flush=True) if os.environ.get("CTS_MGPU_VERBOSE") else None
    nx, nsteps, dt = 16, 3, 0.01
    ok = True
    # (name, newton, nonlinear, krylov, rows per rank, levels); 40 rows per rank: the halo exchange of dss! overlaps the
    # interior rows of T_exp! (cts_column_dss_T_exp_overlapped), 8 rows per rank: the sequential path
    for name, newton, nonlinear, krylov, ny_rank, nlev in (("ARS343", dict(max_iters=1), False, False, 8, 64),
                                                           ("ARS343", dict(max_iters=1), False, False, 40, 16),
                                                           ("SSP333", dict(max_iters=2), True, True, 8, 16)):
        # (16 levels for the Krylov case: the frozen-coefficient preconditioner is only effective in the mildly stiff regime)
        ny_g = ny_rank * world
        part = cts.SlabPartition(ny_g, world, rank)
        K, T0, S = M.column_inputs(nlev, nx, ny_g, part.row0, part.ny_local, 1)
        dev = lambda a: cts.B200Vector.from_host(a, ctx)
        op = cts.ColumnOperator(nlev, nx, part.ny_local, dev(K), S_lev=dev(S), nu=1.0, halo=1, nonlinear=nonlinear, ctx=ctx)
        u = dev(T0)
        op.dss(u, None, 0.0)
        off, cnt = part.owned_window(nx, nlev)
        layout = part.reduction_layout(nx, nlev)
        ctx.set_reduction_layout(*layout)
        nm_kw = dict(newton)
        if krylov:
            nm_kw["krylov_method"] = cts.KrylovMethod(jacobian_free_jvp=cts.ForwardDiffJVP(), forcing_term=cts.ConstantForcing(1e-3), itmax=30)
        integ = cts.init(cts.ODEProblem(op.ode_function(), u, (0.0, 1e9), None),
                         cts.IMEXAlgorithm(getattr(cts, name)(), cts.NewtonsMethod(**nm_kw)), dt=dt, save=False)
        for k in range(nsteps):
            cts.step_(integ)
            log(name, 'step', k, 'done')
        nrm = u.norm()  # all-reduced over owned rows
        log(name, 'norm', nrm)
        own = u.tensor[off:off + cnt].contiguous()
        parts = [torch.empty_like(own) for _ in range(world)] if rank == 0 else None
        torch.cuda.synchronize()
        dist.gather(own, parts, dst=0)
        torch.cuda.synchronize()
        log(name, 'gathered')
        ctx.set_reduction_window(0, 0)
        if rank == 0:
            got = torch.cat(parts).cpu().numpy()
            # oracle: the global problem in one process (periodic ghost rows, halo = 1)
            Kg, Tg, Sg = M.column_inputs(nlev, nx, ny_g, 0, ny_g, 1)
            model = M.ColumnModel(nlev, nx, ny_g, Kg, S_lev=Sg, nu=1.0, halo=1, nonlinear=nonlinear)
            uo = Tg.copy()
            model.dss(uo, None, 0.0)
            R.REDUCTION_WINDOW = (nx * nlev, ny_g * nx * nlev)  # norms/dots over owned rows only, like the shards
            chunk_saved, R.REDUCTION_CHUNK = R.REDUCTION_CHUNK, layout[4]
            o_kw = dict(newton)
            if krylov:
                o_kw["krylov_method"] = R.KrylovMethod(jacobian_free_jvp=R.ForwardDiffJVP(), forcing_term=R.ConstantForcing(1e-3), itmax=30)
            _, step = R.make_stepper(model.ode_function(), R.imex_algorithm(name, R.NewtonsMethod(**o_kw)), uo)
            t = 0.0
            for _ in range(nsteps):
                step(uo, None, t, dt)
                t += dt
            R.REDUCTION_WINDOW = None
            onorm = None
            want = uo[nx * nlev:-nx * nlev]
            err = float(np.max(np.abs(got - want)) / np.max(np.abs(want)))
            exact = bool(np.array_equal(got, want))
            R.REDUCTION_WINDOW = (nx * nlev, ny_g * nx * nlev)
            onorm = R.norm2(uo)   # the oracle's tree over the owned rows == the sharded tree, bit for bit
            R.REDUCTION_WINDOW = None
            R.REDUCTION_CHUNK = chunk_saved
            nerr = abs(nrm - onorm) / onorm
            good = exact and nrm == onorm
            ok = ok and good
            st = integ.cache.stats()
            good = good and np.isfinite(got).all() and (not krylov or st["krylov_iters"] < 30 * st["newton_iters"])
            ok = ok and good
            print(f"[{world} ranks, halo={ctx.halo_path()}, graph_launches={integ.cache.graph_stats()['graph_launches']}] {name} rows/rank={ny_rank} nonlinear={nonlinear} krylov={krylov}: rel err {err:.3e} bit-exact={exact} krylov_iters={st['krylov_iters']} "
                  f"norm rel err {nerr:.2e} -> {'OK' if good else 'FAIL'}", flush=True)
        dist.barrier()
    # ---- EveryXWallTimeSeconds (Callbacks.jl:40-90): every rank has its own (skewed) clock; the max all-reduce
    # (`cts_allreduce_max_f64_host`) makes all ranks fire on the same steps -- those of the fastest clock
    part = cts.SlabPartition(8 * world, world, rank)
    K, T0, S = M.column_inputs(16, nx, 8 * world, part.row0, part.ny_local, 1)
    op = cts.ColumnOperator(16, nx, part.ny_local, cts.B200Vector.from_host(K, ctx), S_lev=cts.B200Vector.from_host(S, ctx),
                            nu=1.0, halo=1, ctx=ctx)
    u = cts.B200Vector.from_host(T0, ctx)
    op.dss(u, None, 0.0)
    ticks = iter(np.arange(0.0, 1000.0, 0.1 * (rank + 1)))   # rank r's clock advances 0.1*(r+1) s per query
    hits = []
    cb = cts.EveryXWallTimeSeconds(lambda integ: hits.append(integ.step), 1.0, ctx, clock=lambda: float(next(ticks)))
    integ = cts.init(cts.ODEProblem(op.ode_function(), u, (0.0, 0.2), None),
                     cts.IMEXAlgorithm(cts.ARS343(), cts.NewtonsMethod(max_iters=1)), dt=0.01, callback=cb)
    cts.solve_(integ)
    fast = 0.1 * world                                        # the max over ranks is the fastest clock
    expect, nxt = [], fast * 0 + 1.0
    for step in range(1, 21):
        now = fast * step
        if now >= nxt:
            while now >= nxt:
                nxt += 1.0
            expect.append(step)
    got = torch.tensor(hits + [-1] * (32 - len(hits)), device=f"cuda:{local}")
    allhits = [torch.empty_like(got) for _ in range(world)] if rank == 0 else None
    dist.gather(got, allhits, dst=0)
    if rank == 0:
        same = all(bool(torch.equal(h, allhits[0])) for h in allhits)
        good = same and hits == expect
        ok = ok and good
        print(f"[{world} ranks] EveryXWallTimeSeconds with skewed clocks: hits {hits} expected {expect} identical on all ranks={same} -> {'OK' if good else 'FAIL'}", flush=True)
    dist.barrier()
    flag = torch.tensor([1 if ok else 0], device=f"cuda:{local}")
    dist.broadcast(flag, 0)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
