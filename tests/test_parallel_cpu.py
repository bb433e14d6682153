"""Host-side logic of the N>1 path on CPU: world_size-2/4 gloo processes run the ORACLE on column slabs with a
halo exchange over torch.distributed and must reproduce the single-process result bit for bit.  This covers the
slab partition, neighbour/ring logic, ghost-row layout and global-index hashing used by the GPU path (the GPU
halo exchange itself is tested in tests/test_multi_gpu.py)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import cts_models as M
import cts_ref as R

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, nlev, nx, ny_g, nsteps, ret):
    sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from cts_b200.parallel import SlabPartition  # host-side partition logic of the product
    part = SlabPartition(ny_g, world, rank)
    K, T0, S = M.column_inputs(nlev, nx, ny_g, part.row0, part.ny_local, 1)
    assert np.array_equal(part.local_rows(1), np.arange(part.row0 - 1, part.row0 + part.ny_local + 1) % ny_g)

    def exchange(first_row, last_row):
        # ring: my first owned row -> lo neighbour's upper ghost, my last owned row -> hi neighbour's lower ghost
        send_lo, send_hi = torch.from_numpy(first_row.copy()), torch.from_numpy(last_row.copy())
        recv_lo, recv_hi = torch.empty_like(send_lo), torch.empty_like(send_hi)
        if world == 1:
            return last_row, first_row
        reqs = [dist.isend(send_lo, part.lo, tag=1), dist.isend(send_hi, part.hi, tag=2),
                dist.irecv(recv_hi, part.hi, tag=1), dist.irecv(recv_lo, part.lo, tag=2)]
        for r in reqs:
            r.wait()
        return recv_lo.numpy(), recv_hi.numpy()

    model = M.ColumnModel(nlev, nx, part.ny_local, K, S_lev=S, nu=1.0, halo=1, exchange=exchange)
    u = T0.copy()
    model.dss(u, None, 0.0)
    _, step = R.make_stepper(model.ode_function(), R.imex_algorithm("ARS343", R.NewtonsMethod(max_iters=1)), u)
    t = 0.0
    for _ in range(nsteps):
        step(u, None, t, 0.02)
        t += 0.02
    off, cnt = part.owned_window(nx, nlev)
    own = torch.from_numpy(u[off:off + cnt].copy())
    parts = [torch.empty_like(own) for _ in range(world)] if rank == 0 else None
    dist.gather(own, parts, dst=0)
    if rank == 0:
        ret.put(torch.cat(parts).numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_sharded_oracle_matches_single_process(world):
    nlev, nx, ny_g, nsteps = 8, 4, 8, 3
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, 29600 + world, nlev, nx, ny_g, nsteps, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = q.get(timeout=300)
    for p in procs:
        p.join(timeout=300)
        assert p.exitcode == 0
    K, T0, S = M.column_inputs(nlev, nx, ny_g, 0, ny_g, 1)
    model = M.ColumnModel(nlev, nx, ny_g, K, S_lev=S, nu=1.0, halo=1)
    u = T0.copy()
    model.dss(u, None, 0.0)
    _, step = R.make_stepper(model.ode_function(), R.imex_algorithm("ARS343", R.NewtonsMethod(max_iters=1)), u)
    t = 0.0
    for _ in range(nsteps):
        step(u, None, t, 0.02)
        t += 0.02
    np.testing.assert_array_equal(got, u[nx * nlev:-nx * nlev])


def test_slab_partition_properties():
    sys.path.insert(0, ROOT)
    from cts_b200.parallel import SlabPartition
    for world in (1, 2, 4, 8):
        rows = np.concatenate([SlabPartition(1024 * world, world, r).local_rows(0) for r in range(world)])
        assert np.array_equal(rows, np.arange(1024 * world))  # every row owned exactly once, in order
        for r in range(world):
            p = SlabPartition(1024 * world, world, r)
            assert SlabPartition(1024 * world, world, p.hi).lo == r and SlabPartition(1024 * world, world, p.lo).hi == r
            assert p.local_rows(1)[0] == (p.row0 - 1) % (1024 * world)
    with pytest.raises(ValueError):
        SlabPartition(10, 4, 0)
