#!/usr/bin/env python
"""bench.py -- IMEX step DOF-stage updates/s of the B200-native ClimaTimeSteppers state-update path.

    python bench.py --gpus N --steps K --warmup W            # our arm (N>1: launched by torchrun)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (oracle port) on host cores

Workload (BASELINE.json configs[2], the configuration the metric/target is quoted on): ARS343 IMEX on the
column-stacked vertical-diffusion problem, 2^21 columns x 64 levels = 2^27 DOF in Float64,
NewtonsMethod(max_iters = 1, update_j = UpdateEvery(NewTimeStep)), batched tridiagonal Wfact/ldiv!, explicit
source + 5-point horizontal coupling, dss! = ghost-row halo exchange.
A "step" is one `step!` of the integrator; the metric is N_dof * s / t_step with s = 4 stages.
Multi-GPU (BASELINE.json configs[4]): STRONG scaling -- the same 2^21 columns (a 2^11 x 2^10 horizontal grid) are
split into N contiguous slabs of 2^10/N rows, one per GPU; dss! = NVLink peer-memory halo kernel (NCCL send/recv if the
peers cannot be mapped).  The weak-scaling figure (2^27 DOF per GPU) is reported beside it in `extras.weak_scaling`.
Prints ONE JSON line on rank 0 (contract in the task statement; extra keys: roofline, cpu_baseline, kernels, ...).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

NLEV, NX, NY = 64, 2 ** 11, 2 ** 10          # the GLOBAL problem: 2^21 columns x 64 levels (NY rows of NX columns)
DT, NU = 0.02, 1.0
STAGES = 4                                   # ARS343
METRIC, UNIT = "imex_step_dof_stage_updates_per_s", "DOF-stage/s"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""

    def __init__(self, device=0):
        self.rows, self.stop_evt, self.proc, self.device = [], threading.Event(), None, device

    def start(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.device)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
            return
        def pump():
            for line in self.proc.stdout:
                if self.stop_evt.is_set():
                    break
                self.rows.append((time.time(), [x.strip() for x in line.split(",")]))
        self.t = threading.Thread(target=pump, daemon=True)
        self.t.start()

    def count_since(self, t0):
        return sum(1 for ts, r in list(self.rows) if ts >= t0 and r and r[0].isdigit())

    def stop(self, t0=0.0):
        """Summarise the samples that arrived after wall-clock t0 (the start of the timed region)."""
        self.stop_evt.set()
        if self.proc:
            self.proc.terminate()
        rows = [r for ts, r in self.rows if ts >= t0]
        sm = sorted(int(r[0]) for r in rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in rows if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in rows if len(r) >= 6 for i in range(4) if r[2 + i].lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
def host_mem_gib():
    try:
        with open("/proc/meminfo") as f:
            for line in f:
                if line.startswith("MemAvailable:"):
                    return int(line.split()[1]) / 2 ** 20
    except Exception:
        pass
    return 0.0


def bind_to_gpu_numa_node(local):
    """Best effort: run this rank's host threads on the CPUs of the NUMA node its GPU hangs off, so that the pinned staging
    buffers of the end-to-end measurement (first touched by this process) are node-local and N ranks do not all stream
    through one memory controller.  Returns what was found (reported in the JSON line); never fails."""
    info = {"gpu_numa_node": None, "bound_cpus": None}
    try:
        import torch
        try:
            pr = torch.cuda.get_device_properties(local)
            bdf = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        except AttributeError:
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = vis.split(",")[local] if vis else str(local)
            out = subprocess.run(["nvidia-smi", "--query-gpu=pci.bus_id", "--format=csv,noheader", "-i", idx],
                                 capture_output=True, text=True, timeout=20).stdout.strip().lower()
            dom, rest = out.split(":", 1)
            bdf = f"{dom[-4:]}:{rest}"
        with open(f"/sys/bus/pci/devices/{bdf}/numa_node") as f:
            node = int(f.read().strip())
        info["gpu_numa_node"] = node
        if node < 0:
            return info
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            spec = f.read().strip()
        cpus = set()
        for part in spec.split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = cpus & set(os.sched_getaffinity(0))
        if allowed and allowed != set(os.sched_getaffinity(0)):
            os.sched_setaffinity(0, allowed)
            info["bound_cpus"] = len(allowed)
    except Exception as e:  # no sysfs entry, no such property, container without the right to set affinity ...
        info["error"] = repr(e)[:120]
    return info


def cpu_reference(steps, warmup, sample_cols_log2=18, threads=None):
    """The reference's pass structure (oracle/cts_oracle.c, OpenMP over all host cores) on 2^sample_cols_log2 columns x 64
    levels of the same workload (21 = the full BASELINE configuration).  `threads = 1` gives the single-core figure: the
    reference on `Base.Array` is a single-threaded broadcast loop, the all-cores number is the generous one."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import numpy as np
    import cts_oracle as CO
    import cts_ref as R
    lib = CO.load()
    nx = NX
    ny = max(1, (1 << sample_cols_log2) // nx)

    # all the host threads this process may use (torchrun exports OMP_NUM_THREADS=1 to its workers: not a property of the box)
    all_threads = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    lib.oracle_set_num_threads(int(threads) if threads is not None else all_threads)
    cores = lib.oracle_num_threads()
    ncol = (ny + 2) * nx
    K = CO.fill_hash(ncol, 2, 0, 0.25, 2.0)
    z = (np.arange(NLEV) + 1) / (NLEV + 1)
    S = 5.0 * np.exp(-((z - 0.3) ** 2) / 0.01)
    co = CO.ColumnARK(NLEV, nx, ny, K, S, NU, 1, tab=R.tableau("ARS343"), max_iters=1, update_j="NewTimeStep")
    u = CO.fill_hash(ncol * NLEV, 3)
    co.dss(u)
    t = 0.0
    for _ in range(warmup):
        co.step(u, t, DT)
        t += DT
    t0 = time.perf_counter()
    for _ in range(steps):
        co.step(u, t, DT)
        t += DT
    el = time.perf_counter() - t0
    n_owned = ny * nx * NLEV
    assert np.isfinite(u).all()
    if threads is not None:
        lib.oracle_set_num_threads(all_threads)
    return {"value": n_owned * STAGES * steps / el, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{ny * nx} columns x {NLEV} levels ({n_owned} DOF), {steps} steps after {warmup} warm-up, "
                      f"oracle/cts_oracle.c (reference pass structure, OpenMP, -O3 -march=native -ffp-contract=off)",
            "ms_per_step": 1e3 * el / steps}


def run_reference(args):
    """The reference's own pass structure on the host cores, on the SAME configuration as the GPU arm (all 2^21 columns,
    same --steps/--warmup) when the box has the memory for it (17 arrays of 1 GiB); otherwise a bounded sample whose size
    is stated in `config`.  ~0.8 s per step on 16 cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    full = host_mem_gib() >= 40 and not args.reference_sample
    log2cols = 21 if full else 18
    steps = max(1, min(args.steps, 30 if full else 60))
    warm = max(1, min(args.warmup, 5))
    cb = cpu_reference(steps, warm, sample_cols_log2=log2cols)
    cfg = workload_config(args.gpus, stored_w=True)
    cfg["reference_arm"] = ("full configuration: 2^21 columns x 64 levels" if full else
                            f"bounded sample: 2^{log2cols} columns x 64 levels of the same workload (host memory)")
    cfg["reference_columns"] = 1 << log2cols
    one = cpu_reference(1, 1, sample_cols_log2=16, threads=1)
    line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": steps, "warmup": warm, "ms_per_step": cb["ms_per_step"], "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": cfg,
            "single_core": {"value": one["value"], "cores": one["cores"], "sample": one["sample"],
                            "note": "the reference's Base.Array path is single-threaded; `value` above is the generous all-cores OpenMP figure"},
            "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def workload_config(n_gpus, stored_w, ny_local=None):
    ny_local = NY // n_gpus if ny_local is None else ny_local
    dof_gpu = NX * ny_local * NLEV
    return {"workload": "ARS343 IMEX + NewtonsMethod(max_iters=1, update_j=NewTimeStep), column vertical diffusion "
                        f"{NX * NY} columns x {NLEV} levels ({NX * NY * NLEV} DOF{', = 2^27' if NY == 1024 else ''}) in total, Float64, batched "
                        f"tridiagonal W ({'stored dl/d/du' if stored_w else 'matrix-free'}), explicit source + 5-point horizontal "
                        f"coupling, dss! = ghost-row halo exchange; split into {n_gpus} slab(s) of {ny_local} rows x {NX} columns",
            "baseline_config": "BASELINE.json configs[2] (1 GPU) / configs[4] (the same problem sharded over N GPUs)",
            "columns_total": NX * NY, "columns_per_gpu": NX * ny_local, "levels": NLEV, "dof_per_gpu": dof_gpu,
            "stages_per_step": STAGES, "dt": DT, "global_dof": NX * NY * NLEV, "parallelism": f"column slabs x{n_gpus}",
            "cache_policy": (f"every operand array is {dof_gpu * 8 / 2 ** 20:.0f} MiB per GPU; one step streams "
                             f"{dof_gpu * 328 / 2 ** 20:.0f} MiB through HBM "
                             + (">> 126 MB L2, so no L2 flush is needed between steps" if dof_gpu * 328 > 8 * 126e6
                                else "(NOT large against the 126 MB L2: numbers are cache-assisted)"))}


# ------------------------------------------------------------------------------------------------
def build_problem(cts, ctx, part, matrix_free, nx=None, nlev=None, alg=None, nonlinear=False, dt=None, foreign=False):
    """The column problem on this rank's slab of `part` (hashed inputs keyed by GLOBAL indices, so every sharding sees
    the same field)."""
    import numpy as np
    nx = NX if nx is None else nx
    nlev = NLEV if nlev is None else nlev
    nyl = part.ny_local
    ncol = (nyl + 2) * nx
    K = cts.B200Vector.zeros(ncol, np.float64, ctx)
    u = cts.B200Vector.zeros(ncol * nlev, np.float64, ctx)
    rows = part.local_rows(1)
    # contiguous runs of global rows (ghost rows wrap around): fill per run
    r = 0
    while r < len(rows):
        e = r
        while e + 1 < len(rows) and rows[e + 1] == rows[e] + 1:
            e += 1
        g0, cnt = int(rows[r]), e - r + 1
        kv = cts.B200Vector(K.tensor[r * nx:(r + cnt) * nx], ctx)
        kv.fill_hash(2, g0 * nx, 0.25, 2.0)
        uv = cts.B200Vector(u.tensor[r * nx * nlev:(r + cnt) * nx * nlev], ctx)
        uv.fill_hash(3, g0 * nx * nlev)
        r = e + 1
    z = (np.arange(nlev) + 1) / (nlev + 1)
    S = cts.B200Vector.from_host(5.0 * np.exp(-((z - 0.3) ** 2) / 0.01), ctx)
    op = cts.ColumnOperator(nlev, nx, nyl, K, S_lev=S, nu=NU, halo=1, nonlinear=nonlinear, ctx=ctx)
    if matrix_free:
        op.set_matrix_free(True)
    if alg is None:
        alg = cts.IMEXAlgorithm(cts.ARS343(), cts.NewtonsMethod(max_iters=1, update_j=cts.UpdateEvery(cts.NewTimeStep)))
    prob = cts.ODEProblem(op.foreign_ode_function() if foreign else op.ode_function(), u, (0.0, 1.0e9), None)
    integ = cts.init(prob, alg, dt=DT if dt is None else dt, save=False)
    return op, integ, u


def sharded_parity(cts, ctx, ctx_single, dist, rank, world, local):
    """Driver-visible parity of the N-rank path: a miniature of the benchmark problem (same code path: fused stages,
    overlapped peer-memory halos, CUDA-graph steps) and a sharded Newton-Krylov (SSP333 + JFNK/GMRES) case are stepped on
    the N slabs and, on rank 0, as ONE slab on a communicator-free context; the owned rows must agree bit for bit."""
    import torch
    out = {"vs": "the same global problem stepped as one slab on rank 0 (which tests/ compare with the oracle)", "cases": {}}
    nx, nlev, nyl = 64, 64, 32
    cases = (("ars343_fused", None, False, 3, 0.02),
             ("ssp333_jfnk", cts.IMEXAlgorithm(cts.SSP333(), cts.NewtonsMethod(max_iters=2, krylov_method=cts.KrylovMethod(
                 jacobian_free_jvp=cts.ForwardDiffJVP(), forcing_term=cts.ConstantForcing(1e-3), itmax=30))), True, 2, 0.002))
    ok_all = True
    for name, alg, nonlinear, nsteps, dt in cases:
        nlev_c = 16 if nonlinear else nlev
        part = cts.SlabPartition(nyl * world, world, rank)
        lay = part.reduction_layout(nx, nlev_c)
        ctx.set_reduction_layout(*lay)
        op, integ, u = build_problem(cts, ctx, part, False, nx=nx, nlev=nlev_c, alg=alg, nonlinear=nonlinear, dt=dt)
        op.dss(u, None, 0.0)
        for _ in range(nsteps):
            cts.step_(integ)
        nrm = u.norm()
        own = u.tensor[lay[0]:lay[0] + lay[1]].contiguous()
        parts = [torch.empty_like(own) for _ in range(world)] if rank == 0 else None
        torch.cuda.synchronize()
        dist.gather(own, parts, dst=0)
        torch.cuda.synchronize()
        ctx.set_reduction_window(0, 0)
        res = None
        if rank == 0:
            whole = cts.SlabPartition(nyl * world, 1, 0)
            lay1 = whole.reduction_layout(nx, nlev_c)
            ctx_single.set_reduction_layout(*lay1)
            op1, integ1, u1 = build_problem(cts, ctx_single, whole, False, nx=nx, nlev=nlev_c, alg=alg, nonlinear=nonlinear, dt=dt)
            op1.dss(u1, None, 0.0)
            for _ in range(nsteps):
                cts.step_(integ1)
            nrm1 = u1.norm()
            ctx_single.set_reduction_window(0, 0)
            want = u1.tensor[lay1[0]:lay1[0] + lay1[1]]
            got = torch.cat(parts)
            exact = bool(torch.equal(got, want)) and bool(torch.isfinite(got).all())
            err = float((got - want).abs().max() / want.abs().max())
            res = {"bit_exact": exact, "max_rel_err": err, "norm_equal": nrm == nrm1, "steps": nsteps,
                   "global_dof": int(want.numel()), "graph_launches": integ.cache.graph_stats()["graph_launches"],
                   "krylov_iters": integ.cache.stats()["krylov_iters"]}
            ok_all = ok_all and exact and nrm == nrm1
            del op1, integ1, u1
        out["cases"][name] = res
        del op, integ, u
        dist.barrier()   # rank 0 was busy with the single-slab run: do not let the other ranks run ahead into the next case
    out["ok"] = ok_all
    out["halo_path"] = ctx.halo_path()
    return out


def stage_bytes(stored_w):
    """Algorithmic bytes per DOF of the fused column stage launches of one ARS343 step:
    stage i reads base + nnz(a_exp[i,:]) + nnz(a_imp[i,:]) arrays (+ dl,d,du when W is stored) and writes U and T_imp[i].
    With update_j = NewTimeStep the first implicit stage WRITES dl,d,du (it folds Wfact in) instead of reading them:
    same byte count."""
    w = 8
    reads = [1 + 1 + 0, 1 + 2 + 1, 1 + 3 + 2]   # stages 2,3,4
    return [(r + 2 + (3 if stored_w else 0)) * w for r in reads]


def time_steps(cts, stream, integ, steps, barrier):
    import torch
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    for _ in range(steps):
        cts.step_(integ)
    ev1.record(stream)
    barrier()
    return ev0.elapsed_time(ev1)


def run_ours(args):
    import numpy as np
    import torch
    import cts_b200 as cts

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the B200 path has no CPU fallback")
    if NY % world or NY // world < 32:
        raise SystemExit(f"bench.py: {NY} rows cannot be split into {world} slabs of >= 32 rows")
    torch.cuda.set_device(local)
    all_cpus = set(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else None
    numa = bind_to_gpu_numa_node(local) if world > 1 else {"gpu_numa_node": None, "bound_cpus": None}
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    stream = torch.cuda.Stream()
    with torch.cuda.stream(stream):
        ctx = cts.Context(local)           # enqueues on `stream`
        if world > 1:
            cts.attach_communicator(ctx, dist)
            ctx.enable_peer_halo(NX * NLEV * 8)   # collective: CUDA IPC mailboxes of the ring neighbours
        # STRONG scaling: the 2^21-column problem is split into `world` slabs (weak scaling: extras.weak_scaling)
        part = cts.SlabPartition(NY, world, rank)
        stored_w = not args.matrix_free
        op, integ, u = build_problem(cts, ctx, part, args.matrix_free)
        if args.no_graph:
            integ.cache.set_graph(False)
        op.dss(u, None, 0.0)
        n_owned = op.ndof_owned
        state_bytes = int(u.tensor.numel() * 8)

        def barrier():
            torch.cuda.synchronize()
            if dist is not None:
                dist.barrier()
            torch.cuda.synchronize()

        def max_over_ranks(x):
            if dist is None:
                return x
            t = torch.tensor([x], dtype=torch.float64, device=f"cuda:{local}")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())

        sampler = ClockSampler(local)
        if rank == 0:
            sampler.start()  # started before the warm-up so that nvidia-smi is already streaming when the timing starts
        for _ in range(args.warmup):
            cts.step_(integ)
        # ---- the timed region: K steps, barrier + synchronize on both sides, CUDA events on the launching stream
        barrier()
        t_wall0 = time.time()
        l0, bytes0 = ctx.launch_count(), ctx.bytes_count()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record(stream)
        for _ in range(args.steps):
            cts.step_(integ)
        ev1.record(stream)
        barrier()
        ms_noprof = max_over_ranks(ev0.elapsed_time(ev1))
        launches = ctx.launch_count() - l0
        counted_bytes_per_step = (ctx.bytes_count() - bytes0) / args.steps
        # ---- the same K steps again with CUDA events around every kernel family (external event nodes when the step is a
        # graph): per-kernel table and the roofline's launch durations
        ctx.prof_read()
        ctx.prof_enable(True)
        ms = max_over_ranks(time_steps(cts, stream, integ, args.steps, barrier))
        prof = ctx.prof_read()
        ctx.prof_enable(False)
        stats = integ.cache.stats()
        graph_stats = integ.cache.graph_stats()
        # A short timed region (K steps of ~10 ms) can end before nvidia-smi's 100 ms sampler has ticked: keep the
        # same load running (untimed, all ranks in lockstep) until at least 3 samples have been taken under load.
        trailing = 0
        while True:
            need = torch.tensor([1.0 if (rank == 0 and sampler.proc and sampler.count_since(t_wall0) < 3 and trailing < 4000) else 0.0],
                                device=f"cuda:{local}")
            if dist is not None:
                dist.broadcast(need, 0)
            if need.item() == 0.0:
                break
            for _ in range(20):
                cts.step_(integ)
            torch.cuda.synchronize()
            trailing += 20
        clocks = sampler.stop(t_wall0) if rank == 0 else None
        if clocks is not None:
            clocks["window"] = f"timed region + instrumented pass + {trailing} trailing load steps (untimed)"
        finite = bool(torch.isfinite(u.tensor).all())

        # ---- end-to-end through the public API with HOST buffers: H2D of the state, step!, D2H of the result ----
        e2e_steps = max(1, min(args.steps, 5))
        host = torch.empty(u.tensor.numel(), dtype=torch.float64).pin_memory()
        host.copy_(u.tensor)
        barrier()
        ev0.record(stream)
        nbytes = u.tensor.numel() * 8
        for _ in range(e2e_steps):  # every call below is a C-ABI entry point of libcts_sm100a.so on a HOST buffer
            ctx.check(ctx.lib.cts_memcpy_h2d(ctx.handle, u.ptr, host.data_ptr(), nbytes), "h2d")
            cts.step_(integ)                                                    # cts_imex_step
            ctx.check(ctx.lib.cts_memcpy_d2h_sync(ctx.handle, host.data_ptr(), u.ptr, nbytes), "d2h")
        ev1.record(stream)
        barrier()
        ms_e2e = max_over_ranks(ev0.elapsed_time(ev1))

        # ---- the same end-to-end work, software-pipelined: independent problem instances (the upload replaces the whole
        # state every step anyway) alternate between two device buffers, so the H2D of step k+1 and the D2H of step k-1
        # overlap the kernels of step k (three streams, full-duplex PCIe).  Every step still uploads its input from
        # pinned host memory and downloads its result; all copies and the step are C-ABI calls.
        ms_pipe = None
        try:
            s_in, s_out = torch.cuda.Stream(device=local), torch.cuda.Stream(device=local)
            ctx_in, ctx_out = cts.Context(local, s_in.cuda_stream), cts.Context(local, s_out.cuda_stream)
            host_in, host_out = host, torch.empty_like(host).pin_memory()
            bufs = [u, cts.B200Vector.zeros(u.tensor.numel(), np.float64, ctx)]
            ev_in = [torch.cuda.Event() for _ in range(2)]
            ev_done = [torch.cuda.Event() for _ in range(2)]
            ev_out = [torch.cuda.Event() for _ in range(2)]
            pipe_steps = 2 * e2e_steps

            def upload(k):
                b = k % 2
                s_in.wait_event(ev_out[b])          # the previous occupant of this buffer has been downloaded
                ctx_in.check(ctx_in.lib.cts_memcpy_h2d(ctx_in.handle, bufs[b].ptr, host_in.data_ptr(), nbytes), "h2d")
                ev_in[b].record(s_in)

            barrier()
            for b in range(2):
                ev_out[b].record(s_out)
            t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0.record(stream)
            s_in.wait_event(t0)
            upload(0)
            for k in range(pipe_steps):
                b = k % 2
                if k + 1 < pipe_steps:
                    upload(k + 1)
                stream.wait_event(ev_in[b])
                integ.u = bufs[b]
                cts.step_(integ)                                                # cts_imex_step on bufs[b]
                ev_done[b].record(stream)
                s_out.wait_event(ev_done[b])
                ctx_out.check(ctx_out.lib.cts_memcpy_d2h(ctx_out.handle, host_out.data_ptr(), bufs[b].ptr, nbytes), "d2h")
                ev_out[b].record(s_out)
            stream.wait_event(ev_out[(pipe_steps - 1) % 2])
            t1.record(stream)
            barrier()
            ms_pipe = max_over_ranks(t0.elapsed_time(t1))
            integ.u = u
            finite = finite and bool(torch.isfinite(host_out).all())
            del bufs, host_out
        except Exception as e:  # the pipelined variant is an extra; the sequential number above is the contract's e2e
            sys.stderr.write(f"pipelined e2e skipped: {e!r}\n")
            integ.u = u

        # ---- extras ------------------------------------------------------------------------------------------
        extras = {}
        halo_path = ctx.halo_path()
        if world > 1:
            # (a) the same slab stepped WITHOUT a communicator (periodic local ghost copies): what one GPU does on its
            #     share of the problem when nothing has to be exchanged -> exposed communication = difference
            ctx1 = cts.Context(local)
            op1, integ1, u1 = build_problem(cts, ctx1, cts.SlabPartition(NY // world, 1, 0), args.matrix_free)
            if args.no_graph:
                integ1.cache.set_graph(False)
            op1.dss(u1, None, 0.0)
            for _ in range(args.warmup):
                cts.step_(integ1)
            ms_local = max_over_ranks(time_steps(cts, stream, integ1, args.steps, barrier))
            del op1, integ1, u1
            extras["local_slab"] = {"ms_per_step": ms_local / args.steps,
                                    "note": "this rank's slab with periodic local ghost copies instead of the neighbour exchange (max over ranks)"}
            extras["exposed_comm_ms_per_step"] = (ms_noprof - ms_local) / args.steps
            # (b) driver-visible parity of the sharded path
            extras["sharded_parity"] = sharded_parity(cts, ctx, ctx1, dist, rank, world, local)
            # (c) weak scaling: 2^27 DOF per GPU (the global grid grows with N), as reported in round 1
            if not args.no_weak:
                del integ, op, u
                torch.cuda.empty_cache()
                opw, integw, uw = build_problem(cts, ctx, cts.SlabPartition(NY * world, world, rank), args.matrix_free)
                opw.dss(uw, None, 0.0)
                wsteps = max(3, min(args.steps, 10))
                for _ in range(3):
                    cts.step_(integw)
                ms_w = max_over_ranks(time_steps(cts, stream, integw, wsteps, barrier))
                extras["weak_scaling"] = {"value": opw.ndof_owned * world * STAGES * wsteps / (ms_w * 1e-3), "unit": UNIT,
                                          "ms_per_step": ms_w / wsteps, "steps": wsteps, "dof_per_gpu": opw.ndof_owned,
                                          "global_dof": opw.ndof_owned * world}
                del opw, integw, uw
        elif not args.no_extras:
            del integ
            extras = run_extras(cts, ctx, stream, op, u)

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return
    peak, peak_src = peaks()
    ndof_local = (NY // world + 2) * NX * NLEV
    # `value`: the K timed steps without the per-kernel event brackets; the instrumented pass of the same K steps (run
    # immediately before, same state) provides the per-kernel table and the roofline's launch durations
    value = n_owned * world * STAGES * args.steps / (ms_noprof * 1e-3)
    # roofline of the dominant kernel (fused column stage): algorithmic bytes / measured launch time
    fs_ms, fs_n = prof.get("fused_stage", (0.0, 0))
    per_step = sum(stage_bytes(stored_w)) * ndof_local
    roof = None
    if fs_n:
        achieved = per_step * (fs_n / 3.0) / (fs_ms * 1e-3) / 1e9
        roof = {"bound": "hbm", "kernel": "column_fused_stage_reg_kernel (K1+K2+T_imp+K3+K4+K5+K6 of one implicit stage; the first one also writes W)",
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "peak_source": peak_src,
                "nominal_peak": 8000.0, "frac_of_nominal": achieved / 8000.0,
                "traffic": ncu_traffic(stored_w, world),
                "algorithmic_bytes_per_launch": per_step / 3.0,
                "launches": fs_n, "avg_launch_ms": fs_ms / fs_n,
                "algorithmic_bytes_per_dof_per_launch": stage_bytes(stored_w),
                "share_of_step": fs_ms / ms}
    kernels = {k: {"ms_per_step": v[0] / args.steps, "launches_per_step": v[1] / args.steps} for k, v in prof.items()}
    step_bytes = step_bytes_per_dof(stored_w) * ndof_local
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_noprof / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": workload_config(world, stored_w),
            "roofline": roof,
            "step_hbm": {"algorithmic_bytes_per_dof_step": step_bytes_per_dof(stored_w),
                         "counted_bytes_per_dof_step": counted_bytes_per_step / ndof_local,
                         "achieved_GBs_per_gpu": step_bytes / (ms_noprof / args.steps * 1e-3) / 1e9,
                         "frac_of_peak": step_bytes / (ms_noprof / args.steps * 1e-3) / 1e9 / peak},
            "kernels": kernels,
            "instrumented_pass": {"ms_per_step": ms / args.steps, "steps": args.steps,
                                  "note": "the same K steps with CUDA events around every kernel family (source of `kernels` and `roofline`)"},
            "halo": {"path": halo_path, "exchanges_per_step": 4,
                     "overlap": "both exchanges of every stage and of the step end run on a second stream beside the interior rows"},
            "cuda_graph": graph_stats,
            "e2e": {"value": n_owned * world * STAGES * e2e_steps / (ms_e2e * 1e-3), "unit": UNIT,
                    "h2d_bytes_per_step": state_bytes * world, "d2h_bytes_per_step": state_bytes * world,
                    "ms_per_step": ms_e2e / e2e_steps, "steps": e2e_steps,
                    "pipelined": None if ms_pipe is None else {
                        "value": n_owned * world * STAGES * pipe_steps / (ms_pipe * 1e-3), "ms_per_step": ms_pipe / pipe_steps,
                        "steps": pipe_steps, "note": "same per-step H2D + cts_imex_step + D2H, independent instances double-buffered "
                        "on three streams so the copies overlap the kernels"},
                    "api": "C ABI: cts_memcpy_h2d (pinned host state) + cts_imex_step (via cts_b200.step_) + cts_memcpy_d2h_sync, every step"},
            "gpu_launches": launches, "clocks": clocks, "solver_stats": stats, "finite": finite, "host_numa": numa}
    if all_cpus is not None:
        try:
            os.sched_setaffinity(0, all_cpus)   # the CPU baseline below uses every host core again
        except Exception:
            pass
    if not args.no_cpu_baseline:   # rank 0, every N (the CPU arm is the same fixed global problem)
        cb = cpu_reference(2, 1)
        line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
        one = cpu_reference(1, 1, sample_cols_log2=16, threads=1)   # the faithful single-threaded figure, smaller sample
        line["cpu_baseline"]["single_core"] = {"value": one["value"], "cores": one["cores"], "sample": one["sample"]}
    else:
        line["cpu_baseline"] = None
    if extras:
        line["extras"] = extras
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


def ncu_traffic(stored_w, world):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the fused stage (mean of the step's 3 launches), read from
    the committed summary of the `ncu --set full` capture of this command (profiles/summarize.py traffic -> JSON)."""
    p = os.path.join(ROOT, "profiles", "fused_stage_traffic.json")
    try:
        with open(p) as f:
            t = json.load(f)
        key = f"{'stored' if stored_w else 'matrix_free'}_n{world}"
        return t.get(key, {}).get("traffic_bytes_per_launch")
    except Exception:
        return None


def step_bytes_per_dof(stored_w):
    """Whole fused ARS343 step: 3 fused stages + 3 T_exp! (1 read, 1 write) + the final update with the 4th T_exp! folded
    in (reads u, U_4, 2 T_exp, 3 T_imp; writes u: 8w).  Stored W: 328 B/DOF (Wfact is folded into the first implicit
    stage, which writes W instead of reading it; SURVEY §8d's "fused minimum" of 344 B/DOF still stores and re-reads
    T_exp[4]); matrix-free: 256 B/DOF."""
    return sum(stage_bytes(stored_w)) + 8 * 8 + 3 * 2 * 8


def timed(stream, fn, reps, warm=2):
    import torch
    for _ in range(warm):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    stream.synchronize()
    a.record(stream)
    for _ in range(reps):
        fn()
    b.record(stream)
    stream.synchronize()
    return a.elapsed_time(b) / reps


def run_extras(cts, ctx, stream, op, u):
    """Other single-GPU configs / kernels, each as achieved algorithmic GB/s vs the measured HBM peak."""
    import ctypes as C
    import numpy as np
    import torch
    from cts_b200 import _lib as L
    peak, _ = peaks()
    out = {}
    n = 1 << 27
    # --- the drop-in (generic-hook) path on the default workload: what a model gets whose T_imp!/Wfact/ldiv! the engine
    # cannot recognise.  (a) the library's hooks with operator fusion switched off = the reference's pass structure, pass by
    # pass (K1, K2, T_imp!, K3, K4, K5, K6 per implicit stage; 93w = 744 B/DOF/step); (b) the same hooks as
    # opaque host closures (Python trampolines -> C ABI): same kernels, plus the host round trip of every callback.
    part = cts.SlabPartition(NY, 1, 0)
    for key, foreign in (("C3_generic_hooks", False), ("C3_foreign_hooks", True)):
        op_g, integ_g, u_g = build_problem(cts, ctx, part, False, foreign=foreign)
        if not foreign:
            integ_g.cache.set_fusion(False)
        op_g.dss(u_g, None, 0.0)
        l0 = ctx.launch_count()
        ms = timed(stream, lambda: cts.step_(integ_g), 3, warm=2)
        launches = (ctx.launch_count() - l0) / 5
        model = 744
        out[key] = {"ms_per_step": ms, "dof_stage_per_s": op_g.ndof_owned * STAGES / (ms * 1e-3), "bytes_per_dof_step": model,
                    "GBs": model * op_g.ndof / (ms * 1e-3) / 1e9, "frac": model * op_g.ndof / (ms * 1e-3) / 1e9 / peak,
                    "launches_per_step": launches, "fused_stage_launches": integ_g.cache.stats()["fused_stage_launches"],
                    "passes_per_implicit_stage": 7}
        del op_g, integ_g, u_g
        torch.cuda.empty_cache()
    # --- K1 stand-alone: ARS343 final update shape (7 reads + 1 write) and stage-2 shape (2 reads + 1 write)
    vs = [cts.B200Vector.zeros(n, np.float64, ctx).fill_hash(10 + k) for k in range(7)]
    dst = cts.B200Vector.zeros(n, np.float64, ctx)
    for nt in (1, 3, 6):
        g1 = [(0.01 * (k + 1), vs[1 + k]) for k in range(nt)]
        ms = timed(stream, lambda: cts.axpy_n(dst, vs[0], g1[:nt // 2 + nt % 2], g1[nt // 2 + nt % 2:]), 10)
        gb = (2 + nt) * 8 * n / (ms * 1e-3) / 1e9
        out[f"K1_axpy_n_{nt}terms"] = {"ms": ms, "GBs": gb, "frac": gb / peak, "bytes_per_dof": (2 + nt) * 8}
    # --- K4 stand-alone: batched tridiagonal solve (5w)
    W = L.Tridiag(NLEV, n // NLEV, vs[1].ptr, vs[2].ptr, vs[3].ptr)
    vs[2].tensor.add_(-4.0)  # diagonally dominant
    ms = timed(stream, lambda: ctx.check(ctx.lib.cts_tridiag_solve_batched(ctx.handle, 1, C.byref(W), vs[4].ptr, dst.ptr)), 10)
    gb = 5 * 8 * n / (ms * 1e-3) / 1e9
    out["K4_tridiag_solve_batched"] = {"ms": ms, "GBs": gb, "frac": gb / peak, "bytes_per_dof": 40}
    # --- reductions
    ms = timed(stream, lambda: vs[0].norm(), 10)
    out["R1_nrm2"] = {"ms": ms, "GBs": 8 * n / (ms * 1e-3) / 1e9, "frac": 8 * n / (ms * 1e-3) / 1e9 / peak}
    ms = timed(stream, lambda: vs[0].dot(vs[1]), 10)
    out["R3_dot"] = {"ms": ms, "GBs": 16 * n / (ms * 1e-3) / 1e9, "frac": 16 * n / (ms * 1e-3) / 1e9 / peak}
    # --- config C2: LSRK54CarpenterKennedy on 2^27-element upwind advection, fused and un-fused
    del vs[3:]
    adv = cts.AdvectionOperator(n, ctx=ctx)
    ua = cts.B200Vector.zeros(n, np.float64, ctx).fill_hash(1)
    integ = cts.init(cts.ODEProblem(adv.ode_function(), ua, (0.0, 1e9), None), cts.LSRK54CarpenterKennedy(), dt=0.5 / n, save=False)
    for fused in (True, False):
        integ.cache.set_fusion(fused)
        ms = timed(stream, lambda: cts.step_(integ), 5)
        model = 5 * 4 * 8 if fused else 30 * 8   # bytes/DOF/step: 5 fused stages of 4w (the last one in place) | 5 x 6w
        out[f"C2_lsrk54_advection_{'fused' if fused else 'unfused'}"] = {
            "ms_per_step": ms, "dof_stage_per_s": n * 5 / (ms * 1e-3), "GBs": model * n / (ms * 1e-3) / 1e9,
            "frac": model * n / (ms * 1e-3) / 1e9 / peak, "bytes_per_dof_step": model}
    del integ, adv, ua, vs, dst
    torch.cuda.empty_cache()
    # --- config C4: SSP333 (SSPRK path) + JFNK/GMRES, Newton max_iters = 2, nonlinear column diffusion, f64 and f32
    for dtype, name in ((np.float64, "f64"), (np.float32, "f32")):
        out[f"C4_ssp333_jfnk_{name}"] = run_c4(cts, ctx, stream, dtype)
        # the same config with the operator declared matrix-free: the Krylov tile kernels generate W(x) themselves (bit-identical)
        out[f"C4_ssp333_jfnk_{name}_matrix_free"] = run_c4(cts, ctx, stream, dtype, matrix_free=True)
        torch.cuda.empty_cache()
    out["C3_ars343_column_f32"] = run_c3_f32(cts, ctx, stream)
    torch.cuda.empty_cache()
    for fused in (True, False):
        out[f"F2_rosenbrock_sspknoth_column_{'fused' if fused else 'unfused'}"] = run_rosenbrock(cts, ctx, stream, fused)
        torch.cuda.empty_cache()
    # --- the default workload with W regenerated in the stage kernel from K_col instead of stored (bit-identical results,
    # 3w B/DOF less per implicit stage; `--matrix-free` runs the whole bench line this way)
    part = cts.SlabPartition(NY, 1, 0)
    op_mf, integ_mf, u_mf = build_problem(cts, ctx, part, True)
    op_mf.dss(u_mf, None, 0.0)
    ms = timed(stream, lambda: cts.step_(integ_mf), 5, warm=3)
    model = step_bytes_per_dof(False)
    out["C3_ars343_column_matrix_free"] = {"ms_per_step": ms, "dof_stage_per_s": op_mf.ndof_owned * STAGES / (ms * 1e-3),
                                           "bytes_per_dof_step": model, "GBs": model * op_mf.ndof / (ms * 1e-3) / 1e9,
                                           "frac": model * op_mf.ndof / (ms * 1e-3) / 1e9 / peak}
    del integ_mf, op_mf, u_mf
    torch.cuda.empty_cache()
    return out


def run_c3_f32(cts, ctx, stream):
    """The default workload (C3: ARS343, 2^21 columns x 64 levels, stored W, halo dss!) with a Float32 state: element-wise
    arithmetic in Float64 with Float64 tableau coefficients, rounded on store; the solve is Float32."""
    import numpy as np
    K = cts.B200Vector.zeros((NY + 2) * NX, np.float32, ctx).fill_hash(2, 0, 0.25, 2.0)
    u = cts.B200Vector.zeros((NY + 2) * NX * NLEV, np.float32, ctx).fill_hash(3)
    z = (np.arange(NLEV) + 1) / (NLEV + 1)
    S = cts.B200Vector.from_host((5.0 * np.exp(-((z - 0.3) ** 2) / 0.01)).astype(np.float32), ctx)
    op = cts.ColumnOperator(NLEV, NX, NY, K, S_lev=S, nu=NU, halo=1, ctx=ctx)
    alg = cts.IMEXAlgorithm(cts.ARS343(), cts.NewtonsMethod(max_iters=1, update_j=cts.UpdateEvery(cts.NewTimeStep)))
    integ = cts.init(cts.ODEProblem(op.ode_function(), u, (0.0, 1.0e9), None), alg, dt=DT, save=False)
    op.dss(u, None, 0.0)
    ms = timed(stream, lambda: cts.step_(integ), 5, warm=3)
    import torch
    n = NX * NY * NLEV
    model = step_bytes_per_dof(True) // 2
    res = {"ms_per_step": ms, "dof_stage_per_s": n * STAGES / (ms * 1e-3), "bytes_per_dof_step": model,
           "GBs": model * op.ndof / (ms * 1e-3) / 1e9, "frac": model * op.ndof / (ms * 1e-3) / 1e9 / peaks()[0],
           "finite": bool(torch.isfinite(u.tensor).all())}
    del integ, op, u, K
    return res


def run_rosenbrock(cts, ctx, stream, fused):
    """RosenbrockAlgorithm(tableau(SSPKnoth())) (rosenbrock.jl:136-239) on the default column workload
    (2^21 columns x 64 levels, Float64, explicit source + horizontal coupling, halo dss!).  `fused`: T_imp! + the `.+=` chain of
    the stage right-hand side + ldiv! run as one tile kernel per stage (cts_column_rosenbrock_stage); otherwise the reference's
    pass structure.  Bytes = those of the kernels that ran (cts_bytes_count)."""
    import numpy as np
    import torch
    part = cts.SlabPartition(NY, 1, 0)
    op, _, u = build_problem(cts, ctx, part, False)
    integ = cts.init(cts.ODEProblem(op.ode_function(), u, (0.0, 1.0e9), None), cts.RosenbrockAlgorithm(cts.tableau(cts.SSPKnoth())),
                     dt=DT, save=False)
    integ.cache.set_fusion(fused)
    op.dss(u, None, 0.0)
    cts.step_(integ)
    steps = 3
    b0, l0 = ctx.bytes_count(), ctx.launch_count()
    ms = timed(stream, lambda: cts.step_(integ), steps, warm=0)
    nbytes, launches = (ctx.bytes_count() - b0) / steps, (ctx.launch_count() - l0) / steps
    peak = peaks()[0]
    gbs = nbytes / (ms * 1e-3) / 1e9
    res = {"ms_per_step": ms, "dof_stage_per_s": op.ndof_owned * 3 / (ms * 1e-3), "stages_per_step": 3,
           "bytes_per_dof_step": nbytes / op.ndof, "GBs": gbs, "frac": gbs / peak, "launches_per_step": launches,
           "fused_stage_launches_per_step": integ.cache.stats()["fused_stage_launches"] / (steps + 1),
           "finite": bool(torch.isfinite(u.tensor).all())}
    del integ, op, u
    return res


def run_c4(cts, ctx, stream, dtype, matrix_free=False):
    """BASELINE configs[3] as written: IMEXAlgorithm(SSP333(), NewtonsMethod(; max_iters = 2, krylov_method = KrylovMethod(;
    jacobian_free_jvp = ForwardDiffJVP(), forcing_term = ConstantForcing(rtol)))) on d/dz(K(1+T^2) dT/dz), 2^27 DOF, the
    frozen-coefficient W (batched tridiagonal) as left preconditioner, GMRES `memory = 20` and `itmax` at their defaults
    (newtons_method.jl:433).  The pass count is data dependent, so the algorithmic bytes are those of the kernels that
    actually ran (`cts_bytes_count`: every launch adds (arrays read + written) x n x w, the per-kernel figures of
    SURVEY §8d); achieved GB/s = those bytes / step time."""
    import numpy as np
    nlev, nx, ny = NLEV, NX, NY
    K = cts.B200Vector.zeros(nx * ny, dtype, ctx).fill_hash(2, 0, 0.25, 2.0)
    # smooth state + 2 % hashed perturbation (white noise would make the frozen-coefficient preconditioner useless)
    u = cts.B200Vector.zeros(nx * ny * nlev, dtype, ctx).fill_hash(3, 0, 0.02, 25.0)
    op = cts.ColumnOperator(nlev, nx, ny, K, nu=0.0, halo=0, nonlinear=True, ctx=ctx).set_matrix_free(matrix_free)
    rtol = 1e-2
    # ForwardDiffStepSize3 shrinks like 1/sqrt(n): at n = 2^27 the Float32 perturbation eps*v_i falls below eps(Float32)*x_i
    # and the JVP vanishes; `step_adjustment` (ForwardDiffJVP's own knob, newtons_method.jl:152-168) restores it.
    adj = 4096.0 if np.dtype(dtype) == np.float32 else 1.0
    km = cts.KrylovMethod(jacobian_free_jvp=cts.ForwardDiffJVP(step_adjustment=adj), forcing_term=cts.ConstantForcing(rtol))
    alg = cts.IMEXAlgorithm(cts.SSP333(), cts.NewtonsMethod(max_iters=2, krylov_method=km))
    integ = cts.init(cts.ODEProblem(op.ode_function(explicit=False), u, (0.0, 1e9), None), alg, dt=2e-3, save=False)
    cts.step_(integ)
    s0 = integ.cache.stats()
    steps = 2
    ctx.prof_read()
    ctx.prof_enable(True)
    b0, l0 = ctx.bytes_count(), ctx.launch_count()
    ms = timed(stream, lambda: cts.step_(integ), steps, warm=0)
    nbytes, launches = (ctx.bytes_count() - b0) / steps, (ctx.launch_count() - l0) / steps
    prof = ctx.prof_read()
    ctx.prof_enable(False)
    s1 = integ.cache.stats()
    kit = (s1["krylov_iters"] - s0["krylov_iters"]) / steps
    nit = (s1["newton_iters"] - s0["newton_iters"]) / steps
    n = nx * ny * nlev
    import torch
    finite = bool(torch.isfinite(u.tensor).all())
    peak = peaks()[0]
    gbs = nbytes / (ms * 1e-3) / 1e9
    res = {"ms_per_step": ms, "dof_stage_per_s": n * 3 / (ms * 1e-3), "krylov_iters_per_step": kit, "newton_iters_per_step": nit,
           "rtol": rtol, "dt": 2e-3, "gmres_memory": km.memory, "gmres_itmax": "default (2n)", "finite": finite,
           "jvp_step_adjustment": adj,
           "algorithmic_bytes_per_step": nbytes, "bytes_per_dof_step": nbytes / n, "GBs": gbs, "frac": gbs / peak,
           "kernels_ms_per_step": {k: round(v[0] / steps, 3) for k, v in prof.items() if v[1]},
           "launches_per_step": launches,
           "note": "bytes = sum over the executed kernels (cts_bytes_count); the Newton-Krylov pass count is data dependent"}
    del integ, op, u, K
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--matrix-free", action="store_true", help="regenerate W = dtγJ - I inside the fused stage kernel")
    ap.add_argument("--no-extras", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="plain stream launches instead of the per-step CUDA graph")
    ap.add_argument("--no-weak", action="store_true", help="N > 1: skip the weak-scaling extra (2^27 DOF per GPU)")
    ap.add_argument("--reference-sample", action="store_true", help="--impl reference: time a 2^18-column sample instead of the full 2^21")
    ap.add_argument("--rows", type=int, default=0,
                    help="global rows of 2048 columns (default 1024 = BASELINE's 2^27 DOF; e.g. 128 = one rank's share at N = 8)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.rows:
        global NY
        NY = int(args.rows)
        args.no_extras = True
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
