"""GPU parity for the round-2 machinery: chunked shard-invariant reductions, the CUDA-graph step, the halo/compute
overlap (fork/join with the communication stream) on one rank, and unaligned states (fallback off the fused stage)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import cts_models as M  # noqa: E402
from ref_problems import R  # noqa: E402


@pytest.fixture(scope="module")
def cts():
    import cts_b200
    return cts_b200


def _column_problem(cts, ctx, nlev, nx, ny, dtype=np.float64, nu=1.0):
    K, T0, S = M.column_inputs(nlev, nx, ny, 0, ny, 1)
    K, T0, S = K.astype(dtype), T0.astype(dtype), S.astype(dtype)
    dev = lambda a: cts.B200Vector.from_host(a, ctx)
    op = cts.ColumnOperator(nlev, nx, ny, dev(K), S_lev=dev(S), nu=nu, halo=1, ctx=ctx)
    u = dev(T0)
    op.dss(u, None, 0.0)
    model = M.ColumnModel(nlev, nx, ny, K, S_lev=S, nu=nu, halo=1)
    uo = T0.copy()
    model.dss(uo, None, 0.0)
    return op, u, model, uo


# ------------------------------------------------------------------ reductions ------------------------------
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("n,chunk", [(1, 0), (7, 0), (8191, 0), (8192, 0), (8193, 0), (100003, 0), (1 << 20, 0), (4096 * 3, 64),
                                     (5000, 4), (1 << 18, 1 << 14)])
def test_chunked_reduction_tree_equals_oracle(cts, dtype, n, chunk):
    # csrc/reduce.cu vs cts_ref.tree_sum: norm / dot / sum(1+|x|) bit for bit, default and explicit chunk sizes
    ctx = cts.Context.default()
    x = (M.fill_hash(n, 11) - 0.5).astype(dtype)
    y = (M.fill_hash(n, 12) - 0.25).astype(dtype)
    dx, dy = cts.B200Vector.from_host(x), cts.B200Vector.from_host(y)
    old = R.REDUCTION_CHUNK
    try:
        if chunk:
            ctx.set_reduction_layout(0, n, 0, n, chunk)
            R.REDUCTION_CHUNK = chunk
        assert dx.norm() == R.norm2(x)
        assert dx.dot(dy) == R.dot(x, y)
        assert dx.sum1pabs() == R.sum1pabs(x)
    finally:
        R.REDUCTION_CHUNK = old
        ctx.set_reduction_window(0, 0)


def test_reduction_is_independent_of_alignment(cts):
    # the accumulation order is defined on element indices, not on 16-byte vectors of the pointer: a view that starts
    # 8 bytes off gives the same bits as an aligned copy
    n = 50001
    x = M.fill_hash(n + 1, 5) - 0.5
    base = cts.B200Vector.from_host(x)
    shifted = cts.B200Vector(base.tensor[1:], base.ctx)
    aligned = cts.B200Vector.from_host(x[1:])
    assert shifted.ptr % 16 == 8 and aligned.ptr % 16 == 0
    assert shifted.norm() == aligned.norm() == R.norm2(x[1:])
    assert shifted.dot(shifted) == aligned.dot(aligned)


def test_layout_rejects_partial_chunks(cts):
    ctx = cts.Context.default()
    with pytest.raises(Exception):
        ctx.set_reduction_layout(0, 100, 50, 1000, 64)   # global offset not a multiple of the chunk
    with pytest.raises(Exception):
        ctx.set_reduction_layout(0, 100, 0, 1000, 64)    # owned range ends inside a chunk that is not the last
    ctx.set_reduction_window(0, 0)


def test_windowed_layout_matches_oracle_window(cts):
    # the owned window of a slab (ghost rows excluded) with the chunk a row divides: same bits as the oracle's window
    ctx = cts.Context.default()
    nx, nlev, ny = 4, 16, 6
    part = cts.SlabPartition(ny, 1, 0)
    lay = part.reduction_layout(nx, nlev)
    assert lay == (64, 384, 0, 384, 64)
    x = M.fill_hash((ny + 2) * nx * nlev, 3) - 0.5
    dx = cts.B200Vector.from_host(x)
    old = R.REDUCTION_CHUNK
    try:
        ctx.set_reduction_layout(*lay)
        R.REDUCTION_WINDOW, R.REDUCTION_CHUNK = (lay[0], lay[1]), lay[4]
        assert dx.norm() == R.norm2(x)
        assert dx.sum1pabs() == R.sum1pabs(x)
    finally:
        R.REDUCTION_WINDOW, R.REDUCTION_CHUNK = None, old
        ctx.set_reduction_window(0, 0)


# ------------------------------------------------------------------ CUDA-graph step + overlap ----------------
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("ny", [8, 40])
def test_graph_step_on_private_stream_is_bit_exact(cts, dtype, ny):
    # a capturable (non-default) stream: every step is captured, applied with cudaGraphExecUpdate and launched as one graph;
    # ny = 40 also takes the fork/join overlap of dss! with the interior rows (T_exp! and the final update)
    import torch
    stream = torch.cuda.Stream()
    with torch.cuda.stream(stream):
        ctx = cts.Context(torch.cuda.current_device())
        nlev, nx = 16, 8
        op, u, model, uo = _column_problem(cts, ctx, nlev, nx, ny, dtype)
        alg = cts.IMEXAlgorithm(cts.ARS343(), cts.NewtonsMethod(max_iters=1, update_j=cts.UpdateEvery(cts.NewTimeStep)))
        integ = cts.init(cts.ODEProblem(op.ode_function(), u, (0.0, 1e9), None), alg, dt=0.01, save=False)
        _, step = R.make_stepper(model.ode_function(), R.imex_algorithm("ARS343", R.NewtonsMethod(
            max_iters=1, update_j=R.UpdateEvery("NewTimeStep"))), uo)
        t = 0.0
        try:
            for k, dt in enumerate((0.01, 0.01, 0.004, 0.01)):   # a changed dt changes kernel arguments, not the topology
                cts.set_dt_(integ, dt)
                cts.step_(integ)
                step(uo, None, t, dt)
                t += dt
                np.testing.assert_array_equal(u.to_host(), uo)
            gs = integ.cache.graph_stats()
            assert gs["graph_launches"] == 4 and gs["instantiations"] == 1, gs
            # same results without graphs
            integ.cache.set_graph(False)
            cts.step_(integ)
            step(uo, None, t, 0.01)
            np.testing.assert_array_equal(u.to_host(), uo)
            assert integ.cache.graph_stats()["graph_launches"] == 4
        finally:   # library objects that enqueue on `stream` must go before the stream does
            ctx.sync()
            del integ, op, u


def test_overlap_on_default_stream_is_bit_exact(cts):
    # legacy default stream: not capturable -> plain launches, but the overlapped dss!/T_exp!/final-update path still runs
    ctx = cts.Context.default()
    op, u, model, uo = _column_problem(cts, ctx, 16, 8, 40)
    alg = cts.IMEXAlgorithm(cts.ARS343(), cts.NewtonsMethod(max_iters=1))
    integ = cts.init(cts.ODEProblem(op.ode_function(), u, (0.0, 1e9), None), alg, dt=0.01, save=False)
    _, step = R.make_stepper(model.ode_function(), R.imex_algorithm("ARS343", R.NewtonsMethod(max_iters=1)), uo)
    for k in range(3):
        cts.step_(integ)
        step(uo, None, 0.01 * k, 0.01)
    np.testing.assert_array_equal(u.to_host(), uo)
    assert integ.cache.graph_stats()["graph_launches"] == 0
    st = integ.cache.stats()
    assert st["fused_stage_launches"] == 9 and st["dss_calls"] == 12


def test_unaligned_state_falls_back_to_the_generic_path(cts):
    # a state that starts 8 bytes off a 16-byte boundary cannot take the fused stage; the step must still
    # succeed (pass by pass) and give the oracle's bits
    ctx = cts.Context.default()
    nlev, nx, ny = 16, 4, 6
    K, T0, S = M.column_inputs(nlev, nx, ny, 0, ny, 1)
    dev = lambda a: cts.B200Vector.from_host(a, ctx)
    op = cts.ColumnOperator(nlev, nx, ny, dev(K), S_lev=dev(S), nu=1.0, halo=1, ctx=ctx)
    big = cts.B200Vector.zeros(T0.size + 1, np.float64, ctx)
    u = cts.B200Vector(big.tensor[1:], ctx)
    assert u.ptr % 16 == 8
    u.tensor.copy_(dev(T0).tensor)
    op.dss(u, None, 0.0)
    model = M.ColumnModel(nlev, nx, ny, K, S_lev=S, nu=1.0, halo=1)
    uo = T0.copy()
    model.dss(uo, None, 0.0)
    alg = cts.IMEXAlgorithm(cts.ARS343(), cts.NewtonsMethod(max_iters=1, update_j=cts.UpdateEvery(cts.NewTimeStep)))
    integ = cts.init(cts.ODEProblem(op.ode_function(), u, (0.0, 1e9), None), alg, dt=0.01, save=False)
    _, step = R.make_stepper(model.ode_function(), R.imex_algorithm("ARS343", R.NewtonsMethod(
        max_iters=1, update_j=R.UpdateEvery("NewTimeStep"))), uo)
    for k in range(2):
        cts.step_(integ)
        step(uo, None, 0.01 * k, 0.01)
    np.testing.assert_array_equal(u.to_host(), uo)


# ------------------------------------------------------------------ cast_tableau_to_state_eltype ------
@pytest.mark.parametrize("dt32", [False, True])
@pytest.mark.parametrize("fusion", [False, True])
@pytest.mark.parametrize("name,constraint", [("ARS343", None), ("ARS222", None), ("SSP333", "SSP"), ("SSP333", "Unconstrained"),
                                             ("SSP222", "SSP")])
def test_downcast_tableau_float32_state(cts, name, constraint, fusion, dt32):
    # imex_tableaus.jl:127-133 + imex_ark.jl:59-61,120-122: with cast_tableau_to_state_eltype the coefficients are rounded to
    # Float32; with a Float64 dt every product dt*Float32(a) and every broadcast is still Float64 arithmetic, with a
    # Float32 dt everything is native Float32 -- except the Shu-Osher blends of the SSPRK path, whose β stays Float64
    import test_gpu_parity as P
    model, op, T0 = P._column_pair(cts, nlev=16, nx=4, ny=6, dtype=np.float32)
    dt = np.float32(0.02) if dt32 else 0.02
    alg_o = R.imex_algorithm(name, R.NewtonsMethod(max_iters=1), constraint, cast_tableau_to_state_eltype=True)
    alg_c = cts.IMEXAlgorithm(getattr(cts, name)(), cts.NewtonsMethod(max_iters=1), constraint, cast_tableau_to_state_eltype=True)
    uo = T0.copy()
    model.dss(uo, None, 0.0)
    _, step_o = R.make_stepper(model.ode_function(), alg_o, uo)
    uc = cts.B200Vector.from_host(T0)
    op.dss(uc, None, 0.0)
    integ = cts.init(cts.ODEProblem(op.ode_function(), uc, (0.0, 1e6), None), alg_c, dt=dt)
    integ.cache.set_fusion(fusion)
    t = 0.0
    for _ in range(3):
        step_o(uo, None, t, dt)
        cts.step_(integ)
        t = t + dt
    assert uo.dtype == np.float32
    np.testing.assert_array_equal(uc.to_host(), uo)
    assert (integ.cache.stats()["fused_stage_launches"] > 0) == (fusion and not dt32)
    # and the downcast is not a no-op: the Float64 tableau gives different bits
    uo2 = T0.copy()
    model.dss(uo2, None, 0.0)
    _, step2 = R.make_stepper(model.ode_function(), R.imex_algorithm(name, R.NewtonsMethod(max_iters=1), constraint), uo2)
    t = 0.0
    for _ in range(3):
        step2(uo2, None, t, dt)
        t = t + dt
    assert not np.array_equal(uo2, uo)


# ------------------------------------------------------------------ saving / callbacks boundary ------
def test_saveat_async_d2h_host_saver(cts):
    # integrators.jl:495-526 with a host `save_func`: every save is a stream-ordered cts_memcpy_d2h into its own pinned
    # buffer (no stall of the step loop); after one synchronisation the saved states equal device-side copies
    ctx = cts.Context.default()
    op, u, model, uo = _column_problem(cts, ctx, 16, 8, 8)
    alg = cts.IMEXAlgorithm(cts.ARS343(), cts.NewtonsMethod(max_iters=1))
    saver = cts.HostSaver(ctx)
    integ = cts.init(cts.ODEProblem(op.ode_function(), u, (0.0, 0.1), None), alg, dt=0.01, saveat=(0.0, 0.03, 0.05, 0.1),
                     save_func=saver)
    cts.solve_(integ)
    host = saver.wait()
    assert integ.sol.t == [0.0, 0.03, 0.05, 0.1] and len(host) == 4
    # oracle at the same save times
    _, step = R.make_stepper(model.ode_function(), R.imex_algorithm("ARS343", R.NewtonsMethod(max_iters=1)), uo)
    want, t = [uo.copy()], 0.0
    for k in range(10):
        step(uo, None, t, 0.01)
        t += 0.01
        if k + 1 in (3, 5, 10):
            want.append(uo.copy())
    for a, b in zip(host, want):
        np.testing.assert_array_equal(a, b)


def test_wall_time_callback_single_rank(cts):
    # Callbacks.jl:40-90 on one rank: cts_allreduce_max_f64_host is the identity, the callback fires on the fake clock
    import ctypes as C
    ctx = cts.Context.default()
    v = C.c_double(123.5)
    ctx.check(ctx.lib.cts_allreduce_max_f64_host(ctx.handle, C.byref(v)))
    assert v.value == 123.5
    op, u, model, uo = _column_problem(cts, ctx, 16, 4, 6)
    clock = iter(np.arange(0.0, 100.0, 0.4))          # +0.4 s per query
    hits = []
    cb = cts.EveryXWallTimeSeconds(lambda integ: hits.append(integ.step), 1.0, ctx, clock=lambda: float(next(clock)))
    integ = cts.init(cts.ODEProblem(op.ode_function(), u, (0.0, 0.1), None),
                     cts.IMEXAlgorithm(cts.ARS343(), cts.NewtonsMethod(max_iters=1)), dt=0.01, callback=cb)
    cts.solve_(integ)
    assert hits == [3, 5, 8, 10]   # clock 0.0 at init -> next = 1.0; queries at 0.4, 0.8, 1.2 (step 3) ...


# ------------------------------------------------------------------ generic (foreign-hook) path -------------------
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("halo", [0, 1])
def test_foreign_hooks_generic_path_bit_exact(cts, dtype, halo):
    # user closures the engine cannot recognise: pass-by-pass path through host callbacks.  Without a dss! hook (halo = 0)
    # `temp = U` rides with the stage assembly and `x .-= Δx` + `T_imp = (U - temp)/dtγ` are one pass (5 passes per
    # implicit stage instead of 7); results are the oracle's bits either way
    import test_gpu_parity as P
    model, op, T0 = P._column_pair(cts, nlev=16, nx=4, ny=6, halo=halo, dtype=dtype, nu=0.7 if halo else 0.0)
    alg = cts.IMEXAlgorithm(cts.ARS343(), cts.NewtonsMethod(max_iters=1))
    uo = T0.copy()
    fo = model.ode_function()
    if halo:
        model.dss(uo, None, 0.0)
    _, step = R.make_stepper(fo, R.imex_algorithm("ARS343", R.NewtonsMethod(max_iters=1)), uo)
    res = []
    for fusion in (True, False):
        uc = cts.B200Vector.from_host(T0)
        if halo:
            op.dss(uc, None, 0.0)
        integ = cts.init(cts.ODEProblem(op.foreign_ode_function(), uc, (0.0, 1e9), None), alg, dt=0.01, save=False)
        integ.cache.set_fusion(fusion)
        for k in range(2):
            cts.step_(integ)
        assert integ.cache.stats()["fused_stage_launches"] == 0     # never the recognised-operator kernel
        res.append(uc.to_host())
    for k in range(2):
        step(uo, None, 0.01 * k, 0.01)
    np.testing.assert_array_equal(res[0], uo)
    np.testing.assert_array_equal(res[1], uo)


# ------------------------------------------------------------------ pipelined fused-stage kernel (opt-in) ----------
@pytest.mark.parametrize("dtype", ["float64", "float32"])
def test_pipelined_fused_stage_kernel_is_bit_identical(dtype):
    """`CTS_FUSED_PIPE=1` selects the persistent, warp-specialised variant of the fused column stage (operand / W / x
    rings over mbarriers) and `CTS_FUSED_SP=1` the software-pipelined TMA-only variant (lanes of solver warps + data threads,
    hardware named barriers), csrc/ops.cu.  Same device functions, same roundings: two ARS343 steps on several waves of
    tiles (stored W with the Wfact-emitting first stage, and matrix-free) must equal the default kernel's result bit for bit."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = r"""
import sys, hashlib, numpy as np
sys.path.insert(0, %r)
import cts_b200 as cts
dtype = np.dtype(%r)
out = []
for mf in (False, True):
    nlev, nx, ny = 64, 256, 150          # 152 x 256 = 38912 columns: > 148 tiles per CTA wave, partial last wave
    K = cts.B200Vector.zeros((ny + 2) * nx, dtype).fill_hash(2, 0, 0.25, 2.0)
    z = (np.arange(nlev) + 1) / (nlev + 1)
    S = cts.B200Vector.from_host((5.0 * np.exp(-((z - 0.3) ** 2) / 0.01)).astype(dtype))
    op = cts.ColumnOperator(nlev, nx, ny, K, S_lev=S, nu=1.0, halo=1).set_matrix_free(mf)
    u = cts.B200Vector.zeros(op.ndof, dtype).fill_hash(3)
    op.dss(u, None, 0.0)
    integ = cts.init(cts.ODEProblem(op.ode_function(), u, (0.0, 1e9), None),
                     cts.IMEXAlgorithm(cts.ARS343(), cts.NewtonsMethod(max_iters=1, update_j=cts.UpdateEvery(cts.NewTimeStep))),
                     dt=0.02, save=False)
    for _ in range(2):
        cts.step_(integ)
    h = u.to_host()
    assert np.isfinite(h).all()
    assert integ.cache.stats()["fused_stage_launches"] == 6
    out.append(hashlib.sha256(h.tobytes()).hexdigest())
print("HASH", *out)
""" % (root, dtype)
    hashes = {}
    for name, extra in (("default", {}), ("pipe", {"CTS_FUSED_PIPE": "1"}), ("sp", {"CTS_FUSED_SP": "1"}),
                        ("cluster", {"CTS_FUSED_CLUSTER": "1"})):
        env = dict(os.environ, **extra)
        r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr[-2000:]
        hashes[name] = [ln for ln in r.stdout.splitlines() if ln.startswith("HASH")][-1]
    assert hashes["default"] == hashes["pipe"] == hashes["sp"] == hashes["cluster"]


# ------------------------------------------------------------------ config C4 at size ------------------------------
@pytest.mark.parametrize("fuse_mvp", [False, True])
@pytest.mark.parametrize("dtype,tol,adj", [(np.float64, 1e-12, 1.0), (np.float32, 1e-5, 64.0)])
def test_c4_ssp333_jfnk_at_2pow20_dof(cts, dtype, tol, adj, fuse_mvp, monkeypatch):
    """BASELINE configs[3] in the exact set-up bench.py times (`run_c4`): SSP333 (SSPRK path) + NewtonsMethod(max_iters = 2,
    KrylovMethod(ForwardDiffJVP, ConstantForcing(1e-2))), nonlinear column diffusion, frozen-coefficient W as left
    preconditioner, smooth state with a 2 % hashed perturbation -- on 2^14 columns x 64 levels = 2^20 DOF instead of 2^27,
    against the numpy oracle (imex_ssprk.jl:95-220, newtons_method.jl:465-509,609-665; GMRES per Krylov.jl v0.10.6).
    Float64 is bit-exact (one reduction tree on both sides), Float32 within the north-star tolerance.
    `fuse_mvp`: the opt-in one-pass JVP + preconditioner solve of every Krylov iteration (cts_column_jvp_precond)."""
    monkeypatch.setenv("CTS_JVP_PRECOND_FUSION", "1" if fuse_mvp else "0")
    nlev, nx, ny = 64, 128, 128
    K = M.fill_hash(nx * ny, 2, 0, 0.25, 2.0, dtype=dtype)
    T0 = M.fill_hash(nx * ny * nlev, 3, 0, 0.02, 25.0, dtype=dtype)
    model = M.ColumnModel(nlev, nx, ny, K, S_lev=None, nu=0.0, halo=0, nonlinear=True)
    km_o = R.KrylovMethod(jacobian_free_jvp=R.ForwardDiffJVP(step_adjustment=adj), forcing_term=R.ConstantForcing(1e-2))
    alg_o = R.imex_algorithm("SSP333", R.NewtonsMethod(max_iters=2, krylov_method=km_o), None)
    uo = T0.copy()
    cache_o, step_o = R.make_stepper(model.ode_function(explicit=False), alg_o, uo)

    op = cts.ColumnOperator(nlev, nx, ny, cts.B200Vector.from_host(K), nu=0.0, halo=0, nonlinear=True)
    km = cts.KrylovMethod(jacobian_free_jvp=cts.ForwardDiffJVP(step_adjustment=adj), forcing_term=cts.ConstantForcing(1e-2))
    alg = cts.IMEXAlgorithm(cts.SSP333(), cts.NewtonsMethod(max_iters=2, krylov_method=km))
    u = cts.B200Vector.from_host(T0)
    integ = cts.init(cts.ODEProblem(op.ode_function(explicit=False), u, (0.0, 1e9), None), alg, dt=2e-3, save=False)
    t = 0.0
    for _ in range(2):
        step_o(uo, None, t, 2e-3)
        cts.step_(integ)
        t += 2e-3
    got = u.to_host()
    assert np.isfinite(got).all()
    st = integ.cache.stats()
    assert st["krylov_iters"] == sum(s["niter"] for s in cache_o["nm_cache"]["stats"]) > 0
    err = float(np.max(np.abs(got.astype(np.float64) - uo.astype(np.float64))) / np.max(np.abs(uo)))
    assert err <= 2 * tol
    if dtype == np.float64:
        np.testing.assert_array_equal(got, uo)


def test_release_after_context_destroy_is_safe(cts):
    """Finalizer order is unspecified in garbage-collected hosts (Julia finalizers, Python __del__): freeing a device
    allocation or destroying a cache after its context has been destroyed must not touch the dead context."""
    import ctypes as C
    lib = cts._lib.load()
    ctx = C.c_void_p()
    assert lib.cts_ctx_create(C.byref(ctx), 0, None) == 0
    p = C.c_void_p()
    assert lib.cts_alloc(ctx, 1 << 20, C.byref(p)) == 0
    assert lib.cts_ctx_destroy(ctx) == 0
    assert lib.cts_free(ctx, p) == 0          # plain cudaFree, no use of the dead context
    assert lib.cts_ctx_destroy(ctx) != 0      # double destroy is refused, not a double free


# ------------------------------------------------------------------ integrator steps with the other Krylov options --------
@pytest.mark.parametrize("step_kind", [1, 2, 3])
@pytest.mark.parametrize("forcing", ["constant", "eisenstat_walker"])
def test_imex_steps_with_eisenstat_walker_and_step_sizes(cts, step_kind, forcing):
    """SSP333 + NewtonsMethod(max_iters = 3) + KrylovMethod on the nonlinear column operator, stepped on the device and by the
    oracle with `EisenstatWalkerForcing` (newtons_method.jl:248-280) and `ForwardDiffStepSize1/2/3` (:107-135): same Krylov
    iteration counts, Float64 results bit for bit (the forcing term only consumes norms of the shared reduction tree)."""
    nlev, nx, ny = 16, 6, 5
    dtype = np.float64
    K = M.fill_hash(nx * ny, 2, 0, 0.25, 2.0, dtype=dtype)
    T0 = M.fill_hash(nx * ny * nlev, 3, 0, 0.05, 2.0, dtype=dtype)
    model = M.ColumnModel(nlev, nx, ny, K, S_lev=None, nu=0.0, halo=0, nonlinear=True)
    f_o = R.EisenstatWalkerForcing() if forcing == "eisenstat_walker" else R.ConstantForcing(1e-4)
    f_c = cts.EisenstatWalkerForcing() if forcing == "eisenstat_walker" else cts.ConstantForcing(1e-4)
    step_c = {1: cts.ForwardDiffStepSize1, 2: cts.ForwardDiffStepSize2, 3: cts.ForwardDiffStepSize3}[step_kind]()
    km_o = R.KrylovMethod(jacobian_free_jvp=R.ForwardDiffJVP(default_step=step_kind), forcing_term=f_o)
    km_c = cts.KrylovMethod(jacobian_free_jvp=cts.ForwardDiffJVP(default_step=step_c), forcing_term=f_c)
    uo = T0.copy()
    cache_o, step_o = R.make_stepper(model.ode_function(explicit=False),
                                     R.imex_algorithm("SSP333", R.NewtonsMethod(max_iters=3, krylov_method=km_o), None), uo)
    op = cts.ColumnOperator(nlev, nx, ny, cts.B200Vector.from_host(K), nu=0.0, halo=0, nonlinear=True)
    u = cts.B200Vector.from_host(T0)
    integ = cts.init(cts.ODEProblem(op.ode_function(explicit=False), u, (0.0, 1e9), None),
                     cts.IMEXAlgorithm(cts.SSP333(), cts.NewtonsMethod(max_iters=3, krylov_method=km_c)), dt=5e-3, save=False)
    t = 0.0
    for _ in range(3):
        step_o(uo, None, t, 5e-3)
        cts.step_(integ)
        t += 5e-3
    got = u.to_host()
    assert np.isfinite(got).all()
    assert integ.cache.stats()["krylov_iters"] == sum(s["niter"] for s in cache_o["nm_cache"]["stats"]) > 0
    np.testing.assert_array_equal(got, uo)


# ------------------------------------------------------------------ K9 + K5 in one pass ---------------------------
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("n,k,keep_dx,offset", [(4096, 1, True, 0), (4096, 2, False, 0), (100003, 3, True, 0), (8192, 20, False, 0),
                                                (8192, 32, True, 0), (4099, 2, True, 1), (64, 5, False, 0)])
def test_lincomb_sub_equals_solution_update_then_k5(cts, dtype, n, k, keep_dx, offset):
    """`cts_lincomb_sub` (GMRES solution update fused with `x .-= Δx`, newtons_method.jl:508,651) against the oracle's two
    statements -- the sequential Float64 fold rounded to FT, then the subtraction in FT -- and against the two separate
    passes (`cts_lincomb`, `cts_sub_inplace`): bit for bit, vector and scalar paths (odd n, an 8- or 4-byte offset view)."""
    import ctypes as C
    from cts_b200 import _lib as L
    ctx = cts.Context.default()
    ft = L.CTS_F64 if dtype == np.float64 else L.CTS_F32
    Vh = [(M.fill_hash(n + offset, 30 + j) - 0.5).astype(dtype) for j in range(k)]
    xh = (M.fill_hash(n + offset, 5) * 3.0).astype(dtype)
    coef_h = [0.7 * (-1) ** j / (j + 1) for j in range(k)]
    acc = np.zeros(n, np.float64)
    for c, v in zip(coef_h, Vh):
        acc = acc + c * v[offset:].astype(np.float64)
    d_want = acc.astype(dtype)
    x_want = (xh[offset:] - d_want).astype(dtype)

    item = np.dtype(dtype).itemsize
    Vd = [cts.B200Vector.from_host(v, ctx) for v in Vh]
    coef = (C.c_double * k)(*coef_h)
    ptrs = (C.c_void_p * k)(*[v.ptr + offset * item for v in Vd])
    # fused
    x1 = cts.B200Vector.from_host(xh, ctx)
    dx1 = cts.B200Vector.zeros(n + offset, dtype, ctx)
    ctx.check(ctx.lib.cts_lincomb_sub(ctx.handle, ft, n, x1.ptr + offset * item, (dx1.ptr + offset * item) if keep_dx else None, k, coef, ptrs))
    # two passes
    x2 = cts.B200Vector.from_host(xh, ctx)
    dx2 = cts.B200Vector.zeros(n + offset, dtype, ctx)
    ctx.check(ctx.lib.cts_lincomb(ctx.handle, ft, n, dx2.ptr + offset * item, k, coef, ptrs, 0))
    ctx.check(ctx.lib.cts_sub_inplace(ctx.handle, ft, n, x2.ptr + offset * item, dx2.ptr + offset * item))
    np.testing.assert_array_equal(x1.to_host()[offset:], x_want)
    np.testing.assert_array_equal(x2.to_host()[offset:], x_want)
    np.testing.assert_array_equal(x1.to_host()[:offset], xh[:offset])
    if keep_dx:
        np.testing.assert_array_equal(dx1.to_host()[offset:], d_want)
    else:
        assert not dx1.to_host().any()


# ------------------------------------------------------------------ matrix-free Newton-Krylov -----------------------
@pytest.mark.parametrize("dtype,adj", [(np.float64, 1.0), (np.float32, 64.0)])
@pytest.mark.parametrize("nlev,nx,ny", [(64, 128, 16), (32, 37, 5), (16, 9, 3)])
def test_matrix_free_newton_krylov_is_bit_identical(cts, dtype, adj, nlev, nx, ny):
    """`ColumnOperator.set_matrix_free(True)` on BASELINE configs[3] (SSP333 on the SSPRK path, NewtonsMethod(max_iters = 2,
    KrylovMethod(ForwardDiffJVP, ConstantForcing)), nonlinear column diffusion, W as left preconditioner): the Krylov tile
    kernels compute the rows of W = dtγ J(x) - I from the Newton iterate (column_wfact_kernel's expressions) instead of
    reading a stored W, and no Wfact pass runs.  With update_j = UpdateEvery(NewNewtonIteration) (the default,
    newtons_method.jl:571,626) that is the stored-W algorithm bit for bit; with UpdateEvery(NewNewtonSolve) the stored W is
    NOT the one at x, so the engine must keep the stored path (same result as without the flag)."""
    def run(mf, update_j=None):
        ctx = cts.Context.default()
        K = cts.B200Vector.zeros(ny * nx, dtype, ctx).fill_hash(2, 0, 0.25, 2.0)
        op = cts.ColumnOperator(nlev, nx, ny, K, nu=0.0, halo=0, nonlinear=True).set_matrix_free(mf)
        u = cts.B200Vector.zeros(op.ndof, dtype, ctx).fill_hash(3, 0, 0.02, 25.0)
        km = cts.KrylovMethod(jacobian_free_jvp=cts.ForwardDiffJVP(step_adjustment=adj), forcing_term=cts.ConstantForcing(1e-2))
        kw = {} if update_j is None else {"update_j": update_j}
        alg = cts.IMEXAlgorithm(cts.SSP333(), cts.NewtonsMethod(max_iters=2, krylov_method=km, **kw))
        integ = cts.init(cts.ODEProblem(op.ode_function(explicit=False), u, (0.0, 1e9), None), alg, dt=2e-3, save=False)
        b0 = ctx.bytes_count()
        for _ in range(2):
            cts.step_(integ)
        return u.to_host(), integ.cache.stats(), ctx.bytes_count() - b0

    stored, st_s, bytes_s = run(False)
    free, st_f, bytes_f = run(True)
    assert np.isfinite(stored).all()
    np.testing.assert_array_equal(free, stored)
    for key in ("newton_iters", "krylov_iters", "wfact_evals", "ldiv_calls", "jvp_evals"):
        assert st_f[key] == st_s[key], key
    assert st_s["krylov_iters"] > 0 and bytes_f < 0.9 * bytes_s
    # W refreshed once per Newton solve only: the matrix-free kernels do not apply
    upd = cts.UpdateEvery(cts.NewNewtonSolve)
    stored2, _, bytes_s2 = run(False, upd)
    free2, _, bytes_f2 = run(True, upd)
    np.testing.assert_array_equal(free2, stored2)
    assert bytes_f2 >= bytes_s2   # (a matrix-free operator also keeps off the stored-W JVP + solve tile kernel)
