# CTSB200.jl -- Julia glue that makes ClimaTimeSteppers.jl step `B200Vector` states through libcts_sm100a.so.
#
# Julia is not available in the build image, so this file is UNTESTED there; it is kept mechanical: every `ccall`
# below has a line-for-line twin in the tested Python/ctypes binding (`climatimesteppers.jl_b200/_lib.py`,
# `solvers.py`).  The reference's plug-in points it implements (SURVEY §8b):
#   * solver dispatch      `init_cache(prob, alg; dt)` / `step_u!(integrator, cache)`   src/integrators.jl:250,441-444
#   * state array type     `zero`, `copy`, `length`, `eltype`, `norm`                   src/solvers/imex_ark.jl:44-48
#   * Jacobian operator    `zero(jac_prototype)`, `ldiv!(x, W, b)`                       src/nl_solvers/newtons_method.jl:596,633
#   * user hooks           called back from the library through `@cfunction` pointers, at the reference's call sites
module CTSB200

import ClimaTimeSteppers as CTS
import LinearAlgebra

const libcts = "libcts_sm100a"          # on LD_LIBRARY_PATH / next to this file
const CTS_F32, CTS_F64 = Cint(0), Cint(1)
ft(::Type{Float32}) = CTS_F32
ft(::Type{Float64}) = CTS_F64

check(ctx, rc) = rc == 0 || error("libcts_sm100a: status $rc: " * unsafe_string(ccall((:cts_last_error, libcts), Cstring, (Ptr{Cvoid},), ctx)))

# ------------------------------------------------------------------------------------------------ context
mutable struct Context
    handle::Ptr{Cvoid}
    function Context(device::Integer = 0, stream::Ptr{Cvoid} = C_NULL)
        h = Ref{Ptr{Cvoid}}(C_NULL)
        rc = ccall((:cts_ctx_create, libcts), Cint, (Ref{Ptr{Cvoid}}, Cint, Ptr{Cvoid}), h, device, stream)
        rc == 0 || error("cts_ctx_create failed with status $rc (no GPU => CTS_ENODEVICE: there is no CPU fallback)")
        finalizer(c -> ccall((:cts_ctx_destroy, libcts), Cint, (Ptr{Cvoid},), c.handle), new(h[]))
    end
end

# ------------------------------------------------------------------------------------------------ state type
"""Flat device vector; the state array type the B200 `step_u!` methods dispatch on."""
mutable struct B200Vector{FT} <: AbstractVector{FT}
    ctx::Context
    ptr::Ptr{Cvoid}          # device pointer
    n::Int
    owned::Bool
end
function B200Vector{FT}(ctx::Context, n::Integer) where {FT}
    p = Ref{Ptr{Cvoid}}(C_NULL)
    check(ctx.handle, ccall((:cts_alloc, libcts), Cint, (Ptr{Cvoid}, Csize_t, Ref{Ptr{Cvoid}}), ctx.handle, n * sizeof(FT), p))
    v = B200Vector{FT}(ctx, p[], n, true)
    finalizer(x -> x.owned && ccall((:cts_free, libcts), Cint, (Ptr{Cvoid}, Ptr{Cvoid}), x.ctx.handle, x.ptr), v)
end
B200Vector(ctx::Context, host::Vector{FT}) where {FT} = (v = B200Vector{FT}(ctx, length(host));
    check(ctx.handle, ccall((:cts_memcpy_h2d, libcts), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{FT}, Csize_t), ctx.handle, v.ptr, host, sizeof(host))); v)
Base.size(v::B200Vector) = (v.n,)
Base.eltype(::B200Vector{FT}) where {FT} = FT
Base.zero(v::B200Vector{FT}) where {FT} = B200Vector{FT}(v.ctx, v.n)                     # cts_alloc zero-fills
function Base.copy(v::B200Vector{FT}) where {FT}
    w = B200Vector{FT}(v.ctx, v.n)
    check(v.ctx.handle, ccall((:cts_memcpy_d2d, libcts), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Csize_t), v.ctx.handle, w.ptr, v.ptr, v.n * sizeof(FT)))
    w
end
function Base.Array(v::B200Vector{FT}) where {FT}
    h = Vector{FT}(undef, v.n)
    check(v.ctx.handle, ccall((:cts_memcpy_d2h_sync, libcts), Cint, (Ptr{Cvoid}, Ptr{FT}, Ptr{Cvoid}, Csize_t), v.ctx.handle, h, v.ptr, sizeof(h)))
    h
end
function LinearAlgebra.norm(v::B200Vector{FT}) where {FT}                                 # newtons_method.jl:108,120; convergence_checker.jl:81-82
    out = Ref{Cdouble}(0)
    check(v.ctx.handle, ccall((:cts_nrm2, libcts), Cint, (Ptr{Cvoid}, Cint, Csize_t, Ptr{Cvoid}, Ref{Cdouble}), v.ctx.handle, ft(FT), v.n, v.ptr, out))
    FT(out[])
end
function LinearAlgebra.dot(x::B200Vector{FT}, y::B200Vector{FT}) where {FT}
    out = Ref{Cdouble}(0)
    check(x.ctx.handle, ccall((:cts_dot, libcts), Cint, (Ptr{Cvoid}, Cint, Csize_t, Ptr{Cvoid}, Ptr{Cvoid}, Ref{Cdouble}), x.ctx.handle, ft(FT), x.n, x.ptr, y.ptr, out))
    FT(out[])
end
view_of(ctx, ptr, n, ::Type{FT}) where {FT} = B200Vector{FT}(ctx, ptr, n, false)           # non-owning view handed to hooks

# ------------------------------------------------------------------------------------------------ C structs (include/cts.h)
struct UpdateHandler
    every_signal::Cint
    fn::Ptr{Cvoid}
    ud::Ptr{Cvoid}
end
signal_id(::Type{CTS.NewTimeStep}) = Cint(0)
signal_id(::Type{CTS.NewNewtonSolve}) = Cint(1)
signal_id(::Type{CTS.NewNewtonIteration}) = Cint(2)
signal_id(::Type{CTS.WithDSS}) = Cint(3)
signal_id(::Type{CTS.EndOfStage}) = Cint(4)
signal_id(::Type{CTS.EndOfStep}) = Cint(5)
UpdateHandler(::CTS.UpdateEvery{U}) where {U} = UpdateHandler(signal_id(U), C_NULL, C_NULL)
# (UpdateEveryN / UpdateEveryDt travel as a host callback `fn(ud, signal, t)` built with @cfunction, as in solvers.py)

struct OdeFunction
    ud::Ptr{Cvoid}
    T_exp_T_lim::Ptr{Cvoid}; has_lim::Cint
    T_imp::Ptr{Cvoid}; Wfact::Ptr{Cvoid}; ldiv::Ptr{Cvoid}; jac_mul::Ptr{Cvoid}; W::Ptr{Cvoid}
    lim::Ptr{Cvoid}; dss::Ptr{Cvoid}; constrain_state::Ptr{Cvoid}; cache::Ptr{Cvoid}; cache_imp::Ptr{Cvoid}
    initialize_imp::Ptr{Cvoid}; initialize_imp_is_nothing::Cint
    update_cache::UpdateHandler; update_constrain_state::UpdateHandler
end
struct ImexTableau
    s::Cint
    a_exp::NTuple{64, Cdouble}; b_exp::NTuple{8, Cdouble}; c_exp::NTuple{8, Cdouble}
    a_imp::NTuple{64, Cdouble}; b_imp::NTuple{8, Cdouble}; c_imp::NTuple{8, Cdouble}
    cast_to_state_eltype::Cint
end
function ImexTableau(tab::CTS.IMEXTableau, cast::Bool)
    s = length(tab.b_exp)
    rowmajor(a) = ntuple(k -> (i = (k - 1) ÷ s + 1; j = (k - 1) % s + 1; k <= s * s ? Cdouble(a[i, j]) : 0.0), 64)
    vec8(b) = ntuple(k -> k <= s ? Cdouble(b[k]) : 0.0, 8)
    ImexTableau(s, rowmajor(tab.a_exp), vec8(tab.b_exp), vec8(tab.c_exp), rowmajor(tab.a_imp), vec8(tab.b_imp), vec8(tab.c_imp), cast)
end
struct NewtonConfig
    present::Cint; max_iters::Cint; update_j::UpdateHandler
    use_krylov::Cint; jacobian_free_jvp::Cint; jvp_step_kind::Cint; jvp_step_adjustment::Cdouble
    disable_preconditioner::Cint; memory::Cint; itmax::Cint
    forcing_kind::Cint; forcing_rtol::Cdouble
    ew_initial_rtol::Cdouble; ew_gamma::Cdouble; ew_alpha::Cdouble; ew_min_rtol_threshold::Cdouble; ew_max_rtol::Cdouble
    batch_gram_schmidt::Cint
end
NewtonConfig(::Nothing) = NewtonConfig(0, 0, UpdateHandler(2, C_NULL, C_NULL), 0, 0, 3, 1.0, 0, 20, 0, 0, 0.0, 0.5, 1.0, 2.0, 0.1, 0.9, 0)
function NewtonConfig(nm::CTS.NewtonsMethod)
    km = nm.krylov_method
    jf = !isnothing(km) && !isnothing(km.jacobian_free_jvp)
    rtol = !isnothing(km) && km.forcing_term isa CTS.ConstantForcing ? Cdouble(km.forcing_term.rtol) : 0.0
    ew = !isnothing(km) && km.forcing_term isa CTS.EisenstatWalkerForcing ? km.forcing_term : CTS.EisenstatWalkerForcing()
    NewtonConfig(1, nm.max_iters, UpdateHandler(nm.update_j), !isnothing(km), jf,
        jf ? (km.jacobian_free_jvp.default_step isa CTS.ForwardDiffStepSize1 ? 1 : km.jacobian_free_jvp.default_step isa CTS.ForwardDiffStepSize2 ? 2 : 3) : 3,
        jf ? Cdouble(km.jacobian_free_jvp.step_adjustment) : 1.0, !isnothing(km) && km.disable_preconditioner,
        isnothing(km) ? 20 : get(km.kwargs, :memory, 20), 0, (!isnothing(km) && km.forcing_term isa CTS.EisenstatWalkerForcing) ? 1 : 0, rtol,
        ew.initial_rtol, ew.γ, ew.α, ew.min_rtol_threshold, ew.max_rtol, 0)
end

# ------------------------------------------------------------------------------------------------ hooks -> @cfunction
# One closure environment per cache; `ud` is unused because Julia closures capture `p`, the views and the Jacobian.
mutable struct HookEnv{F, P, FT}
    f::F; p::P; ctx::Context; n::Int; jac::Any; err::Any
end
V(env::HookEnv{F, P, FT}, ptr) where {F, P, FT} = view_of(env.ctx, ptr, env.n, FT)
guard(env, body) = try body(); Cint(0) catch e; env.err = e; Cint(1) end   # exceptions never cross the C ABI

function hook_table(env::HookEnv)
    f = env.f
    c(fun, ret, args) = fun === nothing ? C_NULL : Base.unsafe_convert(Ptr{Cvoid}, fun)
    exp_lim = isnothing(f.T_exp_T_lim!) ? C_NULL :
        @cfunction($((ud, de, dl, u, t) -> guard(env, () -> f.T_exp_T_lim!(V(env, de), V(env, dl), V(env, u), env.p, t))), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Cdouble))
    timp = isnothing(f.T_imp!) ? C_NULL :
        @cfunction($((ud, du, u, t) -> guard(env, () -> f.T_imp!(V(env, du), V(env, u), env.p, t))), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Cdouble))
    wfact = isnothing(env.jac) ? C_NULL :
        @cfunction($((ud, W, u, dtγ, t) -> guard(env, () -> f.T_imp!.Wfact(env.jac, V(env, u), env.p, dtγ, t))), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Cdouble, Cdouble))
    ldiv = isnothing(env.jac) ? C_NULL :
        @cfunction($((ud, x, W, b) -> guard(env, () -> LinearAlgebra.ldiv!(V(env, x), env.jac, V(env, b)))), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}))
    state(h) = h isa Returns{Nothing} ? C_NULL :
        @cfunction($((ud, u, t) -> guard(env, () -> h(V(env, u), env.p, t))), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Cdouble))
    init_imp = (isnothing(f.initialize_imp!) || f.initialize_imp! isa Returns{Nothing}) ? C_NULL :
        @cfunction($((ud, u, dtγ) -> guard(env, () -> f.initialize_imp!(V(env, u), env.p, dtγ))), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Cdouble))
    lim = f.lim! isa Returns{Nothing} ? C_NULL :
        @cfunction($((ud, u, t, uref) -> guard(env, () -> f.lim!(V(env, u), env.p, t, V(env, uref)))), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Cdouble, Ptr{Cvoid}))
    ptr(x) = x === C_NULL ? C_NULL : Base.unsafe_convert(Ptr{Cvoid}, x)
    keep = (exp_lim, timp, wfact, ldiv, init_imp, lim, state(f.dss!), state(f.constrain_state!), state(f.cache!), state(f.cache_imp!))
    tbl = OdeFunction(C_NULL, ptr(exp_lim), CTS.has_T_lim(f), ptr(timp), ptr(wfact), ptr(ldiv), C_NULL,
        isnothing(env.jac) ? C_NULL : pointer_from_objref(env.jac), ptr(lim), ptr(keep[7]), ptr(keep[8]), ptr(keep[9]), ptr(keep[10]),
        ptr(init_imp), isnothing(f.initialize_imp!), UpdateHandler(f.update_cache), UpdateHandler(f.update_constrain_state))
    tbl, keep
end

# ------------------------------------------------------------------------------------------------ solver dispatch
mutable struct B200IMEXCache
    handle::Ptr{Cvoid}; env::HookEnv; keep::Any
end
has_jac(T) = !isnothing(T) && hasfield(typeof(T), :Wfact) && !isnothing(T.Wfact) && !isnothing(T.jac_prototype)   # imex_ark.jl:3-7

function CTS.init_cache(prob::CTS.ODEProblem{<:Any, <:B200Vector{FT}}, alg::CTS.IMEXAlgorithm; kwargs...) where {FT}
    u0, f = prob.u0, prob.f
    jac = (has_jac(f.T_imp!) && !isnothing(alg.newtons_method)) ? zero(f.T_imp!.jac_prototype) : nothing     # newtons_method.jl:596
    env = HookEnv{typeof(f), typeof(prob.p), FT}(f, prob.p, u0.ctx, length(u0), jac, nothing)
    tbl, keep = hook_table(env)
    tab = ImexTableau(alg.tableau, alg.options.cast_tableau_to_state_eltype)
    ncfg = NewtonConfig(alg.newtons_method)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    check(u0.ctx.handle, ccall((:cts_imex_cache_create, libcts), Cint,
        (Ptr{Cvoid}, Cint, Csize_t, Ref{ImexTableau}, Cint, Ref{OdeFunction}, Ref{NewtonConfig}, Ref{Ptr{Cvoid}}),
        u0.ctx.handle, ft(FT), length(u0), tab, alg.constraint isa CTS.SSP ? 1 : 0, tbl, ncfg, h))
    configure_newton(u0.ctx, h[], alg.newtons_method)
    c = B200IMEXCache(h[], env, keep)
    finalizer(x -> ccall((:cts_imex_cache_destroy, libcts), Cint, (Ptr{Cvoid},), x.handle), c)
end

# ConvergenceChecker / LineSearch (convergence_checker.jl, convergence_condition.jl, line_search.jl): condition trees are
# flattened to postfix order for `cts_convergence_checker`
struct CondNode
    kind::Cint; nchildren::Cint; p0::Cdouble; p1::Cdouble
end
struct ConvergenceCheckerC
    n_norm::Cint; norm_cond::NTuple{16, CondNode}; n_comp::Cint; comp_cond::NTuple{16, CondNode}; combiner_or::Cint
end
cond_nodes(c::CTS.MaximumError) = [CondNode(0, 0, c.max_err, 0)]
cond_nodes(c::CTS.MaximumRelativeError) = [CondNode(1, 0, c.max_rel_err, 0)]
cond_nodes(c::CTS.MaximumErrorReduction) = [CondNode(2, 0, c.max_reduction, 0)]
cond_nodes(c::CTS.MinimumRateOfConvergence) = [CondNode(3, 0, c.rate, c.order)]
cond_nodes(c::CTS.MultipleConditions) =
    vcat(mapreduce(cond_nodes, vcat, reverse(collect(c.conditions))), [CondNode(c.condition_combiner === all ? 4 : 5, length(c.conditions), 0, 0)])
cond_nodes(::Nothing) = CondNode[]
pad16(v) = ntuple(k -> k <= length(v) ? v[k] : CondNode(0, 0, 0, 0), 16)
function configure_newton(ctx, cache_handle, nm)
    isnothing(nm) && return
    (isnothing(nm.convergence_checker) && isnothing(nm.line_search)) && return
    h = Ref{Ptr{Cvoid}}(C_NULL)
    check(ctx.handle, ccall((:cts_imex_cache_newton, libcts), Cint, (Ptr{Cvoid}, Ref{Ptr{Cvoid}}), cache_handle, h))
    if !isnothing(nm.convergence_checker)
        cc = nm.convergence_checker
        cc.norm === LinearAlgebra.norm || error("the B200 path evaluates LinearAlgebra.norm only")
        nn, nc = cond_nodes(cc.norm_condition), cond_nodes(cc.component_condition)
        chk = ConvergenceCheckerC(length(nn), pad16(nn), length(nc), pad16(nc), cc.condition_combiner === (|) ? 1 : 0)
        check(ctx.handle, ccall((:cts_newton_cache_set_convergence_checker, libcts), Cint, (Ptr{Cvoid}, Ref{ConvergenceCheckerC}), h[], chk))
    end
    isnothing(nm.line_search) || check(ctx.handle, ccall((:cts_newton_cache_set_line_search, libcts), Cint, (Ptr{Cvoid}, Cint), h[], 1))
end

function CTS.step_u!(integrator, cache::B200IMEXCache)                                   # integrators.jl:441-444
    (; u, t, dt) = integrator
    rc = ccall((:cts_imex_step, libcts), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Cdouble, Cdouble, Cint),
        cache.handle, u.ptr, Float64(t), Float64(dt), dt isa Float32)
    isnothing(cache.env.err) || (e = cache.env.err; cache.env.err = nothing; throw(e))
    check(u.ctx.handle, rc)
    return u
end

mutable struct B200LSRKCache
    handle::Ptr{Cvoid}; keep::Any
end
function CTS.init_cache(prob::CTS.ODEProblem{<:Any, <:B200Vector{FT}}, alg::CTS.LowStorageRungeKutta2N; kwargs...) where {FT}
    u0 = prob.u0
    tab = CTS.tableau(alg, FT)                                                            # lsrk.jl:51
    A, B, C = Cdouble.(collect(tab.A)), Cdouble.(collect(tab.B)), Cdouble.(collect(tab.C))
    cb = @cfunction($((ud, du, u, t, α, β) -> (prob.f(view_of(u0.ctx, du, length(u0), FT), view_of(u0.ctx, u, length(u0), FT), prob.p, t, α, FT(β)); Cint(0))),
        Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Cdouble, Cdouble, Cdouble))
    h = Ref{Ptr{Cvoid}}(C_NULL)
    check(u0.ctx.handle, ccall((:cts_lsrk_cache_create, libcts), Cint,
        (Ptr{Cvoid}, Cint, Csize_t, Cint, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cvoid}, Ptr{Cvoid}, Ref{Ptr{Cvoid}}),
        u0.ctx.handle, ft(FT), length(u0), length(A), A, B, C, cb, C_NULL, h))
    finalizer(x -> ccall((:cts_lsrk_cache_destroy, libcts), Cint, (Ptr{Cvoid},), x.handle), B200LSRKCache(h[], cb))
end
CTS.step_u!(integrator, cache::B200LSRKCache) =
    (check(integrator.u.ctx.handle, ccall((:cts_lsrk_step, libcts), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Cdouble, Cdouble, Cint),
        cache.handle, integrator.u.ptr, Float64(integrator.t), Float64(integrator.dt), integrator.dt isa Float32)); integrator.u)

# Rosenbrock (rosenbrock.jl:100-124, 136-239)
struct RosenbrockTableauC
    s::Cint; A::NTuple{64, Cdouble}; alpha::NTuple{64, Cdouble}; C::NTuple{64, Cdouble}; Gamma::NTuple{64, Cdouble}; m::NTuple{8, Cdouble}
end
function RosenbrockTableauC(tab::CTS.RosenbrockTableau{N}) where {N}
    pad(M) = ntuple(k -> (i = (k - 1) ÷ N + 1; j = (k - 1) % N + 1; k <= N * N ? Cdouble(M[i, j]) : 0.0), 64)   # row-major s x s
    RosenbrockTableauC(N, pad(tab.A), pad(tab.α), pad(tab.C), pad(tab.Γ), ntuple(k -> k <= N ? Cdouble(tab.m[k]) : 0.0, 8))
end
mutable struct B200RosenbrockCache
    handle::Ptr{Cvoid}; env::HookEnv; keep::Any
end
function CTS.init_cache(prob::CTS.ODEProblem{<:Any, <:B200Vector{FT}}, alg::CTS.RosenbrockAlgorithm; kwargs...) where {FT}
    u0, f = prob.u0, prob.f
    jac = isnothing(f.T_imp!) ? nothing : f.T_imp!.jac_prototype                          # the prototype itself (:117)
    env = HookEnv{typeof(f), typeof(prob.p), FT}(f, prob.p, u0.ctx, length(u0), jac, nothing)
    tbl, keep = hook_table(env)
    tg = (isnothing(f.T_imp!) || isnothing(f.T_imp!.tgrad)) ? C_NULL :
         @cfunction($((ud, dT, u, t) -> (f.T_imp!.tgrad(view_of(u0.ctx, dT, length(u0), FT), view_of(u0.ctx, u, length(u0), FT), prob.p, t); Cint(0))),
                    Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Cdouble))
    h = Ref{Ptr{Cvoid}}(C_NULL)
    check(u0.ctx.handle, ccall((:cts_rosenbrock_cache_create, libcts), Cint,
        (Ptr{Cvoid}, Cint, Csize_t, Ref{RosenbrockTableauC}, Ref{OdeFunction}, Ptr{Cvoid}, Ref{Ptr{Cvoid}}),
        u0.ctx.handle, ft(FT), length(u0), RosenbrockTableauC(alg.tableau), tbl, tg, h))
    c = B200RosenbrockCache(h[], env, (keep, tg))
    finalizer(x -> ccall((:cts_rosenbrock_cache_destroy, libcts), Cint, (Ptr{Cvoid},), x.handle), c)
end
function CTS.step_u!(integrator, cache::B200RosenbrockCache)
    (; u, t, dt) = integrator
    rc = ccall((:cts_rosenbrock_step, libcts), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Cdouble, Cdouble), cache.handle, u.ptr, Float64(t), Float64(dt))
    isnothing(cache.env.err) || (e = cache.env.err; cache.env.err = nothing; throw(e))
    check(u.ctx.handle, rc)
    return u
end

# ------------------------------------------------------------------------------------------------ column tridiagonal W
"""Batched tridiagonal `jac_prototype` for column-stacked states: `zero` + `ldiv!` (cts_tridiag_solve_batched)."""
struct Tridiag
    nlev::Cint; ncols::Csize_t; dl::Ptr{Cvoid}; d::Ptr{Cvoid}; du::Ptr{Cvoid}
end
struct B200ColumnTridiag{FT}
    dl::B200Vector{FT}; d::B200Vector{FT}; du::B200Vector{FT}; nlev::Int
end
Base.zero(W::B200ColumnTridiag{FT}) where {FT} = B200ColumnTridiag{FT}(zero(W.dl), zero(W.d), zero(W.du), W.nlev)
function LinearAlgebra.ldiv!(x::B200Vector{FT}, W::B200ColumnTridiag{FT}, b::B200Vector{FT}) where {FT}
    t = Tridiag(W.nlev, length(b) ÷ W.nlev, W.dl.ptr, W.d.ptr, W.du.ptr)
    check(x.ctx.handle, ccall((:cts_tridiag_solve_batched, libcts), Cint, (Ptr{Cvoid}, Cint, Ref{Tridiag}, Ptr{Cvoid}, Ptr{Cvoid}),
        x.ctx.handle, ft(FT), t, b.ptr, x.ptr))
    x
end

end # module
