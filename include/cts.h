/* cts.h -- C ABI of libcts_sm100a.so: the B200-native (sm_100a) per-step state-update path of
 * CliMA/ClimaTimeSteppers.jl (reference v0.10.0).
 *
 * The reference is pure Julia and has no FFI; its plug-in points are Julia multiple dispatch on
 *   - `init_cache(prob, alg; dt)` / `step_u!(integrator, cache)`        src/integrators.jl:250,441-444
 *   - `ClimaODEFunction` hooks (T_exp_T_lim!, T_imp!, dss!, cache!, ...)  src/functions.jl:81-144
 *   - `T_imp!.jac_prototype` with `zero`/`ldiv!`, `T_imp!.Wfact`          src/nl_solvers/newtons_method.jl:596,633
 *   - array arithmetic on the state type (`copyto!(::Broadcasted)`, `norm`, `sum`).
 * Every entry point below replaces one of those call sites (cited per function).  A Julia binding
 * (`ccall((:sym, "libcts_sm100a"), Cint, (...), ...)`) is shown in INTEGRATION.md; the tested
 * binding in this repo is the Python/ctypes mirror of the same interface.
 *
 * Conventions
 *   - Plain C types only (pointers, sizes, doubles).  No exceptions cross the ABI.
 *   - Every function returns 0 on success or a negative cts_status; cts_last_error(ctx) gives text.
 *   - All buffers are DEVICE pointers owned by the caller unless stated otherwise.  Work is enqueued
 *     on the context's CUDA stream and is asynchronous w.r.t. the host unless the function returns a
 *     host scalar (reductions) or its name ends in `_sync`.
 *   - A context is bound to one GPU and one host thread (one process per GPU).
 *   - There is no CPU fallback: every compute entry point fails with CTS_ENODEVICE without a GPU.
 */
#ifndef CTS_H
#define CTS_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CTS_ABI_VERSION 1
#define CTS_MAX_STAGES 8   /* largest reference tableau: ARK548L2SA2, s = 8 (imex_tableaus.jl:958) */
#define CTS_MAX_TERMS 16   /* operands of one fused increment: <= 2*s (fused_increment.jl:115-172)    */

typedef struct cts_ctx cts_ctx;

typedef enum { CTS_F32 = 0, CTS_F64 = 1 } cts_dtype;

typedef enum {
  CTS_OK = 0,
  CTS_EINVAL = -1,     /* bad argument                                    */
  CTS_ECUDA = -2,      /* CUDA runtime error (text in cts_last_error)     */
  CTS_ENCCL = -3,      /* NCCL error                                      */
  CTS_ENODEVICE = -4,  /* no usable GPU: the library has no CPU path      */
  CTS_ECALLBACK = -5,  /* a user hook returned non-zero                   */
  CTS_EUNSUPPORTED = -6
} cts_status;

/* ------------------------------------------------------------------ context / memory --------- */
const char* cts_version(void);
int cts_abi_version(void);
/* `cuda_stream`: a cudaStream_t to enqueue on (0x1 = cudaStreamLegacy, the default stream), or NULL to create a
 * private non-blocking stream. */
int cts_ctx_create(cts_ctx** out, int device, void* cuda_stream);
int cts_ctx_destroy(cts_ctx* ctx);
int cts_ctx_sync(cts_ctx* ctx);
const char* cts_last_error(cts_ctx* ctx);
void* cts_ctx_stream(cts_ctx* ctx);
/* number of kernels this context has launched so far (evidence for bench.py's gpu_launches) */
int cts_launch_count(cts_ctx* ctx, uint64_t* out);
/* algorithmic HBM bytes of the kernels this context has launched so far: (arrays read + arrays written) x elements x
 * sizeof(FT) per pass, the per-kernel figures of the design document -- the numerator of bench.py's achieved GB/s for
 * workloads whose pass count is data dependent (Newton-Krylov) */
int cts_bytes_count(cts_ctx* ctx, uint64_t* out);
/* NVTX ranges for the host side of the reference's annotated functions (`NVTX.@annotate function solve!`,
 * integrators.jl:345; the library itself opens "step_u!" (integrators.jl:442), "solve_newton!" (newtons_method.jl:609) and
 * "solve_krylov!" (newtons_method.jl:465) around its own entry points).  No-ops unless a profiler is attached. */
int cts_nvtx_range_push(const char* name);
int cts_nvtx_range_pop(void);
/* `zero(u0)` in init_cache (imex_ark.jl:44-48): zero-filled device allocation */
int cts_alloc(cts_ctx* ctx, size_t bytes, void** dptr);
int cts_free(cts_ctx* ctx, void* dptr);
int cts_memset_zero(cts_ctx* ctx, void* dptr, size_t bytes);
int cts_memcpy_h2d(cts_ctx* ctx, void* dst_dev, const void* src_host, size_t bytes);  /* stream-ordered; host buffer may be pageable (then synchronous) */
int cts_memcpy_d2h_sync(cts_ctx* ctx, void* dst_host, const void* src_dev, size_t bytes);
int cts_memcpy_d2h(cts_ctx* ctx, void* dst_host, const void* src_dev, size_t bytes);       /* stream-ordered (pinned host buffer); the saving callback's async D2H */
int cts_memcpy_d2d(cts_ctx* ctx, void* dst_dev, const void* src_dev, size_t bytes);

/* ------------------------------------------------------------------ per-kernel timing ---------- */
/* The counterpart of the reference's `benchmark_step` table (ext/ClimaTimeSteppersBenchmarkToolsExt.jl:42-146):
 * when enabled, the step engine brackets every kernel family it launches with CUDA events on the context's
 * stream; cts_prof_read synchronises, returns the accumulated device time (ms) and launch count per
 * category since the last read, and resets the accumulators. */
typedef enum {
  CTS_PROF_STAGE_ASSEMBLY = 0, /* K1 stage value (compute_stage_value!)            */
  CTS_PROF_FINAL_UPDATE = 1,   /* K1 final update of u                             */
  CTS_PROF_FUSED_STAGE = 2,    /* fused K1+Newton+K6 column stage                  */
  CTS_PROF_T_EXP = 3,          /* T_exp_T_lim! hook                                */
  CTS_PROF_T_IMP = 4,          /* T_imp! hook                                      */
  CTS_PROF_WFACT = 5,          /* Wfact hook                                       */
  CTS_PROF_LDIV = 6,           /* ldiv! hook (K4)                                  */
  CTS_PROF_NEWTON_EW = 7,      /* K2, K3, K5, K6, K7, K8 element-wise passes       */
  CTS_PROF_DSS = 8,            /* dss! hook (halo exchange)                        */
  CTS_PROF_KRYLOV = 9,         /* GMRES vector work (dots, norms, axpys)           */
  CTS_PROF_LSRK_FUSED = 10,    /* fused LSRK stage (K10f + K10)                    */
  CTS_PROF_LSRK_RHS = 11,      /* user f(du,u,p,t,1,A) (K10f)                      */
  CTS_PROF_LSRK_UPDATE = 12,   /* u += dtB*du (K10)                                */
  CTS_PROF_OTHER_HOOKS = 13,   /* cache!/constrain_state!/lim!/initialize_imp!     */
  CTS_PROF_NCAT = 14
} cts_prof_category;
int cts_prof_enable(cts_ctx* ctx, int enabled);
int cts_prof_read(cts_ctx* ctx, double* ms_per_category /*[CTS_PROF_NCAT]*/, uint64_t* launches_per_category /*[CTS_PROF_NCAT]*/);

/* ------------------------------------------------------------------ element-wise kernels ----- */
/* K1/K2/K11  fused multi-operand increment -- replaces `assign_with_increments!`,
 * `assign_fused_increment!`, `fused_increment!` (src/utilities/fused_increment.jl:183-236) and the
 * Shu-Osher updates of src/solvers/imex_ssprk.jl:112-126,155-169:
 *     out[i] = ((c0*base[i] + G1[i]) + G2[i]),   Gk = left fold  sum_j coef[j]*terms[j][i]
 * c0 == 1 emits no multiply; a group with zero terms is skipped (NullBroadcasted); n_grp1+n_grp2 == 0
 * is a copy (K2, `@. temp = U`, imex_ark.jl:209).  `coef[j]` is the host-rounded scalar fl(dt*a_ij)
 * (fused_increment.jl:58).  `out2` (optional) receives the same values (fuses `temp = U`).
 * `compute_native` != 0 evaluates in the state precision (Float32 dt / cast tableau), otherwise every
 * expression is evaluated in Float64 and rounded on store (Float64 tableau).
 * `out` may alias `base` or any term.  No FMA contraction: results are bit-identical to the oracle. */
int cts_axpy_n(cts_ctx* ctx, cts_dtype ft, size_t n, void* out, void* out2, double c0, const void* base,
               int n_grp1, int n_grp2, const double* coef /*host*/, const void* const* terms /*host array of device ptrs*/,
               int compute_native);

/* K11 -- one Shu-Osher stage assembly of IMEX-SSPRK in a single pass; replaces the chained broadcasts
 * `@. U_exp += dt*T_exp` (imex_ssprk.jl:117,162), `@. U_exp = (1-β)*u + β*U_exp` (:165), `@. U = U_exp + inc_imp` (:169)
 * and `@. temp = U` (:190), or with final_update != 0 `@. u = (1-β)*u + β*U_exp + inc_imp` (:120-122).  Each operation
 * is rounded exactly as in the separate passes.  U_exp_out / U_out / temp_out may be NULL (not both of the first two);
 * T_exp may be NULL (no explicit tendency); in-place use (U_exp_out == U_exp, U_out == u) is allowed. */
int cts_ssprk_stage(cts_ctx* ctx, cts_dtype ft, size_t n, void* U_exp_out, void* U_out, void* temp_out, const void* U_exp,
                    const void* T_exp, const void* u, double dt, int dt_native, double beta, double one_minus_beta,
                    int n_imp, const double* coef, const void* const* terms, int compute_native, int final_update);

/* out = ((((base + c1*x1) + c2*x2) + ...) + cn*xn) [* scale]: the chains of `y .+= c .* x` broadcasts (and the trailing
 * `y .*= s`) of the Rosenbrock step (rosenbrock.jl:178-214,231-233) in one pass.  Float64 arithmetic with Float64
 * coefficients; a Float32 state is rounded after every term, as each of the reference's passes stores it.
 * base == NULL starts from zero (`fill!(y, 0)`); out may alias base. */
int cts_axpy_chain(cts_ctx* ctx, cts_dtype ft, size_t n, void* out, const void* base, int nterms, const double* coef,
                   const void* const* terms, double scale, int has_scale);
/* K3  `@. residual = U0 + dtγ * residual - U'` (imex_ark.jl:297); r holds T_imp(U') on entry */
int cts_newton_residual(cts_ctx* ctx, cts_dtype ft, size_t n, void* r, const void* U0, double dtgamma, const void* Up);
/* K5  `x .-= Δx` (newtons_method.jl:651) */
int cts_sub_inplace(cts_ctx* ctx, cts_dtype ft, size_t n, void* x, const void* dx);
/* K6  `@. T_imp[i] = (U - temp) / dtγ` (imex_ark.jl:269, imex_ssprk.jl:216) */
int cts_timp_from_stage(cts_ctx* ctx, cts_dtype ft, size_t n, void* T, const void* U, const void* temp, double dtgamma);
/* K7  `@. x2 = x + ε * Δx`; K8 `@. jΔx = (f2 - f) / ε` (newtons_method.jl:178,181); ε has the state eltype */
int cts_jvp_perturb(cts_ctx* ctx, cts_dtype ft, size_t n, void* x2, const void* x, double eps, const void* dx);
int cts_jvp_quotient(cts_ctx* ctx, cts_dtype ft, size_t n, void* out, const void* f2, const void* f, double eps);
/* K10 `u .+= (dt * B[stage]) .* du` (lsrk.jl:65); dtB = fl(dt*B) in promote(typeof(dt), FT) */
int cts_lsrk_update(cts_ctx* ctx, cts_dtype ft, size_t n, void* u, double dtB, const void* du, int compute_native);
/* y = a*x + b*y  (GMRES basis updates; Krylov.jl kaxpy!/kscal!) evaluated in Float64, rounded on store */
int cts_axpby(cts_ctx* ctx, cts_dtype ft, size_t n, double a, const void* x, double b, void* y);
/* y = x / denom (Arnoldi normalisation V[k] = q / ||q||), evaluated in Float64, rounded on store */
int cts_div_scalar(cts_ctx* ctx, cts_dtype ft, size_t n, void* y, const void* x, double denom);
/* out = sum_j coef[j]*V[j] (GMRES solution update x = V y): sequential Float64 fold, rounded once on store;
 * accumulate != 0 starts the fold from the current contents of `out` */
int cts_lincomb(cts_ctx* ctx, cts_dtype ft, size_t n, void* out, int k, const double* coef /*host*/,
                const void* const* V /*host array of device ptrs*/, int accumulate);

/* K9 + K5 in one pass (newtons_method.jl:508 `Δx .= Krylov.solution(solver)` and :651 `x .-= Δx`): d = sum_j coef[j]*V[j]
 * (the same fold and rounding as cts_lincomb), x -= d in the state precision; `dx` (may be NULL) receives d when a
 * convergence checker or the line search reads Δx afterwards.  1 <= k <= 32. */
int cts_lincomb_sub(cts_ctx* ctx, cts_dtype ft, size_t n, void* x, void* dx, int k, const double* coef /*host*/,
                    const void* const* V /*host array of device ptrs*/);

/* ------------------------------------------------------------------ reductions (R1-R4) ------- */
/* `norm`, `sum(x_i -> 1 + abs(x_i), x)` and Krylov dots (newtons_method.jl:108,120,135,265; convergence_checker.jl:81-82).
 * Deterministic tree in Float64 defined on fixed chunks of the element index (csrc/reduce.cu): independent of the SM
 * count, of the operands' alignment and -- with cts_set_reduction_layout -- of the number of ranks.  These synchronise
 * the stream (host scalar). */
int cts_dot(cts_ctx* ctx, cts_dtype ft, size_t n, const void* x, const void* y, double* host_out);
int cts_nrm2(cts_ctx* ctx, cts_dtype ft, size_t n, const void* x, double* host_out);
int cts_sum1pabs(cts_ctx* ctx, cts_dtype ft, size_t n, const void* x, double* host_out);
/* k dots against one vector in one pass + one all-reduce: out[j] = <V[j], q> (classical Gram-Schmidt sweep) */
int cts_multi_dot(cts_ctx* ctx, cts_dtype ft, size_t n, int k, const void* const* V /*host array*/, const void* q, double* host_out);
/* Producer + reduction in one pass, bit-identical to the two separate calls (same reduction tree):
 *   cts_axpby_dot:       y = a*x + b*y, then *host_out = z ? dot(y, z) : sum(y_i^2)     (MGS update + next dot / norm)
 *   cts_div_scalar_nrm2: y = x/denom,   then *host_norm = norm(y)                        (basis vector + its norm for jvp!) */
int cts_axpby_dot(cts_ctx* ctx, cts_dtype ft, size_t n, double a, const void* x, double b, void* y, const void* z_or_null,
                  double* host_out);
int cts_div_scalar_nrm2(cts_ctx* ctx, cts_dtype ft, size_t n, void* y, const void* x, double denom, double* host_norm);
/* restrict reductions to the owned (non-ghost) element range [offset, offset+count) of each vector */
int cts_set_reduction_window(cts_ctx* ctx, size_t offset, size_t count);
/* Shard-invariant reductions of a distributed vector: this rank's window [local_offset, local_offset+count) is the slice
 * [global_offset, global_offset+count) of a global index space of `global_count` elements, cut into chunks of `chunk`
 * elements (a power of two, 0 = default 8192; every rank must own whole chunks).  The per-chunk partial sums are
 * exchanged and folded in global chunk order on every rank, so dot/nrm2/sum1pabs -- and everything downstream of them
 * (JVP step sizes, GMRES, convergence checks) -- are bit-identical for 1, 2, 4 or 8 ranks.  `length(x)` in
 * ForwardDiffStepSize3 / GMRES itmax (newtons_method.jl:135,445) becomes `global_count`. */
int cts_set_reduction_layout(cts_ctx* ctx, size_t local_offset, size_t count, size_t global_offset, size_t global_count,
                             size_t chunk);
int cts_reduction_global_length(cts_ctx* ctx, size_t n, size_t* out);

/* ------------------------------------------------------------------ batched tridiagonal (K4) -- */
/* The reference has no tridiagonal type of its own; `ldiv!(Δx, W, f)` (newtons_method.jl:633) dispatches on
 * the user's jac_prototype.  This is that `ldiv!` for a column-stacked state (level index fastest):
 * ncols independent nlev x nlev systems, two-sided (twisted) Thomas elimination, no pivoting
 * (W = dtγ J - I is strictly diagonally dominant for diffusion).  dl[0] and du[nlev-1] of each column
 * are ignored.  `x` may alias `rhs`. */
typedef struct {
  int nlev;
  size_t ncols;
  void* dl; /* sub-diagonal,   n = ncols*nlev */
  void* d;  /* diagonal                        */
  void* du; /* super-diagonal                  */
} cts_tridiag;
int cts_tridiag_solve_batched(cts_ctx* ctx, cts_dtype ft, const cts_tridiag* W, const void* rhs, void* x);
/* `mul!(y, W, x)` (needed by GMRES without a JVP, newtons_method.jl:481) */
int cts_tridiag_matvec(cts_ctx* ctx, cts_dtype ft, const cts_tridiag* W, const void* x, void* y);

/* ------------------------------------------------------------------ user hooks ---------------- */
/* Function-pointer forms of the ClimaODEFunction hooks (src/functions.jl:93-129).  `ud` carries the
 * parameter object `p`.  All pointers are device pointers into buffers the host allocated.  Return 0. */
typedef int (*cts_tendency_fn)(void* ud, void* du, const void* u, double t);                       /* T_imp!(du,u,p,t) / T_exp!      */
typedef int (*cts_exp_lim_fn)(void* ud, void* du_exp, void* du_lim, const void* u, double t);      /* T_exp_T_lim!                   */
typedef int (*cts_wfact_fn)(void* ud, void* W, const void* u, double dtgamma, double t);           /* Wfact(W,u,p,dtγ,t)            */
typedef int (*cts_ldiv_fn)(void* ud, void* x, const void* W, const void* b);                       /* ldiv!(x, W, b)                 */
typedef int (*cts_state_fn)(void* ud, void* u, double t);                                          /* dss!, cache!, cache_imp!, constrain_state! */
typedef int (*cts_init_imp_fn)(void* ud, void* u, double dtgamma);                                 /* initialize_imp!(u,p,dtγ)       */
typedef int (*cts_lim_fn)(void* ud, void* u, double t, const void* u_ref);                         /* lim!(u,p,t,u_ref)              */
typedef int (*cts_incr_fn)(void* ud, void* du, const void* u, double t, double alpha, double beta);/* IncrementingODEFunction        */

/* update signals (src/utilities/update_signal_handler.jl:29-86) */
typedef enum {
  CTS_SIG_NEW_TIMESTEP = 0, CTS_SIG_NEW_NEWTON_SOLVE = 1, CTS_SIG_NEW_NEWTON_ITERATION = 2,
  CTS_SIG_WITH_DSS = 3, CTS_SIG_END_OF_STAGE = 4, CTS_SIG_END_OF_STEP = 5
} cts_signal;
/* `needs_update!(handler, signal)`; fn == NULL means UpdateEvery(every_signal) (subtype rule :96-99) */
typedef struct {
  int every_signal;                                     /* cts_signal for the built-in UpdateEvery          */
  int (*fn)(void* ud, int signal, double t);            /* optional host handler (UpdateEveryN/UpdateEveryDt) */
  void* ud;
} cts_update_handler;

typedef struct {
  void* ud;                       /* passed to every hook (the `p` of the reference)                   */
  cts_exp_lim_fn T_exp_T_lim;     /* NULL == nothing                                                    */
  int has_lim;                    /* functions.jl:150 (_has_lim)                                        */
  cts_tendency_fn T_imp;          /* NULL == nothing                                                    */
  cts_wfact_fn Wfact;             /* NULL == no Jacobian (has_jac false, imex_ark.jl:3-7)               */
  cts_ldiv_fn ldiv;               /* ldiv! of the jac_prototype type                                    */
  cts_ldiv_fn jac_mul;            /* mul!(y, W, x) (only for GMRES without JVP), may be NULL            */
  void* W;                        /* `zero(jac_prototype)` owned by the Newton cache (newtons_method.jl:596) */
  cts_lim_fn lim;                 /* NULL == Returns(nothing)                                           */
  cts_state_fn dss;               /* NULL == Returns(nothing)                                           */
  cts_state_fn constrain_state;
  cts_state_fn cache;
  cts_state_fn cache_imp;
  cts_init_imp_fn initialize_imp;
  int initialize_imp_is_nothing;  /* `initialize_imp! = nothing` (imex_ark.jl:211) vs the default Returns(nothing) */
  cts_update_handler update_cache;            /* default UpdateEvery(EndOfStage)  functions.jl:104 */
  cts_update_handler update_constrain_state;  /* default UpdateEvery(EndOfStep)   functions.jl:105 */
} cts_ode_function;

/* IMEX tableau as runtime data (imex_tableaus.jl:35-79); row-major s x s matrices. */
typedef struct {
  int s;
  double a_exp[CTS_MAX_STAGES * CTS_MAX_STAGES];
  double b_exp[CTS_MAX_STAGES];
  double c_exp[CTS_MAX_STAGES];
  double a_imp[CTS_MAX_STAGES * CTS_MAX_STAGES];
  double b_imp[CTS_MAX_STAGES];
  double c_imp[CTS_MAX_STAGES];
  int cast_to_state_eltype;  /* cast_tableau_to_state_eltype (imex_tableaus.jl:98-105) */
} cts_imex_tableau;

/* NewtonsMethod / KrylovMethod configuration (newtons_method.jl:419-438, 568-581) */
typedef enum { CTS_FORCING_CONSTANT = 0, CTS_FORCING_EISENSTAT_WALKER = 1 } cts_forcing_kind;
typedef struct {
  int present;                /* 0: newtons_method === nothing (ExplicitAlgorithm)              */
  int max_iters;
  cts_update_handler update_j;
  int use_krylov;             /* krylov_method !== nothing                                       */
  int jacobian_free_jvp;      /* ForwardDiffJVP                                                  */
  int jvp_step_kind;          /* ForwardDiffStepSize 1, 2 or 3 (default 3)                       */
  double jvp_step_adjustment;
  int disable_preconditioner;
  int memory;                 /* GMRES memory (default 20); vectors beyond it are appended       */
  int itmax;                  /* 0 -> 2n                                                         */
  int forcing_kind;
  double forcing_rtol;        /* ConstantForcing(rtol)                                           */
  double ew_initial_rtol, ew_gamma, ew_alpha, ew_min_rtol_threshold, ew_max_rtol;
  int batch_gram_schmidt;     /* 0: modified GS (Krylov.jl order); 1: one fused sweep (classical GS) */
} cts_newton_config;

/* `wfact_evals` counts the points at which the reference calls `Wfact` (async_utils.jl:25, newtons_method.jl:620,626), also
 * when the work was folded into a stage kernel or done in-kernel by a matrix-free operator (no separate launch shows up in
 * the profile then).  In a Newton-Krylov solve `ldiv_calls`/`jvp_evals`/`f_evals` count the preconditioner solves and
 * operator evaluations whether or not they shared a tile kernel; a stage that ran on the fused column fast path adds one to
 * `fused_stage_launches` and `newton_iters` and nothing to `f_evals`/`ldiv_calls`. */
typedef struct {
  uint64_t newton_iters, krylov_iters, f_evals, jvp_evals, wfact_evals, ldiv_calls, dss_calls;
  uint64_t fused_stage_launches;  /* stages that ran on the fused column fast path */
} cts_solver_stats;

/* ------------------------------------------------------------------ solvers ------------------- */
typedef struct cts_imex_cache cts_imex_cache;
typedef struct cts_lsrk_cache cts_lsrk_cache;

/* `init_cache(prob, alg::IMEXAlgorithm{Unconstrained})` (imex_ark.jl:35-66) and `{SSP}` (imex_ssprk.jl:33-93).
 * Buffers are allocated zero-filled on the device (`zero(u0)`); counts follow test/unit/cache_sizes.jl.
 * constraint: 0 = Unconstrained (IMEXARKCache), 1 = SSP (IMEXSSPRKCache).  Errors mirror the reference:
 * non-low-storage tableau with SSP -> CTS_EINVAL; non-SDIRK tableau with a NewTimeStep Jacobian update -> CTS_EINVAL. */
int cts_imex_cache_create(cts_ctx* ctx, cts_dtype ft, size_t n, const cts_imex_tableau* tab, int constraint,
                          const cts_ode_function* f, const cts_newton_config* newton, cts_imex_cache** out);
int cts_imex_cache_destroy(cts_imex_cache* c);
/* `step_u!(integrator, cache)` (imex_ark.jl:69-104 / imex_ssprk.jl:95-139): advances `u` in place by dt.
 * `dt_is_f32` != 0 tells the library the integrator's dt has Float32 type (promotion rules). */
int cts_imex_step(cts_imex_cache* c, void* u, double t, double dt, int dt_is_f32);
/* named buffer access for hooks/tests: "U","temp","T_exp<i>","T_lim<i>","T_imp<i>","U_exp","U_lim","T_exp","T_lim","dx","f" (1-based i) */
int cts_imex_cache_buffer(cts_imex_cache* c, const char* name, void** dptr);
int cts_imex_cache_nbuffers(cts_imex_cache* c, int* n_state_sized);
int cts_imex_cache_stats(cts_imex_cache* c, cts_solver_stats* out);
/* 1: let the engine use fused column kernels when it recognises library operators (default 1) */
int cts_imex_cache_set_fusion(cts_imex_cache* c, int enabled);

/* 1 (default unless the environment variable CTS_NO_GRAPH is set): a step whose hooks are all library operators (the fused
 * column path) is captured into a CUDA graph every call, applied to the instantiated graph with cudaGraphExecUpdate and
 * launched as one graph; needs a capturable (non-default) stream, otherwise plain launches are used. */
int cts_imex_cache_set_graph(cts_imex_cache* c, int enabled);
int cts_imex_cache_graph_stats(cts_imex_cache* c, uint64_t* graph_launches, uint64_t* instantiations);

/* `init_cache(prob, alg::LowStorageRungeKutta2N)` (lsrk.jl:49-52) + `step_u!` (lsrk.jl:56-67).
 * A, B, C are given in Float64 holding values already rounded to the state eltype (lsrk.jl:51). */
int cts_lsrk_cache_create(cts_ctx* ctx, cts_dtype ft, size_t n, int nstages, const double* A, const double* B,
                          const double* C, cts_incr_fn f, void* ud, cts_lsrk_cache** out);
int cts_lsrk_cache_destroy(cts_lsrk_cache* c);
int cts_lsrk_step(cts_lsrk_cache* c, void* u, double t, double dt, int dt_is_f32);
int cts_lsrk_cache_buffer(cts_lsrk_cache* c, const char* name, void** dptr); /* "du" */
int cts_lsrk_cache_set_fusion(cts_lsrk_cache* c, int enabled);

/* ---- RosenbrockAlgorithm (src/solvers/rosenbrock.jl) -------------------------------------------
 * Transformed tableau as runtime data (rosenbrock.jl:25-46): row-major s x s matrices A = αΓ⁻¹, α, C = diag(Γ⁻¹) − Γ⁻¹, Γ
 * and the weights m = bΓ⁻¹ (the host computes them as RosenbrockTableau's constructor does). */
typedef struct {
  int s;
  double A[CTS_MAX_STAGES * CTS_MAX_STAGES];
  double alpha[CTS_MAX_STAGES * CTS_MAX_STAGES];
  double C[CTS_MAX_STAGES * CTS_MAX_STAGES];
  double Gamma[CTS_MAX_STAGES * CTS_MAX_STAGES];
  double m[CTS_MAX_STAGES];
} cts_rosenbrock_tableau;
typedef struct cts_rosenbrock_cache cts_rosenbrock_cache;
/* `init_cache(prob, alg::RosenbrockAlgorithm)` (rosenbrock.jl:100-124): U, fU, fU_imp, fU_exp, fU_lim, ∂Y∂t and k[1..s],
 * zero-filled; `f->W` is the user's jac_prototype itself (the reference does not copy it here).  `tgrad` is
 * `T_imp!.tgrad` (may be NULL).  `step_u!` (rosenbrock.jl:136-239): Wfact once per step with dtγ = dt·Γ[1,1], one
 * `ldiv!` per stage, `dss!` on every stage value but the first and on the final state. */
int cts_rosenbrock_cache_create(cts_ctx* ctx, cts_dtype ft, size_t n, const cts_rosenbrock_tableau* tab,
                                const cts_ode_function* f, cts_tendency_fn tgrad, cts_rosenbrock_cache** out);
int cts_rosenbrock_cache_destroy(cts_rosenbrock_cache* c);
int cts_rosenbrock_step(cts_rosenbrock_cache* c, void* u, double t, double dt);
/* "U","fU","fU_imp","fU_exp","fU_lim","dYdt","k<i>" (1-based i) */
int cts_rosenbrock_cache_buffer(cts_rosenbrock_cache* c, const char* name, void** dptr);
int cts_rosenbrock_cache_stats(cts_rosenbrock_cache* c, cts_solver_stats* out);
/* Developer switch: 0 = always the reference's pass structure (T_imp!, `.+=` chain, ldiv! as separate passes); default 1 =
 * one tile kernel per stage when T_imp!/Wfact/ldiv! are the library's column operator (rosenbrock.jl:164-224). */
int cts_rosenbrock_cache_set_fusion(cts_rosenbrock_cache* c, int enabled);
int cts_lsrk_cache_stats(cts_lsrk_cache* c, cts_solver_stats* out);

/* Newton's method on its own (`solve_newton!`, newtons_method.jl:609-665) for unit tests of the solver:
 * f!(f, x) and j!(j, x) are host callbacks; x is updated in place. */
typedef int (*cts_residual_fn)(void* ud, void* f, const void* x);
typedef int (*cts_jac_fn)(void* ud, void* W, const void* x);
typedef int (*cts_prepare_fn)(void* ud, void* x);
typedef struct cts_newton_cache cts_newton_cache;
int cts_newton_cache_create(cts_ctx* ctx, cts_dtype ft, size_t n, const cts_newton_config* cfg, void* W,
                            cts_ldiv_fn ldiv, cts_ldiv_fn jac_mul, void* ud, cts_newton_cache** out);
int cts_newton_cache_destroy(cts_newton_cache* c);
int cts_newton_solve(cts_newton_cache* c, void* x, cts_residual_fn f, cts_jac_fn j, cts_prepare_fn prepare, void* ud);
int cts_newton_cache_stats(cts_newton_cache* c, cts_solver_stats* out);

/* ConvergenceChecker (src/utilities/convergence_checker.jl:32-108) with ConvergenceCondition trees
 * (convergence_condition.jl:41-143) in postfix order: leaves first, then the MultipleConditions node that combines the
 * `nchildren` results below it.  p0 = max_err | max_rel_err | max_reduction | rate; p1 = order (MinimumRateOfConvergence).
 * `norm` is LinearAlgebra.norm (the deterministic device reduction, all-reduced over ranks); custom norms are a host-side
 * concern of the binding.  Condition caches start as NaN (the reference leaves them uninitialised), so a condition that
 * reads its cache before the first update is false. */
typedef enum {
  CTS_COND_MAX_ERROR = 0, CTS_COND_MAX_REL_ERROR = 1, CTS_COND_MAX_ERROR_REDUCTION = 2, CTS_COND_MIN_RATE = 3,
  CTS_COND_ALL = 4, CTS_COND_ANY = 5
} cts_cond_kind;
typedef struct { int kind; int nchildren; double p0, p1; } cts_cond_node;
#define CTS_MAX_COND_NODES 16
typedef struct {
  int n_norm;                                   /* 0: norm_condition === nothing      */
  cts_cond_node norm_cond[CTS_MAX_COND_NODES];
  int n_comp;                                   /* 0: component_condition === nothing */
  cts_cond_node comp_cond[CTS_MAX_COND_NODES];
  int combiner_or;                              /* condition_combiner: 0 = &, 1 = |   */
} cts_convergence_checker;
/* `NewtonsMethod(; convergence_checker, line_search)` (newtons_method.jl:568-581): NULL / 0 = nothing.  Allocates the
 * checker cache (`allocate_cache`, convergence_checker.jl:45-60).  A Newton solve with either of them set always runs
 * the generic iteration (no fused column stage). */
int cts_newton_cache_set_convergence_checker(cts_newton_cache* c, const cts_convergence_checker* checker);
int cts_newton_cache_set_line_search(cts_newton_cache* c, int enabled);   /* LineSearch(), line_search.jl:23-47 */
/* `is_converged!(checker, cache, val, err, iter)` on its own (test/utils/convergence_checker.jl) */
int cts_newton_cache_is_converged(cts_newton_cache* c, const void* val, const void* err, int iter, int* converged);
/* the Newton cache embedded in an IMEX cache (NULL when the algorithm has no NewtonsMethod) */
int cts_imex_cache_newton(cts_imex_cache* c, cts_newton_cache** out);
/* all(component_condition) over |val|, |err| (convergence_checker.jl:62-73) and the cache update (:97-105);
 * caches[k] is the per-element cache of node k (NULL for nodes without one) */
int cts_component_check(cts_ctx* ctx, cts_dtype ft, size_t n, const void* val, const void* err, int nnodes,
                        const cts_cond_node* nodes, void* const* caches, int iter, double* dev_scratch, int* all_converged);
int cts_component_update(cts_ctx* ctx, cts_dtype ft, size_t n, const void* val, const void* err, int nnodes,
                         const cts_cond_node* nodes, void* const* caches, int iter);

/* ------------------------------------------------------------------ library-native operators -- */
/* These are the "user model" pieces (layer L5 of the reference) of the synthetic benchmark problems of
 * SURVEY §8d, provided as device code so the whole step stays on the GPU.  Each exposes hook functions
 * with the typedefs above (pass the operator handle as `ud`).  When the solver recognises them it may use
 * fused kernels that produce bit-identical results to calling the hooks one by one. */
typedef struct cts_op_advection cts_op_advection;       /* C2: 1-D periodic upwind advection, du = α·rhs(u) + β·du */
int cts_op_advection_create(cts_ctx* ctx, cts_dtype ft, size_t n, double c, double dx, cts_op_advection** out);
int cts_op_advection_destroy(cts_op_advection* op);
int cts_op_advection_incr(void* op, void* du, const void* u, double t, double alpha, double beta); /* cts_incr_fn */

/* C3/C4/C5 column model: state is [rows][nx][nlev] (level fastest); rows = ny_local + 2*halo (halo = 0 or 1).
 *   T_imp(U)  = K_col d2U/dz2 (3-point, Dirichlet 0)                       linear   (nonlinear = 0)
 *             = d/dz( K_col (1 + Ubar^2) dU/dz )                           nonlinear (nonlinear = 1)
 *   Wfact     : W = dtγ J - I as a batched tridiagonal (frozen coefficients in the nonlinear case)
 *   T_exp(U)  = S[lev]*(1 + 0.5 sin(2π t)) + nu * (5-point horizontal Laplacian over (nx, rows)), owned rows only
 *   dss!(U)   : refresh the ghost rows from their owners (local periodic copy on 1 rank, NCCL send/recv otherwise) */
typedef struct cts_op_column cts_op_column;
typedef struct {
  int nlev;
  size_t nx;        /* columns per horizontal row (fast horizontal axis, periodic)             */
  size_t ny_local;  /* owned rows on this rank                                                  */
  int halo;         /* 0: no ghost rows (no horizontal coupling possible), 1: one ghost row each side */
  double dz;
  double nu;        /* horizontal diffusivity (0 disables the stencil)                          */
  int nonlinear;
  const void* K_col;  /* device, one value per local column incl. ghosts: (ny_local+2*halo)*nx   */
  const void* S_lev;  /* device, nlev source profile values (may be NULL -> no source)          */
} cts_column_desc;
int cts_op_column_create(cts_ctx* ctx, cts_dtype ft, const cts_column_desc* desc, cts_op_column** out);
int cts_op_column_destroy(cts_op_column* op);
size_t cts_op_column_ndof(const cts_op_column* op);   /* local DOFs incl. ghosts */
int cts_op_column_make_W(cts_op_column* op, cts_tridiag** W);  /* zero(jac_prototype): allocates dl,d,du */
int cts_op_column_free_W(cts_op_column* op, cts_tridiag* W);
int cts_op_column_T_imp(void* op, void* du, const void* u, double t);                         /* cts_tendency_fn */
int cts_op_column_T_exp(void* op, void* du_exp, void* du_lim, const void* u, double t);       /* cts_exp_lim_fn  */
int cts_op_column_Wfact(void* op, void* W, const void* u, double dtgamma, double t);          /* cts_wfact_fn    */
int cts_op_column_ldiv(void* op, void* x, const void* W, const void* b);                      /* cts_ldiv_fn     */
int cts_op_column_jac_mul(void* op, void* y, const void* W, const void* x);                   /* cts_ldiv_fn     */
int cts_op_column_dss(void* op, void* u, double t);                                           /* cts_state_fn    */
/* 1: kernels that know the operator regenerate W = dtγJ - I instead of reading dl,d,du: the fused implicit stage (linear
 * operator, from K_col) and the Newton-Krylov preconditioner solves (linear or nonlinear operator, frozen coefficients at the
 * Newton iterate; used when update_j = UpdateEvery(NewNewtonIteration), where the stored W would be the same bits and its
 * Wfact pass is skipped).  Results are bit-identical to the stored-W run. */
int cts_op_column_set_matrix_free(cts_op_column* op, int enabled);

/* ------------------------------------------------------------------ multi-GPU ---------------- */
/* One process per GPU.  The unique id is created on rank 0 and distributed by the host runtime
 * (MPI / torch.distributed / ClimaComms); NCCL is dlopen'ed on first use. */
#define CTS_NCCL_UNIQUE_ID_BYTES 128
int cts_comm_get_unique_id(void* out_128_bytes);
int cts_comm_init(cts_ctx* ctx, int nranks, int rank, const void* unique_id_128_bytes);
int cts_comm_destroy(cts_ctx* ctx);
int cts_comm_rank(cts_ctx* ctx, int* rank, int* nranks);
/* halo exchange of `count` elements: send `send_lo` to rank-1 (their `recv_hi`), `send_hi` to rank+1 (periodic ring) */
int cts_halo_exchange(cts_ctx* ctx, cts_dtype ft, size_t count, const void* send_lo, const void* send_hi,
                      void* recv_lo, void* recv_hi);
/* Collective over the communicator: sets up the NVLink peer-memory path of the halo exchange for rows of up to
 * `max_row_bytes` (CUDA IPC mailboxes mapped by the two ring neighbours; one kernel per exchange pushes the boundary rows
 * with 16-byte stores over NVLink and signals with system-scope atomics -- no NCCL call on the data path).  If any rank
 * cannot map its neighbours every rank keeps the NCCL send/recv path.  cts_comm_halo_path: 0 = single rank (local
 * copies), 1 = NCCL send/recv, 2 = peer-memory kernel. */
int cts_comm_enable_peer_halo(cts_ctx* ctx, size_t max_row_bytes);
int cts_comm_halo_path(cts_ctx* ctx, int* path);
/* in-stream sum all-reduce of `count` doubles on the device */
int cts_allreduce_sum_f64(cts_ctx* ctx, double* dev_buf, size_t count);
int cts_allreduce_max_f64_host(cts_ctx* ctx, double* host_val); /* EveryXWallTimeSeconds' allreduce(max), Callbacks.jl:59,72 */

/* ------------------------------------------------------------------ synthetic inputs --------- */
/* counter-based hash of the GLOBAL index (splitmix64(seed ^ idx) -> uniform [0,1)), SURVEY §8d: identical data
 * on CPU, 1 GPU and any sharding.  out[i] = scale * (hash(seed, offset + i) + shift)  */
int cts_fill_hash(cts_ctx* ctx, cts_dtype ft, size_t n, void* out, uint64_t seed, uint64_t global_offset, double scale, double shift);

#ifdef __cplusplus
}
#endif
#endif /* CTS_H */
