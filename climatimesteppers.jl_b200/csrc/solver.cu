// Host-side step engine: `step_u!` for IMEX-ARK (imex_ark.jl:69-308), IMEX-SSPRK (imex_ssprk.jl:95-220) and
// LSRK-2N (lsrk.jl:56-67), `solve_newton!` / `solve_krylov!` (newtons_method.jl:465-665) and GMRES
// (Krylov.jl v0.10 `gmres!`, restated from its published algorithm).
//
// The engine walks the stages exactly where the reference does, calling the user's hooks through C
// function pointers between library kernels.  All state arithmetic happens in the kernels of
// elementwise.cu / reduce.cu / tridiag.cu / ops.cu; this file only sequences them (O(1) host work per
// call, everything stream-ordered; only reductions feeding host control flow synchronise).
//
// Fast paths: when the hooks ARE the library's own operators (compared by function address) the engine
// substitutes fused kernels that are bit-identical to the hook sequence:
//   - IMEX stage with the linear column operator, direct Newton, max_iters = 1  -> cts_column_fused_stage
//   - LSRK stage with the advection operator                                     -> cts_advection_fused_stage
#include <algorithm>
#include <cmath>
#include <cstring>
#include <functional>
#include <limits>

#include "cts_internal.h"

namespace {

// ---- update-signal handlers (update_signal_handler.jl:88-99) -----------------------------------
bool signal_is_a(int signal, int target) {
  if (signal == target) return true;
  // EndOfStep <: EndOfStage <: WithDSS
  if (signal == CTS_SIG_END_OF_STEP) return target == CTS_SIG_END_OF_STAGE || target == CTS_SIG_WITH_DSS;
  if (signal == CTS_SIG_END_OF_STAGE) return target == CTS_SIG_WITH_DSS;
  return false;
}
bool needs_update(const cts_update_handler& h, int signal, double t = 0.0) {
  if (h.fn) return h.fn(h.ud, signal, t) != 0;
  return signal_is_a(signal, h.every_signal);
}
bool fires_on_new_timestep(const cts_update_handler& h) {
  // async_utils.jl:85-88: a host handler (UpdateEveryN/UpdateEveryDt) declares its signal in every_signal too
  return h.every_signal == CTS_SIG_NEW_TIMESTEP;
}

#define HOOK(ctx, expr)                                                        \
  do {                                                                         \
    int _r = (expr);                                                           \
    if (_r != 0) return cts_fail((ctx), CTS_ECALLBACK, std::string("hook failed: ") + #expr + " -> " + std::to_string(_r)); \
  } while (0)

}  // namespace

// =============================================================================================
// Newton / Krylov
// =============================================================================================
struct cts_newton_cache {
  cts_ctx* ctx = nullptr;
  cts_dtype ft = CTS_F64;
  size_t n = 0;
  cts_newton_config cfg{};
  void* dx = nullptr;
  void* f = nullptr;
  void* W = nullptr;
  cts_ldiv_fn ldiv = nullptr;
  cts_ldiv_fn jac_mul = nullptr;
  void* ud = nullptr;
  // Krylov workspace
  void* x2 = nullptr;
  void* f2 = nullptr;
  void* w = nullptr;
  void* q = nullptr;
  std::vector<void*> V;
  double ew_prev_norm_f = 0.0, ew_prev_rtol = 0.0;
  cts_solver_stats stats{};
  // Set by solve_implicit_equation! while the stage equation's T_imp! is the library's column operator and no hook
  // runs before f!: the residual (and the JVP built on it) is then one fused kernel instead of T_imp! + K3 (+ K7, K8).
  cts_op_column* fused_op = nullptr;
  const void* fused_U0 = nullptr;
  double fused_dtg = 0.0;
  // ConvergenceChecker (convergence_checker.jl) and LineSearch (line_search.jl)
  bool has_checker = false;
  cts_convergence_checker checker{};
  double norm_cache[CTS_MAX_COND_NODES] = {};   // per-node scalar caches of the norm condition
  void* comp_cache[CTS_MAX_COND_NODES] = {};    // per-node element caches of the component condition
  double* dev_scratch = nullptr;
  bool line_search = false;
  // values GMRES / jvp! would otherwise recompute (bit-identical reuse)
  const void* known_norm_ptr = nullptr;  // a basis vector whose norm came out of the kernel that produced it
  double known_norm = 0.0;
  bool fuse_reductions = true;           // cleared together with the engine's fusion switch (reference pass structure)
  bool fused_wfact_is_lib = false;       // the Wfact hook of the current solve is the library column operator's
  bool xstat_valid = false;              // sum(1+|x|) (or norm(x)) of the current solve_krylov! call
  double xstat = 0.0;
  // one-iteration direct solve whose `x .-= Δx` (K5) the caller fuses with what follows (K6): solve_newton leaves Δx
  // in `dx` and x untouched
  bool defer_update = false;
};

namespace {

using ResidualFn = std::function<int(void* f, const void* x)>;
using JacFn = std::function<int(void* W, const void* x)>;
using PrepareFn = std::function<int(void* x)>;

int newton_alloc(cts_newton_cache* c) {
  size_t bytes = c->n * cts_sizeof(c->ft);
  CTS_TRY(cts_alloc(c->ctx, bytes, &c->dx));
  CTS_TRY(cts_alloc(c->ctx, bytes, &c->f));
  if (c->cfg.use_krylov) {
    CTS_TRY(cts_alloc(c->ctx, bytes, &c->w));
    CTS_TRY(cts_alloc(c->ctx, bytes, &c->q));
    if (c->cfg.jacobian_free_jvp) {
      CTS_TRY(cts_alloc(c->ctx, bytes, &c->x2));
      CTS_TRY(cts_alloc(c->ctx, bytes, &c->f2));
    }
    int mem = c->cfg.memory > 0 ? c->cfg.memory : 20;
    for (int i = 0; i < mem; ++i) {
      void* v = nullptr;
      CTS_TRY(cts_alloc(c->ctx, bytes, &v));
      c->V.push_back(v);
    }
  }
  return CTS_OK;
}

void newton_free(cts_newton_cache* c) {
  for (void* p : {c->dx, c->f, c->x2, c->f2, c->w, c->q})
    if (p) cts_free(c->ctx, p);
  for (void* v : c->V) cts_free(c->ctx, v);
  c->V.clear();
  for (int k = 0; k < CTS_MAX_COND_NODES; ++k)
    if (c->comp_cache[k]) cts_free(c->ctx, c->comp_cache[k]);
  if (c->dev_scratch) cts_free(c->ctx, c->dev_scratch);
}

double eps_of(cts_dtype ft) { return ft == CTS_F64 ? std::numeric_limits<double>::epsilon() : (double)std::numeric_limits<float>::epsilon(); }
double round_to(cts_dtype ft, double v) { return ft == CTS_F64 ? v : (double)(float)v; }

size_t global_length(const cts_ctx* ctx, size_t n) { return cts_global_length(ctx, n); }

// Krylov.jl sym_givens (real case): [c s; s -c][a; b] = [rho; 0]
void sym_givens(double a, double b, double& c, double& s, double& rho) {
  if (b == 0) {
    c = (a == 0) ? 1.0 : std::copysign(1.0, a);
    s = 0.0;
    rho = std::fabs(a);
  } else if (a == 0) {
    c = 0.0;
    s = std::copysign(1.0, b);
    rho = std::fabs(b);
  } else if (std::fabs(b) > std::fabs(a)) {
    double t = a / b;
    s = std::copysign(1.0, b) / std::sqrt(1 + t * t);
    c = s * t;
    rho = b / s;
  } else {
    double t = b / a;
    c = std::copysign(1.0, a) / std::sqrt(1 + t * t);
    s = c * t;
    rho = a / c;
  }
}

// ForwardDiffJVP step (newtons_method.jl:107-135,177)
// (norm(Δx) is reused when GMRES already produced it together with the basis vector; sum(1+|x|) / norm(x) are the same
//  for every JVP of one solve_krylov! call -- x does not change -- and are computed once per call)
int jvp_step(cts_newton_cache* c, const void* dx, const void* x, double* eps_out) {
  cts_ctx* ctx = c->ctx;
  double e = 0, ndx = 0;
  const double se = std::sqrt(eps_of(c->ft));
  if (c->known_norm_ptr == dx && dx)
    ndx = c->known_norm;
  else
    CTS_TRY(cts_nrm2(ctx, c->ft, c->n, dx, &ndx));
  if (c->cfg.jvp_step_kind == 1) {
    e = se / ndx;
  } else if (c->cfg.jvp_step_kind == 2) {
    if (!c->xstat_valid || !c->fuse_reductions) {
      CTS_TRY(cts_nrm2(ctx, c->ft, c->n, x, &c->xstat));
      c->xstat_valid = true;
    }
    e = std::sqrt(eps_of(c->ft) * (1 + c->xstat)) / ndx;
  } else {
    if (!c->xstat_valid || !c->fuse_reductions) {
      CTS_TRY(cts_sum1pabs(ctx, c->ft, c->n, x, &c->xstat));
      c->xstat_valid = true;
    }
    e = se * c->xstat / ((double)global_length(ctx, c->n) * ndx);
  }
  double adj = c->cfg.jvp_step_adjustment == 0.0 ? 1.0 : c->cfg.jvp_step_adjustment;
  *eps_out = round_to(c->ft, round_to(c->ft, adj) * round_to(c->ft, e));
  return CTS_OK;
}

int get_rtol(cts_newton_cache* c, const void* f, int n_iter, double* rtol_out) {
  const cts_newton_config& k = c->cfg;
  if (k.forcing_kind == CTS_FORCING_CONSTANT) {
    *rtol_out = round_to(c->ft, k.forcing_rtol);
    return CTS_OK;
  }
  double norm_f = 0;
  CTS_TRY(cts_nrm2(c->ctx, c->ft, c->n, f, &norm_f));
  double rtol;
  if (n_iter == 0) {
    rtol = k.ew_initial_rtol;
  } else {
    rtol = k.ew_gamma * std::pow(norm_f / c->ew_prev_norm_f, k.ew_alpha);
    double min_rtol = k.ew_gamma * std::pow(c->ew_prev_rtol, k.ew_alpha);
    if (min_rtol > k.ew_min_rtol_threshold) rtol = std::max(rtol, min_rtol);
  }
  rtol = std::min(rtol, k.ew_max_rtol);
  c->ew_prev_norm_f = norm_f;
  c->ew_prev_rtol = rtol;
  *rtol_out = rtol;
  return CTS_OK;
}

// GMRES (left preconditioned, no restart; vectors beyond `memory` are appended)
// `matvec_precond` (optional): q = M^-1 (A v) in one pass (the column operator's JVP + tridiagonal solve); same roundings
// as matvec followed by precond, `w` is then never written.
int gmres(cts_newton_cache* c, const std::function<int(void* out, const void* v)>& matvec,
          const std::function<int(void* out, const void* w)>* precond, const void* b, void* x, double rtol,
          const std::function<int(void* out, const void* v)>* matvec_precond = nullptr, void* newton_x = nullptr,
          bool keep_dx = true, bool* applied = nullptr) {
  cts_ctx* ctx = c->ctx;
  const cts_dtype ft = c->ft;
  const size_t n = c->n, bytes = n * cts_sizeof(ft);
  void* w = c->w;
  void* q = c->q;
  // r0 = M^-1 b
  const void* r0 = b;
  if (precond) {
    CTS_TRY((*precond)(q, b));
    r0 = q;
  }
  double beta = 0;
  CTS_TRY(cts_nrm2(ctx, ft, n, r0, &beta));
  if (beta == 0) {  // "x = 0 is a zero-residual solution"
    CTS_TRY(cts_memset_zero(ctx, x, bytes));
    return CTS_OK;
  }
  double rNorm = beta;
  const double eps_tol = 0.0 + rtol * beta;
  const double btol = std::pow(eps_of(ft), 0.75);
  const size_t itmax = c->cfg.itmax > 0 ? (size_t)c->cfg.itmax : 2 * global_length(ctx, n);
  auto ensure_vec = [&](size_t k) -> int {
    while (c->V.size() <= k) {
      void* v = nullptr;
      CTS_TRY(cts_alloc(ctx, bytes, &v));
      c->V.push_back(v);
    }
    return CTS_OK;
  };
  CTS_TRY(ensure_vec(0));
  const bool fuse_red = c->fuse_reductions;
  std::vector<double> vnorm(1, 0.0);  // norm(V[k]) for jvp!'s step size, produced together with V[k]
  if (fuse_red && c->cfg.jacobian_free_jvp) {
    cts_prof_scope ps(ctx, CTS_PROF_KRYLOV);
    CTS_TRY(cts_div_scalar_nrm2(ctx, ft, n, c->V[0], r0, rNorm, &vnorm[0]));
  } else {
    cts_prof_scope ps(ctx, CTS_PROF_KRYLOV);
    CTS_TRY(cts_div_scalar(ctx, ft, n, c->V[0], r0, rNorm));
  }
  std::vector<double> z{beta}, cs, ss;
  std::vector<std::vector<double>> R;
  bool solved = rNorm <= eps_tol, breakdown = false;
  size_t k = 0;
  while (!(solved || k >= itmax || breakdown)) {
    ++k;
    c->stats.krylov_iters++;
    if (fuse_red && c->cfg.jacobian_free_jvp) {
      c->known_norm_ptr = c->V[k - 1];
      c->known_norm = vnorm[k - 1];
    }
    if (matvec_precond && precond) {
      CTS_TRY((*matvec_precond)(q, c->V[k - 1]));
      c->known_norm_ptr = nullptr;
    } else {
      CTS_TRY(matvec(w, c->V[k - 1]));
      c->known_norm_ptr = nullptr;
      if (precond)
        CTS_TRY((*precond)(q, w));
      else
        CTS_TRY(cts_memcpy_d2d(ctx, q, w, bytes));
    }
    std::vector<double> col(k);
    double Hbis_fused = 0;
    if (c->cfg.batch_gram_schmidt) {  // classical GS: one fused sweep of k dots, then one k-term update
      std::vector<const void*> Vp(c->V.begin(), c->V.begin() + k);
      CTS_TRY(cts_multi_dot(ctx, ft, n, (int)k, Vp.data(), q, col.data()));
      for (size_t i0 = 0; i0 < k; i0 += CTS_MAX_TERMS) {
        int kk = (int)std::min<size_t>(CTS_MAX_TERMS, k - i0);
        double coef[CTS_MAX_TERMS];
        for (int i = 0; i < kk; ++i) coef[i] = -col[i0 + i];
        CTS_TRY(cts_axpy_n(ctx, ft, n, q, nullptr, 1.0, q, kk, 0, coef, Vp.data() + i0, 0));
      }
    } else if (fuse_red) {  // modified GS, each update fused with the reduction that follows it (next dot, or the norm)
      cts_prof_scope ps(ctx, CTS_PROF_KRYLOV);
      CTS_TRY(cts_dot(ctx, ft, n, c->V[0], q, &col[0]));
      for (size_t i = 0; i < k; ++i) {
        double red = 0;
        CTS_TRY(cts_axpby_dot(ctx, ft, n, -col[i], c->V[i], 1.0, q, i + 1 < k ? c->V[i + 1] : nullptr, &red));
        if (i + 1 < k)
          col[i + 1] = red;
        else
          Hbis_fused = std::sqrt(red);
      }
    } else {  // modified GS (Krylov.jl order): k dependent dots
      cts_prof_scope ps(ctx, CTS_PROF_KRYLOV);
      for (size_t i = 0; i < k; ++i) {
        CTS_TRY(cts_dot(ctx, ft, n, c->V[i], q, &col[i]));
        CTS_TRY(cts_axpby(ctx, ft, n, -col[i], c->V[i], 1.0, q));
      }
    }
    double Hbis = 0;
    if (fuse_red && !c->cfg.batch_gram_schmidt)
      Hbis = Hbis_fused;
    else
      CTS_TRY(cts_nrm2(ctx, ft, n, q, &Hbis));
    for (size_t i = 0; i + 1 < k; ++i) {
      double rtmp = cs[i] * col[i] + ss[i] * col[i + 1];
      col[i + 1] = ss[i] * col[i] - cs[i] * col[i + 1];
      col[i] = rtmp;
    }
    double cg, sg, rho;
    sym_givens(col[k - 1], Hbis, cg, sg, rho);
    col[k - 1] = rho;
    cs.push_back(cg);
    ss.push_back(sg);
    R.push_back(col);
    double zeta = sg * z[k - 1];
    z[k - 1] = cg * z[k - 1];
    rNorm = std::fabs(zeta);
    solved = rNorm <= eps_tol || rNorm + 1 <= 1;
    breakdown = Hbis <= btol;
    if (!(solved || k >= itmax || breakdown)) {
      CTS_TRY(ensure_vec(k));
      vnorm.push_back(0.0);
      cts_prof_scope ps(ctx, CTS_PROF_KRYLOV);
      if (fuse_red && c->cfg.jacobian_free_jvp)
        CTS_TRY(cts_div_scalar_nrm2(ctx, ft, n, c->V[k], q, Hbis, &vnorm[k]));
      else
        CTS_TRY(cts_div_scalar(ctx, ft, n, c->V[k], q, Hbis));
      z.push_back(zeta);
    }
  }
  std::vector<double> y(k, 0.0);
  for (size_t ii = k; ii-- > 0;) {
    double acc = z[ii];
    for (size_t j = ii + 1; j < k; ++j) acc -= R[j][ii] * y[j];
    y[ii] = (std::fabs(R[ii][ii]) <= btol) ? 0.0 : acc / R[ii][ii];
  }
  std::vector<const void*> Vp(c->V.begin(), c->V.begin() + k);
  if (newton_x && applied && k >= 1 && k <= 32) {
    // K9 + K5 (newtons_method.jl:508,651) in one pass: the Newton iterate takes the update straight from the basis; Δx is
    // written only for a convergence checker / line search (same fold, same roundings as cts_lincomb + cts_sub_inplace)
    cts_prof_scope ps(ctx, CTS_PROF_NEWTON_EW);
    CTS_TRY(cts_lincomb_sub(ctx, ft, n, newton_x, keep_dx ? x : nullptr, (int)k, y.data(), Vp.data()));
    *applied = true;
    return CTS_OK;
  }
  CTS_TRY(cts_lincomb(ctx, ft, n, x, (int)k, y.data(), Vp.data(), 0));
  return CTS_OK;
}

// Matrix-free Newton-Krylov on the library column operator (cts_op_column_set_matrix_free): W is never stored; the
// preconditioner solves generate its rows from the Newton iterate inside the tile kernel.  That is the stored-W algorithm
// bit for bit exactly when the reference would have refreshed W at this very x before every solve_krylov! call, i.e. for
// update_j = UpdateEvery(NewNewtonIteration) (the default, newtons_method.jl:571,626).
bool krylov_matrix_free(const cts_newton_cache* c, const void* x, const void* f);

// solve_krylov! (newtons_method.jl:465-509)
int solve_krylov(cts_newton_cache* c, void* dx, const void* x, const ResidualFn& f_, const void* f, int n_iter,
                 const PrepareFn* prepare, void* newton_x = nullptr, bool keep_dx = true, bool* applied = nullptr) {
  cts_nvtx_scope nvtx("solve_krylov!");  // newtons_method.jl:465
  cts_ctx* ctx = c->ctx;
  const bool jfree = c->cfg.jacobian_free_jvp != 0;
  c->xstat_valid = false;  // x is fixed for the whole GMRES solve; with the reference pass structure recompute per JVP
  struct XstatScope {
    cts_newton_cache* c;
    ~XstatScope() { c->xstat_valid = false; }
  } xstat_scope{c};
  std::function<int(void*, const void*)> matvec = [&](void* out, const void* v) -> int {
    if (!jfree) {
      if (!c->jac_mul || !c->W) return cts_fail(ctx, CTS_EINVAL, "GMRES without JVP needs mul!(y, W, x)");
      HOOK(ctx, c->jac_mul(c->ud, out, c->W, v));
      return CTS_OK;
    }
    double e = 0;  // jvp! :173-182
    CTS_TRY(jvp_step(c, v, x, &e));
    if (c->fused_op) {  // K7 + T_imp! + K3 + K8 in one pass (x2 and f2 are never materialised)
      cts_prof_scope ps(ctx, CTS_PROF_NEWTON_EW);
      CTS_TRY(cts_column_fused_residual(c->fused_op, out, x, v, e, c->fused_U0, f, c->fused_dtg));
      c->stats.f_evals++;
      c->stats.jvp_evals++;
      return CTS_OK;
    }
    {
      cts_prof_scope ps(ctx, CTS_PROF_NEWTON_EW);
      CTS_TRY(cts_jvp_perturb(ctx, c->ft, c->n, c->x2, x, e, v));
    }
    if (prepare) CTS_TRY((*prepare)(c->x2));
    CTS_TRY(f_(c->f2, c->x2));
    c->stats.f_evals++;
    c->stats.jvp_evals++;
    cts_prof_scope ps(ctx, CTS_PROF_NEWTON_EW);
    CTS_TRY(cts_jvp_quotient(ctx, c->ft, c->n, out, c->f2, f, e));
    return CTS_OK;
  };
  const bool use_M = !(c->cfg.disable_preconditioner || !c->W || !jfree);
  const bool kmf = use_M && krylov_matrix_free(c, x, f);
  std::function<int(void*, const void*)> precond = [&](void* out, const void* w) -> int {
    cts_prof_scope ps(ctx, CTS_PROF_LDIV);
    if (kmf && cts_aligned16(out) && cts_aligned16(w)) {
      CTS_TRY(cts_column_precond_matrix_free(c->fused_op, out, x, w, c->fused_dtg));
    } else {
      if (kmf) return cts_fail(ctx, CTS_EUNSUPPORTED, "matrix-free Krylov preconditioner needs 16-byte aligned vectors");
      HOOK(ctx, c->ldiv(c->ud, out, c->W, w));
    }
    c->stats.ldiv_calls++;
    return CTS_OK;
  };
  // library column operator + its batched tridiagonal W as preconditioner: JVP and ldiv! of a Krylov iteration run as one
  // tile kernel (8w instead of 5w + 5w, bit-identical).  Default: on for Float64 (since the tile kernels hold six CTAs per
  // SM it beats the two full-occupancy passes: C4 at 2^27 DOF 39.5 vs 40.4 ms/step), off for Float32 (23.7 vs 23.1 ms:
  // twice the elements per byte through the nonlinear operator's Float64 arithmetic); CTS_JVP_PRECOND_FUSION=0/1 overrides
  // (read per solve).
  const char* mvp_env = getenv("CTS_JVP_PRECOND_FUSION");
  const bool mvp_on = mvp_env ? mvp_env[0] == '1' : c->ft == CTS_F64;
  const bool fuse_mvp = kmf || (mvp_on && use_M && c->fused_op && c->ldiv == cts_op_column_ldiv && c->ud == (void*)c->fused_op &&
                        cts_column_rhs_solve_fusable(c->fused_op) && cts_aligned16(x) && cts_aligned16(f) && cts_aligned16(c->q));
  std::function<int(void*, const void*)> matvec_precond = [&](void* out, const void* v) -> int {
    double e = 0;
    CTS_TRY(jvp_step(c, v, x, &e));
    cts_prof_scope ps(ctx, CTS_PROF_LDIV);
    int r = cts_column_jvp_precond(c->fused_op, out, x, v, e, c->fused_U0, f, c->fused_dtg, kmf ? nullptr : static_cast<const cts_tridiag*>(c->W));
    if (r == CTS_EUNSUPPORTED && kmf) return cts_fail(ctx, CTS_EUNSUPPORTED, "matrix-free Krylov matvec: unsupported operand");
    if (r == CTS_EUNSUPPORTED) {  // e.g. a W of another shape: the two separate passes
      CTS_TRY(cts_column_fused_residual(c->fused_op, c->w, x, v, e, c->fused_U0, f, c->fused_dtg));
      HOOK(ctx, c->ldiv(c->ud, out, c->W, c->w));
    } else {
      CTS_TRY(r);
    }
    c->stats.f_evals++;
    c->stats.jvp_evals++;
    c->stats.ldiv_calls++;
    return CTS_OK;
  };
  double rtol = 0;
  CTS_TRY(get_rtol(c, f, n_iter, &rtol));
  return gmres(c, matvec, use_M ? &precond : nullptr, f, dx, rtol, fuse_mvp ? &matvec_precond : nullptr, newton_x, keep_dx, applied);
}

bool krylov_matrix_free(const cts_newton_cache* c, const void* x, const void* f) {
  if (!c->fused_op || !c->cfg.use_krylov || !c->cfg.jacobian_free_jvp || c->cfg.disable_preconditioner || !c->W) return false;
  if (c->cfg.update_j.fn || c->cfg.update_j.every_signal != CTS_SIG_NEW_NEWTON_ITERATION) return false;
  if (c->ldiv != cts_op_column_ldiv || c->ud != (void*)c->fused_op || !c->fused_wfact_is_lib) return false;
  return cts_column_krylov_matrix_free(c->fused_op) && cts_aligned16(x) && cts_aligned16(f) && cts_aligned16(c->q);
}

// solve_newton! (newtons_method.jl:609-665); no line search, no convergence checker
// has_converged / updated_cache of a condition tree on scalars (the norm condition), convergence_condition.jl:41-143
bool norm_tree_converged(const cts_newton_cache* c, double val, double err, int iter) {
  bool st[CTS_MAX_COND_NODES];
  int sp = 0;
  for (int k = 0; k < c->checker.n_norm; ++k) {
    const cts_cond_node& nd = c->checker.norm_cond[k];
    bool r;
    switch (nd.kind) {
      case CTS_COND_MAX_ERROR: r = err <= nd.p0; break;
      case CTS_COND_MAX_REL_ERROR: r = err <= nd.p0 * val; break;
      case CTS_COND_MAX_ERROR_REDUCTION: r = iter >= 1 && err <= c->norm_cache[k]; break;
      case CTS_COND_MIN_RATE: r = iter >= 1 && err >= c->norm_cache[k]; break;
      default:
        r = nd.kind == CTS_COND_ALL;
        for (int j = 0; j < nd.nchildren; ++j) {
          const bool v = st[--sp];
          r = nd.kind == CTS_COND_ALL ? (r && v) : (r || v);
        }
    }
    st[sp++] = r;
  }
  return st[sp - 1];
}

bool tree_needs_update(const cts_cond_node* nodes, int n, int iter) {
  for (int k = 0; k < n; ++k)
    if ((nodes[k].kind == CTS_COND_MAX_ERROR_REDUCTION && iter == 0) || nodes[k].kind == CTS_COND_MIN_RATE) return true;
  return false;
}

void norm_tree_update(cts_newton_cache* c, double err, int iter) {
  auto store = [&](double v) { return c->ft == CTS_F32 ? (double)(float)v : v; };  // the cache has the state eltype
  for (int k = 0; k < c->checker.n_norm; ++k) {
    const cts_cond_node& nd = c->checker.norm_cond[k];
    if (nd.kind == CTS_COND_MAX_ERROR_REDUCTION && iter == 0) c->norm_cache[k] = store(nd.p0 * err);
    if (nd.kind == CTS_COND_MIN_RATE) {
      double p = nd.p1 == 1.0 ? err : nd.p1 == 2.0 ? err * err : nd.p1 == 3.0 ? err * err * err : std::pow(err, nd.p1);
      c->norm_cache[k] = store(nd.p0 * p);
    }
  }
}

// is_converged!(alg, cache, val, err, iter)  convergence_checker.jl:75-108
int is_converged(cts_newton_cache* c, const void* val, const void* err, int iter, bool* out) {
  *out = false;
  if (!c->has_checker) return CTS_OK;
  cts_ctx* ctx = c->ctx;
  const cts_convergence_checker& k = c->checker;
  bool converged = false;
  double norm_val = 0, norm_err = 0;
  auto component = [&](bool* r) -> int {
    int all = 0;
    CTS_TRY(cts_component_check(ctx, c->ft, c->n, val, err, k.n_comp, k.comp_cond, c->comp_cache, iter, c->dev_scratch, &all));
    *r = all != 0;
    return CTS_OK;
  };
  if (k.n_norm == 0) {
    CTS_TRY(component(&converged));
  } else {
    CTS_TRY(cts_nrm2(ctx, c->ft, c->n, val, &norm_val));
    CTS_TRY(cts_nrm2(ctx, c->ft, c->n, err, &norm_err));
    converged = norm_tree_converged(c, norm_val, norm_err, iter);
    if (k.n_comp > 0) {
      bool cc = false;
      if (!k.combiner_or) {
        if (converged) {
          CTS_TRY(component(&cc));
          converged = cc;
        }
      } else if (!converged) {
        CTS_TRY(component(&cc));
        converged = cc;
      }
    }
  }
  if (!converged) {  // only update caches if they will be needed for future iters
    if (k.n_norm > 0 && tree_needs_update(k.norm_cond, k.n_norm, iter)) norm_tree_update(c, norm_err, iter);
    if (k.n_comp > 0 && tree_needs_update(k.comp_cond, k.n_comp, iter))
      CTS_TRY(cts_component_update(ctx, c->ft, c->n, val, err, k.n_comp, k.comp_cond, c->comp_cache, iter));
  }
  *out = converged;
  return CTS_OK;
}

// line_search!(alg, x, Δx, f, f!, prepare_for_f!)  line_search.jl:28-47
int line_search(cts_newton_cache* c, void* x, const ResidualFn& f_, const PrepareFn* prepare) {
  cts_ctx* ctx = c->ctx;
  double normf0 = 0, normf = 0;
  CTS_TRY(cts_nrm2(ctx, c->ft, c->n, c->f, &normf0));
  if (prepare) CTS_TRY((*prepare)(x));
  CTS_TRY(f_(c->f, x));
  c->stats.f_evals++;
  CTS_TRY(cts_nrm2(ctx, c->ft, c->n, c->f, &normf));
  double alpha = 0.5;
  for (int i = 1; ((normf > normf0) || !std::isfinite(normf)) && i <= 5; ++i) {
    {
      cts_prof_scope ps(ctx, CTS_PROF_NEWTON_EW);
      double coef[1] = {alpha};
      const void* terms[1] = {c->dx};
      CTS_TRY(cts_axpy_n(ctx, c->ft, c->n, x, nullptr, 1.0, x, 1, 0, coef, terms, 0));  // `x .+= α * Δx`
    }
    if (prepare) CTS_TRY((*prepare)(x));
    CTS_TRY(f_(c->f, x));
    c->stats.f_evals++;
    CTS_TRY(cts_nrm2(ctx, c->ft, c->n, c->f, &normf));
    alpha /= 2;
  }
  return CTS_OK;
}

int solve_newton(cts_newton_cache* c, void* x, const ResidualFn& f_, const JacFn* j_, const PrepareFn* prepare) {
  cts_nvtx_scope nvtx("solve_newton!");  // newtons_method.jl:609
  cts_ctx* ctx = c->ctx;
  const cts_newton_config& cfg = c->cfg;
  const bool has_j = c->W != nullptr && j_ != nullptr;
  if (has_j && needs_update(cfg.update_j, CTS_SIG_NEW_NEWTON_SOLVE)) {
    CTS_TRY((*j_)(c->W, x));
    c->stats.wfact_evals++;
  }
  CTS_TRY(f_(c->f, x));
  c->stats.f_evals++;
  for (int n = 1; n <= cfg.max_iters; ++n) {
    c->stats.newton_iters++;
    if (has_j && needs_update(cfg.update_j, CTS_SIG_NEW_NEWTON_ITERATION)) {
      // matrix-free Krylov: the kernels of this iteration generate W(x) themselves, the stored copy is not refreshed
      if (!krylov_matrix_free(c, x, c->f)) CTS_TRY((*j_)(c->W, x));
      c->stats.wfact_evals++;
    }
    bool k5_done = false;
    const bool defer = c->defer_update && cfg.max_iters == 1 && !c->line_search && !c->has_checker;  // K5 fused by the caller
    if (!cfg.use_krylov) {
      if (!c->ldiv || !c->W) return cts_fail(ctx, CTS_EINVAL, "direct Newton needs ldiv! and a Jacobian");
      {
        cts_prof_scope ps(ctx, CTS_PROF_LDIV);
        HOOK(ctx, c->ldiv(c->ud, c->dx, c->W, c->f));
      }
      c->stats.ldiv_calls++;
    } else if (c->fuse_reductions && !defer) {
      // GMRES hands the update over fused with K5 (x is read-only inside solve_krylov! until then)
      CTS_TRY(solve_krylov(c, c->dx, x, f_, c->f, n - 1, prepare, x, c->line_search || c->has_checker, &k5_done));
    } else {
      CTS_TRY(solve_krylov(c, c->dx, x, f_, c->f, n - 1, prepare));
    }
    if (defer) return CTS_OK;
    if (!k5_done) {
      cts_prof_scope ps(ctx, CTS_PROF_NEWTON_EW);
      CTS_TRY(cts_sub_inplace(ctx, c->ft, c->n, x, c->dx));  // K5
    }
    if (c->line_search) CTS_TRY(line_search(c, x, f_, prepare));  // :652
    bool converged = false;
    CTS_TRY(is_converged(c, x, c->dx, n, &converged));  // :656
    if (converged) break;
    if (n < cfg.max_iters && !c->line_search) {  // :658-660
      if (prepare) CTS_TRY((*prepare)(x));
      CTS_TRY(f_(c->f, x));
      c->stats.f_evals++;
    }
  }
  return CTS_OK;
}

}  // namespace

extern "C" {

int cts_newton_cache_create(cts_ctx* ctx, cts_dtype ft, size_t n, const cts_newton_config* cfg, void* W, cts_ldiv_fn ldiv,
                            cts_ldiv_fn jac_mul, void* ud, cts_newton_cache** out) {
  CTS_CHECK_CTX(ctx);
  if (!cfg || !out || n == 0) return cts_fail(ctx, CTS_EINVAL, "cts_newton_cache_create: bad argument");
  // newtons_method.jl:585-588: a Jacobian is required unless the Krylov method is Jacobian free
  if (!W && !(cfg->use_krylov && cfg->jacobian_free_jvp))
    return cts_fail(ctx, CTS_EINVAL, "NewtonsMethod needs a jac_prototype unless a JacobianFreeJVP is used");
  auto* c = new cts_newton_cache();
  c->ctx = ctx;
  c->ft = ft;
  c->n = n;
  c->cfg = *cfg;
  c->W = W;
  c->ldiv = ldiv;
  c->jac_mul = jac_mul;
  c->ud = ud;
  int r = newton_alloc(c);
  if (r) {
    newton_free(c);
    delete c;
    return r;
  }
  *out = c;
  return CTS_OK;
}

int cts_newton_cache_destroy(cts_newton_cache* c) {
  if (!c) return CTS_EINVAL;
  newton_free(c);
  delete c;
  return CTS_OK;
}

int cts_newton_solve(cts_newton_cache* c, void* x, cts_residual_fn f, cts_jac_fn j, cts_prepare_fn prepare, void* ud) {
  if (!c || !x || !f) return CTS_EINVAL;
  cts_ctx* ctx = c->ctx;
  ResidualFn f_ = [&](void* fo, const void* xx) -> int {
    HOOK(ctx, f(ud, fo, xx));
    return CTS_OK;
  };
  JacFn j_ = [&](void* W, const void* xx) -> int {
    HOOK(ctx, j(ud, W, xx));
    return CTS_OK;
  };
  PrepareFn p_ = [&](void* xx) -> int {
    HOOK(ctx, prepare(ud, xx));
    return CTS_OK;
  };
  return solve_newton(c, x, f_, j ? &j_ : nullptr, prepare ? &p_ : nullptr);
}

int cts_newton_cache_stats(cts_newton_cache* c, cts_solver_stats* out) {
  if (!c || !out) return CTS_EINVAL;
  *out = c->stats;
  return CTS_OK;
}

int cts_newton_cache_set_convergence_checker(cts_newton_cache* c, const cts_convergence_checker* checker) {
  if (!c) return CTS_EINVAL;
  cts_ctx* ctx = c->ctx;
  for (int k = 0; k < CTS_MAX_COND_NODES; ++k) {
    if (c->comp_cache[k]) cts_free(ctx, c->comp_cache[k]);
    c->comp_cache[k] = nullptr;
    c->norm_cache[k] = std::nan("");
  }
  c->has_checker = false;
  if (!checker) return CTS_OK;
  if (checker->n_norm == 0 && checker->n_comp == 0)
    return cts_fail(ctx, CTS_EINVAL, "ConvergenceChecker must have at least one ConvergenceCondition");
  if (checker->n_norm < 0 || checker->n_norm > CTS_MAX_COND_NODES || checker->n_comp < 0 || checker->n_comp > CTS_MAX_COND_NODES)
    return cts_fail(ctx, CTS_EINVAL, "ConvergenceChecker: too many condition nodes");
  c->checker = *checker;
  if (!c->dev_scratch) CTS_TRY(cts_alloc(ctx, 64, (void**)&c->dev_scratch));
  const size_t es = cts_sizeof(c->ft);
  for (int k = 0; k < checker->n_comp; ++k) {
    const int kind = checker->comp_cond[k].kind;
    if (kind != CTS_COND_MAX_ERROR_REDUCTION && kind != CTS_COND_MIN_RATE) continue;
    CTS_TRY(cts_alloc(ctx, c->n * es, &c->comp_cache[k]));
    CTS_CUDA(ctx, cudaMemsetAsync(c->comp_cache[k], 0xff, c->n * es, ctx->stream));  // all-ones bit pattern == NaN
  }
  c->has_checker = true;
  return CTS_OK;
}

int cts_newton_cache_set_line_search(cts_newton_cache* c, int enabled) {
  if (!c) return CTS_EINVAL;
  c->line_search = enabled != 0;
  return CTS_OK;
}

int cts_newton_cache_is_converged(cts_newton_cache* c, const void* val, const void* err, int iter, int* converged) {
  if (!c || !val || !err || !converged) return CTS_EINVAL;
  bool r = false;
  CTS_TRY(is_converged(c, val, err, iter, &r));
  *converged = r ? 1 : 0;
  return CTS_OK;
}

}  // extern "C"

// =============================================================================================
// IMEX ARK / SSPRK
// =============================================================================================
struct cts_imex_cache {
  cts_ctx* ctx = nullptr;
  cts_dtype ft = CTS_F64;
  size_t n = 0;
  cts_imex_tableau tab{};
  int constraint = 0;
  cts_ode_function f{};
  cts_newton_config ncfg{};
  cts_newton_cache* nm = nullptr;
  bool has_gamma = false;
  double gamma = 0.0;
  bool fusion = true;
  bool w_pending = false;   // a due Wfact(dt·γ) folded into the next fused stage launch (which writes W itself)
  double w_pending_dtg = 0.0;
  const void* w_pending_u = nullptr;
  double w_pending_t = 0.0;
  // ARK
  void* U = nullptr;
  void* temp = nullptr;
  void* T_exp[CTS_MAX_STAGES] = {};
  void* T_lim[CTS_MAX_STAGES] = {};
  void* T_imp[CTS_MAX_STAGES] = {};
  // SSPRK
  void* U_exp = nullptr;
  void* U_lim = nullptr;
  void* T_exp1 = nullptr;
  void* T_lim1 = nullptr;
  double beta[CTS_MAX_STAGES] = {};
  int nbuffers = 0;
  cts_solver_stats stats{};
  // CUDA graph of one step (cts_imex_step): captured every step (the host still runs the stage loop, so times,
  // coefficients and the Jacobian policy are exact), applied to the instantiated graph with cudaGraphExecUpdate and
  // launched as ONE graph -- the kernels of a step, the halo exchanges and the fork/join with the communication stream
  // then run back to back without per-launch gaps.
  bool dt_f32 = false;             // typeof(dt) of the current step is Float32 (cts_imex_step's dt_is_f32)
  int graph_mode = -1;             // -1: decide on first use (on unless CTS_NO_GRAPH is set), 0: off, 1: on
  cudaGraphExec_t graph_exec = nullptr;
  uint64_t graph_launches = 0, graph_reinstantiations = 0;
};

namespace {

inline double A(const double* a, int s, int i, int j) { return a[i * s + j]; }

bool col_nonzero(const double* a, int s, int j) {
  for (int i = 0; i < s; ++i)
    if (A(a, s, i, j) != 0) return true;
  return false;
}

bool is_fsal(const cts_imex_tableau& t) {
  for (int j = 0; j < t.s; ++j)
    if (t.b_exp[j] != A(t.a_exp, t.s, t.s - 1, j) || t.b_imp[j] != A(t.a_imp, t.s, t.s - 1, j)) return false;
  return true;
}

int alloc_state(cts_imex_cache* c, void** p) {
  CTS_TRY(cts_alloc(c->ctx, c->n * cts_sizeof(c->ft), p));
  c->nbuffers++;
  return CTS_OK;
}

// Broadcast arithmetic is native Float32 only when the state, the (downcast) tableau AND dt are Float32; a Float64 dt or a
// Float64 tableau promotes every expression to Float64, rounded on store.
int native_flag(const cts_imex_cache* c) { return (c->tab.cast_to_state_eltype && c->ft == CTS_F32 && c->dt_f32) ? 1 : 0; }
// coefficient fl(dt * a) in promote(typeof(dt), eltype(tableau)) (fused_increment.jl:58,115-116)
double coef_of(const cts_imex_cache* c, double dt, double a) {
  if (native_flag(c)) return (double)((float)dt * (float)a);
  return dt * a;  // (a is already Float32-valued when the tableau was downcast)
}
// t + dt*c and float(dt)*a_ii (imex_ark.jl:120-122) in the same promoted type
double time_of(const cts_imex_cache* c, double t, double dt, double cc) {
  if (native_flag(c)) return (double)((float)t + (float)dt * (float)cc);
  return t + dt * cc;
}

struct Increment {  // a fused_raw_increment: the non-zero (coef, tendency) pairs, NullBroadcasted when n == 0
  int n = 0;
  double coef[CTS_MAX_STAGES];
  const void* term[CTS_MAX_STAGES];
};

Increment raw_increment(const cts_imex_cache* c, double dt, const double* coefs, int count, void* const* tend) {
  Increment inc;
  for (int j = 0; j < count; ++j) {
    if (coefs[j] == 0) continue;
    inc.coef[inc.n] = coef_of(c, dt, coefs[j]);
    inc.term[inc.n] = tend[j];
    inc.n++;
  }
  return inc;
}

// assign_with_increments!(out, base, inc_exp, inc_imp)  (fused_increment.jl:220-236)
int assign_with_increments(cts_imex_cache* c, void* out, void* out2, double c0, const void* base, const Increment& g1,
                           const Increment& g2, int cat = CTS_PROF_STAGE_ASSEMBLY) {
  cts_prof_scope ps(c->ctx, cat);
  double coef[CTS_MAX_TERMS];
  const void* terms[CTS_MAX_TERMS];
  int k = 0;
  for (int j = 0; j < g1.n; ++j, ++k) {
    coef[k] = g1.coef[j];
    terms[k] = g1.term[j];
  }
  for (int j = 0; j < g2.n; ++j, ++k) {
    coef[k] = g2.coef[j];
    terms[k] = g2.term[j];
  }
  if (k == 0 && out == base && c0 == 1.0 && !out2) return CTS_OK;  // `@. u = u`
  return cts_axpy_n(c->ctx, c->ft, c->n, out, out2, c0, base, g1.n, g2.n, coef, terms, native_flag(c));
}

int maybe_update_jacobian(cts_imex_cache* c, const void* u, double t, double dt, bool defer = false,
                          bool matrix_free = false) {  // async_utils.jl:11-31
  if (!c->f.T_imp || !c->ncfg.present || !c->nm || !c->nm->W) return CTS_OK;
  if (!needs_update(c->ncfg.update_j, CTS_SIG_NEW_TIMESTEP, t)) return CTS_OK;
  if (matrix_free) return CTS_OK;  // the handler has seen the signal; W is regenerated inside the stage kernels
  if (defer) {  // the first fused stage of this step regenerates W in registers and stores it: no separate pass
    c->w_pending = true;
    c->w_pending_dtg = dt * c->gamma;
    c->w_pending_u = u;
    c->w_pending_t = t;
    c->stats.wfact_evals++;
    return CTS_OK;
  }
  cts_prof_scope ps(c->ctx, CTS_PROF_WFACT);
  HOOK(c->ctx, c->f.Wfact(c->f.ud, c->nm->W, u, dt * c->gamma, t));
  c->stats.wfact_evals++;
  return CTS_OK;
}

int flush_pending_jacobian(cts_imex_cache* c) {  // a deferred Wfact that no fused stage consumed
  if (!c->w_pending) return CTS_OK;
  c->w_pending = false;
  cts_prof_scope ps(c->ctx, CTS_PROF_WFACT);
  HOOK(c->ctx, c->f.Wfact(c->f.ud, c->nm->W, c->w_pending_u, c->w_pending_dtg, c->w_pending_t));
  return CTS_OK;
}

int call_dss(cts_imex_cache* c, void* u, double t) {
  if (c->f.dss) {
    cts_prof_scope ps(c->ctx, CTS_PROF_DSS);
    HOOK(c->ctx, c->f.dss(c->f.ud, u, t));
    c->stats.dss_calls++;
  }
  return CTS_OK;
}
int call_state(cts_imex_cache* c, cts_state_fn fn, void* u, double t) {
  if (fn) {
    cts_prof_scope ps(c->ctx, CTS_PROF_OTHER_HOOKS);
    HOOK(c->ctx, fn(c->f.ud, u, t));
  }
  return CTS_OK;
}

// solve_implicit_equation! (imex_ark.jl:283-308)
int solve_implicit_equation(cts_imex_cache* c, void* U, const void* U0, double t_imp, double dtg) {
  cts_ctx* ctx = c->ctx;
  // library column operator as T_imp! and nothing to run before f! (cache_imp! is the reference's prepare_for_f!):
  // residual and JVP go through the fused kernel
  cts_op_column* rop = nullptr;
  if (c->fusion && !native_flag(c) && c->f.T_imp == cts_op_column_T_imp && !c->f.cache_imp && c->f.ud) {
    auto* op = static_cast<cts_op_column*>(c->f.ud);
    if (cts_op_column_ctx(op) == ctx && cts_op_column_dtype(op) == c->ft && cts_op_column_ndof(op) == c->n &&
        cts_column_residual_fusable(op) && cts_aligned16(U) && cts_aligned16(U0) && cts_aligned16(c->nm->f) &&
        cts_aligned16(c->nm->dx))
      rop = op;
  }
  ResidualFn f_ = [&](void* res, const void* Up) -> int {
    if (rop && cts_aligned16(res) && cts_aligned16(Up)) {
      cts_prof_scope ps(ctx, CTS_PROF_NEWTON_EW);
      return cts_column_fused_residual(rop, res, Up, nullptr, 0.0, U0, nullptr, dtg);
    }
    {
      cts_prof_scope ps(ctx, CTS_PROF_T_IMP);
      HOOK(ctx, c->f.T_imp(c->f.ud, res, Up, t_imp));
    }
    cts_prof_scope ps(ctx, CTS_PROF_NEWTON_EW);
    return cts_newton_residual_ex(ctx, c->ft, c->n, res, U0, dtg, Up, native_flag(c));  // K3
  };
  struct FusedScope {  // the JVP of solve_krylov! sees the operator only for the duration of this solve
    cts_newton_cache* nm;
    ~FusedScope() { nm->fused_op = nullptr; }
  } fused_scope{c->nm};
  c->nm->fused_op = rop;
  c->nm->fused_U0 = U0;
  c->nm->fused_dtg = dtg;
  c->nm->fused_wfact_is_lib = c->f.Wfact == cts_op_column_Wfact;
  JacFn j_ = [&](void* W, const void* Up) -> int {
    cts_prof_scope ps(ctx, CTS_PROF_WFACT);
    HOOK(ctx, c->f.Wfact(c->f.ud, W, Up, dtg, t_imp));
    return CTS_OK;
  };
  PrepareFn p_ = [&](void* Up) -> int { return call_state(c, c->f.cache_imp, Up, t_imp); };
  return solve_newton(c->nm, U, f_, c->f.Wfact ? &j_ : nullptr, &p_);
}

// Is the implicit stage eligible for the fused column kernel?
cts_op_column* fusable_column_op(const cts_imex_cache* c) {
  const cts_ode_function& f = c->f;
  if (!c->fusion || !c->ncfg.present || c->ncfg.max_iters != 1 || c->ncfg.use_krylov) return nullptr;
  if (native_flag(c)) return nullptr;  // the column kernels evaluate K3/K6 with a Float64 dtγ
  if (c->nm && (c->nm->has_checker || c->nm->line_search)) return nullptr;  // is_converged! / line_search! must run
  if (f.T_imp != cts_op_column_T_imp || f.Wfact != cts_op_column_Wfact || f.ldiv != cts_op_column_ldiv) return nullptr;
  if (f.has_lim) return nullptr;
  if (f.dss && f.dss != cts_op_column_dss) return nullptr;  // the library dss! only rewrites ghost rows
  if (f.constrain_state || f.cache || f.cache_imp || f.initialize_imp) return nullptr;
  auto* op = static_cast<cts_op_column*>(f.ud);
  if (!op || !cts_column_fusable(op) || cts_op_column_ctx(op) != c->ctx || cts_op_column_dtype(op) != c->ft) return nullptr;
  if (cts_op_column_ndof(op) != c->n) return nullptr;
  // a matrix-free stage always uses W(dtγ of the stage); with a host handler (UpdateEveryN/UpdateEveryDt) the policy may ask
  // for a stale W instead, which only the stored form can honour
  if (cts_op_column_matrix_free(op) && c->ncfg.update_j.fn) return nullptr;
  return op;
}

int fused_implicit_stage(cts_imex_cache* c, cts_op_column* op, void* U, double c0, const void* base, const Increment& g1,
                         const Increment& g2, double dtg, double t_imp, void* T_imp_out) {
  // Jacobian refresh exactly when the reference would do it inside solve_newton! (the linear operator's W
  // does not depend on the state, so the stage value is not needed to build it).
  const bool mf = cts_op_column_matrix_free(op);
  bool emit_W = false;
  const bool upd_solve = needs_update(c->ncfg.update_j, CTS_SIG_NEW_NEWTON_SOLVE);     // newtons_method.jl:620
  const bool upd_iter = needs_update(c->ncfg.update_j, CTS_SIG_NEW_NEWTON_ITERATION);  // :626 (both are always sent)
  if (upd_solve || upd_iter) {
    emit_W = !mf;  // W(dtγ of this stage) is produced and stored by the stage kernel itself
    c->w_pending = false;
    c->stats.wfact_evals++;
  } else if (c->w_pending) {
    if (dtg == c->w_pending_dtg) {
      emit_W = true;
      c->w_pending = false;
    } else {
      CTS_TRY(flush_pending_jacobian(c));
    }
  }
  cts_fused_stage_args a{};
  a.base = base;
  a.c0 = c0;
  a.n_grp1 = g1.n;
  a.n_grp2 = g2.n;
  int k = 0;
  for (int j = 0; j < g1.n; ++j, ++k) {
    a.coef[k] = g1.coef[j];
    a.terms[k] = g1.term[j];
  }
  for (int j = 0; j < g2.n; ++j, ++k) {
    a.coef[k] = g2.coef[j];
    a.terms[k] = g2.term[j];
  }
  a.compute_native = native_flag(c);
  a.dtgamma = dtg;
  a.W = static_cast<const cts_tridiag*>(c->nm->W);
  a.emit_W = emit_W ? 1 : 0;
  a.U = U;
  a.T_imp = T_imp_out;
  a.temp = nullptr;
  {
    cts_prof_scope ps(c->ctx, CTS_PROF_FUSED_STAGE);
    CTS_TRY(cts_column_fused_stage(op, &a));
  }
  c->stats.fused_stage_launches++;
  c->stats.newton_iters++;
  return CTS_OK;
}

// Can the last stage's T_exp! be folded into the final update?  (library column operator as T_exp_T_lim!, no limiter,
// the library halo dss!, and a tableau whose last explicit tendency feeds only the final update)
cts_op_column* final_fusable_column_op(const cts_imex_cache* c) {
  const cts_ode_function& f = c->f;
  const cts_imex_tableau& tab = c->tab;
  const int s = tab.s;
  if (!c->fusion || f.has_lim || f.T_exp_T_lim != cts_op_column_T_exp || !f.ud) return nullptr;
  if (tab.b_exp[s - 1] == 0 || col_nonzero(tab.a_exp, s, s - 1)) return nullptr;
  auto* op = static_cast<cts_op_column*>(f.ud);
  if (cts_op_column_ctx(op) != c->ctx || cts_op_column_dtype(op) != c->ft || cts_op_column_ndof(op) != c->n) return nullptr;
  if (!cts_column_final_fusable(op)) return nullptr;
  if (f.dss ? f.dss != cts_op_column_dss : cts_op_column_has_halo(op)) return nullptr;  // ghost rows are left to dss!
  return op;
}

// ---------------------------------------------------------------------------------------------
// step_u!(integrator, cache::IMEXARKCache)   imex_ark.jl:69-272
// ---------------------------------------------------------------------------------------------
int ark_step(cts_imex_cache* c, void* u, double t, double dt) {
  const cts_imex_tableau& tab = c->tab;
  const cts_ode_function& f = c->f;
  const int s = tab.s;
  const bool has_T_exp = f.T_exp_T_lim != nullptr;
  const bool has_T_imp = f.T_imp != nullptr;
  cts_op_column* fuse_op = fusable_column_op(c);
  if (fuse_op && !cts_aligned16(u)) fuse_op = nullptr;  // (cache buffers are cudaMalloc'ed; only the caller's u can be off)
  cts_op_column* final_op = final_fusable_column_op(c);
  bool fold_last_texp = false;
  bool last_dss_deferred = false;  // dss!(U_s) travels with the overlapped final update
  const void* U_last = nullptr;
  double t_exp_last = 0.0, t_imp_last = 0.0;
  // (a matrix-free fused stage regenerates W in the kernel: the stored copy is never read)
  CTS_TRY(maybe_update_jacobian(c, u, t, dt, fuse_op != nullptr, fuse_op && cts_op_column_matrix_free(fuse_op)));

  for (int i = 0; i < s; ++i) {
    const double t_exp = time_of(c, t, dt, tab.c_exp[i]);
    const double t_imp = time_of(c, t, dt, tab.c_imp[i]);
    const double dtg = coef_of(c, dt, A(tab.a_imp, s, i, i));
    const bool no_imp = !has_T_imp || dtg == 0;
    const bool has_exp_terms = col_nonzero(tab.a_exp, s, i) || tab.b_exp[i] != 0;
    const bool has_imp_terms = col_nonzero(tab.a_imp, s, i) || tab.b_imp[i] != 0;
    void* U = c->U;

    // compute_stage_value!  :140-165
    Increment inc_exp = has_T_exp ? raw_increment(c, dt, &tab.a_exp[i * s], i, c->T_exp) : Increment{};
    Increment inc_imp = has_T_imp ? raw_increment(c, dt, &tab.a_imp[i * s], i, c->T_imp) : Increment{};

    bool stage_done = false, overlap_dss = false, temp_done = false, timp_done = false;
    if (fuse_op && !no_imp) {
      // K1 + Newton + K6 in one kernel; the pre-Newton dss! calls only rewrite ghost rows that the column
      // solve never reads, so one post-Newton dss! leaves every owned DOF bit-identical.
      CTS_TRY(fused_implicit_stage(c, fuse_op, U, 1.0, u, inc_exp, inc_imp, dtg, t_imp, has_imp_terms ? c->T_imp[i] : nullptr));
      // multi-GPU: when the explicit tendency of this stage follows, its interior rows are evaluated while the halo
      // exchange of dss!(U) is in flight (cts_column_dss_T_exp_overlapped); otherwise dss! runs here
      const bool can_overlap = has_exp_terms && f.T_exp_T_lim == cts_op_column_T_exp && f.dss == cts_op_column_dss &&
                               cts_column_overlap_ok(fuse_op);
      overlap_dss = can_overlap && !(i == s - 1 && final_op);
      // last stage with the folded final update: halo(U_s), the boundary strips of the update and halo(u) run beside
      // the interior rows of the update (cts_column_final_update_overlapped)
      last_dss_deferred = can_overlap && i == s - 1 && final_op == fuse_op && U != u;
      if (!overlap_dss && !last_dss_deferred) CTS_TRY(call_dss(c, U, t_imp));
      stage_done = true;
    } else if (fuse_op && no_imp && i == 0 && inc_exp.n == 0 && inc_imp.n == 0) {
      U = u;  // `@. U = u` followed only by tendency evaluations: alias instead of copying
    } else if (f.has_lim) {
      Increment inc_lim = raw_increment(c, dt, &tab.a_exp[i * s], i, c->T_lim);
      CTS_TRY(assign_with_increments(c, U, nullptr, 1.0, u, inc_lim, Increment{}));  // assign_fused_increment! :159
      if (i != 0 && f.lim) HOOK(c->ctx, f.lim(f.ud, U, t_exp, u));
      CTS_TRY(assign_with_increments(c, U, nullptr, 1.0, U, inc_exp, inc_imp));
    } else {
      // `temp = U` (:209) rides along when no hook may touch U between the assembly and the copy
      temp_done = c->fusion && !no_imp && !f.dss && !f.constrain_state && !f.cache_imp && !f.initialize_imp && c->nm != nullptr;
      CTS_TRY(assign_with_increments(c, U, temp_done ? c->temp : nullptr, 1.0, u, inc_exp, inc_imp));
    }

    // solve_stage_implicit!  :174-242
    if (!stage_done) {
      const int top_sig = no_imp ? CTS_SIG_END_OF_STAGE : CTS_SIG_WITH_DSS;
      if (i != 0) {
        CTS_TRY(call_dss(c, U, t_exp));
        if (needs_update(f.update_constrain_state, top_sig)) CTS_TRY(call_state(c, f.constrain_state, U, t_exp));
        if (no_imp && needs_update(f.update_cache, CTS_SIG_END_OF_STAGE)) CTS_TRY(call_state(c, f.cache, U, t_exp));
      }
      if (!no_imp) {
        if (!c->ncfg.present || !c->nm) return cts_fail(c->ctx, CTS_EINVAL, "implicit stage without a NewtonsMethod");
        if (!temp_done) {
          cts_prof_scope ps(c->ctx, CTS_PROF_NEWTON_EW);
          CTS_TRY(cts_axpy_n(c->ctx, c->ft, c->n, c->temp, nullptr, 1.0, U, 0, 0, nullptr, nullptr, 0));  // K2 :209
        }
        if (!f.initialize_imp_is_nothing) {
          if (f.initialize_imp) HOOK(c->ctx, f.initialize_imp(f.ud, U, dtg));
          CTS_TRY(call_dss(c, U, t_exp));
          if (needs_update(f.update_constrain_state, CTS_SIG_WITH_DSS)) CTS_TRY(call_state(c, f.constrain_state, U, t_exp));
          CTS_TRY(call_state(c, f.cache_imp, U, t_imp));
        } else if (i != 0) {
          CTS_TRY(call_state(c, f.cache_imp, U, t_imp));
        }
        // K5 + K6 in one pass: a one-iteration direct solve, no dss!/constrain_state!/cache! between `x .-= Δx` (:651) and
        // `T_imp[i] = (U - temp)/dtγ` (:269), and temp == the pre-update U bit for bit
        const bool fuse56 = c->fusion && has_imp_terms && c->ncfg.max_iters == 1 && !c->ncfg.use_krylov && !c->nm->has_checker &&
                            !c->nm->line_search && !f.dss && !f.constrain_state && !f.cache && !f.initialize_imp && U != c->T_imp[i];
        c->nm->defer_update = fuse56;
        const int rsolve = solve_implicit_equation(c, U, c->temp, t_imp, dtg);
        c->nm->defer_update = false;
        CTS_TRY(rsolve);
        if (fuse56) {
          cts_prof_scope ps(c->ctx, CTS_PROF_NEWTON_EW);
          CTS_TRY(cts_newton_update_timp(c->ctx, c->ft, c->n, U, c->T_imp[i], c->nm->dx, dtg, native_flag(c)));
          timp_done = true;
        }
        CTS_TRY(call_dss(c, U, t_imp));
        if (!(i == s - 1 && is_fsal(tab))) {
          if (needs_update(f.update_constrain_state, CTS_SIG_END_OF_STAGE)) CTS_TRY(call_state(c, f.constrain_state, U, t_imp));
          if (needs_update(f.update_cache, CTS_SIG_END_OF_STAGE)) CTS_TRY(call_state(c, f.cache, U, t_imp));
        }
      }
    }

    // evaluate_stage_tendencies!  :252-272
    if (has_exp_terms && f.T_exp_T_lim) {
      if (i == s - 1 && final_op && U != u) {
        fold_last_texp = true;  // evaluated inside the final update kernel (its only reader)
        U_last = U;
        t_exp_last = t_exp;
        t_imp_last = t_imp;
      } else if (overlap_dss) {
        cts_prof_scope ps(c->ctx, CTS_PROF_T_EXP);
        CTS_TRY(cts_column_dss_T_exp_overlapped(fuse_op, c->T_exp[i], U, t_imp, t_exp));
        c->stats.dss_calls++;
      } else {
        cts_prof_scope ps(c->ctx, CTS_PROF_T_EXP);
        HOOK(c->ctx, f.T_exp_T_lim(f.ud, c->T_exp[i], c->T_lim[i], U, t_exp));
      }
    }
    if (has_imp_terms && has_T_imp && !stage_done && !timp_done) {
      if (dtg == 0) {
        cts_prof_scope ps(c->ctx, CTS_PROF_T_IMP);
        HOOK(c->ctx, f.T_imp(f.ud, c->T_imp[i], U, t_imp));
      } else {
        cts_prof_scope ps(c->ctx, CTS_PROF_NEWTON_EW);
        CTS_TRY(cts_timp_from_stage_ex(c->ctx, c->ft, c->n, c->T_imp[i], U, c->temp, dtg, native_flag(c)));  // K6 :269
      }
    }
  }

  // final update :79-101
  CTS_TRY(flush_pending_jacobian(c));
  const double t_final = t + dt;
  Increment inc_exp = has_T_exp ? raw_increment(c, dt, tab.b_exp, s, c->T_exp) : Increment{};
  Increment inc_imp = has_T_imp ? raw_increment(c, dt, tab.b_imp, s, c->T_imp) : Increment{};
  if (f.has_lim) {
    Increment inc_lim = raw_increment(c, dt, tab.b_exp, s, c->T_lim);
    CTS_TRY(assign_with_increments(c, c->temp, nullptr, 1.0, u, inc_lim, Increment{}));
    if (f.lim) HOOK(c->ctx, f.lim(f.ud, c->temp, t_final, u));
    CTS_TRY(assign_with_increments(c, u, nullptr, 1.0, c->temp, inc_exp, inc_imp, CTS_PROF_FINAL_UPDATE));
  } else if (fold_last_texp) {
    cts_final_update_args a{};
    a.u = u;
    a.U_last = U_last;
    a.t_exp_last = t_exp_last;
    a.coef_last = coef_of(c, dt, tab.b_exp[s - 1]);
    Increment others = raw_increment(c, dt, tab.b_exp, s - 1, c->T_exp);
    a.n_grp1 = others.n;
    a.n_grp2 = inc_imp.n;
    int k = 0;
    for (int j = 0; j < others.n; ++j, ++k) {
      a.coef[k] = others.coef[j];
      a.terms[k] = others.term[j];
    }
    for (int j = 0; j < inc_imp.n; ++j, ++k) {
      a.coef[k] = inc_imp.coef[j];
      a.terms[k] = inc_imp.term[j];
    }
    a.compute_native = native_flag(c);
    int r;
    if (last_dss_deferred && cts_column_final_update_supported(final_op, &a)) {
      cts_prof_scope ps(c->ctx, CTS_PROF_FINAL_UPDATE);
      CTS_TRY(cts_column_final_update_overlapped(final_op, &a));  // dss!(U_s); update; dss!(u)
      c->stats.dss_calls += 2;
      return CTS_OK;  // (no constrain_state!/cache! hooks on this path: fusable_column_op)
    }
    if (last_dss_deferred) CTS_TRY(call_dss(c, const_cast<void*>(U_last), t_imp_last));
    {
      cts_prof_scope ps(c->ctx, CTS_PROF_FINAL_UPDATE);
      r = cts_column_fused_final_update(final_op, &a);
    }
    if (r == CTS_EUNSUPPORTED) {  // alignment / operand count: the two-pass form
      {
        cts_prof_scope ps(c->ctx, CTS_PROF_T_EXP);
        HOOK(c->ctx, f.T_exp_T_lim(f.ud, c->T_exp[s - 1], c->T_lim[s - 1], const_cast<void*>(U_last), t_exp_last));
      }
      CTS_TRY(assign_with_increments(c, u, nullptr, 1.0, u, inc_exp, inc_imp, CTS_PROF_FINAL_UPDATE));
    } else {
      CTS_TRY(r);
    }
  } else {
    CTS_TRY(assign_with_increments(c, u, nullptr, 1.0, u, inc_exp, inc_imp, CTS_PROF_FINAL_UPDATE));
  }
  CTS_TRY(call_dss(c, u, t_final));
  if (needs_update(f.update_constrain_state, CTS_SIG_END_OF_STEP)) CTS_TRY(call_state(c, f.constrain_state, u, t_final));
  if (needs_update(f.update_cache, CTS_SIG_END_OF_STEP)) CTS_TRY(call_state(c, f.cache, u, t_final));
  return CTS_OK;
}

// ---------------------------------------------------------------------------------------------
// step_u!(integrator, cache::IMEXSSPRKCache)   imex_ssprk.jl:95-220
// ---------------------------------------------------------------------------------------------
// `@. x = x + dt * T` in promote(typeof(dt), FT)   (imex_ssprk.jl:112,117,158,163)
int axpy_dt(cts_imex_cache* c, void* out, const void* x, double dt, const void* T, int dt_is_f32) {
  cts_prof_scope ps(c->ctx, CTS_PROF_STAGE_ASSEMBLY);
  double coef[1] = {dt};
  const void* terms[1] = {T};
  int native = (dt_is_f32 && c->ft == CTS_F32) ? 1 : 0;
  return cts_axpy_n(c->ctx, c->ft, c->n, out, nullptr, 1.0, x, 1, 0, coef, terms, native);
}

int ssprk_step(cts_imex_cache* c, void* u, double t, double dt, int dt_is_f32) {
  const cts_imex_tableau& tab = c->tab;
  const cts_ode_function& f = c->f;
  const int s = tab.s;
  const bool has_T_exp = f.T_exp_T_lim != nullptr;
  const bool has_T_imp = f.T_imp != nullptr;
  const int native = native_flag(c);
  cts_op_column* fuse_op = fusable_column_op(c);
  if (fuse_op && !cts_aligned16(u)) fuse_op = nullptr;
  // β is formed from the tableau as given (imex_ssprk.jl:48-49: Float64 for the named methods), so the Shu-Osher blends
  // are Float64 broadcasts whatever the state and dt types are; only the implicit increment can be native Float32
  auto beta_c = [&](double b) { return b; };
  CTS_TRY(maybe_update_jacobian(c, u, t, dt, fuse_op != nullptr, fuse_op && cts_op_column_matrix_free(fuse_op)));

  for (int i = 0; i < s; ++i) {
    const double t_exp = time_of(c, t, dt, tab.c_exp[i]);
    const double t_imp = time_of(c, t, dt, tab.c_imp[i]);
    const double aii = A(tab.a_imp, s, i, i);
    const double dtg = coef_of(c, dt, aii);
    Increment inc_imp = has_T_imp ? raw_increment(c, dt, &tab.a_imp[i * s], i, c->T_imp) : Increment{};
    const bool no_imp = !has_T_imp || aii == 0;
    const bool has_imp_terms = col_nonzero(tab.a_imp, s, i) || tab.b_imp[i] != 0;
    const bool fuse_column = fuse_op && !no_imp;
    // `temp = U` (:190) can ride along with the assembly when no hook touches U in between
    const bool temp_rides = !no_imp && !fuse_column && i != 0 && !f.dss && !f.constrain_state && !f.cache_imp;
    bool assembled = false, temp_done = false;
    if (i == 0) {
      cts_prof_scope ps(c->ctx, CTS_PROF_STAGE_ASSEMBLY);
      CTS_TRY(cts_axpy_n(c->ctx, c->ft, c->n, c->U_exp, nullptr, 1.0, u, 0, 0, nullptr, nullptr, 0));  // :155
    } else if (c->beta[i - 1] != 0) {
      const double b = beta_c(c->beta[i - 1]);
      const double omb = 1.0 - b;
      if (!f.has_lim && c->fusion && inc_imp.n <= 8) {
        // K11: `U_exp += dt*T_exp` :162, `U_exp = (1-β)u + βU_exp` :165, `U = U_exp + inc_imp` :169 (and `temp = U`)
        // in one pass; the fused column stage assembles U itself, so only U_exp is produced for it.
        cts_prof_scope ps(c->ctx, CTS_PROF_STAGE_ASSEMBLY);
        CTS_TRY(cts_ssprk_stage(c->ctx, c->ft, c->n, c->U_exp, fuse_column ? nullptr : c->U, temp_rides ? c->temp : nullptr,
                                c->U_exp, has_T_exp ? c->T_exp1 : nullptr, u, dt, (dt_is_f32 && c->ft == CTS_F32) ? 1 : 0, b,
                                omb, fuse_column ? 0 : inc_imp.n, inc_imp.coef, inc_imp.term, native, 0));
        assembled = !fuse_column;
        temp_done = temp_rides;
      } else {
        if (f.has_lim) {
          CTS_TRY(axpy_dt(c, c->U_lim, c->U_exp, dt, c->T_lim1, dt_is_f32));
          if (f.lim) HOOK(c->ctx, f.lim(f.ud, c->U_lim, t_exp, c->U_exp));
          CTS_TRY(cts_axpy_n(c->ctx, c->ft, c->n, c->U_exp, nullptr, 1.0, c->U_lim, 0, 0, nullptr, nullptr, 0));
        }
        if (has_T_exp) CTS_TRY(axpy_dt(c, c->U_exp, c->U_exp, dt, c->T_exp1, dt_is_f32));
        // `@. U_exp = (1 - β) * u + β * U_exp`  :165
        double coef[1] = {b};
        const void* terms[1] = {c->U_exp};
        cts_prof_scope ps(c->ctx, CTS_PROF_STAGE_ASSEMBLY);
        CTS_TRY(cts_axpy_n(c->ctx, c->ft, c->n, c->U_exp, nullptr, omb, u, 1, 0, coef, terms, 0));
        // NB: c0 == 1 would skip the multiply; (1-β) == 1 only when β == 0, which this branch excludes.
      }
    }
    bool stage_done = false;
    if (fuse_column) {
      CTS_TRY(fused_implicit_stage(c, fuse_op, c->U, 1.0, c->U_exp, inc_imp, Increment{}, dtg, t_imp,
                                   has_imp_terms ? c->T_imp[i] : nullptr));
      CTS_TRY(call_dss(c, c->U, t_imp));
      stage_done = true;
    } else if (!assembled) {
      CTS_TRY(assign_with_increments(c, c->U, nullptr, 1.0, c->U_exp, inc_imp, Increment{}));  // :169
    }
    if (!stage_done) {
      const int top_sig = no_imp ? CTS_SIG_END_OF_STAGE : CTS_SIG_WITH_DSS;
      if (i != 0) {
        CTS_TRY(call_dss(c, c->U, t_exp));
        if (needs_update(f.update_constrain_state, top_sig)) CTS_TRY(call_state(c, f.constrain_state, c->U, t_exp));
        if (no_imp && needs_update(f.update_cache, CTS_SIG_END_OF_STAGE)) CTS_TRY(call_state(c, f.cache, c->U, t_exp));
      }
      if (!no_imp) {
        if (!c->ncfg.present || !c->nm) return cts_fail(c->ctx, CTS_EINVAL, "implicit stage without a NewtonsMethod");
        if (i != 0) CTS_TRY(call_state(c, f.cache_imp, c->U, t_imp));
        if (!temp_done) {
          cts_prof_scope ps(c->ctx, CTS_PROF_NEWTON_EW);
          CTS_TRY(cts_axpy_n(c->ctx, c->ft, c->n, c->temp, nullptr, 1.0, c->U, 0, 0, nullptr, nullptr, 0));  // :190
        }
        CTS_TRY(solve_implicit_equation(c, c->U, c->temp, t_imp, dtg));
        CTS_TRY(call_dss(c, c->U, t_imp));
        if (needs_update(f.update_constrain_state, CTS_SIG_END_OF_STAGE)) CTS_TRY(call_state(c, f.constrain_state, c->U, t_imp));
        if (needs_update(f.update_cache, CTS_SIG_END_OF_STAGE)) CTS_TRY(call_state(c, f.cache, c->U, t_imp));
      }
    }
    if (c->beta[i] != 0 && f.T_exp_T_lim) {
      cts_prof_scope ps(c->ctx, CTS_PROF_T_EXP);
      HOOK(c->ctx, f.T_exp_T_lim(f.ud, c->T_exp1, c->T_lim1, c->U, t_exp));
    }
    if (has_imp_terms && has_T_imp && !stage_done) {
      if (aii == 0) {
        cts_prof_scope ps(c->ctx, CTS_PROF_T_IMP);
        HOOK(c->ctx, f.T_imp(f.ud, c->T_imp[i], c->U, t_imp));
      } else {
        cts_prof_scope ps(c->ctx, CTS_PROF_NEWTON_EW);
        CTS_TRY(cts_timp_from_stage_ex(c->ctx, c->ft, c->n, c->T_imp[i], c->U, c->temp, dtg, native_flag(c)));
      }
    }
  }

  CTS_TRY(flush_pending_jacobian(c));
  const double t_final = t + dt;
  Increment inc_imp = has_T_imp ? raw_increment(c, dt, tab.b_imp, s, c->T_imp) : Increment{};
  if (c->beta[s - 1] != 0 && !f.has_lim && c->fusion && inc_imp.n <= 8) {
    // K11, final form: `U_exp += dt*T_exp` :117 and `u = (1-β)u + βU_exp + inc_imp` :120-122 in one pass
    const double b = beta_c(c->beta[s - 1]);
    const double omb = 1.0 - b;
    cts_prof_scope ps(c->ctx, CTS_PROF_FINAL_UPDATE);
    CTS_TRY(cts_ssprk_stage(c->ctx, c->ft, c->n, nullptr, u, nullptr, c->U_exp, has_T_exp ? c->T_exp1 : nullptr, u, dt,
                            (dt_is_f32 && c->ft == CTS_F32) ? 1 : 0, b, omb, inc_imp.n, inc_imp.coef, inc_imp.term, native, 1));
  } else if (c->beta[s - 1] != 0) {
    if (f.has_lim) {
      CTS_TRY(axpy_dt(c, c->U_lim, c->U_exp, dt, c->T_lim1, dt_is_f32));
      if (f.lim) HOOK(c->ctx, f.lim(f.ud, c->U_lim, t_final, c->U_exp));
      CTS_TRY(cts_axpy_n(c->ctx, c->ft, c->n, c->U_exp, nullptr, 1.0, c->U_lim, 0, 0, nullptr, nullptr, 0));
    }
    if (has_T_exp) CTS_TRY(axpy_dt(c, c->U_exp, c->U_exp, dt, c->T_exp1, dt_is_f32));
    // `@. u = (1 - β) * u + β * U_exp + inc_imp`  :120-122: the blend in Float64, the increment in its own type -> K11's
    // final form without the explicit tendency (already added above)
    const double b = beta_c(c->beta[s - 1]);
    const double omb = 1.0 - b;
    cts_prof_scope ps(c->ctx, CTS_PROF_FINAL_UPDATE);
    CTS_TRY(cts_ssprk_stage(c->ctx, c->ft, c->n, nullptr, u, nullptr, c->U_exp, nullptr, u, dt, 0, b, omb, inc_imp.n,
                            inc_imp.coef, inc_imp.term, native, 1));
  } else if (inc_imp.n > 0) {
    CTS_TRY(assign_with_increments(c, u, nullptr, 1.0, u, inc_imp, Increment{}, CTS_PROF_FINAL_UPDATE));
  }
  CTS_TRY(call_dss(c, u, t_final));
  if (needs_update(f.update_constrain_state, CTS_SIG_END_OF_STEP)) CTS_TRY(call_state(c, f.constrain_state, u, t_final));
  if (needs_update(f.update_cache, CTS_SIG_END_OF_STEP)) CTS_TRY(call_state(c, f.cache, u, t_final));
  return CTS_OK;
}

void imex_free(cts_imex_cache* c) {
  auto fr = [&](void* p) {
    if (p) cts_free(c->ctx, p);
  };
  fr(c->U);
  fr(c->temp);
  fr(c->U_exp);
  fr(c->U_lim);
  fr(c->T_exp1);
  fr(c->T_lim1);
  for (int i = 0; i < CTS_MAX_STAGES; ++i) {
    fr(c->T_exp[i]);
    fr(c->T_lim[i]);
    fr(c->T_imp[i]);
  }
  if (c->nm) cts_newton_cache_destroy(c->nm);
}

}  // namespace

extern "C" {

int cts_imex_cache_create(cts_ctx* ctx, cts_dtype ft, size_t n, const cts_imex_tableau* tab, int constraint,
                          const cts_ode_function* f, const cts_newton_config* newton, cts_imex_cache** out) {
  CTS_CHECK_CTX(ctx);
  if (!tab || !f || !out || n == 0 || tab->s < 1 || tab->s > CTS_MAX_STAGES)
    return cts_fail(ctx, CTS_EINVAL, "cts_imex_cache_create: bad argument");
  const int s = tab->s;
  // IMEXTableau asserts (imex_tableaus.jl:71-72): a_exp strictly lower, a_imp lower triangular
  for (int i = 0; i < s; ++i)
    for (int j = i; j < s; ++j)
      if (A(tab->a_exp, s, i, j) != 0 || (j > i && A(tab->a_imp, s, i, j) != 0))
        return cts_fail(ctx, CTS_EINVAL, "IMEXTableau: a_exp must be strictly lower and a_imp lower triangular");
  auto* c = new cts_imex_cache();
  c->ctx = ctx;
  c->ft = ft;
  c->n = n;
  c->tab = *tab;
  c->constraint = constraint;
  c->f = *f;
  if (newton) c->ncfg = *newton;
  auto fail = [&](int code, const char* msg) {
    imex_free(c);
    delete c;
    return cts_fail(ctx, code, msg);
  };
  // γ detection (imex_ark.jl:49-51) + check_sdirk_compat (async_utils.jl:93-98)
  {
    double uniq[CTS_MAX_STAGES];
    int nu = 0;
    for (int i = 0; i < s; ++i) {
      double g = A(tab->a_imp, s, i, i);
      if (g == 0) continue;
      bool seen = false;
      for (int k = 0; k < nu; ++k) seen = seen || uniq[k] == g;
      if (!seen) uniq[nu++] = g;
    }
    c->has_gamma = nu == 1;
    c->gamma = nu == 1 ? uniq[0] : 0.0;
    if (!c->has_gamma && c->ncfg.present && fires_on_new_timestep(c->ncfg.update_j))
      return fail(CTS_EINVAL, "the tableau has implicit stages with distinct coefficients (it is not SDIRK): do not update on the NewTimeStep signal");
  }
  int r = CTS_OK;
  auto need = [&](void** p) {
    if (r == CTS_OK) r = alloc_state(c, p);
  };
  if (constraint == 0) {
    need(&c->U);
    need(&c->temp);
    for (int i = 0; i < s; ++i) {
      if (col_nonzero(tab->a_exp, s, i) || tab->b_exp[i] != 0) {
        need(&c->T_exp[i]);
        if (f->has_lim) need(&c->T_lim[i]);  // (the reference allocates these even without a limiter, imex_ark.jl:45)
      }
      if (col_nonzero(tab->a_imp, s, i) || tab->b_imp[i] != 0) need(&c->T_imp[i]);
    }
  } else {
    // Shu-Osher β and the low-storage check (imex_ssprk.jl:48-65)
    double ahat[(CTS_MAX_STAGES + 1) * CTS_MAX_STAGES];
    for (int i = 0; i < s; ++i)
      for (int j = 0; j < s; ++j) ahat[i * s + j] = A(tab->a_exp, s, i, j);
    for (int j = 0; j < s; ++j) ahat[s * s + j] = tab->b_exp[j];
    for (int i = 0; i < s; ++i) c->beta[i] = ahat[(i + 1) * s + i];
    for (int i = 0; i < s; ++i) {
      double prod = 1.0;
      for (int k = i; k < s; ++k) {
        prod = (k == i) ? c->beta[k] : prod * c->beta[k];
        if (ahat[(k + 1) * s + i] != prod)
          return fail(CTS_EINVAL, "The SSP IMEXAlgorithm only supports a low-storage IMEX SSPRK tableau (imex_ssprk.jl:51-64)");
      }
    }
    need(&c->U);
    need(&c->U_exp);
    need(&c->T_exp1);
    need(&c->temp);
    if (f->has_lim) {
      need(&c->U_lim);
      need(&c->T_lim1);
    }
    for (int i = 0; i < s; ++i)
      if (col_nonzero(tab->a_imp, s, i) || tab->b_imp[i] != 0) need(&c->T_imp[i]);
  }
  if (r != CTS_OK) {
    imex_free(c);
    delete c;
    return r;
  }
  if (f->T_imp && c->ncfg.present) {
    const bool has_jac = f->Wfact != nullptr && f->W != nullptr;
    r = cts_newton_cache_create(ctx, ft, n, &c->ncfg, has_jac ? f->W : nullptr, f->ldiv, f->jac_mul, f->ud, &c->nm);
    if (r != CTS_OK) {
      std::string msg = ctx->err;
      imex_free(c);
      delete c;
      return cts_fail(ctx, r, msg);
    }
    c->nbuffers += 2;  // Δx, f
  }
  // downcast_tableau (imex_tableaus.jl:127-133, imex_ark.jl:59-61, imex_ssprk.jl:75-78): with cast_tableau_to_state_eltype
  // and a Float32 state every coefficient is rounded to Float32 -- AFTER γ and β were formed from the tableau as given
  if (c->tab.cast_to_state_eltype && ft == CTS_F32) {
    auto rnd = [](double* a, int m) {
      for (int k = 0; k < m; ++k) a[k] = (double)(float)a[k];
    };
    rnd(c->tab.a_exp, CTS_MAX_STAGES * CTS_MAX_STAGES);
    rnd(c->tab.a_imp, CTS_MAX_STAGES * CTS_MAX_STAGES);
    rnd(c->tab.b_exp, CTS_MAX_STAGES);
    rnd(c->tab.b_imp, CTS_MAX_STAGES);
    rnd(c->tab.c_exp, CTS_MAX_STAGES);
    rnd(c->tab.c_imp, CTS_MAX_STAGES);
  }
  *out = c;
  return CTS_OK;
}

int cts_imex_cache_destroy(cts_imex_cache* c) {
  if (!c) return CTS_EINVAL;
  if (c->graph_exec) {
    if (cts_ctx_alive(c->ctx)) cudaStreamSynchronize(c->ctx->stream);
    cudaGraphExecDestroy(c->graph_exec);
  }
  imex_free(c);
  delete c;
  return CTS_OK;
}

// A step may be captured when everything it enqueues is the library's own stream-ordered work: the fused column path
// (all hooks are library operators, direct Newton with max_iters = 1: no host-visible scalars, no user callbacks).
static bool graph_capturable(cts_imex_cache* c, const void* u) {
  if (!cts_aligned16(u)) return false;
  if (c->graph_mode == -1) c->graph_mode = getenv("CTS_NO_GRAPH") ? 0 : 1;
  if (c->graph_mode != 1 || c->constraint != 0) return false;
  if (c->ctx->stream == nullptr || c->ctx->stream == cudaStreamLegacy || c->ctx->stream == cudaStreamPerThread) return false;  // not capturable
  cts_op_column* op = fusable_column_op(c);
  if (!op) return false;
  const cts_ode_function& f = c->f;
  if (f.T_exp_T_lim && f.T_exp_T_lim != cts_op_column_T_exp) return false;
  if (f.dss && f.dss != cts_op_column_dss) return false;
  if (f.lim) return false;
  const cts_update_handler* hs[3] = {&c->ncfg.update_j, &f.update_cache, &f.update_constrain_state};
  for (auto* h : hs)
    if (h->fn) return false;  // host handlers may do anything
  if (c->ctx->nranks > 1 && c->ctx->peer_mode != 1) return false;  // NCCL send/recv stays outside graphs
  cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(c->ctx->stream, &st) != cudaSuccess || st != cudaStreamCaptureStatusNone) return false;  // the caller captures
  return true;
}

static int graphed_ark_step(cts_imex_cache* c, void* u, double t, double dt) {
  cts_ctx* ctx = c->ctx;
  CTS_TRY(cts_ctx_ensure_comm_stream(ctx));  // created outside the capture
  if (cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeRelaxed) != cudaSuccess) {
    cudaGetLastError();
    c->graph_mode = 0;  // this stream cannot be captured: plain launches from now on
    return ark_step(c, u, t, dt);
  }
  ctx->capturing = true;
  const int r = ark_step(c, u, t, dt);
  ctx->capturing = false;
  cudaGraph_t graph = nullptr;
  const cudaError_t e = cudaStreamEndCapture(ctx->stream, &graph);
  if (r != CTS_OK) {
    if (graph) cudaGraphDestroy(graph);
    cudaGetLastError();
    return r;
  }
  if (e != cudaSuccess || !graph) {
    cudaGetLastError();
    return cts_fail(ctx, CTS_ECUDA, std::string("cudaStreamEndCapture: ") + cudaGetErrorString(e));
  }
  if (c->graph_exec) {
    cudaGraphExecUpdateResultInfo info;
    if (cudaGraphExecUpdate(c->graph_exec, graph, &info) != cudaSuccess) {  // topology changed (dt, policy, buffers)
      cudaGetLastError();
      cudaGraphExecDestroy(c->graph_exec);
      c->graph_exec = nullptr;
    }
  }
  if (!c->graph_exec) {
    const cudaError_t ei = cudaGraphInstantiate(&c->graph_exec, graph, 0);
    if (ei != cudaSuccess) {
      cudaGraphDestroy(graph);
      c->graph_exec = nullptr;
      return cts_fail(ctx, CTS_ECUDA, std::string("cudaGraphInstantiate: ") + cudaGetErrorString(ei));
    }
    c->graph_reinstantiations++;
  }
  cudaGraphDestroy(graph);
  CTS_CUDA(ctx, cudaGraphLaunch(c->graph_exec, ctx->stream));
  c->graph_launches++;
  return CTS_OK;
}

int cts_imex_step(cts_imex_cache* c, void* u, double t, double dt, int dt_is_f32) {
  if (!c || !u) return CTS_EINVAL;
  cts_nvtx_scope nvtx("step_u!");  // integrators.jl:442
  c->dt_f32 = dt_is_f32 != 0;
  if (c->constraint == 0) {
    if (graph_capturable(c, u)) {
      // collective set-up (IPC mailboxes) cannot happen inside a capture: do it now if the step will exchange halos
      return graphed_ark_step(c, u, t, dt);
    }
    return ark_step(c, u, t, dt);
  }
  return ssprk_step(c, u, t, dt, dt_is_f32);
}

int cts_imex_cache_set_graph(cts_imex_cache* c, int enabled) {
  if (!c) return CTS_EINVAL;
  c->graph_mode = enabled ? 1 : 0;
  return CTS_OK;
}

int cts_imex_cache_graph_stats(cts_imex_cache* c, uint64_t* launches, uint64_t* instantiations) {
  if (!c) return CTS_EINVAL;
  if (launches) *launches = c->graph_launches;
  if (instantiations) *instantiations = c->graph_reinstantiations;
  return CTS_OK;
}

int cts_imex_cache_buffer(cts_imex_cache* c, const char* name, void** dptr) {
  if (!c || !name || !dptr) return CTS_EINVAL;
  *dptr = nullptr;
  std::string s(name);
  auto idx = [&](const std::string& prefix) -> int {
    if (s.size() <= prefix.size() || s.compare(0, prefix.size(), prefix) != 0) return -1;
    int i = atoi(s.c_str() + prefix.size());
    return (i >= 1 && i <= CTS_MAX_STAGES) ? i - 1 : -1;
  };
  if (s == "U") *dptr = c->U;
  else if (s == "temp") *dptr = c->temp;
  else if (s == "U_exp") *dptr = c->U_exp;
  else if (s == "U_lim") *dptr = c->U_lim;
  else if (s == "T_exp") *dptr = c->T_exp1;
  else if (s == "T_lim") *dptr = c->T_lim1;
  else if (s == "dx") *dptr = c->nm ? c->nm->dx : nullptr;
  else if (s == "f") *dptr = c->nm ? c->nm->f : nullptr;
  else if (idx("T_exp") >= 0) *dptr = c->T_exp[idx("T_exp")];
  else if (idx("T_lim") >= 0) *dptr = c->T_lim[idx("T_lim")];
  else if (idx("T_imp") >= 0) *dptr = c->T_imp[idx("T_imp")];
  else return cts_fail(c->ctx, CTS_EINVAL, "unknown buffer name");
  return CTS_OK;
}

int cts_imex_cache_newton(cts_imex_cache* c, cts_newton_cache** out) {
  if (!c || !out) return CTS_EINVAL;
  *out = c->nm;
  return CTS_OK;
}

int cts_imex_cache_nbuffers(cts_imex_cache* c, int* n) {
  if (!c || !n) return CTS_EINVAL;
  *n = c->nbuffers;
  return CTS_OK;
}

int cts_imex_cache_stats(cts_imex_cache* c, cts_solver_stats* out) {
  if (!c || !out) return CTS_EINVAL;
  *out = c->stats;
  if (c->nm) {
    const cts_solver_stats& n = c->nm->stats;
    out->newton_iters += n.newton_iters;
    out->krylov_iters += n.krylov_iters;
    out->f_evals += n.f_evals;
    out->jvp_evals += n.jvp_evals;
    out->wfact_evals += n.wfact_evals;
    out->ldiv_calls += n.ldiv_calls;
  }
  return CTS_OK;
}

int cts_imex_cache_set_fusion(cts_imex_cache* c, int enabled) {
  if (!c) return CTS_EINVAL;
  c->fusion = enabled != 0;
  if (c->nm) c->nm->fuse_reductions = c->fusion;
  return CTS_OK;
}

}  // extern "C"

// =============================================================================================
// LSRK 2N   (lsrk.jl:49-67)
// =============================================================================================
struct cts_lsrk_cache {
  cts_ctx* ctx = nullptr;
  cts_dtype ft = CTS_F64;
  size_t n = 0;
  int nstages = 0;
  std::vector<double> A, B, C;
  cts_incr_fn f = nullptr;
  void* ud = nullptr;
  void* du = nullptr;
  void* scratch = nullptr;  // ping-pong partner of u for the fused out-of-place stage (lazy)
  bool fusion = true;
  cts_solver_stats stats{};
};

extern "C" {

int cts_lsrk_cache_create(cts_ctx* ctx, cts_dtype ft, size_t n, int nstages, const double* A, const double* B, const double* C,
                          cts_incr_fn f, void* ud, cts_lsrk_cache** out) {
  CTS_CHECK_CTX(ctx);
  if (!out || !A || !B || !C || !f || nstages < 1 || n == 0) return cts_fail(ctx, CTS_EINVAL, "cts_lsrk_cache_create: bad argument");
  auto* c = new cts_lsrk_cache();
  c->ctx = ctx;
  c->ft = ft;
  c->n = n;
  c->nstages = nstages;
  c->A.assign(A, A + nstages);
  c->B.assign(B, B + nstages);
  c->C.assign(C, C + nstages);
  c->f = f;
  c->ud = ud;
  int r = cts_alloc(ctx, n * cts_sizeof(ft), &c->du);  // `du = zero(prob.u0)` lsrk.jl:50
  if (r) {
    delete c;
    return r;
  }
  *out = c;
  return CTS_OK;
}

int cts_lsrk_cache_destroy(cts_lsrk_cache* c) {
  if (!c) return CTS_EINVAL;
  cts_free(c->ctx, c->du);
  if (c->scratch) cts_free(c->ctx, c->scratch);
  delete c;
  return CTS_OK;
}

int cts_lsrk_step(cts_lsrk_cache* c, void* u, double t, double dt, int dt_is_f32) {
  if (!c || !u) return CTS_EINVAL;
  cts_nvtx_scope nvtx("step_u!");  // integrators.jl:442
  cts_ctx* ctx = c->ctx;
  const bool f32 = c->ft == CTS_F32;
  const int native = (f32 && dt_is_f32) ? 1 : 0;
  auto dtB_of = [&](int s) { return native ? (double)((float)dt * (float)c->B[s]) : dt * c->B[s]; };
  auto time_of = [&](int s) { return native ? (double)((float)t + (float)c->C[s] * (float)dt) : t + c->C[s] * dt; };

  cts_op_advection* adv = nullptr;
  if (c->fusion && c->f == cts_op_advection_incr) {
    auto* op = static_cast<cts_op_advection*>(c->ud);
    if (op && cts_op_advection_ctx(op) == ctx && cts_op_advection_dtype(op) == c->ft && cts_op_advection_n(op) == c->n) adv = op;
  }
  int s0 = 0, s_inplace = -1;
  if (adv) {
    if (!c->scratch) CTS_TRY(cts_alloc(ctx, c->n * cts_sizeof(c->ft), &c->scratch));
    // Out-of-place fused stages ping-pong between u and scratch and must end in u (the caller's array,
    // integrators.jl:236): an even number of them; with an odd stage count the LAST stage runs in place on u (same
    // 4w B/DOF), or -- for a state the vector kernel cannot take -- the first stage runs un-fused.
    if (c->nstages % 2) {
      if (cts_advection_inplace_supported(adv, u, c->du))
        s_inplace = c->nstages - 1;
      else
        s0 = 1;
    }
  }
  for (int s = 0; s < c->nstages; ++s) {
    if (adv && s >= s0) {
      const bool from_u = ((s - s0) % 2) == 0;
      void* src = s == s_inplace ? u : (from_u ? u : c->scratch);
      void* dst = s == s_inplace ? u : (from_u ? c->scratch : u);
      cts_prof_scope ps(ctx, CTS_PROF_LSRK_FUSED);
      CTS_TRY(cts_advection_fused_stage(adv, dst, src, c->du, c->A[s], dtB_of(s), native));
      c->stats.fused_stage_launches++;
    } else {
      {
        cts_prof_scope ps(ctx, CTS_PROF_LSRK_RHS);
        HOOK(ctx, c->f(c->ud, c->du, u, time_of(s), 1.0, c->A[s]));  // lsrk.jl:64
      }
      c->stats.f_evals++;
      cts_prof_scope ps(ctx, CTS_PROF_LSRK_UPDATE);
      CTS_TRY(cts_lsrk_update(ctx, c->ft, c->n, u, dtB_of(s), c->du, native));  // lsrk.jl:65
    }
  }
  return CTS_OK;
}

int cts_lsrk_cache_buffer(cts_lsrk_cache* c, const char* name, void** dptr) {
  if (!c || !name || !dptr) return CTS_EINVAL;
  if (strcmp(name, "du") == 0) {
    *dptr = c->du;
    return CTS_OK;
  }
  return CTS_EINVAL;
}

int cts_lsrk_cache_set_fusion(cts_lsrk_cache* c, int enabled) {
  if (!c) return CTS_EINVAL;
  c->fusion = enabled != 0;
  return CTS_OK;
}

int cts_lsrk_cache_stats(cts_lsrk_cache* c, cts_solver_stats* out) {
  if (!c || !out) return CTS_EINVAL;
  *out = c->stats;
  return CTS_OK;
}

}  // extern "C"

// =============================================================================================
// RosenbrockAlgorithm (src/solvers/rosenbrock.jl:85-239)
// =============================================================================================
struct cts_rosenbrock_cache {
  cts_ctx* ctx = nullptr;
  cts_dtype ft = CTS_F64;
  size_t n = 0;
  cts_rosenbrock_tableau tab{};
  cts_ode_function f{};
  cts_tendency_fn tgrad = nullptr;
  void* U = nullptr;
  void* fU = nullptr;
  void* fU_imp = nullptr;
  void* fU_exp = nullptr;
  void* fU_lim = nullptr;
  void* dYdt = nullptr;
  void* k[CTS_MAX_STAGES] = {};
  bool fusion = true;
  cts_solver_stats stats{};
};

namespace {

int ros_state(cts_rosenbrock_cache* c, cts_state_fn fn, void* u, double t, int cat = CTS_PROF_OTHER_HOOKS) {
  if (fn) {
    cts_prof_scope ps(c->ctx, cat);
    HOOK(c->ctx, fn(c->f.ud, u, t));
  }
  return CTS_OK;
}

void ros_free(cts_rosenbrock_cache* c) {
  void* bufs[] = {c->U, c->fU, c->fU_imp, c->fU_exp, c->fU_lim, c->dYdt};
  for (void* b : bufs)
    if (b) cts_free(c->ctx, b);
  for (int i = 0; i < CTS_MAX_STAGES; ++i)
    if (c->k[i]) cts_free(c->ctx, c->k[i]);
}

}  // namespace

extern "C" {

int cts_rosenbrock_cache_create(cts_ctx* ctx, cts_dtype ft, size_t n, const cts_rosenbrock_tableau* tab,
                                const cts_ode_function* f, cts_tendency_fn tgrad, cts_rosenbrock_cache** out) {
  CTS_CHECK_CTX(ctx);
  if (!tab || !f || !out || tab->s < 1 || tab->s > CTS_MAX_STAGES) return cts_fail(ctx, CTS_EINVAL, "cts_rosenbrock_cache_create: bad arguments");
  if (f->T_imp && (!f->Wfact || !f->ldiv || !f->W)) return cts_fail(ctx, CTS_EINVAL, "Rosenbrock: T_imp! needs Wfact, ldiv! and a jac_prototype");
  auto* c = new cts_rosenbrock_cache();
  c->ctx = ctx;
  c->ft = ft;
  c->n = n;
  c->tab = *tab;
  c->f = *f;
  c->tgrad = f->T_imp ? tgrad : nullptr;
  const size_t bytes = n * cts_sizeof(ft);
  void** bufs[] = {&c->U, &c->fU, &c->fU_imp, &c->fU_exp, &c->fU_lim, &c->dYdt};  // rosenbrock.jl:100-108
  int r = CTS_OK;
  for (void** b : bufs)
    if (r == CTS_OK) r = cts_alloc(ctx, bytes, b);
  for (int i = 0; i < tab->s && r == CTS_OK; ++i) r = cts_alloc(ctx, bytes, &c->k[i]);
  if (r != CTS_OK) {
    ros_free(c);
    delete c;
    return r;
  }
  *out = c;
  return CTS_OK;
}

int cts_rosenbrock_cache_destroy(cts_rosenbrock_cache* c) {
  if (!c) return CTS_EINVAL;
  ros_free(c);
  delete c;
  return CTS_OK;
}

int cts_rosenbrock_cache_buffer(cts_rosenbrock_cache* c, const char* name, void** dptr) {
  if (!c || !name || !dptr) return CTS_EINVAL;
  std::string s(name);
  *dptr = nullptr;
  if (s == "U") *dptr = c->U;
  else if (s == "fU") *dptr = c->fU;
  else if (s == "fU_imp") *dptr = c->fU_imp;
  else if (s == "fU_exp") *dptr = c->fU_exp;
  else if (s == "fU_lim") *dptr = c->fU_lim;
  else if (s == "dYdt") *dptr = c->dYdt;
  else if (s.size() > 1 && s[0] == 'k') {
    int i = atoi(s.c_str() + 1);
    if (i >= 1 && i <= c->tab.s) *dptr = c->k[i - 1];
  }
  return *dptr ? CTS_OK : cts_fail(c->ctx, CTS_EINVAL, "unknown buffer name");
}

int cts_rosenbrock_cache_set_fusion(cts_rosenbrock_cache* c, int enabled) {
  if (!c) return CTS_EINVAL;
  c->fusion = enabled != 0;
  return CTS_OK;
}

int cts_rosenbrock_cache_stats(cts_rosenbrock_cache* c, cts_solver_stats* out) {
  if (!c || !out) return CTS_EINVAL;
  *out = c->stats;
  return CTS_OK;
}

// step_u!(int, cache::RosenbrockCache)  rosenbrock.jl:136-239.  The reference's chains of `.+=` broadcasts become one
// chained pass each (cts_axpy_chain: same operations, same order, same roundings); terms whose coefficient is exactly
// zero are not read (the reference multiplies them by zero).
int cts_rosenbrock_step(cts_rosenbrock_cache* c, void* u, double t, double dt) {
  if (!c || !u) return CTS_EINVAL;
  cts_nvtx_scope nvtx("step_u!");  // integrators.jl:442
  cts_ctx* ctx = c->ctx;
  const cts_rosenbrock_tableau& tab = c->tab;
  const cts_ode_function& f = c->f;
  const int s = tab.s;
  const double dtg = dt * tab.Gamma[0];  // "only valid when Γ[i, i] is constant" :151-153
  // Library column operator as T_imp!/Wfact/ldiv! (no ∂Y∂t, no limiter): T_imp!, the `.+=` chain of the right-hand side and
  // ldiv! of a stage are ONE tile kernel (cts_column_rosenbrock_stage; bit-identical to the three passes).
  cts_op_column* rop = nullptr;
  if (c->fusion && f.T_imp == cts_op_column_T_imp && f.Wfact == cts_op_column_Wfact && f.ldiv == cts_op_column_ldiv && !c->tgrad &&
      !f.has_lim && f.ud && f.W) {
    auto* op = static_cast<cts_op_column*>(f.ud);
    if (cts_op_column_ctx(op) == ctx && cts_op_column_dtype(op) == c->ft && cts_op_column_ndof(op) == c->n &&
        cts_column_rhs_solve_fusable(op))
      rop = op;
  }
  if (f.T_imp) {
    cts_prof_scope ps(ctx, CTS_PROF_WFACT);
    HOOK(ctx, f.Wfact(f.ud, f.W, u, dtg, t));
    c->stats.wfact_evals++;
  }
  if (c->tgrad) {
    cts_prof_scope ps(ctx, CTS_PROF_OTHER_HOOKS);
    HOOK(ctx, c->tgrad(f.ud, c->dYdt, u, t));
  }
  for (int i = 0; i < s; ++i) {
    double ai = 0.0, gi = 0.0;
    for (int j = 0; j < i; ++j) ai += tab.alpha[i * s + j];
    for (int j = 0; j <= i; ++j) gi += tab.Gamma[i * s + j];
    const double ti = t + ai * dt;
    double coef[2 * CTS_MAX_STAGES + 4];
    const void* terms[2 * CTS_MAX_STAGES + 4];
    // `U .= u; U .+= A[i, j] .* k[j]`  :178-181  (stage 1: U is only read, so it aliases u)
    void* U = c->U;
    int nt = 0;
    for (int j = 0; j < i; ++j)
      if (tab.A[i * s + j] != 0) {
        coef[nt] = tab.A[i * s + j];
        terms[nt++] = c->k[j];
      }
    if (i == 0) {
      U = u;
    } else {
      cts_prof_scope ps(ctx, CTS_PROF_STAGE_ASSEMBLY);
      CTS_TRY(cts_axpy_chain(ctx, c->ft, c->n, U, u, nt, coef, terms, 1.0, 0));
    }
    if (i != 0) {
      if (f.dss) {
        CTS_TRY(ros_state(c, f.dss, U, ti, CTS_PROF_DSS));
        c->stats.dss_calls++;
      }
      if (needs_update(f.update_constrain_state, CTS_SIG_END_OF_STAGE)) CTS_TRY(ros_state(c, f.constrain_state, U, ti));
      if (needs_update(f.update_cache, CTS_SIG_END_OF_STAGE)) CTS_TRY(ros_state(c, f.cache, U, ti));
    }
    nt = 0;
    if (rop) {
      if (f.T_exp_T_lim) {
        cts_prof_scope ps(ctx, CTS_PROF_T_EXP);
        HOOK(ctx, f.T_exp_T_lim(f.ud, c->fU_exp, c->fU_lim, U, ti));
      }
      for (int j = 0; j < i; ++j)
        if (tab.C[i * s + j] != 0) {
          coef[nt] = tab.C[i * s + j] / dt;
          terms[nt++] = c->k[j];
        }
      int r;
      {
        cts_prof_scope ps(ctx, CTS_PROF_FUSED_STAGE);
        r = cts_column_rosenbrock_stage(rop, c->k[i], U, f.T_exp_T_lim ? c->fU_exp : nullptr, nt, coef, terms, dtg,
                                        static_cast<const cts_tridiag*>(f.W));
      }
      if (r == CTS_OK) {
        c->stats.ldiv_calls++;
        c->stats.fused_stage_launches++;
        continue;
      }
      if (r != CTS_EUNSUPPORTED) return r;
      nt = 0;  // unaligned buffers or too many terms: the generic passes below (T_exp! is simply evaluated again)
    }
    if (f.T_imp) {
      cts_prof_scope ps(ctx, CTS_PROF_T_IMP);
      HOOK(ctx, f.T_imp(f.ud, c->fU_imp, U, ti));
      coef[nt] = 1.0;
      terms[nt++] = c->fU_imp;
    }
    if (f.T_exp_T_lim) {
      cts_prof_scope ps(ctx, CTS_PROF_T_EXP);
      HOOK(ctx, f.T_exp_T_lim(f.ud, c->fU_exp, c->fU_lim, U, ti));
      coef[nt] = 1.0;
      terms[nt++] = c->fU_exp;
      if (f.has_lim) {
        coef[nt] = 1.0;
        terms[nt++] = c->fU_lim;
      }
    }
    if (c->tgrad) {
      coef[nt] = gi * dt;  // `γi .* float(dt) .* ∂Y∂t`
      terms[nt++] = c->dYdt;
    }
    for (int j = 0; j < i; ++j)
      if (tab.C[i * s + j] != 0) {
        coef[nt] = tab.C[i * s + j] / dt;
        terms[nt++] = c->k[j];
      }
    void* rhs = f.T_imp ? c->fU : c->k[i];
    {
      // `fill!(fU, 0); fU .+= ...; fU .*= -dtγ`  :164,198-214 ; without T_imp! `k[i] .= .-fU` folds into the scale
      cts_prof_scope ps(ctx, CTS_PROF_NEWTON_EW);
      if (f.T_imp) {
        CTS_TRY(cts_axpy_chain(ctx, c->ft, c->n, rhs, nullptr, nt, coef, terms, -dtg, 1));
      } else {
        CTS_TRY(cts_axpy_chain(ctx, c->ft, c->n, c->fU, nullptr, nt, coef, terms, -dtg, 1));
        double m1[1] = {-1.0};
        const void* t1[1] = {c->fU};
        CTS_TRY(cts_axpy_chain(ctx, c->ft, c->n, c->k[i], nullptr, 1, m1, t1, 1.0, 0));
      }
    }
    if (f.T_imp) {
      cts_prof_scope ps(ctx, CTS_PROF_LDIV);
      HOOK(ctx, f.ldiv(f.ud, c->k[i], f.W, c->fU));
      c->stats.ldiv_calls++;
    }
  }
  {
    // `u .+= m[i] .* k[i]`  :231-233
    double coef[CTS_MAX_STAGES];
    const void* terms[CTS_MAX_STAGES];
    int nt = 0;
    for (int i = 0; i < s; ++i)
      if (tab.m[i] != 0) {
        coef[nt] = tab.m[i];
        terms[nt++] = c->k[i];
      }
    cts_prof_scope ps(ctx, CTS_PROF_FINAL_UPDATE);
    CTS_TRY(cts_axpy_chain(ctx, c->ft, c->n, u, u, nt, coef, terms, 1.0, 0));
  }
  if (f.dss) {
    CTS_TRY(ros_state(c, f.dss, u, t + dt, CTS_PROF_DSS));
    c->stats.dss_calls++;
  }
  if (needs_update(f.update_constrain_state, CTS_SIG_END_OF_STEP)) CTS_TRY(ros_state(c, f.constrain_state, u, t + dt));
  if (needs_update(f.update_cache, CTS_SIG_END_OF_STEP)) CTS_TRY(ros_state(c, f.cache, u, t + dt));
  return CTS_OK;
}

}  // extern "C"
