// Internal declarations shared by the translation units of libcts_sm100a.so.
#pragma once
#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>

#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

#include "../../include/cts.h"

struct cts_peer_halo;  // comm.cu

struct cts_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  int num_sms = 148;
  std::string err;
  uint64_t launches = 0;
  uint64_t bytes = 0;  // algorithmic HBM bytes of the kernels launched so far: (arrays read + arrays written) x n x sizeof(FT)
  // reduction scratch (reduce.cu): per-chunk partial sums, [k][n_chunks_global] (grown on demand)
  double* d_partials = nullptr;   // this rank's chunk partials (entries of chunks owned by other ranks stay zero)
  double* d_partials_g = nullptr; // all ranks' chunk partials after the exchange (multi-rank layouts only)
  size_t partials_cap = 0;        // capacity of each of the two arrays, in doubles
  double* d_result = nullptr;     // [kMaxMulti]
  unsigned int* d_counter = nullptr;
  double* h_result = nullptr;     // pinned [kMaxMulti]
  bool red_window = false;        // reductions cover [red_offset, red_offset + red_count) of every vector
  size_t red_offset = 0, red_count = 0;
  // shard-invariant layout (cts_set_reduction_layout): the window is the slice [red_g0, red_g0 + red_count) of a
  // global index space of red_G elements that is cut into fixed chunks of red_chunk elements
  bool red_layout = false;
  size_t red_g0 = 0, red_G = 0;
  unsigned red_chunk = 0;         // 0 -> kReduceChunk
  // device-side pointer table scratch for kernels taking many operands
  // communicator (comm.cpp)
  void* nccl_comm = nullptr;
  int rank = 0, nranks = 1;
  struct cts_peer_halo* peer = nullptr;        // NVLink peer-memory mailboxes of the halo exchange (comm.cu)
  int peer_mode = 0;                           // 0: not set up, 1: peer-memory halo kernel, -1: NCCL send/recv
  cudaStream_t comm_stream = nullptr;          // halo exchanges that overlap compute run here (created on first use)
  cudaEvent_t ev_compute = nullptr, ev_comm = nullptr;
  bool capturing = false;  // the context's stream is being captured into a CUDA graph (cts_imex_step)
  // per-kernel timing (cts_prof_*)
  bool prof_on = false;
  struct ProfRec {
    int cat;
    cudaEvent_t a, b;
  };
  std::vector<ProfRec> prof_recs;
  std::vector<cudaEvent_t> prof_pool;
};

// RAII bracket: records a start/stop event pair on the context's stream around a kernel family.
// NVTX ranges with the reference's names (`NVTX.@annotate` on solve!, step_u!, solve_newton!, solve_krylov!:
// integrators.jl:345,442, newtons_method.jl:465,609).  nvtx3 is header-only: without a profiler attached a range is one
// load and a predictable branch.
struct cts_nvtx_scope {
  explicit cts_nvtx_scope(const char* name) { nvtxRangePushA(name); }
  ~cts_nvtx_scope() { nvtxRangePop(); }
  cts_nvtx_scope(const cts_nvtx_scope&) = delete;
  cts_nvtx_scope& operator=(const cts_nvtx_scope&) = delete;
};

struct cts_prof_scope {
  cts_ctx* ctx;
  int idx = -1;
  cts_prof_scope(cts_ctx* c, int cat) : ctx(c) {
    if (!c || !c->prof_on) return;
    auto get = [&]() {
      cudaEvent_t e;
      if (!c->prof_pool.empty()) {
        e = c->prof_pool.back();
        c->prof_pool.pop_back();
      } else {
        cudaEventCreate(&e);
      }
      return e;
    };
    cts_ctx::ProfRec r{cat, get(), get()};
    cudaEventRecordWithFlags(r.a, c->stream, c->capturing ? cudaEventRecordExternal : cudaEventRecordDefault);
    c->prof_recs.push_back(r);
    idx = (int)c->prof_recs.size() - 1;
  }
  ~cts_prof_scope() {
    if (idx >= 0)
      cudaEventRecordWithFlags(ctx->prof_recs[idx].b, ctx->stream, ctx->capturing ? cudaEventRecordExternal : cudaEventRecordDefault);
  }
};

constexpr int kReduceChunk = 8192;   // default elements per reduction chunk (one CTA per chunk): part of the result's definition
constexpr int kReduceThreads = 256;
constexpr int kMaxMulti = 8;

#define CTS_CHECK_CTX(ctx) \
  do {                     \
    if (!(ctx)) return CTS_EINVAL; \
  } while (0)

bool cts_ctx_alive(const cts_ctx* ctx);  // false once cts_ctx_destroy has run (ctx.cu: live-context registry)

inline int cts_fail(cts_ctx* ctx, int code, const std::string& msg) {
  if (ctx) ctx->err = msg;
  return code;
}

#define CTS_CUDA(ctx, expr)                                                                      \
  do {                                                                                           \
    cudaError_t _e = (expr);                                                                     \
    if (_e != cudaSuccess)                                                                       \
      return cts_fail((ctx), CTS_ECUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));     \
  } while (0)

#define CTS_LAUNCH_CHECK(ctx)                                                                    \
  do {                                                                                           \
    (ctx)->launches++;                                                                           \
    cudaError_t _e = cudaGetLastError();                                                         \
    if (_e != cudaSuccess)                                                                       \
      return cts_fail((ctx), CTS_ECUDA, std::string("kernel launch: ") + cudaGetErrorString(_e) + \
                                            " at " + __FILE__ + ":" + std::to_string(__LINE__)); \
  } while (0)

// algorithmic bytes of a pass over `arrays` state-sized arrays of n elements of w bytes (SURVEY §8d's per-kernel figures)
#define CTS_COUNT_BYTES(ctx, arrays, n, w) ((ctx)->bytes += (uint64_t)(arrays) * (uint64_t)(n) * (uint64_t)(w))

#define CTS_TRY(expr)        \
  do {                       \
    int _r = (expr);         \
    if (_r != 0) return _r;  \
  } while (0)

// length(x) of a distributed vector (newtons_method.jl:135,445): the global domain of the layout, else rank-local * ranks
inline size_t cts_global_length(const cts_ctx* ctx, size_t n) {
  if (ctx->red_window && ctx->red_layout) return ctx->red_G;
  return (ctx->red_window ? ctx->red_count : n) * (size_t)ctx->nranks;
}
int cts_allreduce_sum_f64_oop(cts_ctx* ctx, const double* send, double* recv, size_t count);

inline size_t cts_sizeof(cts_dtype ft) { return ft == CTS_F64 ? 8 : 4; }
inline bool cts_aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

int cts_newton_residual_ex(cts_ctx* ctx, cts_dtype ft, size_t n, void* r, const void* U0, double dtgamma, const void* Up, int native);
int cts_timp_from_stage_ex(cts_ctx* ctx, cts_dtype ft, size_t n, void* T, const void* U, const void* temp, double dtgamma, int native);
// K5 + K6 in one pass: U = U - dx ; T_imp = (U_new - U_old)/dtγ  (one-iteration Newton, nothing between K5 and K6)
int cts_newton_update_timp(cts_ctx* ctx, cts_dtype ft, size_t n, void* U, void* T_imp, const void* dx, double dtgamma, int native);
// ---- internal entry points used across translation units ------------------------------------
// fused LSRK stage with the advection stencil (ops.cu); out-of-place in u.
int cts_advection_fused_stage(cts_op_advection* op, void* u_out, const void* u_in, void* du, double A, double dtB,
                              int compute_native);
// fused implicit column stage (ops.cu):  U0 = K1(base, terms) ; Newton(1 iter, direct) ; writes U, T_imp_i [, temp]
struct cts_fused_stage_args {
  const void* base;
  double c0;
  int n_grp1, n_grp2;
  double coef[CTS_MAX_TERMS];
  const void* terms[CTS_MAX_TERMS];
  int compute_native;
  double dtgamma;
  const cts_tridiag* W;  // stored W (ignored when matrix free)
  int emit_W;            // stored W is stale: regenerate it in the kernel for dtgamma and write it to W (no read)
  void* U;               // out: stage value
  void* T_imp;           // out: (U - U0)/dtγ, may be null
  void* temp;            // out (optional): U0
};
int cts_column_fused_stage(cts_op_column* op, const cts_fused_stage_args* a);
bool cts_column_fusable(const cts_op_column* op);
// final update of an ARK step fused with the last stage's T_exp! (column operator recognised as T_exp_T_lim!)
struct cts_final_update_args {
  void* u;                 // in/out: state (base of the update)
  const void* U_last;      // last stage value, already dss!-ed
  double t_exp_last;       // t + dt*c_exp[s]
  double coef_last;        // dt*b_exp[s]
  int n_grp1, n_grp2;      // other explicit terms (j < s), implicit terms
  double coef[CTS_MAX_TERMS];
  const void* terms[CTS_MAX_TERMS];
  int compute_native;
};
bool cts_column_final_fusable(const cts_op_column* op);
bool cts_op_column_has_halo(const cts_op_column* op);
// dss!(U) followed by T_exp!(du, U): the halo exchange runs on the communication stream while the rows that do not touch
// a ghost row are evaluated; the boundary strips follow once the ghosts have arrived.  Same results as the two hooks.
bool cts_column_overlap_ok(const cts_op_column* op);
int cts_column_dss_T_exp_overlapped(cts_op_column* op, void* du_exp, void* U, double t_dss, double t_exp);
int cts_halo_exchange_on(cts_ctx* ctx, cudaStream_t stream, cts_dtype ft, size_t count, const void* send_lo,
                         const void* send_hi, void* recv_lo, void* recv_hi);
int cts_column_fused_final_update(cts_op_column* op, const cts_final_update_args* a);
bool cts_column_final_update_supported(const cts_op_column* op, const cts_final_update_args* a);
// dss!(U_last); final update; dss!(u) with both halo exchanges and the boundary strips on the communication stream
int cts_column_final_update_overlapped(cts_op_column* op, const cts_final_update_args* a);
int cts_ctx_ensure_comm_stream(cts_ctx* ctx);
// Newton residual (dx == NULL) or forward-difference JVP of the stage equation for the column operator in one pass
bool cts_column_residual_fusable(const cts_op_column* op);
int cts_column_fused_residual(cts_op_column* op, void* out, const void* x, const void* dx, double eps, const void* U0,
                              const void* f, double dtgamma);
// right-hand side + column solve in one pass (ops.cu): the left-preconditioned Jacobian-free GMRES matvec
//   q = W \ ((f(x + eps v) - f(x))/eps)   and the Rosenbrock stage   k_i = W \ (-dtγ (T_imp(U) [+ fU_exp] + sum_j coef_j terms_j))
bool cts_column_rhs_solve_fusable(const cts_op_column* op);
bool cts_column_krylov_matrix_free(const cts_op_column* op);
int cts_column_precond_matrix_free(cts_op_column* op, void* q, const void* x, const void* b, double dtg);
int cts_column_jvp_precond(cts_op_column* op, void* q, const void* x, const void* v, double eps, const void* U0, const void* f,
                           double dtgamma, const cts_tridiag* W);
int cts_column_rosenbrock_stage(cts_op_column* op, void* k_out, const void* U, const void* fU_exp, int nt, const double* coef,
                                const void* const* terms, double dtgamma, const cts_tridiag* W);
bool cts_op_column_is_linear(const cts_op_column* op);
bool cts_op_column_matrix_free(const cts_op_column* op);
cts_ctx* cts_op_column_ctx(const cts_op_column* op);
cts_ctx* cts_op_advection_ctx(const cts_op_advection* op);
cts_dtype cts_op_column_dtype(const cts_op_column* op);
cts_dtype cts_op_advection_dtype(const cts_op_advection* op);
size_t cts_op_advection_n(const cts_op_advection* op);
bool cts_advection_inplace_supported(const cts_op_advection* op, const void* u, const void* du);  // u_out == u_in allowed
