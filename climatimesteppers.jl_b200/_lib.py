"""ctypes binding of ``libcts_sm100a.so`` (C ABI declared in ``include/cts.h``).

This is the tested stand-in for the Julia ``ccall`` binding shown in INTEGRATION.md: every call
below has a one-line Julia equivalent.  There is no CPU fallback -- if the shared library is
missing or no GPU is present the product fails loudly.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("CTS_LIB_PATH") or os.path.join(_HERE, "libcts_sm100a.so")  # override for A/B builds

CTS_F32, CTS_F64 = 0, 1
CTS_MAX_STAGES, CTS_MAX_TERMS = 8, 16
(SIG_NEW_TIMESTEP, SIG_NEW_NEWTON_SOLVE, SIG_NEW_NEWTON_ITERATION, SIG_WITH_DSS, SIG_END_OF_STAGE,
 SIG_END_OF_STEP) = range(6)
STATUS = {0: "CTS_OK", -1: "CTS_EINVAL", -2: "CTS_ECUDA", -3: "CTS_ENCCL", -4: "CTS_ENODEVICE",
          -5: "CTS_ECALLBACK", -6: "CTS_EUNSUPPORTED"}

PROF_CATEGORIES = ["stage_assembly", "final_update", "fused_stage", "T_exp", "T_imp", "Wfact", "ldiv", "newton_ew", "dss",
                   "krylov", "lsrk_fused", "lsrk_rhs", "lsrk_update", "other_hooks"]
c_void_pp = C.POINTER(C.c_void_p)
c_double_p = C.POINTER(C.c_double)

# ---- hook typedefs (include/cts.h "user hooks") ---------------------------------------------
TENDENCY_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double)
EXP_LIM_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double)
WFACT_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_double)
LDIV_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p)
STATE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_double)
INIT_IMP_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_double)
LIM_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_double, C.c_void_p)
INCR_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_double)
UPDATE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int, C.c_double)
RESIDUAL_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_void_p)
JAC_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_void_p)
PREPARE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p)


class UpdateHandler(C.Structure):
    _fields_ = [("every_signal", C.c_int), ("fn", UPDATE_FN), ("ud", C.c_void_p)]


class OdeFunction(C.Structure):
    _fields_ = [("ud", C.c_void_p), ("T_exp_T_lim", EXP_LIM_FN), ("has_lim", C.c_int), ("T_imp", TENDENCY_FN),
                ("Wfact", WFACT_FN), ("ldiv", LDIV_FN), ("jac_mul", LDIV_FN), ("W", C.c_void_p), ("lim", LIM_FN),
                ("dss", STATE_FN), ("constrain_state", STATE_FN), ("cache", STATE_FN), ("cache_imp", STATE_FN),
                ("initialize_imp", INIT_IMP_FN), ("initialize_imp_is_nothing", C.c_int),
                ("update_cache", UpdateHandler), ("update_constrain_state", UpdateHandler)]


class ImexTableau(C.Structure):
    _fields_ = [("s", C.c_int), ("a_exp", C.c_double * 64), ("b_exp", C.c_double * 8), ("c_exp", C.c_double * 8),
                ("a_imp", C.c_double * 64), ("b_imp", C.c_double * 8), ("c_imp", C.c_double * 8),
                ("cast_to_state_eltype", C.c_int)]


class RosenbrockTableau(C.Structure):
    _fields_ = [("s", C.c_int), ("A", C.c_double * 64), ("alpha", C.c_double * 64), ("C", C.c_double * 64),
                ("Gamma", C.c_double * 64), ("m", C.c_double * 8)]


COND_MAX_ERROR, COND_MAX_REL_ERROR, COND_MAX_ERROR_REDUCTION, COND_MIN_RATE, COND_ALL, COND_ANY = range(6)
MAX_COND_NODES = 16


class CondNode(C.Structure):
    _fields_ = [("kind", C.c_int), ("nchildren", C.c_int), ("p0", C.c_double), ("p1", C.c_double)]


class ConvergenceCheckerC(C.Structure):
    _fields_ = [("n_norm", C.c_int), ("norm_cond", CondNode * MAX_COND_NODES), ("n_comp", C.c_int),
                ("comp_cond", CondNode * MAX_COND_NODES), ("combiner_or", C.c_int)]


class NewtonConfig(C.Structure):
    _fields_ = [("present", C.c_int), ("max_iters", C.c_int), ("update_j", UpdateHandler), ("use_krylov", C.c_int),
                ("jacobian_free_jvp", C.c_int), ("jvp_step_kind", C.c_int), ("jvp_step_adjustment", C.c_double),
                ("disable_preconditioner", C.c_int), ("memory", C.c_int), ("itmax", C.c_int),
                ("forcing_kind", C.c_int), ("forcing_rtol", C.c_double), ("ew_initial_rtol", C.c_double),
                ("ew_gamma", C.c_double), ("ew_alpha", C.c_double), ("ew_min_rtol_threshold", C.c_double),
                ("ew_max_rtol", C.c_double), ("batch_gram_schmidt", C.c_int)]


class SolverStats(C.Structure):
    _fields_ = [(k, C.c_uint64) for k in ("newton_iters", "krylov_iters", "f_evals", "jvp_evals", "wfact_evals",
                                           "ldiv_calls", "dss_calls", "fused_stage_launches")]

    def as_dict(self):
        return {k: int(getattr(self, k)) for k, _ in self._fields_}


class Tridiag(C.Structure):
    _fields_ = [("nlev", C.c_int), ("ncols", C.c_size_t), ("dl", C.c_void_p), ("d", C.c_void_p), ("du", C.c_void_p)]


class ColumnDesc(C.Structure):
    _fields_ = [("nlev", C.c_int), ("nx", C.c_size_t), ("ny_local", C.c_size_t), ("halo", C.c_int), ("dz", C.c_double),
                ("nu", C.c_double), ("nonlinear", C.c_int), ("K_col", C.c_void_p), ("S_lev", C.c_void_p)]


# (name, restype, argtypes) for every symbol declared in include/cts.h
_P, _I, _D, _SZ, _U64 = C.c_void_p, C.c_int, C.c_double, C.c_size_t, C.c_uint64
SIGNATURES = [
    ("cts_version", C.c_char_p, []),
    ("cts_abi_version", _I, []),
    ("cts_ctx_create", _I, [c_void_pp, _I, _P]),
    ("cts_ctx_destroy", _I, [_P]),
    ("cts_ctx_sync", _I, [_P]),
    ("cts_last_error", C.c_char_p, [_P]),
    ("cts_ctx_stream", _P, [_P]),
    ("cts_launch_count", _I, [_P, C.POINTER(_U64)]),
    ("cts_nvtx_range_push", _I, [C.c_char_p]),
    ("cts_nvtx_range_pop", _I, []),
    ("cts_prof_enable", _I, [_P, _I]),
    ("cts_prof_read", _I, [_P, c_double_p, C.POINTER(_U64)]),
    ("cts_alloc", _I, [_P, _SZ, c_void_pp]),
    ("cts_free", _I, [_P, _P]),
    ("cts_memset_zero", _I, [_P, _P, _SZ]),
    ("cts_memcpy_h2d", _I, [_P, _P, _P, _SZ]),
    ("cts_memcpy_d2h_sync", _I, [_P, _P, _P, _SZ]),
    ("cts_memcpy_d2h", _I, [_P, _P, _P, _SZ]),
    ("cts_memcpy_d2d", _I, [_P, _P, _P, _SZ]),
    ("cts_axpy_n", _I, [_P, _I, _SZ, _P, _P, _D, _P, _I, _I, c_double_p, c_void_pp, _I]),
    ("cts_newton_cache_set_convergence_checker", _I, [_P, _P]),
    ("cts_newton_cache_set_line_search", _I, [_P, _I]),
    ("cts_newton_cache_is_converged", _I, [_P, _P, _P, _I, _P]),
    ("cts_imex_cache_newton", _I, [_P, _P]),
    ("cts_component_check", _I, [_P, _I, _SZ, _P, _P, _I, _P, _P, _I, _P, _P]),
    ("cts_component_update", _I, [_P, _I, _SZ, _P, _P, _I, _P, _P, _I]),
    ("cts_axpby_dot", _I, [_P, _I, _SZ, _D, _P, _D, _P, _P, _P]),
    ("cts_div_scalar_nrm2", _I, [_P, _I, _SZ, _P, _P, _D, _P]),
    ("cts_axpy_chain", _I, [_P, _I, _SZ, _P, _P, _I, c_double_p, c_void_pp, _D, _I]),
    ("cts_rosenbrock_cache_create", _I, [_P, _I, _SZ, _P, _P, TENDENCY_FN, _P]),
    ("cts_rosenbrock_cache_destroy", _I, [_P]),
    ("cts_rosenbrock_step", _I, [_P, _P, _D, _D]),
    ("cts_rosenbrock_cache_buffer", _I, [_P, C.c_char_p, _P]),
    ("cts_rosenbrock_cache_stats", _I, [_P, _P]),
    ("cts_rosenbrock_cache_set_fusion", _I, [_P, _I]),
    ("cts_ssprk_stage", _I, [_P, _I, _SZ, _P, _P, _P, _P, _P, _P, _D, _I, _D, _D, _I, c_double_p, c_void_pp, _I, _I]),
    ("cts_newton_residual", _I, [_P, _I, _SZ, _P, _P, _D, _P]),
    ("cts_sub_inplace", _I, [_P, _I, _SZ, _P, _P]),
    ("cts_timp_from_stage", _I, [_P, _I, _SZ, _P, _P, _P, _D]),
    ("cts_jvp_perturb", _I, [_P, _I, _SZ, _P, _P, _D, _P]),
    ("cts_jvp_quotient", _I, [_P, _I, _SZ, _P, _P, _P, _D]),
    ("cts_lsrk_update", _I, [_P, _I, _SZ, _P, _D, _P, _I]),
    ("cts_axpby", _I, [_P, _I, _SZ, _D, _P, _D, _P]),
    ("cts_div_scalar", _I, [_P, _I, _SZ, _P, _P, _D]),
    ("cts_lincomb", _I, [_P, _I, _SZ, _P, _I, c_double_p, c_void_pp, _I]),
    ("cts_lincomb_sub", _I, [_P, _I, _SZ, _P, _P, _I, c_double_p, c_void_pp]),
    ("cts_dot", _I, [_P, _I, _SZ, _P, _P, c_double_p]),
    ("cts_nrm2", _I, [_P, _I, _SZ, _P, c_double_p]),
    ("cts_sum1pabs", _I, [_P, _I, _SZ, _P, c_double_p]),
    ("cts_multi_dot", _I, [_P, _I, _SZ, _I, c_void_pp, _P, c_double_p]),
    ("cts_imex_cache_set_graph", _I, [_P, _I]),
    ("cts_imex_cache_graph_stats", _I, [_P, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    ("cts_comm_enable_peer_halo", _I, [_P, _SZ]),
    ("cts_comm_halo_path", _I, [_P, C.POINTER(C.c_int)]),
    ("cts_bytes_count", _I, [_P, C.POINTER(C.c_uint64)]),
    ("cts_set_reduction_window", _I, [_P, _SZ, _SZ]),
    ("cts_set_reduction_layout", _I, [_P, _SZ, _SZ, _SZ, _SZ, _SZ]),
    ("cts_reduction_global_length", _I, [_P, _SZ, C.POINTER(C.c_size_t)]),
    ("cts_tridiag_solve_batched", _I, [_P, _I, C.POINTER(Tridiag), _P, _P]),
    ("cts_tridiag_matvec", _I, [_P, _I, C.POINTER(Tridiag), _P, _P]),
    ("cts_imex_cache_create", _I, [_P, _I, _SZ, C.POINTER(ImexTableau), _I, C.POINTER(OdeFunction),
                                   C.POINTER(NewtonConfig), c_void_pp]),
    ("cts_imex_cache_destroy", _I, [_P]),
    ("cts_imex_step", _I, [_P, _P, _D, _D, _I]),
    ("cts_imex_cache_buffer", _I, [_P, C.c_char_p, c_void_pp]),
    ("cts_imex_cache_nbuffers", _I, [_P, C.POINTER(_I)]),
    ("cts_imex_cache_stats", _I, [_P, C.POINTER(SolverStats)]),
    ("cts_imex_cache_set_fusion", _I, [_P, _I]),
    ("cts_lsrk_cache_create", _I, [_P, _I, _SZ, _I, c_double_p, c_double_p, c_double_p, _P, _P, c_void_pp]),
    ("cts_lsrk_cache_destroy", _I, [_P]),
    ("cts_lsrk_step", _I, [_P, _P, _D, _D, _I]),
    ("cts_lsrk_cache_buffer", _I, [_P, C.c_char_p, c_void_pp]),
    ("cts_lsrk_cache_set_fusion", _I, [_P, _I]),
    ("cts_lsrk_cache_stats", _I, [_P, C.POINTER(SolverStats)]),
    ("cts_newton_cache_create", _I, [_P, _I, _SZ, C.POINTER(NewtonConfig), _P, _P, _P, _P, c_void_pp]),
    ("cts_newton_cache_destroy", _I, [_P]),
    ("cts_newton_solve", _I, [_P, _P, RESIDUAL_FN, JAC_FN, PREPARE_FN, _P]),
    ("cts_newton_cache_stats", _I, [_P, C.POINTER(SolverStats)]),
    ("cts_op_advection_create", _I, [_P, _I, _SZ, _D, _D, c_void_pp]),
    ("cts_op_advection_destroy", _I, [_P]),
    ("cts_op_advection_incr", _I, [_P, _P, _P, _D, _D, _D]),
    ("cts_op_column_create", _I, [_P, _I, C.POINTER(ColumnDesc), c_void_pp]),
    ("cts_op_column_destroy", _I, [_P]),
    ("cts_op_column_ndof", _SZ, [_P]),
    ("cts_op_column_make_W", _I, [_P, C.POINTER(C.POINTER(Tridiag))]),
    ("cts_op_column_free_W", _I, [_P, C.POINTER(Tridiag)]),
    ("cts_op_column_T_imp", _I, [_P, _P, _P, _D]),
    ("cts_op_column_T_exp", _I, [_P, _P, _P, _P, _D]),
    ("cts_op_column_Wfact", _I, [_P, _P, _P, _D, _D]),
    ("cts_op_column_ldiv", _I, [_P, _P, _P, _P]),
    ("cts_op_column_jac_mul", _I, [_P, _P, _P, _P]),
    ("cts_op_column_dss", _I, [_P, _P, _D]),
    ("cts_op_column_set_matrix_free", _I, [_P, _I]),
    ("cts_comm_get_unique_id", _I, [_P]),
    ("cts_comm_init", _I, [_P, _I, _I, _P]),
    ("cts_comm_destroy", _I, [_P]),
    ("cts_comm_rank", _I, [_P, C.POINTER(_I), C.POINTER(_I)]),
    ("cts_halo_exchange", _I, [_P, _I, _SZ, _P, _P, _P, _P]),
    ("cts_allreduce_sum_f64", _I, [_P, _P, _SZ]),
    ("cts_allreduce_max_f64_host", _I, [_P, c_double_p]),
    ("cts_fill_hash", _I, [_P, _I, _SZ, _P, _U64, _U64, _D, _D]),
]

_lib = None


class CTSError(RuntimeError):
    pass


def load():
    """Load the shared library (once).  Raises if it has not been built -- there is no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise CTSError(f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                       "(make -C climatimesteppers.jl_b200/csrc). The B200 path has no CPU fallback.")
    lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    for name, res, args in SIGNATURES:
        fn = getattr(lib, name)  # AttributeError if a declared symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def fn_addr(name: str) -> int:
    """Address of an exported function (library-native hooks are recognised by address)."""
    return C.cast(getattr(load(), name), C.c_void_p).value


def check(ctx_handle, status: int, what: str = ""):
    if status != 0:
        msg = load().cts_last_error(ctx_handle) if ctx_handle else b""
        raise CTSError(f"{what}: {STATUS.get(status, status)}: {msg.decode(errors='replace') if msg else ''}")
