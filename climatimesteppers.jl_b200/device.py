"""Device context and the state-array type of the B200 path.

``B200Vector`` plays the role of the user's state array type in the reference (the thing the
package calls ``zero``, ``copy``, ``norm``, ``length`` and broadcasts on -- SURVEY §8b "State array
type").  Memory is a ``torch`` CUDA tensor (torch is plumbing here: allocator + streams); every
arithmetic operation goes through ``libcts_sm100a.so``.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np

from . import _lib as L


def _torch():
    import torch
    return torch


class Context:
    """One GPU + one stream (``cts_ctx``).  By default it enqueues on torch's current stream so that
    torch copies, CUDA events and library kernels are ordered with each other."""

    _default: Optional["Context"] = None

    def __init__(self, device: Optional[int] = None, stream: Optional[int] = None):
        torch = _torch()
        self.lib = L.load()
        if not torch.cuda.is_available():
            raise L.CTSError("no CUDA device: the B200 path has no CPU fallback (CTS_ENODEVICE)")
        self.device = torch.cuda.current_device() if device is None else int(device)
        torch.cuda.set_device(self.device)
        if stream is None:
            # torch reports its default stream as handle 0; the library must then enqueue on that SAME stream
            # (cudaStreamLegacy = 0x1), not on a private one, or torch fills/copies would race with our kernels.
            stream = torch.cuda.current_stream(self.device).cuda_stream or 0x1
        self.stream = stream
        h = C.c_void_p()
        L.check(None, self.lib.cts_ctx_create(C.byref(h), self.device, C.c_void_p(stream)), "cts_ctx_create")
        self.handle = h
        self.nranks, self.rank = 1, 0

    @classmethod
    def default(cls) -> "Context":
        if cls._default is None:
            cls._default = cls()
        return cls._default

    def check(self, status, what=""):
        L.check(self.handle, status, what)

    def sync(self):
        self.check(self.lib.cts_ctx_sync(self.handle), "cts_ctx_sync")

    def launch_count(self) -> int:
        out = C.c_uint64()
        self.check(self.lib.cts_launch_count(self.handle, C.byref(out)))
        return int(out.value)

    def bytes_count(self) -> int:
        """Algorithmic HBM bytes of all kernels launched so far (cts_bytes_count)."""
        out = C.c_uint64()
        self.check(self.lib.cts_bytes_count(self.handle, C.byref(out)))
        return int(out.value)

    def prof_enable(self, enabled: bool = True):
        self.check(self.lib.cts_prof_enable(self.handle, int(enabled)))

    def prof_read(self) -> dict:
        """{category: (device ms, launches)} since the last read (the per-hook table of `benchmark_step`)."""
        n = len(L.PROF_CATEGORIES)
        ms, cnt = (C.c_double * n)(), (C.c_uint64 * n)()
        self.check(self.lib.cts_prof_read(self.handle, ms, cnt))
        return {k: (ms[i], int(cnt[i])) for i, k in enumerate(L.PROF_CATEGORIES) if cnt[i]}

    def set_reduction_window(self, offset: int, count: int):
        self.check(self.lib.cts_set_reduction_window(self.handle, offset, count))

    def set_reduction_layout(self, local_offset: int, count: int, global_offset: int, global_count: int, chunk: int = 0):
        """Shard-invariant reductions (cts_set_reduction_layout): the owned window is a slice of a global index space cut
        into fixed chunks; 1, 2, 4 or 8 ranks give bit-identical norms/dots."""
        self.check(self.lib.cts_set_reduction_layout(self.handle, local_offset, count, global_offset, global_count, chunk),
                   "cts_set_reduction_layout")

    def halo_path(self) -> str:
        """How `dss!` halo exchanges travel: local copies (1 rank), NCCL send/recv, or the NVLink peer-memory kernel."""
        p = C.c_int()
        self.check(self.lib.cts_comm_halo_path(self.handle, C.byref(p)), "cts_comm_halo_path")
        return {0: "local", 1: "nccl", 2: "peer"}[p.value]

    def enable_peer_halo(self, max_row_bytes: int):
        self.check(self.lib.cts_comm_enable_peer_halo(self.handle, int(max_row_bytes)), "cts_comm_enable_peer_halo")
        return self.halo_path()

    def close(self):
        if getattr(self, "handle", None):
            self.lib.cts_ctx_destroy(self.handle)
            self.handle = None


_DT = {np.dtype(np.float32): L.CTS_F32, np.dtype(np.float64): L.CTS_F64}


class _RawCuda:
    """Exposes a raw device pointer through ``__cuda_array_interface__`` so torch can view it."""

    def __init__(self, ptr, n, dtype):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": np.dtype(dtype).str, "data": (int(ptr), False),
                                         "version": 3, "strides": None}


class B200Vector:
    """Flat device vector of Float32/Float64."""

    __slots__ = ("ctx", "tensor", "dtype", "_keep")

    def __init__(self, tensor, ctx: Optional[Context] = None, keep=None):
        torch = _torch()
        assert tensor.is_cuda and tensor.is_contiguous() and tensor.dim() == 1
        self.ctx = ctx or Context.default()
        self.tensor = tensor
        self.dtype = np.dtype(np.float64 if tensor.dtype == torch.float64 else np.float32)
        self._keep = keep

    # -- construction ---------------------------------------------------------------------------
    @classmethod
    def from_host(cls, arr, ctx: Optional[Context] = None) -> "B200Vector":
        torch = _torch()
        ctx = ctx or Context.default()
        a = np.ascontiguousarray(arr).ravel()
        assert a.dtype in _DT
        t = torch.from_numpy(a).to(f"cuda:{ctx.device}")
        return cls(t, ctx)

    @classmethod
    def zeros(cls, n: int, dtype=np.float64, ctx: Optional[Context] = None) -> "B200Vector":
        torch = _torch()
        ctx = ctx or Context.default()
        tdt = torch.float64 if np.dtype(dtype) == np.float64 else torch.float32
        return cls(torch.zeros(n, dtype=tdt, device=f"cuda:{ctx.device}"), ctx)

    @classmethod
    def from_ptr(cls, ptr: int, n: int, dtype, ctx: Context, keep=None) -> "B200Vector":
        """View of a library-owned buffer (cache arrays returned by ``cts_imex_cache_buffer``)."""
        torch = _torch()
        t = torch.as_tensor(_RawCuda(ptr, n, dtype), device=f"cuda:{ctx.device}")
        return cls(t, ctx, keep)

    # -- the array interface the reference relies on -------------------------------------------
    @property
    def ptr(self) -> int:
        return self.tensor.data_ptr()

    @property
    def ft(self) -> int:
        return _DT[self.dtype]

    def __len__(self):
        return self.tensor.numel()

    def zero(self) -> "B200Vector":  # Base.zero
        return B200Vector.zeros(len(self), self.dtype, self.ctx)

    def copy(self) -> "B200Vector":  # Base.copy (save_func default, integrators.jl:197)
        return B200Vector(self.tensor.clone(), self.ctx)

    def copyto(self, other: "B200Vector"):  # `u .= other`
        self.tensor.copy_(other.tensor)
        return self

    def to_host(self) -> np.ndarray:
        self.ctx.sync()
        return self.tensor.cpu().numpy()

    def norm(self) -> float:  # LinearAlgebra.norm
        out = C.c_double()
        self.ctx.check(self.ctx.lib.cts_nrm2(self.ctx.handle, self.ft, len(self), self.ptr, C.byref(out)), "cts_nrm2")
        return out.value

    def dot(self, other: "B200Vector") -> float:
        out = C.c_double()
        self.ctx.check(self.ctx.lib.cts_dot(self.ctx.handle, self.ft, len(self), self.ptr, other.ptr, C.byref(out)), "cts_dot")
        return out.value

    def sum1pabs(self) -> float:  # sum(x_i -> 1 + abs(x_i), x)
        out = C.c_double()
        self.ctx.check(self.ctx.lib.cts_sum1pabs(self.ctx.handle, self.ft, len(self), self.ptr, C.byref(out)))
        return out.value

    def fill_hash(self, seed: int, global_offset: int = 0, scale: float = 1.0, shift: float = 0.0):
        self.ctx.check(self.ctx.lib.cts_fill_hash(self.ctx.handle, self.ft, len(self), self.ptr, seed, global_offset,
                                                  scale, shift), "cts_fill_hash")
        return self


def axpy_n(out: B200Vector, base: B200Vector, grp1=(), grp2=(), c0: float = 1.0, out2: Optional[B200Vector] = None,
           compute_native: bool = False):
    """``out = (c0*base + sum grp1) + sum grp2`` -- the ``copyto!(dest, ::Broadcasted)`` of a fused increment
    (fused_increment.jl:220-236).  ``grp*`` are sequences of ``(coef, B200Vector)``."""
    ctx = out.ctx
    terms = list(grp1) + list(grp2)
    k = len(terms)
    coef = (C.c_double * max(k, 1))(*[float(c) for c, _ in terms])
    ptrs = (C.c_void_p * max(k, 1))(*[v.ptr for _, v in terms])
    ctx.check(ctx.lib.cts_axpy_n(ctx.handle, out.ft, len(out), out.ptr, out2.ptr if out2 is not None else None,
                                 float(c0), base.ptr, len(grp1), len(grp2), coef, ptrs, int(compute_native)), "cts_axpy_n")
    return out
