"""
Host-side mirror of the AbstractNFFTs plan API for the B200 backend.

The reference's host language is Julia, which is not installed in this image; the Julia glue that a
maintainer would load is `nfft.jl_b200/julia/B200NFFT.jl` (same C ABI).  This module is the executable
twin: same names, argument meaning and error behaviour as
    plan_nfft / nodes! / mul! / adjoint(p) * x / size_in / size_out / convolve! / convolve_transpose! /
    deconvolve! / deconvolve_transpose! / PrecomputeFlags / TimingStats
(/root/reference/AbstractNFFTs/src/interface.jl:170-243, derived.jl:7-36, misc.jl:13-81,
 /root/reference/src/implementation.jl:73-193), so that the parity tests read like the reference's tests.

Array conventions (identical memory to Julia):
  nodes  k     shape (D, M)      -- k[:, j] is node j          (Julia Matrix D x M)
  image  f     shape N (+ (B,))  -- f[i1, i2, ...] with i1 fastest in memory (Fortran order)
  output fHat  shape (M,) (+ (B,))
numpy arrays are treated as HOST buffers (copied through the library's staging buffers), torch CUDA
tensors as DEVICE buffers (zero copy; they must have the Fortran-order strides `empty_image()` gives).
"""
from __future__ import annotations

import ctypes as C
import dataclasses
import enum
import time

import numpy as np

from . import _lib

try:  # torch is plumbing only (device memory, streams)
    import torch
except Exception:  # pragma: no cover
    torch = None


class ArgumentError(ValueError):
    """Julia's ArgumentError"""


class DimensionMismatch(ValueError):
    """Julia's DimensionMismatch"""


# getWindow symbols (src/windowFunctions.jl:4-19) in the order of the NFFTB200_* window enum
WINDOWS = ("kaiser_bessel", "gauss", "spline", "kaiser_bessel_rev", "cosh_type", "exp_sqrt")   # exp_sqrt: not in the reference


class PrecomputeFlags(enum.IntEnum):
    """AbstractNFFTs/src/misc.jl:13-18"""
    FULL = 1
    TENSOR = 2
    LINEAR = 3
    POLYNOMIAL = 4


FULL, TENSOR, LINEAR, POLYNOMIAL = (PrecomputeFlags.FULL, PrecomputeFlags.TENSOR,
                                    PrecomputeFlags.LINEAR, PrecomputeFlags.POLYNOMIAL)


@dataclasses.dataclass
class TimingStats:
    """AbstractNFFTs/src/misc.jl:22-42 (seconds)"""
    pre: float = 0.0
    conv: float = 0.0
    fft: float = 0.0
    deconv: float = 0.0
    conv_adjoint: float = 0.0
    fft_adjoint: float = 0.0
    deconv_adjoint: float = 0.0


@dataclasses.dataclass
class NFFTParams:
    """src/implementation.jl:3-14"""
    m: int
    σ: float
    reltol: float
    window: str
    LUTSize: int
    precompute: PrecomputeFlags
    sortNodes: bool
    storeDeconvolutionIdx: bool
    blocking: bool
    blockSize: tuple

    @property
    def sigma(self):
        return self.σ


_STATUS_EXC = {1: ArgumentError, 2: ArgumentError, 3: DimensionMismatch, 4: NotImplementedError,
               5: RuntimeError, 6: RuntimeError, 7: MemoryError, 8: ArgumentError, 9: RuntimeError}

HOST, DEVICE, HOST_ASYNC = 0, 1, 2


def _check(handle, status):
    if status != 0:
        L = _lib.lib()
        msg = L.nfftb200_last_error(handle)
        msg = msg.decode() if msg else ""
        raise _STATUS_EXC.get(status, RuntimeError)(
            f"{L.nfftb200_status_string(status).decode()}: {msg}")


def accuracyParams(m=None, σ=None, reltol=None):
    """AbstractNFFTs/src/misc.jl:66-81"""
    mo, so, ro = C.c_int(), C.c_double(), C.c_double()
    _lib.lib().nfftb200_accuracy_params(int(m or 0), float(σ or 0.0), float(reltol or 0.0),
                                        C.byref(mo), C.byref(so), C.byref(ro))
    return mo.value, so.value, ro.value


def _is_torch(x):
    return torch is not None and isinstance(x, torch.Tensor)


def _fortran_strides(shape):
    s, acc = [], 1
    for n in shape:
        s.append(acc)
        acc *= n
    return tuple(s)


def _normalise_dims(dims, D):
    """dims kwarg of plan_nfft (an Integer or a UnitRange, src/precomputation.jl:36-50) -> tuple of 1-based ints"""
    if dims is None:
        return tuple(range(1, D + 1))
    if isinstance(dims, (int, np.integer)):
        dims = (int(dims),)
    dims = tuple(int(d) for d in dims)
    if not dims or dims != tuple(range(dims[0], dims[-1] + 1)) or dims[0] < 1 or dims[-1] > D:
        raise ArgumentError(f"dims = {dims} must be a range inside 1:{D}")
    return dims


def _dir_to_internal(x, npre, nlead):
    """directional layout -> batched layout: array of size (pre..., lead..., post...) -> (lead..., B) with the batch
    index running over (pre..., post...) in column-major order (pre fastest).  numpy or torch; returns a
    Fortran-ordered array (a copy unless the data already lies that way)."""
    nd = len(x.shape)
    order = list(range(npre, npre + nlead)) + list(range(npre)) + list(range(npre + nlead, nd))
    lead = tuple(x.shape[a] for a in order[:nlead])
    B = 1
    for a in order[nlead:]:
        B *= int(x.shape[a])
    if _is_torch(x):
        xr = x.permute(*order[::-1]).contiguous()              # C-contiguous over the reversed dims == column-major
        return xr.view((B,) + lead[::-1]).permute(*range(nlead, -1, -1))
    return np.asfortranarray(np.reshape(np.transpose(x, order), lead + (B,), order="F"))


def _dir_from_internal(y, pre, post, out):
    """inverse of _dir_to_internal: y of size (lead..., B) is written into out of size (pre..., lead..., post...)"""
    nlead = len(y.shape) - 1
    npre = len(pre)
    lead = tuple(y.shape[:nlead])
    nd = npre + nlead + len(post)
    order = list(range(npre, npre + nlead)) + list(range(npre)) + list(range(npre + nlead, nd))
    inv = [order.index(a) for a in range(nd)]
    full = lead + tuple(pre) + tuple(post)
    if _is_torch(y):
        yr = y.permute(*range(nlead, -1, -1)).contiguous().view(full[::-1]).permute(*range(nd - 1, -1, -1))
        out.copy_(yr.permute(*inv))
    else:
        out[...] = np.transpose(np.reshape(y, full, order="F"), inv)
    return out


class _Buf:
    """pointer + location of a caller array; keeps converted copies alive"""

    def __init__(self, ptr, where, keep, copied_from=None):
        self.ptr, self.where, self.keep, self.copied_from = ptr, where, keep, copied_from


class B200NFFTPlan:
    """NFFTPlan{T,D,1} on one B200 (src/implementation.jl:16-43): opaque C handle + mirrored fields."""

    def __init__(self, k, N, *, m=None, σ=None, sigma=None, reltol=None, window="kaiser_bessel",
                 precompute=POLYNOMIAL, ntransforms=1, blockSize=None, dims=None, device=None,
                 sortNodes=False, storeDeconvolutionIdx=False, blocking=True, fftflags=None,
                 LUTSize=0, timing=None, stream="current", shard=None, process_group=None):
        t0 = time.perf_counter()
        if σ is None:
            σ = sigma
        if isinstance(N, (int, np.integer)):
            N = (int(N),)
        N = tuple(int(n) for n in N)
        D = len(N)
        dims_t = _normalise_dims(dims, D)
        self._dir = None
        self._user_N = N
        if dims_t != tuple(range(1, D + 1)):
            # directional plan (src/directional.jl, test/accuracy.jl:83-163): the transform runs over dims of an
            # array of size N, every other dimension is a batch.  A transform over the LEADING dims 1:D-1 is exactly
            # a batched plan with ntransforms = N[D] (batch slowest, no copies); any other dims are brought to that
            # layout by one permuting copy on the way in and one on the way out (_dir_to_internal/_from_internal).
            if int(ntransforms) != 1:
                raise ArgumentError("a directional plan (dims=...) cannot also be batched (ntransforms > 1)")
            pre, post = N[:dims_t[0] - 1], N[dims_t[-1]:]
            if not (len(pre) == 0 and len(post) == 1):
                self._dir = (tuple(pre), tuple(post))
            ntransforms = int(np.prod(pre + post))
            N = N[dims_t[0] - 1:dims_t[-1]]
            D = len(N)
        window = str(window).lstrip(":")
        if window not in WINDOWS:
            raise NotImplementedError(f"Window {window} not yet implemented!")      # src/windowFunctions.jl:16
        k_is_torch = _is_torch(k)
        kshape = tuple(k.shape)
        if len(kshape) == 1:
            kshape = (1, kshape[0])                                  # derived.jl:23-27
        if kshape[0] != D:
            raise ArgumentError(f"Nodes x have dimension {kshape[0]} != {D}")   # precomputation.jl:19-21
        if k_is_torch:
            T = np.float32 if k.dtype == torch.float32 else np.float64
        else:
            k = np.asarray(k)
            T = np.float32 if k.dtype == np.float32 else np.float64
        self.T = T
        self.cT = np.complex64 if T == np.float32 else np.complex128
        m_, σ_, reltol_ = accuracyParams(m, σ, reltol)
        self._requested = (m_, σ_)
        self._L = _lib.lib()
        if device is None:
            device = k.device.index if (k_is_torch and k.is_cuda) else (
                torch.cuda.current_device() if (torch is not None and torch.cuda.is_available()) else 0)
        self.device = int(device)
        self.N = N
        self.D = D
        self.dims = range(dims_t[0], dims_t[-1] + 1)
        # ---- multi-GPU (one process per GPU): shard = None | "batch" | "nodes"
        self.shard = shard
        self.rank, self.world = 0, 1
        self.global_ntransforms = int(ntransforms)
        if shard is not None:
            import torch.distributed as dist
            self.rank, self.world = dist.get_rank(process_group), dist.get_world_size(process_group)
            if shard == "batch":
                lo, hi = shard_batch(int(ntransforms), self.rank, self.world)
                self.batch_range = (lo, hi)
                ntransforms = hi - lo
            elif shard != "nodes":
                raise ArgumentError("shard must be None, 'batch' or 'nodes'")
        self.ntransforms = int(ntransforms)
        self._h = C.c_void_p()
        Narr = (C.c_int64 * D)(*N)
        bs = (C.c_int64 * D)(*[int(b) for b in blockSize]) if blockSize is not None else None
        st = self._L.nfftb200_plan_create(C.byref(self._h), D, Narr, 0 if T == np.float32 else 1, m_, σ_, WINDOWS.index(window),
                                          int(precompute), self.ntransforms, bs, self.device)
        _check(None, st)
        Nt = (C.c_int64 * D)()
        bso = (C.c_int64 * D)()
        nt, lut, sg, M = C.c_int64(), C.c_int64(), C.c_double(), C.c_int64()
        self._L.nfftb200_get_info(self._h, Nt, bso, C.byref(nt), C.byref(lut), C.byref(sg), C.byref(M))
        self.Ñ = tuple(Nt)
        self.num_tiles = nt.value
        self.params = NFFTParams(m=m_, σ=sg.value, reltol=reltol_, window=window, LUTSize=lut.value,
                                 precompute=PrecomputeFlags(int(precompute)), sortNodes=bool(sortNodes),
                                 storeDeconvolutionIdx=bool(storeDeconvolutionIdx), blocking=bool(blocking),
                                 blockSize=tuple(bso))
        self._follow_torch_stream = stream == "current" and torch is not None and torch.cuda.is_available()
        self._bound_stream = None
        self._user_dims = dims
        self._rebind_stream()
        self._timing_on = False
        self._async_keep = []
        self.J = 0
        self.NOut = (0,)
        self.k = None
        if shard == "nodes" and self.world > 1:
            self.comm_init(broadcast_unique_id(self._L, self.rank, process_group, self.device), self.rank, self.world, 2)
        elif shard == "batch":
            _check(self._h, self._L.nfftb200_comm_init(self._h, None, self.rank, self.world, 1))
        self.nodes_(k)
        if timing is not None:
            timing.pre = time.perf_counter() - t0                      # src/NFFT.jl:51-56

    # aliases for ASCII-only callers
    @property
    def Nt(self):
        return self.Ñ

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass

    def destroy(self):
        h = getattr(self, "_h", None)
        if h is not None and h.value:
            self._L.nfftb200_destroy(h)
            self._h = C.c_void_p()

    def copy(self):
        """Base.copy(p::NFFTPlan) (src/implementation.jl:45-66): an independent plan with the same parameters and
        nodes and its own grid, FFT plans, scratch and stream state -- what a second task needs, since a plan is not
        re-entrant.  (Sharded plans are collective objects and cannot be copied.)"""
        if self.shard is not None:
            raise ArgumentError("a sharded plan cannot be copied")
        directional = self._user_dims is not None and tuple(self._user_N) != tuple(self.N)
        q = B200NFFTPlan(self.k, self._user_N if directional else self.N, m=self._requested[0], σ=self._requested[1],
                         window=self.params.window, precompute=self.params.precompute, blockSize=self.params.blockSize,
                         ntransforms=1 if directional else self.ntransforms, dims=self._user_dims if directional else None,
                         device=self.device, sortNodes=False,
                         storeDeconvolutionIdx=self.params.storeDeconvolutionIdx, blocking=self.params.blocking)
        q.params = dataclasses.replace(self.params)
        return q

    __copy__ = copy

    def __repr__(self):
        return (f"B200NFFTPlan with {self.J} sampling points for an input array of size{self.N} and an "
                f"output array of size{self.NOut} with dims {tuple(self.dims)}")

    # ---- interface.jl:170-211 -----------------------------------------------------------------
    def size_in(self):
        return self._user_shape(self.N)

    def size_out(self):
        return self._user_shape(self.NOut)

    def _user_shape(self, lead):
        """shape the caller sees: (lead..., B) for batched plans, (pre..., lead..., post...) for directional ones"""
        if self._dir is None:
            return self._bshape(lead)
        return self._dir[0] + tuple(lead) + self._dir[1]

    def _dir_call(self, fn, out, x, lead_out, lead_in, **kw):
        """run a batched transform on directional data: permute in, transform, permute out"""
        if tuple(x.shape) != self._user_shape(lead_in) or tuple(out.shape) != self._user_shape(lead_out):
            raise DimensionMismatch("Data is not consistent with NFFTPlan")
        dev = _is_torch(x) and x.is_cuda
        if dev != (_is_torch(out) and out.is_cuda):
            raise ArgumentError("input and output must both be host (numpy) or both device (CUDA tensor) arrays")
        xi = _dir_to_internal(x if (dev or not _is_torch(x)) else x.numpy(), len(self._dir[0]), len(lead_in))
        if not dev:
            xi = np.asfortranarray(xi, dtype=self.cT)
        yi = _empty_fortran(self._bshape(lead_out), self.cT, self.device, dev)
        fn(yi, xi, _internal=True, **kw)
        _dir_from_internal(yi, self._dir[0], self._dir[1], out.numpy() if (_is_torch(out) and not dev) else out)
        return out

    def adjoint(self):
        return AdjointPlan(self)

    @property
    def H(self):
        return AdjointPlan(self)

    def set_stream(self, cuda_stream: int):
        _check(self._h, self._L.nfftb200_set_stream(self._h, C.c_void_p(cuda_stream)))
        self._bound_stream = cuda_stream
        if not getattr(self, "_in_rebind", False):
            self._follow_torch_stream = False          # an explicit stream: stop following torch's current stream

    def _rebind_stream(self):
        """stream="current" plans follow torch's CURRENT stream: it is looked up again at every call and the plan is
        re-bound when it changed (nfftb200_set_stream synchronises the old stream first), so tensors produced on the
        caller's stream are never consumed on a stale one"""
        if not self._follow_torch_stream:
            return
        cur = torch.cuda.current_stream(self.device).cuda_stream
        if cur != self._bound_stream:
            self._in_rebind = True
            try:
                self.set_stream(cur)
            finally:
                self._in_rebind = False

    def sync(self):
        _check(self._h, self._L.nfftb200_sync(self._h))
        self._async_keep = []

    def set_kernel_mode(self, mode: int):
        _check(self._h, self._L.nfftb200_set_kernel_mode(self._h, int(mode)))

    def kernel_times(self):
        """device seconds of the last spread stage (kernel + gather), interp kernel, grid memset and gather pass"""
        t = (C.c_double * 4)()
        _check(self._h, self._L.nfftb200_get_kernel_times(self._h, t))
        return {"spread": t[0], "interp": t[1], "memset": t[2], "gather": t[3]}

    def enable_timing(self, on=True):
        self._L.nfftb200_set_timing(self._h, int(on))
        self._timing_on = bool(on)

    def launch_count(self) -> int:
        n = C.c_int64()
        self._L.nfftb200_get_launch_count(self._h, C.byref(n))
        return n.value

    # ---- nodes!(p, k)  src/implementation.jl:108-141 --------------------------------------------
    def nodes_(self, k):
        self._rebind_stream()
        D = self.D
        if _is_torch(k) and k.is_cuda:
            kk = k if k.dim() == 2 else k.reshape(1, -1)
            if kk.shape[0] != D:
                raise ArgumentError(f"Nodes x have dimension {kk.shape[0]} != {D}")
            want = torch.float32 if self.T == np.float32 else torch.float64
            kt = kk.to(want).t().contiguous()                         # memory: node-major == Julia D x M
            M = kt.shape[0]
            st = self._L.nfftb200_set_nodes(self._h, C.c_void_p(kt.data_ptr()), M, DEVICE)
            _check(self._h, st)
            self.k = k
        else:
            ka = np.asarray(k.cpu().numpy() if _is_torch(k) else k)
            if ka.ndim == 1:
                ka = ka.reshape(1, -1)
            if ka.shape[0] != D:
                raise ArgumentError(f"Nodes x have dimension {ka.shape[0]} != {D}")
            if self.params.sortNodes:                                   # precomputation.jl:52-54
                order = np.lexsort(ka[::-1])
                ka[...] = ka[:, order]
            kf = np.asfortranarray(ka, dtype=self.T)
            M = kf.shape[1]
            st = self._L.nfftb200_set_nodes(self._h, kf.ctypes.data_as(C.c_void_p), M, HOST)
            _check(self._h, st)
            self.k = ka
        self.J = int(M)
        self.NOut = (self.J,)
        return self

    # ---- plan internals exposed for parity checks ----------------------------------------------
    def permutation(self):
        """concat_l nodesInBlock[l] (0-based) and the tile prefix sums."""
        perm = np.empty(self.J, dtype=np.int64)
        ts = np.empty(self.num_tiles + 1, dtype=np.int64)
        _check(self._h, self._L.nfftb200_get_permutation(self._h, perm.ctypes.data_as(C.c_void_p),
                                                         ts.ctypes.data_as(C.c_void_p)))
        return perm, ts

    def table(self, which):
        n = C.c_int64()
        self._L.nfftb200_get_table(self._h, which, None, 0, C.byref(n))
        out = np.empty(n.value, dtype=np.float64)
        self._L.nfftb200_get_table(self._h, which, out.ctypes.data_as(C.c_void_p), n.value, C.byref(n))
        return out

    @property
    def windowHatInvLUT(self):
        t = self.table(0)
        out, o = [], 0
        for n in self.N:
            out.append(t[o:o + n].astype(self.T))
            o += n
        return out

    @property
    def windowPolyInterp(self):
        t = self.table(1)
        m = self.params.m
        return t.reshape((2 * m + 1, 2 * m), order="F").astype(self.T) if t.size else t.reshape(0, 0)

    @property
    def windowLinInterp(self):
        return self.table(2).astype(self.T)

    @property
    def tmpVec(self):
        """the plan's device grid as a torch tensor view (shape Ñ (+B), Fortran strides)"""
        ptr = C.c_void_p()
        self._L.nfftb200_get_grid(self._h, C.byref(ptr))
        shape = self.Ñ + ((self.ntransforms,) if self.ntransforms > 1 else ())
        return _device_view(ptr.value, shape, self.cT, self.device, self)

    # ---- buffers ---------------------------------------------------------------------------------
    def _bshape(self, base):
        return tuple(base) + ((self.ntransforms,) if self.ntransforms > 1 else ())

    def empty_image(self, device=True, batch=True, dtype=None):
        return _empty_fortran(self._bshape(self.N) if batch else self.N, dtype or self.cT, self.device, device)

    def empty_out(self, device=True, batch=True, dtype=None):
        return _empty_fortran(self._bshape(self.NOut) if batch else self.NOut, dtype or self.cT, self.device, device)

    def empty_grid(self, device=True, dtype=None):
        return _empty_fortran(self.Ñ, dtype or self.cT, self.device, device)

    def _in(self, x, shape, dtype, what):
        """read-only argument -> _Buf"""
        if tuple(x.shape) != tuple(shape):
            raise DimensionMismatch(f"{what}: size {tuple(x.shape)} != {tuple(shape)}")
        if _is_torch(x) and x.is_cuda:
            want = _torch_dtype(dtype)
            if x.dtype != want:
                x = x.to(want)
            if tuple(x.stride()) != _fortran_strides(shape) and x.numel() > 1:
                xf = _empty_fortran(shape, dtype, self.device, True)
                xf.copy_(x)
                x = xf
            return _Buf(x.data_ptr(), DEVICE, x)
        a = np.asarray(x.cpu().numpy() if _is_torch(x) else x)
        a = np.asfortranarray(a, dtype=dtype)
        return _Buf(a.ctypes.data, HOST, a)

    def _out(self, x, shape, dtype, what):
        """mutated argument -> _Buf (must already have the right dtype and layout)"""
        if tuple(x.shape) != tuple(shape):
            raise DimensionMismatch(f"{what}: size {tuple(x.shape)} != {tuple(shape)}")
        if _is_torch(x) and x.is_cuda:
            if x.dtype != _torch_dtype(dtype) or (tuple(x.stride()) != _fortran_strides(shape) and x.numel() > 1):
                raise ArgumentError(f"{what}: output tensor must be {dtype} with Fortran-order strides "
                                    "(use plan.empty_image()/empty_out()/empty_grid())")
            return _Buf(x.data_ptr(), DEVICE, x)
        if not isinstance(x, np.ndarray):
            raise ArgumentError(f"{what}: output must be a numpy array or a CUDA tensor")
        if x.dtype == dtype and (x.flags.f_contiguous or x.ndim <= 1 and x.flags.c_contiguous) and x.flags.writeable:
            return _Buf(x.ctypes.data, HOST, x)
        tmp = np.empty(shape, dtype=dtype, order="F")                  # converted on the way back
        return _Buf(tmp.ctypes.data, HOST, tmp, copied_from=x)

    @staticmethod
    def _finish(buf):
        if buf.copied_from is not None:
            buf.copied_from[...] = buf.keep

    def _timing(self, timing):
        self._rebind_stream()
        on = timing is not None or self._timing_on
        if on != self._timing_on:
            self._L.nfftb200_set_timing(self._h, int(on))
            self._timing_on = on

    def _fill_timing(self, timing):
        if timing is None:
            return
        t = (C.c_double * 7)()
        self._L.nfftb200_get_timing(self._h, t)
        pre = timing.pre
        (timing.pre, timing.conv, timing.fft, timing.deconv, timing.conv_adjoint, timing.fft_adjoint,
         timing.deconv_adjoint) = list(t)
        timing.pre = pre or timing.pre

    # ---- mul!  src/implementation.jl:155-193 -------------------------------------------------------
    def _async_host(self, bi, bo):
        """asynchronous host mode: page-locked numpy buffers, no conversions; they belong to the plan until sync()"""
        if bi.where != HOST or bo.where != HOST or bo.copied_from is not None:
            raise ArgumentError("async_host needs numpy (page-locked) buffers of the plan's dtype in Fortran order")
        self._async_keep += [bi.keep, bo.keep]
        return HOST_ASYNC

    def mul_forward(self, fHat, f, timing=None, verbose=False, async_host=False, _internal=False):
        if self._dir is not None and not _internal:
            if async_host:
                raise ArgumentError("async_host is not available for plans that permute their data (dims=...)")
            return self._dir_call(self.mul_forward, fHat, f, self.NOut, self.N, timing=timing, verbose=verbose)
        # consistencyCheck, src/utils.jl:98-105
        if tuple(f.shape) != self._bshape(self.N) or tuple(fHat.shape) != self._bshape(self.NOut):
            raise DimensionMismatch("Data is not consistent with NFFTPlan")
        self._timing(timing)
        bi = self._in(f, self._bshape(self.N), self.cT, "f")
        bo = self._out(fHat, self._bshape(self.NOut), self.cT, "fHat")
        where = DEVICE if (bi.where == DEVICE and bo.where == DEVICE) else HOST
        bi, bo = self._same_side(bi, bo, where)
        if async_host:
            where = self._async_host(bi, bo)
        _check(self._h, self._L.nfftb200_exec_forward(self._h, C.c_void_p(bi.ptr), C.c_void_p(bo.ptr), where))
        self._finish(bo)
        self._fill_timing(timing)
        if verbose and timing is not None:
            print(f"Timing: deconv={timing.deconv} fft={timing.fft} conv={timing.conv}")
        return fHat

    def mul_adjoint(self, f, fHat, timing=None, verbose=False, async_host=False, _internal=False):
        if self._dir is not None and not _internal:
            if async_host:
                raise ArgumentError("async_host is not available for plans that permute their data (dims=...)")
            return self._dir_call(self.mul_adjoint, f, fHat, self.N, self.NOut, timing=timing, verbose=verbose)
        if tuple(f.shape) != self._bshape(self.N) or tuple(fHat.shape) != self._bshape(self.NOut):
            raise DimensionMismatch("Data is not consistent with NFFTPlan")
        self._timing(timing)
        bi = self._in(fHat, self._bshape(self.NOut), self.cT, "fHat")
        bo = self._out(f, self._bshape(self.N), self.cT, "f")
        where = DEVICE if (bi.where == DEVICE and bo.where == DEVICE) else HOST
        bi, bo = self._same_side(bi, bo, where)
        if async_host:
            where = self._async_host(bi, bo)
        _check(self._h, self._L.nfftb200_exec_adjoint(self._h, C.c_void_p(bi.ptr), C.c_void_p(bo.ptr), where))
        self._finish(bo)
        self._fill_timing(timing)
        if verbose and timing is not None:
            print(f"Timing: conv={timing.conv_adjoint} fft={timing.fft_adjoint} deconv={timing.deconv_adjoint}")
        return f

    @staticmethod
    def _same_side(bi, bo, where):
        if bi.where != bo.where:
            raise ArgumentError("input and output must both be host (numpy) or both device (CUDA tensor) arrays")
        return bi, bo

    # allocating versions, derived.jl:174-208
    def __mul__(self, f):
        dev = _is_torch(f) and f.is_cuda
        out = _empty_fortran(self.size_out(), self.cT, self.device, dev)
        return self.mul_forward(out, f)

    __matmul__ = __mul__

    # ---- optional operators, interface.jl:217-243 ------------------------------------------------
    def _real_or_complex(self, x):
        if _is_torch(x):
            return x.is_complex()
        return np.iscomplexobj(x)

    def convolve_(self, g, fHat):
        self._rebind_stream()
        """convolve!(p, g, fHat) -> fHat  (src/convolution.jl:20-53)"""
        if tuple(g.shape) != self.Ñ:
            raise DimensionMismatch(f"size(g)={tuple(g.shape)} ≠ Ñ = {self.Ñ}")
        if tuple(fHat.shape) != (self.J,):
            raise DimensionMismatch(f"size(fHat)={tuple(fHat.shape)} ≠ J = {self.J}")
        cin, cout = self._real_or_complex(g), self._real_or_complex(fHat)
        if cin and not cout:
            raise ArgumentError("Complex input g requires Complex output fHat")
        cplx = cout
        dt = self.cT if cplx else self.T
        bi = self._in(g if (cin or not cplx) else _to_complex(g), self.Ñ, dt, "g")
        bo = self._out(fHat, (self.J,), dt, "fHat")
        self._same_side(bi, bo, None)
        _check(self._h, self._L.nfftb200_convolve(self._h, C.c_void_p(bi.ptr), C.c_void_p(bo.ptr), int(cplx), bi.where))
        self._finish(bo)
        return fHat

    def convolve_transpose_(self, fHat, g):
        self._rebind_stream()
        """convolve_transpose!(p, fHat, g) -> g  (src/convolution.jl:115-149)"""
        if tuple(g.shape) != self.Ñ:
            raise DimensionMismatch(f"size(g)={tuple(g.shape)} ≠ Ñ = {self.Ñ}")
        if tuple(fHat.shape) != (self.J,):
            raise DimensionMismatch(f"size(fHat)={tuple(fHat.shape)} ≠ J = {self.J}")
        cin, cout = self._real_or_complex(fHat), self._real_or_complex(g)
        if cin and not cout:
            raise ArgumentError("Complex input fHat requires Complex output g")
        cplx = cout
        dt = self.cT if cplx else self.T
        bi = self._in(fHat if (cin or not cplx) else _to_complex(fHat), (self.J,), dt, "fHat")
        bo = self._out(g, self.Ñ, dt, "g")
        self._same_side(bi, bo, None)
        _check(self._h, self._L.nfftb200_convolve_transpose(self._h, C.c_void_p(bi.ptr), C.c_void_p(bo.ptr),
                                                            int(cplx), bi.where))
        self._finish(bo)
        return g

    def deconvolve_(self, f, g):
        self._rebind_stream()
        bi = self._in(f, self.N, self.cT, "f")
        bo = self._out(g, self.Ñ, self.cT, "g")
        self._same_side(bi, bo, None)
        _check(self._h, self._L.nfftb200_deconvolve(self._h, C.c_void_p(bi.ptr), C.c_void_p(bo.ptr), bi.where))
        self._finish(bo)
        return g

    def deconvolve_transpose_(self, g, f):
        self._rebind_stream()
        bi = self._in(g, self.Ñ, self.cT, "g")
        bo = self._out(f, self.N, self.cT, "f")
        self._same_side(bi, bo, None)
        _check(self._h, self._L.nfftb200_deconvolve_transpose(self._h, C.c_void_p(bi.ptr), C.c_void_p(bo.ptr), bi.where))
        self._finish(bo)
        return f

    def fft_(self, direction):
        self._rebind_stream()
        _check(self._h, self._L.nfftb200_fft(self._h, int(direction)))

    # ---- multi-GPU -----------------------------------------------------------------------------------
    @property
    def fused_peer_spread(self) -> bool:
        """True if the node-sharded adjoint uses the fused spread + slab gather over peer memory"""
        return bool(self._L.nfftb200_comm_is_fused(self._h) & 1)

    @property
    def fused_peer_interp(self) -> bool:
        """True if the node-sharded forward interpolates from the ranks' slabs over peer memory (no all-gather)"""
        return bool(self._L.nfftb200_comm_is_fused(self._h) & 2)

    def comm_init(self, unique_id: bytes, rank: int, nranks: int, mode: int):
        buf = C.create_string_buffer(unique_id, 128)
        _check(self._h, self._L.nfftb200_comm_init(self._h, buf, rank, nranks, mode))


class AdjointPlan:
    """Adjoint{Complex{T}, <:Plan} (AbstractNFFTs/src/interface.jl:118-119)"""

    def __init__(self, parent):
        self.parent = parent

    def size_in(self):
        return self.parent.size_out()

    def size_out(self):
        return self.parent.size_in()

    def adjoint(self):
        return self.parent

    def __mul__(self, fHat):
        p = self.parent
        dev = _is_torch(fHat) and fHat.is_cuda
        out = _empty_fortran(p.size_in(), p.cT, p.device, dev)
        return p.mul_adjoint(out, fHat)

    __matmul__ = __mul__

    def copy(self):
        return AdjointPlan(self.parent.copy())

    __copy__ = copy

    def __repr__(self):
        return "Adjoint of " + repr(self.parent)


# ---- free functions with the reference's names -----------------------------------------------------
def plan_nfft(k, N, **kw):
    """plan_nfft(k, N; m, σ, reltol, window, precompute, blockSize, ntransforms, timing, ...)
    (AbstractNFFTs/src/derived.jl:7-36, src/NFFT.jl:49-58)"""
    return B200NFFTPlan(k, N, **kw)


def nodes_(p, k):
    return p.nodes_(k)


def size_in(p):
    return p.size_in()


def size_out(p):
    return p.size_out()


def adjoint(p):
    return p.adjoint()


def mul_(out, p, x, timing=None, verbose=False, async_host=False):
    """mul!(fHat, p, f) / mul!(f, adjoint(p), fHat).  async_host=True queues upload, transform and download of
    page-locked numpy buffers and returns at once (NFFTB200_HOST_ASYNC); call p.sync() before touching them."""
    if isinstance(p, AdjointPlan):
        return p.parent.mul_adjoint(out, x, timing=timing, verbose=verbose, async_host=async_host)
    return p.mul_forward(out, x, timing=timing, verbose=verbose, async_host=async_host)


def convolve_(p, g, fHat):
    return p.convolve_(g, fHat)


def convolve_transpose_(p, fHat, g):
    return p.convolve_transpose_(fHat, g)


def deconvolve_(p, f, g):
    return p.deconvolve_(f, g)


def deconvolve_transpose_(p, g, f):
    return p.deconvolve_transpose_(g, f)


def sdc(p, iters=20, device=False):
    """NFFTTools.sdc (NFFTTools/src/samplingDensity.jl:59-155): Pipe-Menon density compensation weights through
    the real-valued convolve_transpose!/convolve! pair, followed by the least-squares global scaling.  The whole
    iteration runs inside the library on the device (nfftb200_sdc); the result is a numpy vector, or a CUDA tensor
    with device=True."""
    if device:
        w = torch.empty(p.J, dtype=_torch_dtype(p.T), device=f"cuda:{p.device}")
        _check(p._h, p._L.nfftb200_sdc(p._h, int(iters), C.c_void_p(w.data_ptr()), DEVICE))
        return w
    w = np.empty(p.J, dtype=p.T)
    try:
        _check(p._h, p._L.nfftb200_sdc(p._h, int(iters), C.c_void_p(w.ctypes.data), HOST))
    except ArgumentError as e:
        if "non-positive weights" in str(e):
            raise ValueError("non-positive weights") from e       # the reference throws this string
        raise
    return w


def sdc_host_loop(p, iters=20):
    """The same algorithm written against the public operators (host work buffers), as NFFTTools does it for any
    AbstractNFFTPlan; kept as the cross-check of the native path."""
    T, J = p.T, p.J
    weights = np.ones(J, dtype=T)
    tmp = np.empty(J, dtype=T)
    workg = np.empty(p.Ñ, dtype=T, order="F")
    scaling = None
    for i in range(iters):
        convolve_transpose_(p, weights, workg)
        if i == 0:
            scaling = workg.max()
        workg /= scaling
        convolve_(p, workg, tmp)
        tmp /= scaling
        if np.any(tmp <= 0):
            raise ValueError("non-positive weights")
        weights /= tmp
    u = np.ones(p.N, dtype=p.cT, order="F")
    workf = (p * u) * weights
    v = p.adjoint() * workf
    c = float(np.real(v.sum()) / np.sum(np.abs(v) ** 2))
    return (weights * T(c)).astype(T)


def nfft(k, f, **kw):
    """derived.jl:140-145"""
    p = plan_nfft(k, tuple(f.shape), **kw)
    return p * f


def nfft_adjoint(k, N, fHat, **kw):
    """derived.jl:147-153"""
    p = plan_nfft(k, N, **kw)
    return p.adjoint() * fHat


# ---- multi-GPU helpers --------------------------------------------------------------------------------
def shard_batch(B, rank, world):
    """transforms [lo, hi) owned by `rank` under batch sharding"""
    if B % world:
        raise ArgumentError(f"ntransforms={B} is not divisible by the number of ranks {world}")
    per = B // world
    return rank * per, (rank + 1) * per


def partition_tiles(tile_start, nranks):
    """tile-aligned node ranges for node sharding; returns nranks+1 tile boundaries"""
    ts = np.ascontiguousarray(tile_start, dtype=np.int64)
    out = np.empty(nranks + 1, dtype=np.int64)
    st = _lib.lib().nfftb200_partition_tiles(ts.ctypes.data_as(C.c_void_p), ts.size - 1, int(nranks),
                                             out.ctypes.data_as(C.c_void_p))
    _check(None, st)
    return out


def broadcast_unique_id(L, rank, process_group=None, device=0):
    """rank 0 creates the 128-byte ncclUniqueId, torch.distributed carries it to the other ranks"""
    import torch.distributed as dist
    buf = C.create_string_buffer(128)
    if rank == 0:
        _check(None, L.nfftb200_comm_unique_id(buf))
    backend = dist.get_backend(process_group)
    dev = f"cuda:{device}" if "nccl" in str(backend) else "cpu"
    t = torch.tensor(list(buf.raw), dtype=torch.uint8, device=dev)
    dist.broadcast(t, src=0, group=process_group)
    return bytes(t.cpu().tolist())


# ---- helpers ----------------------------------------------------------------------------------------
def _torch_dtype(dt):
    return {np.float32: torch.float32, np.float64: torch.float64, np.complex64: torch.complex64,
            np.complex128: torch.complex128}[np.dtype(dt).type]


def _to_complex(x):
    if _is_torch(x):
        return x.to(torch.complex64 if x.dtype == torch.float32 else torch.complex128)
    return np.asarray(x).astype(np.complex64 if np.asarray(x).dtype == np.float32 else np.complex128)


def _empty_fortran(shape, dtype, device, on_device):
    shape = tuple(int(s) for s in shape)
    if on_device:
        t = torch.empty(shape[::-1], dtype=_torch_dtype(dtype), device=f"cuda:{device}")
        return t.permute(*range(len(shape) - 1, -1, -1)) if len(shape) > 1 else t
    return np.empty(shape, dtype=dtype, order="F")


class _CudaArrayHolder:
    def __init__(self, ptr, shape, dtype, owner):
        typestr = np.dtype(dtype).str
        n = int(np.prod(shape))
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 2, "strides": None}
        self.owner = owner


def _device_view(ptr, shape, dtype, device, owner):
    holder = _CudaArrayHolder(ptr, shape, dtype, owner)
    flat = torch.as_tensor(holder, device=f"cuda:{device}")
    shape = tuple(int(s) for s in shape)
    return flat.view(shape[::-1]).permute(*range(len(shape) - 1, -1, -1)) if len(shape) > 1 else flat
