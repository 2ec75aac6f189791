"""Toeplitz (Gram) operator of NFFTTools on the device -- the executable Python mirror of
NFFTTools/src/Toeplitz.jl (file:line under /root/reference):

    calculateToeplitzKernel(shape, tr; m=4, σ=2.0, window=:kaiser_bessel)      :86-93
    calculateToeplitzKernel!(f, p, tr, fftplan)                                :131-137
    convolveToeplitzKernel!(y, λ, fftplan, ifftplan, xOS1, xOS2)               :230-244

numpy arrays are HOST buffers, torch CUDA tensors DEVICE buffers (Fortran strides), exactly like plan.py.
Everything runs in libnfftb200.so (nfftb200_toeplitz_*); there is no CPU path."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from .plan import (DEVICE, HOST, ArgumentError, B200NFFTPlan, DimensionMismatch, _check, _empty_fortran,
                   _fortran_strides, _is_torch, _torch_dtype, plan_nfft)

try:
    import torch
except Exception:  # pragma: no cover
    torch = None


def _tcheck(L, handle, st):
    # every nfftb200_toeplitz_* failure also records its message as the thread's last error
    _check(None, st)


def calculateToeplitzKernel_(f, p: B200NFFTPlan, tr):
    """calculateToeplitzKernel!(f, p, tr, fftplan): nodes!(p, tr); f = FFT(fftshift(adjoint(p) * ones)).
    `p` is a plan on the 2x oversampled image grid; `f` (size p.N, Complex{T}) is overwritten and returned."""
    p.nodes_(tr)
    if tuple(f.shape) != tuple(p.N):
        raise DimensionMismatch(f"Toeplitz kernel has size {tuple(f.shape)} != {tuple(p.N)}")
    buf = p._out(f, p.N, p.cT, "f")
    _check(p._h, p._L.nfftb200_toeplitz_kernel(p._h, C.c_void_p(buf.ptr), buf.where))
    p._finish(buf)
    return f


def calculateToeplitzKernel(shape, tr, *, m=4, σ=None, sigma=None, window="kaiser_bessel", **kw):
    """calculateToeplitzKernel(shape, tr; m=4, σ=2.0, window, kwargs...) -> array of size 2 .* shape.
    The result lives where `tr` lives (numpy -> numpy, CUDA tensor -> CUDA tensor)."""
    if σ is None:
        σ = 2.0 if sigma is None else sigma
    shape = tuple(int(s) for s in shape)
    shape_os = tuple(2 * s for s in shape)
    p = plan_nfft(tr, shape_os, m=m, σ=σ, window=window, **kw)
    on_dev = _is_torch(tr) and tr.is_cuda
    lam = _empty_fortran(shape_os, p.cT, p.device, on_dev)
    _check(p._h, p._L.nfftb200_toeplitz_kernel(p._h, C.c_void_p(lam.data_ptr() if on_dev else lam.ctypes.data),
                                               DEVICE if on_dev else HOST))
    if on_dev:
        p.sync()
    p.destroy()
    return lam


class ToeplitzOperator:
    """The pre-planned form of convolveToeplitzKernel!: owns fftplan, ifftplan, xOS1/xOS2 (Toeplitz.jl:230-235)
    and a device copy of λ.  `ntransforms=B` applies the same kernel to y of size (shape..., B)."""

    def __init__(self, λ, *, ntransforms=1, device=None, stream="current"):
        self._L = _lib.lib()
        self._h = C.c_void_p()
        os_shape = tuple(int(s) for s in λ.shape)
        if any(s % 2 for s in os_shape):
            raise ArgumentError("the Toeplitz kernel must have size 2 .* shape")
        self.shape = tuple(s // 2 for s in os_shape)
        self.D = len(self.shape)
        self.ntransforms = int(ntransforms)
        on_dev = _is_torch(λ) and λ.is_cuda
        if on_dev:
            single = λ.dtype == torch.complex64
        else:
            λ = np.asarray(λ)
            single = λ.dtype == np.complex64
        self.cT = np.complex64 if single else np.complex128
        if device is None:
            device = λ.device.index if on_dev else (
                torch.cuda.current_device() if (torch is not None and torch.cuda.is_available()) else 0)
        self.device = int(device)
        shp = (C.c_int64 * self.D)(*self.shape)
        _tcheck(self._L, None, self._L.nfftb200_toeplitz_create(C.byref(self._h), self.D, shp, 0 if single else 1,
                                                                self.ntransforms, self.device))
        if stream == "current" and torch is not None and torch.cuda.is_available():
            self._L.nfftb200_toeplitz_set_stream(self._h, C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream))
        self.set_kernel(λ)

    def set_kernel(self, λ):
        os_shape = tuple(2 * s for s in self.shape)
        if tuple(λ.shape) != os_shape:
            raise DimensionMismatch(f"Toeplitz kernel has size {tuple(λ.shape)} != {os_shape}")
        if _is_torch(λ) and λ.is_cuda:
            if λ.dtype != _torch_dtype(self.cT):
                λ = λ.to(_torch_dtype(self.cT))
            if tuple(λ.stride()) != _fortran_strides(os_shape):
                lf = _empty_fortran(os_shape, self.cT, self.device, True)
                lf.copy_(λ)
                λ = lf
            ptr, where = λ.data_ptr(), DEVICE
        else:
            λ = np.asfortranarray(np.asarray(λ), dtype=self.cT)
            ptr, where = λ.ctypes.data, HOST
        _tcheck(self._L, self._h, self._L.nfftb200_toeplitz_set_kernel(self._h, C.c_void_p(ptr), where))
        self._keep = λ
        return self

    def _yshape(self):
        return self.shape + ((self.ntransforms,) if self.ntransforms > 1 else ())

    def apply_(self, y):
        """y <- crop(IFFT(λ .* FFT(pad(y)))) in place; returns y"""
        shape = self._yshape()
        if tuple(y.shape) != shape:
            raise DimensionMismatch(f"y has size {tuple(y.shape)} != {shape}")
        if _is_torch(y) and y.is_cuda:
            if y.dtype != _torch_dtype(self.cT) or (tuple(y.stride()) != _fortran_strides(shape) and y.numel() > 1):
                raise ArgumentError("y must be a Complex{T} CUDA tensor with Fortran-order strides")
            _tcheck(self._L, self._h, self._L.nfftb200_toeplitz_apply(self._h, C.c_void_p(y.data_ptr()), DEVICE))
            return y
        if not isinstance(y, np.ndarray):
            raise ArgumentError("y must be a numpy array or a CUDA tensor")
        if y.dtype == self.cT and (y.flags.f_contiguous or (y.ndim <= 1 and y.flags.c_contiguous)) and y.flags.writeable:
            _tcheck(self._L, self._h, self._L.nfftb200_toeplitz_apply(self._h, C.c_void_p(y.ctypes.data), HOST))
            return y
        tmp = np.asfortranarray(y, dtype=self.cT).copy(order="F")
        _tcheck(self._L, self._h, self._L.nfftb200_toeplitz_apply(self._h, C.c_void_p(tmp.ctypes.data), HOST))
        y[...] = tmp
        return y

    __call__ = apply_

    def sync(self):
        self._L.nfftb200_toeplitz_sync(self._h)

    def destroy(self):
        if getattr(self, "_h", None) is not None and self._h:
            self._L.nfftb200_toeplitz_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass


def convolveToeplitzKernel_(y, λ, op: ToeplitzOperator | None = None):
    """convolveToeplitzKernel!(y, λ[, fftplan, ifftplan, xOS1, xOS2]).  Without `op` the plans and work arrays are
    built per call (the reference's default arguments do the same); pass a ToeplitzOperator to reuse them."""
    if op is None:
        op = ToeplitzOperator(λ)
        out = op.apply_(y)
        if _is_torch(y) and y.is_cuda:
            op.sync()
        op.destroy()
        return out
    return op.apply_(y)
