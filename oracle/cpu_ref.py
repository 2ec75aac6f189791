"""
oracle/cpu_ref.py -- ctypes driver of oracle/_build/libnfft_ref.so (the C/OpenMP restatement of
the reference's default blocked CPU path) + scipy.fft (pocketfft) for the FFT stage.

TEST INFRASTRUCTURE ONLY: used by tests/ (cross-check of the numpy oracle) and by bench.py's
`cpu_baseline` / `--impl reference` legs.  cpu_baseline.kind = "port" -- this is the
reference *algorithm* restated, not NFFT.jl itself (no julia binary exists in this image).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import time

import numpy as np
from scipy import fft as _sfft

from . import nfft_oracle as O

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force=False):
    so = os.path.join(_HERE, "_build", "libnfft_ref.so")
    if force or not os.path.exists(so):
        subprocess.check_call(["make", "-C", _HERE], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.ref_num_threads.restype = C.c_int
        _LIB.ref_set_num_threads.argtypes = [C.c_int]
    return _LIB


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class CpuRefPlan:
    """Blocked POLYNOMIAL NFFTPlan restated (src/implementation.jl:73-193)."""

    def __init__(self, k, N, m=None, sigma=None, reltol=None, blockSize=None, workers=None):
        k = np.asarray(k)
        self.T = k.dtype.type
        self.suf = "_f32" if self.T == np.float32 else "_f64"
        self.cT = np.complex64 if self.T == np.float32 else np.complex128
        self.p = O.init_params(N, self.T, m, sigma, reltol, O.POLYNOMIAL, blockSize)
        self.D = self.p.D
        self.workers = workers or os.cpu_count()
        self.P = np.asfortranarray(O.precompute_poly_interp(self.p))       # (2m+1) x 2m, column-major
        self.lut = np.concatenate(O.window_hat_inv_lut(self.p)).astype(self.T)
        self.Nt64 = np.array(self.p.Nt, dtype=np.int64)
        self.N64 = np.array(self.p.N, dtype=np.int64)
        self.bs64 = np.array(self.p.blockSize, dtype=np.int64)
        self.t_pre = self.set_nodes(k)

    def set_nodes(self, k):
        t0 = time.perf_counter()
        O.check_nodes(k)
        self.k = np.ascontiguousarray(k, dtype=self.T)                    # (M, D) == Julia D x M
        self.M = self.k.shape[0]
        nb = int(np.prod([-(-self.p.Nt[d] // self.p.blockSize[d]) for d in range(self.D)]))
        self.perm = np.empty(self.M, dtype=np.int64)
        self.blockStart = np.empty(nb + 1, dtype=np.int64)
        self.xs = np.empty_like(self.k)
        L = lib()
        getattr(L, "ref_precompute_blocks" + self.suf)(
            _p(self.k), C.c_int(self.D), C.c_int64(self.M), _p(self.Nt64), _p(self.bs64),
            _p(self.perm), _p(self.blockStart), _p(self.xs))
        self.y = np.empty((self.M, self.D), dtype=np.int32)
        self.t = np.empty((self.M, self.D), dtype=self.T)
        getattr(L, "ref_precompute_idx" + self.suf)(
            _p(self.xs), C.c_int(self.D), C.c_int64(self.M), _p(self.Nt64), _p(self.bs64),
            C.c_int(self.p.m), _p(self.perm), _p(self.blockStart), _p(self.y), _p(self.t))
        return time.perf_counter() - t0

    def _args(self):
        return (C.c_int(self.D), C.c_int64(self.M), _p(self.Nt64), _p(self.bs64), C.c_int(self.p.m),
                _p(self.P), _p(self.perm), _p(self.blockStart), _p(self.y), _p(self.t))

    def convolve(self, g):
        g = np.asfortranarray(g, dtype=self.cT)
        out = np.empty(self.M, dtype=self.cT)
        getattr(lib(), "ref_convolve_blocking" + self.suf)(_p(g), _p(out), *self._args())
        return out

    def convolve_transpose(self, fHat):
        fHat = np.ascontiguousarray(fHat, dtype=self.cT)
        g = np.empty(self.p.Nt, dtype=self.cT, order="F")
        getattr(lib(), "ref_convolve_transpose_blocking" + self.suf)(_p(fHat), _p(g), *self._args())
        return g

    def deconvolve(self, f):
        f = np.asfortranarray(f, dtype=self.cT)
        g = np.empty(self.p.Nt, dtype=self.cT, order="F")
        getattr(lib(), "ref_deconvolve" + self.suf)(_p(f), _p(g), C.c_int(self.D), _p(self.N64),
                                                   _p(self.Nt64), _p(self.lut), C.c_int(0))
        return g

    def deconvolve_transpose(self, g):
        g = np.asfortranarray(g, dtype=self.cT)
        f = np.empty(self.p.N, dtype=self.cT, order="F")
        getattr(lib(), "ref_deconvolve" + self.suf)(_p(f), _p(g), C.c_int(self.D), _p(self.N64),
                                                   _p(self.Nt64), _p(self.lut), C.c_int(1))
        return f

    def forward(self, f, timing=None):
        t0 = time.perf_counter()
        g = self.deconvolve(f)
        t1 = time.perf_counter()
        g = _sfft.fftn(g, workers=self.workers, overwrite_x=True)
        t2 = time.perf_counter()
        out = self.convolve(g)
        t3 = time.perf_counter()
        if timing is not None:
            timing.update(deconv=t1 - t0, fft=t2 - t1, conv=t3 - t2)
        return out

    def adjoint(self, fHat, timing=None):
        t0 = time.perf_counter()
        g = self.convolve_transpose(fHat)
        t1 = time.perf_counter()
        g = _sfft.ifftn(g, norm="forward", workers=self.workers, overwrite_x=True)
        t2 = time.perf_counter()
        f = self.deconvolve_transpose(g)
        t3 = time.perf_counter()
        if timing is not None:
            timing.update(conv_adjoint=t1 - t0, fft_adjoint=t2 - t1, deconv_adjoint=t3 - t2)
        return f
