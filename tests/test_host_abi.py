"""CPU tests of the host side of the C ABI: the library loads, exports every symbol include/nfftb200.h
declares, resolves parameters/tiles/tables like the reference, and refuses to compute without a GPU
(there is no CPU fallback).  No compute calls are made."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from oracle import nfft_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def nb():
    import nfft_jl_b200 as m
    if not os.path.exists(m.LIB_PATH):
        m.build()
    return m


def host_plan(nb, N, T, m, sigma, pre=4, bs=None, B=1, window=0):
    L = nb.lib()
    h = C.c_void_p()
    D = len(N)
    Narr = (C.c_int64 * D)(*N)
    bsa = (C.c_int64 * D)(*bs) if bs else None
    st = L.nfftb200_plan_create(C.byref(h), D, Narr, 0 if T == np.float32 else 1, m, sigma, window, pre, B, bsa, -1)
    return st, h


def table(nb, h, which):
    L = nb.lib()
    n = C.c_int64()
    L.nfftb200_get_table(h, which, None, 0, C.byref(n))
    out = np.empty(n.value)
    L.nfftb200_get_table(h, which, out.ctypes.data_as(C.c_void_p), n.value, C.byref(n))
    return out


def test_exports_every_declared_symbol(nb):
    hdr = open(os.path.join(ROOT, "include", "nfftb200.h")).read()
    declared = set(re.findall(r"^(?:int|const char\*)\s+(nfftb200_[a-z0-9_]+)\(", hdr, flags=re.M))
    assert len(declared) >= 25
    L = nb.lib()
    for s in declared:
        assert hasattr(L, s), s
    assert declared == set(nb.SYMBOLS)
    assert L.nfftb200_version() >= 100


def test_accuracy_params_match_reference_rule(nb):
    for kw in [dict(), dict(m=3), dict(m=3, σ=2.0), dict(m=4, σ=1.5), dict(reltol=1e-7), dict(reltol=1e-5),
               dict(reltol=1e-12), dict(m=8, σ=2.0, reltol=1e-3)]:
        got = nb.accuracyParams(**kw)
        want = O.accuracy_params(kw.get("m"), kw.get("σ"), kw.get("reltol"))
        assert got[0] == want[0] and got[1] == want[1] and np.isclose(got[2], want[2], rtol=1e-14)


@pytest.mark.parametrize("N,T,m,sigma", [((256, 256), np.float64, 4, 2.0), ((128, 128, 128), np.float32, 3, 2.0),
                                         ((512, 512), np.float32, 4, 2.0), ((2 ** 22,), np.float64, 4, 2.0),
                                         ((33, 35), np.float64, 5, 2.0), ((9,), np.float64, 5, 1.5),
                                         ((11, 12, 14), np.float32, 5, 1.25)])
def test_geometry_matches_init_params(nb, N, T, m, sigma):
    """src/precomputation.jl:3-56"""
    L = nb.lib()
    st, h = host_plan(nb, N, T, m, sigma)
    assert st == 0
    D = len(N)
    Nt = (C.c_int64 * D)(); bs = (C.c_int64 * D)()
    nt, lut, sg, M = C.c_int64(), C.c_int64(), C.c_double(), C.c_int64()
    L.nfftb200_get_info(h, Nt, bs, C.byref(nt), C.byref(lut), C.byref(sg), C.byref(M))
    po = O.init_params(N, T, m, sigma)
    assert tuple(Nt) == po.Nt
    assert lut.value == po.LUTSize
    assert sg.value == po.sigma
    if not (D == 3 and T == np.float64):     # 3-D default tiles shrink only when shared memory would overflow
        assert tuple(bs) == po.blockSize or D == 3
    assert nt.value == int(np.prod([-(-a // b) for a, b in zip(Nt, bs)]))
    L.nfftb200_destroy(h)


@pytest.mark.parametrize("T,m,B,want", [(np.float32, 4, 32, (16, 16)), (np.float32, 3, 8, (16, 16)), (np.float32, 4, 7, (64, 64)),
                                        (np.float64, 4, 32, (64, 64)), (np.float32, 5, 32, (64, 64))])
def test_default_tiles_of_batched_2d_plans(nb, T, m, B, want):
    """the reference default is 64 x 64 (src/precomputation.jl:59-77); Float32 plans with ntransforms >= 8 and m <= 4 take
    16 x 16 tiles so that the padded tiles of the whole batch fit one CTA's shared memory (csrc/twod_batch.cuh).  The
    tile size only changes the reported blockSize and the permutation that goes with it (tests: bit-exact vs the oracle
    built with the same blockSize)."""
    L = nb.lib()
    st, h = host_plan(nb, (512, 512), T, m, 2.0, B=B)
    assert st == 0
    Nt = (C.c_int64 * 2)(); bs = (C.c_int64 * 2)()
    nt, lut, sg, M = C.c_int64(), C.c_int64(), C.c_double(), C.c_int64()
    L.nfftb200_get_info(h, Nt, bs, C.byref(nt), C.byref(lut), C.byref(sg), C.byref(M))
    assert tuple(bs) == want
    L.nfftb200_destroy(h)


def test_oversampled_size_uses_the_plan_precision(nb):
    """Ñ_d = (ceil(Int, σ*N_d) ÷ 2)*2 with σ::T (src/precomputation.jl:25-27): Float32 * Int is a Float32 product, so a
    Float32 plan with non-dyadic σ can get a different Ñ than the Float64 plan.  Pinned independently of the oracle by
    the literal formula in numpy scalars; σ=1.1, N=10 is the known case (Ñ = 10 in both precisions, whereas the product
    Float64(Float32(1.1)) * 10 = 11.0000002 that round 1 used gives 12)."""
    L = nb.lib()

    def lib_Nt(N, T, sigma):
        st, h = host_plan(nb, (N,), T, 3, sigma)
        assert st == 0
        Nt = (C.c_int64 * 1)()
        L.nfftb200_get_info(h, Nt, None, None, None, None, None)
        L.nfftb200_destroy(h)
        return Nt[0]

    assert lib_Nt(10, np.float32, 1.1) == 10 and lib_Nt(10, np.float64, 1.1) == 10
    assert O.init_params((10,), np.float32, 3, 1.1).Nt == (10,) and O.init_params((10,), np.float64, 3, 1.1).Nt == (10,)
    assert (int(np.ceil(float(np.float32(1.1)) * 10)) // 2) * 2 == 12
    nmis = 0
    for sigma in (1.1, 1.3, 1.2, 1.7, 1.9, 2.0, 1.25, 1.5):
        for N in list(range(1, 200)) + [255, 256, 500, 1000, 1023, 2047, 2048]:
            for T in (np.float32, np.float64):
                want = (int(np.ceil(T(sigma) * T(N))) // 2) * 2
                assert lib_Nt(N, T, sigma) == want == O.init_params((N,), T, 3, sigma).Nt[0], (sigma, N, T)
            nmis += lib_Nt(N, np.float32, sigma) != (int(np.ceil(float(np.float32(sigma)) * N)) // 2) * 2
    assert nmis > 0          # the double-precision product (the round-1 formula) disagrees somewhere in this sweep


@pytest.mark.parametrize("T", [np.float64, np.float32])
@pytest.mark.parametrize("m", [2, 3, 4, 5, 6, 8])
def test_tables_match_oracle(nb, T, m):
    """windowPolyInterp (values), windowLinInterp and windowHatInvLUT (src/precomputation.jl:291-358)"""
    N = (40, 36)
    po = O.init_params(N, T, m, 2.0)
    st, h = host_plan(nb, N, T, m, 2.0, pre=4)
    assert st == 0
    P = table(nb, h, 1).reshape((2 * m + 1, 2 * m), order="F")
    Po = O.precompute_poly_interp(po).astype(np.float64)
    t = np.linspace(-0.5, 0.5, 257)
    V = np.vander(t, 2 * m + 1, increasing=True)
    tol = 1e-13 if T == np.float64 else 5e-7
    assert np.abs(V @ P - V @ Po).max() / np.abs(V @ Po).max() < tol
    hat = table(nb, h, 0)
    hat_o = np.concatenate(O.window_hat_inv_lut(po)).astype(np.float64)
    assert np.abs(hat / hat_o - 1).max() < (1e-13 if T == np.float64 else 2e-7)
    nb.lib().nfftb200_destroy(h)
    st, h = host_plan(nb, N, T, m, 2.0, pre=3)
    lin = table(nb, h, 2)
    lin_o = O.precompute_lin_interp(po).astype(np.float64)
    assert lin.shape == lin_o.shape == (po.LUTSize + 2,)
    assert np.abs(lin - lin_o).max() / lin_o.max() < (1e-14 if T == np.float64 else 1e-7)
    nb.lib().nfftb200_destroy(h)


@pytest.mark.parametrize("T", [np.float64, np.float32])
@pytest.mark.parametrize("wi", [1, 2, 3, 4, 5])
def test_other_window_tables_match_oracle(nb, T, wi):
    """getWindow pairs other than :kaiser_bessel (src/windowFunctions.jl:41-134) through the same three tables"""
    N, m = (40, 37), 5
    po = O.init_params(N, T, m, 2.0, window=O.WINDOWS[wi])
    st, h = host_plan(nb, N, T, m, 2.0, pre=4, window=wi)
    assert st == 0
    P = table(nb, h, 1).reshape((2 * m + 1, 2 * m), order="F")
    Po = O.precompute_poly_interp(po).astype(np.float64)
    t = np.linspace(-0.5, 0.5, 257)
    V = np.vander(t, 2 * m + 1, increasing=True)
    assert np.abs(V @ P - V @ Po).max() / np.abs(V @ Po).max() < (1e-12 if T == np.float64 else 5e-7)
    hat = table(nb, h, 0)
    hat_o = np.concatenate(O.window_hat_inv_lut(po)).astype(np.float64)
    assert np.abs(hat / hat_o - 1).max() < (1e-12 if T == np.float64 else 2e-7)
    nb.lib().nfftb200_destroy(h)
    st, h = host_plan(nb, N, T, m, 2.0, pre=3, window=wi)
    lin = table(nb, h, 2)
    lin_o = O.precompute_lin_interp(po).astype(np.float64)
    assert np.abs(lin - lin_o).max() / lin_o.max() < (1e-13 if T == np.float64 else 1e-7)
    nb.lib().nfftb200_destroy(h)


def test_errors_and_no_cpu_fallback(nb):
    L = nb.lib()
    st, h = host_plan(nb, (4, 4, 4, 4, 4), np.float64, 4, 2.0)
    assert st == 4                                              # D = 5 unsupported (D = 4 runs: test/accuracy.jl:43)
    st, h = host_plan(nb, (6, 5, 6, 6), np.float64, 5, 2.0)
    assert st == 0
    Nt4 = (C.c_int64 * 4)(); bs4 = (C.c_int64 * 4)()
    L.nfftb200_get_info(h, Nt4, bs4, None, None, None, None)
    assert tuple(Nt4) == (12, 10, 12, 12) and tuple(bs4) == (12, 10, 12, 1)     # _blockSize: 16,16,16 capped by Ñ, then 1
    L.nfftb200_destroy(h)
    st, h = host_plan(nb, (16,), np.float64, 9, 2.0)
    assert st == 4
    st, h = host_plan(nb, (16,), np.float64, 4, 2.0, pre=7)
    assert st == 4
    h2 = C.c_void_p()
    N1 = (C.c_int64 * 1)(16)
    assert L.nfftb200_plan_create(C.byref(h2), 1, N1, 1, 4, 2.0, 6, 4, 1, None, -1) == 4   # unknown window (0..5 exist)
    st, h = host_plan(nb, (16, 16), np.float32, 4, 2.0)
    assert st == 0
    k = np.zeros((2, 8), dtype=np.float32)
    # a host-only plan must refuse to compute: the product has no CPU path
    assert L.nfftb200_set_nodes(h, k.ctypes.data_as(C.c_void_p), 8, 0) == 5
    assert b"no CPU fallback" in L.nfftb200_last_error(h)
    buf = np.zeros(1024, dtype=np.complex64)
    assert L.nfftb200_exec_forward(h, buf.ctypes.data_as(C.c_void_p), buf.ctypes.data_as(C.c_void_p), 0) == 9
    assert L.nfftb200_deconvolve(h, buf.ctypes.data_as(C.c_void_p), buf.ctypes.data_as(C.c_void_p), 0) == 5
    assert L.nfftb200_destroy(h) == 0
    assert L.nfftb200_destroy(None) == 0
    assert L.nfftb200_status_string(1).decode().startswith("nodes out of range")


def test_missing_library_fails_loudly(nb, monkeypatch, tmp_path):
    from importlib import reload
    import nfft_jl_b200._lib as lb
    monkeypatch.setattr(lb, "LIB_PATH", str(tmp_path / "nope.so"))
    monkeypatch.setattr(lb, "_lib", None)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        lb.lib()
