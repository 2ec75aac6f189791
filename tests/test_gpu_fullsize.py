"""Parity at the FULL sizes of BASELINE.json's configs: the CUDA path (through the C ABI) against the C/OpenMP twin of
the reference's blocked algorithm (oracle/cpu_ref.CpuRefPlan, /root/reference/src/convolution.jl:229-492 restated) on
identical PCG64-seeded inputs -- every node value and the whole image, not a subsample.

Tolerances are the north_star's: node permutation bit-exact, relative L2 <= 1e-12 (Float64) / <= 1e-5 (Float32).
The twin is itself cross-checked against the numpy oracle (1e-16, identical permutation) by the CPU suite
(tests/test_oracle_pins.py); it is used here because the numpy oracle needs minutes at these sizes and the twin seconds.
Configs: C1 2-D 256^2 / 65 536 / m=4 / F64;  C2 3-D 128^3 / 2^21 / m=3 / F32 (default kernels AND
the register-window kernels of kernel_mode 7);  C3 radial 1024 spokes x 1024 samples, N=512^2, ntransforms=32, density
weighted adjoint, per transform;  C4 1-D N=2^22 / M=2^25 / F64;  C5 3-D 256^3 / M=2^27 / F32 (M=2^25 when the host has
less than 48 GB free: the twin needs ~12 GB for 2^27 nodes)."""
import numpy as np
import pytest

from oracle import nfft_oracle as O

pytestmark = pytest.mark.gpu

TOL = {np.float64: 1e-12, np.float32: 1e-5}


def rel(a, b):
    a = np.asarray(a).ravel()
    b = np.asarray(b).ravel()
    num = np.linalg.norm((a.astype(np.complex128) - b.astype(np.complex128)))
    return float(num / np.linalg.norm(b.astype(np.complex128)))


@pytest.fixture(scope="module")
def nb():
    import nfft_jl_b200 as m
    m.lib()          # fails loudly if libnfftb200.so is missing
    return m


@pytest.fixture(scope="module")
def twin():
    from oracle import cpu_ref
    cpu_ref.lib().ref_set_num_threads(len(__import__("os").sched_getaffinity(0)))
    return cpu_ref.CpuRefPlan


def gpu_plan(nb, k, N, m, **kw):
    import torch
    return nb.plan_nfft(torch.from_numpy(np.ascontiguousarray(k.T)).cuda(), N, m=m, σ=2.0, **kw)


def check_pair(nb, twin, k, N, m, T, modes=(0,), seeds=(2, 3)):
    p = gpu_plan(nb, k, N, m)
    pc = twin(k, N, m=m, sigma=2.0, blockSize=p.params.blockSize)
    assert p.Ñ == pc.p.Nt
    perm, ts = p.permutation()
    assert np.array_equal(perm, pc.perm), "node permutation is not bit-exact"
    assert np.array_equal(ts, pc.blockStart)
    f = O.random_complex(N, T, seeds[0])
    fh = O.random_complex(k.shape[0], T, seeds[1])
    want_f, want_a = pc.forward(f), pc.adjoint(fh)
    errs = {}
    for mode in modes:
        p.set_kernel_mode(mode)
        errs[mode] = (rel(p * f, want_f), rel(p.adjoint() * fh, want_a))
        assert errs[mode][0] <= TOL[T] and errs[mode][1] <= TOL[T], (mode, errs)
    return errs


def test_c1_full(nb, twin):
    T = np.float64
    k = O.random_nodes(65536, 2, T, seed=1)
    check_pair(nb, twin, k, (256, 256), 4, T)


@pytest.mark.parametrize("mode", [0, 7, 9])
def test_c2_full(nb, twin, mode):
    T = np.float32
    k = O.random_nodes(2 ** 21, 3, T, seed=1)
    check_pair(nb, twin, k, (128, 128, 128), 3, T, modes=(mode,))


def test_c3_full_radial_batched_density_weighted(nb, twin):
    """1024 spokes x 1024 samples, 32 coils: forward of 32 images and the density-compensated adjoint of 32 data
    vectors, each transform compared with the twin run on that transform alone"""
    T = np.float32
    N, B = (512, 512), 32
    k = O.radial_nodes(1024, 1024, T)
    M = k.shape[0]
    r = np.hypot(k[:, 0].astype(np.float64), k[:, 1].astype(np.float64))
    w = np.maximum(r, 1.0 / 2048)
    w = (w / w.sum() * M).astype(T)                                  # ramp density weights, mean 1
    p = gpu_plan(nb, k, N, 4, ntransforms=B)
    pc = twin(k, N, m=4, sigma=2.0, blockSize=p.params.blockSize)
    assert np.array_equal(p.permutation()[0], pc.perm)
    f = O.random_complex(N + (B,), T, 4)
    data = O.random_complex((M, B), T, 5)
    wdata = np.asfortranarray(data * w[:, None])
    fwd = np.asarray(p * f)
    adj = np.asarray(p.adjoint() * wdata)
    worst = 0.0
    for b in range(B):
        ef = rel(fwd[:, b], pc.forward(np.asfortranarray(f[..., b])))
        ea = rel(adj[..., b], pc.adjoint(np.ascontiguousarray(wdata[:, b])))
        worst = max(worst, ef, ea)
        assert ef <= TOL[T] and ea <= TOL[T], (b, ef, ea)
    print("C3 worst rel-L2 over 32 transforms:", worst)


def test_c4_full(nb, twin):
    T = np.float64
    k = O.random_nodes(2 ** 25, 1, T, seed=1)
    check_pair(nb, twin, k, (2 ** 22,), 4, T)


def test_c5_full(nb, twin):
    import psutil
    T = np.float32
    logM = 27 if psutil.virtual_memory().available > 48 * 2 ** 30 else 25
    rng = np.random.default_rng(1)
    k = np.empty((2 ** logM, 3), dtype=T)
    step = 2 ** 22                                                   # generate in pieces: no 3 GB float64 temporary
    for i in range(0, k.shape[0], step):
        k[i:i + step] = (rng.random((step, 3)) - 0.5).astype(T)
    errs = check_pair(nb, twin, k, (256, 256, 256), 3, T)
    print("C5 M=2^%d rel-L2 (forward, adjoint):" % logM, errs[0])
