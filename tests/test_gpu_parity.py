"""GPU parity tests proper: the CUDA path (through the C ABI, via the Python mirror of the reference's
plan API) against the CPU oracle on identical seeded inputs.

Tolerances (BASELINE.json north_star): node permutation bit-exact; relative L2 <= 1e-12 (Float64) /
<= 1e-5 (Float32) versus the reference restatement with the MATCHED window mode
(POLYNOMIAL/TENSOR <-> polynomial, LINEAR <-> LUT, FULL <-> exact window); and within the reference's own
test tolerance versus the NDFT (test/accuracy.jl:46: 1e-7 for Kaiser-Bessel at m=5, sigma=2)."""
import numpy as np
import pytest

from oracle import nfft_oracle as O

pytestmark = pytest.mark.gpu

TOL = {np.float64: 1e-12, np.float32: 1e-5}


def rel(a, b):
    a = np.asarray(a).ravel().astype(np.complex128)
    b = np.asarray(b).ravel().astype(np.complex128)
    return float(np.linalg.norm(a - b) / np.linalg.norm(b))


@pytest.fixture(scope="module")
def nb():
    import nfft_jl_b200 as m
    m.lib()          # fails loudly if libnfftb200.so is missing
    return m


SHAPES = [(255,), (31, 33), (11, 12, 14), (64, 64, 64), (256, 256), (4096,), (6, 5, 6, 6)]


@pytest.mark.parametrize("N", SHAPES)
@pytest.mark.parametrize("T", [np.float64, np.float32])
def test_permutation_bit_exact(nb, N, T):
    D = len(N)
    M = int(np.prod(N)) if np.prod(N) < 100000 else 100003
    k = O.random_nodes(M, D, T, seed=11)
    k[:5] = 0.5
    k[5:10] = -0.5
    k[10:12] = 0.0
    k[12:14] = T(-1e-12)      # tiny negative -> shifts to exactly 1 -> 1-eps (src/utils.jl:35-40)
    p = nb.plan_nfft(k.T, N, m=4, σ=2.0)
    perm, ts = p.permutation()
    po = O.init_params(N, T, 4, 2.0, blockSize=p.params.blockSize)
    perm_o, counts, _ = O.precompute_blocks(k, po)
    assert np.array_equal(perm, perm_o)
    assert np.array_equal(np.diff(ts), counts)


@pytest.mark.parametrize("bs", [(8, 8), (16, 4), (64, 64), (7, 5)])
def test_permutation_custom_block_size(nb, bs):
    N = (40, 36)
    k = O.random_nodes(5000, 2, np.float32, seed=5)
    p = nb.plan_nfft(k.T, N, m=3, σ=2.0, blockSize=bs)
    perm, ts = p.permutation()
    po = O.init_params(N, np.float32, 3, 2.0, blockSize=bs)
    perm_o, counts, _ = O.precompute_blocks(k, po)
    assert np.array_equal(perm, perm_o)
    assert np.array_equal(np.diff(ts), counts)


@pytest.mark.parametrize("N", [(255,), (31, 33), (11, 12, 14), (32, 32, 32), (6, 5, 6, 6)])
@pytest.mark.parametrize("T", [np.float64, np.float32])
@pytest.mark.parametrize("pre", [O.POLYNOMIAL, O.LINEAR, O.FULL, O.TENSOR])
@pytest.mark.parametrize("kernel_mode", [0, 1, 2])
def test_forward_adjoint_vs_oracle(nb, N, T, pre, kernel_mode):
    D = len(N)
    M = int(np.prod(N))
    m = 5
    k = O.random_nodes(M, D, T, seed=1)
    p = nb.plan_nfft(k.T, N, m=m, σ=2.0, precompute=nb.PrecomputeFlags(pre))
    p.set_kernel_mode(kernel_mode)
    po = O.OraclePlan(k, N, m=m, sigma=2.0, precompute=pre, blocking=True, blockSize=p.params.blockSize)
    fHat = O.random_complex(M, T, 2)
    f = O.random_complex(N, T, 3)
    tol = TOL[T]
    out_adj = p.adjoint() * fHat
    assert out_adj.shape == tuple(N)
    assert rel(out_adj, po.adjoint(fHat)) < tol
    out_fwd = p * f
    assert out_fwd.shape == (M,)
    assert rel(out_fwd, po.forward(f)) < tol
    if T == np.float64:       # the reference's own bar vs the NDFT (test/accuracy.jl:46)
        assert rel(out_adj, O.ndft_adjoint(k, N, fHat)) < 1e-7
        assert rel(out_fwd, O.ndft(k, f)) < 1e-7


WINDOW_EPS = {"kaiser_bessel": 1e-7, "cosh_type": 1e-7, "gauss": 1e-3, "kaiser_bessel_rev": 1e-6, "spline": 1e-4,
              "exp_sqrt": 1e-7}      # exp_sqrt: the north_star's on-the-fly option, not a reference window


@pytest.mark.parametrize("N", [(255,), (31, 33), (11, 12, 14), (6, 5, 6, 6)])
@pytest.mark.parametrize("T", [np.float64, np.float32])
@pytest.mark.parametrize("pre", [O.POLYNOMIAL, O.LINEAR, O.FULL, O.TENSOR])
@pytest.mark.parametrize("window", ["cosh_type", "gauss", "kaiser_bessel_rev", "spline", "exp_sqrt"])
def test_other_windows_vs_oracle(nb, N, T, pre, window):
    """the window x precompute matrix of test/accuracy.jl:41-73 (windows of src/windowFunctions.jl:41-134):
    matched-mode parity with the oracle and the reference's per-window tolerance versus the NDFT"""
    D, M, m = len(N), int(np.prod(N)), 5
    k = O.random_nodes(M, D, T, seed=1)
    p = nb.plan_nfft(k.T, N, m=m, σ=2.0, precompute=nb.PrecomputeFlags(pre), window=window)
    assert p.params.window == window
    po = O.OraclePlan(k, N, m=m, sigma=2.0, precompute=pre, blockSize=p.params.blockSize, window=window)
    fHat = O.random_complex(M, T, 2)
    f = O.random_complex(N, T, 3)
    out_adj = p.adjoint() * fHat
    out_fwd = p * f
    assert rel(out_adj, po.adjoint(fHat)) < TOL[T]
    assert rel(out_fwd, po.forward(f)) < TOL[T]
    if T == np.float64:
        assert rel(out_adj, O.ndft_adjoint(k, N, fHat)) < WINDOW_EPS[window]
        assert rel(out_fwd, O.ndft(k, f)) < WINDOW_EPS[window]
    with pytest.raises(NotImplementedError):
        nb.plan_nfft(k.T, N, m=m, σ=2.0, window="hann")          # src/windowFunctions.jl:16


@pytest.mark.parametrize("m", [2, 3, 4, 6, 8])
@pytest.mark.parametrize("T", [np.float64, np.float32])
def test_kernel_widths_3d(nb, m, T):
    if m == 8 and T == np.float32:
        pytest.skip("phi(0)^3 ~ (4.7e14)^3 overflows Float32 (SURVEY App. B.4); same in the reference")
    N = (20, 18, 16)
    M = 6000
    k = O.random_nodes(M, 3, T, seed=4)
    p = nb.plan_nfft(k.T, N, m=m, σ=2.0)
    po = O.OraclePlan(k, N, m=m, sigma=2.0, blockSize=p.params.blockSize)
    fHat = O.random_complex(M, T, 2)
    f = O.random_complex(N, T, 3)
    assert rel(p.adjoint() * fHat, po.adjoint(fHat)) < TOL[T]
    assert rel(p * f, po.forward(f)) < TOL[T]


@pytest.mark.parametrize("N,T", [((9, 8), np.float32), ((9, 8), np.float64), ((12, 10, 8), np.float32),
                                 ((50,), np.float64)])
@pytest.mark.parametrize("cplx", [False, True])
def test_convolve_operators_real_and_complex(nb, N, T, cplx):
    """test/convolve.jl:57-146: convolve!/convolve_transpose! in isolation, real and complex data"""
    D = len(N)
    J = 101
    k = O.random_nodes(J, D, T, seed=7)
    p = nb.plan_nfft(k.T, N, m=5, σ=2.0, precompute=nb.LINEAR)
    po = O.OraclePlan(k, N, m=5, sigma=2.0, precompute=O.LINEAR, blockSize=p.params.blockSize)
    rng = np.random.default_rng(3)
    cT = p.cT if cplx else T
    g = np.asfortranarray((rng.random(p.Ñ) + (1j * rng.random(p.Ñ) if cplx else 0)).astype(cT))
    fh = np.zeros(J, dtype=cT)
    assert nb.convolve_(p, g, fh) is fh
    ref = po.convolve(g)
    assert rel(fh, ref) < TOL[T] * 10
    v = (rng.random(J) + (1j * rng.random(J) if cplx else 0)).astype(cT)
    gg = np.zeros(p.Ñ, dtype=cT, order="F")
    assert nb.convolve_transpose_(p, v, gg) is gg
    assert rel(gg, po.convolve_transpose(v)) < TOL[T] * 10


def test_convolve_throws(nb):
    """test/convolve.jl:43-54"""
    N = (9, 8)
    J = 101
    k = O.random_nodes(J, 2, np.float32, seed=7)
    p = nb.plan_nfft(k.T, N, m=5, σ=2.0, precompute=nb.LINEAR)
    Nt = p.Ñ
    with pytest.raises(nb.ArgumentError):
        nb.convolve_(p, np.ones(Nt, np.complex64, order="F"), np.zeros(J, np.float32))
    with pytest.raises(nb.ArgumentError):
        nb.convolve_transpose_(p, np.ones(J, np.complex64), np.zeros(Nt, np.float32, order="F"))
    with pytest.raises(nb.DimensionMismatch):
        nb.convolve_(p, np.ones(Nt, np.complex64, order="F"), np.zeros(2, np.complex64))
    with pytest.raises(nb.DimensionMismatch):
        nb.convolve_(p, np.ones(N, np.complex64, order="F"), np.zeros(J, np.complex64))
    with pytest.raises(nb.DimensionMismatch):
        nb.convolve_transpose_(p, np.ones(2, np.complex64), np.zeros(Nt, np.complex64, order="F"))
    with pytest.raises(nb.DimensionMismatch):
        nb.convolve_transpose_(p, np.ones(J, np.complex64), np.zeros(N, np.complex64, order="F"))
    with pytest.raises(nb.DimensionMismatch):
        nb.mul_(np.zeros(J + 1, np.complex64), p, np.zeros(N, np.complex64, order="F"))


def test_constructor_errors_and_nodes(nb):
    """test/constructors.jl:15,35-38,42-70"""
    with pytest.raises(nb.ArgumentError):
        nb.plan_nfft(np.zeros((1, 4)), (2, 2))
    with pytest.raises(nb.ArgumentError):
        nb.plan_nfft(np.array([[-0.6, 0.9], [0.5, -0.5]]), (4, 4))
    with pytest.raises(nb.ArgumentError):
        nb.plan_nfft(np.array([[-0.3, 0.3], [0.3, np.nan]]), (4, 4))
    Nx = 32
    rng = np.random.default_rng(0)
    trj1 = rng.random((2, 1000)) - 0.5
    trj2 = rng.random((2, 1000)) - 0.5
    p1 = nb.plan_nfft(trj1, (Nx, Nx))
    p2 = nb.plan_nfft(trj2, (Nx, Nx))
    assert p1.params.m == 5 and p1.params.σ == 2.0           # no kwargs => reltol=1e-9 (misc.jl:75-78)
    nb.nodes_(p2, trj1)
    assert np.array_equal(p1.permutation()[0], p2.permutation()[0])
    f = O.random_complex((Nx, Nx), np.float64, 1)
    assert np.array_equal(p1 * f, p2 * f)
    assert p1.size_in() == (Nx, Nx) and p1.size_out() == (1000,)
    assert p1.adjoint().size_in() == (1000,)
    # nodes! with a different node count
    nb.nodes_(p2, trj2[:, :500])
    assert p2.size_out() == (500,)
    out = p2 * f
    assert rel(out, O.ndft(trj2[:, :500].T, f)) < 1e-7


def test_issue_106_lut_boundary(nb):
    """test/issues.jl:1-17: LINEAR LUT must not go out of bounds for this Float32 node"""
    T = np.float32
    trj = np.full((1, 2), 0.008333333, dtype=T)
    p = nb.plan_nfft(trj, (240,), precompute=nb.LINEAR)
    lam = p.adjoint() * np.ones(2, dtype=np.complex64)
    assert np.all(np.isfinite(lam))
    po = O.OraclePlan(trj.T, (240,), precompute=O.LINEAR)
    assert rel(lam, po.adjoint(np.ones(2, dtype=np.complex64))) < 1e-5


def test_batched_equals_loop(nb):
    """ntransforms = B: same as looping mul! over the batch with one plan (accuracy.jl:123-163)"""
    N = (16, 12, 10)
    M, B = 3000, 3
    T = np.float32
    k = O.random_nodes(M, 3, T, seed=9)
    pb = nb.plan_nfft(k.T, N, m=3, σ=2.0, ntransforms=B)
    p1 = nb.plan_nfft(k.T, N, m=3, σ=2.0)
    f = O.random_complex(N + (B,), T, 4)
    fh = O.random_complex((M, B), T, 5)
    out = pb * f
    adj = pb.adjoint() * fh
    assert out.shape == (M, B) and adj.shape == N + (B,)
    for b in range(B):
        assert rel(out[:, b], p1 * np.asfortranarray(f[..., b])) < 1e-6
        assert rel(adj[..., b], p1.adjoint() * np.ascontiguousarray(fh[:, b])) < 1e-6


def test_sdc_known_answer(nb):
    """test/samplingDensity.jl:10-27: nodes on the 9x8 grid => weights == 1/72 (Pipe-Menon via the
    real-valued convolve_transpose!/convolve! pair, NFFTTools/src/samplingDensity.jl:93-118)"""
    N = (9, 8)
    T = np.float32
    x = (np.arange(N[0]) / N[0] - 0.5).astype(T)
    y = (np.arange(N[1]) / N[1] - 0.5).astype(T)
    nodes = np.array([[a, b] for b in y for a in x], dtype=T).T
    for pre in (nb.LINEAR, nb.FULL, nb.TENSOR, nb.POLYNOMIAL):
        p = nb.plan_nfft(nodes, N, m=5, σ=2.0, precompute=pre)
        J = p.J
        w = np.ones(J, dtype=T)
        g = np.zeros(p.Ñ, dtype=T, order="F")
        tmp = np.zeros(J, dtype=T)
        scaling = None
        for i in range(10):
            nb.convolve_transpose_(p, w, g)
            if i == 0:
                scaling = g.max()
            g /= scaling
            nb.convolve_(p, g, tmp)
            tmp /= scaling
            assert np.all(tmp > 0)
            w /= tmp
        u = np.ones(N, dtype=np.complex64, order="F")
        wf = (p * u) * w
        v = p.adjoint() * wf
        c = np.real(v.sum()) / np.sum(np.abs(v) ** 2)
        w = w * T(c)
        assert np.all(w > 0)
        assert np.allclose(w, 1 / 72, rtol=1e-4)


def test_device_buffers_and_stage_ops(nb):
    import torch
    N = (24, 20, 18)
    M = 5000
    T = np.float32
    k = O.random_nodes(M, 3, T, seed=2)
    p = nb.plan_nfft(torch.from_numpy(k.T.copy()).cuda(), N, m=4, σ=2.0)
    po = O.OraclePlan(k, N, m=4, sigma=2.0, blockSize=p.params.blockSize)
    f = O.random_complex(N, T, 3)
    fd = p.empty_image()
    fd.copy_(torch.from_numpy(np.ascontiguousarray(f)).cuda())
    out = p.empty_out()
    nb.mul_(out, p, fd)
    assert rel(out.cpu().numpy(), po.forward(f)) < 1e-5
    # stage-by-stage against the oracle: D, F, B
    g = p.empty_grid()
    nb.deconvolve_(p, fd, g)
    g_o = po.deconvolve(f)
    assert rel(g.cpu().numpy(), g_o) < 1e-6
    tv = p.tmpVec
    tv.copy_(g)
    p.fft_(-1)
    from scipy import fft as sfft
    assert rel(tv.cpu().numpy(), sfft.fftn(g_o)) < 1e-5
    back = p.empty_image()
    nb.deconvolve_transpose_(p, g, back)
    assert rel(back.cpu().numpy(), po.deconvolve_transpose(g_o)) < 1e-6
    ts = nb.TimingStats()
    nb.mul_(out, p, fd, timing=ts)
    assert ts.conv > 0 and ts.fft > 0 and ts.deconv > 0
    nb.mul_(fd, p.adjoint(), out, timing=ts)
    assert ts.conv_adjoint > 0 and ts.fft_adjoint > 0 and ts.deconv_adjoint > 0
    assert p.launch_count() > 0


@pytest.mark.parametrize("N,m,T", [((96, 80), 4, np.float64), ((96, 80), 4, np.float32), ((128, 128), 3, np.float32),
                                   ((100, 72), 5, np.float64), ((256, 256), 4, np.float64), ((64, 96), 2, np.float32)])
@pytest.mark.parametrize("pre", [O.POLYNOMIAL, O.LINEAR])
def test_tiled_2d_kernels(nb, N, m, T, pre):
    """grids large enough that the tiled 2-D kernels (twod.cu) run instead of the generic fallback"""
    M = 30011
    k = O.random_nodes(M, 2, T, seed=12)
    k[:3] = 0.5; k[3:6] = -0.5; k[6] = 0.0
    p = nb.plan_nfft(k.T, N, m=m, σ=2.0, precompute=nb.PrecomputeFlags(pre))
    po = O.OraclePlan(k, N, m=m, sigma=2.0, precompute=pre, blockSize=p.params.blockSize)
    fHat = O.random_complex(M, T, 2)
    f = O.random_complex(N, T, 3)
    assert rel(p.adjoint() * fHat, po.adjoint(fHat)) < TOL[T]
    assert rel(p * f, po.forward(f)) < TOL[T]
    p.set_kernel_mode(1)                                     # and the generic kernels agree with the tiled ones
    assert rel(p.adjoint() * fHat, po.adjoint(fHat)) < TOL[T]


def test_radial_batched_2d(nb):
    """C3-like: radial trajectory (all spokes cross the centre tile), ntransforms = 4, density-weighted adjoint"""
    T = np.float32
    N = (128, 128)
    k = O.radial_nodes(96, 256, T)
    M, B = k.shape[0], 4
    w = np.maximum(np.abs(np.tile((np.arange(256) - 128) / 256, 96)), 1 / 512).astype(T)
    p = nb.plan_nfft(k.T, N, m=4, σ=2.0, ntransforms=B)
    po = O.OraclePlan(k, N, m=4, sigma=2.0, blockSize=p.params.blockSize)
    fh = O.random_complex((M, B), T, 5) * w[:, None]
    f = O.random_complex(N + (B,), T, 6)
    adj = p.adjoint() * fh
    out = p * f
    for b in range(B):
        assert rel(adj[..., b], po.adjoint(np.ascontiguousarray(fh[:, b]))) < 1e-5
        assert rel(out[:, b], po.forward(np.asfortranarray(f[..., b]))) < 1e-5


@pytest.mark.parametrize("N,T", [((1 << 16,), np.float64), ((5000,), np.float32), ((3000,), np.float64)])
def test_tiled_1d_kernels(nb, N, T):
    """dense 1-D node sets (several nodes per cell) through the output-stationary spreader (oned.cu)"""
    M = 8 * N[0] + 13
    k = O.random_nodes(M, 1, T, seed=8)
    k[:4, 0] = [0.5, -0.5, 0.0, -1e-9]
    p = nb.plan_nfft(k.T, N, m=4, σ=2.0)
    po = O.OraclePlan(k, N, m=4, sigma=2.0, blockSize=p.params.blockSize)
    fHat = O.random_complex(M, T, 2)
    f = O.random_complex(N, T, 3)
    assert rel(p.adjoint() * fHat, po.adjoint(fHat)) < TOL[T]
    assert rel(p * f, po.forward(f)) < TOL[T]


import glob as _glob
import os as _os


@pytest.mark.parametrize("path", sorted(_glob.glob(_os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "golden", "*.npz"))))
def test_against_committed_golden_vectors(nb, path):
    """tests/golden/*.npz (oracle-generated, script committed): permutation bit-exact, outputs within tolerance"""
    z = np.load(path)
    k = z["k"]
    T = k.dtype.type
    if "toeplitz_lambda" in z:         # NFFTTools fixtures: Toeplitz kernel / apply and sdc
        shape = tuple(int(n) for n in z["shape"])
        lam = nb.calculateToeplitzKernel(shape, k.T, m=4, σ=2.0)
        assert rel(lam, z["toeplitz_lambda"]) < (1e-11 if T == np.float64 else 2e-5)
        y = np.asfortranarray(z["y"]).copy(order="F")
        nb.convolveToeplitzKernel_(y, np.asfortranarray(z["toeplitz_lambda"]))
        assert rel(y, z["toeplitz_out"]) < TOL[T]
        p = nb.plan_nfft(k.T, shape, m=4, σ=2.0, blockSize=tuple(int(b) for b in z["blockSize"]))
        assert np.abs(nb.sdc(p, iters=10) / z["sdc"] - 1).max() < (1e-9 if T == np.float64 else 1e-3)
        return
    N = tuple(int(n) for n in z["N"])
    window = str(z["window"]) if "window" in z else "kaiser_bessel"
    p = nb.plan_nfft(k.T, N, m=int(z["m"]), σ=2.0, precompute=nb.PrecomputeFlags(int(z["pre"])), window=window,
                     blockSize=tuple(int(b) for b in z["blockSize"]))
    assert np.array_equal(p.permutation()[0], z["perm"])
    assert rel(p * z["f"], z["forward"]) < TOL[T]
    assert rel(p.adjoint() * z["fHat"], z["adjoint"]) < TOL[T]


def test_full_size_properties_c2(nb):
    """BASELINE configs[1] at full size (3-D 128^3, M = 2^21, m = 3, Float32), through size-independent properties:
    adjointness <A f, y> == <f, A^H y>, linearity, determinism of the adjoint (no atomics => bit-identical reruns),
    and agreement with the NDFT on a subsample of nodes / image points."""
    import torch
    T = np.float32
    N, M = (128, 128, 128), 2 ** 21
    k = O.random_nodes(M, 3, T, seed=1)
    p = nb.plan_nfft(torch.from_numpy(np.ascontiguousarray(k.T)).cuda(), N, m=3, σ=2.0)
    perm, ts = p.permutation()
    assert np.array_equal(np.sort(perm), np.arange(M)) and ts[-1] == M
    key = O.tile_keys(O.shift_nodes(k), O.init_params(N, T, 3, 2.0, blockSize=p.params.blockSize))[0]
    assert np.all(np.diff(key[perm]) >= 0)                                  # sorted by tile ...
    same = key[perm][1:] == key[perm][:-1]
    assert np.all(np.diff(perm)[same] > 0)                                  # ... and stable inside a tile
    f = O.random_complex(N, T, 2)
    y = O.random_complex(M, T, 3)
    Af = p * f
    AHy = p.adjoint() * y
    lhs = np.vdot(y.astype(np.complex128), Af.astype(np.complex128))
    rhs = np.vdot(AHy.astype(np.complex128).ravel(), f.astype(np.complex128).ravel())
    assert abs(lhs - rhs) / abs(lhs) < 1e-5
    f2 = O.random_complex(N, T, 4)
    lin = p * np.asfortranarray(2 * f + 3j * f2)
    assert rel(lin, 2 * Af + 3j * (p * f2)) < 1e-5
    assert np.array_equal(AHy, p.adjoint() * y)                             # deterministic adjoint
    sub = np.arange(0, M, M // 64)[:64]
    assert rel(Af[sub], O.ndft(k[sub].astype(np.float64), f)) < 3e-5        # reference error level at m = 3, Float32


def test_edge_cases(nb):
    """empty / tiny / degenerate inputs the reference handles (ragged tiles, odd sizes, sigma != 2, clustered nodes)"""
    T = np.float64
    # no nodes at all
    p = nb.plan_nfft(np.zeros((2, 0)), (16, 16), m=3, σ=2.0)
    assert p.size_out() == (0,)
    out = p.adjoint() * np.zeros(0, dtype=np.complex128)
    assert out.shape == (16, 16) and not np.any(out)
    assert (p * O.random_complex((16, 16), T, 1)).shape == (0,)
    # one node, odd sizes, grid not a multiple of the tile, sigma = 1.5 and 1.25
    for N, sig, m in [((33,), 1.5, 3), ((17, 21), 1.25, 4), ((9, 11, 13), 1.5, 2), ((70, 70), 2.0, 4), ((20, 18, 40), 2.0, 3)]:
        for M in (1, 257):
            k = O.random_nodes(M, len(N), T, seed=M)
            p = nb.plan_nfft(k.T, N, m=m, σ=sig)
            po = O.OraclePlan(k, N, m=m, sigma=sig, blockSize=p.params.blockSize)
            assert p.Ñ == po.Nt and p.params.σ == po.p.sigma
            assert np.array_equal(p.permutation()[0], po.perm)
            fh = O.random_complex(M, T, 2); f = O.random_complex(N, T, 3)
            assert rel(p.adjoint() * fh, po.adjoint(fh)) < 1e-12
            assert rel(p * f, po.forward(f)) < 1e-12
    # all nodes in one spot (every work item of one tile, one warp's sub-tile) and on exact cell boundaries
    for N in [(64, 64), (32, 32, 32), (4096,)]:
        D = len(N)
        M = 9000
        k = np.full((M, D), 0.123, dtype=np.float32) + (np.arange(M)[:, None] % 7) * np.float32(1e-4)
        k[::3] = np.float32(0.25)                      # exactly on a grid cell
        p = nb.plan_nfft(k.T, N, m=4 if D < 3 else 3, σ=2.0)
        po = O.OraclePlan(k, N, m=4 if D < 3 else 3, sigma=2.0, blockSize=p.params.blockSize)
        fh = O.random_complex(M, np.float32, 2); f = O.random_complex(N, np.float32, 3)
        assert np.array_equal(p.permutation()[0], po.perm)
        assert rel(p.adjoint() * fh, po.adjoint(fh)) < 1e-5
        assert rel(p * f, po.forward(f)) < 1e-5


@pytest.mark.parametrize("N", [(2048,), (96, 80)])
def test_batched_1d_2d_and_regrow_nodes(nb, N):
    T = np.float32
    D, B = len(N), 3
    M1, M2 = 5000, 12000
    k1 = O.random_nodes(M1, D, T, seed=1); k2 = O.random_nodes(M2, D, T, seed=2)
    p = nb.plan_nfft(k1.T, N, m=4, σ=2.0, ntransforms=B)
    for k, M in ((k1, M1), (k2, M2), (k1, M1)):            # nodes! with growing and shrinking node counts
        nb.nodes_(p, k.T)
        po = O.OraclePlan(k, N, m=4, sigma=2.0, blockSize=p.params.blockSize)
        f = O.random_complex(N + (B,), T, 4); fh = O.random_complex((M, B), T, 5)
        out = p * f; adj = p.adjoint() * fh
        for b in range(B):
            assert rel(out[:, b], po.forward(np.asfortranarray(f[..., b]))) < 1e-5
            assert rel(adj[..., b], po.adjoint(np.ascontiguousarray(fh[:, b]))) < 1e-5


def test_sdc_function_and_directional_plan(nb):
    """§8f-1: nb.sdc == NFFTTools.sdc known answer (test/samplingDensity.jl:10-27);  §8f-2: dims=1:2 of an (N1,N2,B)
    array is the batched plan (test/accuracy.jl:123-163: equals looping the lower-dimensional NFFT)"""
    N, T = (9, 8), np.float32
    x = (np.arange(N[0]) / N[0] - 0.5).astype(T)
    y = (np.arange(N[1]) / N[1] - 0.5).astype(T)
    nodes = np.array([[a, b] for b in y for a in x], dtype=T).T
    p = nb.plan_nfft(nodes, N, m=5, σ=2.0)
    w = nb.sdc(p, iters=10)
    assert w.dtype == T and np.all(w > 0) and np.allclose(w, 1 / 72, rtol=1e-4)
    po = O.OraclePlan(nodes.T, N, m=5, sigma=2.0, blockSize=p.params.blockSize)
    assert np.allclose(w, O.sdc(po, iters=10), rtol=1e-4)
    # directional
    N3 = (24, 20, 5)
    k = O.random_nodes(700, 2, np.float64, seed=3)
    pd = nb.plan_nfft(k.T, N3, m=5, σ=2.0, dims=range(1, 3))
    assert pd.size_in() == N3 and pd.size_out() == (700, 5)
    f = O.random_complex(N3, np.float64, 4)
    out = pd * f
    p2 = O.OraclePlan(k, N3[:2], m=5, sigma=2.0, blockSize=pd.params.blockSize)
    for b in range(5):
        assert rel(out[:, b], p2.forward(np.asfortranarray(f[..., b]))) < 1e-12
    with pytest.raises(nb.ArgumentError):
        nb.plan_nfft(k.T, N3, dims=(1, 3))                     # not a range


def test_multi_gpu_node_and_batch_sharding():
    """node sharding (reduce-scatter + slab FFT + all-gather) and batch sharding on 2 GPUs vs the oracle; skipped on
    single-GPU boxes (run manually with `gpurun --gpus 2`)"""
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29561",
                          _os.path.join(root, "tests", "mgpu_check.py")], capture_output=True, text=True, timeout=600)
    assert "MGPU_CHECK PASS" in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]


def _approx(a, b, rtol):
    """Julia's isapprox on arrays: norm(a-b) <= rtol * max(norm(a), norm(b))"""
    a = np.asarray(a).ravel().astype(np.complex128)
    b = np.asarray(b).ravel().astype(np.complex128)
    return np.linalg.norm(a - b) <= rtol * max(np.linalg.norm(a), np.linalg.norm(b))


@pytest.mark.parametrize("shape,T,M", [((32, 32), np.float64, 1000), ((33, 33), np.float64, 1000),
                                       ((32, 33), np.float32, 1000), ((32, 32, 32), np.float32, 1000),
                                       ((100,), np.float64, 500)])
def test_toeplitz_kernel(nb, shape, T, M):
    """test/testToeplitz.jl:10-44,67-87: calculateToeplitzKernel / calculateToeplitzKernel! vs the explicit kernel
    (rtol 1e-6 F64 / 1e-5 F32) and vs the oracle's restatement of NFFTTools/src/Toeplitz.jl:86-93"""
    D = len(shape)
    k = O.random_nodes(M, D, T, seed=21)
    cT = np.complex64 if T == np.float32 else np.complex128
    Ka = nb.calculateToeplitzKernel(shape, k.T, m=4, σ=2.0)
    assert Ka.dtype == cT and Ka.shape == tuple(2 * s for s in shape)
    rt = 1e-6 if T == np.float64 else 1e-5
    assert _approx(Ka, O.calculate_toeplitz_kernel_explicit(shape, k), rt)
    assert rel(Ka, O.calculate_toeplitz_kernel(shape, k, m=4, sigma=2.0)) < (1e-11 if T == np.float64 else 2e-5)
    # the less-allocating constructor: an existing plan with other nodes, re-noded in place
    p = nb.plan_nfft(O.random_nodes(77, D, T, seed=22).T, tuple(2 * s for s in shape), m=4, σ=2.0)
    Kb = np.empty(tuple(2 * s for s in shape), dtype=cT, order="F")
    assert nb.calculateToeplitzKernel_(Kb, p, k.T) is Kb
    assert rel(Kb, Ka) < (1e-13 if T == np.float64 else 1e-6)    # work-item split may differ after nodes!
    with pytest.raises(nb.DimensionMismatch):
        nb.calculateToeplitzKernel_(np.empty(shape, dtype=cT, order="F"), p, k.T)


def test_toeplitz_convolve(nb):
    """test/testToeplitz.jl:47-58: convolveToeplitzKernel!(x, K) == nfft_adjoint(trj, N, nfft(trj, x)), rtol 1e-5,
    host and device buffers, one-shot and pre-planned (batched) operator"""
    import torch
    Nx, T = 32, np.float32
    k = O.random_nodes(10000, 2, T, seed=23)
    Ka = nb.calculateToeplitzKernel((Nx, Nx), k.T, m=4, σ=2.0)
    x = O.random_complex((Nx, Nx), T, 24)
    xN = nb.nfft_adjoint(k.T, (Nx, Nx), nb.nfft(k.T, x))
    y = x.copy(order="F")
    assert nb.convolveToeplitzKernel_(y, Ka) is y
    assert _approx(y, xN, 1e-5)
    assert rel(y, O.convolve_toeplitz_kernel(x, Ka)) < 1e-5
    # device buffers + operator re-use, batch of 3 images
    kd = torch.from_numpy(np.ascontiguousarray(k)).cuda().T
    Kd = nb.calculateToeplitzKernel((Nx, Nx), kd, m=4, σ=2.0)
    assert Kd.is_cuda and rel(Kd.cpu().numpy(), Ka) < 1e-6
    op = nb.ToeplitzOperator(Kd, ntransforms=3)
    xb = np.stack([O.random_complex((Nx, Nx), T, 30 + i) for i in range(3)], axis=-1)
    yb = torch.empty((3, Nx, Nx), dtype=torch.complex64, device="cuda").permute(2, 1, 0)
    yb.copy_(torch.from_numpy(xb))
    for _ in range(2):                                    # applying twice == applying the Gram operator twice
        op.apply_(yb)
    ref = np.stack([O.convolve_toeplitz_kernel(O.convolve_toeplitz_kernel(xb[..., i], Ka), Ka) for i in range(3)], axis=-1)
    assert rel(yb.cpu().numpy(), ref) < 2e-5
    with pytest.raises(nb.DimensionMismatch):
        op.apply_(y)
    # Float64, 3-D, rectangular
    k3 = O.random_nodes(3000, 3, np.float64, seed=25)
    K3 = nb.calculateToeplitzKernel((12, 10, 9), k3.T, m=6, σ=2.0)
    x3 = O.random_complex((12, 10, 9), np.float64, 26)
    p3 = nb.plan_nfft(k3.T, (12, 10, 9), m=6, σ=2.0)
    y3 = x3.copy(order="F")
    nb.convolveToeplitzKernel_(y3, K3)
    assert rel(y3, p3.adjoint() * (p3 * x3)) < 1e-9
    assert rel(y3, O.convolve_toeplitz_kernel(x3, K3)) < 1e-12


def test_async_host_mode_pipelines_and_matches(nb):
    """NFFTB200_HOST_ASYNC: queued forward/adjoint calls on page-locked host buffers give bit-identical results to the
    synchronous host calls, also when the same plan is re-used back to back with different inputs"""
    import torch
    N, T, M = (24, 20, 18), np.float32, 9000
    k = O.random_nodes(M, 3, T, seed=41)
    p = nb.plan_nfft(k.T, N, m=3, σ=2.0)
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
    fs = [np.asfortranarray(pin(np.ascontiguousarray(O.random_complex(N, T, 50 + i).T)).T) for i in range(3)]
    fhs = [pin(O.random_complex(M, T, 60 + i)) for i in range(3)]
    ref_f = [p * f for f in fs]
    ref_a = [p.adjoint() * fh for fh in fhs]
    out_f = [pin(np.zeros(M, dtype=np.complex64)) for _ in range(3)]
    out_a = [np.asfortranarray(pin(np.zeros(N[::-1], dtype=np.complex64)).T) for _ in range(3)]
    for rep in range(2):
        for i in range(3):
            nb.mul_(out_f[i], p, fs[i], async_host=True)
            nb.mul_(out_a[i], p.adjoint(), fhs[i], async_host=True)
        p.sync()
        for i in range(3):
            assert np.array_equal(out_f[i], ref_f[i]) and np.array_equal(out_a[i], ref_a[i])
            out_f[i][...] = 0; out_a[i][...] = 0
    with pytest.raises(nb.ArgumentError):                     # conversions are not possible behind an asynchronous call
        nb.mul_(np.zeros(M, dtype=np.complex128), p, fs[0], async_host=True)
    # nodes! with more nodes while nothing is in flight: staging buffers grow
    k2 = O.random_nodes(2 * M, 3, T, seed=42)
    nb.nodes_(p, k2.T)
    fh2 = pin(O.random_complex(2 * M, T, 70))
    o2 = np.asfortranarray(pin(np.zeros(N[::-1], dtype=np.complex64)).T)
    nb.mul_(o2, p.adjoint(), fh2, async_host=True)
    p.sync()
    assert np.array_equal(o2, p.adjoint() * fh2)


def test_copy_gives_independent_plan(nb):
    """test/constructors.jl:17-28: copy(p) has the same parameters and results, its own grid/scratch; copy(adjoint(p))
    does not error"""
    import copy as _copy
    k = O.random_nodes(500, 2, np.float64, seed=81)
    p = nb.plan_nfft(k.T, (12, 14), m=4, σ=2.0, window="cosh_type", precompute=nb.LINEAR)
    q = _copy.copy(p)
    assert q.params == p.params and q.N == p.N and q.Ñ == p.Ñ and q.J == p.J
    assert q.tmpVec.data_ptr() != p.tmpVec.data_ptr()
    f = O.random_complex((12, 14), np.float64, 82)
    assert np.array_equal(p * f, q * f)
    p.destroy()
    fh = O.random_complex(500, np.float64, 83)
    qa = q.adjoint().copy()
    assert rel(qa * fh, q.adjoint() * fh) < 1e-14           # small 2-D plans spread with global REDs: order may differ
    assert "Adjoint of B200NFFTPlan with 500 sampling points" in repr(qa)


@pytest.mark.parametrize("N", [(8, 12), (10, 8, 14)])
def test_directional_plans_any_dims(nb, N):
    """test/accuracy.jl:83-163: an NFFT along dims=d (or dims=d:d+1) equals the lower-dimensional NFFT of every slice
    along those dims, for every position of the transformed dims (leading, middle, trailing), numpy and CUDA buffers"""
    import itertools
    import torch
    D, T = len(N), np.float64
    J = 60
    cases = [(d,) for d in range(1, D + 1)] + ([(d, d + 1) for d in range(1, D)] if D == 3 else [])
    for dims in cases:
        nd = len(dims)
        k = O.random_nodes(J, nd, T, seed=90 + dims[0])
        f = O.random_complex(N, T, 91)
        p_dir = nb.plan_nfft(k.T if nd > 1 else k[:, 0], N, dims=dims if nd > 1 else dims[0], m=5, σ=2.0)
        pre, post = N[:dims[0] - 1], N[dims[-1]:]
        assert p_dir.size_in() == tuple(N) and p_dir.size_out() == pre + (J,) + post
        fHat_dir = p_dir * f
        g_dir = p_dir.adjoint() * fHat_dir
        p = nb.plan_nfft(k.T, N[dims[0] - 1:dims[-1]], m=5, σ=2.0)
        fHat = np.zeros_like(fHat_dir)
        g = np.zeros_like(g_dir)
        for Ipost in itertools.product(*[range(n) for n in post]):
            for Ipre in itertools.product(*[range(n) for n in pre]):
                idxf = Ipre + (slice(None),) * nd + Ipost
                idxh = Ipre + (slice(None),) + Ipost
                fHat[idxh] = p * np.asfortranarray(f[idxf])
                g[idxf] = p.adjoint() * np.ascontiguousarray(fHat_dir[idxh])
        assert rel(fHat_dir, fHat) < 1e-13 and rel(g_dir, g) < 1e-13
        # device buffers
        fd = torch.from_numpy(np.ascontiguousarray(f)).cuda()
        hd = p_dir * fd
        assert hd.is_cuda and rel(hd.cpu().numpy(), fHat) < 1e-13
        assert rel((p_dir.adjoint() * hd).cpu().numpy(), g) < 1e-13
    with pytest.raises(nb.ArgumentError):
        nb.plan_nfft(O.random_nodes(J, 1, T, seed=1)[:, 0], N, dims=1, ntransforms=2, m=5, σ=2.0)


@pytest.mark.parametrize("D,T", [(1, np.float64), (2, np.float32), (2, np.float64), (3, np.float32)])
def test_native_sdc_matches_operator_loop(nb, D, T):
    """nfftb200_sdc (the whole Pipe-Menon iteration + scaling on the device) against the same algorithm written with
    the public operators as NFFTTools does (NFFTTools/src/samplingDensity.jl:93-155) and against the oracle, on a
    radial-like non-uniform node set; device output variant"""
    N = {1: (64,), 2: (24, 20), 3: (12, 10, 8)}[D]
    rng = np.random.default_rng(5)
    M = 4 * int(np.prod(N))
    r = rng.random(M) ** 1.5 * 0.5                        # denser near the centre
    d = rng.standard_normal((M, D)); d /= np.linalg.norm(d, axis=1, keepdims=True)
    k = (r[:, None] * d).astype(T)
    p = nb.plan_nfft(k.T, N, m=4, σ=2.0)
    w = nb.sdc(p, iters=12)
    assert w.dtype == T and w.shape == (M,) and np.all(w > 0)
    w_loop = nb.sdc_host_loop(p, iters=12)
    tol = 2e-4 if T == np.float32 else 1e-10
    assert np.abs(w / w_loop - 1).max() < tol
    po = O.OraclePlan(k, N, m=4, sigma=2.0, blockSize=p.params.blockSize)
    assert np.abs(w / O.sdc(po, iters=12) - 1).max() < (1e-3 if T == np.float32 else 1e-9)
    wd = nb.sdc(p, iters=12, device=True)
    assert wd.is_cuda and np.abs(wd.cpu().numpy() / w - 1).max() < tol
    pb = nb.plan_nfft(k.T, N, m=4, σ=2.0, ntransforms=2)
    with pytest.raises(NotImplementedError):
        nb.sdc(pb)


def test_plan_follows_torch_stream_and_copy_keeps_dims(nb):
    """round-1 ADVICE: (1) a stream="current" plan re-binds to torch's current stream at every call, so work queued on a
    side stream (inputs produced there, outputs consumed there) is ordered correctly; (2) copy() of a directional plan
    keeps dims and the user-facing sizes; (3) a batch that would overflow the grid dimension limit fails at plan time."""
    import copy as _copy
    import torch
    T = np.float32
    N, M = (24, 20, 16), 9000
    k = O.random_nodes(M, 3, T, seed=5)
    kd = torch.from_numpy(np.ascontiguousarray(k.T)).cuda()
    p = nb.plan_nfft(kd, N, m=3, σ=2.0)
    f = torch.from_numpy(O.random_complex(N, T, 6)).cuda()
    want = p * f
    torch.cuda.synchronize()
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        g = (f * 2.0)                                  # produced on the side stream ...
        out = p.empty_out()
        nb.mul_(out, p, g.permute(2, 1, 0).contiguous().permute(2, 1, 0))        # ... consumed by the plan on the same stream
        got = out / 2.0
    side.synchronize()
    assert rel(got.cpu().numpy(), want.cpu().numpy() if hasattr(want, "cpu") else want) < 1e-6
    # directional copy
    N3 = (10, 8, 14)
    k2 = O.random_nodes(300, 2, np.float64, seed=7)
    pd = nb.plan_nfft(k2.T, N3, m=4, σ=2.0, dims=range(2, 4))
    qd = _copy.copy(pd)
    assert qd.size_in() == pd.size_in() and qd.size_out() == pd.size_out()
    x = O.random_complex(N3, np.float64, 8)
    assert np.array_equal(pd * x, qd * x)
    # grid-dimension limit is a plan-time error, not a launch failure
    with pytest.raises(Exception):
        nb.plan_nfft(kd, N, m=3, σ=2.0, ntransforms=70000)
