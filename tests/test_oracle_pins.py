"""CPU tests that pin the oracle (oracle/) at everything the reference's own tests pin for this path
(SURVEY.md 8c).  No GPU, no /root/reference access."""
import glob
import os

import numpy as np
import pytest

from oracle import nfft_oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))


def rel(a, b):
    a = np.asarray(a).ravel()
    b = np.asarray(b).ravel()
    return float(np.linalg.norm(a - b) / np.linalg.norm(b))


def test_accuracy_params_rule():
    """AbstractNFFTs/src/misc.jl:44-81 incl. 'm alone is ignored' """
    assert O.accuracy_params() == (5, 2.0, 1e-9)
    assert O.accuracy_params(m=3) == (5, 2.0, 1e-9)
    assert O.accuracy_params(m=3, sigma=2.0)[:2] == (3, 2.0)
    assert O.accuracy_params(reltol=1e-7)[0] == 4
    assert O.accuracy_params(reltol=1e-5)[0] == 3


def test_init_params():
    """src/precomputation.jl:14-29, :59-77"""
    p = O.init_params((128, 128, 128), np.float32, 3, 2.0)
    assert p.Nt == (256, 256, 256) and p.LUTSize == 2 ** 9 * 3 and p.blockSize == (16, 16, 16)
    p = O.init_params((256, 256), np.float64, 4, 2.0)
    assert p.Nt == (512, 512) and p.LUTSize == 2 ** 14 * 4 and p.blockSize == (64, 64)
    p = O.init_params((2 ** 22,), np.float64, 4, 2.0)
    assert p.Nt == (2 ** 23,) and p.blockSize == (1024,)
    p = O.init_params((33, 35), np.float64, 5, 2.0)
    assert p.Nt == (66, 70) and p.LUTSize == 2 ** 17 * 5
    p = O.init_params((9,), np.float64, 5, 1.5)          # ceil(13.5)=14 -> 14 ; sigma re-derived
    assert p.Nt == (14,) and abs(p.sigma - 14 / 9) < 1e-15


def test_shift_and_check_nodes():
    """src/utils.jl:32-55"""
    for T in (np.float32, np.float64):
        k = np.array([[-0.5], [0.5], [0.0], [-1e-30], [0.25], [-0.25]], dtype=T)
        s = O.shift_nodes(k)
        assert s[0, 0] == T(0.5) and s[1, 0] == T(0.5) and s[2, 0] == 0
        assert s[3, 0] == T(1) - np.finfo(T).eps          # tiny negative -> 1 -> 1-eps
        assert s[5, 0] == T(0.75)
        assert np.all((s >= 0) & (s < 1))
    with pytest.raises(ValueError):
        O.check_nodes(np.array([[-0.6, 0.9]]))
    with pytest.raises(ValueError):
        O.check_nodes(np.array([[0.3, np.nan]]))
    O.check_nodes(np.array([[0.5, -0.5]]))                # closed interval is legal


@pytest.mark.parametrize("N", [(255,), (31, 33), (11, 12, 14), (6, 5, 6, 6)])
@pytest.mark.parametrize("pre,blocking", [(O.LINEAR, True), (O.LINEAR, False), (O.FULL, False), (O.TENSOR, True),
                                          (O.POLYNOMIAL, True), (O.POLYNOMIAL, False)])
def test_vs_ndft_reference_tolerance(N, pre, blocking):
    """test/accuracy.jl:41-73: Kaiser-Bessel, m=5, sigma=2 => rel-L2 < 1e-7 for every mode, D=1..4"""
    D, M = len(N), int(np.prod(N))
    k = O.random_nodes(M, D, np.float64, seed=1)
    p = O.OraclePlan(k, N, m=5, sigma=2.0, precompute=pre, blocking=blocking)
    fHat = O.random_complex(M, np.float64, 2)
    f = O.ndft_adjoint(k, N, fHat)
    assert rel(p.adjoint(fHat), f) < 1e-7
    assert rel(p.forward(f), O.ndft(k, f)) < 1e-7


def test_published_error_band():
    """benchmark/paper/img/accuracy_m_D2.tex:40-47,85-92 (2-D, sigma=2, rel l-inf): the restatement must
    land in the same decade band as the published NFFT.jl errors for m = 3..6."""
    pub_fwd = {3: 1.65e-6, 4: 2.71e-8, 5: 1.44e-10, 6: 2.84e-12}
    pub_adj = {3: 1.20e-5, 4: 1.10e-7, 5: 2.11e-9, 6: 2.41e-11}
    N = (48, 48)
    M = 2304
    k = O.random_nodes(M, 2, np.float64, seed=3)
    fHat = O.random_complex(M, np.float64, 4)
    f = O.random_complex(N, np.float64, 5)
    t_fwd = O.ndft(k, f)
    t_adj = O.ndft_adjoint(k, N, fHat)
    for m in (3, 4, 5, 6):
        p = O.OraclePlan(k, N, m=m, sigma=2.0)
        e_f = np.abs(p.forward(f) - t_fwd).max() / np.abs(t_fwd).max()
        e_a = np.abs(p.adjoint(fHat) - t_adj).max() / np.abs(t_adj).max()
        assert pub_fwd[m] / 30 < e_f < pub_fwd[m] * 30, (m, e_f)
        assert pub_adj[m] / 100 < e_a < pub_adj[m] * 30, (m, e_a)


def test_modes_agree_like_published():
    """accuracy_m_pre_D2.tex:40-95: FULL/LINEAR/TENSOR/POLYNOMIAL give the same error to ~2 digits;
    SURVEY 7 'hard parts': POLYNOMIAL differs from exact by ~1e-8*peak at m=4, LINEAR by ~5e-10*peak"""
    p = O.init_params((64,), np.float64, 4, 2.0)
    P = O.precompute_poly_interp(p)
    t = np.linspace(-0.5, 0.5, 501)
    V = np.vander(t, 9, increasing=True)
    ex = np.stack([O.window_kaiser_bessel((-(l - 0.5) + 4) + t, 4, p.b) for l in range(1, 9)], axis=1)
    e = np.abs(V @ P - ex).max() / ex.max()
    assert 3e-9 < e < 3e-8
    lut = O.precompute_lin_interp(p)
    assert lut.shape == (p.LUTSize + 2,)
    x = np.linspace(0, 4, 100001)[:-1]
    idx = x * (p.LUTSize / 4)
    i0 = np.floor(idx).astype(int)
    lerp = lut[i0] + (idx - i0) * (lut[i0 + 1] - lut[i0])
    e = np.abs(lerp - O.window_kaiser_bessel(x, 4, p.b)).max() / ex.max()
    assert e < 2e-9


WINDOW_EPS = {"kaiser_bessel": 1e-7, "cosh_type": 1e-7, "gauss": 1e-3, "kaiser_bessel_rev": 1e-6, "spline": 1e-4}


@pytest.mark.parametrize("N", [(255,), (31, 33), (11, 12, 14)])
@pytest.mark.parametrize("window", ["cosh_type", "gauss", "kaiser_bessel_rev", "spline"])
@pytest.mark.parametrize("pre,blocking", [(O.LINEAR, True), (O.FULL, False), (O.POLYNOMIAL, True)])
def test_windows_vs_ndft_reference_tolerance(N, window, pre, blocking):
    """test/accuracy.jl:41-73: every window of getWindow at m=5, sigma=2 against the NDFT with the reference's
    per-window tolerance table (eps = [1e-7, 1e-7, 1e-3, 1e-6, 1e-4], test/accuracy.jl:46-47); the Chebyshev-30
    deconvolution LUT of the reference (src/precomputation.jl:351) is used here"""
    D, M = len(N), int(np.prod(N))
    k = O.random_nodes(M, D, np.float64, seed=1)
    p = O.OraclePlan(k, N, m=5, sigma=2.0, precompute=pre, blocking=blocking, window=window, cheb30=True)
    fHat = O.random_complex(M, np.float64, 2)
    f = O.ndft_adjoint(k, N, fHat)
    assert rel(p.adjoint(fHat), f) < WINDOW_EPS[window]
    assert rel(p.forward(f), O.ndft(k, f)) < WINDOW_EPS[window]


def _approx(a, b, rtol):
    """Julia's isapprox on arrays: norm(a-b) <= rtol * max(norm(a), norm(b))"""
    a = np.asarray(a).ravel().astype(np.complex128)
    b = np.asarray(b).ravel().astype(np.complex128)
    return np.linalg.norm(a - b) <= rtol * max(np.linalg.norm(a), np.linalg.norm(b))


def test_toeplitz_kernel_matches_explicit():
    """test/testToeplitz.jl:10-33,36-44,67-87: NFFT-based Toeplitz kernel vs the explicit one, rtol 1e-6 (F64) /
    1e-5 (F32), square, rectangular and 3-D"""
    rng = np.random.default_rng(5)
    for Nx in (32, 33):
        k = rng.random((1000, 2)) - 0.5
        Ka = O.calculate_toeplitz_kernel((Nx, Nx), k, m=4, sigma=2.0)
        assert Ka.dtype == np.complex128 and Ka.shape == (2 * Nx, 2 * Nx)
        assert _approx(Ka, O.calculate_toeplitz_kernel_explicit((Nx, Nx), k), 1e-6)
    k = (rng.random((1000, 2)) - 0.5).astype(np.float32)
    Ka = O.calculate_toeplitz_kernel((32, 33), k, m=4, sigma=2.0)
    assert Ka.dtype == np.complex64 and Ka.shape == (64, 66)
    assert _approx(Ka, O.calculate_toeplitz_kernel_explicit((32, 33), k), 1e-5)
    k = (rng.random((1000, 3)) - 0.5).astype(np.float32)
    assert _approx(O.calculate_toeplitz_kernel((16, 16, 16), k), O.calculate_toeplitz_kernel_explicit((16, 16, 16), k), 1e-5)


def test_toeplitz_convolve_equals_gram():
    """test/testToeplitz.jl:47-58: convolveToeplitzKernel!(x, K) == nfft_adjoint(nfft(x)), rtol 1e-5 (F32)"""
    rng = np.random.default_rng(6)
    Nx = 32
    k = (rng.random((10000, 2)) - 0.5).astype(np.float32)
    Ka = O.calculate_toeplitz_kernel((Nx, Nx), k, m=4, sigma=2.0)
    x = O.random_complex((Nx, Nx), np.float32, 7)
    p = O.OraclePlan(k, (Nx, Nx))                      # nfft/nfft_adjoint defaults: reltol 1e-9 -> m=5
    xN = p.adjoint(p.forward(x))
    assert _approx(O.convolve_toeplitz_kernel(x, Ka), xN, 1e-5)


def test_unknown_window_errors():
    """src/windowFunctions.jl:16"""
    with pytest.raises(ValueError):
        O.init_params((16,), np.float64, 4, 2.0, window="hann")


def test_window_hat_cheb30_vs_exact():
    """src/precomputation.jl:347-358: the 30-point Chebyshev interpolant equals the exact 1/phi_hat to
    <= 1e-13 relative for the BASELINE configurations (so evaluating exactly is inside the 1e-12 budget)"""
    for N, T, m in [((256, 256), np.float64, 4), ((128, 128, 128), np.float32, 3), ((512, 512), np.float32, 4)]:
        p = O.init_params(N, np.float64, m, 2.0)
        a = np.concatenate(O.window_hat_inv_lut(p, cheb30=True))
        b = np.concatenate(O.window_hat_inv_lut(p, cheb30=False))
        assert np.abs(a / b - 1).max() < 1e-13
    for window in O.WINDOWS[1:]:
        p = O.init_params((255, 64), np.float64, 5, 2.0, window=window)
        a = np.concatenate(O.window_hat_inv_lut(p, cheb30=True))
        b = np.concatenate(O.window_hat_inv_lut(p, cheb30=False))
        assert np.abs(a / b - 1).max() < 1e-12


def test_sdc_known_answer():
    """test/samplingDensity.jl:10-27: nodes on the 9x8 grid, Float32, m=5 => all weights == 1/72"""
    N = (9, 8)
    T = np.float32
    x = (np.arange(N[0]) / N[0] - 0.5).astype(T)
    y = (np.arange(N[1]) / N[1] - 0.5).astype(T)
    nodes = np.array([[a, b] for b in y for a in x], dtype=T)
    for pre in (O.LINEAR, O.FULL, O.TENSOR, O.POLYNOMIAL):
        p = O.OraclePlan(nodes, N, m=5, sigma=2.0, precompute=pre)
        w = O.sdc(p, iters=10)
        assert w.dtype == T and np.all(w > 0)
        assert np.allclose(w, 1 / 72, rtol=1e-4)


def test_issue_106_lut_bounds():
    """test/issues.jl:1-17"""
    T = np.float32
    trj = np.full((2, 1), 0.008333333, dtype=T)
    for blocking in (False, True):
        p = O.OraclePlan(trj, (240,), precompute=O.LINEAR, blocking=blocking)
        lam = p.adjoint(np.ones(2, dtype=np.complex64))
        assert np.all(np.isfinite(lam))


def test_nodes_bang_equals_fresh_plan():
    """test/constructors.jl:42-70"""
    rng = np.random.default_rng(0)
    t1 = rng.random((1000, 2)) - 0.5
    t2 = rng.random((1000, 2)) - 0.5
    p1 = O.OraclePlan(t1, (32, 32))
    p2 = O.OraclePlan(t2, (32, 32))
    p2.set_nodes(t1)
    assert np.array_equal(p1.perm, p2.perm)
    f = O.random_complex((32, 32), np.float64, 1)
    assert np.array_equal(p1.forward(f), p2.forward(f))
    with pytest.raises(ValueError):
        O.OraclePlan(np.zeros((4, 1)), (2, 2))            # size(k,1) != D -> ArgumentError


def test_permutation_definition():
    """src/precomputation.jl:487-504: tiles column-major, ascending j inside a tile"""
    N = (8, 8)
    k = np.array([[0.4, 0.4], [-0.5, -0.5], [0.01, 0.0], [-0.49, -0.5], [0.4, 0.41]], dtype=np.float64)
    p = O.init_params(N, np.float64, 2, 2.0, blockSize=(8, 8))     # Nt=16 -> 2x2 tiles
    perm, counts, ks = O.precompute_blocks(k, p)
    # shifted: (0.4,0.4)->tile(0,0); (-0.5,-0.5)->(0.5,0.5)->tile(1,1); (0.01,0)->(0,0); (-0.49,-0.5)->(0.51,0.5)->(1,1)
    assert perm.tolist() == [0, 2, 4, 1, 3]
    assert counts.tolist() == [3, 0, 0, 2]


@pytest.mark.parametrize("N,T,m", [((255,), np.float64, 5), ((31, 33), np.float64, 4), ((11, 12, 14), np.float32, 3)])
def test_c_twin_matches_numpy_oracle(N, T, m):
    """the C/OpenMP restatement of the blocked algorithm (used as bench cpu_baseline) == numpy oracle"""
    from oracle.cpu_ref import CpuRefPlan
    D, M = len(N), int(np.prod(N))
    k = O.random_nodes(M, D, T, seed=1)
    po = O.OraclePlan(k, N, m=m, sigma=2.0)
    pc = CpuRefPlan(k, N, m=m, sigma=2.0)
    assert np.array_equal(po.perm, pc.perm)
    fHat = O.random_complex(M, T, 2)
    f = O.random_complex(N, T, 3)
    tol = 1e-13 if T == np.float64 else 1e-5
    assert rel(pc.adjoint(fHat), po.adjoint(fHat)) < tol
    assert rel(pc.forward(f), po.forward(f)) < tol


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(HERE, "golden", "*.npz"))))
def test_golden_fixtures(path):
    """committed vectors (tests/golden/make_golden.py): the oracle must keep reproducing them, and they
    must stay within the reference's tolerance of the NDFT"""
    z = np.load(path)
    k = z["k"]
    tol = 1e-13 if k.dtype == np.float64 else 1e-6
    if "toeplitz_lambda" in z:         # NFFTTools fixtures: Toeplitz kernel / apply (test/testToeplitz.jl) and sdc
        shape = tuple(int(n) for n in z["shape"])
        lam = O.calculate_toeplitz_kernel(shape, k, m=4, sigma=2.0)
        assert rel(lam, z["toeplitz_lambda"]) < tol
        assert _approx(z["toeplitz_lambda"], z["toeplitz_explicit"], 1e-6 if k.dtype == np.float64 else 1e-5)
        assert rel(O.convolve_toeplitz_kernel(z["y"], lam), z["toeplitz_out"]) < 10 * tol
        p = O.OraclePlan(k, shape, m=4, sigma=2.0)
        assert np.abs(O.sdc(p, iters=10) / z["sdc"] - 1).max() < (1e-10 if k.dtype == np.float64 else 1e-4)
        return
    N = tuple(int(n) for n in z["N"])
    window = str(z["window"]) if "window" in z else "kaiser_bessel"
    p = O.OraclePlan(k, N, m=int(z["m"]), sigma=2.0, precompute=int(z["pre"]), window=window)
    assert np.array_equal(p.perm, z["perm"])
    assert rel(p.forward(z["f"]), z["forward"]) < tol
    assert rel(p.adjoint(z["fHat"]), z["adjoint"]) < tol
    bar = {3: 3e-5, 4: 1e-6, 5: WINDOW_EPS[window]}[int(z["m"])]
    assert rel(z["forward"], z["ndft"]) < bar
    assert rel(z["adjoint"], z["ndft_adjoint"]) < bar
