"""
oracle/nfft_oracle.py -- CPU restatement of JuliaMath/NFFT.jl's NFFT hot path.

TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference leg may import this module.  The product path
(nfft.jl_b200/) never imports it and has no CPU fallback.

PARITY STATUS: *parity unpinned at the library boundaries*.  The reference is pure
Julia (no julia binary in this image or on the GPU box) and stores no golden vectors
(SURVEY.md section 8c), so this restatement cannot be checked bit-for-bit against the
reference.  It is pinned at the pipeline level by everything the reference's own tests
pin: tolerance vs. the NDFT for every precompute mode (test/accuracy.jl:41-73), the
closed-form sdc known answer 1/prod(N) (test/samplingDensity.jl:10-27), the issue-106
LUT boundary regression (test/issues.jl:1-17), nodes!==fresh plan
(test/constructors.jl:42-70) and the published error-vs-m band
(benchmark/paper/img/accuracy_m_D2.tex:40-47,85-92).  See tests/test_oracle_pins.py.

Third-party arithmetic that is not under /root/reference and is substituted here:
  FFTW (FFTW.jl compat 1.5)            -> scipy.fft (pocketfft); DFT is unique.
  SpecialFunctions.besseli             -> scipy.special.i0
  BasicInterpolators.ChebyshevInterpolator(f,1,N,30) (compat 0.6.5/0.7)
                                        -> 30-point Chebyshev-Lobatto interpolant restated
                                           below (cheb30=True) or the exact function.
  LAPACK least squares `\\`             -> numpy.linalg.lstsq

All citations are file:line under /root/reference.  Arrays follow the Julia layout
conventions translated to numpy: nodes k have shape (M, D) (row j = node j, i.e. the
Julia D x M column-major buffer), image arrays f have Julia shape N stored with the
FIRST dimension fastest, which in numpy is an array of shape N[::-1] in C order.  To keep
the oracle readable all public functions take/return arrays indexed [i1, i2, ..., iD] in
*Fortran order* (np.asfortranarray) so f[i1,i2,i3] means exactly what it means in Julia.
"""
from __future__ import annotations

import dataclasses
import numpy as np
from scipy import fft as _sfft
from scipy import special as _special

# AbstractNFFTs/src/misc.jl:13-18
FULL, TENSOR, LINEAR, POLYNOMIAL = 1, 2, 3, 4


# --------------------------------------------------------------------------------------
# parameters
# --------------------------------------------------------------------------------------
def reltol_to_params(reltol):
    """AbstractNFFTs/src/misc.jl:44-48"""
    w = int(np.ceil(np.log(1.0 / reltol) / np.log(10.0))) + 1
    return w // 2, 2.0


def params_to_reltol(m, sigma):
    """AbstractNFFTs/src/misc.jl:50-53"""
    return 10.0 ** (-(2 * m - 1))


def accuracy_params(m=None, sigma=None, reltol=None):
    """AbstractNFFTs/src/misc.jl:66-81 -- `m` alone is ignored."""
    if reltol is not None:
        m_, s_ = reltol_to_params(reltol)
        return m_, s_, reltol
    if m is not None and sigma is not None:
        return int(m), float(sigma), params_to_reltol(m, sigma)
    r = 1e-9
    m_, s_ = reltol_to_params(r)
    return m_, s_, r


def default_block_size(Nt, d):
    """src/precomputation.jl:59-77 (d is 0-based here)."""
    D = len(Nt)
    if D == 1:
        return min(1024, Nt[d])
    if D == 2:
        return min(64, Nt[d])
    return min(16, Nt[d]) if d < 3 else 1


@dataclasses.dataclass
class Params:
    T: type
    N: tuple
    Nt: tuple            # oversampled grid size (N-tilde)
    m: int
    sigma: float         # effective sigma, stored in T  (src/precomputation.jl:29)
    reltol: float
    LUTSize: int
    blockSize: tuple
    precompute: int
    b: float             # Kaiser-Bessel shape parameter, evaluated in T
    window: str = "kaiser_bessel"

    @property
    def D(self):
        return len(self.N)


def init_params(N, T=np.float64, m=None, sigma=None, reltol=None, precompute=POLYNOMIAL,
                blockSize=None, window="kaiser_bessel"):
    """src/precomputation.jl:3-56 (dims = 1:D only)."""
    N = tuple(int(n) for n in N)
    m, sigma, reltol = accuracy_params(m, sigma, reltol)
    m2K = [1, 3, 7, 9, 14, 17, 20, 23, 24]
    K = m2K[min(m + 1, len(m2K)) - 1]
    LUTSize = (2 ** K) * m
    # params.sigma is stored as T before it is used (NFFTParams{T,D}.sigma::T)
    sig_T = T(sigma)
    # the product sigma*N[d] is a T product (Float32 * Int -> Float32 in Julia), src/precomputation.jl:25-27
    Nt = tuple((int(np.ceil(float(sig_T * T(n)))) // 2) * 2 for n in N)
    sig_eff = T(Nt[0] / N[0])
    if blockSize is None:
        blockSize = tuple(default_block_size(Nt, d) for d in range(len(N)))
    blockSize = tuple(int(b) for b in blockSize)
    # b = pi*(2-1/sigma) with sigma::T  => evaluated in T (src/windowFunctions.jl:23)
    b = T(np.pi) * (T(2) - T(1) / sig_eff) if T == np.float32 else np.pi * (2.0 - 1.0 / float(sig_eff))
    if window not in WINDOWS:
        raise ValueError("Window %s not yet implemented!" % window)       # src/windowFunctions.jl:16
    return Params(T=T, N=N, Nt=Nt, m=m, sigma=float(sig_eff), reltol=reltol, LUTSize=LUTSize,
                  blockSize=blockSize, precompute=precompute, b=float(T(b)), window=window)


# --------------------------------------------------------------------------------------
# window pair  (src/windowFunctions.jl:21-39)
# --------------------------------------------------------------------------------------
def window_kaiser_bessel(x, m, b, dtype=np.float64):
    """phi in *grid units* (the reference called with N-tilde = 1), evaluated in `dtype`.
    src/windowFunctions.jl:21-34."""
    x = np.asarray(x, dtype=dtype)
    mm = dtype(m)
    bb = dtype(b)
    ax = np.abs(x)
    inside = ax < mm
    arg = np.sqrt(np.where(inside, mm * mm - x * x, dtype(1)))
    y_in = np.sinh(bb * arg) / (arg * dtype(np.pi))
    y = np.where(inside, y_in, np.where(ax > mm, dtype(0), bb / dtype(np.pi)))
    return y.astype(dtype)


def window_kaiser_bessel_hat(n, Nt, m, b):
    """src/windowFunctions.jl:36-39 (always evaluated in Float64, see module docstring)."""
    n = np.asarray(n, dtype=np.float64)
    return _special.i0(m * np.sqrt(float(b) ** 2 - (2.0 * np.pi * n / Nt) ** 2))


# --------------------------------------------------------------------------------------
# the other window pairs (src/windowFunctions.jl:41-134), all in *grid units* x = N-tilde*k
# --------------------------------------------------------------------------------------
WINDOWS = ("kaiser_bessel", "gauss", "spline", "kaiser_bessel_rev", "cosh_type", "exp_sqrt")


def window_kaiser_bessel_rev(x, m, b, dtype=np.float64):
    """src/windowFunctions.jl:41-50: 0.5/m * I0(m b sqrt(1-(x/m)^2)) for |x| < m, else 0."""
    x = np.asarray(x, dtype=dtype)
    mm, bb = dtype(m), dtype(b)
    inside = np.abs(x) < mm
    arg = mm * bb * np.sqrt(np.where(inside, dtype(1) - (x / mm) ** 2, dtype(0)))
    y = dtype(0.5) / mm * _special.i0(arg.astype(np.float64)).astype(dtype)
    return np.where(inside, y, dtype(0)).astype(dtype)


def window_kaiser_bessel_rev_hat(n, Nt, m, b):
    """src/windowFunctions.jl:52-57: real(sinc(sqrt(complex((2 pi m n/Nt)^2-(m b)^2))/pi)),
    i.e. sinh(a)/a below the cut-off and sin(a)/a above it."""
    n = np.asarray(n, dtype=np.float64)
    q = (2.0 * np.pi * m * n / Nt) ** 2 - (m * float(b)) ** 2
    a = np.sqrt(np.abs(q))
    safe = np.where(a == 0, 1.0, a)
    return np.where(a == 0, 1.0, np.where(q < 0, np.sinh(safe) / safe, np.sin(safe) / safe))


def window_gauss(x, m, dtype=np.float64):
    """src/windowFunctions.jl:60-68 with b = m/pi."""
    x = np.asarray(x, dtype=dtype)
    b = dtype(m) / dtype(np.pi)
    y = dtype(1) / np.sqrt(dtype(np.pi) * b) * np.exp(-(x * x) / b)
    return np.where(np.abs(x) < dtype(m), y, dtype(0)).astype(dtype)


def window_gauss_hat(n, Nt, m):
    """src/windowFunctions.jl:70-73."""
    n = np.asarray(n, dtype=np.float64)
    return np.exp(-((np.pi * n / Nt) ** 2) * (m / np.pi))


def cbspline(order, x):
    """Cardinal B-spline of the given order on the knots 0..order (src/windowFunctions.jl:75-86),
    the same Cox-de Boor recursion evaluated bottom-up."""
    x = np.asarray(x)
    dtype = x.dtype.type
    a = [((x - j >= 0) & (x - j < 1)).astype(dtype) for j in range(order)]
    for n in range(2, order + 1):
        a = [((x - j) / dtype(n - 1) * a[j] + (dtype(n) - (x - j)) / dtype(n - 1) * a[j + 1]).astype(dtype)
             for j in range(order - n + 1)]
    return a[0]


def window_spline(x, m, dtype=np.float64):
    """src/windowFunctions.jl:88-95: cbspline(2m, x + m) for |x| < m."""
    x = np.asarray(x, dtype=dtype)
    inside = np.abs(x) < dtype(m)
    y = cbspline(2 * m, np.where(inside, x + dtype(m), dtype(0)))
    return np.where(inside, y, dtype(0)).astype(dtype)


def window_spline_hat(n, Nt, m):
    """src/windowFunctions.jl:97-99 (Julia sinc(x) = sin(pi x)/(pi x) == np.sinc)."""
    n = np.asarray(n, dtype=np.float64)
    return np.sinc(n / Nt) ** (2 * m)


def window_cosh_type(x, m, sigma, dtype=np.float64):
    """src/windowFunctions.jl:104-116."""
    x = np.asarray(x, dtype=dtype)
    beta = dtype(np.pi) * dtype(m) * (dtype(2) - dtype(1) / dtype(sigma))
    inside = np.abs(x) < dtype(m)
    alpha = np.sqrt(np.where(inside, dtype(1) - (x / dtype(m)) ** 2, dtype(1)))
    y = dtype(1) / (np.cosh(beta) - dtype(1)) * (np.cosh(beta * alpha) - dtype(1)) / alpha
    return np.where(inside, y, dtype(0)).astype(dtype)


def window_cosh_type_hat(n, Nt, m, sigma):
    """src/windowFunctions.jl:118-134."""
    n = np.asarray(n, dtype=np.float64)
    beta = np.pi * m * (2.0 - 1.0 / sigma)
    gamma = beta / (2 * np.pi)
    zeta = np.pi / (np.cosh(beta) - 1.0) * m
    arg = m * n / Nt
    two = 2 * np.pi * arg
    lo = _special.i0(np.sqrt(np.maximum(beta ** 2 - two ** 2, 0.0))) - _special.j0(two)
    hi = _special.j0(np.sqrt(np.maximum(two ** 2 - beta ** 2, 0.0))) - _special.j0(two)
    eq = 1.0 - _special.j0(beta)
    a = np.abs(arg)
    return zeta * np.where(a < gamma, lo, np.where(a > gamma, hi, eq))


def exp_sqrt_beta(m, sigma, dtype=np.float64):
    """shape parameter of the exp-sqrt ("exponential of semicircle") window; NOT a reference window: the north_star's
    on-the-fly option, beta = 0.97 pi 2m (1 - 1/(2 sigma))"""
    return dtype(0.97) * dtype(np.pi) * dtype(2 * m) * (dtype(1) - dtype(0.5) / dtype(sigma))


def window_exp_sqrt(x, m, sigma, dtype=np.float64):
    x = np.asarray(x, dtype=dtype)
    beta = exp_sqrt_beta(m, sigma, dtype)
    inside = np.abs(x) < dtype(m)
    a = np.sqrt(np.where(inside, dtype(1) - (x / dtype(m)) ** 2, dtype(1)))
    return np.where(inside, np.exp(beta * (a - dtype(1))), dtype(0)).astype(dtype)


def window_exp_sqrt_hat(n, Nt, m, sigma, beta=None):
    """phi_hat(n) = int_{-m}^{m} phi(x) cos(2 pi n x / Nt) dx by Gauss-Legendre quadrature (no closed form)"""
    n = np.atleast_1d(np.asarray(n, dtype=np.float64))
    gx, gw = np.polynomial.legendre.leggauss(96)
    x = 0.5 * m * (gx + 1.0)
    beta = float(exp_sqrt_beta(m, sigma)) if beta is None else float(beta)
    phi = np.exp(beta * (np.sqrt(1.0 - (x / m) ** 2) - 1.0))
    return m * (np.cos(2.0 * np.pi * np.outer(n, x) / Nt) * (gw * phi)).sum(axis=1)


def window_eval(p: "Params", x, dtype=np.float64):
    """getWindow(window)[1] evaluated in grid units (src/windowFunctions.jl:4-19)."""
    w = p.window
    if w == "kaiser_bessel":
        return window_kaiser_bessel(x, p.m, p.b, dtype)
    if w == "kaiser_bessel_rev":
        return window_kaiser_bessel_rev(x, p.m, p.b, dtype)
    if w == "gauss":
        return window_gauss(x, p.m, dtype)
    if w == "spline":
        return window_spline(x, p.m, dtype)
    if w == "cosh_type":
        return window_cosh_type(x, p.m, p.sigma, dtype)
    if w == "exp_sqrt":
        return window_exp_sqrt(x, p.m, p.sigma, dtype)
    raise ValueError("Window %s not yet implemented!" % w)


def window_hat_eval(p: "Params", n, Nt):
    """getWindow(window)[2] (always Float64)."""
    w = p.window
    if w == "kaiser_bessel":
        return window_kaiser_bessel_hat(n, Nt, p.m, p.b)
    if w == "kaiser_bessel_rev":
        return window_kaiser_bessel_rev_hat(n, Nt, p.m, p.b)
    if w == "gauss":
        return window_gauss_hat(n, Nt, p.m)
    if w == "spline":
        return window_spline_hat(n, Nt, p.m)
    if w == "cosh_type":
        return window_cosh_type_hat(n, Nt, p.m, p.sigma)
    if w == "exp_sqrt":
        beta = float(exp_sqrt_beta(p.m, p.T(p.sigma), p.T)) if p.T == np.float32 else None   # the Float32 plan's beta, as the device has it
        return window_exp_sqrt_hat(n, Nt, p.m, p.sigma, beta).reshape(np.shape(n))
    raise ValueError("Window %s not yet implemented!" % w)


def index_offset(N):
    """src/precomputation.jl:345 (returns the 1-based offset used by the reference)."""
    return (-1 - N // 2) if N % 2 == 0 else (-1 - (N - 1) // 2)


def _cheb_lobatto_interp(fun, a, b, n, xq):
    """n-point Chebyshev-Lobatto interpolant of fun on [a,b], evaluated at xq with the
    barycentric formula (mathematically the same polynomial BasicInterpolators builds)."""
    kk = np.arange(n)
    xi = np.cos(np.pi * kk / (n - 1))
    xs = 0.5 * (a + b) + 0.5 * (b - a) * xi
    fs = fun(xs)
    w = (-1.0) ** kk
    w[0] *= 0.5
    w[-1] *= 0.5
    xq = np.asarray(xq, dtype=np.float64)
    diff = xq[:, None] - xs[None, :]
    exact = np.isclose(diff, 0.0, atol=0.0, rtol=0.0)
    diff[exact] = 1.0
    num = (w / diff * fs).sum(axis=1)
    den = (w / diff).sum(axis=1)
    out = num / den
    rows, cols = np.nonzero(exact)
    out[rows] = fs[cols]
    return out


def window_hat_inv_lut(p: Params, cheb30=False):
    """src/precomputation.jl:347-358: LUT_d[j] = 1/win_hat(j + indexOffset(N_d)), j=1..N_d."""
    luts = []
    for d in range(p.D):
        N, Nt = p.N[d], p.Nt[d]
        j = np.arange(1, N + 1, dtype=np.float64)
        if cheb30 and N > 1:
            kap = lambda x: window_hat_eval(p, x + index_offset(N), Nt)
            vals = _cheb_lobatto_interp(kap, 1.0, float(N), 30, j)
        else:
            vals = window_hat_eval(p, j + index_offset(N), Nt)
        luts.append((1.0 / vals).astype(p.T))
    return luts


# --------------------------------------------------------------------------------------
# window tables (src/precomputation.jl:291-320)
# --------------------------------------------------------------------------------------
def precompute_lin_interp(p: Params):
    """src/precomputation.jl:291-300: K+2 samples of phi on [0, m + m/K]."""
    K = p.LUTSize
    step = p.m / K
    y = np.arange(K + 2, dtype=np.float64) * step
    return window_eval(p, y, np.float64).astype(p.T)


def precompute_poly_interp(p: Params):
    """src/precomputation.jl:302-320: (2m+1) x 2m coefficient matrix, column l (0-based here)
    approximates phi(m - (l+1) + 0.5 + t), t in [-1/2, 1/2], by least squares on 2(2m+1)
    equispaced samples."""
    m = p.m
    deg = 2 * m + 1
    K = 2 * m
    ns = 2 * deg
    t = np.linspace(-0.5, 0.5, ns)
    V = np.vander(t, deg, increasing=True)           # ns x deg  (== V' in the reference)
    P = np.empty((deg, K), dtype=np.float64)
    for l in range(1, K + 1):
        y = (-(l - 0.5) + m) + t
        samples = window_eval(p, y, np.float64)
        P[:, l - 1] = np.linalg.lstsq(V, samples, rcond=None)[0]
    return P.astype(p.T)


# --------------------------------------------------------------------------------------
# nodes (src/utils.jl:32-55)
# --------------------------------------------------------------------------------------
def check_nodes(k):
    """src/utils.jl:46-55 -- raises ValueError (the ArgumentError of the reference)."""
    if not np.all(np.abs(k) <= 0.5):
        raise ValueError("Nodes k need to be within the range [-1/2, 1/2)")


def shift_nodes(k):
    """src/utils.jl:32-44 on a copy (src/precomputation.jl:462-463), in the dtype of k."""
    T = k.dtype.type
    ks = k.copy()
    neg = ks < T(0)
    ks[neg] += T(1)
    one = ks == T(1)
    ks[one] -= np.finfo(T).eps
    return ks


def tile_keys(ks, p: Params):
    """src/precomputation.jl:494-496: idx_d = unsafe_trunc(Int, k[d,j]*Nt[d]) // blockSize[d]
    (0-based here), linearised column-major over numBlocks = ceil(Nt/blockSize)."""
    T = p.T
    nb = [-(-p.Nt[d] // p.blockSize[d]) for d in range(p.D)]
    key = np.zeros(ks.shape[0], dtype=np.int64)
    stride = 1
    for d in range(p.D):
        kscale = ks[:, d] * T(p.Nt[d])                 # product rounded in T
        c = np.trunc(kscale).astype(np.int64)
        key += (c // p.blockSize[d]) * stride
        stride *= nb[d]
    return key, nb


def precompute_blocks(k, p: Params):
    """src/precomputation.jl:459-520 -> (perm, counts): perm = concatenation of nodesInBlock[l]
    over l in column-major tile order, 0-based node ids, ascending j inside a tile."""
    ks = shift_nodes(np.ascontiguousarray(k, dtype=p.T))
    key, nb = tile_keys(ks, p)
    perm = np.argsort(key, kind="stable")
    counts = np.bincount(key, minlength=int(np.prod(nb)))
    return perm.astype(np.int64), counts.astype(np.int64), ks


# --------------------------------------------------------------------------------------
# per-node taps: cells and weights for every precompute mode
# --------------------------------------------------------------------------------------
def _taps_blocked(ks_d, Nt_d, p: Params, tables):
    """Blocked formulation on *shifted* nodes (src/precomputation.jl:524-554 and :170-222).
    Returns (cells (M,2m) int64 0-based wrapped, weights (M,2m) T)."""
    T = p.T
    m = p.m
    kscale = ks_d * T(Nt_d)
    c = np.trunc(kscale).astype(np.int64)
    off = c - m + 1                                       # 0-based first tap cell
    taps = np.arange(2 * m, dtype=np.int64)
    cells = (off[:, None] + taps[None, :]) % Nt_d
    if p.precompute in (POLYNOMIAL, TENSOR):
        # idx = kscale - off - m + 1 - 0.5 (:545) == frac - 1/2 exactly
        idx = (((kscale - off.astype(T)) - T(m)) + T(1)).astype(np.float64) - 0.5
        idx = idx.astype(T)
        P = tables["poly"]                               # (2m+1, 2m) in T
        w = np.zeros((ks_d.shape[0], 2 * m), dtype=T)
        for l in range(2 * m):
            acc = np.full(ks_d.shape[0], P[-1, l], dtype=T)
            for r in range(P.shape[0] - 2, -1, -1):      # Horner == evalpoly (:215-222)
                acc = (acc * idx + P[r, l]).astype(T)
            w[:, l] = acc
    elif p.precompute == LINEAR:
        scale = p.LUTSize // m
        idx = ((kscale - off.astype(T)) * T(scale)).astype(T)          # :543
        idx_int = np.floor(idx).astype(np.int64)
        alpha = (idx - idx_int.astype(T)).astype(T)
        lut = tables["lin"]
        i1 = np.abs(idx_int[:, None] - taps[None, :] * scale)           # :194 (0-based)
        i2 = np.abs(idx_int[:, None] - taps[None, :] * scale + 1)       # :195
        w = (lut[i1] + alpha[:, None] * (lut[i2] - lut[i1])).astype(T)  # :197
    elif p.precompute == FULL:
        # exact window at the blocked coordinates: distance = (kscale - off) - l
        dist = (kscale - off.astype(T))[:, None] - taps[None, :].astype(T)
        w = window_eval(p, dist, T)
    else:
        raise ValueError("precompute mode not supported")
    return cells, w


def _taps_nonblocked(k_d, Nt_d, p: Params, tables):
    """Non-blocking formulation on raw nodes (src/precomputation.jl:124-166)."""
    T = p.T
    m = p.m
    kscale = k_d * T(Nt_d)
    off = np.floor(kscale).astype(np.int64) - m + 1
    taps = np.arange(2 * m, dtype=np.int64)
    cells = (off[:, None] + taps[None, :]) % Nt_d
    if p.precompute == FULL:
        # win((kscale-(l-1)-off)/Nt, Nt, m, sigma) in T (:131)
        x = ((kscale[:, None] - taps[None, :].astype(T)) - off[:, None].astype(T)).astype(T)
        if p.window != "kaiser_bessel":
            return cells, window_eval(p, x, T)
        kk = (x / T(Nt_d)).astype(T)
        m_by_N = T(m) / T(Nt_d)
        ak = np.abs(kk)
        inside = ak < m_by_N
        arg = np.sqrt(np.where(inside, T(m) ** 2 - T(Nt_d) ** 2 * kk * kk, T(1))).astype(T)
        y_in = np.sinh(T(p.b) * arg) / (arg * T(np.pi))
        w = np.where(inside, y_in, np.where(ak > m_by_N, T(0), T(p.b) / T(np.pi))).astype(T)
    elif p.precompute == LINEAR:
        idx = (((kscale - off.astype(T)) * T(p.LUTSize)) / T(m)).astype(T)   # :145
        scale = p.LUTSize // m
        idx_int = np.floor(idx).astype(np.int64)
        alpha = (idx - idx_int.astype(T)).astype(T)
        lut = tables["lin"]
        i1 = np.abs(idx_int[:, None] - taps[None, :] * scale)
        i2 = np.abs(idx_int[:, None] - taps[None, :] * scale + 1)
        w = (lut[i1] + alpha[:, None] * (lut[i2] - lut[i1])).astype(T)
    else:
        idx = (((kscale - off.astype(T)) - T(m)) + T(0.5)).astype(T)        # :161
        P = tables["poly"]
        w = np.zeros((k_d.shape[0], 2 * m), dtype=T)
        for l in range(2 * m):
            acc = np.full(k_d.shape[0], P[-1, l], dtype=T)
            for r in range(P.shape[0] - 2, -1, -1):
                acc = (acc * idx + P[r, l]).astype(T)
            w[:, l] = acc
    return cells, w


# --------------------------------------------------------------------------------------
# the plan
# --------------------------------------------------------------------------------------
class OraclePlan:
    """Restatement of NFFTPlan (src/implementation.jl:16-141) for dims = 1:D."""

    def __init__(self, k, N, T=None, m=None, sigma=None, reltol=None, precompute=POLYNOMIAL,
                 blocking=True, blockSize=None, cheb30=False, window="kaiser_bessel"):
        k = np.asarray(k)
        if k.ndim == 1:
            k = k[:, None]                                # derived.jl:23-27
        T = T or k.dtype.type
        if isinstance(N, (int, np.integer)):
            N = (int(N),)
        if k.shape[1] != len(N):
            raise ValueError("Nodes x have dimension %d != %d" % (k.shape[1], len(N)))  # :19-21
        self.p = init_params(N, T, m, sigma, reltol, precompute, blockSize, window)
        self.T = T
        self.blocking = blocking
        self.cheb30 = cheb30
        self.tables = {}
        if precompute == LINEAR:
            self.tables["lin"] = precompute_lin_interp(self.p)
        elif precompute in (POLYNOMIAL, TENSOR):
            self.tables["poly"] = precompute_poly_interp(self.p)
        self.windowHatInvLUT = window_hat_inv_lut(self.p, cheb30)
        self.set_nodes(k)

    # nodes!  (src/implementation.jl:108-141)
    def set_nodes(self, k):
        k = np.ascontiguousarray(k, dtype=self.T)
        check_nodes(k)
        self.k = k
        self.M = k.shape[0]
        self.perm, self.counts, self.ks = precompute_blocks(k, self.p)
        p = self.p
        self.cells, self.w = [], []
        for d in range(p.D):
            if self.blocking and p.precompute != FULL:
                c, w = _taps_blocked(self.ks[:, d], p.Nt[d], p, self.tables)
            elif self.blocking:
                c, w = _taps_blocked(self.ks[:, d], p.Nt[d], p, self.tables)
            else:
                c, w = _taps_nonblocked(k[:, d], p.Nt[d], p, self.tables)
            self.cells.append(c)
            self.w.append(w)
        return self

    @property
    def N(self):
        return self.p.N

    @property
    def Nt(self):
        return self.p.Nt

    # ---- B and B^H  (src/convolution.jl:72-99, :176-200) -------------------------------
    def _flat_taps(self):
        p = self.p
        D, L = p.D, 2 * p.m
        lin = self.cells[0]
        wt = self.w[0]
        stride = p.Nt[0]
        for d in range(1, D):
            lin = (lin[:, :, None] + stride * self.cells[d][:, None, :]).reshape(self.M, -1)
            wt = (wt[:, :, None] * self.w[d][:, None, :]).reshape(self.M, -1)
            stride *= p.Nt[d]
        return lin, wt

    def convolve(self, g):
        """fHat[j] = sum_taps prod_d w_d * g[cells]  (g is Fortran-indexed, shape Nt)."""
        lin, wt = self._flat_taps()
        gf = np.asarray(g).reshape(-1, order="F")
        out_t = np.result_type(gf.dtype, self.T)
        return (wt.astype(out_t) * gf[lin]).sum(axis=1).astype(out_t)

    def convolve_transpose(self, fHat, out_dtype=None):
        lin, wt = self._flat_taps()
        fHat = np.asarray(fHat)
        out_t = out_dtype or np.result_type(fHat.dtype, self.T)
        g = np.zeros(int(np.prod(self.p.Nt)), dtype=out_t)
        np.add.at(g, lin.ravel(), (wt.astype(out_t) * fHat[:, None].astype(out_t)).ravel())
        return g.reshape(self.p.Nt, order="F")

    # ---- D and D^H  (src/deconvolution.jl:22-92) ---------------------------------------
    def _grid_index(self):
        p = self.p
        idx = []
        for d in range(p.D):
            n = np.arange(p.N[d]) - p.N[d] // 2
            idx.append(n % p.Nt[d])
        return idx

    def _lut_product(self, v):
        for d in range(self.p.D):
            shape = [1] * self.p.D
            shape[d] = self.p.N[d]
            v = v * self.windowHatInvLUT[d].reshape(shape)
        return v

    def deconvolve(self, f):
        cT = np.complex64 if self.T == np.float32 else np.complex128
        g = np.zeros(self.p.Nt, dtype=cT, order="F")
        v = self._lut_product(np.asarray(f).astype(cT))
        g[np.ix_(*self._grid_index())] = v
        return g

    def deconvolve_transpose(self, g):
        cT = np.complex64 if self.T == np.float32 else np.complex128
        v = np.asarray(g)[np.ix_(*self._grid_index())].astype(cT)
        return np.asfortranarray(self._lut_product(v).astype(cT))

    # ---- mul!  (src/implementation.jl:155-193) ----------------------------------------
    def forward(self, f, workers=1):
        g = self.deconvolve(f)
        g = _sfft.fftn(g, workers=workers)                 # sign -1, unnormalised
        return self.convolve(g)

    def adjoint(self, fHat, workers=1):
        g = self.convolve_transpose(np.asarray(fHat))
        cT = np.complex64 if self.T == np.float32 else np.complex128
        g = _sfft.ifftn(g.astype(cT), norm="forward", workers=workers)   # bfft: sign +1, unnormalised
        return self.deconvolve_transpose(g)


# --------------------------------------------------------------------------------------
# NDFT ground truth (src/direct.jl:99-151)
# --------------------------------------------------------------------------------------
def _freqs(N):
    return [np.arange(n) - n // 2 for n in N]


def ndft(k, f, chunk=2048):
    k = np.asarray(k, dtype=np.float64)
    if k.ndim == 1:
        k = k[:, None]
    f = np.asarray(f)
    N = f.shape
    out = np.zeros(k.shape[0], dtype=np.complex128)
    fr = _freqs(N)
    for s in range(0, k.shape[0], chunk):
        kk = k[s:s + chunk]
        acc = f.astype(np.complex128)
        # contract one dimension at a time: out_j = sum_n f[n] prod_d e^{-2 pi i n_d k_jd}
        e = [np.exp(-2j * np.pi * np.outer(kk[:, d], fr[d])) for d in range(len(N))]
        if len(N) == 1:
            out[s:s + chunk] = e[0] @ acc
        elif len(N) == 2:
            out[s:s + chunk] = np.einsum("ja,jb,ab->j", e[0], e[1], acc, optimize=True)
        elif len(N) == 3:
            out[s:s + chunk] = np.einsum("ja,jb,jc,abc->j", e[0], e[1], e[2], acc, optimize=True)
        elif len(N) == 4:     # test/accuracy.jl:43 runs N = (6, 5, 6, 6)
            out[s:s + chunk] = np.einsum("ja,jb,jc,jd,abcd->j", e[0], e[1], e[2], e[3], acc, optimize=True)
        else:
            raise ValueError("D<=4")
    return out


def ndft_adjoint(k, N, fHat, chunk=2048):
    k = np.asarray(k, dtype=np.float64)
    if k.ndim == 1:
        k = k[:, None]
    if isinstance(N, (int, np.integer)):
        N = (int(N),)
    fHat = np.asarray(fHat).astype(np.complex128)
    out = np.zeros(N, dtype=np.complex128)
    fr = _freqs(N)
    for s in range(0, k.shape[0], chunk):
        kk = k[s:s + chunk]
        v = fHat[s:s + chunk]
        e = [np.exp(2j * np.pi * np.outer(kk[:, d], fr[d])) for d in range(len(N))]
        if len(N) == 1:
            out += v @ e[0]
        elif len(N) == 2:
            out += np.einsum("j,ja,jb->ab", v, e[0], e[1], optimize=True)
        elif len(N) == 3:
            out += np.einsum("j,ja,jb,jc->abc", v, e[0], e[1], e[2], optimize=True)
        elif len(N) == 4:
            out += np.einsum("j,ja,jb,jc,jd->abcd", v, e[0], e[1], e[2], e[3], optimize=True)
        else:
            raise ValueError("D<=4")
    return np.asfortranarray(out)


# --------------------------------------------------------------------------------------
# NFFTTools.sdc  (NFFTTools/src/samplingDensity.jl:93-155)
# --------------------------------------------------------------------------------------
def sdc(plan: OraclePlan, iters=20):
    T = plan.T
    w = np.ones(plan.M, dtype=T)
    scaling = None
    for i in range(iters):
        g = plan.convolve_transpose(w, out_dtype=T)
        if i == 0:
            scaling = g.max()
        g = g / scaling
        tmp = plan.convolve(g) / scaling
        if np.any(tmp <= 0):
            raise ValueError("non-positive weights")
        w = (w / tmp).astype(T)
    cT = np.complex64 if T == np.float32 else np.complex128
    u = np.ones(plan.N, dtype=cT, order="F")
    wf = plan.forward(u) * w
    v = plan.adjoint(wf)
    c = np.real(v.sum()) / np.sum(np.abs(v) ** 2)
    return (w * T(c)).astype(T)


# --------------------------------------------------------------------------------------
# Toeplitz (Gram) operator  (NFFTTools/src/Toeplitz.jl)
# --------------------------------------------------------------------------------------
def calculate_toeplitz_kernel(shape, k, m=4, sigma=2.0, window="kaiser_bessel", precompute=POLYNOMIAL):
    """calculateToeplitzKernel (NFFTTools/src/Toeplitz.jl:86-93): adjoint NFFT of ones on the 2x oversampled
    image grid, fftshift, unnormalised forward FFT.  k: (M, D)."""
    k = np.asarray(k)
    shape_os = tuple(2 * int(s) for s in shape)
    p = OraclePlan(k, shape_os, m=m, sigma=sigma, window=window, precompute=precompute)
    cT = np.complex64 if p.T == np.float32 else np.complex128
    eig = p.adjoint(np.ones(k.shape[0], dtype=cT))
    return np.asfortranarray(_sfft.fftn(np.fft.fftshift(eig)).astype(cT))


def calculate_toeplitz_kernel_explicit(shape, k):
    """calculateToeplitzKernel_explicit / getMatrixElement (NFFTTools/src/Toeplitz.jl:145-168):
    lambda[i] = sum_j exp(-2 pi i k_j . (shape_os/2 - i)), i 0-based, then fftshift and FFT."""
    k = np.asarray(k)
    shape_os = tuple(2 * int(s) for s in shape)
    cT = np.complex64 if k.dtype == np.float32 else np.complex128
    lam = ndft_adjoint(k.astype(np.float64), shape_os, np.ones(k.shape[0], dtype=np.complex128))
    return np.asfortranarray(_sfft.fftn(np.fft.fftshift(lam)).astype(cT))


def convolve_toeplitz_kernel(y, lam):
    """convolveToeplitzKernel! (NFFTTools/src/Toeplitz.jl:230-244): zero-pad, FFT, multiply, normalised IFFT,
    crop.  Returns the new y (same dtype)."""
    y = np.asarray(y)
    x = np.zeros(lam.shape, dtype=lam.dtype, order="F")
    sl = tuple(slice(0, n) for n in y.shape)
    x[sl] = y
    x = _sfft.fftn(x).astype(lam.dtype) * lam
    x = _sfft.ifftn(x).astype(lam.dtype)
    return np.asfortranarray(x[sl].astype(y.dtype))


# --------------------------------------------------------------------------------------
# synthetic inputs
# --------------------------------------------------------------------------------------
def random_nodes(M, D, T, seed=1):
    rng = np.random.default_rng(seed)
    return (rng.random((M, D)) - 0.5).astype(T)


def radial_nodes(nspokes, nsamples, T):
    s = np.arange(nspokes)
    th = np.pi * s / nspokes
    r = (np.arange(nsamples) - nsamples // 2) / nsamples
    kx = np.outer(np.cos(th), r).ravel()
    ky = np.outer(np.sin(th), r).ravel()
    k = np.stack([kx, ky], axis=1)
    k = np.clip(k, -0.5, 0.5)
    return k.astype(T)


def random_complex(shape, T, seed=2, order="F"):
    rng = np.random.default_rng(seed)
    cT = np.complex64 if T == np.float32 else np.complex128
    a = (rng.standard_normal(shape) + 1j * rng.standard_normal(shape)).astype(cT)
    return np.asfortranarray(a) if order == "F" and a.ndim > 1 else a
