"""Regenerates tests/golden/*.npz.

The reference (pure Julia) cannot run in this image and stores no golden vectors, so these
fixtures are produced by the CPU oracle (oracle/nfft_oracle.py) on seeded inputs.  They pin the ORACLE against
silent drift and give the GPU tests a committed vector to compare with; they are not reference outputs.
Run from the repo root:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import nfft_oracle as O  # noqa: E402

CASES = [
    ("d1_f64_m4", (64,), np.float64, 4, O.POLYNOMIAL, 200, "kaiser_bessel"),
    ("d2_f64_m4", (16, 12), np.float64, 4, O.POLYNOMIAL, 300, "kaiser_bessel"),
    ("d2_f32_m4_lin", (16, 12), np.float32, 4, O.LINEAR, 300, "kaiser_bessel"),
    ("d3_f32_m3", (12, 10, 8), np.float32, 3, O.POLYNOMIAL, 400, "kaiser_bessel"),
    ("d3_f64_m5_full", (8, 8, 8), np.float64, 5, O.FULL, 250, "kaiser_bessel"),
    # the other window pairs (src/windowFunctions.jl:41-134)
    ("d2_f64_m5_cosh", (16, 12), np.float64, 5, O.POLYNOMIAL, 300, "cosh_type"),
    ("d1_f64_m5_gauss_full", (64,), np.float64, 5, O.FULL, 200, "gauss"),
    ("d2_f32_m5_spline_lin", (16, 12), np.float32, 5, O.LINEAR, 300, "spline"),
    ("d3_f64_m5_kbrev", (8, 8, 8), np.float64, 5, O.TENSOR, 250, "kaiser_bessel_rev"),
]

# Toeplitz operator and sdc (NFFTTools): (name, shape, T, M)
TOOLS = [("tools_d2_f64", (12, 10), np.float64, 400), ("tools_d3_f32", (8, 6, 6), np.float32, 500)]


def main():
    out = os.path.dirname(os.path.abspath(__file__))
    for name, N, T, m, pre, M, window in CASES:
        k = O.random_nodes(M, len(N), T, seed=21)
        k[0] = 0.5
        k[1] = -0.5
        p = O.OraclePlan(k, N, m=m, sigma=2.0, precompute=pre, window=window)
        f = O.random_complex(N, T, 22)
        fh = O.random_complex(M, T, 23)
        np.savez_compressed(os.path.join(out, name + ".npz"), N=np.array(N), m=m, pre=pre, k=k, f=f, fHat=fh,
                            blockSize=np.array(p.p.blockSize), perm=p.perm, forward=p.forward(f),
                            adjoint=p.adjoint(fh), ndft=O.ndft(k, f), ndft_adjoint=O.ndft_adjoint(k, N, fh),
                            window=np.array(window))
        print("wrote", name)
    for name, shape, T, M in TOOLS:
        k = O.random_nodes(M, len(shape), T, seed=31)
        lam = O.calculate_toeplitz_kernel(shape, k, m=4, sigma=2.0)
        y = O.random_complex(shape, T, 32)
        p = O.OraclePlan(k, shape, m=4, sigma=2.0)
        np.savez_compressed(os.path.join(out, name + ".npz"), shape=np.array(shape), k=k, toeplitz_lambda=lam,
                            toeplitz_explicit=O.calculate_toeplitz_kernel_explicit(shape, k), y=y,
                            toeplitz_out=O.convolve_toeplitz_kernel(y, lam), sdc=O.sdc(p, iters=10),
                            blockSize=np.array(p.p.blockSize))
        print("wrote", name)


if __name__ == "__main__":
    main()
