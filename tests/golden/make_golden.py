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
    ("d1_f64_m4", (64,), np.float64, 4, O.POLYNOMIAL, 200),
    ("d2_f64_m4", (16, 12), np.float64, 4, O.POLYNOMIAL, 300),
    ("d2_f32_m4_lin", (16, 12), np.float32, 4, O.LINEAR, 300),
    ("d3_f32_m3", (12, 10, 8), np.float32, 3, O.POLYNOMIAL, 400),
    ("d3_f64_m5_full", (8, 8, 8), np.float64, 5, O.FULL, 250),
]


def main():
    out = os.path.dirname(os.path.abspath(__file__))
    for name, N, T, m, pre, M in CASES:
        k = O.random_nodes(M, len(N), T, seed=21)
        k[0] = 0.5
        k[1] = -0.5
        p = O.OraclePlan(k, N, m=m, sigma=2.0, precompute=pre)
        f = O.random_complex(N, T, 22)
        fh = O.random_complex(M, T, 23)
        np.savez_compressed(os.path.join(out, name + ".npz"), N=np.array(N), m=m, pre=pre, k=k, f=f, fHat=fh,
                            blockSize=np.array(p.p.blockSize), perm=p.perm, forward=p.forward(f),
                            adjoint=p.adjoint(fh), ndft=O.ndft(k, f), ndft_adjoint=O.ndft_adjoint(k, N, fh))
        print("wrote", name)


if __name__ == "__main__":
    main()
