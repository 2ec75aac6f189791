"""Small problems through the round-2 kernels for compute-sanitizer (memcheck / racecheck / synccheck):
   compute-sanitizer --tool racecheck python scripts/sanitize_small.py"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import nfft_jl_b200 as nb
from oracle import nfft_oracle as O

def rel(a, b):
    a = np.asarray(a).ravel().astype(np.complex128); b = np.asarray(b).ravel().astype(np.complex128)
    return float(np.linalg.norm(a - b) / np.linalg.norm(b))

T = np.float32
ONLY = sys.argv[1] if len(sys.argv) > 1 else "all"          # 2d | 3d | all
# batched 2-D: batch-stationary kernels (window adjoint, direct forward, tile gather), 32 and 12 transforms
for B, N, M in ((32, (48, 40), 3000), (12, (33, 36), 1500)) if ONLY in ("all", "2d") else ():
    k = O.random_nodes(M, 2, T, seed=B)
    p = nb.plan_nfft(np.ascontiguousarray(k.T), N, m=4, σ=2.0, ntransforms=B)
    f = O.random_complex(tuple(N) + (B,), T, 4); fh = O.random_complex((M, B), T, 5)
    a = np.asarray(p * f); c = np.asarray(p.adjoint() * fh)
    p.set_kernel_mode(9)
    print("2-D batched B=%d: forward %.2e adjoint %.2e vs per-transform kernels" % (B, rel(a, p * f), rel(c, p.adjoint() * fh)), flush=True)
# 3-D lean kernels (wide TMA layout + compact layout)
k = O.random_nodes(20000, 3, T, seed=3)
for mode in (0, 12) if ONLY in ("all", "3d") else ():
    p = nb.plan_nfft(np.ascontiguousarray(k.T), (32, 32, 32), m=3, σ=2.0)
    p.set_kernel_mode(mode)
    f = O.random_complex((32, 32, 32), T, 4); fh = O.random_complex(20000, T, 5)
    a = np.asarray(p * f); c = np.asarray(p.adjoint() * fh)
    p.set_kernel_mode(9)
    print("3-D lean mode %d: forward %.2e adjoint %.2e vs round-1 kernels" % (mode, rel(a, p * f), rel(c, p.adjoint() * fh)), flush=True)
print("done")
