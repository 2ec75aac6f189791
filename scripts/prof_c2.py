"""Tiny driver for ncu captures: builds the C2 plan (or a named config) and runs forward+adjoint a few times."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import nfft_jl_b200 as nb
from oracle import nfft_oracle as O

cfg = sys.argv[1] if len(sys.argv) > 1 else "C2"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
mode = int(sys.argv[3]) if len(sys.argv) > 3 else 0            # kernel_mode (7 = register-footprint kernels)
CFG = {"C2": ((128, 128, 128), 2 ** 21, 3, np.float32), "C1": ((256, 256), 65536, 4, np.float64),
       "C4s": ((2 ** 20,), 2 ** 23, 4, np.float64), "C3s": ((512, 512), 2 ** 20, 4, np.float32)}
N, M, m, T = CFG[cfg]
k = O.random_nodes(M, len(N), T, seed=1)
p = nb.plan_nfft(torch.from_numpy(np.ascontiguousarray(k.T)).cuda(), N, m=m, σ=2.0)
p.set_kernel_mode(mode)
f = p.empty_image(); f.fill_(1.0)
fh = p.empty_out(); fh.fill_(1.0)
fo = p.empty_image(); fho = p.empty_out()
for _ in range(reps):
    nb.mul_(fho, p, f)
    nb.mul_(fo, p.adjoint(), fh)
torch.cuda.synchronize()
print("done", cfg, "kernel_mode", mode)
