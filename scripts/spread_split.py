"""Splits the C2 adjoint convolution stage into spread kernel and gather pass (CUDA events inside the library)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import nfft_jl_b200 as nb
from oracle import nfft_oracle as O

N, M, m, T = (128, 128, 128), 2 ** 21, 3, np.float32
mode = int(sys.argv[1]) if len(sys.argv) > 1 else 0
k = O.random_nodes(M, 3, T, seed=1)
p = nb.plan_nfft(torch.from_numpy(np.ascontiguousarray(k.T)).cuda(), N, m=m, σ=2.0)
p.set_kernel_mode(mode)
fh = p.empty_out(); fh.fill_(1.0); fo = p.empty_image()
f = p.empty_image(); f.fill_(1.0); fho = p.empty_out()
ts = nb.TimingStats()
acc = np.zeros(4); n = 0
for i in range(13):
    nb.mul_(fo, p.adjoint(), fh, timing=ts)
    kt = p.kernel_times()
    nb.mul_(fho, p, f, timing=ts)
    kt2 = p.kernel_times()
    if i >= 3:
        acc += [ts.conv_adjoint * 1e6, (kt['spread'] - kt['gather']) * 1e6, ts.conv * 1e6, kt2['interp'] * 1e6]; n += 1
acc /= n
print(f"mode {mode}: conv_adjoint {acc[0]:.1f} us = spread kernel {acc[1]:.1f} + gather/rest {acc[0]-acc[1]:.1f};  conv {acc[2]:.1f} us (interp kernel {acc[3]:.1f})")
