"""Quick timing of the spread / interp kernels of a kernel mode on C2 (and optionally the C5 density): prints one line.
usage: python scripts/time8.py [mode] [cfg ...]   cfg in C2, C5d (128^3 M=2^24)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import nfft_jl_b200 as nb
from oracle import nfft_oracle as O
mode = int(sys.argv[1]) if len(sys.argv) > 1 else 8
cfgs = sys.argv[2:] or ["C2"]
CFG = {"C2": ((128, 128, 128), 2 ** 21), "C5d": ((128, 128, 128), 2 ** 24), "C5": ((256, 256, 256), 2 ** 27)}
for c in cfgs:
    N, M = CFG[c]
    k = O.random_nodes(M, 3, np.float32, seed=1)
    p = nb.plan_nfft(torch.from_numpy(np.ascontiguousarray(k.T)).cuda(), N, m=3, σ=2.0)
    p.set_kernel_mode(mode)
    fh = p.empty_out(); fh.fill_(1.0); fo = p.empty_image(); f = p.empty_image(); f.fill_(1.0); fho = p.empty_out()
    ts = nb.TimingStats(); acc = np.zeros(4); n = 0
    for i in range(13):
        nb.mul_(fo, p.adjoint(), fh, timing=ts); kt = p.kernel_times()
        nb.mul_(fho, p, f, timing=ts)
        if i >= 3:
            acc += [ts.conv_adjoint * 1e6, (kt["spread"] - kt["gather"]) * 1e6, kt["gather"] * 1e6, ts.conv * 1e6]; n += 1
    acc /= n
    print(f"{os.environ.get('NFFTB200_LIB','default')} mode {mode} {c}: conv_adjoint {acc[0]:.1f} us (spread kernel {acc[1]:.1f}, gather {acc[2]:.1f}), conv {acc[3]:.1f} us", flush=True)
    del p
