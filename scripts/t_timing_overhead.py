import sys, os
sys.path.insert(0, "/root/repo")
import numpy as np, torch
import nfft_jl_b200 as nb
from oracle import nfft_oracle as O
k = O.random_nodes(2**21, 3, np.float32, seed=1)
p = nb.plan_nfft(torch.from_numpy(np.ascontiguousarray(k.T)).cuda(), (128,128,128), m=3, σ=2.0)
f = p.empty_image(); fh = p.empty_out(); fo = p.empty_image(); fho = p.empty_out(); f.fill_(1.0); fh.fill_(1.0)
flush = torch.empty(256*2**20, dtype=torch.uint8, device="cuda")
ts = nb.TimingStats()
for mode in ("timing on", "timing off", "timing on", "timing off"):
    p.enable_timing(mode == "timing on")
    for _ in range(5): nb.mul_(fho, p, f); nb.mul_(fo, p.adjoint(), fh)
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(20):
        flush.zero_()
        s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
        s.record()
        if mode == "timing on":
            nb.mul_(fho, p, f, timing=ts); nb.mul_(fo, p.adjoint(), fh, timing=ts)
        else:
            nb.mul_(fho, p, f); nb.mul_(fo, p.adjoint(), fh)
        e.record(); e.synchronize(); tot += s.elapsed_time(e)
    print(mode, "%.1f us/step" % (tot / 20 * 1e3), flush=True)
