"""C4 (1-D N=2^22, M=2^25, m=4, Float64) with the nodes given in random order and in ascending order: with sorted nodes
the caller order equals the plan order, so the fHat gather / scatter through the permutation is coalesced -- what the
1-D kernels cost without the random 16-byte accesses the reference contract (fHat[j] in caller order) implies."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import nfft_jl_b200 as nb
g = torch.Generator(device="cuda"); g.manual_seed(1)
k = torch.rand((1, 2 ** 25), generator=g, device="cuda", dtype=torch.float64) - 0.5
for name, kk in (("random", k), ("sorted", torch.sort(k, dim=1).values.contiguous())):
    p = nb.plan_nfft(kk, (2 ** 22,), m=4, σ=2.0)
    f = p.empty_image(); fh = p.empty_out(); fo = p.empty_image(); fho = p.empty_out(); f.fill_(1.0); fh.fill_(1.0)
    ts = nb.TimingStats(); acc = np.zeros(2); n = 0
    for i in range(8):
        nb.mul_(fho, p, f, timing=ts); c = ts.conv
        nb.mul_(fo, p.adjoint(), fh, timing=ts)
        if i >= 3: acc += [c * 1e6, ts.conv_adjoint * 1e6]; n += 1
    print("C4 nodes %s: interp %.0f us, spread %.0f us" % (name, acc[0] / n, acc[1] / n), flush=True)
    del p
