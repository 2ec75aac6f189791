"""Stage timings (library CUDA events) of a named BASELINE config on one GPU.  usage: time_cfg.py C3 [mode ...]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench, nfft_jl_b200 as nb
cfg = bench.CONFIGS[sys.argv[1]]
modes = [int(a) for a in sys.argv[2:]] or [0]
kd = bench.device_nodes(cfg, torch)
kw = dict(m=cfg["m"], σ=2.0)
if cfg["B"] > 1: kw["ntransforms"] = cfg["B"]
if os.environ.get("BS"): kw["blockSize"] = tuple(int(v) for v in os.environ["BS"].split(","))
p = nb.plan_nfft(kd, cfg["N"], **kw)
f = p.empty_image(); fh = p.empty_out(); fo = p.empty_image(); fho = p.empty_out()
f.fill_(1.0); fh.fill_(1.0)
for mode in modes:
    p.set_kernel_mode(mode)
    ts = nb.TimingStats(); acc = np.zeros(4); n = 0
    for i in range(9):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True); e2 = torch.cuda.Event(enable_timing=True)
        e0.record(); nb.mul_(fho, p, f, timing=ts); e1.record(); tc = ts.conv
        nb.mul_(fo, p.adjoint(), fh, timing=ts); e2.record(); torch.cuda.synchronize()
        if i >= 3:
            acc += [e0.elapsed_time(e1) * 1e3, tc * 1e6, e1.elapsed_time(e2) * 1e3, ts.conv_adjoint * 1e6]; n += 1
    acc /= n
    print(f"{sys.argv[1]} tiles {tuple(p.params.blockSize)} mode {mode}: forward {acc[0]:.0f} us (conv {acc[1]:.0f}), adjoint {acc[2]:.0f} us (conv {acc[3]:.0f})", flush=True)
