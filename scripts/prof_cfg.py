"""ncu driver: one forward + adjoint of a named BASELINE config (bench.CONFIGS) on one GPU.  usage: prof_cfg.py C3 [mode]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench, nfft_jl_b200 as nb
cfg = bench.CONFIGS[sys.argv[1]]
mode = int(sys.argv[2]) if len(sys.argv) > 2 else 0
kd = bench.device_nodes(cfg, torch)
kw = dict(m=cfg["m"], σ=2.0)
if cfg["B"] > 1: kw["ntransforms"] = cfg["B"]
p = nb.plan_nfft(kd, cfg["N"], **kw); p.set_kernel_mode(mode)
f = p.empty_image(); fh = p.empty_out(); fo = p.empty_image(); fho = p.empty_out()
f.fill_(1.0); fh.fill_(1.0)
for _ in range(2):
    nb.mul_(fho, p, f); nb.mul_(fo, p.adjoint(), fh)
torch.cuda.synchronize(); print("done", sys.argv[1])
