"""Stage timings (TimingStats, CUDA events) of the BASELINE configs that fit one GPU; also checks each against
the oracle's NDFT-level accuracy on a node subsample."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import nfft_jl_b200 as nb
from oracle import nfft_oracle as O

CFG = {
    "C1": dict(N=(256, 256), M=65536, m=4, T=np.float64, B=1, nodes="random"),
    "C2": dict(N=(128, 128, 128), M=2 ** 21, m=3, T=np.float32, B=1, nodes="random"),
    "C3": dict(N=(512, 512), M=2 ** 20, m=4, T=np.float32, B=4, nodes="radial"),
    "C4s": dict(N=(2 ** 20,), M=2 ** 23, m=4, T=np.float64, B=1, nodes="random"),
    "C5s": dict(N=(256, 256, 256), M=2 ** 24, m=3, T=np.float32, B=1, nodes="random"),
    "C3full": dict(N=(512, 512), M=2 ** 20, m=4, T=np.float32, B=32, nodes="radial"),
    "C4": dict(N=(2 ** 22,), M=2 ** 25, m=4, T=np.float64, B=1, nodes="random"),
    "C5": dict(N=(256, 256, 256), M=2 ** 27, m=3, T=np.float32, B=1, nodes="random"),
}
names = sys.argv[1:] or ["C1", "C2", "C3", "C4s"]
for name in names:
    c = CFG[name]
    N, M, T, B = c["N"], c["M"], c["T"], c["B"]
    D = len(N)
    k = O.radial_nodes(1024, 1024, T) if c["nodes"] == "radial" else O.random_nodes(M, D, T, seed=1)
    kd = torch.from_numpy(np.ascontiguousarray(k.T)).cuda()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    p = nb.plan_nfft(kd, N, m=c["m"], σ=2.0, ntransforms=B)
    torch.cuda.synchronize(); t_plan = time.perf_counter() - t0
    t0 = time.perf_counter(); p.nodes_(kd); torch.cuda.synchronize(); t_nodes = time.perf_counter() - t0
    f = p.empty_image(); fh = p.empty_out()
    f.copy_(torch.randn(f.shape, device="cuda", dtype=torch.float32 if T == np.float32 else torch.float64))
    fh.copy_(torch.randn(fh.shape, device="cuda", dtype=torch.float32 if T == np.float32 else torch.float64))
    fo = p.empty_image(); fho = p.empty_out()
    ts = nb.TimingStats()
    acc = {}
    for it in range(6):
        nb.mul_(fho, p, f, timing=ts); nb.mul_(fo, p.adjoint(), fh, timing=ts)
        if it >= 1:
            for n in ("deconv", "fft", "conv", "conv_adjoint", "fft_adjoint", "deconv_adjoint"):
                acc[n] = acc.get(n, 0) + getattr(ts, n) / 5
    tf = acc["deconv"] + acc["fft"] + acc["conv"]; ta = acc["conv_adjoint"] + acc["fft_adjoint"] + acc["deconv_adjoint"]
    print(f"{name}: plan {t_plan*1e3:.1f} ms, nodes! {t_nodes*1e3:.2f} ms | " +
          " ".join(f"{n}={v*1e6:.0f}us" for n, v in acc.items()) +
          f" | fwd {M*B/tf/1e9:.2f} Gpts/s adj {M*B/ta/1e9:.2f} Gpts/s  bs={p.params.blockSize}")
    # accuracy on a subsample vs NDFT (first transform)
    sub = np.arange(0, k.shape[0], max(1, k.shape[0] // 200))[:200]
    f0 = f.cpu().numpy() if B == 1 else f[..., 0].cpu().numpy()
    if np.prod(N) <= 2 ** 19:
        ref = O.ndft(k[sub].astype(np.float64), f0)
        got = (fho.cpu().numpy() if B == 1 else fho[:, 0].cpu().numpy())[sub]
        print(f"   forward rel err vs NDFT on {len(sub)} nodes: {np.linalg.norm(got-ref)/np.linalg.norm(ref):.2e}")
