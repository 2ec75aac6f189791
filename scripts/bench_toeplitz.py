"""Toeplitz (Gram) operator on the C3 geometry (2-D 512x512, radial 1024x1024 nodes, 32 coil images, Float32):
kernel construction, batched apply, the NFFT-based Gram product it replaces, and the CPU restatement (scipy.fft on
all cores) timed beside it.  Prints one JSON line."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import nfft_jl_b200 as nb
from oracle import nfft_oracle as O

N, B, T = (512, 512), 32, np.float32
k = O.radial_nodes(1024, 1024, T)
M = k.shape[0]
kd = torch.from_numpy(np.ascontiguousarray(k.T)).cuda()


def dev_time(fn, reps=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); e1.synchronize()
    return e0.elapsed_time(e1) / reps * 1e-3


t0 = time.perf_counter()
lam = nb.calculateToeplitzKernel(N, kd, m=4, σ=2.0)
torch.cuda.synchronize()
t_kernel_first = time.perf_counter() - t0
p2 = nb.plan_nfft(kd, tuple(2 * n for n in N), m=4, σ=2.0)
lam2 = p2.empty_image()
L = nb.lib()
import ctypes as C
t_kernel = dev_time(lambda: L.nfftb200_toeplitz_kernel(p2._h, C.c_void_p(lam2.data_ptr()), 1), reps=5, warm=2)
op = nb.ToeplitzOperator(lam, ntransforms=B)
y = torch.empty((B,) + N[::-1], dtype=torch.complex64, device="cuda").permute(2, 1, 0)
y.copy_(torch.randn(y.shape, dtype=torch.complex64, device="cuda"))
y0 = y.clone()
t_apply = dev_time(lambda: op.apply_(y))
# the NFFT pair the operator replaces: adjoint(p) * (p * x) on the same 32 images
p = nb.plan_nfft(kd, N, m=4, σ=2.0, ntransforms=B)
x = p.empty_image(); x.copy_(y0)
fh = p.empty_out(); xo = p.empty_image()
t_gram = dev_time(lambda: (nb.mul_(fh, p, x), nb.mul_(xo, p.adjoint(), fh)))
# parity on one image: Toeplitz apply == NFFT Gram product (reference tolerance 1e-5 * conditioning of m=4)
y.copy_(y0); op.apply_(y); nb.mul_(fh, p, x); nb.mul_(xo, p.adjoint(), fh)
a, b = y[..., 0].cpu().numpy().astype(np.complex128), xo[..., 0].cpu().numpy().astype(np.complex128)
err = float(np.linalg.norm(a - b) / np.linalg.norm(b))
# CPU restatement (oracle, scipy.fft with all workers) on a bounded sample: 4 images
lam_h = lam.cpu().numpy()
xs = [np.asfortranarray(y0[..., i].cpu().numpy()) for i in range(4)]
from scipy import fft as sfft
ncpu = os.cpu_count() or 1
t0 = time.perf_counter()
with sfft.set_workers(ncpu):
    for xi in xs:
        O.convolve_toeplitz_kernel(xi, lam_h)
t_cpu = (time.perf_counter() - t0) / 4
obytes = 8 * 4 * N[0] * N[1]           # one oversampled complex64 array
print(json.dumps({
    "config": "C3 geometry: 2-D 512x512, 1024x1024 radial nodes, m=4, sigma=2, Float32, 32 images",
    "toeplitz_kernel_ms": t_kernel * 1e3, "toeplitz_kernel_first_call_ms_incl_plan": t_kernel_first * 1e3,
    "toeplitz_apply_ms_32_images": t_apply * 1e3, "toeplitz_apply_images_per_s": B / t_apply,
    "nfft_gram_ms_32_images": t_gram * 1e3, "speedup_vs_nfft_gram": t_gram / t_apply,
    "apply_algorithmic_bytes": 2 * B * 8 * N[0] * N[1] + obytes, "apply_hbm_gbs_algorithmic": (2 * B * 8 * N[0] * N[1] + obytes) / t_apply / 1e9,
    "apply_vs_nfft_gram_rel_l2": err,
    "cpu_port_ms_per_image": t_cpu * 1e3, "cpu_cores": ncpu, "gpu_ms_per_image": t_apply / B * 1e3}))
