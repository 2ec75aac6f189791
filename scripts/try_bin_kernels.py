"""Opt-in register-footprint kernels (kernel_mode 7: csrc/spread_bin.cuh, csrc/interp_bin.cuh) on a real GPU: parity of
the adjoint and of the forward transform against the default tiled kernels (kernel_mode 0) and against the oracle,
then the spread-stage and interpolation times of both modes.

    python scripts/try_bin_kernels.py            # parity cases only (seconds)
    python scripts/try_bin_kernels.py --time     # + C2 / C5-density timings (CUDA events inside the library)

Prints one JSON line per case and appends them to gpurun_out/bin_kernels.jsonl.  Exit code 1 if a parity case fails."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import nfft_jl_b200 as nb
from oracle import nfft_oracle as O

TOL = {np.float32: 1e-5, np.float64: 1e-12}
out_lines = []
MODES = (7, 9)            # register-window kernels checked against mode 0: 7 = in-kernel bin sort, 9 = round-1 sub-tile kernels; mode 0 is the default (the lean kernels)
for _a in sys.argv[1:]:
    if _a.startswith("--modes="):
        MODES = tuple(int(x) for x in _a.split("=")[1].split(","))


def emit(d):
    print(json.dumps(d), flush=True)
    out_lines.append(d)


def rel(a, b):
    a = np.asarray(a).ravel().astype(np.complex128)
    b = np.asarray(b).ravel().astype(np.complex128)
    return float(np.linalg.norm(a - b) / np.linalg.norm(b))


def nodes(M, T, seed, cluster=0):
    k = O.random_nodes(M, 3, T, seed=seed)
    if cluster:
        k[:cluster] = (k[:cluster] * T(0.02)).astype(T)         # dense bins around the origin: rounds, chunks, split items
    return k


def parity_case(name, N, M, T, m, B=1, cluster=0, blockSize=None, oracle=True):
    k = nodes(M, T, 3, cluster)
    kw = dict(m=m, σ=2.0, ntransforms=B) if B > 1 else dict(m=m, σ=2.0)
    if blockSize is not None:
        kw["blockSize"] = blockSize
    p = nb.plan_nfft(torch.from_numpy(np.ascontiguousarray(k.T)).cuda(), N, **kw)
    shape = (M, B) if B > 1 else (M,)
    fh = O.random_complex(shape, T, 5)
    f = O.random_complex(tuple(N) + ((B,) if B > 1 else ()), T, 6)
    res, fwd = {}, {}
    launches = {}
    for mode in (0,) + MODES:
        p.set_kernel_mode(mode)
        l0 = p.launch_count()
        res[mode] = np.array(p.adjoint() * fh)
        launches[mode] = p.launch_count() - l0
        again = np.array(p.adjoint() * fh)
        assert np.array_equal(res[mode], again), f"adjoint, mode {mode}: not bit-reproducible"
        fwd[mode] = np.array(p * f)
        assert np.array_equal(fwd[mode], np.array(p * f)), f"forward, mode {mode}: not bit-reproducible"
    d = {"case": name, "N": list(N), "M": M, "dtype": np.dtype(T).name, "m": m, "B": B, "launches": launches}
    ok = True
    po = O.OraclePlan(k, N, m=m, sigma=2.0, blockSize=p.params.blockSize) if (oracle and B == 1) else None
    ref_a = po.adjoint(fh) if po else None
    ref_f = po.forward(f) if po else None
    for mode in MODES:
        e = rel(res[mode], res[0])
        ef = rel(fwd[mode], fwd[0])
        d[f"rel_mode{mode}_vs_mode0"] = e
        d[f"fwd_rel_mode{mode}_vs_mode0"] = ef
        ok = ok and e <= TOL[T] and ef <= TOL[T]
        if po:
            d[f"rel_mode{mode}_vs_oracle"] = rel(res[mode], ref_a)
            d[f"fwd_rel_mode{mode}_vs_oracle"] = rel(fwd[mode], ref_f)
            ok = ok and d[f"rel_mode{mode}_vs_oracle"] <= TOL[T] and d[f"fwd_rel_mode{mode}_vs_oracle"] <= TOL[T]
    d["ok"] = bool(ok)
    emit(d)
    return ok


def time_case(name, N, M, T, m):
    k = nodes(M, T, 1)
    p = nb.plan_nfft(torch.from_numpy(np.ascontiguousarray(k.T)).cuda(), N, m=m, σ=2.0)
    fh = p.empty_out(); fh.fill_(1.0)
    fo = p.empty_image()
    f = p.empty_image(); f.fill_(1.0)
    fho = p.empty_out()
    ts = nb.TimingStats()
    d = {"case": name, "N": list(N), "M": M, "dtype": np.dtype(T).name, "m": m}
    for mode in (0,) + MODES:
        p.set_kernel_mode(mode)
        acc = np.zeros(3); n = 0
        for i in range(13):
            nb.mul_(fo, p.adjoint(), fh, timing=ts)
            kt = p.kernel_times()
            nb.mul_(fho, p, f, timing=ts)
            if i >= 3:
                acc += [ts.conv_adjoint * 1e6, (kt["spread"] - kt["gather"]) * 1e6, ts.conv * 1e6]; n += 1
        acc /= n
        d[f"mode{mode}_conv_adjoint_us"] = round(float(acc[0]), 1)
        d[f"mode{mode}_spread_kernel_us"] = round(float(acc[1]), 1)
        d[f"mode{mode}_conv_us"] = round(float(acc[2]), 1)
    emit(d)


def main():
    ok = True
    f32, f64 = np.float32, np.float64
    ok &= parity_case("f32 m=3 uniform", (32, 32, 32), 20000, f32, 3)
    ok &= parity_case("f32 m=3 clustered", (48, 32, 40), 30000, f32, 3, cluster=6000)
    ok &= parity_case("f64 m=3 clustered", (32, 32, 32), 20000, f64, 3, cluster=3000)
    ok &= parity_case("f32 m=2", (32, 32, 32), 20000, f32, 2)
    ok &= parity_case("f32 m=4 (W=10)", (32, 32, 32), 20000, f32, 4)
    ok &= parity_case("f64 m=4 (falls back to the default kernel)", (32, 32, 32), 8000, f64, 4)
    ok &= parity_case("f32 m=3 B=3", (32, 32, 32), 20000, f32, 3, B=3)
    ok &= parity_case("f32 m=3 thin tiles", (32, 32, 32), 20000, f32, 3, blockSize=(16, 16, 8))
    ok &= parity_case("f32 m=3 64^3 (C2 density)", (64, 64, 64), 2 ** 18, f32, 3, oracle=False)
    if "--time" in sys.argv:
        time_case("C2 128^3 M=2^21", (128, 128, 128), 2 ** 21, f32, 3)
        time_case("C5 density 128^3 M=2^24", (128, 128, 128), 2 ** 24, f32, 3)
        time_case("C2 geometry Float64", (128, 128, 128), 2 ** 21, f64, 3)
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/bin_kernels.jsonl", "a") as fh:
        for d in out_lines:
            fh.write(json.dumps(d) + "\n")
    print("PARITY OK" if ok else "PARITY FAILED")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
