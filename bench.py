#!/usr/bin/env python
"""bench.py -- NFFT + adjoint nonuniform points/s on B200 (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A *step* is one forward NFFT (mul!(fHat,p,f)) plus one adjoint NFFT (mul!(f,adjoint(p),fHat)) over the
workload's M nodes.  Workload at every N: BASELINE.json configs[1] -- 3-D N=(128,128,128), M=2^21 uniform
random nodes, m=3, sigma=2, Float32, Kaiser-Bessel, POLYNOMIAL window (the reference default).  With N GPUs
the job is ntransforms=N batched transforms sharing the nodes, one transform per rank (batch sharding,
SURVEY 8e-a: no collective on the data path), so per-GPU work is fixed ("weak" scaling) and
value = N * 2*M / t_step.

`--impl reference` times the CPU restatement of the reference's default blocked algorithm
(oracle/cpu_ref.py: C/OpenMP port + pocketfft; NFFT.jl itself cannot run here: no julia binary).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

# torchrun exports OMP_NUM_THREADS=1 to every rank.  pocketfft (scipy.fft) sizes its worker pool from that variable when
# it is first imported, which made the reference arm's two FFTs 5x slower under torchrun than standalone (round-1
# verdict: 181 -> 509 ms/step).  The reference arm is meant to use every host core, so undo it before numpy/scipy load.
if "reference" in sys.argv:
    try:
        os.environ["OMP_NUM_THREADS"] = str(len(os.sched_getaffinity(0)))
    except Exception:
        os.environ.pop("OMP_NUM_THREADS", None)

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOAD = dict(name="C2: 3D NFFT N=(128,128,128), M=2^21 random nodes, m=3, sigma=2, Float32",
                N=(128, 128, 128), M=2 ** 21, m=3, sigma=2.0, T=np.float32)
METRIC = "NFFT+adjoint nonuniform pts/sec"
UNIT = "pts/s"


def random_nodes(M, D, T, seed=1):
    """SURVEY 8(d) synthetic nodes: uniform in [-1/2, 1/2), PCG64, cast once to T; shape (M, D)"""
    rng = np.random.default_rng(seed)
    return (rng.random((M, D)) - 0.5).astype(T)


def random_complex(shape, T, seed=2):
    rng = np.random.default_rng(seed)
    cT = np.complex64 if T == np.float32 else np.complex128
    a = (rng.standard_normal(shape) + 1j * rng.standard_normal(shape)).astype(cT)
    return np.asfortranarray(a) if a.ndim > 1 else a


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def algorithmic_bytes(D, s, M, gsz, B=1):
    """coords + int32 permutation + node values + grid, per spread or interp launch"""
    return D * s * M + 4 * M + 2 * s * M * B + 2 * s * gsz * B


class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for n, v in zip(names, r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def bind_to_gpu_numa(local_rank):
    """pin this rank to the CPU cores next to its GPU (NVML's affinity mask) BEFORE any page-locked buffer is allocated, so
    the H2D/D2H staging buffers of the end-to-end path are first-touched on the GPU's own NUMA node; with 8 ranks on one
    box the round-1 run reached only ~17 GB/s per GPU through remote-node buffers"""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = {i * 64 + b for i, w in enumerate(mask) for b in range(64) if (int(w) >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return None


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def config_dict(world, kernel_mode=0):
    """the `config` both arms print (identical keys and values, so the driver can tell they ran the same thing)"""
    w = WORKLOAD
    return {"workload": w["name"], "ntransforms": world, "sharding": "batch (one transform per GPU)" if world > 1 else "none",
            "precompute": "POLYNOMIAL", "window": "kaiser_bessel", "m": w["m"], "sigma": w["sigma"], "blockSize": [16, 16, 16],
            "nodes": "uniform random, PCG64 seed 1", "step": "1 forward + 1 adjoint NFFT; value = n_gpus*2*M/t_step"}


def run_reference(args, rank, world, emit):
    if rank != 0:
        return
    from oracle.cpu_ref import CpuRefPlan, lib      # the reference arm is the one place that executes oracle/
    w = WORKLOAD
    lib().ref_set_num_threads(host_cores())         # torchrun sets OMP_NUM_THREADS=1: use every core we may run on
    cores = lib().ref_num_threads()
    k = random_nodes(w["M"], 3, w["T"], seed=1)
    t0 = time.perf_counter()
    p = CpuRefPlan(k, w["N"], m=w["m"], sigma=w["sigma"], workers=cores)
    t_plan = time.perf_counter() - t0
    f = random_complex(w["N"], w["T"], 2)
    fh = random_complex(w["M"], w["T"], 3)
    for _ in range(args.warmup):
        p.forward(f); p.adjoint(fh)
    times = []
    for _ in range(args.steps):
        t0 = time.perf_counter()
        p.forward(f); p.adjoint(fh)
        times.append(time.perf_counter() - t0)
    t = float(np.mean(times))
    val = 2 * w["M"] / t
    cb = {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
          "sample": f"full workload, {args.steps} steps of forward+adjoint (M=2^21 each); plan {t_plan:.2f}s excluded"}
    emit({
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_dict(args.gpus),
        "note": "reference algorithm restated in C/OpenMP + pocketfft (Julia unavailable); every step is the full workload on the host",
        "cpu_baseline": cb,
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})


# ---- the other BASELINE configs (timed after the headline; parity for all of them is in tests/test_gpu_fullsize.py) ----
CONFIGS = {
    "C1": dict(desc="2D N=(256,256), M=65536 random, m=4, Float64", N=(256, 256), M=65536, m=4, T=np.float64, B=1, nodes="random"),
    "C3": dict(desc="2D radial 1024 spokes x 1024 samples, N=(512,512), m=4, Float32, ntransforms=32, density-weighted adjoint",
               N=(512, 512), M=2 ** 20, m=4, T=np.float32, B=32, nodes="radial"),
    "C4": dict(desc="1D N=2^22, M=2^25 random, m=4, Float64", N=(2 ** 22,), M=2 ** 25, m=4, T=np.float64, B=1, nodes="random"),
    "C5": dict(desc="3D N=(256,256,256), M=2^27 random, m=3, Float32", N=(256, 256, 256), M=2 ** 27, m=3, T=np.float32, B=1, nodes="random"),
}


def radial_nodes(nspokes, nsamples, T):
    """spoke s at angle pi*s/nspokes, sample i at radius (i - n/2)/n"""
    th = np.pi * np.arange(nspokes) / nspokes
    r = (np.arange(nsamples) - nsamples // 2) / nsamples
    k = np.stack([np.outer(np.cos(th), r).ravel(), np.outer(np.sin(th), r).ravel()], axis=1)
    return np.clip(k, -0.5, 0.5).astype(T)


def device_nodes(cfg, torch):
    """D x M nodes on the device: radial trajectory, or uniform random from a seeded device generator (identical on every rank)"""
    tt = torch.float32 if cfg["T"] == np.float32 else torch.float64
    if cfg["nodes"] == "radial":
        return torch.from_numpy(np.ascontiguousarray(radial_nodes(1024, 1024, cfg["T"]).T)).cuda()
    g = torch.Generator(device="cuda")
    g.manual_seed(1)
    return torch.rand((len(cfg["N"]), cfg["M"]), generator=g, device="cuda", dtype=tt) - 0.5


def time_config(nb, torch, dist, cfg, world=1, shard=None, kernel_mode=0, reps=5, check=False):
    """forward / adjoint device times (CUDA events, max over ranks) and TimingStats phases of one config"""
    N, M, B, T = cfg["N"], cfg["M"], cfg["B"], cfg["T"]
    tt = torch.float32 if T == np.float32 else torch.float64
    kd = device_nodes(cfg, torch)
    kw = dict(m=cfg["m"], σ=2.0)
    if B > 1:
        kw["ntransforms"] = B
    if shard:
        kw["shard"] = shard
    p = nb.plan_nfft(kd, N, **kw)
    p.set_kernel_mode(kernel_mode)
    g = torch.Generator(device="cuda")
    g.manual_seed(7)
    f = p.empty_image(); fh = p.empty_out(); fo = p.empty_image(); fho = p.empty_out()
    ct = torch.complex64 if T == np.float32 else torch.complex128
    f.copy_(torch.randn(f.shape, generator=g, device="cuda", dtype=ct))
    fh.copy_(torch.randn(fh.shape, generator=g, device="cuda", dtype=ct))
    if cfg["nodes"] == "radial":            # density-compensated adjoint: ramp weights on the data
        r = torch.sqrt((kd.to(torch.float64) ** 2).sum(0)).clamp_min(1.0 / 2048)
        w = (r / r.mean()).to(tt)
        fh.mul_(w.reshape(-1, *([1] * (fh.dim() - 1))))

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    ts = nb.TimingStats()
    p.enable_timing(False)
    for _ in range(2):
        nb.mul_(fho, p, f); nb.mul_(fo, p.adjoint(), fh)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    tf = ta = 0.0
    ph = {n: 0.0 for n in ("deconv", "fft", "conv", "conv_adjoint", "fft_adjoint", "deconv_adjoint")}
    for _ in range(reps):                   # totals: no per-phase events (see the headline pass)
        sync()
        ev[0].record(); nb.mul_(fho, p, f); ev[1].record(); nb.mul_(fo, p.adjoint(), fh); ev[2].record()
        ev[2].synchronize()
        tf += ev[0].elapsed_time(ev[1]) / reps; ta += ev[1].elapsed_time(ev[2]) / reps
    for _ in range(reps):                   # phases: the library's events
        sync()
        nb.mul_(fho, p, f, timing=ts); nb.mul_(fo, p.adjoint(), fh, timing=ts)
        for n in ph:
            ph[n] += getattr(ts, n) / reps
    p.enable_timing(False)
    t = torch.tensor([tf, ta] + [ph[n] * 1e3 for n in ph], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    t = t.tolist()
    D, s = len(N), (4 if T == np.float32 else 8)
    gsz = int(np.prod(p.Ñ))
    ab = algorithmic_bytes(D, s, M, gsz, B)
    peak = peaks()[0]
    out = {"forward_ms": t[0], "adjoint_ms": t[1], "forward_pts_per_s": M * B / t[0] * 1e3, "adjoint_pts_per_s": M * B / t[1] * 1e3,
           "phases_us": {n: t[2 + i] * 1e3 for i, n in enumerate(ph)}, "kernel_mode": kernel_mode}
    if world == 1:      # one GPU: the conv phases are the interpolation / spread(+gather) kernels alone
        out["interp_hbm_frac"] = ab / (out["phases_us"]["conv"] * 1e-6) / 1e9 / peak
        out["spread_hbm_frac"] = ab / (out["phases_us"]["conv_adjoint"] * 1e-6) / 1e9 / peak
        out["algorithmic_bytes_per_launch"] = ab
    else:
        out["fused_peer_spread"] = bool(p.fused_peer_spread and kernel_mode != 6)
        out["fused_peer_interp"] = bool(p.fused_peer_interp and kernel_mode != 6)
    if check:           # adjointness <A f, y> == <f, A^H y>: size independent, involves every stage and every exchange
        y, Af, fc, AHy = fh, fho, f, fo
        lhs = torch.sum(torch.conj(y).to(torch.complex128) * Af.to(torch.complex128))      # other ranks' entries of A f are zero
        lhs = torch.view_as_real(lhs.reshape(1)).clone()
        if world > 1:
            dist.all_reduce(lhs)
        lhs = torch.view_as_complex(lhs)[0]
        rhs = torch.sum(torch.conj(AHy).to(torch.complex128) * fc.to(torch.complex128))
        out["adjointness_rel_err"] = float(abs(lhs - rhs) / abs(lhs))
    del p
    torch.cuda.empty_cache()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--kernel-mode", type=int, default=0)
    ap.add_argument("--headline-only", action="store_true", help="skip the other-config / node-sharded blocks")
    ap.add_argument("--block-size", type=str, default="", help="override the plan tile, e.g. 16,16,16 (experiments)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    # library chatter (e.g. "NCCL version ...") must not reach stdout: the only stdout line is the JSON result
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(obj):
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        print(json.dumps(obj), flush=True)

    if args.impl == "reference":
        run_reference(args, rank, world, emit)         # a step (forward + adjoint of the full C2 workload) is ~0.2 s of CPU
        return

    numa_cores = bind_to_gpu_numa(local_rank) if world > 1 else None
    import torch
    import torch.distributed as dist
    import nfft_jl_b200 as nb

    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))
    w = WORKLOAD
    N, M, T = w["N"], w["M"], w["T"]
    k = random_nodes(M, 3, T, seed=1)                       # same nodes on every rank (shared by the batch)
    kd = torch.from_numpy(np.ascontiguousarray(k.T)).cuda()
    bsz = tuple(int(x) for x in args.block_size.split(",")) if args.block_size else None
    if world > 1:      # batched plan with ntransforms = world, one transform per rank, no data-path collective
        p = nb.plan_nfft(kd, N, m=w["m"], σ=w["sigma"], ntransforms=world, shard="batch", blockSize=bsz)
        assert p.ntransforms == 1
    else:
        p = nb.plan_nfft(kd, N, m=w["m"], σ=w["sigma"], blockSize=bsz)
    p.set_kernel_mode(args.kernel_mode)
    f_h = random_complex(N, T, 100 + rank)
    fh_h = random_complex(M, T, 200 + rank)
    f = p.empty_image(); f.copy_(torch.from_numpy(np.ascontiguousarray(f_h.T)).cuda().permute(2, 1, 0))
    fh = p.empty_out(); fh.copy_(torch.from_numpy(fh_h).cuda())
    f_out = p.empty_image()
    fh_out = p.empty_out()
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")   # > 126 MB L2

    def step():
        nb.mul_(fh_out, p, f)
        nb.mul_(f_out, p.adjoint(), fh)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # Two passes of K steps.  The headline pass runs the plan as a user runs it (no per-phase events): with the library's
    # timing on, every mul! ends with a host-side wait for its phase events before the next call can be queued, which
    # measured 7.6 % on this step (1295 vs 1203 us, scripts/t_timing_overhead.py).  The instrumented pass that follows
    # collects the phase and kernel durations (TimingStats, roofline) with those events enabled.
    p.enable_timing(False)
    sampler = ClockSampler(local_rank)     # samples SM clocks / throttle reasons from warm-up to the end of the e2e loop
    sampler.start()
    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    l0 = p.launch_count()
    barrier()
    for s, e in ev:
        flush.zero_()                                       # L2 flush between timed iterations (untimed)
        s.record()
        step()
        e.record()
        e.synchronize()
    barrier()
    launches = p.launch_count() - l0
    t_step = sum(s.elapsed_time(e) for s, e in ev) / args.steps * 1e-3
    tt = torch.tensor([t_step], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    t_step = float(tt.item())
    value = world * 2 * M / t_step

    # instrumented pass: same K steps, same flush, the library's per-phase CUDA events on the plan's stream
    p.enable_timing(True)
    ts = nb.TimingStats()
    phases = {n: 0.0 for n in ("deconv", "fft", "conv", "conv_adjoint", "fft_adjoint", "deconv_adjoint")}
    kt = {"spread": 0.0, "interp": 0.0, "memset": 0.0, "gather": 0.0}
    t_instr = 0.0
    for s, e in ev:
        flush.zero_()
        s.record()
        nb.mul_(fh_out, p, f, timing=ts)
        nb.mul_(f_out, p.adjoint(), fh, timing=ts)
        e.record()
        e.synchronize()
        t_instr += s.elapsed_time(e) * 1e-3 / args.steps
        for n in phases:
            phases[n] += getattr(ts, n)
        for n, v in p.kernel_times().items():
            kt[n] += v
    p.enable_timing(False)
    barrier()

    # ---- end-to-end through the public API with pinned HOST buffers (H2D + D2H inside the timed region)
    pin = lambda a: torch.from_numpy(a).pin_memory().numpy()
    f_p = np.asfortranarray(pin(np.ascontiguousarray(f_h.T)).T)
    fh_p = pin(fh_h)
    fo_p = np.asfortranarray(pin(np.empty(N[::-1], dtype=p.cT)).T)
    fho_p = pin(np.empty(M, dtype=p.cT))
    # (a) synchronous calls: every mul! returns with its result on the host
    for _ in range(2):
        nb.mul_(fho_p, p, f_p); nb.mul_(fo_p, p.adjoint(), fh_p)
    ref_fwd, ref_adj = fho_p.copy(), fo_p.copy()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        nb.mul_(fho_p, p, f_p)
        nb.mul_(fo_p, p.adjoint(), fh_p)
    barrier()
    t_e2e_sync = (time.perf_counter() - t0) / args.steps
    # (b) the asynchronous host API (NFFTB200_HOST_ASYNC): the same 2*steps calls with the same per-step uploads
    # and downloads, queued back to back so that the copies of one call overlap the kernels of its neighbours;
    # the region ends when the last result has landed in host memory (p.sync())
    fho_p[...] = 0; fo_p[...] = 0
    for _ in range(2):
        nb.mul_(fho_p, p, f_p, async_host=True); nb.mul_(fo_p, p.adjoint(), fh_p, async_host=True)
    p.sync()
    assert np.array_equal(fho_p, ref_fwd) and np.array_equal(fo_p, ref_adj), "async host path differs from the synchronous one"
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        nb.mul_(fho_p, p, f_p, async_host=True)
        nb.mul_(fo_p, p.adjoint(), fh_p, async_host=True)
    p.sync()
    barrier()
    t_e2e = (time.perf_counter() - t0) / args.steps
    clocks = sampler.stop()
    te = torch.tensor([t_e2e], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    t_e2e = float(te.item())
    csz = 8
    e2e = {"value": world * 2 * M / t_e2e, "unit": UNIT, "h2d_bytes_per_step": int((np.prod(N) + M) * csz),
           "d2h_bytes_per_step": int((np.prod(N) + M) * csz), "ms_per_step": t_e2e * 1e3,
           "mode": "asynchronous host API (NFFTB200_HOST_ASYNC): uploads/downloads of neighbouring calls overlap the kernels",
           "sync_calls_value": world * 2 * M / t_e2e_sync, "sync_calls_ms_per_step": t_e2e_sync * 1e3,
           "numa_bound_cores_per_rank": numa_cores}

    gsz = int(np.prod(p.Ñ))
    del p, f, fh, f_out, fh_out, flush
    torch.cuda.empty_cache()
    extra = {}
    if not args.headline_only:
        if world == 1:
            # the other BASELINE configs on one GPU (radial C3 included), same kernels as the headline
            oc = {}
            for name in ("C1", "C3", "C4", "C5"):
                cfg = CONFIGS[name]
                r = time_config(nb, torch, dist, cfg, reps=5, check=True)
                r["config"] = cfg["desc"]
                oc[name] = r
            extra["other_configs"] = oc
        else:
            # node sharding: the tile-sorted node list is cut into `world` ranges; fused = spread + slab gather
            # and slab-direct interpolation over CUDA-IPC peer memory, kernel_mode 6 = NCCL reduce-scatter / all-gather baseline
            ns = {"n_gpus": world, "scaling": "strong", "note": "phases: conv_adjoint = spread + gather/reduce-scatter, fft_adjoint = slab FFT incl. "
                  "all-to-all, deconv_adjoint = crop + image all-reduce; forward mirrored (conv = grid exchange + interpolation)"}
            for name in ("C5", "C4"):
                cfg = CONFIGS[name]
                ns[name] = {"config": cfg["desc"],
                            "fused": time_config(nb, torch, dist, cfg, world, "nodes", 0, reps=4, check=True),
                            "nccl": time_config(nb, torch, dist, cfg, world, "nodes", 6, reps=4)}
            extra["node_sharded"] = ns

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_src = peaks()
    abytes = algorithmic_bytes(3, 4, M, gsz)
    t_spread = kt["spread"] / args.steps
    t_interp = kt["interp"] / args.steps
    traffic, traffic_src = None, None
    tkey = {0: "lean", 8: "lean", 7: "bin", 9: "sub"}.get(args.kernel_mode)
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as fh_:
            tj = json.load(fh_)
        traffic = tj["spread_stage_dram_bytes_per_launch"].get(tkey)
        traffic_src = tj["source"]
    except Exception:
        pass
    spread_kernel = {"lean": "k_spread_lean<3,8>", "bin": "k_spread_bin3d<float,3,8>"}.get(tkey, "k_spread_sub3d<float,3>")
    interp_kernel = {"lean": "k_interp_lean<3,8>", "bin": "k_interp_bin3d<float,3,8>"}.get(tkey, "k_interp_row3d<float,3>")
    roofline = {"bound": "hbm", "kernel": f"adjoint gridding: {spread_kernel} + k_gather_cols3d (convolve_transpose!)",
                "achieved": abytes / t_spread / 1e9, "peak": peak, "peak_source": peak_src, "unit": "GB/s",
                "frac": abytes / t_spread / 1e9 / peak, "traffic": traffic, "traffic_source": traffic_src,
                "algorithmic_bytes_per_launch": abytes, "us_per_launch": t_spread * 1e6,
                "note": "3-D spreading is issue/shared-memory bound (216 complex FMAs per node): the FP32 floor of this launch is 24 us, "
                        "the HBM floor 28 us; the HBM fraction is the contract metric (DESIGN.md 3)",
                "interp": {"kernel": f"{interp_kernel} (convolve!)", "achieved": abytes / t_interp / 1e9,
                           "frac": abytes / t_interp / 1e9 / peak, "us_per_launch": t_interp * 1e6}}
    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": t_step * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_dict(world),
        "kernel_mode": args.kernel_mode, "l2": "flushed between timed iterations (256 MiB write, untimed)",
        "ms_per_step_instrumented": t_instr * 1e3,
        "timing_note": "value / ms_per_step: K steps without per-phase events; phases_us, roofline kernel durations and "
                       "ms_per_step_instrumented: a second pass of K steps with the library's phase events enabled",
        "forward_pts_per_s": M / sum(phases[n] for n in ("deconv", "fft", "conv")) * args.steps,
        "adjoint_pts_per_s": M / sum(phases[n] for n in ("conv_adjoint", "fft_adjoint", "deconv_adjoint")) * args.steps,
        "phases_us": {n: v / args.steps * 1e6 for n, v in phases.items()},
        "memset_us": kt["memset"] / args.steps * 1e6,
        "roofline": roofline, "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches),
    }
    out.update(extra)
    if world == 1 and not args.no_cpu_baseline:
        from oracle.cpu_ref import CpuRefPlan, lib
        lib().ref_set_num_threads(host_cores())
        cores = lib().ref_num_threads()
        pc = CpuRefPlan(k, N, m=w["m"], sigma=w["sigma"], workers=cores)
        pc.forward(f_h); pc.adjoint(fh_h)
        t0 = time.perf_counter()
        nrep = 5
        for _ in range(nrep):
            pc.forward(f_h); pc.adjoint(fh_h)
        tc = (time.perf_counter() - t0) / nrep
        out["cpu_baseline"] = {"value": 2 * M / tc, "unit": UNIT, "cores": cores, "kind": "port",
                               "sample": f"full workload, {nrep} steps of forward+adjoint after 1 warm-up",
                               "ms_per_step": tc * 1e3}
    if world > 1:
        dist.destroy_process_group()
    emit(out)


if __name__ == "__main__":
    main()
