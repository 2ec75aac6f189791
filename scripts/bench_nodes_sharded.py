"""Node-sharded throughput at P GPUs: adjoint = local spread -> ncclReduceScatter -> slab FFT -> crop;
forward = slab FFT -> ncclAllGather -> local interpolation.  Launch with torchrun; prints one JSON line on rank 0.
  python -m torch.distributed.run --nproc-per-node P --master-addr 127.0.0.1 scripts/bench_nodes_sharded.py C5s"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import nfft_jl_b200 as nb
from oracle import nfft_oracle as O

CFG = {"C4s": ((2 ** 21,), 2 ** 24, 4, np.float64), "C5s": ((256, 256, 256), 2 ** 25, 3, np.float32),
       "C2": ((128, 128, 128), 2 ** 21, 3, np.float32)}
name = sys.argv[1] if len(sys.argv) > 1 else "C5s"
kmode = int(sys.argv[2]) if len(sys.argv) > 2 else 0      # 6 = NCCL reduce-scatter baseline instead of the fused peer gather; 7 = register-footprint kernels on the fused paths
N, M, m, T = CFG[name]
rank, world, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{lr}"))
real_stdout = os.dup(1); os.dup2(2, 1)
k = O.random_nodes(M, len(N), T, seed=1)
kd = torch.from_numpy(np.ascontiguousarray(k.T)).cuda()
p = nb.plan_nfft(kd, N, m=m, σ=2.0, shard="nodes" if world > 1 else None)
p.set_kernel_mode(kmode)
f = p.empty_image(); fh = p.empty_out(); fo = p.empty_image(); fho = p.empty_out()
f.copy_(torch.randn(f.shape, device="cuda", dtype=torch.float32 if T == np.float32 else torch.float64))
fh.copy_(torch.randn(fh.shape, device="cuda", dtype=torch.float32 if T == np.float32 else torch.float64))
def sync():
    if world > 1: dist.barrier()
    torch.cuda.synchronize()
for _ in range(3):
    nb.mul_(fho, p, f); nb.mul_(fo, p.adjoint(), fh)
sync()
reps = 5
e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
tf = ta = 0.0
for _ in range(reps):
    sync(); e[0].record(); nb.mul_(fho, p, f); e[1].record(); nb.mul_(fo, p.adjoint(), fh); e[2].record(); e[2].synchronize()
    tf += e[0].elapsed_time(e[1]) / reps; ta += e[1].elapsed_time(e[2]) / reps
t = torch.tensor([tf, ta], device="cuda", dtype=torch.float64)
if world > 1: dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    os.dup2(real_stdout, 1)
    print(json.dumps({"config": name, "N": N, "M": M, "n_gpus": world, "sharding": "nodes" if world > 1 else "none", "kernel_mode": kmode,
                      "fused_peer_spread": bool(world > 1 and p.fused_peer_spread and kmode != 6),
                      "fused_peer_interp": bool(world > 1 and p.fused_peer_interp and kmode != 6),
                      "forward_ms": t[0].item(), "adjoint_ms": t[1].item(),
                      "forward_pts_per_s": M / t[0].item() * 1e3, "adjoint_pts_per_s": M / t[1].item() * 1e3}), flush=True)
if world > 1: dist.destroy_process_group()
