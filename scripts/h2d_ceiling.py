"""Host-transfer ceiling of the end-to-end path: every rank copies 32 MiB host->device and 32 MiB device->host per
"step" (the bytes bench.py's e2e leg moves per rank and step), concurrently on two streams from page-locked buffers,
with nothing else running.  Launch under torchrun; rank 0 prints one JSON line with the per-rank and aggregate rates.
  python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 scripts/h2d_ceiling.py"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import bench

rank, world, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
bound = bench.bind_to_gpu_numa(lr) if "--bind" in sys.argv else None
torch.cuda.set_device(lr)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{lr}"))
n = 32 * 2 ** 20
hsrc = torch.empty(n, dtype=torch.uint8).pin_memory(); hdst = torch.empty(n, dtype=torch.uint8).pin_memory()
dsrc = torch.empty(n, dtype=torch.uint8, device="cuda"); ddst = torch.empty(n, dtype=torch.uint8, device="cuda")
hsrc.fill_(1)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def step():
    with torch.cuda.stream(s1): ddst.copy_(hsrc, non_blocking=True)
    with torch.cuda.stream(s2): hdst.copy_(dsrc, non_blocking=True)
def sync():
    torch.cuda.synchronize()
    if world > 1: dist.barrier()
for _ in range(5): step()
sync()
reps = 50
t0 = time.perf_counter()
for _ in range(reps): step()
torch.cuda.synchronize()
t = (time.perf_counter() - t0) / reps
tt = torch.tensor([t], device="cuda", dtype=torch.float64)
if world > 1: dist.all_reduce(tt, op=dist.ReduceOp.MAX)
if rank == 0:
    t = tt.item()
    print(json.dumps({"n_gpus": world, "numa_bound_cores": bound, "ms_per_step_copy_only": t * 1e3, "per_rank_GBs_each_direction": n / t / 1e9,
                      "aggregate_GBs_both_directions": 2 * n * world / t / 1e9,
                      "note": "32 MiB H2D + 32 MiB D2H per rank and step, concurrent, page-locked, max over ranks"}), flush=True)
if world > 1: dist.destroy_process_group()
