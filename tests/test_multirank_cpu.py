"""N>1 host logic on CPU: world_size-2 gloo processes exercise the pieces of the multi-GPU path that do not
need a GPU -- the ncclUniqueId-style broadcast plumbing, the tile-aligned node partition (C ABI
nfftb200_partition_tiles) and the batch split."""
import os
import subprocess
import sys
import textwrap

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = textwrap.dedent("""
    import os, sys
    sys.path.insert(0, %r)
    import numpy as np, torch, torch.distributed as dist
    import nfft_jl_b200 as nb
    from oracle import nfft_oracle as O
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    # 1. a 128-byte id created on rank 0 reaches every rank (the path the ncclUniqueId takes)
    t = torch.arange(128, dtype=torch.uint8) if rank == 0 else torch.zeros(128, dtype=torch.uint8)
    dist.broadcast(t, src=0)
    assert t.tolist() == list(range(128))
    # 2. tile-aligned node partition: identical on all ranks, covers all tiles, balanced
    N, T = (64, 64), np.float32
    k = O.random_nodes(50000, 2, T, seed=1)
    p = O.init_params(N, T, 4, 2.0)
    perm, counts, _ = O.precompute_blocks(k, p)
    ts = np.concatenate([[0], np.cumsum(counts)])
    cut = nb.partition_tiles(ts, world)
    allc = [torch.zeros(world + 1, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(allc, torch.from_numpy(cut))
    assert all(torch.equal(allc[0], c) for c in allc)
    assert cut[0] == 0 and cut[-1] == len(counts) and np.all(np.diff(cut) >= 0)
    mine = ts[cut[rank + 1]] - ts[cut[rank]]
    tot = torch.tensor([int(mine)]); dist.all_reduce(tot)
    assert tot.item() == 50000
    assert abs(mine - 50000 / world) <= counts.max()
    # 3. batch split
    lo, hi = nb.shard_batch(32, rank, world)
    assert (hi - lo) == 32 // world and lo == rank * (32 // world)
    try:
        nb.shard_batch(33, rank, world); raise SystemExit("expected ArgumentError")
    except nb.ArgumentError:
        pass
    dist.barrier(); dist.destroy_process_group()
    sys.stdout.write("RANK_OK" + str(rank) + chr(10)); sys.stdout.flush()
""") % ROOT


def test_world_size_2_gloo(tmp_path):
    import nfft_jl_b200 as nb
    if not os.path.exists(nb.LIB_PATH):
        nb.build()
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29533", str(script)],
                         capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert out.stdout.count("RANK_OK") == 2


def test_partition_edge_cases():
    import nfft_jl_b200 as nb
    ts = np.array([0, 0, 0, 10, 10, 10], dtype=np.int64)        # all nodes in one tile
    cut = nb.partition_tiles(ts, 4)
    assert cut[0] == 0 and cut[-1] == 5 and np.all(np.diff(cut) >= 0)
    sizes = [ts[cut[r + 1]] - ts[cut[r]] for r in range(4)]
    assert sum(sizes) == 10
    cut = nb.partition_tiles(np.array([0, 3, 6, 9, 12], dtype=np.int64), 2)
    assert cut.tolist() == [0, 2, 4]
    cut = nb.partition_tiles(np.array([0], dtype=np.int64), 3)  # no tiles
    assert cut.tolist() == [0, 0, 0, 0]
