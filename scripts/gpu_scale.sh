#!/bin/bash
# multi-GPU run on one box: correctness (tests/mgpu_check.py) and bench.py at the given rank counts
#   gpurun --gpus 8 --timeout 1500 -- 'bash scripts/gpu_scale.sh 8 4'
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
port=29600
for n in "$@"; do
  port=$((port+1))
  echo "== mgpu_check N=$n"
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port tests/mgpu_check.py 2>&1 | grep "MGPU_CHECK\|FAIL" | tail -3
  port=$((port+1))
  echo "== bench N=$n"
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/bench_n$n.json 2> gpurun_out/bench_n$n.err
  tail -2 gpurun_out/bench_n$n.err
  port=$((port+1))
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port bench.py --impl reference --gpus $n --steps 5 --warmup 2 > gpurun_out/bench_ref_n$n.json 2>/dev/null
  python - <<PY
import json
d = json.load(open("gpurun_out/bench_n$n.json"))
print({k: d[k] for k in ("n_gpus", "value", "ms_per_step")}, "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e"].get("numa_bound_cores_per_rank"))
for c in ("C5", "C4"):
    for m in ("fused", "nccl"):
        v = d["node_sharded"][c][m]
        print(c, m, {a: (round(b, 3) if isinstance(b, float) and b > 1e-3 else b) for a, b in v.items() if a != "phases_us"}, {a: round(b, 1) for a, b in v["phases_us"].items()})
try:
    r = json.load(open("gpurun_out/bench_ref_n$n.json")); print("reference arm:", r["value"], r["ms_per_step"], r["cpu_baseline"]["cores"])
except Exception as e:
    print("reference arm failed", e)
PY
done
