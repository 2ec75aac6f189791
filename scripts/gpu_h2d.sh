#!/bin/bash
cd "$(dirname "$0")/.."
port=29700
for n in 1 2 4 8; do
  for b in "" "--bind"; do
    port=$((port+1))
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port scripts/h2d_ceiling.py $b 2>/dev/null | tail -1
  done
done
nvidia-smi topo -m 2>/dev/null | head -14
lscpu | grep -i "numa\|socket\|model name" | head -8
