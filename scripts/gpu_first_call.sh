#!/bin/bash
# One gpurun call that answers "do the kernel_mode-7 kernels (spread_bin.cuh / interp_bin.cuh) work on a B200, and how
# fast are they": parity + timings of both modes, a bench line per mode, and one ncu capture of the mode-7 kernels.
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash scripts/gpu_first_call.sh'
# Everything lands in gpurun_out/ (merged back by gpurun); nothing here is a bench value of record -- numbers printed
# under ncu never are.
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
{
  echo "== parity + timings (scripts/try_bin_kernels.py --time)"
  timeout 600 python scripts/try_bin_kernels.py --time
  echo "exit $?"
  echo "== bench, default kernels"
  timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline
  echo "== bench, kernel_mode 7"
  timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --kernel-mode 7
} > gpurun_out/first_call.log 2>&1
# launch list (cheap) and one full capture of the two new kernels
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/mode7_launches.csv \
  python scripts/prof_c2.py C2 2 7 >> gpurun_out/first_call.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_(spread|interp)_bin3d' -c 2 \
  -o gpurun_out/mode7 -f python scripts/prof_c2.py C2 1 7 >> gpurun_out/first_call.log 2>&1
tail -40 gpurun_out/first_call.log
# follow-up on two GPUs (separate call, gpurun --gpus 2): node sharding with the mode-7 kernels on the fused peer paths
#   python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 scripts/bench_nodes_sharded.py C5s 7
