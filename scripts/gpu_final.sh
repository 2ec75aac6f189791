#!/bin/bash
# Round-2 evidence (1 GPU).  Outputs land in gpurun_out/ (merged back only up to 64 MiB per call, hence the stages):
#   bash scripts/gpu_final.sh bench   bench line, reference arm, ncu launch list of the bench command
#   bash scripts/gpu_final.sh c2      ncu --set full capture of the three hot kernels of C2
#   bash scripts/gpu_final.sh c3      ncu --set full capture of the batched 2-D kernels on C3
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
case "${1:-bench}" in
bench)
  timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err
  tail -2 gpurun_out/r02_bench_n1.err
  timeout 300 python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/r02_bench_reference_arm.json 2>/dev/null
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --headline-only > /dev/null 2>&1
  python -c "
import json; d=json.load(open('gpurun_out/r02_bench_n1.json'))
print({k: d[k] for k in ('value','ms_per_step','phases_us')}); print(d['roofline']); print(d['e2e']); print(d['cpu_baseline']); print(d['clocks'])"
  ;;
c2)
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_spread_lean|k_gather_cols3d|k_interp_lean' -c 3 \
    -o gpurun_out/r02_hot -f python scripts/prof_c2.py C2 1 0 2>&1 | tail -2
  ;;
c3)
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_spread_win2d|k_interp_batch2d|k_gather_tile2d' -c 3 \
    -o gpurun_out/r02_hot_c3 -f python scripts/prof_cfg.py C3 2>&1 | tail -1
  ;;
esac
