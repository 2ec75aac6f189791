#!/bin/bash
# Round-2 evidence in one gpurun call (1 GPU): bench line, reference arm, ncu launch list of the bench command and one
# ncu --set full capture of the three hot kernels of C2.  Outputs land in gpurun_out/ (copied to profiles/ afterwards).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err
tail -2 gpurun_out/r02_bench_n1.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/r02_bench_reference_arm.json 2>/dev/null
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline --headline-only > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_spread_lean|k_gather_cols3d|k_interp_lean' -c 3 \
  -o gpurun_out/r02_hot -f python scripts/prof_c2.py C2 1 0 2>&1 | tail -2
python -c "
import json; d=json.load(open('gpurun_out/r02_bench_n1.json'))
print({k: d[k] for k in ('value','ms_per_step','phases_us')}); print(d['roofline']); print(d['e2e']); print(d['cpu_baseline']); print(d['clocks'])"
