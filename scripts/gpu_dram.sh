#!/bin/bash
# DRAM bytes + duration of the spread-stage kernels on C2 for a given library / kernel mode (ncu, metrics only)
cd "$(dirname "$0")/.."
NFFTB200_LIB=${1:-} timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct --clock-control none \
  -k regex:'k_spread|k_gather' -c 3 --csv python scripts/prof_c2.py C2 1 ${2:-0} 2>/dev/null | grep -v "^==" | python -c "
import csv,sys
r=[x for x in csv.reader(sys.stdin) if len(x)>10]
h=r[0]
for x in r[1:]:
    print(x[h.index('Kernel Name')][:40], x[h.index('Metric Name')], x[h.index('Metric Value')], x[h.index('Metric Unit')])
"
