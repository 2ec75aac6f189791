"""Print the last launches of an ncu `--metrics gpu__time_duration.sum --csv` log.  usage: ncu_tail.py [file] [n]"""
import csv, sys
f = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/c3_launches.csv"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 8
rows = [r for r in csv.reader(open(f)) if len(r) > 5]
h = rows[0]; ik = h.index("Kernel Name"); iv = h.index("Metric Value")
for r in rows[-n:]:
    print(f"{float(r[iv]) / 1e3:9.1f} us  {r[ik][:90]}")
