"""Top SASS lines by stall samples of one kernel in an ncu report.  usage: ncu_hot.py REPORT KERNEL_REGEX [N]"""
import csv, io, subprocess, sys, collections
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv", "-k", "regex:" + kern], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw))); h = rows[0]
for r in rows[2:3]:
    print(r[h.index("Kernel Name")][:80])
    for w in ["gpu__time_duration.sum", "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_elapsed",
              "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
              "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
              "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "dram__bytes_read.sum", "dram__bytes_write.sum"]:
        if w in h: print("  ", w, r[h.index(w)][:24])
    for i, c in enumerate(h):
        if "pcsamp_warps_issue_stalled" in c and not c.endswith("not_issued") and float(r[i] or 0) > 1500:
            print("      ", c.replace("smsp__pcsamp_warps_issue_stalled_", ""), r[i])
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "-k", "regex:" + kern], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = None; lines = []
for r in rows:
    if not r: continue
    if "Instructions Executed" in r:
        if hdr is not None: break          # second copy of the listing
        hdr = r; ia = hdr.index("Instructions Executed"); isrc = hdr.index("Source"); ismp = hdr.index("# Samples"); continue
    if hdr and len(r) > ia and r[ia].isdigit():
        lines.append((len(lines), int(r[ia]), r[isrc], int(r[ismp]) if r[ismp].isdigit() else 0))
tots = sum(l[3] for l in lines)
print("sass lines", len(lines), "samples", tots)
for idx, c, sx, sm in sorted(lines, key=lambda t: -t[3])[:top]:
    print(f"{sm:6d} {100.0 * sm / max(tots, 1):5.1f}%  [{idx:5d}] x{c:>10d}  {sx[:90]}")
