"""Opcode histogram (warp instructions per node) + headline counters of one kernel in an ncu report.
usage: python scripts/ncu_ops.py REP KERNEL_REGEX NODES"""
import collections, csv, io, subprocess, sys
rep, kern, nodes = sys.argv[1], sys.argv[2], float(eval(sys.argv[3]))
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "-k", "regex:" + kern], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = next(r for r in rows if r and r[0] == "Address")
ia, isrc = hdr.index("Instructions Executed"), hdr.index("Source")
iw = hdr.index("L1 Wavefronts Shared") if "L1 Wavefronts Shared" in hdr else 0
tot = 0; by = collections.Counter(); wf = 0
for r in rows:
    if len(r) <= ia or not r[ia].isdigit(): continue
    n = int(r[ia]); tot += n
    t = r[isrc].split()
    op = t[1] if t[0].startswith('@') else t[0]
    by[op.split('.')[0]] += n
    if r[iw].isdigit(): wf += int(r[iw])
print(f"total warp-instr {tot}  = {tot/nodes:.1f}/node ; shared wavefronts {wf/nodes:.1f}/node")
print("  ".join(f"{op} {n/nodes:.1f}" for op, n in by.most_common(24)))
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv", "-k", "regex:" + kern], capture_output=True, text=True).stdout
rr = list(csv.reader(io.StringIO(raw)))
h = rr[0]
for name in ("gpu__time_duration.sum", "sm__inst_executed.avg.per_cycle_active", "smsp__issue_active.avg.pct", "dram__bytes_read.sum", "dram__bytes_write.sum",
             "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
             "smsp__average_warp_latency_issue_stalled_barrier.pct", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
             "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
             "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
             "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
             "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio"):
    if name in h:
        i = h.index(name)
        print(f"  {name} = {rr[2][i]} {rr[1][i]}")
