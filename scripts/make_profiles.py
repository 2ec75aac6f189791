"""Turns an ncu report (gpurun_out/*.ncu-rep) into the tracked summaries under profiles/:
   profiles/<tag>_kernels.csv   one row per captured kernel with the counters DESIGN.md / bench.py cite
   profiles/<tag>_hot_<kernel>.txt   hot SASS regions (instruction share, stall-sample share, smem wavefronts)
   profiles/ncu_traffic.json    DRAM bytes per launch of the dominant kernel (bench.py roofline.traffic)
usage: python scripts/make_profiles.py gpurun_out/prof.ncu-rep r01"""
import csv, io, json, os, subprocess, sys

rep, tag = sys.argv[1], sys.argv[2]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out = os.path.join(ROOT, "profiles")
os.makedirs(out, exist_ok=True)
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
KEYS = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
        "smsp__inst_executed.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__occupancy_limit_shared_mem", "launch__grid_size", "launch__block_size",
        "smsp__inst_executed_op_shared_atom.sum", "smsp__inst_executed_op_global_red.sum", "sm__cycles_elapsed.max"]
idx = [(k, hdr.index(k)) for k in KEYS if k in hdr]
with open(os.path.join(out, f"{tag}_kernels.csv"), "w", newline="") as fh:
    w = csv.writer(fh)
    w.writerow([k for k, _ in idx]); w.writerow([units[i] for _, i in idx])
    traffic = {}
    for r in rows[2:]:
        w.writerow([r[i] for _, i in idx])
        name = r[hdr.index("Kernel Name")]
        def val(k):
            i = hdr.index(k); v = float(r[i].replace(",", "")); u = units[i]
            return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
        import re
        short = re.search(r"(k_[a-z0-9_]+)", name).group(1)
        traffic[short] = val("dram__bytes_read.sum") + val("dram__bytes_write.sum")
if "k_spread_lean" in traffic:          # round 2 default kernels: same keys bench.py reads
    tj = {"source": f"{os.path.basename(rep)} (ncu --set full, {tag}; profiles/{tag}_kernels.csv)",
          "spread_stage_dram_bytes_per_launch": {"lean": traffic["k_spread_lean"] + traffic.get("k_gather_cols3d", 0), "sub": 828784896.0, "bin": None},
          "per_kernel_dram_bytes_per_launch": traffic}
    json.dump(tj, open(os.path.join(out, "ncu_traffic.json"), "w"), indent=1)
elif "k_spread_sub3d" in traffic or "k_spread_bin3d" in traffic:
    tj = {"source": os.path.basename(rep), "per_kernel_dram_bytes_per_launch": traffic,
          "spread_dram_bytes_per_launch": traffic.get("k_spread_sub3d", 0) + traffic.get("k_gather_tiles3d", 0) + traffic.get("k_gather_cols3d", 0)}
    if "k_spread_bin3d" in traffic:      # kernel_mode 7 (bench.py --kernel-mode 7 reads this key)
        tj["spread_bin_dram_bytes_per_launch"] = traffic["k_spread_bin3d"] + traffic.get("k_gather_tiles3d", 0) + traffic.get("k_gather_cols3d", 0)
    json.dump(tj, open(os.path.join(out, "ncu_traffic.json"), "w"), indent=1)
for kern in sorted(traffic):
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kern}"],
                         capture_output=True, text=True).stdout
    tmp = os.path.join("/tmp", f"src_{kern}.csv"); open(tmp, "w").write(src)
    txt = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "ncu_regions.py"), tmp], capture_output=True, text=True).stdout
    open(os.path.join(out, f"{tag}_hot_{kern}.txt"), "w").write(f"# hot SASS regions of {kern} ({os.path.basename(rep)}; ncu --set full --import-source on)\n" + txt)
print("wrote", os.listdir(out))
