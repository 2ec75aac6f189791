"""Per-source-line summary of an ncu report (read on the CPU box): warp-instructions executed and stall samples per
CUDA source line, from `ncu -i REP --page source --csv --print-source cuda,sass -k regex:KERNEL`.
usage: python scripts/ncu_lines.py REP KERNEL_REGEX [top_n]"""
import csv
import io
import subprocess
import sys


def main():
    rep, kern = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "-k",
                          "regex:" + kern], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    fpath, hdr, lines = None, None, []
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            fpath = r[1].split("/")[-1]
        elif r[0] == "Line No":
            hdr = r
        elif r[0] not in ("", "Function Name", "Kernel Name") and hdr and r[0].isdigit():
            d = dict(zip(hdr[4:], r[4:]))
            num = lambda v: int(v) if str(v).isdigit() else 0
            lines.append((fpath, int(r[0]), r[1].strip(), num(d.get("# Samples", 0)),
                          num(d.get("Instructions Executed", 0)), d))
    tot_s = sum(l[3] for l in lines) or 1
    tot_i = sum(l[4] for l in lines) or 1
    print(f"total samples {tot_s}, warp-instructions {tot_i}")
    print("-- by samples")
    for f, n, src, s, i, d in sorted(lines, key=lambda l: -l[3])[:top]:
        st = {k: int(v) for k, v in d.items() if k.startswith("stall_") and "Not Issued" not in k and v.isdigit() and int(v)}
        st = sorted(st.items(), key=lambda kv: -kv[1])[:3]
        print(f"{100*s/tot_s:5.1f}% smp {100*i/tot_i:5.1f}% ins  {f}:{n}  {src[:90]}   {st}")
    print("-- by instructions")
    for f, n, src, s, i, d in sorted(lines, key=lambda l: -l[4])[:top]:
        print(f"{100*i/tot_i:5.1f}% ins {100*s/tot_s:5.1f}% smp  {f}:{n}  {src[:100]}")


if __name__ == "__main__":
    main()
