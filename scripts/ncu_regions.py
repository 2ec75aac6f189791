"""Aggregates an `ncu --page source --csv` dump into hot SASS regions (instructions, samples, smem wavefronts)."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
starts = [i for i, r in enumerate(rows) if r and r[0] == 'Kernel Name']
end = starts[1] if len(starts) > 1 else len(rows)
hdr = rows[1]; data = rows[2:end]
ia = hdr.index('Instructions Executed'); isamp = hdr.index('# Samples'); isrc = hdr.index('Source')
iw = hdr.index('L1 Wavefronts Shared') if 'L1 Wavefronts Shared' in hdr else isamp; iwi = hdr.index('L1 Wavefronts Shared Ideal') if 'L1 Wavefronts Shared Ideal' in hdr else isamp
tot = sum(int(r[ia]) for r in data); tots = sum(int(r[isamp]) for r in data)
print("total inst", tot, "n sass", len(data), "samples", tots)
runs = []; cur = None
for k, r in enumerate(data):
    c = int(r[ia])
    if cur and abs(c - cur[2]) <= 0.02 * max(c, cur[2], 1):
        cur[1] = k; cur[3] += c; cur[4] += int(r[isamp]); cur[5] += int(r[iw]); cur[6] += int(r[iwi])
    else:
        cur = [k, k, c, c, int(r[isamp]), int(r[iw]), int(r[iwi])]; runs.append(cur)
for a, b, c, s, smp, w, wi in runs:
    if s > 0.01 * tot or smp > 0.02 * tots:
        print(f"sass[{a:4d}-{b:4d}] n={b-a+1:4d} exec/inst={c:>10d} inst={s/tot*100:5.1f}% samples={smp/tots*100:5.1f}% smem_wf={w} ideal={wi}")
if len(sys.argv) > 3:
    for r in data[int(sys.argv[2]):int(sys.argv[3]) + 1]:
        print(r[isrc].strip()[:100], r[isamp])
