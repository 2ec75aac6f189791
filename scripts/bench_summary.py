import json, sys
for f in sys.argv[1:]:
    try:
        d = json.load(open(f))
    except Exception as e:
        print(f, "unreadable", e); continue
    print(f"{f}: {d['value']/1e9:.3f} Gpts/s  step {d['ms_per_step']:.3f} ms  bs={d['config']['blockSize']}  "
          + " ".join(f"{k}={v:.0f}" for k, v in d['phases_us'].items())
          + f"  spread {d['roofline']['us_per_launch']:.0f}us ({d['roofline']['frac']*100:.1f}% HBM)  interp {d['roofline']['interp']['us_per_launch']:.0f}us ({d['roofline']['interp']['frac']*100:.1f}%)  e2e {d['e2e']['value']/1e9:.2f}")
