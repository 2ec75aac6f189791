"""Independent check of the spreader's colour dependency table (csrc/spread_lean.cuh: LeanNeed / lean_make_need).
Rebuilds the table with the kernel's formula and verifies it against the explicit cell sets of the bins' read-modify-write
windows: bins of one colour never overlap, and every pair of overlapping bins is ordered by the table.  CPU only."""
import itertools

def check(MT, W=8):
    L = 2 * MT; G = W - L + 1; S = (W + G - 1) // G; S3 = S ** 3
    need = {}
    for a in range(8):
        for b in range(8):
            for c in range(S3):
                n = 0
                for e in range(c):
                    if a == b: break
                    hit = True; cc = c; ee = e
                    for d in range(3):
                        oa = W * ((a >> d) & 1) + G * (cc % S); ob = W * ((b >> d) & 1) + G * (ee % S); cc //= S; ee //= S
                        if abs(oa - ob) >= W: hit = False
                    if hit: n = e + 1
                need[(a, b, c)] = n
    def cells(a, c):
        rng = []
        for d in range(3):
            k = c % S; c //= S
            o = 1 + W * ((a >> d) & 1) + G * k
            rng.append(range(o, o + W))
        return set(itertools.product(*rng))
    cs = {(a, c): cells(a, c) for a in range(8) for c in range(S3)}
    bad = 0
    for a in range(8):
        for b in range(8):
            if a == b: continue
            for c in range(S3):
                assert not (cs[(a, c)] & cs[(b, c)]), "bins of one colour overlap"
                for e in range(S3):
                    if e == c or not (cs[(a, c)] & cs[(b, e)]): continue
                    if e < c and need[(a, b, c)] < e + 1: bad += 1      # (a, c) must wait for (b, e)
                    if e > c and need[(b, a, e)] < c + 1: bad += 1      # (b, e) must wait for (a, c)
    return bad

if __name__ == "__main__":
    for MT in (2, 3):
        print("m =", MT, "violations:", check(MT))
