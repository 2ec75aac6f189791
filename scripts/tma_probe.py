"""Exercises the TMA tensor-map staging path of the 3-D interpolator (tiles whose origin is 16-byte aligned:
Float64 always, Float32 for even m) and the TMA bulk store of the spreader; checks against the oracle."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import nfft_jl_b200 as nb
from oracle import nfft_oracle as O
for N, M, T, m in [((32, 32, 32), 20000, np.float32, 4), ((32, 32, 32), 20000, np.float32, 3), ((40, 40, 40), 20000, np.float64, 3),
                   ((32, 32, 32), 20000, np.float32, 2)]:
    k = O.random_nodes(M, 3, T, seed=1)
    p = nb.plan_nfft(k.T, N, m=m, σ=2.0)
    po = O.OraclePlan(k, N, m=m, sigma=2.0, blockSize=p.params.blockSize)
    f = O.random_complex(N, T, 2); fh = O.random_complex(M, T, 3)
    out = p * f; ref = po.forward(f)
    adj = p.adjoint() * fh; refa = po.adjoint(fh)
    print(N, T.__name__, "m", m, "bs", p.params.blockSize, "forward rel", np.linalg.norm(out - ref) / np.linalg.norm(ref),
          "adjoint rel", np.linalg.norm(adj - refa) / np.linalg.norm(refa))
