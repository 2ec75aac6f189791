"""GPU parity of the register-window kernels (kernel_mode 7: csrc/spread_bin.cuh, csrc/interp_bin.cuh) against the
default tiled kernels (kernel_mode 0) and the oracle: Float32/Float64, m = 2, 3, 4, clustered nodes (several chunks per
tile, split work items), batches, thin tiles.  They were seen green on a B200 at the end of round 1 (an XPASS of the
then non-strict xfail), so they are ordinary tests now; C2 at full size in both modes is in test_gpu_fullsize.py.
Every result must also be bit-reproducible run to run (no atomics on the data path)."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "scripts"))

f32, f64 = np.float32, np.float64
CASES = [
    ("f32 m=3 uniform", (32, 32, 32), 20000, f32, 3, {}),
    ("f32 m=3 clustered", (48, 32, 40), 30000, f32, 3, dict(cluster=6000)),
    ("f64 m=3 clustered", (32, 32, 32), 20000, f64, 3, dict(cluster=3000)),
    ("f32 m=2", (32, 32, 32), 20000, f32, 2, {}),
    ("f32 m=4 (W=10)", (32, 32, 32), 20000, f32, 4, {}),
    ("f64 m=4 (falls back to the default kernel)", (32, 32, 32), 8000, f64, 4, {}),
    ("f32 m=3 B=3", (32, 32, 32), 20000, f32, 3, dict(B=3)),
    ("f32 m=3 thin tiles", (32, 32, 32), 20000, f32, 3, dict(blockSize=(16, 16, 8))),
    ("f32 m=3 64^3 (C2 density)", (64, 64, 64), 2 ** 18, f32, 3, dict(oracle=False)),
]


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_bin_kernels_parity(case):
    import try_bin_kernels as tb
    name, N, M, T, m, kw = case
    assert tb.parity_case(name, N, M, T, m, **kw), tb.out_lines[-1]
