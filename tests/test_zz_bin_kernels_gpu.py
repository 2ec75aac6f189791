"""GPU check of the OPT-IN register-footprint kernels (kernel_mode 7, csrc/spread_bin.cuh and csrc/interp_bin.cuh):
adjoint and forward parity against the default kernels and the oracle (scripts/try_bin_kernels.py).  The kernel
sources are verified on the host by tests/test_emu_bin_kernels_cpu.py; they had no hardware run yet when they were committed (the round's GPU budget was
spent), so the check runs in a child process with a time limit -- a fault or hang in the experimental kernel cannot
take the default-path tests with it -- and is a non-strict xfail until it has been seen green on a B200.  The default
path (kernel_mode 0) never executes these kernels."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.xfail(strict=False, reason="experimental kernel_mode 7: verified by host emulation only so far")
def test_bin_kernels_parity_on_gpu():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "try_bin_kernels.py")], cwd=ROOT,
                         capture_output=True, text=True, timeout=600)
    print(out.stdout[-3000:])
    assert out.returncode == 0 and "PARITY OK" in out.stdout, (out.stdout + out.stderr)[-3000:]
