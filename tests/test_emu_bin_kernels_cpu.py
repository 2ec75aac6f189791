"""The opt-in register-footprint spreader (csrc/spread_bin.cuh) and interpolator (csrc/interp_bin.cuh) of kernel_mode 7, compiled for the HOST and run
one OS thread per CUDA thread (tests/emu): every padded tile it writes is compared with a direct evaluation,
two runs must agree bitwise, and the vector accesses must be aligned.  Test infrastructure only -- the product
path is the CUDA build of the same header."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build(tmp_path, extra=()):
    exe = str(tmp_path / "emu_bin_kernels")
    cmd = ["g++", "-std=c++17", "-O1", "-ffp-contract=off", "-w", "-pthread", *extra,
           "-I/usr/local/cuda/include", "-I/usr/include",
           os.path.join(ROOT, "tests", "emu", "emu_bin_kernels.cpp"), "-o", exe]
    subprocess.check_call(cmd)
    return exe


def test_bin_kernel_sources_on_host(tmp_path):
    if not os.path.exists("/usr/local/cuda/include/cuda_runtime.h"):
        pytest.skip("CUDA headers not installed")
    out = subprocess.run([_build(tmp_path)], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "ALL OK" in out.stdout and "FAIL" not in out.stdout, out.stdout
    assert out.stdout.count(": OK") >= 17, out.stdout


@pytest.mark.skipif(os.environ.get("NFFTB_EMU_TSAN", "0") != "1", reason="slow; set NFFTB_EMU_TSAN=1")
def test_bin_kernel_sources_race_free(tmp_path):
    out = subprocess.run([_build(tmp_path, ("-g", "-fsanitize=thread"))], capture_output=True, text=True, timeout=1800)
    assert out.returncode == 0 and "ThreadSanitizer" not in out.stdout + out.stderr, (out.stdout + out.stderr)[-4000:]
