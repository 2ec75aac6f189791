"""The spreader's exact colour dependency table (csrc/spread_lean.cuh: lean_make_need) against the explicit cell sets of
the bins' read-modify-write windows: a missing dependency would be a silent data race on the GPU, so the formula is
checked on the host for both footprints the kernel is instantiated for."""
import importlib.util, os

import pytest


def load():
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "scripts", "check_lean_need.py")
    spec = importlib.util.spec_from_file_location("check_lean_need", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.mark.parametrize("m", [2, 3])
def test_every_overlapping_pair_of_bins_is_ordered(m):
    assert load().check(m) == 0


def test_formula_in_the_kernel_source_is_the_one_checked():
    """the three lines that decide an overlap must be present verbatim in the kernel header"""
    src = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "nfft.jl_b200", "csrc", "spread_lean.cuh")).read()
    assert "const int oa = W * ((a >> d) & 1) + N::G * (cc % N::S), ob = W * ((b >> d) & 1) + N::G * (ee % N::S);" in src
    assert "if (dist >= W) hit = false;" in src
    assert "if (hit) need = e + 1;" in src
