"""ctypes binding of libnfftb200.so (include/nfftb200.h).  There is NO fallback: if the CUDA library is
missing or a call fails, an exception is raised."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("NFFTB200_LIB") or os.path.join(_HERE, "libnfftb200.so")   # NFFTB200_LIB: another build of the same ABI

# every symbol include/nfftb200.h declares
SYMBOLS = [
    "nfftb200_accuracy_params", "nfftb200_plan_create", "nfftb200_destroy", "nfftb200_set_nodes",
    "nfftb200_get_permutation", "nfftb200_get_info", "nfftb200_get_table", "nfftb200_exec_forward",
    "nfftb200_exec_adjoint", "nfftb200_convolve", "nfftb200_convolve_transpose", "nfftb200_deconvolve",
    "nfftb200_deconvolve_transpose", "nfftb200_get_grid", "nfftb200_fft", "nfftb200_set_timing",
    "nfftb200_get_timing", "nfftb200_get_kernel_times", "nfftb200_set_kernel_mode", "nfftb200_get_launch_count", "nfftb200_set_stream",
    "nfftb200_sync", "nfftb200_comm_unique_id", "nfftb200_partition_tiles", "nfftb200_comm_init", "nfftb200_comm_is_fused", "nfftb200_last_error",
    "nfftb200_status_string", "nfftb200_version",
    "nfftb200_sdc", "nfftb200_toeplitz_kernel", "nfftb200_toeplitz_create", "nfftb200_toeplitz_set_kernel", "nfftb200_toeplitz_apply",
    "nfftb200_toeplitz_set_stream", "nfftb200_toeplitz_sync", "nfftb200_toeplitz_destroy", "nfftb200_toeplitz_last_error",
]

_lib = None


def build(force: bool = False) -> str:
    """Compile csrc/*.cu for sm_100a into libnfftb200.so (in-tree)."""
    if force or not os.path.exists(LIB_PATH):
        subprocess.check_call(["bash", os.path.join(_HERE, "csrc", "build.sh")])
    return LIB_PATH


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: the B200 NFFT backend has no CPU fallback. "
                "Build it with `python -c 'import __graft_entry__ as g; g.build()'`.")
        L = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
        for s in SYMBOLS:
            getattr(L, s)  # raises AttributeError if the ABI is incomplete
        L.nfftb200_last_error.restype = C.c_char_p
        L.nfftb200_last_error.argtypes = [C.c_void_p]
        L.nfftb200_status_string.restype = C.c_char_p
        L.nfftb200_plan_create.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_void_p, C.c_int, C.c_int,
                                           C.c_double, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int]
        L.nfftb200_destroy.argtypes = [C.c_void_p]
        L.nfftb200_set_nodes.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int]
        L.nfftb200_get_permutation.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.nfftb200_get_info.argtypes = [C.c_void_p] + [C.c_void_p] * 6
        L.nfftb200_get_table.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int64, C.c_void_p]
        for n in ("nfftb200_exec_forward", "nfftb200_exec_adjoint", "nfftb200_deconvolve",
                  "nfftb200_deconvolve_transpose"):
            getattr(L, n).argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        for n in ("nfftb200_convolve", "nfftb200_convolve_transpose"):
            getattr(L, n).argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        L.nfftb200_get_grid.argtypes = [C.c_void_p, C.POINTER(C.c_void_p)]
        L.nfftb200_fft.argtypes = [C.c_void_p, C.c_int]
        L.nfftb200_set_timing.argtypes = [C.c_void_p, C.c_int]
        L.nfftb200_get_timing.argtypes = [C.c_void_p, C.c_void_p]
        L.nfftb200_get_kernel_times.argtypes = [C.c_void_p, C.c_void_p]
        L.nfftb200_set_kernel_mode.argtypes = [C.c_void_p, C.c_int]
        L.nfftb200_get_launch_count.argtypes = [C.c_void_p, C.c_void_p]
        L.nfftb200_set_stream.argtypes = [C.c_void_p, C.c_void_p]
        L.nfftb200_sync.argtypes = [C.c_void_p]
        L.nfftb200_comm_unique_id.argtypes = [C.c_void_p]
        L.nfftb200_partition_tiles.argtypes = [C.c_void_p, C.c_int64, C.c_int, C.c_void_p]
        L.nfftb200_comm_init.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int]
        L.nfftb200_comm_is_fused.argtypes = [C.c_void_p]
        L.nfftb200_accuracy_params.argtypes = [C.c_int, C.c_double, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p]
        L.nfftb200_sdc.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int]
        L.nfftb200_toeplitz_kernel.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.nfftb200_toeplitz_create.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int]
        L.nfftb200_toeplitz_set_kernel.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.nfftb200_toeplitz_apply.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.nfftb200_toeplitz_set_stream.argtypes = [C.c_void_p, C.c_void_p]
        L.nfftb200_toeplitz_sync.argtypes = [C.c_void_p]
        L.nfftb200_toeplitz_destroy.argtypes = [C.c_void_p]
        L.nfftb200_toeplitz_last_error.argtypes = [C.c_void_p]
        L.nfftb200_toeplitz_last_error.restype = C.c_char_p
        _lib = L
    return _lib
