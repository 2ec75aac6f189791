"""nfft.jl_b200 -- B200-native NFFT backend behind the AbstractNFFTs plan API.

Only what the hot path needs lives here: `csrc/` (hand-written CUDA for sm_100a + the C ABI declared in
include/nfftb200.h), `julia/B200NFFT.jl` (the glue a Julia user loads) and `plan.py` (the executable
Python mirror of that glue).  Import it as `nfft_jl_b200` (see the shim at the repo root)."""
from ._lib import LIB_PATH, SYMBOLS, build, lib
from .plan import (FULL, LINEAR, POLYNOMIAL, TENSOR, AdjointPlan, ArgumentError, B200NFFTPlan,
                   DimensionMismatch, NFFTParams, PrecomputeFlags, TimingStats, accuracyParams, adjoint,
                   convolve_, convolve_transpose_, deconvolve_, deconvolve_transpose_, mul_, nfft,
                   nfft_adjoint, nodes_, partition_tiles, plan_nfft, sdc, sdc_host_loop, shard_batch, size_in,
                   size_out)
from .toeplitz import (ToeplitzOperator, calculateToeplitzKernel, calculateToeplitzKernel_,
                       convolveToeplitzKernel_)

__all__ = ["calculateToeplitzKernel", "calculateToeplitzKernel_", "convolveToeplitzKernel_", "ToeplitzOperator",
           "plan_nfft", "nodes_", "mul_", "adjoint", "size_in", "size_out", "convolve_", "convolve_transpose_",
           "deconvolve_", "deconvolve_transpose_", "nfft", "nfft_adjoint", "sdc", "sdc_host_loop", "PrecomputeFlags", "FULL", "TENSOR",
           "LINEAR", "POLYNOMIAL", "TimingStats", "NFFTParams", "B200NFFTPlan", "AdjointPlan", "ArgumentError",
           "DimensionMismatch", "accuracyParams", "shard_batch", "partition_tiles", "build", "lib", "LIB_PATH", "SYMBOLS"]
