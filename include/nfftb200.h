/*
 * nfftb200.h -- C ABI of the B200-native NFFT backend (libnfftb200.so).
 *
 * This is the drop-in boundary for the NFFT hot path of JuliaMath/NFFT.jl: the Julia glue
 * (nfft.jl_b200/julia/B200NFFT.jl) `ccall`s exactly these entry points, and the Python
 * mirror (nfft.jl_b200/plan.py) binds them with ctypes.  Plain pointers and sizes only; no
 * C++/torch types.  Every function returns an nfftb200_status (0 = ok) and never throws.
 *
 * Conventions (all taken from the reference, /root/reference):
 *   - arrays are Julia column-major: nodes k are D x M (one node's D coordinates contiguous),
 *     image f has N[0] fastest, the oversampled grid g has Nt[0] fastest;
 *   - complex values are interleaved (re, im) in the plan's precision T;
 *   - nodes live in [-1/2, 1/2] (src/utils.jl:46-55);
 *   - batched transforms (ntransforms = B) put the batch in the slowest dimension:
 *     f is (N..., B), fHat is (M, B)  (test/accuracy.jl:123-163, src/directional.jl);
 *   - no 1/prod(Nt) normalisation anywhere (adjoint uses bfft, src/implementation.jl:88,182).
 *
 * Each entry point cites the reference interface it replaces.
 */
#ifndef NFFTB200_H
#define NFFTB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct nfftb200_plan nfftb200_plan;

typedef enum {
    NFFTB200_OK = 0,
    NFFTB200_BAD_NODE_RANGE = 1,  /* ArgumentError, src/utils.jl:50 */
    NFFTB200_BAD_DIM = 2,         /* ArgumentError, src/precomputation.jl:19-21 */
    NFFTB200_SIZE_MISMATCH = 3,   /* DimensionMismatch, src/utils.jl:101-103 */
    NFFTB200_UNSUPPORTED = 4,     /* error(...), src/precomputation.jl:401, windowFunctions.jl:16 */
    NFFTB200_CUDA_ERROR = 5,
    NFFTB200_NCCL_ERROR = 6,
    NFFTB200_OOM = 7,
    NFFTB200_BAD_ARGUMENT = 8,    /* ArgumentError (e.g. complex input with real output, src/convolution.jl:52,148) */
    NFFTB200_NO_NODES = 9
} nfftb200_status;

/* element precision T of the plan (type parameter T of NFFTPlan{T,D,R}) */
enum { NFFTB200_F32 = 0, NFFTB200_F64 = 1 };

/* AbstractNFFTs.PrecomputeFlags, AbstractNFFTs/src/misc.jl:13-18 -- same numeric values.
 * On this backend they select how window weights are produced inside the kernels:
 *   FULL       -> exact Kaiser-Bessel evaluated on the fly (what precomputeB stores, precomputation.jl:124-134)
 *   LINEAR     -> the reference's lookup table with linear interpolation (precomputation.jl:291-300, 179-201)
 *   POLYNOMIAL -> the reference's piecewise polynomial, Horner/FMA (precomputation.jl:302-320, 215-222)
 *   TENSOR     -> numerically identical to POLYNOMIAL (precomputation.jl:556-591); evaluated on the fly */
enum { NFFTB200_FULL = 1, NFFTB200_TENSOR = 2, NFFTB200_LINEAR = 3, NFFTB200_POLYNOMIAL = 4 };

/* window pairs of getWindow (src/windowFunctions.jl:4-19): :kaiser_bessel (default, :21-39), :gauss (:60-73),
 * :spline (:75-99), :kaiser_bessel_rev (:41-57), :cosh_type (:104-134) */
enum {
    NFFTB200_KAISER_BESSEL = 0,
    NFFTB200_GAUSS = 1,
    NFFTB200_SPLINE = 2,
    NFFTB200_KAISER_BESSEL_REV = 3,
    NFFTB200_COSH_TYPE = 4,
    /* not in the reference: the "exponential of semicircle" window exp(beta (sqrt(1 - (x/m)^2) - 1)), beta =
     * 0.97 pi 2m (1 - 1/(2 sigma)), evaluated on the fly in FULL mode; its Fourier coefficients have no closed form and
     * are computed by Gauss-Legendre quadrature at plan time */
    NFFTB200_EXP_SQRT = 5
};

/* where caller buffers live.  NFFTB200_HOST calls return when the result is in the caller's buffer.
 * NFFTB200_HOST_ASYNC (exec_forward / exec_adjoint only) takes page-locked host buffers and returns at once: the
 * upload, the transform and the download are queued on three internal streams, so that back-to-back calls overlap
 * their copies with each other's kernels; the buffers belong to the library until nfftb200_sync returns. */
enum { NFFTB200_HOST = 0, NFFTB200_DEVICE = 1, NFFTB200_HOST_ASYNC = 2 };

/* sharding mode for multi-GPU plans (new; no reference counterpart) */
enum { NFFTB200_SHARD_NONE = 0, NFFTB200_SHARD_BATCH = 1, NFFTB200_SHARD_NODES = 2 };

/* Resolve (m, sigma, reltol) exactly like accuracyParams, AbstractNFFTs/src/misc.jl:66-81.
 * Pass m <= 0 / sigma <= 0 / reltol <= 0 for "keyword not given". */
int nfftb200_accuracy_params(int m_in, double sigma_in, double reltol_in,
                             int* m_out, double* sigma_out, double* reltol_out);

/* NFFTPlan(k, N; m, sigma, window, precompute, blockSize) without nodes
 * (src/implementation.jl:73-106, src/precomputation.jl:3-56).  `block_size` may be NULL
 * (backend default tile) or D entries (the reference's blockSize kwarg, precomputation.jl:32-33).
 * `device` is the CUDA ordinal.  ntransforms >= 1 is the batch size B. */
int nfftb200_plan_create(nfftb200_plan** out, int D, const int64_t* N, int dtype, int m,
                         double sigma, int window, int precompute, int ntransforms,
                         const int64_t* block_size, int device);

/* finalizer (Wrappers/FINUFFT.jl:52-56 precedent); idempotent on NULL */
int nfftb200_destroy(nfftb200_plan* p);

/* nodes!(p, k)  (src/implementation.jl:108-141): checkNodes, shiftNodes!, tile binning
 * (_precomputeBlocks, src/precomputation.jl:487-520) as a stable counting sort on the GPU.
 * k: D x M values of T, host or device. Keeps the grid and the cuFFT plans. */
int nfftb200_set_nodes(nfftb200_plan* p, const void* k, int64_t M, int where);

/* concat_l nodesInBlock[l] (src/precomputation.jl:501-504), 0-based node ids, int64[M] on host.
 * tile_start: int64[ntiles+1] prefix sums (may be NULL). Bit-exact contract vs the reference. */
int nfftb200_get_permutation(nfftb200_plan* p, int64_t* perm, int64_t* tile_start);

/* plan geometry: Nt[D] (p.Ñ), block_size[D], num_tiles, LUTSize, effective sigma, M */
int nfftb200_get_info(nfftb200_plan* p, int64_t* Nt, int64_t* block_size, int64_t* num_tiles,
                      int64_t* lut_size, double* sigma_eff, int64_t* M);

/* copies of the host-built tables, for parity checks (src/precomputation.jl:291-358):
 * which = 0: windowHatInvLUT (sum N_d values), 1: windowPolyInterp ((2m+1)*2m, column-major),
 * 2: windowLinInterp (LUTSize+2).  Values are returned as double. Returns count in *n. */
int nfftb200_get_table(nfftb200_plan* p, int which, double* out, int64_t cap, int64_t* n);

/* mul!(fHat, p, f)  (src/implementation.jl:155-172): fill+deconvolve -> FFT -> convolve */
int nfftb200_exec_forward(nfftb200_plan* p, const void* f, void* fHat, int where);
/* mul!(f, adjoint(p), fHat)  (src/implementation.jl:176-193) */
int nfftb200_exec_adjoint(nfftb200_plan* p, const void* fHat, void* f, int where);

/* AbstractNFFTs.convolve!(p, g, fHat) (src/convolution.jl:20-45) and
 * convolve_transpose!(p, fHat, g) (:115-140).  is_complex = 0 selects real data of type T
 * (density weights, NFFTTools/src/samplingDensity.jl:93-118). One transform (no batch). */
int nfftb200_convolve(nfftb200_plan* p, const void* g, void* fHat, int is_complex, int where);
int nfftb200_convolve_transpose(nfftb200_plan* p, const void* fHat, void* g, int is_complex, int where);

/* AbstractNFFTs.deconvolve!(p, f, g) incl. the preceding fill! (src/implementation.jl:159-160,
 * src/deconvolution.jl:2-45) and deconvolve_transpose!(p, g, f) (:49-92). */
int nfftb200_deconvolve(nfftb200_plan* p, const void* f, void* g, int where);
int nfftb200_deconvolve_transpose(nfftb200_plan* p, const void* g, void* f, int where);

/* p.tmpVec (src/implementation.jl:26): device pointer to the plan's B * prod(Nt) complex grid */
int nfftb200_get_grid(nfftb200_plan* p, void** device_ptr);

/* run only the in-place FFT of the plan's grid; direction -1 forward / +1 backward
 * (p.forwardFFT * tmpVec, p.backwardFFT * tmpVec; src/implementation.jl:161,182) */
int nfftb200_fft(nfftb200_plan* p, int direction);

/* TimingStats (AbstractNFFTs/src/misc.jl:22-30): pre, conv, fft, deconv, conv_adjoint,
 * fft_adjoint, deconv_adjoint in seconds, from CUDA events. enable=1 turns event recording on. */
int nfftb200_set_timing(nfftb200_plan* p, int enable);
int nfftb200_get_timing(nfftb200_plan* p, double out[7]);

/* device time of the last spread stage (spread kernel + gather pass), the last interpolation kernel, the last grid
 * memset and the gather pass alone (seconds, CUDA events on the plan's stream; needs set_timing(1)):
 * out = {spread, interp, memset, gather} */
int nfftb200_get_kernel_times(nfftb200_plan* p, double out[4]);

/* kernel-selection knob for benchmarking/tests:
 *   0 = auto: tiled shared-memory kernels where they apply.  Float32 3-D plans with m <= 3 run the (tile, bin)-ordered
 *       register-window kernels (csrc/spread_lean.cuh, csrc/interp_lean.cuh: no floating-point atomics, interior tiles
 *       staged by a TMA tensor map, LINEAR table staged by a bulk TMA copy); other 3-D plans the warp-private sub-tile
 *       spreader (per-tile sub-grids stored with TMA bulk copies) and the row-per-lane interpolator; a gather pass sums
 *       the per-tile sub-grids; 1-D plans work on a plan-time cell order; 3-D FFT pruned to the z-planes that carry
 *       image frequencies; D = 4 runs on the generic kernels.  Float32 2-D plans with ntransforms >= 8 (default tile
 *       16 x 16) run the batch-stationary kernels of csrc/twod_batch.cuh: one CTA keeps the padded tile of 16 / 32
 *       transforms in shared memory, producer warps evaluate the window taps once per node for the whole batch,
 *   1 = force the generic warp-per-node kernels (global vector REDs),
 *   2 = tiled spreader with the halo flushed by vector REDs instead of scratch + gather,
 *   3 = auto, but the interpolator never uses the TMA tensor-map load,
 *   4 = auto, but the full (unpruned) cuFFT plan,
 *   5 = auto, but the per-cell form of the gather pass,
 *   6 = node-sharded plans: NCCL reduce-scatter of full-grid replicas instead of the fused spread + slab gather
 *       over peer memory (the baseline the fused path is measured against).
 *   7 = 3-D complex transforms, m <= 4: the opt-in register-footprint kernels (csrc/spread_bin.cuh,
 *       csrc/interp_bin.cuh) -- the nodes of a tile are counting-sorted into bins inside the CTA and the footprints
 *       of a bin are summed in registers; falls back to mode 0 where they do not apply (node-sharded plans use the
 *       spreader for the peer scratch and the slab-direct form of the interpolator).  Experimental: verified by
 *       host emulation of the kernel sources (tests/emu); measured on B200 in round 2 (661 / 491 us on C2).
 *       Batched Float32 2-D plans: the batch-stationary kernels without the register windows of the adjoint.
 *   8 = force the (tile, bin)-ordered register-window kernels (csrc/spread_lean.cuh, csrc/interp_lean.cuh): Float32,
 *       3-D, m = 2 or 3, tiles of at most 16 cells.  Mode 0 already selects them where they apply.
 *   9 = the round-1 default: warp-private sub-tile spreader / row-per-lane interpolator for every 3-D plan; the
 *       per-transform kernels for batched 2-D plans.
 *  13 = mode 8 with thread-block clusters of two x-adjacent tiles that merge their shared halo through distributed
 *       shared memory and write one 38-column block to the scratch (experiment).
 *  12 = mode 8 with the compact tile layout in the interpolator (no TMA box load, 8-byte window loads).
 *  11 = mode 8 with the gather pass fused into the spreader ("last arriver gathers", csrc/spread_lean.cuh); measured
 *       slower than spread + separate gather on B200, kept as an experiment. */
int nfftb200_set_kernel_mode(nfftb200_plan* p, int mode);
/* number of kernels + library calls this plan has launched so far */
int nfftb200_get_launch_count(nfftb200_plan* p, int64_t* n);

/* stream control: calls with device buffers are asynchronous on the plan's stream */
int nfftb200_set_stream(nfftb200_plan* p, void* cuda_stream);
int nfftb200_sync(nfftb200_plan* p);

/* ---- multi-GPU (one process per GPU; new) -------------------------------
 * nccl_unique_id: the 128-byte ncclUniqueId produced on rank 0 (nfftb200_comm_unique_id)
 * and distributed by the host language (torch.distributed / MPI.jl).  mode selects
 * NFFTB200_SHARD_BATCH (each rank owns ntransforms/nranks transforms, no exec collective)
 * or NFFTB200_SHARD_NODES (each rank owns a node range; adjoint = local spread ->
 * ncclReduceScatter over slabs of the last grid dim -> slab FFT; forward = slab FFT ->
 * ncclAllGather -> local interpolation).  Where it applies (see nfftb200_comm_is_fused) the adjoint's spread and
 * reduce-scatter are one fused step over peer memory: every rank's per-tile scratch is exported with CUDA IPC and
 * each rank gathers its slab straight from the owners of the covering tiles. */
int nfftb200_comm_unique_id(void* out128);
/* host logic of the node sharding: tile-aligned cut of the sorted node list into nranks ranges of ~M/nranks
 * nodes; tile_start has ntiles+1 prefix sums, out receives nranks+1 tile boundaries */
int nfftb200_partition_tiles(const int64_t* tile_start, int64_t ntiles, int nranks, int64_t* out);
int nfftb200_comm_init(nfftb200_plan* p, const void* nccl_unique_id, int rank, int nranks, int mode);
/* bit 0: the node-sharded adjoint of this plan runs the fused spread + slab gather over peer memory (3-D tiled
 * plans whose slabs are whole tile layers, <= 8 ranks, CUDA IPC available) instead of ncclReduceScatter;
 * bit 1: the forward interpolates straight from the ranks' z-slabs over peer memory instead of ncclAllGather */
int nfftb200_comm_is_fused(nfftb200_plan* p);

/* ---- sampling density compensation ------------------------------------------------------------------
 * NFFTTools.sdc(p; iters) (NFFTTools/src/samplingDensity.jl:59-155): `iters` Pipe-Menon iterations on real data
 * through convolve_transpose!/convolve!, then the global scaling c = real(sum v)/sum|v|^2, v = A' D(w) A 1.
 * Everything runs on the device; weights receives M values of T (host or device).  Plan with ntransforms = 1.
 * Returns NFFTB200_BAD_ARGUMENT with "non-positive weights" where the reference throws that string. */
int nfftb200_sdc(nfftb200_plan* p, int iters, void* weights, int where);

/* ---- Toeplitz (Gram) operator, the iterative-reconstruction caller either side of the path --------------
 * nfftb200_toeplitz_kernel: calculateToeplitzKernel!(f, p, tr, fftplan) after its nodes!(p, tr)
 * (NFFTTools/src/Toeplitz.jl:131-137, and :86-93): lambda = FFT(fftshift(adjoint(p) * ones)).  The plan's image
 * size is 2*shape and it needs ntransforms = 1; lambda receives prod(p.N) complex values, host or device. */
int nfftb200_toeplitz_kernel(nfftb200_plan* p, void* lambda, int where);

/* convolveToeplitzKernel!(y, lambda, fftplan, ifftplan, xOS1, xOS2) (NFFTTools/src/Toeplitz.jl:230-244): the
 * operator object owns the two cuFFT plans and the oversampled work array.  shape = size(y) (D entries),
 * lambda has size 2*shape; ntransforms > 1 applies the same kernel to y of size (shape..., B) (batch slowest).
 * apply overwrites y with crop(IFFT(lambda .* FFT(pad(y)))), IFFT normalised like plan_ifft. */
typedef struct nfftb200_toeplitz nfftb200_toeplitz;
int nfftb200_toeplitz_create(nfftb200_toeplitz** out, int D, const int64_t* shape, int dtype, int ntransforms,
                             int device);
int nfftb200_toeplitz_set_kernel(nfftb200_toeplitz* t, const void* lambda, int where);
int nfftb200_toeplitz_apply(nfftb200_toeplitz* t, void* y, int where);
int nfftb200_toeplitz_set_stream(nfftb200_toeplitz* t, void* cuda_stream);
int nfftb200_toeplitz_sync(nfftb200_toeplitz* t);
int nfftb200_toeplitz_destroy(nfftb200_toeplitz* t);
const char* nfftb200_toeplitz_last_error(nfftb200_toeplitz* t);

const char* nfftb200_last_error(nfftb200_plan* p);
const char* nfftb200_status_string(int status);
int nfftb200_version(void);

#ifdef __cplusplus
}
#endif
#endif /* NFFTB200_H */
