// common.cuh -- plan state, error plumbing and device helpers shared by all kernels of
// libnfftb200.so (B200 / sm_100a).  See DESIGN.md for the data layout in HBM.
#pragma once

#include <cuda_runtime.h>
#include <cufft.h>
#include <stdint.h>
#include <string>
#include <vector>

#include "../../include/nfftb200.h"

#define NFFTB_MAX_D 4            // D = 4 runs on the generic kernels (test/accuracy.jl:43); the tiled kernels are 1-D / 2-D / 3-D
#define NFFTB_MAX_M 8            // taps per dim = 2m <= 16
#define NFFTB_G1D 256            // cells per CTA of the 1-D output-stationary spreader

// ---------------------------------------------------------------------------------------
// complex value types
// ---------------------------------------------------------------------------------------
template <typename T> struct Cplx;
template <> struct Cplx<float> { using type = float2; };
template <> struct Cplx<double> { using type = double2; };

template <typename T> __host__ __device__ inline typename Cplx<T>::type make_c(T a, T b);
template <> __host__ __device__ inline float2 make_c<float>(float a, float b) { return make_float2(a, b); }
template <> __host__ __device__ inline double2 make_c<double>(double a, double b) { return make_double2(a, b); }

// ---------------------------------------------------------------------------------------
// window parameters handed to the kernels by value
// ---------------------------------------------------------------------------------------
template <typename T> struct WinDev {
    int m;             // kernel half width; 2m taps per dimension
    int mode;          // NFFTB200_FULL / LINEAR / POLYNOMIAL (TENSOR is mapped to POLYNOMIAL)
    int lin_scale;     // LUTSize / m
    T b;               // Kaiser-Bessel shape parameter pi*(2-1/sigma) in T
    int window;        // NFFTB200_KAISER_BESSEL ... NFFTB200_COSH_TYPE (used by the FULL mode only)
    T beta;            // cosh_type: pi*m*(2-1/sigma) in T
    const T* poly;     // (2m+1) x 2m, column-major (column = tap)
    const T* lin;      // LUTSize + 2
};

struct GeomDev {
    int D;
    int Nt[NFFTB_MAX_D];      // oversampled grid (padded with 1)
    int N[NFFTB_MAX_D];       // image size (padded with 1)
    int bs[NFFTB_MAX_D];      // tile size (padded with 1)
    int nb[NFFTB_MAX_D];      // tiles per dim (padded with 1)
    unsigned inv_bs[NFFTB_MAX_D];   // ceil(2^32 / bs[d]) for division-free u / bs[d]
    long long gsz;            // prod(Nt)
    long long fsz;            // prod(N)
};

// tile scratch of every rank as seen from this process (node sharding over peer memory, comm.cu / spread.cu)
#define NFFTB_MAX_PEERS 8
struct PeerTab {
    const void* base[NFFTB_MAX_PEERS];   // rank r's scratch (own pointer or CUDA-IPC mapping)
    int cut[NFFTB_MAX_PEERS + 1];        // rank r owns tiles [cut[r], cut[r+1])
    int item_lo[NFFTB_MAX_PEERS];        // first work item of rank r (scratch slot 0)
    int n;                               // ranks
};

// z-slabs of the oversampled grid as seen from this process: plane z lives on rank z / planes at local plane z % planes
struct SlabTab {
    const void* base[NFFTB_MAX_PEERS];
    int planes;                          // planes per slab (Nt[2] / ranks)
    int n;                               // ranks (0 = not used)
};

// ---------------------------------------------------------------------------------------
// the plan (host side)
// ---------------------------------------------------------------------------------------
struct nfftb200_plan {
    int D = 0;
    int dtype = NFFTB200_F32;
    int m = 4;
    double sigma = 2.0;       // effective sigma = Nt[0]/N[0] rounded to T
    double reltol = 1e-9;
    double b = 0.0;           // window shape parameter as evaluated in T
    double beta = 0.0;        // cosh_type shape parameter pi*m*(2-1/sigma) as evaluated in T
    int window = NFFTB200_KAISER_BESSEL;
    int precompute = NFFTB200_POLYNOMIAL;
    int B = 1;                // ntransforms
    int device = 0;
    int64_t N[NFFTB_MAX_D] = {1, 1, 1, 1};
    int64_t Nt[NFFTB_MAX_D] = {1, 1, 1, 1};
    int64_t bs[NFFTB_MAX_D] = {1, 1, 1, 1};
    int64_t nb[NFFTB_MAX_D] = {1, 1, 1, 1};
    int64_t ntiles = 1;
    int64_t gsz = 1, fsz = 1;
    int64_t lut_size = 0;
    int64_t M = 0;
    bool have_nodes = false;
    int kernel_mode = 0;

    cudaStream_t stream = nullptr;
    bool own_stream = false;
    cufftHandle fft = 0;
    bool have_fft = false;
    // pruned 3-D FFT: 2-D FFTs only on the z-planes that hold image frequencies + 1-D FFT along z
    cufftHandle fft_xy = 0, fft_z = 0;
    bool have_pruned = false;
    cufftHandle fft_d4 = 0;            // D = 4: 1-D transform along the 4th dimension (cuFFT plans have rank <= 3)
    bool have_fft_d4 = false;
    int64_t zlo_planes = 0, zhi_planes = 0;     // planes [0, zlo) and [Nt2 - zhi, Nt2) are the non-zero ones

    // host copies of the tables (double), for get_table and for re-upload
    std::vector<double> h_hat_inv;   // concatenated over d
    std::vector<double> h_poly;      // (2m+1)*2m column-major
    std::vector<double> h_lin;       // LUTSize+2

    // device tables in T
    void* d_hat_inv = nullptr;
    void* d_poly = nullptr;
    void* d_lin = nullptr;

    // device grid: B * gsz complex T  (p.tmpVec)
    void* d_grid = nullptr;

    // node state (device)
    void* d_xs = nullptr;            // shifted nodes in sorted order, D x M of T
    int32_t* d_perm = nullptr;       // sorted position -> original node id
    int32_t* d_tile_start = nullptr; // ntiles + 1
    std::vector<int32_t> h_tile_start;  // host copy (tile-aligned sharding, launch ranges)
    // work items: a tile with more than item_cap nodes is split (after the sort) into several items, each a
    // (tile, node range) triple processed by its own CTA; d_tile_items[t]..[t+1] are tile t's items
    int32_t* d_items = nullptr;      // 3 * nitems: tile, n_lo, n_hi
    int32_t* d_item_stride = nullptr;// nitems: 1 = the item is the contiguous node range [n_lo, n_hi); s > 1 (2-D plans) = it takes
                                     // every s-th node n_lo, n_lo + s, ... < n_hi, so that the items of a crowded tile each
                                     // see a uniform sample of its nodes (a contiguous slice of a radial trajectory is a
                                     // wedge that lands in a few of the spreader's warp-private sub-tiles)
    int32_t* d_tile_items = nullptr; // ntiles + 1
    std::vector<int32_t> h_tile_items;
    int64_t nitems = 0, cap_items = 0;
    int64_t cap_nodes = 0;
    int max_neigh_1d = 0;            // 1-D: max nodes a tile's output-stationary spreader must bucket
    // second, finer plan-time order used by the register-window kernels of kernel_mode 8 (spread_lean.cuh /
    // interp_lean.cuh): inside every tile the nodes are grouped by (owning warp = octant of bins, colour = bin within
    // the octant); the reported permutation (d_perm, the bit-exact contract) is untouched.  Built lazily (sort.cu).
    void* d_xs2 = nullptr;           // shifted nodes in (tile, bin) order, D x M
    int32_t* d_perm2 = nullptr;      // (tile, bin)-sorted position -> caller's node id
    int32_t* d_bin_start = nullptr;  // ntiles * NQ + 1 absolute start positions
    bool have_bins = false;
    int32_t* d_inv2 = nullptr;       // 1-D: inverse of perm2 (caller index -> position in the cell order), built on first use
    bool have_inv2 = false;
    int64_t cap_inv2 = 0;
    int bins_nq = 0;                 // NQ = 8 * S^3 bins per tile the table was built for
    int64_t cap_bins_nodes = 0, cap_bin_tab = 0;
    // fused spread + gather (lean.cu): work items expected at every output block (the items of its distinct neighbour
    // tiles) and the per-exec arrival counters
    int32_t* d_expect = nullptr;     // ntiles
    int32_t* d_ready = nullptr;      // B * ntiles, zeroed before every fused spread
    std::vector<int32_t> h_expect;
    int64_t cap_ready = 0;
    int32_t* d_pair_items = nullptr; // cluster-pair experiment (kernel_mode 13): identity item table of the x-pairs of tiles
    int64_t cap_pair_items = 0;

    // sort scratch
    uint32_t* d_keys[2] = {nullptr, nullptr};
    int32_t* d_vals[2] = {nullptr, nullptr};
    uint32_t* d_hist = nullptr;
    int64_t cap_hist = 0;
    int* d_flag = nullptr;

    // staging buffers for host-pointer calls
    void* d_stage_f = nullptr;  int64_t cap_stage_f = 0;     // bytes
    void* d_stage_h = nullptr;  int64_t cap_stage_h = 0;     // bytes
    void* d_stage_k = nullptr;  int64_t cap_stage_k = 0;     // bytes
    void* d_stage_g = nullptr;  int64_t cap_stage_g = 0;     // bytes

    // asynchronous host-buffer mode (where = NFFTB200_HOST_ASYNC): uploads, compute and downloads run on three
    // streams chained by events, each direction with its own input and output staging buffer, so that the copies of
    // one call overlap the kernels of the neighbouring calls; nfftb200_sync waits for everything
    struct AsyncSlot { void* d = nullptr; int64_t cap = 0; cudaEvent_t free_ev = nullptr; bool used = false; };
    cudaStream_t s_up = nullptr, s_down = nullptr;
    cudaEvent_t e_up = nullptr, e_done = nullptr;
    AsyncSlot a_in[2], a_out[2];      // [0] forward, [1] adjoint

    // Toeplitz kernel construction (toeplitz.cu): image-sized cuFFT plan and work arrays, kept across calls
    cufftHandle fft_img = 0;
    bool have_fft_img = false;
    void* d_toep[3] = {nullptr, nullptr, nullptr};     // ones (M), image (fsz), shifted image (fsz)
    int64_t cap_toep[3] = {0, 0, 0};

    // timing
    bool timing = false;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    double t[7] = {0, 0, 0, 0, 0, 0, 0};
    int pending = 0;                 // 0 none, 1 forward, 2 adjoint
    cudaEvent_t evk[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};  // memset | spread stage | interp kernel | [5] gather start
    bool have_gather_ev = false;
    int pending_k = 0;               // bit 0: spread events valid, bit 1: interp events valid
    double tk[4] = {0, 0, 0, 0};     // spread stage (kernel + gather), interp kernel, grid memset, gather pass

    int64_t launches = 0;
    std::string err;

    // multi-GPU
    void* nccl_comm = nullptr;
    int rank = 0, nranks = 1, shard_mode = NFFTB200_SHARD_NONE;
    int64_t node_lo = 0, node_hi = 0;  // sorted-position range owned by this rank (SHARD_NODES)
    int b_lo = 0, b_hi = 1;            // transform range owned by this rank (SHARD_BATCH)
    void* d_slab = nullptr; int64_t cap_slab = 0;
    void* d_tilebuf = nullptr; int64_t cap_tilebuf = 0;   // per-tile padded sub-grids ("blocks") of the spreader

    size_t esz() const { return dtype == NFFTB200_F32 ? 4 : 8; }
};

// ---------------------------------------------------------------------------------------
// error plumbing
// ---------------------------------------------------------------------------------------
int nfftb_fail(nfftb200_plan* p, int code, const std::string& msg);

#define CUDA_TRY(p, call)                                                                  \
    do {                                                                                   \
        cudaError_t e__ = (call);                                                          \
        if (e__ != cudaSuccess)                                                            \
            return nfftb_fail((p), e__ == cudaErrorMemoryAllocation ? NFFTB200_OOM         \
                                                                    : NFFTB200_CUDA_ERROR, \
                              std::string(#call) + ": " + cudaGetErrorString(e__));        \
    } while (0)

#define CUFFT_TRY(p, call)                                                                 \
    do {                                                                                   \
        cufftResult r__ = (call);                                                          \
        if (r__ != CUFFT_SUCCESS)                                                          \
            return nfftb_fail((p), NFFTB200_CUDA_ERROR,                                    \
                              std::string(#call) + ": cufft error " + std::to_string((int)r__)); \
    } while (0)

#define ST_TRY(call)                      \
    do {                                  \
        int s__ = (call);                 \
        if (s__ != NFFTB200_OK) return s__; \
    } while (0)

template <typename T> inline GeomDev make_geom(const nfftb200_plan* p)
{
    GeomDev g;
    g.D = p->D;
    for (int d = 0; d < NFFTB_MAX_D; d++) {
        g.Nt[d] = (int)p->Nt[d]; g.N[d] = (int)p->N[d]; g.bs[d] = (int)p->bs[d]; g.nb[d] = (int)p->nb[d];
        g.inv_bs[d] = (unsigned)((0x100000000ull + (unsigned long long)p->bs[d] - 1) / (unsigned long long)p->bs[d]);
    }
    g.gsz = p->gsz; g.fsz = p->fsz;
    return g;
}

// the tiled kernels evaluate only the Kaiser-Bessel window exactly (FULL); the other windows' exact forms (Bessel I0,
// B-spline recursion, ...) live in the generic kernels so that the hot kernels carry no function call
inline bool nfftb_tiled_ok(const nfftb200_plan* p)
{
    return p->kernel_mode != 1 && !(p->precompute == NFFTB200_FULL && p->window != NFFTB200_KAISER_BESSEL);
}

// kernel modes that run the (tile, bin)-ordered register-window kernels of lean.cu where they apply: 0 = auto, 3 = auto
// without the TMA tensor-map load, 8 = explicit, 11 = fused spread + gather experiment, 12 = compact tile layout, 13 = cluster-pair DSMEM halo experiment
inline bool nfftb_lean_mode(const nfftb200_plan* p)
{
    return p->kernel_mode == 0 || p->kernel_mode == 3 || p->kernel_mode == 8 || (p->kernel_mode >= 11 && p->kernel_mode <= 13);
}

template <typename T> inline WinDev<T> make_win(const nfftb200_plan* p)
{
    WinDev<T> w;
    w.m = p->m;
    w.mode = (p->precompute == NFFTB200_TENSOR) ? NFFTB200_POLYNOMIAL : p->precompute;
    w.lin_scale = (int)(p->lut_size / p->m);
    w.b = (T)p->b;
    w.window = p->window;
    w.beta = (T)p->beta;
    w.poly = (const T*)p->d_poly;
    w.lin = (const T*)p->d_lin;
    return w;
}

// ---------------------------------------------------------------------------------------
// stage launchers implemented in the .cu files
// ---------------------------------------------------------------------------------------
int nfftb_sort_nodes(nfftb200_plan* p, const void* d_k);                       // sort.cu
int nfftb_ensure_bins(nfftb200_plan* p, int W, int G);
int nfftb_ensure_bins_2d(nfftb200_plan* p);                                      // sort.cu: 2-D order by 8 x 8-cell bin inside a tile (xs2, perm2)
int nfftb_ensure_cells_1d(nfftb200_plan* p);                                     // sort.cu: 1-D cell order (xs2, perm2, cell_start in d_bin_start)
// lean.cu (kernel_mode 8, Float32 3-D): -1 when the kernels do not apply; scratch_override / slabs select the node-sharded forms
int nfftb_spread_lean(nfftb200_plan* p, const void* fhat, void* g, void* scratch_override, int B, int t_lo, int t_hi);
int nfftb_interp_lean(nfftb200_plan* p, const void* g, void* fhat, int B, int t_lo, int t_hi, const SlabTab* slabs);
int nfftb_gather_scratch(nfftb200_plan* p, const void* scratch, void* g, int B, int t_lo, int t_hi, int item_lo, int item_hi,
                         const GeomDev* geo_override = nullptr, const int32_t* items_override = nullptr);                          // spread.cu                          // sort.cu: (tile, bin) order for kernel_mode 8
int nfftb_deconvolve(nfftb200_plan* p, const void* d_f, void* d_g, int B);      // deconv.cu
int nfftb_deconvolve_transpose(nfftb200_plan* p, const void* d_g, void* d_f, int B);
// t_lo/t_hi: half-open range of reference tiles ("blocks") whose nodes are processed
int nfftb_spread(nfftb200_plan* p, const void* d_fhat, void* d_g, int B, int is_complex,
                 int64_t t_lo, int64_t t_hi);                                   // spread.cu
int nfftb_interp(nfftb200_plan* p, const void* d_g, void* d_fhat, int B, int is_complex,
                 int64_t t_lo, int64_t t_hi);                                   // interp.cu
int nfftb_build_tables(nfftb200_plan* p);
size_t nfftb_spread3d_smem(int dtype, int m, const int64_t* bs);                // spread.cu
int nfftb_peer_tile_cells(nfftb200_plan* p);                                     // spread.cu (node sharding over peer memory)
int nfftb_peer_spread(nfftb200_plan* p, const void* d_fhat, void* d_scratch, int64_t t_lo, int64_t t_hi);
int nfftb_peer_gather(nfftb200_plan* p, void* d_slab, int layer_lo, int nlayers, const PeerTab& pt);
int nfftb_peer_interp(nfftb200_plan* p, const SlabTab& st, void* d_fhat, int64_t t_lo, int64_t t_hi);       // interp.cu
int nfftb_comm_after_nodes(nfftb200_plan* p);                                    // comm.cu
int nfftb_spread_2d(nfftb200_plan* p, const void* fhat, void* g, int B, int t_lo, int t_hi);                    // twod.cu
int nfftb_interp_2d(nfftb200_plan* p, const void* g, void* fhat, int B, int t_lo, int t_hi);
int nfftb_spread_1d(nfftb200_plan* p, const void* fhat, void* g, int B, int is_complex, int t_lo, int t_hi);   // oned.cu
int nfftb_interp_1d(nfftb200_plan* p, const void* g, void* fhat, int B, int is_complex, long long i_lo, long long i_hi);                                       // tables.cpp
