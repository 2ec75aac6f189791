// oned.cu -- 1-D kernels (D = 1; BASELINE config C4: N = 2^22, M = 2^25, Float64, ~4 nodes per grid cell).
//
//  * k_spread_1d: OUTPUT-STATIONARY adjoint gridding.  One CTA per reference tile (blockSize cells).  The tile's
//    nodes are bucketed by grid cell in shared memory (counting sort; the stable order inside a cell is restored
//    with a rank by original position, so the summation order is deterministic).  Then every thread owns ONE
//    grid cell of the tile, keeps its accumulator in registers and, for each of the 2m taps l, walks the nodes of
//    cell u + m - 1 - l and adds w_l(frac) * fHat.  The number of (node, tap) pairs is exactly the reference's
//    (2m per node, /root/reference/src/convolution.jl:465-492); there is no read-modify-write, no atomic, no
//    memset: each grid cell is written once.  Neighbour tiles' boundary nodes are read through the same buckets
//    (halo of m cells on each side, periodic).
//  * k_interp_1d: thread per node, weights in registers, 2m loads from the (L1/L2-resident) grid.
#include <algorithm>
#include <type_traits>

#include "common.cuh"
#include "tile3d.cuh"
#include "window.cuh"

namespace {

constexpr int O1_THREADS = 256;

// Nodes relevant to a tile: those with cell c in [t0 - m, t0 + bs + m - 1) (periodic).  They live in up to three
// contiguous sorted ranges (previous tile's tail cells, own tile, next tile's head cells); since nodes are only
// sorted by tile, the neighbour tiles are scanned completely and filtered by cell.
template <typename T, int MT, bool CPLX>
__global__ void __launch_bounds__(O1_THREADS)
k_spread_1d(const void* __restrict__ fhat_, void* __restrict__ g_, const T* __restrict__ xs,
            const int32_t* __restrict__ perm, const int32_t* __restrict__ tile_start, int tile_lo, int tile_hi,
            long long M, GeomDev geo, WinDev<T> win, const __grid_constant__ PolyParam<T, MT> pp, int cap, int S)
{
    using C = typename Cplx<T>::type;
    using V = typename std::conditional<CPLX, C, T>::type;
    constexpr int L = 2 * MT;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int bs = geo.bs[0], Nt = geo.Nt[0], nb = geo.nb[0];
    const int W = NFFTB_G1D + L;                           // bucketed cells: [u0 - m, u0 + G + m) of the sub-block
    int* cnt = reinterpret_cast<int*>(smem_raw);           // [W + 1] counts -> offsets
    int* fill = cnt + W + 1;                               // [W]
    const size_t head = (sizeof(int) * (size_t)(2 * W + 1) + 15) & ~(size_t)15;
    V* s_v = reinterpret_cast<V*>(smem_raw + head);        // [cap] node values
    T* s_t = reinterpret_cast<T*>(s_v + cap);              // [cap] frac - 1/2 (POLYNOMIAL) or frac + m - 1
    int* s_i = reinterpret_cast<int*>(s_t + cap);          // [cap] sorted position (restores a stable order)

    const int tile = blockIdx.x / S, sub = blockIdx.x - tile * S;
    const int tlen = min(bs, Nt - tile * bs);              // core length of the tile
    if (sub * NFFTB_G1D >= tlen) return;
    const int t0 = tile * bs + sub * NFFTB_G1D;            // first cell of this sub-block
    const int len = min(NFFTB_G1D, tlen - sub * NFFTB_G1D);
    const int b = blockIdx.y;
    const V* fhat = reinterpret_cast<const V*>(fhat_) + (long long)b * M;
    V* g = reinterpret_cast<V*>(g_) + (long long)b * geo.gsz;

    const int tp = tile == 0 ? nb - 1 : tile - 1, tn = tile == nb - 1 ? 0 : tile + 1;
    auto in_range = [&](int t) { return t >= tile_lo && t < tile_hi; };
    // candidate ranges (sorted positions); a tile that is its own neighbour (nb <= 2) is visited once
    int r_lo[3], r_hi[3], nr = 0;
    const bool need_p = sub * NFFTB_G1D - MT < 0, need_n = sub * NFFTB_G1D + len + MT - 2 >= tlen;
    if (in_range(tile)) { r_lo[nr] = tile_start[tile]; r_hi[nr] = tile_start[tile + 1]; nr++; }
    if (need_p && tp != tile && in_range(tp)) { r_lo[nr] = tile_start[tp]; r_hi[nr] = tile_start[tp + 1]; nr++; }
    if (need_n && tn != tile && !(need_p && tn == tp) && in_range(tn)) { r_lo[nr] = tile_start[tn]; r_hi[nr] = tile_start[tn + 1]; nr++; }

    // bucket index of a node cell c relative to this tile, or -1 (periodic distance, window [-m, len + m - 1))
    auto bucket = [&](int c) {
        int d = c - t0;
        if (d >= Nt - MT) d -= Nt;                         // wrapped from the left neighbour
        if (d < -MT) d += Nt;                              // wrapped from the right neighbour
        return (d >= -MT && d <= len + MT - 2) ? d + MT : -1;
    };

    for (int q = threadIdx.x; q <= W; q += O1_THREADS) cnt[q] = 0;
    __syncthreads();
    // pass 1: count
    for (int r = 0; r < nr; r++)
        for (int i = r_lo[r] + threadIdx.x; i < r_hi[r]; i += O1_THREADS) {
            T ks;
            const int bk = bucket(node_cell<T>(xs[i], Nt, ks));
            if (bk >= 0) atomicAdd(&cnt[bk + 1], 1);
        }
    __syncthreads();
    if (threadIdx.x < 32) {                                // inclusive scan of cnt[0..W] by the first warp
        int carry = 0;
        for (int base = 0; base <= W; base += 32) {
            const int q = base + threadIdx.x;
            int v = q <= W ? cnt[q] : 0;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, v, o); if ((int)threadIdx.x >= o) v += t; }
            if (q <= W) cnt[q] = carry + v;
            carry += __shfl_sync(0xffffffffu, v, 31);
        }
    }
    __syncthreads();
    const int total = cnt[W];
    for (int q = threadIdx.x; q < W; q += O1_THREADS) fill[q] = 0;
    __syncthreads();
    const bool fits = total <= cap;
    if (fits) {
        // pass 2: scatter (order inside a bucket is arbitrary here, fixed up below by sorting on s_i)
        for (int r = 0; r < nr; r++)
            for (int i = r_lo[r] + threadIdx.x; i < r_hi[r]; i += O1_THREADS) {
                T ks;
                const int c = node_cell<T>(xs[i], Nt, ks);
                const int bk = bucket(c);
                if (bk >= 0) {
                    const int pos = cnt[bk] + atomicAdd(&fill[bk], 1);
                    const int off = c - MT + 1;
                    const T d0 = sub_rn(ks, (T)off);
                    // POLYNOMIAL: frac - 1/2 (evalpoly argument); otherwise keep d0 = frac + m - 1
                    s_t[pos] = (win.mode == NFFTB200_POLYNOMIAL) ? sub_rn(add_rn(sub_rn(d0, (T)MT), (T)1), (T)0.5) : d0;
                    s_v[pos] = fhat[perm[i]];
                    s_i[pos] = i;
                }
            }
        __syncthreads();
        // restore ascending sorted position inside every bucket (insertion sort; buckets hold a handful of nodes)
        for (int q = threadIdx.x; q < W; q += O1_THREADS) {
            const int a = cnt[q], e = cnt[q + 1];
            for (int x = a + 1; x < e; x++) {
                const int ki = s_i[x]; const T kt = s_t[x]; const V kv = s_v[x];
                int y = x - 1;
                while (y >= a && s_i[y] > ki) { s_i[y + 1] = s_i[y]; s_t[y + 1] = s_t[y]; s_v[y + 1] = s_v[y]; y--; }
                s_i[y + 1] = ki; s_t[y + 1] = kt; s_v[y + 1] = kv;
            }
        }
        __syncthreads();
    }
    // gather: thread u owns grid cell t0 + u
    for (int u = threadIdx.x; u < len; u += O1_THREADS) {
        T ax = 0, ay = 0;
        if (fits) {
#pragma unroll
            for (int l = 0; l < L; l++) {
                // tap l of a node in cell c hits cell c - m + 1 + l  =>  c = u + m - 1 - l, bucket = c - t0 + m
                const int bk = u + 2 * MT - 1 - l;
                for (int x = cnt[bk]; x < cnt[bk + 1]; x++) {
                    const T t = s_t[x];
                    T w;
                    if (win.mode == NFFTB200_POLYNOMIAL) {
                        constexpr int deg = L + 1;
                        w = pp.c[l * deg + deg - 1];
#pragma unroll
                        for (int r = deg - 2; r >= 0; r--) w = tfma(w, t, pp.c[l * deg + r]);
                    } else if (win.mode == NFFTB200_LINEAR) {
                        const T idx = mul_rn(t, (T)win.lin_scale);
                        const int ii = (int)idx;
                        const T alpha = sub_rn(idx, (T)ii);
                        int a1 = ii - l * win.lin_scale, a2 = a1 + 1;
                        a1 = a1 < 0 ? -a1 : a1; a2 = a2 < 0 ? -a2 : a2;
                        const T v1 = win.lin[a1], v2 = win.lin[a2];
                        w = add_rn(v1, mul_rn(alpha, sub_rn(v2, v1)));
                    } else {
                        w = kb_exact<T>(sub_rn(t, (T)l), MT, win.b);
                    }
                    if constexpr (CPLX) { ax = tfma(w, s_v[x].x, ax); ay = tfma(w, s_v[x].y, ay); }
                    else ax = tfma(w, s_v[x], ax);
                }
            }
        }
        if constexpr (CPLX) g[t0 + u] = make_c<T>(ax, ay); else g[t0 + u] = ax;
    }
}


// Cell-ordered form of the output-stationary spreader (round 2): the plan keeps the nodes sorted by grid cell
// (sort.cu: cells1d_impl), so the nodes that reach a sub-block of NFFTB_G1D cells -- cells [t0 - m, t0 + len + m - 1),
// periodic -- are at most two contiguous ranges of (xs2, perm2).  They are copied into shared memory with coalesced
// loads (no counting sort, no shared atomics, no insertion sort, no scan of whole neighbour tiles), `off` holds the start
// of every cell inside the staged list, and the gather loop is the one of k_spread_1d: thread u owns grid cell t0 + u.
// [n_lo, n_hi) restricts the nodes to a rank's range under node sharding.
template <typename T, int MT, bool CPLX>
__global__ void __launch_bounds__(O1_THREADS)
k_spread_1d_cells(const void* __restrict__ fhat_, void* __restrict__ g_, const T* __restrict__ xs2,
                  const int32_t* __restrict__ perm2, const int32_t* __restrict__ cell_start, int n_lo, int n_hi,
                  long long M, GeomDev geo, WinDev<T> win, const __grid_constant__ PolyParam<T, MT> pp, int cap)
{
    using C = typename Cplx<T>::type;
    using V = typename std::conditional<CPLX, C, T>::type;
    constexpr int L = 2 * MT;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int Nt = geo.Nt[0];
    constexpr int W = NFFTB_G1D + L;                       // window cells [t0 - m, t0 + G + m) of the sub-block (+1 spare)
    int* off = reinterpret_cast<int*>(smem_raw);           // [W + 1] start of window cell q inside the staged list
    const size_t head = (sizeof(int) * (size_t)(W + 2) + 15) & ~(size_t)15;
    V* s_v = reinterpret_cast<V*>(smem_raw + head);        // [cap] node values
    T* s_t = reinterpret_cast<T*>(s_v + cap);              // [cap] frac - 1/2 (POLYNOMIAL) or frac + m - 1

    const int t0 = blockIdx.x * NFFTB_G1D;                 // sub-blocks tile the whole grid (tiles are runs of cells)
    const int len = min(NFFTB_G1D, Nt - t0);
    const int b = blockIdx.y;
    const V* fhat = reinterpret_cast<const V*>(fhat_) + (long long)b * M;
    V* g = reinterpret_cast<V*>(g_) + (long long)b * geo.gsz;
    const int nw = len + L - 1;                            // window cells that can reach this sub-block
    // window cell q <-> grid cell (t0 - m + q) mod Nt; its nodes are [cs(c), cs(c + 1)) clipped to [n_lo, n_hi)
    auto clip = [&](int v) { return min(max(v, n_lo), n_hi); };
    // the window is [a0, a0 + nw) with a0 = t0 - m; split at the periodic boundary into <= 2 runs of cells
    const int a0 = t0 - MT;
    int run_c[2], run_n[2], nrun = 0;
    if (a0 < 0) { run_c[nrun] = a0 + Nt; run_n[nrun] = min(-a0, nw); nrun++; if (nw > -a0) { run_c[nrun] = 0; run_n[nrun] = nw + a0; nrun++; } }
    else if (a0 + nw > Nt) { run_c[nrun] = a0; run_n[nrun] = Nt - a0; nrun++; run_c[nrun] = 0; run_n[nrun] = a0 + nw - Nt; nrun++; }
    else { run_c[nrun] = a0; run_n[nrun] = nw; nrun++; }
    int base[2], cnt[2];
    for (int r = 0; r < nrun; r++) {
        base[r] = clip(cell_start[run_c[r]]);
        cnt[r] = clip(cell_start[run_c[r] + run_n[r]]) - base[r];
    }
    const int total = cnt[0] + (nrun > 1 ? cnt[1] : 0);
    const bool fits = total <= cap;
    // offsets of the window cells in the staged list
    for (int q = threadIdx.x; q <= nw; q += O1_THREADS) {
        int r = 0, qq = q, pre = 0;
        if (nrun > 1 && q >= run_n[0]) { r = 1; qq = q - run_n[0]; pre = cnt[0]; }
        off[q] = pre + clip(cell_start[run_c[r] + qq]) - base[r];
    }
    if (fits) {
        for (int r = 0, pre = 0; r < nrun; pre += cnt[r], r++)
            for (int x = threadIdx.x; x < cnt[r]; x += O1_THREADS) {
                const long long i = (long long)base[r] + x;
                T ks;
                const int c = node_cell<T>(xs2[i], Nt, ks);
                const T d0 = sub_rn(ks, (T)(c - MT + 1));
                s_t[pre + x] = (win.mode == NFFTB200_POLYNOMIAL) ? sub_rn(add_rn(sub_rn(d0, (T)MT), (T)1), (T)0.5) : d0;
                s_v[pre + x] = fhat[perm2[i]];
            }
    }
    __syncthreads();
    for (int u = threadIdx.x; u < len; u += O1_THREADS) {
        T ax = 0, ay = 0;
        if (fits) {
#pragma unroll
            for (int l = 0; l < L; l++) {
                const int bk = u + 2 * MT - 1 - l;         // window cell whose tap l lands on grid cell t0 + u
                for (int x = off[bk]; x < off[bk + 1]; x++) {
                    const T t = s_t[x];
                    T w;
                    if (win.mode == NFFTB200_POLYNOMIAL) {
                        constexpr int deg = L + 1;
                        w = pp.c[l * deg + deg - 1];
#pragma unroll
                        for (int r = deg - 2; r >= 0; r--) w = tfma(w, t, pp.c[l * deg + r]);
                    } else if (win.mode == NFFTB200_LINEAR) {
                        const T idx = mul_rn(t, (T)win.lin_scale);
                        const int ii = (int)idx;
                        const T alpha = sub_rn(idx, (T)ii);
                        int a1 = ii - l * win.lin_scale, a2 = a1 + 1;
                        a1 = a1 < 0 ? -a1 : a1; a2 = a2 < 0 ? -a2 : a2;
                        const T v1 = win.lin[a1], v2 = win.lin[a2];
                        w = add_rn(v1, mul_rn(alpha, sub_rn(v2, v1)));
                    } else {
                        w = kb_exact<T>(sub_rn(t, (T)l), MT, win.b);
                    }
                    if constexpr (CPLX) { ax = tfma(w, s_v[x].x, ax); ay = tfma(w, s_v[x].y, ay); }
                    else ax = tfma(w, s_v[x], ax);
                }
            }
        }
        if constexpr (CPLX) g[t0 + u] = make_c<T>(ax, ay); else g[t0 + u] = ax;
    }
}

template <typename T, int MT, bool CPLX>
__global__ void __launch_bounds__(256)
k_interp_1d(const void* __restrict__ g_, void* __restrict__ fhat_, const T* __restrict__ xs,
            const int32_t* __restrict__ perm, long long i_lo, long long i_hi, long long M, GeomDev geo,
            WinDev<T> win, const __grid_constant__ PolyParam<T, MT> pp, int B)
{
    using C = typename Cplx<T>::type;
    using V = typename std::conditional<CPLX, C, T>::type;
    constexpr int L = 2 * MT;
    const int Nt = geo.Nt[0];
    for (long long i = i_lo + blockIdx.x * (long long)blockDim.x + threadIdx.x; i < i_hi; i += (long long)gridDim.x * blockDim.x) {
        T ks;
        const int c = node_cell<T>(xs[i], Nt, ks);
        T w[L];
        eval_taps<T, MT>(win, pp, ks, c, w);
        int cell = c - MT + 1;
        cell = cell < 0 ? cell + Nt : cell;
        const long long j = perm ? (long long)perm[i] : i;       // perm == nullptr: results in plan (cell) order
        for (int b = 0; b < B; b++) {
            const V* g = reinterpret_cast<const V*>(g_) + (long long)b * geo.gsz;
            T ax = 0, ay = 0;
            int cc = cell;
#pragma unroll
            for (int l = 0; l < L; l++) {
                if constexpr (CPLX) { const C v = g[cc]; ax = tfma(w[l], v.x, ax); ay = tfma(w[l], v.y, ay); }
                else ax = tfma(w[l], g[cc], ax);
                cc = (cc + 1 == Nt) ? 0 : cc + 1;
            }
            if constexpr (CPLX) reinterpret_cast<C*>(fhat_)[(long long)b * M + j] = make_c<T>(ax, ay);
            else reinterpret_cast<T*>(fhat_)[(long long)b * M + j] = ax;
        }
    }
}

// Second pass of the two-pass forward (large node sets): caller index j <- position inv[j] of the cell order, a random
// 16-byte GATHER with coalesced stores.  Measured on C4 (2^25 Float64 nodes): the direct form spends 1.0 of its 1.46 ms
// in the random 16-byte SCATTER fHat[perm[i]] (435 us with sorted nodes, i.e. an identity permutation): every
// partial-sector write is a read-modify-write.  Interpolating into plan order (coalesced) and gathering afterwards costs
// 0.43 + 0.83 = 1.26 ms.  A version bucketed by the high bits of the caller index (8-bit radix partition at plan time,
// interpolator writes into the bucketed order, second pass scatters inside one bucket's 1/256 of fHat) was measured too:
// 1.66 ms -- the cost is per 16-byte access (about 40 G accesses/s either way), not DRAM locality.
template <typename V>
__global__ void __launch_bounds__(256)
k_unpermute_1d(const V* __restrict__ tmp, V* __restrict__ fhat, const int32_t* __restrict__ inv, long long M, int B)
{
    for (long long j = blockIdx.x * (long long)blockDim.x + threadIdx.x; j < M; j += (long long)gridDim.x * blockDim.x) {
        const long long q = inv[j];
        for (int b = 0; b < B; b++) fhat[(long long)b * M + j] = tmp[(long long)b * M + q];
    }
}
__global__ void __launch_bounds__(256) k_invert_perm_1d(const int32_t* __restrict__ perm2, int32_t* __restrict__ inv, long long M)
{
    for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < M; q += (long long)gridDim.x * blockDim.x) inv[perm2[q]] = (int32_t)q;
}

template <typename T, int MT, bool CPLX>
int spread1d_launch(nfftb200_plan* p, const void* fhat, void* g, int B, int t_lo, int t_hi, int cap)
{
    using C = typename Cplx<T>::type;
    const int W = NFFTB_G1D + 2 * MT;
    const int S = (int)((p->bs[0] + NFFTB_G1D - 1) / NFFTB_G1D);
    const size_t vsz = CPLX ? sizeof(C) : sizeof(T);
    const size_t smem = ((sizeof(int) * (size_t)(2 * W + 1) + 15) & ~(size_t)15) + (sizeof(T) + vsz + sizeof(int)) * (size_t)cap + 16;
    auto kern = k_spread_1d<T, MT, CPLX>;
    CUDA_TRY(p, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((unsigned)(p->ntiles * S), B);
    kern<<<grid, O1_THREADS, smem, p->stream>>>(fhat, g, (const T*)p->d_xs, p->d_perm, p->d_tile_start, t_lo, t_hi,
                                               p->M, make_geom<T>(p), make_win<T>(p), make_poly_param<T, MT>(p), cap, S);
    p->launches++;
    CUDA_TRY(p, cudaGetLastError());
    return NFFTB200_OK;
}

template <typename T, int MT, bool CPLX>
int spread1d_cells_launch(nfftb200_plan* p, const void* fhat, void* g, int B, int t_lo, int t_hi, int cap)
{
    using C = typename Cplx<T>::type;
    constexpr int W = NFFTB_G1D + 2 * MT;
    const size_t vsz = CPLX ? sizeof(C) : sizeof(T);
    const size_t smem = ((sizeof(int) * (size_t)(W + 2) + 15) & ~(size_t)15) + (sizeof(T) + vsz) * (size_t)cap + 16;
    auto kern = k_spread_1d_cells<T, MT, CPLX>;
    CUDA_TRY(p, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((unsigned)((p->Nt[0] + NFFTB_G1D - 1) / NFFTB_G1D), B);
    kern<<<grid, O1_THREADS, smem, p->stream>>>(fhat, g, (const T*)p->d_xs2, p->d_perm2, p->d_bin_start, p->h_tile_start[(size_t)t_lo],
                                               p->h_tile_start[(size_t)t_hi], p->M, make_geom<T>(p), make_win<T>(p),
                                               make_poly_param<T, MT>(p), cap);
    p->launches++;
    CUDA_TRY(p, cudaGetLastError());
    return NFFTB200_OK;
}

}  // namespace

// returns -1 when the 1-D kernels do not apply (then the generic kernels run)
template <typename T>
static int spread1d_impl(nfftb200_plan* p, const void* fhat, void* g, int B, int is_complex, int t_lo, int t_hi)
{
    if (p->m < 2 || p->m > 6 || p->bs[0] < p->m || p->Nt[0] < p->bs[0] + 2 * p->m) return -1;
    // every tile core must be at least m long so that only the immediate neighbours reach into a tile
    if (p->Nt[0] - (p->nb[0] - 1) * p->bs[0] < p->m) return -1;
    // shared-memory capacity: the largest per-tile neighbourhood, counted at nodes! (k_count_neigh_1d)
    const int64_t worst = p->max_neigh_1d;
    const size_t vsz = (is_complex ? 2 : 1) * sizeof(T);
    const int64_t W = NFFTB_G1D + 2 * p->m;
    const int64_t cap_max = ((int64_t)200 * 1024 - (int64_t)sizeof(int) * (2 * W + 4)) / (int64_t)(sizeof(T) + vsz + sizeof(int)) - 4;
    if (worst > cap_max) return -1;
    const int cap = (int)std::max<int64_t>(worst, 32);
    // default: the cell-ordered form (plan-time sort by grid cell); kernel_mode 9 keeps the round-1 bucketing kernel.
    // The window of a sub-block is the same either way, so the capacity counted at nodes! (max_neigh_1d) applies.
    if (p->kernel_mode != 9 && p->bs[0] % NFFTB_G1D == 0 && nfftb_ensure_cells_1d(p) == NFFTB200_OK) {
#define GOC(MM)                                                                                        \
    case MM:                                                                                           \
        return is_complex ? spread1d_cells_launch<T, MM, true>(p, fhat, g, B, t_lo, t_hi, cap)         \
                          : spread1d_cells_launch<T, MM, false>(p, fhat, g, B, t_lo, t_hi, cap);
        switch (p->m) { GOC(2) GOC(3) GOC(4) GOC(5) GOC(6) default: break; }
#undef GOC
    }
#define GO(MM)                                                                                         \
    case MM:                                                                                           \
        return is_complex ? spread1d_launch<T, MM, true>(p, fhat, g, B, t_lo, t_hi, cap)               \
                          : spread1d_launch<T, MM, false>(p, fhat, g, B, t_lo, t_hi, cap);
    switch (p->m) { GO(2) GO(3) GO(4) GO(5) GO(6) default: return -1; }
#undef GO
}

template <typename T>
static int interp1d_impl(nfftb200_plan* p, const void* g, void* fhat, int B, int is_complex, long long i_lo, long long i_hi)
{
    if (p->m < 2 || p->m > 6 || p->Nt[0] < 2 * p->m) return -1;
    const long long n = i_hi - i_lo;
    const int blocks = (int)std::min<long long>((n + 255) / 256, 148 * 16);
    // cell-ordered nodes (plan-time, sort.cu): neighbouring threads read neighbouring grid cells -- coalesced instead of
    // 32 cache lines per load; the range [i_lo, i_hi) is tile aligned, so it names the same nodes in either order
    const bool cells = p->kernel_mode != 9 && nfftb_ensure_cells_1d(p) == NFFTB200_OK;
    const T* xs_use = (const T*)(cells ? p->d_xs2 : p->d_xs);
    const int32_t* perm_use = cells ? p->d_perm2 : p->d_perm;
    // two passes when one transform's results exceed what L2 can merge (64 MiB) and the whole node set is processed:
    // interpolate into a plan-ordered buffer (coalesced stores), then gather it into the caller's order
    const size_t vsz = is_complex ? 2 * sizeof(T) : sizeof(T);
    bool two_pass = cells && i_lo == 0 && i_hi == p->M && (size_t)p->M * vsz >= ((size_t)64 << 20);
    void* out = fhat;
    if (two_pass) {
        const int64_t need = (int64_t)((size_t)p->M * vsz * (size_t)B);
        if (need > p->cap_tilebuf) {
            if (p->d_tilebuf) cudaFree(p->d_tilebuf);
            p->d_tilebuf = nullptr; p->cap_tilebuf = 0;
            if (cudaMalloc(&p->d_tilebuf, (size_t)need) != cudaSuccess) { cudaGetLastError(); two_pass = false; }
            else p->cap_tilebuf = need;
        }
        if (two_pass && !p->have_inv2) {
            if (p->M > p->cap_inv2) {
                if (p->d_inv2) cudaFree(p->d_inv2);
                p->d_inv2 = nullptr; p->cap_inv2 = 0;
                if (cudaMalloc((void**)&p->d_inv2, (size_t)p->M * 4) != cudaSuccess) { cudaGetLastError(); two_pass = false; }
                else p->cap_inv2 = p->M;
            }
            if (two_pass) {
                k_invert_perm_1d<<<148 * 16, 256, 0, p->stream>>>(p->d_perm2, p->d_inv2, p->M);
                p->launches++;
                p->have_inv2 = true;
            }
        }
        if (two_pass) { out = p->d_tilebuf; perm_use = nullptr; }
    }
#define GO(MM)                                                                                                   \
    case MM:                                                                                                     \
        if (is_complex) k_interp_1d<T, MM, true><<<blocks, 256, 0, p->stream>>>(g, out, xs_use, perm_use, i_lo, i_hi, p->M, make_geom<T>(p), make_win<T>(p), make_poly_param<T, MM>(p), B); \
        else k_interp_1d<T, MM, false><<<blocks, 256, 0, p->stream>>>(g, out, xs_use, perm_use, i_lo, i_hi, p->M, make_geom<T>(p), make_win<T>(p), make_poly_param<T, MM>(p), B); \
        break;
    switch (p->m) { GO(2) GO(3) GO(4) GO(5) GO(6) default: return -1; }
#undef GO
    p->launches++;
    if (two_pass) {
        using C = typename Cplx<T>::type;
        if (is_complex) k_unpermute_1d<C><<<148 * 16, 256, 0, p->stream>>>((const C*)out, (C*)fhat, p->d_inv2, p->M, B);
        else k_unpermute_1d<T><<<148 * 16, 256, 0, p->stream>>>((const T*)out, (T*)fhat, p->d_inv2, p->M, B);
        p->launches++;
    }
    CUDA_TRY(p, cudaGetLastError());
    return NFFTB200_OK;
}

int nfftb_spread_1d(nfftb200_plan* p, const void* fhat, void* g, int B, int is_complex, int t_lo, int t_hi)
{
    return p->dtype == NFFTB200_F32 ? spread1d_impl<float>(p, fhat, g, B, is_complex, t_lo, t_hi)
                                    : spread1d_impl<double>(p, fhat, g, B, is_complex, t_lo, t_hi);
}
int nfftb_interp_1d(nfftb200_plan* p, const void* g, void* fhat, int B, int is_complex, long long i_lo, long long i_hi)
{
    return p->dtype == NFFTB200_F32 ? interp1d_impl<float>(p, g, fhat, B, is_complex, i_lo, i_hi)
                                    : interp1d_impl<double>(p, g, fhat, B, is_complex, i_lo, i_hi);
}
