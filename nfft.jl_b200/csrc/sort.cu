// sort.cu -- K1 + K2: node range check, shift, tile key, and a *stable* LSD radix
// (counting) sort of node ids by tile key.  Output contract (bit-exact vs the reference):
//   perm = concat over tiles l (column-major tile order) of nodesInBlock[l], ascending j
//   inside a tile          -- /root/reference/src/precomputation.jl:487-504 (_precomputeBlocks)
//   key  = sum_d (unsafe_trunc(Int, k'[d,j]*Nt[d]) / blockSize[d]) * stride_d on the
//   shifted nodes k'       -- /root/reference/src/utils.jl:32-44, precomputation.jl:495
// The sort is 8 bits per pass; every pass is stable (warp match + per-warp running counts +
// per-CTA digit offsets from a global scan), so the composition is the stable counting sort
// the reference performs serially.
#include "common.cuh"
#include "window.cuh"

namespace {

constexpr int SORT_THREADS = 256;
constexpr int SORT_ITEMS = 16;
constexpr int SORT_TILE = SORT_THREADS * SORT_ITEMS;   // 4096 keys per CTA

template <typename T>
__global__ void k_node_keys(const T* __restrict__ k, long long M, GeomDev g,
                            uint32_t* __restrict__ keys, int* __restrict__ flag)
{
    long long j = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (; j < M; j += stride) {
        uint32_t key = 0, mul = 1;
        bool bad = false;
#pragma unroll
        for (int d = 0; d < NFFTB_MAX_D; d++) {
            if (d < g.D) {
                T v = k[j * g.D + d];
                if (!(fabs(v) <= (T)0.5)) bad = true;       // checkNodes (NaN fails too)
                v = shift_node<T>(v);
                T ks;
                int c = node_cell<T>(v, g.Nt[d], ks);
                c = min(max(c, 0), g.Nt[d] - 1);            // only reachable for rejected nodes
                key += (uint32_t)(c / g.bs[d]) * mul;
                mul *= (uint32_t)g.nb[d];
            }
        }
        if (bad) atomicOr(flag, 1);
        keys[j] = key;
    }
}

__global__ void k_radix_hist(const uint32_t* __restrict__ keys, long long M, int shift,
                             uint32_t* __restrict__ hist, int nCTA)
{
    __shared__ uint32_t h[256];
    h[threadIdx.x] = 0;
    __syncthreads();
    const long long base = (long long)blockIdx.x * SORT_TILE;
#pragma unroll
    for (int r = 0; r < SORT_ITEMS; r++) {
        long long idx = base + r * SORT_THREADS + threadIdx.x;
        if (idx < M) atomicAdd(&h[(keys[idx] >> shift) & 255u], 1u);
    }
    __syncthreads();
    hist[(long long)threadIdx.x * nCTA + blockIdx.x] = h[threadIdx.x];
}

// exclusive scan of n uint32 in place: (A) per-CTA scan of 4096-element chunks + chunk sums, (B) scan of the
// chunk sums by one CTA, (C) add the chunk offsets.  Coalesced accesses throughout.
constexpr int SCAN_CHUNK = 4096;     // 1024 threads x 4

__device__ __forceinline__ uint32_t block_exclusive_scan_1024(uint32_t v, uint32_t* warp_sums, uint32_t& total)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) warp_sums[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        uint32_t w = warp_sums[lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, w, o);
            if (lane >= o) w += t;
        }
        warp_sums[lane] = w;
    }
    __syncthreads();
    total = warp_sums[31];
    const uint32_t base = warp == 0 ? 0u : warp_sums[warp - 1];
    __syncthreads();
    return base + inc - v;
}

__global__ void __launch_bounds__(1024) k_scan_chunks(uint32_t* __restrict__ a, long long n, uint32_t* __restrict__ sums)
{
    __shared__ uint32_t ws[32];
    const long long base = (long long)blockIdx.x * SCAN_CHUNK + threadIdx.x * 4;
    uint32_t v[4];
#pragma unroll
    for (int k = 0; k < 4; k++) v[k] = (base + k < n) ? a[base + k] : 0u;
    const uint32_t mine = v[0] + v[1] + v[2] + v[3];
    uint32_t total;
    uint32_t ex = block_exclusive_scan_1024(mine, ws, total);
#pragma unroll
    for (int k = 0; k < 4; k++) { if (base + k < n) a[base + k] = ex; ex += v[k]; }
    if (threadIdx.x == 0) sums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(1024) k_scan_sums(uint32_t* __restrict__ sums, int n)
{
    __shared__ uint32_t ws[32];
    uint32_t carry = 0;
    for (int base = 0; base < n; base += 1024) {
        const int i = base + threadIdx.x;
        const uint32_t v = i < n ? sums[i] : 0u;
        uint32_t total;
        const uint32_t ex = block_exclusive_scan_1024(v, ws, total);
        if (i < n) sums[i] = carry + ex;
        carry += total;
        __syncthreads();
    }
}

__global__ void __launch_bounds__(1024) k_scan_add(uint32_t* __restrict__ a, long long n, const uint32_t* __restrict__ sums)
{
    const uint32_t off = sums[blockIdx.x];
    const long long base = (long long)blockIdx.x * SCAN_CHUNK + threadIdx.x * 4;
#pragma unroll
    for (int k = 0; k < 4; k++) if (base + k < n) a[base + k] += off;
}

__global__ void __launch_bounds__(SORT_THREADS)
k_radix_scatter(const uint32_t* __restrict__ keys_in, const int32_t* __restrict__ vals_in,
                uint32_t* __restrict__ keys_out, int32_t* __restrict__ vals_out, long long M,
                int shift, const uint32_t* __restrict__ hist, int nCTA)
{
    constexpr int NW = SORT_THREADS / 32;
    __shared__ uint32_t wh[NW][256];
    __shared__ uint32_t wbase[NW][256];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < NW * 256; i += SORT_THREADS) (&wh[0][0])[i] = 0;
    __syncthreads();

    uint32_t myk[SORT_ITEMS];
    int32_t myv[SORT_ITEMS];
    uint32_t rank[SORT_ITEMS];
    const long long base = (long long)blockIdx.x * SORT_TILE + (long long)warp * (32 * SORT_ITEMS);
    const uint32_t lt = (1u << lane) - 1u;
#pragma unroll
    for (int r = 0; r < SORT_ITEMS; r++) {
        const long long idx = base + r * 32 + lane;
        const bool valid = idx < M;
        const uint32_t key = valid ? keys_in[idx] : 0u;
        myk[r] = key;
        myv[r] = valid ? (vals_in ? vals_in[idx] : (int32_t)idx) : 0;
        const uint32_t digit = valid ? ((key >> shift) & 255u) : 256u;
        const uint32_t peers = __match_any_sync(0xffffffffu, digit);
        const uint32_t rk = __popc(peers & lt);
        uint32_t pre = 0;
        if (valid) pre = wh[warp][digit];
        __syncwarp();
        if (valid && rk == 0) wh[warp][digit] = pre + __popc(peers);
        __syncwarp();
        rank[r] = pre + rk;
    }
    __syncthreads();
    {
        const int d = threadIdx.x;                        // SORT_THREADS == 256 digits
        uint32_t run = hist[(long long)d * nCTA + blockIdx.x];
#pragma unroll
        for (int w = 0; w < NW; w++) { wbase[w][d] = run; run += wh[w][d]; }
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < SORT_ITEMS; r++) {
        const long long idx = base + r * 32 + lane;
        if (idx < M) {
            const uint32_t digit = (myk[r] >> shift) & 255u;
            const uint32_t pos = wbase[warp][digit] + rank[r];
            keys_out[pos] = myk[r];
            vals_out[pos] = myv[r];
        }
    }
}

__global__ void k_iota(int32_t* v, long long M)
{
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i < M) v[i] = (int32_t)i;
}

// tile_start[t] = first sorted position whose key >= t  (keys sorted ascending)
__global__ void k_tile_start(const uint32_t* __restrict__ keys, long long M, long long ntiles,
                             int32_t* __restrict__ tile_start)
{
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i > M) return;
    const long long prev = (i == 0) ? -1 : (long long)keys[i - 1];
    const long long cur = (i == M) ? ntiles : (long long)keys[i];
    for (long long t = prev + 1; t <= cur; t++) tile_start[t] = (int32_t)i;
}

// shifted nodes in sorted order (physical reorder: removes the "expensive because of cache
// misses" gather of /root/reference/src/precomputation.jl:537 from every exec)
template <typename T>
__global__ void k_gather_nodes(const T* __restrict__ k, const int32_t* __restrict__ perm,
                               long long M, int D, T* __restrict__ xs)
{
    long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (q >= M * D) return;
    const long long i = q / D;
    const int d = (int)(q - i * D);
    xs[q] = shift_node<T>(k[(long long)perm[i] * D + d]);
}

// 1-D only: the output-stationary spreader works on sub-blocks of NFFTB_G1D cells inside a tile; the number of
// nodes one sub-block must bucket = its own nodes + the nodes within m cells on either side.  The maximum over
// all sub-blocks sizes the kernel's shared-memory buffers.
template <typename T>
__global__ void k_count_neigh_1d(const T* __restrict__ xs, long long M, GeomDev g, int m, int S, int* __restrict__ cnt)
{
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= M) return;
    T ks;
    const int c = node_cell<T>(xs[i], g.Nt[0], ks);
    const int bs = g.bs[0], nb = g.nb[0], Nt = g.Nt[0];
    auto block_of = [&](int cell) {                  // sub-block index of a (wrapped) cell
        cell = cell < 0 ? cell + Nt : (cell >= Nt ? cell - Nt : cell);
        const int t = cell / bs;
        return t * S + (cell - t * bs) / NFFTB_G1D;
    };
    // every sub-block whose cells the node's 2m taps reach (cells c - m + 1 .. c + m, periodic) buckets the node; a
    // short trailing sub-block can make that three blocks, so walk the tap span instead of testing its two ends
    int last = -1, first = -1;
    for (int d = -(m - 1); d <= m; d++) {
        const int bk = block_of(c + d);
        if (bk != last && bk != first) { atomicAdd(&cnt[bk], 1); if (first < 0) first = bk; last = bk; }
    }
    (void)nb;
}
__global__ void k_max_int(const int* __restrict__ a, long long n, int* __restrict__ out)
{
    int mx = 0;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) mx = max(mx, a[i]);
    atomicMax(out, mx);
}

template <typename T> int sort_impl(nfftb200_plan* p, const void* d_k)
{
    const long long M = p->M;
    GeomDev g = make_geom<T>(p);
    cudaStream_t s = p->stream;
    CUDA_TRY(p, cudaMemsetAsync(p->d_flag, 0, sizeof(int), s));
    if (M > 0) {
        int blocks = (int)std::min<long long>((M + 255) / 256, 148 * 16);
        k_node_keys<T><<<blocks, 256, 0, s>>>((const T*)d_k, M, g, p->d_keys[0], p->d_flag);
        p->launches++;
    }
    int flag = 0;
    CUDA_TRY(p, cudaMemcpyAsync(&flag, p->d_flag, sizeof(int), cudaMemcpyDeviceToHost, s));
    CUDA_TRY(p, cudaStreamSynchronize(s));
    if (flag) return nfftb_fail(p, NFFTB200_BAD_NODE_RANGE,
                                "Nodes k need to be within the range [-1/2, 1/2]");

    int bits = 0;
    while ((1ll << bits) < p->ntiles) bits++;
    const int passes = (bits + 7) / 8;
    const int nCTA = (int)((M + SORT_TILE - 1) / SORT_TILE);
    int cur = 0;
    if (passes == 0 || M == 0) {
        if (M > 0) { k_iota<<<(unsigned)((M + 255) / 256), 256, 0, s>>>(p->d_vals[0], M); p->launches++; }
    } else {
        for (int ps = 0; ps < passes; ps++) {
            const int shift = 8 * ps;
            k_radix_hist<<<nCTA, SORT_THREADS, 0, s>>>(p->d_keys[cur], M, shift, p->d_hist, nCTA);
            {
                const long long hn = 256ll * nCTA;
                const int nchunks = (int)((hn + SCAN_CHUNK - 1) / SCAN_CHUNK);
                uint32_t* sums = p->d_hist + hn;                     // scratch right after the histogram
                k_scan_chunks<<<nchunks, 1024, 0, s>>>(p->d_hist, hn, sums);
                k_scan_sums<<<1, 1024, 0, s>>>(sums, nchunks);
                k_scan_add<<<nchunks, 1024, 0, s>>>(p->d_hist, hn, sums);
                p->launches += 2;
            }
            k_radix_scatter<<<nCTA, SORT_THREADS, 0, s>>>(
                p->d_keys[cur], ps == 0 ? nullptr : p->d_vals[cur], p->d_keys[cur ^ 1],
                p->d_vals[cur ^ 1], M, shift, p->d_hist, nCTA);
            p->launches += 3;
            cur ^= 1;
        }
    }
    p->d_perm = p->d_vals[cur];
    k_tile_start<<<(unsigned)((M + 1 + 255) / 256), 256, 0, s>>>(p->d_keys[cur], M, p->ntiles,
                                                                 p->d_tile_start);
    p->launches++;
    if (M > 0) {
        const long long n = M * p->D;
        k_gather_nodes<T><<<(unsigned)((n + 255) / 256), 256, 0, s>>>((const T*)d_k, p->d_perm, M,
                                                                    p->D, (T*)p->d_xs);
        p->launches++;
    }
    p->max_neigh_1d = 0;
    if (p->D == 1 && M > 0) {
        int* d_cnt = nullptr;
        const int S = (int)((p->bs[0] + NFFTB_G1D - 1) / NFFTB_G1D);
        const long long nblk = p->ntiles * S;
        CUDA_TRY(p, cudaMalloc(&d_cnt, sizeof(int) * (size_t)(nblk + 1)));
        CUDA_TRY(p, cudaMemsetAsync(d_cnt, 0, sizeof(int) * (size_t)(nblk + 1), s));
        k_count_neigh_1d<T><<<(unsigned)((M + 255) / 256), 256, 0, s>>>((const T*)p->d_xs, M, g, p->m, S, d_cnt);
        k_max_int<<<64, 256, 0, s>>>(d_cnt, nblk, d_cnt + nblk);
        int mx = 0;
        CUDA_TRY(p, cudaMemcpyAsync(&mx, d_cnt + nblk, sizeof(int), cudaMemcpyDeviceToHost, s));
        CUDA_TRY(p, cudaStreamSynchronize(s));
        cudaFree(d_cnt);
        p->max_neigh_1d = mx;
        p->launches += 2;
    }
    CUDA_TRY(p, cudaGetLastError());
    return NFFTB200_OK;
}


// ---------------------------------------------------------------------------------------------------------------
// (tile, bin) order for the register-window kernels of kernel_mode 8.  One CTA per tile re-orders the tile's nodes
// (already contiguous and in ascending caller index after the radix sort) by the bin key
//     q = octant * S^3 + colour,  octant = sum_d (b_d / S) 2^d,  colour = sum_d (b_d % S) S^d,
// b_d = bin of the node's first tap along d (bins of G, ..., G, W-(S-1)G first-tap positions per period of W, see
// bin_of in bin_common.cuh).  Inside a bin the nodes are ordered by the 4-cell half of the period their x cell lies in, then
// by ascending caller index (stable).  Free of data-dependent atomics: every
// warp ranks a contiguous part of the tile with __match_any_sync, two passes (count, then place).
// ---------------------------------------------------------------------------------------------------------------
constexpr int BINQ_MAX = 216;      // 8 * 3^3

__device__ __forceinline__ int binq_1d(int lc, int W, int G, int S)
{
    const int per = lc / W, r = (lc - per * W) / G;
    return S * per + (r < S - 1 ? r : S - 1);
}

template <typename T>
__global__ void __launch_bounds__(256)
k_bin_order(const T* __restrict__ xs, const int32_t* __restrict__ perm, const int32_t* __restrict__ tile_start,
            GeomDev g, int W, int G, int S, int NQ, T* __restrict__ xs2, int32_t* __restrict__ perm2,
            int32_t* __restrict__ bin_start, long long ntiles)
{
    // counters per (bin, 4-cell half of the bin period along x): the order inside a bin is by that half, which is what
    // lets the 10-cell interpolation windows (interp_lean.cuh: LEAN_WIDE_WX) see their nodes contiguously as well
    constexpr int NK_MAX = 2 * BINQ_MAX;
    __shared__ int wh[8][NK_MAX];
    __shared__ int tot[NK_MAX + 1];
    const int NK = 2 * NQ;
    const int t = blockIdx.x;
    const int lo = tile_start[t], hi = tile_start[t + 1];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int tx = t % g.nb[0], ty = (t / g.nb[0]) % g.nb[1], tz = t / (g.nb[0] * g.nb[1]);
    const int c0[3] = {tx * g.bs[0], ty * g.bs[1], tz * g.bs[2]};
    for (int i = threadIdx.x; i < 8 * NK_MAX; i += 256) (&wh[0][0])[i] = 0;
    __syncthreads();
    const int n = hi - lo;
    const int per_warp = ((n + 8 * 32 - 1) / (8 * 32)) * 32;          // whole groups of 32 per warp
    const int w_lo = lo + warp * per_warp, w_hi = min(hi, w_lo + per_warp);
    const unsigned lt = (1u << lane) - 1u;
    auto key_of = [&](int i) {
        int o = 0, c = 0, sm = 1, half = 0;
#pragma unroll
        for (int d = 0; d < 3; d++) {
            T ks;
            const int cell = node_cell<T>(xs[(long long)i * 3 + d], g.Nt[d], ks);
            if (d == 0) half = ((cell - c0[0]) >> 2) & 1;
            const int b = binq_1d(cell - c0[d], W, G, S);
            const int od = b / S;
            o += od << d;
            c += (b - od * S) * sm;
            sm *= S;
        }
        return 2 * (o * sm + c) + half;
    };
    for (int b0 = w_lo; b0 < w_hi; b0 += 32) {                         // pass 1: counts per (warp, bin)
        const int i = b0 + lane;
        const bool on = i < w_hi;
        const int q = on ? key_of(i) : -1;
        const unsigned peers = __match_any_sync(0xffffffffu, q);
        if (on && (peers & lt) == 0) wh[warp][q] += __popc(peers);
        __syncwarp();
    }
    __syncthreads();
    for (int k = threadIdx.x; k < NK; k += 256) {                       // exclusive offsets of the warps inside a key
        int run = 0;
        for (int w = 0; w < 8; w++) { const int c = wh[w][k]; wh[w][k] = run; run += c; }
        tot[k] = run;
    }
    __syncthreads();
    if (warp == 0) {                                                   // exclusive scan of the bin totals (<= 216)
        constexpr int KPL = (NK_MAX + 31) / 32;
        int c[KPL], sum = 0;
#pragma unroll
        for (int k = 0; k < KPL; k++) { const int idx = lane * KPL + k; c[k] = idx < NK ? tot[idx] : 0; sum += c[k]; }
        int incl = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
        int run = incl - sum;
#pragma unroll
        for (int k = 0; k < KPL; k++) { const int idx = lane * KPL + k; if (idx < NK) tot[idx] = run; run += c[k]; }
    }
    __syncthreads();
    for (int q = threadIdx.x; q < NQ; q += 256) bin_start[(long long)t * NQ + q] = lo + tot[2 * q];
    if (t == ntiles - 1 && threadIdx.x == 0) bin_start[ntiles * NQ] = hi;
    for (int b0 = w_lo; b0 < w_hi; b0 += 32) {                         // pass 2: place
        const int i = b0 + lane;
        const bool on = i < w_hi;
        const int q = on ? key_of(i) : -1;
        const unsigned peers = __match_any_sync(0xffffffffu, q);
        const int r = __popc(peers & lt);
        int base = 0;
        if (on) base = wh[warp][q];
        __syncwarp();
        if (on) {
            const long long pos = (long long)lo + tot[q] + base + r;
            xs2[pos * 3 + 0] = xs[(long long)i * 3 + 0];
            xs2[pos * 3 + 1] = xs[(long long)i * 3 + 1];
            xs2[pos * 3 + 2] = xs[(long long)i * 3 + 2];
            perm2[pos] = perm[i];
            if (r == 0) wh[warp][q] = base + __popc(peers);
        }
        __syncwarp();
    }
}

template <typename T> int bins_impl(nfftb200_plan* p, int W, int G)
{
    const int S = (W + G - 1) / G, NQ = 8 * S * S * S;
    if (p->D != 3 || NQ > BINQ_MAX) return nfftb_fail(p, NFFTB200_UNSUPPORTED, "bin order: unsupported geometry");
    for (int d = 0; d < 3; d++)
        if (p->bs[d] > 2 * W) return nfftb_fail(p, NFFTB200_UNSUPPORTED, "bin order: tile wider than two bin periods");
    if (p->have_bins && p->bins_nq == NQ) return NFFTB200_OK;
    const int64_t M = std::max<int64_t>(p->M, 1);
    if (M > p->cap_bins_nodes) {
        if (p->d_xs2) cudaFree(p->d_xs2);
        if (p->d_perm2) cudaFree(p->d_perm2);
        p->d_xs2 = nullptr; p->d_perm2 = nullptr; p->cap_bins_nodes = 0;
        CUDA_TRY(p, cudaMalloc(&p->d_xs2, (size_t)M * 3 * sizeof(T)));
        CUDA_TRY(p, cudaMalloc((void**)&p->d_perm2, (size_t)M * 4));
        p->cap_bins_nodes = M;
    }
    const int64_t tab = p->ntiles * NQ + 1;
    if (tab > p->cap_bin_tab) {
        if (p->d_bin_start) cudaFree(p->d_bin_start);
        p->d_bin_start = nullptr; p->cap_bin_tab = 0;
        CUDA_TRY(p, cudaMalloc((void**)&p->d_bin_start, (size_t)tab * 4));
        p->cap_bin_tab = tab;
    }
    k_bin_order<T><<<(unsigned)p->ntiles, 256, 0, p->stream>>>((const T*)p->d_xs, p->d_perm, p->d_tile_start, make_geom<T>(p), W, G, S,
                                                              NQ, (T*)p->d_xs2, p->d_perm2, p->d_bin_start, p->ntiles);
    p->launches++;
    CUDA_TRY(p, cudaGetLastError());
    {   // expected arrivals per output block: the work items of its distinct neighbour tiles (periodic)
        const int64_t nb0 = p->nb[0], nb1 = p->nb[1], nb2 = p->nb[2];
        p->h_expect.assign((size_t)p->ntiles, 0);
        auto offs = [](int64_t nb, int (&o)[3]) { int n = 0; if (nb >= 3) { o[n++] = -1; o[n++] = 0; o[n++] = 1; } else if (nb == 2) { o[n++] = -1; o[n++] = 0; } else o[n++] = 0; return n; };
        int ox[3], oy[3], oz[3];
        const int nx = offs(nb0, ox), ny = offs(nb1, oy), nz = offs(nb2, oz);
        for (int64_t tz = 0; tz < nb2; tz++)
            for (int64_t ty = 0; ty < nb1; ty++)
                for (int64_t tx = 0; tx < nb0; tx++) {
                    int32_t e = 0;
                    for (int c = 0; c < nz; c++)
                        for (int b = 0; b < ny; b++)
                            for (int a = 0; a < nx; a++) {
                                const int64_t t = (((tz + oz[c] + nb2) % nb2) * nb1 + (ty + oy[b] + nb1) % nb1) * nb0 + (tx + ox[a] + nb0) % nb0;
                                e += p->h_tile_items[(size_t)t + 1] - p->h_tile_items[(size_t)t];
                            }
                    p->h_expect[(size_t)((tz * nb1 + ty) * nb0 + tx)] = e;
                }
        if (p->d_expect) cudaFree(p->d_expect);
        p->d_expect = nullptr;
        CUDA_TRY(p, cudaMalloc((void**)&p->d_expect, sizeof(int32_t) * (size_t)p->ntiles));
        CUDA_TRY(p, cudaMemcpyAsync(p->d_expect, p->h_expect.data(), sizeof(int32_t) * (size_t)p->ntiles, cudaMemcpyHostToDevice, p->stream));
    }
    p->have_bins = true;
    p->bins_nq = NQ;
    return NFFTB200_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// 1-D: cell order.  The tile-sorted node list is sorted once more, stably, by grid cell (a global LSD radix sort on
// the cell index: cells of a tile are contiguous, so the result is still tile-major and tile_start stays valid).
// The 1-D kernels then see the nodes of any run of cells as ONE contiguous range: the interpolator's grid reads and
// the spreader's staging become coalesced, and the spreader needs no per-CTA bucketing at all.
// ---------------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void k_cell_keys_1d(const T* __restrict__ xs, long long M, int Nt, uint32_t* __restrict__ keys)
{
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= M) return;
    T ks;
    keys[i] = (uint32_t)node_cell<T>(xs[i], Nt, ks);
}

template <typename T>
__global__ void k_gather_sorted_1d(const T* __restrict__ xs, const int32_t* __restrict__ perm, const int32_t* __restrict__ order,
                                   long long M, T* __restrict__ xs2, int32_t* __restrict__ perm2)
{
    const long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (q >= M) return;
    const int32_t i = order[q];
    xs2[q] = xs[i];
    perm2[q] = perm[i];
}

template <typename T> int cells1d_impl(nfftb200_plan* p)
{
    if (p->D != 1) return nfftb_fail(p, NFFTB200_UNSUPPORTED, "cell order: 1-D plans only");
    if (p->have_bins && p->bins_nq == -1) return NFFTB200_OK;
    const int64_t M = std::max<int64_t>(p->M, 1);
    if (M > p->cap_bins_nodes) {
        if (p->d_xs2) cudaFree(p->d_xs2);
        if (p->d_perm2) cudaFree(p->d_perm2);
        p->d_xs2 = nullptr; p->d_perm2 = nullptr; p->cap_bins_nodes = 0;
        CUDA_TRY(p, cudaMalloc(&p->d_xs2, (size_t)M * 3 * sizeof(T)));
        CUDA_TRY(p, cudaMalloc((void**)&p->d_perm2, (size_t)M * 4));
        p->cap_bins_nodes = M;
    }
    const int64_t tab = p->Nt[0] + 1;
    if (tab > p->cap_bin_tab) {
        if (p->d_bin_start) cudaFree(p->d_bin_start);
        p->d_bin_start = nullptr; p->cap_bin_tab = 0;
        CUDA_TRY(p, cudaMalloc((void**)&p->d_bin_start, (size_t)tab * 4));
        p->cap_bin_tab = tab;
    }
    cudaStream_t s = p->stream;
    if (p->M > 0) {
        // scratch for the second sort: key / value ping-pong buffers, freed again (plan-time only)
        uint32_t* k2[2] = {nullptr, nullptr};
        int32_t* v2[2] = {nullptr, nullptr};
        for (int i = 0; i < 2; i++) {
            if (cudaMalloc((void**)&k2[i], (size_t)M * 4) != cudaSuccess || cudaMalloc((void**)&v2[i], (size_t)M * 4) != cudaSuccess) {
                cudaGetLastError();
                for (int j = 0; j < 2; j++) { if (k2[j]) cudaFree(k2[j]); if (v2[j]) cudaFree(v2[j]); }
                return nfftb_fail(p, NFFTB200_OOM, "cell order: out of device memory");
            }
        }
        k_cell_keys_1d<T><<<(unsigned)((p->M + 255) / 256), 256, 0, s>>>((const T*)p->d_xs, p->M, (int)p->Nt[0], k2[0]);
        int bits = 0;
        while ((1ll << bits) < p->Nt[0]) bits++;
        const int passes = (bits + 7) / 8;
        const int nCTA = (int)((p->M + SORT_TILE - 1) / SORT_TILE);
        int cur = 0;
        for (int ps = 0; ps < passes; ps++) {
            const int shift = 8 * ps;
            k_radix_hist<<<nCTA, SORT_THREADS, 0, s>>>(k2[cur], p->M, shift, p->d_hist, nCTA);
            const long long hn = 256ll * nCTA;
            const int nchunks = (int)((hn + SCAN_CHUNK - 1) / SCAN_CHUNK);
            uint32_t* sums = p->d_hist + hn;
            k_scan_chunks<<<nchunks, 1024, 0, s>>>(p->d_hist, hn, sums);
            k_scan_sums<<<1, 1024, 0, s>>>(sums, nchunks);
            k_scan_add<<<nchunks, 1024, 0, s>>>(p->d_hist, hn, sums);
            k_radix_scatter<<<nCTA, SORT_THREADS, 0, s>>>(k2[cur], ps == 0 ? nullptr : v2[cur], k2[cur ^ 1], v2[cur ^ 1], p->M, shift, p->d_hist, nCTA);
            p->launches += 5;
            cur ^= 1;
        }
        if (passes == 0) { k_iota<<<(unsigned)((p->M + 255) / 256), 256, 0, s>>>(v2[cur], p->M); }
        k_gather_sorted_1d<T><<<(unsigned)((p->M + 255) / 256), 256, 0, s>>>((const T*)p->d_xs, p->d_perm, v2[cur], p->M, (T*)p->d_xs2, p->d_perm2);
        k_tile_start<<<(unsigned)((p->M + 1 + 255) / 256), 256, 0, s>>>(k2[cur], p->M, p->Nt[0], p->d_bin_start);
        p->launches += 3;
        CUDA_TRY(p, cudaStreamSynchronize(s));
        for (int i = 0; i < 2; i++) { cudaFree(k2[i]); cudaFree(v2[i]); }
    } else {
        CUDA_TRY(p, cudaMemsetAsync(p->d_bin_start, 0, (size_t)tab * 4, s));
    }
    CUDA_TRY(p, cudaGetLastError());
    p->have_bins = true;
    p->bins_nq = -1;                  // marks the 1-D cell order
    return NFFTB200_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// 2-D: bin order for the batch-stationary spreader (twod_batch.cuh: k_spread_win2d).  Inside every tile the nodes are
// grouped by bin = (8 cells along x) x (one row of cells) (key = lc1 * nq0 + lc0 / 8, ascending caller index inside a
// bin): all nodes of a bin have their 2m <= 8 taps per dimension inside one register window of 16 columns x 8 rows.  Same two-pass
// match_any ranking as k_bin_order; tile_start stays valid, the work items (n_lo, n_hi, stride) apply unchanged.
// ---------------------------------------------------------------------------------------------------------------
constexpr int BINQ2D_MAX = 128;    // tiles of at most 32 x 32 cells: 4 column groups x 32 rows

template <typename T>
__global__ void __launch_bounds__(256)
k_bin_order2d(const T* __restrict__ xs, const int32_t* __restrict__ perm, const int32_t* __restrict__ tile_start, GeomDev g,
              int nq0, int NQ, T* __restrict__ xs2, int32_t* __restrict__ perm2)
{
    __shared__ int wh[8][BINQ2D_MAX];
    __shared__ int tot[BINQ2D_MAX + 1];
    const int t = blockIdx.x;
    const int lo = tile_start[t], hi = tile_start[t + 1];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int tx = t % g.nb[0], ty = t / g.nb[0];
    const int c0 = tx * g.bs[0], c1 = ty * g.bs[1];
    for (int q = threadIdx.x; q < 8 * BINQ2D_MAX; q += 256) (&wh[0][0])[q] = 0;
    __syncthreads();
    const int n = hi - lo;
    const int per_warp = ((n + 8 * 32 - 1) / (8 * 32)) * 32;          // whole groups of 32 per warp
    const int w_lo = lo + warp * per_warp, w_hi = min(hi, w_lo + per_warp);
    const unsigned lt = (1u << lane) - 1u;
    auto key_of = [&](int i) {
        T ks;
        const int l0 = node_cell<T>(xs[(long long)i * 2 + 0], g.Nt[0], ks) - c0;
        const int l1 = node_cell<T>(xs[(long long)i * 2 + 1], g.Nt[1], ks) - c1;
        return l1 * nq0 + (l0 >> 3);
    };
    for (int b0 = w_lo; b0 < w_hi; b0 += 32) {                         // pass 1: counts per (warp, bin)
        const int i = b0 + lane;
        const bool on = i < w_hi;
        const int q = on ? key_of(i) : -1;
        const unsigned peers = __match_any_sync(0xffffffffu, q);
        if (on && (peers & lt) == 0) wh[warp][q] += __popc(peers);
        __syncwarp();
    }
    __syncthreads();
    if (threadIdx.x < NQ) {                                            // exclusive offsets of the warps inside a bin
        int run = 0;
        for (int w = 0; w < 8; w++) { const int c = wh[w][threadIdx.x]; wh[w][threadIdx.x] = run; run += c; }
        tot[threadIdx.x] = run;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int run = 0;
        for (int q = 0; q < NQ; q++) { const int c = tot[q]; tot[q] = run; run += c; }
    }
    __syncthreads();
    for (int b0 = w_lo; b0 < w_hi; b0 += 32) {                         // pass 2: place
        const int i = b0 + lane;
        const bool on = i < w_hi;
        const int q = on ? key_of(i) : -1;
        const unsigned peers = __match_any_sync(0xffffffffu, q);
        const int r = __popc(peers & lt);
        int base = 0;
        if (on) base = wh[warp][q];
        __syncwarp();
        if (on) {
            const long long pos = (long long)lo + tot[q] + base + r;
            xs2[pos * 2 + 0] = xs[(long long)i * 2 + 0];
            xs2[pos * 2 + 1] = xs[(long long)i * 2 + 1];
            perm2[pos] = perm[i];
            if (r == 0) wh[warp][q] = base + __popc(peers);
        }
        __syncwarp();
    }
}

template <typename T> int bins2d_impl(nfftb200_plan* p)
{
    if (p->D != 2 || p->bs[0] > 32 || p->bs[1] > 32) return nfftb_fail(p, NFFTB200_UNSUPPORTED, "2-D bin order: unsupported geometry");
    const int nq0 = (int)(p->bs[0] + 7) / 8, NQ = nq0 * (int)p->bs[1];
    if (p->have_bins && p->bins_nq == -100 - NQ) return NFFTB200_OK;
    const int64_t M = std::max<int64_t>(p->M, 1);
    if (M > p->cap_bins_nodes) {
        if (p->d_xs2) cudaFree(p->d_xs2);
        if (p->d_perm2) cudaFree(p->d_perm2);
        p->d_xs2 = nullptr; p->d_perm2 = nullptr; p->cap_bins_nodes = 0;
        CUDA_TRY(p, cudaMalloc(&p->d_xs2, (size_t)M * 3 * sizeof(T)));
        CUDA_TRY(p, cudaMalloc((void**)&p->d_perm2, (size_t)M * 4));
        p->cap_bins_nodes = M;
    }
    k_bin_order2d<T><<<(unsigned)p->ntiles, 256, 0, p->stream>>>((const T*)p->d_xs, p->d_perm, p->d_tile_start, make_geom<T>(p), nq0, NQ,
                                                                (T*)p->d_xs2, p->d_perm2);
    p->launches++;
    CUDA_TRY(p, cudaGetLastError());
    p->have_bins = true;
    p->bins_nq = -100 - NQ;           // marks the 2-D bin order
    return NFFTB200_OK;
}

}  // namespace

int nfftb_ensure_bins_2d(nfftb200_plan* p)
{
    return p->dtype == NFFTB200_F32 ? bins2d_impl<float>(p) : bins2d_impl<double>(p);
}

int nfftb_ensure_cells_1d(nfftb200_plan* p)
{
    return p->dtype == NFFTB200_F32 ? cells1d_impl<float>(p) : cells1d_impl<double>(p);
}

int nfftb_ensure_bins(nfftb200_plan* p, int W, int G)
{
    return p->dtype == NFFTB200_F32 ? bins_impl<float>(p, W, G) : bins_impl<double>(p, W, G);
}

int nfftb_sort_nodes(nfftb200_plan* p, const void* d_k)
{
    return p->dtype == NFFTB200_F32 ? sort_impl<float>(p, d_k) : sort_impl<double>(p, d_k);
}
