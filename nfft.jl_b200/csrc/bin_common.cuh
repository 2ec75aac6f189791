// bin_common.cuh -- pieces shared by the opt-in register-footprint kernels of kernel_mode 7 (spread_bin.cuh,
// interp_bin.cuh): the bin geometry of a tile, the bank-conflict-free pitches of the padded tile in shared memory and
// the stable counting sort of a staged chunk of nodes by bin.  Device code only; tests/emu also compiles it for the
// host (NFFTB_EMU) to check the indexing without a GPU.
#pragma once
#include "common.cuh"
#include "window.cuh"
#include "tile3d.cuh"

// alignment checks of the vector accesses: active only in the host emulation build (tests/emu defines it)
#ifndef NFFTB_EMU_ALIGNED
#define NFFTB_EMU_ALIGNED(ptr, bytes)
#endif

#define NFFTB_BIN_MAXKEYS 512
#define NFFTB_BIN_WARPS 8
#define NFFTB_BIN_ROUND 8          // nodes whose weights are evaluated together (lane = node * 3 + dim)

template <typename T> struct BinChunk { static constexpr int value = sizeof(T) == 4 ? 640 : 512; };

struct BinGeom {
    int G;             // first-tap positions per bin and dimension (the last bin of a period of W holds fewer)
    int S;             // colour stride in bins: windows of bins i and i + S are disjoint
    int nbin[3];       // bins per dimension (bin_of(bs - 1) + 1)
    int maxit;         // spreader: bins of one colour a warp may have to take, ceil(max bins per colour / warps)
    int nkeys;         // spreader sort keys: (warp, colour, turn) = warps * S^3 * maxit  (<= NFFTB_BIN_MAXKEYS)
    int ikeys;         // interpolator sort keys: (warp, turn) = warps * ceil(bins / warps)  (<= NFFTB_BIN_MAXKEYS)
    int PXp, PL;       // row pitch / plane pitch of the padded tile in shared memory (cells), bank-conflict free
    int PNs;           // cells of the shared-memory tile (PL * PZ, even)
};

// bank-conflict degree of the read-modify-write of one pass: lane r touches cell (z*PL + y*PXp + i), i uniform
inline int bin_conflict_degree(int W, int cell_bytes, int PXp, int PL)
{
    const int group = 128 / cell_bytes;                   // lanes served by one wavefront (16 for 8 B, 8 for 16 B)
    const int wpc = cell_bytes / 4;                       // banks per cell
    int worst = 1;
    const int rows = W * W, passes = (rows + 31) / 32;
    for (int p = 0; p < passes; p++)
        for (int g0 = 0; g0 < 32; g0 += group) {
            int cnt[32] = {0};
            for (int l = g0; l < g0 + group; l++) {
                const int r = l + 32 * p;
                if (r >= rows) continue;
                const long long cell = (long long)(r / W) * PL + (long long)(r % W) * PXp;
                const int bank = (int)((cell * wpc) % 32);
                cnt[bank]++;
            }
            for (int b = 0; b < 32; b++) worst = cnt[b] > worst ? cnt[b] : worst;
        }
    return worst;
}

// Bin layout along one dimension: first-tap positions are grouped with period W into S = ceil(W/G) bins of G, ..., G and
// W - (S-1)G positions, so that bins b and b + S start exactly W positions apart (their W-wide windows are disjoint,
// which is what the colouring needs) and the bins of one colour have the same size in every period (balanced warps).
template <int W, int G> __host__ __device__ __forceinline__ int bin_of(int lc)
{
    constexpr int S = (W + G - 1) / G;
    const int per = lc / W, r = (lc - per * W) / G;
    return S * per + (r < S - 1 ? r : S - 1);
}
template <int W, int G> __host__ __device__ __forceinline__ int bin_first(int b)
{
    constexpr int S = (W + G - 1) / G;
    const int per = b / S;
    return W * per + G * (b - per * S);
}

// bins / pitches for a W-wide window and a 2m-tap footprint; false if the tile has too many bins
template <typename T, int MT, int W> inline bool bin_make_geom(const int* bs, BinGeom& bg)
{
    using C = typename Cplx<T>::type;
    constexpr int L = 2 * MT, G = W - L + 1;
    static_assert(G >= 1, "window narrower than the footprint");
    // the pitch search costs ~50 us of host time: remember the last tile size per instantiation and thread
    thread_local int last_bs[3] = {-1, -1, -1};
    thread_local BinGeom last_bg;
    thread_local bool last_ok = false;
    if (last_bs[0] == bs[0] && last_bs[1] == bs[1] && last_bs[2] == bs[2]) { bg = last_bg; return last_ok; }
    struct Remember {
        const int* bs; BinGeom& bg; bool ok = false;
        ~Remember() { last_bs[0] = bs[0]; last_bs[1] = bs[1]; last_bs[2] = bs[2]; last_bg = bg; last_ok = ok; }
    } rem{bs, bg};
    bg.G = G;
    bg.S = (W + G - 1) / G;
    int nbins = 1, percol = 1;
    for (int d = 0; d < 3; d++) {
        bg.nbin[d] = bin_of<W, G>(bs[d] - 1) + 1;
        nbins *= bg.nbin[d];
        percol *= (bg.nbin[d] + bg.S - 1) / bg.S;               // bins of the fullest colour along d
    }
    bg.maxit = (percol + NFFTB_BIN_WARPS - 1) / NFFTB_BIN_WARPS;
    bg.nkeys = NFFTB_BIN_WARPS * bg.S * bg.S * bg.S * bg.maxit;
    bg.ikeys = NFFTB_BIN_WARPS * ((nbins + NFFTB_BIN_WARPS - 1) / NFFTB_BIN_WARPS);
    if (bg.nkeys > NFFTB_BIN_MAXKEYS || bg.ikeys > NFFTB_BIN_MAXKEYS) return false;
    const int PX = bs[0] + L, PY = bs[1] + L, PZ = bs[2] + L;
    int best = 1 << 30, bdeg = 1 << 30;
    bg.PXp = PX; bg.PL = PX * PY;
    for (int px = PX; px < PX + 16; px++)
        for (int pl = px * PY; pl < px * PY + 32; pl++) {
            const int deg = bin_conflict_degree(W, (int)sizeof(C), px, pl);
            const int size = pl * PZ;
            if (deg < bdeg || (deg == bdeg && size < best)) { bdeg = deg; best = size; bg.PXp = px; bg.PL = pl; }
        }
    bg.PNs = (bg.PL * PZ + 1) & ~1;
    rem.ok = true;
    return true;
}

// Stable counting sort of a staged chunk by key (a key names a bin and, through its leading digits, the warp that
// will process it, so that a warp's nodes end up contiguous).  On entry key[0, nc) holds the key of every staged node and
// cntw[NWARP * nkeys] is zero (both visible to the CTA); on exit order[bin_start[k] .. bin_start[k+1]) lists the
// chunk-local ids of bin k in ascending order and cntw is dead.  No atomics: every warp ranks its own contiguous
// range with __match_any_sync, so the order -- and with it every floating-point sum -- is reproducible.
template <int CH, int NWARP>
__device__ __forceinline__ void bin_sort_chunk(int nc, int nkeys, const unsigned short* key, unsigned char* rnk,
                                               unsigned short* cntw, unsigned short* bin_start, unsigned short* order)
{
    constexpr int NTHR = NWARP * 32, CW = CH / NWARP;       // CW: nodes ranked by one warp
    static_assert(CH % NWARP == 0 && CW <= 255, "rank must fit a byte");
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned lt = (1u << lane) - 1u;
    // (1) every warp ranks its own contiguous range of the chunk
    {
        const int q_end = min(nc, (warp + 1) * CW);
        unsigned short* mycnt = cntw + warp * nkeys;
        for (int b0 = warp * CW; b0 < q_end; b0 += 32) {
            const int q = b0 + lane;
            const bool on = q < q_end;
            const unsigned kk = on ? (unsigned)key[q] : 0xffffu;
            const unsigned peers = __match_any_sync(0xffffffffu, kk);
            const int r = __popc(peers & lt);
            int base = 0;
            if (on) base = mycnt[kk];
            __syncwarp();
            if (on) {
                rnk[q] = (unsigned char)(base + r);
                if (r == 0) mycnt[kk] = (unsigned short)(base + __popc(peers));
            }
            __syncwarp();
        }
    }
    __syncthreads();
    // (2) per bin: exclusive offsets of the warps' ranges, and the bin's node count
    for (int k = threadIdx.x; k < nkeys; k += NTHR) {
        int run = 0;
        for (int w = 0; w < NWARP; w++) {
            const int c = cntw[w * nkeys + k];
            cntw[w * nkeys + k] = (unsigned short)run;
            run += c;
        }
        bin_start[k] = (unsigned short)run;
    }
    __syncthreads();
    // (3) exclusive scan of the bin counts by warp 0 (NFFTB_BIN_MAXKEYS / 32 bins per lane)
    if (warp == 0) {
        constexpr int KPL = NFFTB_BIN_MAXKEYS / 32;
        int c[KPL], sum = 0;
#pragma unroll
        for (int k = 0; k < KPL; k++) {
            const int idx = lane * KPL + k;
            c[k] = idx < nkeys ? (int)bin_start[idx] : 0;
            sum += c[k];
        }
        int incl = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        int run = incl - sum;
#pragma unroll
        for (int k = 0; k < KPL; k++) {
            const int idx = lane * KPL + k;
            if (idx < nkeys) bin_start[idx] = (unsigned short)run;
            run += c[k];
        }
        if (lane == 31) bin_start[nkeys] = (unsigned short)incl;          // total = nc
    }
    __syncthreads();
    // (4) scatter
    for (int q = threadIdx.x; q < nc; q += NTHR) {
        const int kk = key[q];
        order[bin_start[kk] + cntw[(q / CW) * nkeys + kk] + rnk[q]] = (unsigned short)q;
    }
    __syncthreads();
}

// zero n values (n * sizeof(T) a multiple of 16, p 16-byte aligned) with one warp
template <typename T> __device__ __forceinline__ void bin_zero_warp(T* p, int n, int lane)
{
    NFFTB_EMU_ALIGNED(p, 16);
    uint4* z = reinterpret_cast<uint4*>(p);
    const int n16 = (int)(n * sizeof(T) / 16);
    for (int i = lane; i < n16; i += 32) z[i] = make_uint4(0, 0, 0, 0);
}

// one cell global -> shared without register staging (LDGSTS); the host emulation copies
template <typename C> __device__ __forceinline__ void bin_copy_cell_async(C* dst, const C* src)
{
#ifdef NFFTB_EMU
    *dst = *src;
#else
    cp_async_cell(dst, src);
#endif
}
__device__ __forceinline__ void bin_copy_wait()
{
#ifndef NFFTB_EMU
    asm volatile("cp.async.wait_all;" ::: "memory");
#endif
}

// W consecutive weights of a record (16-byte aligned) -> registers, with the widest loads the width allows
template <typename T, int W> __device__ __forceinline__ void bin_load_row(const T* __restrict__ p, T (&w)[W])
{
    if constexpr (sizeof(T) == 4 && W % 4 == 0) {
        NFFTB_EMU_ALIGNED(p, 16);
#pragma unroll
        for (int k = 0; k < W / 4; k++) {
            const float4 v = reinterpret_cast<const float4*>(p)[k];
            w[4 * k] = v.x; w[4 * k + 1] = v.y; w[4 * k + 2] = v.z; w[4 * k + 3] = v.w;
        }
    } else if constexpr (W % 2 == 0) {
        using V2 = typename Cplx<T>::type;
        NFFTB_EMU_ALIGNED(p, sizeof(V2));
#pragma unroll
        for (int k = 0; k < W / 2; k++) {
            const V2 v = reinterpret_cast<const V2*>(p)[k];
            w[2 * k] = v.x; w[2 * k + 1] = v.y;
        }
    } else {
#pragma unroll
        for (int i = 0; i < W; i++) w[i] = p[i];
    }
}

// One x-row of a bin window in registers: W complex cells.  The Float32 specialisation keeps every cell as a
// (re, im) float2 -- the layout of the tile in shared memory -- and uses the packed FP32 instructions of sm_100
// (__ffma2_rn / __fmul2_rn / __fadd2_rn, SASS FFMA2 / FMUL2 / FADD2: two IEEE operations per issue slot, the scalar
// weight enters as a broadcast operand).  The register kernels are bound by FMA issue, so this halves their inner
// loops; results are bit-identical to the scalar form (each half of a packed operation is a plain fma / add).
template <typename T, int W> struct BinRow {
    using C = typename Cplx<T>::type;
    T r[W], i[W];
    __device__ __forceinline__ void zero()
    {
#pragma unroll
        for (int k = 0; k < W; k++) { r[k] = (T)0; i[k] = (T)0; }
    }
    __device__ __forceinline__ void set(int k, C c) { r[k] = c.x; i[k] = c.y; }
    // cell += row[k]
    __device__ __forceinline__ void add_to(C& cell, int k) const { cell.x += r[k]; cell.y += i[k]; }
    // row += wx * (wy * vz)
    __device__ __forceinline__ void axpy(const T (&wx)[W], T wy, C vz)
    {
        const T fr = wy * vz.x, fi = wy * vz.y;
#pragma unroll
        for (int k = 0; k < W; k++) { r[k] = tfma(wx[k], fr, r[k]); i[k] = tfma(wx[k], fi, i[k]); }
    }
    // s += wyz * sum_k wx[k] * row[k]
    __device__ __forceinline__ void dot_acc(const T (&wx)[W], T wyz, C& s) const
    {
        T a = (T)0, b = (T)0;
#pragma unroll
        for (int k = 0; k < W; k++) { a = tfma(wx[k], r[k], a); b = tfma(wx[k], i[k], b); }
        s.x = tfma(wyz, a, s.x); s.y = tfma(wyz, b, s.y);
    }
};

#ifdef NFFTB_EMU
inline float2 __ffma2_rn(float2 a, float2 b, float2 c) { return make_float2(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y)); }
inline float2 __fmul2_rn(float2 a, float2 b) { return make_float2(a.x * b.x, a.y * b.y); }
inline float2 __fadd2_rn(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
#endif

template <int W> struct BinRow<float, W> {
    float2 c[W];
    __device__ __forceinline__ void zero()
    {
#pragma unroll
        for (int k = 0; k < W; k++) c[k] = make_float2(0.f, 0.f);
    }
    __device__ __forceinline__ void set(int k, float2 v) { c[k] = v; }
    __device__ __forceinline__ void add_to(float2& cell, int k) const { cell = __fadd2_rn(cell, c[k]); }
    __device__ __forceinline__ void axpy(const float (&wx)[W], float wy, float2 vz)
    {
        const float2 f = __fmul2_rn(make_float2(wy, wy), vz);
#pragma unroll
        for (int k = 0; k < W; k++) c[k] = __ffma2_rn(make_float2(wx[k], wx[k]), f, c[k]);
    }
    __device__ __forceinline__ void dot_acc(const float (&wx)[W], float wyz, float2& s) const
    {
        float2 a = make_float2(0.f, 0.f);
#pragma unroll
        for (int k = 0; k < W; k++) a = __ffma2_rn(make_float2(wx[k], wx[k]), c[k], a);
        s = __ffma2_rn(make_float2(wyz, wyz), a, s);
    }
};
