// twod.cu -- tiled 2-D kernels (BASELINE configs C1: 256^2 Float64 m=4, C3: 512^2 radial Float32 m=4 batched).
// Same construction as the 3-D kernels (spread.cu / interp.cu), specialised to a (2m x 2m) footprint:
//  * spreader: one CTA per reference tile (64 x 64), 8 warps own 4 x 2 warp-private padded sub-tiles in shared
//    memory; lane (yt, q) of a warp owns cells q and q+m of footprint row yt (for m = 4 all 32 lanes are busy and
//    the padded row stride keeps every half-warp on distinct banks); lane-per-node weights; fixed-order merge of
//    the sub-tiles; plain stores of the merged padded tile to the per-tile scratch; k_gather_tiles2d sums the
//    <= 4 tiles covering a grid cell and writes it once.  No atomics, deterministic, no memset.
//  * interpolator: padded tile staged with cp.async, same lane mapping, 4 nodes reduced per shuffle butterfly.
#include <algorithm>
#include <type_traits>

#include "common.cuh"
#include "tile3d.cuh"
#include "window.cuh"

namespace {

constexpr int T2_WARPS = 8;
constexpr int T2_THREADS = T2_WARPS * 32;
constexpr int T2_CHUNK = 2048;

template <typename T, int MT> struct Lane2 {
    static constexpr int L = 2 * MT;
    static constexpr int LC = MT;                 // lane columns per row; a lane owns cells q and q + LC
    static constexpr int RPI = 32 / LC;           // rows per iteration
    static constexpr int NIT = (L + RPI - 1) / RPI;
    static constexpr int RW = ((2 * L + 2) + 3) & ~3;    // record: wx | wy | v
};

__host__ __device__ inline int pad_stride(int q) { while ((q & 7) != 4) q++; return q; }

template <typename T, int MT> struct Sub2 {
    static constexpr int L = 2 * MT;
    int SX, SY, QX, QXP, QY, QN;
    __host__ __device__ Sub2(const int* bs)
    {
        SX = (bs[0] + 3) / 4; SY = (bs[1] + 1) / 2;
        QX = SX + L; QXP = pad_stride(QX); QY = SY + L; QN = QXP * QY;
    }
    __host__ __device__ size_t bytes() const
    {
        return sizeof(typename Cplx<T>::type) * (size_t)T2_WARPS * QN + sizeof(T) * T2_WARPS * 32 * Lane2<T, MT>::RW +
               sizeof(int) * T2_WARPS * 32 + sizeof(unsigned short) * T2_WARPS * 64 + T2_CHUNK;
    }
};

#include "twod_batch.cuh"

template <typename T> __device__ __forceinline__ T wsum(T v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ------------------------------------------------------------------------------------------------------
template <typename T, int MT>
__global__ void __launch_bounds__(T2_THREADS)
k_spread_sub2d(const typename Cplx<T>::type* __restrict__ fhat, typename Cplx<T>::type* __restrict__ scratch,
               const T* __restrict__ xs, const int32_t* __restrict__ perm, const int32_t* __restrict__ tile_start,
               const int32_t* __restrict__ item_stride,
               int tile_lo, long long M, GeomDev geo, WinDev<T> win, const __grid_constant__ PolyParam<T, MT> pp)
{
    using C = typename Cplx<T>::type;
    using LN = Lane2<T, MT>;
    constexpr int L = 2 * MT, LC = LN::LC, RPI = LN::RPI, NIT = LN::NIT, RW = LN::RW;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const Sub2<T, MT> lay(geo.bs);
    const int QXP = lay.QXP, QN = lay.QN, SX = lay.SX, SY = lay.SY, QX = lay.QX, QY = lay.QY;
    C* sub = reinterpret_cast<C*>(smem_raw);
    T* rec_w = reinterpret_cast<T*>(sub + T2_WARPS * QN);
    int* rec_b = reinterpret_cast<int*>(rec_w + T2_WARPS * 32 * RW);
    unsigned short* list = reinterpret_cast<unsigned short*>(rec_b + T2_WARPS * 32);
    unsigned char* oct = reinterpret_cast<unsigned char*>(list + T2_WARPS * 64);

    const int32_t* item = tile_start + 3 * (size_t)(tile_lo + blockIdx.x);     // work item (tile, node range)
    const int tile_id = item[0];
    const int n_lo = item[1], n_hi = item[2];
    const int stride = item_stride[tile_lo + blockIdx.x];                       // node q of the item is n_lo + q * stride
    const int n_item = (n_hi - n_lo + stride - 1) / stride;
    const int tx = tile_id % geo.nb[0], ty = tile_id / geo.nb[0];
    const int cx0 = tx * geo.bs[0], cy0 = ty * geo.bs[1];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int ob0 = warp & 3, ob1 = warp >> 2;
    const int PX = geo.bs[0] + L, PY = geo.bs[1] + L;
    fhat += (long long)blockIdx.y * M;
    scratch += ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * ((size_t)PX * PY);
    C* mysub = sub + warp * QN;
    T* myrec = rec_w + warp * 32 * RW;
    int* mybase = rec_b + warp * 32;
    unsigned short* mylist = list + warp * 64;

    {
        uint4* z = reinterpret_cast<uint4*>(sub);
        const int n16 = (int)((sizeof(C) * (size_t)T2_WARPS * QN) / 16);
        for (int q = threadIdx.x; q < n16; q += T2_THREADS) z[q] = make_uint4(0, 0, 0, 0);
    }
    const int lq = lane % LC, lr = lane / LC;              // lane column / row inside an iteration
    const bool lane_on = lane < RPI * LC;
    const unsigned lt = (1u << lane) - 1u;

    auto process_round = [&](int cbase, int nn) {
        if (lane < nn) {
            const long long i = (long long)n_lo + (long long)(cbase + mylist[lane]) * stride;
            T ks0, ks1;
            const int c0 = node_cell<T>(xs[i * 2 + 0], geo.Nt[0], ks0);
            const int c1 = node_cell<T>(xs[i * 2 + 1], geo.Nt[1], ks1);
            T w0[L], w1[L];
            eval_taps<T, MT>(win, pp, ks0, c0, w0);
            eval_taps<T, MT>(win, pp, ks1, c1, w1);
            const C v = fhat[perm[i]];
            const int px = c0 - cx0 - ob0 * SX + 1, py = c1 - cy0 - ob1 * SY + 1;
            mybase[lane] = py * QXP + px;
            T* dst = myrec + lane * RW;
#pragma unroll
            for (int k = 0; k < L; k++) { dst[k] = w0[k]; dst[L + k] = w1[k]; }
            dst[2 * L] = v.x; dst[2 * L + 1] = v.y;
        }
        __syncwarp();
        for (int n = 0; n < nn; n++) {
            const T* rw = myrec + n * RW;
            C* p0 = mysub + mybase[n];
            const T vx = rw[2 * L], vy = rw[2 * L + 1];
            if (lane_on) {
                const T wa = rw[lq], wb = rw[lq + LC];
#pragma unroll
                for (int it = 0; it < NIT; it++) {
                    const int yt = lr + it * RPI;
                    if (yt < L) {
                        const T wy = rw[L + yt];
                        const T a = wy * vx, b = wy * vy;
                        C* p = p0 + yt * QXP + lq;
                        C c0v = p[0], c1v = p[LC];
                        c0v.x = tfma(wa, a, c0v.x); c0v.y = tfma(wa, b, c0v.y);
                        c1v.x = tfma(wb, a, c1v.x); c1v.y = tfma(wb, b, c1v.y);
                        p[0] = c0v; p[LC] = c1v;
                    }
                }
            }
            __syncwarp();
        }
    };

    for (int cbase = 0; cbase < n_item; cbase += T2_CHUNK) {                    // cbase: item-local index of the chunk
        const int nc = min(T2_CHUNK, n_item - cbase);
        __syncthreads();
        for (int q = threadIdx.x; q < nc; q += T2_THREADS) {
            const long long i = (long long)n_lo + (long long)(cbase + q) * stride;
            T ks;
            const int l0 = node_cell<T>(xs[i * 2 + 0], geo.Nt[0], ks) - cx0;
            const int l1 = node_cell<T>(xs[i * 2 + 1], geo.Nt[1], ks) - cy0;
            oct[q] = (unsigned char)(min(l0 / SX, 3) + 4 * (l1 >= SY));
        }
        __syncthreads();
        int cnt = 0;
        for (int base = 0; base < nc; base += 32) {
            const int idx = base + lane;
            const bool mine = idx < nc && oct[idx] == warp;
            const unsigned mask = __ballot_sync(0xffffffffu, mine);
            if (mine) mylist[cnt + __popc(mask & lt)] = (unsigned short)idx;
            cnt += __popc(mask);
            __syncwarp();
            if (cnt >= 32) {
                process_round(cbase, 32);
                const int rest = cnt - 32;
                unsigned short tmp = 0;
                if (lane < rest) tmp = mylist[32 + lane];
                __syncwarp();
                if (lane < rest) mylist[lane] = tmp;
                __syncwarp();
                cnt = rest;
            }
        }
        if (cnt > 0) process_round(cbase, cnt);
    }
    __syncthreads();
    // merge: y overlap of (ox,0) into (ox,1), then x overlaps of (ox,oy) into (ox+1,oy) on the live rows
    {
        const unsigned inv = fastdiv_inv(QX);
        for (int q = threadIdx.x; q < 4 * L * QX; q += T2_THREADS) {
            const int rr = (int)fastdiv(q, inv), x = q - rr * QX;
            const int ox_ = rr / L, r = rr - ox_ * L;
            const C a = sub[ox_ * QN + (SY + r) * QXP + x];
            C* d = sub + (ox_ + 4) * QN + r * QXP + x;
            C c = *d; c.x += a.x; c.y += a.y; *d = c;
        }
    }
    __syncthreads();
    {
        const unsigned invL = fastdiv_inv(L), invPY = fastdiv_inv(PY);
        for (int q = threadIdx.x; q < 3 * PY * L; q += T2_THREADS) {
            const int rr = (int)fastdiv(q, invL), xx = q - rr * L;
            const int ox_ = (int)fastdiv(rr, invPY), y = rr - ox_ * PY;          // fold ox_ -> ox_ + 1
            const int oy_ = y >= SY;
            const int ro = (y - oy_ * SY) * QXP;
            const C a = sub[(ox_ + 4 * oy_) * QN + ro + SX + xx];
            C* d = sub + (ox_ + 1 + 4 * oy_) * QN + ro + xx;
            C c = *d; c.x += a.x; c.y += a.y; *d = c;
        }
    }
    __syncthreads();
    // NOTE: the three x-folds touch disjoint cells only if SX >= 2m (checked on the host)
    for (int y = warp; y < PY; y += T2_WARPS) {
        const int oy_ = y >= SY;
        const C* prow = sub + 4 * oy_ * QN + (y - oy_ * SY) * QXP;
        C* srow = scratch + (size_t)y * PX;
        for (int x = lane; x < PX; x += 32) {
            const int ox_ = min(x / SX, 3);
            srow[x] = prow[ox_ * QN + x - ox_ * SX];
        }
    }
}

template <typename T, int MT>
__global__ void __launch_bounds__(256)
k_gather_tiles2d(const typename Cplx<T>::type* __restrict__ scratch, typename Cplx<T>::type* __restrict__ g,
                 const int32_t* __restrict__ tile_start, int tile_lo, int tile_hi, int item_lo, int item_hi, GeomDev geo)
{
    using C = typename Cplx<T>::type;
    constexpr int L = 2 * MT;
    const int PX = geo.bs[0] + L, PY = geo.bs[1] + L;
    const size_t PN = (size_t)PX * PY;
    const int u0 = blockIdx.x * blockDim.x + threadIdx.x;
    const int u1 = blockIdx.y, b = blockIdx.z;
    if (u0 >= geo.Nt[0]) return;
    scratch += (size_t)b * (item_hi - item_lo) * PN;
    auto cover = [&](int u, int d, int (&tt)[3], int (&pc)[3]) -> int {
        const int bs = geo.bs[d], nb = geo.nb[d], Nt = geo.Nt[d];
        const int t = u / bs, l = u - t * bs;
        const int len = (t == nb - 1) ? Nt - t * bs : bs;
        int n = 0;
        tt[n] = t; pc[n] = l + MT; n++;
        if (l < MT) {
            const int tp = t == 0 ? nb - 1 : t - 1;
            const int lenp = (tp == nb - 1) ? Nt - tp * bs : bs;
            tt[n] = tp; pc[n] = l + MT + lenp; n++;
        }
        if (l >= len - MT) { tt[n] = t == nb - 1 ? 0 : t + 1; pc[n] = l + MT - len; n++; }
        return n;
    };
    int tx[3], px[3], ty[3], py[3];
    const int nx = cover(u0, 0, tx, px), ny = cover(u1, 1, ty, py);
    T ax = 0, ay = 0;
    for (int iy = 0; iy < ny; iy++)
        for (int ix = 0; ix < nx; ix++) {
            const int tile = ty[iy] * geo.nb[0] + tx[ix];
            if (tile < tile_lo || tile >= tile_hi) continue;
            for (int it = tile_start[tile]; it < tile_start[tile + 1]; it++) {       // tile_start = d_tile_items here
                const C c = scratch[(size_t)(it - item_lo) * PN + (size_t)py[iy] * PX + px[ix]];
                ax += c.x; ay += c.y;
            }
        }
    g[(size_t)b * geo.gsz + (size_t)u1 * geo.Nt[0] + u0] = make_c<T>(ax, ay);
}

// ------------------------------------------------------------------------------------------------------
template <typename T, int MT> struct Interp2 {
    static constexpr int L = 2 * MT;
    static constexpr int RW = ((2 * L) + 3) & ~3;
    int PX, PXP, PY, PN;
    __host__ __device__ Interp2(const int* bs) { PX = bs[0] + L; PXP = pad_stride(PX); PY = bs[1] + L; PN = PXP * PY; }
    __host__ __device__ size_t bytes() const
    {
        return sizeof(typename Cplx<T>::type) * (size_t)(PN + T2_WARPS * 32) + sizeof(T) * T2_WARPS * 32 * RW + sizeof(int) * T2_WARPS * 64;
    }
};

template <typename T, int MT>
__global__ void __launch_bounds__(T2_THREADS)
k_interp_row2d(const typename Cplx<T>::type* __restrict__ g, typename Cplx<T>::type* __restrict__ fhat,
               const T* __restrict__ xs, const int32_t* __restrict__ perm, const int32_t* __restrict__ tile_start,
               const int32_t* __restrict__ item_stride,
               int tile_lo, long long M, GeomDev geo, WinDev<T> win, const __grid_constant__ PolyParam<T, MT> pp)
{
    using C = typename Cplx<T>::type;
    using LN = Lane2<T, MT>;
    constexpr int L = 2 * MT, LC = LN::LC, RPI = LN::RPI, NIT = LN::NIT, RW = Interp2<T, MT>::RW;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const Interp2<T, MT> lay(geo.bs);
    const int PX = lay.PX, PXP = lay.PXP, PY = lay.PY;
    C* tile = reinterpret_cast<C*>(smem_raw);
    C* res = tile + lay.PN;
    T* rec_w = reinterpret_cast<T*>(res + T2_WARPS * 32);
    int* rec_i = reinterpret_cast<int*>(rec_w + T2_WARPS * 32 * RW);

    const int32_t* item = tile_start + 3 * (size_t)(tile_lo + blockIdx.x);     // work item (tile, node range)
    const int tile_id = item[0];
    const int n_lo = item[1], n_hi = item[2];
    const int tx = tile_id % geo.nb[0], ty = tile_id / geo.nb[0];
    const int x0 = tx * geo.bs[0] - MT, y0 = ty * geo.bs[1] - MT;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    g += (long long)blockIdx.y * geo.gsz;
    fhat += (long long)blockIdx.y * M;
    {
        const bool fw = PX <= geo.Nt[0] && PY <= geo.Nt[1];
        for (int y = warp; y < PY; y += T2_WARPS) {
            const C* src = g + (size_t)wrapc(y0 + y, geo.Nt[1], fw) * geo.Nt[0];
            C* dst = tile + y * PXP;
            for (int x = lane; x < PX; x += 32) cp_async_cell(dst + x, src + wrapc(x0 + x, geo.Nt[0], fw));
        }
    }
    const int lq = lane % LC, lr = lane / LC;
    const bool lane_on = lane < RPI * LC;
    T* myrec = rec_w + warp * 32 * RW;
    int* myint = rec_i + warp * 64;
    C* myres = res + warp * 32;
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncthreads();

    const int stride = item_stride[tile_lo + blockIdx.x];                       // node q of the item is n_lo + q * stride
    const int n_item = (n_hi - n_lo + stride - 1) / stride;
    for (int rbase = 32 * warp; rbase < n_item; rbase += 32 * T2_WARPS) {
        const int nn = min(32, n_item - rbase);
        if (lane < nn) {
            const long long i = (long long)n_lo + (long long)(rbase + lane) * stride;
            T ks0, ks1;
            const int c0 = node_cell<T>(xs[i * 2 + 0], geo.Nt[0], ks0);
            const int c1 = node_cell<T>(xs[i * 2 + 1], geo.Nt[1], ks1);
            T w0[L], w1[L];
            eval_taps<T, MT>(win, pp, ks0, c0, w0);
            eval_taps<T, MT>(win, pp, ks1, c1, w1);
            myint[2 * lane] = (c1 - MT + 1 - y0) * PXP + (c0 - MT + 1 - x0);
            myint[2 * lane + 1] = perm[i];
            T* dst = myrec + lane * RW;
#pragma unroll
            for (int k = 0; k < L; k++) { dst[k] = w0[k]; dst[L + k] = w1[k]; }
        }
        __syncwarp();
        for (int n4 = 0; n4 < nn; n4 += 4) {
            T px[4], py[4];
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const int n = n4 + q;
                T sx = 0, sy = 0;
                if (n < nn && lane_on) {
                    const T* rw = myrec + n * RW;
                    const C* p0 = tile + myint[2 * n];
                    const T wa = rw[lq], wb = rw[lq + LC];
#pragma unroll
                    for (int it = 0; it < NIT; it++) {
                        const int yt = lr + it * RPI;
                        if (yt < L) {
                            const C* p = p0 + yt * PXP + lq;
                            const C c0v = p[0], c1v = p[LC];
                            const T wy = rw[L + yt];
                            const T a = tfma(wb, c1v.x, wa * c0v.x), b = tfma(wb, c1v.y, wa * c0v.y);
                            sx = tfma(wy, a, sx); sy = tfma(wy, b, sy);
                        }
                    }
                }
                px[q] = sx; py[q] = sy;
            }
            const bool hi16 = lane & 16, hi8 = lane & 8;
            T ax = hi16 ? px[1] : px[0], bx = hi16 ? px[0] : px[1];
            T ay = hi16 ? py[1] : py[0], by = hi16 ? py[0] : py[1];
            T cx = hi16 ? px[3] : px[2], dx = hi16 ? px[2] : px[3];
            T cy = hi16 ? py[3] : py[2], dy = hi16 ? py[2] : py[3];
            ax += __shfl_xor_sync(0xffffffffu, bx, 16); ay += __shfl_xor_sync(0xffffffffu, by, 16);
            cx += __shfl_xor_sync(0xffffffffu, dx, 16); cy += __shfl_xor_sync(0xffffffffu, dy, 16);
            T ex = hi8 ? cx : ax, fx = hi8 ? ax : cx;
            T ey = hi8 ? cy : ay, fy = hi8 ? ay : cy;
            ex += __shfl_xor_sync(0xffffffffu, fx, 8); ey += __shfl_xor_sync(0xffffffffu, fy, 8);
#pragma unroll
            for (int o = 4; o > 0; o >>= 1) {
                ex += __shfl_xor_sync(0xffffffffu, ex, o);
                ey += __shfl_xor_sync(0xffffffffu, ey, o);
            }
            if ((lane & 7) == 0) {
                const int q = (hi16 ? 1 : 0) + (hi8 ? 2 : 0);
                if (n4 + q < nn) myres[n4 + q] = make_c<T>(ex, ey);
            }
        }
        __syncwarp();
        if (lane < nn) fhat[myint[2 * lane + 1]] = myres[lane];
        __syncwarp();
    }
}

// Float32 gather, two x-adjacent cells per thread (16-byte loads and stores).  Valid when m, the tile width and the grid
// width are even: both cells then have the same covering tiles and every access is 16-byte aligned.  All item ranges of
// the (up to four) covering tiles are fetched before the first scratch load is issued.
template <int MT>
__global__ void __launch_bounds__(128)
k_gather_pairs2d(const float2* __restrict__ scratch, float2* __restrict__ g, const int32_t* __restrict__ tile_items, int tile_lo,
                 int tile_hi, int item_lo, int item_hi, GeomDev geo)
{
    constexpr int L = 2 * MT;
    const int PX = geo.bs[0] + L, PY = geo.bs[1] + L;
    const size_t PN = (size_t)PX * PY;
    const int u0 = 2 * (blockIdx.x * blockDim.x + threadIdx.x);
    const int u1 = blockIdx.y, b = blockIdx.z;
    if (u0 >= geo.Nt[0]) return;
    scratch += (size_t)b * (item_hi - item_lo) * PN;
    auto cover = [&](int u, int d, int (&tt)[2], int (&pc)[2]) -> int {
        const int bs = geo.bs[d], nb = geo.nb[d], Nt = geo.Nt[d];
        const int t = u / bs, l = u - t * bs;
        const int len = (t == nb - 1) ? Nt - t * bs : bs;
        int n = 0;
        tt[n] = t; pc[n] = l + MT; n++;
        if (l < MT) {
            const int tp = t == 0 ? nb - 1 : t - 1;
            const int lenp = (tp == nb - 1) ? Nt - tp * bs : bs;
            tt[n] = tp; pc[n] = l + MT + lenp; n++;
        } else if (l >= len - MT) {
            tt[n] = t == nb - 1 ? 0 : t + 1; pc[n] = l + MT - len; n++;
        }
        return n;
    };
    int tx[2], px[2], ty[2], py[2];
    const int nx = cover(u0, 0, tx, px), ny = cover(u1, 1, ty, py);
    int i0[4], i1[4], off[4], np = 0;
    for (int iy = 0; iy < ny; iy++)
        for (int ix = 0; ix < nx; ix++) {
            const int tile = ty[iy] * geo.nb[0] + tx[ix];
            if (tile < tile_lo || tile >= tile_hi) continue;
            i0[np] = tile_items[tile]; i1[np] = tile_items[tile + 1];
            off[np] = py[iy] * PX + px[ix];
            np++;
        }
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int q = 0; q < np; q++)
        for (int it = i0[q]; it < i1[q]; it++) {
            const float4 c = *reinterpret_cast<const float4*>(scratch + (size_t)(it - item_lo) * PN + off[q]);
            acc.x += c.x; acc.y += c.y; acc.z += c.z; acc.w += c.w;
        }
    *reinterpret_cast<float4*>(g + (size_t)b * geo.gsz + (size_t)u1 * geo.Nt[0] + u0) = acc;
}

// Float32 gather for batched plans, one CTA per output tile and group of transforms: the covering pieces of a pair of
// cells (offsets into the padded tiles of the tile itself and of its neighbours) are worked out once and reused for
// every transform of the group, whose loads are independent of each other (four transforms in flight per thread when
// every covering tile has a single work item, the common case).  Same validity conditions as k_gather_pairs2d.
template <int MT>
__global__ void __launch_bounds__(128)
k_gather_tile2d(const float2* __restrict__ scratch, float2* __restrict__ g, const int32_t* __restrict__ tile_items, int tile_lo,
                int tile_hi, int item_lo, int item_hi, int B, int bper, GeomDev geo)
{
    constexpr int L = 2 * MT;
    const int PX = geo.bs[0] + L, PY = geo.bs[1] + L;
    const size_t PN = (size_t)PX * PY, SB = (size_t)(item_hi - item_lo) * PN;
    const int tile = blockIdx.x, tx = tile % geo.nb[0], ty = tile / geo.nb[0];
    const int b_lo = blockIdx.y * bper, b_hi = min(B, b_lo + bper);
    auto tlen = [&](int t, int d) { return t == geo.nb[d] - 1 ? geo.Nt[d] - t * geo.bs[d] : geo.bs[d]; };
    const int len0 = tlen(tx, 0), len1 = tlen(ty, 1), hp = len0 >> 1;
    auto cover = [&](int t, int l, int len, int d, int (&tt)[2], int (&pc)[2]) -> int {
        const int nb = geo.nb[d];
        int n = 0;
        tt[n] = t; pc[n] = l + MT; n++;
        if (l < MT) {
            const int tp = t == 0 ? nb - 1 : t - 1;
            tt[n] = tp; pc[n] = l + MT + tlen(tp, d); n++;
        } else if (l >= len - MT) {
            tt[n] = t == nb - 1 ? 0 : t + 1; pc[n] = l + MT - len; n++;
        }
        return n;
    };
    for (int idx = threadIdx.x; idx < hp * len1; idx += blockDim.x) {
        const int l1 = idx / hp, l0 = 2 * (idx - l1 * hp);
        int txs[2], pxs[2], tys[2], pys[2];
        const int nx = cover(tx, l0, len0, 0, txs, pxs), ny = cover(ty, l1, len1, 1, tys, pys);
        int i0[4], i1[4], np = 0;
        size_t off[4];
        bool single = true;
        for (int iy = 0; iy < ny; iy++)
            for (int ix = 0; ix < nx; ix++) {
                const int t = tys[iy] * geo.nb[0] + txs[ix];
                if (t < tile_lo || t >= tile_hi) continue;
                i0[np] = tile_items[t]; i1[np] = tile_items[t + 1];
                if (i1[np] == i0[np]) continue;                                  // empty tile: no scratch block
                single = single && i1[np] == i0[np] + 1;
                off[np] = (size_t)(i0[np] - item_lo) * PN + (size_t)pys[iy] * PX + pxs[ix];
                np++;
            }
        float2* out = g + (size_t)(ty * geo.bs[1] + l1) * geo.Nt[0] + tx * geo.bs[0] + l0;
        if (single) {
            int b = b_lo;
            for (; b + 4 <= b_hi; b += 4) {
                float4 c[4][4];
#pragma unroll
                for (int u = 0; u < 4; u++)
#pragma unroll
                    for (int q = 0; q < 4; q++)
                        c[u][q] = q < np ? *reinterpret_cast<const float4*>(scratch + (size_t)(b + u) * SB + off[q]) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    float4 a = c[u][0];
#pragma unroll
                    for (int q = 1; q < 4; q++) { a.x += c[u][q].x; a.y += c[u][q].y; a.z += c[u][q].z; a.w += c[u][q].w; }
                    *reinterpret_cast<float4*>(out + (size_t)(b + u) * geo.gsz) = a;
                }
            }
            for (; b < b_hi; b++) {
                float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
                for (int q = 0; q < np; q++) {
                    const float4 c = *reinterpret_cast<const float4*>(scratch + (size_t)b * SB + off[q]);
                    a.x += c.x; a.y += c.y; a.z += c.z; a.w += c.w;
                }
                *reinterpret_cast<float4*>(out + (size_t)b * geo.gsz) = a;
            }
        } else {
            for (int b = b_lo; b < b_hi; b++) {
                float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
                for (int q = 0; q < np; q++)
                    for (int it = 0; it < i1[q] - i0[q]; it++) {
                        const float4 c = *reinterpret_cast<const float4*>(scratch + (size_t)b * SB + off[q] + (size_t)it * PN);
                        a.x += c.x; a.y += c.y; a.z += c.z; a.w += c.w;
                    }
                *reinterpret_cast<float4*>(out + (size_t)b * geo.gsz) = a;
            }
        }
    }
}

// batch-stationary kernels: transforms per warp (4, 2, 1 -> 32, 16, 8 transforms per CTA), or 0 if they do not apply
template <int MT> int batch2d_tpw(const nfftb200_plan* p, const GeomDev& geo, int B)
{
    if (B < 8 || p->kernel_mode == 9 || p->kernel_mode == 1) return 0;
    const Batch2<MT> lay(geo.bs);
    for (int tpw = 4; tpw >= 1; tpw >>= 1) {
        if (tpw > 1 && 8 * (tpw / 2) >= B) continue;                 // a smaller CTA batch already covers B
        if (lay.bytes(8 * tpw) <= 227 * 1024) return tpw;
    }
    return 0;
}

template <int MT, int TPW>
int spread_batch2d_go(nfftb200_plan* p, const void* fhat, int B, int item_lo, int item_hi, const GeomDev& geo)
{
    const Batch2<MT> lay(geo.bs);
    const size_t smem = lay.bytes(8 * TPW);
    auto kern = k_spread_batch2d<MT, TPW>;
    CUDA_TRY(p, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid(item_hi - item_lo, (B + 8 * TPW - 1) / (8 * TPW));
    kern<<<grid, TB_THREADS, smem, p->stream>>>((const float2*)fhat, (float2*)p->d_tilebuf, (const float*)p->d_xs, p->d_perm, p->d_items,
                                                p->d_item_stride, item_lo, p->M, B, geo, make_win<float>(p), make_poly_param<float, MT>(p));
    return NFFTB200_OK;
}
// register-window adjoint: tiles of 8..32 cells per dimension (so that a 16-cell window fits the padded tile)
template <int MT> bool win2d_ok(const GeomDev& geo)
{
    const Batch2<MT> lay(geo.bs);
    return geo.bs[0] + 2 * MT >= TW_WX && geo.bs[1] + 2 * MT >= TW_WY && geo.bs[0] <= 32 && geo.bs[1] <= 32 &&
           Win2<MT>::bytes(lay) <= 227 * 1024;
}
template <int MT>
int spread_win2d_go(nfftb200_plan* p, const void* fhat, int B, int item_lo, int item_hi, const GeomDev& geo)
{
    const Batch2<MT> lay(geo.bs);
    const size_t smem = Win2<MT>::bytes(lay);
    if (nfftb_ensure_bins_2d(p) != NFFTB200_OK) return -1;
    auto kern = k_spread_win2d<MT>;
    CUDA_TRY(p, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    dim3 grid(item_hi - item_lo, (B + TW_BC - 1) / TW_BC);
    kern<<<grid, TW_THREADS, smem, p->stream>>>((const float2*)fhat, (float2*)p->d_tilebuf, (const float*)p->d_xs2, p->d_perm2, p->d_items,
                                                p->d_item_stride, item_lo, p->M, B, geo, make_win<float>(p), make_poly_param<float, MT>(p));
    return NFFTB200_OK;
}

template <int MT, int TPW, int PXP>
int interp_batch2d_pitch(nfftb200_plan* p, const void* g, void* fhat, int B, int item_lo, int item_hi, const GeomDev& geo)
{
    const Batch2<MT> lay(geo.bs);
    const size_t smem = lay.bytes(8 * TPW);
    auto kern = k_interp_batch2d<MT, TPW, PXP>;
    CUDA_TRY(p, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid(item_hi - item_lo, (B + 8 * TPW - 1) / (8 * TPW));
    kern<<<grid, TF_THREADS, smem, p->stream>>>((const float2*)g, (float2*)fhat, (const float*)p->d_xs, p->d_perm, p->d_items,
                                                p->d_item_stride, item_lo, p->M, B, geo, make_win<float>(p), make_poly_param<float, MT>(p));
    return NFFTB200_OK;
}
template <int MT, int TPW>
int interp_batch2d_go(nfftb200_plan* p, const void* g, void* fhat, int B, int item_lo, int item_hi, const GeomDev& geo)
{
    const Batch2<MT> lay(geo.bs);
    if (lay.PXp == 24) return interp_batch2d_pitch<MT, TPW, 24>(p, g, fhat, B, item_lo, item_hi, geo);
    if (lay.PXp == 40) return interp_batch2d_pitch<MT, TPW, 40>(p, g, fhat, B, item_lo, item_hi, geo);
    return interp_batch2d_pitch<MT, TPW, 0>(p, g, fhat, B, item_lo, item_hi, geo);
}

template <typename T, int MT>
int spread2d_launch(nfftb200_plan* p, const void* fhat, void* g, int B, int t_lo, int t_hi)
{
    using C = typename Cplx<T>::type;
    GeomDev geo = make_geom<T>(p);
    int tpw = 0;
    Sub2<T, MT> lay(geo.bs);
    const size_t smem = lay.bytes();
    const bool sub_ok = smem <= 227 * 1024 && lay.SX >= 2 * MT && lay.SY >= 2 * MT;
    if constexpr (std::is_same<T, float>::value && MT <= 4) {
        tpw = batch2d_tpw<MT>(p, geo, B);
        // without the register windows (tile too large / too small for them) the per-transform kernel is the faster
        // adjoint: the read-modify-write form is kept for tiles it cannot take, and as kernel_mode 7
        if (tpw && !win2d_ok<MT>(geo) && sub_ok && p->kernel_mode != 7) tpw = 0;
    }
    if (tpw == 0 && !sub_ok) return -1;
    for (int d = 0; d < 2; d++) {
        const int last = geo.Nt[d] - (geo.nb[d] - 1) * geo.bs[d];
        if (geo.bs[d] < MT || last < MT) return -1;
    }
    const size_t PN = (size_t)(geo.bs[0] + 2 * MT) * (geo.bs[1] + 2 * MT);
    const int item_lo = p->h_tile_items[(size_t)t_lo], item_hi = p->h_tile_items[(size_t)t_hi];
    const int64_t need = (int64_t)(sizeof(C) * PN * (size_t)std::max(1, item_hi - item_lo) * B);
    if (need > p->cap_tilebuf) {
        if (p->d_tilebuf) cudaFree(p->d_tilebuf);
        p->d_tilebuf = nullptr; p->cap_tilebuf = 0;
        if (cudaMalloc(&p->d_tilebuf, (size_t)need) != cudaSuccess) { cudaGetLastError(); return -1; }
        p->cap_tilebuf = need;
    }
    if constexpr (std::is_same<T, float>::value && MT <= 4) {
        if (tpw && item_hi > item_lo) {
            const int st = (win2d_ok<MT>(geo) && p->kernel_mode != 7) ? spread_win2d_go<MT>(p, fhat, B, item_lo, item_hi, geo)
                         : tpw == 4 ? spread_batch2d_go<MT, 4>(p, fhat, B, item_lo, item_hi, geo)
                         : tpw == 2 ? spread_batch2d_go<MT, 2>(p, fhat, B, item_lo, item_hi, geo)
                                    : spread_batch2d_go<MT, 1>(p, fhat, B, item_lo, item_hi, geo);
            if (st != NFFTB200_OK) return st;
        }
    }
    if (tpw == 0 && item_hi > item_lo) {
        auto kern = k_spread_sub2d<T, MT>;
        CUDA_TRY(p, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        dim3 grid(item_hi - item_lo, B);
        kern<<<grid, T2_THREADS, smem, p->stream>>>((const C*)fhat, (C*)p->d_tilebuf, (const T*)p->d_xs, p->d_perm,
                                                   p->d_items, p->d_item_stride, item_lo, p->M, geo, make_win<T>(p), make_poly_param<T, MT>(p));
    }
    bool pairs = false;
    if constexpr (std::is_same<T, float>::value && MT % 2 == 0) {
        // every tile at least 2m wide (so a cell has at most one neighbour tile per dimension), everything even
        pairs = geo.bs[0] % 2 == 0 && geo.Nt[0] % 2 == 0 && geo.bs[0] >= 2 * MT && geo.bs[1] >= 2 * MT &&
                geo.Nt[0] - (geo.nb[0] - 1) * geo.bs[0] >= 2 * MT && geo.Nt[1] - (geo.nb[1] - 1) * geo.bs[1] >= 2 * MT && p->kernel_mode != 9;
        if (pairs && B >= 8 && p->ntiles <= 0x7fffffff / 8 && p->kernel_mode != 7) {
            const int bper = 8;
            dim3 gg((unsigned)p->ntiles, (B + bper - 1) / bper);
            k_gather_tile2d<MT><<<gg, 128, 0, p->stream>>>((const float2*)p->d_tilebuf, (float2*)g, p->d_tile_items, t_lo, t_hi, item_lo, item_hi,
                                                           B, bper, geo);
        } else if (pairs) {
            dim3 gg((geo.Nt[0] / 2 + 127) / 128, geo.Nt[1], B);
            k_gather_pairs2d<MT><<<gg, 128, 0, p->stream>>>((const float2*)p->d_tilebuf, (float2*)g, p->d_tile_items, t_lo, t_hi, item_lo, item_hi, geo);
        }
    }
    if (!pairs) {
        int bx = 32;
        while (bx < 256 && bx < geo.Nt[0]) bx <<= 1;
        dim3 gg((geo.Nt[0] + bx - 1) / bx, geo.Nt[1], B);
        k_gather_tiles2d<T, MT><<<gg, bx, 0, p->stream>>>((const C*)p->d_tilebuf, (C*)g, p->d_tile_items, t_lo, t_hi, item_lo, item_hi, geo);
    }
    p->launches += 2;
    CUDA_TRY(p, cudaGetLastError());
    return NFFTB200_OK;
}

template <typename T, int MT>
int interp2d_launch(nfftb200_plan* p, const void* g, void* fhat, int B, int t_lo, int t_hi)
{
    using C = typename Cplx<T>::type;
    GeomDev geo = make_geom<T>(p);
    const int item_lo = p->h_tile_items[(size_t)t_lo], item_hi = p->h_tile_items[(size_t)t_hi];
    if constexpr (std::is_same<T, float>::value && MT <= 4) {
        const int tpw = batch2d_tpw<MT>(p, geo, B);
        if (tpw) {
            if (item_hi == item_lo) return NFFTB200_OK;
            const int st = tpw == 4 ? interp_batch2d_go<MT, 4>(p, g, fhat, B, item_lo, item_hi, geo)
                         : tpw == 2 ? interp_batch2d_go<MT, 2>(p, g, fhat, B, item_lo, item_hi, geo)
                                    : interp_batch2d_go<MT, 1>(p, g, fhat, B, item_lo, item_hi, geo);
            if (st != NFFTB200_OK) return st;
            p->launches++;
            CUDA_TRY(p, cudaGetLastError());
            return NFFTB200_OK;
        }
    }
    Interp2<T, MT> lay(geo.bs);
    const size_t smem = lay.bytes();
    if (smem > 227 * 1024) return -1;
    auto kern = k_interp_row2d<T, MT>;
    CUDA_TRY(p, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (item_hi == item_lo) return NFFTB200_OK;
    dim3 grid(item_hi - item_lo, B);
    kern<<<grid, T2_THREADS, smem, p->stream>>>((const C*)g, (C*)fhat, (const T*)p->d_xs, p->d_perm, p->d_items, p->d_item_stride,
                                               item_lo, p->M, geo, make_win<T>(p), make_poly_param<T, MT>(p));
    p->launches++;
    CUDA_TRY(p, cudaGetLastError());
    return NFFTB200_OK;
}

}  // namespace

#define DISPATCH_M(FN, ...)                                 \
    switch (p->m) {                                         \
        case 2: return FN<T, 2>(__VA_ARGS__);               \
        case 3: return FN<T, 3>(__VA_ARGS__);               \
        case 4: return FN<T, 4>(__VA_ARGS__);               \
        case 5: return FN<T, 5>(__VA_ARGS__);               \
        case 6: return FN<T, 6>(__VA_ARGS__);               \
        default: return -1;                                 \
    }
template <typename T> static int spread2d_T(nfftb200_plan* p, const void* fhat, void* g, int B, int t_lo, int t_hi)
{
    DISPATCH_M(spread2d_launch, p, fhat, g, B, t_lo, t_hi)
}
template <typename T> static int interp2d_T(nfftb200_plan* p, const void* g, void* fhat, int B, int t_lo, int t_hi)
{
    DISPATCH_M(interp2d_launch, p, g, fhat, B, t_lo, t_hi)
}
// return -1 when the tiled 2-D kernels do not apply (complex data only)
int nfftb_spread_2d(nfftb200_plan* p, const void* fhat, void* g, int B, int t_lo, int t_hi)
{
    return p->dtype == NFFTB200_F32 ? spread2d_T<float>(p, fhat, g, B, t_lo, t_hi) : spread2d_T<double>(p, fhat, g, B, t_lo, t_hi);
}
int nfftb_interp_2d(nfftb200_plan* p, const void* g, void* fhat, int B, int t_lo, int t_hi)
{
    return p->dtype == NFFTB200_F32 ? interp2d_T<float>(p, g, fhat, B, t_lo, t_hi) : interp2d_T<double>(p, g, fhat, B, t_lo, t_hi);
}
