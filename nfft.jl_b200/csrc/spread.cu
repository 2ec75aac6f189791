// spread.cu -- K3: adjoint gridding  g[(off+l) mod Nt] += prod_d w_d[l_d] * fHat[j]
//   replaces convolve_transpose! -> _convolve_transpose_blocking! -> fillBlock!/fillOneNode!/addBlock!
//   (/root/reference/src/convolution.jl:115-140, :356-492).
//
// Two implementations:
//  * k_spread_generic: warp per node, taps strided over lanes, global vector REDs.  Any D<=3, any m<=8,
//    real or complex data.  Correctness baseline / fallback.
//  * k_spread_tile3d: one CTA per reference tile ("block"), padded sub-grid (bs+2m)^3 in shared memory.
//    Shared-memory float atomics are CAS spin loops on sm_100a (ATOMS.CAST.SPIN), so instead of
//    atomics every z-plane of the padded tile is OWNED by one warp (plane z -> warp z mod 8): for each
//    node the <=2m owner warps each add one (2m x 2m) plane of the footprint with plain LDS/FFMA/STS --
//    conflict-free by construction, deterministic order inside the tile.  Window weights are computed
//    once per (node, dim, tap) by all threads into a double-buffered shared-memory record array.
//    The finished tile (incl. halo) is flushed with one vector RED per cell (REDG.ADD.F32x2).
#include <type_traits>

#include "common.cuh"
#include "window.cuh"
#include "tile3d.cuh"
#include "spread_bin.cuh"
#include "gather3d.cuh"

namespace {

template <typename T> __device__ __forceinline__ void red_add(T* p, T v) { atomicAdd(p, v); }
__device__ __forceinline__ void red_add_c(float2* p, float2 v) { atomicAdd(p, v); }   // REDG.ADD.F32x2
__device__ __forceinline__ void red_add_c(double2* p, double2 v)
{
    atomicAdd(&p->x, v.x);
    atomicAdd(&p->y, v.y);
}

__device__ __forceinline__ int wrap(int v, int n)
{
    v %= n;
    return v < 0 ? v + n : v;
}

// ---------------------------------------------------------------------------------------
// generic: warp per node
// ---------------------------------------------------------------------------------------
template <typename T, bool CPLX>
__global__ void __launch_bounds__(256)
k_spread_generic(const void* __restrict__ fhat_, void* __restrict__ g_, const T* __restrict__ xs,
                 const int32_t* __restrict__ perm, long long i_lo, long long i_hi, long long M,
                 GeomDev geo, WinDev<T> win, int B)
{
    using C = typename Cplx<T>::type;
    __shared__ T s_w[8][NFFTB_MAX_D][2 * NFFTB_MAX_M];
    __shared__ int s_c[8][NFFTB_MAX_D];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int L = 2 * win.m, D = geo.D;
    const int ntaps = (D == 1) ? L : (D == 2 ? L * L : (D == 3 ? L * L * L : L * L * L * L));
    const long long nwarps = (long long)gridDim.x * 8;
    for (long long i = i_lo + (long long)blockIdx.x * 8 + warp; i < i_hi; i += nwarps) {
        __syncwarp();
        for (int q = lane; q < D * L; q += 32) {
            const int d = q / L, l = q - d * L;
            T ks;
            const int c = node_cell<T>(xs[i * D + d], geo.Nt[d], ks);
            s_w[warp][d][l] = node_tap<T>(win, ks, c, l);
            if (l == 0) s_c[warp][d] = c - win.m + 1;
        }
        __syncwarp();
        const long long j = perm[i];
        for (int q = lane; q < ntaps; q += 32) {
            int l0 = q % L, r = q / L;
            int l1 = r % L, l2 = (r / L) % L, l3 = r / (L * L);
            T w = s_w[warp][0][l0];
            long long cell = wrap(s_c[warp][0] + l0, geo.Nt[0]);
            if (D > 1) { w *= s_w[warp][1][l1]; cell += (long long)wrap(s_c[warp][1] + l1, geo.Nt[1]) * geo.Nt[0]; }
            if (D > 2) { w *= s_w[warp][2][l2]; cell += (long long)wrap(s_c[warp][2] + l2, geo.Nt[2]) * geo.Nt[0] * geo.Nt[1]; }
            if (D > 3) { w *= s_w[warp][3][l3]; cell += (long long)wrap(s_c[warp][3] + l3, geo.Nt[3]) * geo.Nt[0] * geo.Nt[1] * geo.Nt[2]; }
            for (int b = 0; b < B; b++) {
                if (CPLX) {
                    const C v = ((const C*)fhat_)[b * M + j];
                    red_add_c((C*)g_ + b * geo.gsz + cell, make_c<T>(w * v.x, w * v.y));
                } else {
                    const T v = ((const T*)fhat_)[b * M + j];
                    red_add<T>((T*)g_ + b * geo.gsz + cell, w * v);
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------
// tiled 3-D spreader: warp-private padded sub-tiles, row-per-lane accumulation
// ---------------------------------------------------------------------------------------
// nodes staged per pass (coordinates, values, octant ids in shared memory).  Float32: 768, so that a default 16^3
// tile of the C2 density (512 +- 23 nodes) is always ONE pass; Float64 keeps 512 (its sub-tiles leave less room).
template <typename T> struct SSChunk { static constexpr int value = sizeof(T) == 4 ? 768 : 512; };

// NW = 8: octants 2x2x2, one CTA per SM;  NW = 4: quadrants 2x2x1 (for tiles that are half as thick), two CTAs
// per SM so that one CTA's shared-memory-bound accumulation overlaps the other's weight/merge/flush phases.
template <typename T, int MT, int NW> struct SubLayout {
    using RG = RowGeom<T, MT>;
    static constexpr int L = 2 * MT;
    static constexpr int RW = ((RG::NWX + L + 2 * L) + 3) & ~3;     // record: wx(shifted) | wy | wz*v
    int SX, SY, SZ, QX, QY, QZ, QN;
    __host__ __device__ SubLayout(const int* bs)
    {
        SX = (bs[0] + 1) / 2; SY = (bs[1] + 1) / 2; SZ = (NW == 8) ? (bs[2] + 1) / 2 : bs[2];
        QX = SX + L; QY = SY + L; QZ = SZ + L;
        QN = (QX * QY * QZ + 2 * RG::VPC + 1) & ~1;                 // tail pad for the widened last row
    }
    __host__ __device__ size_t bytes() const
    {
        return sizeof(typename Cplx<T>::type) * (size_t)NW * QN + sizeof(T) * NW * 32 * RW +
               sizeof(int) * NW * 32 + sizeof(unsigned short) * NW * 64 + SSChunk<T>::value +
               sizeof(T) * 3 * SSChunk<T>::value + sizeof(typename Cplx<T>::type) * SSChunk<T>::value + 16;
    }
};

template <typename T, int MT, bool SCRATCH, int NW>
__global__ void __launch_bounds__(NW * 32, NW == 4 ? 2 : 1)
k_spread_sub3d(const typename Cplx<T>::type* __restrict__ fhat, typename Cplx<T>::type* __restrict__ g,
               typename Cplx<T>::type* __restrict__ scratch,
               const T* __restrict__ xs, const int32_t* __restrict__ perm,
               const int32_t* __restrict__ tile_start, int tile_lo, long long M, GeomDev geo,
               WinDev<T> win, const __grid_constant__ PolyParam<T, MT> pp)
{
    using C = typename Cplx<T>::type;
    using RG = RowGeom<T, MT>;
    using SL = SubLayout<T, MT, NW>;
    constexpr int L = 2 * MT, VPC = RG::VPC, NV = RG::NV, NWX = RG::NWX, RW = SL::RW;
    constexpr int SS_WARPS = NW, SS_THREADS = NW * 32, SS_CHUNK = SSChunk<T>::value;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const SL lay(geo.bs);
    const int QX = lay.QX, QY = lay.QY, QN = lay.QN;
    C* sub = reinterpret_cast<C*>(smem_raw);                                    // [8][QN]
    T* rec_w = reinterpret_cast<T*>(sub + SS_WARPS * QN);                       // [8][32][RW]
    int* rec_b = reinterpret_cast<int*>(rec_w + SS_WARPS * 32 * RW);            // [8][32]
    unsigned short* list = reinterpret_cast<unsigned short*>(rec_b + SS_WARPS * 32);   // [8][64]
    unsigned char* oct = reinterpret_cast<unsigned char*>(list + SS_WARPS * 64);      // [SS_CHUNK]
    C* s_v = reinterpret_cast<C*>(oct + SS_CHUNK);                                     // [SS_CHUNK] node values (staged)
    T* s_x = reinterpret_cast<T*>(s_v + SS_CHUNK);                                     // [SS_CHUNK][3] shifted coordinates (staged)

    // blockIdx.x walks work items (tile, node range); tile_start is the item table here
    const int32_t* item = tile_start + 3 * (size_t)(tile_lo + blockIdx.x);
    const int tile_id = item[0];
    const int n_lo = item[1], n_hi = item[2];
    const int tx = tile_id % geo.nb[0];
    const int ty = (tile_id / geo.nb[0]) % geo.nb[1];
    const int tz = tile_id / (geo.nb[0] * geo.nb[1]);
    const int cx0 = tx * geo.bs[0], cy0 = ty * geo.bs[1], cz0 = tz * geo.bs[2];  // first core cell
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int ob0 = warp & 1, ob1 = (warp >> 1) & 1, ob2 = warp >> 2;
    fhat += (long long)blockIdx.y * M;
    g += (long long)blockIdx.y * geo.gsz;
    if (SCRATCH) {
        const int PXs = geo.bs[0] + L, PYs = geo.bs[1] + L, PZs = geo.bs[2] + L;
        scratch += ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * ((size_t)PXs * PYs * PZs);
    }
    C* mysub = sub + warp * QN;
    T* myrec = rec_w + warp * 32 * RW;
    int* mybase = rec_b + warp * 32;
    unsigned short* mylist = list + warp * 64;

    // per-lane constant row geometry
    int rowoff[RG::FULL_IT > 0 ? RG::FULL_IT : 1], wyo[RG::FULL_IT > 0 ? RG::FULL_IT : 1],
        vzo[RG::FULL_IT > 0 ? RG::FULL_IT : 1];
#pragma unroll
    for (int it = 0; it < RG::FULL_IT; it++) {
        const int r = lane + 32 * it, t = r / L, yt = r - t * L;
        rowoff[it] = (t * QY + yt) * QX; wyo[it] = NWX + yt; vzo[it] = NWX + L + 2 * t;
    }
    int rem_off = 0, rem_wy = 0, rem_vz = 0, rem_u = 0;
    bool rem_on = false;
    if (RG::REM > 0) {
        const int rr = RG::SPLIT ? lane / NV : lane;
        rem_u = RG::SPLIT ? lane - rr * NV : 0;
        rem_on = rr < RG::REM;
        const int r = RG::FULL_IT * 32 + (rem_on ? rr : 0), t = r / L, yt = r - t * L;
        rem_off = (t * QY + yt) * QX + rem_u * VPC; rem_wy = NWX + yt; rem_vz = NWX + L + 2 * t;
    }
    const unsigned lt = (1u << lane) - 1u;

    auto process_round = [&](int cbase, int nn) {
        // ---- phase A: lane-per-node weights and record
        if (lane < nn) {
            const int q = mylist[lane];                       // chunk-local node index: everything is staged in smem
            T ks0, ks1, ks2;
            const int c0 = node_cell<T>(s_x[q * 3 + 0], geo.Nt[0], ks0);
            const int c1 = node_cell<T>(s_x[q * 3 + 1], geo.Nt[1], ks1);
            const int c2 = node_cell<T>(s_x[q * 3 + 2], geo.Nt[2], ks2);
            T w0[L], w1[L], w2[L];
            eval_taps<T, MT>(win, pp, ks0, c0, w0);
            eval_taps<T, MT>(win, pp, ks1, c1, w1);
            eval_taps<T, MT>(win, pp, ks2, c2, w2);
            const C v = s_v[q];
            const int px = c0 - cx0 - ob0 * lay.SX + 1, py = c1 - cy0 - ob1 * lay.SY + 1,
                      pz = c2 - cz0 - ob2 * lay.SZ + 1;                          // first tap, padded sub-tile coords
            const int s = (VPC == 2) ? (px & 1) : 0;
            mybase[lane] = (((pz * QY + py) * QX + (px - s)) << 1) | s;      // low bit: row starts on an odd cell
            T r[RW];
#pragma unroll
            for (int k = 0; k < NWX; k++) {
                const T lo = (k < L) ? w0[k < L ? k : 0] : (T)0;
                const T hi = (k >= 1 && k - 1 < L) ? w0[(k >= 1 && k - 1 < L) ? k - 1 : 0] : (T)0;
                r[k] = s ? hi : lo;
            }
#pragma unroll
            for (int k = 0; k < L; k++) r[NWX + k] = w1[k];
#pragma unroll
            for (int k = 0; k < L; k++) { r[NWX + L + 2 * k] = w2[k] * v.x; r[NWX + L + 2 * k + 1] = w2[k] * v.y; }
#pragma unroll
            for (int k = NWX + 3 * L; k < RW; k++) r[k] = (T)0;
            T* dst = myrec + lane * RW;
#pragma unroll
            for (int k = 0; k < RW; k++) dst[k] = r[k];
        }
        __syncwarp();
        // ---- phase B: one node at a time, every lane owns one x-row of the footprint.  The record of
        //      node n+1 is fetched while node n is accumulated (software pipelining of the dependent loads).
        constexpr int NIT = RG::FULL_IT > 0 ? RG::FULL_IT : 1;
        T wx[NWX], fa[NIT], fb[NIT], ra = 0, rb = 0, rwu[VPC];
        int base;
        auto fetch = [&](int n, T (&wx_)[NWX], T (&fa_)[NIT], T (&fb_)[NIT], T& ra_, T& rb_, T (&wu_)[VPC], int& base_) {
            const T* rw = myrec + n * RW;
            base_ = mybase[n];
#pragma unroll
            for (int k = 0; k < NWX; k++) wx_[k] = rw[k];
#pragma unroll
            for (int it = 0; it < RG::FULL_IT; it++) {
                const T wyv = rw[wyo[it]];
                fa_[it] = wyv * rw[vzo[it]]; fb_[it] = wyv * rw[vzo[it] + 1];
            }
            if (RG::REM > 0) {
                const T wyv = rw[rem_wy];
                ra_ = wyv * rw[rem_vz]; rb_ = wyv * rw[rem_vz + 1];
                if (RG::SPLIT) {
#pragma unroll
                    for (int k = 0; k < VPC; k++) wu_[k] = rw[rem_u * VPC + k];
                }
            }
        };
        fetch(0, wx, fa, fb, ra, rb, rwu, base);
        for (int n = 0; n < nn; n++) {
            C* p0 = mysub + (base >> 1);
            T wx2[NWX], fa2[NIT], fb2[NIT], ra2 = 0, rb2 = 0, rwu2[VPC];
            int base2 = base;
            if (RG::FULL_IT == 0) fetch(n + 1 < nn ? n + 1 : n, wx2, fa2, fb2, ra2, rb2, rwu2, base2);
            // Float32 rows that start on an even cell need one 16-byte unit less (the widened unit would only
            // carry zero weights): node-uniform branch, saves a quarter of the shared-memory traffic of the row
            auto full_iters = [&](auto nvu_tag) {
                constexpr int NVU = decltype(nvu_tag)::value;
#pragma unroll
                for (int it = 0; it < RG::FULL_IT; it++) {
                    Unit<T> U[NVU];
#pragma unroll
                    for (int u = 0; u < NVU; u++) U[u].load(p0 + rowoff[it] + u * VPC);
                    if (it == 0) fetch(n + 1 < nn ? n + 1 : n, wx2, fa2, fb2, ra2, rb2, rwu2, base2);
#pragma unroll
                    for (int u = 0; u < NVU; u++) { U[u].axpy(&wx[u * VPC], fa[it], fb[it]); U[u].store(p0 + rowoff[it] + u * VPC); }
                }
            };
            if (VPC == 2 && NV > 1 && !(base & 1)) full_iters(std::integral_constant<int, (NV > 1 ? NV - 1 : 1)>{});
            else full_iters(std::integral_constant<int, NV>{});
            if (RG::REM > 0 && rem_on) {
                if (RG::SPLIT) { Unit<T> U; U.load(p0 + rem_off); U.axpy(rwu, ra, rb); U.store(p0 + rem_off); }
                else {
#pragma unroll
                    for (int u = 0; u < NV; u++) { Unit<T> U; U.load(p0 + rem_off + u * VPC); U.axpy(&wx[u * VPC], ra, rb); U.store(p0 + rem_off + u * VPC); }
                }
            }
#pragma unroll
            for (int k = 0; k < NWX; k++) wx[k] = wx2[k];
#pragma unroll
            for (int it = 0; it < NIT; it++) { fa[it] = fa2[it]; fb[it] = fb2[it]; }
            ra = ra2; rb = rb2; base = base2;
#pragma unroll
            for (int k = 0; k < VPC; k++) rwu[k] = rwu2[k];
            __syncwarp();
        }
    };

    for (int cbase = n_lo; cbase < n_hi; cbase += SS_CHUNK) {
        const int nc = min(SS_CHUNK, n_hi - cbase);
        __syncthreads();                                   // staging arrays free
        // stage the chunk: issue all global loads (coordinates, permutation -> value gather) first, zero the
        // sub-tiles while they are in flight (first chunk), then consume them -- phase A has no global loads left
        constexpr int NPT = SS_CHUNK / SS_THREADS;         // nodes per thread
        T rx[NPT][3];
        C rv[NPT];
#pragma unroll
        for (int k = 0; k < NPT; k++) {
            const int q = threadIdx.x + k * SS_THREADS;
            if (q < nc) {
                const long long i = (long long)cbase + q;
                rx[k][0] = xs[i * 3 + 0]; rx[k][1] = xs[i * 3 + 1]; rx[k][2] = xs[i * 3 + 2];
                rv[k] = fhat[perm[i]];
            }
        }
        if (cbase == n_lo) {   // zero all sub-tiles with 16-byte stores
            uint4* z = reinterpret_cast<uint4*>(sub);
            const int n16 = (int)((sizeof(C) * (size_t)SS_WARPS * QN) / 16);
            for (int q = threadIdx.x; q < n16; q += SS_THREADS) z[q] = make_uint4(0, 0, 0, 0);
        }
#pragma unroll
        for (int k = 0; k < NPT; k++) {
            const int q = threadIdx.x + k * SS_THREADS;
            if (q < nc) {
                T ks;
                const int l0 = node_cell<T>(rx[k][0], geo.Nt[0], ks) - cx0;
                const int l1 = node_cell<T>(rx[k][1], geo.Nt[1], ks) - cy0;
                const int l2 = node_cell<T>(rx[k][2], geo.Nt[2], ks) - cz0;
                oct[q] = (unsigned char)((l0 >= lay.SX) + 2 * (l1 >= lay.SY) + ((NW == 8) ? 4 * (l2 >= lay.SZ) : 0));
                s_x[q * 3 + 0] = rx[k][0]; s_x[q * 3 + 1] = rx[k][1]; s_x[q * 3 + 2] = rx[k][2];
                s_v[q] = rv[k];
            }
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

    // ---- hierarchical merge of the 8 warp-private sub-tiles (fixed order => deterministic), then flush.
    //      Octant o = ox + 2*oy + 4*oz covers padded-tile cells [o_d*S_d, o_d*S_d + S_d + 2m).  The overlap of
    //      the low octant is folded into the high one, one dimension at a time (bulk shared-memory adds):
    //        z: planes [SZ,QZ) of (ox,oy,0)  -> planes [0,2m) of (ox,oy,1)
    //        y: rows   [SY,QY) of (ox,0,oz)  -> rows   [0,2m) of (ox,1,oz)   (live planes only)
    //        x: cells  [SX,QX) of (0,oy,oz)  -> cells  [0,2m) of (1,oy,oz)   (live rows only)
    //      after which every padded-tile cell lives in exactly one sub-tile.
    {
        const int SX = lay.SX, SY = lay.SY, SZ = lay.SZ, QZ = lay.QZ;
        const int PX = geo.bs[0] + L, PY = geo.bs[1] + L, PZ = geo.bs[2] + L;
        // every fold adds whole runs of cells; runs are moved as 16-byte units when the geometry keeps them aligned
        // (always true for the default tiles: even QX, even SX)
        constexpr int VC = 16 / (int)sizeof(C);                   // cells per 16-byte unit
        const bool vec = (VC == 1) || ((QX % 2 == 0) && (SX % 2 == 0) && (L % 2 == 0));
        auto fold = [&](C* d, const C* a) {                       // *d += *a for one 16-byte unit
            if (VC == 2) {
                float4 x = *reinterpret_cast<const float4*>(a), y = *reinterpret_cast<float4*>(d);
                y.x += x.x; y.y += x.y; y.z += x.z; y.w += x.w;
                *reinterpret_cast<float4*>(d) = y;
            } else { C c = *d; c.x += a->x; c.y += a->y; *d = c; }
        };
        auto fold1 = [&](C* d, const C* a) { C c = *d; c.x += a->x; c.y += a->y; *d = c; };
        if (NW == 8) {   // z
            const int blk = L * QY * QX;
            if (vec) {
                const int nb = blk / VC;
                for (int q = threadIdx.x; q < 4 * nb; q += SS_THREADS) {
                    const int col = q / nb, r = (q - col * nb) * VC;
                    fold(sub + (col + 4) * QN + r, sub + col * QN + SZ * QY * QX + r);
                }
            } else {
                for (int q = threadIdx.x; q < 4 * blk; q += SS_THREADS) {
                    const int col = q / blk, r = q - col * blk;
                    fold1(sub + (col + 4) * QN + r, sub + col * QN + SZ * QY * QX + r);
                }
            }
        }
        if (NW == 8) __syncthreads();
        {   // y: live planes of octant oz: oz==0 -> [0,SZ), oz==1 -> [0,QZ)   (NW == 4: a single z layer, all planes live)
            const int blk = L * QX, npl = (NW == 8) ? SZ + QZ : QZ;
            const int step = vec ? VC : 1, nb = blk / step;
            const unsigned inv = fastdiv_inv(nb);
            for (int q = threadIdx.x; q < 2 * npl * nb; q += SS_THREADS) {
                const int pb = (int)fastdiv(q, inv), r = (q - pb * nb) * step;
                const int ox_ = pb >= npl, pl = pb - ox_ * npl;
                const int oz_ = (NW == 8) ? (pl >= SZ) : 0, zz = pl - oz_ * SZ;
                const int o = ox_ + 4 * oz_;
                const C* a = sub + o * QN + (zz * QY + SY) * QX + r;
                C* d = sub + (o + 2) * QN + zz * QY * QX + r;
                if (vec) fold(d, a); else fold1(d, a);
            }
        }
        __syncthreads();
        {   // x: live rows: merged (y,z) coordinates of the padded tile
            const int step = vec ? VC : 1, nx = L / step;
            const unsigned invL = fastdiv_inv(nx), invPY = fastdiv_inv(PY);
            for (int q = threadIdx.x; q < PY * PZ * nx; q += SS_THREADS) {
                const int row = (int)fastdiv(q, invL), xx = (q - row * nx) * step;
                const int z = (int)fastdiv(row, invPY), y = row - z * PY;
                const int oy_ = y >= SY, oz_ = (NW == 8) ? (z >= SZ) : 0;
                const int o = 2 * oy_ + 4 * oz_;
                const int ro = ((z - oz_ * SZ) * QY + (y - oy_ * SY)) * QX;
                const C* a = sub + o * QN + ro + SX + xx;
                C* d = sub + (o + 1) * QN + ro + xx;
                if (vec) fold(d, a); else fold1(d, a);
            }
        }
        __syncthreads();
        if (SCRATCH) {
            // the merged padded tile ("blocks[l]") leaves through the TMA engine: every (z,y) row is two bulk
            // asynchronous copies shared -> global (cells [0,SX) from the low-x octant, [SX,PX) from the high-x
            // one), issued by one thread per row; no register staging, no per-cell index math
            const unsigned b0 = (unsigned)(SX * sizeof(C)), b1 = (unsigned)((PX - SX) * sizeof(C));
            const bool bulk_ok = (b0 % 16 == 0) && (b1 % 16 == 0) && ((PX * sizeof(C)) % 16 == 0) && ((QX * sizeof(C)) % 16 == 0) &&
                                 ((QN * sizeof(C)) % 16 == 0);
            if (bulk_ok) {
                fence_async_smem();
                for (int row = threadIdx.x; row < PY * PZ; row += SS_THREADS) {
                    const int z = row / PY, y = row - z * PY;
                    const int oz_ = (NW == 8) ? (z >= SZ) : 0, oy_ = y >= SY;
                    const C* pa = sub + (4 * oz_ + 2 * oy_) * QN + ((z - oz_ * SZ) * QY + (y - oy_ * SY)) * QX;
                    C* sr = scratch + (size_t)row * PX;
                    bulk_store_s2g(sr, pa, b0);
                    bulk_store_s2g(sr + SX, pa + QN, b1);
                }
                bulk_commit_wait_read();
                return;
            }
        }
        // plain path (RED halo flush, or scratch rows that are not 16-byte granular)
        const bool fw = PX <= geo.Nt[0] && PY <= geo.Nt[1] && PZ <= geo.Nt[2];
        const int xa = lane, xb = lane + 32;
        const bool on0 = xa < PX, on1 = xb < PX;
        const int so0 = (xa < SX) ? xa : QN + xa - SX, so1 = (xb < SX) ? xb : QN + xb - SX;
        const int xg0 = wrapc(cx0 - MT + xa, geo.Nt[0], fw), xg1 = wrapc(cx0 - MT + xb, geo.Nt[0], fw);
        for (int z = 0; z < PZ; z++) {
            const int oz_ = (NW == 8) ? (z >= SZ) : 0;
            const C* pz_ = sub + 4 * oz_ * QN + (z - oz_ * SZ) * QY * QX;
            const unsigned gz = (unsigned)wrapc(cz0 - MT + z, geo.Nt[2], fw) * geo.Nt[1];
#pragma unroll
            for (int k = 0; k < 32 / SS_WARPS; k++) {
                const int y = warp + SS_WARPS * k;
                if (y < PY) {
                    const int oy_ = y >= SY;
                    const C* pa = pz_ + 2 * oy_ * QN + (y - oy_ * SY) * QX;
                    if (SCRATCH) {
                        C* sr = scratch + (z * PY + y) * PX;
                        if (on0) sr[xa] = pa[so0];
                        if (on1) sr[xb] = pa[so1];
                    } else {
                        C* gr = g + (gz + wrapc(cy0 - MT + y, geo.Nt[1], fw)) * (unsigned)geo.Nt[0];
                        if (on0) { const C c = pa[so0]; if (c.x != (T)0 || c.y != (T)0) red_add_c(gr + xg0, c); }
                        if (on1) { const C c = pa[so1]; if (c.x != (T)0 || c.y != (T)0) red_add_c(gr + xg1, c); }
                    }
                }
            }
        }
    }
}

template <typename T, int MT>
__global__ void __launch_bounds__(256)
k_gather_tiles3d(const typename Cplx<T>::type* __restrict__ scratch, typename Cplx<T>::type* __restrict__ g,
                 const int32_t* __restrict__ tile_items, int tile_lo, int tile_hi, int item_lo, int item_hi, GeomDev geo)
{
    using C = typename Cplx<T>::type;
    constexpr int L = 2 * MT;
    const int PX = geo.bs[0] + L, PY = geo.bs[1] + L, PZ = geo.bs[2] + L;
    const unsigned PN = (unsigned)PX * PY * PZ;
    const int u0 = (blockIdx.x * blockDim.x + threadIdx.x) * 2;
    const int u1 = blockIdx.y, u2 = blockIdx.z % geo.Nt[2], b = blockIdx.z / geo.Nt[2];
    if (u0 >= geo.Nt[0]) return;
    scratch += (size_t)b * (size_t)(item_hi - item_lo) * PN;
    // candidates along one dimension: (tile coordinate * tmul, padded offset * pmul)
    auto cover = [&](int u, int d, int tmul, int pmul, int (&tt)[3], int (&po)[3]) -> int {
        const int bs = geo.bs[d], nb = geo.nb[d], Nt = geo.Nt[d];
        const int t = (int)fastdiv((unsigned)u, geo.inv_bs[d]), l = u - t * bs;
        const int len = (t == nb - 1) ? Nt - t * bs : bs;
        int n = 0;
        tt[n] = t * tmul; po[n] = (l + MT) * pmul; n++;
        if (l < MT) {
            const int tp = t == 0 ? nb - 1 : t - 1;
            const int lenp = (tp == nb - 1) ? Nt - tp * bs : bs;
            tt[n] = tp * tmul; po[n] = (l + MT + lenp) * pmul; n++;
        }
        if (l >= len - MT) { tt[n] = (t == nb - 1 ? 0 : t + 1) * tmul; po[n] = (l + MT - len) * pmul; n++; }
        return n;
    };
    int ty[3], oy[3], tz[3], oz[3];
    const int ny = cover(u1, 1, geo.nb[0], PX, ty, oy);
    const int nz = cover(u2, 2, geo.nb[0] * geo.nb[1], PX * PY, tz, oz);
    // x: both cells of the pair lie in the same tile (bs[0] even, u0 even)
    const int bs0 = geo.bs[0], nb0 = geo.nb[0];
    const int t = (int)fastdiv((unsigned)u0, geo.inv_bs[0]), l = u0 - t * bs0;
    const int len = (t == nb0 - 1) ? geo.Nt[0] - t * bs0 : bs0;
    const int tp = t == 0 ? nb0 - 1 : t - 1, tn = t == nb0 - 1 ? 0 : t + 1;
    const int lenp = (tp == nb0 - 1) ? geo.Nt[0] - tp * bs0 : bs0;
    const bool p0 = l < MT, p1 = l + 1 < MT;                       // previous tile's high halo covers cell 0 / 1
    const bool n0 = l >= len - MT, n1 = l + 1 >= len - MT && l + 1 < len;   // next tile's low halo
    const bool c1 = l + 1 < len;                                   // second cell inside this tile's core
    T a0x = 0, a0y = 0, a1x = 0, a1y = 0;
    auto add_tile = [&](int tile, int off, bool w0, bool w1) {
        if (tile < tile_lo || tile >= tile_hi) return;
        const int ia = tile_items[tile], ib = tile_items[tile + 1];
        for (int it = ia; it < ib; it++) {
            const C* sp = scratch + ((long long)(it - item_lo) * (long long)PN + off);
            if (w0) { const C c = sp[0]; a0x += c.x; a0y += c.y; }
            if (w1) { const C c = sp[1]; a1x += c.x; a1y += c.y; }
        }
    };
    for (int iz = 0; iz < nz; iz++)
        for (int iy = 0; iy < ny; iy++) {
            const int tyz = tz[iz] + ty[iy];
            const int oyz = oz[iz] + oy[iy];
            add_tile(tyz + t, oyz + l + MT, true, c1);
            if (p0) add_tile(tyz + tp, oyz + l + MT + lenp, true, p1);
            if (n1 || n0) add_tile(tyz + tn, oyz + l + MT - len, n0, n1);
        }
    // a second cell beyond the core of a partial last tile belongs to the next tile (only if bs[0] is odd there)
    C* dst = g + (size_t)b * geo.gsz + ((size_t)u2 * geo.Nt[1] + u1) * geo.Nt[0] + u0;
    if (sizeof(T) == 4) *reinterpret_cast<float4*>(dst) = make_float4((float)a0x, (float)a0y, (float)a1x, (float)a1y);
    else { dst[0] = make_c<T>(a0x, a0y); dst[1] = make_c<T>(a1x, a1y); }
}

// returns -1 if the tiled kernel does not apply, else a status; *wrote_all = true if every grid cell was
// written by the gather pass (no memset needed)
template <typename T, int MT, int NW>
int launch_tile3d_nw(nfftb200_plan* p, const void* fhat, void* g, int B, int t_lo, int t_hi)
{
    using C = typename Cplx<T>::type;
    constexpr int SS_THREADS = NW * 32;
    GeomDev geo = make_geom<T>(p);
    SubLayout<T, MT, NW> lay(geo.bs);
    const size_t smem = lay.bytes();
    if (smem > 227 * 1024 || geo.bs[0] + 2 * MT > 64 || geo.bs[1] + 2 * MT > 32) return -1;
    bool use_scratch = p->kernel_mode != 2;
    for (int d = 0; d < 3; d++) {
        const int last = geo.Nt[d] - (geo.nb[d] - 1) * geo.bs[d];
        if (geo.bs[d] < MT || last < MT) use_scratch = false;        // halos would reach past the neighbour
        if (d == 0 && ((geo.bs[0] & 1) || (last & 1))) use_scratch = false;   // the gather works on x cell pairs
    }
    const size_t PN = (size_t)(geo.bs[0] + 2 * MT) * (geo.bs[1] + 2 * MT) * (geo.bs[2] + 2 * MT);
    const cudaStream_t st = p->stream;
    const int item_lo = p->h_tile_items[(size_t)t_lo], item_hi = p->h_tile_items[(size_t)t_hi];
    if (item_hi == item_lo) use_scratch = false;
    dim3 grid(item_hi - item_lo, B);
    if (use_scratch) {
        const int64_t need = (int64_t)(sizeof(C) * PN * (size_t)(item_hi - item_lo) * B);
        if (need > p->cap_tilebuf) {
            if (p->d_tilebuf) cudaFree(p->d_tilebuf);
            p->d_tilebuf = nullptr; p->cap_tilebuf = 0;
            if (cudaMalloc(&p->d_tilebuf, (size_t)need) != cudaSuccess) { cudaGetLastError(); use_scratch = false; }
            else p->cap_tilebuf = need;
        }
    }
    struct KT {
        nfftb200_plan* p;
        explicit KT(nfftb200_plan* q) : p(q) { if (p->timing) cudaEventRecord(p->evk[1], p->stream); }
        ~KT() { if (p->timing) { cudaEventRecord(p->evk[2], p->stream); p->pending_k |= 1; } }
    };
    p->have_gather_ev = false;
    if (use_scratch) {
        if (p->timing) cudaEventRecord(p->evk[0], st);
        KT kt(p);
        auto kern = k_spread_sub3d<T, MT, true, NW>;
        CUDA_TRY(p, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<grid, SS_THREADS, smem, st>>>((const C*)fhat, (C*)g, (C*)p->d_tilebuf, (const T*)p->d_xs, p->d_perm,
                                            p->d_items, item_lo, p->M, geo, make_win<T>(p), make_poly_param<T, MT>(p));
        if (p->timing) { cudaEventRecord(p->evk[5], st); p->have_gather_ev = true; }
        const int units = geo.Nt[0] / 2;                                   // Nt[0] is even; one thread per cell pair
        int bx = 32;
        if (p->kernel_mode != 5 && geo.bs[2] == 16 && geo.Nt[2] % 16 == 0 && 16 >= 2 * MT) {
            while (bx < 128 && bx < units) bx <<= 1;
            dim3 gc((units + bx - 1) / bx, geo.Nt[1], geo.nb[2] * B * 2);
            k_gather_cols3d<T, MT, 16, false, 2><<<gc, bx, 0, st>>>((const C*)p->d_tilebuf, (C*)g, p->d_tile_items, t_lo, t_hi, item_lo, item_hi, geo, PeerTab{}, 0);
            p->launches += 2;
            CUDA_TRY(p, cudaGetLastError());
            return NFFTB200_OK;
        }
        while (bx < 256 && bx < units) bx <<= 1;
        dim3 gg((units + bx - 1) / bx, geo.Nt[1], geo.Nt[2] * B);
        k_gather_tiles3d<T, MT><<<gg, bx, 0, st>>>((const C*)p->d_tilebuf, (C*)g, p->d_tile_items, t_lo, t_hi, item_lo, item_hi, geo);
        p->launches += 2;
    } else {
        if (p->timing) cudaEventRecord(p->evk[0], st);
        CUDA_TRY(p, cudaMemsetAsync(g, 0, sizeof(C) * (size_t)p->gsz * B, st));
        KT kt(p);
        auto kern = k_spread_sub3d<T, MT, false, NW>;
        CUDA_TRY(p, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        if (item_hi > item_lo)
            kern<<<grid, SS_THREADS, smem, st>>>((const C*)fhat, (C*)g, nullptr, (const T*)p->d_xs, p->d_perm,
                                                p->d_items, item_lo, p->M, geo, make_win<T>(p), make_poly_param<T, MT>(p));
        p->launches += 2;
    }
    CUDA_TRY(p, cudaGetLastError());
    return NFFTB200_OK;
}

// kernel_mode 7: register-footprint spreader (spread_bin.cuh) + the same scratch layout and gather pass.  Returns -1
// when it does not apply (the caller then runs the default tiled kernel).
template <typename T, int MT, int W>
int launch_bin3d(nfftb200_plan* p, const void* fhat, void* g, int B, int t_lo, int t_hi)
{
    using C = typename Cplx<T>::type;
    using BL = BinLayout<T, MT, W>;
    GeomDev geo = make_geom<T>(p);
    BinGeom bg;
    if (!BL::make(geo.bs, bg)) return -1;
    const size_t smem = BL::bytes(bg);
    if (smem > 227 * 1024) return -1;
    for (int d = 0; d < 3; d++) {
        const int last = geo.Nt[d] - (geo.nb[d] - 1) * geo.bs[d];
        if (geo.bs[d] < MT || last < MT) return -1;                   // halos would reach past the neighbour
        if (d == 0 && ((geo.bs[0] & 1) || (last & 1))) return -1;     // the gather works on x cell pairs
    }
    const size_t PN = (size_t)(geo.bs[0] + 2 * MT) * (geo.bs[1] + 2 * MT) * (geo.bs[2] + 2 * MT);
    const cudaStream_t st = p->stream;
    const int item_lo = p->h_tile_items[(size_t)t_lo], item_hi = p->h_tile_items[(size_t)t_hi];
    if (item_hi == item_lo) return -1;
    const int64_t need = (int64_t)(sizeof(C) * PN * (size_t)(item_hi - item_lo) * B);
    if (need > p->cap_tilebuf) {
        if (p->d_tilebuf) cudaFree(p->d_tilebuf);
        p->d_tilebuf = nullptr; p->cap_tilebuf = 0;
        if (cudaMalloc(&p->d_tilebuf, (size_t)need) != cudaSuccess) { cudaGetLastError(); return -1; }
        p->cap_tilebuf = need;
    }
    p->have_gather_ev = false;
    if (p->timing) { cudaEventRecord(p->evk[0], st); cudaEventRecord(p->evk[1], st); }
    auto kern = k_spread_bin3d<T, MT, W>;
    CUDA_TRY(p, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    // two CTAs per SM need the largest shared-memory carve-out (a hint; the kernel is correct without it)
    cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    kern<<<dim3(item_hi - item_lo, B), NFFTB_BIN_WARPS * 32, smem, st>>>((const C*)fhat, (C*)p->d_tilebuf, (const T*)p->d_xs, p->d_perm,
                                                                        p->d_items, item_lo, p->M, geo, make_win<T>(p),
                                                                        make_poly_param<T, MT>(p), bg);
    if (p->timing) { cudaEventRecord(p->evk[5], st); p->have_gather_ev = true; }
    const int units = geo.Nt[0] / 2;
    int bx = 32;
    if (geo.bs[2] == 16 && geo.Nt[2] % 16 == 0 && 16 >= 2 * MT) {
        while (bx < 128 && bx < units) bx <<= 1;
        dim3 gc((units + bx - 1) / bx, geo.Nt[1], geo.nb[2] * B * 2);
        k_gather_cols3d<T, MT, 16, false, 2><<<gc, bx, 0, st>>>((const C*)p->d_tilebuf, (C*)g, p->d_tile_items, t_lo, t_hi, item_lo, item_hi, geo, PeerTab{}, 0);
    } else {
        while (bx < 256 && bx < units) bx <<= 1;
        dim3 gg((units + bx - 1) / bx, geo.Nt[1], geo.Nt[2] * B);
        k_gather_tiles3d<T, MT><<<gg, bx, 0, st>>>((const C*)p->d_tilebuf, (C*)g, p->d_tile_items, t_lo, t_hi, item_lo, item_hi, geo);
    }
    if (p->timing) { cudaEventRecord(p->evk[2], st); p->pending_k |= 1; }
    p->launches += 2;
    CUDA_TRY(p, cudaGetLastError());
    return NFFTB200_OK;
}


// ---- node sharding over peer memory (comm.cu): spread the own tile range into an IPC-exported scratch, then
// gather the own z-slab from every rank's scratch
template <typename T, int MT>
int peer_cells(nfftb200_plan* p)
{
    GeomDev geo = make_geom<T>(p);
    SubLayout<T, MT, 8> lay(geo.bs);
    if (lay.bytes() > 227 * 1024 || geo.bs[0] + 2 * MT > 64 || geo.bs[1] + 2 * MT > 32) return 0;
    if (geo.bs[2] != 16 || geo.Nt[2] % 16 != 0 || 16 < 2 * MT) return 0;
    for (int d = 0; d < 3; d++) {
        const int last = geo.Nt[d] - (geo.nb[d] - 1) * geo.bs[d];
        if (geo.bs[d] < MT || last < MT) return 0;
        if (d == 0 && ((geo.bs[0] & 1) || (last & 1))) return 0;
    }
    return (geo.bs[0] + 2 * MT) * (geo.bs[1] + 2 * MT) * (geo.bs[2] + 2 * MT);
}

template <typename T, int MT>
int peer_spread(nfftb200_plan* p, const void* fhat, void* scratch, int t_lo, int t_hi)
{
    using C = typename Cplx<T>::type;
    GeomDev geo = make_geom<T>(p);
    SubLayout<T, MT, 8> lay(geo.bs);
    const size_t smem = lay.bytes();
    const int item_lo = p->h_tile_items[(size_t)t_lo], item_hi = p->h_tile_items[(size_t)t_hi];
    if (item_hi == item_lo) return NFFTB200_OK;
    if (sizeof(T) == 4 && nfftb_lean_mode(p)) {   // (tile, bin)-ordered register windows, same scratch layout
        const int r = nfftb_spread_lean(p, fhat, nullptr, scratch, 1, t_lo, t_hi);
        if (r >= 0) return r;
    }
    if (p->timing) { cudaEventRecord(p->evk[0], p->stream); cudaEventRecord(p->evk[1], p->stream); }
    if constexpr (MT <= 3) {          // opt-in register-footprint spreader (same scratch layout, same peer gather)
        BinGeom bg;
        using BL = BinLayout<T, MT, 8>;
        if (p->kernel_mode == 7 && BL::make(geo.bs, bg) && BL::bytes(bg) <= 227 * 1024) {
            auto kb = k_spread_bin3d<T, MT, 8>;
            CUDA_TRY(p, cudaFuncSetAttribute(kb, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BL::bytes(bg)));
            cudaFuncSetAttribute(kb, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
            kb<<<dim3(item_hi - item_lo, 1), NFFTB_BIN_WARPS * 32, BL::bytes(bg), p->stream>>>((const C*)fhat, (C*)scratch, (const T*)p->d_xs,
                                                                                             p->d_perm, p->d_items, item_lo, p->M, geo,
                                                                                             make_win<T>(p), make_poly_param<T, MT>(p), bg);
            if (p->timing) { cudaEventRecord(p->evk[2], p->stream); p->pending_k |= 1; }
            p->launches++;
            CUDA_TRY(p, cudaGetLastError());
            return NFFTB200_OK;
        }
    }
    auto kern = k_spread_sub3d<T, MT, true, 8>;
    CUDA_TRY(p, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<dim3(item_hi - item_lo, 1), 256, smem, p->stream>>>((const C*)fhat, nullptr, (C*)scratch, (const T*)p->d_xs, p->d_perm,
                                                               p->d_items, item_lo, p->M, geo, make_win<T>(p), make_poly_param<T, MT>(p));
    if (p->timing) { cudaEventRecord(p->evk[2], p->stream); p->pending_k |= 1; }
    p->launches++;
    CUDA_TRY(p, cudaGetLastError());
    return NFFTB200_OK;
}

template <typename T, int MT>
int peer_gather(nfftb200_plan* p, void* slab, int layer_lo, int nlayers, const PeerTab& pt)
{
    using C = typename Cplx<T>::type;
    GeomDev geo = make_geom<T>(p);
    const int units = geo.Nt[0] / 2;
    int bx = 32;
    while (bx < 128 && bx < units) bx <<= 1;
    dim3 gc((units + bx - 1) / bx, geo.Nt[1], nlayers * 2);
    k_gather_cols3d<T, MT, 16, true, 2><<<gc, bx, 0, p->stream>>>(nullptr, (C*)slab, p->d_tile_items, 0, 0, 0, 0, geo, pt, layer_lo);
    p->launches++;
    CUDA_TRY(p, cudaGetLastError());
    return NFFTB200_OK;
}

// thin tiles (bs_z <= 8) run with 4 warps / 2 CTAs per SM, thick ones with 8 warps / 1 CTA per SM
template <typename T, int MT>
int launch_tile3d(nfftb200_plan* p, const void* fhat, void* g, int B, int t_lo, int t_hi)
{
    if (p->bs[2] <= 8) {
        const int r = launch_tile3d_nw<T, MT, 4>(p, fhat, g, B, t_lo, t_hi);
        if (r >= 0) return r;
    }
    return launch_tile3d_nw<T, MT, 8>(p, fhat, g, B, t_lo, t_hi);
}

template <typename T>
int spread_impl(nfftb200_plan* p, const void* fhat, void* g, int B, int is_complex, int t_lo, int t_hi,
                long long i_lo, long long i_hi)
{
    const size_t cell = is_complex ? 2 * sizeof(T) : sizeof(T);
    if (nfftb_tiled_ok(p) && p->D == 1) {        // output-stationary 1-D spreader: writes every cell, no memset
        if (p->timing) { cudaEventRecord(p->evk[0], p->stream); cudaEventRecord(p->evk[1], p->stream); }
        const int r = nfftb_spread_1d(p, fhat, g, B, is_complex, t_lo, t_hi);
        if (r >= 0) {
            if (p->timing) { cudaEventRecord(p->evk[2], p->stream); p->pending_k |= 1; }
            return r;
        }
    }
    if (i_hi > i_lo && nfftb_tiled_ok(p) && is_complex && p->D == 2) {
        if (p->timing) { cudaEventRecord(p->evk[0], p->stream); cudaEventRecord(p->evk[1], p->stream); }
        const int r = nfftb_spread_2d(p, fhat, g, B, t_lo, t_hi);
        if (r >= 0) {
            if (p->timing) { cudaEventRecord(p->evk[2], p->stream); p->pending_k |= 1; }
            return r;
        }
    }
    if (sizeof(T) == 4 && i_hi > i_lo && nfftb_tiled_ok(p) && is_complex && p->D == 3 && nfftb_lean_mode(p)) {
        const int r = nfftb_spread_lean(p, fhat, g, nullptr, B, t_lo, t_hi);    // lean.cu: (tile, bin)-ordered register windows
        if (r >= 0) return r;
    }
    if (i_hi > i_lo && nfftb_tiled_ok(p) && is_complex && p->D == 3 && p->kernel_mode == 7) {
        int r = -1;                                    // opt-in register-footprint spreader (spread_bin.cuh)
        switch (p->m) {
            case 2: r = launch_bin3d<T, 2, 8>(p, fhat, g, B, t_lo, t_hi); break;
            case 3: r = launch_bin3d<T, 3, 8>(p, fhat, g, B, t_lo, t_hi); break;
            case 4: if constexpr (sizeof(T) == 4) r = launch_bin3d<T, 4, 10>(p, fhat, g, B, t_lo, t_hi); break;   // Float64: tile + records exceed 227 KB
            default: break;
        }
        if (r >= 0) return r;
    }
    if (i_hi > i_lo && nfftb_tiled_ok(p) && is_complex && p->D == 3) {
        int r = -1;
        switch (p->m) {
            case 2: r = launch_tile3d<T, 2>(p, fhat, g, B, t_lo, t_hi); break;
            case 3: r = launch_tile3d<T, 3>(p, fhat, g, B, t_lo, t_hi); break;
            case 4: r = launch_tile3d<T, 4>(p, fhat, g, B, t_lo, t_hi); break;
            case 5: r = launch_tile3d<T, 5>(p, fhat, g, B, t_lo, t_hi); break;
            case 6: r = launch_tile3d<T, 6>(p, fhat, g, B, t_lo, t_hi); break;
            default: break;
        }
        if (r >= 0) return r;
    }
    // generic path: zero the grid (memset(g), /root/reference/src/convolution.jl:358), then global REDs
    if (p->timing) cudaEventRecord(p->evk[0], p->stream);
    CUDA_TRY(p, cudaMemsetAsync(g, 0, cell * (size_t)p->gsz * B, p->stream));
    p->launches++;
    if (i_hi <= i_lo) return NFFTB200_OK;
    struct KernelTimer {
        nfftb200_plan* p;
        explicit KernelTimer(nfftb200_plan* q) : p(q) { if (p->timing) cudaEventRecord(p->evk[1], p->stream); }
        ~KernelTimer() { if (p->timing) { cudaEventRecord(p->evk[2], p->stream); p->pending_k |= 1; } }
    } kt(p);
    const long long n = i_hi - i_lo;
    const int blocks = (int)std::min<long long>((n + 7) / 8, 148 * 32);
    if (is_complex)
        k_spread_generic<T, true><<<blocks, 256, 0, p->stream>>>(fhat, g, (const T*)p->d_xs, p->d_perm, i_lo,
                                                                 i_hi, p->M, make_geom<T>(p), make_win<T>(p), B);
    else
        k_spread_generic<T, false><<<blocks, 256, 0, p->stream>>>(fhat, g, (const T*)p->d_xs, p->d_perm, i_lo,
                                                                  i_hi, p->M, make_geom<T>(p), make_win<T>(p), B);
    p->launches++;
    CUDA_TRY(p, cudaGetLastError());
    return NFFTB200_OK;
}

}  // namespace

// shared memory the tiled 3-D spreader needs for a given tile (0 = no tiled kernel for this m)
size_t nfftb_spread3d_smem(int dtype, int m, const int64_t* bs)
{
    int b[3] = {(int)bs[0], (int)bs[1], (int)bs[2]};
#define CASE_M(MM)                                                                                        \
    case MM:                                                                                              \
        if (b[2] <= 8) return dtype == NFFTB200_F32 ? SubLayout<float, MM, 4>(b).bytes() : SubLayout<double, MM, 4>(b).bytes(); \
        return dtype == NFFTB200_F32 ? SubLayout<float, MM, 8>(b).bytes() : SubLayout<double, MM, 8>(b).bytes();
    switch (m) {
        CASE_M(2) CASE_M(3) CASE_M(4) CASE_M(5) CASE_M(6)
        default: return 0;
    }
#undef CASE_M
}

#define PEER_DISPATCH(FN, ...)                                                                     \
    switch (p->m) {                                                                                \
        case 2: return p->dtype == NFFTB200_F32 ? FN<float, 2>(__VA_ARGS__) : FN<double, 2>(__VA_ARGS__); \
        case 3: return p->dtype == NFFTB200_F32 ? FN<float, 3>(__VA_ARGS__) : FN<double, 3>(__VA_ARGS__); \
        case 4: return p->dtype == NFFTB200_F32 ? FN<float, 4>(__VA_ARGS__) : FN<double, 4>(__VA_ARGS__); \
        case 5: return p->dtype == NFFTB200_F32 ? FN<float, 5>(__VA_ARGS__) : FN<double, 5>(__VA_ARGS__); \
        case 6: return p->dtype == NFFTB200_F32 ? FN<float, 6>(__VA_ARGS__) : FN<double, 6>(__VA_ARGS__); \
        default: break;                                                                            \
    }

// cells of one padded tile sub-grid if the scratch + column-gather spreader applies to this plan, else 0
int nfftb_peer_tile_cells(nfftb200_plan* p)
{
    if (p->D != 3 || (p->precompute == NFFTB200_FULL && p->window != NFFTB200_KAISER_BESSEL)) return 0;
    PEER_DISPATCH(peer_cells, p)
    return 0;
}

int nfftb_peer_spread(nfftb200_plan* p, const void* d_fhat, void* d_scratch, int64_t t_lo, int64_t t_hi)
{
    PEER_DISPATCH(peer_spread, p, d_fhat, d_scratch, (int)t_lo, (int)t_hi)
    return nfftb_fail(p, NFFTB200_UNSUPPORTED, "peer spread: unsupported m");
}

int nfftb_peer_gather(nfftb200_plan* p, void* d_slab, int layer_lo, int nlayers, const PeerTab& pt)
{
    PEER_DISPATCH(peer_gather, p, d_slab, layer_lo, nlayers, pt)
    return nfftb_fail(p, NFFTB200_UNSUPPORTED, "peer gather: unsupported m");
}

// gather pass over a scratch of padded tiles written by any of the 3-D spreaders (used by lean.cu)
int nfftb_gather_scratch(nfftb200_plan* p, const void* scratch, void* g, int B, int t_lo, int t_hi, int item_lo, int item_hi,
                         const GeomDev* geo_override, const int32_t* items_override)
{
    if (p->dtype != NFFTB200_F32) return -1;
    using T = float;
    using C = float2;
    GeomDev geo = geo_override ? *geo_override : make_geom<T>(p);
    const int32_t* d_tile_items = items_override ? items_override : p->d_tile_items;
    const int units = geo.Nt[0] / 2;
    int bx = 32;
    const cudaStream_t st = p->stream;
#define GS_CASE(MT)                                                                                                              \
    case MT:                                                                                                                     \
        if (geo.bs[2] == 16 && geo.Nt[2] % 16 == 0 && 16 >= 2 * MT) {                                                            \
            while (bx < 128 && bx < units) bx <<= 1;                                                                             \
            dim3 gc((units + bx - 1) / bx, geo.Nt[1], geo.nb[2] * B * 2);                                                        \
            k_gather_cols3d<T, MT, 16, false, 2><<<gc, bx, 0, st>>>((const C*)scratch, (C*)g, d_tile_items, t_lo, t_hi, item_lo, item_hi, geo, PeerTab{}, 0); \
        } else {                                                                                                                 \
            while (bx < 256 && bx < units) bx <<= 1;                                                                             \
            dim3 gg((units + bx - 1) / bx, geo.Nt[1], geo.Nt[2] * B);                                                            \
            k_gather_tiles3d<T, MT><<<gg, bx, 0, st>>>((const C*)scratch, (C*)g, d_tile_items, t_lo, t_hi, item_lo, item_hi, geo); \
        }                                                                                                                        \
        break;
    switch (p->m) {
        GS_CASE(2) GS_CASE(3)
        default: return -1;
    }
#undef GS_CASE
    p->launches++;
    CUDA_TRY(p, cudaGetLastError());
    return NFFTB200_OK;
}

int nfftb_spread(nfftb200_plan* p, const void* d_fhat, void* d_g, int B, int is_complex, int64_t t_lo,
                 int64_t t_hi)
{
    const long long i_lo = p->h_tile_start[t_lo], i_hi = p->h_tile_start[t_hi];
    return p->dtype == NFFTB200_F32
               ? spread_impl<float>(p, d_fhat, d_g, B, is_complex, (int)t_lo, (int)t_hi, i_lo, i_hi)
               : spread_impl<double>(p, d_fhat, d_g, B, is_complex, (int)t_lo, (int)t_hi, i_lo, i_hi);
}
