// interp.cu -- K4: forward interpolation  fHat[j] = sum_l prod_d w_d[l_d] * g[(off+l) mod Nt]
//   replaces convolve! -> _convolve_blocking! -> toBlock!/calcOneBlock!/calcOneNode!
//   (/root/reference/src/convolution.jl:20-45, :229-344).
//
//  * k_interp_generic: warp per node straight from global memory; any D<=3, m<=8, real/complex.
//  * k_interp_tile3d: one CTA per reference tile; the padded sub-grid (bs+2m)^3 is staged in shared
//    memory with coalesced row copies (periodic wrap resolved per row, as toBlock! does), window
//    weights are computed once per (node,dim,tap) into shared-memory records, then one warp per node
//    gathers: lanes own the (y,x) taps of a plane and loop over the 2m z-planes (bank-conflict-free
//    because a plane footprint spans 2m consecutive words per row), warp-shuffle reduction at the end.
#include <cuda.h>

#include <cstring>

#include "common.cuh"
#include "window.cuh"
#include "tile3d.cuh"
#include "interp_bin.cuh"

namespace {

__device__ __forceinline__ int wrap(int v, int n)
{
    v %= n;
    return v < 0 ? v + n : v;
}

template <typename T> __device__ __forceinline__ T warp_sum(T v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

template <typename T, bool CPLX>
__global__ void __launch_bounds__(256)
k_interp_generic(const void* __restrict__ g_, void* __restrict__ fhat_, const T* __restrict__ xs,
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
        for (int b = 0; b < B; b++) {
            T ax = 0, ay = 0;
            for (int q = lane; q < ntaps; q += 32) {
                int l0 = q % L, r = q / L;
                int l1 = r % L, l2 = (r / L) % L, l3 = r / (L * L);
                T w = s_w[warp][0][l0];
                long long cell = wrap(s_c[warp][0] + l0, geo.Nt[0]);
                if (D > 1) { w *= s_w[warp][1][l1]; cell += (long long)wrap(s_c[warp][1] + l1, geo.Nt[1]) * geo.Nt[0]; }
                if (D > 2) { w *= s_w[warp][2][l2]; cell += (long long)wrap(s_c[warp][2] + l2, geo.Nt[2]) * geo.Nt[0] * geo.Nt[1]; }
                if (D > 3) { w *= s_w[warp][3][l3]; cell += (long long)wrap(s_c[warp][3] + l3, geo.Nt[3]) * geo.Nt[0] * geo.Nt[1] * geo.Nt[2]; }
                if (CPLX) {
                    const C v = ((const C*)g_)[b * geo.gsz + cell];
                    ax = tfma(w, v.x, ax); ay = tfma(w, v.y, ay);
                } else {
                    ax = tfma(w, ((const T*)g_)[b * geo.gsz + cell], ax);
                }
            }
            ax = warp_sum<T>(ax);
            if (CPLX) ay = warp_sum<T>(ay);
            if (lane == 0) {
                if (CPLX) ((C*)fhat_)[b * M + j] = make_c<T>(ax, ay);
                else ((T*)fhat_)[b * M + j] = ax;
            }
        }
    }
}

constexpr int TI_WARPS = 8;
constexpr int TI_THREADS = TI_WARPS * 32;

template <typename T, int MT> struct InterpLayout {
    using RG = RowGeom<T, MT>;
    static constexpr int L = 2 * MT;
    static constexpr int RW = ((RG::NWX + 2 * L) + 3) & ~3;         // record: wx(shifted) | wy | wz
    int PX, PY, PZ, PN;
    __host__ __device__ InterpLayout(const int* bs)
    {
        PX = bs[0] + L; PY = bs[1] + L; PZ = bs[2] + L;
        PN = (PX * PY * PZ + 2 * RG::VPC + 1) & ~1;
    }
    __host__ __device__ size_t bytes() const
    {
        return sizeof(typename Cplx<T>::type) * (size_t)(PN + TI_WARPS * 32) + sizeof(T) * TI_WARPS * 32 * RW +
               sizeof(int) * TI_WARPS * 64 + 16;
    }
};

// PEER = true is the multi-GPU form (node sharding, comm.cu): the grid is never assembled; every z plane of a tile
// is staged with cp.async straight from the slab of the rank that holds it (own memory or a CUDA-IPC mapping read
// over NVLink), which replaces the all-gather of the grid by the halo planes the own tiles actually touch.
template <typename T, int MT, bool PEER>
__global__ void __launch_bounds__(TI_THREADS)
k_interp_row3d(const typename Cplx<T>::type* __restrict__ g, typename Cplx<T>::type* __restrict__ fhat,
               const T* __restrict__ xs, const int32_t* __restrict__ perm,
               const int32_t* __restrict__ tile_start, int tile_lo, long long M, GeomDev geo,
               WinDev<T> win, const __grid_constant__ PolyParam<T, MT> pp, const __grid_constant__ CUtensorMap tmap, int use_tma,
               const __grid_constant__ SlabTab slabs)
{
    using C = typename Cplx<T>::type;
    using RG = RowGeom<T, MT>;
    using IL = InterpLayout<T, MT>;
    constexpr int L = 2 * MT, VPC = RG::VPC, NV = RG::NV, NWX = RG::NWX, RW = IL::RW;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const IL lay(geo.bs);
    const int PX = lay.PX, PY = lay.PY, PZ = lay.PZ;
    C* tile = reinterpret_cast<C*>(smem_raw);                                   // [PN]
    C* res = tile + lay.PN;                                                     // [8][32]
    T* rec_w = reinterpret_cast<T*>(res + TI_WARPS * 32);                       // [8][32][RW]
    int* rec_i = reinterpret_cast<int*>(rec_w + TI_WARPS * 32 * RW);            // [8][32][2]
    unsigned long long* mbar = reinterpret_cast<unsigned long long*>(rec_i + TI_WARPS * 64);

    const int32_t* item = tile_start + 3 * (size_t)(tile_lo + blockIdx.x);     // work item (tile, node range)
    const int tile_id = item[0];
    const int n_lo = item[1], n_hi = item[2];
    const int tx = tile_id % geo.nb[0];
    const int ty = (tile_id / geo.nb[0]) % geo.nb[1];
    const int tz = tile_id / (geo.nb[0] * geo.nb[1]);
    const int x0 = tx * geo.bs[0] - MT, y0 = ty * geo.bs[1] - MT, z0 = tz * geo.bs[2] - MT;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (!PEER) g += (long long)blockIdx.y * geo.gsz;
    fhat += (long long)blockIdx.y * M;

    // toBlock!: interior tiles (no periodic wrap) are staged by ONE TMA tensor-map load of the (PX,PY,PZ) box
    // (cp.async.bulk.tensor + mbarrier); tiles that wrap are staged row by row with cp.async.
    // (measured on B200: a box whose innermost start coordinate is not 16-byte aligned raises "illegal
    //  instruction", so Float32 tiles qualify only when their origin x0 = tx*bs - m is even, i.e. m even)
    const bool interior = !PEER && use_tma && x0 >= 0 && y0 >= 0 && z0 >= 0 && x0 + PX <= geo.Nt[0] && y0 + PY <= geo.Nt[1] &&
                          z0 + PZ <= geo.Nt[2] && ((x0 * (int)sizeof(C)) & 15) == 0;
    if (interior) {
        if (threadIdx.x == 0) {
            mbar_init(mbar, 1);
            mbar_expect_tx(mbar, (unsigned)(sizeof(C) * PX * PY * PZ));
            tma_load_4d(tile, &tmap, mbar, 2 * x0, y0, z0, (int)blockIdx.y);
        }
    } else {
        const bool fw = PX <= geo.Nt[0] && PY <= geo.Nt[1] && PZ <= geo.Nt[2];
        const int xg0 = wrapc(x0 + lane, geo.Nt[0], fw), xg1 = wrapc(x0 + lane + 32, geo.Nt[0], fw);
        const bool on0 = lane < PX, on1 = lane + 32 < PX;
        for (int z = 0; z < PZ; z++) {
            unsigned gz = (unsigned)wrapc(z0 + z, geo.Nt[2], fw);
            const C* gb = g;
            if (PEER) {
                const unsigned owner = gz / (unsigned)slabs.planes;
                gb = (const C*)slabs.base[owner];
                gz -= owner * (unsigned)slabs.planes;
            }
            gz *= geo.Nt[1];
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const int y = warp + TI_WARPS * k;
                if (y < PY) {
                    const C* src = gb + (gz + wrapc(y0 + y, geo.Nt[1], fw)) * (unsigned)geo.Nt[0];
                    C* dst = tile + (z * PY + y) * PX + lane;
                    if (on0) cp_async_cell(dst, src + xg0);
                    if (on1) cp_async_cell(dst + 32, src + xg1);
                }
            }
        }
    }
    if (threadIdx.x < lay.PN - PX * PY * PZ) tile[PX * PY * PZ + threadIdx.x] = make_c<T>(0, 0);

    int rowoff[RG::FULL_IT > 0 ? RG::FULL_IT : 1], wyo[RG::FULL_IT > 0 ? RG::FULL_IT : 1],
        wzo[RG::FULL_IT > 0 ? RG::FULL_IT : 1];
#pragma unroll
    for (int it = 0; it < RG::FULL_IT; it++) {
        const int r = lane + 32 * it, t = r / L, yt = r - t * L;
        rowoff[it] = (t * PY + yt) * PX; wyo[it] = NWX + yt; wzo[it] = NWX + L + t;
    }
    int rem_off = 0, rem_wy = 0, rem_wz = 0, rem_u = 0;
    bool rem_on = false;
    if (RG::REM > 0) {
        const int rr = RG::SPLIT ? lane / NV : lane;
        rem_u = RG::SPLIT ? lane - rr * NV : 0;
        rem_on = rr < RG::REM;
        const int r = RG::FULL_IT * 32 + (rem_on ? rr : 0), t = r / L, yt = r - t * L;
        rem_off = (t * PY + yt) * PX + rem_u * VPC; rem_wy = NWX + yt; rem_wz = NWX + L + t;
    }
    T* myrec = rec_w + warp * 32 * RW;
    int* myint = rec_i + warp * 64;
    C* myres = res + warp * 32;
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncthreads();
    if (interior) mbar_wait(mbar, 0);

    for (int rbase = n_lo + 32 * warp; rbase < n_hi; rbase += 32 * TI_WARPS) {
        const int nn = min(32, n_hi - rbase);
        // ---- phase A: lane-per-node weights
        if (lane < nn) {
            const long long i = (long long)rbase + lane;
            T ks0, ks1, ks2;
            const int c0 = node_cell<T>(xs[i * 3 + 0], geo.Nt[0], ks0);
            const int c1 = node_cell<T>(xs[i * 3 + 1], geo.Nt[1], ks1);
            const int c2 = node_cell<T>(xs[i * 3 + 2], geo.Nt[2], ks2);
            T w0[L], w1[L], w2[L];
            eval_taps<T, MT>(win, pp, ks0, c0, w0);
            eval_taps<T, MT>(win, pp, ks1, c1, w1);
            eval_taps<T, MT>(win, pp, ks2, c2, w2);
            const int px = c0 - MT + 1 - x0, py = c1 - MT + 1 - y0, pz = c2 - MT + 1 - z0;
            const int s = (VPC == 2) ? (px & 1) : 0;
            myint[2 * lane] = (pz * PY + py) * PX + (px - s);
            myint[2 * lane + 1] = perm[i];
            T* dst = myrec + lane * RW;
#pragma unroll
            for (int k = 0; k < NWX; k++) {
                const T lo = (k < L) ? w0[k < L ? k : 0] : (T)0;
                const T hi = (k >= 1 && k - 1 < L) ? w0[(k >= 1 && k - 1 < L) ? k - 1 : 0] : (T)0;
                dst[k] = s ? hi : lo;
            }
#pragma unroll
            for (int k = 0; k < L; k++) { dst[NWX + k] = w1[k]; dst[NWX + L + k] = w2[k]; }
        }
        __syncwarp();
        // ---- phase B: gather, one node at a time, every lane owns one x-row of the footprint; the record
        //      of node n+1 is fetched while node n is reduced (software pipelining)
        constexpr int NIT = RG::FULL_IT > 0 ? RG::FULL_IT : 1;
        T wx[NWX], wq[NIT], rq = 0, rwu[VPC];
        int base;
        auto fetch = [&](int n, T (&wx_)[NWX], T (&wq_)[NIT], T& rq_, T (&wu_)[VPC], int& base_) {
            const T* rw = myrec + n * RW;
            base_ = myint[2 * n];
#pragma unroll
            for (int k = 0; k < NWX; k++) wx_[k] = rw[k];
#pragma unroll
            for (int it = 0; it < RG::FULL_IT; it++) wq_[it] = rw[wyo[it]] * rw[wzo[it]];
            if (RG::REM > 0) {
                rq_ = rw[rem_wy] * rw[rem_wz];
                if (RG::SPLIT) {
#pragma unroll
                    for (int k = 0; k < VPC; k++) wu_[k] = rw[rem_u * VPC + k];
                }
            }
        };
        fetch(0, wx, wq, rq, rwu, base);
        // nodes are reduced in groups of 4: after two exchange steps each quarter-warp holds one node, so the
        // 5-step butterfly costs 12 shuffles per 4 nodes instead of 40
        for (int n4 = 0; n4 < nn; n4 += 4) {
            T px[4], py[4];
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const int n = n4 + q;
                T sx = 0, sy = 0;
                if (n < nn) {
                    const C* p0 = tile + base;
                    T wx2[NWX], wq2[NIT], rq2 = 0, rwu2[VPC];
                    int base2 = base;
                    if (RG::FULL_IT == 0) fetch(n + 1 < nn ? n + 1 : n, wx2, wq2, rq2, rwu2, base2);
#pragma unroll
                    for (int it = 0; it < RG::FULL_IT; it++) {
                        const C* p = p0 + rowoff[it];
                        Unit<T> U[NV];
#pragma unroll
                        for (int u = 0; u < NV; u++) U[u].load(p + u * VPC);
                        if (it == 0) fetch(n + 1 < nn ? n + 1 : n, wx2, wq2, rq2, rwu2, base2);
                        T a = 0, b = 0;
#pragma unroll
                        for (int u = 0; u < NV; u++) U[u].dot(&wx[u * VPC], a, b);
                        sx = tfma(wq[it], a, sx); sy = tfma(wq[it], b, sy);
                    }
                    if (RG::REM > 0 && rem_on) {
                        const C* p = p0 + rem_off;
                        T a = 0, b = 0;
                        if (RG::SPLIT) { Unit<T> U; U.load(p); U.dot(rwu, a, b); }
                        else {
#pragma unroll
                            for (int u = 0; u < NV; u++) { Unit<T> U; U.load(p + u * VPC); U.dot(&wx[u * VPC], a, b); }
                        }
                        sx = tfma(rq, a, sx); sy = tfma(rq, b, sy);
                    }
#pragma unroll
                    for (int k = 0; k < NWX; k++) wx[k] = wx2[k];
#pragma unroll
                    for (int it = 0; it < NIT; it++) wq[it] = wq2[it];
                    rq = rq2; base = base2;
#pragma unroll
                    for (int k = 0; k < VPC; k++) rwu[k] = rwu2[k];
                }
                px[q] = sx; py[q] = sy;
            }
            // exchange step 1 (xor 16): lower half keeps nodes {0,2}, upper half nodes {1,3}
            const bool hi16 = lane & 16, hi8 = lane & 8;
            T ax = hi16 ? px[1] : px[0], bx = hi16 ? px[0] : px[1];
            T ay = hi16 ? py[1] : py[0], by = hi16 ? py[0] : py[1];
            T cx = hi16 ? px[3] : px[2], dx = hi16 ? px[2] : px[3];
            T cy = hi16 ? py[3] : py[2], dy = hi16 ? py[2] : py[3];
            ax += __shfl_xor_sync(0xffffffffu, bx, 16); ay += __shfl_xor_sync(0xffffffffu, by, 16);
            cx += __shfl_xor_sync(0xffffffffu, dx, 16); cy += __shfl_xor_sync(0xffffffffu, dy, 16);
            // exchange step 2 (xor 8): quarter q = (lane>>3) holds node (hi16 ? 1 : 0) + 2*(hi8 ? 1 : 0)
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

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// tensor map of the (batched) grid seen as floats/doubles: dims (2*Nt0, Nt1, Nt2, B), box (2*PX, PY, PZ, 1)
template <typename T> bool make_grid_tensor_map(CUtensorMap* tm, const void* g, const GeomDev& geo, int B, int PX, int PY, int PZ)
{
    static PFN_encodeTiled fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_encodeTiled>(ptr);
        else
            cudaGetLastError();
    }
    if (!fn) return false;
    if (2 * PX > 256 || PY > 256 || PZ > 256 || ((uintptr_t)g & 15)) return false;
    const cuuint64_t csz = 2 * sizeof(T);
    cuuint64_t dims[4] = {(cuuint64_t)2 * geo.Nt[0], (cuuint64_t)geo.Nt[1], (cuuint64_t)geo.Nt[2], (cuuint64_t)B};
    cuuint64_t strides[3] = {(cuuint64_t)geo.Nt[0] * csz, (cuuint64_t)geo.Nt[0] * geo.Nt[1] * csz, (cuuint64_t)geo.gsz * csz};
    cuuint32_t box[4] = {(cuuint32_t)(2 * PX), (cuuint32_t)PY, (cuuint32_t)PZ, 1};
    cuuint32_t es[4] = {1, 1, 1, 1};
    if ((strides[0] & 15) || ((box[0] * sizeof(T)) & 15)) return false;
    const CUresult r = fn(tm, sizeof(T) == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 4,
                          const_cast<void*>(g), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

template <typename T, int MT>
int peer_interp(nfftb200_plan* p, const SlabTab& st, void* fhat, int t_lo, int t_hi)
{
    using C = typename Cplx<T>::type;
    GeomDev geo = make_geom<T>(p);
    InterpLayout<T, MT> lay(geo.bs);
    const size_t smem = lay.bytes();
    if (smem > 227 * 1024 || geo.bs[0] + 2 * MT > 64 || geo.bs[1] + 2 * MT > 32) return -1;
    const int item_lo = p->h_tile_items[(size_t)t_lo], item_hi = p->h_tile_items[(size_t)t_hi];
    if (item_hi == item_lo) return NFFTB200_OK;
    if (sizeof(T) == 4 && nfftb_lean_mode(p)) {   // (tile, bin)-ordered register windows, slab-direct form
        if (p->timing) cudaEventRecord(p->evk[3], p->stream);
        const int r = nfftb_interp_lean(p, nullptr, fhat, 1, t_lo, t_hi, &st);
        if (r >= 0) {
            if (p->timing) { cudaEventRecord(p->evk[4], p->stream); p->pending_k |= 2; }
            return r;
        }
    }
    if constexpr (MT <= 3) {          // opt-in register-window interpolator, slab-direct form
        BinGeom bg;
        using IL = InterpBinLayout<T, MT, 8>;
        if (p->kernel_mode == 7 && IL::make(geo.bs, bg) && IL::bytes(bg) <= 227 * 1024) {
            auto kb = k_interp_bin3d<T, MT, 8, true>;
            CUDA_TRY(p, cudaFuncSetAttribute(kb, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)IL::bytes(bg)));
            cudaFuncSetAttribute(kb, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
            if (p->timing) cudaEventRecord(p->evk[3], p->stream);
            kb<<<dim3(item_hi - item_lo, 1), NFFTB_BIN_WARPS * 32, IL::bytes(bg), p->stream>>>(nullptr, (C*)fhat, (const T*)p->d_xs, p->d_perm,
                                                                                             p->d_items, item_lo, p->M, geo, make_win<T>(p),
                                                                                             make_poly_param<T, MT>(p), bg, st);
            if (p->timing) { cudaEventRecord(p->evk[4], p->stream); p->pending_k |= 2; }
            p->launches++;
            CUDA_TRY(p, cudaGetLastError());
            return NFFTB200_OK;
        }
    }
    auto kern = k_interp_row3d<T, MT, true>;
    CUDA_TRY(p, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CUtensorMap tmap;
    std::memset(&tmap, 0, sizeof(tmap));
    if (p->timing) cudaEventRecord(p->evk[3], p->stream);
    kern<<<dim3(item_hi - item_lo, 1), TI_THREADS, smem, p->stream>>>(nullptr, (C*)fhat, (const T*)p->d_xs, p->d_perm,
                                                                     p->d_items, item_lo, p->M, geo, make_win<T>(p),
                                                                     make_poly_param<T, MT>(p), tmap, 0, st);
    if (p->timing) { cudaEventRecord(p->evk[4], p->stream); p->pending_k |= 2; }
    p->launches++;
    CUDA_TRY(p, cudaGetLastError());
    return NFFTB200_OK;
}

template <typename T, int MT>
int launch_tile3d(nfftb200_plan* p, const void* g, void* fhat, int B, int t_lo, int t_hi)
{
    using C = typename Cplx<T>::type;
    GeomDev geo = make_geom<T>(p);
    InterpLayout<T, MT> lay(geo.bs);
    const size_t smem = lay.bytes();
    if (smem > 227 * 1024 || geo.bs[0] + 2 * MT > 64 || geo.bs[1] + 2 * MT > 32) return -1;
    auto kern = k_interp_row3d<T, MT, false>;
    CUDA_TRY(p, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int item_lo = p->h_tile_items[(size_t)t_lo], item_hi = p->h_tile_items[(size_t)t_hi];
    if (item_hi == item_lo) return NFFTB200_OK;
    dim3 grid(item_hi - item_lo, B);
    CUtensorMap tmap;
    std::memset(&tmap, 0, sizeof(tmap));
    const int use_tma = (p->kernel_mode != 3 && make_grid_tensor_map<T>(&tmap, g, geo, B, lay.PX, lay.PY, lay.PZ)) ? 1 : 0;
    kern<<<grid, TI_THREADS, smem, p->stream>>>((const C*)g, (C*)fhat, (const T*)p->d_xs, p->d_perm,
                                               p->d_items, item_lo, p->M, geo, make_win<T>(p),
                                               make_poly_param<T, MT>(p), tmap, use_tma, SlabTab{});
    p->launches++;
    CUDA_TRY(p, cudaGetLastError());
    return NFFTB200_OK;
}

// kernel_mode 7: register-window interpolator (interp_bin.cuh); -1 when it does not apply
template <typename T, int MT, int W>
int launch_bin3d(nfftb200_plan* p, const void* g, void* fhat, int B, int t_lo, int t_hi)
{
    using C = typename Cplx<T>::type;
    using IL = InterpBinLayout<T, MT, W>;
    GeomDev geo = make_geom<T>(p);
    BinGeom bg;
    if (!IL::make(geo.bs, bg)) return -1;
    const size_t smem = IL::bytes(bg);
    if (smem > 227 * 1024 || geo.bs[0] + 2 * MT > 64) return -1;
    auto kern = k_interp_bin3d<T, MT, W>;
    CUDA_TRY(p, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    // two CTAs per SM need the largest shared-memory carve-out (a hint; the kernel is correct without it)
    cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    const int item_lo = p->h_tile_items[(size_t)t_lo], item_hi = p->h_tile_items[(size_t)t_hi];
    if (item_hi == item_lo) return NFFTB200_OK;
    kern<<<dim3(item_hi - item_lo, B), NFFTB_BIN_WARPS * 32, smem, p->stream>>>((const C*)g, (C*)fhat, (const T*)p->d_xs, p->d_perm,
                                                                               p->d_items, item_lo, p->M, geo, make_win<T>(p),
                                                                               make_poly_param<T, MT>(p), bg, SlabTab{});
    p->launches++;
    CUDA_TRY(p, cudaGetLastError());
    return NFFTB200_OK;
}


template <typename T>
int interp_impl(nfftb200_plan* p, const void* g, void* fhat, int B, int is_complex, int t_lo, int t_hi,
                long long i_lo, long long i_hi)
{
    if (i_hi <= i_lo) return NFFTB200_OK;
    struct KernelTimer {
        nfftb200_plan* p;
        explicit KernelTimer(nfftb200_plan* q) : p(q) { if (p->timing) cudaEventRecord(p->evk[3], p->stream); }
        ~KernelTimer() { if (p->timing) { cudaEventRecord(p->evk[4], p->stream); p->pending_k |= 2; } }
    } kt(p);
    if (nfftb_tiled_ok(p) && p->D == 1) {
        const int r = nfftb_interp_1d(p, g, fhat, B, is_complex, i_lo, i_hi);
        if (r >= 0) return r;
    }
    if (nfftb_tiled_ok(p) && is_complex && p->D == 2) {
        const int r = nfftb_interp_2d(p, g, fhat, B, t_lo, t_hi);
        if (r >= 0) return r;
    }
    if (sizeof(T) == 4 && nfftb_tiled_ok(p) && is_complex && p->D == 3 && nfftb_lean_mode(p)) {
        const int r = nfftb_interp_lean(p, g, fhat, B, t_lo, t_hi, nullptr);      // lean.cu: (tile, bin)-ordered register windows
        if (r >= 0) return r;
    }
    if (nfftb_tiled_ok(p) && is_complex && p->D == 3 && p->kernel_mode == 7) {
        int r = -1;                                    // opt-in register-window interpolator (interp_bin.cuh)
        switch (p->m) {
            case 2: r = launch_bin3d<T, 2, 8>(p, g, fhat, B, t_lo, t_hi); break;
            case 3: r = launch_bin3d<T, 3, 8>(p, g, fhat, B, t_lo, t_hi); break;
            case 4: if constexpr (sizeof(T) == 4) r = launch_bin3d<T, 4, 10>(p, g, fhat, B, t_lo, t_hi); break;   // Float64: tile + records exceed 227 KB
            default: break;
        }
        if (r >= 0) return r;
    }
    if (nfftb_tiled_ok(p) && is_complex && p->D == 3) {
        int r = -1;
        switch (p->m) {
            case 2: r = launch_tile3d<T, 2>(p, g, fhat, B, t_lo, t_hi); break;
            case 3: r = launch_tile3d<T, 3>(p, g, fhat, B, t_lo, t_hi); break;
            case 4: r = launch_tile3d<T, 4>(p, g, fhat, B, t_lo, t_hi); break;
            case 5: r = launch_tile3d<T, 5>(p, g, fhat, B, t_lo, t_hi); break;
            case 6: r = launch_tile3d<T, 6>(p, g, fhat, B, t_lo, t_hi); break;
            default: break;
        }
        if (r >= 0) return r;
    }
    const long long n = i_hi - i_lo;
    const int blocks = (int)std::min<long long>((n + 7) / 8, 148 * 32);
    if (is_complex)
        k_interp_generic<T, true><<<blocks, 256, 0, p->stream>>>(g, fhat, (const T*)p->d_xs, p->d_perm, i_lo,
                                                                 i_hi, p->M, make_geom<T>(p), make_win<T>(p), B);
    else
        k_interp_generic<T, false><<<blocks, 256, 0, p->stream>>>(g, fhat, (const T*)p->d_xs, p->d_perm, i_lo,
                                                                  i_hi, p->M, make_geom<T>(p), make_win<T>(p), B);
    p->launches++;
    CUDA_TRY(p, cudaGetLastError());
    return NFFTB200_OK;
}

}  // namespace

// interpolate the own tile range reading the grid from the ranks' z-slabs; -1 if the tiled kernel does not apply
int nfftb_peer_interp(nfftb200_plan* p, const SlabTab& st, void* d_fhat, int64_t t_lo, int64_t t_hi)
{
    if (p->D != 3 || (p->precompute == NFFTB200_FULL && p->window != NFFTB200_KAISER_BESSEL)) return -1;
#define PI_CASE(MM) case MM: return p->dtype == NFFTB200_F32 ? peer_interp<float, MM>(p, st, d_fhat, (int)t_lo, (int)t_hi) \
                                                               : peer_interp<double, MM>(p, st, d_fhat, (int)t_lo, (int)t_hi);
    switch (p->m) {
        PI_CASE(2) PI_CASE(3) PI_CASE(4) PI_CASE(5) PI_CASE(6)
        default: break;
    }
#undef PI_CASE
    return -1;
}

int nfftb_interp(nfftb200_plan* p, const void* d_g, void* d_fhat, int B, int is_complex, int64_t t_lo,
                 int64_t t_hi)
{
    const long long i_lo = p->h_tile_start[t_lo], i_hi = p->h_tile_start[t_hi];
    return p->dtype == NFFTB200_F32
               ? interp_impl<float>(p, d_g, d_fhat, B, is_complex, (int)t_lo, (int)t_hi, i_lo, i_hi)
               : interp_impl<double>(p, d_g, d_fhat, B, is_complex, (int)t_lo, (int)t_hi, i_lo, i_hi);
}
