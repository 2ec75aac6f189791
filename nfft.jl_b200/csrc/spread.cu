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
#include "common.cuh"
#include "window.cuh"

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
    const int ntaps = (D == 1) ? L : (D == 2 ? L * L : L * L * L);
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
            int l1 = r % L, l2 = r / L;
            T w = s_w[warp][0][l0];
            long long cell = wrap(s_c[warp][0] + l0, geo.Nt[0]);
            if (D > 1) { w *= s_w[warp][1][l1]; cell += (long long)wrap(s_c[warp][1] + l1, geo.Nt[1]) * geo.Nt[0]; }
            if (D > 2) { w *= s_w[warp][2][l2]; cell += (long long)wrap(s_c[warp][2] + l2, geo.Nt[2]) * geo.Nt[0] * geo.Nt[1]; }
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
// tiled 3-D spreader
// ---------------------------------------------------------------------------------------
constexpr int TS_WARPS = 8;
constexpr int TS_THREADS = TS_WARPS * 32;
constexpr int TS_CHUNK = 128;        // nodes per record chunk

template <typename T, int MT> struct TileSmem {
    static constexpr int L = 2 * MT;
    // byte offsets inside dynamic shared memory
    static size_t tile_bytes(int PX, int PY, int PZ) { return sizeof(typename Cplx<T>::type) * (size_t)PX * PY * PZ; }
    static size_t rec_bytes()
    {
        return 2 * (sizeof(T) * TS_CHUNK * 3 * L + sizeof(typename Cplx<T>::type) * TS_CHUNK +
                    2 * sizeof(int) * TS_CHUNK);
    }
};

template <typename T, int MT>
__global__ void __launch_bounds__(TS_THREADS)
k_spread_tile3d(const typename Cplx<T>::type* __restrict__ fhat, typename Cplx<T>::type* __restrict__ g,
                const T* __restrict__ xs, const int32_t* __restrict__ perm,
                const int32_t* __restrict__ tile_start, int tile_lo, long long M, GeomDev geo,
                WinDev<T> win)
{
    using C = typename Cplx<T>::type;
    constexpr int L = 2 * MT;
    constexpr int NIT = (L * L + 31) / 32;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int PX = geo.bs[0] + L, PY = geo.bs[1] + L, PZ = geo.bs[2] + L;
    const int ncell = PX * PY * PZ;
    C* tile = reinterpret_cast<C*>(smem_raw);
    T* s_w = reinterpret_cast<T*>(tile + ncell);                       // [2][CHUNK][3L]
    C* s_v = reinterpret_cast<C*>(s_w + 2 * TS_CHUNK * 3 * L);         // [2][CHUNK]
    int* s_base = reinterpret_cast<int*>(s_v + 2 * TS_CHUNK);          // [2][CHUNK]
    int* s_oz = s_base + 2 * TS_CHUNK;                                 // [2][CHUNK]

    const int tile_id = tile_lo + blockIdx.x;
    const int b = blockIdx.y;
    const int n_lo = tile_start[tile_id], n_hi = tile_start[tile_id + 1];
    if (n_hi == n_lo) return;                                          // grid is pre-zeroed
    const int tx = tile_id % geo.nb[0];
    const int ty = (tile_id / geo.nb[0]) % geo.nb[1];
    const int tz = tile_id / (geo.nb[0] * geo.nb[1]);
    const int x0 = tx * geo.bs[0] - MT, y0 = ty * geo.bs[1] - MT, z0 = tz * geo.bs[2] - MT;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    fhat += (long long)b * M;
    g += (long long)b * geo.gsz;

    for (int q = threadIdx.x; q < ncell; q += TS_THREADS) tile[q] = make_c<T>(0, 0);

    // per-lane constant footprint offsets
    int coff[NIT], xo[NIT], yo[NIT];
#pragma unroll
    for (int it = 0; it < NIT; it++) {
        const int q = lane + 32 * it;
        const int yt = q / L, xt = q - yt * L;
        yo[it] = yt; xo[it] = xt; coff[it] = yt * PX + xt;
    }

    auto phase_a = [&](int buf, int c_lo, int nc) {
        T* w = s_w + buf * TS_CHUNK * 3 * L;
        for (int q = threadIdx.x; q < nc * 3 * L; q += TS_THREADS) {
            const int n = q / (3 * L), r = q - n * (3 * L);
            const int d = r / L, l = r - d * L;
            T ks;
            const int c = node_cell<T>(xs[(long long)(c_lo + n) * 3 + d], geo.Nt[d], ks);
            w[q] = node_tap<T>(win, ks, c, l);
        }
        for (int n = threadIdx.x; n < nc; n += TS_THREADS) {
            const long long i = c_lo + n;
            T ks;
            const int cx = node_cell<T>(xs[i * 3 + 0], geo.Nt[0], ks);
            const int cy = node_cell<T>(xs[i * 3 + 1], geo.Nt[1], ks);
            const int cz = node_cell<T>(xs[i * 3 + 2], geo.Nt[2], ks);
            const int ox = cx - MT + 1 - x0, oy = cy - MT + 1 - y0, oz = cz - MT + 1 - z0;
            s_base[buf * TS_CHUNK + n] = (oz * PY + oy) * PX + ox;
            s_oz[buf * TS_CHUNK + n] = oz;
            s_v[buf * TS_CHUNK + n] = fhat[perm[i]];
        }
    };

    phase_a(0, n_lo, min(TS_CHUNK, n_hi - n_lo));
    __syncthreads();
    int buf = 0;
    for (int c_lo = n_lo; c_lo < n_hi; c_lo += TS_CHUNK, buf ^= 1) {
        const int nc = min(TS_CHUNK, n_hi - c_lo);
        const int nxt = c_lo + TS_CHUNK;
        if (nxt < n_hi) phase_a(buf ^ 1, nxt, min(TS_CHUNK, n_hi - nxt));
        // phase B: plane-owner accumulation
        const T* w = s_w + buf * TS_CHUNK * 3 * L;
        const C* vv = s_v + buf * TS_CHUNK;
        const int* bb = s_base + buf * TS_CHUNK;
        const int* zz = s_oz + buf * TS_CHUNK;
        for (int n = 0; n < nc; n++) {
            const int oz = zz[n];
            for (int t = (warp - oz) & (TS_WARPS - 1); t < L; t += TS_WARPS) {
                const T* wn = w + n * 3 * L;
                const T wz = wn[2 * L + t];
                const C v = vv[n];
                const T vzx = v.x * wz, vzy = v.y * wz;
                C* row = tile + bb[n] + t * PY * PX;
#pragma unroll
                for (int it = 0; it < NIT; it++) {
                    if (lane + 32 * it < L * L) {
                        const T wxy = wn[xo[it]] * wn[L + yo[it]];
                        C cur = row[coff[it]];
                        cur.x = tfma(wxy, vzx, cur.x);
                        cur.y = tfma(wxy, vzy, cur.y);
                        row[coff[it]] = cur;
                    }
                }
            }
        }
        __syncthreads();
    }

    // flush padded tile (halo included) with one vector RED per cell
    for (int q = threadIdx.x; q < ncell; q += TS_THREADS) {
        const int x = q % PX, r = q / PX;
        const int y = r % PY, z = r / PY;
        const C v = tile[q];
        if (v.x == (T)0 && v.y == (T)0) continue;
        const long long gi = ((long long)wrap(z0 + z, geo.Nt[2]) * geo.Nt[1] + wrap(y0 + y, geo.Nt[1])) * geo.Nt[0] +
                             wrap(x0 + x, geo.Nt[0]);
        red_add_c(g + gi, v);
    }
}

template <typename T, int MT>
int launch_tile3d(nfftb200_plan* p, const void* fhat, void* g, int B, int t_lo, int t_hi)
{
    using C = typename Cplx<T>::type;
    const int L = 2 * MT;
    const int PX = (int)p->bs[0] + L, PY = (int)p->bs[1] + L, PZ = (int)p->bs[2] + L;
    const size_t smem = TileSmem<T, MT>::tile_bytes(PX, PY, PZ) + TileSmem<T, MT>::rec_bytes();
    if (smem > 227 * 1024) return -1;
    auto kern = k_spread_tile3d<T, MT>;
    CUDA_TRY(p, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid(t_hi - t_lo, B);
    kern<<<grid, TS_THREADS, smem, p->stream>>>((const C*)fhat, (C*)g, (const T*)p->d_xs, p->d_perm,
                                               p->d_tile_start, t_lo, p->M, make_geom<T>(p), make_win<T>(p));
    p->launches++;
    CUDA_TRY(p, cudaGetLastError());
    return NFFTB200_OK;
}

template <typename T>
int spread_impl(nfftb200_plan* p, const void* fhat, void* g, int B, int is_complex, int t_lo, int t_hi,
                long long i_lo, long long i_hi)
{
    // zero the grid (memset(g), /root/reference/src/convolution.jl:358)
    const size_t cell = is_complex ? 2 * sizeof(T) : sizeof(T);
    if (p->timing) cudaEventRecord(p->evk[0], p->stream);
    CUDA_TRY(p, cudaMemsetAsync(g, 0, cell * (size_t)p->gsz * B, p->stream));
    p->launches++;
    if (i_hi <= i_lo) return NFFTB200_OK;
    struct KernelTimer {
        nfftb200_plan* p;
        explicit KernelTimer(nfftb200_plan* q) : p(q) { if (p->timing) cudaEventRecord(p->evk[1], p->stream); }
        ~KernelTimer() { if (p->timing) { cudaEventRecord(p->evk[2], p->stream); p->pending_k |= 1; } }
    } kt(p);
    if (p->kernel_mode == 0 && is_complex && p->D == 3) {
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

int nfftb_spread(nfftb200_plan* p, const void* d_fhat, void* d_g, int B, int is_complex, int64_t t_lo,
                 int64_t t_hi)
{
    const long long i_lo = p->h_tile_start[t_lo], i_hi = p->h_tile_start[t_hi];
    return p->dtype == NFFTB200_F32
               ? spread_impl<float>(p, d_fhat, d_g, B, is_complex, (int)t_lo, (int)t_hi, i_lo, i_hi)
               : spread_impl<double>(p, d_fhat, d_g, B, is_complex, (int)t_lo, (int)t_hi, i_lo, i_hi);
}
