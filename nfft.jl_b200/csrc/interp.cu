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
#include "common.cuh"
#include "window.cuh"

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
        for (int b = 0; b < B; b++) {
            T ax = 0, ay = 0;
            for (int q = lane; q < ntaps; q += 32) {
                int l0 = q % L, r = q / L;
                int l1 = r % L, l2 = r / L;
                T w = s_w[warp][0][l0];
                long long cell = wrap(s_c[warp][0] + l0, geo.Nt[0]);
                if (D > 1) { w *= s_w[warp][1][l1]; cell += (long long)wrap(s_c[warp][1] + l1, geo.Nt[1]) * geo.Nt[0]; }
                if (D > 2) { w *= s_w[warp][2][l2]; cell += (long long)wrap(s_c[warp][2] + l2, geo.Nt[2]) * geo.Nt[0] * geo.Nt[1]; }
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
constexpr int TI_CHUNK = 128;

template <typename T, int MT>
__global__ void __launch_bounds__(TI_THREADS)
k_interp_tile3d(const typename Cplx<T>::type* __restrict__ g, typename Cplx<T>::type* __restrict__ fhat,
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
    int* s_base = reinterpret_cast<int*>(s_w + 2 * TI_CHUNK * 3 * L);  // [2][CHUNK]
    int* s_j = s_base + 2 * TI_CHUNK;                                  // [2][CHUNK]

    const int tile_id = tile_lo + blockIdx.x;
    const int b = blockIdx.y;
    const int n_lo = tile_start[tile_id], n_hi = tile_start[tile_id + 1];
    if (n_hi == n_lo) return;
    const int tx = tile_id % geo.nb[0];
    const int ty = (tile_id / geo.nb[0]) % geo.nb[1];
    const int tz = tile_id / (geo.nb[0] * geo.nb[1]);
    const int x0 = tx * geo.bs[0] - MT, y0 = ty * geo.bs[1] - MT, z0 = tz * geo.bs[2] - MT;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    g += (long long)b * geo.gsz;
    fhat += (long long)b * M;

    // toBlock!: stage the padded tile, x fastest (coalesced within each row segment)
    for (int q = threadIdx.x; q < ncell; q += TI_THREADS) {
        const int x = q % PX, r = q / PX;
        const int y = r % PY, z = r / PY;
        const long long gi = ((long long)wrap(z0 + z, geo.Nt[2]) * geo.Nt[1] + wrap(y0 + y, geo.Nt[1])) * geo.Nt[0] +
                             wrap(x0 + x, geo.Nt[0]);
        tile[q] = g[gi];
    }

    int coff[NIT], xo[NIT], yo[NIT];
#pragma unroll
    for (int it = 0; it < NIT; it++) {
        const int q = lane + 32 * it;
        const int yt = q / L, xt = q - yt * L;
        yo[it] = yt; xo[it] = xt; coff[it] = yt * PX + xt;
    }

    auto phase_a = [&](int buf, int c_lo, int nc) {
        T* w = s_w + buf * TI_CHUNK * 3 * L;
        for (int q = threadIdx.x; q < nc * 3 * L; q += TI_THREADS) {
            const int n = q / (3 * L), r = q - n * (3 * L);
            const int d = r / L, l = r - d * L;
            T ks;
            const int c = node_cell<T>(xs[(long long)(c_lo + n) * 3 + d], geo.Nt[d], ks);
            w[q] = node_tap<T>(win, ks, c, l);
        }
        for (int n = threadIdx.x; n < nc; n += TI_THREADS) {
            const long long i = c_lo + n;
            T ks;
            const int cx = node_cell<T>(xs[i * 3 + 0], geo.Nt[0], ks);
            const int cy = node_cell<T>(xs[i * 3 + 1], geo.Nt[1], ks);
            const int cz = node_cell<T>(xs[i * 3 + 2], geo.Nt[2], ks);
            const int ox = cx - MT + 1 - x0, oy = cy - MT + 1 - y0, oz = cz - MT + 1 - z0;
            s_base[buf * TI_CHUNK + n] = (oz * PY + oy) * PX + ox;
            s_j[buf * TI_CHUNK + n] = perm[i];
        }
    };

    phase_a(0, n_lo, min(TI_CHUNK, n_hi - n_lo));
    __syncthreads();
    int buf = 0;
    for (int c_lo = n_lo; c_lo < n_hi; c_lo += TI_CHUNK, buf ^= 1) {
        const int nc = min(TI_CHUNK, n_hi - c_lo);
        const int nxt = c_lo + TI_CHUNK;
        if (nxt < n_hi) phase_a(buf ^ 1, nxt, min(TI_CHUNK, n_hi - nxt));
        const T* w = s_w + buf * TI_CHUNK * 3 * L;
        const int* bb = s_base + buf * TI_CHUNK;
        const int* jj = s_j + buf * TI_CHUNK;
        for (int n = warp; n < nc; n += TI_WARPS) {
            const T* wn = w + n * 3 * L;
            const C* base = tile + bb[n];
            T sx = 0, sy = 0;
#pragma unroll
            for (int it = 0; it < NIT; it++) {
                if (lane + 32 * it < L * L) {
                    T ax = 0, ay = 0;
#pragma unroll
                    for (int t = 0; t < L; t++) {
                        const C v = base[t * PY * PX + coff[it]];
                        const T wz = wn[2 * L + t];
                        ax = tfma(wz, v.x, ax);
                        ay = tfma(wz, v.y, ay);
                    }
                    const T wxy = wn[xo[it]] * wn[L + yo[it]];
                    sx = tfma(wxy, ax, sx);
                    sy = tfma(wxy, ay, sy);
                }
            }
            sx = warp_sum<T>(sx);
            sy = warp_sum<T>(sy);
            if (lane == 0) fhat[jj[n]] = make_c<T>(sx, sy);
        }
        __syncthreads();
    }
}

template <typename T, int MT>
int launch_tile3d(nfftb200_plan* p, const void* g, void* fhat, int B, int t_lo, int t_hi)
{
    using C = typename Cplx<T>::type;
    const int L = 2 * MT;
    const int PX = (int)p->bs[0] + L, PY = (int)p->bs[1] + L, PZ = (int)p->bs[2] + L;
    const size_t smem = sizeof(C) * (size_t)PX * PY * PZ +
                        2 * (sizeof(T) * TI_CHUNK * 3 * L + 2 * sizeof(int) * TI_CHUNK);
    if (smem > 227 * 1024) return -1;
    auto kern = k_interp_tile3d<T, MT>;
    CUDA_TRY(p, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid(t_hi - t_lo, B);
    kern<<<grid, TI_THREADS, smem, p->stream>>>((const C*)g, (C*)fhat, (const T*)p->d_xs, p->d_perm,
                                               p->d_tile_start, t_lo, p->M, make_geom<T>(p), make_win<T>(p));
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
    if (p->kernel_mode == 0 && is_complex && p->D == 3) {
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

int nfftb_interp(nfftb200_plan* p, const void* d_g, void* d_fhat, int B, int is_complex, int64_t t_lo,
                 int64_t t_hi)
{
    const long long i_lo = p->h_tile_start[t_lo], i_hi = p->h_tile_start[t_hi];
    return p->dtype == NFFTB200_F32
               ? interp_impl<float>(p, d_g, d_fhat, B, is_complex, (int)t_lo, (int)t_hi, i_lo, i_hi)
               : interp_impl<double>(p, d_g, d_fhat, B, is_complex, (int)t_lo, (int)t_hi, i_lo, i_hi);
}
