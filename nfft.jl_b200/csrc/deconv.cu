// deconv.cu -- K5 / K6: the diagonal stage D of the NFFT (apodisation + zero padding + fftshift).
//   forward  g[(n mod Nt)] = f[n] * prod_d LUT_d[n_d], every other grid cell = 0
//            -- replaces fill!(tmpVec,0) + deconvolve!, /root/reference/src/implementation.jl:159-160,
//               /root/reference/src/deconvolution.jl:22-45: ONE pass that writes each grid cell once.
//   adjoint  f[n] = g[(n mod Nt)] * prod_d LUT_d[n_d]
//            -- deconvolve_transpose!, /root/reference/src/deconvolution.jl:69-92.
// n_d = i_d - N_d/2 (integer division) for image index i_d in [0, N_d); multiplication order
// ((f*L1)*L2)*L3 as in the reference.  HBM-bound: 2s*(prod N + prod Nt)*B bytes per call.
#include <algorithm>

#include "common.cuh"

namespace {

// image index of grid cell u along one dim, or -1 if the cell is zero padding
__device__ __forceinline__ int img_index(int u, int N, int Nt)
{
    const int Na = N / 2, Nb = (N + 1) / 2;
    if (u < Nb) return u + Na;
    if (u >= Nt - Na) return u - (Nt - Na);
    return -1;
}
__device__ __forceinline__ int grid_index(int i, int N, int Nt)
{
    const int n = i - N / 2;
    return n < 0 ? n + Nt : n;
}

// One thread writes one 16-byte unit of the grid (two Float32 cells or one Float64 cell); row
// quantities (u1,u2 -> i1,i2, LUT factors) are block-uniform.  blockIdx.z = u2 + Nt2 * batch.
// One thread writes 16 bytes of DECONV_ROWS (8, or 1 for small / 1-D grids) consecutive grid rows: 65 536 blocks of 128
// threads were block-scheduling bound on C2 (50 us for 134 MB of plain stores).
template <typename T, int DECONV_ROWS>
__global__ void __launch_bounds__(256)
k_deconv_fwd(const typename Cplx<T>::type* __restrict__ f, typename Cplx<T>::type* __restrict__ g,
             GeomDev geo, const T* __restrict__ lut)
{
    using C = typename Cplx<T>::type;
    constexpr int VPC = 16 / (int)sizeof(C);
    const int u0 = (blockIdx.x * blockDim.x + threadIdx.x) * VPC;
    const int u2 = blockIdx.z % geo.Nt[2];
    const int r3 = blockIdx.z / geo.Nt[2];
    const int u3 = r3 % geo.Nt[3];                   // Nt[3] = 1 unless D = 4
    const int b = r3 / geo.Nt[3];
    if (u0 >= geo.Nt[0]) return;
    int i2 = geo.D > 2 ? img_index(u2, geo.N[2], geo.Nt[2]) : 0;
    const int i3 = geo.D > 3 ? img_index(u3, geo.N[3], geo.Nt[3]) : 0;
    if (i3 < 0) i2 = -1;
    const T s2 = (geo.D > 2 && i2 >= 0) ? lut[geo.N[0] + geo.N[1] + i2] : (T)1;
    const T s3 = (geo.D > 3 && i3 >= 0) ? lut[geo.N[0] + geo.N[1] + geo.N[2] + i3] : (T)1;
    int i0s[VPC];
    T s0s[VPC];
#pragma unroll
    for (int k = 0; k < VPC; k++) {
        i0s[k] = i2 >= 0 ? img_index(u0 + k, geo.N[0], geo.Nt[0]) : -1;
        s0s[k] = i0s[k] >= 0 ? lut[i0s[k]] : (T)0;
    }
    C out[DECONV_ROWS][VPC];
#pragma unroll
    for (int r = 0; r < DECONV_ROWS; r++) {
        const int u1 = blockIdx.y * DECONV_ROWS + r;
#pragma unroll
        for (int k = 0; k < VPC; k++) out[r][k] = make_c<T>(0, 0);
        const int i1 = (geo.D > 1 && u1 < geo.Nt[1]) ? img_index(u1, geo.N[1], geo.Nt[1]) : (u1 < geo.Nt[1] ? 0 : -1);
        if (i1 >= 0 && i2 >= 0) {
            const T s1 = geo.D > 1 ? lut[geo.N[0] + i1] : (T)1;
            const C* src = f + (size_t)b * geo.fsz + (((size_t)i3 * geo.N[2] + i2) * geo.N[1] + i1) * geo.N[0];
#pragma unroll
            for (int k = 0; k < VPC; k++)
                if (i0s[k] >= 0) {
                    C v = src[i0s[k]];                               // same product order as the one-row form
                    v.x *= s0s[k]; v.y *= s0s[k];
                    if (geo.D > 1) { v.x *= s1; v.y *= s1; }
                    if (geo.D > 2) { v.x *= s2; v.y *= s2; }
                    if (geo.D > 3) { v.x *= s3; v.y *= s3; }
                    out[r][k] = v;
                }
        }
    }
#pragma unroll
    for (int r = 0; r < DECONV_ROWS; r++) {
        const int u1 = blockIdx.y * DECONV_ROWS + r;
        if (u1 >= geo.Nt[1]) break;
        C* dst = g + (size_t)b * geo.gsz + (((size_t)u3 * geo.Nt[2] + u2) * geo.Nt[1] + u1) * geo.Nt[0] + u0;
        if (VPC == 2) {
            *reinterpret_cast<float4*>(dst) = make_float4((float)out[r][0].x, (float)out[r][0].y, (float)out[r][VPC - 1].x, (float)out[r][VPC - 1].y);
        } else {
            dst[0] = out[r][0];
        }
    }
}

template <typename T>
__global__ void __launch_bounds__(256)
k_deconv_adj(const typename Cplx<T>::type* __restrict__ g, typename Cplx<T>::type* __restrict__ f,
             GeomDev geo, const T* __restrict__ lut, int B)
{
    using C = typename Cplx<T>::type;
    const int i0 = blockIdx.x * blockDim.x + threadIdx.x;
    const int i1 = blockIdx.y * blockDim.y + threadIdx.y;
    const int i2 = blockIdx.z % geo.N[2], i3 = blockIdx.z / geo.N[2];       // N[3] = 1 unless D = 4
    if (i0 >= geo.N[0] || i1 >= geo.N[1]) return;
    const int u0 = grid_index(i0, geo.N[0], geo.Nt[0]);
    const int u1 = geo.D > 1 ? grid_index(i1, geo.N[1], geo.Nt[1]) : 0;
    const int u2 = geo.D > 2 ? grid_index(i2, geo.N[2], geo.Nt[2]) : 0;
    const int u3 = geo.D > 3 ? grid_index(i3, geo.N[3], geo.Nt[3]) : 0;
    const long long gq = (((long long)u3 * geo.Nt[2] + u2) * geo.Nt[1] + u1) * geo.Nt[0] + u0;
    const long long fq = (((long long)i3 * geo.N[2] + i2) * geo.N[1] + i1) * geo.N[0] + i0;
    const T s0 = lut[i0];
    const T s1 = geo.D > 1 ? lut[geo.N[0] + i1] : (T)1;
    const T s2 = geo.D > 2 ? lut[geo.N[0] + geo.N[1] + i2] : (T)1;
    const T s3 = geo.D > 3 ? lut[geo.N[0] + geo.N[1] + geo.N[2] + i3] : (T)1;
    for (int b = 0; b < B; b++) {
        C v = g[b * geo.gsz + gq];
        v.x *= s0; v.y *= s0;
        if (geo.D > 1) { v.x *= s1; v.y *= s1; }
        if (geo.D > 2) { v.x *= s2; v.y *= s2; }
        if (geo.D > 3) { v.x *= s3; v.y *= s3; }
        f[b * geo.fsz + fq] = v;
    }
}

// ---- slab variants for the node-sharded multi-GPU path (comm.cu).  The transposed slab holds, for the grid
// dimension D-2 ("mid"), only the range [mid_off, mid_off + Ms); layout [outer = dim D-1][Ms][inner = dims < D-2].
template <typename T>
__global__ void k_slab_deconv_fwd(const typename Cplx<T>::type* __restrict__ f, typename Cplx<T>::type* __restrict__ slab,
                                  GeomDev geo, const T* __restrict__ lut, int mid_off, int Ms)
{
    using C = typename Cplx<T>::type;
    const int D = geo.D;
    const long long inner = D == 3 ? geo.Nt[0] : 1;
    const long long n = inner * Ms * geo.Nt[D - 1];
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(e % inner);
        long long r = e / inner;
        const int ml = (int)(r % Ms);
        const int o = (int)(r / Ms);
        int u[3] = {0, 0, 0};
        if (D == 3) { u[0] = i; u[1] = mid_off + ml; u[2] = o; } else { u[0] = mid_off + ml; u[1] = o; }
        C v = make_c<T>(0, 0);
        int idx[3] = {0, 0, 0};
        bool in = true;
        for (int d = 0; d < D; d++) { idx[d] = img_index(u[d], geo.N[d], geo.Nt[d]); in = in && idx[d] >= 0; }
        if (in) {
            v = f[((long long)idx[2] * geo.N[1] + idx[1]) * geo.N[0] + idx[0]];
            int lo = 0;
            for (int d = 0; d < D; d++) { const T s = lut[lo + idx[d]]; v.x *= s; v.y *= s; lo += geo.N[d]; }
        }
        slab[e] = v;
    }
}

template <typename T>
__global__ void k_slab_deconv_adj(const typename Cplx<T>::type* __restrict__ slab, typename Cplx<T>::type* __restrict__ f,
                                  GeomDev geo, const T* __restrict__ lut, int mid_off, int Ms)
{
    using C = typename Cplx<T>::type;
    const int D = geo.D;
    const long long inner = D == 3 ? geo.Nt[0] : 1;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < geo.fsz; e += (long long)gridDim.x * blockDim.x) {
        int idx[3];
        long long r = e;
        idx[0] = (int)(r % geo.N[0]); r /= geo.N[0];
        idx[1] = (int)(r % geo.N[1]); idx[2] = (int)(r / geo.N[1]);
        int u[3] = {0, 0, 0};
        for (int d = 0; d < D; d++) u[d] = grid_index(idx[d], geo.N[d], geo.Nt[d]);
        const int um = u[D - 2] - mid_off;
        if (um < 0 || um >= Ms) continue;                         // another rank owns this part of the image
        const long long se = ((long long)u[D - 1] * Ms + um) * inner + (D == 3 ? u[0] : 0);
        C v = slab[se];
        int lo = 0;
        for (int d = 0; d < D; d++) { const T s = lut[lo + idx[d]]; v.x *= s; v.y *= s; lo += geo.N[d]; }
        f[e] = v;
    }
}

inline void launch_dims(int n0, int n1, int n2, dim3& grid, dim3& block)
{
    int bx = 32;
    while (bx < 256 && bx < n0) bx <<= 1;
    int by = 256 / bx;
    if (by > n1) by = n1 < 1 ? 1 : n1;
    block = dim3(bx, by, 1);
    grid = dim3((n0 + bx - 1) / bx, (n1 + by - 1) / by, n2);
}

template <typename T> int deconv_impl(nfftb200_plan* p, const void* src, void* dst, int B, bool adj)
{
    using C = typename Cplx<T>::type;
    GeomDev geo = make_geom<T>(p);
    dim3 grid, block;
    if (!adj) {
        constexpr int VPC = 16 / (int)sizeof(C);
        const int units = (geo.Nt[0] + VPC - 1) / VPC;                      // Nt[0] is even
        int bx = 32;
        while (bx < 256 && bx < units) bx <<= 1;
        const long long blocks8 = (long long)((units + bx - 1) / bx) * ((geo.Nt[1] + 7) / 8) * geo.Nt[2] * geo.Nt[3] * B;
        if (geo.Nt[1] >= 8 && blocks8 >= 8 * 148) {
            grid = dim3((units + bx - 1) / bx, (geo.Nt[1] + 7) / 8, geo.Nt[2] * geo.Nt[3] * B);
            k_deconv_fwd<T, 8><<<grid, bx, 0, p->stream>>>((const C*)src, (C*)dst, geo, (const T*)p->d_hat_inv);
        } else {
            grid = dim3((units + bx - 1) / bx, geo.Nt[1], geo.Nt[2] * geo.Nt[3] * B);
            k_deconv_fwd<T, 1><<<grid, bx, 0, p->stream>>>((const C*)src, (C*)dst, geo, (const T*)p->d_hat_inv);
        }
    } else {
        launch_dims(geo.N[0], geo.N[1], geo.N[2] * geo.N[3], grid, block);
        k_deconv_adj<T><<<grid, block, 0, p->stream>>>((const C*)src, (C*)dst, geo,
                                                       (const T*)p->d_hat_inv, B);
    }
    p->launches++;
    CUDA_TRY(p, cudaGetLastError());
    return NFFTB200_OK;
}

}  // namespace

int nfftb_slab_deconvolve(nfftb200_plan* p, const void* d_f, void* d_slab, int64_t mid_off, int64_t Ms)
{
    const long long n = p->gsz / p->nranks;
    const int blocks = (int)std::min<long long>((n + 255) / 256, 148 * 32);
    if (p->dtype == NFFTB200_F32)
        k_slab_deconv_fwd<float><<<blocks, 256, 0, p->stream>>>((const float2*)d_f, (float2*)d_slab, make_geom<float>(p), (const float*)p->d_hat_inv, (int)mid_off, (int)Ms);
    else
        k_slab_deconv_fwd<double><<<blocks, 256, 0, p->stream>>>((const double2*)d_f, (double2*)d_slab, make_geom<double>(p), (const double*)p->d_hat_inv, (int)mid_off, (int)Ms);
    p->launches++;
    CUDA_TRY(p, cudaGetLastError());
    return NFFTB200_OK;
}
int nfftb_slab_deconvolve_transpose(nfftb200_plan* p, const void* d_slab, void* d_f, int64_t mid_off, int64_t Ms)
{
    const int blocks = (int)std::min<long long>((p->fsz + 255) / 256, 148 * 32);
    if (p->dtype == NFFTB200_F32)
        k_slab_deconv_adj<float><<<blocks, 256, 0, p->stream>>>((const float2*)d_slab, (float2*)d_f, make_geom<float>(p), (const float*)p->d_hat_inv, (int)mid_off, (int)Ms);
    else
        k_slab_deconv_adj<double><<<blocks, 256, 0, p->stream>>>((const double2*)d_slab, (double2*)d_f, make_geom<double>(p), (const double*)p->d_hat_inv, (int)mid_off, (int)Ms);
    p->launches++;
    CUDA_TRY(p, cudaGetLastError());
    return NFFTB200_OK;
}

int nfftb_deconvolve(nfftb200_plan* p, const void* d_f, void* d_g, int B)
{
    return p->dtype == NFFTB200_F32 ? deconv_impl<float>(p, d_f, d_g, B, false)
                                    : deconv_impl<double>(p, d_f, d_g, B, false);
}
int nfftb_deconvolve_transpose(nfftb200_plan* p, const void* d_g, void* d_f, int B)
{
    return p->dtype == NFFTB200_F32 ? deconv_impl<float>(p, d_g, d_f, B, true)
                                    : deconv_impl<double>(p, d_g, d_f, B, true);
}
