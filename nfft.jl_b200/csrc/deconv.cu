// deconv.cu -- K5 / K6: the diagonal stage D of the NFFT (apodisation + zero padding + fftshift).
//   forward  g[(n mod Nt)] = f[n] * prod_d LUT_d[n_d], every other grid cell = 0
//            -- replaces fill!(tmpVec,0) + deconvolve!, /root/reference/src/implementation.jl:159-160,
//               /root/reference/src/deconvolution.jl:22-45: ONE pass that writes each grid cell once.
//   adjoint  f[n] = g[(n mod Nt)] * prod_d LUT_d[n_d]
//            -- deconvolve_transpose!, /root/reference/src/deconvolution.jl:69-92.
// n_d = i_d - N_d/2 (integer division) for image index i_d in [0, N_d); multiplication order
// ((f*L1)*L2)*L3 as in the reference.  HBM-bound: 2s*(prod N + prod Nt)*B bytes per call.
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

template <typename T>
__global__ void __launch_bounds__(256)
k_deconv_fwd(const typename Cplx<T>::type* __restrict__ f, typename Cplx<T>::type* __restrict__ g,
             GeomDev geo, const T* __restrict__ lut, int B)
{
    using C = typename Cplx<T>::type;
    const int u0 = blockIdx.x * blockDim.x + threadIdx.x;
    const int u1 = blockIdx.y * blockDim.y + threadIdx.y;
    const int u2 = blockIdx.z;
    if (u0 >= geo.Nt[0] || u1 >= geo.Nt[1]) return;
    const int i0 = img_index(u0, geo.N[0], geo.Nt[0]);
    const int i1 = geo.D > 1 ? img_index(u1, geo.N[1], geo.Nt[1]) : 0;
    const int i2 = geo.D > 2 ? img_index(u2, geo.N[2], geo.Nt[2]) : 0;
    const long long gq = ((long long)u2 * geo.Nt[1] + u1) * geo.Nt[0] + u0;
    const bool in = (i0 >= 0) && (i1 >= 0) && (i2 >= 0);
    T s0 = 0, s1 = 1, s2 = 1;
    long long fq = 0;
    if (in) {
        s0 = lut[i0];
        if (geo.D > 1) s1 = lut[geo.N[0] + i1];
        if (geo.D > 2) s2 = lut[geo.N[0] + geo.N[1] + i2];
        fq = ((long long)i2 * geo.N[1] + i1) * geo.N[0] + i0;
    }
    for (int b = 0; b < B; b++) {
        C v = make_c<T>(0, 0);
        if (in) {
            v = f[b * geo.fsz + fq];
            v.x *= s0; v.y *= s0;
            if (geo.D > 1) { v.x *= s1; v.y *= s1; }
            if (geo.D > 2) { v.x *= s2; v.y *= s2; }
        }
        g[b * geo.gsz + gq] = v;
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
    const int i2 = blockIdx.z;
    if (i0 >= geo.N[0] || i1 >= geo.N[1]) return;
    const int u0 = grid_index(i0, geo.N[0], geo.Nt[0]);
    const int u1 = geo.D > 1 ? grid_index(i1, geo.N[1], geo.Nt[1]) : 0;
    const int u2 = geo.D > 2 ? grid_index(i2, geo.N[2], geo.Nt[2]) : 0;
    const long long gq = ((long long)u2 * geo.Nt[1] + u1) * geo.Nt[0] + u0;
    const long long fq = ((long long)i2 * geo.N[1] + i1) * geo.N[0] + i0;
    const T s0 = lut[i0];
    const T s1 = geo.D > 1 ? lut[geo.N[0] + i1] : (T)1;
    const T s2 = geo.D > 2 ? lut[geo.N[0] + geo.N[1] + i2] : (T)1;
    for (int b = 0; b < B; b++) {
        C v = g[b * geo.gsz + gq];
        v.x *= s0; v.y *= s0;
        if (geo.D > 1) { v.x *= s1; v.y *= s1; }
        if (geo.D > 2) { v.x *= s2; v.y *= s2; }
        f[b * geo.fsz + fq] = v;
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
        launch_dims(geo.Nt[0], geo.Nt[1], geo.Nt[2], grid, block);
        k_deconv_fwd<T><<<grid, block, 0, p->stream>>>((const C*)src, (C*)dst, geo,
                                                       (const T*)p->d_hat_inv, B);
    } else {
        launch_dims(geo.N[0], geo.N[1], geo.N[2], grid, block);
        k_deconv_adj<T><<<grid, block, 0, p->stream>>>((const C*)src, (C*)dst, geo,
                                                       (const T*)p->d_hat_inv, B);
    }
    p->launches++;
    CUDA_TRY(p, cudaGetLastError());
    return NFFTB200_OK;
}

}  // namespace

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
