// sdc.cu -- sampling density compensation weights, entirely on the device:
//   sdc / sdc!   /root/reference/NFFTTools/src/samplingDensity.jl:59-155  (Pipe & Menon iteration on REAL data
//   through convolve_transpose!/convolve!, then the least-squares global scaling c = real(sum v) / sum |v|^2 with
//   v = A' D(w) A 1).  The two convolution operators are the plan's own kernels (real-valued instantiation); the
//   element-wise steps and reductions between them are the small kernels below, so that no iteration touches the
//   host.  The reference's `any(<=(0), weights_tmp) && throw` is evaluated once, after the last iteration.
#include "common.cuh"

namespace {

constexpr int RB = 256;        // threads of the reduction kernels
constexpr int RG = 592;        // blocks: 4 per SM

template <typename T> __global__ void k_fill(T* __restrict__ v, long long n, T val)
{
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) v[i] = val;
}

template <typename C> __global__ void k_fill_c(C* __restrict__ v, long long n)
{
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) { v[i].x = 1; v[i].y = 0; }
}

// partial[b] = max over the block's share; a second launch with one block finishes into out[0]
template <typename T> __global__ void k_max(const T* __restrict__ v, long long n, T* __restrict__ partial)
{
    __shared__ T sh[RB];
    T m = -INFINITY;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) m = fmax(m, v[i]);
    sh[threadIdx.x] = m;
    __syncthreads();
    for (int s = RB / 2; s > 0; s >>= 1) {
        if (threadIdx.x < s) sh[threadIdx.x] = fmax(sh[threadIdx.x], sh[threadIdx.x + s]);
        __syncthreads();
    }
    if (threadIdx.x == 0) partial[blockIdx.x] = sh[0];
}

// workg ./= scaling_factor   (scaling_factor lives on the device)
template <typename T> __global__ void k_div(T* __restrict__ v, long long n, const T* __restrict__ scal)
{
    const T s = scal[0];
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) v[i] = v[i] / s;
}

// weights_tmp ./= scaling_factor; any(<=0) -> flag; weights ./= weights_tmp
template <typename T> __global__ void k_update(T* __restrict__ w, T* __restrict__ tmp, long long n, const T* __restrict__ scal, int* __restrict__ flag)
{
    const T s = scal[0];
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const T t = tmp[i] / s;
        tmp[i] = t;
        if (!(t > (T)0)) *flag = 1;
        w[i] = w[i] / t;
    }
}

// workf .*= weights
template <typename T, typename C> __global__ void k_apply_w(C* __restrict__ f, const T* __restrict__ w, long long n)
{
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) { f[i].x *= w[i]; f[i].y *= w[i]; }
}

// partial sums of real(v) and |v|^2 in double
template <typename C> __global__ void k_sums(const C* __restrict__ v, long long n, double* __restrict__ partial)
{
    __shared__ double s0[RB], s1[RB];
    double a = 0, b = 0;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const double x = v[i].x, y = v[i].y;
        a += x; b += x * x + y * y;
    }
    s0[threadIdx.x] = a; s1[threadIdx.x] = b;
    __syncthreads();
    for (int s = RB / 2; s > 0; s >>= 1) {
        if (threadIdx.x < s) { s0[threadIdx.x] += s0[threadIdx.x + s]; s1[threadIdx.x] += s1[threadIdx.x + s]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) { partial[2 * blockIdx.x] = s0[0]; partial[2 * blockIdx.x + 1] = s1[0]; }
}

// c = real(sum v) / sum |v|^2 from the partials (one block), then weights .*= c (all blocks recompute c: tiny)
template <typename T> __global__ void k_scale_by_c(T* __restrict__ w, long long n, const double* __restrict__ partial, int np)
{
    __shared__ double c_sh;
    if (threadIdx.x == 0) {
        double a = 0, b = 0;
        for (int i = 0; i < np; i++) { a += partial[2 * i]; b += partial[2 * i + 1]; }
        c_sh = a / b;
    }
    __syncthreads();
    const T c = (T)c_sh;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) w[i] *= c;
}

template <typename T>
int sdc_impl(nfftb200_plan* p, int iters, void* weights_out, int where)
{
    using C = typename Cplx<T>::type;
    const cudaStream_t st = p->stream;
    const long long M = p->M, G = p->gsz, F = p->fsz;
    T *w = nullptr, *tmp = nullptr, *g = nullptr, *scal = nullptr;
    C *wf = nullptr, *wv = nullptr;
    double* part = nullptr;
    int* flag = nullptr;
    void* all[8] = {};
    auto cleanup = [&]() { for (void* b : all) if (b) cudaFree(b); };
    auto alloc = [&](void** ptr, size_t bytes, int slot) -> bool {
        if (cudaMalloc(ptr, bytes ? bytes : 16) != cudaSuccess) { cudaGetLastError(); return false; }
        all[slot] = *ptr;
        return true;
    };
    if (!alloc((void**)&w, sizeof(T) * M, 0) || !alloc((void**)&tmp, sizeof(T) * M, 1) || !alloc((void**)&g, sizeof(T) * G, 2) ||
        !alloc((void**)&wf, sizeof(C) * M, 3) || !alloc((void**)&wv, sizeof(C) * F, 4) || !alloc((void**)&scal, sizeof(T) * (RG + 1), 5) ||
        !alloc((void**)&part, sizeof(double) * 2 * RG, 6) || !alloc((void**)&flag, sizeof(int), 7)) {
        cleanup();
        return nfftb_fail(p, NFFTB200_OOM, "sdc: out of device memory");
    }
    int rc = NFFTB200_OK;
    auto run = [&]() -> int {
        CUDA_TRY(p, cudaMemsetAsync(flag, 0, sizeof(int), st));
        k_fill<T><<<RG, RB, 0, st>>>(w, M, (T)1);
        for (int i = 0; i < iters; i++) {
            ST_TRY(nfftb_spread(p, w, g, 1, 0, 0, p->ntiles));                 // convolve_transpose!(p, weights, workg)
            if (i == 0) {                                                      // scaling_factor = maximum(workg)
                k_max<T><<<RG, RB, 0, st>>>(g, G, scal + 1);
                k_max<T><<<1, RB, 0, st>>>(scal + 1, RG, scal);
            }
            k_div<T><<<RG, RB, 0, st>>>(g, G, scal);
            ST_TRY(nfftb_interp(p, g, tmp, 1, 0, 0, p->ntiles));               // convolve!(p, workg, weights_tmp)
            k_update<T><<<RG, RB, 0, st>>>(w, tmp, M, scal, flag);
            p->launches += 2 + (i == 0 ? 2 : 0);
        }
        // scale such that A' D(w) A 1 ~ 1
        k_fill_c<C><<<RG, RB, 0, st>>>(wv, F);
        ST_TRY(nfftb200_exec_forward(p, wv, wf, NFFTB200_DEVICE));
        k_apply_w<T, C><<<RG, RB, 0, st>>>(wf, w, M);
        ST_TRY(nfftb200_exec_adjoint(p, wf, wv, NFFTB200_DEVICE));
        k_sums<C><<<RG, RB, 0, st>>>(wv, F, part);
        k_scale_by_c<T><<<RG, RB, 0, st>>>(w, M, part, RG);
        p->launches += 5;
        CUDA_TRY(p, cudaGetLastError());
        int h_flag = 0;
        CUDA_TRY(p, cudaMemcpyAsync(&h_flag, flag, sizeof(int), cudaMemcpyDeviceToHost, st));
        CUDA_TRY(p, cudaMemcpyAsync(weights_out, w, sizeof(T) * (size_t)M,
                                    where == NFFTB200_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, st));
        CUDA_TRY(p, cudaStreamSynchronize(st));
        if (h_flag) return nfftb_fail(p, NFFTB200_BAD_ARGUMENT, "non-positive weights");   // samplingDensity.jl:114
        return NFFTB200_OK;
    };
    rc = run();
    if (rc != NFFTB200_OK) cudaStreamSynchronize(st);
    cleanup();
    return rc;
}

}  // namespace

extern "C" int nfftb200_sdc(nfftb200_plan* p, int iters, void* weights, int where)
{
    if (!p || !p->have_nodes) return nfftb_fail(p, NFFTB200_NO_NODES, "plan has no nodes");
    if (!weights || iters < 0) return nfftb_fail(p, NFFTB200_BAD_ARGUMENT, "sdc: bad argument");
    if (p->B != 1 || p->shard_mode == NFFTB200_SHARD_NODES)
        return nfftb_fail(p, NFFTB200_UNSUPPORTED, "sdc needs an unsharded plan with ntransforms = 1");
    if (p->M == 0) return NFFTB200_OK;
    int prev = -1;
    cudaGetDevice(&prev);
    if (prev != p->device) cudaSetDevice(p->device);
    const int rc = p->dtype == NFFTB200_F32 ? sdc_impl<float>(p, iters, weights, where) : sdc_impl<double>(p, iters, weights, where);
    if (prev >= 0 && prev != p->device) cudaSetDevice(prev);
    return rc;
}
