// toeplitz.cu -- the Gram (Toeplitz) operator either side of the NFFT path, on the device:
//   calculateToeplitzKernel!   /root/reference/NFFTTools/src/Toeplitz.jl:131-137
//       lambda = FFT(fftshift(adjoint(p) * ones)) on the 2x oversampled image grid of the plan
//   convolveToeplitzKernel!    /root/reference/NFFTTools/src/Toeplitz.jl:230-244
//       y <- crop(IFFT(lambda .* FFT(zero-pad(y))))       (IFFT normalised by 1/prod(2N), plan_ifft)
// The FFTs are cuFFT; pad, multiply and crop+scale are single-pass kernels written here.
#include <cstring>

#include "common.cuh"

struct nfftb200_toeplitz {
    int D = 0;
    int dtype = NFFTB200_F32;
    int B = 1;
    int device = 0;
    int64_t shape[NFFTB_MAX_D] = {1, 1, 1};     // image size
    int64_t os[NFFTB_MAX_D] = {1, 1, 1};        // 2 * shape
    int64_t isz = 1, osz = 1;
    cudaStream_t stream = nullptr;
    cufftHandle fft = 0;
    bool have_fft = false;
    void* d_lambda = nullptr;     // osz complex
    void* d_x = nullptr;          // B * osz complex (xOS1 / xOS2 of the reference, in place)
    void* d_y = nullptr;          // staging for host callers, B * isz complex
    bool have_kernel = false;
    int64_t launches = 0;
    std::string err;
    size_t esz() const { return dtype == NFFTB200_F32 ? 4 : 8; }
};

namespace {

int tfail(nfftb200_toeplitz* t, int code, const std::string& msg)
{
    if (t) t->err = msg;
    return nfftb_fail(nullptr, code, msg);
}

#define TCUDA_TRY(t, call)                                                                        \
    do {                                                                                          \
        cudaError_t e__ = (call);                                                                 \
        if (e__ != cudaSuccess)                                                                   \
            return tfail((t), e__ == cudaErrorMemoryAllocation ? NFFTB200_OOM : NFFTB200_CUDA_ERROR, \
                         std::string(#call) + ": " + cudaGetErrorString(e__));                    \
    } while (0)

struct Dev {
    int prev = -1;
    explicit Dev(int dev)
    {
        cudaGetDevice(&prev);
        if (prev != dev) cudaSetDevice(dev); else prev = -1;
    }
    ~Dev() { if (prev >= 0) cudaSetDevice(prev); }
};

struct Shape3 { int n[3]; int o[3]; };

// fHat = ones (OnesVector, Toeplitz.jl:250-268)
template <typename C> __global__ void k_fill_ones(C* __restrict__ v, long long n)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { v[i].x = 1; v[i].y = 0; }
}

// out[(i + n/2) mod n] = in[i] per dimension (fftshift)
template <typename C> __global__ void k_fftshift(const C* __restrict__ in, C* __restrict__ out, Shape3 s, long long total)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int i0 = (int)(i % s.n[0]);
    const long long r = i / s.n[0];
    const int i1 = (int)(r % s.n[1]), i2 = (int)(r / s.n[1]);
    int j0 = i0 + s.n[0] / 2, j1 = i1 + s.n[1] / 2, j2 = i2 + s.n[2] / 2;
    if (j0 >= s.n[0]) j0 -= s.n[0];
    if (j1 >= s.n[1]) j1 -= s.n[1];
    if (j2 >= s.n[2]) j2 -= s.n[2];
    out[((long long)j2 * s.n[1] + j1) * s.n[0] + j0] = in[i];
}

// xOS1 = 0; xOS1[CartesianIndices(y)] = y    (one pass over the oversampled array, every cell written once)
template <typename C> __global__ void k_toep_pad(const C* __restrict__ y, C* __restrict__ x, Shape3 s, long long osz,
                                                 long long isz, long long total)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const long long b = i / osz, c = i - b * osz;
    const int i0 = (int)(c % s.o[0]);
    const long long r = c / s.o[0];
    const int i1 = (int)(r % s.o[1]), i2 = (int)(r / s.o[1]);
    C v; v.x = 0; v.y = 0;
    if (i0 < s.n[0] && i1 < s.n[1] && i2 < s.n[2]) v = y[b * isz + ((long long)i2 * s.n[1] + i1) * s.n[0] + i0];
    x[i] = v;
}

// xOS2 .*= lambda
template <typename C> __global__ void k_toep_mul(C* __restrict__ x, const C* __restrict__ lam, long long osz, long long total)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const C a = x[i], l = lam[i % osz];
    C o;
    o.x = a.x * l.x - a.y * l.y;
    o.y = a.x * l.y + a.y * l.x;
    x[i] = o;
}

// y = (1/prod(2N)) * xOS1[CartesianIndices(y)]   (the scaling of plan_ifft)
template <typename C, typename T> __global__ void k_toep_crop(const C* __restrict__ x, C* __restrict__ y, Shape3 s,
                                                              long long osz, long long isz, long long total, T scale)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const long long b = i / isz, c = i - b * isz;
    const int i0 = (int)(c % s.n[0]);
    const long long r = c / s.n[0];
    const int i1 = (int)(r % s.n[1]), i2 = (int)(r / s.n[1]);
    const C v = x[b * osz + ((long long)i2 * s.o[1] + i1) * s.o[0] + i0];
    C o; o.x = v.x * scale; o.y = v.y * scale;
    y[i] = o;
}

inline unsigned nblk(long long n) { return (unsigned)((n + 255) / 256); }

}  // namespace

extern "C" {

int nfftb200_toeplitz_kernel(nfftb200_plan* p, void* lambda, int where)
{
    if (!p || !p->have_nodes) return nfftb_fail(p, NFFTB200_NO_NODES, "plan has no nodes");
    if (!lambda) return nfftb_fail(p, NFFTB200_BAD_ARGUMENT, "lambda == NULL");
    if (p->B != 1) return nfftb_fail(p, NFFTB200_UNSUPPORTED, "Toeplitz kernel needs a plan with ntransforms = 1");
    if (p->D > 3) return nfftb_fail(p, NFFTB200_UNSUPPORTED, "Toeplitz kernel: only D = 1, 2, 3 are supported");
    Dev guard(p->device);
    const size_t csz = 2 * p->esz();
    // work arrays and the image-sized FFT plan live in the plan (the kernel is often rebuilt for new trajectories)
    const int64_t need[3] = {std::max<int64_t>(p->M, 1) * (int64_t)csz, p->fsz * (int64_t)csz,
                             where == NFFTB200_DEVICE ? 0 : p->fsz * (int64_t)csz};
    for (int i = 0; i < 3; i++) {
        if (need[i] <= p->cap_toep[i]) continue;
        if (p->d_toep[i]) { cudaFree(p->d_toep[i]); p->d_toep[i] = nullptr; p->cap_toep[i] = 0; }
        CUDA_TRY(p, cudaMalloc(&p->d_toep[i], (size_t)need[i]));
        p->cap_toep[i] = need[i];
    }
    if (!p->have_fft_img) {
        long long n[3];
        for (int d = 0; d < p->D; d++) n[d] = p->N[p->D - 1 - d];
        size_t ws = 0;
        CUFFT_TRY(p, cufftCreate(&p->fft_img));
        p->have_fft_img = true;
        CUFFT_TRY(p, cufftMakePlanMany64(p->fft_img, p->D, n, nullptr, 1, p->fsz, nullptr, 1, p->fsz,
                                         p->dtype == NFFTB200_F32 ? CUFFT_C2C : CUFFT_Z2Z, 1, &ws));
    }
    CUFFT_TRY(p, cufftSetStream(p->fft_img, p->stream));
    void* d_ones = p->d_toep[0];
    void* d_img = p->d_toep[1];
    void* d_sh = where == NFFTB200_DEVICE ? lambda : p->d_toep[2];
    if (p->M > 0) {
        if (p->dtype == NFFTB200_F32) k_fill_ones<float2><<<nblk(p->M), 256, 0, p->stream>>>((float2*)d_ones, p->M);
        else k_fill_ones<double2><<<nblk(p->M), 256, 0, p->stream>>>((double2*)d_ones, p->M);
        p->launches++;
    }
    ST_TRY(nfftb200_exec_adjoint(p, d_ones, d_img, NFFTB200_DEVICE));
    Shape3 s;
    for (int d = 0; d < 3; d++) { s.n[d] = (int)p->N[d]; s.o[d] = (int)p->N[d]; }
    if (p->dtype == NFFTB200_F32) k_fftshift<float2><<<nblk(p->fsz), 256, 0, p->stream>>>((const float2*)d_img, (float2*)d_sh, s, p->fsz);
    else k_fftshift<double2><<<nblk(p->fsz), 256, 0, p->stream>>>((const double2*)d_img, (double2*)d_sh, s, p->fsz);
    // fftplan * fftshift(eigMat): unnormalised forward DFT of the image-sized array
    if (p->dtype == NFFTB200_F32) CUFFT_TRY(p, cufftExecC2C(p->fft_img, (cufftComplex*)d_sh, (cufftComplex*)d_sh, CUFFT_FORWARD));
    else CUFFT_TRY(p, cufftExecZ2Z(p->fft_img, (cufftDoubleComplex*)d_sh, (cufftDoubleComplex*)d_sh, CUFFT_FORWARD));
    p->launches += 2;
    CUDA_TRY(p, cudaGetLastError());
    if (where != NFFTB200_DEVICE) {
        CUDA_TRY(p, cudaMemcpyAsync(lambda, d_sh, (size_t)p->fsz * csz, cudaMemcpyDeviceToHost, p->stream));
        CUDA_TRY(p, cudaStreamSynchronize(p->stream));
    }
    return NFFTB200_OK;
}

int nfftb200_toeplitz_create(nfftb200_toeplitz** out, int D, const int64_t* shape, int dtype, int ntransforms, int device)
{
    if (!out) return tfail(nullptr, NFFTB200_BAD_ARGUMENT, "out == NULL");
    *out = nullptr;
    if (D < 1 || D > 3) return tfail(nullptr, NFFTB200_UNSUPPORTED, "only D = 1, 2, 3 are supported");
    if (dtype != NFFTB200_F32 && dtype != NFFTB200_F64) return tfail(nullptr, NFFTB200_UNSUPPORTED, "dtype");
    if (ntransforms < 1) return tfail(nullptr, NFFTB200_BAD_ARGUMENT, "ntransforms must be >= 1");
    if (device < 0) return tfail(nullptr, NFFTB200_CUDA_ERROR, "Toeplitz operator needs a CUDA device: no CPU fallback exists");
    nfftb200_toeplitz* t = new nfftb200_toeplitz();
    t->D = D; t->dtype = dtype; t->B = ntransforms; t->device = device;
    for (int d = 0; d < D; d++) {
        if (shape[d] < 1 || shape[d] > (1 << 29)) { delete t; return tfail(nullptr, NFFTB200_BAD_ARGUMENT, "bad shape"); }
        t->shape[d] = shape[d]; t->os[d] = 2 * shape[d];
        t->isz *= shape[d]; t->osz *= 2 * shape[d];
    }
    Dev guard(device);
    const size_t csz = 2 * t->esz();
    auto bail = [&](int code, const std::string& msg) { nfftb200_toeplitz_destroy(t); return tfail(nullptr, code, msg); };
    if (cudaMalloc(&t->d_lambda, (size_t)t->osz * csz) != cudaSuccess ||
        cudaMalloc(&t->d_x, (size_t)t->osz * t->B * csz) != cudaSuccess ||
        cudaMalloc(&t->d_y, (size_t)t->isz * t->B * csz) != cudaSuccess) {
        cudaGetLastError();
        return bail(NFFTB200_OOM, "toeplitz_create: out of device memory");
    }
    long long n[3];
    for (int d = 0; d < D; d++) n[d] = t->os[D - 1 - d];
    size_t ws = 0;
    if (cufftCreate(&t->fft) != CUFFT_SUCCESS) return bail(NFFTB200_CUDA_ERROR, "cufftCreate failed");
    t->have_fft = true;
    if (cufftMakePlanMany64(t->fft, D, n, nullptr, 1, t->osz, nullptr, 1, t->osz,
                            dtype == NFFTB200_F32 ? CUFFT_C2C : CUFFT_Z2Z, t->B, &ws) != CUFFT_SUCCESS)
        return bail(NFFTB200_CUDA_ERROR, "cufftMakePlanMany64 failed");
    cufftSetStream(t->fft, t->stream);
    *out = t;
    return NFFTB200_OK;
}

int nfftb200_toeplitz_destroy(nfftb200_toeplitz* t)
{
    if (!t) return NFFTB200_OK;
    {
        Dev guard(t->device);
        if (t->have_fft) cufftDestroy(t->fft);
        cudaFree(t->d_lambda); cudaFree(t->d_x); cudaFree(t->d_y);
    }
    delete t;
    return NFFTB200_OK;
}

int nfftb200_toeplitz_set_stream(nfftb200_toeplitz* t, void* cuda_stream)
{
    if (!t) return NFFTB200_BAD_ARGUMENT;
    Dev guard(t->device);
    TCUDA_TRY(t, cudaStreamSynchronize(t->stream));
    t->stream = (cudaStream_t)cuda_stream;
    cufftSetStream(t->fft, t->stream);
    return NFFTB200_OK;
}

int nfftb200_toeplitz_set_kernel(nfftb200_toeplitz* t, const void* lambda, int where)
{
    if (!t || !lambda) return tfail(t, NFFTB200_BAD_ARGUMENT, "toeplitz_set_kernel: NULL argument");
    Dev guard(t->device);
    TCUDA_TRY(t, cudaMemcpyAsync(t->d_lambda, lambda, (size_t)t->osz * 2 * t->esz(),
                                 where == NFFTB200_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, t->stream));
    if (where != NFFTB200_DEVICE) TCUDA_TRY(t, cudaStreamSynchronize(t->stream));
    t->have_kernel = true;
    return NFFTB200_OK;
}

int nfftb200_toeplitz_apply(nfftb200_toeplitz* t, void* y, int where)
{
    if (!t || !y) return tfail(t, NFFTB200_BAD_ARGUMENT, "toeplitz_apply: NULL argument");
    if (!t->have_kernel) return tfail(t, NFFTB200_BAD_ARGUMENT, "toeplitz_apply: no kernel set");
    Dev guard(t->device);
    const size_t csz = 2 * t->esz();
    const size_t ybytes = (size_t)t->isz * t->B * csz;
    void* dy = y;
    if (where != NFFTB200_DEVICE) {
        dy = t->d_y;
        TCUDA_TRY(t, cudaMemcpyAsync(dy, y, ybytes, cudaMemcpyHostToDevice, t->stream));
    }
    Shape3 s;
    for (int d = 0; d < 3; d++) { s.n[d] = (int)t->shape[d]; s.o[d] = (int)t->os[d]; }
    const long long tot_os = t->osz * t->B, tot_is = t->isz * t->B;
    cufftResult r;
    if (t->dtype == NFFTB200_F32) {
        float2* x = (float2*)t->d_x;
        k_toep_pad<float2><<<nblk(tot_os), 256, 0, t->stream>>>((const float2*)dy, x, s, t->osz, t->isz, tot_os);
        r = cufftExecC2C(t->fft, x, x, CUFFT_FORWARD);
        k_toep_mul<float2><<<nblk(tot_os), 256, 0, t->stream>>>(x, (const float2*)t->d_lambda, t->osz, tot_os);
        if (r == CUFFT_SUCCESS) r = cufftExecC2C(t->fft, x, x, CUFFT_INVERSE);
        k_toep_crop<float2, float><<<nblk(tot_is), 256, 0, t->stream>>>(x, (float2*)dy, s, t->osz, t->isz, tot_is,
                                                                       (float)(1.0 / (double)t->osz));
    } else {
        double2* x = (double2*)t->d_x;
        k_toep_pad<double2><<<nblk(tot_os), 256, 0, t->stream>>>((const double2*)dy, x, s, t->osz, t->isz, tot_os);
        r = cufftExecZ2Z(t->fft, x, x, CUFFT_FORWARD);
        k_toep_mul<double2><<<nblk(tot_os), 256, 0, t->stream>>>(x, (const double2*)t->d_lambda, t->osz, tot_os);
        if (r == CUFFT_SUCCESS) r = cufftExecZ2Z(t->fft, x, x, CUFFT_INVERSE);
        k_toep_crop<double2, double><<<nblk(tot_is), 256, 0, t->stream>>>(x, (double2*)dy, s, t->osz, t->isz, tot_is,
                                                                         1.0 / (double)t->osz);
    }
    t->launches += 5;
    if (r != CUFFT_SUCCESS) return tfail(t, NFFTB200_CUDA_ERROR, "toeplitz_apply: cufft error " + std::to_string((int)r));
    TCUDA_TRY(t, cudaGetLastError());
    if (where != NFFTB200_DEVICE) {
        TCUDA_TRY(t, cudaMemcpyAsync(y, dy, ybytes, cudaMemcpyDeviceToHost, t->stream));
        TCUDA_TRY(t, cudaStreamSynchronize(t->stream));
    }
    return NFFTB200_OK;
}

int nfftb200_toeplitz_sync(nfftb200_toeplitz* t)
{
    if (!t) return NFFTB200_BAD_ARGUMENT;
    Dev guard(t->device);
    TCUDA_TRY(t, cudaStreamSynchronize(t->stream));
    return NFFTB200_OK;
}

const char* nfftb200_toeplitz_last_error(nfftb200_toeplitz* t) { return t ? t->err.c_str() : nfftb200_last_error(nullptr); }

}  // extern "C"
