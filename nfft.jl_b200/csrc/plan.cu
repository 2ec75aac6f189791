// plan.cu -- the C ABI of libnfftb200.so (include/nfftb200.h): plan life cycle, nodes!, mul! forward /
// adjoint drivers and the optional AbstractNFFTs operators, mirroring NFFTPlan
// (/root/reference/src/implementation.jl:16-193) with the FFT delegated to cuFFT.
#include <algorithm>
#include <cmath>
#include <cstring>

#include "common.cuh"

int nfftb_comm_exec_adjoint(nfftb200_plan* p, const void* d_fhat, void* d_f);   // comm.cu
int nfftb_comm_exec_forward(nfftb200_plan* p, const void* d_f, void* d_fhat);
void nfftb_comm_destroy(nfftb200_plan* p);
void nfftb_comm_set_stream(nfftb200_plan* p);

static thread_local std::string g_last_error;

int nfftb_fail(nfftb200_plan* p, int code, const std::string& msg)
{
    if (p) p->err = msg;
    g_last_error = msg;
    return code;
}

namespace {

struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev)
    {
        if (dev < 0) return;                      // host-only plan: never touch the CUDA runtime
        cudaGetDevice(&prev);
        if (prev != dev) cudaSetDevice(dev); else prev = -1;
    }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

int ensure(nfftb200_plan* p, void** ptr, int64_t* cap, int64_t bytes)
{
    if (bytes <= *cap) return NFFTB200_OK;
    if (*ptr) { cudaFree(*ptr); *ptr = nullptr; *cap = 0; }
    CUDA_TRY(p, cudaMalloc(ptr, (size_t)bytes));
    *cap = bytes;
    return NFFTB200_OK;
}

int upload_table(nfftb200_plan* p, const std::vector<double>& h, void** d)
{
    if (*d) { cudaFree(*d); *d = nullptr; }
    if (h.empty()) return NFFTB200_OK;
    CUDA_TRY(p, cudaMalloc(d, h.size() * p->esz() + 16));          // slack: bulk (TMA) copies of a table are 16-byte granular
    CUDA_TRY(p, cudaMemset(*d, 0, h.size() * p->esz() + 16));
    if (p->dtype == NFFTB200_F32) {
        std::vector<float> t(h.size());
        for (size_t i = 0; i < h.size(); i++) t[i] = (float)h[i];
        CUDA_TRY(p, cudaMemcpy(*d, t.data(), t.size() * 4, cudaMemcpyHostToDevice));
    } else {
        CUDA_TRY(p, cudaMemcpy(*d, h.data(), h.size() * 8, cudaMemcpyHostToDevice));
    }
    return NFFTB200_OK;
}

// reference defaults (src/precomputation.jl:59-77: 1-D 1024, 2-D 64x64, 3-D 16^3); in 3-D the tile is
// shrunk only when the warp-private sub-tiles of the spreader would not fit 227 KB of shared memory
void default_tiles(nfftb200_plan* p)
{
    const int D = p->D;
    for (int d = 0; d < D; d++) {
        int64_t v = (D == 1) ? 1024 : (D == 2 ? 64 : (d < 3 ? 16 : 1));      // _blockSize, src/precomputation.jl:59-77
        // batched Float32 2-D plans: 16 x 16 tiles, so that the padded tiles of 32 transforms fit one CTA's shared memory
        // (the batch-stationary kernels of twod_batch.cuh)
        if (D == 2 && p->dtype == NFFTB200_F32 && p->B >= 8 && p->m <= 4) v = 16;
        p->bs[d] = std::min<int64_t>(v, p->Nt[d]);
    }
    if (D == 3) {
        static const int cand[][3] = {{16, 16, 16}, {16, 16, 8}, {16, 8, 8}, {8, 8, 8}, {8, 8, 4}, {8, 4, 4}, {4, 4, 4}};
        for (auto& c : cand) {
            int64_t bs[3];
            for (int d = 0; d < 3; d++) bs[d] = std::min<int64_t>(c[d], p->Nt[d]);
            const size_t need = nfftb_spread3d_smem(p->dtype, p->m, bs);
            if (need == 0) break;                       // no tiled kernel for this m: keep the reference default
            if (need <= 227 * 1024) { for (int d = 0; d < 3; d++) p->bs[d] = bs[d]; break; }
        }
    }
}

int make_fft(nfftb200_plan* p)
{
    const cufftType ty = p->dtype == NFFTB200_F32 ? CUFFT_C2C : CUFFT_Z2Z;
    CUFFT_TRY(p, cufftCreate(&p->fft));
    p->have_fft = true;
    size_t ws = 0;
    if (p->D == 4) {
        // cuFFT plans have rank <= 3: a batched 3-D transform over dims 1..3 of every (u4, transform) slice, then a strided
        // 1-D transform along dim 4
        const long long vol3 = p->Nt[0] * p->Nt[1] * p->Nt[2];
        long long n3[3] = {p->Nt[2], p->Nt[1], p->Nt[0]};
        CUFFT_TRY(p, cufftMakePlanMany64(p->fft, 3, n3, nullptr, 1, vol3, nullptr, 1, vol3, ty, p->Nt[3] * p->B, &ws));
        CUFFT_TRY(p, cufftSetStream(p->fft, p->stream));
        long long n1[1] = {p->Nt[3]}, emb[1] = {p->Nt[3]};
        CUFFT_TRY(p, cufftCreate(&p->fft_d4));
        p->have_fft_d4 = true;
        CUFFT_TRY(p, cufftMakePlanMany64(p->fft_d4, 1, n1, emb, vol3, 1, emb, vol3, 1, ty, vol3, &ws));
        CUFFT_TRY(p, cufftSetStream(p->fft_d4, p->stream));
        return NFFTB200_OK;
    }
    long long nn[3] = {1, 1, 1};
    for (int d = 0; d < p->D; d++) nn[d] = p->Nt[p->D - 1 - d];      // column-major -> cuFFT row-major
    CUFFT_TRY(p, cufftMakePlanMany64(p->fft, p->D, nn, nullptr, 1, p->gsz, nullptr, 1, p->gsz, ty, p->B, &ws));
    CUFFT_TRY(p, cufftSetStream(p->fft, p->stream));
    return NFFTB200_OK;
}

// Pruned 3-D FFT (single GPU): the zero padding of D means that, in the forward transform, only the z-planes
// u2 in [0, ceil(N2/2)) U [Nt2 - N2/2, Nt2) are non-zero before the FFT, and in the adjoint only those planes are
// read after it.  So the 2-D (x,y) FFTs run on those planes only (two contiguous batches) and the 1-D FFT along z
// runs on everything: 2/3 of the work of the full 3-D transform for sigma = 2.
int make_fft_pruned(nfftb200_plan* p)
{
    p->have_pruned = false;
    if (p->D != 3) return NFFTB200_OK;
    const int64_t Na = p->N[2] / 2, Nb = (p->N[2] + 1) / 2;
    if (Na + Nb > (3 * p->Nt[2]) / 4 || Na < 1) return NFFTB200_OK;       // not enough padding to pay off
    p->zlo_planes = Nb; p->zhi_planes = Na;
    const cufftType ty = p->dtype == NFFTB200_F32 ? CUFFT_C2C : CUFFT_Z2Z;
    const long long plane = p->Nt[0] * p->Nt[1];
    size_t ws = 0;
    long long n2[2] = {p->Nt[1], p->Nt[0]};
    if (cufftCreate(&p->fft_xy) != CUFFT_SUCCESS) return NFFTB200_OK;
    // one plan with batch = max(Na, Nb) planes; the shorter range re-uses it through a second plan only if sizes differ
    if (Na != Nb || cufftMakePlanMany64(p->fft_xy, 2, n2, nullptr, 1, plane, nullptr, 1, plane, ty, Nb, &ws) != CUFFT_SUCCESS) {
        cufftDestroy(p->fft_xy); p->fft_xy = 0;
        return NFFTB200_OK;                                                 // odd N2: keep the full plan
    }
    long long n1[1] = {p->Nt[2]}, emb[1] = {p->Nt[2]};
    if (cufftCreate(&p->fft_z) != CUFFT_SUCCESS ||
        cufftMakePlanMany64(p->fft_z, 1, n1, emb, plane, 1, emb, plane, 1, ty, plane, &ws) != CUFFT_SUCCESS) {
        cufftDestroy(p->fft_xy); p->fft_xy = 0;
        if (p->fft_z) { cufftDestroy(p->fft_z); p->fft_z = 0; }
        return NFFTB200_OK;
    }
    cufftSetStream(p->fft_xy, p->stream);
    cufftSetStream(p->fft_z, p->stream);
    p->have_pruned = true;
    return NFFTB200_OK;
}

template <typename CT, typename F> int pruned_exec(nfftb200_plan* p, CT* grid, int dir, F exec)
{
    const int cdir = dir < 0 ? CUFFT_FORWARD : CUFFT_INVERSE;
    const long long plane = p->Nt[0] * p->Nt[1];
    for (int b = 0; b < p->B; b++) {
        CT* gb = grid + (size_t)b * p->gsz;
        CT* hi = gb + (size_t)(p->Nt[2] - p->zhi_planes) * plane;
        if (dir < 0) {       // forward: 2-D on the non-zero planes, then z
            CUFFT_TRY(p, exec(p->fft_xy, gb, gb, cdir));
            CUFFT_TRY(p, exec(p->fft_xy, hi, hi, cdir));
            CUFFT_TRY(p, exec(p->fft_z, gb, gb, cdir));
        } else {             // adjoint: z on everything, then 2-D on the planes that are read
            CUFFT_TRY(p, exec(p->fft_z, gb, gb, cdir));
            CUFFT_TRY(p, exec(p->fft_xy, gb, gb, cdir));
            CUFFT_TRY(p, exec(p->fft_xy, hi, hi, cdir));
        }
        p->launches += 3;
    }
    return NFFTB200_OK;
}

// pruned = true only from the mul! drivers (deconvolve -> FFT -> ... ), where the zero structure is known
int run_fft(nfftb200_plan* p, void* grid, int dir, bool pruned = false)
{
    if (pruned && p->have_pruned && p->kernel_mode != 4) {
        if (p->dtype == NFFTB200_F32) return pruned_exec<cufftComplex>(p, (cufftComplex*)grid, dir, cufftExecC2C);
        return pruned_exec<cufftDoubleComplex>(p, (cufftDoubleComplex*)grid, dir, cufftExecZ2Z);
    }
    const int cdir = dir < 0 ? CUFFT_FORWARD : CUFFT_INVERSE;
    if (p->dtype == NFFTB200_F32)
        CUFFT_TRY(p, cufftExecC2C(p->fft, (cufftComplex*)grid, (cufftComplex*)grid, cdir));
    else
        CUFFT_TRY(p, cufftExecZ2Z(p->fft, (cufftDoubleComplex*)grid, (cufftDoubleComplex*)grid, cdir));
    p->launches++;
    if (p->have_fft_d4) {
        for (int b = 0; b < p->B; b++) {
            if (p->dtype == NFFTB200_F32) {
                cufftComplex* gb = (cufftComplex*)grid + (size_t)b * p->gsz;
                CUFFT_TRY(p, cufftExecC2C(p->fft_d4, gb, gb, cdir));
            } else {
                cufftDoubleComplex* gb = (cufftDoubleComplex*)grid + (size_t)b * p->gsz;
                CUFFT_TRY(p, cufftExecZ2Z(p->fft_d4, gb, gb, cdir));
            }
            p->launches++;
        }
    }
    return NFFTB200_OK;
}

void rec(nfftb200_plan* p, int i)
{
    if (p->timing) cudaEventRecord(p->ev[i], p->stream);
}

int stage_in(nfftb200_plan* p, const void* src, int64_t bytes, int where, void** slot, int64_t* cap,
             const void** out)
{
    if (where == NFFTB200_DEVICE) { *out = src; return NFFTB200_OK; }
    ST_TRY(ensure(p, slot, cap, bytes));
    CUDA_TRY(p, cudaMemcpyAsync(*slot, src, (size_t)bytes, cudaMemcpyHostToDevice, p->stream));
    *out = *slot;
    return NFFTB200_OK;
}

int stage_out_buf(nfftb200_plan* p, void* dst, int64_t bytes, int where, void** slot, int64_t* cap, void** out)
{
    if (where == NFFTB200_DEVICE) { *out = dst; return NFFTB200_OK; }
    ST_TRY(ensure(p, slot, cap, bytes));
    *out = *slot;
    return NFFTB200_OK;
}

int stage_back(nfftb200_plan* p, void* dst, const void* dsrc, int64_t bytes, int where)
{
    if (where == NFFTB200_DEVICE) return NFFTB200_OK;
    CUDA_TRY(p, cudaMemcpyAsync(dst, dsrc, (size_t)bytes, cudaMemcpyDeviceToHost, p->stream));
    CUDA_TRY(p, cudaStreamSynchronize(p->stream));
    return NFFTB200_OK;
}

// ---- asynchronous host-buffer mode -----------------------------------------------------------------------
int async_setup(nfftb200_plan* p, int dir, int64_t bytes_in, int64_t bytes_out)
{
    if (!p->s_up) {
        CUDA_TRY(p, cudaStreamCreateWithFlags(&p->s_up, cudaStreamNonBlocking));
        CUDA_TRY(p, cudaStreamCreateWithFlags(&p->s_down, cudaStreamNonBlocking));
        CUDA_TRY(p, cudaEventCreateWithFlags(&p->e_up, cudaEventDisableTiming));
        CUDA_TRY(p, cudaEventCreateWithFlags(&p->e_done, cudaEventDisableTiming));
        for (int i = 0; i < 2; i++) {
            CUDA_TRY(p, cudaEventCreateWithFlags(&p->a_in[i].free_ev, cudaEventDisableTiming));
            CUDA_TRY(p, cudaEventCreateWithFlags(&p->a_out[i].free_ev, cudaEventDisableTiming));
        }
    }
    nfftb200_plan::AsyncSlot* slots[2] = {&p->a_in[dir], &p->a_out[dir]};
    const int64_t need[2] = {bytes_in, bytes_out};
    for (int i = 0; i < 2; i++) {
        if (need[i] <= slots[i]->cap) continue;
        // growing a staging buffer: nothing may still be using the old one
        CUDA_TRY(p, cudaStreamSynchronize(p->s_up));
        CUDA_TRY(p, cudaStreamSynchronize(p->stream));
        CUDA_TRY(p, cudaStreamSynchronize(p->s_down));
        ST_TRY(ensure(p, &slots[i]->d, &slots[i]->cap, need[i]));
        slots[i]->used = false;
    }
    return NFFTB200_OK;
}

// upload `src` into the direction's input slot on the upload stream; the compute stream waits for it and for the
// previous download out of the output slot
int async_begin(nfftb200_plan* p, int dir, const void* src, int64_t bytes_in)
{
    nfftb200_plan::AsyncSlot& in = p->a_in[dir];
    nfftb200_plan::AsyncSlot& out = p->a_out[dir];
    if (in.used) CUDA_TRY(p, cudaStreamWaitEvent(p->s_up, in.free_ev, 0));
    CUDA_TRY(p, cudaMemcpyAsync(in.d, src, (size_t)bytes_in, cudaMemcpyHostToDevice, p->s_up));
    CUDA_TRY(p, cudaEventRecord(p->e_up, p->s_up));
    CUDA_TRY(p, cudaStreamWaitEvent(p->stream, p->e_up, 0));
    if (out.used) CUDA_TRY(p, cudaStreamWaitEvent(p->stream, out.free_ev, 0));
    return NFFTB200_OK;
}

// the transform is queued: release the input slot, download the output slot on the download stream
int async_end(nfftb200_plan* p, int dir, void* dst, int64_t bytes_out)
{
    nfftb200_plan::AsyncSlot& in = p->a_in[dir];
    nfftb200_plan::AsyncSlot& out = p->a_out[dir];
    CUDA_TRY(p, cudaEventRecord(in.free_ev, p->stream));
    CUDA_TRY(p, cudaEventRecord(p->e_done, p->stream));
    CUDA_TRY(p, cudaStreamWaitEvent(p->s_down, p->e_done, 0));
    CUDA_TRY(p, cudaMemcpyAsync(dst, out.d, (size_t)bytes_out, cudaMemcpyDeviceToHost, p->s_down));
    CUDA_TRY(p, cudaEventRecord(out.free_ev, p->s_down));
    in.used = out.used = true;
    return NFFTB200_OK;
}

}  // namespace

extern "C" {

int nfftb200_version(void) { return 100; }

const char* nfftb200_status_string(int s)
{
    switch (s) {
        case NFFTB200_OK: return "ok";
        case NFFTB200_BAD_NODE_RANGE: return "nodes out of range [-1/2, 1/2]";
        case NFFTB200_BAD_DIM: return "node dimension does not match plan dimension";
        case NFFTB200_SIZE_MISMATCH: return "data is not consistent with the plan";
        case NFFTB200_UNSUPPORTED: return "unsupported option";
        case NFFTB200_CUDA_ERROR: return "CUDA error";
        case NFFTB200_NCCL_ERROR: return "NCCL error";
        case NFFTB200_OOM: return "out of device memory";
        case NFFTB200_BAD_ARGUMENT: return "bad argument";
        case NFFTB200_NO_NODES: return "plan has no nodes";
        default: return "unknown status";
    }
}

const char* nfftb200_last_error(nfftb200_plan* p) { return p ? p->err.c_str() : g_last_error.c_str(); }

int nfftb200_accuracy_params(int m_in, double sigma_in, double reltol_in, int* m_out, double* sigma_out,
                             double* reltol_out)
{
    // accuracyParams / reltolToParams / paramsToReltol, AbstractNFFTs/src/misc.jl:44-81
    int m;
    double sigma, reltol;
    auto from_reltol = [](double r, int& mm, double& ss) {
        const int w = (int)std::ceil(std::log(1.0 / r) / std::log(10.0)) + 1;
        mm = w / 2;
        ss = 2.0;
    };
    if (reltol_in > 0) {
        reltol = reltol_in;
        from_reltol(reltol, m, sigma);
    } else if (m_in > 0 && sigma_in > 0) {
        m = m_in; sigma = sigma_in;
        reltol = std::pow(10.0, -(2.0 * m - 1.0));
    } else {
        reltol = 1e-9;
        from_reltol(reltol, m, sigma);
    }
    if (m_out) *m_out = m;
    if (sigma_out) *sigma_out = sigma;
    if (reltol_out) *reltol_out = reltol;
    return NFFTB200_OK;
}

int nfftb200_plan_create(nfftb200_plan** out, int D, const int64_t* N, int dtype, int m, double sigma,
                         int window, int precompute, int ntransforms, const int64_t* block_size, int device)
{
    if (!out) return nfftb_fail(nullptr, NFFTB200_BAD_ARGUMENT, "out == NULL");
    *out = nullptr;
    if (D < 1 || D > NFFTB_MAX_D) return nfftb_fail(nullptr, NFFTB200_UNSUPPORTED, "only D = 1, 2, 3, 4 are supported");
    if (dtype != NFFTB200_F32 && dtype != NFFTB200_F64) return nfftb_fail(nullptr, NFFTB200_UNSUPPORTED, "dtype");
    if (window < NFFTB200_KAISER_BESSEL || window > NFFTB200_EXP_SQRT)
        return nfftb_fail(nullptr, NFFTB200_UNSUPPORTED, "Window not yet implemented!");
    if (precompute < NFFTB200_FULL || precompute > NFFTB200_POLYNOMIAL)
        return nfftb_fail(nullptr, NFFTB200_UNSUPPORTED, "precompute flag not supported");
    if (m < 1 || m > NFFTB_MAX_M) return nfftb_fail(nullptr, NFFTB200_UNSUPPORTED, "m must be in 1..8");
    if (!(sigma >= 1.0)) return nfftb_fail(nullptr, NFFTB200_BAD_ARGUMENT, "sigma must be >= 1");
    if (ntransforms < 1) return nfftb_fail(nullptr, NFFTB200_BAD_ARGUMENT, "ntransforms must be >= 1");

    nfftb200_plan* p = new nfftb200_plan();
    p->D = D; p->dtype = dtype; p->m = m; p->precompute = precompute; p->B = ntransforms; p->device = device; p->window = window;
    // initParams, src/precomputation.jl:14-29
    static const int m2K[9] = {1, 3, 7, 9, 14, 17, 20, 23, 24};
    p->lut_size = ((int64_t)1 << m2K[std::min(m + 1, 9) - 1]) * m;
    // params.σ*N[d] is evaluated in T: σ::Float32 times an Int is a Float32 product (src/precomputation.jl:25-27), so
    // for non-dyadic σ the Float32 plan's Ñ can differ from the Float64 one (σ=1.1f, N=10: 10 vs 12)
    for (int d = 0; d < D; d++) {
        if (N[d] < 1) { delete p; return nfftb_fail(nullptr, NFFTB200_BAD_ARGUMENT, "N[d] must be >= 1"); }
        p->N[d] = N[d];
        const double prod = dtype == NFFTB200_F32 ? (double)((float)sigma * (float)N[d]) : sigma * (double)N[d];
        p->Nt[d] = ((int64_t)std::ceil(prod) / 2) * 2;
        if (p->Nt[d] > (1 << 30)) { delete p; return nfftb_fail(nullptr, NFFTB200_UNSUPPORTED, "grid dimension too large"); }
    }
    if (dtype == NFFTB200_F32) {
        const float s = (float)((double)p->Nt[0] / (double)p->N[0]);
        p->sigma = s;
        p->b = (double)((float)M_PI * (2.0f - 1.0f / s));       // pi*(2-1/sigma) evaluated in Float32
        p->beta = (double)((float)M_PI * (float)m * (2.0f - 1.0f / s));
    } else {
        p->sigma = (double)p->Nt[0] / (double)p->N[0];
        p->b = M_PI * (2.0 - 1.0 / p->sigma);
        p->beta = M_PI * (double)m * (2.0 - 1.0 / p->sigma);
    }
    if (window == NFFTB200_EXP_SQRT) {       // exp-sqrt shape parameter (same expression in T as the device sees it)
        if (dtype == NFFTB200_F32) p->beta = (double)(0.97f * (float)M_PI * (float)(2 * m) * (1.0f - 0.5f / (float)p->sigma));
        else p->beta = 0.97 * M_PI * (double)(2 * m) * (1.0 - 0.5 / p->sigma);
    }
    if (block_size) {
        for (int d = 0; d < D; d++) {
            if (block_size[d] < 1) { delete p; return nfftb_fail(nullptr, NFFTB200_BAD_ARGUMENT, "blockSize must be >= 1"); }
            p->bs[d] = block_size[d];
        }
    } else {
        default_tiles(p);
    }
    p->ntiles = 1; p->gsz = 1; p->fsz = 1;
    for (int d = 0; d < D; d++) {
        p->nb[d] = (p->Nt[d] + p->bs[d] - 1) / p->bs[d];
        p->ntiles *= p->nb[d];
        p->gsz *= p->Nt[d];
        p->fsz *= p->N[d];
    }
    if (p->ntiles >= ((int64_t)1 << 31)) { delete p; return nfftb_fail(nullptr, NFFTB200_UNSUPPORTED, "too many tiles"); }
    // the batch rides on gridDim.y / gridDim.z (limit 65535) in the deconvolve, gather and tile kernels
    if ((int64_t)ntransforms > 65535 || (D >= 3 && p->Nt[2] * p->Nt[3] * (int64_t)ntransforms > 65535)) {
        delete p;
        return nfftb_fail(nullptr, NFFTB200_UNSUPPORTED, "ntransforms too large for one plan (grid dimension limit 65535): split the batch");
    }
    nfftb_build_tables(p);
    if (device < 0) {   // host-only plan (parameters + tables, for CPU-side checks); every exec call fails
        *out = p;
        return NFFTB200_OK;
    }

    DeviceGuard guard(device);
    int st = [&]() -> int {
        CUDA_TRY(p, cudaStreamCreateWithFlags(&p->stream, cudaStreamNonBlocking));
        p->own_stream = true;
        for (int i = 0; i < 4; i++) CUDA_TRY(p, cudaEventCreate(&p->ev[i]));
        for (int i = 0; i < 6; i++) CUDA_TRY(p, cudaEventCreate(&p->evk[i]));
        ST_TRY(upload_table(p, p->h_hat_inv, &p->d_hat_inv));
        ST_TRY(upload_table(p, p->h_poly, &p->d_poly));
        ST_TRY(upload_table(p, p->h_lin, &p->d_lin));
        CUDA_TRY(p, cudaMalloc(&p->d_grid, (size_t)p->gsz * p->B * 2 * p->esz()));
        CUDA_TRY(p, cudaMalloc(&p->d_tile_start, sizeof(int32_t) * (size_t)(p->ntiles + 1)));
        CUDA_TRY(p, cudaMalloc(&p->d_flag, sizeof(int)));
        ST_TRY(make_fft(p));
        ST_TRY(make_fft_pruned(p));
        return NFFTB200_OK;
    }();
    if (st != NFFTB200_OK) {
        g_last_error = p->err;
        nfftb200_destroy(p);
        return st;
    }
    *out = p;
    return NFFTB200_OK;
}

int nfftb200_destroy(nfftb200_plan* p)
{
    if (!p) return NFFTB200_OK;
    DeviceGuard guard(p->device);
    if (p->s_up) cudaStreamSynchronize(p->s_up);
    if (p->stream) cudaStreamSynchronize(p->stream);
    if (p->s_down) cudaStreamSynchronize(p->s_down);
    for (int i = 0; i < 2; i++) {
        if (p->a_in[i].d) cudaFree(p->a_in[i].d);
        if (p->a_out[i].d) cudaFree(p->a_out[i].d);
        if (p->a_in[i].free_ev) cudaEventDestroy(p->a_in[i].free_ev);
        if (p->a_out[i].free_ev) cudaEventDestroy(p->a_out[i].free_ev);
    }
    if (p->have_fft_img) cufftDestroy(p->fft_img);
    for (void* b : p->d_toep) if (b) cudaFree(b);
    if (p->e_up) cudaEventDestroy(p->e_up);
    if (p->e_done) cudaEventDestroy(p->e_done);
    if (p->s_up) cudaStreamDestroy(p->s_up);
    if (p->s_down) cudaStreamDestroy(p->s_down);
    nfftb_comm_destroy(p);
    if (p->have_fft) cufftDestroy(p->fft);
    if (p->have_fft_d4) cufftDestroy(p->fft_d4);
    if (p->have_pruned) { cufftDestroy(p->fft_xy); cufftDestroy(p->fft_z); }
    void* bufs[] = {p->d_hat_inv, p->d_poly, p->d_lin, p->d_grid, p->d_xs, p->d_tile_start, p->d_keys[0],
                    p->d_keys[1], p->d_vals[0], p->d_vals[1], p->d_hist, p->d_flag, p->d_stage_f,
                    p->d_stage_h, p->d_stage_k, p->d_stage_g, p->d_slab, p->d_tilebuf, p->d_items, p->d_tile_items,
                    p->d_xs2, p->d_perm2, p->d_bin_start, p->d_expect, p->d_ready, p->d_pair_items, p->d_item_stride, p->d_inv2};
    for (void* b : bufs) if (b) cudaFree(b);
    for (int i = 0; i < 4; i++) if (p->ev[i]) cudaEventDestroy(p->ev[i]);
    for (int i = 0; i < 6; i++) if (p->evk[i]) cudaEventDestroy(p->evk[i]);
    if (p->own_stream && p->stream) cudaStreamDestroy(p->stream);
    delete p;
    return NFFTB200_OK;
}

int nfftb200_set_nodes(nfftb200_plan* p, const void* k, int64_t M, int where)
{
    if (!p) return NFFTB200_BAD_ARGUMENT;
    if (M < 0 || M >= ((int64_t)1 << 31)) return nfftb_fail(p, NFFTB200_BAD_ARGUMENT, "M out of range");
    if (p->device < 0) return nfftb_fail(p, NFFTB200_CUDA_ERROR, "host-only plan (device < 0): no CPU fallback exists");
    DeviceGuard guard(p->device);
    if (p->timing) cudaEventRecord(p->ev[0], p->stream);
    p->have_nodes = false;
    p->have_bins = false;
    p->have_inv2 = false;
    if (M > p->cap_nodes) {
        void* old[] = {p->d_xs, p->d_keys[0], p->d_keys[1], p->d_vals[0], p->d_vals[1]};
        for (void* b : old) if (b) cudaFree(b);
        p->d_xs = nullptr; p->d_keys[0] = p->d_keys[1] = nullptr; p->d_vals[0] = p->d_vals[1] = nullptr;
        p->cap_nodes = 0;
        const size_t n = (size_t)std::max<int64_t>(M, 1);
        CUDA_TRY(p, cudaMalloc(&p->d_xs, n * p->D * p->esz()));
        for (int i = 0; i < 2; i++) {
            CUDA_TRY(p, cudaMalloc((void**)&p->d_keys[i], n * 4));
            CUDA_TRY(p, cudaMalloc((void**)&p->d_vals[i], n * 4));
        }
        p->cap_nodes = M;
    }
    const int64_t nCTA = (M + 4095) / 4096;
    ST_TRY(ensure(p, (void**)&p->d_hist, &p->cap_hist, (std::max<int64_t>(1, nCTA) * 256 + std::max<int64_t>(1, nCTA) / 16 + 64) * 4));
    p->M = M;
    const void* dk = nullptr;
    ST_TRY(stage_in(p, k, M * p->D * (int64_t)p->esz(), where, &p->d_stage_k, &p->cap_stage_k, &dk));
    ST_TRY(nfftb_sort_nodes(p, dk));
    p->h_tile_start.resize((size_t)p->ntiles + 1);
    CUDA_TRY(p, cudaMemcpyAsync(p->h_tile_start.data(), p->d_tile_start, sizeof(int32_t) * (size_t)(p->ntiles + 1),
                                cudaMemcpyDeviceToHost, p->stream));
    CUDA_TRY(p, cudaStreamSynchronize(p->stream));
    {   // work items (load balancing of crowded tiles, e.g. the k-space centre of radial trajectories)
        const int64_t cap = p->D == 3 ? 4096 : 2048;
        std::vector<int32_t> items, strides;
        p->h_tile_items.assign((size_t)p->ntiles + 1, 0);
        for (int64_t t = 0; t < p->ntiles; t++) {
            p->h_tile_items[(size_t)t] = (int32_t)(items.size() / 3);
            const int64_t lo = p->h_tile_start[(size_t)t], hi = p->h_tile_start[(size_t)t + 1];
            if (hi == lo) continue;
            const int64_t parts = (hi - lo + cap - 1) / cap;
            for (int64_t q = 0; q < parts; q++) {
                items.push_back((int32_t)t);
                if (p->D == 2 && parts > 1) {            // strided split (see d_item_stride)
                    items.push_back((int32_t)(lo + q));
                    items.push_back((int32_t)hi);
                    strides.push_back((int32_t)parts);
                } else {
                    items.push_back((int32_t)(lo + (hi - lo) * q / parts));
                    items.push_back((int32_t)(lo + (hi - lo) * (q + 1) / parts));
                    strides.push_back(1);
                }
            }
        }
        p->h_tile_items[(size_t)p->ntiles] = (int32_t)(items.size() / 3);
        p->nitems = (int64_t)items.size() / 3;
        if (!p->d_tile_items) CUDA_TRY(p, cudaMalloc((void**)&p->d_tile_items, sizeof(int32_t) * (size_t)(p->ntiles + 1)));
        if (p->nitems > p->cap_items) {
            if (p->d_items) cudaFree(p->d_items);
            if (p->d_item_stride) cudaFree(p->d_item_stride);
            p->d_items = nullptr; p->d_item_stride = nullptr; p->cap_items = 0;
            CUDA_TRY(p, cudaMalloc((void**)&p->d_items, sizeof(int32_t) * 3 * (size_t)p->nitems));
            CUDA_TRY(p, cudaMalloc((void**)&p->d_item_stride, sizeof(int32_t) * (size_t)p->nitems));
            p->cap_items = p->nitems;
        }
        if (p->nitems > 0) {
            CUDA_TRY(p, cudaMemcpyAsync(p->d_items, items.data(), sizeof(int32_t) * items.size(), cudaMemcpyHostToDevice, p->stream));
            CUDA_TRY(p, cudaMemcpyAsync(p->d_item_stride, strides.data(), sizeof(int32_t) * strides.size(), cudaMemcpyHostToDevice, p->stream));
        }
        CUDA_TRY(p, cudaMemcpyAsync(p->d_tile_items, p->h_tile_items.data(), sizeof(int32_t) * (size_t)(p->ntiles + 1),
                                    cudaMemcpyHostToDevice, p->stream));
    }
    if (p->timing) cudaEventRecord(p->ev[1], p->stream);
    CUDA_TRY(p, cudaStreamSynchronize(p->stream));
    if (p->timing) { float ms = 0; cudaEventElapsedTime(&ms, p->ev[0], p->ev[1]); p->t[0] = ms * 1e-3; }
    p->have_nodes = true;
    // node sharding: (re)size and re-export the tile scratch for the new node set -- collective over the ranks
    if (p->shard_mode == NFFTB200_SHARD_NODES) ST_TRY(nfftb_comm_after_nodes(p));
    return NFFTB200_OK;
}

int nfftb200_get_permutation(nfftb200_plan* p, int64_t* perm, int64_t* tile_start)
{
    if (!p || !p->have_nodes) return nfftb_fail(p, NFFTB200_NO_NODES, "plan has no nodes");
    DeviceGuard guard(p->device);
    if (perm && p->M > 0) {
        std::vector<int32_t> h((size_t)p->M);
        CUDA_TRY(p, cudaMemcpyAsync(h.data(), p->d_perm, 4 * (size_t)p->M, cudaMemcpyDeviceToHost, p->stream));
        CUDA_TRY(p, cudaStreamSynchronize(p->stream));
        for (int64_t i = 0; i < p->M; i++) perm[i] = h[(size_t)i];
    }
    if (tile_start)
        for (int64_t i = 0; i <= p->ntiles; i++) tile_start[i] = p->h_tile_start[(size_t)i];
    return NFFTB200_OK;
}

int nfftb200_get_info(nfftb200_plan* p, int64_t* Nt, int64_t* block_size, int64_t* num_tiles, int64_t* lut_size,
                      double* sigma_eff, int64_t* M)
{
    if (!p) return NFFTB200_BAD_ARGUMENT;
    for (int d = 0; d < p->D; d++) {
        if (Nt) Nt[d] = p->Nt[d];
        if (block_size) block_size[d] = p->bs[d];
    }
    if (num_tiles) *num_tiles = p->ntiles;
    if (lut_size) *lut_size = p->lut_size;
    if (sigma_eff) *sigma_eff = p->sigma;
    if (M) *M = p->M;
    return NFFTB200_OK;
}

int nfftb200_get_table(nfftb200_plan* p, int which, double* out, int64_t cap, int64_t* n)
{
    if (!p) return NFFTB200_BAD_ARGUMENT;
    const std::vector<double>* h = which == 0 ? &p->h_hat_inv : (which == 1 ? &p->h_poly : (which == 2 ? &p->h_lin : nullptr));
    if (!h) return nfftb_fail(p, NFFTB200_BAD_ARGUMENT, "unknown table");
    if (n) *n = (int64_t)h->size();
    if (out) {
        const size_t c = std::min<size_t>(h->size(), (size_t)std::max<int64_t>(cap, 0));
        // values as the kernels see them (rounded to T)
        for (size_t i = 0; i < c; i++) out[i] = p->dtype == NFFTB200_F32 ? (double)(float)(*h)[i] : (*h)[i];
    }
    return NFFTB200_OK;
}

int nfftb200_exec_forward(nfftb200_plan* p, const void* f, void* fHat, int where)
{
    if (!p || !p->have_nodes) return nfftb_fail(p, NFFTB200_NO_NODES, "plan has no nodes");
    if (!f || !fHat) return nfftb_fail(p, NFFTB200_SIZE_MISMATCH, "Data is not consistent with NFFTPlan");
    DeviceGuard guard(p->device);
    const int64_t csz = 2 * (int64_t)p->esz();
    const void* df = nullptr;
    void* dh = nullptr;
    const bool async = where == NFFTB200_HOST_ASYNC;
    if (async) {
        ST_TRY(async_setup(p, 0, p->fsz * p->B * csz, p->M * p->B * csz));
        ST_TRY(async_begin(p, 0, f, p->fsz * p->B * csz));
        df = p->a_in[0].d; dh = p->a_out[0].d;
    } else {
        ST_TRY(stage_in(p, f, p->fsz * p->B * csz, where, &p->d_stage_f, &p->cap_stage_f, &df));
        ST_TRY(stage_out_buf(p, fHat, p->M * p->B * csz, where, &p->d_stage_h, &p->cap_stage_h, &dh));
    }
    if (p->shard_mode == NFFTB200_SHARD_NODES) {
        ST_TRY(nfftb_comm_exec_forward(p, df, dh));
    } else {
        rec(p, 0);
        ST_TRY(nfftb_deconvolve(p, df, p->d_grid, p->B));
        rec(p, 1);
        ST_TRY(run_fft(p, p->d_grid, -1, true));
        rec(p, 2);
        ST_TRY(nfftb_interp(p, p->d_grid, dh, p->B, 1, 0, p->ntiles));
        rec(p, 3);
        p->pending = 1;
    }
    if (async) return async_end(p, 0, fHat, p->M * p->B * csz);
    return stage_back(p, fHat, dh, p->M * p->B * csz, where);
}

int nfftb200_exec_adjoint(nfftb200_plan* p, const void* fHat, void* f, int where)
{
    if (!p || !p->have_nodes) return nfftb_fail(p, NFFTB200_NO_NODES, "plan has no nodes");
    if (!f || !fHat) return nfftb_fail(p, NFFTB200_SIZE_MISMATCH, "Data is not consistent with NFFTPlan");
    DeviceGuard guard(p->device);
    const int64_t csz = 2 * (int64_t)p->esz();
    const void* dh = nullptr;
    void* df = nullptr;
    const bool async = where == NFFTB200_HOST_ASYNC;
    if (async) {
        ST_TRY(async_setup(p, 1, p->M * p->B * csz, p->fsz * p->B * csz));
        ST_TRY(async_begin(p, 1, fHat, p->M * p->B * csz));
        dh = p->a_in[1].d; df = p->a_out[1].d;
    } else {
        ST_TRY(stage_in(p, fHat, p->M * p->B * csz, where, &p->d_stage_h, &p->cap_stage_h, &dh));
        ST_TRY(stage_out_buf(p, f, p->fsz * p->B * csz, where, &p->d_stage_f, &p->cap_stage_f, &df));
    }
    if (p->shard_mode == NFFTB200_SHARD_NODES) {
        ST_TRY(nfftb_comm_exec_adjoint(p, dh, df));
    } else {
        rec(p, 0);
        ST_TRY(nfftb_spread(p, dh, p->d_grid, p->B, 1, 0, p->ntiles));
        rec(p, 1);
        ST_TRY(run_fft(p, p->d_grid, +1, true));
        rec(p, 2);
        ST_TRY(nfftb_deconvolve_transpose(p, p->d_grid, df, p->B));
        rec(p, 3);
        p->pending = 2;
    }
    if (async) return async_end(p, 1, f, p->fsz * p->B * csz);
    return stage_back(p, f, df, p->fsz * p->B * csz, where);
}

int nfftb200_convolve(nfftb200_plan* p, const void* g, void* fHat, int is_complex, int where)
{
    if (!p || !p->have_nodes) return nfftb_fail(p, NFFTB200_NO_NODES, "plan has no nodes");
    if (!g || !fHat) return nfftb_fail(p, NFFTB200_SIZE_MISMATCH, "size(g) != Nt or size(fHat) != J");
    DeviceGuard guard(p->device);
    const int64_t csz = (is_complex ? 2 : 1) * (int64_t)p->esz();
    const void* dg = nullptr;
    void* dh = nullptr;
    ST_TRY(stage_in(p, g, p->gsz * csz, where, &p->d_stage_g, &p->cap_stage_g, &dg));
    ST_TRY(stage_out_buf(p, fHat, p->M * csz, where, &p->d_stage_h, &p->cap_stage_h, &dh));
    ST_TRY(nfftb_interp(p, dg, dh, 1, is_complex, 0, p->ntiles));
    return stage_back(p, fHat, dh, p->M * csz, where);
}

int nfftb200_convolve_transpose(nfftb200_plan* p, const void* fHat, void* g, int is_complex, int where)
{
    if (!p || !p->have_nodes) return nfftb_fail(p, NFFTB200_NO_NODES, "plan has no nodes");
    if (!g || !fHat) return nfftb_fail(p, NFFTB200_SIZE_MISMATCH, "size(g) != Nt or size(fHat) != J");
    DeviceGuard guard(p->device);
    const int64_t csz = (is_complex ? 2 : 1) * (int64_t)p->esz();
    const void* dh = nullptr;
    void* dg = nullptr;
    ST_TRY(stage_in(p, fHat, p->M * csz, where, &p->d_stage_h, &p->cap_stage_h, &dh));
    ST_TRY(stage_out_buf(p, g, p->gsz * csz, where, &p->d_stage_g, &p->cap_stage_g, &dg));
    ST_TRY(nfftb_spread(p, dh, dg, 1, is_complex, 0, p->ntiles));
    return stage_back(p, g, dg, p->gsz * csz, where);
}

int nfftb200_deconvolve(nfftb200_plan* p, const void* f, void* g, int where)
{
    if (!p) return NFFTB200_BAD_ARGUMENT;
    if (p->device < 0) return nfftb_fail(p, NFFTB200_CUDA_ERROR, "host-only plan (device < 0): no CPU fallback exists");
    if (!g || !f) return nfftb_fail(p, NFFTB200_SIZE_MISMATCH, "Data is not consistent with NFFTPlan");
    DeviceGuard guard(p->device);
    const int64_t csz = 2 * (int64_t)p->esz();
    const void* df = nullptr;
    void* dg = nullptr;
    ST_TRY(stage_in(p, f, p->fsz * csz, where, &p->d_stage_f, &p->cap_stage_f, &df));
    ST_TRY(stage_out_buf(p, g, p->gsz * csz, where, &p->d_stage_g, &p->cap_stage_g, &dg));
    ST_TRY(nfftb_deconvolve(p, df, dg, 1));
    return stage_back(p, g, dg, p->gsz * csz, where);
}

int nfftb200_deconvolve_transpose(nfftb200_plan* p, const void* g, void* f, int where)
{
    if (!p) return NFFTB200_BAD_ARGUMENT;
    if (p->device < 0) return nfftb_fail(p, NFFTB200_CUDA_ERROR, "host-only plan (device < 0): no CPU fallback exists");
    if (!g || !f) return nfftb_fail(p, NFFTB200_SIZE_MISMATCH, "Data is not consistent with NFFTPlan");
    DeviceGuard guard(p->device);
    const int64_t csz = 2 * (int64_t)p->esz();
    const void* dg = nullptr;
    void* df = nullptr;
    ST_TRY(stage_in(p, g, p->gsz * csz, where, &p->d_stage_g, &p->cap_stage_g, &dg));
    ST_TRY(stage_out_buf(p, f, p->fsz * csz, where, &p->d_stage_f, &p->cap_stage_f, &df));
    ST_TRY(nfftb_deconvolve_transpose(p, dg, df, 1));
    return stage_back(p, f, df, p->fsz * csz, where);
}

int nfftb200_get_grid(nfftb200_plan* p, void** device_ptr)
{
    if (!p || !device_ptr) return NFFTB200_BAD_ARGUMENT;
    *device_ptr = p->d_grid;
    return NFFTB200_OK;
}

int nfftb200_fft(nfftb200_plan* p, int direction)
{
    if (!p) return NFFTB200_BAD_ARGUMENT;
    if (p->device < 0) return nfftb_fail(p, NFFTB200_CUDA_ERROR, "host-only plan (device < 0): no CPU fallback exists");
    DeviceGuard guard(p->device);
    return run_fft(p, p->d_grid, direction);
}

int nfftb200_set_timing(nfftb200_plan* p, int enable)
{
    if (!p) return NFFTB200_BAD_ARGUMENT;
    p->timing = enable != 0;
    p->pending = 0;
    return NFFTB200_OK;
}

int nfftb200_get_timing(nfftb200_plan* p, double out[7])
{
    if (!p || !out) return NFFTB200_BAD_ARGUMENT;
    DeviceGuard guard(p->device);
    if (p->timing && p->pending) {
        CUDA_TRY(p, cudaEventSynchronize(p->ev[3]));
        float a = 0, b = 0, c = 0;
        cudaEventElapsedTime(&a, p->ev[0], p->ev[1]);
        cudaEventElapsedTime(&b, p->ev[1], p->ev[2]);
        cudaEventElapsedTime(&c, p->ev[2], p->ev[3]);
        if (p->pending == 1) { p->t[3] = a * 1e-3; p->t[2] = b * 1e-3; p->t[1] = c * 1e-3; }   // deconv, fft, conv
        else { p->t[4] = a * 1e-3; p->t[5] = b * 1e-3; p->t[6] = c * 1e-3; }                  // conv_adj, fft_adj, deconv_adj
        p->pending = 0;
    }
    for (int i = 0; i < 7; i++) out[i] = p->t[i];
    return NFFTB200_OK;
}

int nfftb200_get_kernel_times(nfftb200_plan* p, double out[4])
{
    if (!p || !out) return NFFTB200_BAD_ARGUMENT;
    DeviceGuard guard(p->device);
    if (p->timing && (p->pending_k & 1)) {
        CUDA_TRY(p, cudaEventSynchronize(p->evk[2]));
        float a = 0, b = 0;
        cudaEventElapsedTime(&a, p->evk[0], p->evk[1]);
        cudaEventElapsedTime(&b, p->evk[1], p->evk[2]);
        p->tk[2] = a * 1e-3; p->tk[0] = b * 1e-3;
        p->tk[3] = 0;
        if (p->have_gather_ev) { float c = 0; cudaEventElapsedTime(&c, p->evk[5], p->evk[2]); p->tk[3] = c * 1e-3; }
    }
    if (p->timing && (p->pending_k & 2)) {
        CUDA_TRY(p, cudaEventSynchronize(p->evk[4]));
        float a = 0;
        cudaEventElapsedTime(&a, p->evk[3], p->evk[4]);
        p->tk[1] = a * 1e-3;
    }
    p->pending_k = 0;
    for (int i = 0; i < 4; i++) out[i] = p->tk[i];
    return NFFTB200_OK;
}

int nfftb200_set_kernel_mode(nfftb200_plan* p, int mode)
{
    if (!p) return NFFTB200_BAD_ARGUMENT;
    p->kernel_mode = mode;
    return NFFTB200_OK;
}

int nfftb200_get_launch_count(nfftb200_plan* p, int64_t* n)
{
    if (!p || !n) return NFFTB200_BAD_ARGUMENT;
    *n = p->launches;
    return NFFTB200_OK;
}

int nfftb200_set_stream(nfftb200_plan* p, void* cuda_stream)
{
    if (!p) return NFFTB200_BAD_ARGUMENT;
    if (p->device < 0) return NFFTB200_OK;
    DeviceGuard guard(p->device);
    cudaStreamSynchronize(p->stream);
    if (p->own_stream) { cudaStreamDestroy(p->stream); p->own_stream = false; }
    p->stream = (cudaStream_t)cuda_stream;
    CUFFT_TRY(p, cufftSetStream(p->fft, p->stream));
    if (p->have_pruned) { CUFFT_TRY(p, cufftSetStream(p->fft_xy, p->stream)); CUFFT_TRY(p, cufftSetStream(p->fft_z, p->stream)); }
    if (p->have_fft_img) CUFFT_TRY(p, cufftSetStream(p->fft_img, p->stream));
    if (p->have_fft_d4) CUFFT_TRY(p, cufftSetStream(p->fft_d4, p->stream));
    nfftb_comm_set_stream(p);
    return NFFTB200_OK;
}

int nfftb200_sync(nfftb200_plan* p)
{
    if (!p) return NFFTB200_BAD_ARGUMENT;
    if (p->device < 0) return NFFTB200_OK;
    DeviceGuard guard(p->device);
    if (p->s_up) CUDA_TRY(p, cudaStreamSynchronize(p->s_up));
    CUDA_TRY(p, cudaStreamSynchronize(p->stream));
    if (p->s_down) CUDA_TRY(p, cudaStreamSynchronize(p->s_down));
    return NFFTB200_OK;
}

}  // extern "C"
