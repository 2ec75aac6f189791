// tile3d.cuh -- pieces shared by the tiled 3-D spread and interpolation kernels.
//
// "Row-per-lane" footprint layout: a node's (2m)^3 footprint is (2m)^2 x-rows of 2m cells.  Lane r of a
// warp owns row (t, yt) = (r / 2m, r % 2m) and walks it with 16-byte shared-memory accesses (one unit =
// two Float32 cells or one Float64 cell).  For Float32 a row that starts at an odd cell is widened to the
// aligned pair below it and the x weights are shifted by one with a zero in front, so every access is an
// aligned LDS.128/STS.128; row strides of 2*(S+2m) words keep the 8 lanes of a quarter-warp on distinct banks
// (verified for the default tiles).
//
// Window weights are evaluated once per node by ONE lane (lane-per-node, all 32 lanes busy), with the
// polynomial coefficients passed by value in the kernel parameter block so that the Horner FMAs read them
// straight from the constant bank (/root/reference/src/precomputation.jl:215-222 evalpoly == fma Horner).
#pragma once
#include "common.cuh"
#include "window.cuh"

template <typename T, int MT> struct PolyParam {
    T c[(2 * MT + 1) * 2 * MT];          // column-major (2m+1) x 2m, column = tap
};

template <typename T, int MT> inline PolyParam<T, MT> make_poly_param(const nfftb200_plan* p)
{
    PolyParam<T, MT> pp;
    const size_t n = (size_t)(2 * MT + 1) * 2 * MT;
    for (size_t i = 0; i < n; i++) pp.c[i] = i < p->h_poly.size() ? (T)p->h_poly[i] : (T)0;
    return pp;
}

template <typename T> struct Tile3 {
    using C = typename Cplx<T>::type;
    static constexpr int VPC = 16 / (int)sizeof(C);      // cells per 16-byte unit
};

// units per x-row and length of the (shifted, zero padded) x-weight vector
template <typename T, int MT> struct RowGeom {
    static constexpr int L = 2 * MT;
    static constexpr int VPC = Tile3<T>::VPC;
    static constexpr int NV = (VPC == 2) ? (L + 2) / 2 : L;
    static constexpr int NWX = NV * VPC;
    static constexpr int NROW = L * L;
    static constexpr int FULL_IT = NROW / 32;             // iterations with all 32 lanes on full rows
    static constexpr int REM = NROW % 32;                 // leftover rows
    static constexpr bool SPLIT = (REM > 0) && (REM * NV <= 32);   // leftover rows split into units
};

// all 2m tap weights of one dimension for the calling lane's node (compile-time unrolled)
template <typename T, int MT>
__device__ __forceinline__ void eval_taps(const WinDev<T>& win, const PolyParam<T, MT>& pp, T kscale, int c,
                                          T (&w)[2 * MT])
{
    constexpr int L = 2 * MT, deg = L + 1;
    const int off = c - MT + 1;
    const T d0 = sub_rn(kscale, (T)off);
    if (win.mode == NFFTB200_POLYNOMIAL) {
        const T x = sub_rn(add_rn(sub_rn(d0, (T)MT), (T)1), (T)0.5);
#pragma unroll
        for (int l = 0; l < L; l++) {
            T acc = pp.c[l * deg + deg - 1];
#pragma unroll
            for (int r = deg - 2; r >= 0; r--) acc = tfma(acc, x, pp.c[l * deg + r]);
            w[l] = acc;
        }
    } else if (win.mode == NFFTB200_LINEAR) {
        const T idx = mul_rn(d0, (T)win.lin_scale);
        const int ii = (int)idx;
        const T alpha = sub_rn(idx, (T)ii);
#pragma unroll
        for (int l = 0; l < L; l++) {
            int a1 = ii - l * win.lin_scale;
            int a2 = a1 + 1;
            a1 = a1 < 0 ? -a1 : a1;
            a2 = a2 < 0 ? -a2 : a2;
            const T v1 = win.lin[a1], v2 = win.lin[a2];
            w[l] = add_rn(v1, mul_rn(alpha, sub_rn(v2, v1)));
        }
    } else {
#pragma unroll
        for (int l = 0; l < L; l++) w[l] = kb_exact<T>(sub_rn(d0, (T)l), MT, win.b);   // other windows: generic kernels (nfftb_tiled_ok)
    }
}

__device__ __forceinline__ int wrapi(int v, int n)
{
    v %= n;
    return v < 0 ? v + n : v;
}
// periodic wrap of v in [-n, 2n) without a division (fast == true), else the general modulo
__device__ __forceinline__ int wrapc(int v, int n, bool fast)
{
    if (fast) {
        v += (v < 0) ? n : 0;
        v -= (v >= n) ? n : 0;
        return v;
    }
    return wrapi(v, n);
}
// exact row / d for row * d < 2^32, with inv = ceil(2^32 / d)
__device__ __forceinline__ unsigned fastdiv(unsigned row, unsigned inv) { return __umulhi(row, inv); }
__device__ __forceinline__ unsigned fastdiv_inv(unsigned d) { return (unsigned)((0x100000000ull + d - 1) / d); }

// asynchronous global -> shared copy of one complex cell (LDGSTS; no register staging, every row of the
// tile is in flight at once)
__device__ __forceinline__ void cp_async_cell(float2* dst, const float2* src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
// two Float32 cells: both addresses 16-byte aligned
__device__ __forceinline__ void cp_async_pair(float2* dst, const float2* src)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_cell(double2* dst, const double2* src)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}

// bulk asynchronous copy shared -> global through the TMA engine (SASS: UBLKCP); size and both addresses must be
// multiples of 16 bytes.  Completion is tracked with bulk async-groups.
__device__ __forceinline__ void bulk_store_s2g(void* gdst, const void* ssrc, unsigned bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst),
                 "r"((unsigned)__cvta_generic_to_shared(ssrc)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit_wait_read()
{
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
// make generic-proxy shared-memory writes visible to the async proxy (TMA) -- call after the CTA barrier
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- TMA tensor-map tile load (cp.async.bulk.tensor, SASS: UTMALDG) + mbarrier completion -------------------
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@!p bra WAIT_LOOP;\n"
        "}\n" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(parity) : "memory");
}
// 4-D box (x, y, z, batch) of the grid -> dense shared-memory tile
__device__ __forceinline__ void tma_load_4d(void* sdst, const void* tmap, unsigned long long* bar, int c0, int c1, int c2, int c3)
{
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                 ::"r"((unsigned)__cvta_generic_to_shared(sdst)), "l"(tmap), "r"((unsigned)__cvta_generic_to_shared(bar)),
                   "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}

// 16-byte shared-memory unit <-> registers
template <typename T> struct Unit;
template <> struct Unit<float> {
    float4 v;
    __device__ __forceinline__ void load(const float2* p) { v = *reinterpret_cast<const float4*>(p); }
    __device__ __forceinline__ void store(float2* p) const { *reinterpret_cast<float4*>(p) = v; }
    // cell k of the unit += w[k] * (a, b)
    __device__ __forceinline__ void axpy(const float* w, float a, float b)
    {
        v.x = fmaf(w[0], a, v.x); v.y = fmaf(w[0], b, v.y);
        v.z = fmaf(w[1], a, v.z); v.w = fmaf(w[1], b, v.w);
    }
    // (a, b) += sum_k w[k] * cell k
    __device__ __forceinline__ void dot(const float* w, float& a, float& b) const
    {
        a = fmaf(w[0], v.x, a); b = fmaf(w[0], v.y, b);
        a = fmaf(w[1], v.z, a); b = fmaf(w[1], v.w, b);
    }
};
template <> struct Unit<double> {
    double2 v;
    __device__ __forceinline__ void load(const double2* p) { v = *p; }
    __device__ __forceinline__ void store(double2* p) const { *p = v; }
    __device__ __forceinline__ void axpy(const double* w, double a, double b)
    {
        v.x = fma(w[0], a, v.x); v.y = fma(w[0], b, v.y);
    }
    __device__ __forceinline__ void dot(const double* w, double& a, double& b) const
    {
        a = fma(w[0], v.x, a); b = fma(w[0], v.y, b);
    }
};
