// window.cuh -- device-side node -> (cell, tap weights) evaluation, bit-faithful to the
// reference's *blocked* formulas on shifted nodes:
//   kscale/off/idx      /root/reference/src/precomputation.jl:536-547  (_precomputeIdxInBlock)
//   POLYNOMIAL Horner   /root/reference/src/precomputation.jl:215-222  (evalpoly == fma Horner)
//   LINEAR LUT lerp     /root/reference/src/precomputation.jl:179-201  (shiftedWindowEntries)
//   FULL exact window   /root/reference/src/windowFunctions.jl:21-34   (window_kaiser_bessel)
// Rounding-critical expressions use the _rn intrinsics so that nvcc's default FMA contraction
// cannot change them.
#pragma once
#include "common.cuh"

__device__ __forceinline__ float mul_rn(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ double mul_rn(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ float add_rn(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ double add_rn(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ float sub_rn(float a, float b) { return __fadd_rn(a, -b); }
__device__ __forceinline__ double sub_rn(double a, double b) { return __dadd_rn(a, -b); }
__device__ __forceinline__ float tfma(float a, float b, float c) { return fmaf(a, b, c); }
__device__ __forceinline__ double tfma(double a, double b, double c) { return fma(a, b, c); }
__device__ __forceinline__ float tsqrt(float a) { return sqrtf(a); }
__device__ __forceinline__ double tsqrt(double a) { return sqrt(a); }
__device__ __forceinline__ float tsinh(float a) { return sinhf(a); }
__device__ __forceinline__ double tsinh(double a) { return sinh(a); }
template <typename T> __device__ __forceinline__ T teps();
template <> __device__ __forceinline__ float teps<float>() { return 1.1920928955078125e-07f; }
template <> __device__ __forceinline__ double teps<double>() { return 2.220446049250313e-16; }

// shiftNodes! for one coordinate (src/utils.jl:32-44)
template <typename T> __device__ __forceinline__ T shift_node(T k)
{
    if (k < (T)0) k = add_rn(k, (T)1);
    if (k == (T)1) k = sub_rn(k, teps<T>());
    return k;
}

// kscale = k * Nt rounded in T; returns c = unsafe_trunc(Int, kscale)
template <typename T> __device__ __forceinline__ int node_cell(T ks, int Nt, T& kscale)
{
    kscale = mul_rn(ks, (T)Nt);
    return (int)kscale;
}

template <typename T> __device__ __forceinline__ T kb_exact(T x, int m, T b)
{
    const T mm = (T)m;
    const T ax = fabs(x);
    if (ax < mm) {
        T arg = tsqrt(mm * mm - x * x);
        return tsinh(b * arg) / (arg * (T)3.141592653589793238462643383279502884);
    } else if (ax > mm) {
        return (T)0;
    }
    return b / (T)3.141592653589793238462643383279502884;
}

__device__ __forceinline__ float texp(float a) { return expf(a); }
__device__ __forceinline__ double texp(double a) { return exp(a); }
__device__ __forceinline__ float tcosh(float a) { return coshf(a); }
__device__ __forceinline__ double tcosh(double a) { return cosh(a); }
__device__ __forceinline__ float ti0(float a) { return cyl_bessel_i0f(a); }
__device__ __forceinline__ double ti0(double a) { return cyl_bessel_i0(a); }

// exact window in grid units for the FULL mode, any window of getWindow
// (/root/reference/src/windowFunctions.jl:41-116); every window but kaiser_bessel is 0 for |x| >= m
template <typename T> __device__ __noinline__ T win_other(T x, int m, const WinDev<T>& w)
{
    const T mm = (T)m;
    const T pi = (T)3.141592653589793238462643383279502884;
    if (!(fabs(x) < mm)) return (T)0;
    if (w.window == NFFTB200_GAUSS) {
        const T b = mm / pi;
        return (T)1 / tsqrt(pi * b) * texp(-(x * x) / b);
    }
    if (w.window == NFFTB200_KAISER_BESSEL_REV) {
        const T q = x / mm;
        return (T)0.5 / mm * ti0(mm * w.b * tsqrt((T)1 - q * q));
    }
    if (w.window == NFFTB200_EXP_SQRT) {
        const T q = x / mm;
        return texp(w.beta * (tsqrt((T)1 - q * q) - (T)1));
    }
    if (w.window == NFFTB200_COSH_TYPE) {
        const T q = x / mm;
        const T alpha = tsqrt((T)1 - q * q);
        return (T)1 / (tcosh(w.beta) - (T)1) * (tcosh(w.beta * alpha) - (T)1) / alpha;
    }
    // spline: cardinal B-spline of order 2m at x + m, Cox-de Boor bottom-up
    const T k = x + mm;
    const int order = 2 * m;
    T a[2 * NFFTB_MAX_M];
    for (int j = 0; j < 2 * NFFTB_MAX_M; j++) {
        const T kj = k - (T)j;
        a[j] = (j < order && kj >= (T)0 && kj < (T)1) ? (T)1 : (T)0;
    }
    for (int n = 2; n <= order; n++)
        for (int j = 0; j + n <= order; j++) {
            const T kj = k - (T)j;
            a[j] = kj / (T)(n - 1) * a[j] + ((T)n - kj) / (T)(n - 1) * a[j + 1];
        }
    return a[0];
}

template <typename T> __device__ __forceinline__ T win_exact(T x, int m, const WinDev<T>& w)
{
    if (w.window == NFFTB200_KAISER_BESSEL) return kb_exact<T>(x, m, w.b);
    return win_other<T>(x, m, w);
}

// weights of the 2m taps of one dimension; tap l sits at cell (c - m + 1 + l)
template <typename T>
__device__ __forceinline__ void node_taps(const WinDev<T>& w, T kscale, int c, T* __restrict__ out)
{
    const int m = w.m;
    const int L = 2 * m;
    const int off = c - m + 1;
    const T d0 = sub_rn(kscale, (T)off);                  // frac + m - 1, exact
    if (w.mode == NFFTB200_POLYNOMIAL) {
        const T x = sub_rn(add_rn(sub_rn(d0, (T)m), (T)1), (T)0.5);   // frac - 1/2
        const int deg = L + 1;
        for (int l = 0; l < L; l++) {
            const T* cf = w.poly + l * deg;
            T acc = cf[deg - 1];
            for (int r = deg - 2; r >= 0; r--) acc = tfma(acc, x, cf[r]);
            out[l] = acc;
        }
    } else if (w.mode == NFFTB200_LINEAR) {
        const T idx = mul_rn(d0, (T)w.lin_scale);
        const int ii = (int)idx;                          // idx >= 0 -> floor
        const T alpha = sub_rn(idx, (T)ii);
        for (int l = 0; l < L; l++) {
            int a1 = ii - l * w.lin_scale;
            int a2 = a1 + 1;
            a1 = a1 < 0 ? -a1 : a1;
            a2 = a2 < 0 ? -a2 : a2;
            const T v1 = w.lin[a1], v2 = w.lin[a2];
            out[l] = add_rn(v1, mul_rn(alpha, sub_rn(v2, v1)));
        }
    } else {
        for (int l = 0; l < L; l++) out[l] = win_exact<T>(sub_rn(d0, (T)l), m, w);
    }
}

// single tap (used when a thread owns one (node, dim, tap) triple)
template <typename T>
__device__ __forceinline__ T node_tap(const WinDev<T>& w, T kscale, int c, int l)
{
    const int m = w.m;
    const int off = c - m + 1;
    const T d0 = sub_rn(kscale, (T)off);
    if (w.mode == NFFTB200_POLYNOMIAL) {
        const T x = sub_rn(add_rn(sub_rn(d0, (T)m), (T)1), (T)0.5);
        const int deg = 2 * m + 1;
        const T* cf = w.poly + l * deg;
        T acc = cf[deg - 1];
        for (int r = deg - 2; r >= 0; r--) acc = tfma(acc, x, cf[r]);
        return acc;
    } else if (w.mode == NFFTB200_LINEAR) {
        const T idx = mul_rn(d0, (T)w.lin_scale);
        const int ii = (int)idx;
        const T alpha = sub_rn(idx, (T)ii);
        int a1 = ii - l * w.lin_scale;
        int a2 = a1 + 1;
        a1 = a1 < 0 ? -a1 : a1;
        a2 = a2 < 0 ? -a2 : a2;
        const T v1 = w.lin[a1], v2 = w.lin[a2];
        return add_rn(v1, mul_rn(alpha, sub_rn(v2, v1)));
    }
    return win_exact<T>(sub_rn(d0, (T)l), m, w);
}
