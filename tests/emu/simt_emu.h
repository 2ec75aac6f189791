// simt_emu.h -- TEST INFRASTRUCTURE: runs a CUDA kernel body on the host, one OS thread per CUDA thread, one CTA at a
// time, so that the indexing / synchronisation logic of a kernel can be checked in the CPU-only container.  It is not
// a product path (nothing in nfft.jl_b200/ includes it) and says nothing about performance.
//
// Include AFTER <cuda_runtime.h> (vector types) and BEFORE the kernel header.  Provides: threadIdx/blockIdx/
// blockDim/gridDim (thread-local), __syncthreads/__syncwarp, the warp collectives the kernels use
// (__match_any_sync, __ballot_sync, __shfl_up_sync, __shfl_xor_sync), the rounding intrinsics and a launch helper.
#pragma once
#include <pthread.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <functional>
#include <thread>
#include <vector>

#ifdef __CUDACC__
#error "simt_emu.h is for host-only builds"
#endif

#define __launch_bounds__(...)
#define NFFTB_EMU 1
#ifndef __noinline__
#define __noinline__ __attribute__((noinline))
#endif
#include <cstdio>
#include <cstdlib>
// kernels mark the pointers of their vector accesses with this; the device build compiles it away
#define NFFTB_EMU_ALIGNED(ptr, bytes)                                                                              \
    do {                                                                                                           \
        if (((uintptr_t)(ptr)) % (bytes) != 0) {                                                                   \
            fprintf(stderr, "%s:%d: %s is not %d-byte aligned\n", __FILE__, __LINE__, #ptr, (int)(bytes));          \
            abort();                                                                                               \
        }                                                                                                          \
    } while (0)
using std::max;
using std::min;

namespace emu {
struct Dim3 { unsigned x = 1, y = 1, z = 1; };
inline thread_local Dim3 t_threadIdx, t_blockIdx;
inline Dim3 g_blockDim, g_gridDim;

struct WarpState {
    pthread_barrier_t bar;
    unsigned long long vals[32];
};
inline pthread_barrier_t g_cta_bar;
inline std::vector<WarpState> g_warps;
inline WarpState& my_warp() { return g_warps[t_threadIdx.x >> 5]; }
inline int my_lane() { return (int)(t_threadIdx.x & 31); }
}  // namespace emu

#define threadIdx emu::t_threadIdx
#define blockIdx emu::t_blockIdx
#define blockDim emu::g_blockDim
#define gridDim emu::g_gridDim

// dynamic shared memory of the emulated CTA (the kernels declare `extern __shared__ ... smem_raw[]`)
alignas(128) unsigned char smem_raw[232 * 1024];

inline void __syncthreads() { pthread_barrier_wait(&emu::g_cta_bar); }
inline void __syncwarp(unsigned = 0xffffffffu) { pthread_barrier_wait(&emu::my_warp().bar); }

// full-mask warp collectives (all 32 lanes must call)
inline unsigned __match_any_sync(unsigned, unsigned v)
{
    auto& w = emu::my_warp();
    w.vals[emu::my_lane()] = v;
    pthread_barrier_wait(&w.bar);
    unsigned m = 0;
    for (int l = 0; l < 32; l++) m |= (w.vals[l] == v) ? (1u << l) : 0u;
    pthread_barrier_wait(&w.bar);
    return m;
}
inline unsigned __ballot_sync(unsigned, int pred)
{
    auto& w = emu::my_warp();
    w.vals[emu::my_lane()] = pred ? 1 : 0;
    pthread_barrier_wait(&w.bar);
    unsigned m = 0;
    for (int l = 0; l < 32; l++) m |= w.vals[l] ? (1u << l) : 0u;
    pthread_barrier_wait(&w.bar);
    return m;
}
inline int __shfl_up_sync(unsigned, int v, int delta)
{
    auto& w = emu::my_warp();
    const int lane = emu::my_lane();
    w.vals[lane] = (unsigned long long)(long long)v;
    pthread_barrier_wait(&w.bar);
    const int r = lane >= delta ? (int)(long long)w.vals[lane - delta] : v;
    pthread_barrier_wait(&w.bar);
    return r;
}

template <typename V> inline V __shfl_xor_sync(unsigned, V v, int lane_mask)
{
    static_assert(sizeof(V) <= 8, "shuffle emulation moves at most 8 bytes");
    auto& w = emu::my_warp();
    const int lane = emu::my_lane();
    unsigned long long bits = 0;
    std::memcpy(&bits, &v, sizeof(V));
    w.vals[lane] = bits;
    pthread_barrier_wait(&w.bar);
    const unsigned long long got = w.vals[lane ^ lane_mask];
    pthread_barrier_wait(&w.bar);
    V r;
    std::memcpy(&r, &got, sizeof(V));
    return r;
}

inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline unsigned __umulhi(unsigned a, unsigned b) { return (unsigned)(((unsigned long long)a * b) >> 32); }
// the host build uses -ffp-contract=off, so plain operators are the _rn forms
inline float __fmul_rn(float a, float b) { return a * b; }
inline float __fadd_rn(float a, float b) { return a + b; }
inline double __dmul_rn(double a, double b) { return a * b; }
inline double __dadd_rn(double a, double b) { return a + b; }
inline float cyl_bessel_i0f(float x) { return std::cyl_bessel_i(0.0f, x); }
inline double cyl_bessel_i0(double x) { return std::cyl_bessel_i(0.0, x); }

namespace emu {
// run body() as a grid of CTAs with `threads` threads each (threads % 32 == 0), CTAs one after the other
inline void launch(Dim3 grid, unsigned threads, const std::function<void()>& body)
{
    g_blockDim = Dim3{threads, 1, 1};
    g_gridDim = grid;
    const unsigned nw = threads / 32;
    g_warps = std::vector<WarpState>(nw);
    for (auto& w : g_warps) pthread_barrier_init(&w.bar, nullptr, 32);
    pthread_barrier_init(&g_cta_bar, nullptr, threads);
    for (unsigned by = 0; by < grid.y; by++)
        for (unsigned bx = 0; bx < grid.x; bx++) {
            std::vector<std::thread> th;
            th.reserve(threads);
            for (unsigned t = 0; t < threads; t++)
                th.emplace_back([&, t, bx, by] {
                    t_threadIdx = Dim3{t, 0, 0};
                    t_blockIdx = Dim3{bx, by, 0};
                    body();
                });
            for (auto& x : th) x.join();
        }
    for (auto& w : g_warps) pthread_barrier_destroy(&w.bar);
    pthread_barrier_destroy(&g_cta_bar);
}
}  // namespace emu
