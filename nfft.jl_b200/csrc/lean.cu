// lean.cu -- launchers of the kernel_mode-8 kernels (spread_lean.cuh / interp_lean.cuh): Float32, 3-D, m = 2 or 3,
// tiles of at most 16 cells.  Own translation unit so that the hot kernels rebuild in seconds.
#include "common.cuh"
#include "window.cuh"
#include "tile3d.cuh"
#include "spread_lean.cuh"

namespace {

template <int MT, int W> bool lean_geom_ok(const GeomDev& geo)
{
    for (int d = 0; d < 3; d++) {
        if (geo.bs[d] > 2 * W || geo.bs[d] + 2 * MT < W) return false;
    }
    return true;
}

template <int MT, int W>
int spread_lean(nfftb200_plan* p, const void* fhat, void* g, void* scratch_override, int B, int t_lo, int t_hi)
{
    using T = float;
    using C = float2;
    using SLy = LeanSpreadLayout<MT, W>;
    GeomDev geo = make_geom<T>(p);
    BinGeom bg;
    if (!lean_geom_ok<MT, W>(geo) || !SLy::make(geo.bs, bg)) return -1;
    const size_t smem = SLy::bytes(bg);
    if (smem > 227 * 1024) return -1;
    for (int d = 0; d < 3; d++) {
        const int last = geo.Nt[d] - (geo.nb[d] - 1) * geo.bs[d];
        if (geo.bs[d] < MT || last < MT) return -1;                   // halos would reach past the neighbour
        if (d == 0 && ((geo.bs[0] & 1) || (last & 1))) return -1;     // the gather works on x cell pairs
    }
    if (nfftb_ensure_bins(p, W, LeanGeom<MT, W>::G) != NFFTB200_OK) return -1;
    const size_t PN = (size_t)(geo.bs[0] + 2 * MT) * (geo.bs[1] + 2 * MT) * (geo.bs[2] + 2 * MT);
    const cudaStream_t st = p->stream;
    const int item_lo = p->h_tile_items[(size_t)t_lo], item_hi = p->h_tile_items[(size_t)t_hi];
    if (item_hi == item_lo) return scratch_override ? NFFTB200_OK : -1;
    void* scratch = scratch_override;
    if (!scratch) {
        const int64_t need = (int64_t)(sizeof(C) * PN * (size_t)(item_hi - item_lo) * B);
        if (need > p->cap_tilebuf) {
            if (p->d_tilebuf) cudaFree(p->d_tilebuf);
            p->d_tilebuf = nullptr; p->cap_tilebuf = 0;
            if (cudaMalloc(&p->d_tilebuf, (size_t)need) != cudaSuccess) { cudaGetLastError(); return -1; }
            p->cap_tilebuf = need;
        }
        scratch = p->d_tilebuf;
    }
    p->have_gather_ev = false;
    // fused "last arriver gathers" form (opt-in, kernel_mode 11): the whole grid in one launch, 16-cell-thick tile layers.
    // Measured on B200 (profiles/r02_fused_gather.txt): C2 1073 us against 528 + 108 us for spread + separate gather, and
    // MORE DRAM traffic (866 MB vs 834 MB): the scratch tiles a CTA gathers were written 30-70 us earlier by other SMs
    // but the reads still go to DRAM, and a 256-thread CTA cannot cover that latency.  Kept for the record.
    const bool fuse = !scratch_override && p->kernel_mode == 11 && t_lo == 0 && t_hi == p->ntiles && geo.bs[2] == 16 &&
                      geo.Nt[2] % 16 == 0 && 16 >= 2 * MT;
    if (fuse) {
        const int64_t need = (int64_t)sizeof(int32_t) * p->ntiles * B;
        if (need > p->cap_ready) {
            if (p->d_ready) cudaFree(p->d_ready);
            p->d_ready = nullptr; p->cap_ready = 0;
            CUDA_TRY(p, cudaMalloc((void**)&p->d_ready, (size_t)need));
            p->cap_ready = need;
        }
        if (p->timing) { cudaEventRecord(p->evk[0], st); cudaEventRecord(p->evk[1], st); }
        CUDA_TRY(p, cudaMemsetAsync(p->d_ready, 0, (size_t)need, st));
        LeanFuse fz{(C*)g, p->d_tile_items, item_hi, (int)p->ntiles, p->d_ready, p->d_expect};
        auto kf = k_spread_lean<MT, W, true>;
        CUDA_TRY(p, cudaFuncSetAttribute(kf, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        cudaFuncSetAttribute(kf, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        kf<<<dim3(item_hi - item_lo, B), NFFTB_BIN_WARPS * 32, smem, st>>>((const C*)fhat, (C*)scratch, (const T*)p->d_xs2, p->d_perm2,
                                                                          p->d_bin_start, p->d_items, item_lo, p->M, geo, make_win<T>(p),
                                                                          make_poly_param<T, MT>(p), bg, fz);
        k_zero_empty_blocks<MT><<<dim3((unsigned)p->ntiles, B), 256, 0, st>>>((C*)g, p->d_expect, geo);
        p->launches += 3;
        if (p->timing) { cudaEventRecord(p->evk[2], st); p->pending_k |= 1; }
        CUDA_TRY(p, cudaGetLastError());
        return NFFTB200_OK;
    }
    if (p->timing) { cudaEventRecord(p->evk[0], st); cudaEventRecord(p->evk[1], st); }
    auto kern = k_spread_lean<MT, W, false>;
    CUDA_TRY(p, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    kern<<<dim3(item_hi - item_lo, B), NFFTB_BIN_WARPS * 32, smem, st>>>((const C*)fhat, (C*)scratch, (const T*)p->d_xs2, p->d_perm2,
                                                                        p->d_bin_start, p->d_items, item_lo, p->M, geo, make_win<T>(p),
                                                                        make_poly_param<T, MT>(p), bg, LeanFuse{});
    p->launches++;
    CUDA_TRY(p, cudaGetLastError());
    if (scratch_override) {                                           // node sharding: the peer gather follows separately
        if (p->timing) { cudaEventRecord(p->evk[2], st); p->pending_k |= 1; }
        return NFFTB200_OK;
    }
    if (p->timing) { cudaEventRecord(p->evk[5], st); p->have_gather_ev = true; }
    ST_TRY(nfftb_gather_scratch(p, scratch, g, B, t_lo, t_hi, item_lo, item_hi));
    if (p->timing) { cudaEventRecord(p->evk[2], st); p->pending_k |= 1; }
    return NFFTB200_OK;
}

template <int MT, int W>
int interp_lean(nfftb200_plan* p, const void* g, void* fhat, int B, int t_lo, int t_hi, const SlabTab* slabs)
{
    using T = float;
    using C = float2;
    using ILy = LeanInterpLayout<MT, W>;
    GeomDev geo = make_geom<T>(p);
    BinGeom bg;
    if (!lean_geom_ok<MT, W>(geo) || !ILy::make(geo.bs, bg)) return -1;
    const size_t smem = ILy::bytes(bg);
    if (smem > 227 * 1024 || geo.bs[0] + 2 * MT > 64) return -1;
    if (nfftb_ensure_bins(p, W, LeanGeom<MT, W>::G) != NFFTB200_OK) return -1;
    const int item_lo = p->h_tile_items[(size_t)t_lo], item_hi = p->h_tile_items[(size_t)t_hi];
    if (item_hi == item_lo) return NFFTB200_OK;
    if (slabs) {
        auto kern = k_interp_lean<MT, W, true>;
        CUDA_TRY(p, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        kern<<<dim3(item_hi - item_lo, 1), NFFTB_BIN_WARPS * 32, smem, p->stream>>>(nullptr, (C*)fhat, (const T*)p->d_xs2, p->d_perm2,
                                                                                   p->d_bin_start, p->d_items, item_lo, p->M, geo,
                                                                                   make_win<T>(p), make_poly_param<T, MT>(p), bg, *slabs);
    } else {
        auto kern = k_interp_lean<MT, W, false>;
        CUDA_TRY(p, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        kern<<<dim3(item_hi - item_lo, B), NFFTB_BIN_WARPS * 32, smem, p->stream>>>((const C*)g, (C*)fhat, (const T*)p->d_xs2, p->d_perm2,
                                                                                   p->d_bin_start, p->d_items, item_lo, p->M, geo,
                                                                                   make_win<T>(p), make_poly_param<T, MT>(p), bg, SlabTab{});
    }
    p->launches++;
    CUDA_TRY(p, cudaGetLastError());
    return NFFTB200_OK;
}

}  // namespace

int nfftb_spread_lean(nfftb200_plan* p, const void* fhat, void* g, void* scratch_override, int B, int t_lo, int t_hi)
{
    if (p->dtype != NFFTB200_F32 || p->D != 3) return -1;
    switch (p->m) {
        case 2: return spread_lean<2, 8>(p, fhat, g, scratch_override, B, t_lo, t_hi);
        case 3: return spread_lean<3, 8>(p, fhat, g, scratch_override, B, t_lo, t_hi);
        default: return -1;
    }
}

int nfftb_interp_lean(nfftb200_plan* p, const void* g, void* fhat, int B, int t_lo, int t_hi, const SlabTab* slabs)
{
    if (p->dtype != NFFTB200_F32 || p->D != 3) return -1;
    switch (p->m) {
        case 2: return interp_lean<2, 8>(p, g, fhat, B, t_lo, t_hi, slabs);
        case 3: return interp_lean<3, 8>(p, g, fhat, B, t_lo, t_hi, slabs);
        default: return -1;
    }
}
