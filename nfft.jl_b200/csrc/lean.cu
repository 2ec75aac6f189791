// lean.cu -- launchers of the kernel_mode-8 kernels (spread_lean.cuh / interp_lean.cuh): Float32, 3-D, m = 2 or 3,
// tiles of at most 16 cells.  Own translation unit so that the hot kernels rebuild in seconds.
#include <cuda.h>

#include <cstring>
#include <vector>

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
    const int lut_floats = (p->precompute == NFFTB200_LINEAR && p->lut_size + 2 <= 4096) ? (int)p->lut_size + 2 : 0;
    const size_t smem = SLy::bytes(bg, lut_floats);
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
                                                                          make_poly_param<T, MT>(p), bg, fz, lut_floats);
        k_zero_empty_blocks<MT><<<dim3((unsigned)p->ntiles, B), 256, 0, st>>>((C*)g, p->d_expect, geo);
        p->launches += 3;
        if (p->timing) { cudaEventRecord(p->evk[2], st); p->pending_k |= 1; }
        CUDA_TRY(p, cudaGetLastError());
        return NFFTB200_OK;
    }
    // cluster-pair experiment (kernel_mode 13): needs one work item per tile, every tile non-empty, whole tiles, an even
    // number of tiles along x, the whole grid in one launch
    bool pair = !scratch_override && p->kernel_mode == 13 && t_lo == 0 && t_hi == p->ntiles && (geo.nb[0] % 2 == 0) &&
                geo.Nt[0] % (2 * geo.bs[0]) == 0 && p->nitems == p->ntiles && geo.bs[2] == 16 && geo.Nt[2] % 16 == 0;
    if (pair) {
        const int npairs = (int)(p->ntiles / 2);
        if (npairs + 1 > p->cap_pair_items) {
            if (p->d_pair_items) cudaFree(p->d_pair_items);
            p->d_pair_items = nullptr; p->cap_pair_items = 0;
            CUDA_TRY(p, cudaMalloc((void**)&p->d_pair_items, sizeof(int32_t) * (size_t)(npairs + 1)));
            std::vector<int32_t> iota((size_t)npairs + 1);
            for (int i = 0; i <= npairs; i++) iota[(size_t)i] = i;
            CUDA_TRY(p, cudaMemcpy(p->d_pair_items, iota.data(), sizeof(int32_t) * iota.size(), cudaMemcpyHostToDevice));
            p->cap_pair_items = npairs + 1;
        }
        if (p->timing) { cudaEventRecord(p->evk[0], st); cudaEventRecord(p->evk[1], st); }
        auto kp = k_spread_lean<MT, W, false, true>;
        CUDA_TRY(p, cudaFuncSetAttribute(kp, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        cudaFuncSetAttribute(kp, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(item_hi - item_lo, B);
        cfg.blockDim = dim3(NFFTB_BIN_WARPS * 32);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        CUDA_TRY(p, cudaLaunchKernelEx(&cfg, kp, (const C*)fhat, (C*)scratch, (const T*)p->d_xs2, (const int32_t*)p->d_perm2,
                                       (const int32_t*)p->d_bin_start, (const int32_t*)p->d_items, item_lo, (long long)p->M, geo, make_win<T>(p),
                                       make_poly_param<T, MT>(p), bg, LeanFuse{}, lut_floats));
        p->launches++;
        if (p->timing) { cudaEventRecord(p->evk[5], st); p->have_gather_ev = true; }
        GeomDev gp = geo;
        gp.bs[0] = 2 * geo.bs[0]; gp.nb[0] = geo.nb[0] / 2;
        gp.inv_bs[0] = (unsigned)((0x100000000ull + (unsigned long long)gp.bs[0] - 1) / (unsigned long long)gp.bs[0]);
        ST_TRY(nfftb_gather_scratch(p, scratch, g, B, 0, npairs, 0, npairs, &gp, p->d_pair_items));
        if (p->timing) { cudaEventRecord(p->evk[2], st); p->pending_k |= 1; }
        return NFFTB200_OK;
    }
    if (p->timing) { cudaEventRecord(p->evk[0], st); cudaEventRecord(p->evk[1], st); }
    auto kern = k_spread_lean<MT, W, false>;
    CUDA_TRY(p, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    kern<<<dim3(item_hi - item_lo, B), NFFTB_BIN_WARPS * 32, smem, st>>>((const C*)fhat, (C*)scratch, (const T*)p->d_xs2, p->d_perm2,
                                                                        p->d_bin_start, p->d_items, item_lo, p->M, geo, make_win<T>(p),
                                                                        make_poly_param<T, MT>(p), bg, LeanFuse{}, lut_floats);
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

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// tensor map of the (batched) Float32 grid seen as floats: dims (2*Nt0, Nt1, Nt2, B), box (2*box_cells, PY, PZ, 1)
bool lean_tensor_map(CUtensorMap* tm, const void* g, const GeomDev& geo, int B, int box_cells, int PY, int PZ)
{
    static PFN_encodeTiled fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_encodeTiled>(ptr);
        else
            cudaGetLastError();
    }
    if (!fn || 2 * box_cells > 256 || PY > 256 || PZ > 256 || ((uintptr_t)g & 15)) return false;
    cuuint64_t dims[4] = {(cuuint64_t)2 * geo.Nt[0], (cuuint64_t)geo.Nt[1], (cuuint64_t)geo.Nt[2], (cuuint64_t)B};
    cuuint64_t strides[3] = {(cuuint64_t)geo.Nt[0] * 8, (cuuint64_t)geo.Nt[0] * geo.Nt[1] * 8, (cuuint64_t)geo.gsz * 8};
    cuuint32_t box[4] = {(cuuint32_t)(2 * box_cells), (cuuint32_t)PY, (cuuint32_t)PZ, 1};
    cuuint32_t es[4] = {1, 1, 1, 1};
    if (strides[0] & 15) return false;
    return fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<void*>(g), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int MT, int W, bool PEER, bool WIDE>
int interp_lean_launch(nfftb200_plan* p, const void* g, void* fhat, int B, int item_lo, int item_hi, const GeomDev& geo, const BinGeom& bg,
                       size_t smem, const SlabTab& slabs, int lut_floats)
{
    using T = float;
    using C = float2;
    CUtensorMap tmap;
    std::memset(&tmap, 0, sizeof(tmap));
    int use_tma = 0;
    if (WIDE && !PEER && p->kernel_mode != 3)
        use_tma = lean_tensor_map(&tmap, g, geo, B, LEAN_WIDE_PITCH, geo.bs[1] + 2 * MT, geo.bs[2] + 2 * MT) ? 1 : 0;
    auto kern = k_interp_lean<MT, W, PEER, WIDE>;
    CUDA_TRY(p, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    kern<<<dim3(item_hi - item_lo, B), NFFTB_BIN_WARPS * 32, smem, p->stream>>>((const C*)g, (C*)fhat, (const T*)p->d_xs2, p->d_perm2,
                                                                               p->d_bin_start, p->d_items, item_lo, p->M, geo,
                                                                               make_win<T>(p), make_poly_param<T, MT>(p), bg, slabs, tmap,
                                                                               use_tma, lut_floats);
    p->launches++;
    CUDA_TRY(p, cudaGetLastError());
    return NFFTB200_OK;
}

template <int MT, int W>
int interp_lean(nfftb200_plan* p, const void* g, void* fhat, int B, int t_lo, int t_hi, const SlabTab* slabs)
{
    using ILy = LeanInterpLayout<MT, W>;
    GeomDev geo = make_geom<float>(p);
    if (!lean_geom_ok<MT, W>(geo) || geo.bs[0] + 2 * MT > 64) return -1;
    // LINEAR windows: the table (LUTSize + 2 entries, padded to 16 bytes by upload_table) is staged in shared memory
    const int lut_floats = (p->precompute == NFFTB200_LINEAR && p->lut_size + 2 <= 12288) ? (int)p->lut_size + 2 : 0;
    constexpr size_t kTwoPerSM = 113 * 1024;            // two CTAs per SM: (228 KB - 2 x 1 KB reserved) / 2
    BinGeom bg;
    // wide (TMA box) layout: 3.5 % faster on C2 (512 nodes per tile) but 6 % slower at 4096 nodes per tile, where the
    // tile staging no longer matters and the smaller L1 (218 KB of the 256 KB array go to shared memory) does; measured
    const bool sparse = p->M <= 1536 * p->ntiles || p->kernel_mode == 3;
    bool wide = p->kernel_mode != 12 && sparse && ILy::make(geo.bs, bg, true) && ILy::bytes(bg, lut_floats) <= kTwoPerSM;
    if (!wide && !ILy::make(geo.bs, bg, false)) return -1;
    const size_t smem = ILy::bytes(bg, lut_floats);
    if (smem > 227 * 1024) return -1;
    if (nfftb_ensure_bins(p, W, LeanGeom<MT, W>::G) != NFFTB200_OK) return -1;
    const int item_lo = p->h_tile_items[(size_t)t_lo], item_hi = p->h_tile_items[(size_t)t_hi];
    if (item_hi == item_lo) return NFFTB200_OK;
    if (slabs) {
        return wide ? interp_lean_launch<MT, W, true, true>(p, nullptr, fhat, 1, item_lo, item_hi, geo, bg, smem, *slabs, lut_floats)
                    : interp_lean_launch<MT, W, true, false>(p, nullptr, fhat, 1, item_lo, item_hi, geo, bg, smem, *slabs, lut_floats);
    }
    return wide ? interp_lean_launch<MT, W, false, true>(p, g, fhat, B, item_lo, item_hi, geo, bg, smem, SlabTab{}, lut_floats)
                : interp_lean_launch<MT, W, false, false>(p, g, fhat, B, item_lo, item_hi, geo, bg, smem, SlabTab{}, lut_floats);
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
