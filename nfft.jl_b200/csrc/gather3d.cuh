// gather3d.cuh -- the gather pass of the 3-D spreaders (addBlock! without its global lock,
// /root/reference/src/convolution.jl:371-443): every grid cell sums, in a fixed order, the <= 8 padded tiles that cover it
// and is written exactly once.  Shared by spread.cu (stand-alone pass, peer pass) and lean.cu (fused into the spreader).
#pragma once
#include <type_traits>

#include "common.cuh"
#include "tile3d.cuh"

// Column form of the gather for the default 16-cell-thick tiles: one thread owns the cell pair (u0, u0+1) x u1 for
// ALL z of one tile layer, so the x/y candidate search and the work-item lookups are done once per 16 output
// cells and the accumulators stay in registers; the z halos of the layers below/above are added with compile-time
// z ranges.  ~16x fewer instructions per output cell than the per-cell kernel.
// PEER = true is the multi-GPU form (node sharding, comm.cu): the tile scratch of every rank is mapped into this
// process (CUDA IPC over NVLink), a tile's sub-grid is read from the rank that owns the tile, and only the z layers
// [tz0, tz0 + gridDim.z) of this rank's slab are produced -- the spread's halo exchange and the reduce-scatter
// in one pass, with no atomics and a fixed summation order.
// ZS = 2 splits a column between two threads (half a tile layer each): half the accumulator registers, twice the
// resident threads to cover the DRAM latency; the per-cell summation order does not change.
// One column: the cell pair (u0, u0 + 1) x u1 for the planes [zh * BSZ / ZS, (zh + 1) * BSZ / ZS) of tile layer tzc, transform b.
// CG = true reads the scratch with ld.global.cg (the fused spreader of lean.cu gathers tiles that other SMs wrote during
// the same kernel).
template <typename T, int MT, int BSZ, bool PEER, int ZS, bool CG = false>
__device__ __forceinline__ void
gather_cols3d_column(const typename Cplx<T>::type* __restrict__ scratch, typename Cplx<T>::type* __restrict__ g,
                     const int32_t* __restrict__ tile_items, int tile_lo, int tile_hi, int item_lo, int item_hi, const GeomDev& geo,
                     const PeerTab& pt, int tz0, int u0, int u1, int zh, int tzc, int b)
{
    using C = typename Cplx<T>::type;
    constexpr int L = 2 * MT;
    const int PX = geo.bs[0] + L, PY = geo.bs[1] + L;
    constexpr int PZ = BSZ + L;
    const int plane = PX * PY;
    const unsigned PN = (unsigned)plane * PZ;
    constexpr int HZ = BSZ / ZS;                           // planes per thread
    static_assert(BSZ % ZS == 0 && HZ >= MT, "half layers must hold a halo");
    if (u0 >= geo.Nt[0]) return;
    if (!PEER) scratch += (size_t)b * (size_t)(item_hi - item_lo) * PN;
    auto cover = [&](int u, int d, int tmul, int pmul, int (&tt)[3], int (&po)[3]) -> int {
        const int bs = geo.bs[d], nb = geo.nb[d], Nt = geo.Nt[d];
        const int t = (int)fastdiv((unsigned)u, geo.inv_bs[d]), l = u - t * bs;
        const int len = (t == nb - 1) ? Nt - t * bs : bs;
        int n = 0;
        tt[n] = t * tmul; po[n] = (l + MT) * pmul; n++;
        if (l < MT) {
            const int tp = t == 0 ? nb - 1 : t - 1;
            const int lenp = (tp == nb - 1) ? Nt - tp * bs : bs;
            tt[n] = tp * tmul; po[n] = (l + MT + lenp) * pmul; n++;
        }
        if (l >= len - MT) { tt[n] = (t == nb - 1 ? 0 : t + 1) * tmul; po[n] = (l + MT - len) * pmul; n++; }
        return n;
    };
    int ty[3], oy[3];
    const int ny = cover(u1, 1, geo.nb[0], PX, ty, oy);
    const int bs0 = geo.bs[0], nb0 = geo.nb[0];
    const int t = (int)fastdiv((unsigned)u0, geo.inv_bs[0]), l = u0 - t * bs0;
    const int len = (t == nb0 - 1) ? geo.Nt[0] - t * bs0 : bs0;
    const int tp = t == 0 ? nb0 - 1 : t - 1, tn = t == nb0 - 1 ? 0 : t + 1;
    const int lenp = (tp == nb0 - 1) ? geo.Nt[0] - tp * bs0 : bs0;
    // x candidates: (tile, offset, cell-0 valid, cell-1 valid)
    int xt[3], xo[3]; bool xw0[3], xw1[3]; int nx = 0;
    xt[nx] = t; xo[nx] = l + MT; xw0[nx] = true; xw1[nx] = true; nx++;
    if (l < MT) { xt[nx] = tp; xo[nx] = l + MT + lenp; xw0[nx] = true; xw1[nx] = l + 1 < MT; nx++; }
    if (l + 1 >= len - MT) { xt[nx] = tn; xo[nx] = l + MT - len; xw0[nx] = l >= len - MT; xw1[nx] = true; nx++; }
    // z layers: own, below (its high halo covers my planes [0,m)), above (its low halo covers [BSZ-m, BSZ))
    const int nb2 = geo.nb[2], tmz = geo.nb[0] * geo.nb[1];
    const int tzp = tzc == 0 ? nb2 - 1 : tzc - 1, tzn = tzc == nb2 - 1 ? 0 : tzc + 1;
    T ax0[HZ], ay0[HZ], ax1[HZ], ay1[HZ];
#pragma unroll
    for (int k = 0; k < HZ; k++) { ax0[k] = ay0[k] = ax1[k] = ay1[k] = 0; }
    for (int iy = 0; iy < ny; iy++)
        for (int ix = 0; ix < nx; ix++) {
            const int txy = ty[iy] + xt[ix];
            const int oxy = oy[iy] + xo[ix];
            const bool w0 = xw0[ix], w1 = xw1[ix];
            auto add_layer = [&](int tzl, auto lo_tag, auto cnt_tag, int pz0) {
                constexpr int LO = decltype(lo_tag)::value, CNT = decltype(cnt_tag)::value;
                const int tile = tzl * tmz + txy;
                const C* sbase = scratch;
                int ilo = item_lo;
                if (PEER) {
                    int r = 0;
                    while (r + 1 < pt.n && tile >= pt.cut[r + 1]) r++;       // owner of the tile
                    sbase = (const C*)pt.base[r];
                    ilo = pt.item_lo[r];
                } else if (tile < tile_lo || tile >= tile_hi) return;
                const int ia = tile_items[tile], ib = tile_items[tile + 1];
                for (int it = ia; it < ib; it++) {
                    const C* sp = sbase + ((long long)(it - ilo) * (long long)PN + (long long)pz0 * plane + oxy);
#pragma unroll
                    for (int k = 0; k < CNT; k++) {
                        if (w0) { const C c = CG ? __ldcg(sp + (size_t)k * plane) : sp[(size_t)k * plane]; ax0[LO + k] += c.x; ay0[LO + k] += c.y; }
                        if (w1) { const C c = CG ? __ldcg(sp + (size_t)k * plane + 1) : sp[(size_t)k * plane + 1]; ax1[LO + k] += c.x; ay1[LO + k] += c.y; }
                    }
                }
            };
            add_layer(tzc, std::integral_constant<int, 0>{}, std::integral_constant<int, HZ>{}, MT + zh * HZ);
            if (zh == 0) add_layer(tzp, std::integral_constant<int, 0>{}, std::integral_constant<int, MT>{}, MT + BSZ);
            if (zh == ZS - 1) add_layer(tzn, std::integral_constant<int, HZ - MT>{}, std::integral_constant<int, MT>{}, 0);
        }
    C* dst = g + (size_t)b * geo.gsz + ((size_t)((tzc - (PEER ? tz0 : 0)) * BSZ + zh * HZ) * geo.Nt[1] + u1) * geo.Nt[0] + u0;
    const size_t gplane = (size_t)geo.Nt[0] * geo.Nt[1];
#pragma unroll
    for (int k = 0; k < HZ; k++) {
        if (sizeof(T) == 4) *reinterpret_cast<float4*>(dst + k * gplane) = make_float4((float)ax0[k], (float)ay0[k], (float)ax1[k], (float)ay1[k]);
        else { dst[k * gplane] = make_c<T>(ax0[k], ay0[k]); dst[k * gplane + 1] = make_c<T>(ax1[k], ay1[k]); }
    }
}


template <typename T, int MT, int BSZ, bool PEER, int ZS>
__global__ void __launch_bounds__(128)
k_gather_cols3d(const typename Cplx<T>::type* __restrict__ scratch, typename Cplx<T>::type* __restrict__ g,
                const int32_t* __restrict__ tile_items, int tile_lo, int tile_hi, int item_lo, int item_hi, GeomDev geo,
                const __grid_constant__ PeerTab pt, int tz0)
{
    const int u0 = (blockIdx.x * blockDim.x + threadIdx.x) * 2;
    const int u1 = blockIdx.y;
    const int zh = (int)(blockIdx.z % ZS), zl = (int)(blockIdx.z / ZS);
    const int tzc = PEER ? zl + tz0 : zl % geo.nb[2];
    const int b = PEER ? 0 : zl / geo.nb[2];
    gather_cols3d_column<T, MT, BSZ, PEER, ZS>(scratch, g, tile_items, tile_lo, tile_hi, item_lo, item_hi, geo, pt, tz0, u0, u1, zh, tzc, b);
}

