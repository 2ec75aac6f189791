// gather.cuh -- addBlock! without the lock (/root/reference/src/convolution.jl:371-443), tile-centric form.
// One CTA owns the CORE region of one tile of the grid.  It first writes the core from the tile's own partial
// padded tiles (work items), then, neighbour by neighbour in a fixed order, adds the halo parts of the <= 3^D - 1
// neighbouring tiles that overlap the core (read-add-write by the owning CTA only: no atomics, deterministic,
// every grid cell is written by exactly one CTA, so the grid needs no memset).  All control flow is CTA-uniform;
// per-cell work is one index decode, one scratch load and one grid load/store.
// Preconditions (checked on the host): every tile core is at least m cells long in every dimension.
#pragma once
#include "common.cuh"
#include "tile3d.cuh"

template <typename T, int MT, int D>
__global__ void __launch_bounds__(256)
k_gather_core(const typename Cplx<T>::type* __restrict__ scratch, typename Cplx<T>::type* __restrict__ g,
              const int32_t* __restrict__ tile_items, int tile_lo, int tile_hi, int item_lo, int item_hi, GeomDev geo)
{
    using C = typename Cplx<T>::type;
    constexpr int L = 2 * MT;
    int P[3] = {1, 1, 1}, tc[3] = {0, 0, 0}, len[3] = {1, 1, 1}, c0[3] = {0, 0, 0};
    size_t PN = 1;
    int r = blockIdx.x;
#pragma unroll
    for (int d = 0; d < D; d++) {
        P[d] = geo.bs[d] + L; PN *= (size_t)P[d];
        tc[d] = r % geo.nb[d]; r /= geo.nb[d];
        c0[d] = tc[d] * geo.bs[d];
        len[d] = min(geo.bs[d], geo.Nt[d] - c0[d]);
    }
    const int tile = blockIdx.x;
    g += (size_t)blockIdx.y * geo.gsz;
    scratch += (size_t)blockIdx.y * (size_t)(item_hi - item_lo) * PN;
    const size_t gs1 = geo.Nt[0], gs2 = (size_t)geo.Nt[0] * geo.Nt[1];
    const size_t ps1 = P[0], ps2 = (size_t)P[0] * P[1];
    C* gcore = g + (size_t)c0[2] * gs2 + (size_t)c0[1] * gs1 + c0[0];

    // phase 0: own tile (offset 0,0,0) initialises the core; then the 3^D - 1 neighbours add their halos
    constexpr int NNB = (D == 3) ? 27 : 9;
    bool first = true;
    for (int nbi = 0; nbi < NNB; nbi++) {
        // order: own tile first, then the neighbours in lexicographic (dz, dy, dx) order
        int code = nbi == 0 ? (NNB / 2) : (nbi <= NNB / 2 ? nbi - 1 : nbi);
        int dd[3] = {0, 0, 0};
        dd[0] = code % 3 - 1; code /= 3;
        dd[1] = code % 3 - 1; code /= 3;
        if (D == 3) dd[2] = code % 3 - 1;
        int lo[3] = {0, 0, 0}, ext[3] = {1, 1, 1}, po[3] = {0, 0, 0}, nt[3] = {0, 0, 0};
        bool any = true;
#pragma unroll
        for (int d = 0; d < D; d++) {
            const int nb = geo.nb[d];
            if (dd[d] == 0) { lo[d] = 0; ext[d] = len[d]; po[d] = MT; nt[d] = tc[d]; }
            else if (dd[d] < 0) {          // previous tile: its high halo covers my cells [0, m)
                nt[d] = tc[d] == 0 ? nb - 1 : tc[d] - 1;
                const int lenp = (nt[d] == nb - 1) ? geo.Nt[d] - nt[d] * geo.bs[d] : geo.bs[d];
                lo[d] = 0; ext[d] = min(MT, len[d]); po[d] = MT + lenp;
            } else {                        // next tile: its low halo covers my cells [len - m, len)
                nt[d] = tc[d] == nb - 1 ? 0 : tc[d] + 1;
                lo[d] = max(len[d] - MT, 0); ext[d] = len[d] - lo[d]; po[d] = MT - len[d];   // p = l + m - len
            }
            if (ext[d] <= 0) any = false;
        }
        if (!any) continue;
        const int ntile = (D == 3 ? (nt[2] * geo.nb[1] + nt[1]) : nt[1]) * geo.nb[0] + nt[0];
        int it_lo = 0, it_hi = 0;
        if (ntile >= tile_lo && ntile < tile_hi) { it_lo = tile_items[ntile]; it_hi = tile_items[ntile + 1]; }
        const bool own = nbi == 0;
        if (it_hi == it_lo && !(own)) continue;
        const int n01 = ext[0] * ext[1];
        const int ncell = n01 * ext[2];
        const unsigned inv0 = fastdiv_inv(ext[0]), inv01 = fastdiv_inv(n01);
        for (int q = threadIdx.x; q < ncell; q += 256) {
            const int z = (int)fastdiv(q, inv01), r2 = q - z * n01;
            const int y = (int)fastdiv(r2, inv0), x = r2 - y * ext[0];
            const int lx = lo[0] + x, ly = lo[1] + y, lz = lo[2] + z;
            const size_t so = (size_t)(lz + po[2]) * ps2 * (D == 3) + (size_t)(ly + po[1]) * ps1 + (lx + po[0]);
            T ax = 0, ay = 0;
            for (int it = it_lo; it < it_hi; it++) {
                const C c = scratch[(size_t)(it - item_lo) * PN + so];
                ax += c.x; ay += c.y;
            }
            C* gp = gcore + (size_t)lz * gs2 + (size_t)ly * gs1 + lx;
            if (own) *gp = make_c<T>(ax, ay);
            else { C cur = *gp; cur.x += ax; cur.y += ay; *gp = cur; }
        }
        if (first) first = false;
        __syncthreads();
    }
    (void)tile; (void)first;
}
