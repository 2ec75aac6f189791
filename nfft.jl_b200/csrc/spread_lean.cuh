// spread_lean.cuh -- K3, kernel_mode 8: register-window adjoint gridding over the plan-time (tile, bin) order.
//   replaces fillBlock!/fillOneNode! (/root/reference/src/convolution.jl:445-492); hands the same padded tile
//   ("blocks[l]") to the same gather pass (addBlock!, :371-443) as the other 3-D spreaders.
//
// Same arithmetic as spread_bin.cuh (kernel_mode 7): the footprints of a bin's nodes are summed in a W^3 register
// window (lane r owns the x-rows (y, z) = (r % W, r / W + 4p)) and shared memory sees one read-modify-write per bin.
// What changed after the first B200 measurement of mode 7 (profiles/r02_mode7_*: 661 us, 158 warp instructions per
// node, 26 % of the stall samples at the colour barriers):
//   * the nodes arrive grouped by (warp = octant of bins, colour) from the plan (sort.cu: k_bin_order): no staging,
//     no in-kernel sort, and dense data fills the bins (C5: 19 nodes per bin instead of 3 per staged chunk);
//   * the 27 CTA barriers between the colours are replaced by per-warp progress counters: a warp may accumulate the
//     bin of colour c in registers at once, and only its read-modify-write waits until every other warp has
//     finished its colours < c.  Still no atomics on the data path and a fixed summation order (bit-reproducible);
//   * the next round's coordinates and values (a dependent perm -> fHat gather) are fetched one round ahead.
// Float32 only (packed FFMA2 arithmetic); other types keep the default kernels.
#pragma once
#include "bin_common.cuh"
#include "interp_lean.cuh"
#include "gather3d.cuh"
#include <cooperative_groups.h>

#ifndef NFFTB_LEAN_EVICT_LAST
#define NFFTB_LEAN_EVICT_LAST 0     // 1: fused form stores its scratch tile with an L2 evict_last policy
#endif

// Which colours of warp (octant) b must be finished before warp a may read-modify-write its bin of colour c:
//   t[(a * 8 + b) * S^3 + c] = 1 + the last colour c' < c whose bin in octant b overlaps the bin (a, c), 0 if none.
// Bins (a, c) and (b, c') overlap iff their W-wide windows do along every dimension; the window of bin index
// o * S + k (octant half o, colour digit k) starts at W * o + G * k.  Two bins of the same colour never overlap, every
// wait is on an earlier colour, so the order is acyclic.  Waiting for "all warps finished all colours < c" (the first
// form of this kernel) is the upper bound of this table; the exact table lets a warp run ahead of octants it does not
// touch -- 15 % instead of 20 % of the warp time spent waiting in a Poisson model of C2 (scripts in DESIGN.md 3.0).
template <int MT, int W> struct LeanNeed {
    static constexpr int L = 2 * MT, G = W - L + 1, S = (W + G - 1) / G, S3 = S * S * S;
    static constexpr int BYTES = (8 * 8 * S3 + 15) & ~15;
    unsigned char t[BYTES];
};
template <int MT, int W> constexpr LeanNeed<MT, W> lean_make_need()
{
    using N = LeanNeed<MT, W>;
    N r{};
    for (int a = 0; a < 8; a++)
        for (int b = 0; b < 8; b++)
            for (int c = 0; c < N::S3; c++) {
                int need = 0;
                for (int e = 0; e < c && a != b; e++) {
                    bool hit = true;
                    int cc = c, ee = e;
                    for (int d = 0; d < 3; d++) {
                        const int oa = W * ((a >> d) & 1) + N::G * (cc % N::S), ob = W * ((b >> d) & 1) + N::G * (ee % N::S);
                        cc /= N::S; ee /= N::S;
                        const int dist = oa > ob ? oa - ob : ob - oa;
                        if (dist >= W) hit = false;
                    }
                    if (hit) need = e + 1;
                }
                r.t[(a * 8 + b) * N::S3 + c] = (unsigned char)need;
            }
    return r;
}
__device__ const LeanNeed<2, 8> g_lean_need2 = lean_make_need<2, 8>();
__device__ const LeanNeed<3, 8> g_lean_need3 = lean_make_need<3, 8>();
template <int MT> __device__ __forceinline__ const unsigned char* lean_need_table()
{
    if constexpr (MT == 2) return g_lean_need2.t;
    else return g_lean_need3.t;
}

template <int MT, int W> struct LeanSpreadLayout {
    static constexpr int RW = 4 * W + 4;                     // record: wx[W] | wy[W] | (wz * v)[W] re, im | window origin (3 ints)
    static bool make(const int* bs, BinGeom& bg) { return bin_make_geom<float, MT, W>(bs, bg); }
    static size_t bytes(const BinGeom& bg, int lut_floats)
    {
        return sizeof(float2) * (size_t)bg.PNs + sizeof(float) * NFFTB_BIN_WARPS * NFFTB_BIN_ROUND * RW + 64 +
               LeanNeed<MT, W>::BYTES + sizeof(float) * (size_t)((lut_floats + 3) & ~3) + 16 + 16;
    }
};

__device__ __forceinline__ int lean_ld_acquire(const int* p)
{
    int v;
    asm volatile("ld.acquire.cta.shared.b32 %0, [%1];" : "=r"(v) : "r"((unsigned)__cvta_generic_to_shared(p)) : "memory");
    return v;
}
__device__ __forceinline__ void lean_st_release(int* p, int v)
{
    asm volatile("st.release.cta.shared.b32 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(p)), "r"(v) : "memory");
}

// what the fused form needs on top: the grid, the item table of the gather and the arrival counters
struct LeanFuse {
    float2* g;                   // oversampled grid (all transforms)
    const int32_t* tile_items;   // ntiles + 1
    int item_hi;                 // one past the last launched work item
    int ntiles;
    int* ready;                  // [B][ntiles] arrivals per output block, zero at launch
    const int* expect;           // [ntiles] work items among the block's distinct neighbour tiles
};

// FUSE = true: "last arriver gathers".  After its padded tile is in the scratch, a CTA bumps the arrival counter of the
// (up to) 27 output blocks its tile overlaps; whoever completes a block's count sums that block right away from the
// scratch tiles -- written moments ago by neighbouring CTAs, so the reads hit L2 instead of DRAM -- and writes the 16^3
// grid cells once.  Integer counters only: no floating-point atomics, no waiting (nothing to deadlock on), and the
// per-cell summation order is the gather pass's fixed order whoever runs it.  Replaces the separate gather launch.
// PAIR = true (kernel_mode 13, experiment): thread-block clusters of two x-adjacent tiles exchange their shared halo through
// distributed shared memory -- the north_star's "clusters with DSMEM for halo accumulation".  After both tiles are
// complete (cluster barrier) each CTA adds, from its partner's shared memory, the m halo columns of the partner that fall
// into its own core, and the pair writes ONE merged padded block of 2*bs + 2m columns to the scratch instead of two
// padded tiles (-13.6 % scratch volume for m = 3); the gather pass then sees tiles of 32 x 16 x 16 cells.
template <int MT, int W, bool FUSE, bool PAIR = false>
__global__ void __launch_bounds__(NFFTB_BIN_WARPS * 32, 2)
k_spread_lean(const float2* __restrict__ fhat, float2* __restrict__ scratch, const float* __restrict__ xs2,
              const int32_t* __restrict__ perm2, const int32_t* __restrict__ bin_start, const int32_t* __restrict__ items,
              int item_lo, long long M, GeomDev geo, WinDev<float> win, const __grid_constant__ PolyParam<float, MT> pp,
              BinGeom bg, const __grid_constant__ LeanFuse fz, int lut_floats)
{
    using T = float;
    using C = float2;
    using LG = LeanGeom<MT, W>;
    constexpr int L = LG::L, G = LG::G, S3 = LG::S3, NQ = LG::NQ, RW = LeanSpreadLayout<MT, W>::RW;
    constexpr int NWARP = NFFTB_BIN_WARPS, NTHR = NWARP * 32, RND = NFFTB_BIN_ROUND, NP = 2;

    extern __shared__ __align__(16) unsigned char smem_raw[];
    C* P = reinterpret_cast<C*>(smem_raw);                                          // [PZ][PL] padded tile
    T* rec = reinterpret_cast<T*>(P + bg.PNs);                                      // [NWARP][RND][RW]
    int* done = reinterpret_cast<int*>(rec + NWARP * RND * RW);                     // [NWARP] colours finished per warp (64 bytes)
    unsigned char* need_s = reinterpret_cast<unsigned char*>(done + 16);            // [NWARP][NWARP][S3] colour dependencies
    T* lut = reinterpret_cast<T*>(need_s + LeanNeed<MT, W>::BYTES);                 // [lut_floats, rounded up to 4]: LINEAR window table
    unsigned long long* mbar = reinterpret_cast<unsigned long long*>(lut + ((lut_floats + 3) & ~3));

    const int32_t* item = items + 3 * (size_t)(item_lo + blockIdx.x);
    const int tile_id = item[0];
    const int n_lo = item[1], n_hi = item[2];
    const int tx = tile_id % geo.nb[0];
    const int ty = (tile_id / geo.nb[0]) % geo.nb[1];
    const int tz = tile_id / (geo.nb[0] * geo.nb[1]);
    const int cx0 = tx * geo.bs[0], cy0 = ty * geo.bs[1], cz0 = tz * geo.bs[2];     // first core cell
    const int PX = geo.bs[0] + L, PY = geo.bs[1] + L, PZ = geo.bs[2] + L;
    const int PXp = bg.PXp, PL = bg.PL;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    fhat += (long long)blockIdx.y * M;
    const float2* scratch_base = scratch;
    if (PAIR) scratch += ((size_t)blockIdx.y * (gridDim.x >> 1) + (blockIdx.x >> 1)) * ((size_t)(2 * geo.bs[0] + L) * PY * PZ);
    else scratch += ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * ((size_t)PX * PY * PZ);
    T* myrec = rec + warp * RND * RW;

    // the S^3 + 1 bin boundaries of this warp's octant: one load per lane, handed out by shuffles
    const int q0 = tile_id * NQ + warp * S3;
    static_assert(S3 + 1 <= 32, "one bin boundary per lane");
    const int bsl = min(max(bin_start[q0 + min(lane, S3)], n_lo), n_hi);
    const int nl0 = __shfl_sync(0xffffffffu, bsl, 0), nl1 = __shfl_sync(0xffffffffu, bsl, S3);
    const int rowy = lane & (W - 1), rowz = lane >> 3;                       // row (y, z + 4p) of the window
    const int wn = lane / 3, wd = lane - 3 * wn;                             // lane = (node of the round, dimension)
    const int wNt = wd == 0 ? geo.Nt[0] : (wd == 1 ? geo.Nt[1] : geo.Nt[2]);
    const int wc0 = wd == 0 ? cx0 : (wd == 1 ? cy0 : cz0);

    if (lut_floats > 0 && threadIdx.x == 0) {                                // stage the window table: one bulk (TMA) copy
        const unsigned bytes = (unsigned)(sizeof(T) * ((lut_floats + 3) & ~3));
        mbar_init(mbar, 1);
        mbar_expect_tx(mbar, bytes);
        lean_bulk_g2s(lut, win.lin, bytes, mbar);
    }
    // first round's inputs in flight while the tile is zeroed
    T xnext = (T)0;
    C vnext = make_float2(0.f, 0.f);
    if (wn < RND && nl0 + wn < nl1) {
        xnext = xs2[(long long)(nl0 + wn) * 3 + wd];
        if (wd == 2) vnext = fhat[perm2[nl0 + wn]];
    }
    {
        uint4* z = reinterpret_cast<uint4*>(P);
        const int n16 = (int)((sizeof(C) * (size_t)bg.PNs) / 16);
        for (int q = threadIdx.x; q < n16; q += NTHR) z[q] = make_uint4(0, 0, 0, 0);
        if (threadIdx.x < NWARP) done[threadIdx.x] = 0;
        const uint4* nt = reinterpret_cast<const uint4*>(lean_need_table<MT>());
        for (int q = threadIdx.x; q < LeanNeed<MT, W>::BYTES / 16; q += NTHR) reinterpret_cast<uint4*>(need_s)[q] = nt[q];
    }
    __syncthreads();
    WinDev<T> winl = win;
    if (lut_floats > 0) { lean_mbar_wait(mbar, 0); winl.lin = lut; }

    int rbase = nl0 - RND;                                                   // list index of the resident round
    int pos = nl0;                                                           // next node of this warp's list
    for (int c = 0; c < S3; c++) {
        const int hi = __shfl_sync(0xffffffffu, bsl, c + 1);
        if (hi > pos) {                                                      // warp-uniform: bin of colour c is not empty
            BinRow<T, W> acc[NP];
#pragma unroll
            for (int p = 0; p < NP; p++) acc[p].zero();
            int o0 = 0, o1 = 0, o2 = 0;                                      // window origin of the bin, padded-tile coordinates
            const int lo = pos;
            for (int i = lo; i < hi;) {
                if (i >= rbase + RND) {                                      // warp-uniform: next round of records
                    __syncwarp();                                            // the previous round has been read
                    rbase = i;
                    const T x = xnext;
                    const C v = vnext;
                    {
                        const int nb0 = rbase + RND;
                        if (wn < RND && nb0 + wn < nl1) {
                            xnext = xs2[(long long)(nb0 + wn) * 3 + wd];
                            if (wd == 2) vnext = fhat[perm2[nb0 + wn]];
                        }
                    }
                    if (wn < RND && rbase + wn < nl1) {                      // weights of (node wn of the round, dimension wd)
                        T ks;
                        const int cc = node_cell<T>(x, wNt, ks);
                        T w[L];
                        eval_taps<T, MT>(winl, pp, ks, cc, w);
                        const int lc = cc - wc0;                             // first tap at padded coordinate lc + 1
                        const int wo = 1 + bin_first<W, G>(bin_of<W, G>(lc));
                        const int dl = lc + 1 - wo;                          // first tap inside the window, in [0, G)
                        T* rn = myrec + wn * RW;
                        reinterpret_cast<int*>(rn + 4 * W)[wd] = wo;
                        if (wd < 2) {
#pragma unroll
                            for (int l = 0; l < L; l++) rn[wd * W + dl + l] = w[l];
#pragma unroll
                            for (int j = 0; j < W - L; j++) rn[wd * W + (j < dl ? j : j + L)] = (T)0;
                        } else {
                            C* rz = reinterpret_cast<C*>(rn + 2 * W);
#pragma unroll
                            for (int l = 0; l < L; l++) rz[dl + l] = make_float2(w[l] * v.x, w[l] * v.y);
#pragma unroll
                            for (int j = 0; j < W - L; j++) rz[j < dl ? j : j + L] = make_float2(0.f, 0.f);
                        }
                    }
                    __syncwarp();
                }
                const T* rn = myrec + (i - rbase) * RW;
                if (i == lo) {
                    const int4 org = *reinterpret_cast<const int4*>(rn + 4 * W);
                    o0 = org.x; o1 = org.y; o2 = org.z;
                }
                // the bin's nodes inside the resident round: a tight run with one pointer increment per node
                const T* rend = myrec + (min(hi, rbase + RND) - rbase) * RW;
                i = min(hi, rbase + RND);
                for (; rn < rend; rn += RW) {
                    T wx[W];
                    bin_load_row<T, W>(rn, wx);
                    const T wy = rn[W + rowy];
#pragma unroll
                    for (int p = 0; p < NP; p++) {
                        const C vz = reinterpret_cast<const C*>(rn + 2 * W)[rowz + 4 * p];
                        acc[p].axpy(wx, wy, vz);
                    }
                }
            }
            pos = hi;
            // the overlapping bins of earlier colours must be finished before this window is read-modify-written
            if (c > 0) {
                const int need = (lane < NWARP) ? (int)need_s[(warp * NWARP + lane) * S3 + c] : 0;
                for (;;) {
                    const int d = (lane < NWARP) ? lean_ld_acquire(done + lane) : S3;
                    if (__all_sync(0xffffffffu, d >= need)) break;
                }
            }
            // one read-modify-write of the window (cells beyond the padded tile carry zero weights only)
#pragma unroll
            for (int p = 0; p < NP; p++) {
                const int Y = o1 + rowy, Z = o2 + rowz + 4 * p;
                if (Y < PY && Z < PZ) {
                    C* row = P + (Z * PL + Y * PXp + o0);
                    if (o0 + W <= PX) {                                      // warp-uniform: all but the last bin of a row
#pragma unroll
                        for (int k = 0; k < W; k++) { C cv = row[k]; acc[p].add_to(cv, k); row[k] = cv; }
                    } else {
#pragma unroll
                        for (int k = 0; k < W; k++)
                            if (o0 + k < PX) { C cv = row[k]; acc[p].add_to(cv, k); row[k] = cv; }
                    }
                }
            }
            __threadfence_block();
        }
        __syncwarp();
        if (lane == 0) lean_st_release(done + warp, c + 1);
    }
    __syncthreads();

    if constexpr (PAIR) {
        namespace cg = cooperative_groups;
        cg::cluster_group cl = cg::this_cluster();
        cl.sync();                                                           // both tiles of the pair are complete
        const int r = (int)cl.block_rank();                                  // 0: low-x tile, 1: high-x tile
        const C* Pp = cl.map_shared_rank(P, r ^ 1);                          // the partner's tile, read over DSMEM
        const int bs0 = geo.bs[0];
        const int mine0 = r == 0 ? bs0 : MT, theirs0 = r == 0 ? 0 : bs0 + MT;   // my core columns next to the partner <- its halo
        for (int idx = threadIdx.x; idx < PZ * PY * MT; idx += NTHR) {
            const int i = idx % MT, row = idx / MT;
            const int z = row / PY, y = row - z * PY;
            const int o = z * PL + y * PXp;
            const C a = Pp[o + theirs0 + i];
            C b = P[o + mine0 + i];
            b.x += a.x; b.y += a.y;
            P[o + mine0 + i] = b;
        }
        cl.sync();                                                           // the partner has read my halo; my sums are visible
        // merged block [PZ][PY][2*bs0 + 2m]: the low tile owns block columns [0, bs0 + m), the high tile the rest
        const int PXB = 2 * bs0 + L, ncol = bs0 + MT;
        const int x_lo = r == 0 ? 0 : MT, bcol0 = r == 0 ? 0 : bs0 + MT;      // first local column I write, where it goes
        for (int idx = threadIdx.x; idx < PZ * PY * ncol; idx += NTHR) {
            const int i = idx % ncol, row = idx / ncol;
            const int z = row / PY, y = row - z * PY;
            scratch[(size_t)row * PXB + bcol0 + i] = P[z * PL + y * PXp + x_lo + i];
        }
        return;
    }
    // ---- flush the padded tile to its scratch slot, dense [PZ][PY][PX] (what the gather pass reads)
    if ((PX & 1) == 0) {
        const int hx = PX >> 1;
        const unsigned inv_hx = fastdiv_inv(hx), inv_py = fastdiv_inv(PY);
        for (int idx = threadIdx.x; idx < PZ * PY * hx; idx += NTHR) {
            const int row = (int)fastdiv(idx, inv_hx), u = idx - row * hx;
            const int z = (int)fastdiv(row, inv_py), y = row - z * PY;
            const C* src = P + (z * PL + y * PXp + 2 * u);
            const C a = src[0], b = src[1];
#if NFFTB_LEAN_EVICT_LAST
            if (FUSE) {
                unsigned long long pol;
                asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
                asm volatile("st.global.L2::cache_hint.v4.f32 [%0], {%1, %2, %3, %4}, %5;" ::"l"(scratch + ((size_t)row * PX + 2 * u)),
                             "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "l"(pol) : "memory");
            } else
#endif
            *reinterpret_cast<float4*>(scratch + ((size_t)row * PX + 2 * u)) = make_float4(a.x, a.y, b.x, b.y);
        }
    } else {
        const unsigned inv_px = fastdiv_inv(PX), inv_py = fastdiv_inv(PY);
        for (int idx = threadIdx.x; idx < PZ * PY * PX; idx += NTHR) {
            const int row = (int)fastdiv(idx, inv_px), x = idx - row * PX;
            const int z = (int)fastdiv(row, inv_py), y = row - z * PY;
            scratch[idx] = P[z * PL + y * PXp + x];
        }
    }
    if constexpr (FUSE) {
        __shared__ int s_glist[27];
        __shared__ int s_gcount;
        __threadfence();                                                     // the tile is visible device-wide ...
        if (threadIdx.x == 0) s_gcount = 0;
        __syncthreads();                                                     // ... before any arrival is counted
        const int nb0 = geo.nb[0], nb1 = geo.nb[1], nb2 = geo.nb[2];
        if (threadIdx.x < 27) {
            const int ox = (int)threadIdx.x % 3 - 1, oy = ((int)threadIdx.x / 3) % 3 - 1, oz = (int)threadIdx.x / 9 - 1;
            auto distinct = [](int off, int nb) { return nb >= 3 || (nb == 2 ? off <= 0 : off == 0); };
            if (distinct(ox, nb0) && distinct(oy, nb1) && distinct(oz, nb2)) {
                const int nx = (tx + ox + nb0) % nb0, ny = (ty + oy + nb1) % nb1, nz = (tz + oz + nb2) % nb2;
                const int nbt = (nz * nb1 + ny) * nb0 + nx;
                const int old = atomicAdd(fz.ready + (size_t)blockIdx.y * fz.ntiles + nbt, 1);
                if (old + 1 == fz.expect[nbt]) s_glist[atomicAdd(&s_gcount, 1)] = nbt;
            }
        }
        __syncthreads();
        const int ng = s_gcount;
        const int hx = geo.bs[0] >> 1;
        const int xp = (int)threadIdx.x % hx, yy = ((int)threadIdx.x / hx) % geo.bs[1], zh = (int)threadIdx.x / (hx * geo.bs[1]);
        for (int k = 0; k < ng; k++) {
            __threadfence();
            const int t = s_glist[k];
            const int gx = t % nb0, gy = (t / nb0) % nb1, gz = t / (nb0 * nb1);
            const int u0 = gx * geo.bs[0] + 2 * xp, u1 = gy * geo.bs[1] + yy;
            if (zh < 2 && u1 < geo.Nt[1])
                gather_cols3d_column<T, MT, 16, false, 2, true>(scratch_base, fz.g, fz.tile_items, 0, fz.ntiles, item_lo, fz.item_hi, geo,
                                                                PeerTab{}, 0, u0, u1, zh, gz, (int)blockIdx.y);
        }
    }
}

// blocks no work item reaches (expect == 0) are never gathered by the fused spreader: zero them
template <int MT>
__global__ void __launch_bounds__(256) k_zero_empty_blocks(float2* __restrict__ g, const int* __restrict__ expect, GeomDev geo)
{
    const int t = blockIdx.x;
    if (expect[t] != 0) return;
    const int tx = t % geo.nb[0], ty = (t / geo.nb[0]) % geo.nb[1], tz = t / (geo.nb[0] * geo.nb[1]);
    g += (size_t)blockIdx.y * geo.gsz;
    const int n = geo.bs[0] * geo.bs[1] * geo.bs[2];
    for (int q = threadIdx.x; q < n; q += 256) {
        const int x = tx * geo.bs[0] + q % geo.bs[0], y = ty * geo.bs[1] + (q / geo.bs[0]) % geo.bs[1], z = tz * geo.bs[2] + q / (geo.bs[0] * geo.bs[1]);
        if (x < geo.Nt[0] && y < geo.Nt[1] && z < geo.Nt[2]) g[((size_t)z * geo.Nt[1] + y) * geo.Nt[0] + x] = make_float2(0.f, 0.f);
    }
}
