// interp_lean.cuh -- K4, kernel_mode 8: register-window forward interpolation over the plan-time (tile, bin) order.
//   replaces toBlock!/calcOneBlock!/calcOneNode! (/root/reference/src/convolution.jl:229-344).
//
// Same arithmetic as interp_bin.cuh (kernel_mode 7): the W^3 window of a bin lives in registers (lane r owns the x-rows
// (y, z) = (r % W, r / W + 4p)), a node is a register dot product with its window-aligned weights, 8 nodes are folded
// by one halving butterfly.  What changed after the first B200 measurement of mode 7 (profiles/r02_mode7_*):
//   * the nodes arrive already grouped by (warp = octant of bins, bin) from the plan (sort.cu: k_bin_order), so the
//     per-chunk staging, key computation, counting sort and the `order` indirection are gone -- a warp reads its
//     contiguous node list straight from global memory, one round of 8 nodes ahead of the one it is working on;
//   * denser data now fills the bins (C5: 19 nodes per bin instead of 3 per chunk), so the window load amortises;
//   * window origins are clamped into the padded tile, so no load is bounds-checked.
// Float32 only (packed FFMA2 arithmetic); other types keep the default kernels.
#pragma once
#include <cuda.h>

#include <type_traits>

#include "bin_common.cuh"
#include "interp_bin.cuh"

template <int MT, int W> struct LeanGeom {
    static constexpr int L = 2 * MT;
    static constexpr int G = W - L + 1;
    static constexpr int S = (W + G - 1) / G;
    static constexpr int S3 = S * S * S;
    static constexpr int NQ = 8 * S3;                        // bins per tile: 8 octants x S^3 colours
    static_assert(W == 8, "lane = (y, z) row mapping below assumes W = 8 (two passes of 32 rows)");
};

// Two shared-memory layouts of the padded tile:
//   compact: row / plane pitches searched for conflict-free 8-byte window loads (bin_make_geom), cp.async staging;
//   WIDE   : rows of LEAN_WIDE_PITCH = 26 cells starting at the even cell x0 - (x0 & 1), planes of 26 * PY cells -- the
//            dense image of ONE TMA tensor-map box (cp.async.bulk.tensor, SASS UTMALDG), which is how interior tiles
//            are staged (the box start must be 16-byte aligned, hence the even start; tiles that wrap periodically are
//            staged row by row with cp.async into the same layout).  A pitch of 26 cells = 13 sixteen-byte granules keeps
//            the 8 lanes of a quarter-warp (8 consecutive y rows) on 8 distinct granules, so the window is loaded with
//            conflict-free LDS.128.  In this layout the x extent of the register window is LEAN_WIDE_WX = 10 cells starting
//            on an EVEN slot: 5 aligned LDS.128 per row, no parity cases, and -- because 10 cells hold the 2m <= 6 taps of
//            every node whose first tap lies in a run of 4 cells -- only TWO windows per 8-cell octant along x instead of
//            three (the plan-time order refines every x bin by that 4-cell half, sort.cu: k_bin_order).  Measured before
//            this form (ncu source view, C2): the 8-cell window cost 30 issue slots per node (LDS + 32 register moves
//            that reconcile the even / odd start cases at the join) against 36 for the arithmetic itself.
constexpr int LEAN_WIDE_PITCH = 26;
constexpr int LEAN_WIDE_WX = 10;

template <int MT, int W> struct LeanInterpLayout {
    static constexpr int RW = 4 * W;                         // record: wx[W] | wy[W] | wz[W] | window origin (3 ints), pad
    static bool make(const int* bs, BinGeom& bg, bool wide)
    {
        if (!bin_make_geom<float, MT, W>(bs, bg)) return false;
        if (wide) {
            if (bs[0] + 2 * MT + 1 > LEAN_WIDE_PITCH) return false;
            bg.PXp = LEAN_WIDE_PITCH;
            bg.PL = LEAN_WIDE_PITCH * (bs[1] + 2 * MT);
            bg.PNs = bg.PL * (bs[2] + 2 * MT);
        }
        return true;
    }
    // lut_floats: entries of the LINEAR window table staged in shared memory (0: none)
    static size_t bytes(const BinGeom& bg, int lut_floats)
    {
        return sizeof(float2) * (size_t)bg.PNs + sizeof(float) * NFFTB_BIN_WARPS * NFFTB_BIN_ROUND * RW +
               sizeof(float) * NFFTB_BIN_WARPS * 2 * NFFTB_BIN_ROUND + sizeof(float) * (size_t)((lut_floats + 3) & ~3) + 16 + 16;
    }
};

// mbarrier wait with a label that stays unique when the function is inlined several times
__device__ __forceinline__ void lean_mbar_wait(unsigned long long* bar, unsigned parity)
{
    unsigned done = 0;
    const unsigned addr = (unsigned)__cvta_generic_to_shared(bar);
    while (!done) {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                     : "=r"(done) : "r"(addr), "r"(parity) : "memory");
    }
}
// 1-D bulk copy global -> shared through the TMA engine (SASS UBLKCP), completion on an mbarrier; 16-byte granular
__device__ __forceinline__ void lean_bulk_g2s(void* sdst, const void* gsrc, unsigned bytes, unsigned long long* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"((unsigned)__cvta_generic_to_shared(sdst)), "l"(gsrc), "r"(bytes), "r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}

// lut_floats > 0 (LINEAR windows): the interpolation table is staged in shared memory by one bulk copy and eval_taps
// reads it there instead of issuing two global loads per tap.
template <int MT, int W, bool PEER, bool WIDE>
__global__ void __launch_bounds__(NFFTB_BIN_WARPS * 32, 2)
k_interp_lean(const float2* __restrict__ g, float2* __restrict__ fhat, const float* __restrict__ xs2,
              const int32_t* __restrict__ perm2, const int32_t* __restrict__ bin_start, const int32_t* __restrict__ items,
              int item_lo, long long M, GeomDev geo, WinDev<float> win, const __grid_constant__ PolyParam<float, MT> pp,
              BinGeom bg, const __grid_constant__ SlabTab slabs, const __grid_constant__ CUtensorMap tmap, int use_tma,
              int lut_floats)
{
    using T = float;
    using C = float2;
    using LG = LeanGeom<MT, W>;
    constexpr int L = LG::L, G = LG::G, S3 = LG::S3, NQ = LG::NQ, RW = LeanInterpLayout<MT, W>::RW;
    constexpr int NWARP = NFFTB_BIN_WARPS, RND = NFFTB_BIN_ROUND, NP = 2;
    static_assert(RND == 8, "rounds of 8 nodes");
    // record of a node: wx[WXR] | wy[W] | wz[W] | window origin (3 ints); WX = cells of the register window along x
    constexpr int WX = WIDE ? LEAN_WIDE_WX : W, WXR = WIDE ? 12 : W;
    constexpr int OY = WXR, OZ = WXR + W, OO = WXR + 2 * W;
    static_assert(OO + 4 <= RW && (OO % 4) == 0, "record layout");
    static_assert(!WIDE || LEAN_WIDE_WX >= L + 4, "a 10-cell window must hold the taps of 4 consecutive first-tap cells");

    extern __shared__ __align__(128) unsigned char smem_raw[];
    C* P = reinterpret_cast<C*>(smem_raw);                                          // [PZ][PL] padded tile
    T* rec = reinterpret_cast<T*>(P + bg.PNs);                                      // [NWARP][RND][RW]
    T* res = rec + NWARP * RND * RW;                                                // [NWARP][2 * RND]
    T* lut = res + NWARP * 2 * RND;                                                 // [lut_floats, rounded up to 4]
    unsigned long long* mbar = reinterpret_cast<unsigned long long*>(lut + ((lut_floats + 3) & ~3));

    const int32_t* item = items + 3 * (size_t)(item_lo + blockIdx.x);
    const int tile_id = item[0];
    const int n_lo = item[1], n_hi = item[2];
    if (n_hi <= n_lo) return;
    const int tx = tile_id % geo.nb[0];
    const int ty = (tile_id / geo.nb[0]) % geo.nb[1];
    const int tz = tile_id / (geo.nb[0] * geo.nb[1]);
    const int cx0 = tx * geo.bs[0], cy0 = ty * geo.bs[1], cz0 = tz * geo.bs[2];     // first core cell
    const int PX = geo.bs[0] + L, PY = geo.bs[1] + L, PZ = geo.bs[2] + L;
    const int PXp = bg.PXp, PL = bg.PL;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (!PEER) g += (long long)blockIdx.y * geo.gsz;
    fhat += (long long)blockIdx.y * M;
    T* myrec = rec + warp * RND * RW;
    T* myres = res + warp * 2 * RND;

    // ---- toBlock!: the padded tile.  WIDE: interior tiles by one TMA box load, the others (periodic wrap) row by row.
    const int xs_ = WIDE ? ((cx0 - MT) & 1) : 0;                              // tile cell X sits at row offset X + xs_
    bool tma_tile = false;
    if (WIDE && !PEER) {
        const int x0 = cx0 - MT, y0 = cy0 - MT, z0 = cz0 - MT;
        tma_tile = use_tma && x0 >= 0 && y0 >= 0 && z0 >= 0 && x0 + PX <= geo.Nt[0] && y0 + PY <= geo.Nt[1] && z0 + PZ <= geo.Nt[2];
    }
    const bool use_bar = tma_tile || lut_floats > 0;
    if (use_bar && threadIdx.x == 0) {
        mbar_init(mbar, 1);
        unsigned bytes = 0;
        if (tma_tile) bytes += (unsigned)(sizeof(C) * LEAN_WIDE_PITCH * PY * PZ);
        if (lut_floats > 0) bytes += (unsigned)(sizeof(T) * ((lut_floats + 3) & ~3));
        mbar_expect_tx(mbar, bytes);
        if (tma_tile) tma_load_4d(P, &tmap, mbar, 2 * (cx0 - MT - xs_), cy0 - MT, cz0 - MT, (int)blockIdx.y);
        if (lut_floats > 0) lean_bulk_g2s(lut, win.lin, (unsigned)(sizeof(T) * ((lut_floats + 3) & ~3)), mbar);
    }
    if (!tma_tile && WIDE) {
        // all 26 slots of every row, two cells (16 bytes) per copy: slot 0 is the even cell x0 - xs_, the grid width is
        // even, so a pair never straddles the periodic wrap.  16 lanes per row, two rows per warp pass.
        const int x0 = cx0 - MT - xs_, y0 = cy0 - MT, z0 = cz0 - MT;
        const bool fw = LEAN_WIDE_PITCH <= geo.Nt[0] && PY <= geo.Nt[1] && PZ <= geo.Nt[2];
        const int pr = lane & 15, rsel = lane >> 4;
        const bool onp = pr < LEAN_WIDE_PITCH / 2;
        const int xg = wrapc(x0 + 2 * pr, geo.Nt[0], fw);
        for (int z = 0; z < PZ; z++) {
            unsigned gz = (unsigned)wrapc(z0 + z, geo.Nt[2], fw);
            const C* gb = g;
            if (PEER) {                                                     // plane gz lives on rank gz / planes
                const unsigned owner = gz / (unsigned)slabs.planes;
                gb = (const C*)slabs.base[owner];
                gz -= owner * (unsigned)slabs.planes;
            }
            gz *= geo.Nt[1];
            for (int y = 2 * warp + rsel; y < PY; y += 2 * NWARP) {
                const C* src = gb + (size_t)(gz + wrapc(y0 + y, geo.Nt[1], fw)) * (unsigned)geo.Nt[0] + xg;
                if (onp) cp_async_pair(P + (z * PL + y * PXp + 2 * pr), src);
            }
        }
    }
    if (!tma_tile && !WIDE) {
        const int x0 = cx0 - MT, y0 = cy0 - MT, z0 = cz0 - MT;
        const bool fw = PX <= geo.Nt[0] && PY <= geo.Nt[1] && PZ <= geo.Nt[2];
        const int xg0 = wrapc(x0 + lane, geo.Nt[0], fw), xg1 = wrapc(x0 + lane + 32, geo.Nt[0], fw);
        const bool on0 = lane < PX, on1 = lane + 32 < PX;
        for (int z = 0; z < PZ; z++) {
            unsigned gz = (unsigned)wrapc(z0 + z, geo.Nt[2], fw);
            const C* gb = g;
            if (PEER) {                                                     // plane gz lives on rank gz / planes
                const unsigned owner = gz / (unsigned)slabs.planes;
                gb = (const C*)slabs.base[owner];
                gz -= owner * (unsigned)slabs.planes;
            }
            gz *= geo.Nt[1];
            for (int y = warp; y < PY; y += NWARP) {
                const C* src = gb + (size_t)(gz + wrapc(y0 + y, geo.Nt[1], fw)) * (unsigned)geo.Nt[0];
                C* dst = P + (z * PL + y * PXp + lane);
                if (on0) cp_async_cell(dst, src + xg0);
                if (on1) cp_async_cell(dst + 32, src + xg1);
            }
        }
    }

    // this warp's node list: the S^3 bins of octant `warp`, contiguous in the (tile, bin) order
    const int q0 = tile_id * NQ + warp * S3;
    const int nl0 = min(max(bin_start[q0], n_lo), n_hi), nl1 = min(max(bin_start[q0 + S3], n_lo), n_hi);
    const int rowy = lane & (W - 1), rowz = lane >> 3;                       // row (y, z + 4p) of the window
    const int wn = lane / 3, wd = lane - 3 * wn;                             // lane = (node of the round, dimension)
    const int wNt = wd == 0 ? geo.Nt[0] : (wd == 1 ? geo.Nt[1] : geo.Nt[2]);
    const int wc0 = wd == 0 ? cx0 : (wd == 1 ? cy0 : cz0);
    const int wmax = (wd == 0 ? PX : (wd == 1 ? PY : PZ)) - W;               // largest window origin inside the tile
    T xnext = (T)0;
    if (wn < RND && nl0 + wn < nl1) xnext = xs2[(long long)(nl0 + wn) * 3 + wd];
    int jnext = (nl0 + lane < nl1 && lane < RND) ? perm2[nl0 + lane] : 0;

    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncthreads();                                                         // tile resident (cp.async part), mbarrier initialised
    if (use_bar) lean_mbar_wait(mbar, 0);                                    // TMA box and / or the window table have landed
    WinDev<T> winl = win;
    if (lut_floats > 0) winl.lin = lut;

    BinRow<T, WX> win_row[NP];
    int lastwo = -1;                                                         // window origin (dimension wd) of the previous node
    for (int rbase = nl0; rbase < nl1; rbase += RND) {
        const int nn = min(RND, nl1 - rbase);
        const T x = xnext;
        const int jdst = jnext;
        {   // next round's coordinates and destinations: in flight while this round is worked on
            const int nb0 = rbase + RND;
            if (wn < RND && nb0 + wn < nl1) xnext = xs2[(long long)(nb0 + wn) * 3 + wd];
            if (lane < RND && nb0 + lane < nl1) jnext = perm2[nb0 + lane];
        }
        int wo = -1;
        if (wn < nn) {                                                       // weights of (node wn of the round, dimension wd)
            T ks;
            const int c = node_cell<T>(x, wNt, ks);
            T w[L];
            eval_taps<T, MT>(winl, pp, ks, c, w);
            const int lc = c - wc0;                                          // first tap at padded coordinate lc + 1
            wo = min(1 + bin_first<W, G>(bin_of<W, G>(lc)), wmax);
            int ro = wd * W, nz = W - L;                                     // row of the record, zeros around the taps
            if (WIDE) {
                if (wd == 0) {                                               // 10-cell window on an even slot: 4 cells of first taps
                    const int fb = 4 * (lc >> 2) + 1;
                    wo = fb - ((fb + xs_) & 1);
                    nz = WXR - L;
                } else {
                    ro = OY + (wd - 1) * W;
                }
            }
            const int dl = lc + 1 - wo;                                      // first tap inside the window
            reinterpret_cast<int*>(myrec + wn * RW + OO)[wd] = wo;
            T* rn = myrec + wn * RW + ro;                                    // 2m taps at [dl, dl + 2m), zeros elsewhere
#pragma unroll
            for (int l = 0; l < L; l++) rn[dl + l] = w[l];
#pragma unroll
            for (int j = 0; j < WXR - L; j++)
                if (j < nz) rn[j < dl ? j : j + L] = (T)0;
        }
        // which nodes of the round start a new window: bit 3n + d = origin d of node n differs from its predecessor's
        unsigned chm;
        {
            const int up = __shfl_up_sync(0xffffffffu, wo, 3);
            const int pw = wn == 0 ? lastwo : up;
            chm = __ballot_sync(0xffffffffu, wn < nn && wo != pw);
            lastwo = __shfl_sync(0xffffffffu, wo, 3 * (nn - 1) + wd);
        }
        __syncwarp();
        // a full round (all but the last of a warp's list) carries no per-node guards
        auto fold_round = [&](auto full_tag) {
            constexpr bool FULL = decltype(full_tag)::value;
            T v[2 * RND];
#pragma unroll
            for (int n = 0; n < RND; n++) {
                C sn = make_float2(0.f, 0.f);
                if (FULL || n < nn) {                                        // warp-uniform
                    const T* rn = myrec + n * RW;
                    if (chm & (7u << (3 * n))) {                             // warp-uniform: first node of a window
                        const int4 org = *reinterpret_cast<const int4*>(rn + OO);
                        const C* row = P + ((org.z + rowz) * PL + (org.y + rowy) * PXp + org.x + xs_);
#pragma unroll
                        for (int p = 0; p < NP; p++) {
                            if (WIDE) {                                      // org.x + xs_ is even: aligned 16-byte loads
#pragma unroll
                                for (int k = 0; k < WX / 2; k++) {
                                    const float4 u = *reinterpret_cast<const float4*>(row + p * 4 * PL + 2 * k);
                                    win_row[p].set(2 * k, make_float2(u.x, u.y)); win_row[p].set(2 * k + 1, make_float2(u.z, u.w));
                                }
                            } else {
#pragma unroll
                                for (int k = 0; k < WX; k++) win_row[p].set(k, row[p * 4 * PL + k]);
                            }
                        }
                    }
                    T wx[WX];
                    {
                        T wl[WXR];
                        bin_load_row<T, WXR>(rn, wl);
#pragma unroll
                        for (int k = 0; k < WX; k++) wx[k] = wl[k];
                    }
                    const T wy = rn[OY + rowy];
#pragma unroll
                    for (int p = 0; p < NP; p++) {
                        const T wyz = wy * rn[OZ + rowz + 4 * p];
                        win_row[p].dot_acc(wx, wyz, sn);
                    }
                }
                v[2 * n] = sn.x; v[2 * n + 1] = sn.y;
            }
            int idx;
            if (FULL || nn > RND / 2) {
                const T tot = bin_halving_reduce<T, 2 * RND>(v, lane, idx);
                if ((lane & (32 / (2 * RND) - 1)) == 0) myres[idx] = tot;
            } else {
                T h[RND];
#pragma unroll
                for (int k = 0; k < RND; k++) h[k] = v[k];
                const T tot = bin_halving_reduce<T, RND>(h, lane, idx);
                if ((lane & (32 / RND - 1)) == 0) myres[idx] = tot;
            }
        };
        if (nn == RND) fold_round(std::true_type{});
        else fold_round(std::false_type{});
        __syncwarp();
        if (lane < nn) fhat[jdst] = make_float2(myres[2 * lane], myres[2 * lane + 1]);
        __syncwarp();                                                        // records and results free for the next round
    }
}
