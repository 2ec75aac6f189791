// spread_bin.cuh -- K3 variant "register footprint" (kernel_mode 7, opt-in): adjoint gridding of one reference tile
//   replaces fillBlock!/fillOneNode! (/root/reference/src/convolution.jl:445-492) like k_spread_sub3d, and hands the
//   same padded tile ("blocks[l]") to the same gather pass (addBlock!, :371-443).
//
// Why: k_spread_sub3d read-modify-writes (2m)^3 complex cells of shared memory per node, so its accumulate loop is
// bound by the 128 B/clk shared-memory pipe.  Here the footprints of SEVERAL nodes are summed in
// REGISTERS first and shared memory sees one read-modify-write per bin:
//   * the nodes of the tile (already contiguous after the plan-time sort) are counting-sorted inside the CTA into
//     bins of up to G^3 consecutive first-tap positions, G = W - 2m + 1, so that every node of a bin has its (2m)^3
//     taps inside one W^3 window (W = 8 for m <= 3; per dimension the positions are cut with period W into bins of
//     3, 3, 2 for m = 3, see bin_of in bin_common.cuh);
//   * one warp accumulates one bin: lane r owns the x-row (y, z) = (r % W, r / W) of the window (W complex
//     accumulators per pass, ceil(W*W/32) passes), the node's weights are laid out on the window (zeros outside
//     the node's taps) so the inner loop is W*2 FFMA per pass and node with no address arithmetic;
//   * bins whose windows are disjoint run concurrently: bins are coloured by (index mod S) per dimension,
//     S = ceil(W/G), and the S^3 colours are separated by CTA barriers, so no atomics are needed and the
//     summation order is fixed (bit-reproducible, as in the default kernel);
//   * window weights of a bin are evaluated 8 nodes at a time, lane = (node, dimension), with the constant-bank
//     Horner of tile3d.cuh.
// FFMA work grows by W^3/(2m)^3 (2.4x for m = 3), shared-memory traffic per node falls from 2*(2m)^3 cells to
// 2*W^3/(nodes per bin) plus 5 broadcast loads ("per bin" meaning per staged chunk: a chunk is a run of the tile's
// nodes in the plan's order, so it is spread over all bins of the tile).  One padded tile instead of 8 private sub-tiles: 109 KB for
// Float32, two CTAs per SM.
//
// The file holds device code only and is also compiled for the HOST by tests/emu (one OS thread per CUDA thread),
// which checks the kernel's indexing against a direct evaluation without a GPU.
#pragma once
#include "bin_common.cuh"

template <typename T, int MT, int W> struct BinLayout {
    using C = typename Cplx<T>::type;
    static constexpr int L = 2 * MT;
    static constexpr int G = W - L + 1;
    static constexpr int RW = 4 * W + 16 / (int)sizeof(T); // record: wx[W] | wy[W] | (wz * v)[W] re, im | window origin (3 ints)
    static constexpr int ROWS = W * W, NP = (ROWS + 31) / 32;
    static_assert(G >= 1, "window narrower than the footprint");

    static bool make(const int* bs, BinGeom& bg) { return bin_make_geom<T, MT, W>(bs, bg); }
    static size_t bytes(const BinGeom& bg)
    {
        const size_t CH = BinChunk<T>::value;
        size_t b = sizeof(C) * (size_t)bg.PNs;                               // padded tile
        b += sizeof(T) * NFFTB_BIN_WARPS * NFFTB_BIN_ROUND * RW;             // weight records (sort counters alias them)
        b += sizeof(C) * CH + sizeof(T) * 3 * CH;                            // staged values and coordinates
        b += 2 * CH + 2 * CH + CH;                                           // order (u16), key (u16), rank (u8)
        b += 2 * (NFFTB_BIN_MAXKEYS + 8);                                    // bin_start (u16)
        return b + 16;
    }
};

template <typename T, int MT, int W>
__global__ void __launch_bounds__(NFFTB_BIN_WARPS * 32, (sizeof(T) == 4 && W <= 8) ? 2 : 1)
k_spread_bin3d(const typename Cplx<T>::type* __restrict__ fhat, typename Cplx<T>::type* __restrict__ scratch,
               const T* __restrict__ xs, const int32_t* __restrict__ perm, const int32_t* __restrict__ items,
               int item_lo, long long M, GeomDev geo, WinDev<T> win, const __grid_constant__ PolyParam<T, MT> pp,
               BinGeom bg)
{
    using C = typename Cplx<T>::type;
    using BL = BinLayout<T, MT, W>;
    constexpr int L = BL::L, G = BL::G, RW = BL::RW, NP = BL::NP, ROWS = BL::ROWS;
    constexpr int NWARP = NFFTB_BIN_WARPS, NTHR = NWARP * 32, CH = BinChunk<T>::value, RND = NFFTB_BIN_ROUND;
    constexpr int S = (W + G - 1) / G, NPH = S * S * S;     // colours per dimension / per tile
    static_assert(3 * RND <= 32, "one lane per (node, dimension)");
    static_assert(sizeof(unsigned short) * NWARP * NFFTB_BIN_MAXKEYS <= sizeof(T) * NWARP * RND * RW, "counters alias the records");

    extern __shared__ __align__(16) unsigned char smem_raw[];
    C* P = reinterpret_cast<C*>(smem_raw);                                          // [PZ][PL] padded tile
    T* rec = reinterpret_cast<T*>(P + bg.PNs);                                      // [NWARP][RND][RW]
    unsigned short* cntw = reinterpret_cast<unsigned short*>(rec);                  // [NWARP][nkeys]  (sort only)
    C* s_v = reinterpret_cast<C*>(rec + NWARP * RND * RW);                          // [CH]
    T* s_x = reinterpret_cast<T*>(s_v + CH);                                        // [CH][3]
    unsigned short* order = reinterpret_cast<unsigned short*>(s_x + 3 * CH);        // [CH] bin-sorted chunk-local ids
    unsigned short* bin_start = order + CH;                                         // [nkeys + 1]
    unsigned short* key = bin_start + NFFTB_BIN_MAXKEYS + 8;                        // [CH] (warp, colour, turn) of the node's bin
    unsigned char* rnk = reinterpret_cast<unsigned char*>(key + CH);                // [CH]

    const int32_t* item = items + 3 * (size_t)(item_lo + blockIdx.x);
    const int tile_id = item[0];
    const int n_lo = item[1], n_hi = item[2];
    const int tx = tile_id % geo.nb[0];
    const int ty = (tile_id / geo.nb[0]) % geo.nb[1];
    const int tz = tile_id / (geo.nb[0] * geo.nb[1]);
    const int cx0 = tx * geo.bs[0], cy0 = ty * geo.bs[1], cz0 = tz * geo.bs[2];     // first core cell
    const int PX = geo.bs[0] + L, PY = geo.bs[1] + L, PZ = geo.bs[2] + L;
    const int PXp = bg.PXp, PL = bg.PL, nkeys = bg.nkeys;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    fhat += (long long)blockIdx.y * M;
    scratch += ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * ((size_t)PX * PY * PZ);
    T* myrec = rec + warp * RND * RW;
    NFFTB_EMU_ALIGNED(P, 16); NFFTB_EMU_ALIGNED(rec, 16); NFFTB_EMU_ALIGNED(s_v, sizeof(C)); NFFTB_EMU_ALIGNED(s_x, sizeof(T));
    NFFTB_EMU_ALIGNED(order, 2); NFFTB_EMU_ALIGNED(scratch, 16);

    // per-lane constant row geometry: row r = lane + 32 * p of the W x W window
    int rowy[NP], rowz[NP];
#pragma unroll
    for (int p = 0; p < NP; p++) { const int r = lane + 32 * p; rowy[p] = r % W; rowz[p] = r / W; }
    // lane = (node, dimension) of the weight rounds
    const int wn = lane / 3, wd = lane - 3 * wn;
    const int wNt = wd == 0 ? geo.Nt[0] : (wd == 1 ? geo.Nt[1] : geo.Nt[2]);
    const int wc0 = wd == 0 ? cx0 : (wd == 1 ? cy0 : cz0);

    for (int cbase = n_lo; cbase < n_hi; cbase += CH) {
        const int nc = min(CH, n_hi - cbase);
        // ---- stage the chunk: global loads first, the tile is zeroed while they are in flight (first chunk)
        constexpr int NPT = (CH + NTHR - 1) / NTHR;
        T rx[NPT][3];
        C rv[NPT];
#pragma unroll
        for (int k = 0; k < NPT; k++) {
            const int q = threadIdx.x + k * NTHR;
            if (q < nc) {
                const long long i = (long long)cbase + q;
                rx[k][0] = xs[i * 3 + 0]; rx[k][1] = xs[i * 3 + 1]; rx[k][2] = xs[i * 3 + 2];
                rv[k] = fhat[perm[i]];
            }
        }
        if (cbase == n_lo) {
            uint4* z = reinterpret_cast<uint4*>(P);
            const int n16 = (int)((sizeof(C) * (size_t)bg.PNs) / 16);
            for (int q = threadIdx.x; q < n16; q += NTHR) z[q] = make_uint4(0, 0, 0, 0);
        }
        for (int q = threadIdx.x; q < NWARP * nkeys; q += NTHR) cntw[q] = 0;
#pragma unroll
        for (int k = 0; k < NPT; k++) {
            const int q = threadIdx.x + k * NTHR;
            if (q < nc) {
                T ks;
                const int b0 = bin_of<W, G>(node_cell<T>(rx[k][0], geo.Nt[0], ks) - cx0);
                const int b1 = bin_of<W, G>(node_cell<T>(rx[k][1], geo.Nt[1], ks) - cy0);
                const int b2 = bin_of<W, G>(node_cell<T>(rx[k][2], geo.Nt[2], ks) - cz0);
                // colour = (b mod S) per dimension; the bins of a colour are dealt to the warps round-robin ("slot"); the
                // key (warp, colour, turn) makes the nodes of one warp contiguous, in the order the warp needs them
                const int p0 = b0 % S, p1 = b1 % S, p2 = b2 % S;
                const int A0 = (bg.nbin[0] - p0 + S - 1) / S, A1 = (bg.nbin[1] - p1 + S - 1) / S;
                const int slot = ((b2 / S) * A1 + b1 / S) * A0 + b0 / S;
                key[q] = (unsigned short)(((slot % NWARP) * NPH + (p2 * S + p1) * S + p0) * bg.maxit + slot / NWARP);
                s_x[q * 3 + 0] = rx[k][0]; s_x[q * 3 + 1] = rx[k][1]; s_x[q * 3 + 2] = rx[k][2];
                s_v[q] = rv[k];
            }
        }
        __syncthreads();
        bin_sort_chunk<CH, NWARP>(nc, nkeys, key, rnk, cntw, bin_start, order);

        // ---- accumulate: S^3 colours; the bins of one colour have disjoint windows and run on different warps.
        //      This warp's nodes are contiguous in `order` (colour-major), so the weight records are evaluated RND
        //      nodes at a time along that list -- a round may run ahead into the bins of later colours, which keeps
        //      the (node, dimension) lanes busy even when a bin holds two or three nodes.
        const int wl1 = bin_start[(warp + 1) * NPH * bg.maxit];             // end of this warp's node list
        int rbase = -RND;                                                   // list index of the resident round
        for (int ph = 0; ph < NPH; ph++) {
            for (int it = 0; it < bg.maxit; it++) {
                const int kk = (warp * NPH + ph) * bg.maxit + it;
                const int lo = bin_start[kk], hi = bin_start[kk + 1];
                if (hi <= lo) continue;                                   // warp-uniform
                BinRow<T, W> acc[NP];
#pragma unroll
                for (int p = 0; p < NP; p++) acc[p].zero();
                int o0 = 0, o1 = 0, o2 = 0;                               // window origin of the bin, padded-tile coordinates
                for (int i = lo; i < hi; i++) {
                    if (i >= rbase + RND) {                               // warp-uniform: next round of records
                        __syncwarp();                                     // the previous round has been read
                        rbase = i;
                        if (wn < min(RND, wl1 - rbase)) {                 // weights of (node wn of the round, dimension wd)
                            const int q = order[rbase + wn];
                            T ks;
                            const int c = node_cell<T>(s_x[q * 3 + wd], wNt, ks);
                            T w[L];
                            eval_taps<T, MT>(win, pp, ks, c, w);
                            const int lc = c - wc0;                       // first tap at padded coordinate lc + 1
                            const int wo = 1 + bin_first<W, G>(bin_of<W, G>(lc));
                            const int dl = lc + 1 - wo;                   // first tap inside the window, in [0, G)
                            T* rn = myrec + wn * RW;
                            reinterpret_cast<int*>(rn + 4 * W)[wd] = wo;
                            // the 2m taps at window positions [dl, dl + 2m), zeros in the other W - 2m positions
                            if (wd < 2) {
#pragma unroll
                                for (int l = 0; l < L; l++) rn[wd * W + dl + l] = w[l];
#pragma unroll
                                for (int j = 0; j < W - L; j++) rn[wd * W + (j < dl ? j : j + L)] = (T)0;
                            } else {
                                const C v = s_v[q];
                                C* rz = reinterpret_cast<C*>(rn + 2 * W);
#pragma unroll
                                for (int l = 0; l < L; l++) rz[dl + l] = make_c<T>(w[l] * v.x, w[l] * v.y);
#pragma unroll
                                for (int j = 0; j < W - L; j++) rz[j < dl ? j : j + L] = make_c<T>(0, 0);
                            }
                        }
                        __syncwarp();
                    }
                    const T* rn = myrec + (i - rbase) * RW;
                    if (i == lo) {
                        const int* ro = reinterpret_cast<const int*>(rn + 4 * W);
                        o0 = ro[0]; o1 = ro[1]; o2 = ro[2];
                    }
                    T wx[W];
                    bin_load_row<T, W>(rn, wx);
#pragma unroll
                    for (int p = 0; p < NP; p++) {
                        if (NP * 32 == ROWS || lane + 32 * p < ROWS) {
                            const T wy = rn[W + rowy[p]];
                            NFFTB_EMU_ALIGNED(rn + 2 * W, sizeof(C));
                            const C vz = reinterpret_cast<const C*>(rn + 2 * W)[rowz[p]];
                            acc[p].axpy(wx, wy, vz);
                        }
                    }
                }
                // one read-modify-write of the window (cells beyond the padded tile carry zero weights only)
#pragma unroll
                for (int p = 0; p < NP; p++) {
                    const int Y = o1 + rowy[p], Z = o2 + rowz[p];
                    if ((NP * 32 == ROWS || lane + 32 * p < ROWS) && Y < PY && Z < PZ) {
                        C* row = P + (Z * PL + Y * PXp + o0);
                        if (o0 + W <= PX) {                               // warp-uniform: all but the last bin of a row
#pragma unroll
                            for (int i = 0; i < W; i++) { C c = row[i]; acc[p].add_to(c, i); row[i] = c; }
                        } else {
#pragma unroll
                            for (int i = 0; i < W; i++)
                                if (o0 + i < PX) { C c = row[i]; acc[p].add_to(c, i); row[i] = c; }
                        }
                    }
                }
            }
            __syncthreads();
        }
    }
    if (n_hi <= n_lo) {                                                   // empty item: the tile is all zeros
        uint4* z = reinterpret_cast<uint4*>(P);
        const int n16 = (int)((sizeof(C) * (size_t)bg.PNs) / 16);
        for (int q = threadIdx.x; q < n16; q += NTHR) z[q] = make_uint4(0, 0, 0, 0);
        __syncthreads();
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
            C* dst = scratch + ((size_t)row * PX + 2 * u);
            NFFTB_EMU_ALIGNED(dst, 16);
            if (sizeof(T) == 4) *reinterpret_cast<float4*>(dst) = make_float4((float)a.x, (float)a.y, (float)b.x, (float)b.y);
            else { dst[0] = a; dst[1] = b; }
        }
    } else {
        const unsigned inv_px = fastdiv_inv(PX), inv_py = fastdiv_inv(PY);
        for (int idx = threadIdx.x; idx < PZ * PY * PX; idx += NTHR) {
            const int row = (int)fastdiv(idx, inv_px), x = idx - row * PX;
            const int z = (int)fastdiv(row, inv_py), y = row - z * PY;
            scratch[idx] = P[z * PL + y * PXp + x];
        }
    }
}
