// interp_bin.cuh -- K4 variant "register window" (kernel_mode 7, opt-in): forward interpolation of one reference tile
//   replaces toBlock!/calcOneBlock!/calcOneNode! (/root/reference/src/convolution.jl:229-344) like k_interp_row3d.
//
// The transpose of spread_bin.cuh: k_interp_row3d reads (2m)^3 complex cells of shared memory per node; here the
// nodes of the tile are counting-sorted into bins of up to G^3 first-tap positions (bin_common.cuh) with the owning
// warp in the sort key, so every warp walks a contiguous node list, 8 nodes per round: the W^3 window of a bin is
// loaded into REGISTERS when the list enters the bin (lane r owns the x-row (y, z) = (r % W, r / W), W complex cells
// per pass), every node is a register dot product with its window-aligned weights (zeros outside its 2m taps), and
// one butterfly reduction folds the 8 nodes of a round together (2 shuffles per node instead of 10).
// The tile is read-only, so bins need no colouring and no CTA barrier separates them.
#pragma once
#include "bin_common.cuh"

template <typename T, int MT, int W> struct InterpBinLayout {
    using C = typename Cplx<T>::type;
    static constexpr int L = 2 * MT;
    static constexpr int G = W - L + 1;
    static constexpr int RW = 4 * W;                       // record: wx[W] | wy[W] | wz[W] | window origin (3 ints), pad
    static constexpr int ROWS = W * W, NP = (ROWS + 31) / 32;
    static bool make(const int* bs, BinGeom& bg) { return bin_make_geom<T, MT, W>(bs, bg); }
    static size_t bytes(const BinGeom& bg)
    {
        const size_t CH = BinChunk<T>::value;
        size_t b = sizeof(C) * (size_t)bg.PNs;                               // padded tile
        b += sizeof(T) * NFFTB_BIN_WARPS * NFFTB_BIN_ROUND * RW;             // weight records (sort counters alias them)
        b += sizeof(T) * NFFTB_BIN_WARPS * 2 * NFFTB_BIN_ROUND;              // reduced results of a round
        b += sizeof(T) * 3 * CH + sizeof(int) * CH;                          // staged coordinates and destinations
        b += 2 * CH + 2 * CH + CH;                                           // order (u16), key (u16), rank (u8)
        b += 2 * (NFFTB_BIN_MAXKEYS + 8);                                    // bin_start (u16)
        return b + 16;
    }
};

// Sums K per-lane values over the warp by recursive halving: after the step with lane distance o the lanes with
// (lane & o) set keep the upper half of the values.  Returns the warp total of value `idx` (an output, a function of
// the lane); all lanes that share idx return the same total.
template <typename T, int K> __device__ __forceinline__ T bin_halving_reduce(T (&v)[K], int lane, int& idx)
{
    idx = 0;
    int c = K;
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) {
        if (c > 1) {
            const bool up = (lane & o) != 0;
            const int h = c / 2;
#pragma unroll
            for (int j = 0; j < K / 2; j++) {
                if (j < h) {
                    const T keep = up ? v[j + h] : v[j];
                    const T send = up ? v[j] : v[j + h];
                    v[j] = keep + __shfl_xor_sync(0xffffffffu, send, o);
                }
            }
            idx = 2 * idx + (up ? 1 : 0);
            c = h;
        } else {
            v[0] += __shfl_xor_sync(0xffffffffu, v[0], o);
        }
    }
    return v[0];
}

// PEER = true is the node-sharded form (comm.cu): the grid is never assembled, every z plane of the tile is staged
// from the slab of the rank that holds it (own memory or a CUDA-IPC mapping read over NVLink), like k_interp_row3d.
template <typename T, int MT, int W, bool PEER = false>
__global__ void __launch_bounds__(NFFTB_BIN_WARPS * 32, (sizeof(T) == 4 && W <= 8) ? 2 : 1)
k_interp_bin3d(const typename Cplx<T>::type* __restrict__ g, typename Cplx<T>::type* __restrict__ fhat,
               const T* __restrict__ xs, const int32_t* __restrict__ perm, const int32_t* __restrict__ items,
               int item_lo, long long M, GeomDev geo, WinDev<T> win, const __grid_constant__ PolyParam<T, MT> pp,
               BinGeom bg, const __grid_constant__ SlabTab slabs)
{
    using C = typename Cplx<T>::type;
    using IL = InterpBinLayout<T, MT, W>;
    constexpr int L = IL::L, G = IL::G, RW = IL::RW, NP = IL::NP, ROWS = IL::ROWS;
    constexpr int NWARP = NFFTB_BIN_WARPS, NTHR = NWARP * 32, CH = BinChunk<T>::value, RND = NFFTB_BIN_ROUND;
    static_assert(3 * RND <= 32, "one lane per (node, dimension)");
    static_assert(RND == 8, "the reduction handles rounds of 8 (16 values) and of up to 4 (8 values)");
    static_assert(sizeof(unsigned short) * NWARP * NFFTB_BIN_MAXKEYS <= sizeof(T) * NWARP * RND * RW, "counters alias the records");

    extern __shared__ __align__(16) unsigned char smem_raw[];
    C* P = reinterpret_cast<C*>(smem_raw);                                          // [PZ][PL] padded tile
    T* rec = reinterpret_cast<T*>(P + bg.PNs);                                      // [NWARP][RND][RW]
    unsigned short* cntw = reinterpret_cast<unsigned short*>(rec);                  // [NWARP][nkeys]  (sort only)
    T* res = rec + NWARP * RND * RW;                                                // [NWARP][2 * RND]
    T* s_x = res + NWARP * 2 * RND;                                                 // [CH][3]
    int* s_j = reinterpret_cast<int*>(s_x + 3 * CH);                                // [CH] destination (caller's node id)
    unsigned short* order = reinterpret_cast<unsigned short*>(s_j + CH);            // [CH]
    unsigned short* bin_start = order + CH;                                         // [nkeys + 1]
    unsigned short* key = bin_start + NFFTB_BIN_MAXKEYS + 8;                        // [CH] (warp, turn) of the node's bin
    unsigned char* rnk = reinterpret_cast<unsigned char*>(key + CH);                // [CH]

    const int32_t* item = items + 3 * (size_t)(item_lo + blockIdx.x);
    const int tile_id = item[0];
    const int n_lo = item[1], n_hi = item[2];
    const int tx = tile_id % geo.nb[0];
    const int ty = (tile_id / geo.nb[0]) % geo.nb[1];
    const int tz = tile_id / (geo.nb[0] * geo.nb[1]);
    const int cx0 = tx * geo.bs[0], cy0 = ty * geo.bs[1], cz0 = tz * geo.bs[2];     // first core cell
    const int PX = geo.bs[0] + L, PY = geo.bs[1] + L, PZ = geo.bs[2] + L;
    const int PXp = bg.PXp, PL = bg.PL, nkeys = bg.ikeys, turns = bg.ikeys / NWARP;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (!PEER) g += (long long)blockIdx.y * geo.gsz;
    fhat += (long long)blockIdx.y * M;
    T* myrec = rec + warp * RND * RW;
    T* myres = res + warp * 2 * RND;
    NFFTB_EMU_ALIGNED(P, 16); NFFTB_EMU_ALIGNED(rec, 16); NFFTB_EMU_ALIGNED(s_x, sizeof(T)); NFFTB_EMU_ALIGNED(s_j, 4);
    NFFTB_EMU_ALIGNED(order, 2);
    if (n_hi <= n_lo) return;

    // ---- toBlock!: the padded tile (periodic wrap per row), asynchronously; the first chunk is staged and sorted
    //      while the copies are in flight
    {
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
                if (on0) bin_copy_cell_async(dst, src + xg0);
                if (on1) bin_copy_cell_async(dst + 32, src + xg1);
            }
        }
    }

    int rowy[NP], rowz[NP];
#pragma unroll
    for (int p = 0; p < NP; p++) { const int r = lane + 32 * p; rowy[p] = r % W; rowz[p] = r / W; }
    const int wn = lane / 3, wd = lane - 3 * wn;
    const int wNt = wd == 0 ? geo.Nt[0] : (wd == 1 ? geo.Nt[1] : geo.Nt[2]);
    const int wc0 = wd == 0 ? cx0 : (wd == 1 ? cy0 : cz0);

    for (int cbase = n_lo; cbase < n_hi; cbase += CH) {
        const int nc = min(CH, n_hi - cbase);
        for (int q = threadIdx.x; q < NWARP * nkeys; q += NTHR) cntw[q] = 0;
        for (int q = threadIdx.x; q < nc; q += NTHR) {
            const long long i = (long long)cbase + q;
            const T x0 = xs[i * 3 + 0], x1 = xs[i * 3 + 1], x2 = xs[i * 3 + 2];
            T ks;
            const int b0 = bin_of<W, G>(node_cell<T>(x0, geo.Nt[0], ks) - cx0);
            const int b1 = bin_of<W, G>(node_cell<T>(x1, geo.Nt[1], ks) - cy0);
            const int b2 = bin_of<W, G>(node_cell<T>(x2, geo.Nt[2], ks) - cz0);
            // bins are dealt to the warps round-robin; the key (warp, turn) makes the nodes of one warp contiguous
            const int kb = (b2 * bg.nbin[1] + b1) * bg.nbin[0] + b0;
            key[q] = (unsigned short)((kb % NWARP) * turns + kb / NWARP);
            s_x[q * 3 + 0] = x0; s_x[q * 3 + 1] = x1; s_x[q * 3 + 2] = x2;
            s_j[q] = perm[i];
        }
        __syncthreads();
        bin_sort_chunk<CH, NWARP>(nc, nkeys, key, rnk, cntw, bin_start, order);
        if (cbase == n_lo) { bin_copy_wait(); __syncthreads(); }            // tile resident

        // This warp's nodes are contiguous in `order`, bin after bin.  The warp walks that list RND nodes at a time: one
        // evaluation of the weight records, one register dot product per node -- the window registers are reloaded
        // whenever a node's window origin differs from the resident one, i.e. at the first node of every bin -- and
        // one butterfly reduction and one store per round, however many sparse bins the round covers.
        const int wl0 = bin_start[warp * turns], wl1 = bin_start[(warp + 1) * turns];
        BinRow<T, W> win_row[NP];
        int c0 = -1, c1 = -1, c2 = -1;                                      // origin of the resident window
        for (int rbase = wl0; rbase < wl1; rbase += RND) {
            const int nn = min(RND, wl1 - rbase);
            if (wn < nn) {                                                  // weights of (node wn of the round, dimension wd)
                const int q = order[rbase + wn];
                T ks;
                const int c = node_cell<T>(s_x[q * 3 + wd], wNt, ks);
                T w[L];
                eval_taps<T, MT>(win, pp, ks, c, w);
                const int lc = c - wc0;                                     // first tap at padded coordinate lc + 1
                const int wo = 1 + bin_first<W, G>(bin_of<W, G>(lc));
                const int dl = lc + 1 - wo;                                 // first tap inside the window, in [0, G)
                reinterpret_cast<int*>(myrec + wn * RW + 3 * W)[wd] = wo;
                T* rn = myrec + wn * RW + wd * W;                           // 2m taps at [dl, dl + 2m), zeros elsewhere
#pragma unroll
                for (int l = 0; l < L; l++) rn[dl + l] = w[l];
#pragma unroll
                for (int j = 0; j < W - L; j++) rn[j < dl ? j : j + L] = (T)0;
            }
            __syncwarp();
            T v[2 * RND];
#pragma unroll
            for (int n = 0; n < RND; n++) {
                C sn = make_c<T>(0, 0);
                if (n < nn) {                                               // warp-uniform
                    const T* rn = myrec + n * RW;
                    const int* ro = reinterpret_cast<const int*>(rn + 3 * W);
                    const int o0 = ro[0], o1 = ro[1], o2 = ro[2];           // window origin, padded-tile coordinates
                    if (o0 != c0 || o1 != c1 || o2 != c2) {                 // warp-uniform: first node of a bin
                        c0 = o0; c1 = o1; c2 = o2;
                        // the bin's window -> registers (cells beyond the padded tile meet zero weights only)
#pragma unroll
                        for (int p = 0; p < NP; p++) {
                            const int Y = o1 + rowy[p], Z = o2 + rowz[p];
                            const bool rowok = (NP * 32 == ROWS || lane + 32 * p < ROWS) && Y < PY && Z < PZ;
                            const C* row = P + (Z * PL + Y * PXp + o0);
                            if (o0 + W <= PX) {                             // all but the last bin of a row
#pragma unroll
                                for (int k = 0; k < W; k++) {
                                    C c = make_c<T>(0, 0);
                                    if (rowok) c = row[k];
                                    win_row[p].set(k, c);
                                }
                            } else {
#pragma unroll
                                for (int k = 0; k < W; k++) {
                                    C c = make_c<T>(0, 0);
                                    if (rowok && o0 + k < PX) c = row[k];
                                    win_row[p].set(k, c);
                                }
                            }
                        }
                    }
                    T wx[W];
                    bin_load_row<T, W>(rn, wx);
                    const T wy = rn[W + rowy[0]];                           // rowy[p] is the same for every pass when 32 % W == 0
#pragma unroll
                    for (int p = 0; p < NP; p++) {
                        if (NP * 32 == ROWS || lane + 32 * p < ROWS) {
                            const T wyp = (32 % W == 0) ? wy : rn[W + rowy[p]];
                            const T wyz = wyp * rn[2 * W + rowz[p]];
                            win_row[p].dot_acc(wx, wyz, sn);
                        }
                    }
                }
                v[2 * n] = sn.x; v[2 * n + 1] = sn.y;
            }
            // fold the round: value 2n / 2n+1 = real / imaginary part of its node n
            int idx;
            if (nn > RND / 2) {
                const T tot = bin_halving_reduce<T, 2 * RND>(v, lane, idx);
                if ((lane & (32 / (2 * RND) - 1)) == 0) myres[idx] = tot;
            } else {
                T h[RND];
#pragma unroll
                for (int k = 0; k < RND; k++) h[k] = v[k];
                const T tot = bin_halving_reduce<T, RND>(h, lane, idx);
                if ((lane & (32 / RND - 1)) == 0) myres[idx] = tot;
            }
            __syncwarp();
            if (lane < nn) fhat[s_j[order[rbase + lane]]] = make_c<T>(myres[2 * lane], myres[2 * lane + 1]);
            __syncwarp();                                                   // records and results free for the next round
        }
        __syncthreads();                                                    // staging arrays free for the next chunk
    }
}
