// twod_batch.cuh -- batch-stationary 2-D kernels (Float32, ntransforms >= 8; BASELINE config C3: radial 512^2, 32 coils).
//   K3/K4 for a batched plan: replaces the per-transform launches of k_spread_sub2d / k_interp_row2d, i.e. the
//   reference's loop over the trailing dimension around convolve!/convolve_transpose!
//   (/root/reference/src/convolution.jl:229-492 applied to every slice of the batched arrays, NFFT.jl:mul! with dims).
//
// The per-transform kernels evaluate the 2 x 2m window taps of every node once per transform (32 times for C3) and
// keep one padded tile in shared memory.  Here one CTA keeps the padded tile of a whole group of transforms resident
// (planes[t][y][x]; 16 x 16 tiles: 24^2 cells x 32 transforms = 150 KB) and walks the tile's nodes ONCE, warp-specialised:
//   * producer warps evaluate the taps of the next chunk of 64 nodes (lane per node), look up the nodes' caller indices
//     and -- adjoint -- gather the chunk's coefficients, into the other half of a double / triple buffer;
//   * consumer warps own transforms, so warps never share a plane: nothing to colour, no atomics, fixed summation order.
// What the launcher (twod.cu) uses:
//   adjoint : k_spread_win2d    register windows (16 columns x 8 exact rows) over the plan-time bin order of
//                               sort.cu: k_bin_order2d, 16 transforms per CTA, two CTAs per SM; the planes go to the
//                               per-(transform, work item) scratch layout the 2-D gather kernels read
//   forward : k_interp_batch2d  lane (tap column, transform), 8 nodes per halving butterfly, 32 / 16 / 8 transforms per CTA
//   k_spread_batch2d            the first adjoint form (read-modify-write of the planes per node): kernel_mode 7 and tiles
//                               the windows cannot take; kept because it is the simplest statement of the design
// The measured sequence of forms is in DESIGN.md 3.6 and in the comments above each kernel.
#pragma once

constexpr int TB_NCH = 64;            // nodes per chunk
constexpr int TB_WARPS = 9;           // 8 consumers + 1 producer
constexpr int TB_THREADS = TB_WARPS * 32;

template <int MT> struct Batch2 {
    static constexpr int L = 2 * MT;
    static constexpr int REC = 2 * L;                       // wx[L] | wy[L] per node
    int PX, PY, PXp, PP;
    __host__ __device__ static int pad8(int q) { while ((q & 15) != 8) q++; return q; }
    __host__ __device__ Batch2(const int* bs)
    {
        PX = bs[0] + L; PY = bs[1] + L;
        PXp = pad8(bs[0] + 8);                              // rows r, r + 1 of one transform: 64 bytes apart modulo 128; room for 8 columns from any first tap
        PP = pad8(PXp * PY);                                // transforms t, t + 1: likewise
    }
    // planes | 2 x (records, base, caller index) | 2 x values / results
    __host__ __device__ size_t bytes(int BC) const
    {
        return sizeof(float2) * (size_t)PP * BC + 2 * (sizeof(float) * TB_NCH * 16 + 2 * sizeof(int) * TB_NCH) +
               2 * sizeof(float2) * (size_t)TB_NCH * BC;
    }
};

// producer: taps, tile-local base and caller index of the nodes [cbase, cbase + nc) of the work item
template <int MT>
__device__ __forceinline__ void tb_stage_nodes(const float* __restrict__ xs, const int32_t* __restrict__ perm, long long n_lo, int stride,
                                               int cbase, int nc, int lane, int cx0, int cy0, int PXp, const GeomDev& geo,
                                               const WinDev<float>& win, const PolyParam<float, MT>& pp, float* rec, int* base, int* pidx)
{
    constexpr int L = 2 * MT;
    for (int n = lane; n < nc; n += 32) {
        const long long i = n_lo + (long long)(cbase + n) * stride;
        float ks0, ks1;
        const int c0 = node_cell<float>(xs[i * 2 + 0], geo.Nt[0], ks0);
        const int c1 = node_cell<float>(xs[i * 2 + 1], geo.Nt[1], ks1);
        float w0[L], w1[L];
        eval_taps<float, MT>(win, pp, ks0, c0, w0);
        eval_taps<float, MT>(win, pp, ks1, c1, w1);
        base[n] = (c1 - cy0 + 1) * PXp + (c0 - cx0 + 1);    // first tap, padded-tile coordinates
        pidx[n] = perm[i];
        float* dst = rec + n * (2 * L);
#pragma unroll
        for (int k = 0; k < L; k++) { dst[k] = w0[k]; dst[L + k] = w1[k]; }
    }
}

// ------------------------------------------------------------------------------------------------------
template <int MT, int TPW>
__global__ void __launch_bounds__(TB_THREADS, 1)
k_spread_batch2d(const float2* __restrict__ fhat, float2* __restrict__ scratch, const float* __restrict__ xs,
                 const int32_t* __restrict__ perm, const int32_t* __restrict__ items, const int32_t* __restrict__ item_stride,
                 int item_lo, long long M, int B, GeomDev geo, WinDev<float> win, const __grid_constant__ PolyParam<float, MT> pp)
{
    using C = float2;
    constexpr int L = 2 * MT, REC = 2 * L, BC = 8 * TPW, RS = 4 / TPW, NCH = TB_NCH;
    static_assert(L <= 8, "one tap column per lane of an 8-lane group");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const Batch2<MT> lay(geo.bs);
    const int PX = lay.PX, PY = lay.PY, PXp = lay.PXp, PP = lay.PP;
    C* planes = reinterpret_cast<C*>(smem_raw);                                  // [BC][PP]
    C* val = planes + (size_t)PP * BC;                                           // [2][BC][NCH]
    float* rec = reinterpret_cast<float*>(val + 2 * BC * NCH);                   // [2][NCH][REC]
    int* base = reinterpret_cast<int*>(rec + 2 * NCH * REC);                     // [2][NCH]
    int* pidx = base + 2 * NCH;                                                  // [2][NCH]

    const int32_t* item = items + 3 * (size_t)(item_lo + blockIdx.x);
    const int tile_id = item[0];
    const long long n_lo = item[1];
    const int n_hi = item[2];
    const int stride = item_stride[item_lo + blockIdx.x];
    const int n_item = (int)((n_hi - n_lo + stride - 1) / stride);
    const int tx = tile_id % geo.nb[0], ty = tile_id / geo.nb[0];
    const int cx0 = tx * geo.bs[0], cy0 = ty * geo.bs[1];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int b0 = blockIdx.y * BC, nbt = min(BC, B - b0);
    const int nch = (n_item + NCH - 1) / NCH;

    auto stage = [&](int k) {                                                     // producer warp: chunk k -> buffer k & 1
        const int buf = k & 1, cbase = k * NCH, nc = min(NCH, n_item - cbase);
        tb_stage_nodes<MT>(xs, perm, n_lo, stride, cbase, nc, lane, cx0, cy0, PXp, geo, win, pp, rec + buf * NCH * REC,
                           base + buf * NCH, pidx + buf * NCH);
        __syncwarp();
        C* vb = val + buf * BC * NCH;
        const int* pb = pidx + buf * NCH;
#pragma unroll 4
        for (int idx = lane; idx < BC * NCH; idx += 32) {
            const int n = idx & (NCH - 1), t = idx / NCH;
            C v = make_float2(0.f, 0.f);
            if (n < nc && t < nbt) v = fhat[(long long)(b0 + t) * M + pb[n]];
            vb[idx] = v;
        }
    };

    {
        uint4* z = reinterpret_cast<uint4*>(planes);
        const int n16 = (int)((sizeof(C) * (size_t)PP * BC) / 16);
        for (int q = threadIdx.x; q < n16; q += TB_THREADS) z[q] = make_uint4(0, 0, 0, 0);
    }
    if (warp == 8) stage(0);
    __syncthreads();

    const int i = lane & 7, rest = lane >> 3;
    const int tl = warp * TPW + rest / RS, r0 = rest % RS;                       // my transform (CTA-local), my first row
    C* plane = planes + (size_t)(warp < 8 ? tl : 0) * PP + i;
    for (int k = 0; k < nch; k++) {
        if (warp == 8) {
            if (k + 1 < nch) stage(k + 1);
        } else {
            const int buf = k & 1, nc = min(NCH, n_item - k * NCH);
            const float* rb = rec + buf * NCH * REC;
            const int* bb = base + buf * NCH;
            const C* vb = val + (buf * BC + tl) * NCH;
            for (int n = 0; n < nc; n++) {
                const float* rw = rb + n * REC;
                const float wxi = i < L ? rw[i] : 0.f;
                const C v = vb[n];
                const C a = make_float2(wxi * v.x, wxi * v.y);
                C* p = plane + bb[n];
                if (i < L) {
#pragma unroll
                    for (int y = r0; y < L; y += RS) {
                        const float wy = rw[L + y];
                        C c = p[y * PXp];
                        c = __ffma2_rn(make_float2(wy, wy), a, c);
                        p[y * PXp] = c;
                    }
                }
                __syncwarp();                                                    // the next node may touch the same cells from other lanes
            }
        }
        __syncthreads();
    }
    // planes -> scratch [transform][work item][PY][PX] (what k_gather_tiles2d reads)
    const size_t PN = (size_t)PX * PY;
    const unsigned inv_px = fastdiv_inv(PX);
    for (int t = 0; t < nbt; t++) {
        C* dst = scratch + ((size_t)(b0 + t) * gridDim.x + blockIdx.x) * PN;
        const C* src = planes + (size_t)t * PP;
        for (int q = threadIdx.x; q < (int)PN; q += TB_THREADS) {
            const int y = (int)fastdiv(q, inv_px), x = q - y * PX;
            dst[q] = src[y * PXp + x];
        }
    }
}

// ------------------------------------------------------------------------------------------------------
// Adjoint, register-window form (the one the launcher uses when the tile is at least 8 x 8 and at most 32 x 32 cells).
// Measured on C3 with the read-modify-write form above: 4.99 ms, 83 issue slots and 48 shared-memory wavefronts per
// (node, warp), the consumers waiting at the chunk barrier for the single producer warp's coefficient gather.  Here
//   * the nodes of a tile arrive grouped by bin = (8 cells along x) x (one row of cells) (sort.cu: k_bin_order2d); all
//     taps of a bin's nodes lie in one window of 16 columns x 8 rows, which lane (i, t) of a consumer warp keeps in
//     registers as column i of transform t (8 float2).  A node is 8 packed FMAs against its 2m row weights and its
//     zero-padded 16-tap column weight -- no shared-memory traffic except its record (5 wavefronts) -- and the window
//     meets shared memory once per bin.  (The first version used 16 x 16 windows over 8 x 8-cell bins: 16 FMAs per
//     node and lane, half of them against zero weights; FFMA2 issues at half rate, so that form was bound by the FP32
//     pipe: 1.70 ms on C3.);
//   * a warp owns 2 transforms, a CTA 16: the planes take 75 KB, two CTAs share an SM (16 consumer warps);
//   * two producer warps split the taps and the coefficient gather of the next chunk (16 loads in flight per lane).
constexpr int TW_CONS = 8, TW_PROD = 2, TW_THREADS = (TW_CONS + TW_PROD) * 32, TW_BC = 2 * TW_CONS, TW_REC = 28;
constexpr int TW_WX = 16, TW_WY = 8;   // window: columns (lanes), rows (registers)

template <int MT> struct Win2 {
    static size_t bytes(const Batch2<MT>& lay)
    {
        return sizeof(float2) * (size_t)lay.PP * TW_BC + 2 * sizeof(float2) * (size_t)TB_NCH * TW_BC +
               3 * (sizeof(float) * TB_NCH * TW_REC + sizeof(int) * TB_NCH);
    }
};

template <int MT>
__global__ void __launch_bounds__(TW_THREADS, 2)
k_spread_win2d(const float2* __restrict__ fhat, float2* __restrict__ scratch, const float* __restrict__ xs2,
               const int32_t* __restrict__ perm2, const int32_t* __restrict__ items, const int32_t* __restrict__ item_stride,
               int item_lo, long long M, int B, GeomDev geo, WinDev<float> win, const __grid_constant__ PolyParam<float, MT> pp)
{
    using C = float2;
    constexpr int L = 2 * MT, BC = TW_BC, NCH = TB_NCH, REC = TW_REC, WW = TW_WX, WY = TW_WY;
    static_assert(L <= 8, "taps of an 8-cell bin must fit the 16-cell window");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const Batch2<MT> lay(geo.bs);
    const int PX = lay.PX, PY = lay.PY, PXp = lay.PXp, PP = lay.PP;
    C* planes = reinterpret_cast<C*>(smem_raw);                                  // [BC][PP]
    C* val = planes + (size_t)PP * BC;                                           // [2][BC][NCH]
    float* rec = reinterpret_cast<float*>(val + 2 * BC * NCH);                   // [3][NCH][REC]: wx16 | wy8 | bin | window origin | pad
    int* pidx = reinterpret_cast<int*>(rec + 3 * NCH * REC);                     // [3][NCH]

    const int32_t* item = items + 3 * (size_t)(item_lo + blockIdx.x);
    const int tile_id = item[0];
    const long long n_lo = item[1];
    const int n_hi = item[2];
    const int stride = item_stride[item_lo + blockIdx.x];
    const int n_item = (int)((n_hi - n_lo + stride - 1) / stride);
    const int tx = tile_id % geo.nb[0], ty = tile_id / geo.nb[0];
    const int cx0 = tx * geo.bs[0], cy0 = ty * geo.bs[1];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int b0 = blockIdx.y * BC, nbt = min(BC, B - b0);
    const int nch = (n_item + NCH - 1) / NCH;
    const int nq0 = (geo.bs[0] + 7) >> 3;

    // Producer pipeline, one global round trip per warp and pass: warp 8 evaluates the taps (and caller indices) of chunk
    // k + 2 into record buffer (k + 2) % 3 while warp 9 gathers the coefficients of chunk k + 1 -- whose caller indices
    // were staged one pass earlier -- into value buffer (k + 1) & 1, and the consumers work on chunk k.
    auto stage_taps = [&](int k) {
        const int buf = k % 3, cbase = k * NCH, nc = min(NCH, n_item - cbase);
#pragma unroll
        for (int h = 0; h < NCH / 32; h++) {
            const int n = h * 32 + lane;
            float* dst = rec + (buf * NCH + n) * REC;
#pragma unroll
            for (int q = 0; q < (WW + WY) / 4; q++) reinterpret_cast<float4*>(dst)[q] = make_float4(0.f, 0.f, 0.f, 0.f);
            int qbin = -1, org = 0;                                              // past the item: zero weights, no window change
            if (n < nc) {
                const long long i = n_lo + (long long)(cbase + n) * stride;
                float ks0, ks1;
                const int c0 = node_cell<float>(xs2[i * 2 + 0], geo.Nt[0], ks0);
                const int c1 = node_cell<float>(xs2[i * 2 + 1], geo.Nt[1], ks1);
                float w0[L], w1[L];
                eval_taps<float, MT>(win, pp, ks0, c0, w0);
                eval_taps<float, MT>(win, pp, ks1, c1, w1);
                const int l0 = c0 - cx0, l1 = c1 - cy0;
                const int q0 = l0 >> 3;
                const int o0 = min(8 * q0, PX - WW), o1 = min(l1 + 1, PY - WY);   // window origin, padded-tile coordinates
                const int d0 = l0 + 1 - o0, d1 = l1 + 1 - o1;                    // first tap inside the window: [1, 8], [0, 8 - 2m]
#pragma unroll
                for (int l = 0; l < L; l++) { dst[d0 + l] = w0[l]; dst[WW + d1 + l] = w1[l]; }
                org = o1 * PXp + o0;
                qbin = l1 * nq0 + q0;
                pidx[buf * NCH + n] = perm2[i];
            }
            reinterpret_cast<int*>(dst)[WW + WY] = qbin;
            reinterpret_cast<int*>(dst)[WW + WY + 1] = org;
        }
    };
    // coefficient gather: all caller indices, then all global loads, then all stores -- written out so that no load
    // waits behind a shared-memory store it might alias (the first version ran its round trips one after another)
    auto stage_vals = [&](int k) {
        const int nc = min(NCH, n_item - k * NCH);
        C* vb = val + (k & 1) * BC * NCH;
        const int* pb = pidx + (k % 3) * NCH;
        constexpr int U = 16;
#pragma unroll 1
        for (int u0 = 0; u0 < BC * NCH / 32; u0 += U) {
            int pn[U];
#pragma unroll
            for (int u = 0; u < U; u++) {
                const int idx = (u0 + u) * 32 + lane;
                const int n = idx & (NCH - 1), t = idx / NCH;
                pn[u] = (n < nc && t < nbt) ? pb[n] : -1;
            }
            C vv[U];
#pragma unroll
            for (int u = 0; u < U; u++) {
                const int t = ((u0 + u) * 32 + lane) / NCH;
                vv[u] = make_float2(0.f, 0.f);
                if (pn[u] >= 0) vv[u] = fhat[(long long)(b0 + t) * M + pn[u]];
            }
#pragma unroll
            for (int u = 0; u < U; u++) vb[(u0 + u) * 32 + lane] = vv[u];
        }
    };

    {
        uint4* z = reinterpret_cast<uint4*>(planes);
        const int n16 = (int)((sizeof(C) * (size_t)PP * BC) / 16);
        for (int q = threadIdx.x; q < n16; q += TW_THREADS) z[q] = make_uint4(0, 0, 0, 0);
    }
    const int i = lane & 15, tl = 2 * warp + (lane >> 4);                        // my window column, my transform (CTA-local)
    C* plane = planes + (size_t)(warp < TW_CONS ? tl : 0) * PP + i;
    C c[WY];
    int cur = -1, curoff = 0;
    for (int k = -2; k < nch; k++) {
        if (warp == TW_CONS) {
            if (k + 2 < nch) stage_taps(k + 2);
        } else if (warp == TW_CONS + 1) {
            if (k + 1 >= 0 && k + 1 < nch) stage_vals(k + 1);
        } else if (k >= 0) {
            const int nc = min(NCH, n_item - k * NCH);
            const float* rw = rec + (k % 3) * NCH * REC;
            const C* vb = val + ((k & 1) * BC + tl) * NCH;
            auto load_node = [&](const float* r, const C* vp, float& wxi, C& v, float4 (&wy)[WY / 4], int& q, int& org) {
                wxi = r[i];
                v = *vp;
#pragma unroll
                for (int y4 = 0; y4 < WY / 4; y4++) wy[y4] = reinterpret_cast<const float4*>(r + WW)[y4];
                const int2 h = *reinterpret_cast<const int2*>(r + WW + WY);
                q = h.x; org = h.y;
            };
            auto add_node = [&](float wxi, C v, const float4 (&wy)[WY / 4], int q, int org) {
                if (q != cur && q >= 0) {                                        // warp-uniform: the next bin's window
                    if (cur >= 0) {
#pragma unroll
                        for (int y = 0; y < WY; y++) plane[curoff + y * PXp] = c[y];
                    }
                    __syncwarp();                                                // overlapping windows: another lane's column
                    cur = q; curoff = org;
#pragma unroll
                    for (int y = 0; y < WY; y++) c[y] = plane[curoff + y * PXp];
                }
                const C a = make_float2(wxi * v.x, wxi * v.y);
#pragma unroll
                for (int y4 = 0; y4 < WY / 4; y4++) {
                    c[4 * y4 + 0] = __ffma2_rn(make_float2(wy[y4].x, wy[y4].x), a, c[4 * y4 + 0]);
                    c[4 * y4 + 1] = __ffma2_rn(make_float2(wy[y4].y, wy[y4].y), a, c[4 * y4 + 1]);
                    c[4 * y4 + 2] = __ffma2_rn(make_float2(wy[y4].z, wy[y4].z), a, c[4 * y4 + 2]);
                    c[4 * y4 + 3] = __ffma2_rn(make_float2(wy[y4].w, wy[y4].w), a, c[4 * y4 + 3]);
                }
            };
            // two nodes per pass (the chunk buffer is padded to NCH with zero-weight nodes): both records are in
            // registers before the first one is applied
            for (int n = 0; n < nc; n += 2, rw += 2 * REC, vb += 2) {
                float wx0, wx1; C v0, v1; float4 wy0[WY / 4], wy1[WY / 4]; int q0, q1, g0, g1;
                load_node(rw, vb, wx0, v0, wy0, q0, g0);
                load_node(rw + REC, vb + 1, wx1, v1, wy1, q1, g1);
                add_node(wx0, v0, wy0, q0, g0);
                add_node(wx1, v1, wy1, q1, g1);
            }
        }
        __syncthreads();
    }
    if (warp < TW_CONS && cur >= 0) {
#pragma unroll
        for (int y = 0; y < WY; y++) plane[curoff + y * PXp] = c[y];
    }
    __syncthreads();
    // planes -> scratch [transform][work item][PY][PX] (what k_gather_tiles2d reads)
    const size_t PN = (size_t)PX * PY;
    const unsigned inv_px = fastdiv_inv(PX);
    for (int t = 0; t < nbt; t++) {
        C* dst = scratch + ((size_t)(b0 + t) * gridDim.x + blockIdx.x) * PN;
        const C* src = planes + (size_t)t * PP;
        for (int q = threadIdx.x; q < (int)PN; q += TW_THREADS) {
            const int y = (int)fastdiv(q, inv_px), x = q - y * PX;
            dst[q] = src[y * PXp + x];
        }
    }
}

// ------------------------------------------------------------------------------------------------------
// Forward.  First B200 measurement of the straightforward form (one node at a time, 3-stage shuffle sum per node):
// 2.14 ms on C3, 88 issue slots per (node, warp), stalls on fixed-latency dependencies with two warps per scheduler.
// This form works on 8 nodes at a time -- eight independent load / FMA chains -- and folds their partial sums over
// the 8 column lanes with one halving butterfly (14 shuffles instead of 48); the tap rows are read with
// 16-byte loads and the row pitch is a compile-time constant for 16- and 32-cell tiles (PXP; 0 = run time).
// TF_SETS sets of 8 consumer warps share the planes and take alternate groups of 8 nodes (a forward kernel has no write
// conflicts to avoid): twice the warps to hide the shared-memory latency behind.
constexpr int TF_SETS = 2, TF_CONS = 8 * TF_SETS, TF_PROD = 2, TF_THREADS = (TF_CONS + TF_PROD) * 32, TF_REC = 16;

template <int MT, int TPW, int PXP>
__global__ void __launch_bounds__(TF_THREADS, 1)
k_interp_batch2d(const float2* __restrict__ g, float2* __restrict__ fhat, const float* __restrict__ xs,
                 const int32_t* __restrict__ perm, const int32_t* __restrict__ items, const int32_t* __restrict__ item_stride,
                 int item_lo, long long M, int B, GeomDev geo, WinDev<float> win, const __grid_constant__ PolyParam<float, MT> pp)
{
    using C = float2;
    constexpr int L = 2 * MT, REC = TF_REC, BC = 8 * TPW, RS = 4 / TPW, NCH = TB_NCH;
    static_assert(L <= 8, "one tap column per lane of an 8-lane group");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const Batch2<MT> lay(geo.bs);
    const int PX = lay.PX, PY = lay.PY, PXp = PXP ? PXP : lay.PXp, PP = lay.PP;
    C* planes = reinterpret_cast<C*>(smem_raw);                                  // [BC][PP]
    C* res = planes + (size_t)PP * BC;                                           // [2][BC][NCH]
    float* rec = reinterpret_cast<float*>(res + 2 * BC * NCH);                   // [2][NCH][REC]: wx at 0, wy at 8
    int* base = reinterpret_cast<int*>(rec + 2 * NCH * REC);                     // [2][NCH]
    int* pidx = base + 2 * NCH;                                                  // [2][NCH]

    const int32_t* item = items + 3 * (size_t)(item_lo + blockIdx.x);
    const int tile_id = item[0];
    const long long n_lo = item[1];
    const int n_hi = item[2];
    const int stride = item_stride[item_lo + blockIdx.x];
    const int n_item = (int)((n_hi - n_lo + stride - 1) / stride);
    const int tx = tile_id % geo.nb[0], ty = tile_id / geo.nb[0];
    const int cx0 = tx * geo.bs[0], cy0 = ty * geo.bs[1];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int b0 = blockIdx.y * BC, nbt = min(BC, B - b0);
    const int nch = (n_item + NCH - 1) / NCH;

    auto stage = [&](int k) {                                                     // producer warps: taps of chunk k -> buffer k & 1
        const int buf = k & 1, cbase = k * NCH, nc = min(NCH, n_item - cbase);
        for (int n = (warp - TF_CONS) * 32 + lane; n < NCH; n += TF_PROD * 32) {
            float* dst = rec + (buf * NCH + n) * REC;
            if (n < nc) {
                const long long i = n_lo + (long long)(cbase + n) * stride;
                float ks0, ks1;
                const int c0 = node_cell<float>(xs[i * 2 + 0], geo.Nt[0], ks0);
                const int c1 = node_cell<float>(xs[i * 2 + 1], geo.Nt[1], ks1);
                float w0[L], w1[L];
                eval_taps<float, MT>(win, pp, ks0, c0, w0);
                eval_taps<float, MT>(win, pp, ks1, c1, w1);
                base[buf * NCH + n] = (c1 - cy0 + 1) * PXp + (c0 - cx0 + 1);     // first tap, padded-tile coordinates
                pidx[buf * NCH + n] = perm[i];
#pragma unroll
                for (int l = 0; l < 8; l++) { dst[l] = l < L ? w0[l < L ? l : 0] : 0.f; dst[8 + l] = l < L ? w1[l < L ? l : 0] : 0.f; }
            } else {                                                             // the consumers work on whole groups of 4 nodes
                base[buf * NCH + n] = PXp + 1;
#pragma unroll
                for (int q = 0; q < REC / 4; q++) reinterpret_cast<float4*>(dst)[q] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
    };
    auto write_out = [&](int k) {                                                 // producer warps: results of chunk k
        const int buf = k & 1, nc = min(NCH, n_item - k * NCH);
        const C* rb = res + buf * BC * NCH;
        const int* pb = pidx + buf * NCH;
#pragma unroll 4
        for (int idx = (warp - TF_CONS) * 32 + lane; idx < BC * NCH; idx += TF_PROD * 32) {
            const int n = idx & (NCH - 1), t = idx / NCH;
            if (n < nc && t < nbt) fhat[(long long)(b0 + t) * M + pb[n]] = rb[idx];
        }
    };

    if (warp >= TF_CONS) {
        stage(0);
    } else {                                                                     // toBlock!: the padded tiles of my transforms
        const int x0 = cx0 - MT, y0 = cy0 - MT;
        const bool fw = PX <= geo.Nt[0] && PY <= geo.Nt[1];
        const C* gb = g + (long long)b0 * geo.gsz;
        for (int xb = 0; xb < PXp; xb += 32) {                                   // one pass for tiles of up to 24 cells
            const int x = xb + lane;
            const int xg = x < PX ? wrapc(x0 + x, geo.Nt[0], fw) : -1;           // -1: pitch padding, read with zero weights
            for (int y = warp; y < PY; y += TF_CONS) {
                const C* src = gb + (size_t)wrapc(y0 + y, geo.Nt[1], fw) * geo.Nt[0] + xg;
                C* dst = planes + y * PXp + x;
                if (xg >= 0) {
                    for (int t = 0; t < nbt; t++, src += geo.gsz, dst += PP) cp_async_cell(dst, src);
                } else if (x < PXp) {
                    for (int t = 0; t < nbt; t++, dst += PP) *dst = make_float2(0.f, 0.f);
                }
            }
        }
        asm volatile("cp.async.wait_all;" ::: "memory");
    }
    __syncthreads();

    const int i = lane & 7, rest = lane >> 3;
    const int tl = (warp & 7) * TPW + rest / RS, r0 = rest % RS;
    const int set = warp >> 3;
    // transforms past the batch (tl >= nbt) compute on whatever their plane holds; their results are never written
    const C* plane = planes + (size_t)(warp < TF_CONS ? tl : 0) * PP + i;
    const bool up4 = lane & 4, up2 = lane & 2, up1 = lane & 1;
    for (int k = 0; k <= nch; k++) {
        if (warp >= TF_CONS) {
            if (k >= 1) write_out(k - 1);                                        // before buffer (k + 1) & 1 = (k - 1) & 1 is restaged
            asm volatile("bar.sync 1, %0;" ::"n"(TF_PROD * 32) : "memory");
            if (k + 1 < nch) stage(k + 1);
        } else if (k < nch) {
            const int buf = k & 1, nc = min(NCH, n_item - k * NCH);
            const float* rb = rec + (buf * NCH + 8 * set) * REC;
            const int* bb = base + buf * NCH + 8 * set;
            C* ob = res + (buf * BC + tl) * NCH;
            for (int n0 = 8 * set; n0 < nc; n0 += 8 * TF_SETS, rb += 8 * TF_SETS * REC, bb += 8 * TF_SETS) {
                C acc[8];
#pragma unroll
                for (int q = 0; q < 8; q++) {
                    const float* rw = rb + q * REC;
                    const C* p = plane + bb[q];
                    const float4 wa = reinterpret_cast<const float4*>(rw + 8)[0], wb = reinterpret_cast<const float4*>(rw + 8)[1];
                    const float wy[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
                    C a = make_float2(0.f, 0.f);
#pragma unroll
                    for (int y = r0; y < L; y += RS) a = __ffma2_rn(make_float2(wy[y], wy[y]), p[y * PXp], a);
                    const float wxi = rw[i];                                     // zero for columns i >= 2m
                    a.x *= wxi; a.y *= wxi;
#pragma unroll
                    for (int o = 8; o < 8 * RS; o <<= 1) {                       // my row split (lane bits 3, 4)
                        a.x += __shfl_xor_sync(0xffffffffu, a.x, o);
                        a.y += __shfl_xor_sync(0xffffffffu, a.y, o);
                    }
                    acc[q] = a;
                }
                // halving butterfly over the 8 column lanes: lane i ends with the sum of node n0 + i
                C h[4], e[2], f;
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const C kk = up4 ? acc[4 + j] : acc[j], ss = up4 ? acc[j] : acc[4 + j];
                    h[j].x = kk.x + __shfl_xor_sync(0xffffffffu, ss.x, 4); h[j].y = kk.y + __shfl_xor_sync(0xffffffffu, ss.y, 4);
                }
#pragma unroll
                for (int j = 0; j < 2; j++) {
                    const C kk = up2 ? h[2 + j] : h[j], ss = up2 ? h[j] : h[2 + j];
                    e[j].x = kk.x + __shfl_xor_sync(0xffffffffu, ss.x, 2); e[j].y = kk.y + __shfl_xor_sync(0xffffffffu, ss.y, 2);
                }
                {
                    const C kk = up1 ? e[1] : e[0], ss = up1 ? e[0] : e[1];
                    f.x = kk.x + __shfl_xor_sync(0xffffffffu, ss.x, 1); f.y = kk.y + __shfl_xor_sync(0xffffffffu, ss.y, 1);
                }
                if (r0 == 0 && n0 + i < nc) ob[n0 + i] = f;
            }
        }
        __syncthreads();
    }
}

// A register-window forward kernel (the windows of k_spread_win2d, read-only, with a halving butterfly over the 16 column
// lanes shared by 4 nodes) was built and measured on C3: 1.25 ms against 1.06 ms for k_interp_batch2d.  It takes the
// shared-memory traffic away but needs 20 issue slots per (node, transform) instead of 10 -- the column reduction over
// 16 lanes and the half-empty 16-tap column weights -- so the direct form stays.
// A row-window forward kernel (lane (tap row, transform) holds 16 columns of the bin's window: exact rows, 3-stage
// reduction shared by 8 nodes, records only in shared memory) was measured as well: 1.05 ms, 53 issue slots per
// (node, warp) at 1.5 IPC -- no better than the direct form, which needs neither the bin order nor a second layout.
