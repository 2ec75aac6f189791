// emu_bin_kernels.cpp -- TEST INFRASTRUCTURE: runs the opt-in kernel_mode-7 kernels on the host through simt_emu.h:
// k_spread_bin3d (nfft.jl_b200/csrc/spread_bin.cuh), whose every padded tile is compared with a direct evaluation
// (same window weights, double accumulation), and k_interp_bin3d (csrc/interp_bin.cuh), whose every fHat[j] is
// compared with the direct sum over the periodically wrapped grid.  Covers: several bins per warp and colour, bins with more
// nodes than one weight round, tiles with more nodes than one staged chunk, split work items, an empty work item,
// partial last tiles, thin tiles, ntransforms > 1, Float32 and Float64, m = 2, 3, 4, the three window evaluation modes.  Each case runs twice and the two
// results must be bit-identical (the summation order may not depend on thread scheduling).
#include <cuda_runtime.h>

#include <cstdio>
#include <random>

#include "simt_emu.h"
inline size_t __cvta_generic_to_shared(const void* p) { return (size_t)p; }
#include "../../nfft.jl_b200/csrc/spread_bin.cuh"
#include "../../nfft.jl_b200/csrc/interp_bin.cuh"

template <typename T, int MT, int W>
static int run_case(const char* name, const int Nt[3], const int bs[3], int M, int cluster, int B, unsigned seed,
                    int mode = NFFTB200_POLYNOMIAL)
{
    using C = typename Cplx<T>::type;
    using BL = BinLayout<T, MT, W>;
    constexpr int L = 2 * MT, deg = L + 1;
    GeomDev geo{};
    geo.D = 3;
    geo.gsz = 1;
    for (int d = 0; d < 3; d++) {
        geo.Nt[d] = Nt[d]; geo.N[d] = Nt[d] / 2; geo.bs[d] = bs[d]; geo.nb[d] = (Nt[d] + bs[d] - 1) / bs[d];
        geo.inv_bs[d] = (unsigned)((0x100000000ull + bs[d] - 1) / bs[d]);
        geo.gsz *= Nt[d];
    }
    BinGeom bg;
    if (!BL::make(geo.bs, bg)) { printf("%s: layout does not apply\n", name); return 1; }
    const size_t smem = BL::bytes(bg);
    const int deg_conf = bin_conflict_degree(W, (int)sizeof(C), bg.PXp, bg.PL);
    if (smem > 227 * 1024) {       // the launcher falls back to the default kernel in this case
        printf("%s: SKIP, %zu bytes of shared memory (launcher returns -1)\n", name, smem);
        return 0;
    }

    std::mt19937 rng(seed);
    std::uniform_real_distribution<double> U(0.0, 1.0);
    // nodes in [0,1): uniform, plus a cluster inside a 5x5x5-cell corner of tile 0 (dense bins, several chunks)
    std::vector<T> x((size_t)3 * M);
    for (int i = 0; i < M; i++)
        for (int d = 0; d < 3; d++) {
            double v = U(rng);
            if (i < cluster) v = v * 5.0 / Nt[d];
            if (i == cluster) v = 0.0;                                   // a node exactly on the first cell
            if (i == cluster + 1) v = std::nextafter((T)1, (T)0);        // and one in the very last cell
            x[(size_t)3 * i + d] = (T)v;
        }
    auto tile_of = [&](int i) {
        int t[3];
        for (int d = 0; d < 3; d++) { T ks; t[d] = node_cell<T>(x[(size_t)3 * i + d], Nt[d], ks) / bs[d]; }
        return (t[2] * geo.nb[1] + t[1]) * geo.nb[0] + t[0];
    };
    std::vector<int32_t> perm(M);
    for (int i = 0; i < M; i++) perm[i] = i;
    std::stable_sort(perm.begin(), perm.end(), [&](int a, int b) { return tile_of(a) < tile_of(b); });
    std::vector<T> xs((size_t)3 * M);
    for (int i = 0; i < M; i++)
        for (int d = 0; d < 3; d++) xs[(size_t)3 * i + d] = x[(size_t)3 * perm[i] + d];
    // work items: one per non-empty tile; tiles above 700 nodes are split at an odd boundary; one empty item
    std::vector<int32_t> items;
    for (int i = 0; i < M;) {
        const int t = tile_of(perm[i]);
        int j = i;
        while (j < M && tile_of(perm[j]) == t) j++;
        if (j - i > 700) {
            const int mid = i + (j - i) / 2 + 1;
            items.insert(items.end(), {t, i, mid});
            items.insert(items.end(), {t, mid, j});
        } else items.insert(items.end(), {t, i, j});
        i = j;
    }
    items.insert(items.end(), {0, 0, 0});
    const int nitems = (int)items.size() / 3;

    std::vector<C> fhat((size_t)B * M);
    for (auto& v : fhat) { v.x = (T)(U(rng) - 0.5); v.y = (T)(U(rng) - 0.5); }
    PolyParam<T, MT> pp;
    for (int l = 0; l < L; l++)
        for (int r = 0; r < deg; r++) pp.c[l * deg + r] = (T)((U(rng) - 0.3) / (r + 1));
    WinDev<T> win{};
    win.m = MT; win.mode = mode; win.window = NFFTB200_KAISER_BESSEL;
    // LINEAR: a synthetic even lookup table of 2048 intervals (|u| in [0, m] grid units); FULL: exact Kaiser-Bessel
    const int lut = 2048;
    std::vector<T> lin(lut + 2);
    for (int i = 0; i < lut + 2; i++) { const double u = (double)i / lut; lin[i] = (T)(std::exp(-4.0 * u * u) * (1.0 + 0.1 * u)); }
    win.lin = lin.data(); win.lin_scale = lut / MT;
    win.b = (T)(3.141592653589793 * (2.0 - 1.0 / 2.0));

    const int PX = bs[0] + L, PY = bs[1] + L, PZ = bs[2] + L;
    const size_t PN = (size_t)PX * PY * PZ;
    std::vector<C> out[2];
    for (int rep = 0; rep < 2; rep++) {
        out[rep].assign(PN * nitems * B, C{(T)777, (T)777});
        C* scratch = out[rep].data();
        emu::launch(emu::Dim3{(unsigned)nitems, (unsigned)B, 1}, NFFTB_BIN_WARPS * 32, [&] {
            k_spread_bin3d<T, MT, W>(fhat.data(), scratch, xs.data(), perm.data(), items.data(), 0, (long long)M, geo, win, pp, bg);
        });
    }
    if (std::memcmp(out[0].data(), out[1].data(), sizeof(C) * out[0].size()) != 0) {
        printf("%s: FAIL, two runs differ bitwise\n", name);
        return 1;
    }
    // direct evaluation
    double worst = 0, scale = 0;
    std::vector<double> er(PN), ei(PN);
    for (int b = 0; b < B; b++)
        for (int it = 0; it < nitems; it++) {
            std::fill(er.begin(), er.end(), 0.0);
            std::fill(ei.begin(), ei.end(), 0.0);
            const int t = items[3 * it], lo = items[3 * it + 1], hi = items[3 * it + 2];
            const int c0[3] = {(t % geo.nb[0]) * bs[0], ((t / geo.nb[0]) % geo.nb[1]) * bs[1], (t / (geo.nb[0] * geo.nb[1])) * bs[2]};
            for (int i = lo; i < hi; i++) {
                T w[3][L];
                int s[3];
                for (int d = 0; d < 3; d++) {
                    T ks;
                    const int c = node_cell<T>(xs[(size_t)3 * i + d], Nt[d], ks);
                    eval_taps<T, MT>(win, pp, ks, c, w[d]);
                    s[d] = c - c0[d] + 1;
                }
                const C v = fhat[(size_t)b * M + perm[i]];
                for (int l2 = 0; l2 < L; l2++)
                    for (int l1 = 0; l1 < L; l1++)
                        for (int l0 = 0; l0 < L; l0++) {
                            const size_t cell = ((size_t)(s[2] + l2) * PY + (s[1] + l1)) * PX + (s[0] + l0);
                            const double ww = (double)w[0][l0] * (double)w[1][l1] * (double)w[2][l2];
                            er[cell] += ww * v.x; ei[cell] += ww * v.y;
                        }
            }
            const C* got = out[0].data() + ((size_t)b * nitems + it) * PN;
            for (size_t q = 0; q < PN; q++) {
                worst = std::max(worst, std::max(std::fabs(got[q].x - er[q]), std::fabs(got[q].y - ei[q])));
                scale = std::max(scale, std::max(std::fabs(er[q]), std::fabs(ei[q])));
            }
        }
    // ---- forward interpolation from a random grid (B grids), same nodes / items / weights
    double iworst = 0, iscale = 0;
    bool ideterm = true;
    {
        using IL = InterpBinLayout<T, MT, W>;
        BinGeom ig;
        IL::make(geo.bs, ig);
        if (IL::bytes(ig) > 227 * 1024) { printf("%s: interp layout too large\n", name); return 1; }
        std::vector<C> grid((size_t)B * geo.gsz);
        for (auto& v : grid) { v.x = (T)(U(rng) - 0.5); v.y = (T)(U(rng) - 0.5); }
        std::vector<C> fo[2];
        for (int rep = 0; rep < 2; rep++) {
            fo[rep].assign((size_t)B * M, C{(T)777, (T)777});
            C* dst = fo[rep].data();
            emu::launch(emu::Dim3{(unsigned)nitems, (unsigned)B, 1}, NFFTB_BIN_WARPS * 32, [&] {
                k_interp_bin3d<T, MT, W>(grid.data(), dst, xs.data(), perm.data(), items.data(), 0, (long long)M, geo, win, pp, ig, SlabTab{});
            });
        }
        ideterm = std::memcmp(fo[0].data(), fo[1].data(), sizeof(C) * fo[0].size()) == 0;
        // node-sharded form: the first grid cut into z-slabs that live in separate allocations ("ranks"); must
        // reproduce the single-grid result of transform 0 bit for bit
        {
            const int nr = Nt[2] % 4 == 0 ? 4 : 2;
            SlabTab st{};
            st.n = nr; st.planes = Nt[2] / nr;
            const size_t slab_cells = (size_t)st.planes * Nt[1] * Nt[0];
            std::vector<std::vector<C>> slab(nr);
            for (int r = 0; r < nr; r++) {
                slab[r].assign(grid.begin() + r * slab_cells, grid.begin() + (r + 1) * slab_cells);
                st.base[r] = slab[r].data();
            }
            std::vector<C> fp((size_t)M, C{(T)777, (T)777});
            C* dst = fp.data();
            emu::launch(emu::Dim3{(unsigned)nitems, 1, 1}, NFFTB_BIN_WARPS * 32, [&] {
                k_interp_bin3d<T, MT, W, true>(nullptr, dst, xs.data(), perm.data(), items.data(), 0, (long long)M, geo, win, pp, ig, st);
            });
            if (std::memcmp(fp.data(), fo[0].data(), sizeof(C) * (size_t)M) != 0) { printf("%s: slab-direct interpolation differs\n", name); ideterm = false; }
        }
        for (int b = 0; b < B; b++)
            for (int i = 0; i < M; i++) {
                T w[3][L];
                int s[3];
                for (int d = 0; d < 3; d++) {
                    T ks;
                    const int c = node_cell<T>(xs[(size_t)3 * i + d], Nt[d], ks);
                    eval_taps<T, MT>(win, pp, ks, c, w[d]);
                    s[d] = c - MT + 1;
                }
                double ar = 0, ai = 0;
                for (int l2 = 0; l2 < L; l2++)
                    for (int l1 = 0; l1 < L; l1++)
                        for (int l0 = 0; l0 < L; l0++) {
                            const int gx = ((s[0] + l0) % Nt[0] + Nt[0]) % Nt[0], gy = ((s[1] + l1) % Nt[1] + Nt[1]) % Nt[1],
                                      gz = ((s[2] + l2) % Nt[2] + Nt[2]) % Nt[2];
                            const C gv = grid[(size_t)b * geo.gsz + ((size_t)gz * Nt[1] + gy) * Nt[0] + gx];
                            const double ww = (double)w[0][l0] * (double)w[1][l1] * (double)w[2][l2];
                            ar += ww * gv.x; ai += ww * gv.y;
                        }
                const C got = fo[0][(size_t)b * M + perm[i]];
                iworst = std::max(iworst, std::max(std::fabs(got.x - ar), std::fabs(got.y - ai)));
                iscale = std::max(iscale, std::max(std::fabs(ar), std::fabs(ai)));
            }
    }
    const double tol = sizeof(T) == 4 ? 2e-5 : 1e-12;
    // the searched pitches must make the window rows of a pass bank-conflict free
    const bool ok = worst <= tol * scale && scale > 0 && iworst <= tol * iscale && iscale > 0 && ideterm && deg_conf == 1;
    printf("%s: %s  items %d  spread max|err|/max|ref| %.3e  interp %.3e%s  smem %zu B  pitch (%d,%d)  conflict degree %d  bins %dx%dx%d  colours %d\n",
           name, ok ? "OK" : "FAIL", nitems, worst / scale, iworst / iscale, ideterm ? "" : " (NOT reproducible)", smem, bg.PXp, bg.PL, deg_conf, bg.nbin[0], bg.nbin[1], bg.nbin[2],
           bg.S * bg.S * bg.S);
    return ok ? 0 : 1;
}

// bin layout invariants for every tile edge 3..48: every first-tap position falls into a bin whose W-wide window
// holds all its 2m taps, bins are contiguous and ordered, and bins of one colour (index distance a multiple of S) have
// disjoint windows -- the property the barrier-separated colours of the spreader rely on
template <int MT, int W> static int check_layout()
{
    constexpr int L = 2 * MT, G = W - L + 1, S = (W + G - 1) / G;
    int bad = 0;
    for (int bs = 3; bs <= 48; bs++) {
        const int nbin = bin_of<W, G>(bs - 1) + 1;
        int prev = 0;
        for (int lc = 0; lc < bs; lc++) {
            const int b = bin_of<W, G>(lc), first = bin_first<W, G>(b), dl = lc - first;
            if (b < prev || b > prev + 1 || b >= nbin) bad++;
            if (dl < 0 || dl >= G || dl + L > W) bad++;
            prev = b;
        }
        for (int b = 0; b < nbin; b++)
            for (int c = b + S; c < nbin; c += S)
                if (bin_first<W, G>(c) - bin_first<W, G>(b) < W) bad++;
    }
    printf("bin layout m=%d W=%d (G=%d, S=%d): %s\n", MT, W, G, S, bad ? "FAIL" : "OK");
    return bad;
}

int main()
{
    int bad = 0;
    bad += check_layout<2, 8>() + check_layout<3, 8>() + check_layout<4, 10>() + check_layout<1, 8>() + check_layout<5, 12>();
    const int n32[3] = {32, 32, 32}, b16[3] = {16, 16, 16};
    const int n40[3] = {40, 32, 48};
    const int nthin[3] = {32, 32, 16}, bthin[3] = {16, 16, 8};
    bad += run_case<float, 3, 8>("f32 m=3 uniform 32^3 B=2", n32, b16, 3000, 0, 2, 1);
    bad += run_case<float, 3, 8>("f32 m=3 clustered 40x32x48 (partial last tile, chunks, split items)", n40, b16, 5000, 1800, 1, 2);
    bad += run_case<double, 3, 8>("f64 m=3 clustered 32^3", n32, b16, 2500, 700, 1, 3);
    bad += run_case<float, 2, 8>("f32 m=2 uniform 32^3", n32, b16, 2000, 0, 1, 4);
    bad += run_case<float, 4, 10>("f32 m=4 (W=10) 32^3", n32, b16, 2000, 100, 1, 5);
    bad += run_case<double, 4, 10>("f64 m=4 (W=10) 32^3", n32, b16, 1200, 0, 1, 6);
    bad += run_case<float, 3, 8>("f32 m=3 thin tiles 16x16x8", nthin, bthin, 1500, 0, 1, 7);
    // chunk-boundary cases: tile 0 holds exactly one chunk (640 nodes), one chunk + 1, and two split items around 640
    bad += run_case<float, 3, 8>("f32 m=3 tile 0 with exactly 640 nodes", n32, b16, 660, 639, 1, 12);
    bad += run_case<float, 3, 8>("f32 m=3 tile 0 with 641 nodes", n32, b16, 661, 640, 1, 13);
    bad += run_case<float, 3, 8>("f32 m=3 tile 0 with 1282 nodes (items of 642 + 640)", n32, b16, 1300, 1281, 1, 14);
    bad += run_case<double, 3, 8>("f64 m=3 tile 0 with 513 nodes (chunk 512 + 1)", n32, b16, 530, 512, 1, 15);
    bad += run_case<float, 3, 8>("f32 m=3 seven nodes", n32, b16, 7, 0, 1, 16);
    const int nodd[3] = {24, 30, 28}, bodd[3] = {12, 10, 14};
    bad += run_case<float, 3, 8>("f32 m=3 tiles 12x10x14 (not a multiple of the bin period)", nodd, bodd, 2500, 200, 1, 11);
    const int nwide[3] = {64, 32, 32}, bwide[3] = {32, 16, 16};
    bad += run_case<float, 3, 8>("f32 m=3 wide tiles 32x16x16 (two bins per warp and colour)", nwide, bwide, 3000, 300, 1, 10);
    bad += run_case<float, 3, 8>("f32 m=3 LINEAR lookup table", n32, b16, 1500, 0, 1, 8, NFFTB200_LINEAR);
    bad += run_case<double, 3, 8>("f64 m=3 FULL (exact Kaiser-Bessel)", n32, b16, 1500, 0, 1, 9, NFFTB200_FULL);
    printf(bad ? "FAILED\n" : "ALL OK\n");
    return bad ? 1 : 0;
}
