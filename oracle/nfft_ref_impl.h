/*
 * oracle/nfft_ref_impl.h -- precision-generic body of the C restatement (included twice by
 * nfft_ref.c with REAL = float / double and SUF = _f32 / _f64).
 *
 * TEST INFRASTRUCTURE ONLY (see nfft_ref.c).  Restates the reference's *blocked* CPU
 * algorithm, which is its default path (precompute=POLYNOMIAL, blocking=true):
 *   node binning         /root/reference/src/precomputation.jl:487-520  (_precomputeBlocks)
 *   idxInBlock           /root/reference/src/precomputation.jl:524-554
 *   Horner window        /root/reference/src/precomputation.jl:215-222
 *   forward  toBlock!/calcOneNode!          /root/reference/src/convolution.jl:229-344
 *   adjoint  fillOneNode!/locked addBlock!  /root/reference/src/convolution.jl:356-492
 *   deconvolve / transpose                  /root/reference/src/deconvolution.jl:22-92
 * All arrays are Julia column-major: nodes k are D x M, grids have dim 1 fastest,
 * complex numbers are interleaved (re, im).
 */

#define CAT_(a, b) a##b
#define CAT(a, b) CAT_(a, b)
#define FN(name) CAT(name, SUF)

/* shiftNodes! (src/utils.jl:32-44) on a copy + tile key + stable counting sort.
 * perm: M int64 (0-based node ids, tile-major, ascending j inside a tile)
 * blockStart: nblocks+1 prefix sums. xs: shifted nodes (D x M). Serial like the reference
 * (src/precomputation.jl:494-504 has threading commented out). */
void FN(ref_precompute_blocks)(const REAL *k, int D, int64_t M, const int64_t *Nt,
                               const int64_t *bs, int64_t *perm, int64_t *blockStart,
                               REAL *xs)
{
    int64_t nb[3] = {1, 1, 1}, nblocks = 1;
    for (int d = 0; d < D; d++) { nb[d] = (Nt[d] + bs[d] - 1) / bs[d]; nblocks *= nb[d]; }
    for (int64_t i = 0; i < (int64_t)D * M; i++) {
        REAL v = k[i];
        if (v < (REAL)0) v += (REAL)1;
        if (v == (REAL)1) v -= REAL_EPS;
        xs[i] = v;
    }
    int64_t *key = (int64_t *)malloc(sizeof(int64_t) * (size_t)(M > 0 ? M : 1));
    for (int64_t l = 0; l <= nblocks; l++) blockStart[l] = 0;
    for (int64_t j = 0; j < M; j++) {
        int64_t id = 0, stride = 1;
        for (int d = 0; d < D; d++) {
            volatile REAL ks = xs[j * D + d] * (REAL)Nt[d]; /* rounded in REAL, no contraction */
            int64_t c = (int64_t)ks;                         /* unsafe_trunc */
            id += (c / bs[d]) * stride;
            stride *= nb[d];
        }
        key[j] = id;
        blockStart[id + 1]++;
    }
    for (int64_t l = 0; l < nblocks; l++) blockStart[l + 1] += blockStart[l];
    int64_t *cur = (int64_t *)malloc(sizeof(int64_t) * (size_t)nblocks);
    memcpy(cur, blockStart, sizeof(int64_t) * (size_t)nblocks);
    for (int64_t j = 0; j < M; j++) perm[cur[key[j]]++] = j;
    free(cur);
    free(key);
}

/* _precomputeIdxInBlock (POLYNOMIAL flavour): for sorted position i and dim d:
 * y[i*D+d]  = 0-based first-tap index inside the padded tile, t[i*D+d] = frac - 1/2. */
void FN(ref_precompute_idx)(const REAL *xs, int D, int64_t M, const int64_t *Nt,
                            const int64_t *bs, int m, const int64_t *perm,
                            const int64_t *blockStart, int32_t *y, REAL *t)
{
    int64_t nb[3] = {1, 1, 1}, nblocks = 1;
    for (int d = 0; d < D; d++) { nb[d] = (Nt[d] + bs[d] - 1) / bs[d]; nblocks *= nb[d]; }
#pragma omp parallel for schedule(dynamic, 16)
    for (int64_t l = 0; l < nblocks; l++) {
        int64_t lc[3], r = l;
        for (int d = 0; d < D; d++) { lc[d] = r % nb[d]; r /= nb[d]; }
        for (int64_t i = blockStart[l]; i < blockStart[l + 1]; i++) {
            int64_t j = perm[i];
            for (int d = 0; d < D; d++) {
                volatile REAL ks = xs[j * D + d] * (REAL)Nt[d];
                int64_t off = (int64_t)ks - m + 1;                 /* 0-based first tap cell */
                int64_t bo = lc[d] * bs[d] - m;                    /* 0-based cell of padded idx 0 */
                y[i * D + d] = (int32_t)(off - bo);
                REAL a = ks - (REAL)off; a = a - (REAL)m; a = a + (REAL)1;
                t[i * D + d] = (REAL)((double)a - 0.5);
            }
        }
    }
}

static inline void FN(horner_taps)(const REAL *P, int m, REAL x, REAL *w)
{
    int L = 2 * m, deg = 2 * m + 1;
    for (int l = 0; l < L; l++) {
        const REAL *c = P + (size_t)l * deg;          /* column l of the (deg x L) matrix */
        REAL acc = c[deg - 1];
        for (int r = deg - 2; r >= 0; r--) acc = FMA(acc, x, c[r]);
        w[l] = acc;
    }
}

/* convolve! blocked (src/convolution.jl:235-344). g: Nt complex, fHat: M complex. */
void FN(ref_convolve_blocking)(const REAL *g, REAL *fHat, int D, int64_t M, const int64_t *Nt,
                               const int64_t *bs, int m, const REAL *P, const int64_t *perm,
                               const int64_t *blockStart, const int32_t *y, const REAL *t)
{
    int64_t nb[3] = {1, 1, 1}, nblocks = 1, pb[3] = {1, 1, 1}, NT[3] = {1, 1, 1};
    int L[3] = {1, 1, 1};
    for (int d = 0; d < D; d++) {
        nb[d] = (Nt[d] + bs[d] - 1) / bs[d]; nblocks *= nb[d];
        pb[d] = bs[d] + 2 * m; NT[d] = Nt[d]; L[d] = 2 * m;
    }
    size_t bsz = (size_t)(pb[0] * pb[1] * pb[2]);
#pragma omp parallel
    {
        REAL *block = (REAL *)malloc(sizeof(REAL) * 2 * bsz);
        int64_t *ix = (int64_t *)malloc(sizeof(int64_t) * (size_t)(pb[0] + pb[1] + pb[2]));
#pragma omp for schedule(dynamic, 1)
        for (int64_t l = 0; l < nblocks; l++) {
            if (blockStart[l + 1] == blockStart[l]) continue;
            int64_t lc[3] = {0, 0, 0}, r = l;
            for (int d = 0; d < D; d++) { lc[d] = r % nb[d]; r /= nb[d]; }
            int64_t *ixd[3] = {ix, ix + pb[0], ix + pb[0] + pb[1]};
            for (int d = 0; d < 3; d++)
                for (int64_t q = 0; q < pb[d]; q++) {
                    int64_t c = (d < D) ? (lc[d] * bs[d] - m + q) : 0;
                    c %= NT[d]; if (c < 0) c += NT[d];
                    ixd[d][q] = c;
                }
            /* toBlock! */
            for (int64_t q2 = 0; q2 < pb[2]; q2++)
                for (int64_t q1 = 0; q1 < pb[1]; q1++) {
                    const REAL *src = g + 2 * (size_t)((ixd[2][q2] * NT[1] + ixd[1][q1]) * NT[0]);
                    REAL *dst = block + 2 * (size_t)((q2 * pb[1] + q1) * pb[0]);
                    for (int64_t q0 = 0; q0 < pb[0]; q0++) {
                        dst[2 * q0] = src[2 * ixd[0][q0]];
                        dst[2 * q0 + 1] = src[2 * ixd[0][q0] + 1];
                    }
                }
            /* calcOneBlock! */
            for (int64_t i = blockStart[l]; i < blockStart[l + 1]; i++) {
                REAL w[3][16];
                int32_t yy[3] = {0, 0, 0};
                for (int d = 0; d < 3; d++) {
                    if (d < D) { FN(horner_taps)(P, m, t[i * D + d], w[d]); yy[d] = y[i * D + d]; }
                    else w[d][0] = (REAL)1;
                }
                REAL s3r = 0, s3i = 0;
                for (int l2 = 0; l2 < L[2]; l2++) {
                    REAL s2r = 0, s2i = 0;
                    for (int l1 = 0; l1 < L[1]; l1++) {
                        const REAL *row = block + 2 * (size_t)(((yy[2] + l2) * pb[1] + (yy[1] + l1)) * pb[0] + yy[0]);
                        REAL s1r = 0, s1i = 0;
                        for (int l0 = 0; l0 < L[0]; l0++) {
                            s1r += w[0][l0] * row[2 * l0];
                            s1i += w[0][l0] * row[2 * l0 + 1];
                        }
                        s2r += w[1][l1] * s1r; s2i += w[1][l1] * s1i;
                    }
                    s3r += w[2][l2] * s2r; s3i += w[2][l2] * s2i;
                }
                int64_t j = perm[i];
                fHat[2 * j] = s3r; fHat[2 * j + 1] = s3i;
            }
        }
        free(block);
        free(ix);
    }
}

/* convolve_transpose! blocked (src/convolution.jl:356-492): memset g, per-tile private
 * scratch, fillOneNode!, then addBlock! under ONE global lock (:371-373). */
void FN(ref_convolve_transpose_blocking)(const REAL *fHat, REAL *g, int D, int64_t M,
                                         const int64_t *Nt, const int64_t *bs, int m,
                                         const REAL *P, const int64_t *perm,
                                         const int64_t *blockStart, const int32_t *y,
                                         const REAL *t)
{
    int64_t nb[3] = {1, 1, 1}, nblocks = 1, pb[3] = {1, 1, 1}, NT[3] = {1, 1, 1};
    int L[3] = {1, 1, 1};
    size_t gsz = 1;
    for (int d = 0; d < D; d++) {
        nb[d] = (Nt[d] + bs[d] - 1) / bs[d]; nblocks *= nb[d];
        pb[d] = bs[d] + 2 * m; NT[d] = Nt[d]; L[d] = 2 * m; gsz *= (size_t)Nt[d];
    }
    memset(g, 0, sizeof(REAL) * 2 * gsz);
    size_t bsz = (size_t)(pb[0] * pb[1] * pb[2]);
#pragma omp parallel
    {
        REAL *block = (REAL *)malloc(sizeof(REAL) * 2 * bsz);
        int64_t *ix = (int64_t *)malloc(sizeof(int64_t) * (size_t)(pb[0] + pb[1] + pb[2]));
#pragma omp for schedule(dynamic, 1)
        for (int64_t l = 0; l < nblocks; l++) {
            if (blockStart[l + 1] == blockStart[l]) continue;
            memset(block, 0, sizeof(REAL) * 2 * bsz);
            for (int64_t i = blockStart[l]; i < blockStart[l + 1]; i++) {
                REAL w[3][16];
                int32_t yy[3] = {0, 0, 0};
                for (int d = 0; d < 3; d++) {
                    if (d < D) { FN(horner_taps)(P, m, t[i * D + d], w[d]); yy[d] = y[i * D + d]; }
                    else w[d][0] = (REAL)1;
                }
                int64_t j = perm[i];
                REAL vr = fHat[2 * j], vi = fHat[2 * j + 1];
                REAL iwr[16], iwi[16];
                for (int l0 = 0; l0 < L[0]; l0++) { iwr[l0] = w[0][l0] * vr; iwi[l0] = w[0][l0] * vi; }
                for (int l2 = 0; l2 < L[2]; l2++)
                    for (int l1 = 0; l1 < L[1]; l1++) {
                        REAL pw = w[2][l2] * w[1][l1];
                        REAL *row = block + 2 * (size_t)(((yy[2] + l2) * pb[1] + (yy[1] + l1)) * pb[0] + yy[0]);
                        for (int l0 = 0; l0 < L[0]; l0++) {
                            row[2 * l0] += iwr[l0] * pw;
                            row[2 * l0 + 1] += iwi[l0] * pw;
                        }
                    }
            }
            int64_t lc[3] = {0, 0, 0}, r = l;
            for (int d = 0; d < D; d++) { lc[d] = r % nb[d]; r /= nb[d]; }
            int64_t *ixd[3] = {ix, ix + pb[0], ix + pb[0] + pb[1]};
            for (int d = 0; d < 3; d++)
                for (int64_t q = 0; q < pb[d]; q++) {
                    int64_t c = (d < D) ? (lc[d] * bs[d] - m + q) : 0;
                    c %= NT[d]; if (c < 0) c += NT[d];
                    ixd[d][q] = c;
                }
#pragma omp critical(addblock)
            {
                for (int64_t q2 = 0; q2 < pb[2]; q2++)
                    for (int64_t q1 = 0; q1 < pb[1]; q1++) {
                        REAL *dst = g + 2 * (size_t)((ixd[2][q2] * NT[1] + ixd[1][q1]) * NT[0]);
                        const REAL *src = block + 2 * (size_t)((q2 * pb[1] + q1) * pb[0]);
                        for (int64_t q0 = 0; q0 < pb[0]; q0++) {
                            dst[2 * ixd[0][q0]] += src[2 * q0];
                            dst[2 * ixd[0][q0] + 1] += src[2 * q0 + 1];
                        }
                    }
            }
        }
        free(block);
        free(ix);
    }
}

/* fill! + deconvolve! (src/implementation.jl:159-160, src/deconvolution.jl:22-45) and its
 * transpose (:69-92).  lut: concatenated LUT_d (N_0 + N_1 + N_2 entries). dir=0: f->g, 1: g->f */
void FN(ref_deconvolve)(REAL *f, REAL *g, int D, const int64_t *N, const int64_t *Nt,
                        const REAL *lut, int dir)
{
    int64_t n[3] = {1, 1, 1}, nt[3] = {1, 1, 1};
    const REAL *lt[3] = {lut, lut, lut};
    size_t gsz = 1;
    for (int d = 0; d < D; d++) { n[d] = N[d]; nt[d] = Nt[d]; gsz *= (size_t)Nt[d]; }
    if (D > 1) lt[1] = lut + N[0];
    if (D > 2) lt[2] = lut + N[0] + N[1];
    if (dir == 0) {
#pragma omp parallel for
        for (int64_t q = 0; q < (int64_t)(2 * gsz); q++) g[q] = 0;
    }
#pragma omp parallel for collapse(2)
    for (int64_t i2 = 0; i2 < n[2]; i2++)
        for (int64_t i1 = 0; i1 < n[1]; i1++) {
            int64_t u2 = (D > 2) ? ((i2 - n[2] / 2) % nt[2] + nt[2]) % nt[2] : 0;
            int64_t u1 = (D > 1) ? ((i1 - n[1] / 2) % nt[1] + nt[1]) % nt[1] : 0;
            REAL *fr = f + 2 * (size_t)((i2 * n[1] + i1) * n[0]);
            REAL *gr = g + 2 * (size_t)((u2 * nt[1] + u1) * nt[0]);
            for (int64_t i0 = 0; i0 < n[0]; i0++) {
                int64_t u0 = ((i0 - n[0] / 2) % nt[0] + nt[0]) % nt[0];
                REAL s = lt[0][i0];
                if (dir == 0) {
                    REAL vr = fr[2 * i0] * s, vi = fr[2 * i0 + 1] * s;
                    if (D > 1) { vr *= lt[1][i1]; vi *= lt[1][i1]; }
                    if (D > 2) { vr *= lt[2][i2]; vi *= lt[2][i2]; }
                    gr[2 * u0] = vr; gr[2 * u0 + 1] = vi;
                } else {
                    REAL vr = gr[2 * u0] * s, vi = gr[2 * u0 + 1] * s;
                    if (D > 1) { vr *= lt[1][i1]; vi *= lt[1][i1]; }
                    if (D > 2) { vr *= lt[2][i2]; vi *= lt[2][i2]; }
                    fr[2 * i0] = vr; fr[2 * i0 + 1] = vi;
                }
            }
        }
}

#undef FN
#undef CAT
#undef CAT_
