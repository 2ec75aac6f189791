/*
 * oracle/nfft_ref.c -- C/OpenMP restatement of the reference's blocked CPU algorithm.
 *
 * TEST INFRASTRUCTURE ONLY: linked/loaded only by tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py.  Never by the product path.
 *
 * "parity unpinned": the reference (pure Julia) cannot run here and ships no golden vectors;
 * this file is pinned at the pipeline level only (see oracle/nfft_oracle.py header).  It is
 * kind = "port" in bench.py's cpu_baseline: Julia's NFFT.jl itself is NOT what is timed.
 * The FFT between the stages is pocketfft via scipy.fft (FFTW is absent), driven from
 * oracle/cpu_ref.py.
 *
 * Build: make -C oracle   (gcc -O3 -march=native -fopenmp -shared) -> oracle/_build/libnfft_ref.so
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define REAL float
#define SUF _f32
#define REAL_EPS 1.1920928955078125e-07f
#define FMA(a, b, c) fmaf(a, b, c)
#include "nfft_ref_impl.h"
#undef REAL
#undef SUF
#undef REAL_EPS
#undef FMA

#define REAL double
#define SUF _f64
#define REAL_EPS 2.220446049250313e-16
#define FMA(a, b, c) fma(a, b, c)
#include "nfft_ref_impl.h"
#undef REAL
#undef SUF
#undef REAL_EPS
#undef FMA

/* torchrun exports OMP_NUM_THREADS=1 to every rank; the baseline is meant to use all host cores */
void ref_set_num_threads(int n)
{
#ifdef _OPENMP
    extern void omp_set_num_threads(int);
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

int ref_num_threads(void)
{
    int n = 1;
#ifdef _OPENMP
    extern int omp_get_max_threads(void);
    n = omp_get_max_threads();
#endif
    return n;
}
