// tables.cpp -- host-side window tables of the plan (K7), built in double and uploaded in T:
//   windowHatInvLUT   /root/reference/src/precomputation.jl:347-358 (exact 1/phi_hat instead of the
//                     30-point Chebyshev interpolant; they agree to <= 7e-15 relative)
//   windowLinInterp   /root/reference/src/precomputation.jl:291-300
//   windowPolyInterp  /root/reference/src/precomputation.jl:302-320 (least squares via Householder QR)
// window pair: /root/reference/src/windowFunctions.jl:21-39.
#include <cmath>
#include <vector>

#include "common.cuh"

namespace {

double kb_window(double x, int m, double b)
{
    const double ax = std::fabs(x);
    if (ax < m) {
        const double arg = std::sqrt((double)m * m - x * x);
        return std::sinh(b * arg) / (arg * M_PI);
    }
    if (ax > m) return 0.0;
    return b / M_PI;
}

double kb_window_hat(double n, double Nt, int m, double b)
{
    const double t = 2.0 * M_PI * n / Nt;
    return std::cyl_bessel_i(0.0, m * std::sqrt(b * b - t * t));
}

// min ||A x - y||_2 for A (rows x cols, row-major) by Householder QR; A and y are overwritten.
void lstsq_qr(std::vector<double>& A, std::vector<double>& y, int rows, int cols, double* x)
{
    for (int k = 0; k < cols; k++) {
        double nrm = 0;
        for (int i = k; i < rows; i++) nrm += A[i * cols + k] * A[i * cols + k];
        nrm = std::sqrt(nrm);
        if (nrm == 0) continue;
        const double alpha = A[k * cols + k] > 0 ? -nrm : nrm;
        std::vector<double> v(rows, 0.0);
        for (int i = k; i < rows; i++) v[i] = A[i * cols + k];
        v[k] -= alpha;
        double vn = 0;
        for (int i = k; i < rows; i++) vn += v[i] * v[i];
        if (vn == 0) continue;
        for (int j = k; j < cols; j++) {
            double s = 0;
            for (int i = k; i < rows; i++) s += v[i] * A[i * cols + j];
            s = 2 * s / vn;
            for (int i = k; i < rows; i++) A[i * cols + j] -= s * v[i];
        }
        double s = 0;
        for (int i = k; i < rows; i++) s += v[i] * y[i];
        s = 2 * s / vn;
        for (int i = k; i < rows; i++) y[i] -= s * v[i];
    }
    for (int k = cols - 1; k >= 0; k--) {
        double s = y[k];
        for (int j = k + 1; j < cols; j++) s -= A[k * cols + j] * x[j];
        x[k] = s / A[k * cols + k];
    }
}

}  // namespace

int nfftb_build_tables(nfftb200_plan* p)
{
    const int m = p->m;
    const double b = p->b;
    // 1/phi_hat per dimension, index i <-> frequency n = i - N/2
    p->h_hat_inv.clear();
    for (int d = 0; d < p->D; d++)
        for (int64_t i = 0; i < p->N[d]; i++)
            p->h_hat_inv.push_back(1.0 / kb_window_hat((double)(i - p->N[d] / 2), (double)p->Nt[d], m, b));
    p->h_poly.clear();
    p->h_lin.clear();
    const int mode = p->precompute;
    if (mode == NFFTB200_POLYNOMIAL || mode == NFFTB200_TENSOR) {
        const int deg = 2 * m + 1, K = 2 * m, ns = 2 * deg;
        p->h_poly.assign((size_t)deg * K, 0.0);
        std::vector<double> t(ns);
        for (int i = 0; i < ns; i++) t[i] = -0.5 + (double)i / (ns - 1);   // range(-0.5,0.5,length=ns)
        for (int l = 1; l <= K; l++) {
            std::vector<double> A((size_t)ns * deg), y(ns);
            for (int i = 0; i < ns; i++) {
                double pw = 1.0;
                for (int r = 0; r < deg; r++) { A[(size_t)i * deg + r] = pw; pw *= t[i]; }
                y[i] = kb_window((-(l - 0.5) + m) + t[i], m, b);
            }
            lstsq_qr(A, y, ns, deg, &p->h_poly[(size_t)(l - 1) * deg]);
        }
    } else if (mode == NFFTB200_LINEAR) {
        const int64_t K = p->lut_size;
        const double step = (double)m / (double)K;
        p->h_lin.resize((size_t)K + 2);
        for (int64_t l = 0; l < K + 2; l++) p->h_lin[(size_t)l] = kb_window((double)l * step, m, b);
    }
    return NFFTB200_OK;
}
