// tables.cpp -- host-side window tables of the plan (K7), built in double and uploaded in T:
//   windowHatInvLUT   /root/reference/src/precomputation.jl:347-358 (exact 1/phi_hat instead of the
//                     30-point Chebyshev interpolant; they agree to <= 7e-15 relative)
//   windowLinInterp   /root/reference/src/precomputation.jl:291-300
//   windowPolyInterp  /root/reference/src/precomputation.jl:302-320 (least squares via Householder QR)
// window pairs: /root/reference/src/windowFunctions.jl:21-134 (getWindow, :4-19), evaluated in grid units.
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

// cardinal B-spline of the given order on the knots 0..order (windowFunctions.jl:75-86), bottom-up
double cbspline(int order, double k)
{
    std::vector<double> a(order + 1, 0.0);
    for (int j = 0; j < order; j++) a[j] = (k - j >= 0 && k - j < 1) ? 1.0 : 0.0;
    for (int n = 2; n <= order; n++)
        for (int j = 0; j + n <= order; j++)
            a[j] = (k - j) / (n - 1) * a[j] + (n - (k - j)) / (n - 1) * a[j + 1];
    return a[0];
}

// window(x) in grid units; every window but kaiser_bessel is 0 for |x| >= m
double window_eval(const nfftb200_plan* p, double x)
{
    const int m = p->m;
    if (p->window == NFFTB200_KAISER_BESSEL) return kb_window(x, m, p->b);
    if (!(std::fabs(x) < m)) return 0.0;
    switch (p->window) {
    case NFFTB200_GAUSS: {
        const double b = m / M_PI;
        return 1.0 / std::sqrt(M_PI * b) * std::exp(-(x * x) / b);
    }
    case NFFTB200_SPLINE:
        return cbspline(2 * m, x + m);
    case NFFTB200_KAISER_BESSEL_REV: {
        const double q = x / m;
        return 0.5 / m * std::cyl_bessel_i(0.0, m * p->b * std::sqrt(1.0 - q * q));
    }
    case NFFTB200_EXP_SQRT: {
        const double q = x / m;
        return std::exp(p->beta * (std::sqrt(1.0 - q * q) - 1.0));
    }
    default: {   // NFFTB200_COSH_TYPE
        const double beta = M_PI * m * (2.0 - 1.0 / p->sigma);
        const double q = x / m;
        const double alpha = std::sqrt(1.0 - q * q);
        return 1.0 / (std::cosh(beta) - 1.0) * (std::cosh(beta * alpha) - 1.0) / alpha;
    }
    }
}

// Gauss-Legendre nodes / weights on [-1, 1] (Newton on the Legendre recurrence)
void gauss_legendre(int n, std::vector<double>& x, std::vector<double>& w)
{
    x.assign(n, 0.0); w.assign(n, 0.0);
    for (int i = 0; i < n; i++) {
        double z = std::cos(M_PI * (i + 0.75) / (n + 0.5)), pp = 1.0;
        for (int it = 0; it < 100; it++) {
            double p1 = 1.0, p2 = 0.0;
            for (int j = 0; j < n; j++) { const double p3 = p2; p2 = p1; p1 = ((2.0 * j + 1.0) * z * p2 - j * p3) / (j + 1.0); }
            pp = n * (z * p1 - p2) / (z * z - 1.0);
            const double dz = p1 / pp;
            z -= dz;
            if (std::fabs(dz) < 1e-15) break;
        }
        x[i] = z; w[i] = 2.0 / ((1.0 - z * z) * pp * pp);
    }
}

double window_hat_eval(const nfftb200_plan* p, double n, double Nt)
{
    const int m = p->m;
    switch (p->window) {
    case NFFTB200_EXP_SQRT: {
        // phi_hat(n) = int_{-m}^{m} phi(x) cos(2 pi n x / Nt) dx, no closed form: 96-point Gauss-Legendre on [0, m]
        static std::vector<double> gx, gw;
        if (gx.empty()) gauss_legendre(96, gx, gw);
        double s = 0.0;
        for (size_t i = 0; i < gx.size(); i++) {
            const double x = 0.5 * m * (gx[i] + 1.0), q = x / m;
            s += gw[i] * std::exp(p->beta * (std::sqrt(1.0 - q * q) - 1.0)) * std::cos(2.0 * M_PI * n * x / Nt);
        }
        return s * m;      // 2 * (m / 2) * sum
    }
    case NFFTB200_KAISER_BESSEL:
        return kb_window_hat(n, Nt, m, p->b);
    case NFFTB200_GAUSS: {
        const double t = M_PI * n / Nt;
        return std::exp(-(t * t) * (m / M_PI));
    }
    case NFFTB200_SPLINE: {
        const double t = M_PI * n / Nt;
        const double sc = t == 0.0 ? 1.0 : std::sin(t) / t;
        return std::pow(sc, 2 * m);
    }
    case NFFTB200_KAISER_BESSEL_REV: {
        // real(sinc(sqrt(complex(q))/pi)): sinh(a)/a below the cut-off, sin(a)/a above it
        const double t = 2.0 * M_PI * m * n / Nt, mb = m * p->b;
        const double q = t * t - mb * mb;
        const double a = std::sqrt(std::fabs(q));
        if (a == 0.0) return 1.0;
        return q < 0 ? std::sinh(a) / a : std::sin(a) / a;
    }
    default: {   // NFFTB200_COSH_TYPE
        const double beta = M_PI * m * (2.0 - 1.0 / p->sigma);
        const double gamma = beta / (2 * M_PI);
        const double zeta = M_PI / (std::cosh(beta) - 1.0) * m;
        const double arg = m * n / Nt, two = 2 * M_PI * arg;
        const double j0 = std::cyl_bessel_j(0.0, std::fabs(two));
        if (std::fabs(arg) < gamma) return zeta * (std::cyl_bessel_i(0.0, std::sqrt(beta * beta - two * two)) - j0);
        if (std::fabs(arg) > gamma) return zeta * (std::cyl_bessel_j(0.0, std::sqrt(two * two - beta * beta)) - j0);
        return zeta * (1.0 - std::cyl_bessel_j(0.0, beta));
    }
    }
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
    // 1/phi_hat per dimension, index i <-> frequency n = i - N/2
    p->h_hat_inv.clear();
    for (int d = 0; d < p->D; d++)
        for (int64_t i = 0; i < p->N[d]; i++)
            p->h_hat_inv.push_back(1.0 / window_hat_eval(p, (double)(i - p->N[d] / 2), (double)p->Nt[d]));
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
                y[i] = window_eval(p, (-(l - 0.5) + m) + t[i]);
            }
            lstsq_qr(A, y, ns, deg, &p->h_poly[(size_t)(l - 1) * deg]);
        }
    } else if (mode == NFFTB200_LINEAR) {
        const int64_t K = p->lut_size;
        const double step = (double)m / (double)K;
        p->h_lin.resize((size_t)K + 2);
        for (int64_t l = 0; l < K + 2; l++) p->h_lin[(size_t)l] = window_eval(p, (double)l * step);
    }
    return NFFTB200_OK;
}
