/*
 * oracle/brute64.c — fp64 brute-force Gaussian-KDE partial log-sum-exp in plain C (TEST INFRASTRUCTURE ONLY).
 *
 * The C twin of oracle/brute64.py::partial_lse, for parity checks at BASELINE.json's full training-set
 * sizes (d = 16 x 2e6 rows, d = 8 x 4e6 rows) where the numpy version needs minutes.  Same definition
 * (sklearn neighbors/_binary_tree.pxi.tp:376-378 log K = -0.5 d^2 / h^2; :1527-1529 / _kde.py:273-287
 * log p = log_knorm + log sum_j w_j K_j - log sum_j w_j), same arithmetic: direct differences, fp64
 * everywhere, max-shifted log-sum-exp.  tests/test_oracle.py pins it against brute64.py (and through it
 * against the golden vectors of the reference).  Nothing under skysampler_b200/ links or loads this.
 *
 * Build (done by __graft_entry__.build()):  gcc -O2 -fopenmp -shared -fPIC -o oracle/_brute64.so oracle/brute64.c -lm
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

/* per query i:  sum_j w_j exp(-|q_i - x_j|^2 / 2h^2) = s[i] * exp(m[i]),  m[i] = max_j (log w_j - 0.5 d_ij^2 / h^2)
 * over w_j > 0 (m = -inf, s = 0 when no row carries weight).  logw may be NULL (unit weights). */
void brute64_partial_lse(const double* train, const double* logw, int64_t n, int d, double h,
                         const double* query, int64_t m_rows, double* out_m, double* out_s)
{
    const double inv2h2 = 0.5 / (h * h);
#pragma omp parallel
    {
        double* e = (double*)malloc(sizeof(double) * (size_t)(n > 0 ? n : 1));
#pragma omp for schedule(dynamic, 4)
        for (int64_t i = 0; i < m_rows; i++) {
            const double* q = query + (size_t)i * d;
            double mx = -INFINITY;
            for (int64_t j = 0; j < n; j++) {
                const double* x = train + (size_t)j * d;
                double d2 = 0.0;
                for (int k = 0; k < d; k++) {
                    const double t = q[k] - x[k];
                    d2 += t * t;
                }
                const double v = (logw ? logw[j] : 0.0) - d2 * inv2h2;
                e[j] = v;
                if (v > mx) mx = v;
            }
            const double ms = isfinite(mx) ? mx : 0.0;
            double s = 0.0;
            for (int64_t j = 0; j < n; j++) {
                const double a = e[j] - ms;
                if (a > -745.0) s += exp(a);      /* exp(a) == 0 in fp64 below -745.13: skipping is exact */
            }
            out_m[i] = mx;
            out_s[i] = s;
        }
        free(e);
    }
}
