/*
 * disort_oracle.c -- TEST INFRASTRUCTURE ONLY (see disort_oracle.h).
 *
 * Plain-C FP64 restatement of the reference's per-bin discrete-ordinate
 * solve.  Every routine cites the reference lines it follows
 * (/root/reference/disort.f, /root/reference/disutil.f).  It deliberately
 * keeps the reference's numerical choices where they change results:
 * delta-M with f = PMOM(NSTR), DITHER, ABSCUT layer truncation, the PLKAVG
 * series, Newton Gauss nodes, the ASYMTX Hessenberg-QR eigen-solver and
 * LINPACK-style partial-pivot LU (dense and banded).  Single-precision
 * literals of the reference (PI = 2.*ASIN(1.0), C2, SIGMA) are mirrored by
 * rounding through float.
 *
 * Out of scope: IBCND=1, BRDF surfaces, printing.
 *
 * Parity pin: DISORT self-test constants disort.f:6446-6449 (fluxes and the
 * CORINT-corrected intensity) -- tests/test_oracle_golden.py.
 */
#include "disort_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>

/* reference: PI = 2.*ASIN(1.0) evaluated in default REAL (disort.f:441) */
static double ref_pi(void) { return (double)(2.0f * asinf(1.0f)); }

/* R1MACH(4) for REAL(KR) = double: B**(1-DIGITS) (disutil.f:91-92) */
#define R1MACH4 DBL_EPSILON

/* ------------------------------------------------------------------ */
/* QGAUSN  (disort.f:5984-6157): Gauss-Legendre nodes/weights on (0,1) */
/* ------------------------------------------------------------------ */
void sbdo_qgausn(int m, double *gmu, double *gwt)
{
    const double pi = ref_pi();
    const double tol = 10.0 * R1MACH4;
    if (m < 1) return;
    if (m == 1) { gmu[0] = 0.5; gwt[0] = 1.0; return; }
    const double en = m;
    const int np1 = m + 1;
    const double nnp1 = (double)m * np1;
    const double cona = (double)((float)(m - 1) / (float)(8 * m * m * m));
    const int lim = m / 2;
    double p = 0.0, pm1, pm2, tmp = 0.0;
    for (int k = 1; k <= lim; k++) {
        double t = (4 * k - 1) * pi / (4 * m + 2);
        double x = cos(t + cona / tan(t));
        int iter = 0;
        for (;;) {
            iter++;
            pm2 = 1.0; pm1 = x;
            for (int nn = 2; nn <= m; nn++) {
                p = ((2 * nn - 1) * x * pm1 - (nn - 1) * pm2) / nn;
                pm2 = pm1; pm1 = p;
            }
            tmp = 1.0 / (1.0 - x * x);
            double ppr = en * (pm2 - x * p) * tmp;
            double p2pri = (2.0 * x * ppr - nnp1 * p) * tmp;
            double xi = x - (p / ppr) * (1.0 + (p / ppr) * p2pri / (2.0 * ppr));
            if (fabs(xi - x) > tol && iter <= 1000) { x = xi; continue; }
            break;
        }
        gmu[k - 1] = -x;
        gwt[k - 1] = 2.0 / (tmp * (en * pm2) * (en * pm2));
        gmu[np1 - k - 1] = -gmu[k - 1];
        gwt[np1 - k - 1] = gwt[k - 1];
    }
    if (m % 2 != 0) {
        gmu[lim] = 0.0;
        double prod = 1.0;
        for (int k = 3; k <= m; k += 2) prod = prod * k / (k - 1);
        gwt[lim] = 2.0 / (prod * prod);
    }
    for (int k = 0; k < m; k++) {
        gmu[k] = 0.5 * gmu[k] + 0.5;
        gwt[k] = 0.5 * gwt[k];
    }
}

/* ------------------------------------------------------------------ */
/* PLKAVG  (disort.f:5410-5671): band-integrated Planck function       */
/* ------------------------------------------------------------------ */
static double plkf(double x) { return x * x * x / (exp(x) - 1.0); }

double sbdo_plkavg(double wnumlo, double wnumhi, double t, int *warn)
{
    const double a1 = 1. / 3., a2 = -1. / 8., a3 = 1. / 60., a4 = -1. / 5040.,
                 a5 = 1. / 272160., a6 = -1. / 13305600.;
    const double c2 = (double)1.438786f, sigma = (double)5.67032E-8f,
                 vcut = 1.5;
    static const double vcp[7] = { 10.25, (double)5.7f, (double)3.9f,
                                   (double)2.9f, (double)2.3f, (double)1.9f,
                                   0.0 };
    const double pi = ref_pi();
    const double vmax = log(DBL_MAX);
    const double epsil = R1MACH4;
    const double sigdpi = sigma / pi;
    const double conc = 15. / (pi * pi * pi * pi);
    double d[2] = { 0, 0 }, p[2] = { 0, 0 }, v[2];

    if (t < 0.0 || wnumhi <= wnumlo || wnumlo < 0.) return NAN;
    if (t < 1.e-4) return 0.0;
    v[0] = c2 * wnumlo / t;
    v[1] = c2 * wnumhi / t;
    if (v[0] > epsil && v[1] < vmax && (wnumhi - wnumlo) / wnumhi < 1.e-2) {
        /* Simpson rule with convergence test (disort.f:5566-5598) */
        double hh = v[1] - v[0], oldval = 0.0, val = 0.0;
        double val0 = plkf(v[0]) + plkf(v[1]);
        int conv = 0;
        for (int n = 1; n <= 10; n++) {
            double del = hh / (2 * n);
            val = val0;
            for (int k = 1; k <= 2 * n - 1; k++)
                val += 2 * (1 + k % 2) * plkf(v[0] + k * del);
            val = del / 3. * val;
            if (fabs((val - oldval) / val) <= 1.e-6) { conv = 1; break; }
            oldval = val;
        }
        if (!conv && warn) *warn |= 1 << 9;
        return sigdpi * t * t * t * t * conc * val;
    }
    int smallv = 0;
    for (int i = 0; i < 2; i++) {
        if (v[i] < vcut) {
            smallv++;
            double vsq = v[i] * v[i];
            p[i] = conc * vsq * v[i] *
                   (a1 + v[i] * (a2 + v[i] * (a3 + vsq * (a4 + vsq * (a5 + vsq * a6)))));
        } else {
            int mmax = 0;
            do { mmax++; } while (v[i] < vcp[mmax - 1]);
            double ex = exp(-v[i]), exm = 1.0;
            d[i] = 0.0;
            for (int m = 1; m <= mmax; m++) {
                double mv = m * v[i];
                exm = ex * exm;
                d[i] += exm * (6. + mv * (6. + mv * (3. + mv))) /
                        ((double)m * m * m * m);
            }
            d[i] = conc * d[i];
        }
    }
    double r;
    if (smallv == 2) r = p[1] - p[0];
    else if (smallv == 1) r = 1. - p[0] - d[1];
    else r = d[0] - d[1];
    r = sigdpi * t * t * t * t * r;
    if (r == 0.0 && warn) *warn |= 1 << 10;
    return r;
}

/* ------------------------------------------------------------------ */
/* LEPOLY (disort.f:5286-5408): normalised associated Legendre Y_l^m   */
/* ylm is [nmu][ld], l = 0..twonm1; needs the m-1 result in place.     */
/* ------------------------------------------------------------------ */
static void lepoly(int nmu, int m, int ld, int twonm1, const double *mu,
                   double *ylm)
{
#define YLM(l, i) ylm[(size_t)(i) * ld + (l)]
    if (m == 0) {
        for (int i = 0; i < nmu; i++) { YLM(0, i) = 1.0; YLM(1, i) = mu[i]; }
        for (int l = 2; l <= twonm1; l++)
            for (int i = 0; i < nmu; i++)
                YLM(l, i) = ((2 * l - 1) * mu[i] * YLM(l - 1, i) -
                             (l - 1) * YLM(l - 2, i)) / l;
    } else {
        for (int i = 0; i < nmu; i++) {
            YLM(m, i) = -sqrt((double)(2 * m - 1)) / sqrt((double)(2 * m)) *
                        sqrt(1. - mu[i] * mu[i]) * YLM(m - 1, i);
            YLM(m + 1, i) = sqrt((double)(2 * m + 1)) * mu[i] * YLM(m, i);
        }
        for (int l = m + 2; l <= twonm1; l++) {
            double tmp1 = sqrt((double)(l - m)) * sqrt((double)(l + m));
            double tmp2 = sqrt((double)(l - m - 1)) * sqrt((double)(l + m - 1));
            for (int i = 0; i < nmu; i++)
                YLM(l, i) = ((2 * l - 1) * mu[i] * YLM(l - 1, i) -
                             tmp2 * YLM(l - 2, i)) / tmp1;
        }
    }
#undef YLM
}

/* ------------------------------------------------------------------ */
/* ASYMTX (disort.f:873-1656): eigenvalues/vectors of a real general   */
/* matrix with real spectrum (balance, Hessenberg, shifted QR).        */
/* Column-major, 1-based macros; wk needs 2*m doubles.                 */
/* returns IER (0 ok, >0 eigenvalue index that failed; -1 complex 2x2) */
/* ------------------------------------------------------------------ */
int sbdo_asymtx(double *aa, double *evec, double *eval, int m, int ia,
                int ievec, double *wk)
{
#define AA(i, j) aa[((i) - 1) + (size_t)((j) - 1) * ia]
#define EVEC(i, j) evec[((i) - 1) + (size_t)((j) - 1) * ievec]
#define EVAL(i) eval[(i) - 1]
#define WK(i) wk[(i) - 1]
    const double c1 = 0.4375, c2 = 0.5, c3 = 0.75, c4 = 0.95, c5 = 16.0,
                 c6 = 256.0;
    const double tol = R1MACH4;
    double p = 0, q = 0, r = 0, s, t, w, x, y, z, f, g, h, rnorm, row, col,
           repl, scale, uu, vv;
    int i, j, k, l, n, n1, n2, in, lb = 0, ka, ii, kkk, lll, noconv, notlas;

    if (m < 1 || ia < m || ievec < m) return -1;
    if (m == 1) { EVAL(1) = AA(1, 1); EVEC(1, 1) = 1.0; return 0; }
    if (m == 2) {
        double discri = (AA(1, 1) - AA(2, 2)) * (AA(1, 1) - AA(2, 2)) +
                        4. * AA(1, 2) * AA(2, 1);
        if (discri < 0.0) return -1;
        double sgn = 1.0;
        if (AA(1, 1) < AA(2, 2)) sgn = -1.0;
        EVAL(1) = 0.5 * (AA(1, 1) + AA(2, 2) + sgn * sqrt(discri));
        EVAL(2) = 0.5 * (AA(1, 1) + AA(2, 2) - sgn * sqrt(discri));
        EVEC(1, 1) = 1.0; EVEC(2, 2) = 1.0;
        if (AA(1, 1) == AA(2, 2) && (AA(2, 1) == 0.0 || AA(1, 2) == 0.0)) {
            rnorm = fabs(AA(1, 1)) + fabs(AA(1, 2)) + fabs(AA(2, 1)) +
                    fabs(AA(2, 2));
            w = tol * rnorm;
            EVEC(2, 1) = AA(2, 1) / w;
            EVEC(1, 2) = -AA(1, 2) / w;
        } else {
            EVEC(2, 1) = AA(2, 1) / (EVAL(1) - AA(2, 2));
            EVEC(1, 2) = AA(1, 2) / (EVAL(2) - AA(1, 1));
        }
        return 0;
    }

    for (i = 1; i <= m; i++) {
        EVAL(i) = 0.0;
        for (j = 1; j <= m; j++) EVEC(i, j) = 0.0;
        EVEC(i, i) = 1.0;
    }

    /* balance: isolate eigenvalues by row/column permutation (:1043-1122) */
    rnorm = 0.0; l = 1; k = m;
    for (;;) {
        int found = 0;
        kkk = k;
        for (j = kkk; j >= 1; j--) {
            row = 0.0;
            for (i = 1; i <= k; i++) if (i != j) row += fabs(AA(j, i));
            if (row == 0.0) {
                WK(k) = j;
                if (j != k) {
                    for (i = 1; i <= k; i++) { repl = AA(i, j); AA(i, j) = AA(i, k); AA(i, k) = repl; }
                    for (i = l; i <= m; i++) { repl = AA(j, i); AA(j, i) = AA(k, i); AA(k, i) = repl; }
                }
                k--; found = 1; break;
            }
        }
        if (!found) break;
    }
    for (;;) {
        int found = 0;
        lll = l;
        for (j = lll; j <= k; j++) {
            col = 0.0;
            for (i = l; i <= k; i++) if (i != j) col += fabs(AA(i, j));
            if (col == 0.0) {
                WK(l) = j;
                if (j != l) {
                    for (i = 1; i <= k; i++) { repl = AA(i, j); AA(i, j) = AA(i, l); AA(i, l) = repl; }
                    for (i = l; i <= m; i++) { repl = AA(j, i); AA(j, i) = AA(l, i); AA(l, i) = repl; }
                }
                l++; found = 1; break;
            }
        }
        if (!found) break;
    }
    /* balance the submatrix in rows l..k (:1125-1188) */
    for (i = l; i <= k; i++) WK(i) = 1.0;
    do {
        noconv = 0;
        for (i = l; i <= k; i++) {
            col = 0.0; row = 0.0;
            for (j = l; j <= k; j++)
                if (j != i) { col += fabs(AA(j, i)); row += fabs(AA(i, j)); }
            f = 1.0; g = row / c5; h = col + row;
            while (col < g) { f *= c5; col *= c6; }
            g = row * c5;
            while (col >= g) { f /= c5; col /= c6; }
            if ((col + row) / f < c4 * h) {
                WK(i) = WK(i) * f;
                noconv = 1;
                for (j = l; j <= m; j++) AA(i, j) = AA(i, j) / f;
                for (j = 1; j <= k; j++) AA(j, i) = AA(j, i) * f;
            }
        }
    } while (noconv);

    /* reduce to upper Hessenberg by Householder (:1191-1286) */
    if (!(k - 1 < l + 1)) {
        for (n = l + 1; n <= k - 1; n++) {
            h = 0.0; WK(n + m) = 0.0; scale = 0.0;
            for (i = n; i <= k; i++) scale += fabs(AA(i, n - 1));
            if (scale != 0.0) {
                for (i = k; i >= n; i--) {
                    WK(i + m) = AA(i, n - 1) / scale;
                    h += WK(i + m) * WK(i + m);
                }
                g = -copysign(sqrt(h), WK(n + m));
                h = h - WK(n + m) * g;
                WK(n + m) = WK(n + m) - g;
                for (j = n; j <= m; j++) {
                    f = 0.0;
                    for (i = k; i >= n; i--) f += WK(i + m) * AA(i, j);
                    for (i = n; i <= k; i++) AA(i, j) = AA(i, j) - WK(i + m) * f / h;
                }
                for (i = 1; i <= k; i++) {
                    f = 0.0;
                    for (j = k; j >= n; j--) f += WK(j + m) * AA(i, j);
                    for (j = n; j <= k; j++) AA(i, j) = AA(i, j) - WK(j + m) * f / h;
                }
                WK(n + m) = scale * WK(n + m);
                AA(n, n - 1) = scale * g;
            }
        }
        for (n = k - 2; n >= l; n--) {
            n1 = n + 1; n2 = n + 2;
            f = AA(n + 1, n);
            if (f != 0.0) {
                f = f * WK(n + 1 + m);
                for (i = n + 2; i <= k; i++) WK(i + m) = AA(i, n);
                if (n + 1 <= k) {
                    for (j = 1; j <= m; j++) {
                        g = 0.0;
                        for (i = n + 1; i <= k; i++) g += WK(i + m) * EVEC(i, j);
                        g = g / f;
                        for (i = n + 1; i <= k; i++) EVEC(i, j) = EVEC(i, j) + g * WK(i + m);
                    }
                }
            }
        }
        (void)n1; (void)n2;
    }

    /* norm and isolated eigenvalues (:1289-1304) */
    n = 1;
    for (i = 1; i <= m; i++) {
        for (j = n; j <= m; j++) rnorm += fabs(AA(i, j));
        n = i;
        if (i < l || i > k) EVAL(i) = AA(i, i);
    }
    n = k; t = 0.0;

    /* shifted QR sweep for eigenvalue n (:1307-1546) */
    while (n >= l) {
        in = 0; n1 = n - 1; n2 = n - 2;
        for (;;) {
            /* look for a small sub-diagonal element */
            for (i = l; i <= n; i++) {
                lb = n + l - i;
                if (lb == l) break;
                s = fabs(AA(lb - 1, lb - 1)) + fabs(AA(lb, lb));
                if (s == 0.0) s = rnorm;
                if (fabs(AA(lb, lb - 1)) <= tol * s) break;
            }
            x = AA(n, n);
            if (lb == n) {              /* one eigenvalue found */
                AA(n, n) = x + t;
                EVAL(n) = AA(n, n);
                n = n1;
                break;
            }
            y = AA(n1, n1);
            w = AA(n, n1) * AA(n1, n);
            if (lb == n1) {             /* two eigenvalues found */
                p = (y - x) * c2;
                q = p * p + w;
                z = sqrt(fabs(q));
                AA(n, n) = x + t;
                x = AA(n, n);
                AA(n1, n1) = y + t;
                z = p + copysign(z, p);
                EVAL(n1) = x + z;
                EVAL(n) = EVAL(n1);
                if (z != 0.0) EVAL(n) = x - w / z;
                x = AA(n, n1);
                r = sqrt(x * x + z * z);
                p = x / r; q = z / r;
                for (j = n1; j <= m; j++) {
                    z = AA(n1, j);
                    AA(n1, j) = q * z + p * AA(n, j);
                    AA(n, j) = q * AA(n, j) - p * z;
                }
                for (i = 1; i <= n; i++) {
                    z = AA(i, n1);
                    AA(i, n1) = q * z + p * AA(i, n);
                    AA(i, n) = q * AA(i, n) - p * z;
                }
                for (i = l; i <= k; i++) {
                    z = EVEC(i, n1);
                    EVEC(i, n1) = q * z + p * EVEC(i, n);
                    EVEC(i, n) = q * EVEC(i, n) - p * z;
                }
                n = n2;
                break;
            }
            if (in == 30) return n;     /* no convergence */
            if (in == 10 || in == 20) { /* exceptional shift */
                t = t + x;
                for (i = l; i <= n; i++) AA(i, i) = AA(i, i) - x;
                s = fabs(AA(n, n1)) + fabs(AA(n1, n2));
                x = c3 * s; y = x; w = -c1 * s * s;
            }
            in++;
            /* look for two consecutive small sub-diagonal elements */
            for (j = lb; j <= n2; j++) {
                i = n2 + lb - j;
                z = AA(i, i);
                r = x - z; s = y - z;
                p = (r * s - w) / AA(i + 1, i) + AA(i, i + 1);
                q = AA(i + 1, i + 1) - z - r - s;
                r = AA(i + 2, i + 1);
                s = fabs(p) + fabs(q) + fabs(r);
                p /= s; q /= s; r /= s;
                if (i == lb) break;
                uu = fabs(AA(i, i - 1)) * (fabs(q) + fabs(r));
                vv = fabs(p) * (fabs(AA(i - 1, i - 1)) + fabs(z) + fabs(AA(i + 1, i + 1)));
                if (uu <= tol * vv) break;
            }
            AA(i + 2, i) = 0.0;
            for (j = i + 3; j <= n; j++) { AA(j, j - 2) = 0.0; AA(j, j - 3) = 0.0; }
            /* double QR step on rows lb..n, columns i..n */
            for (ka = i; ka <= n1; ka++) {
                notlas = (ka != n1);
                if (ka == i) {
                    s = copysign(sqrt(p * p + q * q + r * r), p);
                    if (lb != i) AA(ka, ka - 1) = -AA(ka, ka - 1);
                } else {
                    p = AA(ka, ka - 1);
                    q = AA(ka + 1, ka - 1);
                    r = 0.0;
                    if (notlas) r = AA(ka + 2, ka - 1);
                    x = fabs(p) + fabs(q) + fabs(r);
                    if (x == 0.0) continue;
                    p /= x; q /= x; r /= x;
                    s = copysign(sqrt(p * p + q * q + r * r), p);
                    AA(ka, ka - 1) = -s * x;
                }
                p = p + s;
                x = p / s; y = q / s; z = r / s;
                q = q / p; r = r / p;
                for (j = ka; j <= m; j++) {     /* row modification */
                    p = AA(ka, j) + q * AA(ka + 1, j);
                    if (notlas) {
                        p = p + r * AA(ka + 2, j);
                        AA(ka + 2, j) = AA(ka + 2, j) - p * z;
                    }
                    AA(ka + 1, j) = AA(ka + 1, j) - p * y;
                    AA(ka, j) = AA(ka, j) - p * x;
                }
                int iimax = (n < ka + 3) ? n : ka + 3;
                for (ii = 1; ii <= iimax; ii++) { /* column modification */
                    p = x * AA(ii, ka) + y * AA(ii, ka + 1);
                    if (notlas) {
                        p = p + z * AA(ii, ka + 2);
                        AA(ii, ka + 2) = AA(ii, ka + 2) - p * r;
                    }
                    AA(ii, ka + 1) = AA(ii, ka + 1) - p * q;
                    AA(ii, ka) = AA(ii, ka) - p;
                }
                for (ii = l; ii <= k; ii++) {   /* accumulate transformations */
                    p = x * EVEC(ii, ka) + y * EVEC(ii, ka + 1);
                    if (notlas) {
                        p = p + z * EVEC(ii, ka + 2);
                        EVEC(ii, ka + 2) = EVEC(ii, ka + 2) - p * r;
                    }
                    EVEC(ii, ka + 1) = EVEC(ii, ka + 1) - p * q;
                    EVEC(ii, ka) = EVEC(ii, ka) - p;
                }
            }
        }
    }

    /* back-substitute for vectors of the upper triangular form (:1551-1609) */
    if (rnorm != 0.0) {
        for (n = m; n >= 1; n--) {
            n2 = n;
            AA(n, n) = 1.0;
            for (i = n - 1; i >= 1; i--) {
                w = AA(i, i) - EVAL(n);
                if (w == 0.0) w = tol * rnorm;
                r = AA(i, n);
                for (j = n2; j <= n - 1; j++) r += AA(i, j) * AA(j, n);
                AA(i, n) = -r / w;
                n2 = i;
            }
        }
        for (i = 1; i <= m; i++)
            if (i < l || i > k)
                for (j = i; j <= m; j++) EVEC(i, j) = AA(i, j);
        if (k != 0) {
            for (j = m; j >= l; j--) {
                for (i = l; i <= k; i++) {
                    z = 0.0;
                    int nmax = (j < k) ? j : k;
                    for (n = l; n <= nmax; n++) z += EVEC(i, n) * AA(n, j);
                    EVEC(i, j) = z;
                }
            }
        }
    }
    for (i = l; i <= k; i++)
        for (j = 1; j <= m; j++) EVEC(i, j) = EVEC(i, j) * WK(i);
    for (i = l - 1; i >= 1; i--) {
        j = (int)WK(i);
        if (i != j)
            for (n = 1; n <= m; n++) { repl = EVEC(i, n); EVEC(i, n) = EVEC(j, n); EVEC(j, n) = repl; }
    }
    for (i = k + 1; i <= m; i++) {
        j = (int)WK(i);
        if (i != j)
            for (n = 1; n <= m; n++) { repl = EVEC(i, n); EVEC(i, n) = EVEC(j, n); EVEC(j, n) = repl; }
    }
    return 0;
#undef AA
#undef EVEC
#undef EVAL
#undef WK
}

/* ------------------------------------------------------------------ */
/* Dense LU with partial pivoting, column-major (SGEFA/SGESL,          */
/* disutil.f:1355-1609).  The reference's SGECO condition estimate only */
/* raises a warning; here the warning is raised on an exact zero pivot. */
/* ------------------------------------------------------------------ */
static int gefa(double *a, int lda, int n, int *ipvt)
{
#define A(i, j) a[(i) + (size_t)(j) * lda]
    int info = 0;
    for (int k = 0; k < n - 1; k++) {
        int l = k; double amax = fabs(A(k, k));
        for (int i = k + 1; i < n; i++)
            if (fabs(A(i, k)) > amax) { amax = fabs(A(i, k)); l = i; }
        ipvt[k] = l;
        if (A(l, k) == 0.0) { info = k + 1; continue; }
        if (l != k) { double t = A(l, k); A(l, k) = A(k, k); A(k, k) = t; }
        double t = -1.0 / A(k, k);
        for (int i = k + 1; i < n; i++) A(i, k) *= t;
        for (int j = k + 1; j < n; j++) {
            double tj = A(l, j);
            if (l != k) { A(l, j) = A(k, j); A(k, j) = tj; }
            for (int i = k + 1; i < n; i++) A(i, j) += tj * A(i, k);
        }
    }
    ipvt[n - 1] = n - 1;
    if (A(n - 1, n - 1) == 0.0) info = n;
    return info;
}

static void gesl(const double *a, int lda, int n, const int *ipvt, double *b)
{
    for (int k = 0; k < n - 1; k++) {
        int l = ipvt[k];
        double t = b[l];
        if (l != k) { b[l] = b[k]; b[k] = t; }
        for (int i = k + 1; i < n; i++) b[i] += t * A(i, k);
    }
    for (int k = n - 1; k >= 0; k--) {
        b[k] /= A(k, k);
        double t = -b[k];
        for (int i = 0; i < k; i++) b[i] += t * A(i, k);
    }
#undef A
}

/* ------------------------------------------------------------------ */
/* Band LU with partial pivoting in LINPACK band storage (SGBFA/SGBSL, */
/* disutil.f:771-1060); abd is [lda][n] column-major, 1-based macros.   */
/* ------------------------------------------------------------------ */
static int gbfa(double *abd, int lda, int n, int ml, int mu, int *ipvt)
{
#define ABD(i, j) abd[((i) - 1) + (size_t)((j) - 1) * lda]
    int m = ml + mu + 1, info = 0;
    int j0 = mu + 2, j1 = ((n < m) ? n : m) - 1;
    for (int jz = j0; jz <= j1; jz++) {
        int i0 = m + 1 - jz;
        for (int i = i0; i <= ml; i++) ABD(i, jz) = 0.0;
    }
    int jz = j1, ju = 0;
    for (int k = 1; k <= n - 1; k++) {
        jz++;
        if (jz <= n) for (int i = 1; i <= ml; i++) ABD(i, jz) = 0.0;
        int lm = (ml < n - k) ? ml : n - k;
        int l = m; double amax = fabs(ABD(m, k));
        for (int i = 1; i <= lm; i++)
            if (fabs(ABD(m + i, k)) > amax) { amax = fabs(ABD(m + i, k)); l = m + i; }
        ipvt[k - 1] = l + k - m;
        if (ABD(l, k) == 0.0) { info = k; continue; }
        if (l != m) { double t = ABD(l, k); ABD(l, k) = ABD(m, k); ABD(m, k) = t; }
        double t = -1.0 / ABD(m, k);
        for (int i = 1; i <= lm; i++) ABD(m + i, k) *= t;
        int cand = mu + ipvt[k - 1];
        if (cand > ju) ju = cand;
        if (ju > n) ju = n;
        int mm = m;
        for (int j = k + 1; j <= ju; j++) {
            l--; mm--;
            double tj = ABD(l, j);
            if (l != mm) { ABD(l, j) = ABD(mm, j); ABD(mm, j) = tj; }
            for (int i = 1; i <= lm; i++) ABD(mm + i, j) += tj * ABD(m + i, k);
        }
    }
    ipvt[n - 1] = n;
    if (ABD(m, n) == 0.0) info = n;
    return info;
}

static void gbsl(const double *abd, int lda, int n, int ml, int mu,
                 const int *ipvt, double *b)
{
    int m = mu + ml + 1;
    if (ml != 0) {
        for (int k = 1; k <= n - 1; k++) {
            int lm = (ml < n - k) ? ml : n - k;
            int l = ipvt[k - 1];
            double t = b[l - 1];
            if (l != k) { b[l - 1] = b[k - 1]; b[k - 1] = t; }
            for (int i = 1; i <= lm; i++) b[k - 1 + i] += t * ABD(m + i, k);
        }
    }
    for (int kb = 1; kb <= n; kb++) {
        int k = n + 1 - kb;
        b[k - 1] /= ABD(m, k);
        int lm = ((k < m) ? k : m) - 1;
        int la = m - lm, lb = k - lm;
        double t = -b[k - 1];
        for (int i = 0; i < lm; i++) b[lb - 1 + i] += t * ABD(la + i, k);
    }
#undef ABD
}

/* ------------------------------------------------------------------ */
/* work space for one DISORT call                                      */
/* ------------------------------------------------------------------ */
typedef struct {
    int N, n, L, NT, NU, ncut, lyrcut;
    double *cmu, *cwt;                 /* [N]  first n: +mu ascending        */
    double *gl;                        /* [L][N+1]                           */
    double *dtaucp, *oprim, *flyr;     /* [L]                                */
    double *taucpr, *expbea, *tauc;    /* [L+1]                              */
    double *pkag;                      /* [L+1]                              */
    double *utau, *utaupr;             /* [NT]                               */
    int *layru;                        /* [NT] 1-based layer index           */
    double *ylm0, *ylmc, *ylmu;        /* [N+1], [N][N+1], [NU][N+1]         */
    double *cc, *evecc, *array;        /* [N][N] CMU ordering                */
    double *amb, *apb, *eval;          /* [n][n], [n]                        */
    double *gc;                        /* [L][N(col jq)][N(row iq)]          */
    double *gu;                        /* [L][N(col)][NU]                    */
    double *kk, *ll, *zz, *zplk0, *zplk1; /* [L][N]                          */
    double *xr0, *xr1;                 /* [L]                                */
    double *zbeam, *z0u, *z1u;         /* [L][NU]                            */
    double *bdr, *bem, *rmu, *emu;     /* surface arrays (SURFAC)            */
    int lamber;                        /* LAMBER                             */
    double *wk, *z0, *z1, *zj, *psi0, *psi1; /* [N+1] scratch                */
    double *cband, *b;                 /* band matrix and rhs                */
    int *ipvt;
    double *uum, *u0c;                 /* [NT][NU], [NT][N]                  */
    double *phirad;
} work_t;

#define GC(iq, jq, lc) w->gc[((size_t)(lc) * N + (jq)) * N + (iq)]
#define GU(iu, jq, lc) w->gu[((size_t)(lc) * N + (jq)) * NU + (iu)]
#define KK(jq, lc) w->kk[(size_t)(lc) * N + (jq)]
#define LL(jq, lc) w->ll[(size_t)(lc) * N + (jq)]
#define ZZ(iq, lc) w->zz[(size_t)(lc) * N + (iq)]
#define ZPLK0(iq, lc) w->zplk0[(size_t)(lc) * N + (iq)]
#define ZPLK1(iq, lc) w->zplk1[(size_t)(lc) * N + (iq)]
#define GLM(l, lc) w->gl[(size_t)(lc) * (N + 1) + (l)]
#define YLMC(l, iq) w->ylmc[(size_t)(iq) * (N + 1) + (l)]
#define YLMU(l, iu) w->ylmu[(size_t)(iu) * (N + 1) + (l)]
#define CC(iq, jq) w->cc[(iq) + (size_t)(jq) * N]
#define EVECC(iq, jq) w->evecc[(iq) + (size_t)(jq) * N]
#define ARR(iq, jq) w->array[(iq) + (size_t)(jq) * N]
#define AMB(iq, jq) w->amb[(iq) + (size_t)(jq) * n]
#define APB(iq, jq) w->apb[(iq) + (size_t)(jq) * n]

/* Per-thread bump arena: one (re)allocation per thread instead of ~50
 * calloc/free pairs per call (those serialise OpenMP threads in mmap). */
static __thread char *tl_base = NULL;
static __thread size_t tl_cap = 0, tl_off = 0;
static __thread int tl_measure = 0;

static void *arena_take(size_t bytes)
{
    bytes = (bytes + 63) & ~(size_t)63;
    if (bytes == 0) bytes = 64;
    size_t at = tl_off;
    tl_off += bytes;
    if (tl_measure) return NULL;
    return tl_base + at;
}
static double *dalloc(size_t k) { return (double *)arena_take(k * sizeof(double)); }
static int *ialloc(size_t k) { return (int *)arena_take(k * sizeof(int)); }

/* ------------------------------------------------------------------ */
/* SOLEIG (disort.f:3099-3320); lc is 0-based layer index              */
/* ------------------------------------------------------------------ */
static int soleig(work_t *w, int mazim, int lc)
{
    const int N = w->N, n = w->n;
    for (int iq = 0; iq < n; iq++) {
        for (int jq = 0; jq < N; jq++) {
            double sum = 0.0;
            for (int l = mazim; l <= N - 1; l++)
                sum += GLM(l, lc) * YLMC(l, iq) * YLMC(l, jq);
            CC(iq, jq) = 0.5 * sum * w->cwt[jq];
        }
        for (int jq = 0; jq < n; jq++) {
            CC(iq + n, jq) = CC(iq, jq + n);
            CC(iq + n, jq + n) = CC(iq, jq);
            double alpha = CC(iq, jq) / w->cmu[iq];
            double beta = CC(iq, jq + n) / w->cmu[iq];
            AMB(iq, jq) = alpha - beta;
            APB(iq, jq) = alpha + beta;
        }
        AMB(iq, iq) -= 1.0 / w->cmu[iq];
        APB(iq, iq) -= 1.0 / w->cmu[iq];
    }
    /* ARRAY is dimensioned (MI,*) in SOLEIG; use leading dim n here */
    double *arr = w->array;
    for (int iq = 0; iq < n; iq++)
        for (int jq = 0; jq < n; jq++) {
            double sum = 0.;
            for (int kq = 0; kq < n; kq++) sum += APB(iq, kq) * AMB(kq, jq);
            arr[iq + (size_t)jq * n] = sum;
        }
    int ier = sbdo_asymtx(arr, w->evecc, w->eval, n, n, N, w->wk);
    if (ier != 0) return SBDO_EIG_NOCONV;
    for (int iq = 0; iq < n; iq++) {
        w->eval[iq] = sqrt(fabs(w->eval[iq]));
        KK(iq + n, lc) = w->eval[iq];
        KK(n - 1 - iq, lc) = -w->eval[iq];
    }
    for (int jq = 0; jq < n; jq++)
        for (int iq = 0; iq < n; iq++) {
            double sum = 0.;
            for (int kq = 0; kq < n; kq++) sum += AMB(iq, kq) * EVECC(kq, jq);
            APB(iq, jq) = sum / w->eval[jq];
        }
    for (int jq = 0; jq < n; jq++)
        for (int iq = 0; iq < n; iq++) {
            double gpplgm = APB(iq, jq), gpmigm = EVECC(iq, jq);
            EVECC(iq, jq) = 0.5 * (gpplgm + gpmigm);
            EVECC(iq + n, jq) = 0.5 * (gpplgm - gpmigm);
            gpplgm = -gpplgm;
            EVECC(iq, jq + n) = 0.5 * (gpplgm + gpmigm);
            EVECC(iq + n, jq + n) = 0.5 * (gpplgm - gpmigm);
            GC(iq + n, jq + n, lc) = EVECC(iq, jq);
            GC(n - 1 - iq, jq + n, lc) = EVECC(iq + n, jq);
            GC(iq + n, n - 1 - jq, lc) = EVECC(iq, jq + n);
            GC(n - 1 - iq, n - 1 - jq, lc) = EVECC(iq + n, jq + n);
        }
    return 0;
}

/* UPBEAM (disort.f:4130-4245) */
static void upbeam(work_t *w, int mazim, int lc, double delm0, double fbeam,
                   double umu0, double pi, int *warn)
{
    const int N = w->N, n = w->n;
    for (int iq = 0; iq < N; iq++) {
        for (int jq = 0; jq < N; jq++) ARR(iq, jq) = -CC(iq, jq);
        ARR(iq, iq) = 1. + w->cmu[iq] / umu0 + ARR(iq, iq);
        double sum = 0.;
        for (int k = mazim; k <= N - 1; k++)
            sum += GLM(k, lc) * YLMC(k, iq) * w->ylm0[k];
        w->zj[iq] = (2. - delm0) * fbeam * sum / (4. * pi);
    }
    if (gefa(w->array, N, N, w->ipvt) != 0) *warn |= 1 << 3;
    gesl(w->array, N, N, w->ipvt, w->zj);
    for (int iq = 0; iq < n; iq++) {
        ZZ(iq + n, lc) = w->zj[iq];
        ZZ(n - 1 - iq, lc) = w->zj[iq + n];
    }
}

/* UPISOT (disort.f:4247-4353) */
static void upisot(work_t *w, int lc, int *warn)
{
    const int N = w->N, n = w->n;
    const double oprim = w->oprim[lc], xr0 = w->xr0[lc], xr1 = w->xr1[lc];
    for (int iq = 0; iq < N; iq++) {
        for (int jq = 0; jq < N; jq++) ARR(iq, jq) = -CC(iq, jq);
        ARR(iq, iq) = 1.0 + ARR(iq, iq);
        w->z1[iq] = (1. - oprim) * xr1;
    }
    if (gefa(w->array, N, N, w->ipvt) != 0) *warn |= 1 << 4;
    gesl(w->array, N, N, w->ipvt, w->z1);
    for (int iq = 0; iq < N; iq++)
        w->z0[iq] = (1. - oprim) * xr0 + w->cmu[iq] * w->z1[iq];
    gesl(w->array, N, N, w->ipvt, w->z0);
    for (int iq = 0; iq < n; iq++) {
        ZPLK0(iq + n, lc) = w->z0[iq];
        ZPLK1(iq + n, lc) = w->z1[iq];
        ZPLK0(n - 1 - iq, lc) = w->z0[iq + n];
        ZPLK1(n - 1 - iq, lc) = w->z1[iq + n];
    }
}

/* TERPEV (disort.f:3920-3978) */
static void terpev(work_t *w, int mazim, int lc)
{
    const int N = w->N, n = w->n, NU = w->NU;
    for (int iq = 0; iq < N; iq++) {
        for (int l = mazim; l <= N - 1; l++) {
            double sum = 0.0;
            for (int jq = 0; jq < N; jq++)
                sum += w->cwt[jq] * YLMC(l, jq) * EVECC(jq, iq);
            w->wk[l] = 0.5 * GLM(l, lc) * sum;
        }
        for (int iu = 0; iu < NU; iu++) {
            double sum = 0.;
            for (int l = mazim; l <= N - 1; l++) sum += w->wk[l] * YLMU(l, iu);
            if (iq < n) GU(iu, iq + n, lc) = sum;
            else GU(iu, N - 1 - iq, lc) = sum;
        }
    }
}

/* TERPSO (disort.f:3980-4128) */
static void terpso(work_t *w, int mazim, int lc, double delm0, double fbeam,
                   int plank, double pi)
{
    const int N = w->N, NU = w->NU;
    if (fbeam > 0.0) {
        for (int iq = mazim; iq <= N - 1; iq++) {
            double psum = 0.;
            for (int jq = 0; jq < N; jq++)
                psum += w->cwt[jq] * YLMC(iq, jq) * w->zj[jq];
            w->psi0[iq] = 0.5 * GLM(iq, lc) * psum;
        }
        double fact = (2. - delm0) * fbeam / (4.0 * pi);
        for (int iu = 0; iu < NU; iu++) {
            double sum = 0.;
            for (int iq = mazim; iq <= N - 1; iq++)
                sum += YLMU(iq, iu) * (w->psi0[iq] + fact * GLM(iq, lc) * w->ylm0[iq]);
            w->zbeam[(size_t)lc * NU + iu] = sum;
        }
    }
    if (plank && mazim == 0) {
        for (int iq = mazim; iq <= N - 1; iq++) {
            double psum0 = 0.0, psum1 = 0.0;
            for (int jq = 0; jq < N; jq++) {
                psum0 += w->cwt[jq] * YLMC(iq, jq) * w->z0[jq];
                psum1 += w->cwt[jq] * YLMC(iq, jq) * w->z1[jq];
            }
            w->psi0[iq] = 0.5 * GLM(iq, lc) * psum0;
            w->psi1[iq] = 0.5 * GLM(iq, lc) * psum1;
        }
        for (int iu = 0; iu < NU; iu++) {
            double sum0 = 0.0, sum1 = 0.0;
            for (int iq = mazim; iq <= N - 1; iq++) {
                sum0 += YLMU(iq, iu) * w->psi0[iq];
                sum1 += YLMU(iq, iu) * w->psi1[iq];
            }
            w->z0u[(size_t)lc * NU + iu] = sum0 + (1. - w->oprim[lc]) * w->xr0[lc];
            w->z1u[(size_t)lc * NU + iu] = sum1 + (1. - w->oprim[lc]) * w->xr1[lc];
        }
    }
}

/* SETMTX (disort.f:2702-2994) + SOLVE0 (disort.f:3322-3637) */
static int setmtx_solve0(work_t *w, int mazim, double delm0, double fbeam,
                         double umu0, double fisot, double tplank,
                         double bplank, double pi, int *warn)
{
    const int N = w->N, n = w->n, ncut = w->ncut, lyrcut = w->lyrcut;
    const int ncd = 3 * n - 1, lda = 3 * ncd + 1, nshift = lda - 2 * N + 1;
    const int ncoltot = N * ncut;
    double *cband = w->cband, *b = w->b;
    memset(cband, 0, sizeof(double) * (size_t)lda * ncoltot);
#define CB(i, j) cband[((i) - 1) + (size_t)((j) - 1) * lda]
    int ncol = 0;
    for (int lc = 1; lc <= ncut; lc++) {
        for (int iq = 1; iq <= n; iq++)
            w->wk[iq - 1] = exp(KK(iq - 1, lc - 1) * w->dtaucp[lc - 1]);
        int jcol = 0;
        for (int iq = 1; iq <= n; iq++) {
            ncol++;
            int irow = nshift - jcol;
            for (int jq = 1; jq <= N; jq++) {
                CB(irow + N, ncol) = GC(jq - 1, iq - 1, lc - 1);
                CB(irow, ncol) = -GC(jq - 1, iq - 1, lc - 1) * w->wk[iq - 1];
                irow++;
            }
            jcol++;
        }
        for (int iq = n + 1; iq <= N; iq++) {
            ncol++;
            int irow = nshift - jcol;
            for (int jq = 1; jq <= N; jq++) {
                CB(irow + N, ncol) = GC(jq - 1, iq - 1, lc - 1) * w->wk[N - iq];
                CB(irow, ncol) = -GC(jq - 1, iq - 1, lc - 1);
                irow++;
            }
            jcol++;
        }
    }
    /* top boundary (:2887-2915) */
    int jcol = 0;
    for (int iq = 1; iq <= n; iq++) {
        double expa = exp(KK(iq - 1, 0) * w->taucpr[1]);
        int irow = nshift - jcol + n;
        for (int jq = n; jq >= 1; jq--) {
            CB(irow, jcol + 1) = GC(jq - 1, iq - 1, 0) * expa;
            irow++;
        }
        jcol++;
    }
    for (int iq = n + 1; iq <= N; iq++) {
        int irow = nshift - jcol + n;
        for (int jq = n; jq >= 1; jq--) {
            CB(irow, jcol + 1) = GC(jq - 1, iq - 1, 0);
            irow++;
        }
        jcol++;
    }
    /* bottom boundary (:2919-2990); Lambertian BDR = albedo for m=0 */
    int nncol = ncol - N;
    jcol = 0;
    const int noreflect = lyrcut || (w->lamber && delm0 == 0.0); /* disort.f:2929 */
    for (int iq = 1; iq <= n; iq++) {
        nncol++;
        int irow = nshift - jcol + N;
        for (int jq = n + 1; jq <= N; jq++) {
            if (noreflect) {
                CB(irow, nncol) = GC(jq - 1, iq - 1, ncut - 1);
            } else {
                double sum = 0.0;
                for (int k = 1; k <= n; k++)
                    sum += w->cwt[k - 1] * w->cmu[k - 1] *
                           w->bdr[(jq - n - 1) * (n + 1) + k] *
                           GC(n - k, iq - 1, ncut - 1);
                CB(irow, nncol) = GC(jq - 1, iq - 1, ncut - 1) - (1. + delm0) * sum;
            }
            irow++;
        }
        jcol++;
    }
    for (int iq = n + 1; iq <= N; iq++) {
        nncol++;
        int irow = nshift - jcol + N;
        double expa = w->wk[N - iq];
        for (int jq = n + 1; jq <= N; jq++) {
            if (noreflect) {
                CB(irow, nncol) = GC(jq - 1, iq - 1, ncut - 1) * expa;
            } else {
                double sum = 0.0;
                for (int k = 1; k <= n; k++)
                    sum += w->cwt[k - 1] * w->cmu[k - 1] *
                           w->bdr[(jq - n - 1) * (n + 1) + k] *
                           GC(n - k, iq - 1, ncut - 1);
                CB(irow, nncol) = (GC(jq - 1, iq - 1, ncut - 1) - (1. + delm0) * sum) * expa;
            }
            irow++;
        }
        jcol++;
    }
#undef CB

    /* right-hand side (SOLVE0 :3429-3599); b is 1-based below */
    memset(b, 0, sizeof(double) * (size_t)ncoltot);
#define B(i) b[(i) - 1]
#define BDR(iq, jq) w->bdr[((iq) - 1) * (n + 1) + (jq)]
    const double *expbea = w->expbea, *taucpr = w->taucpr, *cwt = w->cwt,
                 *cmu = w->cmu;
    if (mazim > 0 && fbeam > 0.0) {
        /* azimuth-dependent case (disort.f:3436-3471) */
        for (int iq = 1; iq <= n; iq++) {
            B(iq) = -ZZ(n - iq, 0);
            if (lyrcut || w->lamber) {
                B(ncol - n + iq) = -ZZ(iq + n - 1, ncut - 1) * expbea[ncut];
            } else {
                double sum = 0.;
                for (int jq = 1; jq <= n; jq++)
                    sum += cwt[jq - 1] * cmu[jq - 1] * BDR(iq, jq) * ZZ(n - jq, ncut - 1) * expbea[ncut];
                B(ncol - n + iq) = sum + (BDR(iq, 0) * umu0 * fbeam / pi - ZZ(iq + n - 1, ncut - 1)) * expbea[ncut];
            }
        }
        int it = n;
        for (int lc = 1; lc <= ncut - 1; lc++)
            for (int iq = 1; iq <= N; iq++) {
                it++;
                B(it) = (ZZ(iq - 1, lc) - ZZ(iq - 1, lc - 1)) * expbea[lc];
            }
    } else if (fbeam == 0.0) {
        for (int iq = 1; iq <= n; iq++)
            B(iq) = -ZPLK0(n - iq, 0) + fisot + tplank;
        if (lyrcut) {
            for (int iq = 1; iq <= n; iq++)
                B(ncol - n + iq) = -ZPLK0(iq + n - 1, ncut - 1) -
                                   ZPLK1(iq + n - 1, ncut - 1) * taucpr[ncut];
        } else {
            for (int iq = 1; iq <= n; iq++) {
                double sum = 0.;
                for (int jq = 1; jq <= n; jq++)
                    sum += cwt[jq - 1] * cmu[jq - 1] * BDR(iq, jq) *
                           (ZPLK0(n - jq, ncut - 1) +
                            ZPLK1(n - jq, ncut - 1) * taucpr[ncut]);
                B(ncol - n + iq) = 2. * sum + w->bem[iq - 1] * bplank -
                                   ZPLK0(iq + n - 1, ncut - 1) -
                                   ZPLK1(iq + n - 1, ncut - 1) * taucpr[ncut];
            }
        }
        int it = n;
        for (int lc = 1; lc <= ncut - 1; lc++)
            for (int iq = 1; iq <= N; iq++) {
                it++;
                B(it) = ZPLK0(iq - 1, lc) - ZPLK0(iq - 1, lc - 1) +
                        (ZPLK1(iq - 1, lc) - ZPLK1(iq - 1, lc - 1)) * taucpr[lc];
            }
    } else {
        for (int iq = 1; iq <= n; iq++)
            B(iq) = -ZZ(n - iq, 0) - ZPLK0(n - iq, 0) + fisot + tplank;
        if (lyrcut) {
            for (int iq = 1; iq <= n; iq++)
                B(ncol - n + iq) = -ZZ(iq + n - 1, ncut - 1) * expbea[ncut] -
                                   ZPLK0(iq + n - 1, ncut - 1) -
                                   ZPLK1(iq + n - 1, ncut - 1) * taucpr[ncut];
        } else {
            for (int iq = 1; iq <= n; iq++) {
                double sum = 0.;
                for (int jq = 1; jq <= n; jq++)
                    sum += cwt[jq - 1] * cmu[jq - 1] * BDR(iq, jq) *
                           (ZZ(n - jq, ncut - 1) * expbea[ncut] +
                            ZPLK0(n - jq, ncut - 1) +
                            ZPLK1(n - jq, ncut - 1) * taucpr[ncut]);
                B(ncol - n + iq) = 2. * sum +
                    (BDR(iq, 0) * umu0 * fbeam / pi - ZZ(iq + n - 1, ncut - 1)) * expbea[ncut] +
                    w->bem[iq - 1] * bplank - ZPLK0(iq + n - 1, ncut - 1) -
                    ZPLK1(iq + n - 1, ncut - 1) * taucpr[ncut];
            }
        }
        int it = n;
        for (int lc = 1; lc <= ncut - 1; lc++)
            for (int iq = 1; iq <= N; iq++) {
                it++;
                B(it) = (ZZ(iq - 1, lc) - ZZ(iq - 1, lc - 1)) * expbea[lc] +
                        ZPLK0(iq - 1, lc) - ZPLK0(iq - 1, lc - 1) +
                        (ZPLK1(iq - 1, lc) - ZPLK1(iq - 1, lc - 1)) * taucpr[lc];
            }
    }
#undef BDR
    if (gbfa(cband, lda, ncol, ncd, ncd, w->ipvt) != 0) {
        *warn |= 1 << 2;
        return SBDO_SINGULAR;
    }
    gbsl(cband, lda, ncol, ncd, ncd, w->ipvt, b);
    for (int lc = 1; lc <= ncut; lc++) {
        int ipnt = lc * N - n;
        for (int iq = 1; iq <= n; iq++) {
            LL(n - iq, lc - 1) = B(ipnt + 1 - iq);
            LL(iq + n - 1, lc - 1) = B(iq + ipnt);
        }
    }
#undef B
    return 0;
}

/* FLUXES (disort.f:1780-2042) */
static void fluxes(work_t *w, double fbeam, double umu0, double pi,
                   const double *ssalb, double *rfldir, double *rfldn,
                   double *flup, double *dfdt, double *uavg)
{
    const int N = w->N, n = w->n, NT = w->NT;
    memset(w->u0c, 0, sizeof(double) * (size_t)NT * N);
    for (int lu = 0; lu < NT; lu++) {
        rfldir[lu] = rfldn[lu] = flup[lu] = dfdt[lu] = uavg[lu] = 0.0;
        int lyu = w->layru[lu];           /* 1-based */
        if (w->lyrcut && lyu > w->ncut) continue;
        int lc = lyu - 1;
        double fact = 0.0, dirint, fldir, fldn = 0.0;
        if (fbeam > 0.0) {
            fact = exp(-w->utaupr[lu] / umu0);
            dirint = fbeam * fact;
            fldir = umu0 * (fbeam * fact);
            rfldir[lu] = umu0 * fbeam * exp(-w->utau[lu] / umu0);
        } else {
            dirint = 0.0; fldir = 0.0; rfldir[lu] = 0.0;
        }
        for (int iq = 0; iq < N; iq++) {
            double zint = 0.0;
            for (int jq = 0; jq < n; jq++)
                zint += GC(iq, jq, lc) * LL(jq, lc) *
                        exp(-KK(jq, lc) * (w->utaupr[lu] - w->taucpr[lyu]));
            for (int jq = n; jq < N; jq++)
                zint += GC(iq, jq, lc) * LL(jq, lc) *
                        exp(-KK(jq, lc) * (w->utaupr[lu] - w->taucpr[lyu - 1]));
            double u = zint;
            if (fbeam > 0.0) u = zint + ZZ(iq, lc) * fact;
            u = u + ZPLK0(iq, lc) + ZPLK1(iq, lc) * w->utaupr[lu];
            w->u0c[(size_t)lu * N + iq] = u;
            if (iq < n) {
                uavg[lu] += w->cwt[n - 1 - iq] * u;
                fldn += w->cwt[n - 1 - iq] * w->cmu[n - 1 - iq] * u;
            } else {
                uavg[lu] += w->cwt[iq - n] * u;
                flup[lu] += w->cwt[iq - n] * w->cmu[iq - n] * u;
            }
        }
        flup[lu] = 2. * pi * flup[lu];
        fldn = 2. * pi * fldn;
        double fdntot = fldn + fldir;
        rfldn[lu] = fdntot - rfldir[lu];
        uavg[lu] = (2. * pi * uavg[lu] + dirint) / (4. * pi);
        double plsorc = w->xr0[lc] + w->xr1[lc] * w->utaupr[lu];
        dfdt[lu] = (1. - ssalb[lc]) * 4. * pi * (uavg[lu] - plsorc);
    }
}

/* CMPINT (disort.f:1658-1778) */
static void cmpint(work_t *w, int mazim, double fbeam, double umu0, int plank)
{
    const int N = w->N, n = w->n, NT = w->NT, NU = w->NU;
    for (int lu = 0; lu < NT; lu++) {
        int lyu = w->layru[lu];
        if (w->lyrcut && lyu > w->ncut) continue;
        int lc = lyu - 1;
        for (int iq = 0; iq < N; iq++) {
            double zint = 0.0;
            for (int jq = 0; jq < n; jq++)
                zint += GC(iq, jq, lc) * LL(jq, lc) *
                        exp(-KK(jq, lc) * (w->utaupr[lu] - w->taucpr[lyu]));
            for (int jq = n; jq < N; jq++)
                zint += GC(iq, jq, lc) * LL(jq, lc) *
                        exp(-KK(jq, lc) * (w->utaupr[lu] - w->taucpr[lyu - 1]));
            double u = zint;
            if (fbeam > 0.0) u = zint + ZZ(iq, lc) * exp(-w->utaupr[lu] / umu0);
            if (plank && mazim == 0)
                u = u + ZPLK0(iq, lc) + ZPLK1(iq, lc) * w->utaupr[lu];
            w->uum[(size_t)lu * NU + iq] = u;
        }
    }
}

/* USRINT (disort.f:4355-4793), Lambertian surface */
static void usrint(work_t *w, int mazim, double delm0, double fbeam,
                   double umu0, double fisot, double tplank, double bplank,
                   int plank, double pi, const double *umu)
{
    const int N = w->N, n = w->n, NT = w->NT, NU = w->NU, L = w->L,
              ncut = w->ncut, lyrcut = w->lyrcut;
    const double *taucpr = w->taucpr, *utaupr = w->utaupr,
                 *expbea = w->expbea, *dtaucp = w->dtaucp;
    double exp0 = 0.0, exp1 = 0.0, exp2 = 0.0; /* SAVEd in the reference */
    for (int lc = 0; lc < ncut; lc++)
        for (int iq = 0; iq < N; iq++)
            for (int iu = 0; iu < NU; iu++) GU(iu, iq, lc) *= LL(iq, lc);
    for (int lu = 0; lu < NT; lu++) {
        if (fbeam > 0.0) exp0 = exp(-utaupr[lu] / umu0);
        int lyu = w->layru[lu];
        for (int iu = 0; iu < NU; iu++) {
            if (lyrcut && lyu > ncut) continue;
            int negumu = umu[iu] < 0.0;
            int lyrstr, lyrend; double sgn;
            if (negumu) { lyrstr = 1; lyrend = lyu - 1; sgn = -1.0; }
            else { lyrstr = lyu + 1; lyrend = ncut; sgn = 1.0; }
            double palint = 0.0, plkint = 0.0;
            for (int lc = lyrstr; lc <= lyrend; lc++) {
                double dtau = dtaucp[lc - 1];
                exp1 = exp((utaupr[lu] - taucpr[lc - 1]) / umu[iu]);
                exp2 = exp((utaupr[lu] - taucpr[lc]) / umu[iu]);
                if (plank && mazim == 0) {
                    double f0n = sgn * (exp1 - exp2);
                    double f1n = sgn * ((taucpr[lc - 1] + umu[iu]) * exp1 -
                                        (taucpr[lc] + umu[iu]) * exp2);
                    plkint += w->z0u[(size_t)(lc - 1) * NU + iu] * f0n +
                              w->z1u[(size_t)(lc - 1) * NU + iu] * f1n;
                }
                if (fbeam > 0.0) {
                    double denom = 1. + umu[iu] / umu0, expn;
                    if (fabs(denom) < 0.0001) expn = (dtau / umu0) * exp0;
                    else expn = (exp1 * expbea[lc - 1] - exp2 * expbea[lc]) * sgn / denom;
                    palint += w->zbeam[(size_t)(lc - 1) * NU + iu] * expn;
                }
                for (int iq = 0; iq < n; iq++) {
                    w->wk[iq] = exp(KK(iq, lc - 1) * dtau);
                    double denom = 1.0 + umu[iu] * KK(iq, lc - 1), expn;
                    if (fabs(denom) < 0.0001) expn = dtau / umu[iu] * exp2;
                    else expn = sgn * (exp1 * w->wk[iq] - exp2) / denom;
                    palint += GU(iu, iq, lc - 1) * expn;
                }
                for (int iq = n; iq < N; iq++) {
                    double denom = 1.0 + umu[iu] * KK(iq, lc - 1), expn;
                    if (fabs(denom) < 0.0001) expn = -dtau / umu[iu] * exp1;
                    else expn = sgn * (exp1 - exp2 * w->wk[N - 1 - iq]) / denom;
                    palint += GU(iu, iq, lc - 1) * expn;
                }
            }
            /* layer containing the level (:4623-4729) */
            double dtau1 = utaupr[lu] - taucpr[lyu - 1];
            double dtau2 = utaupr[lu] - taucpr[lyu];
            int skip = (fabs(dtau1) < 1.e-6 && negumu) ||
                       (fabs(dtau2) < 1.e-6 && !negumu);
            if (!skip) {
                if (negumu) exp1 = exp(dtau1 / umu[iu]);
                else exp2 = exp(dtau2 / umu[iu]);
                if (fbeam > 0.0) {
                    double denom = 1. + umu[iu] / umu0, expn;
                    if (fabs(denom) < 0.0001) expn = (dtau1 / umu0) * exp0;
                    else if (negumu) expn = (exp0 - expbea[lyu - 1] * exp1) / denom;
                    else expn = (exp0 - expbea[lyu] * exp2) / denom;
                    palint += w->zbeam[(size_t)(lyu - 1) * NU + iu] * expn;
                }
                double dtau = dtaucp[lyu - 1];
                for (int iq = 0; iq < n; iq++) {
                    double kq = KK(iq, lyu - 1);
                    double denom = 1. + umu[iu] * kq, expn;
                    if (fabs(denom) < 0.0001) expn = -dtau2 / umu[iu] * exp2;
                    else if (negumu) expn = (exp(-kq * dtau2) - exp(kq * dtau) * exp1) / denom;
                    else expn = (exp(-kq * dtau2) - exp2) / denom;
                    palint += GU(iu, iq, lyu - 1) * expn;
                }
                for (int iq = n; iq < N; iq++) {
                    double kq = KK(iq, lyu - 1);
                    double denom = 1. + umu[iu] * kq, expn;
                    if (fabs(denom) < 0.0001) expn = -dtau1 / umu[iu] * exp1;
                    else if (negumu) expn = (exp(-kq * dtau1) - exp1) / denom;
                    else expn = (exp(-kq * dtau1) - exp(-kq * dtau) * exp2) / denom;
                    palint += GU(iu, iq, lyu - 1) * expn;
                }
                if (plank && mazim == 0) {
                    double expn, fact;
                    if (negumu) { expn = exp1; fact = taucpr[lyu - 1] + umu[iu]; }
                    else { expn = exp2; fact = taucpr[lyu] + umu[iu]; }
                    double f0n = 1. - expn;
                    double f1n = utaupr[lu] + umu[iu] - fact * expn;
                    plkint += w->z0u[(size_t)(lyu - 1) * NU + iu] * f0n +
                              w->z1u[(size_t)(lyu - 1) * NU + iu] * f1n;
                }
            }
            /* boundary contributions (:4735-4781) */
            double bndint = 0.0;
            if (negumu && mazim == 0) {
                bndint = (fisot + tplank) * exp(utaupr[lu] / umu[iu]);
            } else if (!negumu) {
                if (!(lyrcut || (w->lamber && mazim > 0))) {  /* disort.f:4744 */
                    for (int jq = n; jq < N; jq++)
                        w->wk[jq] = exp(-KK(jq, L - 1) * dtaucp[L - 1]);
                    double bnddfu = 0.0;
                    for (int iq = n; iq >= 1; iq--) {
                        double dfuint = 0.0;
                        for (int jq = 0; jq < n; jq++)
                            dfuint += GC(iq - 1, jq, L - 1) * LL(jq, L - 1);
                        for (int jq = n; jq < N; jq++)
                            dfuint += GC(iq - 1, jq, L - 1) * LL(jq, L - 1) * w->wk[jq];
                        if (fbeam > 0.0) dfuint += ZZ(iq - 1, L - 1) * expbea[L];
                        dfuint += delm0 * (ZPLK0(iq - 1, L - 1) + ZPLK1(iq - 1, L - 1) * taucpr[L]);
                        bnddfu += (1. + delm0) * w->rmu[iu * (n + 1) + (n + 1 - iq)] *
                                  w->cmu[n - iq] * w->cwt[n - iq] * dfuint;
                    }
                    double bnddir = 0.0;
                    if (fbeam > 0.0)
                        bnddir = umu0 * fbeam / pi * w->rmu[iu * (n + 1) + 0] * expbea[L];
                    bndint = (bnddfu + bnddir + delm0 * w->emu[iu] * bplank) *
                             exp((utaupr[lu] - taucpr[L]) / umu[iu]);
                }
            }
            w->uum[(size_t)lu * NU + iu] = palint + plkint + bndint;
        }
    }
}

/* XIFUNC (disort.f:4795-4858) */
static double xifunc(double umu1, double umu2, double umu3, double tau)
{
    double x1 = 1. / umu1 - 1. / umu2, x2 = 1. / umu1 - 1. / umu3;
    double exp1 = exp(-tau / umu1);
    if (umu2 == umu3 && umu1 == umu2)
        return tau * tau * exp1 / (2. * umu1 * umu2);
    if (umu2 == umu3 && umu1 != umu2)
        return ((tau - 1. / x1) * exp(-tau / umu2) + exp1 / x1) / (x1 * umu1 * umu2);
    if (umu2 != umu3 && umu1 == umu2)
        return ((exp(-tau / umu3) - exp1) / x2 - tau * exp1) / (x2 * umu1 * umu2);
    if (umu2 != umu3 && umu1 == umu3)
        return ((exp(-tau / umu2) - exp1) / x1 - tau * exp1) / (x1 * umu1 * umu2);
    return ((exp(-tau / umu3) - exp1) / x2 - (exp(-tau / umu2) - exp1) / x1) /
           (x2 * umu1 * umu2);
}

/* SINSCA (disort.f:2996-3097); tau is 0-based level array, layru 1-based */
static double sinsca(double dither, int layru, int nlyr, const double *phase,
                     const double *omega, const double *tau, double umu,
                     double umu0, double utau, double fbeam, double pi)
{
    double s = 0., exp0 = exp(-utau / umu0), exp1;
    if (fabs(umu + umu0) <= dither) {
        for (int lyr = 1; lyr <= layru - 1; lyr++)
            s += omega[lyr - 1] * phase[lyr - 1] * (tau[lyr] - tau[lyr - 1]);
        return fbeam / (4. * pi * umu0) * exp0 *
               (s + omega[layru - 1] * phase[layru - 1] * (utau - tau[layru - 1]));
    }
    if (umu > 0.) {
        for (int lyr = layru; lyr <= nlyr; lyr++) {
            exp1 = exp(-((tau[lyr] - utau) / umu + tau[lyr] / umu0));
            s += omega[lyr - 1] * phase[lyr - 1] * (exp0 - exp1);
            exp0 = exp1;
        }
    } else {
        for (int lyr = layru; lyr >= 1; lyr--) {
            exp1 = exp(-((tau[lyr - 1] - utau) / umu + tau[lyr - 1] / umu0));
            s += omega[lyr - 1] * phase[lyr - 1] * (exp0 - exp1);
            exp0 = exp1;
        }
    }
    return fbeam / (4. * pi * (1. + umu / umu0)) * s;
}

/* SECSCA (disort.f:2299-2452) */
static double secsca(double ctheta, const double *flyr, int layru, int ldp,
                     int nmom, int nstr, const double *pmom,
                     const double *ssalb, const double *dtauc,
                     const double *tauc, double umu, double umu0, double utau,
                     double fbeam, double pi)
{
#define PMOM(k, lc) pmom[(size_t)((lc) - 1) * ldp + (k)]
    const double zero = (double)1E-4f;
    double dtau = utau - tauc[layru - 1];
    double wbar = ssalb[layru - 1] * dtau;
    double fbar = flyr[layru - 1] * wbar;
    double stau = dtau;
    for (int lyr = 1; lyr <= layru - 1; lyr++) {
        wbar += ssalb[lyr - 1] * dtauc[lyr - 1];
        fbar += ssalb[lyr - 1] * dtauc[lyr - 1] * flyr[lyr - 1];
        stau += dtauc[lyr - 1];
    }
    if (wbar <= zero || fbar <= zero || stau <= zero || fbeam <= zero) return 0.0;
    fbar = fbar / wbar;
    wbar = wbar / stau;
    double pspike = 1., gbar = 1., plm1 = 1., plm2 = 0., pl;
    for (int k = 1; k <= nstr - 1; k++) {
        pl = ((2 * k - 1) * ctheta * plm1 - (k - 1) * plm2) / k;
        plm2 = plm1; plm1 = pl;
        pspike += (2. * gbar - gbar * gbar) * (2 * k + 1) * pl;
    }
    for (int k = nstr; k <= nmom; k++) {
        pl = ((2 * k - 1) * ctheta * plm1 - (k - 1) * plm2) / k;
        plm2 = plm1; plm1 = pl;
        dtau = utau - tauc[layru - 1];
        gbar = PMOM(k, layru) * ssalb[layru - 1] * dtau;
        for (int lyr = 1; lyr <= layru - 1; lyr++)
            gbar += PMOM(k, lyr) * ssalb[lyr - 1] * dtauc[lyr - 1];
        if (fbar * wbar * stau <= zero) gbar = 0.0;
        else gbar = gbar / (fbar * wbar * stau);
        pspike += (2. * gbar - gbar * gbar) * (2 * k + 1) * pl;
    }
    double umu0p = umu0 / (1. - fbar * wbar);
    return fbeam / (4. * pi) * (fbar * wbar) * (fbar * wbar) / (1. - fbar * wbar) *
           pspike * xifunc(-umu, umu0p, umu0p, utau);
#undef PMOM
}

/* INTCOR (disort.f:2044-2297): Nakajima-Tanaka TMS/IMS corrections */
static void intcor(work_t *w, double dither, double fbeam, int ldp, int nmom,
                   int nphi, const double *pmom, const double *ssalb,
                   const double *dtauc, const double *umu, double umu0,
                   double pi, double rpd, double *uu)
{
    const int N = w->N, NT = w->NT, NU = w->NU, ncut = w->ncut,
              lyrcut = w->lyrcut;
    double *phasa = (double *)calloc(ncut, sizeof(double)),
           *phast = (double *)calloc(ncut, sizeof(double)),
           *phasm = (double *)calloc(ncut, sizeof(double));
    const double dtheta = 10.;
    double theta0 = 0.0, thetap = 0.0;
#define PMOM(k, lc) pmom[(size_t)(lc) * ldp + (k)]
    for (int iu = 0; iu < NU; iu++) {
        if (umu[iu] < 0.) {
            theta0 = acos(-umu0) / rpd;
            thetap = acos(umu[iu]) / rpd;
        }
        for (int jp = 0; jp < nphi; jp++) {
            double ctheta = -umu0 * umu[iu] +
                sqrt((1. - umu0 * umu0) * (1. - umu[iu] * umu[iu])) * cos(w->phirad[jp]);
            for (int lc = 0; lc < ncut; lc++) { phasa[lc] = 1.; phasm[lc] = 1.; }
            double plm1 = 1., plm2 = 0.;
            for (int k = 1; k <= nmom; k++) {
                double pl = ((2 * k - 1) * ctheta * plm1 - (k - 1) * plm2) / k;
                plm2 = plm1; plm1 = pl;
                for (int lc = 0; lc < ncut; lc++)
                    phasa[lc] += (2 * k + 1) * pl * PMOM(k, lc);
                if (k <= N - 1)
                    for (int lc = 0; lc < ncut; lc++)
                        phasm[lc] += (2 * k + 1) * pl * (PMOM(k, lc) - w->flyr[lc]) /
                                     (1. - w->flyr[lc]);
            }
            for (int lc = 0; lc < ncut; lc++)
                phast[lc] = phasa[lc] / (1. - w->flyr[lc] * ssalb[lc]);
            for (int lu = 0; lu < NT; lu++) {
                if (!lyrcut || w->layru[lu] < ncut) {
                    double ussndm = sinsca(dither, w->layru[lu], ncut, phast, ssalb,
                                           w->taucpr, umu[iu], umu0, w->utaupr[lu], fbeam, pi);
                    double ussp = sinsca(dither, w->layru[lu], ncut, phasm, w->oprim,
                                         w->taucpr, umu[iu], umu0, w->utaupr[lu], fbeam, pi);
                    uu[((size_t)jp * NT + lu) * NU + iu] += ussndm - ussp;
                }
            }
            if (umu[iu] < 0. && fabs(theta0 - thetap) <= dtheta) {
                int ltau = 1;
                if (w->utau[0] <= dither) ltau = 2;
                for (int lu = ltau; lu <= NT; lu++) {
                    if (!lyrcut || w->layru[lu - 1] < ncut) {
                        double duims = secsca(ctheta, w->flyr, w->layru[lu - 1], ldp, nmom, N,
                                              pmom, ssalb, dtauc, w->tauc, umu[iu], umu0,
                                              w->utau[lu - 1], fbeam, pi);
                        uu[((size_t)jp * NT + (lu - 1)) * NU + iu] -= duims;
                    }
                }
            }
        }
    }
#undef PMOM
    free(phasa); free(phast); free(phasm);
}

/* RATIO (disort.f:6159-6266) */
static double ratio(double a, double b)
{
    const double tiny = DBL_MIN, huge = DBL_MAX;
    const double powmax = log10(huge), powmin = log10(tiny);
    double r;
    if (a == 0.0) return (b == 0.0) ? 1.0 : 0.0;
    if (b == 0.0) return copysign(huge, a);
    double absa = fabs(a), absb = fabs(b);
    double powa = log10(absa), powb = log10(absb);
    if (absa < tiny && absb < tiny) r = 1.0;
    else if (powa - powb >= powmax) r = huge;
    else if (powa - powb <= powmin) r = tiny;
    else r = absa / absb;
    if ((a > 0.0 && b < 0.0) || (a < 0.0 && b > 0.0)) r = -r;
    return r;
}

/* ------------------------------------------------------------------ */
/* DISORT main (disort.f:472-871)                                      */
/* ------------------------------------------------------------------ */
/* ------------------------------------------------------------------------------------------
 * BDREF (spectra.f:249-296) and its three models, for non-Lambertian surfaces (LAMBER = 0).
 * The model and its parameters live in file-scope state like the reference's module albblk
 * (spectra.f:20-26); the ocean model's wavelength-dependent terms (INDWAT spectra.f:594,
 * MORCASIWAT spectra.f:467: table look-ups) are handed over by the caller.
 * ------------------------------------------------------------------------------------------ */
static struct {
    int ibdrf;                 /* 1 ocean, 2 Hapke, 3 Ross-Li (spectra.f:139-162) */
    double p[5];               /* sc(1..5) as suralb stores them                  */
    double nr, ni, rsw;        /* ocean: refractive index, sub-surface reflectance */
} g_sfc;

void sbdo_set_bdref(int ibdrf, const double *sc, double nr, double ni, double rsw)
{
    g_sfc.ibdrf = ibdrf;
    for (int i = 0; i < 5; i++) g_sfc.p[i] = sc ? sc[i] : 0.0;
    g_sfc.nr = nr; g_sfc.ni = ni; g_sfc.rsw = rsw;
}

static const double kPiParams = 3.1415926536;      /* params.f:29 */

/* fresnel, spectra.f:1320-1352 */
static double bd_fresnel(double nr, double ni, double coschi, double sinchi)
{
    double a1 = fabs(nr * nr - ni * ni - sinchi * sinchi);
    double a2 = sqrt(pow(nr * nr - ni * ni - sinchi * sinchi, 2.) + 4 * nr * nr * ni * ni);
    double u = sqrt(0.5 * (a1 + a2)), v = sqrt(0.5 * (-a1 + a2));
    double rr2 = ((coschi - u) * (coschi - u) + v * v) / ((coschi + u) * (coschi + u) + v * v);
    double b1 = (nr * nr - ni * ni) * coschi, b2 = 2 * nr * ni * coschi;
    double rl2 = ((b1 - u) * (b1 - u) + (b2 + v) * (b2 + v)) / ((b1 + u) * (b1 + u) + (b2 - v) * (b2 - v));
    return (rr2 + rl2) / 2.;
}

/* sunglint, spectra.f:1224-1316 (wind-direction average of the Cox-Munk distribution) */
static double bd_sunglint(double wndspd, double nr, double ni, double csin, double cvin, double phi)
{
    const double pi = kPiParams;
    double cs = csin > (double)0.05f ? csin : (double)0.05f;
    double cv = cvin > (double)0.05f ? cvin : (double)0.05f;
    double ss = sqrt(1. - cs * cs), sv = sqrt(1. - cv * cv);
    double zx = -sv * sin(pi - phi) / (cs + cv);
    double zy = (ss + sv * cos(pi - phi)) / (cs + cv);
    double tilt = atan(sqrt(zx * zx + zy * zy));
    double sigmac = (double)0.003f + (double)0.00192f * wndspd;
    double sigmau = (double)0.00316f * wndspd;
    double c40 = (double)0.40f, c22 = (double)0.12f, c04 = (double)0.23f;
    double r2 = zx * zx + zy * zy;
    double axe2 = (double).5f * r2 / sigmac, axn2 = (double).5f * r2 / sigmau;
    double q4 = 3 * pow(zx, 4) + 6 * zx * zx * zy * zy + 3 * pow(zy, 4);
    double axe4 = q4 / (8 * sigmac * sigmac), axn4 = q4 / (8 * sigmau * sigmau);
    double axe2xn2 = (pow(zx, 4) + 10 * zx * zx * zy * zy + pow(zy, 4)) / (8 * sigmau * sigmac);
    double coef = 1.;
    coef = coef + c40 / 24. * (axe4 - 6 * axe2 + 3);
    coef = coef + c04 / 24. * (axn4 - 6 * axn2 + 3);
    coef = coef + c22 / 4. * (axe2xn2 - axn2 - axe2 + 1);
    coef = coef / (2. * pi * sqrt(sigmau) * sqrt(sigmac));
    double proba = coef * exp(-(axe2 + axn2) / 2.);
    double cos2chi = cv * cs + sv * ss * cos(pi - phi);
    if (cos2chi > 1.0) cos2chi = 0.99999999999;
    if (cos2chi < -1.0) cos2chi = -0.99999999999;
    double coschi = sqrt(0.5 * (1 + cos2chi)), sinchi = sqrt(0.5 * (1 - cos2chi));
    double r1 = bd_fresnel(nr, ni, coschi, sinchi);
    return pi * r1 * proba / (4. * cs * cv * pow(cos(tilt), 4));
}

/* bdref(wvnmlo, wvnmhi, mur, mui, phir), spectra.f:249 */
static double bdref(double mur, double mui, double phir)
{
    const double pi = kPiParams;
    const double *p = g_sfc.p;
    if (g_sfc.ibdrf == 1) {                 /* seabdrf(wl, mus = mui, muv = mur, phir), spectra.f:421 */
        double wndspd = p[1];
        double wndwt = (double)2.951e-6f * pow(wndspd, (double)3.52f);
        double rfoam = wndwt * (double)0.22f;
        double rgl = bd_sunglint(wndspd, g_sfc.nr, g_sfc.ni, mui, mur, phir);
        return rfoam + (1. - wndwt) * rgl + (1. - rfoam) * g_sfc.rsw;
    }
    if (g_sfc.ibdrf == 2) {                 /* hapkbdrf(ui = mui, ur = mur), spectra.f:298 */
        double hssa = p[0], hasym = p[1], hotspt = p[2], hotwdth = p[3];
        double ui = mui, ur = mur;
        double coss = ui * ur + sqrt(1. - ur * ur) * sqrt(1. - ui * ui) * cos(pi - phir);
        double s = acos(coss);
        double pfun = (1. - hasym * hasym) / pow(1 + hasym * hasym + 2 * hasym * coss, 1.5);
        double pfun0 = (1. - hasym * hasym) / pow(1 + hasym, 3.);
        double b0 = hotspt / (hssa * pfun0);
        double bfun = b0 / (1. + tan(s / 2) / hotwdth);
        double hfunr = (1. + 2 * ur) / (1. + 2. * ur * sqrt(1. - hssa));
        double hfuni = (1. + 2 * ui) / (1. + 2. * ui * sqrt(1. - hssa));
        double bd = (1. + bfun) * pfun + hfunr * hfuni - 1.;
        return (double).25f * hssa * bd / (ur + ui);
    }
    /* rtlsbdrf(mui, mur, phir), spectra.f:350 */
    double rliso = p[0], rlvol = p[1], rlgeo = p[2], rlhot = p[3], rlwdth = p[4];
    double ui = mui > (double).01f ? mui : (double).01f;
    double ur = mur > (double).01f ? mur : (double).01f;
    double cosra = cos(pi - phir);
    double coss = ui * ur + sqrt(1. - ur * ur) * sqrt(1. - ui * ui) * cosra;
    coss = coss < -1. ? -1. : (coss > 1. ? 1. : coss);
    double s = acos(coss), sins = sin(s);
    double f1 = (pi / 2 - s) * coss + sins;
    f1 = f1 / (ui + ur) - pi / 4.;
    double vza = acos(ur), sza = acos(ui);
    double tanvzap = rlwdth * tan(vza), tanszap = rlwdth * tan(sza);
    double vzap, szap;
    if (rlwdth == 1.) { vzap = vza; szap = sza; }
    else { vzap = atan(tanvzap); szap = atan(tanszap); }
    double cossp = cos(szap) * cos(vzap) + sin(szap) * sin(vzap) * cosra;
    cossp = cossp < -1. ? -1. : (cossp > 1. ? 1. : cossp);
    double dd = tanszap * tanszap + tanvzap * tanvzap - 2 * tanszap * tanvzap * cosra;
    double secsum = 1. / cos(szap) + 1. / cos(vzap);
    double cost = rlhot * sqrt(dd + pow(tanszap * tanvzap * sin(pi - phir), 2.));
    cost = cost / secsum;
    cost = cost < -1. ? -1. : (cost > 1. ? 1. : cost);
    double t = acos(cost);
    double f2 = (t - sin(t) * cost) * secsum / pi;
    f2 = f2 - 1. / cos(vzap) + (double).5f * (1. + cossp) / (cos(szap) * cos(vzap));
    return rliso + rlvol * f1 + rlgeo * f2;
}

/* test hook: the model set with sbdo_set_bdref evaluated at one geometry */
double sbdo_bdref_eval(double mur, double mui, double phir) { return bdref(mur, mui, phir); }

#define SBDO_NMUG 50
/* SURFAC, non-Lambertian branch (disort.f:3765-3907): Fourier coefficients of the
 * bidirectional reflectivity at the computational and user angles by a 50-point quadrature
 * in azimuth, directional emissivities from a 25 x 50 quadrature. */
static void surfac_brdf(int n, int NU, int mazim, double delm0, double fbeam, double umu0,
                        double pi, const double *cmu, const double *umu, int user,
                        double *bdr, double *bem, double *rmu, double *emu)
{
    double gmu[SBDO_NMUG], gwt[SBDO_NMUG];
    sbdo_qgausn(SBDO_NMUG / 2, gmu, gwt);
    for (int k = 0; k < SBDO_NMUG / 2; k++) { gmu[k + SBDO_NMUG / 2] = -gmu[k]; gwt[k + SBDO_NMUG / 2] = gwt[k]; }
    for (int iq = 0; iq < n; iq++) {
        for (int jq = 1; jq <= n; jq++) {
            double sum = 0.0;
            for (int k = 0; k < SBDO_NMUG; k++)
                sum += gwt[k] * bdref(cmu[iq], cmu[jq - 1], pi * gmu[k]) * cos(mazim * pi * gmu[k]);
            bdr[iq * (n + 1) + jq] = 0.5 * (2. - delm0) * sum;
        }
        if (fbeam > 0.0) {
            double sum = 0.0;
            for (int k = 0; k < SBDO_NMUG; k++)
                sum += gwt[k] * bdref(cmu[iq], umu0, pi * gmu[k]) * cos(mazim * pi * gmu[k]);
            bdr[iq * (n + 1)] = 0.5 * (2. - delm0) * sum;
        }
    }
    if (mazim == 0)
        for (int iq = 0; iq < n; iq++) {
            double dref = 0.0;
            for (int jg = 0; jg < SBDO_NMUG; jg++) {
                double sum = 0.0;
                for (int k = 0; k < SBDO_NMUG / 2; k++)
                    sum += gwt[k] * gmu[k] * bdref(cmu[iq], gmu[k], pi * gmu[jg]);
                dref += gwt[jg] * sum;
            }
            bem[iq] = 1.0 - dref;
        }
    if (!user) return;
    for (int iu = 0; iu < NU; iu++) {
        if (!(umu[iu] > 0.0)) continue;
        for (int iq = 1; iq <= n; iq++) {
            double sum = 0.0;
            for (int k = 0; k < SBDO_NMUG; k++)
                sum += gwt[k] * bdref(umu[iu], cmu[iq - 1], pi * gmu[k]) * cos(mazim * pi * gmu[k]);
            rmu[iu * (n + 1) + iq] = 0.5 * (2. - delm0) * sum;
        }
        if (fbeam > 0.0) {
            double sum = 0.0;
            for (int k = 0; k < SBDO_NMUG; k++)
                sum += gwt[k] * bdref(umu[iu], umu0, pi * gmu[k]) * cos(mazim * pi * gmu[k]);
            rmu[iu * (n + 1)] = 0.5 * (2. - delm0) * sum;
        }
        if (mazim == 0) {
            double dref = 0.0;
            for (int jg = 0; jg < SBDO_NMUG; jg++) {
                double sum = 0.0;
                for (int k = 0; k < SBDO_NMUG / 2; k++)
                    sum += gwt[k] * gmu[k] * bdref(umu[iu], gmu[k], pi * gmu[jg]);
                dref += gwt[jg] * sum;
            }
            emu[iu] = 1.0 - dref;
        }
    }
}

int sbdo_disort(const sbdo_input *in, const double *dtauc_in,
                const double *ssalb_in, const double *pmom,
                const double *temper, const double *utau_in,
                const double *umu_in, const double *phi, double *rfldir,
                double *rfldn, double *flup, double *dfdt, double *uavg,
                double *uu, double *u0u, int *warn_out)
{
    const int L = in->nlyr, N = in->nstr, n = N / 2, ldp = in->nmom + 1;
    const double pi = ref_pi(), rpd = pi / 180.0;
    double dither = 10. * R1MACH4;
    if (dither < 1.e-10) dither = 10. * dither;
    const double fbeam = in->fbeam, umu0 = in->umu0;
    int warn = 0, status = 0;
    int corint = in->corint;
    const int plank = in->plank, onlyfl = in->onlyfl;

    /* CHEKIN essentials (disort.f:4920-5155) */
    if (N < 4 || N % 2 != 0 || L < 1) return SBDO_BAD_INPUT;
    if (!in->lamber && (g_sfc.ibdrf < 1 || g_sfc.ibdrf > 3)) return SBDO_BAD_INPUT;
    if (in->nmom < N) return SBDO_BAD_INPUT;
    if (fbeam < 0.0 || (fbeam > 0.0 && (umu0 <= 0.0 || umu0 > 1.0))) return SBDO_BAD_INPUT;
    if ((in->lamber && (in->albedo < 0.0 || in->albedo > 1.0)) || in->fisot < 0.0) return SBDO_BAD_INPUT;
    if (plank && (in->wvnmlo < 0.0 || in->wvnmhi <= in->wvnmlo ||
                  in->temis < 0.0 || in->temis > 1.0 || in->btemp < 0.0 ||
                  in->ttemp < 0.0)) return SBDO_BAD_INPUT;
    if (in->usrang && !onlyfl && in->numu < 1) return SBDO_BAD_INPUT;
    if (!onlyfl && in->nphi < 1) return SBDO_BAD_INPUT;

    const int NT = in->usrtau ? in->ntau : L + 1;
    int usrang_eff = in->usrang && !onlyfl;
    /* SETDIS :2655-2669: without user angles the quadrature angles are used */
    const int NU = usrang_eff ? in->numu : N;

    work_t W; memset(&W, 0, sizeof W);
    work_t *w = &W;
    double *dtauc = NULL, *ssalb = NULL, *umu = NULL;
    /* pass 0 measures the arena, pass 1 hands out zeroed memory */
    for (int pass = 0; pass < 2; pass++) {
        tl_measure = (pass == 0);
        tl_off = 0;
        w->N = N; w->n = n; w->L = L; w->NT = NT; w->NU = NU;
        w->cmu = dalloc(N); w->cwt = dalloc(N);
        w->gl = dalloc((size_t)L * (N + 1));
        w->dtaucp = dalloc(L); w->oprim = dalloc(L); w->flyr = dalloc(L);
        w->taucpr = dalloc(L + 1); w->expbea = dalloc(L + 1); w->tauc = dalloc(L + 1);
        w->pkag = dalloc(L + 1);
        w->utau = dalloc(NT); w->utaupr = dalloc(NT);
        w->layru = ialloc(NT);
        w->ylm0 = dalloc(N + 1); w->ylmc = dalloc((size_t)N * (N + 1));
        w->ylmu = dalloc((size_t)NU * (N + 1));
        w->cc = dalloc((size_t)N * N); w->evecc = dalloc((size_t)N * N);
        w->array = dalloc((size_t)N * N);
        w->amb = dalloc((size_t)n * n); w->apb = dalloc((size_t)n * n); w->eval = dalloc(n);
        w->gc = dalloc((size_t)L * N * N);
        w->gu = dalloc((size_t)L * N * NU);
        w->kk = dalloc((size_t)L * N); w->ll = dalloc((size_t)L * N);
        w->zz = dalloc((size_t)L * N); w->zplk0 = dalloc((size_t)L * N);
        w->zplk1 = dalloc((size_t)L * N);
        w->xr0 = dalloc(L); w->xr1 = dalloc(L);
        w->zbeam = dalloc((size_t)L * NU); w->z0u = dalloc((size_t)L * NU);
        w->z1u = dalloc((size_t)L * NU);
        w->bdr = dalloc((size_t)n * (n + 1)); w->bem = dalloc(n);
        w->rmu = dalloc((size_t)NU * (n + 1)); w->emu = dalloc(NU);
        w->wk = dalloc(2 * N + 2); w->z0 = dalloc(N); w->z1 = dalloc(N);
        w->zj = dalloc(N); w->psi0 = dalloc(N + 1); w->psi1 = dalloc(N + 1);
        {
            int ncd = 3 * n - 1, lda = 3 * ncd + 1;
            w->cband = dalloc((size_t)lda * N * L);
            w->b = dalloc((size_t)N * L);
            w->ipvt = ialloc((size_t)N * L + N);
        }
        w->uum = dalloc((size_t)NT * NU); w->u0c = dalloc((size_t)NT * N);
        w->phirad = dalloc(in->nphi > 0 ? in->nphi : 1);
        dtauc = dalloc(L); ssalb = dalloc(L); umu = dalloc(NU);
        if (pass == 0) {
            if (tl_off > tl_cap) {
                free(tl_base);
                tl_cap = tl_off + tl_off / 4;
                tl_base = (char *)malloc(tl_cap);
                if (!tl_base) { tl_cap = 0; return SBDO_BAD_INPUT; }
            }
        }
    }
    /* the cband region is zeroed by setmtx_solve0 itself; zero the rest */
    memset(tl_base, 0, (size_t)((char *)w->cband - tl_base));
    memset(w->b, 0, tl_off - (size_t)((char *)w->b - tl_base));

    /* cumulative optical depth, SSALB dither (disort.f:482-489) */
    for (int lc = 0; lc < L; lc++) {
        ssalb[lc] = ssalb_in[lc];
        if (ssalb[lc] == 1.0) ssalb[lc] = 1.0 - dither;
        dtauc[lc] = dtauc_in[lc];
        w->tauc[lc + 1] = w->tauc[lc] + dtauc[lc];   /* uses unclipped DTAUC */
    }
    /* CHEKIN (disort.f:4938-4982) */
    for (int lc = 0; lc < L; lc++) {
        if (dtauc[lc] < 0.0) dtauc[lc] = 0.0;
        if (ssalb[lc] < 0.0 || ssalb[lc] > 1.0) status = SBDO_BAD_INPUT;
        if (plank && (temper[lc + 1] < 0.0 || (lc == 0 && temper[0] < 0.0)))
            status = SBDO_BAD_INPUT;
        for (int k = 0; k <= in->nmom; k++) {
            double pm = (k == 0) ? 1.0 : pmom[(size_t)lc * ldp + k];
            if (pm < -1.0 || pm > 1.0) status = SBDO_BAD_INPUT;
        }
    }
    if (in->usrtau) {
        for (int lu = 0; lu < NT; lu++) {
            w->utau[lu] = utau_in[lu];
            if (fabs(w->utau[lu] - w->tauc[L]) <= 1.e-4) w->utau[lu] = w->tauc[L];
            if (w->utau[lu] < 0.0 || w->utau[lu] > w->tauc[L]) status = SBDO_BAD_INPUT;
        }
    }
    if (usrang_eff)
        for (int iu = 0; iu < NU; iu++) {
            umu[iu] = umu_in[iu];
            if (umu[iu] < -1.0 || umu[iu] > 1.0 || umu[iu] == 0.0) status = SBDO_BAD_INPUT;
            if (iu > 0 && umu[iu] < umu[iu - 1]) status = SBDO_BAD_INPUT;
        }
    if (status) goto done;

    /* ---------------- SETDIS (disort.f:2454-2700) ---------------- */
    if (!in->usrtau)
        for (int lc = 0; lc <= L; lc++) w->utau[lc] = w->tauc[lc];
    {
        w->expbea[0] = 1.0; w->taucpr[0] = 0.0;
        double abstau = 0.0, yessct = 0.0;
        const double abscut = 10.;
        int ncut = L;
        for (int lc = 0; lc < L; lc++) {
            yessct += ssalb[lc];
            if (abstau < abscut) ncut = lc + 1;
            abstau += (1. - ssalb[lc]) * dtauc[lc];
            double f = pmom[(size_t)lc * ldp + N];
            w->oprim[lc] = ssalb[lc] * (1. - f) / (1. - f * ssalb[lc]);
            w->dtaucp[lc] = (1. - f * ssalb[lc]) * dtauc[lc];
            w->taucpr[lc + 1] = w->taucpr[lc] + w->dtaucp[lc];
            for (int k = 0; k <= N - 1; k++) {
                double pm = (k == 0) ? 1.0 : pmom[(size_t)lc * ldp + k];
                GLM(k, lc) = (2 * k + 1) * w->oprim[lc] * (pm - f) / (1. - f);
            }
            w->flyr[lc] = f;
            w->expbea[lc + 1] = 0.0;
            if (fbeam > 0.0) w->expbea[lc + 1] = exp(-w->taucpr[lc + 1] / umu0);
        }
        w->lyrcut = 0;
        w->lamber = in->lamber;
        if (abstau >= abscut && !plank && L > 1) w->lyrcut = 1;
        if (!w->lyrcut) ncut = L;
        w->ncut = ncut;
        for (int lu = 0; lu < NT; lu++) {
            int lc;
            for (lc = 1; lc <= L; lc++)
                if (w->utau[lu] >= w->tauc[lc - 1] && w->utau[lu] <= w->tauc[lc]) break;
            if (lc > L) lc = L;
            w->utaupr[lu] = w->taucpr[lc - 1] +
                (1. - ssalb[lc - 1] * w->flyr[lc - 1]) * (w->utau[lu] - w->tauc[lc - 1]);
            w->layru[lu] = lc;
        }
        sbdo_qgausn(n, w->cmu, w->cwt);
        for (int iq = 0; iq < n; iq++) { w->cmu[iq + n] = -w->cmu[iq]; w->cwt[iq + n] = w->cwt[iq]; }
        if (fbeam > 0.0) {
            for (int iq = 0; iq < n; iq++)
                if (fabs(umu0 - w->cmu[iq]) / umu0 < 1.e-4) { warn |= 1 << 1; status = SBDO_ANGLE_CLASH; }
            if (status) goto done;
        }
        if (!usrang_eff) {
            for (int iu = 0; iu < n; iu++) umu[iu] = -w->cmu[n - 1 - iu];
            for (int iu = n; iu < N; iu++) umu[iu] = w->cmu[iu - n];
        }
        if (onlyfl || fbeam == 0.0 || yessct == 0.0) corint = 0;
    }

    /* Planck sources (disort.f:556-571) */
    double bplank = 0.0, tplank = 0.0;
    if (plank) {
        tplank = in->temis * sbdo_plkavg(in->wvnmlo, in->wvnmhi, in->ttemp, &warn);
        bplank = sbdo_plkavg(in->wvnmlo, in->wvnmhi, in->btemp, &warn);
        for (int lev = 0; lev <= L; lev++)
            w->pkag[lev] = sbdo_plkavg(in->wvnmlo, in->wvnmhi, temper[lev], &warn);
    }

    /* azimuth loop (disort.f:577-827) */
    int kconv = 0, naz = N - 1;
    if (fbeam == 0.0 || fabs(1. - umu0) < 1.e-5 || onlyfl ||
        (NU == 1 && fabs(1. - umu[0]) < 1.e-5) ||
        (NU == 1 && fabs(1. + umu[0]) < 1.e-5) ||
        (NU == 2 && fabs(1. + umu[0]) < 1.e-5 && fabs(1. - umu[1]) < 1.e-5))
        naz = 0;
    if (!onlyfl && uu) memset(uu, 0, sizeof(double) * (size_t)in->nphi * NT * NU);

    for (int mazim = 0; mazim <= naz; mazim++) {
        double delm0 = (mazim == 0) ? 1.0 : 0.0;
        if (fbeam > 0.0) {
            double angcos = -umu0;
            lepoly(1, mazim, N + 1, N - 1, &angcos, w->ylm0);
        }
        if (!onlyfl && usrang_eff) lepoly(NU, mazim, N + 1, N - 1, umu, w->ylmu);
        lepoly(n, mazim, N + 1, N - 1, w->cmu, w->ylmc);
        {
            double sgn = -1.0;
            for (int l = mazim; l <= N - 1; l++) {
                sgn = -sgn;
                for (int iq = n; iq < N; iq++) YLMC(l, iq) = sgn * YLMC(l, iq - n);
            }
        }
        /* SURFAC (disort.f:3639): Lambertian :3746-3763, :3834-3849; BRDF :3765-3907 */
        if (!w->lyrcut) {
            memset(w->bdr, 0, sizeof(double) * (size_t)n * (n + 1));
            memset(w->bem, 0, sizeof(double) * n);
            if (!onlyfl && usrang_eff) {
                memset(w->emu, 0, sizeof(double) * NU);
                memset(w->rmu, 0, sizeof(double) * (size_t)NU * (n + 1));
            }
            if (!in->lamber) {
                surfac_brdf(n, NU, mazim, delm0, fbeam, umu0, pi, w->cmu, umu, !onlyfl && usrang_eff,
                            w->bdr, w->bem, w->rmu, w->emu);
            } else {
                if (mazim == 0)
                    for (int iq = 0; iq < n; iq++) {
                        w->bem[iq] = 1.0 - in->albedo;
                        for (int jq = 0; jq <= n; jq++) w->bdr[iq * (n + 1) + jq] = in->albedo;
                    }
                if (!onlyfl && usrang_eff)
                    for (int iu = 0; iu < NU; iu++)
                        if (umu[iu] > 0.0 && mazim == 0) {
                            for (int iq = 0; iq <= n; iq++) w->rmu[iu * (n + 1) + iq] = in->albedo;
                            w->emu[iu] = 1.0 - in->albedo;
                        }
            }
        }
        for (int lc = 0; lc < w->ncut; lc++) {
            status = soleig(w, mazim, lc);
            if (status) goto done;
            if (fbeam > 0.0) upbeam(w, mazim, lc, delm0, fbeam, umu0, pi, &warn);
            if (plank && mazim == 0) {
                w->xr1[lc] = 0.0;
                if (w->dtaucp[lc] > 0.0)
                    w->xr1[lc] = (w->pkag[lc + 1] - w->pkag[lc]) / w->dtaucp[lc];
                w->xr0[lc] = w->pkag[lc] - w->xr1[lc] * w->taucpr[lc];
                upisot(w, lc, &warn);
            }
            if (!onlyfl && usrang_eff) {
                terpev(w, mazim, lc);
                terpso(w, mazim, lc, delm0, fbeam, plank, pi);
            }
        }
        status = setmtx_solve0(w, mazim, delm0, fbeam, umu0, in->fisot, tplank,
                               bplank, pi, &warn);
        if (status) goto done;
        if (mazim == 0)
            fluxes(w, fbeam, umu0, pi, ssalb, rfldir, rfldn, flup, dfdt, uavg);
        if (onlyfl) {
            if (u0u) memcpy(u0u, w->u0c, sizeof(double) * (size_t)NT * N);
            break;
        }
        memset(w->uum, 0, sizeof(double) * (size_t)NT * NU);
        if (usrang_eff)
            usrint(w, mazim, delm0, fbeam, umu0, in->fisot, tplank, bplank, plank, pi, umu);
        else
            cmpint(w, mazim, fbeam, umu0, plank);
        if (mazim == 0) {
            for (int lu = 0; lu < NT; lu++)
                for (int iu = 0; iu < NU; iu++) {
                    double v = w->uum[(size_t)lu * NU + iu];
                    if (u0u) u0u[(size_t)lu * NU + iu] = v;
                    for (int j = 0; j < in->nphi; j++)
                        uu[((size_t)j * NT + lu) * NU + iu] = v;
                }
            if (naz > 0)
                for (int j = 0; j < in->nphi; j++)
                    w->phirad[j] = rpd * (phi[j] - in->phi0);
        } else {
            double azerr = 0.0;
            for (int j = 0; j < in->nphi; j++) {
                double cosphi = cos(mazim * w->phirad[j]);
                for (int lu = 0; lu < NT; lu++)
                    for (int iu = 0; iu < NU; iu++) {
                        double azterm = w->uum[(size_t)lu * NU + iu] * cosphi;
                        double *pu = &uu[((size_t)j * NT + lu) * NU + iu];
                        *pu += azterm;
                        double rr = ratio(fabs(azterm), fabs(*pu));
                        if (rr > azerr) azerr = rr;
                    }
            }
            if (azerr <= in->accur) kconv++;
            if (kconv >= 2) break;
        }
    }
    if (corint)
        intcor(w, dither, fbeam, ldp, in->nmom, in->nphi, pmom, ssalb, dtauc,
               umu, umu0, pi, rpd, uu);
done:
    if (warn_out) *warn_out = warn;
    return status;
}

/* ------------------------------------------------------------------ */
/* batched flux-only driver for the CPU baseline                       */
/* ------------------------------------------------------------------ */
#ifdef _OPENMP
#include <omp.h>
#endif
#ifdef __GLIBC__
#include <malloc.h>
#endif
int sbdo_disort_flux_batch(int nbins, int nlyr, int nstr, int nmom,
                           const double *dtauc, const double *ssalb,
                           const double *pmom, const double *fbeam,
                           const double *umu0, const double *albedo,
                           const int *plank, const double *wvnmlo,
                           const double *wvnmhi, const double *btemp,
                           const double *ttemp, const double *temis,
                           const double *fisot, const double *temper,
                           const int *col, double *rfldir, double *rfldn,
                           double *flup, double *dfdt, double *uavg,
                           int *status, int nthreads)
{
    int nbad = 0;
    const int NT = nlyr + 1;
    (void)nthreads;
#ifdef __GLIBC__
    /* keep the per-call work arrays on the heap: mmap/munmap of the band
       matrix on every call serialises the threads in the kernel */
    mallopt(M_MMAP_THRESHOLD, 256 << 20);
    mallopt(M_TRIM_THRESHOLD, 512 << 20);
#endif
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 4) num_threads(nthreads > 0 ? nthreads : 1) reduction(+ : nbad)
#endif
    for (int b = 0; b < nbins; b++) {
        sbdo_input in;
        memset(&in, 0, sizeof in);
        in.nlyr = nlyr; in.nstr = nstr; in.nmom = nmom;
        in.onlyfl = 1; in.lamber = 1; in.plank = plank ? plank[b] : 0;
        in.fbeam = fbeam[b]; in.umu0 = umu0[b]; in.albedo = albedo[b];
        in.fisot = fisot ? fisot[b] : 0.0;
        if (in.plank) {
            in.wvnmlo = wvnmlo[b]; in.wvnmhi = wvnmhi[b];
            in.btemp = btemp[b]; in.ttemp = ttemp[b]; in.temis = temis[b];
        }
        int warn = 0;
        int c = col ? col[b] : 0;
        int st = sbdo_disort(&in, dtauc + (size_t)b * nlyr, ssalb + (size_t)b * nlyr,
                             pmom + (size_t)b * nlyr * (nmom + 1),
                             temper ? temper + (size_t)c * (nlyr + 1) : NULL,
                             NULL, NULL, NULL,
                             rfldir + (size_t)b * NT, rfldn + (size_t)b * NT,
                             flup + (size_t)b * NT, dfdt + (size_t)b * NT,
                             uavg + (size_t)b * NT, NULL, NULL, &warn);
        if (status) status[b] = st;
        if (st) nbad++;
    }
    return nbad;
}
