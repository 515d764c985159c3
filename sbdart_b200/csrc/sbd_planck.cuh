// Band-integrated Planck function shared by the kernels.
#pragma once
#include <math.h>

#include "sbd_internal.h"

namespace sbd {

// 1/m^4 for the exponential series of PLKAVG (disort.f:5628-5633), m = 1..7
static __constant__ double kRm4[8] = { 0.0, 1.0, 1.0 / 16.0, 1.0 / 81.0, 1.0 / 256.0,
                                       1.0 / 625.0, 1.0 / 1296.0, 1.0 / 2401.0 };

// ---- PLKAVG (disort.f:5410-5671), same three regimes --------------------
static __device__ __forceinline__ double plkf(double x) { return x * x * x / (exp(x) - 1.0); }

static __device__ __noinline__ double plkavg_dev(double wnumlo, double wnumhi, double t)
{
    const double a1 = 1. / 3., a2 = -1. / 8., a3 = 1. / 60., a4 = -1. / 5040.,
                 a5 = 1. / 272160., a6 = -1. / 13305600.;
    const double c2 = (double)1.438786f, sigma = (double)5.67032E-8f;
    const double vcp[7] = { 10.25, (double)5.7f, (double)3.9f, (double)2.9f,
                            (double)2.3f, (double)1.9f, 0.0 };
    const double pi = kPiRef;
    const double vmax = 709.782712893384;   // log(DBL_MAX)
    const double epsil = 2.220446049250313e-16;
    const double sigdpi = sigma / pi;
    const double conc = 15. / (pi * pi * pi * pi);
    if (t < 1.e-4) return 0.0;
    double v[2] = { c2 * wnumlo / t, c2 * wnumhi / t };
    if (v[0] > epsil && v[1] < vmax && (wnumhi - wnumlo) / wnumhi < 1.e-2) {
        double hh = v[1] - v[0], oldval = 0.0, val = 0.0;
        double val0 = plkf(v[0]) + plkf(v[1]);
        for (int n = 1; n <= 10; n++) {
            double del = hh / (2 * n);
            val = val0;
            for (int k = 1; k <= 2 * n - 1; k++)
                val += 2 * (1 + (k & 1)) * plkf(v[0] + k * del);
            val = del / 3. * val;
            if (fabs((val - oldval) / val) <= 1.e-6) break;
            oldval = val;
        }
        return sigdpi * t * t * t * t * conc * val;
    }
    double d[2] = { 0, 0 }, p[2] = { 0, 0 };
    int smallv = 0;
    for (int i = 0; i < 2; i++) {
        if (v[i] < 1.5) {
            smallv++;
            double vsq = v[i] * v[i];
            p[i] = conc * vsq * v[i] *
                   (a1 + v[i] * (a2 + v[i] * (a3 + vsq * (a4 + vsq * (a5 + vsq * a6)))));
        } else {
            int mmax = 0;
            do { mmax++; } while (v[i] < vcp[mmax - 1]);
            double ex = exp(-v[i]), exm = 1.0, s = 0.0;
            for (int m = 1; m <= mmax; m++) {
                double mv = m * v[i];
                exm = ex * exm;
                s += exm * (6. + mv * (6. + mv * (3. + mv))) * kRm4[m];   // m <= 7
            }
            d[i] = conc * s;
        }
    }
    double r;
    if (smallv == 2) r = p[1] - p[0];
    else if (smallv == 1) r = 1. - p[0] - d[1];
    else r = d[0] - d[1];
    return sigdpi * t * t * t * t * r;
}


}  // namespace sbd
