// K2: optical-property producer.  One thread per wavelength runs what the
// reference does between `wllimits` and `CALL DISORT` (drt.f:432-533) and
// writes the DISORT inputs of the wavelength's 1 or 3 k-distribution bins
// straight into the batch arrays in HBM:
//   gasset  -> taugas x2 (vertical + slant path), kdistr, taucor   taugas.f:7392,2236,1802,7650
//   taucloud/cloudpar/getmom                                     taucloud.f:10,:6726 disutil.f:2104
//   rayleigh, raysig                                             spectra.f:179-247
//   solirr, salbedo (table lerps), wllimits                      spectra.f:1409,:53 drt.f:1657
//   normom, depthscl (+rolloff)                                  drt.f:1366 taugas.f:7512
// Default-REAL literals of the reference are written as float literals
// widened to double, so the numbers match a gfortran build of the reference.
#include <math.h>

#include "sbd_internal.h"
#include "sbd_optics.cuh"

namespace sbd {

#define F32(x) ((double)(x##f))
constexpr int kMaxZ = 66;

__device__ __forceinline__ double tg(const OpticsTables &T, int id, int i) { return T.base[T.off[id] + i]; }

// locate (drt.f:1181-1235), 1-based result
__device__ int locate_dev(const double *xx, int n, double x)
{
    if (x == xx[0]) return 1;
    if (x == xx[n - 1]) return n - 1;
    int jl = 1, ju = n;
    const bool asc = xx[n - 1] > xx[0];
    while (ju - jl > 1) {
        const int jm = (ju + jl) / 2;
        if (asc == (x > xx[jm - 1])) jl = jm; else ju = jm;
    }
    return jl;
}

__device__ double interp_dev(const double *wlt, const double *tab, int n, double wl)
{
    const int j = locate_dev(wlt, n, wl);
    double wt = (wl - wlt[j - 1]) / (wlt[j] - wlt[j - 1]);
    wt = fmax(0.0, fmin(1.0, wt));
    return tab[j - 1] * (1.0 - wt) + tab[j] * wt;
}

// sint (taugas.f:3854-3871): v1 = -20, dv = 10, npts = 2003
__device__ double sint_dev(const OpticsTables &T, int id, double v)
{
    const int i = (int)((v + 20.) / 10. + F32(1.00001));
    if (i >= 2003) return 0.0;
    double c = tg(T, id, i - 1);
    if (((int)v) % 10 > 0) c = (tg(T, id, i - 1) + tg(T, id, i)) / 2.;
    return c;
}

// cxdta (taugas.f:6416-6456), stateless search
__device__ double cxdta_dev(const OpticsTables &T, int imol, double v)
{
    const int iv = (int)v;
    int ic = 0;
    for (int b = 0;; b++) {
        const int lo = (int)tg(T, T_IWL0 + imol - 1, b), hi = (int)tg(T, T_IWH0 + imol - 1, b);
        if (lo == -999) break;
        if (iv >= lo && iv <= hi) return tg(T, T_CP0 + imol - 1, ic + (iv - lo) / 5);
        ic += (hi - lo) / 5 + 1;
    }
    return -20.0;
}

struct BandState {
    double cps[12], bms[12], bma[12], bmb[12], bmc[12];
    int ibnd[12];
};

// abcdta (taugas.f:6458-6733): band windows come as (imol, iw, lo, hi) quadruples
__device__ void abcdta_dev(const OpticsTables &T, int iv, BandState &s)
{
    for (int m = 1; m <= 11; m++) s.ibnd[m] = -1;
    const int nq = T.len[T_BANDS] / 4;
    for (int q = 0; q < nq; q++) {
        const int imol = (int)tg(T, T_BANDS, 4 * q), iw = (int)tg(T, T_BANDS, 4 * q + 1);
        const int lo = (int)tg(T, T_BANDS, 4 * q + 2), hi = (int)tg(T, T_BANDS, 4 * q + 3);
        if (iv >= lo && iv <= hi) s.ibnd[imol] = iw;     // later windows override earlier ones
    }
    const int base[12] = { 0, 16, 35, 30, 46, 43, 45, 49, 53, 55, 54, 51 };
    for (int m = 1; m <= 11; m++) {
        const int iw = s.ibnd[m];
        if (iw > 0) {
            const int ib = iw - base[m] - 1;
            s.bms[m] = tg(T, T_BMS0 + m - 1, ib);
            if (m == 7 && iv >= 49600 && iv <= 52710) s.bms[m] = F32(.4704);
            s.bma[m] = tg(T, T_BMA0 + m - 1, ib);
            s.bmb[m] = tg(T, T_BMB0 + m - 1, ib);
            s.bmc[m] = tg(T, T_BMC0 + m - 1, ib);
        }
    }
}

__device__ double raysig_dev(double v) { return v * v * v * v / (F32(9.38076e+18) + F32(-1.08426e+09) * v * v); }

// taugas (taugas.f:2236-2534): continuum and band-model depth increments, top-down
__device__ void taugas_dev(const OpticsArgs &a, double wl, double amu0, double *dtauc, double *dtaul,
                           BandState &s)
{
    const OpticsTables &T = a.tab;
    const int nz = a.p.nz;
    const int iv = 5 * ((int)(10000.0 / wl) / 5);
    const double v = 10000. / wl;
    double s0 = sint_dev(T, T_SLF296, v), s1 = sint_dev(T, T_SLF260, v);
    const double fh2o = sint_dev(T, T_FRN296, v);
    const double t0 = 296., t1 = 260.;
    if (s0 > 0.) {
        const double alpha2 = 200. * 200.;
        const double xh2o = 1. - F32(0.2333) * (alpha2 / ((v - 1050.) * (v - 1050.) + alpha2));
        s0 *= xh2o; s1 *= xh2o;
    }
    double radfn0, radfn1;
    if ((v / F32(0.6952)) / t1 <= 87.) {
        double xd = exp(-v / (t0 * F32(0.6952)));
        radfn0 = v * (1. - xd) / (1. + xd);
        xd = exp(-v / (t1 * F32(0.6952)));
        radfn1 = v * (1. - xd) / (1. + xd);
    } else { radfn0 = v; radfn1 = v; }
    const double wfac = F32(1.e-20);
    const double ya = exp(-log(F32(1.025) * F32(3.159e-8)) + F32(2.75e-4) * v);
    const double yb = exp(-log(F32(8.97e-6)) + F32(1.300e-3) * v);
    const double fdg = 1. / (ya + yb);
    // c4dta, hno3, hertda, o2cont, o4cont
    double abn2 = 0.0;
    if (v >= 2080. && v <= 2740.) abn2 = tg(T, T_C4, ((int)v - 2080) / 5);
    double abno3 = 0.0;
    if (v >= 850. && v <= 920.) abno3 = tg(T, T_H1, (int)((v - 845.) / 5.) - 1);
    else if (v >= 1275. && v <= 1350.) abno3 = tg(T, T_H2, (int)((v - 1270.) / 5.) - 1);
    else if (v >= 1675. && v <= 1735.) abno3 = tg(T, T_H3, (int)((v - 1670.) / 5.) - 1);
    double abo2 = 0.0;
    if (v > 36000.) {
        double corr = 0.0;
        if (v <= 40000.) corr = ((40000. - v) / 4000.) * F32(7.917e-27);
        const double rlosch = F32(2.6868e24) * F32(1.0e-5), yr = v / 48811.0, ly = log(yr);
        abo2 = (F32(6.884e-24) * yr * exp(F32(-69.738) * ly * ly) - corr) * rlosch;
    }
    double sigo20 = 0.0, sigo2a = 0.0, sigo2b = 0.0;
    if (v >= 1395 && v <= 1760) {
        const int i = (int)((v - 1395.0) / 5.0 + F32(1.00001));
        double c = 0., aa = 0., b = 0.;
        if (i >= 1 && i <= 74) { c = tg(T, T_O2S0, i - 1); aa = tg(T, T_O2A, i - 1); b = tg(T, T_O2B, i - 1); }
        sigo20 = c / F32(0.20946); sigo2a = aa; sigo2b = aa * aa / 2. + b;
    }
    double sigo4 = 0.0;
    {
        const double wnm = 1000. * wl;
        int inm = (int)wnm;
        const double f = wnm - inm;
        inm = inm - 335 + 1;
        if (inm >= 1 && inm <= 1015) {
            const double fraco2 = F32(.209), fracn2 = F32(.781), effn2 = F32(.2);
            double factor = fraco2 * fraco2;
            if (wl > F32(1.2)) factor = fraco2 * (fraco2 + effn2 * fracn2);
            sigo4 = a.p.xo4 * factor * (tg(T, T_O4SIG, inm - 1) * (1. - f) + tg(T, T_O4SIG, inm) * f);
        }
    }
    double doz1 = 0., doz2 = 0., doz3 = 0.;
    if (v > 40800) {            // o3uv
        const int i0 = (int)((v - 40800.) / 100. + F32(1.00001));
        double c = 0.0;
        if (i0 >= 1 && i0 <= 133) {
            int i = i0;
            const double vr = i * 100. + 40800.;
            if (vr <= v + F32(.1) && vr >= v - F32(.1)) c = tg(T, T_O3UV, i - 1);
            else {
                if (i == 133) i = 132;
                const double am = (tg(T, T_O3UV, i) - tg(T, T_O3UV, i - 1)) / 100.;
                c = am * v + (tg(T, T_O3UV, i - 1) - am * vr);
            }
        }
        doz1 = F32(.269) * c;
    } else if (v > 24370) {     // o3hht (table starts at 27370: zeros in between, as the reference)
        const int i = (int)((v - 27370.) / 5. + F32(1.00001));
        if (i >= 1 && i <= 2687) {
            const double c0 = tg(T, T_O3S0, i - 1);
            doz1 = F32(.269) * c0; doz2 = c0 * tg(T, T_O3S1, i - 1); doz3 = c0 * tg(T, T_O3S2, i - 1);
        }
    } else if (v >= 13000. && v <= 24200) {   // c8dta
        const int ivv = (int)v;
        if (!(ivv > 24200 && ivv < 27500)) {
            double xi = (v - 13000.0) / 200.0 + 1.;
            if (ivv >= 27500) xi = (v - 27500.0) / 500. + 57.;
            const int nn = (int)(xi + F32(1.001));
            const double xd = xi - (double)nn;
            doz1 = tg(T, T_C8, nn - 1) + xd * (tg(T, T_C8, nn - 1) - tg(T, T_C8, nn - 2));
        }
    }
    for (int m = 1; m <= 11; m++) s.cps[m] = cxdta_dev(T, m, v);
    abcdta_dev(T, iv, s);
    if (v > 49600) {            // schrun
        const int i = (int)((v - 49600.) / 5. + F32(1.0001));
        s.cps[7] = (i >= 1 && i <= 423) ? tg(T, T_SHN, i - 1) : -20.;
    }

    // species whose slant-weighted amounts are needed: continua + active bands
    const int cont[13] = { 1, 2, 3, 4, 5, 8, 9, 10, 11, 58, 59, 60, 63 };
    double wc[13], wb[12];
    for (int q = 0; q < 13; q++) wc[q] = 0.0;
    for (int m = 1; m <= 11; m++) wb[m] = 0.0;
    const double re = F32(6371.2);
    const int ld = nz + 1;
    double taucp = 0.0, taulp = 0.0;
    for (int im = 1; im <= nz; im++) {
        const int i = nz - im;      // 0-based level index, from the top
        const double zb = (i == nz - 1) ? a.z[i] : 0.5 * (a.z[i] + a.z[i + 1]);
        const double rr = re / (re + zb);
        const double ramu = 1.0 / sqrt(1. - (1. - amu0 * amu0) * rr * rr);
        for (int q = 0; q < 13; q++) {
            const int k = cont[q];
            wc[q] += (a.uu[k * ld + i] - a.uu[k * ld + i + 1]) * ramu;
        }
        for (int m = 1; m <= 11; m++) {
            const int ib = s.ibnd[m];
            if (ib > 0) wb[m] += (a.uu[ib * ld + i] - a.uu[ib * ld + i + 1]) * ramu;
        }
        // wc: 0:w1 1:w2 2:w3 3:w4 4:w5 5:w8 6:w9 7:w10 8:w11 9:w58 10:w59 11:w60 12:w63
        const double tcunif = sigo4 * wc[2] + abn2 * wc[3] +
                              sigo20 * (wc[12] + sigo2a * (wc[0] - 220 * wc[12]) + sigo2b * wc[1]) +
                              abo2 * wc[9];
        const double tch2o = s0 * radfn0 * (wfac * wc[4]) +
                             ((s1 * radfn1) - (s0 * radfn0)) * (wfac * wc[6]) +
                             (fh2o + fdg) * radfn0 * (wfac * wc[7]);
        const double tco3 = doz1 * wc[5] + doz2 * wc[10] + doz3 * wc[11];
        const double tctrc = abno3 * wc[8];
        const double tauc = tcunif + tch2o + tco3 + tctrc;
        double taul = 0.0;
        for (int m = 1; m <= 11; m++) {
            if (s.ibnd[m] > 0 && s.cps[m] > -20. && wb[m] > 1.e-20) {
                double awl = s.bms[m] * (s.cps[m] + log10(wb[m]));
                awl = fmin(awl, 20.);
                taul += exp10(awl);
            }
        }
        dtauc[im - 1] = tauc - taucp;
        dtaul[im - 1] = taul - taulp;
        taucp = tauc; taulp = taul;
    }
}

// taucor (taugas.f:7650-7692); returns false when the Newton iteration fails
__device__ bool taucor_dev(const double *gwk, const double *tau, double amu, double utau, double &cf)
{
    cf = 1.;
    if (utau > 12.0) return true;
    for (int it = 0; it < 20; it++) {
        double ff = 0.0, fs = 0.0;
        for (int k = 0; k < 3; k++) {
            const double e = gwk[k] * exp(-cf * tau[k] / amu);
            ff += e; fs += e * tau[k];
        }
        const double f = log(ff) + utau;
        if (fabs(f) < F32(0.000001)) return true;
        const double fp = -fs / (ff * amu);
        cf += -f / fp;
    }
    return false;
}

__device__ double rolloff_dev(double wl, double tsc)
{
    double ramp = (F32(4.1) - wl) / (F32(4.1) - F32(3.9));
    ramp = fmax(fmin(1.0, ramp), 0.0);
    return ramp * exp(1. - fmax(tsc, 1.0));
}

// cloudpar (taucloud.f:6726-6762): tables are stored [re][wl]
__device__ void cloudpar_dev(const OpticsTables &T, double wl, double re, double &qc, double &wc, double &gc)
{
    const double wmin = log(F32(0.29)), wmax = log(F32(333.33));
    const double wstep = (wmax - wmin) / (400 - 1);
    const double eps = F32(.000001);
    double fw = 1 + (log(wl) - wmin) / wstep;
    fw = fmin(fmax(fw, 1.0), 400.0 - eps);
    const int iw = (int)fw;
    fw -= iw;
    double fr = 1. + (log(fabs(re)) / log(2.) - 1.) * 2;
    fr = fmin(fmax(fr, 1.0), 13.0 - eps);
    const int ir = (int)fr;
    fr -= ir;
    const int ice = re < 0. ? 3 : 0;
    double out[3];
    for (int q = 0; q < 3; q++) {
        const int id = T_MIE_QQ + ice + q;
        const double a00 = tg(T, id, (ir - 1) * 400 + iw - 1), a10 = tg(T, id, (ir - 1) * 400 + iw);
        const double a01 = tg(T, id, ir * 400 + iw - 1), a11 = tg(T, id, ir * 400 + iw);
        out[q] = a00 * (1. - fw) * (1. - fr) + a10 * fw * (1. - fr) + a01 * (1. - fw) * fr + a11 * fw * fr;
    }
    qc = out[0]; wc = out[1]; gc = out[2];
}

// aerbwi / aestrat (tauaero.f:177-250, :253-403): log-log interpolation of extinction and
// absorption, linear in the asymmetry factor; power-law continuation outside the table.
// `wl_hi_ref` is the wavelength the upper continuation is scaled from (the last point for
// the boundary-layer model, awl(1) for the stratospheric models -- as in the reference).
__device__ void aer_interp_dev(const double *wlb, const double *ext, const double *ab, const double *as,
                               int n, double wl, double abaer, double wl_hi_ref, bool linear_abs_guard,
                               double &extinc, double &wa, double &ga)
{
    wa = 0.0;
    if (wl <= wlb[0]) {
        extinc = ext[0] * pow(wlb[0] / wl, abaer);
        wa = 1. - ab[0] / ext[0];
        ga = as[0];
    } else if (wl >= wlb[n - 1]) {
        extinc = ext[n - 1] * pow(wl_hi_ref / wl, abaer);
        wa = 1. - ab[n - 1] / ext[n - 1];
        ga = as[n - 1];
    } else {
        const int l = locate_dev(wlb, n, wl);
        const double wt = log(wl / wlb[l - 1]) / log(wlb[l] / wlb[l - 1]);
        extinc = ext[l - 1] * pow(ext[l] / ext[l - 1], wt);
        double absorp;
        if (!linear_abs_guard || (ab[l - 1] > 0. && ab[l] > 0.)) absorp = ab[l - 1] * pow(ab[l] / ab[l - 1], wt);
        else absorp = ab[l - 1] * (1. - wt) + ab[l] * wt;
        if (extinc > 0.) wa = fmax(0.0, fmin(1. - absorp / extinc, 1.0));
        ga = (1. - wt) * as[l - 1] + wt * as[l];
    }
}

__global__ void __launch_bounds__(64)
optics_kernel(const OpticsArgs a)
{
    const int il = blockIdx.x * blockDim.x + threadIdx.x;
    const sbd_optics_params &P = a.p;
    if (il >= P.nwl) return;
    const int nz = P.nz, nmom = (P.nstr + 2 < 40) ? P.nstr + 2 : 40, ldp = nmom + 1;

    // ---- wllimits (drt.f:1657-1740)
    double wl, ww1, ww2;
    {
        const double wi = (double)il, nm1 = (double)(P.nwl - 1);
        if (P.wlinc > 1) {
            auto f = [&](double x) { const double xx = x / nm1; // no FMA contraction: int(1e4/wl) below must land on the same side of a table
                // boundary as the host front end when the grid hits exact wavenumbers
                return __dmul_rn(P.wl1, P.wl2) / __dadd_rn(__dmul_rn(1. - xx, P.wl2), __dmul_rn(xx, P.wl1)); };
            wl = f(wi); ww1 = f(wi - .5); ww2 = f(wi + .5);
        } else if (P.wlinc < 0.) {
            const double wr = P.wl2 / P.wl1;
            auto f = [&](double x) { return P.wl1 * pow(wr, x / nm1); };
            wl = f(wi); ww1 = f(wi - .5); ww2 = f(wi + .5);
        } else {
            wl = __dadd_rn(P.wl1, __dmul_rn(wi, P.wlinc));
            ww1 = __dadd_rn(wl, -__dmul_rn(.5, P.wlinc)); ww2 = __dadd_rn(wl, __dmul_rn(.5, P.wlinc));
        }
        if (il == 0 && il != P.nwl - 1) ww1 = wl;
        if (il == P.nwl - 1 && il != 0) ww2 = wl;
        if (ww1 == wl && ww2 == wl) { ww1 = wl - F32(.0005); ww2 = wl + F32(.0005); }
    }
    const double wvnmlo = 10000. / ww2, wvnmhi = 10000. / ww1;
    const double dwl = 10000. / wvnmlo - 10000. / wvnmhi;
    // sza >= 90: the reference overwrites amu0 with 1 inside the first pass of the loop
    // (drt.f:456-459), so gasset sees cos(sza) for the first wavelength only
    double amu0 = (P.night && il > 0) ? 1.0 : P.amu0;

    // ---- gasset (taugas.f:7392-7510)
    double dtcv[kMaxZ], dtlv[kMaxZ], dtls[kMaxZ], dtk[kMaxZ][3], dk2[kMaxZ][3];
    BandState st, st2;
    taugas_dev(a, wl, 1.0, dtcv, dtlv, st);
    if (amu0 > 0.) {
        double dtcs[kMaxZ];
        taugas_dev(a, wl, amu0, dtcs, dtls, st2);
    } else {
        for (int j = 0; j < nz; j++) dtls[j] = dtlv[j];
    }
    int nk = 1;
    double gwk[3] = { 1., 0., 0. };
    double sumlv = 0.0;
    for (int j = 0; j < nz; j++) sumlv += dtlv[j];
    bool usek = false;
    if (!(P.kdist == 0 || sumlv < F32(.01))) {
        // kdistr (taugas.f:1802-1920)
        const int ld = nz + 1;
        double tk[3] = { 0, 0, 0 }, gw[3] = { 0, 0, 0 };
        for (int nn = 0; nn < nz; nn++) {
            const int i = nz - 1 - nn;
            double dt3[3] = { 0, 0, 0 }, tw3[3] = { 0, 0, 0 };
            for (int m = 1; m <= 11; m++) {
                const int ib = st.ibnd[m];
                if (ib < 0) continue;
                const double duu = a.uu[ib * ld + i] - a.uu[ib * ld + i + 1];
                const double cp1 = exp10(st.cps[m]);
                for (int k = 0; k < 3; k++) {
                    const double gk = tg(a.tab, T_KFAC, k) * st.bmc[m];
                    const double dp = k == 0 ? st.bma[m] : (k == 1 ? st.bmb[m] : 1. - st.bma[m] - st.bmb[m]);
                    const double wpth = duu * gk;
                    dt3[k] += wpth * cp1;
                    tw3[k] += wpth * cp1 * dp;
                }
            }
            double wk3[3], sm = 0.0;
            for (int k = 0; k < 3; k++) { wk3[k] = dt3[k] != 0 ? tw3[k] / dt3[k] : 1. / 3.; sm += wk3[k]; }
            for (int k = 0; k < 3; k++) {
                wk3[k] /= sm;
                dtk[nn][k] = dt3[k];
                tk[k] += dt3[k];
                gw[k] += dtlv[nn] * wk3[k];
            }
        }
        if (fmax(tk[0], fmax(tk[1], tk[2])) >= F32(0.01)) {
            nk = 3; usek = true;
            const double wn = gw[0] + gw[1] + gw[2];
            if (wn == 0) { gwk[0] = 1.; gwk[1] = 0.; gwk[2] = 0.; }
            else for (int k = 0; k < 3; k++) gwk[k] = gw[k] / wn;
        }
    }
    // dtauk(:,1:3) -> dtk, dtauk(:,4:6) -> dk2
    if (!usek) {
        for (int j = 0; j < nz; j++) { dtk[j][0] = dtlv[j]; dk2[j][0] = amu0 * dtls[j]; }
    } else {
        for (int j = 0; j < nz; j++)
            for (int k = 0; k < 3; k++) dk2[j][k] = dtk[j][k];
        if (P.kdist >= 2 && amu0 > 0.) {
            double tauls = 0., tglc[3] = { 0, 0, 0 };
            for (int j = 0; j < nz; j++) {
                tauls += dtls[j];
                for (int k = 0; k < 3; k++) tglc[k] += dtk[j][k];
                double cf;
                taucor_dev(gwk, tglc, amu0, tauls, cf);
                for (int k = 0; k < 3; k++) { dk2[j][k] = tglc[k] * (cf - 1.0) + dtk[j][k]; tglc[k] *= cf; }
            }
        }
    }

    if (amu0 <= 0.)      // taugas.f:7491 -- applies to the first slant column whatever nk is
        for (int j = 0; j < nz; j++) dk2[j][0] = dtlv[j];

    // ---- solar flux, surface albedo, Planck switch (drt.f:448-486)
    double flxin = (P.nf == 0) ? dwl : interp_dev(a.wlsun, a.sun, P.nsun, wl) * dwl * P.solfac;
    if (P.night) { flxin = 0.; amu0 = 1.; }
    const int plank = (P.nothrm < 0) ? (wl > 2.) : (P.nothrm == 0);
    double rsfc = interp_dev(a.wlalb, a.alb, P.nalb, wl);
    rsfc = fmax(0.0, fmin(rsfc, 1.0));

    // ---- clouds (taucloud.f:10-140); slot 0 of pmom is the accumulation buffer
    const size_t s0i = (size_t)3 * il;
    double *pm0 = a.pmom + s0i * nz * ldp;
    double taucld[kMaxZ], wcld[kMaxZ];
    int icnt[kMaxZ];
    for (int j = 0; j < nz; j++) {
        taucld[j] = 0.; wcld[j] = 0.; icnt[j] = 0;
        for (int k = 0; k <= nmom; k++) pm0[(size_t)j * ldp + k] = 0.0;
    }
    for (int c = 0; c < P.ncloud; c++) {
        const sbd_cloud_entry ce = a.clouds[c];
        const int j = ce.layer - 1;
        double qc, wc, gc;
        cloudpar_dev(a.tab, wl, ce.reff, qc, wc, gc);
        double gp = 1.0;
        for (int k = 1; k <= nmom; k++) {          // getmom, HG (iphas = 3) or Rayleigh (2)
            gp *= gc;
            const double pmk = (P.imomc == 2) ? (k == 2 ? F32(0.1) : 0.0) : gp;
            pm0[(size_t)j * ldp + k] += pmk;
        }
        wcld[j] += wc;
        icnt[j] += 1;
        if (ce.use_tau) taucld[j] += ce.tcld * qc / ce.q550;
        else if (ce.lwpth != 0.) {
            if (ce.reff < 0.) taucld[j] += F32(-.75) * qc * ce.lwpth / ce.reff / F32(.917);
            else taucld[j] += F32(.75) * qc * ce.lwpth / ce.reff;
        }
    }
    // ---- aerosols (tauaero, tauaero.f:1223-1331): per-wavelength scattering parameters
    double dtaua[kMaxZ], waer[kMaxZ];
    for (int j = 0; j < nz; j++) { dtaua[j] = 0.; waer[j] = 0.; }
    double bl_ext = 0., bl_wa = 0., bl_ga = 0.;
    const double *dtsv = nullptr, *awl = nullptr, *strat = nullptr;
    double st_dt[SBD_NAERZ], st_wa[SBD_NAERZ], st_ga[SBD_NAERZ];
    int st_layer[SBD_NAERZ];
    const int nbl = a.aero ? a.aer.nwlbaer : 0, nst = a.aero ? a.aer.nstrat : 0;
    if (a.aero) {
        const double *wlb = a.aero, *ext = wlb + nbl, *ab = ext + nbl, *as = ab + nbl;
        dtsv = as + nbl; awl = dtsv + nz; strat = awl + SBD_NAERW;
        if (nbl > 0) {
            aer_interp_dev(wlb, ext, ab, as, nbl, wl, a.aer.abaer, wlb[nbl - 1], true, bl_ext, bl_wa, bl_ga);
            if (a.aer.nosct == 1) bl_ext *= 1. - bl_wa;
            if (a.aer.nosct == 3) bl_ext *= 1. - bl_wa * bl_ga;
            if (a.aer.nosct != 0) { bl_wa = 0.; bl_ga = 0.; }
            for (int j = 0; j < nz; j++) { dtaua[j] = bl_ext * dtsv[j]; waer[j] = bl_wa; }
        }
        const int per = (int)(sizeof(sbd_strat_entry) / 8);
        for (int e = 0; e < nst; e++) {
            const double *se = strat + (size_t)e * per;
            double qa;
            aer_interp_dev(awl, se + 2, se + 2 + SBD_NAERW, se + 2 + 2 * SBD_NAERW, SBD_NAERW, wl, a.aer.abaer,
                           awl[0], false, qa, st_wa[e], st_ga[e]);
            st_layer[e] = (int)se[0] - 1;
            st_dt[e] = se[1] * qa;
        }
    }
    // ---- rayleigh (spectra.f:206-247), normom (drt.f:1366-1397)
    double dtaur[kMaxZ];
    {
        const double sig = raysig_dev(10000. / wl);
        const double pz = F32(1013.25), tz = F32(273.15);
        dtaur[0] = sig * (a.p_[nz - 1] / pz) / (a.t[nz - 1] / tz) * 5.;
        for (int i = 2; i <= nz; i++) {
            const int im = nz - i + 1;
            const double rhom = (a.p_[im - 1] / pz) / (a.t[im - 1] / tz);
            const double rhop = (a.p_[im] / pz) / (a.t[im] / tz);
            const double dz = a.z[im] - a.z[im - 1];
            dtaur[i - 1] = (rhom == rhop) ? .5 * sig * dz * (rhom + rhop)
                                          : sig * dz * (rhop - rhom) / log(rhop / rhom);
        }
        if (P.xrsc != 1.0) for (int j = 0; j < nz; j++) dtaur[j] *= P.xrsc;
    }
    for (int j = 0; j < nz; j++) {
        if (icnt[j]) {
            wcld[j] /= icnt[j];
            for (int k = 1; k <= nmom; k++)
                pm0[(size_t)j * ldp + k] = taucld[j] * wcld[j] * pm0[(size_t)j * ldp + k] / icnt[j];
        }
        if (nbl > 0) {                               // boundary-layer aerosol, getmom(imoma)
            const double dab = dtaua[j];
            double gp = 1.0;
            for (int k = 1; k <= nmom; k++) {
                gp *= bl_ga;
                const double pmk = (a.aer.imoma == 2) ? (k == 2 ? F32(0.1) : 0.0) : gp;
                pm0[(size_t)j * ldp + k] += pmk * dab * bl_wa;
            }
        }
        for (int e = 0; e < nst; e++) {              // stratospheric layers, Henyey-Greenstein
            if (st_layer[e] != j) continue;
            double gp = 1.0;
            for (int k = 1; k <= nmom; k++) { gp *= st_ga[e]; pm0[(size_t)j * ldp + k] += gp * st_dt[e] * st_wa[e]; }
            waer[j] = (waer[j] * dtaua[j] + st_wa[e] * st_dt[e]) / (dtaua[j] + st_dt[e]);
            dtaua[j] += st_dt[e];
        }
        pm0[(size_t)j * ldp + 2] += F32(.1) * dtaur[j];
        const double dtsct = taucld[j] * wcld[j] + dtaua[j] * waer[j] + dtaur[j];
        if (dtsct != 0.)
            for (int k = 0; k <= nmom; k++) pm0[(size_t)j * ldp + k] /= dtsct;
        pm0[(size_t)j * ldp] = 1.;
    }

    // ---- depthscl per k term (taugas.f:7512-7621) and the per-bin scalars
    for (int kd = 0; kd < nk; kd++) {
        const size_t slot = s0i + kd;
        double *od = a.dtauc + slot * nz, *os = a.ssalb + slot * nz;
        double wt = gwk[kd];
        double tsc = 0., tglv = 0., tgls = 0.;
        if (P.kdist == 0 || nk == 1) wt = 1.;
        for (int i = 0; i < nz; i++) {
            double dtaug;
            tsc += dtaur[i] + taucld[i] + dtaua[i];
            if (P.kdist == 0 || nk == 1) {
                tglv += dtk[i][0]; tgls += dk2[i][0];
                double afac = 1.;
                if (tglv > F32(.001)) afac = tgls / tglv;
                const double ramp = rolloff_dev(wl, tsc);
                afac = afac * ramp + 1. - ramp;
                dtaug = dtcv[i] + dtk[i][0] * afac;
            } else if (P.kdist == 1) dtaug = dtcv[i] + dtk[i][kd];
            else if (P.kdist == 2) dtaug = dtcv[i] + dk2[i][kd];
            else {
                const double ramp = rolloff_dev(wl, tsc);
                dtaug = dtcv[i] + dtk[i][kd] * (1. - ramp) + dk2[i][kd] * ramp;
            }
            const double dtau = dtaug + taucld[i] + dtaua[i] + dtaur[i];
            od[i] = dtau;
            os[i] = (dtau > 2.2250738585072014e-308) ? (taucld[i] * wcld[i] + dtaua[i] * waer[i] + dtaur[i]) / dtau : 0.0;
        }
        if (kd > 0) {
            double *pk = a.pmom + slot * nz * ldp;
            for (int e = 0; e < nz * ldp; e++) pk[e] = pm0[e];
        }
        sbd_bin b;
        b.fbeam = flxin; b.umu0 = amu0; b.phi0 = P.phi0; b.fisot = P.fisot; b.albedo = rsfc;
        b.btemp = P.btemp; b.ttemp = P.ttemp; b.temis = P.temis; b.wvnmlo = wvnmlo; b.wvnmhi = wvnmhi;
        b.accur = 0.0; b.plank = plank; b.col = 0;
        a.bins[slot] = b;
        a.wt[slot] = wt;
    }
    a.nk[il] = nk;
    a.wl[il] = wl;
    a.dwl[il] = dwl;
}

// bin -> slot map in loop order (wavelength-major, k-terms together)
__global__ void binmap_kernel(const int32_t *nk, int nwl, int32_t *binmap, int32_t *nbins)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    int b = 0;
    for (int il = 0; il < nwl; il++)
        for (int kd = 0; kd < nk[il]; kd++) binmap[b++] = 3 * il + kd;
    *nbins = b;
}

cudaError_t launch_optics(const OpticsArgs &a, cudaStream_t st)
{
    const int threads = 64, blocks = (a.p.nwl + threads - 1) / threads;
    optics_kernel<<<blocks, threads, 0, st>>>(a);
    return cudaGetLastError();
}

cudaError_t launch_binmap(const int32_t *nk, int nwl, int32_t *binmap, int32_t *nbins, cudaStream_t st)
{
    binmap_kernel<<<1, 32, 0, st>>>(nk, nwl, binmap, nbins);
    return cudaGetLastError();
}

}  // namespace sbd
