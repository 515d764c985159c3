// The adding form of the boundary-value problem (see sbd_adding.cu) as building blocks for the
// radiance kernel (sbd_fast.cu, NSTR 4/8/16): layer reflection / transmission operators from the
// eigen-solution, the bottom-up sweep, the top-down sweep that leaves the intensities at every
// interface, and the recovery of the layer solutions (the coefficients LL of disort.f) from them.
#pragma once
#include <cuda_runtime.h>

#include "sbd_devutil.cuh"
#include "sbd_internal.h"

namespace sbd {

template <int n>
struct AddOps {
    static constexpr int arec = 2 * n * n + 2 * n;      // R, T, s_up, s_dn  ->  Y, y, Rb, sb
    static constexpr int r_R = 0, r_T = n * n, r_su = 2 * n * n, r_sd = 2 * n * n + n;
    static constexpr int o_Y = 0, o_y = n * n, o_Rb = n * n + n, o_sb = 2 * n * n + n;
    static constexpr int CG = n >= 10 ? 2 : (n >= 4 ? 4 : n), CW = n / CG;
    static constexpr int p2 = 3 * n * n + 3 * n + 2 * arec;     // shared memory of the sweeps
};

// R, T, s_up, s_dn of one layer (lane g of a group of n lanes = row g / mode g).
//   P[i]: column g of P = L V; kk: eigenvalue; zus / zds: scaled beam particular solution D Z at
//   the upward / downward direction g (without the exponential); qs: scaled thermal vector D q;
//   b1 = dB / dtau'.  mP, mA, mB: three [n][LD] matrices of shared memory (free on entry), sv: 4n
//   doubles.  Writes the record when `active`.  Returns non-zero when an inversion broke down.
template <int n, int LD>
__device__ __forceinline__ int layer_operators(const double (&P)[n], double kk, double dtaucp,
                                               double zus, double zds, double qs, double b1,
                                               double pk_top, double pk_bot, double et, double eb,
                                               bool beam, bool therm, const double *cmu, const double *csq,
                                               double *mP, double *mA, double *mB, double *sv,
                                               double *rec, bool active, int g, bool gact = true)
{
    using AO = AddOps<n>;
    int bad = 0;
    __syncwarp();
#pragma unroll
    for (int i = 0; i < n; i++) if (gact) mP[g * LD + i] = P[i];        // [mode][direction]
    if (gact) { sv[g] = zus; sv[n + g] = zds; sv[2 * n + g] = qs; }
    __syncwarp();
    double bp[n], bm[n];
#pragma unroll
    for (int j = 0; j < n; j++) {
        double acc = 0.0;
#pragma unroll
        for (int i = 0; i < n; i++) acc = fma(P[i], mP[j * LD + i], acc);
        bp[j] = acc; bm[j] = acc;
    }
    {
        const double x = expm1(-kk * dtaucp);               // e^{-2kh} - 1
        const double th = -x / (2.0 + x);                   // tanh(kh)
        const double lam_p = kk * th;
        const double lam_m = (th > 1.0e-280) ? kk / th : 1.0e100;
#pragma unroll
        for (int j = 0; j < n; j++)
            if (j == g) { bp[j] += lam_p; bm[j] += lam_m; }
    }
    // Gauss-Jordan inversion of both SPD matrices, pivot rows broadcast through mB
#pragma unroll
    for (int j = 0; j < n; j++) {
        double *pb = mB + (j & 1) * 2 * n;
        if (g == j && gact) {
#pragma unroll
            for (int c = 0; c < n; c++) { pb[c] = bp[c]; pb[n + c] = bm[c]; }
        }
        __syncwarp();
        double pr[n], qr[n];
#pragma unroll
        for (int c = 0; c < n; c++) { pr[c] = pb[c]; qr[c] = pb[n + c]; }
        const double rp = fast_rcp(pr[j]), rq = fast_rcp(qr[j]);
        if (!(pr[j] > 0.0) || !(qr[j] > 0.0)) bad = 1;
        const double mp = (g == j) ? 1.0 - rp : bp[j] * rp, mq = (g == j) ? 1.0 - rq : bm[j] * rq;
        const double dp = (g == j) ? rp : -mp, dq = (g == j) ? rq : -mq;
#pragma unroll
        for (int c = 0; c < n; c++) {
            bp[c] = (c == j) ? dp : fma(-mp, pr[c], bp[c]);
            bm[c] = (c == j) ? dq : fma(-mq, qr[c], bm[c]);
        }
    }
    __syncwarp();
    {   // U+- = B+-^-1 P^T (row g) -> mA, mB
        double up[n], um[n];
#pragma unroll
        for (int b = 0; b < n; b++) { up[b] = 0.0; um[b] = 0.0; }
#pragma unroll
        for (int j = 0; j < n; j++) {
#pragma unroll
            for (int b = 0; b < n; b++) {
                const double pj = mP[j * LD + b];
                up[b] = fma(bp[j], pj, up[b]);
                um[b] = fma(bm[j], pj, um[b]);
            }
        }
#pragma unroll
        for (int b = 0; b < n; b++) if (gact) { mA[g * LD + b] = up[b]; mB[g * LD + b] = um[b]; }
    }
    __syncwarp();
    double R[n], T[n], mm[n];
    {
        double mp[n];
#pragma unroll
        for (int b = 0; b < n; b++) { mp[b] = 0.0; mm[b] = 0.0; }
#pragma unroll 1
        for (int j = 0; j < n; j++) {
            const double pa = mP[j * LD + g];
#pragma unroll
            for (int b = 0; b < n; b++) {
                mp[b] = fma(pa, mA[j * LD + b], mp[b]);
                mm[b] = fma(pa, mB[j * LD + b], mm[b]);
            }
        }
#pragma unroll
        for (int b = 0; b < n; b++) {
            R[b] = mp[b] + mm[b] - ((b == g) ? 1.0 : 0.0);
            T[b] = mp[b] - mm[b];
        }
    }
    double s_up = 0.0, s_dn = 0.0;
    if (beam) {
        double Rzd = 0.0, Tzu = 0.0, Tzd = 0.0, Rzu = 0.0;
#pragma unroll
        for (int b = 0; b < n; b++) {
            const double zu = sv[b], zd = sv[n + b];
            Rzd = fma(R[b], zd, Rzd); Tzu = fma(T[b], zu, Tzu);
            Tzd = fma(T[b], zd, Tzd); Rzu = fma(R[b], zu, Rzu);
        }
        s_up = et * (zus - Rzd) - eb * Tzu;
        s_dn = eb * (zds - Rzu) - et * Tzd;
    }
    if (therm) {
        double RD = 0.0, TD = 0.0, e = 0.0;
#pragma unroll
        for (int b = 0; b < n; b++) {
            const double Db = cmu[b] * csq[b];
            RD = fma(R[b], Db, RD); TD = fma(T[b], Db, TD);
            e = fma(mm[b], sv[2 * n + b], e);
        }
        e *= 2.0 * b1;
        const double Dg = cmu[g] * csq[g];
        s_up += (Dg - RD) * pk_top - TD * pk_bot + e;
        s_dn += (Dg - RD) * pk_bot - TD * pk_top - e;
    }
    if (active) {
#pragma unroll
        for (int b = 0; b < n; b++) { rec[AO::r_R + g * n + b] = R[b]; rec[AO::r_T + g * n + b] = T[b]; }
        rec[AO::r_su + g] = s_up;
        rec[AO::r_sd + g] = s_dn;
    }
    __syncwarp();
    return bad;
}

// Bottom-up sweep over the layers ncut-1 .. 0 (the whole warp on one layer, see sbd_adding.cu).
// On entry sm[0 .. n*n+n) holds Rb, sb of the bottom boundary; every record is overwritten with
// Y, y (d_below = Y d_above + y) and Rb, sb of the layer's top interface (u = Rb d + sb).
// sm: AddOps<n>::p2 doubles.  Returns false when a pivot vanished.
template <int n>
__device__ __forceinline__ bool adding_sweep_up(double *recs, int ncut, double *sm, int lane)
{
    using AO = AddOps<n>;
    constexpr int CG = AO::CG, CW = AO::CW, REC = AO::arec, KB = n > 8 ? 4 : 3, KM = (1 << KB) - 1;
    const bool act2 = lane < n * CG;
    const int i2 = act2 ? lane / CG : n - 1, p2 = lane % CG, c0 = p2 * CW;
    double *sRb = sm, *ssb = sRb + n * n, *sY = ssb + n, *sy = sY + n * n, *sW = sy + n, *sw = sW + n * n;
    double *rbuf = sw + n;
    warp_copy_async(rbuf, recs + (size_t)(ncut - 1) * REC, REC, lane);
    cp_async_commit();
    __syncwarp();
    bool ok = true;
    for (int lc = ncut - 1; lc >= 0; lc--) {
        const int buf = (ncut - 1 - lc) & 1;
        if (lc > 0) {
            warp_copy_async(rbuf + (buf ^ 1) * REC, recs + (size_t)(lc - 1) * REC, REC, lane);
            cp_async_commit();
            cp_async_wait_one();
        } else {
            cp_async_wait_all();
        }
        __syncwarp();
        const double *rc = rbuf + buf * REC;
        const double *Rr = rc + AO::r_R + i2 * n, *Tr = rc + AO::r_T + i2 * n;
        double b[CW], t[CW], v;
        {
            double r[n];
#pragma unroll
            for (int k = 0; k < n; k++) r[k] = Rr[k];
#pragma unroll
            for (int s = 0; s < CW; s++) b[s] = (c0 + s == i2) ? 1.0 : 0.0;
            v = rc[AO::r_sd + i2];
#pragma unroll
            for (int k = 0; k < n; k++) {
#pragma unroll
                for (int s = 0; s < CW; s++) b[s] = fma(-r[k], sRb[k * n + c0 + s], b[s]);
                v = fma(r[k], ssb[k], v);
            }
#pragma unroll
            for (int s = 0; s < CW; s++) t[s] = Tr[c0 + s];
        }
        unsigned used = 0;
        int myj = 0, sing = 0;
        double myrp = 0.0;
#pragma unroll
        for (int j = 0; j < n; j++) {
            const int pj = j / CW, sj = j % CW;
            const double colv = __shfl_sync(FULLMASK, b[sj], (lane & ~(CG - 1)) | pj);
            const int key = (act2 && !((used >> i2) & 1u)) ? ((__double2hiint(colv) & (0x7fffffff & ~KM)) | (KM - i2)) : -1;
            const int mx = __reduce_max_sync(FULLMASK, key);
            if ((mx >> KB) <= 0) sing = 1;
            const int ip = KM - (mx & KM);
            used |= 1u << ip;
            const int srcl = ip * CG + p2;
            const double rp = fast_rcp(__shfl_sync(FULLMASK, colv, ip * CG));
            const double m = (i2 == ip) ? 0.0 : colv * rp;
            if (i2 == ip) { myj = j; myrp = rp; }
#pragma unroll
            for (int s = 0; s < CW; s++) {
                b[s] = fma(-m, __shfl_sync(FULLMASK, b[s], srcl), b[s]);
                t[s] = fma(-m, __shfl_sync(FULLMASK, t[s], srcl), t[s]);
            }
            v = fma(-m, __shfl_sync(FULLMASK, v, srcl), v);
        }
        if (sing) { ok = false; break; }
        double *orec = recs + (size_t)lc * REC;          // (the layer's R, T are in shared memory)
        if (act2) {
#pragma unroll
            for (int s = 0; s < CW; s++) {
                const double y = t[s] * myrp;
                sY[myj * n + c0 + s] = y; orec[AO::o_Y + myj * n + c0 + s] = y;
            }
            if (p2 == 0) { const double y = v * myrp; sy[myj] = y; orec[AO::o_y + myj] = y; }
        }
        __syncwarp();
        {
            double rb[n], w[CW], wv = ssb[i2];
#pragma unroll
            for (int k = 0; k < n; k++) rb[k] = sRb[i2 * n + k];
#pragma unroll
            for (int s = 0; s < CW; s++) w[s] = 0.0;
#pragma unroll
            for (int k = 0; k < n; k++) {
#pragma unroll
                for (int s = 0; s < CW; s++) w[s] = fma(rb[k], sY[k * n + c0 + s], w[s]);
                wv = fma(rb[k], sy[k], wv);
            }
            if (act2) {
#pragma unroll
                for (int s = 0; s < CW; s++) sW[i2 * n + c0 + s] = w[s];
                if (p2 == 0) sw[i2] = wv;
            }
        }
        __syncwarp();
        {
            double tr[n], nr[CW], ns = rc[AO::r_su + i2];
#pragma unroll
            for (int k = 0; k < n; k++) tr[k] = Tr[k];
#pragma unroll
            for (int s = 0; s < CW; s++) nr[s] = Rr[c0 + s];
#pragma unroll
            for (int k = 0; k < n; k++) {
#pragma unroll
                for (int s = 0; s < CW; s++) nr[s] = fma(tr[k], sW[k * n + c0 + s], nr[s]);
                ns = fma(tr[k], sw[k], ns);
            }
            if (act2) {
#pragma unroll
                for (int s = 0; s < CW; s++) { sRb[i2 * n + c0 + s] = nr[s]; orec[AO::o_Rb + i2 * n + c0 + s] = nr[s]; }
                if (p2 == 0) { ssb[i2] = ns; orec[AO::o_sb + i2] = ns; }
            }
        }
        __syncwarp();
    }
    cp_async_wait_all();
    __syncwarp();
    return ok;
}

// Top-down sweep: the scaled intensities at every interface, levs[lev][0..n) = d (downward),
// levs[lev][n..2n) = u (upward), lev = 0 .. ncut.  d0: the downward intensity at the top boundary
// (scaled), bot: Rb [n][n] and sb [n] of the bottom boundary, u = Rb d + sb (zeros: nothing comes up).
template <int n>
__device__ __forceinline__ void adding_sweep_down(const double *recs, int ncut, double *levs, double d0,
                                                  const double *bot, const double *cmu, const double *csq, int lane)
{
    using AO = AddOps<n>;
    constexpr int REC = AO::arec;
    double d[n];
#pragma unroll
    for (int c = 0; c < n; c++) d[c] = cmu[c] * csq[c] * d0;
    const int row = lane < n ? lane : (lane < 2 * n ? lane - n : 0);
    double dme = cmu[row] * csq[row] * d0;          // lanes < n: d[lane]
    for (int lev = 0; lev <= ncut; lev++) {
        double x;
        {
            // lanes 0..n-1: row of Y (next d); lanes n..2n-1: row of Rb (u at this level; the
            // bottom boundary's own at the last one)
            const double *orec = recs + (size_t)lev * REC;
            const double *mrow = lev < ncut ? orec + (lane < n ? AO::o_Y : AO::o_Rb) + row * n : bot + row * n;
            x = lev < ncut ? orec[(lane < n ? AO::o_y : AO::o_sb) + row] : bot[n * n + row];
#pragma unroll
            for (int c = 0; c < n; c++) x = fma(mrow[c], d[c], x);
        }
        if (lane < n) levs[(size_t)lev * 2 * n + lane] = dme;
        else if (lane < 2 * n) levs[(size_t)lev * 2 * n + lane] = x;        // u_lev
        dme = x;
#pragma unroll
        for (int c = 0; c < n; c++) d[c] = __shfl_sync(FULLMASK, x, c);
    }
    __syncwarp();
}

}  // namespace sbd
