// Register-resident discrete-ordinate kernel for NSTR = 20, 24, 32: one CTA per bin.
//
// Same mathematics and the same three phases as sbd_fast.cu (see the header comment there),
// laid out for systems too large for one warp's registers:
//   phase 1  per-layer eigen / particular solutions: a layer is owned by a group of 16 lanes
//            (lane j holds row / column j of the n x n matrices, n <= 16), so the CTA's
//            2 x WARPS half-warps work on 2 x WARPS layers at once.
//   phase 2  downward elimination with partial pivoting.  The (n+N) x (2N+1) window is tiled
//            2-D over the CTA: 16 row groups x CG = N/4 column groups, a thread holding up to
//            KS = ceil(3n/16) rows x 8 columns (+ right-hand sides) in registers; warp w owns
//            column groups 2w and 2w+1 (one per half-warp).  A pivot step: the warp that owns
//            the pivot column finds the pivot (REDUX + VOTE) and publishes the pivot position,
//            its reciprocal and the column's entries through a double-buffered exchange area in
//            shared memory; after ONE CTA barrier every thread forms its multipliers from
//            that area and fetches its own 8-column slice of the pivot row with width-16
//            shuffles inside its warp.  Rows never travel through shared memory.
//   phase 3  upward back substitution by warp 0 (row per lane), fluxes as dot products with
//            the flux functionals of phase 1; all warps stage the pivot rows (cp.async).
// Layer records, flux records and pivot rows go through a per-CTA scratch slot in global memory.
#include <math.h>
#include <stdlib.h>

#include "sbd_devutil.cuh"
#include "sbd_internal.h"
#include "sbd_planck.cuh"

namespace sbd {

template <int n>
struct WideLayout {
    static constexpr int N = 2 * n;
    static constexpr int G = 16;                   // lanes per layer group / row groups
    static constexpr int CG = N / 4;               // column groups of LC window columns (2N = CG * LC)
    static constexpr int LC = 8;
    static constexpr int WARPS = (CG + 1) / 2;
    static constexpr int NTHR = WARPS * 32;
    static constexpr int KS = (3 * n + 15) / 16;   // row slots per row group
    static constexpr int US = CG * LC + 2;         // stored pivot row: CG slices, rhs, pad
    static constexpr int ublk = N * US;
    // per-layer record (doubles), as in FastLayout
    static constexpr int off_kk = 0;
    static constexpr int off_ek = n;
    static constexpr int off_gp = 2 * n;
    static constexpr int off_gm = 2 * n + n * n;
    static constexpr int off_zz = 2 * n + 2 * n * n;
    static constexpr int off_zp0 = off_zz + N;
    static constexpr int off_xr = off_zp0 + N;
    static constexpr int rec = ((off_xr + 2 + 1) / 2) * 2;
    static constexpr int f_cu = 0;
    static constexpr int f_ek = 3 * N;
    static constexpr int f_kk = 3 * N + n;
    static constexpr int f_sc = 3 * N + 2 * n;
    static constexpr int frec = f_sc + 10;
    __host__ __device__ static size_t slot_doubles(int L) { return (size_t)L * (rec + frec + ublk); }
    // shared memory (doubles)
    static constexpr int cta = 4 * n + N * n + 2;             // cmu cwt csq cdinv, sum(w mu), sum(w), ylm
    static constexpr int tasks = 2 * WARPS;
    static constexpr int task = N + 2 * n * n + 6 * n;        // gl, K, L, vectors, 1/diagonals
    static constexpr int xch = 2 * (KS * G + 4);              // double-buffered pivot exchange
    static constexpr int cmax(int a, int b) { return a > b ? a : b; }
    static constexpr int work = cmax(cmax(tasks * task, 3 * rec + ublk + xch), 2 * (ublk + frec));
    __host__ __device__ static size_t bin_doubles(int L, int NT)
    {
        // y0, solution of the current layer, work area, taucpr / tauc, beam transmissions (2),
        // pk (+2 boundary temperatures), prologue work values, level map, scalars
        size_t d = (size_t)2 * N + work + 4 * (L + 1) + (L + 3) + 3 * L + (NT + 1) / 2 + 8;
        return (d + 1) & ~(size_t)1;
    }
};

template <int n>
__device__ __forceinline__ unsigned long long wide_partners(int g)
{
    // round-robin (tournament) partner of lane g in round r, 4 bits per round (n <= 16, even)
    unsigned long long pk = 0;
#pragma unroll
    for (int r = 0; r < n - 1; r++) {
        int partner;
        if (g >= n - 1) partner = r;
        else if (g == r) partner = n - 1;
        else {
            partner = 2 * r - g + (n - 1);
            if (partner >= n - 1) partner -= n - 1;
            if (partner >= n - 1) partner -= n - 1;
        }
        pk |= (unsigned long long)partner << (4 * r);
    }
    return pk;
}

// sum over the 16 lanes of a layer group
__device__ __forceinline__ double group16_sum(double v)
{
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(FULLMASK, v, o, 16);
    return v;
}

// ---------------------------------------------------------------------------
// phase 1: one layer per group of 16 lanes (lanes g >= n mirror lane n-1 and write nothing)
// ---------------------------------------------------------------------------
template <int n>
__device__ __forceinline__ int wide_phase1(
    const double *__restrict__ dtauc, const double *__restrict__ ssalb,
    const double *__restrict__ pmom, int ldp, int lc, bool active,
    double fbeam, double umu0, bool plank,
    const double *cmu, const double *cwt, const double *csq, const double *cdinv, const double *cylm,
    const double *y0, const double *taucpr, const double *pk,
    double *tsm, double *rec, double *frec, int glane, unsigned long long jpart)
{
    using WL = WideLayout<n>;
    constexpr int N = 2 * n;
    const int g = glane < n ? glane : n - 1;
    const bool wr = active && glane < n;          // this lane publishes results
    double *sgl = tsm, *sK = sgl + N, *sL = sK + n * n,
           *sv = sL + n * n, *srd = sv + 4 * n;    // sv: 4 vectors of n; srd: 1/diag(K), 1/diag(L)

    double ss = ssalb[lc];
    if (ss == 1.0) ss = 1.0 - kDither;
    double dt = dtauc[lc];
    if (dt < 0.0) dt = 0.0;
    const double f = pmom[(size_t)lc * ldp + N];
    const double oprim = ss * (1. - f) * fast_rcp(1. - f * ss);
    const double dtaucp = (1. - f * ss) * dt;
    const double rf = fast_rcp(1. - f);
    for (int l = glane; l < N; l += 16) {
        double pm = (l == 0) ? 1.0 : pmom[(size_t)lc * ldp + l];
        sgl[l] = (2 * l + 1) * oprim * (pm - f) * rf;
    }
    __syncwarp();

    // rows g of Pe~ and Po~ (m = 0)
    double pe[n], po[n];
#pragma unroll
    for (int j = 0; j < n; j++) { pe[j] = 0.0; po[j] = 0.0; }
#pragma unroll 2
    for (int l = 0; l < N; l++) {
        const double t = sgl[l] * cylm[l * n + g];
        if (l & 1) {
#pragma unroll
            for (int j = 0; j < n; j++) po[j] = fma(t, cylm[l * n + j], po[j]);
        } else {
#pragma unroll
            for (int j = 0; j < n; j++) pe[j] = fma(t, cylm[l * n + j], pe[j]);
        }
    }
    const double sqg = csq[g], rmu = fast_rcp(cmu[g]);
#pragma unroll
    for (int j = 0; j < n; j++) {
        const double sc = sqg * csq[j];
        const double dg = (j == g) ? rmu : 0.0;
        pe[j] = dg - sc * pe[j];
        po[j] = dg - sc * po[j];
    }

    // row Cholesky of both operators: Po~ = L L^T, Pe~ = K K^T (lane g = row g)
    int bad = 0;
#pragma unroll
    for (int j = 0; j < n; j++) {
        double nume = pe[j], numo = po[j];
#pragma unroll
        for (int k = 0; k < j; k++) {
            nume = fma(-pe[k], shfl_d(pe[k], j, 16), nume);
            numo = fma(-po[k], shfl_d(po[k], j, 16), numo);
        }
        double pive = shfl_d(nume, j, 16), pivo = shfl_d(numo, j, 16);
        if (!(pivo > 0.0)) { bad = 1; pivo = 1.0; }
        const double floor_e = 1.0e-30;
        if (pive != pive) bad = 1;
        if (!(pive > floor_e)) pive = floor_e;
        const double rie = fast_rsqrt(pive), rio = fast_rsqrt(pivo);
        pe[j] = (g == j) ? pive * rie : ((g > j) ? nume * rie : 0.0);
        po[j] = (g == j) ? pivo * rio : ((g > j) ? numo * rio : 0.0);
    }
    if (glane < n) {
#pragma unroll
        for (int j = 0; j < n; j++) { sK[g * n + j] = pe[j]; sL[g * n + j] = po[j]; }
    }
    __syncwarp();
    // reciprocal diagonals (uniform in the group), kept in shared memory to save registers
    if (glane < n) { srd[g] = 1.0 / sK[g * n + g]; srd[n + g] = 1.0 / sL[g * n + g]; }
    __syncwarp();

    // column g of A = K^T L
    double a[n];
    {
#pragma unroll
        for (int i = 0; i < n; i++) a[i] = 0.0;
#pragma unroll 2
        for (int k = 0; k < n; k++) {
            const double lk = sL[k * n + g];
#pragma unroll
            for (int i = 0; i < n; i++)
                if (i <= k) a[i] = fma(sK[k * n + i], lk, a[i]);
        }
    }

    // One-sided Jacobi (see sbd_fast.cu): only A is rotated
    {
        double own2 = 0.0;
#pragma unroll
        for (int i = 0; i < n; i++) own2 = fma(a[i], a[i], own2);
        for (int sweep = 0; sweep < 60; sweep++) {
            int big = 0;
#pragma unroll 1
            for (int r = 0; r < n - 1; r++) {
                const int partner = (int)((jpart >> (4 * r)) & 15);
                double pa[n];
                double g0 = 0.0, g1 = 0.0;
                const double oth2 = shfl_d(own2, partner, 16);
#pragma unroll
                for (int i = 0; i < n; i++) {
                    pa[i] = shfl_d(a[i], partner, 16);
                    if (i & 1) g1 = fma(a[i], pa[i], g1); else g0 = fma(a[i], pa[i], g0);
                }
                const double gam = g0 + g1;
                const bool lo = g < partner;
                const double gg = gam * gam, ab = own2 * oth2;
                if (gg > 1.0e-24 * ab && gg > 1.0e-290) {
                    if (gg > kJacobiBig * ab) big = 1;
                    const double dl = lo ? 0.5 * (oth2 - own2) : 0.5 * (own2 - oth2);
                    const double rh = fast_rsqrt(fma(dl, dl, gg));
                    const double x = fma(0.5 * fabs(dl), rh, 0.5);
                    const double rc = fast_rsqrt(x);
                    const double cc = x * rc;
                    double sn = gam * (0.5 * rh) * rc;
                    if ((dl < 0.0) != lo) sn = -sn;
                    own2 = fma(sn * rc, gam, own2);
#pragma unroll
                    for (int i = 0; i < n; i++) a[i] = fma(sn, pa[i], cc * a[i]);
                }
            }
            if (!__any_sync(FULLMASK, big)) break;
            own2 = 0.0;
#pragma unroll
            for (int i = 0; i < n; i++) own2 = fma(a[i], a[i], own2);
        }
    }

    double s2 = 0.0;
#pragma unroll
    for (int i = 0; i < n; i++) s2 = fma(a[i], a[i], s2);
    const double kk = sqrt(s2);
    const double ek = exp(-kk * dtaucp);
    // column g of P = K^-T a (in place in a), then Q = L^-T (L^-1 P)
#pragma unroll
    for (int i = n - 1; i >= 0; i--) {
        double acc = a[i];
#pragma unroll
        for (int k = i + 1; k < n; k++) acc = fma(-sK[k * n + i], a[k], acc);
        a[i] = acc * srd[i];
    }
    double Q[n];
    {
#pragma unroll
        for (int i = 0; i < n; i++) {
            double acc = a[i];
#pragma unroll
            for (int k = 0; k < i; k++) acc = fma(-sL[i * n + k], Q[k], acc);
            Q[i] = acc * srd[n + i];
        }
#pragma unroll
        for (int i = n - 1; i >= 0; i--) {
            double acc = Q[i];
#pragma unroll
            for (int k = i + 1; k < n; k++) acc = fma(-sL[k * n + i], Q[k], acc);
            Q[i] = acc * srd[n + i];
        }
    }
    // G+ - G- = D^-1 Q ; G+ + G- = -D^-1 P / k   (disort.f:3273-3301)
    const double rk = fast_rcp(kk);
    double fA = 0.0, fB = 0.0, fC = 0.0;
#pragma unroll
    for (int i = 0; i < n; i++) {
        const double gd = cdinv[i] * Q[i];
        const double gs = -cdinv[i] * a[i] * rk;
        const double gpi = 0.5 * (gs + gd), gmi = 0.5 * (gs - gd);
        const double wm = cwt[i] * cmu[i];
        fA = fma(wm, gpi, fA);
        fB = fma(wm, gmi, fB);
        fC = fma(cwt[i], gs, fC);
        if (wr) {
            rec[WL::off_gp + i * n + g] = gpi;
            rec[WL::off_gm + i * n + g] = gmi;
        }
    }
    if (wr) {
        rec[WL::off_kk + g] = kk;
        rec[WL::off_ek + g] = ek;
        frec[WL::f_cu + n + g] = fA;          frec[WL::f_cu + n - 1 - g] = -fB;
        frec[WL::f_cu + N + n + g] = fB;      frec[WL::f_cu + N + n - 1 - g] = -fA;
        frec[WL::f_cu + 2 * N + n + g] = fC;  frec[WL::f_cu + 2 * N + n - 1 - g] = -fC;
        frec[WL::f_ek + g] = ek;
        frec[WL::f_kk + g] = kk;
    }

    // ---- beam particular solution: spectral form of UPBEAM (disort.f:4130) ----
    double zup = 0.0, zdn = 0.0;
    if (fbeam > 0.0) {
        const double fac = fbeam * (1.0 / (4. * kPiRef));       // (2 - delta_m0) = 1 for m = 0
        const double rmu0 = fast_rcp(umu0);
        double be = 0.0, bo = 0.0;
#pragma unroll 4
        for (int l = 0; l < N; l++) {
            const double t = sgl[l] * cylm[l * n + g] * y0[l];
            if (l & 1) bo += t; else be += t;
        }
        const double bs = 2.0 * fac * sqg * be, bd = 2.0 * fac * sqg * bo;
        __syncwarp();
        if (glane < n) sv[g] = bd;
        __syncwarp();
        double t1 = 0.0;                                    // (K^T b^_d)_g
#pragma unroll 4
        for (int k = 0; k < n; k++) t1 = fma(sK[k * n + g], sv[k], t1);
        if (glane < n) sv[n + g] = t1;
        __syncwarp();
        double t2 = 0.0;                                    // (K K^T b^_d)_g
#pragma unroll 4
        for (int k = 0; k < n; k++) t2 = fma(sK[g * n + k], sv[n + k], t2);
        if (glane < n) sv[2 * n + g] = bs * rmu0 - t2;      // r_g
        __syncwarp();
        double cj = 0.0;                                    // (P^T r)_g / (1/mu0^2 - k^2)
#pragma unroll
        for (int i = 0; i < n; i++) cj = fma(a[i], sv[2 * n + i], cj);
        cj = cj * fast_rcp(rmu0 * rmu0 - s2);
        if (glane < n) { sv[3 * n + g] = cj; sv[g] = cj * kk; }
        __syncwarp();
        // d_i = sum_j Gd(i,j) c_j, s_i = sum_j Gs(i,j) k_j c_j: lane j holds column j of Gd = D^-1 Q and
        // Gs = -D^-1 P / k in registers; the sums over the lanes go to the lane of direction i
        double dv = 0.0, sv2 = 0.0;
        const double cm = glane < n ? cj : 0.0;
#pragma unroll
        for (int i = 0; i < n; i++) {
            const double td = group16_sum(cdinv[i] * Q[i] * cm);
            const double ts = group16_sum(-cdinv[i] * a[i] * cm);       // (k_j c_j) Gs(i,j) = -D^-1 P c_j
            if (i == g) { dv = td; sv2 = ts; }
        }
        sv2 = umu0 * (cdinv[g] * bd + sv2);
        zup = 0.5 * (sv2 + dv);
        zdn = 0.5 * (sv2 - dv);
        __syncwarp();
    }
    // ---- thermal particular solution (UPISOT, disort.f:4247) -----------------
    double xr0 = 0.0, xr1 = 0.0, q = 0.0;
    if (plank) {
        if (dtaucp > 1.0e-200) xr1 = (pk[lc + 1] - pk[lc]) * fast_rcp(dtaucp);
        else if (dtaucp > 0.0) xr1 = (pk[lc + 1] - pk[lc]) / dtaucp;
        xr0 = pk[lc] - xr1 * taucpr[lc];
        // q = D^-1 L^-T L^-1 D 1, element g: two triangular solves, every lane does them all
        double y[n];
#pragma unroll
        for (int i = 0; i < n; i++) {
            double acc = cmu[i] * csq[i];
#pragma unroll
            for (int k = 0; k < i; k++) acc = fma(-sL[i * n + k], y[k], acc);
            y[i] = acc * srd[n + i];
        }
#pragma unroll
        for (int i = n - 1; i >= 0; i--) {
            double acc = y[i];
#pragma unroll
            for (int k = i + 1; k < n; k++) acc = fma(-sL[k * n + i], y[k], acc);
            y[i] = acc * srd[n + i];
            if (i == g) q = cdinv[i] * y[i];
        }
    }
    {
        const double wmg = glane < n ? cwt[g] * cmu[g] : 0.0, wg = glane < n ? cwt[g] : 0.0;
        const double Zu = group16_sum(wmg * zup);
        const double Zd = group16_sum(wmg * zdn);
        const double Za = group16_sum(wg * (zup + zdn));
        const double Q1 = plank ? group16_sum(wmg * q) : 0.0;
        if (active && glane == 0) {
            const double W = cylm[-2], SW = cylm[-1];
            double *sc = frec + WL::f_sc;
            sc[0] = Zu; sc[1] = Zd; sc[2] = Za;
            sc[3] = fma(xr1, Q1, xr0 * W); sc[4] = fma(-xr1, Q1, xr0 * W); sc[5] = 2.0 * xr0 * SW;
            sc[6] = xr0; sc[7] = xr1; sc[8] = 1.0 - ss; sc[9] = 1.0 - ss * f;
        }
    }
    if (wr) {
        rec[WL::off_zz + n + g] = zup;
        rec[WL::off_zz + n - 1 - g] = zdn;
        rec[WL::off_zp0 + n + g] = xr0 + xr1 * q;
        rec[WL::off_zp0 + n - 1 - g] = xr0 - xr1 * q;
        if (g == 0) { rec[WL::off_xr] = xr0; rec[WL::off_xr + 1] = xr1; }
    }
    __syncwarp();
    return (bad && active) ? SBD_BIN_EIG_FAIL : 0;
}

// ---------------------------------------------------------------------------
// phase 2 helpers
// ---------------------------------------------------------------------------
// Rows of the boundary system are assembled by the whole CTA into a staging area in the
// layout the pivot rows are stored in: row r at r*US, window column c at (c % CG) * LC + c / CG,
// right-hand side at CG*LC.  See stage_rows of sbd_fast.cu for the meaning of the arguments.
template <int n>
__device__ __forceinline__ void wide_stage_rows(double *stg, const double *recA, bool bottomA,
                                                const double *recB, int r0, int nrows, double refl,
                                                const double *cwt, const double *cmu, int tid)
{
    using WL = WideLayout<n>;
    constexpr int N = 2 * n, LC = WL::LC, US = WL::US, CG = WL::CG;
    for (int e = tid; e < nrows * 2 * N; e += WL::NTHR) {
        const int rr = e / (2 * N), c = e - rr * (2 * N);
        const bool isB = c >= N;
        const int cc = isB ? c - N : c;
        const bool plus = cc >= n;
        const int j = plus ? cc - n : n - 1 - cc;
        const double *rec = (isB && recB) ? recB : recA;
        const bool bottom = isB ? false : bottomA;
        double fac = (plus == bottom) ? rec[WL::off_ek + j] : 1.0;
        if (!plus) fac = -fac;
        if (isB) fac = recB ? -fac : 0.0;
        double rsum = 0.0;
        if (refl != 0.0) {
            const double *gdn = rec + (plus ? WL::off_gm : WL::off_gp) + j;
#pragma unroll 1
            for (int k = 0; k < n; k++) rsum = fma(cwt[k] * cmu[k], gdn[k * n], rsum);
        }
        const int r = r0 + rr;
        const double *gup = rec + (plus ? WL::off_gp : WL::off_gm) + j;
        const double *gdw = rec + (plus ? WL::off_gm : WL::off_gp) + j;
        const double v = r >= n ? gup[(r - n) * n] : gdw[(n - 1 - r) * n];
        stg[rr * US + (c % CG) * LC + c / CG] = (v - refl * rsum) * fac;
    }
}

// resident CTAs per SM the register budget is sized for (4: 128 registers per thread; measured on the
// NSTR=32 / 65-layer set: 2 CTAs 162 k, 3 CTAs 215 k, 4 CTAs 239 k bins/s -- occupancy beats the spills)
#ifndef SBD_WIDE_MINB
#define SBD_WIDE_MINB 4
#endif
int wide_ctas_per_sm() { return SBD_WIDE_MINB; }

template <int n>
__global__ void __launch_bounds__(WideLayout<n>::NTHR, SBD_WIDE_MINB)
disort_wide_kernel(const LaunchArgs a)
{
    using WL = WideLayout<n>;
    constexpr int N = 2 * n, KS = WL::KS, LC = WL::LC, US = WL::US, CG = WL::CG, NTHR = WL::NTHR;
    const int L = a.d.nlyr;
    const int NT = a.d.ntau > 0 ? a.d.ntau : L + 1;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ldp = a.d.nmom + 1;
    extern __shared__ double smem_wide[];
    double *cmu = smem_wide, *cwt = cmu + n, *csq = cwt + n, *cdinv = csq + n;
    double *cylm = cdinv + n + 2;       // cylm[-2] = sum(w mu), cylm[-1] = sum(w)
    double *bsm = smem_wide + WL::cta;
    double *y0 = bsm, *xsm = y0 + N;
    double *wk = xsm + N;               // 16-byte aligned work area shared by the phases
    double *taucpr = wk + WL::work, *tauc = taucpr + (L + 1);
    double *ebeam = tauc + (L + 1), *edir = ebeam + (L + 1);
    double *pk = edir + (L + 1);
    double *lw = pk + (L + 3);
    int *layru = (int *)(lw + 3 * L);
    int *misc = layru + ((NT + 1) / 2) * 2;       // bin, ncut, lyrcut, monotone, status flags ...

    for (int i = tid; i < n; i += NTHR) {
        double mu = a.quad[i], wt = a.quad[n + i];
        cmu[i] = mu; cwt[i] = wt; csq[i] = sqrt(wt / mu); cdinv[i] = 1.0 / sqrt(wt * mu);
    }
    if (tid == 0) {
        double W = 0.0, SW = 0.0;
        for (int i = 0; i < n; i++) { W += a.quad[n + i] * a.quad[i]; SW += a.quad[n + i]; }
        cylm[-2] = W; cylm[-1] = SW;
    }
    for (int e = tid; e < N * n; e += NTHR) cylm[e] = a.ylmc[e];
    __syncthreads();
    const double Wq = cylm[-2], SWq = cylm[-1];

    double *scr = a.scratch + (size_t)blockIdx.x * a.slot_stride;
    double *recs = scr;                                   // [L][rec]
    double *frecs = scr + (size_t)L * WL::rec;            // [L][frec]
    double *ublk = frecs + (size_t)L * WL::frec;          // [L][N][US]
    const int glane = lane & 15, gid = tid >> 4;          // phase 1: layer group
    const int rg = lane & 15, cg = warp * 2 + (lane >> 4);   // phase 2: 2-D tiling
    const bool cgreal = cg < CG;
    const unsigned long long jpart = wide_partners<n>(glane);

    for (;;) {
        __syncthreads();
        if (tid == 0) misc[0] = atomicAdd(a.work_counter, 1);
        __syncthreads();
        int bin = misc[0];
        if (bin >= (a.redo_consume ? *a.redo_count : (a.nbins_dev ? *a.nbins_dev : a.d.nbins))) break;
        if (a.redo_consume) bin = a.redo_list[bin];       // bins handed over by the adding kernel
        const int src = a.binmap ? a.binmap[bin] : bin;
        const sbd_bin bp = a.bins[src];
        const double *dtauc = a.dtauc + (size_t)src * L;
        const double *ssalb = a.ssalb + (size_t)src * L;
        const double *pmom = a.pmom + (size_t)src * L * ldp;
        const double fbeam = bp.fbeam, umu0 = bp.umu0, albedo = bp.albedo;
        const bool plank = bp.plank != 0;
        double *o_rfldir = a.rfldir ? a.rfldir + (size_t)bin * NT : nullptr;
        double *o_rfldn = a.rfldn ? a.rfldn + (size_t)bin * NT : nullptr;
        double *o_flup = a.flup ? a.flup + (size_t)bin * NT : nullptr;
        double *o_dfdt = a.dfdt ? a.dfdt + (size_t)bin * NT : nullptr;
        double *o_uavg = a.uavg ? a.uavg + (size_t)bin * NT : nullptr;

        // ---- CHEKIN subset (disort.f:4920-5155) and SETDIS prologue (:2546-2605)
        int badl = 0;
        for (int lc = tid; lc < L; lc += NTHR) {
            double s = ssalb[lc];
            if (!(s >= 0.0 && s <= 1.0)) badl = 1;
            if (!(fabs(dtauc[lc]) <= 1.79e308)) badl = 1;
            for (int k = 1; k <= a.d.nmom; k++) {
                double pm = pmom[(size_t)lc * ldp + k];
                if (!(pm >= -1.0 && pm <= 1.0)) badl = 1;
            }
            if (s == 1.0) s = 1.0 - kDither;
            const double dtr = dtauc[lc];
            const double dt = dtr < 0.0 ? 0.0 : dtr;
            const double f = pmom[(size_t)lc * ldp + N];
            lw[lc] = dtr; lw[L + lc] = (1. - s) * dt; lw[2 * L + lc] = (1. - f * s) * dt;
        }
        if (fbeam < 0.0 || (fbeam > 0.0 && !(umu0 > 0.0 && umu0 <= 1.0))) badl = 1;
        if (!(albedo >= 0.0 && albedo <= 1.0) || bp.fisot < 0.0) badl = 1;
        if (plank && (bp.wvnmlo < 0.0 || bp.wvnmhi <= bp.wvnmlo || bp.temis < 0.0 ||
                      bp.temis > 1.0 || bp.btemp < 0.0 || bp.ttemp < 0.0)) badl = 1;
        if (plank && (!a.temper || bp.col < 0 || bp.col >= a.d.ncol)) badl = 1;
        int clash = 0;
        if (fbeam > 0.0 && tid < n && fabs(umu0 - cmu[tid]) / umu0 < 1.e-4) clash = 1;
        const int anybad = __syncthreads_or(badl);
        const int anyclash = __syncthreads_or(clash);
        int status = anybad ? SBD_BIN_BAD_INPUT : (anyclash ? SBD_BIN_ANGLE_CLASH : 0);
        if (tid == 0) {
            double tc = 0.0, tp = 0.0, abstau = 0.0;
            int ncut = L, monotone = 1;
            tauc[0] = 0.0; taucpr[0] = 0.0;
            for (int lc = 0; lc < L; lc++) {
                if (lw[lc] < 0.0) monotone = 0;
                tc += lw[lc];
                if (abstau < 10.0) ncut = lc + 1;
                abstau += lw[L + lc];
                tp += lw[2 * L + lc];
                tauc[lc + 1] = tc; taucpr[lc + 1] = tp;
            }
            const int lyrcut = (abstau >= 10.0 && !plank && L > 1);
            if (!lyrcut) ncut = L;
            misc[1] = ncut; misc[2] = lyrcut; misc[3] = monotone;
        }
        __syncthreads();
        const int ncut = misc[1], lyrcut = misc[2], monotone = misc[3];
        const bool fastmap = (a.d.ntau == 0) && monotone;
        for (int lev = tid; lev <= L; lev += NTHR) {
            ebeam[lev] = fbeam > 0.0 ? exp(-taucpr[lev] / umu0) : 0.0;
            edir[lev] = fbeam > 0.0 ? exp(-tauc[lev] / umu0) : 0.0;
        }
        int badtau = 0;
        for (int lu = tid; lu < NT; lu += NTHR) {
            double ut = a.d.ntau > 0 ? a.utau[(size_t)src * NT + lu] : tauc[lu];
            if (a.d.ntau > 0 && fabs(ut - tauc[L]) <= 1.e-4) ut = tauc[L];
            if (a.d.ntau > 0 && !(ut >= 0.0 && ut <= tauc[L])) badtau = 1;
            int lc;
            if (fastmap && lu >= 1 && tauc[lu - 1] < ut) {
                lc = lu;
            } else {
                for (lc = 1; lc <= L; lc++)
                    if (ut >= tauc[lc - 1] && ut <= tauc[lc]) break;
                if (lc > L) lc = L;
            }
            layru[lu] = lc;
            if (o_rfldir) o_rfldir[lu] = 0.0;
            if (o_rfldn) o_rfldn[lu] = 0.0;
            if (o_flup) o_flup[lu] = 0.0;
            if (o_dfdt) o_dfdt[lu] = 0.0;
            if (o_uavg) o_uavg[lu] = 0.0;
        }
        if (__syncthreads_or(badtau)) status = SBD_BIN_BAD_INPUT;
        double tplank = 0.0, bplank = 0.0;
        if (plank && !status) {
            const double *tp = a.temper + (size_t)bp.col * (L + 1);
            for (int lev = tid; lev <= L + 2; lev += NTHR) {
                const double t = lev <= L ? tp[lev] : (lev == L + 1 ? bp.ttemp : bp.btemp);
                pk[lev] = plkavg_dev(bp.wvnmlo, bp.wvnmhi, t);
            }
        }
        if (tid == 0 && fbeam > 0.0) {   // Y_l^0(-mu0), LEPOLY m = 0
            double x = -umu0;
            y0[0] = 1.0; y0[1] = x;
            double pm2 = 1.0, pm1 = x;
            for (int l = 2; l < N; l++) {
                const double p = ((2 * l - 1) * x * pm1 - (l - 1) * pm2) * (1.0 / l);
                y0[l] = p; pm2 = pm1; pm1 = p;
            }
        }
        __syncthreads();
        if (plank && !status) { tplank = bp.temis * pk[L + 1]; bplank = pk[L + 2]; }

        // ===================== phase 1 =====================================
        if (!status) {
            int st = 0;
            for (int lc0 = 0; lc0 < ncut; lc0 += WL::tasks) {
                int lc = lc0 + gid;
                const bool active = lc < ncut;
                if (!active) lc = ncut - 1;
                st |= wide_phase1<n>(dtauc, ssalb, pmom, ldp, lc, active, fbeam, umu0, plank,
                                     cmu, cwt, csq, cdinv, cylm, y0, taucpr, pk, wk + (size_t)gid * WL::task,
                                     recs + (size_t)lc * WL::rec, frecs + (size_t)lc * WL::frec, glane, jpart);
            }
            __threadfence_block();
            if (__syncthreads_or(st != 0)) status = SBD_BIN_EIG_FAIL;
        }
        __syncthreads();

        // ===================== phase 2: downward elimination ================
        // slot s = k*16 + rg holds a window row; `act` (uniform) marks the live equations
        double *rslot = wk;
        double *stg = wk + 3 * WL::rec;
        double *xch = stg + WL::ublk;            // 2 x { column entries [KS*16], rp, info, pad, pad }
        constexpr int XS = KS * 16 + 4;
        if (!status) {
            double w[KS][LC], rhs[KS];
            for (int i = tid; i < WL::rec / 2; i += NTHR) cp_async16(rslot + 2 * i, recs + 2 * i);
            if (ncut > 1)
                for (int i = tid; i < WL::rec / 2; i += NTHR) cp_async16(rslot + WL::rec + 2 * i, recs + WL::rec + 2 * i);
            cp_async_commit();
            cp_async_wait_all();
            __syncthreads();
            // top boundary rows r = 0..n-1 (disort.f:2887-2915, :3547-3550)
            wide_stage_rows<n>(stg, rslot, false, nullptr, 0, n, 0.0, cwt, cmu, tid);
            if (tid < n)
                stg[tid * US + CG * LC] = bp.fisot + tplank - rslot[WL::off_zz + tid] - rslot[WL::off_zp0 + tid];
            __syncthreads();
#pragma unroll
            for (int k = 0; k < KS; k++) {
                const int r = k * 16 + rg;
                rhs[k] = 0.0;
#pragma unroll
                for (int l = 0; l < LC; l++) w[k][l] = 0.0;
                if (r < n && cgreal) {
#pragma unroll
                    for (int l = 0; l < LC; l++) w[k][l] = stg[r * US + cg * LC + l];
                    rhs[k] = stg[r * US + CG * LC];
                }
            }
            unsigned long long act = (1ull << n) - 1ull;
            constexpr unsigned long long slotmask = (KS * 16 >= 64) ? ~0ull : ((1ull << (KS * 16)) - 1ull);
            __syncthreads();
            int step = 0;
            bool sing = false;
            for (int lc = 0; lc < ncut && !sing; lc++) {
                const bool last = (lc == ncut - 1);
                if (lc + 2 < ncut) {
                    double *dst = rslot + ((lc + 2) % 3) * WL::rec;
                    const double *sp = recs + (size_t)(lc + 2) * WL::rec;
                    for (int i = tid; i < WL::rec / 2; i += NTHR) cp_async16(dst + 2 * i, sp + 2 * i);
                    cp_async_commit();
                }
                const double *rc = rslot + (lc % 3) * WL::rec;
                const double *rn = rslot + ((lc + 1) % 3) * WL::rec;
                const double tb = taucpr[lc + 1];
                const double eb = ebeam[lc + 1];
                const unsigned long long freem = ~act & slotmask;
                const int nnew = last ? n : N;
                unsigned long long newm = freem;
                if (__popcll(freem) > nnew) {          // keep the lowest nnew free slots
                    newm = 0;
                    unsigned long long f = freem;
                    for (int i = 0; i < nnew; i++) { const unsigned long long b = f & (0ull - f); newm |= b; f ^= b; }
                }
                if (!last) {
                    wide_stage_rows<n>(stg, rc, true, rn, 0, N, 0.0, cwt, cmu, tid);
                    if (tid < N)
                        stg[tid * US + CG * LC] = (rn[WL::off_zz + tid] - rc[WL::off_zz + tid]) * eb +
                                                   rn[WL::off_zp0 + tid] - rc[WL::off_zp0 + tid] +
                                                   (rn[WL::off_xr + 1] - rc[WL::off_xr + 1]) * tb;
                } else {
                    wide_stage_rows<n>(stg, rc, true, nullptr, n, n, lyrcut ? 0.0 : 2.0 * albedo, cwt, cmu, tid);
                    if (tid < n) {
                        const int r = n + tid;
                        const double xr1 = rc[WL::off_xr + 1];
                        double v = -rc[WL::off_zz + r] * eb - rc[WL::off_zp0 + r] - xr1 * tb;
                        if (!lyrcut) {
                            double rsum = 0.0;
#pragma unroll 1
                            for (int k = 0; k < n; k++)
                                rsum = fma(cwt[k] * cmu[k], rc[WL::off_zz + n - 1 - k] * eb +
                                                                rc[WL::off_zp0 + n - 1 - k] + xr1 * tb, rsum);
                            v += 2.0 * albedo * rsum + albedo * umu0 * fbeam / kPiRef * eb +
                                 (1.0 - albedo) * bplank;
                        }
                        stg[tid * US + CG * LC] = v;
                    }
                }
                __syncthreads();
#pragma unroll
                for (int k = 0; k < KS; k++) {
                    const int s = k * 16 + rg;
                    if ((newm >> s) & 1ull) {
                        const double *row = stg + __popcll(freem & ((1ull << s) - 1ull)) * US;
                        if (cgreal) {
#pragma unroll
                            for (int l2 = 0; l2 < LC / 2; l2++) {
                                const double2 v = reinterpret_cast<const double2 *>(row + cg * LC)[l2];
                                w[k][2 * l2] = v.x; w[k][2 * l2 + 1] = v.y;
                            }
                        }
                        rhs[k] = row[CG * LC];
                    }
                }
                act |= newm;
                // eliminate the N columns of layer lc
                double *urow = ublk + (size_t)lc * WL::ublk;
#pragma unroll 1
                for (int j = 0; j < N; j++, urow += US, step++) {
                    const int cgj = j % CG;
                    double *xb = xch + (step & 1) * XS;
                    if (warp == (cgj >> 1)) {
                        // pivot search among the live rows of the owner half-warp
                        const bool own = (lane >> 4) == (cgj & 1);
                        int best = -1;
                        double bval = 1.0;
#pragma unroll
                        for (int k = 0; k < KS; k++) {
                            const bool live = own && ((act >> (k * 16 + rg)) & 1ull);
                            const int h = live ? ((__double2hiint(w[k][0]) & 0x7ffffffc) | (KS - 1 - k)) : -1;
                            if (h > best) { best = h; bval = w[k][0]; }
                            if (own) xb[k * 16 + rg] = w[k][0];
                        }
                        const double rloc = fast_rcp(bval);
                        const int mx = __reduce_max_sync(FULLMASK, best);
                        const unsigned who = __ballot_sync(FULLMASK, best == mx);
                        const int pl = __ffs(who) - 1;
                        const double rp = -__shfl_sync(FULLMASK, rloc, pl);
                        if (lane == 0) {
                            xb[KS * 16] = rp;
                            const int kp = KS - 1 - (mx & 3), rgp = pl & 15;
                            reinterpret_cast<int *>(xb + KS * 16 + 1)[0] = ((mx >> 2) <= 0) ? -1 : (kp * 16 + rgp);
                        }
                    }
                    __syncthreads();
                    const int pinfo = reinterpret_cast<const int *>(xb + KS * 16 + 1)[0];
                    if (pinfo < 0) { sing = true; break; }
                    const int kp = pinfo >> 4, rgp = pinfo & 15;
                    const double rp = xb[KS * 16];
                    double m[KS];
#pragma unroll
                    for (int k = 0; k < KS; k++) m[k] = xb[k * 16 + rg] * rp;
                    // this thread's 8-column slice of the pivot row: from lane rgp of its half-warp
                    double p[LC], pr;
                    if (KS == 1 || kp == 0) {
#pragma unroll
                        for (int l = 0; l < LC; l++) p[l] = __shfl_sync(FULLMASK, w[0][l], rgp, 16);
                        pr = __shfl_sync(FULLMASK, rhs[0], rgp, 16);
                    } else if (KS == 2 || kp == 1) {
#pragma unroll
                        for (int l = 0; l < LC; l++) p[l] = __shfl_sync(FULLMASK, w[KS > 1 ? 1 : 0][l], rgp, 16);
                        pr = __shfl_sync(FULLMASK, rhs[KS > 1 ? 1 : 0], rgp, 16);
                    } else {
#pragma unroll
                        for (int l = 0; l < LC; l++) p[l] = __shfl_sync(FULLMASK, w[KS > 2 ? 2 : 0][l], rgp, 16);
                        pr = __shfl_sync(FULLMASK, rhs[KS > 2 ? 2 : 0], rgp, 16);
                    }
                    act &= ~(1ull << pinfo);
                    // the pivot row goes to scratch (row group 0 holds a copy of every slice)
                    if (rg == 0 && cgreal) {
#pragma unroll
                        for (int l2 = 0; l2 < LC / 2; l2++)
                            reinterpret_cast<double2 *>(urow + cg * LC)[l2] = make_double2(p[2 * l2], p[2 * l2 + 1]);
                        if (cg == 0) *reinterpret_cast<double2 *>(urow + CG * LC) = make_double2(pr, 0.0);   // rhs, pad
                    }
#pragma unroll
                    for (int k = 0; k < KS; k++) rhs[k] = fma(m[k], pr, rhs[k]);
                    if (cgj == CG - 1) {
                        // every row drops its leading entry: the current column is always entry 0
#pragma unroll
                        for (int k = 0; k < KS; k++) {
#pragma unroll
                            for (int l = 0; l + 1 < LC; l++) w[k][l] = fma(m[k], p[l + 1], w[k][l + 1]);
                            w[k][LC - 1] = 0.0;
                        }
                    } else {
#pragma unroll
                        for (int k = 0; k < KS; k++) {
#pragma unroll
                            for (int l = 0; l < LC; l++) w[k][l] = fma(m[k], p[l], w[k][l]);
                        }
                    }
                }
                cp_async_wait_all();       // record lc+2 has landed
                __syncthreads();
            }
            if (sing) status = SBD_BIN_SINGULAR;
        }
        cp_async_wait_all();
        __threadfence_block();
        __syncthreads();

        // ===================== phase 3: back substitution + fluxes ===========
        if (!status) {
            constexpr int kSlot = WL::ublk + WL::frec;
            auto fetch_layer = [&](int lyr, int buf) {
                double *dstp = wk + buf * kSlot;
                const double *s1 = ublk + (size_t)lyr * WL::ublk, *s2 = frecs + (size_t)lyr * WL::frec;
                for (int i = tid; i < WL::ublk / 2; i += NTHR) cp_async16(dstp + 2 * i, s1 + 2 * i);
                for (int i = tid; i < WL::frec / 2; i += NTHR) cp_async16(dstp + WL::ublk + 2 * i, s2 + 2 * i);
                cp_async_commit();
            };
            if (tid < N) xsm[tid] = 0.0;
            int lu_next = NT - 1;
            fetch_layer(ncut - 1, 0);
            for (int lc = ncut - 1; lc >= 0; lc--) {
                const int buf = (ncut - 1 - lc) & 1;
                if (lc > 0) { fetch_layer(lc - 1, buf ^ 1); cp_async_wait_one(); }
                else cp_async_wait_all();
                __syncthreads();
                if (warp == 0) {
                    const double *ubuf = wk + buf * kSlot;
                    const double *fr = ubuf + WL::ublk;
                    {
                        // stored row r: window column c sits at (c % CG) * LC + c / CG - r / CG
                        const int row = lane < N ? lane : 0;
                        const double *u = ubuf + row * US - row / CG;
                        const double dinv = fast_rcp(u[(row % CG) * LC + row / CG]);
                        double a4[4] = { ubuf[row * US + CG * LC], 0.0, 0.0, 0.0 };
#pragma unroll 4
                        for (int j = 0; j < N; j++)
                            a4[j & 3] = fma(-u[((N + j) % CG) * LC + (N + j) / CG], xsm[j], a4[j & 3]);
                        double acc = ((a4[0] + a4[1]) + (a4[2] + a4[3])) * dinv;
                        __syncwarp();
                        // x_c = acc of lane c; the chain per step is one shuffle + one FMA
#pragma unroll 4
                        for (int c = N - 1; c >= 0; c--) {
                            const double xc = __shfl_sync(FULLMASK, acc, c);
                            if (lane < c) acc = fma(-u[(c % CG) * LC + c / CG] * dinv, xc, acc);
                        }
                        if (lane < N) xsm[lane] = acc;
                        __syncwarp();
                    }
                    // ---- fluxes at the levels living in this layer (FLUXES, disort.f:1780) ----
                    if (fastmap)
                        while (lu_next >= 0 && layru[lu_next] > lc + 1) lu_next--;
                    for (int lu = fastmap ? lu_next : NT - 1; lu >= 0; lu--) {
                        if (layru[lu] != lc + 1) { if (fastmap) break; else continue; }
                        if (fastmap) lu_next = lu - 1;
                        const bool atbot = (a.d.ntau == 0 && lu == lc + 1);
                        const bool attop = (a.d.ntau == 0 && lu == lc);
                        const double *sc = fr + WL::f_sc;
                        double ut, utp, fact, edr;
                        if (atbot || attop) {
                            ut = tauc[lu]; utp = taucpr[lu]; fact = ebeam[lu]; edr = edir[lu];
                        } else {
                            ut = a.d.ntau > 0 ? a.utau[(size_t)src * NT + lu] : tauc[lu];
                            if (a.d.ntau > 0 && fabs(ut - tauc[L]) <= 1.e-4) ut = tauc[L];
                            utp = taucpr[lc] + sc[9] * (ut - tauc[lc]);
                            fact = fbeam > 0.0 ? exp(-utp / umu0) : 0.0;
                            edr = fbeam > 0.0 ? exp(-ut / umu0) : 0.0;
                        }
                        // S_t = sum_j cu[t][j] x_j f_j, t = up, down, mean: one mode per lane
                        double t0 = 0.0, t1 = 0.0, t2 = 0.0;
                        if (lane < N) {
                            const bool plusm = lane >= n;
                            const int jm = plusm ? lane - n : n - 1 - lane;
                            double fj;
                            if (atbot) fj = plusm ? fr[WL::f_ek + jm] : 1.0;
                            else if (attop) fj = plusm ? 1.0 : fr[WL::f_ek + jm];
                            else {
                                const double k = fr[WL::f_kk + jm];
                                const double d = plusm ? utp - taucpr[lc] : taucpr[lc + 1] - utp;
                                fj = exp(-k * d);
                            }
                            const double xf = xsm[lane] * fj;
                            t0 = fr[WL::f_cu + lane] * xf;
                            t1 = fr[WL::f_cu + N + lane] * xf;
                            t2 = fr[WL::f_cu + 2 * N + lane] * xf;
                        }
#pragma unroll
                        for (int o = 16; o > 0; o >>= 1) {
                            t0 += __shfl_xor_sync(FULLMASK, t0, o);
                            t1 += __shfl_xor_sync(FULLMASK, t1, o);
                            t2 += __shfl_xor_sync(FULLMASK, t2, o);
                        }
                        if (lane == 0) {
                            const double pi = kPiRef;
                            const double sup = t0 + sc[0] * fact + sc[3] + sc[7] * utp * Wq;
                            const double sdn = t1 + sc[1] * fact + sc[4] + sc[7] * utp * Wq;
                            const double sav = t2 + sc[2] * fact + sc[5] + sc[7] * utp * 2.0 * SWq;
                            const double dirint = fbeam * fact;
                            const double fldir = umu0 * (fbeam * fact);
                            const double rfldir = umu0 * fbeam * edr;
                            const double flup = 2. * pi * sup, fldn = 2. * pi * sdn;
                            const double fdntot = fldn + fldir;
                            constexpr double inv4pi = 1.0 / (4. * kPiRef);
                            const double uavg = (2. * pi * sav + dirint) * inv4pi;
                            const double plsorc = sc[6] + sc[7] * utp;
                            if (o_rfldir) o_rfldir[lu] = rfldir;
                            if (o_rfldn) o_rfldn[lu] = fdntot - rfldir;
                            if (o_flup) o_flup[lu] = flup;
                            if (o_uavg) o_uavg[lu] = uavg;
                            if (o_dfdt) o_dfdt[lu] = sc[8] * 4. * pi * (uavg - plsorc);
                        }
                    }
                }
                __syncthreads();     // everyone is done with this half of the double buffer
            }
        }
        cp_async_wait_all();
        if (tid == 0) a.status[bin] = status;
    }
}

// ---- host-side launch helpers ---------------------------------------------
bool wide_supported(int N) { return N == 20 || N == 24 || N == 32; }

template <int n>
static size_t wide_smem_t(int L, int NT) { return 8 * (WideLayout<n>::cta + WideLayout<n>::bin_doubles(L, NT)); }

size_t wide_smem_bytes(int N, int L, int NT)
{
    switch (N) {
    case 20: return wide_smem_t<10>(L, NT);
    case 24: return wide_smem_t<12>(L, NT);
    case 32: return wide_smem_t<16>(L, NT);
    }
    return 0;
}

size_t wide_slot_doubles(int N, int L)
{
    switch (N) {
    case 20: return WideLayout<10>::slot_doubles(L);
    case 24: return WideLayout<12>::slot_doubles(L);
    case 32: return WideLayout<16>::slot_doubles(L);
    }
    return 0;
}

template <int n>
static cudaError_t launch_wide_t(const LaunchArgs &a, int grid, cudaStream_t st)
{
    const int L = a.d.nlyr, NT = a.d.ntau > 0 ? a.d.ntau : L + 1;
    const size_t smem = wide_smem_t<n>(L, NT);
    cudaError_t e = cudaFuncSetAttribute(disort_wide_kernel<n>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    disort_wide_kernel<n><<<grid, WideLayout<n>::NTHR, smem, st>>>(a);
    return cudaGetLastError();
}

cudaError_t launch_wide(const LaunchArgs &a, int grid, cudaStream_t st)
{
    switch (a.d.nstr) {
    case 20: return launch_wide_t<10>(a, grid, st);
    case 24: return launch_wide_t<12>(a, grid, st);
    case 32: return launch_wide_t<16>(a, grid, st);
    }
    return cudaErrorInvalidValue;
}

}  // namespace sbd
