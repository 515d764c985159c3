// Register-resident warp-per-bin discrete-ordinate kernel for NSTR = 4, 8, 16.
//
// Same mathematics as sbd_generic.cu (see the header comment there) but laid
// out for the hardware.  The measured limiters of this kernel are the issue slots
// and the SM's shared-memory / shuffle pipe (MIO), not the FP64 pipe,
// so every phase is organised to move as few words between lanes as possible and
// the warps of a CTA walk through the phases together (one instruction stream in
// the I-cache at a time):
//   phase 1  per-layer eigen / particular solutions.  A warp works on 32/n
//            layers of its bin at once; each layer is owned by a group of n
//            lanes, lane j holding ROW j (Cholesky factors) or COLUMN j
//            (one-sided Jacobi) of the n x n matrices in registers, exchanged
//            with width-n shuffles.  The symmetric eigenproblem
//            T = L^T Pe~ L is solved as the SVD of A = K^T L (Pe~ = K K^T)
//            by one-sided (Hestenes) Jacobi with round-robin pairing: the
//            singular values are the DISORT eigenvalues k_j themselves.
//            Besides the layer record (eigenvectors, particular solutions)
//            the phase emits the FLUX FUNCTIONALS of the layer: the quadrature
//            sums of FLUXES (disort.f:1926-2006) applied to the eigenvectors,
//            so that a level flux is later a dot product with the solution.
//   phase 2  downward elimination of the block-bidiagonal boundary system with
//            partial pivoting.  The (n+N) x (2N+1) window is tiled 2-D over the
//            warp: 8 row groups x 4 column groups, each lane holding up to 3
//            rows x n columns (+ the right-hand sides).  A pivot step moves only
//            the lane's own column slice of the pivot row (one shuffle serves
//            all four column groups) and its three multipliers; rows never
//            travel through shared memory.
//   phase 3  upward back-substitution (row per lane) and the level fluxes as
//            dot products with the flux functionals.
// Layer records, flux records and pivot rows travel between the phases through
// a per-warp scratch slot in global memory, staged with cp.async.
#include <math.h>
#include <stdlib.h>

#include "sbd_internal.h"
#include "sbd_planck.cuh"
#include "sbd_devutil.cuh"
#include "sbd_addops.cuh"

// Staging of scratch data between the phases (global -> shared): bulk asynchronous copies
// (cp.async.bulk: the TMA engine, one instruction by one lane, completion on an mbarrier) or
// cp.async (LDGSTS, 16 bytes per lane and instruction).  Measured on the C2 bench (M bins/s,
// same box, two runs each): LDGSTS everywhere 4.20 / 4.21; bulk copies for the phase-3 pivot rows
// + flux records (4.9 KB per layer) 4.26 / 4.26; for the phase-2 record prefetch (1.4 KB per
// layer) 4.00 / 3.99; for both 4.20 / 4.19.  The phase-3 staging ships as bulk copies.
#ifndef SBD_TMA_P2
#define SBD_TMA_P2 0
#endif
#ifndef SBD_TMA_P3
#define SBD_TMA_P3 1
#endif
#define SBD_USE_TMA (SBD_TMA_P2 || SBD_TMA_P3)

namespace sbd {

#ifdef SBD_PHASE_TIMING
// debug build only: SM-clock ticks spent between the phase barriers, summed over CTAs
__device__ unsigned long long g_phase_ticks[8];
#define SBD_TICK(i)                                                                     \
    do {                                                                                \
        if (threadIdx.x == 0) {                                                         \
            const long long tnow = clock64();                                           \
            atomicAdd(&g_phase_ticks[i], (unsigned long long)(tnow - tphase));          \
            tphase = tnow;                                                              \
        }                                                                               \
    } while (0)
#else
#define SBD_TICK(i) do { } while (0)
#endif


template <int n>
struct FastLayout {
    static constexpr int N = 2 * n;
    // ---- per-layer record for assembling the boundary system (doubles) ----
    static constexpr int off_kk = 0;
    static constexpr int off_ek = n;
    static constexpr int off_gp = 2 * n;
    static constexpr int off_gm = 2 * n + n * n;
    static constexpr int off_zz = 2 * n + 2 * n * n;
    static constexpr int off_zp0 = off_zz + N;
    static constexpr int off_xr = off_zp0 + N;
    static constexpr int rec = ((off_xr + 2 + 1) / 2) * 2;
    // ---- per-layer flux record (phase 3) ----
    static constexpr int f_cu = 0;              // [3][N]: up, down, mean-intensity functionals
    static constexpr int f_ek = 3 * N;          // [n] exp(-k dtau')
    static constexpr int f_kk = 3 * N + n;      // [n] k
    static constexpr int f_sc = 3 * N + 2 * n;  // zzw[3], zp0w[3], xr0, xr1, 1-w, 1-w f
    static constexpr int frec = f_sc + 10;      // even
    // ---- 2-D tiling of the elimination window ----
    static constexpr int KS = (3 * n + 7) / 8;  // row slots per row group (8 groups)
    static constexpr int LC = n;                // window columns per lane (4 column groups)
    static constexpr int LH = n / 2;            // ... of which belong to the current layer
    static constexpr int US = 4 * LC + 2;       // stored pivot row: [cg][l], rhs, pad
    static constexpr int ublk = N * US;
    static constexpr unsigned slotmask = (KS * 8 >= 32) ? 0xffffffffu : ((1u << (KS * 8)) - 1u);
    __host__ __device__ static size_t slot_doubles(int L) { return (size_t)L * (rec + frec + ublk); }
    // ---- shared memory (doubles) ----
    static constexpr int cta = 4 * n + N * n + 2;   // cmu cwt csq cdinv, ylm, sum(w mu), sum(w)
    // radiance runs: Y_l^m at the user cosines of the current azimuth mode, CTA-shared
    __host__ __device__ static size_t cta_doubles(int NU) { return (size_t)cta + (((size_t)N * NU + 1) & ~(size_t)1); }
    // phase 1: a layer is owned by a group of GW lanes (the n first ones hold a row / column /
    // mode each; for n = 10, 12 the rest shadow lane n-1, compute along and store nothing)
    // (n = 10: three groups of 10 lanes, lanes 30 and 31 shadow lane 29; shuffles then name the
    // source lane in the warp)
    static constexpr int GW = n <= 2 ? 2 : (n <= 4 ? 4 : (n <= 8 ? 8 : (n == 10 ? 10 : 16)));
    static constexpr int tasks = 32 / GW;
    // gl, K, L, G1, G2 [n][LD], vectors.  Rows n + 2 doubles apart (row-wise accesses of a layer group's
    // lanes hit different banks), areas padded to 4 (mod 16) doubles (the groups' broadcast loads too)
    static constexpr int LD = n + 2;
    static constexpr int task = ((N + 4 * n * LD + 4 * n + 11) / 16) * 16 + 4;
    // work area shared by the phases: per-task areas (phase 1), 3 records + assembled rows (phase 2),
    // 2 x (pivot rows + flux record) (phase 3)
    static constexpr int cmax(int a, int b) { return a > b ? a : b; }
    static constexpr int work = cmax(cmax(cmax(tasks * task, 3 * rec + ublk), 2 * (ublk + frec)), 3 * SBD_MAX_NLYR);
    // radiance runs also stage the layer record in phase 3 (eigenvectors at the user angles); the
    // adding form has no pivot rows in its phase-3 buffers
    __host__ __device__ static constexpr int stage3(bool add) { return 2 * ((add ? 0 : ublk) + frec + rec); }
    static constexpr int ecols = N + 3;            // eigen-terms + beam, Planck Z0, Z1 sources
    // user-angle work values: E[N][ecols], GU[NU][ecols], running intensity [NU], cos(m dphi) [NPHI], g_l [N],
    // layer solution [N]
    __host__ __device__ static size_t rad_doubles(int NU, int NPHI)
    {
        size_t d = (size_t)N * ecols + (size_t)NU * ecols + NU + NPHI + 2 * N;
        return (d + 1) & ~(size_t)1;
    }
    // work area of a radiance run: the phase-1 task areas (or the sweeps' matrices), later the
    // phase-3 buffers with the user-angle work values behind them
    __host__ __device__ static size_t work_rad(int NU, int NPHI, bool add)
    {
        const size_t base = n > 8 ? (size_t)cmax(tasks * task, AddOps<n>::p2) : (size_t)work;
        const size_t p3 = (size_t)stage3(add) + rad_doubles(NU, NPHI);
        return ((base > p3 ? base : p3) + 1) & ~(size_t)1;
    }
    __host__ __device__ static size_t warp_doubles(int L, int NT, int NU = 0, int NPHI = 0, bool add = false)
    {
        // y0, work area (also the 3 L prologue work values), taucpr/tauc, beam transmissions (2),
        // pk(+2 boundary temps), level map; kept even for 16-byte alignment
        // (+ three mbarriers of the bulk-copy staging and a spare)
        size_t d = (size_t)N + (NU > 0 ? work_rad(NU, NPHI, add) : (size_t)work) + 4 * (L + 1) + (L + 3) + (NT + 1) / 2 + 4;
        return (d + 1) & ~(size_t)1;
    }
    // global scratch per warp: records, flux records, pivot rows (+ the downward-intensity
    // source and transmission of every layer for the top-down pass of radiance runs)
    __host__ __device__ static size_t slot_doubles_rad(int L, int NU) { return slot_doubles(L) + (size_t)2 * L * NU; }
};

// sum over the GW lanes of a layer group (shadow lanes pass zero); g: lane in group, gbase: first lane
template <int GW>
__device__ __forceinline__ double group_sum(double v, int g, int gbase)
{
    if constexpr ((GW & (GW - 1)) == 0) {
#pragma unroll
        for (int o = GW / 2; o > 0; o >>= 1) v += __shfl_xor_sync(FULLMASK, v, o, GW);
        return v;
    } else {
        static_assert(GW == 10, "group of 10 lanes");
        const double t = __shfl_down_sync(FULLMASK, v, 8);
        if (g < 2) v += t;
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) v += __shfl_down_sync(FULLMASK, v, o);
        return __shfl_sync(FULLMASK, v, gbase);
    }
}



// Round-robin (tournament) pairing of the one-sided Jacobi sweeps: the partner of
// lane g in round r, 3 bits per round (4 above n = 8).
template <int n>
__device__ __forceinline__ unsigned long long jacobi_partners(int g)
{
    constexpr int PB = n > 8 ? 4 : 3;
    unsigned long long pk = 0;
#pragma unroll
    for (int r = 0; r < n - 1; r++) {
        int partner;
        if (g == n - 1) partner = r;
        else if (g == r) partner = n - 1;
        else {
            partner = 2 * r - g + (n - 1);
            if (partner >= n - 1) partner -= n - 1;
            if (partner >= n - 1) partner -= n - 1;
        }
        pk |= (unsigned long long)partner << (PB * r);
    }
    return pk;
}

// ---------------------------------------------------------------------------
// phase 1: one layer per group of n lanes
// ---------------------------------------------------------------------------
template <int n, bool ADDREC>
__device__ __forceinline__ int phase1_layers(
    const double *__restrict__ dtauc, const double *__restrict__ ssalb,
    const double *__restrict__ pmom, int ldp, int lc, bool active, int mazim,
    double fbeam, double umu0, bool plank, double delm0,
    const double *cmu, const double *cwt, const double *csq, const double *cdinv, const double *cylm,
    const double *y0, const double *taucpr, const double *pk,
    double *tsm /* per-task shared */, double *rec /* scratch record of this layer */,
    double *frec /* flux record of this layer */, int g /* lane in group (shadow lanes: n-1) */,
    bool gact /* false: shadow lane */, int gbase /* first lane of the group */,
    unsigned long long jpart /* Jacobi partners of this lane, see jacobi_partners */,
    double *arec = nullptr /* ADDREC: the layer's R, T, s_up, s_dn */, const double *ebeam = nullptr)
{
    using FL = FastLayout<n>;
    constexpr int N = 2 * n, LD = FL::LD, GW = FL::GW, PB = n > 8 ? 4 : 3;
    double *sgl = tsm, *sK = sgl + N, *sL = sK + n * LD, *sG1 = sL + n * LD, *sG2 = sG1 + n * LD,
           *sv = sG2 + n * LD;   // sv: 4 vectors of n

    double ss = ssalb[lc];
    if (ss == 1.0) ss = 1.0 - kDither;
    double dt = dtauc[lc];
    if (dt < 0.0) dt = 0.0;
    const double f = pmom[(size_t)lc * ldp + N];
    const double oprim = ss * (1. - f) * fast_rcp(1. - f * ss);
    const double dtaucp = (1. - f * ss) * dt;
    const double rf = fast_rcp(1. - f);
#pragma unroll
    for (int h = 0; h < 2; h++) {
        int l = g + h * n;
        double pm = (l == 0) ? 1.0 : pmom[(size_t)lc * ldp + l];
        if (gact) sgl[l] = (2 * l + 1) * oprim * (pm - f) * rf;
    }
    __syncwarp();

    // rows g of Pe~ and Po~
    double pe[n], po[n];
#pragma unroll
    for (int j = 0; j < n; j++) { pe[j] = 0.0; po[j] = 0.0; }
#pragma unroll
    for (int l = 0; l < N; l++) {
        if (l >= mazim) {
            const double t = sgl[l] * cylm[l * n + g];
            if ((l - mazim) & 1) {
#pragma unroll
                for (int j = 0; j < n; j++) po[j] = fma(t, cylm[l * n + j], po[j]);
            } else {
#pragma unroll
                for (int j = 0; j < n; j++) pe[j] = fma(t, cylm[l * n + j], pe[j]);
            }
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
    double rK[n], rL[n];     // reciprocal diagonals of K and L (uniform in the group)
#pragma unroll
    for (int j = 0; j < n; j++) {
        double nume = pe[j], numo = po[j];
#pragma unroll
        for (int k = 0; k < j; k++) {
            nume = fma(-pe[k], group_get<GW>(pe[k], j, gbase), nume);
            numo = fma(-po[k], group_get<GW>(po[k], j, gbase), numo);
        }
        double pive = group_get<GW>(nume, j, gbase), pivo = group_get<GW>(numo, j, gbase);
        if (!(pivo > 0.0)) { bad = 1; pivo = 1.0; }
        // Pe~ is only semidefinite when w' -> 1: keep the factor real (a NaN pivot is a failure)
        const double floor_e = 1.0e-30;
        if (pive != pive) bad = 1;
        if (!(pive > floor_e)) pive = floor_e;
        const double rie = fast_rsqrt(pive), rio = fast_rsqrt(pivo);
        rK[j] = rie; rL[j] = rio;
        pe[j] = (g == j) ? pive * rie : ((g > j) ? nume * rie : 0.0);
        po[j] = (g == j) ? pivo * rio : ((g > j) ? numo * rio : 0.0);
    }
#pragma unroll
    for (int j = 0; j < n; j++) if (gact) { sK[g * LD + j] = pe[j]; sL[g * LD + j] = po[j]; }
    __syncwarp();

    // column g of A = K^T L
    double a[n];
    {
        double lcol[n];
#pragma unroll
        for (int k = 0; k < n; k++) lcol[k] = sL[k * LD + g];
#pragma unroll
        for (int i = 0; i < n; i++) {
            double acc = 0.0;
#pragma unroll
            for (int k = i; k < n; k++) acc = fma(sK[k * LD + i], lcol[k], acc);
            a[i] = acc;
        }
    }

    // One-sided Jacobi, round-robin pairing; partner columns come by shuffle.
    // Only A is rotated: with A V = U S, the vectors needed downstream are
    // P = L V = K^-T (A V) and Q = L^-T V = (L L^T)^-1 P, so V is never formed.
    if (n > 1) {
        // squared column norm, carried along analytically (a'_pp = a_pp - t a_pq,
        // a'_qq = a_qq + t a_pq); only used for the rotation angles -- the singular
        // value itself is recomputed from the column after convergence
        double own2 = 0.0;
#pragma unroll
        for (int i = 0; i < n; i++) own2 = fma(a[i], a[i], own2);
        for (int sweep = 0; sweep < 40; sweep++) {
            // `big`: some pair of this sweep was still more than 1e-5 away from
            // orthogonal.  The cyclic method converges quadratically, so a sweep
            // without such a pair leaves residual cosines ~1e-10 and no checking
            // sweep is needed.
            int big = 0;
#pragma unroll 1
            for (int r = 0; r < n - 1; r++) {     // not unrolled: keeps the code in the I-cache
                const int partner = (int)(jpart >> (PB * r)) & ((1 << PB) - 1);
                double pa[n];
                double g0 = 0.0, g1 = 0.0;
                const double oth2 = group_get<GW>(own2, partner, gbase);
#pragma unroll
                for (int i = 0; i < n; i++) {
                    pa[i] = group_get<GW>(a[i], partner, gbase);
                    if (i & 1) g1 = fma(a[i], pa[i], g1); else g0 = fma(a[i], pa[i], g0);
                }
                const double gam = g0 + g1;
                const bool lo = g < partner;
                const double gg = gam * gam, ab = own2 * oth2;
                // rotate unless the pair is orthogonal to 1e-12 (eigenvectors then carry
                // errors ~1e-12, far below the 1e-5 target)
                if (gg > 1.0e-24 * ab && gg > 1.0e-290) {
                    if (gg > kJacobiBig * ab) big = 1;
                    // rotation by theta, |theta| <= pi/4: cos 2theta = |dl| / h, sin 2theta = gam / h
                    // with dl = (beta - alpha) / 2, h = sqrt(dl^2 + gam^2); no division:
                    // c = sqrt(x), x = (1 + cos 2theta) / 2, s = sin 2theta / (2 c), t = s / c
                    const double dl = lo ? 0.5 * (oth2 - own2) : 0.5 * (own2 - oth2);
                    const double rh = fast_rsqrt(fma(dl, dl, gg));
                    const double x = fma(0.5 * fabs(dl), rh, 0.5);
                    const double rc = fast_rsqrt(x);
                    const double cc = x * rc;
                    double sn = gam * (0.5 * rh) * rc;       // sign of gam
                    if ((dl < 0.0) != lo) sn = -sn;          // sign(dl), and the low column takes -s
                    // squared norms follow analytically: alpha' = alpha - t gam, beta' = beta + t gam
                    own2 = fma(sn * rc, gam, own2);
#pragma unroll
                    for (int i = 0; i < n; i++) a[i] = fma(sn, pa[i], cc * a[i]);
                }
            }
            if (!__any_sync(FULLMASK, big)) break;
            // refresh the carried norms once per sweep (cheap, stops drift)
            own2 = 0.0;
#pragma unroll
            for (int i = 0; i < n; i++) own2 = fma(a[i], a[i], own2);
        }
    }

    // singular value = DISORT eigenvalue k (disort.f:3264-3269)
    double s2 = 0.0;
#pragma unroll
    for (int i = 0; i < n; i++) s2 = fma(a[i], a[i], s2);
    const double kk = sqrt(s2);
    const double ek = exp(-kk * dtaucp);
    // column g of P = K^-T a, then Q = L^-T (L^-1 P)
    double P[n], Q[n];
#pragma unroll
    for (int i = n - 1; i >= 0; i--) {
        double acc = a[i];
#pragma unroll
        for (int k = i + 1; k < n; k++) acc = fma(-sK[k * LD + i], P[k], acc);
        P[i] = acc * rK[i];
    }
    {
        double y[n];
#pragma unroll
        for (int i = 0; i < n; i++) {
            double acc = P[i];
#pragma unroll
            for (int k = 0; k < i; k++) acc = fma(-sL[i * LD + k], y[k], acc);
            y[i] = acc * rL[i];
        }
#pragma unroll
        for (int i = n - 1; i >= 0; i--) {
            double acc = y[i];
#pragma unroll
            for (int k = i + 1; k < n; k++) acc = fma(-sL[k * LD + i], Q[k], acc);
            Q[i] = acc * rL[i];
        }
    }
    // G+ - G- = D^-1 Q ; G+ + G- = -D^-1 P / k   (disort.f:3273-3301)
    const double rk = fast_rcp(kk);
    double gs[n], gd[n];
    // flux functionals of mode g (FLUXES, disort.f:1926-2006): quadrature sums of the
    // eigenvector over the upward / downward hemispheres and over all directions
    double fA = 0.0, fB = 0.0, fC = 0.0;
#pragma unroll
    for (int i = 0; i < n; i++) {
        gd[i] = cdinv[i] * Q[i];
        gs[i] = -cdinv[i] * P[i] * rk;
        if (gact) {
            sG1[g * LD + i] = gs[i];      // [mode j][direction i]
            sG2[g * LD + i] = gd[i];
        }
        const double gpi = 0.5 * (gs[i] + gd[i]), gmi = 0.5 * (gs[i] - gd[i]);
        const double wm = cwt[i] * cmu[i];
        fA = fma(wm, gpi, fA);
        fB = fma(wm, gmi, fB);
        fC = fma(cwt[i], gs[i], fC);
        if (active && gact) {
            rec[FL::off_gp + i * n + g] = gpi;
            rec[FL::off_gm + i * n + g] = gmi;
        }
    }
    if (active && gact) {
        rec[FL::off_kk + g] = kk;
        rec[FL::off_ek + g] = ek;
        // solution column n+g belongs to +k_g, column n-1-g to -k_g (disort.f:3264-3312)
        frec[FL::f_cu + n + g] = fA;          frec[FL::f_cu + n - 1 - g] = -fB;
        frec[FL::f_cu + N + n + g] = fB;      frec[FL::f_cu + N + n - 1 - g] = -fA;
        frec[FL::f_cu + 2 * N + n + g] = fC;  frec[FL::f_cu + 2 * N + n - 1 - g] = -fC;
        frec[FL::f_ek + g] = ek;
        frec[FL::f_kk + g] = kk;
    }

    // ---- beam particular solution: spectral form of UPBEAM (disort.f:4130) ----
    double zup = 0.0, zdn = 0.0;
    if (fbeam > 0.0) {
        const double fac = (2. - delm0) * fbeam * (1.0 / (4. * kPiRef));
        const double rmu0 = fast_rcp(umu0);
        double be = 0.0, bo = 0.0;
#pragma unroll
        for (int l = 0; l < N; l++) {
            if (l >= mazim) {
                const double t = sgl[l] * cylm[l * n + g] * y0[l];
                if ((l - mazim) & 1) bo += t; else be += t;
            }
        }
        const double bs = 2.0 * fac * sqg * be, bd = 2.0 * fac * sqg * bo;
        if (gact) sv[g] = bd;
        __syncwarp();
        double t1 = 0.0;                                    // (K^T b^_d)_g
#pragma unroll
        for (int k = 0; k < n; k++) t1 = fma(sK[k * LD + g], sv[k], t1);
        if (gact) sv[n + g] = t1;
        __syncwarp();
        double t2 = 0.0;                                    // (K K^T b^_d)_g
#pragma unroll
        for (int k = 0; k < n; k++) t2 = fma(sK[g * LD + k], sv[n + k], t2);
        if (gact) sv[2 * n + g] = bs * rmu0 - t2;           // r_g
        __syncwarp();
        double cj = 0.0;                                    // (P^T r)_g / (1/mu0^2 - k^2)
#pragma unroll
        for (int i = 0; i < n; i++) cj = fma(P[i], sv[2 * n + i], cj);
        cj = cj * fast_rcp(rmu0 * rmu0 - s2);
        if (gact) { sv[3 * n + g] = cj; sv[g] = cj * kk; }
        __syncwarp();
        // d = sum_j Gd(:,j) c_j ; s = mu0 (D^-1 b^_d + sum_j Gs(:,j) k_j c_j), direction i = g
        double dv = 0.0, sv2 = 0.0;
#pragma unroll
        for (int j = 0; j < n; j++) {
            dv = fma(sG2[j * LD + g], sv[3 * n + j], dv);
            sv2 = fma(sG1[j * LD + g], sv[j], sv2);
        }
        sv2 = umu0 * (cdinv[g] * bd + sv2);
        zup = 0.5 * (sv2 + dv);
        zdn = 0.5 * (sv2 - dv);
        __syncwarp();
    }
    // ---- thermal particular solution (UPISOT, disort.f:4247) -----------------
    double xr0 = 0.0, xr1 = 0.0, q = 0.0;
    if (plank && mazim == 0) {
        if (dtaucp > 1.0e-200) xr1 = (pk[lc + 1] - pk[lc]) * fast_rcp(dtaucp);
        else if (dtaucp > 0.0) xr1 = (pk[lc + 1] - pk[lc]) / dtaucp;
        xr0 = pk[lc] - xr1 * taucpr[lc];
        double y[n], z[n];
#pragma unroll
        for (int i = 0; i < n; i++) {
            double acc = cmu[i] * csq[i];      // D_i = sqrt(w mu) = mu sqrt(w/mu)
#pragma unroll
            for (int k = 0; k < i; k++) acc = fma(-sL[i * LD + k], y[k], acc);
            y[i] = acc * rL[i];
        }
#pragma unroll
        for (int i = n - 1; i >= 0; i--) {
            double acc = y[i];
#pragma unroll
            for (int k = i + 1; k < n; k++) acc = fma(-sL[k * LD + i], z[k], acc);
            z[i] = acc * rL[i];
            if (i == g) q = cdinv[i] * z[i];
        }
    }
    // quadrature sums of the particular solutions (same functionals as above)
    {
        const double wmg = gact ? cwt[g] * cmu[g] : 0.0, wg = gact ? cwt[g] : 0.0;
        const double Zu = group_sum<GW>(wmg * zup, g, gbase);
        const double Zd = group_sum<GW>(wmg * zdn, g, gbase);
        const double Za = group_sum<GW>(wg * (zup + zdn), g, gbase);
        const double Q1 = (plank && mazim == 0) ? group_sum<GW>(wmg * q, g, gbase) : 0.0;
        if (active && g == 0) {
            const double W = cylm[-2], SW = cylm[-1];      // sum(w mu), sum(w): see the kernel
            double *sc = frec + FL::f_sc;
            sc[0] = Zu; sc[1] = Zd; sc[2] = Za;
            sc[3] = fma(xr1, Q1, xr0 * W); sc[4] = fma(-xr1, Q1, xr0 * W); sc[5] = 2.0 * xr0 * SW;
            sc[6] = xr0; sc[7] = xr1; sc[8] = 1.0 - ss; sc[9] = 1.0 - ss * f;
        }
    }
    if (active && gact) {
        rec[FL::off_zz + n + g] = zup;
        rec[FL::off_zz + n - 1 - g] = zdn;
        rec[FL::off_zp0 + n + g] = xr0 + xr1 * q;
        rec[FL::off_zp0 + n - 1 - g] = xr0 - xr1 * q;
        if (g == 0) { rec[FL::off_xr] = xr0; rec[FL::off_xr + 1] = xr1; }
    }
    if (ADDREC) {
        // the layer's reflection / transmission operators and sources for the adding sweeps
        // (scaled variables u^ = D u, D = sqrt(w mu) = 1 / cdinv)
        const double Dg = cmu[g] * csq[g];
        const bool therm = plank && mazim == 0;
        bad |= layer_operators<n, LD>(P, kk, dtaucp, zup * Dg, zdn * Dg, q * Dg, xr1,
                                      therm ? pk[lc] : 0.0, therm ? pk[lc + 1] : 0.0,
                                      fbeam > 0.0 ? ebeam[lc] : 0.0, fbeam > 0.0 ? ebeam[lc + 1] : 0.0,
                                      fbeam > 0.0, therm, cmu, csq, sG1, sK, sG2, sv, arec, active && gact, g, gact);
    }
    __syncwarp();
    return (bad && active) ? SBD_BIN_EIG_FAIL : 0;
}

// ---------------------------------------------------------------------------
// phase 2 helpers
// ---------------------------------------------------------------------------
// Rows of the boundary system are assembled by the whole warp into a staging
// area in shared memory, one window column per lane (2N divides 32), in the
// layout the pivot rows are stored in: row r at r*US, window column c at
// (c & 3) * LC + (c >> 2), right-hand side at 4*LC.
//
// Entry (r, c) of a layer's eigenvector matrix GC times the exponential factors
// of the layer bottom or top (disort.f:2846-2882): column n+j belongs to +k_j,
// column n-1-j to -k_j; rows r >= n are the upward directions.  recA/bottomA
// describe the block of the current layer (columns < N), recB (top of the next
// layer, negated; may be null) the block of columns >= N.
// refl != 0: bottom boundary rows r0.. with the Lambertian reflection of the
// downward directions folded in (disort.f:2919-2990).
template <int n>
__device__ __forceinline__ void stage_rows(double *stg, const double *recA, bool bottomA,
                                           const double *recB, int r0, int nrows, double refl,
                                           const double *cwt, const double *cmu, int lane)
{
    using FL = FastLayout<n>;
    constexpr int N = 2 * n, LC = FL::LC, US = FL::US;
    const int c = lane % (2 * N);
    const bool isB = c >= N;
    const int cc = isB ? c - N : c;
    const bool plus = cc >= n;
    const int j = plus ? cc - n : n - 1 - cc;
    const double *rec = (isB && recB) ? recB : recA;
    const bool bottom = isB ? false : bottomA;
    double fac = (plus == bottom) ? rec[FL::off_ek + j] : 1.0;
    if (!plus) fac = -fac;
    if (isB) fac = recB ? -fac : 0.0;
    // reflected part: sum over the downward directions (rows n-1-k, i = k) of this column
    double rsum = 0.0;
    if (refl != 0.0) {
        const double *gdn = rec + (plus ? FL::off_gm : FL::off_gp) + j;
#pragma unroll 1
        for (int k = 0; k < n; k++) rsum = fma(cwt[k] * cmu[k], gdn[k * n], rsum);
    }
    const int pos = (c & 3) * LC + (c >> 2);
    constexpr int rstep = 32 / (2 * N);
    // rows r < n are the downward directions i = n-1-r, rows r >= n the upward ones
    // i = r-n; for a "+k" column the upward rows read G+ and the downward rows G-
    const double *gup = rec + (plus ? FL::off_gp : FL::off_gm) + j;
    const double *gdw = rec + (plus ? FL::off_gm : FL::off_gp) + j;
    const double rfl = refl * rsum;
#pragma unroll 1
    for (int rr = lane / (2 * N); rr < nrows; rr += rstep) {
        const int r = r0 + rr;
        const double v = r >= n ? gup[(r - n) * n] : gdw[(n - 1 - r) * n];
        stg[rr * US + pos] = (v - rfl) * fac;
    }
}

// One pivot step of the 2-D tiled elimination on the leading column of the
// lanes' column slices; cgj = column group that owns it.  The pivot row is
// written to `urow` (global scratch) on the way.  Returns true when no usable
// pivot exists.
// W = slice entries that can still be non-zero: LC, fewer once the slices have slid.
template <int n, int W>
__device__ __forceinline__ bool elim_step(double (&w)[FastLayout<n>::KS][FastLayout<n>::LC],
                                          double (&rhs)[FastLayout<n>::KS], unsigned &act,
                                          double *uslice /* urow + cg*LC */, int cgj, int rg, int cg)
{
    constexpr int KS = FastLayout<n>::KS, LC = FastLayout<n>::LC;
    // bit 8k: slot k of this row group holds a live equation (pivot candidates: the
    // owner lanes' live rows)
    const unsigned cand = (cg == cgj) ? ((act >> rg) & 0x010101u) : 0u;
    // column-j entries of this row group's rows, from the group's owner lane
    double colj[KS];
#pragma unroll
    for (int k = 0; k < KS; k++) colj[k] = __shfl_sync(FULLMASK, w[k][0], (rg << 2) | cgj);
    // pivot: largest |a| by high word (any near-maximal pivot is as stable); the two
    // low bits carry the row slot so that the reduction returns it as well
    int best = -1;
    double bval = 1.0;
#pragma unroll
    for (int k = 0; k < KS; k++) {
        const int h = ((cand >> (8 * k)) & 1u) ? ((__double2hiint(w[k][0]) & 0x7ffffffc) | (KS - 1 - k)) : -1;
        if (h > best) { best = h; bval = w[k][0]; }
    }
    const double rloc = fast_rcp(bval);          // speculative: off the critical path
    const int mx = __reduce_max_sync(FULLMASK, best);
    const unsigned who = __ballot_sync(FULLMASK, best == mx);
    if ((mx >> 2) <= 0) return true;
    const int pl = __ffs(who) - 1, kp = KS - 1 - (mx & 3), rgp = pl >> 2;
    const double rp = -__shfl_sync(FULLMASK, rloc, pl);
    // Every slot is updated, dead ones included: a row that served as a pivot
    // annihilates itself to rounding level (multiplier -1) and only carries finite
    // residue afterwards; it is never a pivot candidate again and is overwritten
    // when the slot receives a new equation.
    double m[KS];
#pragma unroll
    for (int k = 0; k < KS; k++) m[k] = colj[k] * rp;
    // this lane's column slice of the pivot row: one shuffle serves all 4 column groups
    double p[W], pr;
    const int src = (rgp << 2) | cg;
    if (KS == 1 || kp == 0) {
#pragma unroll
        for (int l = 0; l < W; l++) p[l] = __shfl_sync(FULLMASK, w[0][l], src);
        pr = __shfl_sync(FULLMASK, rhs[0], src);
    } else if (KS == 2 || kp == 1) {
#pragma unroll
        for (int l = 0; l < W; l++) p[l] = __shfl_sync(FULLMASK, w[KS > 1 ? 1 : 0][l], src);
        pr = __shfl_sync(FULLMASK, rhs[KS > 1 ? 1 : 0], src);
    } else {
#pragma unroll
        for (int l = 0; l < W; l++) p[l] = __shfl_sync(FULLMASK, w[KS > 2 ? 2 : 0][l], src);
        pr = __shfl_sync(FULLMASK, rhs[KS > 2 ? 2 : 0], src);
    }
    act &= ~(1u << (kp * 8 + rgp));
    // the pivot row goes to scratch (row group 0 holds a copy of every slice)
    if (rg == 0) {
#pragma unroll
        for (int l2 = 0; l2 < W / 2; l2++)
            reinterpret_cast<double2 *>(uslice)[l2] = make_double2(p[2 * l2], p[2 * l2 + 1]);
        if (cg == 0) *reinterpret_cast<double2 *>(uslice + 4 * LC) = make_double2(pr, 0.0);   // rhs, pad
    }
    // update; after the last column group of a slice position every row drops its
    // leading entry, so the current column is always entry 0 of the slice (dead
    // slots slide along: their contents are never used again)
#pragma unroll
    for (int k = 0; k < KS; k++) rhs[k] = fma(m[k], pr, rhs[k]);
    if (cgj == 3) {
#pragma unroll
        for (int k = 0; k < KS; k++) {
#pragma unroll
            for (int l = 0; l + 1 < W; l++) w[k][l] = fma(m[k], p[l + 1], w[k][l + 1]);
            w[k][W - 1] = 0.0;
        }
    } else {
#pragma unroll
        for (int k = 0; k < KS; k++) {
#pragma unroll
            for (int l = 0; l < W; l++) w[k][l] = fma(m[k], p[l], w[k][l]);
        }
    }
    return false;
}

// ---------------------------------------------------------------------------
// radiance runs (RAD): intensities at the user angles
// ---------------------------------------------------------------------------
// Y_l^m(x), l = 0..N-1 (zero below l = m), the functions LEPOLY builds mode by mode
// (disort.f:5286-5408); one lane, N <= 16
__device__ __forceinline__ void lepoly_mode(int m, int N, double x, double *y)
{
    if (m == 0) {
        y[0] = 1.0; y[1] = x;
        for (int l = 2; l < N; l++) y[l] = ((2 * l - 1) * x * y[l - 1] - (l - 1) * y[l - 2]) / l;
        return;
    }
    double d = 1.0;
    const double s = sqrt(1.0 - x * x);
    for (int k = 1; k <= m; k++) d = -sqrt((double)(2 * k - 1)) / sqrt((double)(2 * k)) * s * d;
    for (int l = 0; l < m && l < N; l++) y[l] = 0.0;
    if (m < N) y[m] = d;
    if (m + 1 < N) y[m + 1] = sqrt((double)(2 * m + 1)) * x * d;
    for (int l = m + 2; l < N; l++) {
        const double t1 = sqrt((double)(l - m)) * sqrt((double)(l + m));
        const double t2 = sqrt((double)(l - m - 1)) * sqrt((double)(l + m - 1));
        y[l] = ((2 * l - 1) * x * y[l - 1] - t2 * y[l - 2]) / t1;
    }
}

// User-angle terms of one layer for azimuth mode m (TERPEV disort.f:3920, TERPSO :3980):
// the eigenvectors times the solution coefficients, the beam and the Planck particular
// solutions re-expanded at the user cosines through the Legendre sum,
//   GU[iu][c] = sum_l 1/2 g_l [sum_i w_i Y_l^m(mu_i) V_c(mu_i)] Y_l^m(umu_iu)      (V_c over all N directions)
//               c < N: eigenvector c times x_c; c = N beam (+ the direct term), N+1 / N+2 Planck Z0 / Z1.
// With Y_l^m(-mu) = (-1)^(l-m) Y_l^m(mu) the sum over the N directions of the eigenvector pair +-k_j
// needs only F[l][j] = sum_i w_i Y_l^m(mu_i) (G+ + G-)(i,j) for even l-m, (G+ - G-)(i,j) for odd l-m:
//   column n+j:   x (He + Ho),   column n-1-j:   x (Ho - He),   He / Ho = sum over even / odd l-m of
//   1/2 g_l F[l][j] Y_l^m(umu)  -- half the products of the straightforward form, and the solution
// enters at the end.
// rec: the layer record (shared); sx: the layer's solution (shared); E: work area, N x ecols.
template <int n>
__device__ __forceinline__ void user_terms_fast(
    const double *rec, const double *sx, const double *gl, const double *cwt,
    const double *cylm, const double *y0, const double *ylmu_m, int NU, int mazim,
    bool beam, double fact, bool therm, double oprim, double *E, double *GU, int lane)
{
    using FL = FastLayout<n>;
    constexpr int N = 2 * n, EC = FL::ecols;
    constexpr int JW = (n % 4 == 0) ? 4 : 2, JG = n / JW;      // columns per work item
    const double *gp = rec + FL::off_gp, *gm = rec + FL::off_gm;
    const double *zz = rec + FL::off_zz, *zp0 = rec + FL::off_zp0;
    const double xr0 = rec[FL::off_xr], xr1 = rec[FL::off_xr + 1];
    double *Fs = E, *Eb = E + N * n;         // [N][n] scaled F, [N][3] beam / Planck columns
    for (int it = lane; it < N * JG; it += 32) {
        const int l = it / JG, jq = (it - l * JG) * JW;
        double acc[JW];
#pragma unroll
        for (int q = 0; q < JW; q++) acc[q] = 0.0;
        if (l >= mazim) {
            const bool odd = (l - mazim) & 1;
            const double *yl = cylm + l * n;
#pragma unroll
            for (int i = 0; i < n; i++) {
                const double wy = cwt[i] * yl[i];
                const double2 *a2 = reinterpret_cast<const double2 *>(gp + i * n + jq);
                const double2 *b2 = reinterpret_cast<const double2 *>(gm + i * n + jq);
#pragma unroll
                for (int q2 = 0; q2 < JW / 2; q2++) {
                    const double2 a = a2[q2], b = b2[q2];
                    acc[2 * q2] = fma(wy, odd ? a.x - b.x : a.x + b.x, acc[2 * q2]);
                    acc[2 * q2 + 1] = fma(wy, odd ? a.y - b.y : a.y + b.y, acc[2 * q2 + 1]);
                }
            }
            const double h = 0.5 * gl[l];
#pragma unroll
            for (int q = 0; q < JW; q++) acc[q] *= h;
        }
#pragma unroll
        for (int q = 0; q < JW; q++) Fs[l * n + jq + q] = acc[q];
    }
    for (int e = lane; e < N * 3; e += 32) {
        const int l = e / 3, cc = e - l * 3;
        double acc = 0.0;
        if (l >= mazim) {
            const double sg = ((l - mazim) & 1) ? -1.0 : 1.0;
            const double *yl = cylm + l * n;
            if (cc == 0) {
                if (beam) {
#pragma unroll
                    for (int i = 0; i < n; i++) acc = fma(cwt[i] * yl[i], fma(sg, zz[n - 1 - i], zz[n + i]), acc);
                    acc = 0.5 * gl[l] * acc + fact * gl[l] * y0[l];
                }
            } else if (therm) {
                if (cc == 1) {
#pragma unroll
                    for (int i = 0; i < n; i++) acc = fma(cwt[i] * yl[i], fma(sg, zp0[n - 1 - i], zp0[n + i]), acc);
                } else {
#pragma unroll
                    for (int i = 0; i < n; i++) acc = fma(cwt[i] * yl[i], xr1 + sg * xr1, acc);
                }
                acc *= 0.5 * gl[l];
            }
        }
        Eb[e] = acc;
    }
    __syncwarp();
    for (int it = lane; it < NU * JG; it += 32) {
        const int iu = it / JG, jq = (it - iu * JG) * JW;
        double he[JW], ho[JW];
#pragma unroll
        for (int q = 0; q < JW; q++) { he[q] = 0.0; ho[q] = 0.0; }
        for (int l = mazim; l < N; l += 2) {
            const double yu0 = ylmu_m[l * NU + iu];
            const double2 *f0 = reinterpret_cast<const double2 *>(Fs + l * n + jq);
#pragma unroll
            for (int q2 = 0; q2 < JW / 2; q2++) {
                const double2 f = f0[q2];
                he[2 * q2] = fma(f.x, yu0, he[2 * q2]);
                he[2 * q2 + 1] = fma(f.y, yu0, he[2 * q2 + 1]);
            }
            if (l + 1 < N) {
                const double yu1 = ylmu_m[(l + 1) * NU + iu];
                const double2 *f1 = reinterpret_cast<const double2 *>(Fs + (l + 1) * n + jq);
#pragma unroll
                for (int q2 = 0; q2 < JW / 2; q2++) {
                    const double2 f = f1[q2];
                    ho[2 * q2] = fma(f.x, yu1, ho[2 * q2]);
                    ho[2 * q2 + 1] = fma(f.y, yu1, ho[2 * q2 + 1]);
                }
            }
        }
#pragma unroll
        for (int q = 0; q < JW; q++) {
            const int j = jq + q;
            GU[iu * EC + n + j] = sx[n + j] * (he[q] + ho[q]);
            GU[iu * EC + n - 1 - j] = sx[n - 1 - j] * (ho[q] - he[q]);
        }
    }
    for (int e = lane; e < NU * 3; e += 32) {
        const int iu = e / 3, cc = e - iu * 3;
        double acc = 0.0;
        for (int l = mazim; l < N; l++) acc = fma(Eb[l * 3 + cc], ylmu_m[l * NU + iu], acc);
        if (therm && cc == 1) acc += (1. - oprim) * xr0;
        if (therm && cc == 2) acc += (1. - oprim) * xr1;
        GU[iu * EC + N + cc] = acc;
    }
    __syncwarp();
}

// Source-function integral of one layer for one user angle (USRINT, disort.f:4355-4793,
// in recurrence form): the intensity leaving the layer towards the viewer is
//   I(near boundary) = T I(far boundary) + S,   T = exp(-dtau' / |umu|),
// with S the analytic integral of the layer's source terms; same L'Hospital limits as the
// reference (|1 +- umu k| < 1e-4, |1 + umu/umu0| < 1e-4).  up: umu > 0, the near boundary is
// the layer top; else the layer bottom.  The reference's L'Hospital form of the beam term
// (disort.f:4560-4566: ZBEAM dtau/umu0 exp(-utau'/umu0), the exponential taken at the USER LEVEL
// for every layer of the path) is not a recurrence term: rdenom == 0 leaves it out of S and
// returns its coefficient ZBEAM dtau/umu0 in `bl` -- the caller keeps the running sum.
template <int n>
__device__ __forceinline__ double layer_source(const double *gu /* GU row of this angle */, const double *kk,
                                               const double *ek, double umu, double rmu /* 1 / umu */,
                                               double rdenom /* 1 / (1 + umu/umu0), 0: L'Hospital */,
                                               double t0, double t1,
                                               double eb0, double eb1 /* beam transmission at t0, t1 */,
                                               bool beam, double umu0, bool therm, double &T, double &bl)
{
    constexpr int N = 2 * n;
    const double dtau = t1 - t0;
    const bool up = umu > 0.0;
    T = exp(-dtau * fabs(rmu));
    double s = 0.0;
    bl = 0.0;
    if (beam) {
        if (rdenom == 0.0) bl = gu[N] * (dtau / umu0);
        else s = gu[N] * (up ? (eb0 - T * eb1) * rdenom : (eb1 - T * eb0) * rdenom);
    }
    double sp = 0.0;      // second accumulator: two modes in flight (the reciprocals overlap)
#pragma unroll 2
    for (int j = 0; j < n; j++) {
        const double k = kk[j], wk = ek[j];
        // column n-1-j belongs to -k_j, column n+j to +k_j (disort.f:3264-3312); the divisions by
        // 1 -+ umu k are reciprocal multiplications (the IEEE division costs ~40 instructions each)
        const double dm = 1.0 - umu * k, dp = 1.0 + umu * k;
        double em, ep;
        if (up) {
            em = (fabs(dm) < 0.0001) ? dtau * rmu * T : (wk - T) * fast_rcp(dm);
            ep = (1.0 - T * wk) * fast_rcp(dp);
        } else {
            em = (1.0 - T * wk) * fast_rcp(dm);
            ep = (fabs(dp) < 0.0001) ? -dtau * rmu * T : (wk - T) * fast_rcp(dp);
        }
        s = fma(gu[n - 1 - j], em, s);
        sp = fma(gu[n + j], ep, sp);
    }
    s += sp;
    if (therm) {
        const double f0 = 1.0 - T;
        const double f1 = up ? (t0 + umu) - (t1 + umu) * T : (t1 + umu) - (t0 + umu) * T;
        s = fma(gu[N + 1], f0, s);
        s = fma(gu[N + 2], f1, s);
    }
    return s;
}

// WARPS warps per CTA, 16 warps per SM.  SYNC: the warps of a CTA move through the
// phases together (CTA barriers between them), so that at any time they execute the
// same code and share its instruction-cache lines; a warp that found no bin left
// keeps attending the barriers until every warp of the CTA is out of work.
// RAD: intensities at user angles (a.d.numu > 0): the phases are repeated for every azimuth
// mode, phase 3 also re-expands the layer solutions at the user cosines and integrates the
// source function (levels = layer boundaries, a.d.ntau == 0; always CTA-synchronous).
// ADD (radiance runs): the boundary-value problem in the adding form (sbd_addops.cuh) instead of
// the elimination: two sweeps leave the intensities at every interface, and the layer solutions
// the source-function integration needs follow from them by the eigenvectors' orthogonality.
template <int n, int WARPS, bool SYNC, bool RAD, bool ADD = false>
__global__ void __launch_bounds__(WARPS * 32, n > 8 ? 1 : 16 / WARPS)
disort_fast_kernel(const LaunchArgs a)
{
    using FL = FastLayout<n>;
    using AO = AddOps<n>;
    static_assert(!ADD || RAD, "the adding form of this kernel serves the radiance runs");
    static_assert(!ADD || ((size_t)AO::arec * 64 + 2 * n * 65 + n * n + n <= (size_t)FL::ublk * 64 &&
                           (size_t)AO::arec + 2 * n * 2 + n * n + n <= (size_t)FL::ublk),
                  "sweep records live in the pivot-row area");
    constexpr int N = 2 * n, TASKS = FL::tasks, GW = FL::GW, KS = FL::KS, LC = FL::LC, US = FL::US;
    static_assert(n <= 8 || ADD, "NSTR > 16: adding form only");
    static_assert(!RAD || SYNC, "radiance runs reload the CTA's Legendre table per azimuth mode");
    const int L = a.d.nlyr;
    const int NT = a.d.ntau > 0 ? a.d.ntau : L + 1;
    const int NU = RAD ? a.d.numu : 0, NPHI = RAD ? a.d.nphi : 0;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, warps = blockDim.x >> 5;
    const int ldp = a.d.nmom + 1;
    extern __shared__ double smem_fast[];
    double *cmu = smem_fast, *cwt = cmu + n, *csq = cwt + n, *cdinv = csq + n;
    double *cylm = cdinv + n + 2;       // cylm[-2] = sum(w mu), cylm[-1] = sum(w)
    double *cylmu = smem_fast + FL::cta;
    double *wsm = smem_fast + FL::cta_doubles(NU) + (size_t)warp * FL::warp_doubles(L, NT, NU, NPHI, ADD);
    double *y0 = wsm;
    // 16-byte aligned work area: per-task areas in phase 1, cp.async staging afterwards
    double *tsm_base = y0 + N;
    double *taucpr = tsm_base + (RAD ? FL::work_rad(NU, NPHI, ADD) : (size_t)FL::work), *tauc = taucpr + (L + 1);
    double *ebeam = tauc + (L + 1), *edir = ebeam + (L + 1);   // exp(-tau'/mu0), exp(-tau/mu0)
    double *pk = edir + (L + 1);
    double *lw = tsm_base;                                       // 3 x L prologue work values (the work area is idle then)
    int *layru = (int *)(pk + (L + 3));
#if SBD_USE_TMA
    // mbarriers of the bulk-copy staging: [0], [1] phase-3 double buffer, [2] phase-2 record prefetch
    unsigned long long *mbar = reinterpret_cast<unsigned long long *>(pk + (L + 3) + (NT + 1) / 2);
#endif
    // radiance work values (phase 3 only): behind the phase-3 buffers in the work area
    double *uE = tsm_base + FL::stage3(ADD);
    double *uGU = uE + N * FL::ecols, *uI = uGU + NU * FL::ecols, *cosm = uI + NU, *ugl = cosm + NPHI, *usx = ugl + N;

    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        double mu = a.quad[i], wt = a.quad[n + i];
        cmu[i] = mu; cwt[i] = wt; csq[i] = sqrt(wt / mu); cdinv[i] = 1.0 / sqrt(wt * mu);
    }
    if (threadIdx.x == 0) {
        double W = 0.0, SW = 0.0;
        for (int i = 0; i < n; i++) { W += a.quad[n + i] * a.quad[i]; SW += a.quad[n + i]; }
        cylm[-2] = W; cylm[-1] = SW;
    }
    for (int e = threadIdx.x; e < N * n; e += blockDim.x) cylm[e] = a.ylmc[e];
    __syncthreads();
    const double Wq = cylm[-2], SWq = cylm[-1];

    const int slot = blockIdx.x * warps + warp;
    double *scr = a.scratch + (size_t)slot * a.slot_stride;
    double *recs = scr;                                   // [L][rec]
    double *frecs = scr + (size_t)L * FL::rec;            // [L][frec]
    double *ublk = frecs + (size_t)L * FL::frec;          // [L][N][US]
    double *dscr = ublk + (size_t)L * FL::ublk;           // RAD: [L][2][NU] downward source, transmission
    double *arecs = ublk;                                 // ADD: [L][arec] sweep records, then [L+1][2n] interface intensities
    double *levs = ublk + (size_t)L * AO::arec;
    double *botb = levs + (size_t)(L + 1) * 2 * n;       // ADD: Rb, sb of the bottom boundary
    // layer group of this lane; lanes beyond the last full group shadow its last lane
    const int task = lane / GW < TASKS ? lane / GW : TASKS - 1;
    const bool gact = lane < TASKS * GW && lane % GW < n;
    const int g = gact ? lane % GW : n - 1;
    const int gbase = task * GW;
    // the spectrum path keeps the bin count on the device (a.d.nbins is then an upper bound)
    const int nbins_all = a.redo_consume ? *a.redo_count : (a.nbins_dev ? *a.nbins_dev : a.d.nbins);
    double *tsm = tsm_base + (size_t)task * FL::task;
    const int rg = lane >> 2, cg = lane & 3;              // 2-D tiling of phase 2
    const unsigned long long jpart = jacobi_partners<n>(g);
#if SBD_USE_TMA
    if (lane == 0) { mbar_init(mbar, 1); mbar_init(mbar + 1, 1); mbar_init(mbar + 2, 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncwarp();
    unsigned mphase = 0;      // bit b: parity of the next completion of mbarrier b
    // One lane issues the bulk copies.  Their sources (scratch records written with ordinary
    // stores earlier in the launch) are ordered before them by ONE proxy fence at the start of
    // phases 2 and 3 (after the phase barrier); a destination buffer is reused only after the
    // warp has finished reading it (__syncwarp), as in any mbarrier pipeline.  (A proxy fence
    // before every issue also waits for the lane's outstanding pivot-row stores: -8 %.)
#define SBD_BULK_BEGIN(bar, bytes) mbar_expect_tx(bar, bytes)
#endif

    for (;;) {
        int bin = 0;
        if (lane == 0) bin = atomicAdd(a.work_counter, 1);
        bin = __shfl_sync(FULLMASK, bin, 0);
        const bool have = bin < nbins_all;
        if (SYNC) { if (!__syncthreads_or(have)) break; }
        else if (!have) break;
        if (a.redo_consume && have) bin = a.redo_list[bin];      // bins handed over by the adding kernel
#ifdef SBD_PHASE_TIMING
        long long tphase = clock64();
#endif
        const int src = !have ? 0 : (a.binmap ? a.binmap[bin] : bin);     // input slot of this bin
        const sbd_bin bp = a.bins[src];
        const double *dtauc = a.dtauc + (size_t)src * L;
        const double *ssalb = a.ssalb + (size_t)src * L;
        const double *pmom = a.pmom + (size_t)src * L * ldp;
        const double fbeam = bp.fbeam, umu0 = bp.umu0, albedo = bp.albedo;
        const bool plank = bp.plank != 0;
        double *o_rfldir = (a.rfldir && have) ? a.rfldir + (size_t)bin * NT : nullptr;
        double *o_rfldn = (a.rfldn && have) ? a.rfldn + (size_t)bin * NT : nullptr;
        double *o_flup = (a.flup && have) ? a.flup + (size_t)bin * NT : nullptr;
        double *o_dfdt = (a.dfdt && have) ? a.dfdt + (size_t)bin * NT : nullptr;
        double *o_uavg = (a.uavg && have) ? a.uavg + (size_t)bin * NT : nullptr;

        int status = have ? 0 : -1;
        int surf = -1;
        {   // CHEKIN subset (disort.f:4920-5155)
            int badl = 0;
            for (int lc = lane; lc < L; lc += 32) {
                double s = ssalb[lc];
                if (!(s >= 0.0 && s <= 1.0)) badl = 1;
                if (!(fabs(dtauc[lc]) <= 1.79e308)) badl = 1;      // NaN / Inf optical depth
                for (int k = 1; k <= a.d.nmom; k++) {
                    double pm = pmom[(size_t)lc * ldp + k];
                    if (!(pm >= -1.0 && pm <= 1.0)) badl = 1;
                }
            }
            if (fbeam < 0.0 || (fbeam > 0.0 && !(umu0 > 0.0 && umu0 <= 1.0))) badl = 1;
            // albedo = SBD_SURFACE(s): BRDF surface s (LAMBER = .FALSE.; the adding form carries it)
            if (ADD && albedo < 0.0) {
                surf = (int)(-albedo) - 1;
                if (!a.sf_bdr || surf >= a.sf_count || (double)(surf + 1) != -albedo || a.sf_modes != N) badl = 1;
            } else if (!(albedo >= 0.0 && albedo <= 1.0)) badl = 1;
            if (bp.fisot < 0.0) badl = 1;
            if (plank && (bp.wvnmlo < 0.0 || bp.wvnmhi <= bp.wvnmlo || bp.temis < 0.0 ||
                          bp.temis > 1.0 || bp.btemp < 0.0 || bp.ttemp < 0.0)) badl = 1;
            // device-pointer callers: a Planck bin needs a valid row of temper[ncol][L+1]
            if (plank && (!a.temper || bp.col < 0 || bp.col >= a.d.ncol)) badl = 1;
            if (__any_sync(FULLMASK, badl)) status = SBD_BIN_BAD_INPUT;
            int clash = 0;
            if (fbeam > 0.0 && lane < n && fabs(umu0 - cmu[lane]) / umu0 < 1.e-4) clash = 1;
            if (!status && __any_sync(FULLMASK, clash)) status = SBD_BIN_ANGLE_CLASH;
        }
        int ncut = L, lyrcut = 0, monotone = 1;
        // SETDIS prologue (disort.f:2546-2605): lanes fetch the per-layer terms,
        // lane 0 accumulates them in the reference's order
        for (int lc = lane; lc < L; lc += 32) {
            double s = ssalb[lc];
            if (s == 1.0) s = 1.0 - kDither;
            const double dtr = dtauc[lc];
            const double dt = dtr < 0.0 ? 0.0 : dtr;
            const double f = pmom[(size_t)lc * ldp + N];
            lw[lc] = dtr; lw[L + lc] = (1. - s) * dt; lw[2 * L + lc] = (1. - f * s) * dt;
        }
        __syncwarp();
        if (lane == 0) {
            double tc = 0.0, tp = 0.0, abstau = 0.0;
            tauc[0] = 0.0; taucpr[0] = 0.0;
            for (int lc = 0; lc < L; lc++) {
                if (lw[lc] < 0.0) monotone = 0;
                tc += lw[lc];
                if (abstau < 10.0) ncut = lc + 1;
                abstau += lw[L + lc];
                tp += lw[2 * L + lc];
                tauc[lc + 1] = tc; taucpr[lc + 1] = tp;
            }
            lyrcut = (abstau >= 10.0 && !plank && L > 1);
            if (!lyrcut) ncut = L;
        }
        ncut = __shfl_sync(FULLMASK, ncut, 0);
        lyrcut = __shfl_sync(FULLMASK, lyrcut, 0);
        monotone = __shfl_sync(FULLMASK, monotone, 0);
        const bool fastmap = (a.d.ntau == 0) && monotone;
        __syncwarp();
        // Radiance runs: negative optical depths (legal upstream, taugas.f:7485) make TAUC
        // non-monotone (disort.f:487 accumulates before CHEKIN clips) and the reference then
        // integrates the source function to levels INSIDE earlier layers; this kernel's recurrence
        // has its levels at the layer boundaries, so such bins go to the general kernel
        if (RAD && !monotone && !status && have && a.redo_list && !a.redo_consume) {
            if (lane == 0) {
                const int k = atomicAdd(a.redo_count, 1);
                a.redo_list[k] = bin;
            }
            status = -100;          // parked: no phases, no status write
        }
        // beam transmission to every layer boundary: scaled depth (EXPBEA, disort.f:2592)
        // and true depth (the direct flux that is reported, disort.f:1998)
        for (int lev = lane; lev <= L; lev += 32) {
            ebeam[lev] = fbeam > 0.0 ? exp(-taucpr[lev] / umu0) : 0.0;
            edir[lev] = fbeam > 0.0 ? exp(-tauc[lev] / umu0) : 0.0;
        }
        int badtau = 0;
        for (int lu = lane; lu < NT; lu += 32) {
            double ut = a.d.ntau > 0 ? a.utau[(size_t)src * NT + lu] : tauc[lu];
            if (a.d.ntau > 0 && fabs(ut - tauc[L]) <= 1.e-4) ut = tauc[L];
            if (a.d.ntau > 0 && !(ut >= 0.0 && ut <= tauc[L])) badtau = 1;
            int lc;
            if (fastmap && lu >= 1 && tauc[lu - 1] < ut) {
                lc = lu;          // no earlier layer can contain this boundary
            } else {
                for (lc = 1; lc <= L; lc++)
                    if (ut >= tauc[lc - 1] && ut <= tauc[lc]) break;
                if (lc > L) lc = L;
            }
            layru[lu] = lc;
        }
        if (__any_sync(FULLMASK, badtau)) status = SBD_BIN_BAD_INPUT;
        double tplank = 0.0, bplank = 0.0;
        if (plank && !status) {
            // one list: L+1 level temperatures, then TTEMP, BTEMP (disort.f:556-571)
            const double *tp = a.temper + (size_t)bp.col * (L + 1);
            for (int lev = lane; lev <= L + 2; lev += 32) {
                const double t = lev <= L ? tp[lev] : (lev == L + 1 ? bp.ttemp : bp.btemp);
                pk[lev] = plkavg_dev(bp.wvnmlo, bp.wvnmhi, t);
            }
            __syncwarp();
            tplank = bp.temis * pk[L + 1];
            bplank = pk[L + 2];
        }
        if (lane == 0 && fbeam > 0.0) {   // Y_l^0(-mu0), LEPOLY m = 0
            double x = -umu0;
            y0[0] = 1.0; y0[1] = x;
            double pm2 = 1.0, pm1 = x;
#pragma unroll
            for (int l = 2; l < N; l++) {
                const double p = ((2 * l - 1) * x * pm1 - (l - 1) * pm2) * (1.0 / l);
                y0[l] = p; pm2 = pm1; pm1 = p;
            }
        }
        for (int lu = lane; lu < NT; lu += 32) {
            if (o_rfldir) o_rfldir[lu] = 0.0;
            if (o_rfldn) o_rfldn[lu] = 0.0;
            if (o_flup) o_flup[lu] = 0.0;
            if (o_dfdt) o_dfdt[lu] = 0.0;
            if (o_uavg) o_uavg[lu] = 0.0;
        }
        double *o_uu = nullptr;
        int naz = 0, kconv = 0;
        if (RAD) {
            if (have) {
                o_uu = a.uu + (size_t)bin * NPHI * a.uu_nt * NU;
                for (int e = lane; e < NPHI * a.uu_nt * NU; e += 32) o_uu[e] = 0.0;
            }
            // number of azimuth modes (disort.f:577-586)
            naz = N - 1;
            const double u0 = a.umu[0], u1 = NU > 1 ? a.umu[1] : 0.0;
            if (fbeam == 0.0 || fabs(1. - umu0) < 1.e-5 ||
                (NU == 1 && fabs(1. - u0) < 1.e-5) || (NU == 1 && fabs(1. + u0) < 1.e-5) ||
                (NU == 2 && fabs(1. + u0) < 1.e-5 && fabs(1. - u1) < 1.e-5))
                naz = 0;
        }
        __syncwarp();

        SBD_TICK(0);
      for (int mazim = 0; ; mazim++) {
        if (RAD) {
            // every warp of the CTA works on the same mode: the Legendre table of the mode is
            // CTA-shared (a warp without a mode left keeps attending the barriers)
            const bool mact = !status && mazim <= naz;
            if (!__syncthreads_or(mact)) break;
            for (int e = threadIdx.x; e < N * n; e += blockDim.x) cylm[e] = a.ylmc[(size_t)mazim * N * n + e];
            for (int e = threadIdx.x; e < N * NU; e += blockDim.x) cylmu[e] = a.ylmu[(size_t)mazim * N * NU + e];
            if (lane == 0 && fbeam > 0.0 && mazim > 0) lepoly_mode(mazim, N, -umu0, y0);
            __syncthreads();
        } else if (mazim > 0) break;
        const bool mrun = !RAD || mazim <= naz;      // false: parked, only attends the barriers
        const bool m0 = !RAD || mazim == 0;
        const double delm0 = m0 ? 1.0 : 0.0;
        const double *ylmu_m = RAD ? cylmu : nullptr;       // this mode's table (shared memory)
        // ===================== phase 1 =====================================
        if (mrun && !status) {
            for (int lc0 = 0; lc0 < ncut; lc0 += TASKS) {
                int lc = lc0 + task;
                const bool active = lc < ncut;
                if (!active) lc = ncut - 1;
                int st = phase1_layers<n, ADD>(dtauc, ssalb, pmom, ldp, lc, active, RAD ? mazim : 0, fbeam, umu0,
                                               plank && m0, delm0, cmu, cwt, csq, cdinv, cylm, y0, taucpr, pk, tsm,
                                               recs + (size_t)lc * FL::rec, frecs + (size_t)lc * FL::frec, g, gact, gbase, jpart,
                                               arecs + (size_t)lc * AO::arec, ebeam);
                if (__any_sync(FULLMASK, st != 0)) { status = SBD_BIN_EIG_FAIL; break; }
            }
        }
        __syncwarp();
        __threadfence_block();
        if (SYNC) __syncthreads();
        SBD_TICK(1);

        // ===================== phase 2: downward elimination ================
        // The window rows live in slots (rg, k), k < KS; slot s = k*8 + rg.  `act`
        // (warp-uniform) marks the slots that hold an equation not yet used as a pivot.
        // Layer records are staged global -> shared with cp.async, three slots deep:
        // stage lc uses records lc and lc+1 while record lc+2 is in flight.
        double *rslot = tsm_base;   // phase-1 task areas are idle now
        double *stg = tsm_base + 3 * FL::rec;     // assembled rows of the current layer
        if (ADD && mrun && !status) {
            // adding sweeps: bottom boundary (Lambertian: reflects the m = 0 mode only; a BRDF surface
            // every mode, SURFAC's tables disort.f:3765-3907; nothing comes up at a truncation level),
            // bottom-up operators, top-down intensities.  Scaled: Rb = (1 + delta_0m) D BDR D
            double *sRb = tsm_base, *ssb = sRb + n * n;
            if (surf >= 0 && !lyrcut) {
                const double *bdr = a.sf_bdr + ((size_t)surf * a.sf_modes + mazim) * n * (n + 1);
                const double *bem = a.sf_bem + (size_t)surf * n;
                const double beamfac = umu0 * fbeam / kPiRef * ebeam[ncut];
                for (int e = lane; e < n * n; e += 32) {
                    const int i = e / n, k = e - i * n;
                    sRb[e] = (1.0 + delm0) * (cmu[i] * csq[i]) * bdr[i * (n + 1) + 1 + k] * (cmu[k] * csq[k]);
                }
                if (lane < n)
                    ssb[lane] = cmu[lane] * csq[lane] * (bdr[lane * (n + 1)] * beamfac + delm0 * bem[lane] * bplank);
            } else {
                const bool refl = !lyrcut && m0;
                const double rbB = refl ? 2.0 * albedo : 0.0;
                const double sbB = refl ? albedo * umu0 * fbeam / kPiRef * ebeam[ncut] + (1.0 - albedo) * bplank : 0.0;
                for (int e = lane; e < n * n; e += 32) sRb[e] = rbB * (cmu[e / n] * csq[e / n]) * (cmu[e % n] * csq[e % n]);
                if (lane < n) ssb[lane] = cmu[lane] * csq[lane] * sbB;
            }
            __syncwarp();
            for (int e = lane; e < n * n + n; e += 32) botb[e] = sRb[e];      // kept for the top-down sweep
            __syncwarp();
            if (!adding_sweep_up<n>(arecs, ncut, tsm_base, lane)) status = SBD_BIN_SINGULAR;
            if (!status)
                adding_sweep_down<n>(arecs, ncut, levs, m0 ? bp.fisot + tplank : 0.0, botb, cmu, csq, lane);
        }
        if constexpr (!ADD) if (mrun && !status) {
            double w[KS][LC], rhs[KS];
#if SBD_TMA_P2
            __syncwarp();
            fence_proxy_async();
            if (lane == 0) {
                const int nr = ncut > 1 ? 2 : 1;          // records 0 and 1 are contiguous
                SBD_BULK_BEGIN(mbar + 2, 8u * FL::rec * nr);
                bulk_g2s(rslot, recs, 8u * FL::rec * nr, mbar + 2);
            }
            if (!mbar_wait(mbar + 2, (mphase >> 2) & 1u)) status = SBD_BIN_SINGULAR;    // never expected
            mphase ^= 4u;
            __syncwarp();
#else
            warp_copy_async(rslot, recs, FL::rec, lane);
            if (ncut > 1) warp_copy_async(rslot + FL::rec, recs + FL::rec, FL::rec, lane);
            cp_async_commit();
            cp_async_wait_all();
            __syncwarp();
#endif
            // top boundary rows r = 0..n-1 in slots 0..n-1 (disort.f:2887-2915, :3547-3550)
            stage_rows<n>(stg, rslot, false, nullptr, 0, n, 0.0, cwt, cmu, lane);
            if (lane < n)
                stg[lane * US + 4 * LC] = (m0 ? bp.fisot + tplank : 0.0) - rslot[FL::off_zz + lane] - rslot[FL::off_zp0 + lane];
            __syncwarp();
#pragma unroll
            for (int k = 0; k < KS; k++) {
                const int r = k * 8 + rg;
                if (r < n) {
#pragma unroll
                    for (int l = 0; l < LC; l++) w[k][l] = stg[r * US + cg * LC + l];
                    rhs[k] = stg[r * US + 4 * LC];
                } else {
                    rhs[k] = 0.0;
#pragma unroll
                    for (int l = 0; l < LC; l++) w[k][l] = 0.0;
                }
            }
            unsigned act = (1u << n) - 1u;
            __syncwarp();
            for (int lc = 0; lc < ncut; lc++) {
                const bool last = (lc == ncut - 1);
#if SBD_TMA_P2
                if (lc + 2 < ncut && lane == 0) {      // (the slot was last read two layers ago)
                    SBD_BULK_BEGIN(mbar + 2, 8u * FL::rec);
                    bulk_g2s(rslot + ((lc + 2) % 3) * FL::rec, recs + (size_t)(lc + 2) * FL::rec, 8u * FL::rec, mbar + 2);
                }
#else
                if (lc + 2 < ncut) {
                    warp_copy_async(rslot + ((lc + 2) % 3) * FL::rec, recs + (size_t)(lc + 2) * FL::rec,
                                    FL::rec, lane);
                    cp_async_commit();
                }
#endif
                const double *rc = rslot + (lc % 3) * FL::rec;
                const double *rn = rslot + ((lc + 1) % 3) * FL::rec;
                const double tb = taucpr[lc + 1];
                const double eb = ebeam[lc + 1];
                // free slots take the new equations: N interface rows, or n bottom rows
                const unsigned freem = ~act & FL::slotmask;
                const int nnew = last ? n : N;
                unsigned newm = freem;
                if (__popc(freem) > nnew) {          // keep the lowest nnew free slots
                    newm = 0;
                    unsigned f = freem;
                    for (int i = 0; i < nnew; i++) { const unsigned b = f & (0u - f); newm |= b; f ^= b; }
                }
                if (!last) {
                    // interface lc | lc+1: u_lc(bottom) = u_lc+1(top)  (disort.f:2846-2882, :3585-3593)
                    stage_rows<n>(stg, rc, true, rn, 0, N, 0.0, cwt, cmu, lane);
                    if (lane < N)
                        stg[lane * US + 4 * LC] = (rn[FL::off_zz + lane] - rc[FL::off_zz + lane]) * eb +
                                                  rn[FL::off_zp0 + lane] - rc[FL::off_zp0 + lane] +
                                                  (rn[FL::off_xr + 1] - rc[FL::off_xr + 1]) * tb;
                } else {
                    // bottom boundary, Lambertian m = 0 (disort.f:2919-2990, :3552-3578)
                    // (the Lambertian surface reflects the m = 0 mode only, disort.f:3753-3763)
                    const bool refl = !lyrcut && m0;
                    stage_rows<n>(stg, rc, true, nullptr, n, n, refl ? 2.0 * albedo : 0.0, cwt, cmu, lane);
                    if (lane < n) {
                        const int r = n + lane;
                        const double xr1 = rc[FL::off_xr + 1];
                        double v = -rc[FL::off_zz + r] * eb - rc[FL::off_zp0 + r] - xr1 * tb;
                        if (refl) {
                            double rsum = 0.0;
#pragma unroll 1
                            for (int k = 0; k < n; k++)
                                rsum = fma(cwt[k] * cmu[k], rc[FL::off_zz + n - 1 - k] * eb +
                                                                rc[FL::off_zp0 + n - 1 - k] + xr1 * tb, rsum);
                            v += 2.0 * albedo * rsum + albedo * umu0 * fbeam / kPiRef * eb +
                                 (1.0 - albedo) * bplank;
                        }
                        stg[lane * US + 4 * LC] = v;
                    }
                }
                __syncwarp();
#pragma unroll
                for (int k = 0; k < KS; k++) {
                    const int s = k * 8 + rg;
                    if ((newm >> s) & 1u) {
                        const double *row = stg + __popc(freem & ((1u << s) - 1u)) * US;
#pragma unroll
                        for (int l2 = 0; l2 < LC / 2; l2++) {
                            const double2 v = reinterpret_cast<const double2 *>(row + cg * LC)[l2];
                            w[k][2 * l2] = v.x; w[k][2 * l2 + 1] = v.y;
                        }
                        rhs[k] = row[4 * LC];
                    }
                }
                act |= newm;
                // eliminate the N columns of layer lc: one loop body serves all of them
                // (the slices slide, see elim_step)
                // pivot row j lands in scratch as [cg][l - j/4] slices, then the right-hand side
                double *uslice = ublk + (size_t)lc * FL::ublk + cg * LC;
                bool sing = false;
                // second half of the layer's columns: the slices have slid LH/2 times, their
                // last LH/2 entries are zero in every live row -- a narrower loop body
                constexpr int W2 = ((LC - FL::LH / 2) % 2 == 0 && FL::LH >= 2) ? LC - FL::LH / 2 : LC;
#pragma unroll 1
                for (int j = 0; j < N / 2 && !sing; j++, uslice += US) sing = elim_step<n, LC>(w, rhs, act, uslice, j & 3, rg, cg);
#pragma unroll 1
                for (int j = N / 2; j < N && !sing; j++, uslice += US) sing = elim_step<n, W2>(w, rhs, act, uslice, j & 3, rg, cg);
                if (sing) { status = SBD_BIN_SINGULAR; break; }
#if SBD_TMA_P2
                if (lc + 2 < ncut) {       // record lc+2 has landed
                    if (!mbar_wait(mbar + 2, (mphase >> 2) & 1u)) { status = SBD_BIN_SINGULAR; break; }
                    mphase ^= 4u;
                }
#else
                cp_async_wait_all();       // record lc+2 has landed
#endif
                __syncwarp();
            }
        }
        cp_async_wait_all();
        __syncwarp();
        __threadfence_block();
        if (SYNC) __syncthreads();
        SBD_TICK(2);

        // ===================== phase 3: back substitution + fluxes ===========
        if (mrun && !status) {
            double xs[N];          // solution of the layer below (uniform)
#pragma unroll
            for (int j = 0; j < N; j++) xs[j] = 0.0;
            int lu_next = NT - 1;  // levels are visited bottom-up when the map is monotone
            // pivot rows + flux record of layer lc-1 stream into the other half of a
            // double buffer (cp.async) while layer lc is being solved
            constexpr int UB = ADD ? 0 : FL::ublk;      // pivot-row part of a buffer
            constexpr int kSlot = UB + FL::frec + (RAD ? FL::rec : 0);
            auto fetch_layer = [&](int lyr, int buf) {
                double *dstp = tsm_base + buf * kSlot;
#if SBD_TMA_P3
                if (lane == 0) {
                    SBD_BULK_BEGIN(mbar + buf, 8u * (kSlot - (ADD ? UB : 0)));
                    if (!ADD) bulk_g2s(dstp, ublk + (size_t)lyr * FL::ublk, 8u * FL::ublk, mbar + buf);
                    bulk_g2s(dstp + UB, frecs + (size_t)lyr * FL::frec, 8u * FL::frec, mbar + buf);
                    if (RAD) bulk_g2s(dstp + UB + FL::frec, recs + (size_t)lyr * FL::rec, 8u * FL::rec, mbar + buf);
                }
#else
                if (!ADD) warp_copy_async(dstp, ublk + (size_t)lyr * FL::ublk, FL::ublk, lane);
                warp_copy_async(dstp + UB, frecs + (size_t)lyr * FL::frec, FL::frec, lane);
                if (RAD) warp_copy_async(dstp + UB + FL::frec, recs + (size_t)lyr * FL::rec, FL::rec, lane);
                cp_async_commit();
#endif
            };
            // RAD: upward intensities are carried bottom-up along with the back substitution
            // (one user angle per lane); bnd_up = intensity leaving the surface (disort.f:4747-4778)
            double bnd_up = 0.0, azerr = 0.0;
            const double rpd = kPiRef / 180.0;
            if (RAD) {
                for (int j = lane; j < NPHI; j += 32) cosm[j] = cos(mazim * (rpd * (a.phi[j] - bp.phi0)));
                __syncwarp();
            }
            // adds this mode's term of the azimuth series at (level, angle) (disort.f:767-825)
            auto emit = [&](int lu, int iu, double val) {
                const int slot = a.uu_slot[lu];
                if (slot < 0) return;
                for (int j = 0; j < NPHI; j++) {
                    double *pu = o_uu + ((size_t)j * a.uu_nt + slot) * NU + iu;
                    if (mazim == 0) { *pu = val; continue; }
                    const double azterm = val * cosm[j];
                    const double unew = *pu + azterm;
                    *pu = unew;
                    // convergence test of the series (RATIO, disort.f:6159, and :821): is
                    // |term| / |sum| <= ACCUR everywhere?  Kept as a flag (a comparison instead of a
                    // division per intensity; SBDART runs ACCUR = 0, where both forms say term == 0)
                    const double aa = fabs(azterm), bb = fabs(unew);
                    const bool small = (aa == 0.0) ? (bb != 0.0 || 1.0 <= bp.accur) : (bb != 0.0 && aa <= bp.accur * bb);
                    if (!small) azerr = 1.0;
                }
            };
#if SBD_TMA_P3
            fence_proxy_async();      // pivot rows and flux records: ordinary stores of phases 1 and 2
#endif
            fetch_layer(ncut - 1, 0);
            for (int lc = ncut - 1; lc >= 0; lc--) {
                const int buf = (ncut - 1 - lc) & 1;
                // radiance runs: the layer's single-scattering albedo and moments for g_l, loaded
                // here so that the back substitution hides their latency
                double rad_ss = 0.0, rad_f = 0.0, rad_pm = 1.0;
                if (RAD) {
                    rad_ss = ssalb[lc];
                    rad_f = pmom[(size_t)lc * ldp + N];
                    if (lane > 0 && lane < N) rad_pm = pmom[(size_t)lc * ldp + lane];
                }
#ifdef SBD_PHASE_TIMING
                long long tsub = clock64();
#endif
#if SBD_TMA_P3
                if (lc > 0) fetch_layer(lc - 1, buf ^ 1);
                if (!mbar_wait(mbar + buf, (mphase >> buf) & 1u)) status = SBD_BIN_SINGULAR;   // never expected
                mphase ^= 1u << buf;
#else
                if (lc > 0) { fetch_layer(lc - 1, buf ^ 1); cp_async_wait_one(); }
                else cp_async_wait_all();
#endif
                __syncwarp();
#ifdef SBD_PHASE_TIMING
                if (threadIdx.x == 0) { const long long t = clock64(); atomicAdd(&g_phase_ticks[4], (unsigned long long)(t - tsub)); tsub = t; }
#endif
                const double *ubuf = tsm_base + buf * kSlot;
                const double *fr = ubuf + UB;
                if constexpr (ADD) {
                    // Layer solution from the interface intensities (scaled, u^ = D u): with the
                    // homogeneous parts du = u^ - D p at the layer top and bottom,
                    //   x(+k_j) = -k_j sum_i D_i [G+_ij du+_i - G-_ij du-_i]   at the top,
                    //   x(-k_j) =  k_j sum_i D_i [G+_ij du-_i - G-_ij du+_i]   at the bottom
                    // (orthogonality of the eigenvectors; each family is taken at the boundary where
                    // it is not attenuated).
                    const double *urc = fr + FL::frec;
                    const double *lt = levs + (size_t)lc * 2 * n, *lb = lt + 2 * n;
                    double xp = 0.0, xm = 0.0;
                    if (lane < n) {
                        const double et = ebeam[lc], ebt = ebeam[lc + 1], tt = taucpr[lc], tb = taucpr[lc + 1];
                        const double xr1 = urc[FL::off_xr + 1];
                        double ap = 0.0, am = 0.0;
#pragma unroll
                        for (int i = 0; i < n; i++) {
                            const double Di = cmu[i] * csq[i];
                            const double zu = urc[FL::off_zz + n + i], zd = urc[FL::off_zz + n - 1 - i];
                            const double pu = urc[FL::off_zp0 + n + i], pd = urc[FL::off_zp0 + n - 1 - i];
                            const double gp = urc[FL::off_gp + i * n + lane], gm = urc[FL::off_gm + i * n + lane];
                            const double dut = lt[n + i] - Di * (zu * et + pu + xr1 * tt);
                            const double ddt = lt[i] - Di * (zd * et + pd + xr1 * tt);
                            const double dub = lb[n + i] - Di * (zu * ebt + pu + xr1 * tb);
                            const double ddb = lb[i] - Di * (zd * ebt + pd + xr1 * tb);
                            ap = fma(Di, gp * dut - gm * ddt, ap);
                            am = fma(Di, gp * ddb - gm * dub, am);
                        }
                        const double k = urc[FL::off_kk + lane];
                        xp = -k * ap; xm = k * am;
                    }
#pragma unroll
                    for (int j = 0; j < n; j++) {
                        xs[n + j] = __shfl_sync(FULLMASK, xp, j);
                        xs[n - 1 - j] = __shfl_sync(FULLMASK, xm, j);
                    }
                    if (lane < n) { usx[n + lane] = xp; usx[n - 1 - lane] = xm; }
                } else {
                double acc, dinv;
                double ur[N];      // row `lane` of the upper triangle, pre-divided by the diagonal
                                   // (entries left of the diagonal unused)
                {
                    // stored row `lane`: window column c sits at (c & 3) * LC + (c >> 2) - lane / 4
                    // (the slices had slid lane/4 times when the row became a pivot)
                    const int row = lane < N ? lane : 0;
                    const double *u = ubuf + row * US - (row >> 2);
                    dinv = fast_rcp(u[(row & 3) * LC + (row >> 2)]);
                    // rhs - U[:, N:2N] x_below: four partial sums (a 16-long FMA chain otherwise)
                    double a4[4] = { ubuf[row * US + 4 * LC], 0.0, 0.0, 0.0 };
#pragma unroll
                    for (int j = 0; j < N; j++)
                        a4[j & 3] = fma(-u[((N + j) & 3) * LC + ((N + j) >> 2)], xs[j], a4[j & 3]);
                    acc = ((a4[0] + a4[1]) + (a4[2] + a4[3])) * dinv;
#pragma unroll
                    for (int c = 1; c < N; c++) ur[c] = u[(c & 3) * LC + (c >> 2)] * dinv;
                }
                // x_c = acc of lane c; the chain per step is one shuffle + one FMA
#pragma unroll
                for (int c = N - 1; c >= 0; c--) {
                    const double xc = __shfl_sync(FULLMASK, acc, c);
                    xs[c] = xc;
                    if (lane < c) acc = fma(-ur[c], xc, acc);
                }
                if (RAD) {
                    double v = 0.0;
#pragma unroll
                    for (int j2 = 0; j2 < N; j2++) v = (lane == j2) ? xs[j2] : v;
                    if (lane < N) usx[lane] = v;
                }
                }
#ifdef SBD_PHASE_TIMING
                if (threadIdx.x == 0) { const long long t = clock64(); atomicAdd(&g_phase_ticks[5], (unsigned long long)(t - tsub)); tsub = t; }
#endif
                // ---- fluxes at the levels living in this layer (FLUXES, disort.f:1780) ----
                if (fastmap)
                    while (lu_next >= 0 && layru[lu_next] > lc + 1) lu_next--;
                for (int lu = fastmap ? lu_next : NT - 1; lu >= 0 && m0; lu--) {
                    if (layru[lu] != lc + 1) { if (fastmap) break; else continue; }
                    if (fastmap) lu_next = lu - 1;
                    const bool atbot = (a.d.ntau == 0 && lu == lc + 1);
                    const bool attop = (a.d.ntau == 0 && lu == lc);
                    const double *sc = fr + FL::f_sc;
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
                    // S_t = sum_j cu[t][j] x_j f_j, t = up, down, mean; f_j = attenuation of
                    // mode j between its reference boundary and the level
                    double dot = 0.0;
                    if (atbot || attop) {
                        // lanes 0..2 take one functional each; x is uniform
                        double xe[N];
                        if (atbot) {
#pragma unroll
                            for (int j = 0; j < n; j++) {
                                xe[n + j] = xs[n + j] * fr[FL::f_ek + j];
                                xe[n - 1 - j] = xs[n - 1 - j];
                            }
                        } else {
#pragma unroll
                            for (int j = 0; j < n; j++) {
                                xe[n + j] = xs[n + j];
                                xe[n - 1 - j] = xs[n - 1 - j] * fr[FL::f_ek + j];
                            }
                        }
                        if (lane < 3) {
                            const double2 *cu2 = reinterpret_cast<const double2 *>(fr + FL::f_cu + lane * N);
                            double d0 = 0.0, d1 = 0.0, d2 = 0.0, d3 = 0.0;
#pragma unroll
                            for (int j2 = 0; j2 < n; j2 += 2) {
                                const double2 c = cu2[j2];
                                d0 = fma(c.x, xe[2 * j2], d0);
                                d1 = fma(c.y, xe[2 * j2 + 1], d1);
                                if (j2 + 1 < n) {
                                    const double2 e2 = cu2[j2 + 1];
                                    d2 = fma(e2.x, xe[2 * j2 + 2], d2);
                                    d3 = fma(e2.y, xe[2 * j2 + 3], d3);
                                }
                            }
                            dot = (d0 + d1) + (d2 + d3);
                        }
                    } else {     // level inside a layer (USRTAU): one mode per lane
                        double xl[N];
#pragma unroll
                        for (int j = 0; j < N; j++) xl[j] = xs[j];
                        double t0 = 0.0, t1 = 0.0, t2 = 0.0;
                        if (lane < N) {
                            const bool plusm = lane >= n;
                            const int jm = plusm ? lane - n : n - 1 - lane;
                            const double k = fr[FL::f_kk + jm];
                            const double d = plusm ? utp - taucpr[lc] : taucpr[lc + 1] - utp;
                            const double xf = xl[lane] * exp(-k * d);
                            t0 = fr[FL::f_cu + lane] * xf;
                            t1 = fr[FL::f_cu + N + lane] * xf;
                            t2 = fr[FL::f_cu + 2 * N + lane] * xf;
                        }
#pragma unroll
                        for (int o = 16; o > 0; o >>= 1) {
                            t0 += __shfl_xor_sync(FULLMASK, t0, o);
                            t1 += __shfl_xor_sync(FULLMASK, t1, o);
                            t2 += __shfl_xor_sync(FULLMASK, t2, o);
                        }
                        dot = lane == 0 ? t0 : (lane == 1 ? t1 : t2);
                    }
                    // particular solutions: beam, thermal constant and slope
                    if (lane < 3) {
                        const double wsum = lane < 2 ? Wq : 2.0 * SWq;
                        dot += sc[lane] * fact + sc[3 + lane] + sc[7] * utp * wsum;
                    }
                    const double sdn = __shfl_sync(FULLMASK, dot, 1);
                    const double sav = __shfl_sync(FULLMASK, dot, 2);
                    if (RAD && lu == L && !lyrcut && surf < 0)      // Lambertian surface, m = 0 (disort.f:4747-4778)
                        bnd_up = 2.0 * albedo * sdn + umu0 * fbeam / kPiRef * albedo * fact + (1.0 - albedo) * bplank;
                    if (lane == 0) {
                        const double pi = kPiRef;
                        const double dirint = fbeam * fact;
                        const double fldir = umu0 * (fbeam * fact);
                        const double rfldir = umu0 * fbeam * edr;
                        const double flup = 2. * pi * dot, fldn = 2. * pi * sdn;
                        const double fdntot = fldn + fldir;
                        // (x) / (4 pi) as a multiplication: the IEEE division costs ~40 instructions
                        // on the one active lane; the result differs by at most one ulp
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
#ifdef SBD_PHASE_TIMING
                if (RAD && threadIdx.x == 0) { const long long t = clock64(); atomicAdd(&g_phase_ticks[6], (unsigned long long)(t - tsub)); tsub = t; }
#endif
                if (RAD) {
                    // ---- intensities at the user angles (TERPEV, TERPSO, USRINT) ----
                    const double *urec = fr + FL::frec;
                    const double *sc = fr + FL::f_sc;
                    double ss = rad_ss;
                    if (ss == 1.0) ss = 1.0 - kDither;
                    const double f = rad_f;
                    const double oprim = ss * (1. - f) / (1. - f * ss);
                    (void)sc;
                    if (lane < N)                        // g_l of this layer (delta-M, disort.f:2583)
                        ugl[lane] = (2 * lane + 1) * oprim * (rad_pm - f) / (1. - f);
                    __syncwarp();
                    const bool therm = plank && m0;
                    if (lc == ncut - 1) {                 // intensity entering the bottom layer from below
                        for (int iu = lane; iu < NU; iu += 32) {
                            double v = 0.0;
                            if (a.umu[iu] > 0.0 && !lyrcut) {
                                if (ADD && surf >= 0) {
                                    // BRDF surface (disort.f:4744-4778): RMU-weighted downward intensities
                                    // at the bottom (w mu u = D u^), direct beam, emission
                                    const double *rm = a.sf_rmu + (((size_t)surf * a.sf_modes + mazim) * NU + iu) * (n + 1);
                                    const double *db = levs + (size_t)ncut * 2 * n;
                                    for (int k = 0; k < n; k++) v = fma(rm[1 + k], cmu[k] * csq[k] * db[k], v);
                                    v *= 1.0 + delm0;
                                    if (fbeam > 0.0) v += umu0 * fbeam / kPiRef * rm[0] * ebeam[ncut];
                                    if (m0) v += a.sf_emu[(size_t)surf * NU + iu] * bplank;
                                } else if (ADD && m0) {
                                    // Lambertian surface (disort.f:4747-4778), from the downward
                                    // intensities at the bottom interface (not from the flux of level
                                    // L: with empty layers that level is evaluated in an earlier layer)
                                    const double *db = levs + (size_t)ncut * 2 * n;
                                    double dn = 0.0;
                                    for (int k = 0; k < n; k++) dn = fma(cmu[k] * csq[k], db[k], dn);
                                    v = 2.0 * albedo * dn + umu0 * fbeam / kPiRef * albedo * ebeam[ncut] + (1.0 - albedo) * bplank;
                                } else if (m0) v = bnd_up;
                            }
                            uI[iu] = v;
                        }
                    }
                    // level at the bottom of the bottom layer
                    if (lc == ncut - 1)
                        for (int iu = lane; iu < NU; iu += 32)
                            if (a.umu[iu] > 0.0) emit(ncut, iu, uI[iu]);
                    __syncwarp();
                    user_terms_fast<n>(urec, usx, ugl, cwt, cylm, y0, ylmu_m, NU, mazim, fbeam > 0.0,
                                       (2. - delm0) * fbeam / (4.0 * kPiRef), therm, oprim, uE, uGU, lane);
                    for (int iu = lane; iu < NU; iu += 32) {
                        const double umu = a.umu[iu];
                        const double rmu = fast_rcp(umu), denom = 1. + umu / umu0;
                        double T, bl;
                        const double S = layer_source<n>(uGU + iu * FL::ecols, urec + FL::off_kk, urec + FL::off_ek,
                                                         umu, rmu, fabs(denom) < 0.0001 ? 0.0 : fast_rcp(denom),
                                                         taucpr[lc], taucpr[lc + 1], ebeam[lc], ebeam[lc + 1],
                                                         fbeam > 0.0, umu0, therm, T, bl);
                        if (umu > 0.0) {
                            const double below = T * uI[iu];
                            const double v = below + S;
                            uI[iu] = v;
                            // level lc = top of layer lc.  The reference drops the source integral of
                            // the layer that contains the level when the level lies within 1e-6
                            // (scaled depth) of the boundary the light leaves through
                            // (disort.f:4635-4641); with levels at the layer boundaries that is the
                            // top layer seen from level 0 -- kept so that thin cap layers agree
                            emit(lc, iu, (lc == 0 && taucpr[1] - taucpr[0] < 1.e-6) ? below : v);
                        } else {                          // kept for the top-down pass
                            // (viewing against the beam, |1 + umu/umu0| < 1e-4: the beam coefficient
                            // instead of the transmission, which the pass then recomputes)
                            dscr[((size_t)lc * 2) * NU + iu] = S;
                            dscr[((size_t)lc * 2 + 1) * NU + iu] = (fbeam > 0.0 && fabs(denom) < 0.0001) ? bl : T;
                        }
                    }
                }
                __syncwarp();     // everyone is done with this half of the double buffer
#ifdef SBD_PHASE_TIMING
                if (threadIdx.x == 0) { const long long t = clock64(); atomicAdd(&g_phase_ticks[RAD ? 7 : 6], (unsigned long long)(t - tsub)); tsub = t; }
#endif
            }
            if (RAD) {
                // downward intensities, top-down (disort.f:4735-4741: the top boundary emits
                // FISOT + TPLANK in the m = 0 mode)
                for (int iu = lane; iu < NU; iu += 32) {
                    if (a.umu[iu] > 0.0) continue;
                    double v = m0 ? bp.fisot + tplank : 0.0;
                    emit(0, iu, v);
                    // viewing against the beam: the reference's L'Hospital beam term of the layers above
                    // a level is exp(-tau'_level / umu0) x the sum of their coefficients (disort.f:4560-4566)
                    const bool lh = fbeam > 0.0 && fabs(1. + a.umu[iu] / umu0) < 0.0001;
                    const double armu = fabs(1.0 / a.umu[iu]);
                    double accb = 0.0;
                    for (int lc = 0; lc < ncut; lc++) {
                        const double w1 = dscr[((size_t)lc * 2 + 1) * NU + iu];
                        const double T = lh ? exp(-(taucpr[lc + 1] - taucpr[lc]) * armu) : w1;
                        const double above = T * v;
                        v = above + dscr[((size_t)lc * 2) * NU + iu];
                        // (same rule: a layer thinner than 1e-6 directly above the level, disort.f:4635-4641)
                        const bool thin = taucpr[lc + 1] - taucpr[lc] < 1.e-6;
                        double out = thin ? above : v;
                        if (lh) {
                            out += ebeam[lc + 1] * (thin ? accb : accb + w1);
                            accb += w1;
                        }
                        emit(lc + 1, iu, out);
                    }
                }
                if (mazim > 0) {
                    if (!__any_sync(FULLMASK, azerr != 0.0)) kconv++;
                    if (kconv >= 2) naz = mazim;       // converged: no further modes (disort.f:821-823)
                }
                __syncwarp();
            }
        }
        if (RAD) SBD_TICK(3);
      }   // azimuth modes
        cp_async_wait_all();
        if (lane == 0 && have && status != -100) a.status[bin] = status;
        __syncwarp();
#ifdef SBD_PHASE_TIMING
        if (SYNC) __syncthreads();
        SBD_TICK(3);
#endif
    }
}

#ifdef SBD_PHASE_TIMING
extern "C" void sbd_debug_phase_ticks(unsigned long long *out, int reset)
{
    cudaMemcpyFromSymbol(out, g_phase_ticks, sizeof(g_phase_ticks));
    if (reset) { unsigned long long z[8] = {0, 0, 0, 0, 0, 0, 0, 0}; cudaMemcpyToSymbol(g_phase_ticks, z, sizeof z); }
}
#endif

// ---- host-side launch helpers ---------------------------------------------
// CTA shape: 8 warps with phase barriers by default (measured best); SBD_FAST_WARPS = 4 | 8 and
// SBD_FAST_SYNC = 0 | 1 override it (tuning knobs, not API).  18 warps per SM at 96
// registers were measured as well: the spills cost more than the occupancy gains.
int fast_warps()
{
    const char *e = getenv("SBD_FAST_WARPS");
    const int w = e ? atoi(e) : 8;
    return (w == 4 || w == 8) ? w : 8;
}
static bool fast_sync(int warps)
{
    const char *e = getenv("SBD_FAST_SYNC");
    return e ? atoi(e) != 0 : warps > 4;
}

template <int n, int WARPS, bool SYNC, bool RAD, bool ADD = false>
static cudaError_t launch_fast_k(const LaunchArgs &a, int grid, size_t smem, cudaStream_t st)
{
    cudaError_t e = cudaFuncSetAttribute(disort_fast_kernel<n, WARPS, SYNC, RAD, ADD>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    disort_fast_kernel<n, WARPS, SYNC, RAD, ADD><<<grid, WARPS * 32, smem, st>>>(a);
    return cudaGetLastError();
}

// radiance runs: the adding form unless SBD_RAD_ELIM = 1 asks for the elimination (comparison knob, NSTR <= 16)
static bool rad_adding(int N) { return N > 16 || !getenv("SBD_RAD_ELIM"); }

template <int n>
static cudaError_t launch_fast_t(const LaunchArgs &a, int warps, int grid, cudaStream_t st)
{
    const int L = a.d.nlyr, NT = a.d.ntau > 0 ? a.d.ntau : L + 1;
    const bool radd = rad_adding(a.d.nstr);
    size_t smem = 8 * (FastLayout<n>::cta_doubles(a.d.numu) + (size_t)warps * FastLayout<n>::warp_doubles(L, NT, a.d.numu, a.d.nphi, radd));
    if constexpr (n > 8) {   // NSTR 20/24/32: radiance runs in the adding form, 4-warp CTAs (register budget)
        if (a.d.numu == 0 || warps != 4) return cudaErrorInvalidValue;
        return launch_fast_k<n, 4, true, true, true>(a, grid, smem, st);
    } else {
    if (a.d.numu > 0) {      // radiance runs: CTA-synchronous always
        // adding sweeps + solution recovery; SBD_RAD_ELIM = 1: the elimination (comparison knob)
        const bool add = radd;
        switch (warps) {
        case 4: return add ? launch_fast_k<n, 4, true, true, true>(a, grid, smem, st) : launch_fast_k<n, 4, true, true>(a, grid, smem, st);
        case 8: return add ? launch_fast_k<n, 8, true, true, true>(a, grid, smem, st) : launch_fast_k<n, 8, true, true>(a, grid, smem, st);
        }
        return cudaErrorInvalidValue;
    }
    const bool sync = fast_sync(warps);
    switch (warps) {
    case 4: return sync ? launch_fast_k<n, 4, true, false>(a, grid, smem, st) : launch_fast_k<n, 4, false, false>(a, grid, smem, st);
    case 8: return launch_fast_k<n, 8, true, false>(a, grid, smem, st);
    }
    return cudaErrorInvalidValue;
    }
}

bool fast_supported(int N) { return N == 4 || N == 8 || N == 16; }
// radiance runs (adding form) reach further
bool fast_rad_supported(int N) { return fast_supported(N) || ((N == 20 || N == 24 || N == 32) && !getenv("SBD_RAD_ELIM")); }

size_t fast_slot_doubles(int N, int L, int NU)
{
    switch (N) {
    case 4: return FastLayout<2>::slot_doubles_rad(L, NU);
    case 8: return FastLayout<4>::slot_doubles_rad(L, NU);
    case 16: return FastLayout<8>::slot_doubles_rad(L, NU);
    case 20: return FastLayout<10>::slot_doubles_rad(L, NU);
    case 24: return FastLayout<12>::slot_doubles_rad(L, NU);
    case 32: return FastLayout<16>::slot_doubles_rad(L, NU);
    }
    return 0;
}

size_t fast_smem_bytes(int N, int L, int NT, int warps, int NU, int NPHI)
{
    switch (N) {
    case 4: return 8 * (FastLayout<2>::cta_doubles(NU) + (size_t)warps * FastLayout<2>::warp_doubles(L, NT, NU, NPHI, rad_adding(N)));
    case 8: return 8 * (FastLayout<4>::cta_doubles(NU) + (size_t)warps * FastLayout<4>::warp_doubles(L, NT, NU, NPHI, rad_adding(N)));
    case 16: return 8 * (FastLayout<8>::cta_doubles(NU) + (size_t)warps * FastLayout<8>::warp_doubles(L, NT, NU, NPHI, rad_adding(N)));
    case 20: return 8 * (FastLayout<10>::cta_doubles(NU) + (size_t)warps * FastLayout<10>::warp_doubles(L, NT, NU, NPHI, rad_adding(N)));
    case 24: return 8 * (FastLayout<12>::cta_doubles(NU) + (size_t)warps * FastLayout<12>::warp_doubles(L, NT, NU, NPHI, rad_adding(N)));
    case 32: return 8 * (FastLayout<16>::cta_doubles(NU) + (size_t)warps * FastLayout<16>::warp_doubles(L, NT, NU, NPHI, rad_adding(N)));
    }
    return 0;
}

cudaError_t launch_fast(const LaunchArgs &a, int warps, int grid, cudaStream_t st)
{
    switch (a.d.nstr) {
    case 4: return launch_fast_t<2>(a, warps, grid, st);
    case 8: return launch_fast_t<4>(a, warps, grid, st);
    case 16: return launch_fast_t<8>(a, warps, grid, st);
    case 20: return launch_fast_t<10>(a, warps, grid, st);
    case 24: return launch_fast_t<12>(a, warps, grid, st);
    case 32: return launch_fast_t<16>(a, warps, grid, st);
    }
    return cudaErrorInvalidValue;
}

}  // namespace sbd
