// Flux kernel for NSTR = 4, 8, 16 (levels at the layer boundaries): one warp per bin, the
// boundary-value problem of DISORT solved layer by layer with reflection / transmission
// operators instead of the N L x N L band system (SETMTX / SOLVE0, disort.f:2702, :3322).
//
// Same discrete-ordinate equations, same delta-M scaling, truncation, sources and boundary
// conditions as the reference (tests/adding_model.py states the
// algorithm in numpy and tests/test_adding_math_cpu.py checks it against the CPU restatement of the reference).
//
// In the flux-weighted variables u^ = sqrt(w mu) u the sums s^ = u^+ + u^- and differences
// d^ = u^+ - u^- obey d s^/d tau = Po d^, d d^/d tau = Pe s^ with the symmetric operators of
// the reduced eigenproblem (disort.f:3221-3269).  With Po = L L^T, L^T Pe L = V K^2 V^T (the
// SVD of K^T L, as in sbd_fast.cu) and P = L V, a layer of scaled depth 2h has
//     R + T = I - 2 (I + P diag(coth(kh)/k) P^T)^-1,   R - T = I - 2 (I + P diag(tanh(kh)/k) P^T)^-1,
// evaluated in mode space:  M+- = P (P^T P + diag(k tanh(kh) | k coth(kh)))^-1 P^T,
//     R = -I + M+ + M-,   T = M+ - M-
// (two SPD inversions without pivoting; finite for k -> 0 and h -> 0).  Interaction principle:
//     u+_top = R u-_top + T u+_bot + s_up,     u-_bot = T u-_top + R u+_bot + s_dn
// with the sources from the particular solutions of UPBEAM / UPISOT (disort.f:4130, :4247).
//
//   phase 1  32/n layers at a time, n lanes per layer: eigenproblem, R, T, s_up, s_dn -> scratch
//   phase 2  bottom-up over the layers, the whole warp on one layer (2-D tiles of the n x n
//            matrices): reflection Rb / emission sb of everything below each interface,
//            (I - R Rb) Y = [T | R sb + s_dn] by Gauss-Jordan with partial pivoting in registers
//   phase 3  top-down: downward intensities d <- Y d + y, fluxes at every level as four dot
//            products (FLUXES, disort.f:1926-2006)
// Scratch per resident warp: L (2 n^2 + 2 n) doubles (38 KB at NSTR=16 / 33 layers; all
// resident warps together stay inside the L2).
#include <math.h>
#include <stdlib.h>

#include "sbd_internal.h"
#include "sbd_planck.cuh"
#include "sbd_devutil.cuh"

#ifndef SBD_ADD_ROLL
#define SBD_ADD_ROLL 1       // rolled Legendre / M loops: smaller code, +1.2 % on the C2 bench
#endif
#ifndef SBD_ADD_MEET
#define SBD_ADD_MEET 1       // phases 2 / 3: two adding sweeps that meet in the middle (NSTR <= 16)
#endif
#ifndef SBD_ADD_SYNC
#define SBD_ADD_SYNC 1       // CTA barriers between the phases
#endif
#ifndef SBD_ADD_NOPIV
#define SBD_ADD_NOPIV 0
#endif

namespace sbd {

#ifdef SBD_PHASE_TIMING
__device__ unsigned long long g_add_ticks[8];
#define ADD_TICK(i)                                                                     \
    do {                                                                                \
        if (threadIdx.x == 0) {                                                         \
            const long long tnow = clock64();                                           \
            atomicAdd(&g_add_ticks[i], (unsigned long long)(tnow - tphase));            \
            tphase = tnow;                                                              \
        }                                                                               \
    } while (0)
#else
#define ADD_TICK(i) do { } while (0)
#endif

template <int n>
struct AddLayout {
    static constexpr int N = 2 * n;
    // phase-1 record of a layer (global scratch): R[n][n], T[n][n], s_up[n], s_dn[n]
    static constexpr int r_R = 0, r_T = n * n, r_su = 2 * n * n, r_sd = 2 * n * n + n;
    static constexpr int rec = 2 * n * n + 2 * n;
    // phase-2 record (overwrites the phase-1 record of the layer): Y[n][n], y[n], and the flux
    // functionals of the layer's top interface: fu = D^T Rb, cu = c^T Rb, D^T sb, c^T sb
    static constexpr int o_Y = 0, o_y = n * n, o_fu = n * n + n, o_cu = n * n + 2 * n, o_f0 = n * n + 3 * n;
    static_assert(n * n + 3 * n + 2 <= rec, "phase-2 record must fit into the phase-1 record");
    // 2-D tiling of phase 2: n rows x CG column groups, CW columns per lane (n CG <= 32 lanes)
    static constexpr int CG = n >= 10 ? 2 : (n >= 4 ? 4 : n);
    static constexpr int CW = n / CG;
    // phase 1: a layer is owned by a group of GW lanes (the n first ones hold a row / column /
    // mode each; for n = 10, 12 the rest shadow lane n-1)
    // (n = 10: three groups of 10 lanes, lanes 30 and 31 shadow lane 29)
    static constexpr int GW = n <= 2 ? 2 : (n <= 4 ? 4 : (n <= 8 ? 8 : (n == 10 ? 10 : 16)));
    // phase-1 shared memory per layer group: gl[N], K, L, P, X [n][LD], 4 vectors.  Rows are n + 2
    // doubles apart: row-wise accesses of the group's lanes (stride 16 B x odd) hit different banks
    static constexpr int LD = n + 2;
    static constexpr int tasks = 32 / GW;
    static constexpr int task0 = N + 4 * n * LD + 4 * n;
    // padded to 4 (mod 16) doubles: the groups' areas start 8 banks apart, so the four
    // addresses of a group-wide broadcast load never share a bank
    static constexpr int task = ((task0 + 11) / 16) * 16 + 4;
    // phase 2 as two sweeps that meet in the middle (n <= 8: the work area holds two chains)
    static constexpr bool MEET = (SBD_ADD_MEET != 0) && n <= 8;
    // phase-2 shared memory per chain: Rb, sb, Y, y, W, w, two records
    static constexpr int p2 = (MEET ? 2 : 1) * (3 * n * n + 3 * n + 2 * rec);
    static constexpr int work = tasks * task > p2 ? tasks * task : p2;
    __host__ __device__ static size_t warp_doubles(int L)
    {
        // y0, work area, taucpr / tauc / beam transmissions (2), pk (+2 boundary temperatures),
        // three prologue work values per layer, level map
        // (the prologue's three work values per layer live in the work area, or behind it when the
        // atmosphere has more layers than the area can hold)
        const size_t lwx = (size_t)(3 * L > work ? 3 * L - work : 0);
        size_t d = (size_t)N + work + lwx + 4 * (L + 1) + (L + 3) + (L + 2) / 2 + 2;
        return (d + 1) & ~(size_t)1;
    }
    static constexpr int cta = 4 * n + N * n + 2;     // cmu cwt csq cd, ylm, sum(w mu), sum(w)
    __host__ __device__ static size_t slot_doubles(int L) { return (size_t)L * rec; }
};


template <int n>
__device__ __forceinline__ unsigned long long add_jacobi_partners(int g)
{
    constexpr int PB = n > 8 ? 4 : 3;          // bits per round
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
// phase 1: one layer per group of GW lanes; lane g < n holds row g / column g / mode g
// (gact = false: a shadow of lane n-1 that computes along and stores nothing)
// ---------------------------------------------------------------------------
template <int n>
__device__ __forceinline__ int phase1_adding(
    const double *__restrict__ dtauc, const double *__restrict__ ssalb,
    const double *__restrict__ pmom, int ldp, int lc, bool active,
    double fbeam, double umu0, bool plank,
    const double *cmu, const double *csq, const double *cd, const double *cylm,
    const double *y0, const double *ebeam, const double *pk,
    double *tsm, double *rec, int g, bool gact, int gbase, unsigned long long jpart)
{
    using AL = AddLayout<n>;
    constexpr int N = 2 * n, GW = AL::GW, PB = n > 8 ? 4 : 3, LD = AL::LD;
    double *sgl = tsm, *sK = sgl + N, *sL = sK + n * LD, *sP = sL + n * LD, *sX = sP + n * LD, *sv = sX + n * LD;

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

    // rows g of Pe and Po (even / odd Legendre sums; disort.f:3197-3216 in symmetric form)
    double pe[n], po[n];
#pragma unroll
    for (int j = 0; j < n; j++) { pe[j] = 0.0; po[j] = 0.0; }
#if SBD_ADD_ROLL
#pragma unroll 1
#else
#pragma unroll
#endif
    for (int l2 = 0; l2 < n; l2++) {
        const double te = sgl[2 * l2] * cylm[2 * l2 * n + g], to = sgl[2 * l2 + 1] * cylm[(2 * l2 + 1) * n + g];
#pragma unroll
        for (int j = 0; j < n; j++) {
            pe[j] = fma(te, cylm[2 * l2 * n + j], pe[j]);
            po[j] = fma(to, cylm[(2 * l2 + 1) * n + j], po[j]);
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

    // row Cholesky of both operators: Po = L L^T, Pe = K K^T (lane g = row g)
    int bad = 0;
    double rK[n], rL[n];
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
        // Pe is only semidefinite when w' -> 1: keep the factor real (a NaN pivot is a failure)
        const double floor_e = 1.0e-30;
        if (pive != pive) bad = 1;
        if (!(pive > floor_e)) pive = floor_e;
        const double rie = fast_rsqrt(pive), rio = fast_rsqrt(pivo);
        rK[j] = rie; rL[j] = rio;
        pe[j] = (g == j) ? pive * rie : ((g > j) ? nume * rie : 0.0);
        po[j] = (g == j) ? pivo * rio : ((g > j) ? numo * rio : 0.0);
    }
    if (gact) {
#pragma unroll
        for (int j = 0; j < n; j++) { sK[g * LD + j] = pe[j]; sL[g * LD + j] = po[j]; }
    }
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

    // one-sided Jacobi (round-robin pairing): A V = U S, S = the eigenvalues k (see sbd_fast.cu)
    if (n > 1) {
        double own2 = 0.0;
#pragma unroll
        for (int i = 0; i < n; i++) own2 = fma(a[i], a[i], own2);
        for (int sweep = 0; sweep < 40; sweep++) {
            int big = 0;
#pragma unroll 1
            for (int r = 0; r < n - 1; r++) {
                const int partner = (int)((jpart >> (PB * r)) & ((1u << PB) - 1u));
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
#ifdef SBD_PHASE_TIMING
            if ((threadIdx.x & 31) == 0) atomicAdd(&g_add_ticks[6], 1ull);
#endif
            if (!__any_sync(FULLMASK, big)) break;
            own2 = 0.0;
#pragma unroll
            for (int i = 0; i < n; i++) own2 = fma(a[i], a[i], own2);
        }
#ifdef SBD_PHASE_TIMING
        if ((threadIdx.x & 31) == 0) atomicAdd(&g_add_ticks[7], 1ull);
#endif
    }
    double s2 = 0.0;
#pragma unroll
    for (int i = 0; i < n; i++) s2 = fma(a[i], a[i], s2);
    const double kk = sqrt(s2);
    // column g of P = L V = K^-T (A V)
    double P[n];
#pragma unroll
    for (int i = n - 1; i >= 0; i--) {
        double acc = a[i];
#pragma unroll
        for (int k = i + 1; k < n; k++) acc = fma(-sK[k * LD + i], P[k], acc);
        P[i] = acc * rK[i];
    }
    if (gact) {
#pragma unroll
        for (int i = 0; i < n; i++) sP[g * LD + i] = P[i];     // [mode][direction]
    }

    // ---- particular solutions in the scaled variables (u^ = D u) ------------------------
    // beam (UPBEAM, disort.f:4130, spectral form): with c_j = (P^T r)_j / (1/mu0^2 - k_j^2),
    //   d^ = Po^-1 (P c),  s^ = mu0 (b_d - P c),  z^up = (s^ + d^)/2, z^dn = (s^ - d^)/2
    // thermal (UPISOT, disort.f:4247): p+-(tau) = D B(tau) +- b1 q^, q^ = Po^-1 D 1
    double zup = 0.0, zdn = 0.0;
    const bool beam = fbeam > 0.0;
    double tg = 0.0, bdg = 0.0;
    if (beam) {
        const double fac = fbeam * (1.0 / (4. * kPiRef));
        const double rmu0 = fast_rcp(umu0);
        double be = 0.0, bo = 0.0;
#pragma unroll
        for (int l = 0; l < N; l++) {
            const double t = sgl[l] * cylm[l * n + g] * y0[l];
            if (l & 1) bo += t; else be += t;
        }
        const double bs = 2.0 * fac * sqg * be;
        bdg = 2.0 * fac * sqg * bo;
        if (gact) sv[g] = bdg;
        __syncwarp();
        double t1 = 0.0;                                    // (K^T b_d)_g
#pragma unroll
        for (int k = 0; k < n; k++) t1 = fma(sK[k * LD + g], sv[k], t1);
        if (gact) sv[n + g] = t1;
        __syncwarp();
        double t2 = 0.0;                                    // (K K^T b_d)_g
#pragma unroll
        for (int k = 0; k < n; k++) t2 = fma(sK[g * LD + k], sv[n + k], t2);
        if (gact) sv[2 * n + g] = bs * rmu0 - t2;           // r_g
        __syncwarp();
        double cj = 0.0;
#pragma unroll
        for (int i = 0; i < n; i++) cj = fma(P[i], sv[2 * n + i], cj);
        cj = cj * fast_rcp(rmu0 * rmu0 - s2);
        if (gact) sv[3 * n + g] = cj;
        __syncwarp();
#pragma unroll
        for (int j = 0; j < n; j++) tg = fma(sP[j * LD + g], sv[3 * n + j], tg);     // (P c)_g
        if (gact) sv[g] = tg;
    } else {
        __syncwarp();
    }
    if (gact) sv[n + g] = cmu[g] * csq[g];                  // D_g
    __syncwarp();
    // Po^-1 [P c | D 1]: every lane runs both triangular solves (L y = b, L^T z = y)
    double zq[n];                                            // q^ (all entries, uniform in the group)
    {
        double y1[n], y2[n], z1[n];
#pragma unroll
        for (int i = 0; i < n; i++) {
            double a1 = sv[i], a2 = sv[n + i];
#pragma unroll
            for (int k = 0; k < i; k++) {
                const double lik = sL[i * LD + k];
                a1 = fma(-lik, y1[k], a1);
                a2 = fma(-lik, y2[k], a2);
            }
            y1[i] = a1 * rL[i]; y2[i] = a2 * rL[i];
        }
        double dg = 0.0;
#pragma unroll
        for (int i = n - 1; i >= 0; i--) {
            double a1 = y1[i], a2 = y2[i];
#pragma unroll
            for (int k = i + 1; k < n; k++) {
                const double lki = sL[k * LD + i];
                a1 = fma(-lki, z1[k], a1);
                a2 = fma(-lki, zq[k], a2);
            }
            z1[i] = a1 * rL[i]; zq[i] = a2 * rL[i];
            if (i == g) dg = z1[i];
        }
        if (beam) {
            const double sg = umu0 * (bdg - tg);
            zup = 0.5 * (sg + dg);
            zdn = 0.5 * (sg - dg);
        }
    }

    // ---- B+- = P^T P + diag(k tanh(kh) | k coth(kh)), row g (mode g) ---------------------
    double bp[n], bm[n];
    if constexpr (n > 8) {
        // large n: rolled loops (the code of this phase would not stay in the instruction cache);
        // the row goes through the X area because bp[j] cannot be indexed at run time
#pragma unroll 1
        for (int j = 0; j < n; j++) {
            double acc = 0.0;
#pragma unroll
            for (int i = 0; i < n; i++) acc = fma(P[i], sP[j * LD + i], acc);
            if (gact) sX[g * LD + j] = acc;
        }
        __syncwarp();
#pragma unroll
        for (int j = 0; j < n; j++) { bp[j] = sX[g * LD + j]; bm[j] = bp[j]; }
    } else {
#pragma unroll
        for (int j = 0; j < n; j++) {
            double acc = 0.0;
#pragma unroll
            for (int i = 0; i < n; i++) acc = fma(P[i], sP[j * LD + i], acc);
            bp[j] = acc; bm[j] = acc;
        }
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
    // in-place Gauss-Jordan inversion of both SPD matrices (no pivoting), row per lane.  The pivot
    // rows travel through shared memory (the owner stores 2n doubles, everyone loads them with
    // broadcast reads: half the shared-memory-pipe instructions of 2n double shuffles); two
    // buffers in the vector area, one warp barrier per step.
    __syncwarp();               // (the triangular solves above read the vector area)
#pragma unroll
    for (int j = 0; j < n; j++) {
        double *pb = sv + (j & 1) * 2 * n;
        if (gact && g == j) {
#pragma unroll
            for (int c = 0; c < n; c += 2) {
                reinterpret_cast<double2 *>(pb)[c / 2] = make_double2(bp[c], bp[c + 1]);
                reinterpret_cast<double2 *>(pb + n)[c / 2] = make_double2(bm[c], bm[c + 1]);
            }
        }
        __syncwarp();
        double pr[n], qr[n];
#pragma unroll
        for (int c = 0; c < n; c += 2) {
            const double2 u = reinterpret_cast<const double2 *>(pb)[c / 2], w = reinterpret_cast<const double2 *>(pb + n)[c / 2];
            pr[c] = u.x; pr[c + 1] = u.y; qr[c] = w.x; qr[c + 1] = w.y;
        }
        const double rp = fast_rcp(pr[j]), rq = fast_rcp(qr[j]);
        if (!(pr[j] > 0.0) || !(qr[j] > 0.0)) bad = 1;
        // row j (the pivot row, own row of lane j) becomes row / pivot = row - (1 - 1/pivot) row:
        // one update for every lane, with the multiplier 1 - 1/pivot on lane j
        const double mp = (g == j) ? 1.0 - rp : bp[j] * rp, mq = (g == j) ? 1.0 - rq : bm[j] * rq;
        const double dp = (g == j) ? rp : -mp, dq = (g == j) ? rq : -mq;
#pragma unroll
        for (int c = 0; c < n; c++) {
            bp[c] = (c == j) ? dp : fma(-mp, pr[c], bp[c]);
            bm[c] = (c == j) ? dq : fma(-mq, qr[c], bm[c]);
        }
    }
    __syncwarp();
    // U+- = B+-^-1 P^T (row g), stored over K and X
    {
        double up[n], um[n];
#pragma unroll
        for (int b = 0; b < n; b++) { up[b] = 0.0; um[b] = 0.0; }
        if constexpr (n > 8) {
            // rolled over j: the rows of the inverses wait in the K and X areas
            __syncwarp();       // everyone is done reading K (and X)
            if (gact) {
#pragma unroll
                for (int j = 0; j < n; j++) { sK[g * LD + j] = bp[j]; sX[g * LD + j] = bm[j]; }
            }
            __syncwarp();
#pragma unroll 1
            for (int j = 0; j < n; j++) {
                const double bpj = sK[g * LD + j], bmj = sX[g * LD + j];
#pragma unroll
                for (int b = 0; b < n; b++) {
                    const double pj = sP[j * LD + b];
                    up[b] = fma(bpj, pj, up[b]);
                    um[b] = fma(bmj, pj, um[b]);
                }
            }
        } else {
#pragma unroll
            for (int j = 0; j < n; j++) {
#pragma unroll
                for (int b = 0; b < n; b++) {
                    const double pj = sP[j * LD + b];
                    up[b] = fma(bp[j], pj, up[b]);
                    um[b] = fma(bm[j], pj, um[b]);
                }
            }
        }
        __syncwarp();           // everyone is done reading K
        if (gact) {
#pragma unroll
            for (int b = 0; b < n; b++) { sK[g * LD + b] = up[b]; sX[g * LD + b] = um[b]; }
        }
    }
    if (gact) { sv[g] = zup; sv[n + g] = zdn; }
    __syncwarp();
    // M+- = P U+- (row a = g): R = -I + M+ + M-, T = M+ - M-
    double R[n], T[n], mm[n];
    {
        double mp[n];
#pragma unroll
        for (int b = 0; b < n; b++) { mp[b] = 0.0; mm[b] = 0.0; }
#if SBD_ADD_ROLL
#pragma unroll 1
#else
#pragma unroll
#endif
        for (int j = 0; j < n; j++) {
            const double pa = sP[j * LD + g];
#pragma unroll
            for (int b = 0; b < n; b++) {
                mp[b] = fma(pa, sK[j * LD + b], mp[b]);
                mm[b] = fma(pa, sX[j * LD + b], mm[b]);
            }
        }
#pragma unroll
        for (int b = 0; b < n; b++) {
            R[b] = mp[b] + mm[b] - ((b == g) ? 1.0 : 0.0);
            T[b] = mp[b] - mm[b];
        }
    }
    // sources: s_up = p+_top - R p-_top - T p+_bot,  s_dn = p-_bot - T p-_top - R p+_bot
    double s_up = 0.0, s_dn = 0.0;
    if (beam) {
        double Rzd = 0.0, Tzu = 0.0, Tzd = 0.0, Rzu = 0.0;
#pragma unroll
        for (int b = 0; b < n; b++) {
            const double zu = sv[b], zd = sv[n + b];
            Rzd = fma(R[b], zd, Rzd); Tzu = fma(T[b], zu, Tzu);
            Tzd = fma(T[b], zd, Tzd); Rzu = fma(R[b], zu, Rzu);
        }
        const double et = ebeam[lc], eb = ebeam[lc + 1];
        s_up = et * (zup - Rzd) - eb * Tzu;
        s_dn = eb * (zdn - Rzu) - et * Tzd;
    }
    if (plank) {
        // the b1 q^ terms of the four boundary values combine to b1 (I + R - T) q^ = 2 b1 M- q^,
        // finite for dtau' -> 0 (b1 = dB / dtau' grows, M- ~ h shrinks)
        double b1 = 0.0;
        if (dtaucp > 1.0e-200) b1 = (pk[lc + 1] - pk[lc]) * fast_rcp(dtaucp);
        else if (dtaucp > 0.0) b1 = (pk[lc + 1] - pk[lc]) / dtaucp;
        double RD = 0.0, TD = 0.0, e = 0.0;
#pragma unroll
        for (int b = 0; b < n; b++) {
            const double Db = cmu[b] * csq[b];
            RD = fma(R[b], Db, RD); TD = fma(T[b], Db, TD);
            e = fma(mm[b], zq[b], e);
        }
        e *= 2.0 * b1;
        const double Dg = cmu[g] * csq[g];
        s_up += (Dg - RD) * pk[lc] - TD * pk[lc + 1] + e;
        s_dn += (Dg - RD) * pk[lc + 1] - TD * pk[lc] - e;
    }
    if (active && gact) {
#pragma unroll
        for (int b = 0; b < n; b += 2) {
            reinterpret_cast<double2 *>(rec + AL::r_R + g * n)[b / 2] = make_double2(R[b], R[b + 1]);
            reinterpret_cast<double2 *>(rec + AL::r_T + g * n)[b / 2] = make_double2(T[b], T[b + 1]);
        }
        rec[AL::r_su + g] = s_up;
        rec[AL::r_sd + g] = s_dn;
    }
    __syncwarp();
    return (bad && active) ? SBD_BIN_EIG_FAIL : 0;
}

// WARPS warps per CTA, 16 warps per SM; the warps of a CTA move through the phases together
// (one instruction stream in the I-cache at a time, see sbd_fast.cu).
// (n > 8: 4-warp CTAs, two per SM, up to 255 registers)
template <int n, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, n > 8 ? 2 : 16 / WARPS)
disort_adding_kernel(const LaunchArgs a)
{
    using AL = AddLayout<n>;
    constexpr int N = 2 * n, TASKS = AL::tasks, GW = AL::GW, CG = AL::CG, CW = AL::CW;
    constexpr int KB = n > 8 ? 4 : 3, KM = (1 << KB) - 1;       // pivot key: row index bits
    const int L = a.d.nlyr;
    const int NT = L + 1;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, warps = blockDim.x >> 5;
    const int ldp = a.d.nmom + 1;
    extern __shared__ __align__(16) double smem_add[];
    double *cmu = smem_add, *cwt = cmu + n, *csq = cwt + n, *cd = csq + n;
    double *cylm = cd + n + 2;          // cylm[-2] = sum(w mu), cylm[-1] = sum(w)
    double *wsm = smem_add + AL::cta + (size_t)warp * AL::warp_doubles(L);
    double *y0 = wsm;
    double *work = y0 + N;
    double *lw = work;                  // prologue only: 3 x L work values
    double *taucpr = work + (3 * L > AL::work ? 3 * L : AL::work), *tauc = taucpr + (L + 1);
    double *ebeam = tauc + (L + 1), *edir = ebeam + (L + 1);
    double *pk = edir + (L + 1);
    int *layru = (int *)(pk + (L + 3));

    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        double mu = a.quad[i], wt = a.quad[n + i];
        cmu[i] = mu; cwt[i] = wt; csq[i] = sqrt(wt / mu); cd[i] = sqrt(wt * mu);
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
    double *recs = a.scratch + (size_t)slot * a.slot_stride;      // [L][rec]
    // layer group of this lane; lanes beyond the last full group shadow its last lane
    const int task = lane / GW < TASKS ? lane / GW : TASKS - 1;
    const bool gact = lane < TASKS * GW && (lane % GW) < n;
    const int g = gact ? lane % GW : n - 1, gbase = task * GW;
    const int nbins_all = a.nbins_dev ? *a.nbins_dev : a.d.nbins;
    double *tsm = work + (size_t)task * AL::task;
    const unsigned long long jpart = add_jacobi_partners<n>(g);
    // phase-2 tiling: row i2, column group p2 (lanes beyond n*CG shadow the last row)
    const bool act2 = lane < n * CG;
    const int i2 = act2 ? lane / CG : n - 1, p2 = lane % CG, c0 = p2 * CW;
    // phase-2 shared memory
    double *sRb = work, *ssb = sRb + n * n, *sY = ssb + n, *sy = sY + n * n, *sW = sy + n, *sw = sW + n * n;
    double *rbuf = sw + n;

    for (;;) {
        int bin = 0;
        if (lane == 0) bin = atomicAdd(a.work_counter, 1);
        bin = __shfl_sync(FULLMASK, bin, 0);
        const bool have = bin < nbins_all;
        if (!__syncthreads_or(have)) break;
#ifdef SBD_PHASE_TIMING
        long long tphase = clock64();
#endif
        const int src = !have ? 0 : (a.binmap ? a.binmap[bin] : bin);
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
                if (!(fabs(dtauc[lc]) <= 1.79e308)) badl = 1;
                for (int k = 1; k <= a.d.nmom; k++) {
                    double pm = pmom[(size_t)lc * ldp + k];
                    if (!(pm >= -1.0 && pm <= 1.0)) badl = 1;
                }
            }
            if (fbeam < 0.0 || (fbeam > 0.0 && !(umu0 > 0.0 && umu0 <= 1.0))) badl = 1;
            // albedo = SBD_SURFACE(s): BRDF surface s (LAMBER = .FALSE.)
            if (albedo < 0.0) {
                surf = (int)(-albedo) - 1;
                if (!a.sf_bdr || surf >= a.sf_count || (double)(surf + 1) != -albedo) badl = 1;
            } else if (!(albedo <= 1.0)) badl = 1;
            if (bp.fisot < 0.0) badl = 1;
            if (plank && (bp.wvnmlo < 0.0 || bp.wvnmhi <= bp.wvnmlo || bp.temis < 0.0 ||
                          bp.temis > 1.0 || bp.btemp < 0.0 || bp.ttemp < 0.0)) badl = 1;
            if (plank && (!a.temper || bp.col < 0 || bp.col >= a.d.ncol)) badl = 1;
            if (__any_sync(FULLMASK, badl)) status = SBD_BIN_BAD_INPUT;
            int clash = 0;
            if (fbeam > 0.0 && lane < n && fabs(umu0 - cmu[lane]) / umu0 < 1.e-4) clash = 1;
            if (!status && __any_sync(FULLMASK, clash)) status = SBD_BIN_ANGLE_CLASH;
        }
        int ncut = L, lyrcut = 0, monotone = 1;
        // SETDIS prologue (disort.f:2546-2605), accumulated in the reference's order by lane 0
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
        __syncwarp();
        // Negative optical depths (legal upstream, taugas.f:7485) make TAUC non-monotone
        // (disort.f:487 accumulates before CHEKIN clips), and the reference then evaluates
        // levels INSIDE earlier layers: such bins go to the elimination kernel (sbd_fast.cu),
        // which keeps the layer solutions.
        if (!monotone && !status && have) {
            if (lane == 0) {
                const int k = atomicAdd(a.redo_count, 1);
                a.redo_list[k] = bin;
            }
            status = -100;          // parked: no phases, no status write
        }
        for (int lev = lane; lev <= L; lev += 32) {
            ebeam[lev] = fbeam > 0.0 ? exp(-taucpr[lev] / umu0) : 0.0;
            edir[lev] = fbeam > 0.0 ? exp(-tauc[lev] / umu0) : 0.0;
        }
        // the level belongs to the first layer whose interval contains it (disort.f:2610-2625)
        for (int lu = lane; lu < NT; lu += 32) {
            const double ut = tauc[lu];
            int lc;
            if (lu >= 1 && tauc[lu - 1] < ut) lc = lu;
            else {
                for (lc = 1; lc <= L; lc++)
                    if (ut >= tauc[lc - 1] && ut <= tauc[lc]) break;
                if (lc > L) lc = L;
            }
            layru[lu] = lc;
        }
        double tplank = 0.0, bplank = 0.0;
        if (plank && !status) {
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
        if (status != -100)
            for (int lu = lane; lu < NT; lu += 32) {
                if (o_rfldir) o_rfldir[lu] = 0.0;
                if (o_rfldn) o_rfldn[lu] = 0.0;
                if (o_flup) o_flup[lu] = 0.0;
                if (o_dfdt) o_dfdt[lu] = 0.0;
                if (o_uavg) o_uavg[lu] = 0.0;
            }
        __syncwarp();
        ADD_TICK(0);

        // ===================== phase 1 =====================================
        if (!status) {
            for (int lc0 = 0; lc0 < ncut; lc0 += TASKS) {
                int lc = lc0 + task;
                const bool active = lc < ncut;
                if (!active) lc = ncut - 1;
                int st = phase1_adding<n>(dtauc, ssalb, pmom, ldp, lc, active, fbeam, umu0, plank,
                                          cmu, csq, cd, cylm, y0, ebeam, pk, tsm,
                                          recs + (size_t)lc * AL::rec, g, gact, gbase, jpart);
                if (__any_sync(FULLMASK, st != 0)) { status = SBD_BIN_EIG_FAIL; break; }
            }
        }
        __syncwarp();
        __threadfence_block();
        if (SBD_ADD_SYNC) __syncthreads();
        ADD_TICK(1);

        // ===================== phase 2: bottom-up adding ===================
        // functionals of the bottom interface (level ncut), kept for phase 3 by lanes n, n+1
        double fnB[n], fnB0 = 0.0;
#pragma unroll
        for (int c = 0; c < n; c++) fnB[c] = 0.0;
        if (!status) {
            // bottom boundary (disort.f:2919-2990, :3552-3578): Lambertian or BRDF surface (m = 0
            // tables of SURFAC), or no upwelling radiation at the truncation level.  In the scaled
            // variables: Rb = 2 D BDR D, sb = D (BDR(:,0) mu0 F / pi E + BEM B(Ts))
            if (surf >= 0 && !lyrcut) {
                const double *bdr = a.sf_bdr + (size_t)surf * a.sf_modes * n * (n + 1);
                const double *bem = a.sf_bem + (size_t)surf * n;
                const double beamfac = umu0 * fbeam / kPiRef * ebeam[ncut];
                for (int e = lane; e < n * n; e += 32)
                    sRb[e] = 2.0 * cd[e / n] * bdr[(e / n) * (n + 1) + 1 + e % n] * cd[e % n];
                if (lane < n) ssb[lane] = cd[lane] * (bdr[lane * (n + 1)] * beamfac + bem[lane] * bplank);
                __syncwarp();
                if (lane == n || lane == n + 1) {
                    const double *wv = (lane == n) ? cd : csq;
                    fnB0 = 0.0;
#pragma unroll
                    for (int c = 0; c < n; c++) {
                        double acc = 0.0;
                        for (int k = 0; k < n; k++) acc = fma(wv[k], sRb[k * n + c], acc);
                        fnB[c] = acc;
                        fnB0 = fma(wv[c], ssb[c], fnB0);
                    }
                }
            } else {
                const double refl = lyrcut ? 0.0 : 2.0 * albedo;
                const double emis = lyrcut ? 0.0 : albedo * umu0 * fbeam / kPiRef * ebeam[ncut] + (1.0 - albedo) * bplank;
                for (int e = lane; e < n * n; e += 32) sRb[e] = refl * cd[e / n] * cd[e % n];
                if (lane < n) ssb[lane] = cd[lane] * emis;
                if (lane == n || lane == n + 1) {
                    const double s = (lane == n) ? Wq : SWq;     // D^T D, c^T D
#pragma unroll
                    for (int c = 0; c < n; c++) fnB[c] = refl * s * cd[c];
                    fnB0 = s * emis;
                }
            }
            if constexpr (AL::MEET) {
            // Two adding sweeps that meet in the middle: chain 0 runs bottom-up over the layers
            // ncut-1 .. mid (Rb, sb: reflection / emission of everything below the layer's top
            // interface), chain 1 top-down over the layers 0 .. mid-1 (the mirror image: Ra, sa of
            // everything above the layer's bottom interface; the same step with s_up and s_dn
            // exchanged).  The chains are independent: every lane carries both, which doubles the
            // instruction-level parallelism of the sequential part and halves its length.
            constexpr int CH = 3 * n * n + 3 * n;
            double *const chs[2] = { work, work + CH };
            double *const rb2[2] = { work + 2 * CH, work + 2 * CH + 2 * AL::rec };
            const int mid = ncut / 2, niter = ncut - mid;
            {   // nothing above the top boundary reflects; it emits D (fisot + tplank)
                double *R1 = chs[1], *s1 = R1 + n * n;
                for (int e = lane; e < n * n; e += 32) R1[e] = 0.0;
                if (lane < n) s1[lane] = cd[lane] * (bp.fisot + tplank);
            }
            warp_copy_async(rb2[0], recs + (size_t)(ncut - 1) * AL::rec, AL::rec, lane);
            if (mid > 0) warp_copy_async(rb2[1], recs, AL::rec, lane);
            cp_async_commit();
            __syncwarp();
            for (int k = 0; k < niter; k++) {
                const int buf = k & 1;
                const int lq[2] = { ncut - 1 - k, k };
                const bool actq[2] = { true, k < mid };
                if (k + 1 < niter) {
                    warp_copy_async(rb2[0] + (buf ^ 1) * AL::rec, recs + (size_t)(lq[0] - 1) * AL::rec, AL::rec, lane);
                    if (k + 1 < mid)
                        warp_copy_async(rb2[1] + (buf ^ 1) * AL::rec, recs + (size_t)(lq[1] + 1) * AL::rec, AL::rec, lane);
                    cp_async_commit();
                    cp_async_wait_one();
                } else {
                    cp_async_wait_all();
                }
                __syncwarp();
                // ---- B = I - R Rx (my CW columns of row i2), t = T, v = (R sx + s)_i2 ----
                double b[2][CW], t[2][CW], v[2];
#pragma unroll
                for (int q = 0; q < 2; q++) {
                    const double *rc = rb2[q] + buf * AL::rec;
                    const double *Rr = rc + AL::r_R + i2 * n, *Tr = rc + AL::r_T + i2 * n;
                    const double *Rx = chs[q], *sx = Rx + n * n;
                    double r[n];
#pragma unroll
                    for (int kk = 0; kk < n; kk++) r[kk] = Rr[kk];
#pragma unroll
                    for (int s = 0; s < CW; s++) b[q][s] = (c0 + s == i2) ? 1.0 : 0.0;
                    v[q] = rc[(q == 0 ? AL::r_sd : AL::r_su) + i2];
#pragma unroll
                    for (int kk = 0; kk < n; kk++) {
#pragma unroll
                        for (int s = 0; s < CW; s++) b[q][s] = fma(-r[kk], Rx[kk * n + c0 + s], b[q][s]);
                        v[q] = fma(r[kk], sx[kk], v[q]);
                    }
#pragma unroll
                    for (int s = 0; s < CW; s++) t[q][s] = Tr[c0 + s];
                }
                // ---- Gauss-Jordan with partial pivoting on [B | T | v] of both chains ----
                unsigned used[2] = { 0u, 0u };
                int myj[2] = { 0, 0 }, sing = 0;
                double myrp[2] = { 0.0, 0.0 };
#pragma unroll
                for (int j = 0; j < n; j++) {
                    const int pj = j / CW, sj = j % CW;
#pragma unroll
                    for (int q = 0; q < 2; q++) {
                        const double colv = __shfl_sync(FULLMASK, b[q][sj], (lane & ~(CG - 1)) | pj);
                        const int key = (act2 && !((used[q] >> i2) & 1u))
                                            ? ((__double2hiint(colv) & ~KM & 0x7fffffff) | (KM - i2)) : -1;
                        const int mx = __reduce_max_sync(FULLMASK, key);
                        if (actq[q] && (mx >> KB) <= 0) sing = 1;
                        const int ip = KM - (mx & KM);
                        used[q] |= 1u << ip;
                        const int srcl = ip * CG + p2;
                        const double rp = fast_rcp(__shfl_sync(FULLMASK, colv, ip * CG));
                        const double m = (i2 == ip) ? 0.0 : colv * rp;
                        if (i2 == ip) { myj[q] = j; myrp[q] = rp; }
#pragma unroll
                        for (int s = 0; s < CW; s++) {
                            b[q][s] = fma(-m, __shfl_sync(FULLMASK, b[q][s], srcl), b[q][s]);
                            t[q][s] = fma(-m, __shfl_sync(FULLMASK, t[q][s], srcl), t[q][s]);
                        }
                        v[q] = fma(-m, __shfl_sync(FULLMASK, v[q], srcl), v[q]);
                    }
                }
                if (sing) { status = SBD_BIN_SINGULAR; break; }
                // row i2 now holds row myj of Y = (I - R Rx)^-1 T and of y
#pragma unroll
                for (int q = 0; q < 2; q++) {
                    double *sY = chs[q] + n * n + n, *sy = sY + n * n;
                    double *orec = recs + (size_t)lq[q] * AL::rec;       // (the phase-1 record is in shared memory)
                    if (act2 && actq[q]) {
#pragma unroll
                        for (int s = 0; s < CW; s++) {
                            const double y = t[q][s] * myrp[q];
                            sY[myj[q] * n + c0 + s] = y; orec[AL::o_Y + myj[q] * n + c0 + s] = y;
                        }
                        if (p2 == 0) { const double y = v[q] * myrp[q]; sy[myj[q]] = y; orec[AL::o_y + myj[q]] = y; }
                    }
                }
                __syncwarp();
                // ---- W = Rx [Y | y] + [0 | sx] ----
#pragma unroll
                for (int q = 0; q < 2; q++) {
                    const double *Rx = chs[q], *sx = Rx + n * n, *sY = sx + n, *sy = sY + n * n;
                    double *sW = chs[q] + 2 * n * n + 2 * n, *sw = sW + n * n;
                    double rb[n], w[CW], wv = sx[i2];
#pragma unroll
                    for (int kk = 0; kk < n; kk++) rb[kk] = Rx[i2 * n + kk];
#pragma unroll
                    for (int s = 0; s < CW; s++) w[s] = 0.0;
#pragma unroll
                    for (int kk = 0; kk < n; kk++) {
#pragma unroll
                        for (int s = 0; s < CW; s++) w[s] = fma(rb[kk], sY[kk * n + c0 + s], w[s]);
                        wv = fma(rb[kk], sy[kk], wv);
                    }
                    if (act2 && actq[q]) {
#pragma unroll
                        for (int s = 0; s < CW; s++) sW[i2 * n + c0 + s] = w[s];
                        if (p2 == 0) sw[i2] = wv;
                    }
                }
                __syncwarp();
                // ---- [Rx | sx] <- [R | s'] + T W ----
#pragma unroll
                for (int q = 0; q < 2; q++) {
                    const double *rc = rb2[q] + buf * AL::rec;
                    const double *Rr = rc + AL::r_R + i2 * n, *Tr = rc + AL::r_T + i2 * n;
                    double *Rx = chs[q], *sx = Rx + n * n;
                    const double *sW = chs[q] + 2 * n * n + 2 * n, *sw = sW + n * n;
                    double tr[n], nr[CW], ns = rc[(q == 0 ? AL::r_su : AL::r_sd) + i2];
#pragma unroll
                    for (int kk = 0; kk < n; kk++) tr[kk] = Tr[kk];
#pragma unroll
                    for (int s = 0; s < CW; s++) nr[s] = Rr[c0 + s];
#pragma unroll
                    for (int kk = 0; kk < n; kk++) {
#pragma unroll
                        for (int s = 0; s < CW; s++) nr[s] = fma(tr[kk], sW[kk * n + c0 + s], nr[s]);
                        ns = fma(tr[kk], sw[kk], ns);
                    }
                    if (act2 && actq[q]) {
#pragma unroll
                        for (int s = 0; s < CW; s++) Rx[i2 * n + c0 + s] = nr[s];
                        if (p2 == 0) sx[i2] = ns;
                    }
                }
                __syncwarp();
                // ---- flux functionals of the chains' new interfaces: D^T Rx, c^T Rx, D^T sx, c^T sx ----
                for (int e = lane; e < 2 * (2 * n + 2); e += 32) {
                    const int q = e >= 2 * n + 2 ? 1 : 0, f = e - q * (2 * n + 2);
                    if (q == 1 && k >= mid) continue;
                    const double *Rx = work + q * CH, *sx = Rx + n * n;
                    double *orec = recs + (size_t)(q ? k : ncut - 1 - k) * AL::rec;
                    double acc = 0.0;
                    if (f < 2 * n) {
                        const double *wv = (f < n) ? cd : csq;
                        const int c = f % n;
#pragma unroll
                        for (int kk = 0; kk < n; kk++) acc = fma(wv[kk], Rx[kk * n + c], acc);
                        orec[(f < n ? AL::o_fu : AL::o_cu) + c] = acc;
                    } else {
                        const double *wv = (f == 2 * n) ? cd : csq;
#pragma unroll
                        for (int kk = 0; kk < n; kk++) acc = fma(wv[kk], sx[kk], acc);
                        orec[AL::o_f0 + (f - 2 * n)] = acc;
                    }
                }
            }
            // ---- the chains meet at interface `mid`: (I - Ra Rb) d = Ra sb + sa, u = Rb d + sb ----
            if (!status) {
                __syncwarp();
                const double *Rb = chs[0], *sb = Rb + n * n, *Ra = chs[1], *sa = Ra + n * n;
                double *dm = chs[0] + n * n + n;                 // Y area of chain 0: d_mid, then u_mid
                double b[CW], v, ra[n];
#pragma unroll
                for (int kk = 0; kk < n; kk++) ra[kk] = Ra[i2 * n + kk];
#pragma unroll
                for (int s = 0; s < CW; s++) b[s] = (c0 + s == i2) ? 1.0 : 0.0;
                v = sa[i2];
#pragma unroll
                for (int kk = 0; kk < n; kk++) {
#pragma unroll
                    for (int s = 0; s < CW; s++) b[s] = fma(-ra[kk], Rb[kk * n + c0 + s], b[s]);
                    v = fma(ra[kk], sb[kk], v);
                }
                unsigned used = 0u;
                int myj = 0, sing = 0;
                double myrp = 0.0;
#pragma unroll
                for (int j = 0; j < n; j++) {
                    const int pj = j / CW, sj = j % CW;
                    const double colv = __shfl_sync(FULLMASK, b[sj], (lane & ~(CG - 1)) | pj);
                    const int key = (act2 && !((used >> i2) & 1u))
                                        ? ((__double2hiint(colv) & ~KM & 0x7fffffff) | (KM - i2)) : -1;
                    const int mx = __reduce_max_sync(FULLMASK, key);
                    if ((mx >> KB) <= 0) sing = 1;
                    const int ip = KM - (mx & KM);
                    used |= 1u << ip;
                    const int srcl = ip * CG + p2;
                    const double rp = fast_rcp(__shfl_sync(FULLMASK, colv, ip * CG));
                    const double m = (i2 == ip) ? 0.0 : colv * rp;
                    if (i2 == ip) { myj = j; myrp = rp; }
#pragma unroll
                    for (int s = 0; s < CW; s++) b[s] = fma(-m, __shfl_sync(FULLMASK, b[s], srcl), b[s]);
                    v = fma(-m, __shfl_sync(FULLMASK, v, srcl), v);
                }
                if (sing) status = SBD_BIN_SINGULAR;
                if (act2 && p2 == 0) dm[myj] = v * myrp;
                __syncwarp();
                if (lane < n) {
                    double u = sb[lane];
#pragma unroll
                    for (int kk = 0; kk < n; kk++) u = fma(Rb[lane * n + kk], dm[kk], u);
                    dm[n + lane] = u;
                }
                __syncwarp();
            }
            } else {
            warp_copy_async(rbuf, recs + (size_t)(ncut - 1) * AL::rec, AL::rec, lane);
            cp_async_commit();
            __syncwarp();
            for (int lc = ncut - 1; lc >= 0; lc--) {
                const int buf = (ncut - 1 - lc) & 1;
                if (lc > 0) {
                    warp_copy_async(rbuf + (buf ^ 1) * AL::rec, recs + (size_t)(lc - 1) * AL::rec, AL::rec, lane);
                    cp_async_commit();
                    cp_async_wait_one();
                } else {
                    cp_async_wait_all();
                }
                __syncwarp();
                const double *rc = rbuf + buf * AL::rec;
                const double *Rr = rc + AL::r_R + i2 * n, *Tr = rc + AL::r_T + i2 * n;
                // ---- B = I - R Rb (my CW columns of row i2), t = T, v = (R sb + s_dn)_i2 ----
                double b[CW], t[CW], v;
                {
                    double r[n];
#pragma unroll
                    for (int k = 0; k < n; k++) r[k] = Rr[k];
#pragma unroll
                    for (int s = 0; s < CW; s++) b[s] = (c0 + s == i2) ? 1.0 : 0.0;
                    v = rc[AL::r_sd + i2];
#pragma unroll
                    for (int k = 0; k < n; k++) {
#pragma unroll
                        for (int s = 0; s < CW; s++) b[s] = fma(-r[k], sRb[k * n + c0 + s], b[s]);
                        v = fma(r[k], ssb[k], v);
                    }
#pragma unroll
                    for (int s = 0; s < CW; s++) t[s] = Tr[c0 + s];
                }
                // ---- Gauss-Jordan with partial pivoting on [B | T | v], rows in registers ----
                unsigned used = 0;
                int myj = 0, sing = 0;
                double myrp = 0.0;
#pragma unroll
                for (int j = 0; j < n; j++) {
                    const int pj = j / CW, sj = j % CW;
                    const double colv = __shfl_sync(FULLMASK, b[sj], (lane & ~(CG - 1)) | pj);
#if SBD_ADD_NOPIV
                    const int ip = j;
                    (void)used;
#else
                    const int key = (act2 && !((used >> i2) & 1u))
                                        ? ((__double2hiint(colv) & ~KM & 0x7fffffff) | (KM - i2)) : -1;
                    const int mx = __reduce_max_sync(FULLMASK, key);
                    if ((mx >> KB) <= 0) sing = 1;
                    const int ip = KM - (mx & KM);
                    used |= 1u << ip;
#endif
                    const int srcl = ip * CG + p2;
                    // rows keep their scale: multiplier = column entry / pivot (0 on the pivot row
                    // itself); the pivot row is divided by its pivot once, after the last step
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
#pragma unroll
                for (int s = 0; s < CW; s++) t[s] *= myrp;
                v *= myrp;
                if (sing) { status = SBD_BIN_SINGULAR; break; }
                // row i2 now holds row myj of Y = (I - R Rb)^-1 T and of y
                double *orec = recs + (size_t)lc * AL::rec;      // (the phase-1 record is in shared memory)
                if (act2) {
#pragma unroll
                    for (int s = 0; s < CW; s++) { sY[myj * n + c0 + s] = t[s]; orec[AL::o_Y + myj * n + c0 + s] = t[s]; }
                    if (p2 == 0) { sy[myj] = v; orec[AL::o_y + myj] = v; }
                }
                __syncwarp();
                // ---- W = Rb [Y | y] + [0 | sb] ----
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
                // ---- [Rb | sb] <- [R | s_up] + T W ----
                {
                    double tr[n], nr[CW], ns = rc[AL::r_su + i2];
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
                        for (int s = 0; s < CW; s++) sRb[i2 * n + c0 + s] = nr[s];
                        if (p2 == 0) ssb[i2] = ns;
                    }
                }
                __syncwarp();
                // ---- flux functionals of interface lc: D^T Rb, c^T Rb, D^T sb, c^T sb ----
                for (int e = lane; e < 2 * n + 2; e += 32) {
                    if (e < 2 * n) {
                        const double *wv = (e < n) ? cd : csq;
                        const int c = e % n;
                        double acc = 0.0;
#pragma unroll
                        for (int k = 0; k < n; k++) acc = fma(wv[k], sRb[k * n + c], acc);
                        orec[(e < n ? AL::o_fu : AL::o_cu) + c] = acc;
                    } else {
                        const double *wv = (e == 2 * n) ? cd : csq;
                        double acc = 0.0;
#pragma unroll
                        for (int k = 0; k < n; k++) acc = fma(wv[k], ssb[k], acc);
                        orec[AL::o_f0 + (e - 2 * n)] = acc;
                    }
                }
            }
                    }
        }
        cp_async_wait_all();
        __syncwarp();
        __threadfence_block();
        if (SBD_ADD_SYNC) __syncthreads();
        ADD_TICK(2);

        // ===================== phase 3: intensities at the interfaces + fluxes ======
        if constexpr (AL::MEET) {
        // From interface `mid` the downward intensities are carried down (lanes 0..15: d <- Y d + y
        // with the records of chain 0) and the upward intensities up (lanes 16..31: u <- Y' u + y'
        // with the records of chain 1) at the same time.  In each half, lane r < n holds row r of the
        // propagator, lanes n .. n+3 the four flux functionals of the level; every lane forms s0 + vv . x.
        if (!status) {
            const int half = lane >> 4, hl = lane & 15, hb = lane & 16;
            const int role = hl - n;
            const int mid = ncut / 2, niter = ncut - mid;
            const double *dm = work + n * n + n;
            const double top0 = bp.fisot + tplank;
            double x[n], vv[n], s0 = 0.0;
#pragma unroll
            for (int c = 0; c < n; c++) x[c] = dm[half * n + c];
            // record of iteration t: chain 0: layer mid + t (the surface at level ncut), chain 1: layer
            // mid - t - 1 (the top boundary at level 0: nothing reflects, D (fisot + tplank) comes down)
            auto load_rec = [&](int t) {
                const int lyr = half == 0 ? mid + t : mid - t - 1;
                const bool inrange = half == 0 ? lyr < ncut : lyr >= 0;
                if (inrange) {
                    const double *orec = recs + (size_t)lyr * AL::rec;
                    if (hl < n) {
#pragma unroll
                        for (int c = 0; c < n; c += 2) {
                            const double2 q = reinterpret_cast<const double2 *>(orec + AL::o_Y + hl * n)[c / 2];
                            vv[c] = q.x; vv[c + 1] = q.y;
                        }
                        s0 = orec[AL::o_y + hl];
                    } else if (role == 0 || role == 1) {
#pragma unroll
                        for (int c = 0; c < n; c += 2) {
                            const double2 q = reinterpret_cast<const double2 *>(orec + (role == 0 ? AL::o_fu : AL::o_cu))[c / 2];
                            vv[c] = q.x; vv[c + 1] = q.y;
                        }
                        s0 = orec[AL::o_f0 + role];
                    }
                } else if (role == 0 || role == 1) {
                    if (half == 0) {
#pragma unroll
                        for (int c = 0; c < n; c++) vv[c] = fnB[c];
                        s0 = fnB0;
                    } else {
#pragma unroll
                        for (int c = 0; c < n; c++) vv[c] = 0.0;
                        s0 = (role == 0 ? Wq : SWq) * top0;
                    }
                }
            };
#pragma unroll
            for (int c = 0; c < n; c++) vv[c] = (role == 2) ? cd[c] : ((role == 3) ? csq[c] : 0.0);
            load_rec(0);
            for (int t = 0; t <= niter; t++) {
                const int lev = half == 0 ? mid + t : mid - t;
                double xs = s0;
#pragma unroll
                for (int c = 0; c < n; c++) xs = fma(vv[c], x[c], xs);
                if (t < niter) load_rec(t + 1);               // next record while this level finishes
                // x_R: the reflected side (u+ below mid, u- above), x_V: the carried side
                const double xcr = __shfl_sync(FULLMASK, xs, hb | (n + 1));
                const double xdv = __shfl_sync(FULLMASK, xs, hb | (n + 2));
                const double xcv = __shfl_sync(FULLMASK, xs, hb | (n + 3));
                if (hl == n && (half == 0 || (t > 0 && lev >= 0))) {
                    const double pi = kPiRef;
                    const double fact = ebeam[lev];
                    const double dirint = fbeam * fact;
                    const double fldir = umu0 * (fbeam * fact);
                    const double rfldir = umu0 * fbeam * edir[lev];
                    const double flup = 2. * pi * (half == 0 ? xs : xdv), fldn = 2. * pi * (half == 0 ? xdv : xs);
                    const double fdntot = fldn + fldir;
                    constexpr double inv4pi = 1.0 / (4. * kPiRef);
                    const double uavg = (2. * pi * (xcr + xcv) + dirint) * inv4pi;
                    // the layer the level belongs to (its single-scattering albedo and Planck value)
                    const int lyr = layru[lev] - 1;
                    double ssl = ssalb[lyr];
                    if (ssl == 1.0) ssl = 1.0 - kDither;
                    double plsorc = 0.0;
                    if (plank) plsorc = (lev == 0) ? pk[0] : ((taucpr[lyr + 1] - taucpr[lyr] > 0.0) ? pk[lyr + 1] : pk[lyr]);
                    if (o_rfldir) o_rfldir[lev] = rfldir;
                    if (o_rfldn) o_rfldn[lev] = fdntot - rfldir;
                    if (o_flup) o_flup[lev] = flup;
                    if (o_uavg) o_uavg[lev] = uavg;
                    if (o_dfdt) o_dfdt[lev] = (1.0 - ssl) * 4. * pi * (uavg - plsorc);
                }
                // the new intensities to everyone in the half
#pragma unroll
                for (int c = 0; c < n; c++) x[c] = __shfl_sync(FULLMASK, xs, hb | c);
            }
        }
        } else {
        // top-down: lane r < n: row r of Y (new downward intensity r); lanes n .. n+3: the four flux
        // functionals (D^T u+, c^T u+, D^T u-, c^T u-); every lane forms  s0 + vv . d
        if (!status) {
            double d[n], vv[n], s0 = 0.0;
            const int role = lane - n;
#pragma unroll
            for (int c = 0; c < n; c++) d[c] = cd[c] * (bp.fisot + tplank);
            auto load_rec = [&](int lev) {
                // record `lev` = layer lev+1 (top interface = level lev); lev == ncut: the boundary
                if (lev < ncut) {
                    const double *orec = recs + (size_t)lev * AL::rec;
                    if (lane < n) {
#pragma unroll
                        for (int c = 0; c < n; c += 2) {
                            const double2 q = reinterpret_cast<const double2 *>(orec + AL::o_Y + lane * n)[c / 2];
                            vv[c] = q.x; vv[c + 1] = q.y;
                        }
                        s0 = orec[AL::o_y + lane];
                    } else if (role < 2) {
#pragma unroll
                        for (int c = 0; c < n; c += 2) {
                            const double2 q = reinterpret_cast<const double2 *>(orec + (role == 0 ? AL::o_fu : AL::o_cu))[c / 2];
                            vv[c] = q.x; vv[c + 1] = q.y;
                        }
                        s0 = orec[AL::o_f0 + role];
                    }
                } else if (role == 0 || role == 1) {
#pragma unroll
                    for (int c = 0; c < n; c++) vv[c] = fnB[c];
                    s0 = fnB0;
                }
            };
#pragma unroll
            for (int c = 0; c < n; c++) vv[c] = (role == 2) ? cd[c] : ((role == 3) ? csq[c] : 0.0);
            load_rec(0);
            for (int lev = 0; lev <= ncut; lev++) {
                double x = s0;
#pragma unroll
                for (int c = 0; c < n; c++) x = fma(vv[c], d[c], x);
                if (lev < ncut) load_rec(lev + 1);           // next record while this level finishes
                const double xcu = __shfl_sync(FULLMASK, x, n + 1);
                const double xdn = __shfl_sync(FULLMASK, x, n + 2);
                const double xcd = __shfl_sync(FULLMASK, x, n + 3);
                if (lane == n) {
                    const double pi = kPiRef;
                    const double fact = ebeam[lev];
                    const double dirint = fbeam * fact;
                    const double fldir = umu0 * (fbeam * fact);
                    const double rfldir = umu0 * fbeam * edir[lev];
                    const double flup = 2. * pi * x, fldn = 2. * pi * xdn;
                    const double fdntot = fldn + fldir;
                    constexpr double inv4pi = 1.0 / (4. * kPiRef);
                    const double uavg = (2. * pi * (xcu + xcd) + dirint) * inv4pi;
                    // the layer the level belongs to (its single-scattering albedo and Planck value)
                    const int lyr = layru[lev] - 1;
                    double ssl = ssalb[lyr];
                    if (ssl == 1.0) ssl = 1.0 - kDither;
                    double plsorc = 0.0;
                    if (plank) plsorc = (lev == 0) ? pk[0] : ((taucpr[lyr + 1] - taucpr[lyr] > 0.0) ? pk[lyr + 1] : pk[lyr]);
                    if (o_rfldir) o_rfldir[lev] = rfldir;
                    if (o_rfldn) o_rfldn[lev] = fdntot - rfldir;
                    if (o_flup) o_flup[lev] = flup;
                    if (o_uavg) o_uavg[lev] = uavg;
                    if (o_dfdt) o_dfdt[lev] = (1.0 - ssl) * 4. * pi * (uavg - plsorc);
                }
                // the new downward intensities to everyone
#pragma unroll
                for (int c = 0; c < n; c++) d[c] = __shfl_sync(FULLMASK, x, c);
            }
        }
        }
        if (lane == 0 && have && status != -100) a.status[bin] = status;
        __syncwarp();
#ifdef SBD_PHASE_TIMING
        __syncthreads();
        ADD_TICK(3);
#endif
    }
}

#ifdef SBD_PHASE_TIMING
extern "C" void sbd_debug_add_ticks(unsigned long long *out, int reset)
{
    cudaMemcpyFromSymbol(out, g_add_ticks, sizeof(g_add_ticks));
    if (reset) { unsigned long long z[8] = {0, 0, 0, 0, 0, 0, 0, 0}; cudaMemcpyToSymbol(g_add_ticks, z, sizeof z); }
}
#endif

// ---- host-side launch helpers ---------------------------------------------
bool adding_supported(int N) { return N == 4 || N == 8 || N == 16 || N == 20 || N == 24 || N == 32; }

// resident warps per SM the register budget allows (launch bounds): 16 up to NSTR = 16, 8 above
int adding_warps_per_sm(int N) { return N > 16 ? 8 : 16; }

size_t adding_slot_doubles(int N, int L)
{
    switch (N) {
    case 4: return AddLayout<2>::slot_doubles(L);
    case 8: return AddLayout<4>::slot_doubles(L);
    case 16: return AddLayout<8>::slot_doubles(L);
    case 20: return AddLayout<10>::slot_doubles(L);
    case 24: return AddLayout<12>::slot_doubles(L);
    case 32: return AddLayout<16>::slot_doubles(L);
    }
    return 0;
}

size_t adding_smem_bytes(int N, int L, int warps)
{
    switch (N) {
    case 4: return 8 * (AddLayout<2>::cta + (size_t)warps * AddLayout<2>::warp_doubles(L));
    case 8: return 8 * (AddLayout<4>::cta + (size_t)warps * AddLayout<4>::warp_doubles(L));
    case 16: return 8 * (AddLayout<8>::cta + (size_t)warps * AddLayout<8>::warp_doubles(L));
    case 20: return 8 * (AddLayout<10>::cta + (size_t)warps * AddLayout<10>::warp_doubles(L));
    case 24: return 8 * (AddLayout<12>::cta + (size_t)warps * AddLayout<12>::warp_doubles(L));
    case 32: return 8 * (AddLayout<16>::cta + (size_t)warps * AddLayout<16>::warp_doubles(L));
    }
    return 0;
}

template <int n, int WARPS>
static cudaError_t launch_adding_k(const LaunchArgs &a, int grid, size_t smem, cudaStream_t st)
{
    cudaError_t e = cudaFuncSetAttribute(disort_adding_kernel<n, WARPS>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    disort_adding_kernel<n, WARPS><<<grid, WARPS * 32, smem, st>>>(a);
    return cudaGetLastError();
}

template <int n>
static cudaError_t launch_adding_t(const LaunchArgs &a, int warps, int grid, cudaStream_t st)
{
    const size_t smem = 8 * (AddLayout<n>::cta + (size_t)warps * AddLayout<n>::warp_doubles(a.d.nlyr));
    if (n > 8) return warps == 4 ? launch_adding_k<n, 4>(a, grid, smem, st) : cudaErrorInvalidValue;
    switch (warps) {
    case 4: return launch_adding_k<n, 4>(a, grid, smem, st);
    case 8: return launch_adding_k<(n > 8 ? 8 : n), 8>(a, grid, smem, st);
    }
    return cudaErrorInvalidValue;
}

cudaError_t launch_adding(const LaunchArgs &a, int warps, int grid, cudaStream_t st)
{
    switch (a.d.nstr) {
    case 4: return launch_adding_t<2>(a, warps, grid, st);
    case 8: return launch_adding_t<4>(a, warps, grid, st);
    case 16: return launch_adding_t<8>(a, warps, grid, st);
    case 20: return launch_adding_t<10>(a, warps, grid, st);
    case 24: return launch_adding_t<12>(a, warps, grid, st);
    case 32: return launch_adding_t<16>(a, warps, grid, st);
    }
    return cudaErrorInvalidValue;
}

}  // namespace sbd
