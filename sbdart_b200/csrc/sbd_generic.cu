// Generic warp-per-bin discrete-ordinate kernel (any even NSTR <= 40).
//
// One warp owns one spectral bin and runs the whole solve with its working
// set in shared memory; the per-layer factors needed by the upward sweep are
// spilled to an L2-resident scratch slot owned by the warp.
//
// The mathematics is the DISORT v2 solution (reference disort.f:472-871;
// SURVEY appendix C) but none of the reference's algorithms are kept where a
// GPU-friendlier one gives the same numbers:
//   * SOLEIG/ASYMTX (disort.f:3099, :873): the reduced n x n eigenproblem
//     (alpha+beta)(alpha-beta) e = k^2 e is symmetrised with
//     D = diag(sqrt(w_i mu_i)); a Cholesky factor of the odd-parity operator
//     turns it into a symmetric positive semidefinite problem solved by
//     parallel-order cyclic Jacobi (Stamnes, Tsay, Nakajima 1988).
//   * UPBEAM (disort.f:4130): the N x N LU is replaced by the spectral
//     solution of the same linear system in the eigenbasis already computed.
//   * UPISOT (disort.f:4247): (I-C) 1 = (1-w') 1 holds exactly for the
//     quadrature, so Z1 = XR1 and Z0 needs two triangular solves with the
//     Cholesky factor.
//   * SETMTX/SOLVE0 + SGBCO/SGBSL (disort.f:2702, :3322): the block
//     bidiagonal boundary system is eliminated layer by layer with partial
//     pivoting inside a (n+N) x (2N+1) window; only N pivot rows per layer
//     are kept for the back substitution.
//   * FLUXES (disort.f:1780): evaluated layer by layer during the back
//     substitution, exponentials hoisted out of the double sum.
#include <math.h>
#include <stdio.h>

#include "sbd_internal.h"
#include "sbd_planck.cuh"

namespace sbd {

#define FULLMASK 0xffffffffu

// e / d for 0 <= e < 2^16 with the precomputed magic number ceil(2^32 / d): one IMAD.HI
// instead of the ~20-instruction integer division by a run-time divisor
__device__ __forceinline__ unsigned div_magic(int d) { return 0xFFFFFFFFu / (unsigned)d + 1u; }
__device__ __forceinline__ int fdiv(int e, unsigned magic) { return (int)__umulhi((unsigned)e, magic); }

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULLMASK, v, o);
    return v;
}

// ---- shared-memory carve-up ----------------------------------------------
struct CtaShared {          // same for every warp of the CTA
    double *mu, *wt, *sq, *dinv;         // [n] each
    const double *ylm;                   // [N*n] Y_l^m(mu_i) of the current azimuth mode
};

struct WarpShared {
    double *gl, *y0;                 // [N]
    double *Pe, *Lo, *T, *V, *X, *P; // [n*n] each
    double *kk[2], *ek[2], *Gp[2], *Gm[2], *zz[2], *zp0[2];
    double *W;                       // [(n+N)*(2N+1)]
    double *xn, *xc;                 // [N]
    double *v1, *v2, *v3, *v4;       // [N] scratch vectors
    double *taucpr, *tauc, *pk;      // [L+1]
    int *layru;                      // [NT]
    double *cs;                      // Jacobi rotations [n+2]
    int *pq;                         // [n+2]
};

__host__ __device__ inline size_t cta_shared_doubles(int N)
{
    int n = N / 2;
    return 4 * n + (size_t)N * n;
}

__host__ __device__ inline size_t warp_shared_doubles(int N, int L, int NT)
{
    int n = N / 2;
    // the eigenproblem scratch (Pe, Lo, T, V, X, P) lives inside W, see carve()
    size_t o = 2 * N + 2 * (2 * n + 2 * (size_t)n * n + 2 * N) +
               (size_t)(n + N) * (2 * N + 1) + 2 * N + 4 * N + 3 * (size_t)(L + 1) +
               (n + 2);
    size_t ints = (size_t)NT + (n + 2);
    return o + (ints + 1) / 2 + 2;
}

size_t generic_smem_bytes(int N, int L, int NT, int warps)
{
    return 8 * (cta_shared_doubles(N) + warps * warp_shared_doubles(N, L, NT > 8 ? NT : 8));
}

size_t generic_slot_doubles(int N, int L, int NU)
{
    LayerLayout ll(N, NU);
    return (size_t)ll.stride * L + (size_t)N * L;
}

int generic_pick_warps(int N, int L, int NT, size_t smem_limit)
{
    int w = 8;
    while (w > 1 && generic_smem_bytes(N, L, NT, w) > smem_limit) w--;
    if (generic_smem_bytes(N, L, NT, w) > smem_limit) return 0;
    return w;
}

__device__ void carve(double *base, int N, int L, int NT, WarpShared &w)
{
    int n = N / 2;
    double *p = base;
    w.gl = p; p += N;
    w.y0 = p; p += N;
    for (int s = 0; s < 2; s++) {
        w.kk[s] = p; p += n;
        w.ek[s] = p; p += n;
        w.Gp[s] = p; p += n * n;
        w.Gm[s] = p; p += n * n;
        w.zz[s] = p; p += N;
        w.zp0[s] = p; p += N;
    }
    w.W = p; p += (n + N) * (2 * N + 1);
    {
        // Scratch of solve_layer, aliased onto rows n .. n+N-1 of the elimination window: while
        // a layer's eigenproblem is solved only the n carried rows (rows 0 .. n-1) are live, the
        // interface rows are assembled afterwards from Gp/Gm/kk/ek/zz/zp0; the upward sweep
        // never calls solve_layer.  4 n^2 + 2 n (n|1) <= N (2N+1) doubles.
        double *q = w.W + n * (2 * N + 1);
        w.Pe = q; q += n * n;
        w.Lo = q; q += n * n;
        w.T = q; q += n * (n | 1);     // T and V: odd leading dimension (column walks of the
        w.V = q; q += n * (n | 1);     // Jacobi rotations would otherwise hit one bank pair)
        w.X = q; q += n * n;
        w.P = q; q += n * n;
    }
    w.xn = p; p += N;
    w.xc = p; p += N;
    w.v1 = p; p += N;
    w.v2 = p; p += N;
    w.v3 = p; p += N;
    w.v4 = p; p += N;
    w.taucpr = p; p += L + 1;
    w.tauc = p; p += L + 1;
    w.pk = p; p += L + 1;
    w.cs = p; p += n + 2;
    w.layru = (int *)p;
    w.pq = w.layru + NT;
}


// The structs above hold generic pointers; the device functions below are too large to be
// inlined, so the compiler cannot see that every one of them points into shared memory and
// emits generic LD/ST (slower, and tracked on the long scoreboard).  These hints restore
// LDS/STS.  (Wrapping the dynamically indexed buffer pointers w.Gp[s] ... in a hinting helper
// at every use miscompiled the radiance path with nvcc 12.9, so those stay generic.)
#define SBD_SHARED(p) __builtin_assume(__isShared(p))

__device__ __forceinline__ void hint_shared(const CtaShared &cs)
{
    // cs.ylm is NOT hinted: it points to global memory for azimuth modes m > 0
    SBD_SHARED(cs.mu); SBD_SHARED(cs.wt); SBD_SHARED(cs.sq); SBD_SHARED(cs.dinv);
}
__device__ __forceinline__ void hint_shared(const WarpShared &w)
{
    SBD_SHARED(w.gl); SBD_SHARED(w.y0);
    SBD_SHARED(w.Pe); SBD_SHARED(w.Lo); SBD_SHARED(w.T); SBD_SHARED(w.V); SBD_SHARED(w.X); SBD_SHARED(w.P);
#pragma unroll
    for (int s = 0; s < 2; s++) {
        SBD_SHARED(w.kk[s]); SBD_SHARED(w.ek[s]); SBD_SHARED(w.Gp[s]); SBD_SHARED(w.Gm[s]);
        SBD_SHARED(w.zz[s]); SBD_SHARED(w.zp0[s]);
    }
    SBD_SHARED(w.W); SBD_SHARED(w.xn); SBD_SHARED(w.xc);
    SBD_SHARED(w.v1); SBD_SHARED(w.v2); SBD_SHARED(w.v3); SBD_SHARED(w.v4);
    SBD_SHARED(w.taucpr); SBD_SHARED(w.tauc); SBD_SHARED(w.pk);
    SBD_SHARED(w.layru); SBD_SHARED(w.cs); SBD_SHARED(w.pq);
}

// per-bin scalars held in registers by every lane
struct BinCtx {
    int N, n, L, NT, ncut, lyrcut, plank, mazim;
    double fbeam, umu0, albedo, fisot, tplank, bplank, delm0;
    const double *dtauc, *ssalb, *pmom;
    int ldp;
    // BRDF surface (LAMBER = .FALSE.): SURFAC's tables of the current azimuth mode, or null
    const double *bdr, *bem, *rmu, *emu;
};

// ---------------------------------------------------------------------------
// Per-layer solution for azimuth mode m: eigenvalues k_j, eigenvector blocks
// G+ / G-, beam and thermal particular solutions (reference SOLEIG, UPBEAM,
// UPISOT).  Results go to buffer set `s` in shared memory.
// returns 0 or SBD_BIN_EIG_FAIL (uniform across the warp)
// ---------------------------------------------------------------------------
__device__ int solve_layer(const BinCtx &c, const CtaShared &cs, WarpShared &w,
                           int lc, int s, double &xr0, double &xr1, int lane)
{
    const int N = c.N, n = c.n, m = c.mazim;
    const int ldt = n | 1;                             // leading dimension of T and V
    const unsigned rn = div_magic(n);
    hint_shared(cs); hint_shared(w);
    double ss = c.ssalb[lc];
    if (ss == 1.0) ss = 1.0 - kDither;                 // disort.f:486
    double dt = c.dtauc[lc];
    if (dt < 0.0) dt = 0.0;                            // disort.f:4944
    const double f = c.pmom[(size_t)lc * c.ldp + N];   // delta-M, disort.f:2577
    const double oprim = ss * (1. - f) / (1. - f * ss);
    const double dtaucp = (1. - f * ss) * dt;

    for (int l = lane; l < N; l += 32) {
        double pm = (l == 0) ? 1.0 : c.pmom[(size_t)lc * c.ldp + l];
        w.gl[l] = (2 * l + 1) * oprim * (pm - f) / (1. - f);   // disort.f:2583
    }
    __syncwarp();

    // symmetrised even / odd parity operators
    //   Pe~ = M^-1/2 (I - 2 W^1/2 Se W^1/2) M^-1/2 , Po~ likewise with So
    for (int e = lane; e < n * n; e += 32) {
        int i = fdiv(e, rn), j = e - i * n;
        double se = 0.0, so = 0.0;
        for (int l = m; l < N; l++) {
            double t = w.gl[l] * cs.ylm[l * n + i] * cs.ylm[l * n + j];
            if ((l - m) & 1) so += t; else se += t;
        }
        double dg = (i == j) ? 1.0 / cs.mu[i] : 0.0;
        double sc = cs.sq[i] * cs.sq[j];
        w.Pe[e] = dg - sc * se;
        w.Lo[e] = dg - sc * so;
    }
    __syncwarp();

    // Cholesky Po~ = L L^T (lower triangle of Lo, in place)
    int bad = 0;
    for (int j = 0; j < n; j++) {
        double d = w.Lo[j * n + j];
        if (!(d > 0.0)) { bad = 1; break; }
        d = sqrt(d);
        __syncwarp();
        for (int i = j + lane; i < n; i += 32) w.Lo[i * n + j] = (i == j) ? d : w.Lo[i * n + j] / d;
        __syncwarp();
        int rem = n - j - 1;
        for (int e = lane; e < rem * rem; e += 32) {
            int a = e / rem, b = e - a * rem;
            if (b <= a) {
                int i = j + 1 + a, k = j + 1 + b;
                w.Lo[i * n + k] -= w.Lo[i * n + j] * w.Lo[k * n + j];
            }
        }
        __syncwarp();
    }
    if (bad) return SBD_BIN_EIG_FAIL;

    // X = Pe~ L ;  T = L^T X  (symmetric positive semidefinite, eigenvalues k^2)
    for (int e = lane; e < n * n; e += 32) {
        int i = fdiv(e, rn), j = e - i * n;
        double a = 0.0;
        for (int k = j; k < n; k++) a += w.Pe[i * n + k] * w.Lo[k * n + j];
        w.X[e] = a;
    }
    __syncwarp();
    for (int e = lane; e < n * n; e += 32) {
        int i = fdiv(e, rn), j = e - i * n;
        double a = 0.0;
        for (int k = i; k < n; k++) a += w.Lo[k * n + i] * w.X[k * n + j];
        w.P[e] = a;
    }
    __syncwarp();
    for (int e = lane; e < n * n; e += 32) {
        int i = fdiv(e, rn), j = e - i * n;
        w.T[i * ldt + j] = 0.5 * (w.P[i * n + j] + w.P[j * n + i]);
        w.V[i * ldt + j] = (i == j) ? 1.0 : 0.0;
    }
    __syncwarp();

    // parallel-order cyclic Jacobi on T, eigenvectors accumulated in V
    {
        const int mm = (n & 1) ? n + 1 : n;    // players in the tournament
        const int np = mm / 2;
        double tr = 0.0;
        for (int i = lane; i < n; i += 32) tr += fabs(w.T[i * ldt + i]);
        const double floor_abs = 1.0e-19 * warp_sum(tr);
        int sweep = 0;
        for (; sweep < 60; sweep++) {
            int nrot = 0;
            for (int r = 0; r < mm - 1; r++) {
                int did = 0;
                if (lane < np) {
                    int p, q;
                    if (lane == 0) { p = r; q = mm - 1; }
                    else { p = (r + lane) % (mm - 1); q = (r - lane + (mm - 1)) % (mm - 1); }
                    if (p > q) { int t = p; p = q; q = t; }
                    double cc = 1.0, sn = 0.0;
                    if (q < n) {
                        double apq = w.T[p * ldt + q], app = w.T[p * ldt + p], aqq = w.T[q * ldt + q];
                        if (fabs(apq) > floor_abs &&
                            fabs(apq) > 1.1e-16 * sqrt(fabs(app) * fabs(aqq))) {
                            double th = (aqq - app) / (2.0 * apq);
                            double t = 1.0 / (fabs(th) + sqrt(th * th + 1.0));
                            if (th < 0.0) t = -t;
                            cc = 1.0 / sqrt(t * t + 1.0);
                            sn = t * cc;
                            did = 1;
                        }
                    } else { q = p; }
                    w.cs[2 * lane] = cc; w.cs[2 * lane + 1] = sn;
                    w.pq[2 * lane] = p; w.pq[2 * lane + 1] = q;
                }
                nrot += __popc(__ballot_sync(FULLMASK, did));
                __syncwarp();
                // column rotations of T and V
                for (int e = lane; e < np * n; e += 32) {
                    int k = fdiv(e, rn), i = e - k * n;
                    int p = w.pq[2 * k], q = w.pq[2 * k + 1];
                    if (p != q) {
                        double cc = w.cs[2 * k], sn = w.cs[2 * k + 1];
                        double a = w.T[i * ldt + p], b = w.T[i * ldt + q];
                        w.T[i * ldt + p] = cc * a - sn * b;
                        w.T[i * ldt + q] = sn * a + cc * b;
                        a = w.V[i * ldt + p]; b = w.V[i * ldt + q];
                        w.V[i * ldt + p] = cc * a - sn * b;
                        w.V[i * ldt + q] = sn * a + cc * b;
                    }
                }
                __syncwarp();
                // row rotations of T
                for (int e = lane; e < np * n; e += 32) {
                    int k = fdiv(e, rn), j = e - k * n;
                    int p = w.pq[2 * k], q = w.pq[2 * k + 1];
                    if (p != q) {
                        double cc = w.cs[2 * k], sn = w.cs[2 * k + 1];
                        double a = w.T[p * ldt + j], b = w.T[q * ldt + j];
                        w.T[p * ldt + j] = cc * a - sn * b;
                        w.T[q * ldt + j] = sn * a + cc * b;
                    }
                }
                __syncwarp();
            }
            if (nrot == 0) break;
        }
        if (sweep >= 60) return SBD_BIN_EIG_FAIL;
    }

    // k_j = sqrt(|lambda_j|)  (disort.f:3264-3269)
    for (int j = lane; j < n; j += 32) {
        double k = sqrt(fabs(w.T[j * ldt + j]));
        w.kk[s][j] = k;
        w.ek[s][j] = exp(-k * dtaucp);
    }
    // Q = L^-T V (into X), P = L V
    for (int j = lane; j < n; j += 32) {
        for (int i = n - 1; i >= 0; i--) {
            double a = w.V[i * ldt + j];
            for (int k = i + 1; k < n; k++) a -= w.Lo[k * n + i] * w.X[k * n + j];
            w.X[i * n + j] = a / w.Lo[i * n + i];
        }
    }
    for (int e = lane; e < n * n; e += 32) {
        int i = fdiv(e, rn), j = e - i * n;
        double a = 0.0;
        for (int k = 0; k <= i; k++) a += w.Lo[i * n + k] * w.V[k * ldt + j];
        w.P[e] = a;
    }
    __syncwarp();
    // G+ - G- = e = D^-1 Q ; G+ + G- = (alpha-beta) e / k = -D^-1 P / k
    for (int e = lane; e < n * n; e += 32) {
        int i = fdiv(e, rn), j = e - i * n;
        double gd = cs.dinv[i] * w.X[e];
        double gs = -cs.dinv[i] * w.P[e] / w.kk[s][j];
        w.Gp[s][e] = 0.5 * (gs + gd);
        w.Gm[s][e] = 0.5 * (gs - gd);
    }

    // ---- beam particular solution (reference UPBEAM) ----------------------
    if (c.fbeam > 0.0) {
        const double fac = (2. - c.delm0) * c.fbeam / (4. * kPiRef);
        const double rmu0 = 1.0 / c.umu0;
        // b^_s, b^_d: even / odd parity parts of the source, scaled by sq
        for (int i = lane; i < n; i += 32) {
            double be = 0.0, bo = 0.0;
            for (int l = m; l < N; l++) {
                double t = w.gl[l] * cs.ylm[l * n + i] * w.y0[l];
                if ((l - m) & 1) bo += t; else be += t;
            }
            w.v1[i] = 2.0 * fac * cs.sq[i] * be;   // b^_s
            w.v2[i] = 2.0 * fac * cs.sq[i] * bo;   // b^_d
        }
        __syncwarp();
        // r = b^_s / mu0 - Pe~ b^_d
        for (int i = lane; i < n; i += 32) {
            double a = w.v1[i] * rmu0;
            for (int k = 0; k < n; k++) a -= w.Pe[i * n + k] * w.v2[k];
            w.v3[i] = a;
        }
        __syncwarp();
        // c = V^T L^T r, scaled by 1/(1/mu0^2 - lambda)
        for (int i = lane; i < n; i += 32) {
            double a = 0.0;
            for (int k = i; k < n; k++) a += w.Lo[k * n + i] * w.v3[k];
            w.v4[i] = a;
        }
        __syncwarp();
        for (int j = lane; j < n; j += 32) {
            double a = 0.0;
            for (int k = 0; k < n; k++) a += w.V[k * ldt + j] * w.v4[k];
            double den = rmu0 * rmu0 - w.T[j * ldt + j];
            w.v3[j] = a / den;
        }
        __syncwarp();
        // y = V c
        for (int i = lane; i < n; i += 32) {
            double a = 0.0;
            for (int k = 0; k < n; k++) a += w.V[i * ldt + k] * w.v3[k];
            w.v4[i] = a;
        }
        __syncwarp();
        // d^ = L^-T y (one lane, n is small), s^ = mu0 (b^_d - L y)
        if (lane == 0) {
            for (int i = n - 1; i >= 0; i--) {
                double a = w.v4[i];
                for (int k = i + 1; k < n; k++) a -= w.Lo[k * n + i] * w.v3[k];
                w.v3[i] = a / w.Lo[i * n + i];
            }
        }
        __syncwarp();
        for (int i = lane; i < n; i += 32) {
            double a = 0.0;
            for (int k = 0; k <= i; k++) a += w.Lo[i * n + k] * w.v4[k];
            double sh = c.umu0 * (w.v2[i] - a);
            double dh = w.v3[i];
            double sv = cs.dinv[i] * sh, dv = cs.dinv[i] * dh;
            w.zz[s][n + i] = 0.5 * (sv + dv);       // +mu_i
            w.zz[s][n - 1 - i] = 0.5 * (sv - dv);   // -mu_i
        }
    } else {
        for (int i = lane; i < N; i += 32) w.zz[s][i] = 0.0;
    }

    // ---- thermal particular solution (reference UPISOT) -------------------
    xr0 = 0.0; xr1 = 0.0;
    if (c.plank && m == 0) {
        if (dtaucp > 0.0) xr1 = (w.pk[lc + 1] - w.pk[lc]) / dtaucp;   // disort.f:661-666
        xr0 = w.pk[lc] - xr1 * w.taucpr[lc];
        // q = D^-1 L^-T L^-1 D 1 ;  Z0(+-mu_i) = xr0 +- xr1 q_i ; Z1 = xr1
        __syncwarp();
        if (lane == 0) {
            for (int i = 0; i < n; i++) {
                double a = 1.0 / cs.dinv[i];
                for (int k = 0; k < i; k++) a -= w.Lo[i * n + k] * w.v1[k];
                w.v1[i] = a / w.Lo[i * n + i];
            }
            for (int i = n - 1; i >= 0; i--) {
                double a = w.v1[i];
                for (int k = i + 1; k < n; k++) a -= w.Lo[k * n + i] * w.v2[k];
                w.v2[i] = a / w.Lo[i * n + i];
            }
        }
        __syncwarp();
        for (int i = lane; i < n; i += 32) {
            double q = cs.dinv[i] * w.v2[i];
            w.zp0[s][n + i] = xr0 + xr1 * q;
            w.zp0[s][n - 1 - i] = xr0 - xr1 * q;
        }
    } else {
        for (int i = lane; i < N; i += 32) w.zp0[s][i] = 0.0;
    }
    __syncwarp();
    return 0;
}

// GC(r, j) of the reference (disort.f:3309-3312) from the compact blocks
__device__ __forceinline__ double gc_elem(const double *Gp, const double *Gm,
                                          int n, int r, int j)
{
    SBD_SHARED(Gp); SBD_SHARED(Gm);
    if (r >= n) {
        int i = r - n;
        return (j >= n) ? Gp[i * n + (j - n)] : -Gm[i * n + (n - 1 - j)];
    } else {
        int i = n - 1 - r;
        return (j >= n) ? Gm[i * n + (j - n)] : -Gp[i * n + (n - 1 - j)];
    }
}

// scratch (global) -> shared copy, four loads in flight per lane: written as a plain loop
// the loads are issued one at a time behind their stores (the destination pointers that come
// out of the run-time indexed buffer sets are generic, so the compiler assumes aliasing)
__device__ __forceinline__ void copy_in(double *dst, const double *src, int count, int lane)
{
    int e = lane;
    for (; e + 96 < count; e += 128) {
        // plain (coherent) loads: the scratch slot was written by this warp earlier in the same launch
        const double v0 = src[e], v1 = src[e + 32], v2 = src[e + 64], v3 = src[e + 96];
        dst[e] = v0; dst[e + 32] = v1; dst[e + 64] = v2; dst[e + 96] = v3;
    }
    for (; e < count; e += 32) dst[e] = src[e];
}

// copy a layer record shared -> scratch
__device__ void store_layer(const LayerLayout &ll, double *rec, const WarpShared &w,
                            int s, double xr0, double xr1, int lane)
{
    const int n = ll.n, N = ll.N;
    hint_shared(w);
    for (int e = lane; e < n; e += 32) { rec[ll.off_kk + e] = w.kk[s][e]; rec[ll.off_ek + e] = w.ek[s][e]; }
    for (int e = lane; e < n * n; e += 32) { rec[ll.off_gp + e] = w.Gp[s][e]; rec[ll.off_gm + e] = w.Gm[s][e]; }
    for (int e = lane; e < N; e += 32) { rec[ll.off_zz + e] = w.zz[s][e]; rec[ll.off_zp0 + e] = w.zp0[s][e]; }
    if (lane == 0) { rec[ll.off_xr] = xr0; rec[ll.off_xr + 1] = xr1; }
}

__device__ void load_layer(const LayerLayout &ll, const double *rec, WarpShared &w,
                           int s, double &xr0, double &xr1, int lane)
{
    const int n = ll.n, N = ll.N;
    hint_shared(w);
    copy_in(w.kk[s], rec + ll.off_kk, n, lane);
    copy_in(w.ek[s], rec + ll.off_ek, n, lane);
    copy_in(w.Gp[s], rec + ll.off_gp, n * n, lane);
    copy_in(w.Gm[s], rec + ll.off_gm, n * n, lane);
    copy_in(w.zz[s], rec + ll.off_zz, N, lane);
    copy_in(w.zp0[s], rec + ll.off_zp0, N, lane);
    xr0 = rec[ll.off_xr]; xr1 = rec[ll.off_xr + 1];
}

// Gaussian elimination with partial pivoting of the first `ncols` columns of
// the row-major window W[rows][C]; pivot rows end up in rows 0..ncols-1.
// returns 0 or SBD_BIN_SINGULAR
__device__ int eliminate(double *W, int rows, int C, int ncols, int lane)
{
    SBD_SHARED(W);
    for (int j = 0; j < ncols; j++) {
        // pivot search in column j among rows j..rows-1
        double best = -1.0; int bi = j;
        for (int r = j + lane; r < rows; r += 32) {
            double v = fabs(W[r * C + j]);
            if (v > best) { best = v; bi = r; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            double ob = __shfl_xor_sync(FULLMASK, best, o);
            int oi = __shfl_xor_sync(FULLMASK, bi, o);
            if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
        }
        if (!(best > 0.0)) return SBD_BIN_SINGULAR;
        if (bi != j) {
            for (int cidx = j + lane; cidx < C; cidx += 32) {
                double t = W[j * C + cidx];
                W[j * C + cidx] = W[bi * C + cidx];
                W[bi * C + cidx] = t;
            }
        }
        __syncwarp();
        const double rp = 1.0 / W[j * C + j];
        // lanes over columns, loop over rows; four rows per trip with all loads ahead of the
        // stores (the compiler cannot prove that a store to column cidx leaves column j alone,
        // so the plain loop is one load -> FMA -> store latency chain per row)
        for (int cidx = j + 1 + lane; cidx < C; cidx += 32) {
            const double pv = W[j * C + cidx];
            int r = j + 1;
            for (; r + 3 < rows; r += 4) {
                double *p0 = W + r * C;
                const double m0 = p0[j], m1 = p0[C + j], m2 = p0[2 * C + j], m3 = p0[3 * C + j];
                const double v0 = p0[cidx], v1 = p0[C + cidx], v2 = p0[2 * C + cidx], v3 = p0[3 * C + cidx];
                p0[cidx] = v0 - (m0 * rp) * pv;
                p0[C + cidx] = v1 - (m1 * rp) * pv;
                p0[2 * C + cidx] = v2 - (m2 * rp) * pv;
                p0[3 * C + cidx] = v3 - (m3 * rp) * pv;
            }
            for (; r < rows; r++) {
                double mlt = W[r * C + j] * rp;
                W[r * C + cidx] -= mlt * pv;
            }
        }
        __syncwarp();
    }
    return 0;
}

// Y_l^m(x) for l = 0..N-1 (zero below l = m): what LEPOLY (disort.f:5286) builds
// mode by mode, with the diagonal term in closed-chain form.
__device__ void lepoly_one(int m, int N, double x, double *y)
{
    SBD_SHARED(y);
    if (m == 0) {
        y[0] = 1.0;
        if (N > 1) y[1] = x;
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

// ---------------------------------------------------------------------------
// User-angle quantities of one layer for azimuth mode m (reference TERPEV
// disort.f:3920 and TERPSO disort.f:3980): eigenvectors and particular
// solutions re-expanded at the user cosines through the Legendre sum.
// Expects the layer record in buffer set s of shared memory and x_lc in w.xc.
// Writes GU(iu,col)*LL(col), ZBEAM, Z0U, Z1U into the scratch record.
// ---------------------------------------------------------------------------
__device__ void user_terms(const BinCtx &c, const CtaShared &cs, WarpShared &w,
                           const LayerLayout &ll, double *rec, const double *ylmu_m, int NU,
                           int lc, int s, double xr0, double xr1, int lane)
{
    const int N = c.N, n = c.n, m = c.mazim;
    hint_shared(cs); hint_shared(w);
    double ss = c.ssalb[lc];
    if (ss == 1.0) ss = 1.0 - kDither;
    const double f = c.pmom[(size_t)lc * c.ldp + N];
    const double oprim = ss * (1. - f) / (1. - f * ss);
    for (int l = lane; l < N; l += 32) {
        double pm = (l == 0) ? 1.0 : c.pmom[(size_t)lc * c.ldp + l];
        w.gl[l] = (2 * l + 1) * oprim * (pm - f) / (1. - f);
    }
    __syncwarp();
    // E[l][col] = 0.5 g_l sum_jq w_jq Y_l(mu_jq) EVECC(jq, col)
    double *E = w.W;
    for (int e = lane; e < N * N; e += 32) {
        const int l = e / N, col = e - l * N;
        double acc = 0.0;
        if (l >= m) {
            const double sg = ((l - m) & 1) ? -1.0 : 1.0;
            const bool plus = col >= n;
            const int j = plus ? col - n : n - 1 - col;
            for (int i = 0; i < n; i++) {
                const double gp = w.Gp[s][i * n + j], gm = w.Gm[s][i * n + j];
                const double ev = plus ? (gp + sg * gm) : -(gm + sg * gp);
                acc += cs.wt[i] * cs.ylm[l * n + i] * ev;
            }
            acc *= 0.5 * w.gl[l];
        }
        E[e] = acc;
    }
    __syncwarp();
    for (int e = lane; e < NU * N; e += 32) {
        const int iu = e / N, col = e - iu * N;
        double acc = 0.0;
        for (int l = m; l < N; l++) acc += E[l * N + col] * ylmu_m[l * NU + iu];
        rec[ll.off_gu + e] = acc * w.xc[col];          // GU * LL (disort.f:4496-4506)
    }
    // source terms: psi[l] for beam (v1), Planck Z0 (v2) and Z1 (v3)
    for (int l = lane; l < N; l += 32) {
        double pb = 0.0, p0 = 0.0, p1 = 0.0;
        if (l >= m) {
            const double sg = ((l - m) & 1) ? -1.0 : 1.0;
            for (int i = 0; i < n; i++) {
                const double wy = cs.wt[i] * cs.ylm[l * n + i];
                pb += wy * (w.zz[s][n + i] + sg * w.zz[s][n - 1 - i]);
                p0 += wy * (w.zp0[s][n + i] + sg * w.zp0[s][n - 1 - i]);
                p1 += wy * (xr1 + sg * xr1);
            }
        }
        w.v1[l] = 0.5 * w.gl[l] * pb;
        w.v2[l] = 0.5 * w.gl[l] * p0;
        w.v3[l] = 0.5 * w.gl[l] * p1;
    }
    __syncwarp();
    const double fact = (2. - c.delm0) * c.fbeam / (4.0 * kPiRef);
    for (int iu = lane; iu < NU; iu += 32) {
        double zb = 0.0, z0 = 0.0, z1 = 0.0;
        for (int l = m; l < N; l++) {
            const double yu = ylmu_m[l * NU + iu];
            if (c.fbeam > 0.0) zb += yu * (w.v1[l] + fact * w.gl[l] * w.y0[l]);
            z0 += yu * w.v2[l];
            z1 += yu * w.v3[l];
        }
        rec[ll.off_zb + iu] = zb;
        if (c.plank && m == 0) {
            rec[ll.off_z0u + iu] = z0 + (1. - oprim) * xr0;
            rec[ll.off_z1u + iu] = z1 + (1. - oprim) * xr1;
        } else {
            rec[ll.off_z0u + iu] = 0.0;
            rec[ll.off_z1u + iu] = 0.0;
        }
    }
    __syncwarp();
}

// ---------------------------------------------------------------------------
// Intensity of azimuth mode m at user angle umu and level (lu) by analytic
// integration of the source function layer by layer (reference USRINT,
// disort.f:4355-4793; Lambertian surface).  One call per (level, angle) lane.
// ---------------------------------------------------------------------------
__device__ double usrint_one(const BinCtx &c, const WarpShared &w, const LayerLayout &ll,
                             const double *scr, int NU, int lu, int iu, double umu, double utpr,
                             double fisot, double bnd_up /* surface term without the exp */)
{
    const int N = c.N, n = c.n, m = c.mazim, ncut = c.ncut, L = c.L;
    hint_shared(w);
    const int lyu = w.layru[lu];
    if (c.lyrcut && lyu > ncut) return 0.0;
    const bool negumu = umu < 0.0;
    const bool therm = c.plank && m == 0;
    const double umu0 = c.umu0;
    const double exp0 = (c.fbeam > 0.0) ? exp(-utpr / umu0) : 0.0;
    int lyrstr, lyrend; double sgn;
    if (negumu) { lyrstr = 1; lyrend = lyu - 1; sgn = -1.0; }
    else { lyrstr = lyu + 1; lyrend = ncut; sgn = 1.0; }
    double palint = 0.0, plkint = 0.0;
    double exp1 = 0.0, exp2 = 0.0;
    for (int lc = lyrstr; lc <= lyrend; lc++) {
        const double *rec = scr + (size_t)(lc - 1) * ll.stride;
        const double t0 = w.taucpr[lc - 1], t1 = w.taucpr[lc];
        const double dtau = t1 - t0;
        exp1 = exp((utpr - t0) / umu);
        exp2 = exp((utpr - t1) / umu);
        if (therm) {
            const double f0n = sgn * (exp1 - exp2);
            const double f1n = sgn * ((t0 + umu) * exp1 - (t1 + umu) * exp2);
            plkint += rec[ll.off_z0u + iu] * f0n + rec[ll.off_z1u + iu] * f1n;
        }
        if (c.fbeam > 0.0) {
            const double denom = 1. + umu / umu0;
            double expn;
            if (fabs(denom) < 0.0001) expn = (dtau / umu0) * exp0;
            else expn = (exp1 * exp(-t0 / umu0) - exp2 * exp(-t1 / umu0)) * sgn / denom;
            palint += rec[ll.off_zb + iu] * expn;
        }
        for (int iq = 0; iq < n; iq++) {        // columns of the -k solutions
            const double k = rec[ll.off_kk + n - 1 - iq], wk = rec[ll.off_ek + n - 1 - iq];
            const double denom = 1.0 - umu * k;
            double expn;
            if (fabs(denom) < 0.0001) expn = dtau / umu * exp2;
            else expn = sgn * (exp1 * wk - exp2) / denom;
            palint += rec[ll.off_gu + iu * N + iq] * expn;
        }
        for (int iq = n; iq < N; iq++) {        // columns of the +k solutions
            const double k = rec[ll.off_kk + iq - n], wk = rec[ll.off_ek + iq - n];
            const double denom = 1.0 + umu * k;
            double expn;
            if (fabs(denom) < 0.0001) expn = -dtau / umu * exp1;
            else expn = sgn * (exp1 - exp2 * wk) / denom;
            palint += rec[ll.off_gu + iu * N + iq] * expn;
        }
    }
    // the layer that contains the level (disort.f:4623-4729)
    {
        const double *rec = scr + (size_t)(lyu - 1) * ll.stride;
        const double t0 = w.taucpr[lyu - 1], t1 = w.taucpr[lyu];
        const double dtau1 = utpr - t0, dtau2 = utpr - t1, dtau = t1 - t0;
        const bool skip = (fabs(dtau1) < 1.e-6 && negumu) || (fabs(dtau2) < 1.e-6 && !negumu);
        if (!skip) {
            if (negumu) exp1 = exp(dtau1 / umu); else exp2 = exp(dtau2 / umu);
            if (c.fbeam > 0.0) {
                const double denom = 1. + umu / umu0;
                double expn;
                if (fabs(denom) < 0.0001) expn = (dtau1 / umu0) * exp0;
                else if (negumu) expn = (exp0 - exp(-t0 / umu0) * exp1) / denom;
                else expn = (exp0 - exp(-t1 / umu0) * exp2) / denom;
                palint += rec[ll.off_zb + iu] * expn;
            }
            for (int iq = 0; iq < n; iq++) {
                const double kq = -rec[ll.off_kk + n - 1 - iq];
                const double denom = 1. + umu * kq;
                double expn;
                if (fabs(denom) < 0.0001) expn = -dtau2 / umu * exp2;
                else if (negumu) expn = (exp(-kq * dtau2) - exp(kq * dtau) * exp1) / denom;
                else expn = (exp(-kq * dtau2) - exp2) / denom;
                palint += rec[ll.off_gu + iu * N + iq] * expn;
            }
            for (int iq = n; iq < N; iq++) {
                const double kq = rec[ll.off_kk + iq - n];
                const double denom = 1. + umu * kq;
                double expn;
                if (fabs(denom) < 0.0001) expn = -dtau1 / umu * exp1;
                else if (negumu) expn = (exp(-kq * dtau1) - exp1) / denom;
                else expn = (exp(-kq * dtau1) - exp(-kq * dtau) * exp2) / denom;
                palint += rec[ll.off_gu + iu * N + iq] * expn;
            }
            if (therm) {
                double expn, fact;
                if (negumu) { expn = exp1; fact = t0 + umu; }
                else { expn = exp2; fact = t1 + umu; }
                const double f0n = 1. - expn;
                const double f1n = utpr + umu - fact * expn;
                plkint += rec[ll.off_z0u + iu] * f0n + rec[ll.off_z1u + iu] * f1n;
            }
        }
    }
    // boundary contributions (disort.f:4735-4781)
    double bndint = 0.0;
    if (negumu && m == 0) bndint = (fisot + c.tplank) * exp(utpr / umu);
    else if (!negumu && c.rmu && !c.lyrcut) {
        // BRDF surface (disort.f:4744-4778): RMU-weighted downward intensities + direct beam + emission
        const double *rm = c.rmu + (size_t)iu * (n + 1);
        double bnd = 0.0;
        for (int k = 0; k < n; k++) bnd += rm[1 + k] * w.cs[k];
        bnd *= 1.0 + c.delm0;
        if (c.fbeam > 0.0) bnd += c.umu0 * c.fbeam / kPiRef * rm[0] * exp(-w.taucpr[L] / c.umu0);
        if (m == 0) bnd += c.emu[iu] * c.bplank;
        bndint = bnd * exp((utpr - w.taucpr[L]) / umu);
    } else if (!negumu && !(c.lyrcut || m > 0))
        bndint = bnd_up * exp((utpr - w.taucpr[L]) / umu);
    return palint + plkint + bndint;
}

// ---------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------
extern __shared__ double smem_dyn[];

template <bool SYNC>
__global__ void __launch_bounds__(256, 2)
disort_generic_kernel(const LaunchArgs a)
{
    const int N = a.d.nstr, n = N / 2, L = a.d.nlyr;
    const int NT = a.d.ntau > 0 ? a.d.ntau : L + 1;
    const int NU = a.d.numu;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int warps = blockDim.x >> 5;
    const int ldp = a.d.nmom + 1;
    const LayerLayout ll(N, NU);

    CtaShared cs;
    cs.mu = smem_dyn; cs.wt = cs.mu + n; cs.sq = cs.wt + n; cs.dinv = cs.sq + n;
    double *ylm_s = cs.dinv + n;
    cs.ylm = ylm_s;
    WarpShared w;
    const int NTs = NT > 8 ? NT : 8;
    carve(smem_dyn + cta_shared_doubles(N) + (size_t)warp * warp_shared_doubles(N, L, NTs),
          N, L, NTs, w);
    hint_shared(cs); hint_shared(w);

    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        double mu = a.quad[i], wt = a.quad[n + i];
        cs.mu[i] = mu; cs.wt[i] = wt;
        cs.sq[i] = sqrt(wt / mu);
        cs.dinv[i] = 1.0 / sqrt(wt * mu);
    }
    // flux path: azimuth mode 0 only
    for (int e = threadIdx.x; e < N * n; e += blockDim.x) ylm_s[e] = a.ylmc[e];
    __syncthreads();

    const int slot = blockIdx.x * warps + warp;
    double *scr = a.scratch + (size_t)slot * a.slot_stride;
    double *LLs = scr + (size_t)ll.stride * L;   // [L][N] solution coefficients

    for (;;) {
        int bin = 0;
        if (lane == 0) bin = atomicAdd(a.work_counter, 1);
        bin = __shfl_sync(FULLMASK, bin, 0);
        // The warps of a CTA move through bins, azimuth modes and layers together
        // (CTA barriers with per-warp "active" flags): they then execute the same code
        // at the same time and share its instruction-cache lines -- free-running warps
        // of this 150 KB kernel spend 40 % of their time waiting for instructions.
        const bool have = bin < (a.redo_consume ? *a.redo_count : (a.nbins_dev ? *a.nbins_dev : a.d.nbins));
        if (!(SYNC ? __syncthreads_or(have) : (int)have)) break;
        if (a.redo_consume && have) bin = a.redo_list[bin];      // bins handed over by the radiance register kernel

        const int src = !have ? 0 : (a.binmap ? a.binmap[bin] : bin);     // input slot of this bin
        const sbd_bin bp = a.bins[src];
        BinCtx c;
        c.N = N; c.n = n; c.L = L; c.NT = NT; c.mazim = 0; c.delm0 = 1.0;
        c.fbeam = bp.fbeam; c.umu0 = bp.umu0; c.albedo = bp.albedo; c.fisot = bp.fisot;
        c.plank = bp.plank;
        c.dtauc = a.dtauc + (size_t)src * L;
        c.ssalb = a.ssalb + (size_t)src * L;
        c.pmom = a.pmom + (size_t)src * L * ldp;
        c.ldp = ldp;
        double *o_rfldir = (a.rfldir && have) ? a.rfldir + (size_t)bin * NT : nullptr;
        double *o_rfldn = (a.rfldn && have) ? a.rfldn + (size_t)bin * NT : nullptr;
        double *o_flup = (a.flup && have) ? a.flup + (size_t)bin * NT : nullptr;
        double *o_dfdt = (a.dfdt && have) ? a.dfdt + (size_t)bin * NT : nullptr;
        double *o_uavg = (a.uavg && have) ? a.uavg + (size_t)bin * NT : nullptr;

        int status = have ? 0 : -1;
        int surf = -1;
        c.bdr = c.bem = c.rmu = c.emu = nullptr;
        // ---- input checks (subset of CHEKIN, disort.f:4920-5155) ----------
        {
            int badl = 0;
            for (int lc = lane; lc < L; lc += 32) {
                double s = c.ssalb[lc];
                if (!(s >= 0.0 && s <= 1.0)) badl = 1;
                if (!(fabs(c.dtauc[lc]) <= 1.79e308)) badl = 1;    // NaN / Inf optical depth
                for (int k = 1; k <= a.d.nmom; k++) {
                    double pm = c.pmom[(size_t)lc * ldp + k];
                    if (!(pm >= -1.0 && pm <= 1.0)) badl = 1;
                }
            }
            if (c.fbeam < 0.0 || (c.fbeam > 0.0 && !(c.umu0 > 0.0 && c.umu0 <= 1.0))) badl = 1;
            // albedo = SBD_SURFACE(s): BRDF surface s
            if (c.albedo < 0.0) {
                surf = (int)(-c.albedo) - 1;
                if (!a.sf_bdr || surf >= a.sf_count || (double)(surf + 1) != -c.albedo ||
                    (NU > 0 && a.sf_modes != N)) badl = 1;
            } else if (!(c.albedo <= 1.0)) badl = 1;
            if (c.fisot < 0.0) badl = 1;
            if (c.plank && (bp.wvnmlo < 0.0 || bp.wvnmhi <= bp.wvnmlo || bp.temis < 0.0 ||
                            bp.temis > 1.0 || bp.btemp < 0.0 || bp.ttemp < 0.0)) badl = 1;
            // device-pointer callers: a Planck bin needs a valid row of temper[ncol][L+1]
            if (c.plank && (!a.temper || bp.col < 0 || bp.col >= a.d.ncol)) badl = 1;
            if (__any_sync(FULLMASK, badl)) status = SBD_BIN_BAD_INPUT;
            // beam angle = quadrature angle (disort.f:2641-2650)
            int clash = 0;
            if (c.fbeam > 0.0)
                for (int i = lane; i < n; i += 32)
                    if (fabs(c.umu0 - cs.mu[i]) / c.umu0 < 1.e-4) clash = 1;
            if (!status && __any_sync(FULLMASK, clash)) status = SBD_BIN_ANGLE_CLASH;
        }

        // ---- SETDIS prologue: cumulative depths, truncation (:2546-2605) --
        if (lane == 0) {
            double tc = 0.0, tp = 0.0, abstau = 0.0;
            int ncut = L;
            w.tauc[0] = 0.0; w.taucpr[0] = 0.0;
            for (int lc = 0; lc < L; lc++) {
                double s = c.ssalb[lc];
                if (s == 1.0) s = 1.0 - kDither;
                double dt = c.dtauc[lc];
                tc += dt;                               // TAUC uses unclipped DTAUC
                if (dt < 0.0) dt = 0.0;
                if (abstau < 10.0) ncut = lc + 1;
                abstau += (1. - s) * dt;
                double f = c.pmom[(size_t)lc * ldp + N];
                tp += (1. - f * s) * dt;
                w.tauc[lc + 1] = tc; w.taucpr[lc + 1] = tp;
            }
            int lyrcut = (abstau >= 10.0 && !c.plank && L > 1);
            if (!lyrcut) ncut = L;
            w.pq[0] = ncut; w.pq[1] = lyrcut;
        }
        __syncwarp();
        c.ncut = w.pq[0]; c.lyrcut = w.pq[1];
        __syncwarp();
        const int ncut = c.ncut;

        // level -> layer map and scaled level depths (disort.f:2610-2625)
        int badtau = 0;
        for (int lu = lane; lu < NT; lu += 32) {
            double ut = a.d.ntau > 0 ? a.utau[(size_t)src * NT + lu] : w.tauc[lu];
            if (a.d.ntau > 0 && fabs(ut - w.tauc[L]) <= 1.e-4) ut = w.tauc[L];
            if (a.d.ntau > 0 && !(ut >= 0.0 && ut <= w.tauc[L])) badtau = 1;
            int lc;
            for (lc = 1; lc <= L; lc++)
                if (ut >= w.tauc[lc - 1] && ut <= w.tauc[lc]) break;
            if (lc > L) lc = L;
            w.layru[lu] = lc;
        }
        if (__any_sync(FULLMASK, badtau)) status = SBD_BIN_BAD_INPUT;

        // Planck function at the levels (disort.f:556-571)
        c.tplank = 0.0; c.bplank = 0.0;
        if (c.plank && !status) {
            const double *tp = a.temper + (size_t)bp.col * (L + 1);
            for (int lev = lane; lev <= L; lev += 32)
                w.pk[lev] = plkavg_dev(bp.wvnmlo, bp.wvnmhi, tp[lev]);
            c.tplank = bp.temis * plkavg_dev(bp.wvnmlo, bp.wvnmhi, bp.ttemp);
            c.bplank = plkavg_dev(bp.wvnmlo, bp.wvnmhi, bp.btemp);
        }
        // zero the outputs (ZEROAL, disort.f:518); levels below NCUT stay 0
        for (int lu = lane; lu < NT; lu += 32) {
            if (o_rfldir) o_rfldir[lu] = 0.0;
            if (o_rfldn) o_rfldn[lu] = 0.0;
            if (o_flup) o_flup[lu] = 0.0;
            if (o_dfdt) o_dfdt[lu] = 0.0;
            if (o_uavg) o_uavg[lu] = 0.0;
        }
        double *o_uu = (NU > 0 && a.uu && have) ? a.uu + (size_t)bin * a.d.nphi * a.uu_nt * NU : nullptr;
        if (o_uu)
            for (int e = lane; e < a.d.nphi * a.uu_nt * NU; e += 32) o_uu[e] = 0.0;

        // number of azimuth modes (disort.f:577-586); flux-only runs need m = 0 only
        int naz = 0;
        if (NU > 0) {
            naz = N - 1;
            const double u0 = a.umu[0], u1 = NU > 1 ? a.umu[1] : 0.0;
            if (c.fbeam == 0.0 || fabs(1. - c.umu0) < 1.e-5 ||
                (NU == 1 && fabs(1. - u0) < 1.e-5) || (NU == 1 && fabs(1. + u0) < 1.e-5) ||
                (NU == 2 && fabs(1. + u0) < 1.e-5 && fabs(1. - u1) < 1.e-5))
                naz = 0;
        }
        const int R = n + N, C = 2 * N + 1;
        const unsigned rC = div_magic(C);
        const double *ylm_smem = ylm_s;        // (cs.ylm points at the last mode of the previous bin)
        int kconv = 0;

      for (int mazim = 0; ; mazim++) {
        const bool mact = !status && mazim <= naz;       // this warp still has a mode to do
        if (!(SYNC ? __syncthreads_or(mact) : (int)mact)) break;
        c.mazim = mazim;
        c.delm0 = (mazim == 0) ? 1.0 : 0.0;
        if (surf >= 0 && !status) {       // SURFAC's tables of this mode (disort.f:3765-3907)
            c.bdr = a.sf_bdr + ((size_t)surf * a.sf_modes + mazim) * n * (n + 1);
            c.bem = a.sf_bem + (size_t)surf * n;
            if (NU > 0) {
                c.rmu = a.sf_rmu + ((size_t)surf * a.sf_modes + mazim) * NU * (n + 1);
                c.emu = a.sf_emu + (size_t)surf * NU;
            }
        }
        cs.ylm = (mazim == 0) ? ylm_smem : a.ylmc + (size_t)mazim * N * n;
        const double *ylmu_m = (NU > 0) ? a.ylmu + (size_t)mazim * N * NU : nullptr;
        // Y_l^m(-mu0) (LEPOLY, disort.f:599-605)
        if (lane == 0 && c.fbeam > 0.0) lepoly_one(mazim, N, -c.umu0, w.y0);
        __syncwarp();

        double xr0c = 0, xr1c = 0, xr0n = 0, xr1n = 0;
        int cur = 0;
        double bnd_up = 0.0;

        // ================= downward sweep ================================
        if (mact) status = solve_layer(c, cs, w, 0, cur, xr0c, xr1c, lane);
        if (mact && !status) {
            store_layer(ll, scr, w, cur, xr0c, xr1c, lane);
            // top boundary rows (disort.f:2887-2915, :3547-3550)
            for (int e = lane; e < n * C; e += 32) {
                int r = fdiv(e, rC), j = e - r * C;
                double v;
                if (j < N) {
                    double g = gc_elem(w.Gp[cur], w.Gm[cur], n, r, j);
                    v = (j < n) ? g * w.ek[cur][n - 1 - j] : g;
                } else if (j < 2 * N) {
                    v = 0.0;
                } else {
                    // disort.f:3445 (m > 0) / :3547-3550 (m = 0)
                    v = (mazim == 0 ? c.fisot + c.tplank : 0.0) - w.zz[cur][r] - w.zp0[cur][r];
                }
                w.W[r * C + j] = v;
            }
        }
        for (int lc = 0; ; lc++) {
            const bool lact = mact && !status && lc < ncut - 1;
            if (!(SYNC ? __syncthreads_or(lact) : (int)lact)) break;
            if (!lact) continue;
            const int nxt = cur ^ 1;
            status = solve_layer(c, cs, w, lc + 1, nxt, xr0n, xr1n, lane);
            if (status) continue;
            store_layer(ll, scr + (size_t)(lc + 1) * ll.stride, w, nxt, xr0n, xr1n, lane);
            // interface rows between layer lc and lc+1 (disort.f:2846-2882, :3585-3593)
            const double tb = w.taucpr[lc + 1];
            const double eb = (c.fbeam > 0.0) ? exp(-tb / c.umu0) : 0.0;
            for (int e = lane; e < N * C; e += 32) {
                int r = fdiv(e, rC), j = e - r * C;
                double v;
                if (j < N) {
                    double g = gc_elem(w.Gp[cur], w.Gm[cur], n, r, j);
                    v = (j < n) ? g : g * w.ek[cur][j - n];
                } else if (j < 2 * N) {
                    int jj = j - N;
                    double g = gc_elem(w.Gp[nxt], w.Gm[nxt], n, r, jj);
                    v = (jj < n) ? -g * w.ek[nxt][n - 1 - jj] : -g;
                } else {
                    v = (w.zz[nxt][r] - w.zz[cur][r]) * eb + w.zp0[nxt][r] - w.zp0[cur][r] +
                        (xr1n - xr1c) * tb;
                }
                w.W[(n + r) * C + j] = v;
            }
            __syncwarp();
            status = eliminate(w.W, R, C, N, lane);
            if (status) continue;
            // keep the N pivot rows, carry the n remaining rows
            double *U = scr + (size_t)lc * ll.stride + ll.off_u;
            for (int e = lane; e < N * C; e += 32) U[e] = w.W[e];
            __syncwarp();
            for (int e = lane; e < n * C; e += 32) {
                int r = fdiv(e, rC), j = e - r * C;
                double v;
                if (j < N) v = w.W[(N + r) * C + N + j];
                else if (j < 2 * N) v = 0.0;
                else v = w.W[(N + r) * C + 2 * N];
                // source rows N..R-1 never overlap destination rows 0..n-1
                w.W[r * C + j] = v;
            }
            __syncwarp();
            cur = nxt; xr0c = xr0n; xr1c = xr1n;
        }

        // ================= bottom boundary + last layer ====================
        if (mact && !status) {
            const int lc = ncut - 1;
            const double tb = w.taucpr[ncut];
            const double eb = (c.fbeam > 0.0) ? exp(-tb / c.umu0) : 0.0;
            // Lambertian surface reflects the m = 0 mode only (disort.f:2929-2947); a BRDF
            // surface every mode
            const int refl = !c.lyrcut && (mazim == 0 || c.bdr);
            const double fref = 1.0 + c.delm0;
            // sum_k w_k mu_k GC(-mu_k, j): one value per column, and for the rhs (Lambertian)
            for (int j = lane; j <= N; j += 32) {
                double sacc = 0.0;
                if (refl && !c.bdr) {
                    for (int k = 0; k < n; k++) {
                        double g = (j < N) ? gc_elem(w.Gp[cur], w.Gm[cur], n, n - 1 - k, j)
                                           : (w.zz[cur][n - 1 - k] * eb + w.zp0[cur][n - 1 - k] + xr1c * tb);
                        sacc += cs.wt[k] * cs.mu[k] * g;
                    }
                }
                if (j < N) w.v1[j] = sacc; else w.v2[0] = sacc;   // v2[0]: rhs part
            }
            __syncwarp();
            for (int e = lane; e < n * C; e += 32) {
                int i = fdiv(e, rC), j = e - i * C;
                int r = n + i;
                double v;
                if (j < N) {
                    double g = gc_elem(w.Gp[cur], w.Gm[cur], n, r, j);
                    if (refl && c.bdr) {
                        double sacc = 0.0;
                        for (int k = 0; k < n; k++)
                            sacc += cs.wt[k] * cs.mu[k] * c.bdr[i * (n + 1) + 1 + k] *
                                    gc_elem(w.Gp[cur], w.Gm[cur], n, n - 1 - k, j);
                        g -= fref * sacc;
                    } else if (refl) g -= 2.0 * c.albedo * w.v1[j];
                    v = (j < n) ? g : g * w.ek[cur][j - n];
                } else if (j < 2 * N) {
                    v = 0.0;
                } else {
                    v = -w.zz[cur][r] * eb - (mazim == 0 ? w.zp0[cur][r] + xr1c * tb : 0.0);
                    if (refl && c.bdr) {
                        // SOLVE0 (disort.f:3452-3466 for m > 0, :3552-3578 for m = 0)
                        double sacc = 0.0;
                        for (int k = 0; k < n; k++)
                            sacc += cs.wt[k] * cs.mu[k] * c.bdr[i * (n + 1) + 1 + k] *
                                    (w.zz[cur][n - 1 - k] * eb +
                                     (mazim == 0 ? w.zp0[cur][n - 1 - k] + xr1c * tb : 0.0));
                        v += fref * sacc + c.bdr[i * (n + 1)] * c.umu0 * c.fbeam / kPiRef * eb +
                             c.delm0 * c.bem[i] * c.bplank;
                    } else if (refl)
                        v += 2.0 * c.albedo * w.v2[0] + c.albedo * c.umu0 * c.fbeam / kPiRef * eb +
                             (1.0 - c.albedo) * c.bplank;
                }
                w.W[(n + i) * C + j] = v;
            }
            __syncwarp();
            status = eliminate(w.W, N, C, N, lane);
            (void)lc;
        }

        // ================= upward sweep: back substitution + fluxes ========
        if (mact && !status) {
            for (int lc = ncut - 1; lc >= 0; lc--) {
                if (lc < ncut - 1) {
                    const double *U = scr + (size_t)lc * ll.stride + ll.off_u;
                    copy_in(w.W, U, N * C, lane);
                    load_layer(ll, scr + (size_t)lc * ll.stride, w, cur, xr0c, xr1c, lane);
                    __syncwarp();
                    // rhs -= U[:, N:2N] x_{lc+1}
                    for (int r = lane; r < N; r += 32) {
                        double acc = w.W[r * C + 2 * N];
                        for (int j = 0; j < N; j++) acc -= w.W[r * C + N + j] * w.xn[j];
                        w.W[r * C + 2 * N] = acc;
                    }
                    __syncwarp();
                }
                // upper-triangular solve
                for (int j = N - 1; j >= 0; j--) {
                    double xj = w.W[j * C + 2 * N] / w.W[j * C + j];
                    __syncwarp();
                    if (lane == 0) w.xc[j] = xj;
                    for (int r = lane; r < j; r += 32) w.W[r * C + 2 * N] -= w.W[r * C + j] * xj;
                    __syncwarp();
                }
                for (int j = lane; j < N; j += 32) { LLs[(size_t)lc * N + j] = w.xc[j]; }
                __syncwarp();

                if (NU > 0) {
                    // surface-reflected term of USRINT (disort.f:4747-4778): downward
                    // intensities at the bottom boundary; Lambertian: m = 0 only, one value
                    // for all angles; BRDF: every mode, the weighted intensities are kept
                    // (in the Jacobi work vector, idle during the upward sweep) for RMU
                    if (lc == ncut - 1 && (mazim == 0 || c.bdr) && !c.lyrcut) {
                        const double tb = w.taucpr[ncut];
                        const double eb = (c.fbeam > 0.0) ? exp(-tb / c.umu0) : 0.0;
                        double dn = 0.0;
                        for (int i = lane; i < n; i += 32) {
                            double ud = 0.0;
                            for (int j = 0; j < n; j++)
                                ud += w.Gm[cur][i * n + j] * w.xc[n + j] * w.ek[cur][j] -
                                      w.Gp[cur][i * n + j] * w.xc[n - 1 - j];
                            ud += w.zz[cur][n - 1 - i] * eb + w.zp0[cur][n - 1 - i] + xr1c * tb;
                            dn += cs.wt[i] * cs.mu[i] * ud;
                            if (c.bdr) w.cs[i] = cs.wt[i] * cs.mu[i] * ud;
                        }
                        dn = warp_sum(dn);
                        bnd_up = 2.0 * c.albedo * dn + c.umu0 * c.fbeam / kPiRef * c.albedo * eb +
                                 (1.0 - c.albedo) * c.bplank;
                        __syncwarp();
                    }
                    user_terms(c, cs, w, ll, scr + (size_t)lc * ll.stride, ylmu_m, NU, lc, cur,
                               xr0c, xr1c, lane);
                }

                // fluxes at the levels that live in this layer (FLUXES), m = 0 only
                for (int lu = 0; lu < NT && mazim == 0; lu++) {
                    if (w.layru[lu] != lc + 1) continue;
                    double ut = a.d.ntau > 0 ? a.utau[(size_t)src * NT + lu] : w.tauc[lu];
                    if (a.d.ntau > 0 && fabs(ut - w.tauc[L]) <= 1.e-4) ut = w.tauc[L];
                    double ss = c.ssalb[lc]; if (ss == 1.0) ss = 1.0 - kDither;
                    const double f = c.pmom[(size_t)lc * ldp + N];
                    const double utp = w.taucpr[lc] + (1. - ss * f) * (ut - w.tauc[lc]);
                    // a_j = x+_j exp(-k (t - t_top)), b_j = x-_j exp(-k (t_bot - t))
                    for (int j = lane; j < n; j += 32) {
                        double k = w.kk[cur][j];
                        w.v1[j] = w.xc[n + j] * exp(-k * (utp - w.taucpr[lc]));
                        w.v2[j] = w.xc[n - 1 - j] * exp(-k * (w.taucpr[lc + 1] - utp));
                    }
                    __syncwarp();
                    double fact = 0.0, dirint = 0.0, fldir = 0.0, rfldir = 0.0;
                    if (c.fbeam > 0.0) {
                        fact = exp(-utp / c.umu0);
                        dirint = c.fbeam * fact;
                        fldir = c.umu0 * (c.fbeam * fact);
                        rfldir = c.umu0 * c.fbeam * exp(-ut / c.umu0);
                    }
                    double up = 0.0, dn = 0.0, av = 0.0;
                    for (int i = lane; i < n; i += 32) {
                        double uu = 0.0, ud = 0.0;
                        for (int j = 0; j < n; j++) {
                            double gp = w.Gp[cur][i * n + j], gm = w.Gm[cur][i * n + j];
                            uu += gp * w.v1[j] - gm * w.v2[j];
                            ud += gm * w.v1[j] - gp * w.v2[j];
                        }
                        uu += w.zz[cur][n + i] * fact + w.zp0[cur][n + i] + xr1c * utp;
                        ud += w.zz[cur][n - 1 - i] * fact + w.zp0[cur][n - 1 - i] + xr1c * utp;
                        double wm = cs.wt[i] * cs.mu[i];
                        up += wm * uu; dn += wm * ud; av += cs.wt[i] * (uu + ud);
                    }
                    up = warp_sum(up); dn = warp_sum(dn); av = warp_sum(av);
                    if (lane == 0) {
                        const double pi = kPiRef;
                        double flup = 2. * pi * up, fldn = 2. * pi * dn;
                        double fdntot = fldn + fldir;
                        double uavg = (2. * pi * av + dirint) / (4. * pi);
                        double plsorc = xr0c + xr1c * utp;
                        if (o_rfldir) o_rfldir[lu] = rfldir;
                        if (o_rfldn) o_rfldn[lu] = fdntot - rfldir;
                        if (o_flup) o_flup[lu] = flup;
                        if (o_uavg) o_uavg[lu] = uavg;
                        if (o_dfdt) o_dfdt[lu] = (1. - ss) * 4. * pi * (uavg - plsorc);
                    }
                    __syncwarp();
                }
                for (int j = lane; j < N; j += 32) w.xn[j] = w.xc[j];
                __syncwarp();
            }
        }

        // ================= intensities at the user angles ====================
        if (mact && !status && NU > 0) {
            __threadfence_block();
            const double rpd = kPiRef / 180.0;
            double azerr = 0.0;
            for (int e = lane; e < NT * NU; e += 32) {
                const int lu = e / NU, iu = e - lu * NU;
                const int slot = a.uu_slot[lu];
                if (slot < 0) continue;     // level not wanted (sbd_set_radiance_levels): stays zero
                const int lyu = w.layru[lu];
                double ut = a.d.ntau > 0 ? a.utau[(size_t)src * NT + lu] : w.tauc[lu];
                if (a.d.ntau > 0 && fabs(ut - w.tauc[L]) <= 1.e-4) ut = w.tauc[L];
                double ss = c.ssalb[lyu - 1]; if (ss == 1.0) ss = 1.0 - kDither;
                const double f = c.pmom[(size_t)(lyu - 1) * ldp + N];
                const double utp = w.taucpr[lyu - 1] + (1. - ss * f) * (ut - w.tauc[lyu - 1]);
                const double val = usrint_one(c, w, ll, scr, NU, lu, iu, a.umu[iu], utp, c.fisot, bnd_up);
                // Fourier sum over azimuth (disort.f:767-825)
                for (int j = 0; j < a.d.nphi; j++) {
                    double *pu = o_uu + ((size_t)j * a.uu_nt + slot) * NU + iu;
                    if (mazim == 0) {
                        *pu = val;
                    } else {
                        const double azterm = val * cos(mazim * (rpd * (a.phi[j] - bp.phi0)));
                        const double unew = *pu + azterm;
                        *pu = unew;
                        // RATIO(|AZTERM|, |UU|), disort.f:6159
                        const double aa = fabs(azterm), bb = fabs(unew);
                        double rr;
                        if (aa == 0.0) rr = (bb == 0.0) ? 1.0 : 0.0;
                        else if (bb == 0.0) rr = 1.79e308;
                        else rr = aa / bb;
                        azerr = fmax(azerr, rr);
                    }
                }
            }
            if (mazim > 0) {
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) azerr = fmax(azerr, __shfl_xor_sync(FULLMASK, azerr, o));
                if (azerr <= bp.accur) kconv++;
                if (kconv >= 2) naz = mazim;       // converged: no further modes (disort.f:821-823)
            }
            __syncwarp();
        }
      }   // azimuth modes
        if (lane == 0 && have) a.status[bin] = status;
        __syncwarp();
    }
}

cudaError_t launch_generic(const LaunchArgs &a, int warps, int grid, cudaStream_t st)
{
    size_t smem = generic_smem_bytes(a.d.nstr, a.d.nlyr,
                                     a.d.ntau > 0 ? a.d.ntau : a.d.nlyr + 1, warps);
    // CTA-synchronous loops pay off when several CTAs share an SM (small NSTR: +27 % on
    // the NSTR=8 radiance set); at NSTR=32 one 4-warp CTA fills the SM's shared memory
    // and the barriers only add idle time (-10 %)
    const bool sync = a.d.nstr <= 20;
    auto kern = sync ? disort_generic_kernel<true> : disort_generic_kernel<false>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    kern<<<grid, warps * 32, smem, st>>>(a);
    return cudaGetLastError();
}

}  // namespace sbd
