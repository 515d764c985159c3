// Nakajima-Tanaka intensity corrections (CORINT=.TRUE.): the TMS single-scattering
// correction and the IMS secondary-scattering correction of the aureole, applied to
// the intensities the solve kernel has just written.
//
// Reference: INTCOR disort.f:2044-2297, SINSCA :2996-3097, SECSCA :2299-2452,
// XIFUNC :4795-4858; switched off inside DISORT for flux-only calls, FBEAM = 0 and
// non-scattering media (disort.f:2695-2696).  SBDART passes NMOM = 299 moments when
// the option is on (drt.f:490-491).
//
// Layout: one CTA per bin, one warp per (user angle, azimuth) pair.
//   * phase functions of all layers: lanes own layers, the Legendre recurrence runs in
//     every lane (pmom rows are read k-contiguous);
//   * TMS at a level: the reference evaluates SINSCA twice (exact phase function with the
//     unscaled albedo, delta-M phase function with the scaled one) over the same
//     exponentials; here the two layer weights are subtracted first and ONE lane-parallel
//     sum over layers with a warp reduction gives the correction;
//   * IMS at a level: lanes own the moments k >= NSTR (pmom read coalesced across lanes),
//     the Legendre values of the pair sit in shared memory.
#include <math.h>

#include "sbd_internal.h"

namespace sbd {

namespace {

constexpr int kWarps = 4;

__device__ __forceinline__ double wsum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// XIFUNC (disort.f:4795-4858) for umu2 == umu3, the only form SECSCA uses
__device__ double xifunc_eq(double umu1, double umu2, double tau)
{
    const double exp1 = exp(-tau / umu1);
    if (umu1 == umu2) return tau * tau * exp1 / (2. * umu1 * umu2);
    const double x1 = 1. / umu1 - 1. / umu2;
    return ((tau - 1. / x1) * exp(-tau / umu2) + exp1 / x1) / (x1 * umu1 * umu2);
}

__global__ void __launch_bounds__(kWarps * 32)
intcor_kernel(const LaunchArgs a)
{
    extern __shared__ double sm[];
    const int bin = blockIdx.x;
    const int N = a.d.nstr, L = a.d.nlyr, NT = L + 1, NU = a.d.numu, NP = a.d.nphi;
    const int nmom = a.d.nmom, ldp = nmom + 1;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (bin >= (a.nbins_dev ? *a.nbins_dev : a.d.nbins)) return;
    if (a.status[bin] != 0) return;
    const int src = a.binmap ? a.binmap[bin] : bin;
    const sbd_bin bp = a.bins[src];
    const double fbeam = bp.fbeam, umu0 = bp.umu0;
    if (!(fbeam > 0.0)) return;
    const double *dtauc = a.dtauc + (size_t)src * L, *ssalb = a.ssalb + (size_t)src * L;
    const double *pmom = a.pmom + (size_t)src * L * ldp;
    double *uu = a.uu + (size_t)bin * NP * a.uu_nt * NU;

    // ---- per-bin layer quantities (SETDIS, disort.f:2546-2625)
    double *ss = sm, *dt = ss + L, *tauc = dt + L, *taucpr = tauc + (L + 1), *flyr = taucpr + (L + 1),
           *oprim = flyr + L, *misc = oprim + L;                  // misc: ncut, lyrcut, yessct
    double *wbase = misc + 4;
    double *wdiff = wbase + (size_t)warp * (L + ldp);             // [L]   omega phi_exact - omega' phi_deltaM
    double *plk = wdiff + L;                                      // [nmom+1] P_k(cos scattering angle)
    if (threadIdx.x == 0) {
        double tc = 0., tp = 0., abstau = 0., yes = 0.;
        int ncut = L;
        tauc[0] = 0.; taucpr[0] = 0.;
        for (int lc = 0; lc < L; lc++) {
            double s = ssalb[lc];
            if (s == 1.0) s = 1.0 - kDither;
            double d = dtauc[lc];
            tc += d;
            if (d < 0.0) d = 0.0;
            yes += s;
            if (abstau < 10.0) ncut = lc + 1;
            abstau += (1. - s) * d;
            const double f = pmom[(size_t)lc * ldp + N];
            tp += (1. - f * s) * d;
            ss[lc] = s; dt[lc] = d; flyr[lc] = f;
            oprim[lc] = s * (1. - f) / (1. - f * s);
            tauc[lc + 1] = tc; taucpr[lc + 1] = tp;
        }
        const int lyrcut = (abstau >= 10.0 && !bp.plank && L > 1);
        if (!lyrcut) ncut = L;
        misc[0] = ncut; misc[1] = lyrcut; misc[2] = yes;
    }
    __syncthreads();
    const int ncut = (int)misc[0], lyrcut = (int)misc[1];
    if (misc[2] == 0.0) return;

    const double pi = kPiRef, rpd = pi / 180.0;
    for (int pair = warp; pair < NU * NP; pair += kWarps) {
        const int iu = pair / NP, jp = pair - iu * NP;
        const double umu = a.umu[iu];
        const double ctheta = -umu0 * umu +
            sqrt((1. - umu0 * umu0) * (1. - umu * umu)) * cos(rpd * (a.phi[jp] - bp.phi0));
        // ---- phase functions of the layers this lane owns (disort.f:2165-2200)
        for (int lc = lane; lc < ncut; lc += 32) {
            const double f = flyr[lc];
            const double *pm = pmom + (size_t)lc * ldp;
            double pa = 1., pd = 1., plm1 = 1., plm2 = 0.;
            for (int k = 1; k <= nmom; k++) {
                const double pl = ((2 * k - 1) * ctheta * plm1 - (k - 1) * plm2) / k;
                plm2 = plm1; plm1 = pl;
                const double x = pm[k];
                pa += (2 * k + 1) * pl * x;
                if (k <= N - 1) pd += (2 * k + 1) * pl * (x - f) / (1. - f);
            }
            wdiff[lc] = ss[lc] * (pa / (1. - f * ss[lc])) - oprim[lc] * pd;
        }
        {   // Legendre values for the IMS term
            double plm1 = 1., plm2 = 0.;
            if (lane == 0) plk[0] = 1.;
            for (int k = 1; k <= nmom; k++) {
                const double pl = ((2 * k - 1) * ctheta * plm1 - (k - 1) * plm2) / k;
                plm2 = plm1; plm1 = pl;
                if ((k & 31) == lane) plk[k] = pl;
            }
        }
        __syncwarp();

        // ---- TMS (two SINSCA calls of the reference folded into one sum)
        for (int lu = 0; lu < NT; lu++) {
            const int slot = a.uu_slot[lu];
            if (slot < 0) continue;
            // level -> layer: the first layer whose interval holds the level (disort.f:2610-2625)
            int lyu = 1;
            const double ut = tauc[lu];
            while (lyu < L && !(ut >= tauc[lyu - 1] && ut <= tauc[lyu])) lyu++;
            if (lyrcut && !(lyu < ncut)) continue;
            const double utp = taucpr[lyu - 1] + (1. - ss[lyu - 1] * flyr[lyu - 1]) * (ut - tauc[lyu - 1]);
            const double e0 = exp(-utp / umu0);
            double s = 0.0, pref;
            if (fabs(umu + umu0) <= kDither) {
                for (int lyr = 1 + lane; lyr <= lyu; lyr += 32)
                    s += wdiff[lyr - 1] * (lyr < lyu ? taucpr[lyr] - taucpr[lyr - 1] : utp - taucpr[lyu - 1]);
                pref = fbeam / (4. * pi * umu0) * e0;
            } else if (umu > 0.) {
                for (int lyr = lyu + lane; lyr <= ncut; lyr += 32) {
                    const double ea = (lyr == lyu) ? e0
                        : exp(-((taucpr[lyr - 1] - utp) / umu + taucpr[lyr - 1] / umu0));
                    const double eb = exp(-((taucpr[lyr] - utp) / umu + taucpr[lyr] / umu0));
                    s += wdiff[lyr - 1] * (ea - eb);
                }
                pref = fbeam / (4. * pi * (1. + umu / umu0));
            } else {
                for (int lyr = lyu - lane; lyr >= 1; lyr -= 32) {
                    const double ea = (lyr == lyu) ? e0
                        : exp(-((taucpr[lyr] - utp) / umu + taucpr[lyr] / umu0));
                    const double eb = exp(-((taucpr[lyr - 1] - utp) / umu + taucpr[lyr - 1] / umu0));
                    s += wdiff[lyr - 1] * (ea - eb);
                }
                pref = fbeam / (4. * pi * (1. + umu / umu0));
            }
            s = wsum(s);
            if (lane == 0) uu[((size_t)jp * a.uu_nt + slot) * NU + iu] += pref * s;
        }

        // ---- IMS: aureole within 10 degrees of the direct beam, downward directions
        if (umu < 0.) {
            const double theta0 = acos(-umu0) / rpd, thetap = acos(umu) / rpd;
            if (fabs(theta0 - thetap) <= 10.) {
                for (int lu = 1; lu < NT; lu++) {        // level 0 has utau = 0 <= DITHER: skipped
                    const int slot = a.uu_slot[lu];
                    if (slot < 0) continue;
                    int lyu = 1;
                    const double ut = tauc[lu];
                    while (lyu < L && !(ut >= tauc[lyu - 1] && ut <= tauc[lyu])) lyu++;
                    if (lyrcut && !(lyu < ncut)) continue;
                    const double dlast = ut - tauc[lyu - 1];
                    double wbar = ss[lyu - 1] * dlast, fbar = flyr[lyu - 1] * wbar, stau = dlast;
                    for (int lyr = 1; lyr <= lyu - 1; lyr++) {
                        wbar += ss[lyr - 1] * dt[lyr - 1];
                        fbar += ss[lyr - 1] * dt[lyr - 1] * flyr[lyr - 1];
                        stau += dt[lyr - 1];
                    }
                    const double zero = (double)1E-4f;
                    if (wbar <= zero || fbar <= zero || stau <= zero || fbeam <= zero) continue;
                    fbar = fbar / wbar;
                    wbar = wbar / stau;
                    const double den = fbar * wbar * stau;
                    double ps = 0.0;
                    for (int k = 1 + lane; k <= nmom; k += 32) {
                        double g = 1.0;
                        if (k >= N) {
                            g = pmom[(size_t)(lyu - 1) * ldp + k] * ss[lyu - 1] * dlast;
                            for (int lyr = 1; lyr <= lyu - 1; lyr++)
                                g += pmom[(size_t)(lyr - 1) * ldp + k] * ss[lyr - 1] * dt[lyr - 1];
                            g = (den <= zero) ? 0.0 : g / den;
                        }
                        ps += (2. * g - g * g) * (2 * k + 1) * plk[k];
                    }
                    const double pspike = 1. + wsum(ps);
                    if (lane == 0) {
                        const double umu0p = umu0 / (1. - fbar * wbar);
                        const double d = fbeam / (4. * pi) * (fbar * wbar) * (fbar * wbar) / (1. - fbar * wbar) *
                                         pspike * xifunc_eq(-umu, umu0p, ut);
                        uu[((size_t)jp * a.uu_nt + slot) * NU + iu] -= d;
                    }
                }
            }
        }
        __syncwarp();
    }
}

}  // namespace

size_t intcor_smem_bytes(int L, int nmom)
{
    return 8 * ((size_t)6 * L + 2 + 4 + (size_t)kWarps * (L + nmom + 1));
}

cudaError_t launch_intcor(const LaunchArgs &a, cudaStream_t st)
{
    const size_t smem = intcor_smem_bytes(a.d.nlyr, a.d.nmom);
    cudaError_t e = cudaFuncSetAttribute(intcor_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    intcor_kernel<<<a.d.nbins, kWarps * 32, smem, st>>>(a);
    return cudaGetLastError();
}

}  // namespace sbd
