// C ABI of the B200 batched discrete-ordinate solver (include/sbdart_b200.h).
// Host side only: argument checks, quadrature / Legendre tables, device
// buffers, launches.  There is deliberately no CPU compute path here: if no
// CUDA device is usable every entry returns SBD_ERR_CUDA.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <map>
#include <mutex>
#include <vector>

#include "sbd_handle.h"
#include "sbd_internal.h"

using namespace sbd;

namespace {

// Gauss-Legendre rule on (0,1) with m points, nodes ascending (the rule
// QGAUSN produces, disort.f:5984-6157).  Newton on P_m with the standard
// cosine first guess; nodes come in +- pairs on (-1,1).
void gauss01(int m, double *mu, double *wt)
{
    const double pi = 3.14159265358979323846;
    for (int k = 0; k < (m + 1) / 2; k++) {
        double x = cos(pi * (k + 0.75) / (m + 0.5));   // descending from +1
        double pp = 1.0;
        for (int it = 0; it < 100; it++) {
            double p0 = 1.0, p1 = x;
            for (int l = 2; l <= m; l++) {
                double p2 = ((2 * l - 1) * x * p1 - (l - 1) * p0) / l;
                p0 = p1; p1 = p2;
            }
            if (m == 1) { p1 = x; p0 = 1.0; }
            pp = m * (x * p1 - p0) / (x * x - 1.0);
            double dx = p1 / pp;
            x -= dx;
            if (fabs(dx) < 1e-16) break;
        }
        {   // derivative at the converged node
            double p0 = 1.0, p1 = x;
            for (int l = 2; l <= m; l++) {
                double p2 = ((2 * l - 1) * x * p1 - (l - 1) * p0) / l;
                p0 = p1; p1 = p2;
            }
            pp = m * (x * p1 - p0) / (x * x - 1.0);
        }
        double w = 2.0 / ((1.0 - x * x) * pp * pp);
        // x is the k-th largest node on (-1,1); map to (0,1)
        mu[m - 1 - k] = 0.5 * x + 0.5;
        wt[m - 1 - k] = 0.5 * w;
        mu[k] = 0.5 * (-x) + 0.5;
        wt[k] = 0.5 * w;
    }
}

// normalised associated Legendre functions Y_l^m(x), l = 0..lmax, for
// m = 0..M-1 (the functions LEPOLY builds mode by mode, disort.f:5286-5408):
//   Y_l^m = sqrt((l-m)!/(l+m)!) P_l^m, Y_m^m = -sqrt((2m-1)/(2m)) sqrt(1-x^2) Y_{m-1}^{m-1}
// out[(m*(lmax+1) + l)*nx + i]
void legendre_table(int M, int lmax, int nx, const double *x, double *out)
{
    std::vector<double> diag(nx, 1.0);
    for (int m = 0; m < M; m++) {
        for (int i = 0; i < nx; i++) {
            if (m > 0)
                diag[i] = -sqrt((double)(2 * m - 1)) / sqrt((double)(2 * m)) *
                          sqrt(1.0 - x[i] * x[i]) * diag[i];
            double *o = out + (size_t)m * (lmax + 1) * nx;
            for (int l = 0; l < m && l <= lmax; l++) o[(size_t)l * nx + i] = 0.0;
            if (m <= lmax) o[(size_t)m * nx + i] = diag[i];
            if (m + 1 <= lmax)
                o[(size_t)(m + 1) * nx + i] = sqrt((double)(2 * m + 1)) * x[i] * diag[i];
            for (int l = m + 2; l <= lmax; l++) {
                double t1 = sqrt((double)(l - m)) * sqrt((double)(l + m));
                double t2 = sqrt((double)(l - m - 1)) * sqrt((double)(l + m - 1));
                o[(size_t)l * nx + i] = ((2 * l - 1) * x[i] * o[(size_t)(l - 1) * nx + i] -
                                         t2 * o[(size_t)(l - 2) * nx + i]) / t1;
            }
        }
    }
}

}  // namespace

static int check_dims(const sbd_dims *d)
{
    if (!d) return SBD_ERR_ARG;
    if (d->nbins < 0 || d->nlyr < 1 || d->nlyr > SBD_MAX_NLYR) return SBD_ERR_ARG;
    if (d->nstr < 4 || d->nstr > SBD_MAX_NSTR || (d->nstr & 1)) return SBD_ERR_ARG;
    if (d->nmom < d->nstr) return SBD_ERR_ARG;
    if (d->ntau < 0 || d->numu < 0 || d->nphi < 0 || d->ncol < 0) return SBD_ERR_ARG;
    if (d->numu > 0 && d->nphi < 1) return SBD_ERR_ARG;
    return SBD_SUCCESS;
}

extern "C" int sbd_abi_version(void) { return SBD_ABI_VERSION; }

#ifndef SBD_BUILD_ID
#define SBD_BUILD_ID "unknown"
#endif
extern "C" const char *sbd_build_id(void) { return SBD_BUILD_ID; }

extern "C" const char *sbd_status_string(int code)
{
    switch (code) {
    case SBD_SUCCESS: return "ok";
    case SBD_BIN_ANGLE_CLASH: return "beam angle equals a quadrature angle; change NSTR (disort.f:2645)";
    case SBD_BIN_BAD_INPUT: return "input out of range (CHEKIN, disort.f:4864)";
    case SBD_BIN_EIG_FAIL: return "eigen-solve failed (ASYMTX, disort.f:3254)";
    case SBD_BIN_SINGULAR: return "boundary system singular (SOLVE0, disort.f:3609)";
    case SBD_ERR_CUDA: return "CUDA device unavailable or runtime error";
    case SBD_ERR_ARG: return "bad argument";
    case SBD_ERR_UNSUPPORTED: return "option outside the supported hot path";
    default: return "unknown";
    }
}

extern "C" int sbd_quadrature(int m, double *mu, double *wt)
{
    if (m < 1 || !mu || !wt) return SBD_ERR_ARG;
    gauss01(m, mu, wt);
    return SBD_SUCCESS;
}

extern "C" int sbd_create(sbd_handle **out, int device)
{
    if (!out) return SBD_ERR_ARG;
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) return SBD_ERR_CUDA;
    if (cudaSetDevice(device) != cudaSuccess) return SBD_ERR_CUDA;
    sbd_handle *h = new sbd_handle();
    h->device = device;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) { delete h; return SBD_ERR_CUDA; }
    h->sm_count = prop.multiProcessorCount;
    h->smem_optin = prop.sharedMemPerBlockOptin;
    if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) { delete h; return SBD_ERR_CUDA; }
    if (cudaStreamCreateWithFlags(&h->copy_in, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&h->copy_out, cudaStreamNonBlocking) != cudaSuccess) { delete h; return SBD_ERR_CUDA; }
    for (int i = 0; i < sbd_handle::kMaxChunks; i++) {
        cudaEventCreateWithFlags(&h->ev_in[i], cudaEventDisableTiming);
        cudaEventCreateWithFlags(&h->ev_k[i], cudaEventDisableTiming);
    }
    if (h->counter.reserve(256) != cudaSuccess || h->counter2.reserve(256) != cudaSuccess) { delete h; return SBD_ERR_CUDA; }
    if (cudaStreamCreateWithFlags(&h->stream2, cudaStreamNonBlocking) != cudaSuccess) { delete h; return SBD_ERR_CUDA; }
    cudaEventCreateWithFlags(&h->ev_misc, cudaEventDisableTiming);
    *out = h;
    return SBD_SUCCESS;
}

extern "C" void sbd_destroy(sbd_handle *h)
{
    if (!h) return;
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    for (auto &kv : h->tables) { cudaFree(kv.second.quad); cudaFree(kv.second.ylmc); }
    cudaStreamSynchronize(h->stream2);
    SbdDevBuf *bufs[] = { &h->surfaces, &h->redo, &h->redo2, &h->redo_scratch, &h->redo_scratch2, &h->scratch, &h->counter, &h->scratch2, &h->counter2, &h->ylmu, &h->angles, &h->d_dtauc, &h->d_ssalb,
                       &h->d_pmom, &h->d_bins, &h->d_temper, &h->d_utau, &h->d_out, &h->d_uu,
                       &h->d_status, &h->opt_tables, &h->opt_atm, &h->opt_misc, &h->opt_map, &h->opt_aero,
                       &h->d_uupack, &h->d_sel, &h->d_fluxpack };
    for (SbdDevBuf *b : bufs) b->release();
    if (h->host_stage) cudaFreeHost(h->host_stage);
    for (int i = 0; i < sbd_handle::kMaxChunks; i++) { cudaEventDestroy(h->ev_in[i]); cudaEventDestroy(h->ev_k[i]); }
    cudaEventDestroy(h->ev_misc);
    cudaStreamDestroy(h->stream2);
    cudaStreamDestroy(h->copy_in);
    cudaStreamDestroy(h->copy_out);
    cudaStreamDestroy(h->stream);
    delete h;
}

extern "C" int sbd_synchronize(sbd_handle *h)
{
    if (!h) return SBD_ERR_ARG;
    return cudaStreamSynchronize(h->stream) == cudaSuccess ? SBD_SUCCESS : SBD_ERR_CUDA;
}

extern "C" int sbd_set_radiance_levels(sbd_handle *h, const int32_t *levels, int32_t n)
{
    if (!h || n < 0 || (n > 0 && !levels)) return SBD_ERR_ARG;
    if (n == 0) { h->uu_mask[0] = h->uu_mask[1] = ~0ull; return SBD_SUCCESS; }
    unsigned long long m[2] = { 0, 0 };
    for (int i = 0; i < n; i++) {
        if (levels[i] < 0 || levels[i] >= 128) return SBD_ERR_ARG;
        m[levels[i] >> 6] |= 1ull << (levels[i] & 63);
    }
    h->uu_mask[0] = m[0]; h->uu_mask[1] = m[1];
    return SBD_SUCCESS;
}

extern "C" int sbd_set_radiance_layout(sbd_handle *h, int32_t packed)
{
    if (!h) return SBD_ERR_ARG;
    h->uu_packed = packed != 0;
    return SBD_SUCCESS;
}

// levels selected with sbd_set_radiance_levels, ascending
static std::vector<int32_t> selected_levels(const sbd_handle *h, int NT)
{
    std::vector<int32_t> sel;
    for (int lu = 0; lu < NT; lu++)
        if (lu >= 128 || ((h->uu_mask[lu >> 6] >> (lu & 63)) & 1ull)) sel.push_back(lu);
    return sel;
}

extern "C" int sbd_set_surfaces(sbd_handle *h, int32_t nsurf, int32_t nstr, int32_t nmodes, int32_t numu,
                                const double *bdr, const double *bem, const double *rmu, const double *emu)
{
    if (!h) return SBD_ERR_ARG;
    if (nsurf <= 0) { h->sf_count = 0; return SBD_SUCCESS; }
    if (nstr < 4 || nstr % 2 || nstr > SBD_MAX_NSTR || (nmodes != 1 && nmodes != nstr) || numu < 0 || !bdr || !bem ||
        (numu > 0 && (!rmu || !emu))) return SBD_ERR_ARG;
    const size_t n = nstr / 2;
    const size_t n_bdr = (size_t)nsurf * nmodes * n * (n + 1), n_bem = (size_t)nsurf * n;
    const size_t n_rmu = (size_t)nsurf * nmodes * numu * (n + 1), n_emu = (size_t)nsurf * numu;
    if (cudaSetDevice(h->device) != cudaSuccess) return SBD_ERR_CUDA;
    if (h->surfaces.reserve((n_bdr + n_bem + n_rmu + n_emu) * 8) != cudaSuccess) return SBD_ERR_CUDA;
    double *d = (double *)h->surfaces.p;
    // (launches in flight may still read the previous tables)
    if (cudaStreamSynchronize(h->stream) != cudaSuccess) return SBD_ERR_CUDA;
    if (h->stream2 && cudaStreamSynchronize(h->stream2) != cudaSuccess) return SBD_ERR_CUDA;
    if (cudaMemcpy(d, bdr, n_bdr * 8, cudaMemcpyHostToDevice) != cudaSuccess) return SBD_ERR_CUDA;
    if (cudaMemcpy(d + n_bdr, bem, n_bem * 8, cudaMemcpyHostToDevice) != cudaSuccess) return SBD_ERR_CUDA;
    if (numu > 0) {
        if (cudaMemcpy(d + n_bdr + n_bem, rmu, n_rmu * 8, cudaMemcpyHostToDevice) != cudaSuccess) return SBD_ERR_CUDA;
        if (cudaMemcpy(d + n_bdr + n_bem + n_rmu, emu, n_emu * 8, cudaMemcpyHostToDevice) != cudaSuccess) return SBD_ERR_CUDA;
    }
    h->sf_count = nsurf; h->sf_nstr = nstr; h->sf_modes = nmodes; h->sf_numu = numu;
    return SBD_SUCCESS;
}

extern "C" int sbd_set_corint(sbd_handle *h, int32_t on)
{
    if (!h) return SBD_ERR_ARG;
    h->corint = on != 0;
    return SBD_SUCCESS;
}

extern "C" void *sbd_stream(sbd_handle *h) { return h ? (void *)h->stream : nullptr; }

extern "C" int64_t sbd_kernel_launches(const sbd_handle *h) { return h ? h->launches : 0; }

int sbd_get_tables(sbd_handle *h, int N, SbdTables &t, cudaStream_t st)
{
    auto it = h->tables.find(N);
    if (it != h->tables.end()) { t = it->second; return SBD_SUCCESS; }
    const int n = N / 2;
    std::vector<double> quad(2 * n), ylm((size_t)N * N * n);
    gauss01(n, quad.data(), quad.data() + n);
    legendre_table(N, N - 1, n, quad.data(), ylm.data());
    SbdTables nt;
    if (cudaMalloc(&nt.quad, quad.size() * 8) != cudaSuccess) return SBD_ERR_CUDA;
    if (cudaMalloc(&nt.ylmc, ylm.size() * 8) != cudaSuccess) return SBD_ERR_CUDA;
    cudaMemcpyAsync(nt.quad, quad.data(), quad.size() * 8, cudaMemcpyHostToDevice, st);
    cudaMemcpyAsync(nt.ylmc, ylm.data(), ylm.size() * 8, cudaMemcpyHostToDevice, st);
    if (cudaStreamSynchronize(st) != cudaSuccess) return SBD_ERR_CUDA;
    h->tables[N] = nt;
    t = nt;
    return SBD_SUCCESS;
}

extern "C" int sbd_disort_batch_device(sbd_handle *h, const sbd_dims *dims, const double *dtauc,
                                       const double *ssalb, const double *pmom,
                                       const sbd_bin *bins, const double *temper,
                                       const double *utau, const double *umu, const double *phi,
                                       double *rfldir, double *rfldn, double *flup, double *dfdt,
                                       double *uavg, double *uu, int32_t *status, void *stream)
{
    if (!h) return SBD_ERR_ARG;
    int rc = check_dims(dims);
    if (rc) return rc;
    if (!dtauc || !ssalb || !pmom || !bins || !status) return SBD_ERR_ARG;
    if (dims->ntau > 0 && !utau) return SBD_ERR_ARG;
    const int NU = dims->numu;
    if (NU > 0) {
        // user angles are call-level parameters: HOST arrays even in the device variant
        if (!umu || !phi || !uu) return SBD_ERR_ARG;
        for (int iu = 0; iu < NU; iu++) {      // CHEKIN, disort.f:5022-5036
            if (!(umu[iu] >= -1.0 && umu[iu] <= 1.0) || umu[iu] == 0.0) return SBD_ERR_ARG;
            if (iu > 0 && umu[iu] < umu[iu - 1]) return SBD_ERR_ARG;
        }
        for (int j = 0; j < dims->nphi; j++)
            if (!(phi[j] >= 0.0 && phi[j] <= 360.0)) return SBD_ERR_ARG;
    }
    if (dims->nbins == 0) return SBD_SUCCESS;
    if (cudaSetDevice(h->device) != cudaSuccess) return SBD_ERR_CUDA;
    cudaStream_t st = stream ? (cudaStream_t)stream : h->stream;

    const int N = dims->nstr, L = dims->nlyr;
    const int NT = dims->ntau > 0 ? dims->ntau : L + 1;
    SbdTables tb;
    rc = sbd_get_tables(h, N, tb, st);
    if (rc) return rc;

    size_t smem_limit = h->smem_optin ? h->smem_optin : 48 * 1024;
    // register kernel: NSTR 4/8/16; radiances at the layer boundaries (the mode SBDART uses)
    // BRDF surfaces: the adding kernel (fluxes at the layer boundaries, NSTR 4/8/16) and the
    // general kernel implement them
    const bool brdf = h->sf_count > 0;
    if (brdf && (h->sf_nstr != N || (NU > 0 && (h->sf_modes != N || h->sf_numu != NU)))) return SBD_ERR_ARG;
    // adding kernel: fluxes at the layer boundaries, NSTR 4/8/16/20/24/32.  Tuning / comparison
    // knobs (not API): SBD_FORCE_ELIM = the elimination kernels instead, SBD_FORCE_GENERIC = the
    // general kernel for everything
    const bool adding_ok = adding_supported(N) && NU == 0 && dims->ntau == 0 && !getenv("SBD_FORCE_GENERIC");
    const bool adding = adding_ok && (brdf || !getenv("SBD_FORCE_ELIM"));
    // elimination register kernel: NSTR 4/8/16; radiances at the layer boundaries (the mode SBDART uses)
    // (radiance runs in the adding form: NSTR up to 32, BRDF surfaces too)
    const bool fast_rad = NU > 0 && dims->ntau == 0 && fast_rad_supported(N) && (!brdf || !getenv("SBD_RAD_ELIM"));
    // (a radiance run with very many user angles may not fit the register kernel's shared memory:
    // the general kernel takes it)
    const bool fast_fits = fast_smem_bytes(N, L, NT, 4, NU, dims->nphi) <= smem_limit;
    const bool fast = !adding && (NU > 0 ? fast_rad && fast_fits : (!brdf && fast_supported(N))) && !getenv("SBD_FORCE_GENERIC");
    // CTA-per-bin register kernel: NSTR 20/24/32, fluxes
    const bool wide = !adding && !fast && !brdf && wide_supported(N) && NU == 0 && !getenv("SBD_FORCE_GENERIC") &&
                      wide_smem_bytes(N, L, NT) <= smem_limit;
    int warps, grid;
    size_t slot;
    if (wide) {
        const size_t smem = wide_smem_bytes(N, L, NT);
        int cta_per_sm = (int)((smem_limit + 1024) / (smem + 1024));
        if (cta_per_sm > wide_ctas_per_sm()) cta_per_sm = wide_ctas_per_sm();     // register budget (launch bounds)
        if (cta_per_sm < 1) cta_per_sm = 1;
        grid = h->sm_count * cta_per_sm;
        warps = 1;                                   // scratch slots and work items per CTA
        slot = wide_slot_doubles(N, L);
    } else if (fast || adding) {
        // CTA shape: the preferred one (8 warps) unless 4-warp CTAs keep more warps
        // resident per SM under the shared-memory limit (deep atmospheres); the adding kernel
        // above NSTR = 16 runs 4-warp CTAs (register budget)
        warps = N > 16 ? 4 : fast_warps();
        const int wmax = adding ? adding_warps_per_sm(N) : (N > 16 ? 8 : 16);
        int cta_per_sm = 0;
        for (int wtry = warps; wtry >= 4; wtry /= 2) {
            const size_t smem = adding ? adding_smem_bytes(N, L, wtry) : fast_smem_bytes(N, L, NT, wtry, NU, dims->nphi);
            if (smem > smem_limit) continue;
            int c = (int)((smem_limit + 1024) / (smem + 1024));
            if (c > wmax / wtry) c = wmax / wtry;
            if (c * wtry > cta_per_sm * warps || cta_per_sm == 0) { cta_per_sm = c; warps = wtry; }
        }
        if (cta_per_sm == 0) return SBD_ERR_UNSUPPORTED;
        grid = h->sm_count * cta_per_sm;
        slot = adding ? adding_slot_doubles(N, L) : fast_slot_doubles(N, L, NU);
    } else {
        warps = generic_pick_warps(N, L, NT, smem_limit);
        if (warps == 0) return SBD_ERR_UNSUPPORTED;
        // small NSTR: several 4-warp CTAs per SM; large NSTR (one CTA fills the SM's shared
        // memory): as many warps as fit
        if (warps >= 8) warps = 4;
        size_t smem = generic_smem_bytes(N, L, NT, warps);
        int cta_per_sm = (int)((smem_limit + 1024) / (smem + 1024));
        if (cta_per_sm < 1) cta_per_sm = 1;
        if (cta_per_sm * warps > 16) cta_per_sm = 16 / warps > 0 ? 16 / warps : 1;
        grid = h->sm_count * cta_per_sm;
        slot = generic_slot_doubles(N, L, NU);
    }
    int need = (dims->nbins + warps - 1) / warps;
    if (grid > need) grid = need;

    LaunchArgs a;
    memset(&a, 0, sizeof a);
    a.d = *dims;
    a.dtauc = dtauc; a.ssalb = ssalb; a.pmom = pmom; a.bins = bins; a.temper = temper;
    a.utau = utau;
    if (NU > 0) {
        // Y_l^m at the user cosines for every azimuth mode (LEPOLY, disort.f:608-609)
        std::vector<double> tab((size_t)N * N * NU + NU + dims->nphi);
        legendre_table(N, N - 1, NU, umu, tab.data());
        double *ang = tab.data() + (size_t)N * N * NU;
        for (int iu = 0; iu < NU; iu++) ang[iu] = umu[iu];
        for (int j = 0; j < dims->nphi; j++) ang[NU + j] = phi[j];
        if (h->ylmu.reserve(tab.size() * 8) != cudaSuccess) return SBD_ERR_CUDA;
        if (cudaMemcpyAsync(h->ylmu.p, tab.data(), tab.size() * 8, cudaMemcpyHostToDevice, st) != cudaSuccess)
            return SBD_ERR_CUDA;
        a.ylmu = (const double *)h->ylmu.p;
        a.umu = a.ylmu + (size_t)N * N * NU;
        a.phi = a.umu + NU;
    }
    a.rfldir = rfldir; a.rfldn = rfldn; a.flup = flup; a.dfdt = dfdt; a.uavg = uavg; a.uu = uu;
    a.status = status;
    a.binmap = h->pending_binmap;      // set by the spectrum path for exactly one launch
    a.nbins_dev = h->pending_nbins_dev;
    h->pending_binmap = nullptr;
    h->pending_nbins_dev = nullptr;
    a.quad = tb.quad; a.ylmc = tb.ylmc;
    if (brdf) {
        const size_t n = N / 2;
        const double *d = (const double *)h->surfaces.p;
        a.sf_count = h->sf_count; a.sf_modes = h->sf_modes;
        a.sf_bdr = d;
        a.sf_bem = a.sf_bdr + (size_t)h->sf_count * h->sf_modes * n * (n + 1);
        a.sf_rmu = a.sf_bem + (size_t)h->sf_count * n;
        a.sf_emu = a.sf_rmu + (size_t)h->sf_count * h->sf_modes * h->sf_numu * (n + 1);
    }
    a.nslots = grid * warps;
    a.slot_stride = slot;
    a.nmodes = 1;
    a.uu_mask[0] = h->uu_mask[0]; a.uu_mask[1] = h->uu_mask[1];
    {   // where each output level's intensities go
        const std::vector<int32_t> sel = selected_levels(h, NT);
        const bool packed = (h->uu_packed || h->uu_packed_once) && (int)sel.size() < NT;
        for (int lu = 0; lu < SBD_MAX_NLYR + 2; lu++) a.uu_slot[lu] = -1;
        for (size_t s = 0; s < sel.size(); s++) a.uu_slot[sel[s]] = (short)(packed ? (int)s : sel[s]);
        a.uu_nt = packed ? (int)sel.size() : NT;
    }
    SbdDevBuf &scr = h->scratch_set ? h->scratch2 : h->scratch;
    SbdDevBuf &ctr = h->scratch_set ? h->counter2 : h->counter;
    if (scr.reserve(a.slot_stride * (size_t)a.nslots * 8) != cudaSuccess) return SBD_ERR_CUDA;
    a.scratch = (double *)scr.p;
    a.work_counter = (int *)ctr.p;
    if (cudaMemsetAsync(a.work_counter, 0, 4 * sizeof(int), st) != cudaSuccess) return SBD_ERR_CUDA;
    cudaError_t le;
    if (adding) {
        // bins whose TAUC is not monotone (negative optical depths) are appended to a list and
        // solved by the elimination kernel launched behind (normally the list stays empty and
        // that launch ends at once)
        SbdDevBuf &rl = h->scratch_set ? h->redo2 : h->redo;
        SbdDevBuf &rs = h->scratch_set ? h->redo_scratch2 : h->redo_scratch;
        if (rl.reserve((size_t)dims->nbins * sizeof(int)) != cudaSuccess) return SBD_ERR_CUDA;
        a.redo_count = a.work_counter + 1;
        a.redo_list = (int *)rl.p;
        le = launch_adding(a, warps, grid, st);
        if (le != cudaSuccess) return SBD_ERR_CUDA;
        h->launches += 1;
        LaunchArgs a2 = a;
        a2.redo_consume = true;
        a2.work_counter = a.work_counter + 2;
        if (brdf) {
            // (the elimination kernels are Lambertian: BRDF bins go to the general kernel)
            int w2 = generic_pick_warps(N, L, NT, smem_limit);
            if (w2 == 0) return SBD_ERR_UNSUPPORTED;
            if (w2 >= 8) w2 = 4;
            int g2 = 16;
            if (g2 > (dims->nbins + w2 - 1) / w2) g2 = (dims->nbins + w2 - 1) / w2;
            a2.slot_stride = generic_slot_doubles(N, L, NU);
            a2.nslots = g2 * w2;
            if (rs.reserve(a2.slot_stride * (size_t)a2.nslots * 8) != cudaSuccess) return SBD_ERR_CUDA;
            a2.scratch = (double *)rs.p;
            le = launch_generic(a2, w2, g2, st);
        } else if (fast_supported(N)) {
            int w2 = fast_warps();
            if (fast_smem_bytes(N, L, NT, w2, 0, 0) > smem_limit) w2 = 4;
            if (fast_smem_bytes(N, L, NT, w2, 0, 0) > smem_limit) return SBD_ERR_UNSUPPORTED;
            int g2 = 16;
            if (g2 > (dims->nbins + w2 - 1) / w2) g2 = (dims->nbins + w2 - 1) / w2;
            a2.slot_stride = fast_slot_doubles(N, L, 0);
            a2.nslots = g2 * w2;
            if (rs.reserve(a2.slot_stride * (size_t)a2.nslots * 8) != cudaSuccess) return SBD_ERR_CUDA;
            a2.scratch = (double *)rs.p;
            le = launch_fast(a2, w2, g2, st);
        } else {
            if (wide_smem_bytes(N, L, NT) > smem_limit) return SBD_ERR_UNSUPPORTED;
            int g2 = 16;
            if (g2 > dims->nbins) g2 = dims->nbins;
            a2.slot_stride = wide_slot_doubles(N, L);
            a2.nslots = g2;
            if (rs.reserve(a2.slot_stride * (size_t)a2.nslots * 8) != cudaSuccess) return SBD_ERR_CUDA;
            a2.scratch = (double *)rs.p;
            le = launch_wide(a2, g2, st);
        }
    } else if (fast && NU > 0) {
        // radiance register kernel; bins whose TAUC is not monotone go to the general kernel
        // launched behind (normally the list stays empty and that launch ends at once)
        SbdDevBuf &rl = h->scratch_set ? h->redo2 : h->redo;
        SbdDevBuf &rs = h->scratch_set ? h->redo_scratch2 : h->redo_scratch;
        if (rl.reserve((size_t)dims->nbins * sizeof(int)) != cudaSuccess) return SBD_ERR_CUDA;
        a.redo_count = a.work_counter + 1;
        a.redo_list = (int *)rl.p;
        le = launch_fast(a, warps, grid, st);
        if (le != cudaSuccess) return SBD_ERR_CUDA;
        h->launches += 1;
        LaunchArgs a2 = a;
        a2.redo_consume = true;
        a2.work_counter = a.work_counter + 2;
        int w2 = generic_pick_warps(N, L, NT, smem_limit);
        if (w2 == 0) return SBD_ERR_UNSUPPORTED;
        if (w2 >= 8) w2 = 4;
        int g2 = 16;
        if (g2 > (dims->nbins + w2 - 1) / w2) g2 = (dims->nbins + w2 - 1) / w2;
        a2.slot_stride = generic_slot_doubles(N, L, NU);
        a2.nslots = g2 * w2;
        if (rs.reserve(a2.slot_stride * (size_t)a2.nslots * 8) != cudaSuccess) return SBD_ERR_CUDA;
        a2.scratch = (double *)rs.p;
        le = launch_generic(a2, w2, g2, st);
    } else {
        le = wide ? launch_wide(a, grid, st)
                  : (fast ? launch_fast(a, warps, grid, st) : launch_generic(a, warps, grid, st));
    }
    if (le != cudaSuccess) return SBD_ERR_CUDA;
    h->launches += 1;
    if (NU > 0 && h->corint) {         // INTCOR (disort.f:827-835): layer boundaries only
        if (dims->ntau > 0) return SBD_ERR_UNSUPPORTED;
        if (intcor_smem_bytes(L, dims->nmom) > smem_limit) return SBD_ERR_UNSUPPORTED;
        if (launch_intcor(a, st) != cudaSuccess) return SBD_ERR_CUDA;
        h->launches += 1;
    }
    return SBD_SUCCESS;
}

extern "C" int sbd_disort_batch(sbd_handle *h, const sbd_dims *dims, const double *dtauc,
                                const double *ssalb, const double *pmom, const sbd_bin *bins,
                                const double *temper, const double *utau, const double *umu,
                                const double *phi, double *rfldir, double *rfldn, double *flup,
                                double *dfdt, double *uavg, double *uu, int32_t *status)
{
    if (!h) return SBD_ERR_ARG;
    int rc = check_dims(dims);
    if (rc) return rc;
    if (!dtauc || !ssalb || !pmom || !bins || !status) return SBD_ERR_ARG;
    if (dims->numu > 0 && (!umu || !phi || !uu)) return SBD_ERR_ARG;
    if (dims->nbins == 0) return SBD_SUCCESS;
    if (cudaSetDevice(h->device) != cudaSuccess) return SBD_ERR_CUDA;
    cudaStream_t st = h->stream;
    const size_t B = dims->nbins, L = dims->nlyr, ldp = dims->nmom + 1;
    const size_t NT = dims->ntau > 0 ? dims->ntau : L + 1;
    bool any_plank = false;
    for (size_t b = 0; b < B; b++) {
        if (bins[b].plank) {
            any_plank = true;
            if (bins[b].col < 0 || bins[b].col >= dims->ncol) return SBD_ERR_ARG;
        }
    }
    if (any_plank && !temper) return SBD_ERR_ARG;

#define CK(x) do { if ((x) != cudaSuccess) return SBD_ERR_CUDA; } while (0)
    CK(h->d_dtauc.reserve(B * L * 8));
    CK(h->d_ssalb.reserve(B * L * 8));
    CK(h->d_pmom.reserve(B * L * ldp * 8));
    CK(h->d_bins.reserve(B * sizeof(sbd_bin)));
    CK(h->d_out.reserve(5 * B * NT * 8));
    CK(h->d_status.reserve(B * 4));
    const double *d_temper = nullptr;
    if (temper && dims->ncol > 0) {
        CK(h->d_temper.reserve((size_t)dims->ncol * (L + 1) * 8));
        CK(cudaMemcpyAsync(h->d_temper.p, temper, (size_t)dims->ncol * (L + 1) * 8,
                           cudaMemcpyHostToDevice, st));
        d_temper = (const double *)h->d_temper.p;
    }
    if (dims->ntau > 0) {
        if (!utau) return SBD_ERR_ARG;
        CK(h->d_utau.reserve(B * NT * 8));
    }
    const size_t per = B * NT;
    const size_t nuu1 = (size_t)dims->numu * dims->nphi * NT;      // uu doubles per bin
    // level selection (sbd_set_radiance_levels): the kernels write the selected levels only
    // (packed layout) and only those cross PCIe; with the full layout on the caller's side
    // they are scattered into uu on the host, the other levels of uu are left untouched
    std::vector<int32_t> sel;
    std::vector<double> uustage;
    size_t npack1 = 0;                                             // packed uu doubles per bin
    if (nuu1) {
        sel = selected_levels(h, (int)NT);
        if (sel.size() == NT) sel.clear();                         // every level: plain copy
    }
    if (nuu1 && sel.empty()) CK(h->d_uu.reserve(nuu1 * B * 8));
    const bool caller_packed = h->uu_packed && !sel.empty();
    if (!sel.empty()) {
        npack1 = (size_t)dims->nphi * sel.size() * dims->numu;
        CK(h->d_uupack.reserve(npack1 * B * 8));
        if (!caller_packed) uustage.resize(npack1 * B);
    }
    double *o = (double *)h->d_out.p;
    double *dst[5] = { rfldir, rfldn, flup, dfdt, uavg };
    // flux-level selection (sbd_set_flux_levels): the host arrays are [B][nsel]
    const int nfsel = (int)h->flux_levels.size();
    if (nfsel > 0) {
        for (int i = 0; i < nfsel; i++) if (h->flux_levels[i] >= (int)NT) return SBD_ERR_ARG;
        CK(h->d_sel.reserve((size_t)nfsel * 4));
        CK(cudaMemcpyAsync(h->d_sel.p, h->flux_levels.data(), (size_t)nfsel * 4, cudaMemcpyHostToDevice, st));
        CK(h->d_fluxpack.reserve(5 * B * nfsel * 8));
    }
    // Pipeline over chunks of bins: H2D of chunk c+1 and D2H of chunk c-1 overlap the
    // kernel of chunk c (copy-in, copy-out and two alternating compute streams, events
    // between them).  Only the first copy in and the last copy out are exposed, so the
    // chunks ramp up at the front (4096, 8192, 16384 bins) and down at the back
    // (8192, 4096); the middle is cut into equal parts of at most 32768 bins (27 at most).
    size_t cuts[sbd_handle::kMaxChunks + 1];
    int nchunk = 0;
    cuts[0] = 0;
    {
        const size_t front[3] = { 4096, 8192, 16384 }, back[2] = { 8192, 4096 };
        size_t lo = 0, hi = B;
        int nb_back = 0;
        if (B >= 4 * 4096) {
            for (int i = 0; i < 3 && hi - lo > 2 * front[i]; i++) { lo += front[i]; cuts[++nchunk] = lo; }
            for (int i = 1; i >= 0 && hi - lo > 2 * back[i]; i--) { hi -= back[i]; nb_back++; }
        }
        const size_t rest = hi - lo;
        int parts = (int)((rest + 32767) / 32768);
        if (parts < 1) parts = 1;
        if (parts > sbd_handle::kMaxChunks - 5) parts = sbd_handle::kMaxChunks - 5;
        for (int i = 1; i <= parts; i++) cuts[++nchunk] = lo + rest * i / parts;
        // back ramp: the larger chunk first
        if (nb_back == 2) { cuts[nchunk + 1] = hi + back[0]; nchunk++; }
        if (nb_back >= 1) cuts[++nchunk] = B;
    }
    CK(cudaStreamSynchronize(h->copy_out));
    // flux runs alternate between two compute streams (each with its own scratch set)
    const bool two = nchunk > 1 && dims->numu == 0;
    if (two) {       // the second stream sees the temperature upload
        CK(cudaEventRecord(h->ev_misc, st));
        CK(cudaStreamWaitEvent(h->stream2, h->ev_misc, 0));
    }
    for (int c = 0; c < nchunk; c++) {
        const size_t b0 = cuts[c], b1 = cuts[c + 1], nb = b1 - b0;
        cudaStream_t si = nchunk > 1 ? h->copy_in : st;
        cudaStream_t sk = (two && (c & 1)) ? h->stream2 : st;
        h->scratch_set = (two && (c & 1)) ? 1 : 0;
        CK(cudaMemcpyAsync((double *)h->d_dtauc.p + b0 * L, dtauc + b0 * L, nb * L * 8, cudaMemcpyHostToDevice, si));
        CK(cudaMemcpyAsync((double *)h->d_ssalb.p + b0 * L, ssalb + b0 * L, nb * L * 8, cudaMemcpyHostToDevice, si));
        CK(cudaMemcpyAsync((double *)h->d_pmom.p + b0 * L * ldp, pmom + b0 * L * ldp, nb * L * ldp * 8,
                           cudaMemcpyHostToDevice, si));
        CK(cudaMemcpyAsync((sbd_bin *)h->d_bins.p + b0, bins + b0, nb * sizeof(sbd_bin), cudaMemcpyHostToDevice, si));
        if (dims->ntau > 0)
            CK(cudaMemcpyAsync((double *)h->d_utau.p + b0 * NT, utau + b0 * NT, nb * NT * 8, cudaMemcpyHostToDevice, si));
        if (nchunk > 1) {
            CK(cudaEventRecord(h->ev_in[c], si));
            CK(cudaStreamWaitEvent(sk, h->ev_in[c], 0));
        }
        sbd_dims dc = *dims;
        dc.nbins = (int32_t)nb;
        h->uu_packed_once = !sel.empty();
        rc = sbd_disort_batch_device(
            h, &dc, (const double *)h->d_dtauc.p + b0 * L, (const double *)h->d_ssalb.p + b0 * L,
            (const double *)h->d_pmom.p + b0 * L * ldp, (const sbd_bin *)h->d_bins.p + b0, d_temper,
            dims->ntau > 0 ? (const double *)h->d_utau.p + b0 * NT : nullptr, umu, phi,
            o + b0 * NT, o + per + b0 * NT, o + 2 * per + b0 * NT, o + 3 * per + b0 * NT,
            o + 4 * per + b0 * NT,
            !nuu1 ? nullptr : (sel.empty() ? (double *)h->d_uu.p + b0 * nuu1 : (double *)h->d_uupack.p + b0 * npack1),
            (int32_t *)h->d_status.p + b0, sk);
        h->uu_packed_once = false;
        h->scratch_set = 0;
        if (rc) return rc;
        cudaStream_t so = nchunk > 1 ? h->copy_out : st;
        if (nchunk > 1) {
            CK(cudaEventRecord(h->ev_k[c], sk));
            CK(cudaStreamWaitEvent(so, h->ev_k[c], 0));
        }
        if (nfsel > 0) {
            double *pkc = (double *)h->d_fluxpack.p + 5 * b0 * nfsel;       // this chunk: [5][nb][nsel]
            CK(sbd_launch_pack_flux(o + b0 * NT, pkc, (const int32_t *)h->d_sel.p, nfsel, (int)NT, per, nb, 5, so));
            for (int k = 0; k < 5; k++)
                if (dst[k]) CK(cudaMemcpyAsync(dst[k] + b0 * nfsel, pkc + k * nb * nfsel, nb * nfsel * 8, cudaMemcpyDeviceToHost, so));
        } else {
            for (int k = 0; k < 5; k++)
                if (dst[k]) CK(cudaMemcpyAsync(dst[k] + b0 * NT, o + k * per + b0 * NT, nb * NT * 8, cudaMemcpyDeviceToHost, so));
        }
        if (nuu1 && sel.empty())
            CK(cudaMemcpyAsync(uu + b0 * nuu1, (double *)h->d_uu.p + b0 * nuu1, nb * nuu1 * 8, cudaMemcpyDeviceToHost, so));
        if (nuu1 && !sel.empty())
            CK(cudaMemcpyAsync((caller_packed ? uu : uustage.data()) + b0 * npack1,
                               (double *)h->d_uupack.p + b0 * npack1, nb * npack1 * 8, cudaMemcpyDeviceToHost, so));
        CK(cudaMemcpyAsync(status + b0, (int32_t *)h->d_status.p + b0, nb * 4, cudaMemcpyDeviceToHost, so));
    }
    CK(cudaStreamSynchronize(st));
    if (two) CK(cudaStreamSynchronize(h->stream2));
    if (nchunk > 1) CK(cudaStreamSynchronize(h->copy_out));
    if (!sel.empty() && !caller_packed) {   // scatter the selected levels; the others are not written
        const size_t NU = dims->numu, NP = dims->nphi, ns = sel.size();
        for (size_t b = 0; b < B; b++)
            for (size_t j = 0; j < NP; j++)
                for (size_t s = 0; s < ns; s++)
                    memcpy(uu + ((b * NP + j) * NT + sel[s]) * NU, uustage.data() + ((b * NP + j) * ns + s) * NU, NU * 8);
    }
#undef CK
    return SBD_SUCCESS;
}

// ---------------------------------------------------------------------------
// gfortran-compatible single-call entry (reference call site drt.f:541-546)
// ---------------------------------------------------------------------------
static thread_local int g_last_status = 0;

// BDREF of the host program (spectra.f:249; REAL(KR) FUNCTION with kr = 8, everything by
// reference): resolved from the executable when it exports the symbol (weak), or handed over
// with sbd_set_bdref_callback.
typedef double (*sbd_bdref_fn)(const double *wvnmlo, const double *wvnmhi, const double *mur,
                               const double *mui, const double *phir);
extern "C" double bdref_(const double *, const double *, const double *, const double *, const double *)
    __attribute__((weak));
static sbd_bdref_fn g_bdref = nullptr;
extern "C" void sbd_set_bdref_callback(sbd_bdref_fn fn) { g_bdref = fn; }

// SURFAC for LAMBER = .FALSE. (disort.f:3765-3907) on the host: Fourier coefficients of the
// bidirectional reflectivity by a 50-point azimuth quadrature, directional emissivities by a
// 25 x 50 quadrature.  Output layout: sbd_set_surfaces, one surface.
static void surfac_host(sbd_bdref_fn f, int N, int nmodes, int NU, const double *umu, double fbeam,
                        double umu0, double wlo, double whi, std::vector<double> &bdr,
                        std::vector<double> &bem, std::vector<double> &rmu, std::vector<double> &emu)
{
    constexpr int NMUG = 50;
    const int n = N / 2;
    const double pi = kPiRef;
    double gmu[NMUG], gwt[NMUG];
    gauss01(NMUG / 2, gmu, gwt);
    for (int k = 0; k < NMUG / 2; k++) { gmu[k + NMUG / 2] = -gmu[k]; gwt[k + NMUG / 2] = gwt[k]; }
    std::vector<double> cmu(n), cwt(n);
    gauss01(n, cmu.data(), cwt.data());
    bdr.assign((size_t)nmodes * n * (n + 1), 0.0); bem.assign(n, 0.0);
    rmu.assign((size_t)nmodes * NU * (n + 1), 0.0); emu.assign(NU, 0.0);
    auto bd = [&](double mur, double mui, double phir) { return f(&wlo, &whi, &mur, &mui, &phir); };
    // all modes of one (reflection, incidence) pair share the 50 BDREF values
    auto fourier = [&](double mur, double mui, double *out, size_t stride) {
        double v[NMUG];
        for (int k = 0; k < NMUG; k++) v[k] = bd(mur, mui, pi * gmu[k]);
        for (int m = 0; m < nmodes; m++) {
            double sum = 0.0;
            for (int k = 0; k < NMUG; k++) sum += gwt[k] * v[k] * cos(m * pi * gmu[k]);
            out[m * stride] = 0.5 * (2. - (m == 0 ? 1. : 0.)) * sum;
        }
    };
    auto emiss = [&](double mur) {
        double dref = 0.0;
        for (int jg = 0; jg < NMUG; jg++) {
            double sum = 0.0;
            for (int k = 0; k < NMUG / 2; k++) sum += gwt[k] * gmu[k] * bd(mur, gmu[k], pi * gmu[jg]);
            dref += gwt[jg] * sum;
        }
        return 1.0 - dref;
    };
    for (int iq = 0; iq < n; iq++) {
        for (int jq = 1; jq <= n; jq++) fourier(cmu[iq], cmu[jq - 1], &bdr[(size_t)iq * (n + 1) + jq], (size_t)n * (n + 1));
        if (fbeam > 0.0) fourier(cmu[iq], umu0, &bdr[(size_t)iq * (n + 1)], (size_t)n * (n + 1));
        bem[iq] = emiss(cmu[iq]);
    }
    for (int iu = 0; iu < NU; iu++) {
        if (!(umu[iu] > 0.0)) continue;
        for (int iq = 1; iq <= n; iq++) fourier(umu[iu], cmu[iq - 1], &rmu[(size_t)iu * (n + 1) + iq], (size_t)NU * (n + 1));
        if (fbeam > 0.0) fourier(umu[iu], umu0, &rmu[(size_t)iu * (n + 1)], (size_t)NU * (n + 1));
        emu[iu] = emiss(umu[iu]);
    }
}
static sbd_handle *g_handle = nullptr;
static std::mutex g_mutex;

extern "C" int sbd_disort_last_status(void) { return g_last_status; }

extern "C" void disort_(int *nlyr, double *dtauc, double *ssalb, int *corint, int *nmom,
                        double *pmom, double *temper, double *wvnmlo, double *wvnmhi, int *usrtau,
                        int *ntau, double *utau, int *nstr, int *usrang, int *numu, double *umu,
                        int *nphi, double *phi, int *ibcnd, double *fbeam, double *umu0,
                        double *phi0, double *fisot, int *lamber, double *albedo, double *btemp,
                        double *ttemp, double *temis, int *plank, int *onlyfl, double *accur,
                        int *prnt, char *header, int *maxcly, int *maxulv, int *maxumu,
                        int *maxphi, int *maxmom, double *rfldir, double *rfldn, double *flup,
                        double *dfdt, double *uavg, double *uu, double *albmed, double *trnmed,
                        size_t header_len)
{
    (void)prnt; (void)header; (void)header_len; (void)albmed; (void)trnmed; (void)maxcly;
    std::lock_guard<std::mutex> lock(g_mutex);
    if (!g_handle && sbd_create(&g_handle, 0) != SBD_SUCCESS) {
        fprintf(stderr, "sbdart_b200: disort_: no usable CUDA device\n");
        g_last_status = SBD_ERR_CUDA;
        return;
    }
    const int L = *nlyr, N = *nstr;
    const bool rad = !*onlyfl;
    // outside the hot path: IBCND=1, BRDF surfaces, intensities at the quadrature
    // angles (USRANG=F, never used by SBDART), CORINT with user levels
    const sbd_bdref_fn bdfn = g_bdref ? g_bdref : (sbd_bdref_fn)bdref_;
    if (*ibcnd != 0 || (rad && !*usrang) || (rad && *corint && *usrtau) || (!*lamber && (!bdfn || *corint || *usrtau))) {
        g_last_status = SBD_ERR_UNSUPPORTED;
        return;
    }
    if (rad && (*numu < 1 || *nphi < 1 || *numu > *maxumu || *nphi > *maxphi)) {
        g_last_status = SBD_ERR_ARG;
        return;
    }
    sbd_dims d;
    memset(&d, 0, sizeof d);
    d.nbins = 1; d.nlyr = L; d.nstr = N; d.nmom = *nmom; d.ncol = 1;
    d.ntau = *usrtau ? *ntau : 0;
    if (rad) { d.numu = *numu; d.nphi = *nphi; }
    const int ldp = *maxmom + 1;
    std::vector<double> pm((size_t)L * (*nmom + 1));
    for (int lc = 0; lc < L; lc++) {
        pmom[(size_t)lc * ldp] = 1.0;                         // disort.f:2555
        for (int k = 0; k <= *nmom; k++) pm[(size_t)lc * (*nmom + 1) + k] = pmom[(size_t)lc * ldp + k];
    }
    sbd_bin b;
    memset(&b, 0, sizeof b);
    b.fbeam = *fbeam; b.umu0 = *umu0; b.phi0 = *phi0; b.fisot = *fisot; b.albedo = *albedo;
    b.btemp = *btemp; b.ttemp = *ttemp; b.temis = *temis; b.wvnmlo = *wvnmlo; b.wvnmhi = *wvnmhi;
    b.plank = *plank ? 1 : 0; b.col = 0; b.accur = *accur;
    const int NT = *usrtau ? *ntau : L + 1;
    if (NT > *maxulv) { g_last_status = SBD_ERR_ARG; return; }
    int32_t st = 0;
    std::vector<double> uuc(rad ? (size_t)d.nphi * NT * d.numu : 0);
    if (!*lamber) {           // SURFAC with the host's BDREF (disort.f:3765-3907)
        std::vector<double> bdr, bem, rmu, emu;
        const int nmodes = rad ? N : 1;
        surfac_host(bdfn, N, nmodes, rad ? d.numu : 0, umu, *fbeam, *umu0, *wvnmlo, *wvnmhi, bdr, bem, rmu, emu);
        int rcs = sbd_set_surfaces(g_handle, 1, N, nmodes, rad ? d.numu : 0, bdr.data(), bem.data(),
                                   rad ? rmu.data() : nullptr, rad ? emu.data() : nullptr);
        if (rcs) { g_last_status = rcs; return; }
        b.albedo = SBD_SURFACE(0);
    }
    g_handle->corint = rad && *corint;
    int rc = sbd_disort_batch(g_handle, &d, dtauc, ssalb, pm.data(), &b, temper,
                              *usrtau ? utau : nullptr, rad ? umu : nullptr, rad ? phi : nullptr,
                              rfldir, rfldn, flup, dfdt, uavg, rad ? uuc.data() : nullptr, &st);
    g_handle->corint = false;
    if (!*lamber) sbd_set_surfaces(g_handle, 0, 0, 0, 0, nullptr, nullptr, nullptr, nullptr);
    g_last_status = rc ? rc : st;
    if (rc) return;
    if (rad)      // UU(IU,LU,J), leading dimensions MAXUMU, MAXULV (disort.f:377)
        for (int j = 0; j < d.nphi; j++)
            for (int lu = 0; lu < NT; lu++)
                for (int iu = 0; iu < d.numu; iu++)
                    uu[iu + (size_t)*maxumu * (lu + (size_t)*maxulv * j)] =
                        uuc[((size_t)j * NT + lu) * d.numu + iu];
    // visible argument mutations of the reference
    double tc = 0.0;
    if (!*usrtau) { *ntau = L + 1; utau[0] = 0.0; }
    for (int lc = 0; lc < L; lc++) {
        if (ssalb[lc] == 1.0) ssalb[lc] = 1.0 - kDither;       // disort.f:486
        tc += dtauc[lc];
        if (!*usrtau) utau[lc + 1] = tc;                       // disort.f:2534-2543
        if (dtauc[lc] < 0.0) dtauc[lc] = 0.0;                  // disort.f:4944
    }
    if (st == SBD_BIN_ANGLE_CLASH) { *nstr = -abs(N); return; } // disort.f:2648
    if (!*usrang || (*onlyfl && *maxumu >= N)) {               // disort.f:2655-2669
        std::vector<double> mu(N / 2), wt(N / 2);
        gauss01(N / 2, mu.data(), wt.data());
        *numu = N;
        for (int iu = 0; iu < N / 2; iu++) umu[iu] = -mu[N / 2 - 1 - iu];
        for (int iu = N / 2; iu < N; iu++) umu[iu] = mu[iu - N / 2];
    }
}
