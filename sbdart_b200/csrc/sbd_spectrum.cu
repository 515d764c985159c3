// Whole-spectrum entry points of the C ABI: optical-property producer kernel
// (K2, sbd_optics.cu) followed by the solve kernel on the same stream.
#include <string.h>

#include <vector>

#include "sbd_handle.h"

using namespace sbd;

extern "C" int sbd_optics_upload_tables(sbd_handle *h, const double *tables, int64_t ntables_doubles,
                                        const int32_t *index, int32_t ntab)
{
    if (!h || !tables || !index || ntab != T_COUNT || ntables_doubles <= 0) return SBD_ERR_ARG;
    if (cudaSetDevice(h->device) != cudaSuccess) return SBD_ERR_CUDA;
    if (h->opt_tables.reserve((size_t)ntables_doubles * 8) != cudaSuccess) return SBD_ERR_CUDA;
    if (cudaMemcpyAsync(h->opt_tables.p, tables, (size_t)ntables_doubles * 8, cudaMemcpyHostToDevice,
                        h->stream) != cudaSuccess) return SBD_ERR_CUDA;
    h->opt_index.base = (const double *)h->opt_tables.p;
    for (int i = 0; i < T_COUNT; i++) {
        h->opt_index.off[i] = index[i];
        h->opt_index.len[i] = index[T_COUNT + i];
        if (index[i] < 0 || (int64_t)index[i] + index[T_COUNT + i] > ntables_doubles) return SBD_ERR_ARG;
    }
    if (cudaStreamSynchronize(h->stream) != cudaSuccess) return SBD_ERR_CUDA;
    h->opt_ready = true;
    return SBD_SUCCESS;
}

extern "C" int sbd_spectrum_set_aerosols(sbd_handle *h, const sbd_aerosol_params *p, const double *wlbaer,
                                         const double *aerext, const double *aerabs, const double *aerasm,
                                         const double *dtsv, const double *awl, const sbd_strat_entry *strat)
{
    if (!h) return SBD_ERR_ARG;
    if (!p) { h->aero_on = false; return SBD_SUCCESS; }
    const int n = p->nwlbaer, nz = p->nz, ns = p->nstrat;
    if (n < 0 || n == 1 || n > 150 || nz < 2 || nz > 65 || ns < 0 || ns > SBD_NAERZ) return SBD_ERR_ARG;
    if (n > 0 && (!wlbaer || !aerext || !aerabs || !aerasm || !dtsv)) return SBD_ERR_ARG;
    if (ns > 0 && (!awl || !strat)) return SBD_ERR_ARG;
    if (n > 0 && p->imoma != 2 && p->imoma != 3) return SBD_ERR_UNSUPPORTED;
    for (int i = 0; i < ns; i++)
        if (!(strat[i].layer >= 1 && strat[i].layer <= nz)) return SBD_ERR_ARG;
    if (cudaSetDevice(h->device) != cudaSuccess) return SBD_ERR_CUDA;
    // packed: [wlb n][ext n][abs n][asm n][dtsv nz][awl 47][strat ns]
    const size_t per = sizeof(sbd_strat_entry) / 8;
    std::vector<double> pk(4 * (size_t)n + nz + SBD_NAERW + per * ns, 0.0);
    double *q = pk.data();
    if (n > 0) {
        memcpy(q, wlbaer, 8 * (size_t)n); q += n;
        memcpy(q, aerext, 8 * (size_t)n); q += n;
        memcpy(q, aerabs, 8 * (size_t)n); q += n;
        memcpy(q, aerasm, 8 * (size_t)n); q += n;
        memcpy(q, dtsv, 8 * (size_t)nz);
    }
    q = pk.data() + 4 * (size_t)n + nz;
    if (ns > 0) {
        memcpy(q, awl, 8 * SBD_NAERW); q += SBD_NAERW;
        memcpy(q, strat, sizeof(sbd_strat_entry) * ns);
    }
    if (h->opt_aero.reserve(pk.size() * 8) != cudaSuccess) return SBD_ERR_CUDA;
    if (cudaMemcpyAsync(h->opt_aero.p, pk.data(), pk.size() * 8, cudaMemcpyHostToDevice, h->stream) != cudaSuccess)
        return SBD_ERR_CUDA;
    if (cudaStreamSynchronize(h->stream) != cudaSuccess) return SBD_ERR_CUDA;   // pk is a temporary
    h->aero = *p;
    h->aero_on = true;
    return SBD_SUCCESS;
}

// out[k][b][s] = in[k][b][sel[s]], k = rfldir, rfldn, flup, dfdt, uavg: only the levels chosen with
// sbd_set_flux_levels cross PCIe (SBDART's iout 1 / 10 records read two of the nz+1 levels)
__global__ void pack_flux_levels_kernel(const double *in, double *out, const int32_t *sel, int nsel, int NT,
                                        size_t per_in /* B*NT */, size_t nbins, int narr)
{
    const size_t total = (size_t)narr * nbins * nsel;
    for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
        const int s = (int)(e % nsel);
        const size_t t = e / nsel;
        const size_t b = t % nbins, k = t / nbins;
        out[e] = in[k * per_in + b * NT + sel[s]];
    }
}

cudaError_t sbd_launch_pack_flux(const double *in, double *out, const int32_t *sel, int nsel, int NT,
                                 size_t per_in, size_t nbins, int narr, cudaStream_t st)
{
    const size_t total = (size_t)narr * nbins * nsel;
    if (total == 0) return cudaSuccess;
    const int blocks = (int)((total + 255) / 256 < 2368 ? (total + 255) / 256 : 2368);
    pack_flux_levels_kernel<<<blocks, 256, 0, st>>>(in, out, sel, nsel, NT, per_in, nbins, narr);
    return cudaGetLastError();
}

extern "C" int sbd_set_flux_levels(sbd_handle *h, const int32_t *levels, int32_t n)
{
    if (!h || n < 0 || (n > 0 && !levels)) return SBD_ERR_ARG;
    std::vector<int32_t> v(levels, levels + n);
    for (int i = 0; i < n; i++)
        if (v[i] < 0 || v[i] > SBD_MAX_NLYR || (i > 0 && v[i] <= v[i - 1])) return SBD_ERR_ARG;
    h->flux_levels = v;
    return SBD_SUCCESS;
}

// K2 for ncol columns x nwl wavelengths, bin map, solve, (packed) copies back.  One stream, one
// synchronisation at the end: the number of bins stays on the device (the solve kernel reads
// it there), the host learns it from nk[].
static int spectrum_core(sbd_handle *h, const sbd_optics_params *p, int ncol, const double *z, const double *pr,
                         const double *t, const double *uu, const sbd_cloud_entry *clouds, const double *wlalb,
                         const double *alb, const double *wlsun, const double *sun, int32_t numu,
                         const double *umu, int32_t nphi, const double *phi, int32_t *nk, double *wl, double *dwl,
                         double *wt, int32_t *nbins, double *rfldir, double *rfldn, double *flup, double *uuout,
                         int32_t *status, const sbd_inputs_out *inputs_out)
{
    if (!h || !p || !z || !pr || !t || !uu || !wlalb || !alb || !nk || !wl || !dwl || !wt || !nbins ||
        !status)
        return SBD_ERR_ARG;
    if (!h->opt_ready) return SBD_ERR_ARG;
    if (ncol < 1 || p->nz < 2 || p->nz > 65 || p->nwl < 1 || p->nstr < 4 || p->nstr > SBD_MAX_NSTR || (p->nstr & 1))
        return SBD_ERR_ARG;
    if ((size_t)ncol * p->nwl * 3 > (size_t)1 << 30) return SBD_ERR_ARG;
    if (p->imomc != 2 && p->imomc != 3) return SBD_ERR_UNSUPPORTED;
    if (p->nf != 0 && (!wlsun || !sun || p->nsun < 2)) return SBD_ERR_ARG;
    if (p->ncloud > 0 && !clouds) return SBD_ERR_ARG;
    if (cudaSetDevice(h->device) != cudaSuccess) return SBD_ERR_CUDA;
    cudaStream_t st = h->stream;
    const int nz = p->nz, nwl = p->nwl, nmom = (p->nstr + 2 < 40) ? p->nstr + 2 : 40, ldp = nmom + 1;
    const size_t nitem = (size_t)ncol * nwl, nslot = 3 * nitem;
    const size_t NT = nz + 1;
#define CK(x) do { if ((x) != cudaSuccess) return SBD_ERR_CUDA; } while (0)
    // per-run setup arrays -> device (one packed upload from a pinned staging buffer):
    // z, p, t [ncol][nz]; uu [ncol][64][nz+1]; temper [ncol][nz+1]; btemp, ttemp [ncol]; tables
    const size_t n_col = 3 * (size_t)nz + 64 * (size_t)(nz + 1) + (nz + 1) + 2;
    const size_t n_atm = n_col * ncol + 2 * (size_t)p->nalb + 2 * (size_t)p->nsun;
    const size_t up_bytes = n_atm * 8 + (size_t)p->ncloud * sizeof(sbd_cloud_entry);
    if (h->host_stage_cap < up_bytes) {
        if (h->host_stage) cudaFreeHost(h->host_stage);
        h->host_stage = nullptr; h->host_stage_cap = 0;
        CK(cudaMallocHost(&h->host_stage, up_bytes + up_bytes / 4 + 256));
        h->host_stage_cap = up_bytes + up_bytes / 4 + 256;
    }
    double *q = (double *)h->host_stage;
    double *hz = q, *hp = hz + (size_t)ncol * nz, *ht = hp + (size_t)ncol * nz, *huu = ht + (size_t)ncol * nz;
    double *htemper = huu + (size_t)ncol * 64 * (nz + 1), *hbt = htemper + (size_t)ncol * (nz + 1), *htt = hbt + ncol;
    double *htab = htt + ncol;
    memcpy(hz, z, 8 * (size_t)ncol * nz);
    memcpy(hp, pr, 8 * (size_t)ncol * nz);
    memcpy(ht, t, 8 * (size_t)ncol * nz);
    memcpy(huu, uu, 8 * (size_t)ncol * 64 * (nz + 1));
    for (int c = 0; c < ncol; c++) {
        // temperature profile, top-down with the cap level duplicated (drt.f:330-333)
        const double *tc = t + (size_t)c * nz;
        double *tp = htemper + (size_t)c * (nz + 1);
        tp[0] = tc[nz - 1];
        for (int j = 1; j <= nz; j++) tp[j] = tc[nz - j];
        hbt[c] = p->btemp >= 0. ? p->btemp : tp[nz];          // drt.f:334-335
        htt[c] = p->ttemp >= 0. ? p->ttemp : tp[0];
    }
    {
        double *w = htab;
        memcpy(w, wlalb, 8 * (size_t)p->nalb); w += p->nalb;
        memcpy(w, alb, 8 * (size_t)p->nalb); w += p->nalb;
        if (p->nsun > 0 && wlsun) { memcpy(w, wlsun, 8 * (size_t)p->nsun); w += p->nsun; memcpy(w, sun, 8 * (size_t)p->nsun); }
        if (p->ncloud > 0) memcpy(htab + 2 * (size_t)p->nalb + 2 * (size_t)p->nsun, clouds, (size_t)p->ncloud * sizeof(sbd_cloud_entry));
    }
    CK(h->opt_atm.reserve(up_bytes + 64));
    CK(cudaMemcpyAsync(h->opt_atm.p, h->host_stage, up_bytes, cudaMemcpyHostToDevice, st));
    h->last_h2d_bytes = up_bytes;

    CK(h->d_dtauc.reserve(nslot * nz * 8));
    CK(h->d_ssalb.reserve(nslot * nz * 8));
    CK(h->d_pmom.reserve(nslot * nz * ldp * 8));
    CK(h->d_bins.reserve(nslot * sizeof(sbd_bin)));
    // misc: nbins (int32, padded), nk[nitem] (int32), wl[nwl], dwl[nwl], wt[nslot]
    const size_t misc_d = 2 + nitem / 2 + 1 + 2 * (size_t)nwl + nslot;
    CK(h->opt_misc.reserve(misc_d * 8));
    double *m = (double *)h->opt_misc.p;
    int32_t *d_nbins = (int32_t *)m;
    int32_t *d_nk = (int32_t *)(m + 2);
    double *d_wl = m + 2 + nitem / 2 + 1, *d_dwl = d_wl + nwl, *d_wt = d_dwl + nwl;
    CK(h->opt_map.reserve(nslot * 4));

    OpticsArgs a;
    memset(&a, 0, sizeof a);
    a.p = *p;
    a.tab = h->opt_index;
    a.ncol = ncol;
    const double *d = (const double *)h->opt_atm.p;
    a.z = d; a.p_ = d + (size_t)ncol * nz; a.t = d + 2 * (size_t)ncol * nz; a.uu = d + 3 * (size_t)ncol * nz;
    const double *d_temper = a.uu + (size_t)ncol * 64 * (nz + 1);
    a.btemp = d_temper + (size_t)ncol * (nz + 1); a.ttemp = a.btemp + ncol;
    a.wlalb = a.ttemp + ncol; a.alb = a.wlalb + p->nalb;
    a.wlsun = a.alb + p->nalb; a.sun = a.wlsun + p->nsun;
    a.clouds = (const sbd_cloud_entry *)(a.sun + p->nsun);
    if (h->aero_on) {
        if (h->aero.nz != nz) return SBD_ERR_ARG;
        a.aer = h->aero;
        a.aero = (const double *)h->opt_aero.p;
    }
    a.dtauc = (double *)h->d_dtauc.p; a.ssalb = (double *)h->d_ssalb.p; a.pmom = (double *)h->d_pmom.p;
    a.bins = (sbd_bin *)h->d_bins.p;
    a.nk = d_nk; a.wl = d_wl; a.dwl = d_dwl; a.wt = d_wt;
    // unused k slots must not hold garbage weights
    CK(cudaMemsetAsync(d_wt, 0, nslot * 8, st));
    if (launch_optics(a, st) != cudaSuccess) return SBD_ERR_CUDA;
    // bin list in loop order, on the device (the host rebuilds it from nk for the accumulation)
    if (launch_binmap(d_nk, (int)nitem, (int32_t *)h->opt_map.p, d_nbins, st) != cudaSuccess) return SBD_ERR_CUDA;
    h->launches += 2;

    sbd_dims dims;
    memset(&dims, 0, sizeof dims);
    dims.nbins = (int32_t)nslot;            // upper bound; the kernels read the count from d_nbins
    dims.nlyr = nz; dims.nstr = p->nstr; dims.nmom = nmom; dims.ncol = ncol;
    dims.numu = numu; dims.nphi = nphi;
    const size_t per = nslot * NT;
    const size_t nuu1 = (size_t)numu * nphi * NT;
    CK(h->d_out.reserve(5 * per * 8));
    CK(h->d_status.reserve(nslot * 4));
    if (nuu1) CK(h->d_uu.reserve(nuu1 * nslot * 8));
    double *o = (double *)h->d_out.p;
    h->pending_binmap = (const int32_t *)h->opt_map.p;
    h->pending_nbins_dev = d_nbins;
    int rc = sbd_disort_batch_device(h, &dims, a.dtauc, a.ssalb, a.pmom, a.bins, d_temper, nullptr, umu, phi,
                                     o, o + per, o + 2 * per, o + 3 * per, o + 4 * per,
                                     nuu1 ? (double *)h->d_uu.p : nullptr, (int32_t *)h->d_status.p, st);
    h->pending_binmap = nullptr;
    h->pending_nbins_dev = nullptr;
    if (rc) return rc;
    // results: the whole slot range is copied (bins beyond the count are not meaningful);
    // with a flux-level selection only the chosen levels, packed [bin][nsel]
    size_t d2h = 0;
    const int nsel = (int)h->flux_levels.size();
    if (nsel > 0) {
        for (int i = 0; i < nsel; i++) if (h->flux_levels[i] >= (int)NT) return SBD_ERR_ARG;
        CK(h->d_sel.reserve((size_t)nsel * 4));
        CK(cudaMemcpyAsync(h->d_sel.p, h->flux_levels.data(), (size_t)nsel * 4, cudaMemcpyHostToDevice, st));
        CK(h->d_fluxpack.reserve(3 * nslot * nsel * 8));
        CK(sbd_launch_pack_flux(o, (double *)h->d_fluxpack.p, (const int32_t *)h->d_sel.p, nsel, (int)NT, per, nslot, 3, st));
        h->launches += 1;
        const size_t one = nslot * nsel * 8;
        const double *pk = (const double *)h->d_fluxpack.p;
        h->last_flux_dev = h->d_fluxpack.p; h->last_flux_doubles = 3 * nslot * nsel;
        if (rfldir) { CK(cudaMemcpyAsync(rfldir, pk, one, cudaMemcpyDeviceToHost, st)); d2h += one; }
        if (rfldn) { CK(cudaMemcpyAsync(rfldn, pk + nslot * nsel, one, cudaMemcpyDeviceToHost, st)); d2h += one; }
        if (flup) { CK(cudaMemcpyAsync(flup, pk + 2 * nslot * nsel, one, cudaMemcpyDeviceToHost, st)); d2h += one; }
    } else {
        h->last_flux_dev = o; h->last_flux_doubles = 3 * per;
        if (rfldir) { CK(cudaMemcpyAsync(rfldir, o, per * 8, cudaMemcpyDeviceToHost, st)); d2h += per * 8; }
        if (rfldn) { CK(cudaMemcpyAsync(rfldn, o + per, per * 8, cudaMemcpyDeviceToHost, st)); d2h += per * 8; }
        if (flup) { CK(cudaMemcpyAsync(flup, o + 2 * per, per * 8, cudaMemcpyDeviceToHost, st)); d2h += per * 8; }
    }
    if (nuu1 && uuout) { CK(cudaMemcpyAsync(uuout, h->d_uu.p, nuu1 * nslot * 8, cudaMemcpyDeviceToHost, st)); d2h += nuu1 * nslot * 8; }
    CK(cudaMemcpyAsync(status, h->d_status.p, nslot * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(nk, d_nk, nitem * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(wl, d_wl, (size_t)nwl * 8, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(dwl, d_dwl, (size_t)nwl * 8, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(wt, d_wt, nslot * 8, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(nbins, d_nbins, 4, cudaMemcpyDeviceToHost, st));
    d2h += nslot * 4 + nitem * 4 + 2 * (size_t)nwl * 8 + nslot * 8 + 4;
    if (inputs_out) {
        if (inputs_out->dtauc) CK(cudaMemcpyAsync(inputs_out->dtauc, a.dtauc, nslot * nz * 8, cudaMemcpyDeviceToHost, st));
        if (inputs_out->ssalb) CK(cudaMemcpyAsync(inputs_out->ssalb, a.ssalb, nslot * nz * 8, cudaMemcpyDeviceToHost, st));
        if (inputs_out->pmom) CK(cudaMemcpyAsync(inputs_out->pmom, a.pmom, nslot * nz * ldp * 8, cudaMemcpyDeviceToHost, st));
        if (inputs_out->bins) CK(cudaMemcpyAsync(inputs_out->bins, a.bins, nslot * sizeof(sbd_bin), cudaMemcpyDeviceToHost, st));
    }
    CK(cudaStreamSynchronize(st));
    h->last_d2h_bytes = d2h;
    for (size_t i = 0; i < nitem; i++)
        if (nk[i] < 1 || nk[i] > 3) return SBD_ERR_CUDA;
#undef CK
    return SBD_SUCCESS;
}

extern "C" int sbd_spectrum_run(sbd_handle *h, const sbd_optics_params *p, const double *z,
                                const double *pr, const double *t, const double *uu,
                                const sbd_cloud_entry *clouds, const double *wlalb, const double *alb,
                                const double *wlsun, const double *sun, int32_t numu, const double *umu,
                                int32_t nphi, const double *phi, int32_t *nk, double *wl, double *dwl,
                                double *wt, int32_t *nbins, double *rfldir, double *rfldn, double *flup,
                                double *uuout, int32_t *status, const sbd_inputs_out *inputs_out)
{
    return spectrum_core(h, p, 1, z, pr, t, uu, clouds, wlalb, alb, wlsun, sun, numu, umu, nphi, phi, nk, wl, dwl,
                         wt, nbins, rfldir, rfldn, flup, uuout, status, inputs_out);
}

extern "C" int sbd_spectrum_run_columns(sbd_handle *h, const sbd_optics_params *p, int32_t ncol, const double *z,
                                        const double *pr, const double *t, const double *uu,
                                        const sbd_cloud_entry *clouds, const double *wlalb, const double *alb,
                                        const double *wlsun, const double *sun, int32_t *nk, double *wl,
                                        double *dwl, double *wt, int32_t *nbins, double *rfldir, double *rfldn,
                                        double *flup, int32_t *status)
{
    return spectrum_core(h, p, ncol, z, pr, t, uu, clouds, wlalb, alb, wlsun, sun, 0, nullptr, 0, nullptr, nk, wl,
                         dwl, wt, nbins, rfldir, rfldn, flup, nullptr, status, nullptr);
}

extern "C" int sbd_spectrum_device_fluxes(sbd_handle *h, void **ptr, int64_t *ndoubles)
{
    if (!h || !ptr || !ndoubles) return SBD_ERR_ARG;
    *ptr = h->last_flux_dev;
    *ndoubles = (int64_t)h->last_flux_doubles;
    return SBD_SUCCESS;
}

extern "C" int sbd_last_transfer_bytes(const sbd_handle *h, int64_t *h2d, int64_t *d2h)
{
    if (!h) return SBD_ERR_ARG;
    if (h2d) *h2d = (int64_t)h->last_h2d_bytes;
    if (d2h) *d2h = (int64_t)h->last_d2h_bytes;
    return SBD_SUCCESS;
}
