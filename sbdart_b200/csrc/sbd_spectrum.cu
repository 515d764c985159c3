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

extern "C" int sbd_spectrum_run(sbd_handle *h, const sbd_optics_params *p, const double *z,
                                const double *pr, const double *t, const double *uu,
                                const sbd_cloud_entry *clouds, const double *wlalb, const double *alb,
                                const double *wlsun, const double *sun, int32_t numu, const double *umu,
                                int32_t nphi, const double *phi, int32_t *nk, double *wl, double *dwl,
                                double *wt, int32_t *nbins, double *rfldir, double *rfldn, double *flup,
                                double *uuout, int32_t *status, const sbd_inputs_out *inputs_out)
{
    if (!h || !p || !z || !pr || !t || !uu || !wlalb || !alb || !nk || !wl || !dwl || !wt || !nbins ||
        !status)
        return SBD_ERR_ARG;
    if (!h->opt_ready) return SBD_ERR_ARG;
    if (p->nz < 2 || p->nz > 65 || p->nwl < 1 || p->nstr < 4 || p->nstr > SBD_MAX_NSTR || (p->nstr & 1))
        return SBD_ERR_ARG;
    if (p->imomc != 2 && p->imomc != 3) return SBD_ERR_UNSUPPORTED;
    if (p->nf != 0 && (!wlsun || !sun || p->nsun < 2)) return SBD_ERR_ARG;
    if (p->ncloud > 0 && !clouds) return SBD_ERR_ARG;
    if (cudaSetDevice(h->device) != cudaSuccess) return SBD_ERR_CUDA;
    cudaStream_t st = h->stream;
    const int nz = p->nz, nwl = p->nwl, nmom = (p->nstr + 2 < 40) ? p->nstr + 2 : 40, ldp = nmom + 1;
    const size_t nslot = (size_t)3 * nwl;
#define CK(x) do { if ((x) != cudaSuccess) return SBD_ERR_CUDA; } while (0)
    // per-run setup arrays -> device (one packed upload)
    const size_t n_atm = 3 * (size_t)nz + 64 * (size_t)(nz + 1) + 2 * (size_t)p->nalb + 2 * (size_t)p->nsun;
    std::vector<double> atm(n_atm);
    {
        double *q = atm.data();
        memcpy(q, z, 8 * nz); q += nz;
        memcpy(q, pr, 8 * nz); q += nz;
        memcpy(q, t, 8 * nz); q += nz;
        memcpy(q, uu, 8 * 64 * (size_t)(nz + 1)); q += 64 * (size_t)(nz + 1);
        memcpy(q, wlalb, 8 * (size_t)p->nalb); q += p->nalb;
        memcpy(q, alb, 8 * (size_t)p->nalb); q += p->nalb;
        if (p->nsun > 0 && wlsun) { memcpy(q, wlsun, 8 * (size_t)p->nsun); q += p->nsun; memcpy(q, sun, 8 * (size_t)p->nsun); }
    }
    CK(h->opt_atm.reserve(n_atm * 8 + (size_t)p->ncloud * sizeof(sbd_cloud_entry) + 64));
    CK(cudaMemcpyAsync(h->opt_atm.p, atm.data(), n_atm * 8, cudaMemcpyHostToDevice, st));
    sbd_cloud_entry *d_clouds = (sbd_cloud_entry *)((double *)h->opt_atm.p + n_atm);
    if (p->ncloud > 0)
        CK(cudaMemcpyAsync(d_clouds, clouds, (size_t)p->ncloud * sizeof(sbd_cloud_entry), cudaMemcpyHostToDevice, st));

    CK(h->d_dtauc.reserve(nslot * nz * 8));
    CK(h->d_ssalb.reserve(nslot * nz * 8));
    CK(h->d_pmom.reserve(nslot * nz * ldp * 8));
    CK(h->d_bins.reserve(nslot * sizeof(sbd_bin)));
    // misc: nk[nwl] (int32), wl[nwl], dwl[nwl], wt[3 nwl], temper[nz+1]
    const size_t misc_d = (size_t)nwl / 2 + 1 + 2 * (size_t)nwl + nslot + (nz + 1);
    CK(h->opt_misc.reserve(misc_d * 8));
    double *m = (double *)h->opt_misc.p;
    int32_t *d_nk = (int32_t *)m;
    double *d_wl = m + nwl / 2 + 1, *d_dwl = d_wl + nwl, *d_wt = d_dwl + nwl, *d_temper = d_wt + nslot;

    OpticsArgs a;
    memset(&a, 0, sizeof a);
    a.p = *p;
    a.tab = h->opt_index;
    const double *d = (const double *)h->opt_atm.p;
    a.z = d; a.p_ = d + nz; a.t = d + 2 * nz; a.uu = d + 3 * nz;
    a.wlalb = a.uu + 64 * (size_t)(nz + 1); a.alb = a.wlalb + p->nalb;
    a.wlsun = a.alb + p->nalb; a.sun = a.wlsun + p->nsun;
    a.clouds = d_clouds;
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
    h->launches += 1;
    // temperature profile, top-down with the cap level duplicated (drt.f:330-333)
    std::vector<double> temper(nz + 1);
    temper[0] = t[nz - 1];
    for (int j = 1; j <= nz; j++) temper[j] = t[nz - j];
    CK(cudaMemcpyAsync(d_temper, temper.data(), (nz + 1) * 8, cudaMemcpyHostToDevice, st));
    // bin list in loop order (needs nk on the host anyway for the output accumulation)
    CK(cudaMemcpyAsync(nk, d_nk, (size_t)nwl * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(wl, d_wl, (size_t)nwl * 8, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(dwl, d_dwl, (size_t)nwl * 8, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(wt, d_wt, nslot * 8, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    std::vector<int32_t> map;
    map.reserve(nslot);
    for (int il = 0; il < nwl; il++) {
        if (nk[il] < 1 || nk[il] > 3) return SBD_ERR_CUDA;
        for (int kd = 0; kd < nk[il]; kd++) map.push_back(3 * il + kd);
    }
    const int B = (int)map.size();
    *nbins = B;
    CK(h->opt_map.reserve((size_t)B * 4));
    CK(cudaMemcpyAsync(h->opt_map.p, map.data(), (size_t)B * 4, cudaMemcpyHostToDevice, st));

    sbd_dims dims;
    memset(&dims, 0, sizeof dims);
    dims.nbins = B; dims.nlyr = nz; dims.nstr = p->nstr; dims.nmom = nmom; dims.ncol = 1;
    dims.numu = numu; dims.nphi = nphi;
    const size_t NT = nz + 1, per = (size_t)B * NT;
    const size_t nuu1 = (size_t)numu * nphi * NT;
    CK(h->d_out.reserve(5 * per * 8));
    CK(h->d_status.reserve((size_t)B * 4));
    if (nuu1) CK(h->d_uu.reserve(nuu1 * B * 8));
    double *o = (double *)h->d_out.p;
    h->pending_binmap = (const int32_t *)h->opt_map.p;
    int rc = sbd_disort_batch_device(h, &dims, a.dtauc, a.ssalb, a.pmom, a.bins, d_temper, nullptr, umu, phi,
                                     o, o + per, o + 2 * per, o + 3 * per, o + 4 * per,
                                     nuu1 ? (double *)h->d_uu.p : nullptr, (int32_t *)h->d_status.p, st);
    h->pending_binmap = nullptr;
    if (rc) return rc;
    if (rfldir) CK(cudaMemcpyAsync(rfldir, o, per * 8, cudaMemcpyDeviceToHost, st));
    if (rfldn) CK(cudaMemcpyAsync(rfldn, o + per, per * 8, cudaMemcpyDeviceToHost, st));
    if (flup) CK(cudaMemcpyAsync(flup, o + 2 * per, per * 8, cudaMemcpyDeviceToHost, st));
    if (nuu1 && uuout) CK(cudaMemcpyAsync(uuout, h->d_uu.p, nuu1 * B * 8, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(status, h->d_status.p, (size_t)B * 4, cudaMemcpyDeviceToHost, st));
    if (inputs_out) {
        if (inputs_out->dtauc) CK(cudaMemcpyAsync(inputs_out->dtauc, a.dtauc, nslot * nz * 8, cudaMemcpyDeviceToHost, st));
        if (inputs_out->ssalb) CK(cudaMemcpyAsync(inputs_out->ssalb, a.ssalb, nslot * nz * 8, cudaMemcpyDeviceToHost, st));
        if (inputs_out->pmom) CK(cudaMemcpyAsync(inputs_out->pmom, a.pmom, nslot * nz * ldp * 8, cudaMemcpyDeviceToHost, st));
        if (inputs_out->bins) CK(cudaMemcpyAsync(inputs_out->bins, a.bins, nslot * sizeof(sbd_bin), cudaMemcpyDeviceToHost, st));
    }
    CK(cudaStreamSynchronize(st));
#undef CK
    return SBD_SUCCESS;
}
