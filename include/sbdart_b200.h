/*
 * sbdart_b200.h -- C ABI of the B200 batched discrete-ordinate solver.
 *
 * Drop-in boundary for SBDART's hot path: the body of the wavelength loop
 * (reference drt.f:425-561) calls SUBROUTINE DISORT once per
 * (wavelength, k-distribution term) "bin" (drt.f:541-546, disort.f:1-6).
 * This library replaces that call with
 *   - disort_()            : the gfortran-mangled single-call entry with the
 *                            reference's own 46-argument list, and
 *   - sbd_disort_batch*()  : all bins of a run flattened into one launch.
 * No torch / C++ types cross this boundary: plain pointers and sizes only.
 * The library has no CPU fallback: every entry fails with SBD_ERR_CUDA when
 * no CUDA device is usable.
 */
#ifndef SBDART_B200_H
#define SBDART_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SBD_ABI_VERSION 1

/* hard limits mirrored from the reference (params.f:8-15) */
#define SBD_MAX_NSTR 40
#define SBD_MAX_NLYR 128

/* return codes of the API calls */
enum {
    SBD_SUCCESS = 0,
    SBD_ERR_CUDA = -100,      /* no device / CUDA runtime failure            */
    SBD_ERR_ARG = -101,       /* bad dims or NULL pointer                    */
    SBD_ERR_UNSUPPORTED = -102 /* IBCND=1, USRTAU w/ radiances                 */
};

/* per-bin status written to status[B]; numbering follows the reference's
 * errmsg numbers where one exists (disutil.f:278-325)     */
enum {
    SBD_BIN_OK = 0,
    SBD_BIN_ANGLE_CLASH = 1,  /* disort.f:2645-2650 -> NSTR returned negative;
                                 the host retries with NSTR-2 / NSTR+2 as
                                 drt.f:536-554 does                           */
    SBD_BIN_BAD_INPUT = -1,   /* CHEKIN fatal (disort.f:5155)                 */
    SBD_BIN_EIG_FAIL = -2,    /* eigen-solve failed (disort.f:3254-3261)      */
    SBD_BIN_SINGULAR = -3     /* zero pivot in the boundary system            */
};

/* Shapes of one batched call.  All bins of a call share these. */
typedef struct sbd_dims {
    int32_t nbins;  /* B : number of (wavelength, k-term[, column]) bins     */
    int32_t nlyr;   /* L : NLYR  (SBDART passes nz)     */
    int32_t nstr;   /* N : NSTR, even, 4..SBD_MAX_NSTR                        */
    int32_t nmom;   /* pmom holds moments 0..nmom per layer, nmom >= nstr
                       (SBDART passes nstr+2, drt.f:493)                      */
    int32_t ntau;   /* 0 => USRTAU=.FALSE.: the L+1 layer boundaries
                       (the only mode SBDART uses, drt.f:175-178);
                       >0 => USRTAU=.TRUE. with utau[B][ntau] (flux only)     */
    int32_t numu;   /* 0 => ONLYFL=.TRUE. (fluxes only); >0 => USRANG user
                       polar angles umu[numu] shared by all bins              */
    int32_t nphi;   /* number of azimuths phi[nphi] (numu>0 only)             */
    int32_t ncol;   /* number of temperature profiles temper[ncol][L+1]       */
} sbd_dims;

/* Per-bin scalar arguments of DISORT (disort.f:1-6). */
typedef struct sbd_bin {
    double fbeam;   /* FBEAM  */
    double umu0;    /* UMU0   */
    double phi0;    /* PHI0   (degrees)                                       */
    double fisot;   /* FISOT  */
    double albedo;  /* ALBEDO (Lambertian); SBD_SURFACE(s) selects the BRDF
                       surface s of sbd_set_surfaces (LAMBER = .FALSE.)       */
    double btemp;   /* BTEMP  */
    double ttemp;   /* TTEMP  */
    double temis;   /* TEMIS  */
    double wvnmlo;  /* WVNMLO */
    double wvnmhi;  /* WVNMHI */
    double accur;   /* ACCUR  azimuth-series convergence (SBDART passes 0, drt.f:142) */
    int32_t plank;  /* PLANK  */
    int32_t col;    /* row of temper[][] used by this bin                     */
} sbd_bin;

typedef struct sbd_handle sbd_handle; /* owns a CUDA stream + device scratch */

/* Create / destroy a solver bound to CUDA device `device`. */
int sbd_create(sbd_handle **out, int device);
void sbd_destroy(sbd_handle *h);

/*
 * Batched solve, HOST buffers (the reference-facing call: copies inputs to
 * the device, launches, copies results back; synchronous on return).
 *
 *  dtauc [B][L], ssalb [B][L]     DTAUC, SSALB       (top layer first)
 *  pmom  [B][L][nmom+1]           PMOM(0:nmom, lc)   (moment index fastest)
 *  bins  [B]                      scalars
 *  temper[ncol][L+1]              TEMPER(0:NLYR); may be NULL if no bin has
 *                                 plank set
 *  utau  [B][ntau]                only when dims.ntau > 0, else NULL
 *  umu   [numu], phi [nphi]       only when dims.numu > 0, else NULL
 * Outputs, NT = ntau ? ntau : L+1; any flux pointer may be NULL:
 *  rfldir, rfldn, flup, dfdt, uavg : [B][NT]
 *  uu   [B][nphi][NT][numu]       (numu>0), Fortran UU(iu,lu,j) transposed
 *  status [B]
 * The inputs are const: the in-place edits DISORT makes (SSALB=1 -> 1-DITHER
 * disort.f:486, DTAUC<0 -> 0 disort.f:4944) are applied on the device copy.
 */
int sbd_disort_batch(sbd_handle *h, const sbd_dims *dims, const double *dtauc,
                     const double *ssalb, const double *pmom,
                     const sbd_bin *bins, const double *temper,
                     const double *utau, const double *umu, const double *phi,
                     double *rfldir, double *rfldn, double *flup, double *dfdt,
                     double *uavg, double *uu, int32_t *status);

/* Same, all pointers are DEVICE pointers; enqueued on the handle's stream
 * (or on `stream` if non-NULL, a cudaStream_t passed as void*), returns
 * without synchronising. */
int sbd_disort_batch_device(sbd_handle *h, const sbd_dims *dims,
                            const double *dtauc, const double *ssalb,
                            const double *pmom, const sbd_bin *bins,
                            const double *temper, const double *utau,
                            const double *umu, const double *phi,
                            double *rfldir, double *rfldn, double *flup,
                            double *dfdt, double *uavg, double *uu,
                            int32_t *status, void *stream);

/* Block until everything enqueued on the handle's stream has finished. */
int sbd_synchronize(sbd_handle *h);

/* Levels at which intensities (uu) are wanted by the following batched calls on
 * this handle: indices into the NT output levels, n = 0 restores "all levels".
 * SBDART consumes the radiances of one or two levels only (ntop / nbot,
 * drt.f:1008-1016, :1143-1151) while DISORT's USRINT (disort.f:4355) integrates
 * the source function for every level.  The device-pointer call writes zeros at
 * the unselected levels of uu; the host-buffer call copies back the selected
 * levels only and leaves the rest of the caller's uu untouched (the copy of all
 * L+1 levels would dominate the call).  Fluxes are not affected.  With ACCUR > 0
 * the convergence test of the azimuth series (disort.f:802-823) sees the selected
 * levels only (SBDART runs ACCUR = 0: every mode is summed). */
int sbd_set_radiance_levels(sbd_handle *h, const int32_t *levels, int32_t n);

/* Layout of uu in the following batched calls while a level selection is active:
 * packed = 0 (default) the full layout uu[B][nphi][NT][numu] described above;
 * packed = 1 the selected levels only, ascending: uu[B][nphi][nsel][numu] -- host-buffer and
 * device-pointer calls alike (no zero fill, no scatter: what a caller that reads one or two
 * levels per bin wants).  Without a selection uu is always the full layout. */
int sbd_set_radiance_layout(sbd_handle *h, int32_t packed);

/* Output levels of the flux arrays in the following HOST-buffer calls (sbd_disort_batch,
 * sbd_spectrum_run*): n > 0 strictly ascending indices into the NT output levels -- rfldir,
 * rfldn, flup, dfdt, uavg are then [B][n] and only those levels cross PCIe (SBDART's records
 * read ntop and nbot, drt.f:977-982); n = 0 restores the full [B][NT] layout.  Device-pointer
 * calls are not affected. */
int sbd_set_flux_levels(sbd_handle *h, const int32_t *levels, int32_t n);

/* CORINT of DISORT (disort.f:112-118, INTCOR disort.f:2044-2297) for the following
 * batched radiance calls on this handle: 0 (default) off, 1 on.  When on, the
 * Nakajima-Tanaka TMS and IMS corrections are added to uu after the solve; pmom
 * should then carry the full phase function (SBDART passes nmom = 299,
 * drt.f:490-491).  As in the reference it has no effect on flux-only calls, for
 * fbeam = 0 and for non-scattering media (disort.f:2695-2696). */
int sbd_set_corint(sbd_handle *h, int32_t on);

/* cudaStream_t of the handle (as void*), for event timing by the caller. */
void *sbd_stream(sbd_handle *h);

/* Number of kernels of this library launched through `h` so far. */
int64_t sbd_kernel_launches(const sbd_handle *h);

/* Gauss-Legendre nodes/weights on (0,1) used for NSTR = 2*m streams
 * (same rule as QGAUSN, disort.f:5984). Host-only helper. */
int sbd_quadrature(int m, double *mu, double *wt);

/* ------------------------------------------------------------------------
 * Whole-spectrum path: the optical-property producers of the wavelength loop
 * (gasset/taugas/kdistr/taucor taugas.f:7392,2236,1802,7650; taucloud/cloudpar
 * taucloud.f:10,344; rayleigh spectra.f:206; normom drt.f:1366; depthscl
 * taugas.f:7512; solirr/salbedo/wllimits) run in a GPU kernel that writes the
 * DISORT inputs of every (wavelength, k-term) bin straight into HBM; the solve
 * kernel follows on the same stream.  Only the per-run setup (atmosphere,
 * absorber amounts uu(63,nz) of absint, cloud layer list, tables) comes from
 * the host.  Covers Lambertian surfaces, flat filters (isat 0/-2) and, through
 * sbd_spectrum_set_aerosols, the standard and user aerosol models.
 * ------------------------------------------------------------------------ */
typedef struct sbd_optics_params {
    int32_t nz;        /* levels = DISORT layers            */
    int32_t nwl;       /* wavelengths (setfilt, spectra.f:3370-3384)               */
    int32_t kdist;     /* 0..3 (taugas.f:7441, :7562-7590)                         */
    int32_t nstr;      /* streams; nmom = min(nstr+2, 40) moments are produced      */
    int32_t nf;        /* 0: unit solar flux, else the table wlsun/sun is used      */
    int32_t nothrm;    /* <0: Planck for wl > 2 um; 0: always; >0: never (drt.f:463)*/
    int32_t imomc;     /* cloud phase function model: 3 = Henyey-Greenstein         */
    int32_t ncloud;    /* entries of the cloud layer list                           */
    int32_t nalb, nsun;/* table lengths                                             */
    int32_t night;     /* sza >= 90: flxin = 0, amu0 = 1 (drt.f:456-459)            */
    int32_t pad_;
    double wl1, wl2, wlinc;      /* wavelength grid (wllimits, drt.f:1657)          */
    double amu0, xo4, xrsc, solfac, phi0, fisot, temis, btemp, ttemp;
} sbd_optics_params;

/* One (cloud layer, DISORT layer) pair of taucloud's double loop
 * (taucloud.f:61-124), resolved on the host because it does not depend on
 * wavelength. */
typedef struct sbd_cloud_entry {
    int32_t layer;     /* DISORT layer j, 1-based from the top                      */
    int32_t use_tau;   /* tcloud given: tau = tcld*qc/q550; else LWP form           */
    double reff;       /* effective radius (um); < 0 selects ice                    */
    double tcld;       /* optical depth share of this layer at 0.55 um              */
    double lwpth;      /* liquid water path share (g/m2)                            */
    double q550;       /* extinction efficiency at 0.55 um (first-call cache)       */
} sbd_cloud_entry;

/* Upload the packed table bundle (frontend/device.py); index = offsets then
 * lengths, 2 x ntab int32.  Call once per handle. */
int sbd_optics_upload_tables(sbd_handle *h, const double *tables, int64_t ntables_doubles,
                             const int32_t *index, int32_t ntab);

/* Produce every bin of a run on the device and solve it.  All array arguments
 * are HOST pointers.  Outputs (host): nk[nwl], wl[nwl], dwl[nwl], wt[3*nwl]
 * (slot 3*il+kd), *nbins = sum nk; fluxes compact per bin in loop order:
 * rfldir, rfldn, flup [nbins][nz+1] (buffers sized for 3*nwl bins), uu
 * [nbins][nphi][nz+1][numu] when numu > 0, status[nbins].  If inputs_out is
 * non-NULL the produced DISORT inputs are copied back too (parity tests):
 * dtauc, ssalb [3 nwl][nz], pmom [3 nwl][nz][nmom+1], bins [3 nwl]. */
typedef struct sbd_inputs_out {
    double *dtauc, *ssalb, *pmom;
    sbd_bin *bins;
} sbd_inputs_out;

/* Aerosols of the following sbd_spectrum_run calls (tauaero, tauaero.f:1175-1358).
 * The wavelength-independent part of the reference's first call (denprfl
 * tauaero.f:1361: vertical profile, humidity-interpolated spectral model,
 * normalisation to vis / tbaer; zlayer of the stratospheric layers) is done by
 * the host front end; the producer kernel evaluates aerbwi (tauaero.f:177),
 * aestrat (:253), getmom and the moment mixing per wavelength.  All pointers are
 * HOST pointers; p == NULL switches aerosols off again. */
#define SBD_NAERW 47           /* wavelengths of the standard models (naerw, tauaero.f:21) */
#define SBD_NAERZ 5            /* stratospheric layers (naerz, tauaero.f:19)               */
typedef struct sbd_aerosol_params {
    int32_t nwlbaer;   /* spectral points of the boundary-layer model, 0: none (iaer=0)  */
    int32_t imoma;     /* phase function of the boundary-layer aerosol: 2 or 3 (getmom)  */
    int32_t nosct;     /* 0 normal, 1/2/3 no-scattering variants (tauaero.f:1266-1275)    */
    int32_t nstrat;    /* stratospheric layers in use, <= SBD_NAERZ                       */
    int32_t nz;        /* length of dtsv = levels of the run                              */
    int32_t pad_;
    double abaer;      /* Angstrom exponent outside the tabulated range                   */
} sbd_aerosol_params;

typedef struct sbd_strat_entry {
    double layer;      /* DISORT layer, 1-based from the top (laer, tauaero.f:1216)       */
    double taerst;     /* optical depth at 0.55 um                                        */
    double ext[SBD_NAERW], absb[SBD_NAERW], asym[SBD_NAERW];   /* aerstr(:,1:3,jaer)      */
} sbd_strat_entry;

int sbd_spectrum_set_aerosols(sbd_handle *h, const sbd_aerosol_params *p, const double *wlbaer,
                              const double *aerext, const double *aerabs, const double *aerasm,
                              const double *dtsv, const double *awl, const sbd_strat_entry *strat);

int sbd_spectrum_run(sbd_handle *h, const sbd_optics_params *p, const double *z, const double *pr,
                     const double *t, const double *uu, const sbd_cloud_entry *clouds,
                     const double *wlalb, const double *alb, const double *wlsun, const double *sun,
                     int32_t numu, const double *umu, int32_t nphi, const double *phi,
                     int32_t *nk, double *wl, double *dwl, double *wt, int32_t *nbins,
                     double *rfldir, double *rfldn, double *flup, double *uuout, int32_t *status,
                     const sbd_inputs_out *inputs_out);

/* The same for ncol independent atmospheric columns sharing the spectral grid, clouds,
 * surface and aerosol settings (retrieval batches): z, pr, t are [ncol][nz], uu
 * [ncol][64][nz+1]; p->btemp / p->ttemp < 0 take each column's own surface / top temperature
 * (drt.f:334-335).  One producer launch makes the bins of every (column, wavelength), one
 * solve launch follows on the same stream; nothing but the setup arrays goes to the device and
 * nothing but nk [ncol nwl], wt [3 ncol nwl], the fluxes and status come back, with ONE
 * synchronisation at the end.  Flux-only.  Fluxes and status are in loop order (column, then
 * wavelength, then k-term), sized for 3 ncol nwl bins; with sbd_set_flux_levels they are
 * [bin][nsel]. */
int sbd_spectrum_run_columns(sbd_handle *h, const sbd_optics_params *p, int32_t ncol, const double *z,
                             const double *pr, const double *t, const double *uu,
                             const sbd_cloud_entry *clouds, const double *wlalb, const double *alb,
                             const double *wlsun, const double *sun, int32_t *nk, double *wl, double *dwl,
                             double *wt, int32_t *nbins, double *rfldir, double *rfldn, double *flup,
                             int32_t *status);

/* Device copy of the flux results of the last sbd_spectrum_run* call on this handle
 * ([3][3 ncol nwl][nsel] with a level selection, else [3][3 ncol nwl][nz+1]: rfldir, rfldn,
 * flup), valid until the next call: what a multi-GPU caller hands to its all-gather. */
int sbd_spectrum_device_fluxes(sbd_handle *h, void **ptr, int64_t *ndoubles);

/* Bytes the last sbd_spectrum_run* call moved over PCIe (setup arrays in, results out). */
int sbd_last_transfer_bytes(const sbd_handle *h, int64_t *h2d, int64_t *d2h);

/* Diagnostic: sustained FP64 FMA throughput of the device in TFLOP/s (FMA = 2
 * flops), best of `reps` launches of a register-resident DFMA loop.  Used as
 * the roofline denominator of the solver kernel in bench.py. */
int sbd_measure_fp64_peak(sbd_handle *h, int reps, double *tflops);

const char *sbd_status_string(int code);
int sbd_abi_version(void);
/* Hash of the sources and compiler flags this binary was built from (the build
 * script passes it as -DSBD_BUILD_ID); the Python mirror refuses a library whose
 * id differs from the sources next to it. */
const char *sbd_build_id(void);

/*
 * gfortran-compatible replacement for SUBROUTINE DISORT (disort.f:1-6), call
 * site drt.f:541-546.  Everything by reference, LOGICAL = 4-byte int,
 * arrays column-major with the leading dimensions the caller passes
 * (PMOM(0:MAXMOM,MAXCLY), UU(MAXUMU,MAXULV,MAXPHI)), hidden length of
 * HEADER*127 appended by value.  Reproduces the reference's visible argument
 * mutations: NSTR -> -NSTR on beam/quadrature angle clash, NTAU/UTAU set to
 * the layer boundaries, NUMU/UMU set to the quadrature angles for flux runs,
 * SSALB=1 -> 1-DITHER, DTAUC<0 -> 0, PMOM(0,:) = 1.
 * Uses a process-wide handle on device 0 created on first use.
 */
void disort_(int *nlyr, double *dtauc, double *ssalb, int *corint, int *nmom,
             double *pmom, double *temper, double *wvnmlo, double *wvnmhi,
             int *usrtau, int *ntau, double *utau, int *nstr, int *usrang,
             int *numu, double *umu, int *nphi, double *phi, int *ibcnd,
             double *fbeam, double *umu0, double *phi0, double *fisot,
             int *lamber, double *albedo, double *btemp, double *ttemp,
             double *temis, int *plank, int *onlyfl, double *accur, int *prnt,
             char *header, int *maxcly, int *maxulv, int *maxumu, int *maxphi,
             int *maxmom, double *rfldir, double *rfldn, double *flup,
             double *dfdt, double *uavg, double *uu, double *albmed,
             double *trnmed, size_t header_len);

/*
 * Non-Lambertian surfaces (LAMBER = .FALSE.; replaces SURFAC's calls of the host function
 * BDREF, disort.f:3765-3907, spectra.f:249): the caller supplies SURFAC's tables,
 *   bdr [nsurf][nmodes][n][n+1]   BDR(iq, 0:n) of azimuth mode m (column 0: incidence at UMU0)
 *   bem [nsurf][n]                BEM(iq)
 *   rmu [nsurf][nmodes][numu][n+1], emu [nsurf][numu]   RMU / EMU at the user angles (radiance
 *                                 runs; NULL for flux runs),
 * n = nstr/2, nmodes = 1 (fluxes) or nstr (all azimuth modes).  HOST arrays, copied to the
 * device; they stay set until the next call (nsurf = 0 clears them).  A bin selects surface s
 * with bins[b].albedo = SBD_SURFACE(s).  Flux runs at the layer boundaries with NSTR
 * 4/8/16/20/24/32 and all radiance runs support it.  Launches given their own stream must
 * have finished before the call.
 */
#define SBD_SURFACE(s) (-(double)((s) + 1))
int sbd_set_surfaces(sbd_handle *h, int32_t nsurf, int32_t nstr, int32_t nmodes, int32_t numu,
                     const double *bdr, const double *bem, const double *rmu, const double *emu);

/* LAMBER = .FALSE. through disort_(): the library calls the host program's BDREF
 * (spectra.f:249: REAL(KR) FUNCTION BDREF(WVNMLO, WVNMHI, MUR, MUI, PHIR), kr = 8) -- the
 * symbol `bdref_` of the executable when it is exported (link with -rdynamic), or the function
 * handed over here. */
void sbd_set_bdref_callback(double (*bdref)(const double *wvnmlo, const double *wvnmhi,
                                            const double *mur, const double *mui,
                                            const double *phir));

/* status of the last disort_() call on this thread (SBD_BIN_* or SBD_ERR_*) */
int sbd_disort_last_status(void);

#ifdef __cplusplus
}
#endif
#endif /* SBDART_B200_H */
