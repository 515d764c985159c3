/*
 * disort_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU (plain C, FP64) restatement of the reference's per-bin discrete-ordinate
 * solve (SUBROUTINE DISORT, /root/reference/disort.f:1-871 and helpers).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may call into this; the product (sbdart_b200/) never
 * links, imports or executes it.
 *
 * Parity pin: DISORT's built-in self test (disort.f:6393-6449) -- see
 * tests/test_oracle_golden.py.
 */
#ifndef SBD_DISORT_ORACLE_H
#define SBD_DISORT_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

/* scalar arguments of SUBROUTINE DISORT (disort.f:1-6); logicals are ints */
typedef struct sbdo_input {
    int nlyr;      /* NLYR                                                  */
    int nstr;      /* NSTR  (even, >= 4)                                     */
    int nmom;      /* NMOM; pmom is [nlyr][nmom+1] (moment index fastest)    */
    int usrtau;    /* USRTAU; 0 => ntau = nlyr+1 layer boundaries            */
    int ntau;      /* NTAU  (input when usrtau)                              */
    int usrang;    /* USRANG                                                 */
    int numu;      /* NUMU  (input when usrang)                              */
    int nphi;      /* NPHI                                                   */
    int plank;     /* PLANK                                                  */
    int onlyfl;    /* ONLYFL                                                 */
    int corint;    /* CORINT                                                 */
    int lamber;    /* LAMBER; 0 => BDREF surface set with sbdo_set_bdref()    */
    double fbeam, umu0, phi0, fisot, albedo;
    double btemp, ttemp, temis, wvnmlo, wvnmhi, accur;
} sbdo_input;

/* status codes (mirror the reference's errmsg numbers) */
enum {
    SBDO_OK = 0,
    SBDO_ANGLE_CLASH = 1,      /* disort.f:2645-2650: NSTR returned negative */
    SBDO_BAD_INPUT = -1,       /* CHEKIN fatal, disort.f:5155                */
    SBDO_EIG_NOCONV = -2,      /* ASYMTX, disort.f:3254-3261                 */
    SBDO_SINGULAR = -3         /* exact zero pivot                           */
};

/*
 * One DISORT call.
 *   dtauc[nlyr], ssalb[nlyr], pmom[nlyr][nmom+1], temper[nlyr+1]
 *   utau[ntau]   (read when usrtau, else may be NULL)
 *   umu[numu]    (read when usrang, else may be NULL)
 *   phi[nphi]
 * Outputs (sized for NT = usrtau ? ntau : nlyr+1):
 *   rfldir, rfldn, flup, dfdt, uavg : [NT]
 *   uu  : [nphi][NT][numu]   (only when !onlyfl; may be NULL when onlyfl)
 *   u0u : [NT][NU] azimuthal average; NU = usrang ? numu : nstr; may be NULL
 *   warn: bit mask of non-fatal reference warnings (bit k = errmsg number k)
 */
int sbdo_disort(const sbdo_input *in, const double *dtauc, const double *ssalb,
                const double *pmom, const double *temper, const double *utau,
                const double *umu, const double *phi, double *rfldir,
                double *rfldn, double *flup, double *dfdt, double *uavg,
                double *uu, double *u0u, int *warn);

/* batched flux-only convenience used by bench.py's CPU leg: B independent
 * bins sharing dims; per-bin scalars given as arrays; OpenMP over bins when
 * compiled with -fopenmp.  temper is [ncol][nlyr+1], col[b] selects a row.
 * Outputs are [B][nlyr+1]. Returns number of bins with non-zero status. */
int sbdo_disort_flux_batch(int nbins, int nlyr, int nstr, int nmom,
                           const double *dtauc, const double *ssalb,
                           const double *pmom, const double *fbeam,
                           const double *umu0, const double *albedo,
                           const int *plank, const double *wvnmlo,
                           const double *wvnmhi, const double *btemp,
                           const double *ttemp, const double *temis,
                           const double *fisot, const double *temper,
                           const int *col, double *rfldir, double *rfldn,
                           double *flup, double *dfdt, double *uavg,
                           int *status, int nthreads);

/* Surface model of the following LAMBER = 0 calls (the reference's module albblk,
 * spectra.f:20-26, set by suralb spectra.f:139-162): ibdrf 1 ocean / 2 Hapke / 3 Ross-Li,
 * sc[5] as suralb stores them; nr, ni, rsw: the ocean model's table look-ups (INDWAT,
 * MORCASIWAT) for the wavelength of the call.  Not thread-safe (file-scope state). */
void sbdo_set_bdref(int ibdrf, const double *sc, double nr, double ni, double rsw);
double sbdo_bdref_eval(double mur, double mui, double phir);

/* pieces exported for unit tests */
void sbdo_qgausn(int m, double *gmu, double *gwt);
double sbdo_plkavg(double wnumlo, double wnumhi, double t, int *warn);
int sbdo_asymtx(double *aa, double *evec, double *eval, int m, int ia,
                int ievec, double *wk);

#ifdef __cplusplus
}
#endif
#endif
