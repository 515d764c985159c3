// Layout of the optical-property tables on the device and the launch arguments
// of the producer kernel (K2).  The table bundle is packed by
// sbdart_b200/frontend/device.py in exactly this order.
#pragma once
#include <stdint.h>

#include "../../include/sbdart_b200.h"

namespace sbd {

constexpr int kNumMol = 11;   // h2o co2 o3 n2o co ch4 o2 no so2 no2 nh3 (imol 1..11, taugas.f:2405-2415)

enum OpticsTable : int {
    T_SLF296 = 0, T_SLF260, T_FRN296,           // H2O continua, 2003 points each
    T_C4, T_H1, T_H2, T_H3,                      // N2 continuum, HNO3
    T_O2S0, T_O2A, T_O2B,                        // O2 1395-1760 cm-1
    T_O4SIG,                                     // O4 / O2-N2 collision complex
    T_O3S0, T_O3S1, T_O3S2, T_O3UV, T_C8,        // ozone Hartley-Huggins, UV, Chappuis
    T_SHN,                                       // Schumann-Runge
    T_KFAC,                                      // k-distribution factors (3)
    T_MIE_QQ, T_MIE_WW, T_MIE_GG, T_MIE_QQI, T_MIE_WWI, T_MIE_GGI,   // [13][400] each (re-major)
    T_CP0,                                       // + imol-1 : band-model coefficients C'
    T_IWL0 = T_CP0 + kNumMol,                    // band limits (as doubles, -999 terminated)
    T_IWH0 = T_IWL0 + kNumMol,
    T_BMS0 = T_IWH0 + kNumMol,                   // abcdta parameters per sub-band
    T_BMA0 = T_BMS0 + kNumMol,
    T_BMB0 = T_BMA0 + kNumMol,
    T_BMC0 = T_BMB0 + kNumMol,
    T_BANDS = T_BMC0 + kNumMol,                  // quadruples (imol, iw, lo, hi) of abcdta windows
    T_COUNT
};

struct OpticsTables {
    const double *base;
    int32_t off[T_COUNT];
    int32_t len[T_COUNT];
};

struct OpticsArgs {
    sbd_optics_params p;
    OpticsTables tab;
    int32_t ncol;                                // atmospheric columns (0 / 1: one)
    const double *z, *p_, *t, *uu;               // [ncol][nz] x 3, [ncol][64][nz+1]
    const double *btemp, *ttemp;                 // [ncol] boundary temperatures, or null: p.btemp / p.ttemp
    const sbd_cloud_entry *clouds;               // [p.ncloud]
    const double *wlalb, *alb, *wlsun, *sun;     // surface albedo and solar tables
    // aerosols (sbd_spectrum_set_aerosols): packed [wlb n][ext n][abs n][asm n][dtsv nz][awl 47][strat...]
    sbd_aerosol_params aer;
    const double *aero;                          // nullptr: no aerosols
    // outputs: slot = 3 * (col * nwl + il) + kd
    double *dtauc, *ssalb, *pmom;                // [3 nwl][nz], [3 nwl][nz], [3 nwl][nz][nmom+1]
    sbd_bin *bins;                               // [3 nwl]
    int32_t *nk;                                 // [ncol nwl]
    double *wl, *dwl, *wt;                       // [nwl], [nwl], [3 ncol nwl]
};

cudaError_t launch_optics(const OpticsArgs &a, cudaStream_t st);
cudaError_t launch_binmap(const int32_t *nk, int nitem, int32_t *binmap, int32_t *nbins, cudaStream_t st);

}  // namespace sbd
