// Internal declarations shared by the C-ABI host code and the CUDA kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/sbdart_b200.h"

namespace sbd {

// DITHER of the reference: 10*R1MACH(4), times 10 again because it is
// below 1e-10 (disort.f:442-448) => 100 * 2^-52.
constexpr double kDither = 100.0 * 2.220446049250313e-16;
// PI = 2.*ASIN(1.0) evaluated in default REAL and widened (disort.f:441).
constexpr double kPiRef = 3.1415927410125732;

// Everything a kernel launch needs (device pointers).
struct LaunchArgs {
    sbd_dims d;
    const double *dtauc, *ssalb, *pmom;
    const sbd_bin *bins;
    const double *temper, *utau, *umu, *phi;
    double *rfldir, *rfldn, *flup, *dfdt, *uavg, *uu;
    int32_t *status;
    const double *quad;   // [2n]: mu[0..n), wt[0..n)
    const double *ylmc;   // [M][N][n]   Y_l^m(mu_i)
    const double *ylmu;   // [M][N][numu] Y_l^m(umu_iu)  (radiance runs)
    double *scratch;      // nslots * slot_stride doubles
    size_t slot_stride;
    int nslots;           // number of concurrently resident warps
    int nmodes;           // azimuth modes to run (1 for flux-only)
    int *work_counter;    // dynamic bin scheduler
    const int32_t *binmap; // optional: bin b reads its inputs from slot binmap[b]
    const int32_t *nbins_dev; // optional: number of bins lives on the device (spectrum path)
    // bins the adding kernel hands to the elimination kernel (non-monotone TAUC): the adding
    // kernel appends, the elimination kernel launched behind it works through the list
    int *redo_count;
    int *redo_list;
    bool redo_consume;        // this launch takes its bins from the list
    // BRDF surfaces (sbd_set_surfaces): bdr [ns][modes][n][n+1], bem [ns][n],
    // rmu [ns][modes][numu][n+1], emu [ns][numu]
    const double *sf_bdr, *sf_bem, *sf_rmu, *sf_emu;
    int sf_count, sf_modes;
    unsigned long long uu_mask[2]; // bit lu: intensities wanted at output level lu
    // layout of uu: [bin][nphi][uu_nt][numu]; output level lu lives in slot uu_slot[lu]
    // (-1: not wanted).  Full layout: uu_nt = NT, slot = level; packed: the wanted levels only.
    int uu_nt;
    short uu_slot[SBD_MAX_NLYR + 2];
};

// Per-layer record kept in scratch between the downward elimination sweep
// and the upward back-substitution sweep.
struct LayerLayout {
    int n, N, NU;
    int off_kk, off_ek, off_gp, off_gm, off_zz, off_zp0, off_xr, off_u,
        off_gu, off_zb, off_z0u, off_z1u, stride;
    __host__ __device__ LayerLayout(int N_, int NU_) {
        N = N_; n = N_ / 2; NU = NU_;
        int o = 0;
        off_kk = o; o += n;
        off_ek = o; o += n;
        off_gp = o; o += n * n;
        off_gm = o; o += n * n;
        off_zz = o; o += N;
        off_zp0 = o; o += N;
        off_xr = o; o += 2;
        off_u = o; o += N * (2 * N + 1);
        off_gu = o; o += NU * N;
        off_zb = o; o += NU;
        off_z0u = o; o += NU;
        off_z1u = o; o += NU;
        stride = (o + 1) & ~1;
    }
};

size_t generic_smem_bytes(int N, int L, int NT, int warps);
size_t generic_slot_doubles(int N, int L, int NU);
int generic_pick_warps(int N, int L, int NT, size_t smem_limit);
cudaError_t launch_generic(const LaunchArgs &a, int warps, int grid,
                           cudaStream_t st);

// Nakajima-Tanaka intensity corrections after a radiance launch (sbd_intcor.cu)
size_t intcor_smem_bytes(int L, int nmom);
cudaError_t launch_intcor(const LaunchArgs &a, cudaStream_t st);

// register-resident kernel for NSTR in {4, 8, 16} (sbd_fast.cu)
bool fast_supported(int N);
bool fast_rad_supported(int N);
int fast_warps();
size_t fast_slot_doubles(int N, int L, int NU);
size_t fast_smem_bytes(int N, int L, int NT, int warps, int NU, int NPHI);
cudaError_t launch_fast(const LaunchArgs &a, int warps, int grid, cudaStream_t st);

// adding kernel: NSTR in {4, 8, 16}, fluxes at the layer boundaries (sbd_adding.cu)
bool adding_supported(int N);
int adding_warps_per_sm(int N);
size_t adding_slot_doubles(int N, int L);
size_t adding_smem_bytes(int N, int L, int warps);
cudaError_t launch_adding(const LaunchArgs &a, int warps, int grid, cudaStream_t st);

// one-CTA-per-bin register kernel for NSTR in {20, 24, 32}, fluxes (sbd_wide.cu)
bool wide_supported(int N);
int wide_ctas_per_sm();
size_t wide_smem_bytes(int N, int L, int NT);
size_t wide_slot_doubles(int N, int L);
cudaError_t launch_wide(const LaunchArgs &a, int grid, cudaStream_t st);

}  // namespace sbd

// out[k][b][s] = in[k * per_in + b * NT + sel[s]] for narr arrays (sbd_spectrum.cu)
cudaError_t sbd_launch_pack_flux(const double *in, double *out, const int32_t *sel, int nsel, int NT,
                                 size_t per_in, size_t nbins, int narr, cudaStream_t st);
