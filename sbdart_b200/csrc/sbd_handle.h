// The opaque solver handle of the C ABI: a CUDA stream plus grow-only device buffers.
#pragma once
#include <cuda_runtime.h>

#include <map>
#include <vector>

#include "sbd_internal.h"
#include "sbd_optics.cuh"

struct SbdDevBuf {
    void *p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t bytes)
    {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        size_t want = bytes + bytes / 4 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

struct SbdTables {
    double *quad = nullptr;   // [2n]
    double *ylmc = nullptr;   // [N][N][n]
};

struct sbd_handle {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t copy_in = nullptr, copy_out = nullptr;   // H2D / D2H overlap in the host-buffer call
    static constexpr int kMaxChunks = 32;
    cudaEvent_t ev_in[kMaxChunks] = {}, ev_k[kMaxChunks] = {};
    int sm_count = 0;
    size_t smem_optin = 0;
    int64_t launches = 0;
    std::map<int, SbdTables> tables;          // per NSTR
    SbdDevBuf scratch, counter, ylmu, angles;
    // second scratch set + stream: consecutive chunk kernels of the host-buffer call
    // run on alternating streams, so the next chunk's CTAs fill the SMs the previous
    // chunk's tail leaves idle
    SbdDevBuf scratch2, counter2;
    // adding kernel: list of bins handed to the elimination kernel, and that kernel's scratch
    SbdDevBuf redo, redo2, redo_scratch, redo_scratch2;
    SbdDevBuf surfaces;                        // sbd_set_surfaces
    int sf_count = 0, sf_nstr = 0, sf_modes = 0, sf_numu = 0;
    cudaStream_t stream2 = nullptr;
    cudaEvent_t ev_misc = nullptr;
    int scratch_set = 0;
    bool uu_packed = false;                   // uu holds the selected levels only (sbd_set_radiance_layout)
    bool uu_packed_once = false;              // host-buffer call: one packed device launch, then back to uu_packed
    unsigned long long uu_mask[2] = { ~0ull, ~0ull };   // levels at which uu is wanted                       // set used by the next device-level launch
    // staging for the host-pointer API
    SbdDevBuf d_dtauc, d_ssalb, d_pmom, d_bins, d_temper, d_utau, d_out, d_uu, d_status;
    SbdDevBuf d_uupack, d_sel;                // selected-level intensities, compact, for the D2H copy
    // whole-spectrum path (sbd_spectrum.cu)
    SbdDevBuf opt_tables, opt_atm, opt_misc, opt_map, opt_aero;
    sbd_aerosol_params aero = {};             // aerosols of the next spectrum runs
    bool aero_on = false;
    bool corint = false;                      // INTCOR after radiance launches (sbd_set_corint)
    sbd::OpticsTables opt_index = {};
    bool opt_ready = false;
    const int32_t *pending_binmap = nullptr;   // device bin -> slot map for the next solve launch
    const int32_t *pending_nbins_dev = nullptr; // ... and its bin count, on the device
    std::vector<int32_t> flux_levels;          // sbd_set_flux_levels: host-side flux outputs are [bin][nsel]
    SbdDevBuf d_fluxpack;
    void *host_stage = nullptr;                // pinned staging of the whole-spectrum setup arrays
    size_t host_stage_cap = 0;
    size_t last_h2d_bytes = 0, last_d2h_bytes = 0;
    void *last_flux_dev = nullptr;             // device copy of the last spectrum run's flux results
    size_t last_flux_doubles = 0;
};
