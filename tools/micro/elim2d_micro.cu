// Microbenchmark of the 2-D tiled pivot step of sbd_fast.cu in isolation: wall time
// per step at 1, 2, 4 resident warps per SM sub-partition.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a --expt-relaxed-constexpr \
//        -I../../sbdart_b200/csrc -o elim2d_micro elim2d_micro.cu
#include <cstdio>
#include "../../sbdart_b200/csrc/sbd_fast.cu"

using namespace sbd;
constexpr int n = 8;
using FL = FastLayout<n>;

__device__ __forceinline__ double prand(unsigned a, unsigned b, unsigned c)
{
    unsigned x = a * 2654435761u ^ b * 40503u ^ c * 2246822519u;
    x ^= x >> 15; x *= 2246822519u; x ^= x >> 13;
    return (double)(x & 0xffffff) / 16777216.0 - 0.5;
}


// variants of the pivot step: bit 0 = no REDUX/VOTE (fixed pivot), bit 1 = no scratch store,
// bit 2 = no speculative reciprocal (use 1.0), bit 3 = no colj shuffles, bit 4 = no pivot-row shuffles
template <int VAR, int W>
__device__ __forceinline__ bool step_var(double (&w)[FL::KS][FL::LC], double (&rhs)[FL::KS], unsigned &act,
                                         double *uslice, int cgj, int rg, int cg, int j)
{
    const int lane = threadIdx.x & 31;
    constexpr int KS = FL::KS, LC = FL::LC;
    const unsigned cand = (cg == cgj) ? ((act >> rg) & 0x010101u) : 0u;
    double colj[KS];
#pragma unroll
    for (int k = 0; k < KS; k++) colj[k] = (VAR & 8) ? w[k][0] : __shfl_sync(FULLMASK, w[k][0], (rg << 2) | cgj);
    int best = -1;
    double bval = 1.0;
#pragma unroll
    for (int k = 0; k < KS; k++) {
        const int h = ((cand >> (8 * k)) & 1u) ? ((__double2hiint(w[k][0]) & 0x7ffffffc) | (KS - 1 - k)) : -1;
        if (h > best) { best = h; bval = w[k][0]; }
    }
    const double rloc = (VAR & 4) ? bval * 0.5 : fast_rcp(bval);
    int mx; unsigned who;
    if (VAR & 1) { mx = 0x3ff00000 | (2 - (j % 3)); who = 1u << ((j * 5) & 31); }
    else { mx = __reduce_max_sync(FULLMASK, best); who = __ballot_sync(FULLMASK, best == mx); }
    if ((mx >> 2) <= 0) return true;
    const int pl = __ffs(who) - 1, kp = KS - 1 - (mx & 3), rgp = pl >> 2;
    const double rp = -__shfl_sync(FULLMASK, rloc, pl);
    double m[KS];
#pragma unroll
    for (int k = 0; k < KS; k++) m[k] = colj[k] * rp * ((VAR & 1) ? 1e-3 : 1.0);
    double p[W], pr;
    const int src = (rgp << 2) | cg;
    if (VAR & 16) {
#pragma unroll
        for (int l = 0; l < W; l++) p[l] = w[1][l] * 0.999;
        pr = rhs[1];
    } else if (kp == 0) {
#pragma unroll
        for (int l = 0; l < W; l++) p[l] = __shfl_sync(FULLMASK, w[0][l], src);
        pr = __shfl_sync(FULLMASK, rhs[0], src);
    } else if (kp == 1) {
#pragma unroll
        for (int l = 0; l < W; l++) p[l] = __shfl_sync(FULLMASK, w[1][l], src);
        pr = __shfl_sync(FULLMASK, rhs[1], src);
    } else {
#pragma unroll
        for (int l = 0; l < W; l++) p[l] = __shfl_sync(FULLMASK, w[2][l], src);
        pr = __shfl_sync(FULLMASK, rhs[2], src);
    }
    act &= ~(1u << (kp * 8 + rgp));
    if (!(VAR & 512)) {
    if (VAR & 256) {           // transposed slice layout: one STG.128 writes 64 contiguous bytes
        if (rg == 0) {
            double *rowbase = uslice - cg * LC;
#pragma unroll
            for (int l2 = 0; l2 < W / 2; l2++)
                reinterpret_cast<double2 *>(rowbase + l2 * 8)[cg] = make_double2(p[2 * l2], p[2 * l2 + 1]);
            if (cg == 0) rowbase[4 * LC] = pr;
        }
    } else if (VAR & 64) {            // streaming stores
        if (rg == 0) {
#pragma unroll
            for (int l2 = 0; l2 < W / 2; l2++)
                __stcs(reinterpret_cast<double2 *>(uslice) + l2, make_double2(p[2 * l2], p[2 * l2 + 1]));
            if (cg == 0) __stcs(uslice + 4 * LC, pr);
        }
    } else if (VAR & 128) {    // one 16-lane store: row group q writes pair q of every slice
        double a0 = p[0], a1 = p[1];
        if (rg == 1) { a0 = p[2]; a1 = p[3]; }
        if (rg == 2) { a0 = p[4]; a1 = p[5]; }
        if (rg == 3) { a0 = p[6]; a1 = p[7]; }
        if (rg < 4) reinterpret_cast<double2 *>(uslice)[rg] = make_double2(a0, a1);
        if (lane == 16) uslice[4 * LC] = pr;
    } else if (!(VAR & 2) && rg == 0) {
#pragma unroll
        for (int l2 = 0; l2 < W / 2; l2++)
            reinterpret_cast<double2 *>(uslice)[l2] = make_double2(p[2 * l2], p[2 * l2 + 1]);
        if (cg == 0) uslice[4 * LC] = pr;
    }
    }
#pragma unroll
    for (int k = 0; k < KS; k++) rhs[k] = fma(m[k], pr, rhs[k]);
    if (cgj == 3) {
#pragma unroll
        for (int k = 0; k < KS; k++) {
#pragma unroll
            for (int l = 0; l + 1 < W; l++) w[k][l] = fma(m[k], p[l + 1], w[k][l + 1]);
            w[k][W - 1] = 0.0;
        }
    } else {
#pragma unroll
        for (int k = 0; k < KS; k++) {
#pragma unroll
            for (int l = 0; l < W; l++) w[k][l] = fma(m[k], p[l], w[k][l]);
        }
    }
    if ((VAR & 512) && rg == 0) {     // pivot row stored after the update stream
#pragma unroll
        for (int l2 = 0; l2 < W / 2; l2++)
            reinterpret_cast<double2 *>(uslice)[l2] = make_double2(p[2 * l2], p[2 * l2 + 1]);
        if (cg == 0) uslice[4 * LC] = pr;
    }
    return false;
}

template <int VAR>
__global__ void __launch_bounds__(128, 4) micro(double *out, double *ubuf, int L)
{
    constexpr int KS = FL::KS, LC = FL::LC, N = 2 * n, US = FL::US;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int rg = lane >> 2, cg = lane & 3;
    double w[KS][LC], rhs[KS];
#pragma unroll
    for (int k = 0; k < KS; k++) {
        rhs[k] = 0.0;
#pragma unroll
        for (int l = 0; l < LC; l++) w[k][l] = 0.0;
    }
    // top rows in slots 0..n-1
#pragma unroll
    for (int l = 0; l < LC / 2; l++) w[0][l] = prand(rg, 4 * l + cg, 999) + ((4 * l + cg) == rg ? 2.0 : 0.0);
    rhs[0] = 1.0;
    unsigned act = (1u << n) - 1u;
    double *ul = ubuf + (size_t)(blockIdx.x * 4 + warp) * FL::ublk;
    int bad = 0;
    for (int lc = 0; lc < L; lc++) {
        const unsigned freem = ~act & FL::slotmask;
#pragma unroll
        for (int k = 0; k < KS; k++) {
            const int s = k * 8 + rg;
            if ((freem >> s) & 1u) {
                const int r = __popc(freem & ((1u << s) - 1u));
#pragma unroll
                for (int l = 0; l < LC; l++) {
                    const int c = 4 * l + cg;
                    w[k][l] = prand(r, c, lc) + ((c == r || c == r + N) ? 1.5 : 0.0);
                }
                rhs[k] = prand(r, 77, lc);
            }
        }
        act |= freem;
        if (VAR & 1) act = 0xffffffu;
        double *uslice = ul + cg * LC;
        bool sing = false;
        if (VAR & 32) {
#pragma unroll 1
            for (int j = 0; j < N / 2 && !sing; j++, uslice += US) sing = step_var<VAR, LC>(w, rhs, act, uslice, j & 3, rg, cg, j);
#pragma unroll 1
            for (int j = N / 2; j < N && !sing; j++, uslice += US) sing = step_var<VAR, 6>(w, rhs, act, uslice, j & 3, rg, cg, j);
        } else {
#pragma unroll 1
            for (int j = 0; j < N && !sing; j++, uslice += US) sing = step_var<VAR, LC>(w, rhs, act, uslice, j & 3, rg, cg, j);
        }
        bad |= sing;
    }
    double chk = rhs[0] + rhs[1] + rhs[2];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) chk += __shfl_xor_sync(FULLMASK, chk, o);
    if (lane == 0) out[blockIdx.x * 4 + warp] = chk + bad * 1e30;
}

template <int VAR>
void run(int c)
{
    const int L = 33 * 8;
    const int grid = 148 * c;
    double *out, *ub;
    cudaMalloc(&out, grid * 4 * 8);
    cudaMalloc(&ub, (size_t)grid * 4 * FL::ublk * 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    micro<VAR><<<grid, 128>>>(out, ub, L);
    cudaEventRecord(e0);
    micro<VAR><<<grid, 128>>>(out, ub, L);
    cudaEventRecord(e1);
    cudaError_t e = cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double steps_per_warp = (double)L * 16;
    printf("var %2d warps/SMSP=%d: %s %.3f ms -> %.0f cycles per step per warp, %.1f per step per SM\n",
           VAR, c, cudaGetErrorString(e), ms, ms * 1e-3 * 1.965e9 / steps_per_warp,
           ms * 1e-3 * 1.965e9 / (steps_per_warp * 4 * c));
    cudaFree(out); cudaFree(ub);
}

int main()
{
    for (int c = 4; c <= 4; c *= 2) {
        run<0>(c); run<512>(c); run<2>(c);
    }
    return 0;
}
