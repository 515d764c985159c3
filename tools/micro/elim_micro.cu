// Microbenchmark of the phase-2 column elimination (sbd_fast.cu) in isolation:
// cycles per eliminated column at 1, 2 and 4 resident warps per SM sub-partition.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o elim_micro elim_micro.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define FULLMASK 0xffffffffu
constexpr int n = 8, N = 16, C = 2 * N + 1;

__device__ __forceinline__ double fast_rcp(double x)
{
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    y = fma(y, fma(-x, y, 1.0), y);
    y = fma(y, fma(-x, y, 1.0), y);
    return y;
}

// ---- V0: current kernel's column step ---------------------------------------
template <int W2>
__device__ __forceinline__ int elim_v0(double (&w)[2 * N + 2], double2 *pb, unsigned &act, int &mycol, int j, int lane)
{
    const bool cand = (act >> lane) & 1u;
    const double av = cand ? fabs(w[0]) : -1.0;
    const int hi = cand ? __double2hiint(av) : -1;
    const int mx = __reduce_max_sync(FULLMASK, hi);
    const unsigned who = __ballot_sync(FULLMASK, hi == mx && cand);
    if (mx <= 0 || who == 0) return 1;
    const int pl = __ffs(who) - 1;
    const bool ispiv = (lane == pl);
    if (ispiv) {
#pragma unroll
        for (int c2 = 0; c2 <= W2; c2++) pb[c2] = make_double2(w[2 * c2], w[2 * c2 + 1]);
        mycol = j;
    }
    __syncwarp();
    if (cand && !ispiv) {
        const double2 p0 = pb[0];
        const double mlt = w[0] * fast_rcp(p0.x);
        w[0] = fma(-mlt, p0.y, w[1]);
#pragma unroll
        for (int c2 = 1; c2 <= W2; c2++) {
            const double2 p = pb[c2];
            w[2 * c2 - 1] = fma(-mlt, p.x, w[2 * c2]);
            w[2 * c2] = fma(-mlt, p.y, w[2 * c2 + 1]);
        }
    }
    act &= ~(1u << pl);
    return 0;
}

// ---- V1: look-ahead pivot search, reciprocal off the chain, unconditional update ----
// state carried between columns: pl = pivot lane of the column about to be eliminated,
// rp = 1 / w[0] of this lane (speculative).
template <int W2>
__device__ __forceinline__ int elim_v1(double (&w)[2 * N + 2], double2 *pb, double *prc, unsigned &act,
                                       int &pl, double &rp, int lane)
{
    if (lane == pl) {
#pragma unroll
        for (int c2 = 0; c2 <= W2; c2++) pb[c2] = make_double2(w[2 * c2], w[2 * c2 + 1]);
        *prc = rp;
    }
    act &= ~(1u << pl);
    __syncwarp();
    const double2 p0 = pb[0];
    const double mlt = -w[0] * *prc;
    const double w0 = fma(mlt, p0.y, w[1]);
    // pivot of the next column
    const bool cand = (act >> lane) & 1u;
    const int hi = cand ? __double2hiint(fabs(w0)) : -1;
    const int mx = __reduce_max_sync(FULLMASK, hi);
    const unsigned who = __ballot_sync(FULLMASK, hi == mx);
    rp = fast_rcp(w0);
    w[0] = w0;
#pragma unroll
    for (int c2 = 1; c2 <= W2; c2++) {
        const double2 p = pb[c2];
        w[2 * c2 - 1] = fma(mlt, p.x, w[2 * c2]);
        w[2 * c2] = fma(mlt, p.y, w[2 * c2 + 1]);
    }
    pl = __ffs(who) - 1;
    return mx <= 0;
}


__device__ __forceinline__ void sts128_pred(double2 *p, double x, double y, int pred)
{
    const unsigned a = (unsigned)__cvta_generic_to_shared(p);
    asm volatile("{ .reg .pred q; setp.ne.s32 q, %3, 0; @q st.shared.v2.f64 [%0], {%1, %2}; }" ::"r"(a), "d"(x), "d"(y), "r"(pred) : "memory");
}
// ---- V3: V1 with predicated (non-divergent) publication ----
template <int W2>
__device__ __forceinline__ int elim_v3(double (&w)[2 * N + 2], double2 *pb, double *prc, unsigned &act,
                                       int &pl, double &rp, int lane)
{
    const int ip = lane == pl;
#pragma unroll
    for (int c2 = 0; c2 <= W2; c2++) sts128_pred(pb + c2, w[2 * c2], w[2 * c2 + 1], ip);
    if (ip) *prc = rp;
    act &= ~(1u << pl);
    __syncwarp();
    const double2 p0 = pb[0];
    const double mlt = -w[0] * *prc;
    const double w0 = fma(mlt, p0.y, w[1]);
    const bool cand = (act >> lane) & 1u;
    const int hi = cand ? __double2hiint(fabs(w0)) : -1;
    const int mx = __reduce_max_sync(FULLMASK, hi);
    const unsigned who = __ballot_sync(FULLMASK, hi == mx);
    rp = fast_rcp(w0);
    w[0] = w0;
#pragma unroll
    for (int c2 = 1; c2 <= W2; c2++) {
        const double2 p = pb[c2];
        w[2 * c2 - 1] = fma(mlt, p.x, w[2 * c2]);
        w[2 * c2] = fma(mlt, p.y, w[2 * c2 + 1]);
    }
    pl = __ffs(who) - 1;
    return mx <= 0;
}

// ---- V2: pivot row broadcast by shuffles (no shared memory, no warp barrier) ----
template <int W>   // W = number of live entries to move (<= 2N+1)
__device__ __forceinline__ int elim_v2(double (&w)[2 * N + 2], unsigned &act, int &pl, double &rp, int &mycol, int j, int lane)
{
    const bool ispiv = lane == pl;
    if (ispiv) mycol = j;
    act &= ~(1u << pl);
    const bool upd = (act >> lane) & 1u;
    const double prp = __shfl_sync(FULLMASK, rp, pl);
    const double mlt = -w[0] * prp;
    const double p1 = __shfl_sync(FULLMASK, w[1], pl);
    const double w0 = fma(mlt, p1, w[1]);
    const int hi = upd ? __double2hiint(fabs(w0)) : -1;
    const int mx = __reduce_max_sync(FULLMASK, hi);
    const unsigned who = __ballot_sync(FULLMASK, hi == mx);
    if (upd) { rp = fast_rcp(w0); w[0] = w0; }
#pragma unroll
    for (int c = 2; c < W; c++) {
        const double p = __shfl_sync(FULLMASK, w[c], pl);
        if (upd) w[c - 1] = fma(mlt, p, w[c]);
    }
    pl = __ffs(who) - 1;
    return mx <= 0;
}

__device__ __forceinline__ double prand(unsigned a, unsigned b, unsigned c)
{
    unsigned x = a * 2654435761u ^ b * 40503u ^ c * 2246822519u;
    x ^= x >> 15; x *= 2246822519u; x ^= x >> 13; x *= 3266489917u; x ^= x >> 16;
    return (double)(x & 0xffffff) / 16777216.0 - 0.5;
}

template <int V>
__global__ void __launch_bounds__(128, 4) micro(double *out, long long *cyc, int L, double *ubuf_g)
{
    extern __shared__ double sm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int UB = 17 * 34 + 64;
    double *wsm = sm + warp * (2 * UB + 8);
    double w[C + 1];
#pragma unroll
    for (int c = 0; c < C + 1; c++) w[c] = 0.0;
    if (lane < n) {
#pragma unroll
        for (int c = 0; c < N; c++) w[c] = prand(lane, c, 999) + (c == lane ? 2.0 : 0.0);
        w[2 * N] = 1.0;
    }
    unsigned live = (1u << n) - 1u;
    double chk = 0.0;
    int status = 0;
    const long long t0 = clock64();
    for (int lc = 0; lc < L; lc++) {
        const unsigned freem = ~live & ((1u << (n + N)) - 1u);
        const int rank = __popc(freem & ((1u << lane) - 1u));
        const bool isnew = ((freem >> lane) & 1u) && rank < N;
        if (isnew) {
#pragma unroll
            for (int c = 0; c < 2 * N; c++) w[c] = prand(rank, c, lc) + ((c == rank || c == rank + N) ? 1.5 : 0.0);
            w[2 * N] = prand(rank, 77, lc);
            w[2 * N + 1] = 0.0;
        }
        unsigned act = live | __ballot_sync(FULLMASK, isnew);
        double *ub = wsm + (lc & 1) * UB;
        if (V == 0) {
            int mycol = -1;
            double2 *prow2 = reinterpret_cast<double2 *>(wsm);
#pragma unroll 1
            for (int j = 0; j < N / 2 && !status; j++) status = elim_v0<N>(w, prow2 + (j & 1) * (N + 1), act, mycol, j, lane);
#pragma unroll 1
            for (int j = N / 2; j < N && !status; j++) status = elim_v0<N - N / 4>(w, prow2 + (j & 1) * (N + 1), act, mycol, j, lane);
            if (mycol >= 0) {
                double2 *u2 = reinterpret_cast<double2 *>(ubuf_g + ((size_t)(blockIdx.x * 4 + warp) * 2 + (lc & 1)) * 17 * 34 + mycol * 34);
#pragma unroll
                for (int c2 = 0; c2 <= N; c2++) u2[c2] = make_double2(w[2 * c2], w[2 * c2 + 1]);
                chk += w[0] + w[1] * 0.5 + w[3];
            }
        } else if (V == 2) {
            int mycol = -1;
            const bool cand = (act >> lane) & 1u;
            const int hi = cand ? __double2hiint(fabs(w[0])) : -1;
            const int mx = __reduce_max_sync(FULLMASK, hi);
            int pl = __ffs(__ballot_sync(FULLMASK, hi == mx)) - 1;
            double rp = fast_rcp(w[0]);
            if (mx <= 0) status = 1;
#pragma unroll 1
            for (int j = 0; j < N / 2 && !status; j++) status = elim_v2<2 * N + 1>(w, act, pl, rp, mycol, j, lane);
#pragma unroll 1
            for (int j = N / 2; j < N && !status; j++) status = elim_v2<2 * N + 1 - N / 2>(w, act, pl, rp, mycol, j, lane);
            status = 0;
            if (mycol >= 0) {
                double2 *u2 = reinterpret_cast<double2 *>(ubuf_g + ((size_t)(blockIdx.x * 4 + warp) * 2 + (lc & 1)) * 17 * 34 + mycol * 34);
#pragma unroll
                for (int c2 = 0; c2 <= N; c2++) u2[c2] = make_double2(w[2 * c2], w[2 * c2 + 1]);
                chk += w[0] + w[1] * 0.5 + w[3];
            }
        } else {
            // first pivot of the layer
            const bool cand = (act >> lane) & 1u;
            const int hi = cand ? __double2hiint(fabs(w[0])) : -1;
            const int mx = __reduce_max_sync(FULLMASK, hi);
            int pl = __ffs(__ballot_sync(FULLMASK, hi == mx)) - 1;
            double rp = fast_rcp(w[0]);
            double *prc = wsm + 2 * UB;
            if (mx <= 0) status = 1;
#pragma unroll 1
            for (int j = 0; j < N / 2 && !status; j++) {
                if (lane == pl) chk += w[0] + w[1] * 0.5 + w[3];
                status = (V == 3 ? elim_v3<N>(w, reinterpret_cast<double2 *>(ub + j * 34), prc + (j & 1), act, pl, rp, lane) : elim_v1<N>(w, reinterpret_cast<double2 *>(ub + j * 34), prc + (j & 1), act, pl, rp, lane));
            }
#pragma unroll 1
            for (int j = N / 2; j < N && !status; j++) {
                if (lane == pl) chk += w[0] + w[1] * 0.5 + w[3];
                status = (V == 3 ? elim_v3<N - N / 4>(w, reinterpret_cast<double2 *>(ub + j * 34), prc + (j & 1), act, pl, rp, lane) : elim_v1<N - N / 4>(w, reinterpret_cast<double2 *>(ub + j * 34), prc + (j & 1), act, pl, rp, lane));
            }
            status = 0;   // the look-ahead of the last column sees no candidates
            __syncwarp();
            // pivot rows -> global, coalesced
            double2 *g2 = reinterpret_cast<double2 *>(ubuf_g + ((size_t)(blockIdx.x * 4 + warp) * 2 + (lc & 1)) * 17 * 34);
            const double2 *s2 = reinterpret_cast<const double2 *>(ub);
            for (int i = lane; i < 16 * 17; i += 32) g2[i] = s2[i];
        }
        live = act;
        if ((live >> lane) & 1u) {
            w[2 * N] = w[N];
#pragma unroll
            for (int j = N; j < 2 * N; j++) w[j] = 0.0;
            w[2 * N + 1] = 0.0;
        }
        __syncwarp();
    }
    const long long t1 = clock64();
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) chk += __shfl_xor_sync(FULLMASK, chk, o);
    if (lane == 0) {
        out[blockIdx.x * 4 + warp] = chk + status * 1e30;
        cyc[blockIdx.x * 4 + warp] = t1 - t0;
    }
}

template <int V>
void run(int ctas_per_sm, int L)
{
    int grid = 148 * ctas_per_sm;
    double *out, *ub; long long *cyc;
    cudaMalloc(&out, grid * 4 * 8); cudaMalloc(&cyc, grid * 4 * 8);
    cudaMalloc(&ub, (size_t)grid * 4 * 2 * 17 * 34 * 8);
    size_t smem = 4 * (2 * (17 * 34 + 64) + 8) * 8;
    cudaFuncSetAttribute(micro<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    micro<V><<<grid, 128, smem>>>(out, cyc, L, ub);
    cudaEventRecord(e0);
    micro<V><<<grid, 128, smem>>>(out, cyc, L, ub);
    cudaEventRecord(e1);
    cudaError_t e = cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double h[4]; long long hc[4];
    cudaMemcpy(h, out, 32, cudaMemcpyDeviceToHost); cudaMemcpy(hc, cyc, 32, cudaMemcpyDeviceToHost);
    printf("V%d warps/SMSP=%d: %s  %.3f ms  cycles/column (warp 0) = %.1f  -> SMSP cycles per column = %.1f  chk=%.12g\n",
           V, ctas_per_sm, cudaGetErrorString(e), ms, (double)hc[0] / (L * N), (double)hc[0] / (L * N) / ctas_per_sm, h[0]);
    cudaFree(out); cudaFree(cyc); cudaFree(ub);
}

int main()
{
    const int L = 33 * 8;
    for (int c = 1; c <= 4; c *= 2) { run<0>(c, L); run<1>(c, L); run<2>(c, L); run<3>(c, L); }
    return 0;
}
