// Where do the cycles of one elimination column go?  Builds the column step up
// piece by piece (T1 DFMA stream from registers ... T5 full step) and prints
// cycles per step for 1, 2, 4 warps per SM sub-partition.
#include <cstdio>
#include <cuda_runtime.h>
#define FULLMASK 0xffffffffu
constexpr int N = 16, C = 2 * N + 1;

__device__ __forceinline__ double2 lds128(const double2 *p)
{
    double2 v; unsigned a = (unsigned)__cvta_generic_to_shared(p);
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(a));
    return v;
}
__device__ __forceinline__ double lds64(const double *p)
{
    double v; unsigned a = (unsigned)__cvta_generic_to_shared(p);
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ int lds32(const int *p)
{
    int v; unsigned a = (unsigned)__cvta_generic_to_shared(p);
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ double fast_rcp(double x)
{
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    y = fma(y, fma(-x, y, 1.0), y);
    y = fma(y, fma(-x, y, 1.0), y);
    return y;
}

template <int T>
__global__ void __launch_bounds__(128, 4) k(double *out, long long *cyc, int iters)
{
    extern __shared__ double sm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double2 *pb = reinterpret_cast<double2 *>(sm + warp * 80);
    double w[C + 1];
#pragma unroll
    for (int c = 0; c < C + 1; c++) w[c] = 1.0 + 1e-3 * (lane + c);
    for (int c = lane; c < 40; c += 32) pb[c] = make_double2(1e-3 * c, 2e-3 * c);
    __syncwarp();
    double mlt = 1e-6 * lane;
    int pl = 3;
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < iters; it++) {
        if (T == 1) {          // DFMA stream, pivot row in registers (same value)
            const double px = mlt * 3.0, py = mlt * 5.0;
            w[0] = fma(mlt, py, w[1]);
#pragma unroll
            for (int c2 = 1; c2 <= N; c2++) {
                w[2 * c2 - 1] = fma(mlt, px, w[2 * c2]);
                w[2 * c2] = fma(mlt, py, w[2 * c2 + 1]);
            }
        }
        if (T == 6) {          // DFMA stream in place (no shifting)
            const double px = mlt * 3.0, py = mlt * 5.0;
#pragma unroll
            for (int c2 = 0; c2 <= N; c2++) {
                w[2 * c2] = fma(mlt, px, w[2 * c2]);
                w[2 * c2 + 1] = fma(mlt, py, w[2 * c2 + 1]);
            }
        }

        if (T == 7) {          // pivot row via 64-bit broadcast loads
            const double *pd = reinterpret_cast<const double *>(pb);
            w[0] = fma(mlt, lds64(pd + 1), w[1]);
#pragma unroll
            for (int c = 2; c <= 2 * N + 1; c++) w[c - 1] = fma(mlt, lds64(pd + c), w[c]);
        }
        if (T == 8) {          // pivot row via shuffles from the pivot lane
            double nw[C + 1];
#pragma unroll
            for (int c = 1; c <= 2 * N + 1; c++) nw[c - 1] = fma(mlt, __shfl_sync(FULLMASK, w[c], pl), w[c]);
#pragma unroll
            for (int c = 0; c <= 2 * N; c++) w[c] = nw[c];
        }
        if (T == 9) {          // pivot row via REDUX.OR into uniform registers
            double nw[C + 1];
            const bool ip = lane == pl;
#pragma unroll
            for (int c = 1; c <= 2 * N + 1; c++) {
                const unsigned lo = __reduce_or_sync(FULLMASK, ip ? (unsigned)__double2loint(w[c]) : 0u);
                const unsigned hi = __reduce_or_sync(FULLMASK, ip ? (unsigned)__double2hiint(w[c]) : 0u);
                nw[c - 1] = fma(mlt, __hiloint2double((int)hi, (int)lo), w[c]);
            }
#pragma unroll
            for (int c = 0; c <= 2 * N; c++) w[c] = nw[c];
        }
        if (T == 10) {         // 32-bit broadcast loads (2 per double)
            const int *pi = reinterpret_cast<const int *>(pb);
#pragma unroll
            for (int c = 1; c <= 2 * N + 1; c++) w[c - 1] = fma(mlt, __hiloint2double(lds32(pi + 2 * c + 1), lds32(pi + 2 * c)), w[c]);
        }
        if (T >= 2 && T <= 5) {
            if (T >= 3) {
                if (lane == pl) {
#pragma unroll
                    for (int c2 = 0; c2 <= N; c2++) pb[c2] = make_double2(w[2 * c2], w[2 * c2 + 1]);
                }
                __syncwarp();
            }
            const double2 p0 = lds128(pb);
            if (T >= 4) mlt = -w[0] * fast_rcp(p0.x) * 1e-9;
            w[0] = fma(mlt, p0.y, w[1]);
            if (T >= 5) {
                const int hi = __double2hiint(fabs(w[0])) + lane;
                const int mx = __reduce_max_sync(FULLMASK, hi);
                pl = __ffs(__ballot_sync(FULLMASK, hi == mx)) - 1;
            }
#pragma unroll
            for (int c2 = 1; c2 <= N; c2++) {
                const double2 p = lds128(pb + c2);
                w[2 * c2 - 1] = fma(mlt, p.x, w[2 * c2]);
                w[2 * c2] = fma(mlt, p.y, w[2 * c2 + 1]);
            }
            if (T >= 3) __syncwarp();
        }
    }
    const long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int c = 0; c < C + 1; c++) s += w[c];
    out[blockIdx.x * 128 + threadIdx.x] = s + pl;
    if (lane == 0) cyc[blockIdx.x * 4 + warp] = t1 - t0;
}

template <int T>
void run(int ctas)
{
    const int grid = 148 * ctas, iters = 4000;
    double *out; long long *cyc;
    cudaMalloc(&out, grid * 128 * 8); cudaMalloc(&cyc, grid * 4 * 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<T><<<grid, 128, 4 * 80 * 8>>>(out, cyc, iters);
    cudaEventRecord(e0);
    k<T><<<grid, 128, 4 * 80 * 8>>>(out, cyc, iters);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long hc[4]; cudaMemcpy(hc, cyc, 32, cudaMemcpyDeviceToHost);
    printf("T%d warps/SMSP=%d cycles/step/warp=%.1f  per SMSP=%.1f  wall: %.3f ms = %.1f ns/step/warp\n", T, ctas, (double)hc[0] / iters,
           (double)hc[0] / iters / ctas, ms, ms * 1e6 / iters);
    cudaFree(out); cudaFree(cyc);
}

int main()
{
    for (int c = 1; c <= 4; c *= 4) { run<6>(c); run<2>(c); run<7>(c); run<10>(c); run<8>(c); run<9>(c); run<3>(c); }
    return 0;
}
