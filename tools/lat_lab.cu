// Dev lab: dependent-issue latencies of FP64 ops / warp reductions on B200 (single warp, clock64).
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double* out, long long* cyc, double seed) {
    double x = seed + threadIdx.x * 1e-9, y = 1.0000001;
    long long t0, t1;
    const int N = 512;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) x = fma(x, y, 1e-9);
    t1 = clock64(); if (threadIdx.x == 0) cyc[0] = (t1 - t0);
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) x = x + y;
    t1 = clock64(); if (threadIdx.x == 0) cyc[1] = (t1 - t0);
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) x = __drcp_rn(x) + 1.5;
    t1 = clock64(); if (threadIdx.x == 0) cyc[2] = (t1 - t0);
    unsigned u = (unsigned)threadIdx.x + (unsigned)seed;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) u = __reduce_max_sync(0xffffffffu, u + threadIdx.x) + 1;
    t1 = clock64(); if (threadIdx.x == 0) cyc[3] = (t1 - t0);
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) u = __shfl_sync(0xffffffffu, u, (i + 1) & 31) + 1;
    t1 = clock64(); if (threadIdx.x == 0) cyc[4] = (t1 - t0);
    // 8 independent DFMA chains (throughput with ILP, one warp)
    double a0 = x, a1 = x + 1, a2 = x + 2, a3 = x + 3, a4 = x + 4, a5 = x + 5, a6 = x + 6, a7 = x + 7;
    t0 = clock64();
#pragma unroll 4
    for (int i = 0; i < N; i++) {
        a0 = fma(a0, y, 1e-9); a1 = fma(a1, y, 1e-9); a2 = fma(a2, y, 1e-9); a3 = fma(a3, y, 1e-9);
        a4 = fma(a4, y, 1e-9); a5 = fma(a5, y, 1e-9); a6 = fma(a6, y, 1e-9); a7 = fma(a7, y, 1e-9);
    }
    t1 = clock64(); if (threadIdx.x == 0) cyc[5] = (t1 - t0);
    __shared__ double sm[64];
    sm[threadIdx.x] = x; __syncwarp();
    int idx = threadIdx.x;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) { idx = ((int)sm[idx & 31] & 1) + ((idx + 1) & 31); }
    t1 = clock64(); if (threadIdx.x == 0) cyc[6] = (t1 - t0);
    out[threadIdx.x] = x + u + a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7 + idx;
}
int main() {
    double* out; long long* cyc; cudaMalloc(&out, 256 * 8); cudaMallocManaged(&cyc, 64);
    k<<<1, 32>>>(out, cyc, 1.0); cudaDeviceSynchronize();
    k<<<1, 32>>>(out, cyc, 1.0); cudaDeviceSynchronize();
    const char* names[] = {"DFMA dependent", "DADD dependent", "drcp_rn + DADD", "REDUX.max + IADD", "SHFL + IADD", "8 indep DFMA chains (per 8)", "LDS->F2I->addr chain"};
    for (int i = 0; i < 7; i++) printf("%-32s %.1f cycles/iter\n", names[i], cyc[i] / 512.0);
    return 0;
}
