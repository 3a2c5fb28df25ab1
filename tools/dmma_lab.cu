// Dev lab: what limits sustained DMMA.8x8x4 issue on B200 when the operands vary / come from shared memory.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/dmma_lab tools/dmma_lab.cu
#include <cstdio>
#include <cuda_runtime.h>
typedef double2 cplx;

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// MODE 0: complex 16x32 warp tile, operands fixed in registers
// MODE 1: operands re-read from shared memory every k-step (conflict-free LDS.128)
// MODE 2: MODE 1 + __syncthreads every 4 k-steps
// MODE 3: 3M variant of MODE 1 (3 DMMAs per complex tile product, 2 DADDs per fragment)
template <int MODE, int MI, int NI>
__global__ void __launch_bounds__(256) k(double* out, int iters, double seed) {
    __shared__ cplx Ps[64 * 20];
    __shared__ cplx Ws[16 * 66];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, gid = lane >> 2, tig = lane & 3;
    const int wm = warp >> 1, wn = warp & 1;
    for (int i = tid; i < 64 * 20; i += 256) Ps[i] = make_double2(seed + i * 1e-9, seed - i * 1e-9);
    for (int i = tid; i < 16 * 66; i += 256) Ws[i] = make_double2(seed - i * 1e-9, seed + i * 1e-9);
    __syncthreads();
    double cre[MI][NI][2], cim[MI][NI][2], c3[MI][NI][2];
#pragma unroll
    for (int mi = 0; mi < MI; mi++)
#pragma unroll
        for (int ni = 0; ni < NI; ni++)
#pragma unroll
            for (int e = 0; e < 2; e++) { cre[mi][ni][e] = seed * mi; cim[mi][ni][e] = seed * ni; c3[mi][ni][e] = seed; }
    cplx af[MI], bf[NI];
#pragma unroll
    for (int mi = 0; mi < MI; mi++) af[mi] = Ps[((wm * 16 + mi * 8 + gid) % 64) * 20 + tig];
#pragma unroll
    for (int ni = 0; ni < NI; ni++) bf[ni] = Ws[tig * 66 + (wn * 32 + ni * 8 + gid) % 64];
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int kk = 0; kk < 16; kk += 4) {
            if (MODE >= 1) {
#pragma unroll
                for (int mi = 0; mi < MI; mi++) af[mi] = Ps[((wm * 16 + mi * 8 + gid) % 64) * 20 + kk + tig];
#pragma unroll
                for (int ni = 0; ni < NI; ni++) bf[ni] = Ws[(kk + tig) * 66 + (wn * 32 + ni * 8 + gid) % 64];
            }
            if (MODE == 3) {
                double as[MI], bs[NI];
#pragma unroll
                for (int mi = 0; mi < MI; mi++) as[mi] = af[mi].x + af[mi].y;
#pragma unroll
                for (int ni = 0; ni < NI; ni++) bs[ni] = bf[ni].x + bf[ni].y;
#pragma unroll
                for (int mi = 0; mi < MI; mi++) {
#pragma unroll
                    for (int ni = 0; ni < NI; ni++) dmma884(cre[mi][ni][0], cre[mi][ni][1], af[mi].x, bf[ni].x);
#pragma unroll
                    for (int ni = 0; ni < NI; ni++) dmma884(cim[mi][ni][0], cim[mi][ni][1], af[mi].y, bf[ni].y);
#pragma unroll
                    for (int ni = 0; ni < NI; ni++) dmma884(c3[mi][ni][0], c3[mi][ni][1], as[mi], bs[ni]);
                }
            } else {
#pragma unroll
                for (int mi = 0; mi < MI; mi++) {
                    const double nay = -af[mi].y;
#pragma unroll
                    for (int ni = 0; ni < NI; ni++) dmma884(cre[mi][ni][0], cre[mi][ni][1], af[mi].x, bf[ni].x);
#pragma unroll
                    for (int ni = 0; ni < NI; ni++) dmma884(cim[mi][ni][0], cim[mi][ni][1], af[mi].x, bf[ni].y);
#pragma unroll
                    for (int ni = 0; ni < NI; ni++) dmma884(cre[mi][ni][0], cre[mi][ni][1], nay, bf[ni].y);
#pragma unroll
                    for (int ni = 0; ni < NI; ni++) dmma884(cim[mi][ni][0], cim[mi][ni][1], af[mi].y, bf[ni].x);
                }
            }
        }
        if (MODE == 2) __syncthreads();
    }
    double s = 0;
#pragma unroll
    for (int mi = 0; mi < MI; mi++)
#pragma unroll
        for (int ni = 0; ni < NI; ni++) s += cre[mi][ni][0] + cre[mi][ni][1] + cim[mi][ni][0] + cim[mi][ni][1] + c3[mi][ni][0] + c3[mi][ni][1];
    if (s == 123.456) out[0] = s;
}

template <int MODE, int MI, int NI>
void run(const char* name, int sms, double* out) {
    for (int cps = 1; cps <= 2; cps++) {
        const int iters = 4000;
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        k<MODE, MI, NI><<<sms * cps, 256>>>(out, iters, 1.0); cudaDeviceSynchronize();
        cudaEventRecord(e0); k<MODE, MI, NI><<<sms * cps, 256>>>(out, iters, 1.0); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        const double per = (MODE == 3 ? 3.0 : 4.0) * MI * NI * 4;           // DMMAs per warp per iteration
        const double dm = (double)sms * cps * 8 * iters * per;
        const double alg = (double)sms * cps * 8 * iters * 4.0 * MI * NI * 4 * 512.0;
        printf("%-40s ctas/sm=%d  dmma-rate %6.2f TF/s  algorithmic(4M-equiv) %6.2f TF/s\n", name, cps, dm * 512.0 / ms / 1e9, alg / ms / 1e9);
    }
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    double* out; cudaMalloc(&out, 8);
    run<0, 2, 4>("regs fixed, 16x32 warp tile", p.multiProcessorCount, out);
    run<1, 2, 4>("LDS operands, 16x32", p.multiProcessorCount, out);
    run<2, 2, 4>("LDS operands + barrier/16k, 16x32", p.multiProcessorCount, out);
    run<3, 2, 4>("3M LDS operands, 16x32", p.multiProcessorCount, out);
    run<1, 2, 2>("LDS operands, 16x16", p.multiProcessorCount, out);
    run<1, 4, 2>("LDS operands, 32x16", p.multiProcessorCount, out);
    run<1, 4, 4>("LDS operands, 32x32", p.multiProcessorCount, out);
    run<3, 4, 4>("3M LDS operands, 32x32", p.multiProcessorCount, out);
    return 0;
}
