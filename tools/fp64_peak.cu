// FP64 pipe micro-benchmark for B200 (sm_100a): DMMA.8x8x4 and DFMA issue-rate ceilings.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_peak fp64_peak.cu
// These numbers are the roofline denominators for the complex128 elimination kernels
// (MEASURED_PEAKS.json holds only bf16 + HBM).
#include <cstdio>
#include <cuda_runtime.h>

template <int NACC>
__global__ void __launch_bounds__(256) dmma_kernel(double* out, int iters, double seed) {
    double c[NACC][2];
#pragma unroll
    for (int i = 0; i < NACC; i++) { c[i][0] = seed * i; c[i][1] = seed; }
    double a = seed + threadIdx.x * 1e-9, b = seed - threadIdx.x * 1e-9;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < NACC; i++)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; i++) s += c[i][0] + c[i][1];
    if (s == 123.456) out[0] = s;
}

template <int NACC>
__global__ void __launch_bounds__(256) dfma_kernel(double* out, int iters, double seed) {
    double c[NACC];
#pragma unroll
    for (int i = 0; i < NACC; i++) c[i] = seed * i;
    double a = seed + threadIdx.x * 1e-9, b = seed - threadIdx.x * 1e-9;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < NACC; i++) c[i] = fma(a, c[i], b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; i++) s += c[i];
    if (s == 123.456) out[0] = s;
}

template <typename F>
float time_it(F f) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); cudaDeviceSynchronize();
    cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount;
    double* out; cudaMalloc(&out, 8);
    printf("{\"gpu\": \"%s\", \"sms\": %d,\n", p.name, sms);
    for (int cps = 1; cps <= 4; cps *= 2) {
        int iters = 20000;
        float ms = time_it([&] { dmma_kernel<8><<<sms * cps, 256>>>(out, iters, 1.0); });
        double flops = (double)sms * cps * 8 /*warps*/ * iters * 8 /*acc*/ * 512.0;
        printf(" \"dmma884_tflops_cta%d\": %.2f,\n", cps, flops / ms / 1e9);
    }
    for (int cps = 1; cps <= 4; cps *= 2) {
        int iters = 20000;
        float ms = time_it([&] { dfma_kernel<16><<<sms * cps, 256>>>(out, iters, 1.0); });
        double flops = (double)sms * cps * 256 * (double)iters * 16 * 2.0;
        printf(" \"dfma_tflops_cta%d\": %.2f,\n", cps, flops / ms / 1e9);
    }
    printf(" \"done\": 1}\n");
    return 0;
}
