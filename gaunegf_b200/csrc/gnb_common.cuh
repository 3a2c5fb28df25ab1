// Shared device helpers for the gaunegf_b200 kernels (complex128 as interleaved double2).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

typedef double2 cplx;

#ifndef GNB_NB
#define GNB_NB 32          // elimination block width (pivot block is NB x NB)
#endif
#define GNB_GROUP 256      // rows per tournament group (= threads per tournament CTA)

__host__ __device__ __forceinline__ cplx cmake(double r, double i) { return make_double2(r, i); }
__device__ __forceinline__ cplx cadd(cplx a, cplx b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ cplx csub(cplx a, cplx b) { return make_double2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ cplx cneg(cplx a) { return make_double2(-a.x, -a.y); }
__device__ __forceinline__ cplx cconj(cplx a) { return make_double2(a.x, -a.y); }
__device__ __forceinline__ cplx cmul(cplx a, cplx b) {
    return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
// a - b*c
__device__ __forceinline__ cplx cfnma(cplx a, cplx b, cplx c) {
    double re = fma(-b.x, c.x, a.x);
    re = fma(b.y, c.y, re);
    double im = fma(-b.x, c.y, a.y);
    im = fma(-b.y, c.x, im);
    return make_double2(re, im);
}
// a + b*c
__device__ __forceinline__ cplx cfma(cplx a, cplx b, cplx c) {
    double re = fma(b.x, c.x, a.x);
    re = fma(-b.y, c.y, re);
    double im = fma(b.x, c.y, a.y);
    im = fma(b.y, c.x, im);
    return make_double2(re, im);
}
// Smith's algorithm (the shape of LAPACK's zladiv) : a / b
__device__ __forceinline__ cplx cdiv(cplx a, cplx b) {
    if (fabs(b.x) >= fabs(b.y)) {
        double r = b.y / b.x, d = b.x + b.y * r;
        return make_double2((a.x + a.y * r) / d, (a.y - a.x * r) / d);
    } else {
        double r = b.x / b.y, d = b.y + b.x * r;
        return make_double2((a.x * r + a.y) / d, (a.y * r - a.x) / d);
    }
}
__device__ __forceinline__ double cabs1(cplx a) { return fabs(a.x) + fabs(a.y); }   // izamax metric
__device__ __forceinline__ double cabs2(cplx a) { return hypot(a.x, a.y); }         // numpy abs()

// D(8x8) += A(8x4) * B(4x8), FP64 tensor pipe (SASS DMMA.8x8x4 on sm_100a).
// lane = 4*gid + tig : A holds A[gid][tig], B holds B[tig][gid], C holds C[gid][2*tig + {0,1}].
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
