// Dev lab: batched complex128 rank-K update  C -= P * W  on the FP64 tensor pipe (DMMA.8x8x4), sm_100a.
// Explores CTA tile / stage depth / CTAs-per-SM for the elimination engine's dominant kernel.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o tools/zgemm_lab tools/zgemm_lab.cu
// Run  : tools/zgemm_lab [M=296] [N=1024]
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <vector>
#include <string>

typedef double2 cplx;

struct Args {
    cplx* C; long sC; int ldc;
    const cplx* P; long sP; int ldp;
    const cplx* W; long sW; int ldw;
    int rows, cols, K, nb;
};

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, int src_bytes) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" :: "r"(d), "l"(gsrc), "r"(src_bytes) : "memory");
}
template <int N> __device__ __forceinline__ void cp_wait() { asm volatile("cp.async.wait_group %0;" :: "n"(N) : "memory"); }
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }

template <int BM, int BN, int KC, int ST, int WM, int WN, int MINB, int M3>
__global__ void __launch_bounds__(WM * WN * 32, MINB) zg(Args g, int nti, int ntj, int total) {
    constexpr int NT = WM * WN * 32;
    constexpr int WTM = BM / WM, WTN = BN / WN, MI = WTM / 8, NI = WTN / 8;
    constexpr int PS = KC + ((KC % 8 == 4) ? 0 : (KC % 8 < 4 ? 4 - KC % 8 : 12 - KC % 8));   // == 4 mod 8
    constexpr int WS = BN + 2;                                                             // == 2 mod 8 (BN % 8 == 0)
    constexpr int STAGE = BM * PS + KC * WS;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cplx* sm = reinterpret_cast<cplx*>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int gid = lane >> 2, tig = lane & 3;
    const int wm = warp / WN, wn = warp % WN;
    const int per_mat = nti * ntj;
    const int nch = g.K / KC;
    const int my_tiles = (total - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    const int nitems = my_tiles * nch;

    auto issue = [&](int q) {
        if (q < nitems) {
            const int it = q / nch, ch = q - it * nch;
            const int tile = blockIdx.x + it * gridDim.x;
            const int b = tile / per_mat, rem = tile - b * per_mat;
            const int ti = rem / ntj, tj = rem - ti * ntj;
            const int i0 = ti * BM, j0 = tj * BN, k0 = ch * KC;
            const cplx* Pb = g.P + (long)b * g.sP + (long)i0 * g.ldp + k0;
            const cplx* Wb = g.W + (long)b * g.sW + (long)k0 * g.ldw + j0;
            cplx* Ps = sm + (q % ST) * STAGE;
            cplx* Ws = Ps + BM * PS;
#pragma unroll
            for (int x = 0; x < (BM * KC + NT - 1) / NT; x++) {
                const int idx = tid + x * NT;
                if ((BM * KC) % NT == 0 || idx < BM * KC) {
                    const int r = idx / KC, k = idx - r * KC;
                    cp_async16(&Ps[r * PS + k], Pb + (long)r * g.ldp + k, 16);
                }
            }
#pragma unroll
            for (int x = 0; x < (KC * BN + NT - 1) / NT; x++) {
                const int idx = tid + x * NT;
                if ((KC * BN) % NT == 0 || idx < KC * BN) {
                    const int k = idx / BN, n = idx - k * BN;
                    cp_async16(&Ws[k * WS + n], Wb + (long)k * g.ldw + n, 16);
                }
            }
        }
        cp_commit();
    };

    double cre[MI][NI][2], cim[MI][NI][2];
    double c3[M3 ? MI : 1][M3 ? NI : 1][2];
#pragma unroll
    for (int mi = 0; mi < MI; mi++)
#pragma unroll
        for (int ni = 0; ni < NI; ni++) {
            cre[mi][ni][0] = cre[mi][ni][1] = cim[mi][ni][0] = cim[mi][ni][1] = 0.0;
            if (M3) c3[mi][ni][0] = c3[mi][ni][1] = 0.0;
        }

#pragma unroll
    for (int s = 0; s < ST - 1; s++) issue(s);

    for (int q = 0; q < nitems; q++) {
        cp_wait<ST - 2>();
        __syncthreads();
        issue(q + ST - 1);
        const cplx* Ps = sm + (q % ST) * STAGE;
        const cplx* Ws = Ps + BM * PS;
#pragma unroll
        for (int kk = 0; kk < KC; kk += 4) {
            cplx af[MI], bf[NI];
#pragma unroll
            for (int mi = 0; mi < MI; mi++) af[mi] = Ps[(wm * WTM + mi * 8 + gid) * PS + kk + tig];
#pragma unroll
            for (int ni = 0; ni < NI; ni++) bf[ni] = Ws[(kk + tig) * WS + wn * WTN + ni * 8 + gid];
            if (M3) {
                // 3M: X = ar*br, Y = ai*bi, Z = (ar+ai)*(br+bi);  re = X - Y, im = Z - X - Y
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
        const int it = q / nch, ch = q - it * nch;
        if (ch == nch - 1) {
            const int tile = blockIdx.x + it * gridDim.x;
            const int b = tile / per_mat, rem = tile - b * per_mat;
            const int ti = rem / ntj, tj = rem - ti * ntj;
            cplx* Cb = g.C + (long)b * g.sC + (long)(ti * BM + wm * WTM + gid) * g.ldc + tj * BN + wn * WTN + tig * 2;
#pragma unroll
            for (int mi = 0; mi < MI; mi++) {
                cplx v[NI][2];
#pragma unroll
                for (int ni = 0; ni < NI; ni++) {
                    v[ni][0] = Cb[(long)(mi * 8) * g.ldc + ni * 8];
                    v[ni][1] = Cb[(long)(mi * 8) * g.ldc + ni * 8 + 1];
                }
#pragma unroll
                for (int ni = 0; ni < NI; ni++) {
#pragma unroll
                    for (int e = 0; e < 2; e++) {
                        double re, im;
                        if (M3) { re = cre[mi][ni][e] - cim[mi][ni][e]; im = c3[mi][ni][e] - cre[mi][ni][e] - cim[mi][ni][e]; }
                        else { re = cre[mi][ni][e]; im = cim[mi][ni][e]; }
                        v[ni][e].x -= re; v[ni][e].y -= im;
                        cre[mi][ni][e] = 0.0; cim[mi][ni][e] = 0.0;
                        if (M3) c3[mi][ni][e] = 0.0;
                    }
                    Cb[(long)(mi * 8) * g.ldc + ni * 8] = v[ni][0];
                    Cb[(long)(mi * 8) * g.ldc + ni * 8 + 1] = v[ni][1];
                }
            }
        }
    }
    cp_wait<0>();
}


// ------------------------------------------------------------------------------------------------
// Warp-specialised variant: one producer warp streams P / W row segments into an ST-deep shared-memory
// ring with cp.async.bulk (TMA, mbarrier complete_tx); NCW consumer warps run LDS + DMMA with no
// CTA-wide barrier (full/empty mbarriers per stage) and read-modify-write C in the epilogue.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* b, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* b, int bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* b) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* b, int parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t}" :: "r"(smem_u32(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, int bytes, unsigned long long* b) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(b)) : "memory");
}

template <int BM, int BN, int KC, int ST, int WM, int WN, int MINB, int M3>
__global__ void __launch_bounds__(WM * WN * 32 + 32, MINB) zgw(Args g, int nti, int ntj, int total) {
    constexpr int NCW = WM * WN;
    constexpr int WTM = BM / WM, WTN = BN / WN, MI = WTM / 8, NI = WTN / 8;
    constexpr int PS = KC + ((KC % 8 == 4) ? 0 : (KC % 8 < 4 ? 4 - KC % 8 : 12 - KC % 8));
    constexpr int WS = BN + 2;
    constexpr int STAGE = BM * PS + KC * WS;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cplx* sm = reinterpret_cast<cplx*>(smem_raw);
    unsigned long long* full = reinterpret_cast<unsigned long long*>(sm + ST * STAGE);
    unsigned long long* empty = full + ST;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int per_mat = nti * ntj;
    const int nch = g.K / KC;
    const int my_tiles = (total - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    if (tid == 0) {
        for (int s = 0; s < ST; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], NCW); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (warp == NCW) {
        // ---------------- producer warp ----------------
        int q = 0;
        for (int it = 0; it < my_tiles; it++) {
            const int tile = blockIdx.x + it * gridDim.x;
            const int b = tile / per_mat, rem = tile - b * per_mat;
            const int ti = rem / ntj, tj = rem - ti * ntj;
            const cplx* Pb = g.P + (long)b * g.sP + (long)(ti * BM) * g.ldp;
            const cplx* Wb = g.W + (long)b * g.sW + tj * BN;
            for (int ch = 0; ch < nch; ch++, q++) {
                const int s = q % ST;
                if (q >= ST) mbar_wait(&empty[s], ((q / ST) - 1) & 1);
                cplx* Ps = sm + s * STAGE;
                cplx* Ws = Ps + BM * PS;
                if (lane == 0) mbar_expect_tx(&full[s], (BM * KC + KC * BN) * 16);
                __syncwarp();
                const int k0 = ch * KC;
#pragma unroll
                for (int r = lane; r < BM; r += 32) bulk_g2s(&Ps[r * PS], Pb + (long)r * g.ldp + k0, KC * 16, &full[s]);
                if (lane < KC) bulk_g2s(&Ws[lane * WS], Wb + (long)(k0 + lane) * g.ldw, BN * 16, &full[s]);
            }
        }
        return;
    }
    // ---------------- consumer warps ----------------
    const int gid = lane >> 2, tig = lane & 3;
    const int wm = warp / WN, wn = warp % WN;
    double cre[MI][NI][2], cim[MI][NI][2];
    double c3[M3 ? MI : 1][M3 ? NI : 1][2];
#pragma unroll
    for (int mi = 0; mi < MI; mi++)
#pragma unroll
        for (int ni = 0; ni < NI; ni++) {
            cre[mi][ni][0] = cre[mi][ni][1] = cim[mi][ni][0] = cim[mi][ni][1] = 0.0;
            if (M3) c3[mi][ni][0] = c3[mi][ni][1] = 0.0;
        }
    int q = 0;
    for (int it = 0; it < my_tiles; it++) {
        for (int ch = 0; ch < nch; ch++, q++) {
            const int s = q % ST;
            mbar_wait(&full[s], (q / ST) & 1);
            const cplx* Ps = sm + s * STAGE;
            const cplx* Ws = Ps + BM * PS;
#pragma unroll
            for (int kk = 0; kk < KC; kk += 4) {
                cplx af[MI], bf[NI];
#pragma unroll
                for (int mi = 0; mi < MI; mi++) af[mi] = Ps[(wm * WTM + mi * 8 + gid) * PS + kk + tig];
#pragma unroll
                for (int ni = 0; ni < NI; ni++) bf[ni] = Ws[(kk + tig) * WS + wn * WTN + ni * 8 + gid];
                if (M3) {
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
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[s]);
        }
        {
            const int tile = blockIdx.x + it * gridDim.x;
            const int b = tile / per_mat, rem = tile - b * per_mat;
            const int ti = rem / ntj, tj = rem - ti * ntj;
            cplx* Cb = g.C + (long)b * g.sC + (long)(ti * BM + wm * WTM + gid) * g.ldc + tj * BN + wn * WTN + tig * 2;
#pragma unroll
            for (int mi = 0; mi < MI; mi++) {
                cplx v[NI][2];
#pragma unroll
                for (int ni = 0; ni < NI; ni++) {
                    v[ni][0] = Cb[(long)(mi * 8) * g.ldc + ni * 8];
                    v[ni][1] = Cb[(long)(mi * 8) * g.ldc + ni * 8 + 1];
                }
#pragma unroll
                for (int ni = 0; ni < NI; ni++) {
#pragma unroll
                    for (int e = 0; e < 2; e++) {
                        double re, im;
                        if (M3) { re = cre[mi][ni][e] - cim[mi][ni][e]; im = c3[mi][ni][e] - cre[mi][ni][e] - cim[mi][ni][e]; }
                        else { re = cre[mi][ni][e]; im = cim[mi][ni][e]; }
                        v[ni][e].x -= re; v[ni][e].y -= im;
                        cre[mi][ni][e] = 0.0; cim[mi][ni][e] = 0.0;
                        if (M3) c3[mi][ni][e] = 0.0;
                    }
                    Cb[(long)(mi * 8) * g.ldc + ni * 8] = v[ni][0];
                    Cb[(long)(mi * 8) * g.ldc + ni * 8 + 1] = v[ni][1];
                }
            }
        }
    }
}


// ------------------------------------------------------------------------------------------------
// zgc: warp-specialised like zgw, 1 CTA/SM, and the producer also prefetches the C tile of each CTA tile
// into shared memory (rows padded to CS cplx) so the epilogue never waits on global memory.
// ------------------------------------------------------------------------------------------------
template <int BM, int BN, int KC, int ST, int WM, int WN, int M3>
__global__ void __launch_bounds__(WM * WN * 32 + 32, 1) zgc(Args g, int nti, int ntj, int total) {
    constexpr int NCW = WM * WN;
    constexpr int WTM = BM / WM, WTN = BN / WN, MI = WTM / 8, NI = WTN / 8;
    constexpr int PS = KC + ((KC % 8 == 4) ? 0 : (KC % 8 < 4 ? 4 - KC % 8 : 12 - KC % 8));
    constexpr int WS = BN + 2;
    constexpr int CS = BN + 1;                       // 16 B mod 128 B row shift: conflict-free fragment reads
    constexpr int STAGE = BM * PS + KC * WS;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cplx* sm = reinterpret_cast<cplx*>(smem_raw);
    cplx* Cs = sm + ST * STAGE;
    unsigned long long* full = reinterpret_cast<unsigned long long*>(Cs + BM * CS);
    unsigned long long* empty = full + ST;
    unsigned long long* cfull = empty + ST;
    unsigned long long* cempty = cfull + 1;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int per_mat = nti * ntj;
    const int nch = g.K / KC;
    const int my_tiles = (total - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    if (tid == 0) {
        for (int s = 0; s < ST; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], NCW); }
        mbar_init(cfull, 1); mbar_init(cempty, NCW);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (warp == NCW) {
        int q = 0;
        for (int it = 0; it < my_tiles; it++) {
            const int tile = blockIdx.x + it * gridDim.x;
            const int b = tile / per_mat, rem = tile - b * per_mat;
            const int ti = rem / ntj, tj = rem - ti * ntj;
            const cplx* Pb = g.P + (long)b * g.sP + (long)(ti * BM) * g.ldp;
            const cplx* Wb = g.W + (long)b * g.sW + tj * BN;
            const cplx* Cb = g.C + (long)b * g.sC + (long)(ti * BM) * g.ldc + tj * BN;
            for (int ch = 0; ch < nch; ch++, q++) {
                const int s = q % ST;
                if (q >= ST) mbar_wait(&empty[s], ((q / ST) - 1) & 1);
                cplx* Ps = sm + s * STAGE;
                cplx* Ws = Ps + BM * PS;
                if (lane == 0) mbar_expect_tx(&full[s], (BM * KC + KC * BN) * 16);
                __syncwarp();
                const int k0 = ch * KC;
#pragma unroll
                for (int r = lane; r < BM; r += 32) bulk_g2s(&Ps[r * PS], Pb + (long)r * g.ldp + k0, KC * 16, &full[s]);
                if (lane < KC) bulk_g2s(&Ws[lane * WS], Wb + (long)(k0 + lane) * g.ldw, BN * 16, &full[s]);
                if (ch == 0) {
                    // C tile of this CTA tile (the previous tile's epilogue must have drained the buffer)
                    if (it > 0) mbar_wait(cempty, (it - 1) & 1);
                    if (lane == 0) mbar_expect_tx(cfull, BM * BN * 16);
                    __syncwarp();
#pragma unroll
                    for (int r = lane; r < BM; r += 32) bulk_g2s(&Cs[r * CS], Cb + (long)r * g.ldc, BN * 16, cfull);
                }
            }
        }
        return;
    }
    const int gid = lane >> 2, tig = lane & 3;
    const int wm = warp / WN, wn = warp % WN;
    double cre[MI][NI][2], cim[MI][NI][2];
    double c3[M3 ? MI : 1][M3 ? NI : 1][2];
#pragma unroll
    for (int mi = 0; mi < MI; mi++)
#pragma unroll
        for (int ni = 0; ni < NI; ni++) {
            cre[mi][ni][0] = cre[mi][ni][1] = cim[mi][ni][0] = cim[mi][ni][1] = 0.0;
            if (M3) c3[mi][ni][0] = c3[mi][ni][1] = 0.0;
        }
    int q = 0;
    for (int it = 0; it < my_tiles; it++) {
        for (int ch = 0; ch < nch; ch++, q++) {
            const int s = q % ST;
            mbar_wait(&full[s], (q / ST) & 1);
            const cplx* Ps = sm + s * STAGE;
            const cplx* Ws = Ps + BM * PS;
#pragma unroll
            for (int kk = 0; kk < KC; kk += 4) {
                cplx af[MI], bf[NI];
#pragma unroll
                for (int mi = 0; mi < MI; mi++) af[mi] = Ps[(wm * WTM + mi * 8 + gid) * PS + kk + tig];
#pragma unroll
                for (int ni = 0; ni < NI; ni++) bf[ni] = Ws[(kk + tig) * WS + wn * WTN + ni * 8 + gid];
                if (M3) {
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
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[s]);
        }
        {
            const int tile = blockIdx.x + it * gridDim.x;
            const int b = tile / per_mat, rem = tile - b * per_mat;
            const int ti = rem / ntj, tj = rem - ti * ntj;
            cplx* Cb = g.C + (long)b * g.sC + (long)(ti * BM + wm * WTM + gid) * g.ldc + tj * BN + wn * WTN + tig * 2;
            const cplx* Cl = Cs + (wm * WTM + gid) * CS + wn * WTN + tig * 2;
            mbar_wait(cfull, it & 1);
#pragma unroll
            for (int mi = 0; mi < MI; mi++) {
#pragma unroll
                for (int ni = 0; ni < NI; ni++) {
                    cplx v[2];
                    v[0] = Cl[(mi * 8) * CS + ni * 8];
                    v[1] = Cl[(mi * 8) * CS + ni * 8 + 1];
#pragma unroll
                    for (int e = 0; e < 2; e++) {
                        double re, im;
                        if (M3) { re = cre[mi][ni][e] - cim[mi][ni][e]; im = c3[mi][ni][e] - cre[mi][ni][e] - cim[mi][ni][e]; }
                        else { re = cre[mi][ni][e]; im = cim[mi][ni][e]; }
                        v[e].x -= re; v[e].y -= im;
                        cre[mi][ni][e] = 0.0; cim[mi][ni][e] = 0.0;
                        if (M3) c3[mi][ni][e] = 0.0;
                    }
                    Cb[(long)(mi * 8) * g.ldc + ni * 8] = v[0];
                    Cb[(long)(mi * 8) * g.ldc + ni * 8 + 1] = v[1];
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(cempty);
        }
    }
}


// ------------------------------------------------------------------------------------------------
// zgp: operands pre-packed in global memory as shared-memory images so that ONE bulk copy per operand
// per K chunk feeds the ring:
//   Ppk[b][kc][rb][32][PPS]  (kc = K chunk of 16, rb = row block of 32; a 64-row tile = 2 adjacent blocks)
//   Wpk[b][kc][cb][16][WPS]  (cb = column block of 32)
// CPRE = 1: 1 CTA/SM, the producer also prefetches the C tile into smem; CPRE = 0: C read in the epilogue.
// ------------------------------------------------------------------------------------------------
#define PPS 20
#define WPS 34
struct ArgsP {
    cplx* C; long sC; int ldc;
    const cplx* Ppk; long sP; int nrb;      // nrb = row blocks per chunk
    const cplx* Wpk; long sW; int ncb;      // ncb = column blocks per chunk
    int rows, cols, K, nb;
};
__global__ void k_pack_p(const cplx* P, long sP, int ldp, int rows, int K, cplx* Ppk, long sPk) {
    const int b = blockIdx.y;
    const int nrb = rows / 32;
    for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < (long)rows * K; idx += (long)gridDim.x * blockDim.x) {
        const int r = (int)(idx / K), k = (int)(idx % K);
        Ppk[(long)b * sPk + ((long)(k / 16) * nrb + r / 32) * (32 * PPS) + (r % 32) * PPS + k % 16] = P[(long)b * sP + (long)r * ldp + k];
    }
}
__global__ void k_pack_w(const cplx* W, long sW, int ldw, int cols, int K, cplx* Wpk, long sWk) {
    const int b = blockIdx.y;
    const int ncb = cols / 32;
    for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < (long)cols * K; idx += (long)gridDim.x * blockDim.x) {
        const int k = (int)(idx / cols), c = (int)(idx % cols);
        Wpk[(long)b * sWk + ((long)(k / 16) * ncb + c / 32) * (16 * WPS) + (k % 16) * WPS + c % 32] = W[(long)b * sW + (long)k * ldw + c];
    }
}

template <int ST, int M3, int CPRE, int MINB, int FENCE>
__global__ void __launch_bounds__(288, MINB) zgp(ArgsP g, int nti, int ntj, int total) {
    constexpr int BM = 64, BN = 64, KC = 16, WM = 4, WN = 2, NCW = 8, MI = 2, NI = 4;
    constexpr int PST = 2 * 32 * PPS, WST = 2 * 16 * WPS, STAGE = PST + WST;
    constexpr int CS = BN + 1;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cplx* sm = reinterpret_cast<cplx*>(smem_raw);
    cplx* Cs = sm + ST * STAGE;
    unsigned long long* full = reinterpret_cast<unsigned long long*>(Cs + (CPRE == 1 ? BM * CS : 0));
    unsigned long long* empty = full + ST;
    unsigned long long* cfull = empty + ST;
    unsigned long long* cempty = cfull + 1;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int per_mat = nti * ntj;
    const int nch = g.K / KC;
    const int my_tiles = (total - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    if (tid == 0) {
        for (int s = 0; s < ST; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], NCW); }
        mbar_init(cfull, 1); mbar_init(cempty, NCW);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (warp == NCW) {
        int q = 0;
        for (int it = 0; it < my_tiles; it++) {
            const int tile = blockIdx.x + it * gridDim.x;
            const int b = tile / per_mat, rem = tile - b * per_mat;
            const int ti = rem / ntj, tj = rem - ti * ntj;
            const cplx* Pb = g.Ppk + (long)b * g.sP + (long)(ti * 2) * (32 * PPS);
            const cplx* Wb = g.Wpk + (long)b * g.sW + (long)(tj * 2) * (16 * WPS);
            const cplx* Cb = g.C + (long)b * g.sC + (long)(ti * BM) * g.ldc + tj * BN;
            for (int ch = 0; ch < nch; ch++, q++) {
                const int s = q % ST;
                if (q >= ST) mbar_wait(&empty[s], ((q / ST) - 1) & 1);
                cplx* Ps = sm + s * STAGE;
                cplx* Ws = Ps + PST;
                if (lane == 0) {
                    mbar_expect_tx(&full[s], STAGE * 16);
                    bulk_g2s(Ps, Pb + (long)ch * g.nrb * (32 * PPS), PST * 16, &full[s]);
                    bulk_g2s(Ws, Wb + (long)ch * g.ncb * (16 * WPS), WST * 16, &full[s]);
                }
                if (CPRE == 1 && ch == 0) {
                    if (it > 0) mbar_wait(cempty, (it - 1) & 1);
                    if (lane == 0) mbar_expect_tx(cfull, BM * BN * 16);
                    __syncwarp();
#pragma unroll
                    for (int r = lane; r < BM; r += 32) bulk_g2s(&Cs[r * CS], Cb + (long)r * g.ldc, BN * 16, cfull);
                }
            }
        }
        return;
    }
    const int gid = lane >> 2, tig = lane & 3;
    const int wm = warp / WN, wn = warp % WN;
    double cre[MI][NI][2], cim[MI][NI][2];
    double c3[M3 ? MI : 1][M3 ? NI : 1][2];
#pragma unroll
    for (int mi = 0; mi < MI; mi++)
#pragma unroll
        for (int ni = 0; ni < NI; ni++) {
            cre[mi][ni][0] = cre[mi][ni][1] = cim[mi][ni][0] = cim[mi][ni][1] = 0.0;
            if (M3) c3[mi][ni][0] = c3[mi][ni][1] = 0.0;
        }
    cplx pv[CPRE >= 2 ? (CPRE == 3 ? 1 : MI) : 1][NI][2];
    int q = 0;
    for (int it = 0; it < my_tiles; it++) {
        for (int ch = 0; ch < nch; ch++, q++) {
            const int s = q % ST;
            if (CPRE >= 2 && ch == nch - 1) {
                const int tile = blockIdx.x + it * gridDim.x;
                const int b = tile / per_mat, rem = tile - b * per_mat;
                const int ti = rem / ntj, tj = rem - ti * ntj;
                const cplx* Cb = g.C + (long)b * g.sC + (long)(ti * BM + wm * 16 + gid) * g.ldc + tj * BN + wn * 32 + tig * 2;
#pragma unroll
                for (int mi = 0; mi < (CPRE == 3 ? 1 : MI); mi++)
#pragma unroll
                    for (int ni = 0; ni < NI; ni++) {
                        pv[mi][ni][0] = Cb[(long)(mi * 8) * g.ldc + ni * 8];
                        pv[mi][ni][1] = Cb[(long)(mi * 8) * g.ldc + ni * 8 + 1];
                    }
            }
            mbar_wait(&full[s], (q / ST) & 1);
            const cplx* Ps = sm + s * STAGE + (wm * 16 + gid) * PPS + tig;            // rows 0..63 contiguous (2 blocks of 32)
            const cplx* Ws = sm + s * STAGE + PST + wn * (16 * WPS) + tig * WPS + gid; // column block wn
#pragma unroll
            for (int kk = 0; kk < KC; kk += 4) {
                cplx af[MI], bf[NI];
#pragma unroll
                for (int mi = 0; mi < MI; mi++) af[mi] = Ps[(mi * 8) * PPS + kk];
#pragma unroll
                for (int ni = 0; ni < NI; ni++) bf[ni] = Ws[kk * WPS + ni * 8];
                if (M3) {
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
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[s]);
        }
        {
            const int tile = blockIdx.x + it * gridDim.x;
            const int b = tile / per_mat, rem = tile - b * per_mat;
            const int ti = rem / ntj, tj = rem - ti * ntj;
            cplx* Cb = g.C + (long)b * g.sC + (long)(ti * BM + wm * 16 + gid) * g.ldc + tj * BN + wn * 32 + tig * 2;
            const cplx* Cl = Cs + (wm * 16 + gid) * CS + wn * 32 + tig * 2;
            if (CPRE == 1) mbar_wait(cfull, it & 1);
#pragma unroll
            for (int mi = 0; mi < MI; mi++) {
                cplx v[NI][2];
#pragma unroll
                for (int ni = 0; ni < NI; ni++) {
                    if (CPRE == 1) { v[ni][0] = Cl[(mi * 8) * CS + ni * 8]; v[ni][1] = Cl[(mi * 8) * CS + ni * 8 + 1]; }
                    else if (CPRE == 2 || (CPRE == 3 && mi == 0)) { v[ni][0] = pv[mi][ni][0]; v[ni][1] = pv[mi][ni][1]; }
                    else { v[ni][0] = Cb[(long)(mi * 8) * g.ldc + ni * 8]; v[ni][1] = Cb[(long)(mi * 8) * g.ldc + ni * 8 + 1]; }
                }
#pragma unroll
                for (int ni = 0; ni < NI; ni++) {
#pragma unroll
                    for (int e = 0; e < 2; e++) {
                        double re, im;
                        if (M3) { re = cre[mi][ni][e] - cim[mi][ni][e]; im = c3[mi][ni][e] - cre[mi][ni][e] - cim[mi][ni][e]; }
                        else { re = cre[mi][ni][e]; im = cim[mi][ni][e]; }
                        v[ni][e].x -= re; v[ni][e].y -= im;
                        cre[mi][ni][e] = 0.0; cim[mi][ni][e] = 0.0;
                        if (M3) c3[mi][ni][e] = 0.0;
                    }
                    Cb[(long)(mi * 8) * g.ldc + ni * 8] = v[ni][0];
                    Cb[(long)(mi * 8) * g.ldc + ni * 8 + 1] = v[ni][1];
                }
            }
            if (CPRE == 1) { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); __syncwarp(); if (lane == 0) mbar_arrive(cempty); }
        }
    }
}

__global__ void k_fill(cplx* p, long n, unsigned seed) {
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        unsigned long long h = (i + 1) * 0x9E3779B97F4A7C15ull + seed * 0xD1B54A32D192ED03ull;
        h ^= h >> 29; h *= 0xBF58476D1CE4E5B9ull; h ^= h >> 32;
        p[i] = make_double2(((h & 0xffff) / 65536.0 - 0.5), (((h >> 16) & 0xffff) / 65536.0 - 0.5));
    }
}
__global__ void k_ref(Args g, const cplx* C0, int b, int nsamp, double* maxerr, double* maxref) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nsamp) return;
    const int r = (s * 7919) % g.rows, c = (s * 104729 + 13) % g.cols;
    cplx acc = C0[(long)b * g.sC + (long)r * g.ldc + c];
    for (int k = 0; k < g.K; k++) {
        const cplx p = g.P[(long)b * g.sP + (long)r * g.ldp + k], w = g.W[(long)b * g.sW + (long)k * g.ldw + c];
        acc.x -= p.x * w.x - p.y * w.y; acc.y -= p.x * w.y + p.y * w.x;
    }
    const cplx got = g.C[(long)b * g.sC + (long)r * g.ldc + c];
    const double e = fmax(fabs(got.x - acc.x), fabs(got.y - acc.y));
    atomicMax((unsigned long long*)maxerr, __double_as_longlong(e));
    atomicMax((unsigned long long*)maxref, __double_as_longlong(fabs(acc.x)));
}

static int g_sms = 148;
static cplx *dC, *dC0, *dP, *dW;
static double* dErr;

template <int BM, int BN, int KC, int ST, int WM, int WN, int MINB, int M3, int SPEC = 0>
void run(const char* name, int M, int rows, int cols, int K, int ldc) {
    constexpr int PS = KC + ((KC % 8 == 4) ? 0 : (KC % 8 < 4 ? 4 - KC % 8 : 12 - KC % 8));
    constexpr int WS = BN + 2;
    const size_t smem = (size_t)ST * (BM * PS + KC * WS) * sizeof(cplx) + (SPEC ? 2 * ST * 8 : 0);
    const int nthr = WM * WN * 32 + (SPEC ? 32 : 0);
    if (rows % BM || cols % BN || K % KC) { printf("%-34s skip shape\n", name); return; }
    auto kern = SPEC ? zgw<BM, BN, KC, ST, WM, WN, MINB, M3> : zg<BM, BN, KC, ST, WM, WN, MINB, M3>;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
        printf("%-34s smem %zu too large\n", name, smem); cudaGetLastError(); return;
    }
    int occ = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, nthr, smem);
    Args g{dC, (long)rows * ldc, ldc, dP, (long)rows * K, K, dW, (long)K * cols, cols, rows, cols, K, M};
    const int nti = rows / BM, ntj = cols / BN;
    const long total = (long)M * nti * ntj;
    const int grid = (int)std::min<long>(total, (long)occ * g_sms);
    cudaMemcpy(dC, dC0, (size_t)M * rows * ldc * sizeof(cplx), cudaMemcpyDeviceToDevice);
    kern<<<grid, nthr, smem>>>(g, nti, ntj, (int)total);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%-34s ERROR %s\n", name, cudaGetErrorString(e)); exit(1); }
    cudaMemset(dErr, 0, 16);
    k_ref<<<8, 128>>>(g, dC0, 0, 1024, dErr, dErr + 1);
    k_ref<<<8, 128>>>(g, dC0, M - 1, 1024, dErr, dErr + 1);
    double herr[2];
    cudaMemcpy(herr, dErr, 16, cudaMemcpyDeviceToHost);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 5;
    cudaEventRecord(e0);
    for (int i = 0; i < iters; i++) kern<<<grid, nthr, smem>>>(g, nti, ntj, (int)total);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= iters;
    const double tf = 8.0 * rows * cols * (double)K * M / (ms * 1e-3) / 1e12;
    printf("%-34s K=%4d %4dx%4d occ=%d smem=%6zu  %8.3f ms  %6.2f TF/s  relerr %.1e\n", name, K, rows, cols, occ, smem, ms, tf,
           herr[0] / (herr[1] > 0 ? herr[1] : 1));
    fflush(stdout);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
}

template <int BM, int BN, int KC, int ST, int WM, int WN, int M3>
void runc(const char* name, int M, int rows, int cols, int K, int ldc) {
    constexpr int PS = KC + ((KC % 8 == 4) ? 0 : (KC % 8 < 4 ? 4 - KC % 8 : 12 - KC % 8));
    constexpr int WS = BN + 2;
    const size_t smem = (size_t)(ST * (BM * PS + KC * WS) + BM * (BN + 1)) * sizeof(cplx) + (2 * ST + 2) * 8;
    const int nthr = WM * WN * 32 + 32;
    if (rows % BM || cols % BN || K % KC) { printf("%-34s skip shape\n", name); return; }
    auto kern = zgc<BM, BN, KC, ST, WM, WN, M3>;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
        printf("%-34s smem %zu too large\n", name, smem); cudaGetLastError(); return;
    }
    Args g{dC, (long)rows * ldc, ldc, dP, (long)rows * K, K, dW, (long)K * cols, cols, rows, cols, K, M};
    const int nti = rows / BM, ntj = cols / BN;
    const long total = (long)M * nti * ntj;
    const int grid = (int)std::min<long>(total, (long)g_sms);
    cudaMemcpy(dC, dC0, (size_t)M * rows * ldc * sizeof(cplx), cudaMemcpyDeviceToDevice);
    kern<<<grid, nthr, smem>>>(g, nti, ntj, (int)total);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%-34s ERROR %s\n", name, cudaGetErrorString(e)); exit(1); }
    cudaMemset(dErr, 0, 16);
    k_ref<<<8, 128>>>(g, dC0, 0, 1024, dErr, dErr + 1);
    k_ref<<<8, 128>>>(g, dC0, M - 1, 1024, dErr, dErr + 1);
    double herr[2];
    cudaMemcpy(herr, dErr, 16, cudaMemcpyDeviceToHost);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 5;
    cudaEventRecord(e0);
    for (int i = 0; i < iters; i++) kern<<<grid, nthr, smem>>>(g, nti, ntj, (int)total);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= iters;
    const double tf = 8.0 * rows * cols * (double)K * M / (ms * 1e-3) / 1e12;
    printf("%-34s K=%4d %4dx%4d occ=1 smem=%6zu  %8.3f ms  %6.2f TF/s  relerr %.1e\n", name, K, rows, cols, smem, ms, tf,
           herr[0] / (herr[1] > 0 ? herr[1] : 1));
    fflush(stdout);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
}

static cplx *dPpk, *dWpk;
__global__ void k_tile_err(Args g, const cplx* C0, int b, double* terr) {
    // one block per 64x64 tile of matrix b: max abs error of the tile
    const int ti = blockIdx.y, tj = blockIdx.x;
    double m = 0;
    for (int e = threadIdx.x; e < 64 * 64; e += blockDim.x) {
        const int r = ti * 64 + e / 64, c = tj * 64 + e % 64;
        cplx acc = C0[(long)b * g.sC + (long)r * g.ldc + c];
        for (int k = 0; k < g.K; k++) {
            const cplx p = g.P[(long)b * g.sP + (long)r * g.ldp + k], w = g.W[(long)b * g.sW + (long)k * g.ldw + c];
            acc.x -= p.x * w.x - p.y * w.y; acc.y -= p.x * w.y + p.y * w.x;
        }
        const cplx got = g.C[(long)b * g.sC + (long)r * g.ldc + c];
        m = fmax(m, fmax(fabs(got.x - acc.x), fabs(got.y - acc.y)));
    }
    __shared__ double sh[256];
    sh[threadIdx.x] = m; __syncthreads();
    for (int o = 128; o > 0; o >>= 1) { if (threadIdx.x < o) sh[threadIdx.x] = fmax(sh[threadIdx.x], sh[threadIdx.x + o]); __syncthreads(); }
    if (threadIdx.x == 0) terr[ti * gridDim.x + tj] = sh[0];
}
static int g_diag = 0;
template <int ST, int M3, int CPRE, int MINB = (CPRE == 1 ? 1 : 2), int FENCE = 0>
void runp(const char* name, int M, int rows, int cols, int K, int ldc) {
    constexpr int STAGE = 2 * 32 * PPS + 2 * 16 * WPS;
    const size_t smem = (size_t)(ST * STAGE + (CPRE == 1 ? 64 * 65 : 0)) * sizeof(cplx) + (2 * ST + 2) * 8;
    const int nthr = 288;
    if (rows % 64 || cols % 64 || K % 16) { printf("%-34s skip shape\n", name); return; }
    auto kern = zgp<ST, M3, CPRE, MINB, FENCE>;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
        printf("%-34s smem %zu too large\n", name, smem); cudaGetLastError(); return;
    }
    int occ = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, nthr, smem);
    const long sPk = (long)(K / 16) * (rows / 32) * 32 * PPS, sWk = (long)(K / 16) * (cols / 32) * 16 * WPS;
    k_pack_p<<<dim3(256, M), 256>>>(dP, (long)rows * K, K, rows, K, dPpk, sPk);
    k_pack_w<<<dim3(256, M), 256>>>(dW, (long)K * cols, cols, cols, K, dWpk, sWk);
    ArgsP gp{dC, (long)rows * ldc, ldc, dPpk, sPk, rows / 32, dWpk, sWk, cols / 32, rows, cols, K, M};
    Args g{dC, (long)rows * ldc, ldc, dP, (long)rows * K, K, dW, (long)K * cols, cols, rows, cols, K, M};
    const int nti = rows / 64, ntj = cols / 64;
    const long total = (long)M * nti * ntj;
    const int grid = (int)std::min<long>(total, (long)occ * g_sms);
    cudaMemcpy(dC, dC0, (size_t)M * rows * ldc * sizeof(cplx), cudaMemcpyDeviceToDevice);
    kern<<<grid, nthr, smem>>>(gp, nti, ntj, (int)total);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%-34s ERROR %s\n", name, cudaGetErrorString(e)); exit(1); }
    cudaMemset(dErr, 0, 16);
    k_ref<<<8, 128>>>(g, dC0, 0, 1024, dErr, dErr + 1);
    k_ref<<<8, 128>>>(g, dC0, M - 1, 1024, dErr, dErr + 1);
    double herr[2];
    cudaMemcpy(herr, dErr, 16, cudaMemcpyDeviceToHost);
    if (g_diag && herr[0] > 0) {
        double* dT; cudaMalloc(&dT, nti * ntj * 8);
        std::vector<double> ht(nti * ntj);
        for (int b : {0, 1, M / 2, M - 1}) {
            k_tile_err<<<dim3(ntj, nti), 256>>>(g, dC0, b, dT);
            cudaMemcpy(ht.data(), dT, nti * ntj * 8, cudaMemcpyDeviceToHost);
            int bad = 0;
            for (double x : ht) bad += x > 0;
            printf("  batch %d: %d bad tiles of %d:", b, bad, nti * ntj);
            int shown = 0;
            for (int t = 0; t < nti * ntj && shown < 24; t++) if (ht[t] > 0) { printf(" (%d,%d)", t / ntj, t % ntj); shown++; }
            printf("\n");
        }
        cudaFree(dT);
    }
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 5;
    cudaEventRecord(e0);
    for (int i = 0; i < iters; i++) kern<<<grid, nthr, smem>>>(gp, nti, ntj, (int)total);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= iters;
    const double tf = 8.0 * rows * cols * (double)K * M / (ms * 1e-3) / 1e12;
    printf("%-34s K=%4d %4dx%4d occ=%d smem=%6zu  %8.3f ms  %6.2f TF/s  relerr %.1e\n", name, K, rows, cols, occ, smem, ms, tf,
           herr[0] / (herr[1] > 0 ? herr[1] : 1));
    fflush(stdout);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
}

int main(int argc, char** argv) {
    const int M = argc > 1 ? atoi(argv[1]) : 296;
    const int N = argc > 2 ? atoi(argv[2]) : 1024;
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0); g_sms = p.multiProcessorCount;
    const int KMAX = 512;
    const size_t nC = (size_t)M * N * N, nP = (size_t)M * N * KMAX;
    cudaMalloc(&dC, nC * 16); cudaMalloc(&dC0, nC * 16); cudaMalloc(&dP, nP * 16); cudaMalloc(&dW, nP * 16);
    cudaMalloc(&dErr, 16);
    cudaMalloc(&dPpk, nP * 16 * 5 / 4); cudaMalloc(&dWpk, nP * 16 * 17 / 16);
    k_fill<<<1024, 256>>>(dC0, nC, 1); k_fill<<<1024, 256>>>(dP, nP, 2); k_fill<<<1024, 256>>>(dW, nP, 3);
    cudaDeviceSynchronize();
    printf("zgemm_lab M=%d N=%d sms=%d\n", M, N, g_sms);
    const int Ksel = argc > 3 ? atoi(argv[3]) : 0;      // 0 = sweep
    const int vsel = argc > 4 ? atoi(argv[4]) : -1;     // -1 = all variants
    g_diag = argc > 5 ? atoi(argv[5]) : 0;
    for (int K : {32, 64, 128, 256, 512}) {
        if (Ksel && K != Ksel) continue;
        int v = 0;
#define V(...) if (vsel < 0 || vsel == v++) run<__VA_ARGS__>
        //   BM  BN  KC ST WM WN MINB M3
        V(64, 64, 16, 3, 4, 2, 2, 0)("v0 64x64 kc16 st3 8w x2", M, N, N, K, N);
        V(64, 64, 16, 2, 4, 2, 2, 0)("v1 64x64 kc16 st2 8w x2", M, N, N, K, N);
        V(32, 64, 16, 3, 2, 2, 4, 0)("v2 32x64 kc16 st3 4w x4", M, N, N, K, N);
        V(128, 64, 16, 3, 8, 2, 1, 0)("v3 128x64 kc16 st3 16w(8x2) x1", M, N, N, K, N);
        V(64, 64, 16, 3, 4, 2, 2, 1)("v4 64x64 kc16 st3 8w x2 3M", M, N, N, K, N);
        V(128, 64, 16, 3, 4, 4, 1, 1)("v5 128x64 kc16 st3 16w x1 3M", M, N, N, K, N);
        V(64, 64, 16, 3, 4, 2, 2, 0, 1)("v6 WS 64x64 kc16 st3 8w x2", M, N, N, K, N);
        V(64, 64, 16, 3, 4, 2, 2, 1, 1)("v7 WS 64x64 kc16 st3 8w x2 3M", M, N, N, K, N);
        V(64, 64, 8, 5, 4, 2, 2, 0, 1)("v8 WS 64x64 kc8 st5 8w x2", M, N, N, K, N);
        V(128, 64, 16, 3, 8, 2, 1, 0, 1)("v9 WS 128x64 kc16 st3 16w x1", M, N, N, K, N);
        V(128, 64, 16, 3, 8, 2, 1, 1, 1)("v10 WS 128x64 kc16 st3 16w x1 3M", M, N, N, K, N);
        V(64, 32, 16, 4, 4, 1, 4, 0, 1)("v11 WS 64x32 kc16 st4 4w x4", M, N, N, K, N);
#define VC(...) if (vsel < 0 || vsel == v++) runc<__VA_ARGS__>
        VC(64, 64, 16, 3, 4, 2, 0)("v12 WSC 64x64 kc16 st3 8w", M, N, N, K, N);
        VC(64, 64, 16, 3, 4, 2, 1)("v13 WSC 64x64 kc16 st3 8w 3M", M, N, N, K, N);
        VC(64, 64, 8, 6, 4, 2, 0)("v14 WSC 64x64 kc8 st6 8w", M, N, N, K, N);
        VC(64, 64, 16, 3, 2, 4, 0)("v15 WSC 64x64 kc16 st3 8w(2x4)", M, N, N, K, N);
        VC(64, 64, 16, 3, 4, 4, 0)("v16 WSC 64x64 kc16 st3 16w", M, N, N, K, N);
        VC(64, 64, 16, 3, 4, 4, 1)("v17 WSC 64x64 kc16 st3 16w 3M", M, N, N, K, N);
        VC(128, 32, 16, 3, 8, 1, 0)("v18 WSC 128x32 kc16 st3 8w", M, N, N, K, N);
        VC(96, 64, 16, 2, 6, 2, 0)("v19 WSC 96x64 kc16 st2 12w", M, 960, N, K, N);
#define VP(...) if (vsel < 0 || vsel == v++) runp<__VA_ARGS__>
        VP(3, 0, 1)("v20 PK st3 cpre", M, N, N, K, N);
        VP(4, 0, 1)("v21 PK st4 cpre", M, N, N, K, N);
        VP(3, 1, 1)("v22 PK st3 cpre 3M", M, N, N, K, N);
        VP(4, 1, 1)("v23 PK st4 cpre 3M", M, N, N, K, N);
        VP(3, 0, 0)("v24 PK st3 x2 rmw", M, N, N, K, N);
        VP(2, 0, 0)("v25 PK st2 x2 rmw", M, N, N, K, N);
        VP(3, 0, 0, 1, 0)("v26 PK st3 x1 rmw", M, N, N, K, N);
        VP(3, 0, 0, 2, 1)("v27 PK st3 x2 rmw fence", M, N, N, K, N);
        VP(4, 0, 0, 1, 0)("v28 PK st4 x1 rmw", M, N, N, K, N);
        VP(5, 0, 0, 1, 0)("v29 PK st5 x1 rmw", M, N, N, K, N);
        VP(4, 0, 2, 1, 0)("v30 PK st4 x1 early", M, N, N, K, N);
        VP(5, 0, 2, 1, 0)("v31 PK st5 x1 early", M, N, N, K, N);
        VP(4, 1, 0, 1, 0)("v32 PK st4 x1 rmw 3M", M, N, N, K, N);
        VP(5, 1, 0, 1, 0)("v33 PK st5 x1 rmw 3M", M, N, N, K, N);
        VP(4, 1, 3, 1, 0)("v34 PK st4 x1 early-half 3M", M, N, N, K, N);
        VP(4, 1, 2, 1, 0)("v35 PK st4 x1 early 3M", M, N, N, K, N);
    }
    return 0;
}
