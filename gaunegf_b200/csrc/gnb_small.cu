// Small-orbital-count path: ONE CTA per energy point, the whole matrix in shared memory.
//
// For N <= GNB_SMALL_MAX_N (119: N*(N|1)*16 B + bookkeeping fits the 227 KB of one sm_100a CTA) the kernel
// assembles A = E S - F - Sigma0 - Sigma_k(E) in shared memory, inverts it in place with a partially pivoted
// Gauss-Jordan elimination (every warp finds the pivot of a column redundantly with warp shuffles, so there is no
// broadcast step; two CTA barriers per column), and reduces G = A^-1 on chip:
//   mode GREEN : G written out (consumers that need the whole matrix, utils.inv / integrate.py:67-71)
//   mode DOS   : -Im diag(G)/pi per orbital and its sum (transport.py:183-190)
//   mode T     : Re Tr[Gamma1 G Gamma2 G^H] from the contact blocks of G only (transport.py:150-157); G never leaves
//                the SM.
// Replaces assemble + tournament-pivoted block elimination + reduction launches of the lock-step engines, whose
// per-launch parallelism (energies x 32-column blocks) is too thin at these sizes (BASELINE cfg 1: N = 64).
#include "gnb_common.cuh"
#include "gnb_kernels.h"

namespace {

constexpr double kInvPi = 0.31830988618379067154;

template <int NT>
__global__ void __launch_bounds__(NT) k_small_gj(const GnbSmallArgs a) {
    extern __shared__ __align__(16) unsigned char sm_raw[];
    const int N = a.N, ld = N | 1;             // odd row stride: column walks hit 8 distinct 16-byte bank groups
    cplx* Am = reinterpret_cast<cplx*>(sm_raw);
    int* rowsrc = reinterpret_cast<int*>(Am + (size_t)N * ld);
    int* q = rowsrc + N;
    double* red = reinterpret_cast<double*>(q + N);                // [NT/32] (2N ints: 8-byte aligned)
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    constexpr int NW = NT / 32;
    const int e = blockIdx.x;

    // ---- assemble (integrate.py:70, transport.py:152): coalesced row reads of F, S (L2 resident)
    if (a.Araw) {                                                  // utils.inv: the matrices are given
        const cplx* Ar = a.Araw + (size_t)e * N * N;
        for (int i = warp; i < N; i += NW)
            for (int j = lane; j < N; j += 32) Am[i * ld + j] = Ar[(size_t)i * N + j];
    } else {
        const cplx E = a.E[e];
        const cplx* SB = a.SigB ? a.SigB + (size_t)e * a.strideSigB : nullptr;
        for (int i = warp; i < N; i += NW)
            for (int j = lane; j < N; j += 32) {
                const size_t g = (size_t)i * N + j;
                cplx v = csub(cmul(E, a.S[g]), a.F[g]);
                if (a.Sig0) v = csub(v, a.Sig0[g]);
                if (SB) v = csub(v, SB[g]);
                Am[i * ld + j] = v;
            }
    }
    if (t < N) rowsrc[t] = t;
    __syncthreads();
    for (int cidx = 0; cidx < a.ncontacts; cidx++) {              // contacts may overlap: one at a time
        const GnbSmallContact& ct = a.ct[cidx];
        const cplx* blk = ct.blk + (size_t)e * ct.blk_stride;
        for (int idx = t; idx < ct.nc * ct.nc; idx += NT) {
            const int r = idx / ct.nc, cc = idx - r * ct.nc;
            cplx* p = &Am[ct.inds[r] * ld + ct.inds[cc]];
            *p = csub(*p, blk[idx]);
        }
        __syncthreads();
    }

    // ---- in-place Gauss-Jordan inverse with partial pivoting
    for (int k = 0; k < N; k++) {
        // pivot search on column k, rows k..N-1, redundantly in every warp (max |a|^2, lowest row on ties)
        double best = -1.0;
        int p = k;
        for (int i = k + lane; i < N; i += 32) {
            const cplx v = Am[i * ld + k];
            const double m = v.x * v.x + v.y * v.y;
            if (m > best) { best = m; p = i; }
        }
#pragma unroll
        for (int off = 16; off; off >>= 1) {
            const double ob = __shfl_xor_sync(0xffffffffu, best, off);
            const int op = __shfl_xor_sync(0xffffffffu, p, off);
            if (ob > best || (ob == best && op < p)) { best = ob; p = op; }
        }
        if (best == 0.0 && t == 0) *a.info = 1;                    // exactly singular column (NaNs pass silently)
        const cplx r = cdiv(cmake(1.0, 0.0), Am[p * ld + k]);
        const cplx akk = Am[k * ld + k];                           // becomes the multiplier of the swapped-out row
        cplx rs[4], rk[4];
#pragma unroll
        for (int c4 = 0; c4 < 4; c4++) {
            const int j = lane + 32 * c4;
            if (j < N) {
                rs[c4] = (j == k) ? r : cmul(Am[p * ld + j], r);   // new row k
                rk[c4] = Am[k * ld + j];                           // old row k (moves to row p)
            }
        }
        if (t == 0) { const int s = rowsrc[k]; rowsrc[k] = rowsrc[p]; rowsrc[p] = s; }
        __syncthreads();                                           // rows p and k are read by everybody
        for (int i = warp; i < N; i += NW) {
            if (i == k) {
#pragma unroll
                for (int c4 = 0; c4 < 4; c4++) {
                    const int j = lane + 32 * c4;
                    if (j < N) Am[i * ld + j] = rs[c4];
                }
                continue;
            }
            const cplx f = (i == p) ? akk : Am[i * ld + k];
            __syncwarp();                                          // (i, k) is overwritten by lane k % 32 below
#pragma unroll
            for (int c4 = 0; c4 < 4; c4++) {
                const int j = lane + 32 * c4;
                if (j < N) {
                    cplx base = (i == p) ? rk[c4] : Am[i * ld + j];
                    if (j == k) base = cmake(0.0, 0.0);
                    Am[i * ld + j] = cfnma(base, f, rs[c4]);
                }
            }
        }
        __syncthreads();
    }
    // Am = (P A)^-1 with (P A)[i,:] = A[rowsrc[i],:]  ->  A^-1[:, j] = Am[:, q[j]],  q[rowsrc[i]] = i
    if (t < N) q[rowsrc[t]] = t;
    __syncthreads();

    if (a.mode == GNB_SMALL_GREEN) {
        cplx* G = a.G + (size_t)e * a.strideG;
        for (int i = warp; i < N; i += NW)
            for (int j = lane; j < N; j += 32) G[(size_t)i * a.ldg + j] = Am[i * ld + q[j]];
        return;
    }

    double acc = 0.0;
    if (a.mode == GNB_SMALL_DOS) {
        for (int i = t; i < N; i += NT) {
            const double d = -Am[i * ld + q[i]].y * kInvPi;
            if (a.dos_site) a.dos_site[(size_t)e * N + i] = d;
            acc += d;
        }
    } else {                                                        // GNB_SMALL_T
        // T = sum_{b in C1, d in C2} Y[b,d] conj(W[b,d]),  Y = G12 Gamma2,  W[b,d] = sum_a conj(Gamma1[a,b]) G12[a,d]
        const GnbSmallContact& c1 = a.ct[a.ca];
        const GnbSmallContact& c2 = a.ct[a.cb];
        const cplx* g1 = c1.gam + (size_t)e * c1.gam_stride;
        const cplx* g2 = c2.gam + (size_t)e * c2.gam_stride;
        const int n1 = c1.nc, n2 = c2.nc;
        for (int idx = t; idx < n1 * n2; idx += NT) {
            const int b = idx / n2, d = idx - b * n2;
            const int rb = c1.inds[b] * ld, qd = q[c2.inds[d]];
            cplx Y = cmake(0.0, 0.0), W = cmake(0.0, 0.0);
            for (int cc = 0; cc < n2; cc++) Y = cfma(Y, Am[rb + q[c2.inds[cc]]], g2[cc * n2 + d]);
            for (int aa = 0; aa < n1; aa++) W = cfma(W, cconj(g1[aa * n1 + b]), Am[c1.inds[aa] * ld + qd]);
            acc += Y.x * W.x + Y.y * W.y;
        }
    }
#pragma unroll
    for (int off = 16; off; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
    if (lane == 0) red[warp] = acc;
    __syncthreads();
    if (t == 0) {
        double s = 0.0;
        for (int w = 0; w < NW; w++) s += red[w];                  // fixed order: run-to-run reproducible
        if (a.mode == GNB_SMALL_DOS) a.dos_tot[e] = s;
        else a.T[e] = s;
    }
}

size_t small_smem(int N, int nt) {
    const int ld = N | 1;
    return (size_t)N * ld * sizeof(cplx) + (size_t)(2 * N) * sizeof(int) + (size_t)(nt / 32) * sizeof(double);
}

}  // namespace

cudaError_t gnb_small_init() {
    cudaError_t e;
    if ((e = cudaFuncSetAttribute(k_small_gj<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024))) return e;
    if ((e = cudaFuncSetAttribute(k_small_gj<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024))) return e;
    return cudaFuncSetAttribute(k_small_gj<512>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
}

void gnb_launch_small(cudaStream_t st, const GnbSmallArgs& a) {
    if (a.M <= 0) return;
    const int N = a.N;
    if (N <= 32) k_small_gj<128><<<a.M, 128, small_smem(N, 128), st>>>(a);
    else if (N <= 64) k_small_gj<256><<<a.M, 256, small_smem(N, 256), st>>>(a);
    else k_small_gj<512><<<a.M, 512, small_smem(N, 512), st>>>(a);
}
