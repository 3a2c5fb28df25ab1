// Batched complex128 block-elimination engine for sm_100a (B200).
//
// One launch family serves every large-N Green's-function reduction of the reference
// (utils.py:52-54 `inv`, integrate.py:67-82, transport.py:150-190):
//   * JORDAN mode  : in-place block Gauss-Jordan inverse  G = A^-1            (8 N^3 flops)
//   * FORWARD mode : block Gaussian elimination of [A | B] + unit-block-upper back-substitution,
//                    i.e. only the contact columns of G                      (8/3 N^3 + 8 N^2 m)
// Each block step (width NB = 32) is: tournament pivoting (parallel GEPP over 256-row groups, rows
// held in registers), a row-permutation + triangular solve of the pivot row block, and a rank-NB
// update of the whole batch on the FP64 tensor pipe (mma.sync m8n8k4 -> SASS DMMA.8x8x4; tcgen05
// has no f64 kind).  All energies of a chunk advance in lock-step so that every launch fills the
// 148 SMs.  Matrices are row-major interleaved complex128.
#include <algorithm>
#include <type_traits>
#include "gnb_common.cuh"
#include "gnb_kernels.h"

// ------------------------------------------------------------------------------------------
// Assembly  A_k = E_k * S - F - Sigma0 - SigmaB_k   (integrate.py:70,77; transport.py:153,186)
// HBM-bound: 16 N^2 B written per energy, F/S/Sigma0 stay L2-resident.
// ------------------------------------------------------------------------------------------
// mixr > 0 (mixed layout of the transmission path): the first mixr columns of every row are stored as real
// doubles, the rest as complex128; A is the LOGICAL complex base (A[row * ld + col] is valid for col >= mixr),
// the real view starts mixr * 8 bytes later with a row stride of 2 * ld doubles (see gnb_rec.cu).
__device__ __forceinline__ double* gnb_real_view(cplx* A, int mixr) {
    return reinterpret_cast<double*>(reinterpret_cast<char*>(A) + (size_t)mixr * 8);
}
__device__ __forceinline__ const double* gnb_real_view(const cplx* A, int mixr) {
    return reinterpret_cast<const double*>(reinterpret_cast<const char*>(A) + (size_t)mixr * 8);
}

// One CTA = one matrix row x ASM_EB energies: F, S (and Sigma0) of the row are read ONCE into registers and reused
// for every energy of the group, so the kernel is bound by its HBM writes (16 N ld B per energy), not by L2 reads of
// F and S per energy or by index arithmetic.
#define ASM_EB 16
__global__ void __launch_bounds__(256) k_assemble(cplx* __restrict__ A, long strideA, int ld, int N,
                                                  const cplx* __restrict__ F, const cplx* __restrict__ S,
                                                  const cplx* __restrict__ Sig0,
                                                  const cplx* __restrict__ SigB, long strideSigB,
                                                  const cplx* __restrict__ E, const int* __restrict__ pi, int mixr, int M) {
    __shared__ cplx sE[ASM_EB];
    const int i = blockIdx.x, b0 = blockIdx.y * ASM_EB, nb = min(ASM_EB, M - b0), t = threadIdx.x;
    if (t < nb) sE[t] = E[b0 + t];
    __syncthreads();
    const long rowsrc = (long)(pi ? pi[i] : i) * N;              // symmetric orbital reordering (contacts last)
    double* Ar = gnb_real_view(A, mixr);
    for (int j = t; j < N; j += 256) {
        const long src = rowsrc + (pi ? pi[j] : j);
        const cplx s = S[src], f = F[src];
        const cplx s0 = Sig0 ? Sig0[src] : cmake(0.0, 0.0);
        const bool real_slot = j < mixr;
        for (int bb = 0; bb < nb; bb++) {
            const long b = b0 + bb;
            cplx v = csub(cmul(sE[bb], s), f);
            if (Sig0) v = csub(v, s0);
            if (SigB) v = csub(v, SigB[b * strideSigB + src]);
            if (real_slot) Ar[b * 2 * strideA + (long)i * 2 * ld + j] = v.x;
            else A[b * strideA + (long)i * ld + j] = v;
        }
    }
}

// A[b][inds[p]][inds[q]] -= blk[b][p][q]   (surfG1D.py:372, surfGBethe.py:527 scatter of contact blocks)
__global__ void __launch_bounds__(256) k_scatter_sub(cplx* __restrict__ A, long strideA, int ld,
                                                     const int* __restrict__ inds, int nc,
                                                     const cplx* __restrict__ blk, long strideBlk,
                                                     const int* __restrict__ map) {
    const int b = blockIdx.y;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < nc * nc; idx += gridDim.x * blockDim.x) {
        const int p = idx / nc, q = idx - p * nc;
        const int ip = map ? map[inds[p]] : inds[p], iq = map ? map[inds[q]] : inds[q];
        cplx* a = A + (long)b * strideA + (long)ip * ld + iq;
        *a = csub(*a, blk[(long)b * strideBlk + idx]);
    }
}

// Augmented right-hand side: A[b][i][xoff + c] = (i == cols[c])
__global__ void __launch_bounds__(256) k_set_aug(cplx* __restrict__ A, long strideA, int ld, int N, int xoff,
                                                 const int* __restrict__ cols, int m, const int* __restrict__ map) {
    const int b = blockIdx.y;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < N * m; idx += gridDim.x * blockDim.x) {
        const int i = idx / m, c = idx - i * m;
        const int r = map ? map[cols[c]] : cols[c];
        A[(long)b * strideA + (long)i * ld + xoff + c] = cmake(i == r ? 1.0 : 0.0, 0.0);
    }
}
// identity on the padded diagonal [N, Np)
__global__ void k_pad_diag(cplx* __restrict__ A, long strideA, int ld, int N, int Np) {
    const int b = blockIdx.x, i = N + threadIdx.x;
    if (i < Np) A[(long)b * strideA + (long)i * ld + i] = cmake(1.0, 0.0);
}

__global__ void k_init_perm(int* __restrict__ perm, int stride, int N) {
    const int b = blockIdx.y;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += gridDim.x * blockDim.x)
        perm[(long)b * stride + i] = i;
}

// ------------------------------------------------------------------------------------------
// Tournament pivoting round.  One CTA = one group of <= GROUP candidate rows of one matrix; each
// thread keeps its row of the NB-wide panel in registers and the CTA runs Gaussian elimination
// with partial pivoting (LAPACK izamax metric |re|+|im|) WITHOUT physical swaps: the thread whose
// row is chosen retires.  The w chosen rows go to the next round; the final round (a single group)
// also emits the explicit inverse of the pivot block, the net row moves of this step and updates the
// running row permutation.
//
// The row is ROTATED by one element per pivot step, so the pivot column is always a[0] and the loop
// body has static register indices: the kernel is ~500 instructions instead of a 14 k-instruction
// fully unrolled triangle that missed the instruction cache on every step.  One barrier per step:
// every warp publishes its local winner (value, thread, row) into a double-buffered slot, all threads
// pick the global winner after the barrier.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ cplx crcp_fast(cplx p) {
    // 1 / p with a power-of-two pre-scaling (exact), one real reciprocal
    const double s = fmax(fabs(p.x), fabs(p.y));
    const int e = (__double2hiint(s) >> 20) & 0x7ff;
    const double sc = __hiloint2double((2046 - e) << 20, 0);          // 2^(1023 - e): |p| * sc in [1, 2.83)
    const double x = p.x * sc, y = p.y * sc;
    const double r = __drcp_rn(fma(x, x, y * y)) * sc;
    return cmake(x * r, -y * r);
}

// Scalar traits of the tournament: the NOMINATING rounds run in single precision (half the registers, twice the
// resident groups, no FP64-pipe use); they only pick which rows go on.  The FINAL round, which fixes the pivot
// order, and the pivot-block inverse are always double precision on the original matrix entries.
template <typename R> struct TT;
template <> struct TT<double> {
    typedef double2 C;
    static __device__ __forceinline__ C ld(cplx v) { return v; }
    static __device__ __forceinline__ C zero() { return make_double2(0.0, 0.0); }
    static __device__ __forceinline__ C mul(C a, C b) { return cmul(a, b); }
    static __device__ __forceinline__ C fnma(C a, C b, C c) { return cfnma(a, b, c); }
    static __device__ __forceinline__ C rcp(C p) { return crcp_fast(p); }
    static __device__ __forceinline__ unsigned long long key(C v) {          // order-preserving, > 0
        return (unsigned long long)__double_as_longlong(fabs(v.x) + fabs(v.y)) + 1ull;
    }
};
template <> struct TT<float> {
    typedef float2 C;
    static __device__ __forceinline__ C ld(cplx v) { return make_float2((float)v.x, (float)v.y); }
    static __device__ __forceinline__ C zero() { return make_float2(0.f, 0.f); }
    static __device__ __forceinline__ C mul(C a, C b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
    static __device__ __forceinline__ C fnma(C a, C b, C c) {
        return make_float2(fmaf(b.y, c.y, fmaf(-b.x, c.x, a.x)), fmaf(-b.y, c.x, fmaf(-b.x, c.y, a.y)));
    }
    static __device__ __forceinline__ C rcp(C p) {
        const float s = fmaxf(fabsf(p.x), fabsf(p.y));
        const float sc = 1.0f / s;
        const float x = p.x * sc, y = p.y * sc;
        const float r = sc / fmaf(x, x, y * y);
        return make_float2(x * r, -y * r);
    }
    static __device__ __forceinline__ unsigned long long key(C v) {
        return (unsigned long long)__float_as_uint(fabsf(v.x) + fabsf(v.y)) + 1ull;
    }
};

// Real panels (real F, S, E; columns left of the contact orbitals): the same tournament on the real parts only.
template <typename R> struct TTR;
template <> struct TTR<double> {
    typedef double C;
    static __device__ __forceinline__ C ld(cplx v) { return v.x; }
    static __device__ __forceinline__ C zero() { return 0.0; }
    static __device__ __forceinline__ C mul(C a, C b) { return a * b; }
    static __device__ __forceinline__ C fnma(C a, C b, C c) { return fma(-b, c, a); }
    static __device__ __forceinline__ C rcp(C p) { return 1.0 / p; }
    static __device__ __forceinline__ unsigned long long key(C v) {
        return (unsigned long long)__double_as_longlong(fabs(v)) + 1ull;
    }
};
template <> struct TTR<float> {
    typedef float C;
    static __device__ __forceinline__ C ld(cplx v) { return (float)v.x; }
    static __device__ __forceinline__ C zero() { return 0.f; }
    static __device__ __forceinline__ C mul(C a, C b) { return a * b; }
    static __device__ __forceinline__ C fnma(C a, C b, C c) { return fmaf(-b, c, a); }
    static __device__ __forceinline__ C rcp(C p) { return 1.0f / p; }
    static __device__ __forceinline__ unsigned long long key(C v) { return (unsigned long long)__float_as_uint(fabsf(v)) + 1ull; }
};

// Thread tiling: GROUP rows x 32 columns as 4-row x 8-column register tiles; warp w owns column group w
// (columns 8w .. 8w+7) of all rows, lane l owns rows 4l .. 4l+3.  The warp that owns the pivot column finds
// the pivot with REDUX operations (no cross-warp reduction), publishes the column (for the multipliers)
// and the winner; every warp takes its 8 elements of the pivot row from the lane of its own that holds them.
// One CTA barrier per pivot step; the owning warp searches the NEXT pivot (look-ahead) before it finishes its
// own update.  Inside the owning warp the tile is rotated by one column per step so that the pivot column is
// always local column 0 (static register indices, small code).
template <int GROUP, typename T, bool F64, int MINB>
__global__ void __launch_bounds__(GROUP, MINB)
k_tourn(const cplx* __restrict__ A, long strideA, int ld, int c0, int w, int r0, int n_in,
        const int* __restrict__ cand_in, int cand_in_stride, int* __restrict__ cand_out, int cand_out_stride,
        int final_round, cplx* __restrict__ LU, int* __restrict__ moves, int* __restrict__ perm, int perm_stride,
        int* __restrict__ info, int mixr) {
    static_assert(GROUP == 128, "tile mapping: 4 warps = 4 column groups, 32 lanes x 4 rows");
    typedef typename T::C C;
    const int b = blockIdx.y, g = blockIdx.x, t = threadIdx.x;
    const int lane = t & 31, tc = t >> 5;                  // tc = column group of this warp
    int rows[4];
    C a[4][8];
    const cplx* Ab = A + (long)b * strideA;
#pragma unroll
    for (int rr = 0; rr < 4; rr++) {
        const int i = g * GROUP + 4 * lane + rr;
        const bool valid = i < n_in;
        rows[rr] = valid ? (cand_in ? cand_in[(long)b * cand_in_stride + i] : r0 + i) : -1;
        if (c0 < mixr) {                                     // panel stored as real doubles (mixed layout)
            const double* srcr = gnb_real_view(Ab, mixr) + (long)(valid ? rows[rr] : 0) * 2 * ld + c0 + 8 * tc;
#pragma unroll
            for (int k = 0; k < 8; k++) a[rr][k] = (valid && 8 * tc + k < w) ? T::ld(cmake(srcr[k], 0.0)) : T::zero();
        } else {
            const cplx* src = Ab + (long)(valid ? rows[rr] : 0) * ld + c0 + 8 * tc;
#pragma unroll
            for (int k = 0; k < 8; k++) a[rr][k] = (valid && 8 * tc + k < w) ? T::ld(src[k]) : T::zero();
        }
    }
    const int ngroup = min(GROUP, n_in - g * GROUP);
    const int nsel = min(w, ngroup);

    __shared__ __align__(16) C s_col[2][GROUP];             // pivot column of every row (before scaling), [rr][lane]
    __shared__ int s_widx[2];                               // winner: rr * 32 + lane
    __shared__ int s_wnz[2];                                // winner magnitude: -1 none, 0 exactly zero, 1 positive
    __shared__ __align__(16) C s_prow[4][8];                // per warp: its 8 elements of the pivot row
    __shared__ int s_win[GNB_NB];
    __shared__ __align__(16) cplx s_B[F64 ? GNB_NB : 1][GNB_NB + 1];  // final round: pivot block -> inverse
    unsigned alive = 0;                                     // bit rr: row 4*lane + rr still a candidate
#pragma unroll
    for (int rr = 0; rr < 4; rr++) alive |= (rows[rr] >= 0 ? 1u : 0u) << rr;

    // pivot search of step jn on local column 0 of the owning warp; results go to buffer jn & 1
    auto search = [&](int jn) {
        const int nb_ = jn & 1;
        unsigned long long key = 0ull;
        int krr = 0;
#pragma unroll
        for (int rr = 0; rr < 4; rr++) {
            s_col[nb_][rr * 32 + lane] = a[rr][0];
            const unsigned long long kq = ((alive >> rr) & 1u) ? T::key(a[rr][0]) : 0ull;
            if (kq > key) { key = kq; krr = rr; }            // first maximum wins (izamax)
        }
        unsigned bal;
        if (!F64) {
            const unsigned k32 = (unsigned)key;
            const unsigned kmax = __reduce_max_sync(0xffffffffu, k32);
            bal = __ballot_sync(0xffffffffu, k32 == kmax);
        } else {
            const unsigned hi = (unsigned)(key >> 32), lo = (unsigned)key;
            const unsigned hmax = __reduce_max_sync(0xffffffffu, hi);
            const unsigned lmax = __reduce_max_sync(0xffffffffu, hi == hmax ? lo : 0u);
            bal = __ballot_sync(0xffffffffu, hi == hmax && lo == lmax);
        }
        if (lane == __ffs(bal) - 1) {
            s_widx[nb_] = krr * 32 + lane;
            s_wnz[nb_] = key == 0ull ? -1 : (key == T::key(T::zero()) ? 0 : 1);
            s_win[jn] = krr == 0 ? rows[0] : krr == 1 ? rows[1] : krr == 2 ? rows[2] : rows[3];
        }
    };
    if (tc == 0 && nsel > 0) search(0);

#pragma unroll 1
    for (int j = 0; j < nsel; j++) {
        const int buf = j & 1, jc = j >> 3;
        __syncthreads();
        const int wi = s_widx[buf];
        const int nz = s_wnz[buf];
        if (final_round && nz == 0 && t == 0) *info = 1;      // exactly singular pivot (LAPACK info > 0)
        if (lane == (wi & 31)) {
            const int wr = wi >> 5;
            alive &= ~(1u << wr);
            if (tc >= jc) {                                   // this lane holds the pivot row: publish this warp's 8 elements
#pragma unroll
                for (int k = 0; k < 8; k++) {
                    C v = a[0][k];
                    if (wr == 1) v = a[1][k];
                    if (wr == 2) v = a[2][k];
                    if (wr == 3) v = a[3][k];
                    s_prow[tc][k] = v;
                }
            }
        }
        if (tc < jc) continue;                                // all columns of this warp are eliminated (warp-uniform)
        __syncwarp();
        // LAPACK zgetf2 scales the column by the reciprocal of the pivot
        const C rinv = (nz > 0) ? T::rcp(s_col[buf][wi]) : T::zero();
        C l[4];
#pragma unroll
        for (int rr = 0; rr < 4; rr++) l[rr] = T::mul(s_col[buf][rr * 32 + lane], rinv);
        // look-ahead: the warp that owns the next pivot column updates that column first, searches the next
        // pivot and publishes it, and only then finishes its update -- the search overlaps the other warps' work
        const bool next_owner = (j + 1 < nsel) && (tc == ((j + 1) >> 3));
        if (tc == jc) {                                       // owning warp: update and rotate left by one column
            {
                const C p = s_prow[tc][1];
#pragma unroll
                for (int rr = 0; rr < 4; rr++) a[rr][0] = T::fnma(a[rr][1], l[rr], p);
            }
            if (next_owner) search(j + 1);
#pragma unroll
            for (int k = 1; k < 7; k++) {
                const C p = s_prow[tc][k + 1];
#pragma unroll
                for (int rr = 0; rr < 4; rr++) a[rr][k] = T::fnma(a[rr][k + 1], l[rr], p);
            }
#pragma unroll
            for (int rr = 0; rr < 4; rr++) a[rr][7] = T::zero();
        } else {
            {
                const C p = s_prow[tc][0];
#pragma unroll
                for (int rr = 0; rr < 4; rr++) a[rr][0] = T::fnma(a[rr][0], l[rr], p);
            }
            if (next_owner) search(j + 1);
#pragma unroll
            for (int k = 1; k < 8; k++) {
                const C p = s_prow[tc][k];
#pragma unroll
                for (int rr = 0; rr < 4; rr++) a[rr][k] = T::fnma(a[rr][k], l[rr], p);
            }
        }
        __syncwarp();                                         // s_prow[tc] is rewritten in the next step
    }
    __syncthreads();
    if (!final_round || !F64) {
        if (t < nsel) cand_out[(long)b * cand_out_stride + g * w + t] = s_win[t];
        return;
    }
    // ---- final round: explicit inverse of the pivot block (chosen rows, pivot order) by Gauss-Jordan in
    // shared memory without further pivoting (the row order IS the partial-pivoting order).
    {
        const cplx* Ab = A + (long)b * strideA;
        for (int e = t; e < GNB_NB * GNB_NB; e += GROUP) {
            const int r = e >> 5, c = e & 31;
            cplx v = cmake(r == c ? 1.0 : 0.0, 0.0);
            if (r < w && c < w)
                v = (c0 < mixr) ? cmake(gnb_real_view(Ab, mixr)[(long)s_win[r] * 2 * ld + c0 + c], 0.0)
                                : Ab[(long)s_win[r] * ld + c0 + c];
            s_B[r][c] = v;
        }
        constexpr int PER = GNB_NB * GNB_NB / GROUP;
        for (int k = 0; k < w; k++) {
            __syncthreads();
            const cplx piv = s_B[k][k];
            const cplx rk = (piv.x != 0.0 || piv.y != 0.0) ? crcp_fast(piv) : cmake(0.0, 0.0);
            cplx nv[PER];
#pragma unroll
            for (int q = 0; q < PER; q++) {
                const int e = t + q * GROUP, r = e >> 5, c = e & 31;
                const cplx pk = cmul(s_B[k][c], rk);          // scaled pivot row
                if (r == k) nv[q] = (c == k) ? rk : pk;
                else {
                    const cplx f = s_B[r][k];
                    nv[q] = (c == k) ? cneg(cmul(f, rk)) : cfnma(s_B[r][c], f, pk);
                }
            }
            __syncthreads();
#pragma unroll
            for (int q = 0; q < PER; q++) {
                const int e = t + q * GROUP;
                s_B[e >> 5][e & 31] = nv[q];
            }
        }
        __syncthreads();
        cplx* inv = LU + (long)b * GNB_NB * GNB_NB;
        for (int e = t; e < GNB_NB * GNB_NB; e += GROUP) {
            const int r = e >> 5, c = e & 31;
            inv[e] = (r < w && c < w) ? s_B[r][c] : cmake(0.0, 0.0);
        }
    }
    // ---- net row moves + permutation bookkeeping (warp 0) ---------------------------------------
    if (t < 32) {
        const bool act = t < w;
        const int ch = act ? s_win[t] : -1;                    // chosen global row, pivot order
        const bool in_blk = act && ch < c0 + w;                // ch >= c0 always
        const unsigned chosen_pos = __reduce_or_sync(0xffffffffu, in_blk ? (1u << (ch - c0)) : 0u);
        const unsigned vacmask = __ballot_sync(0xffffffffu, act && !in_blk);
        const unsigned blkmask = (w == 32) ? 0xffffffffu : ((1u << w) - 1u);
        const unsigned dismask = blkmask & ~chosen_pos;        // block rows that were not chosen
        int* mv = moves + (long)b * GNB_MOVES_STRIDE;
        int d2 = -1, s2 = -1;
        if (act) { mv[1 + 2 * t] = c0 + t; mv[2 + 2 * t] = ch; }
        if (act && !in_blk) {
            const int rank = __popc(vacmask & ((1u << t) - 1u));
            const int p = __fns(dismask, 0, rank + 1);
            d2 = ch; s2 = c0 + p;                              // displaced block row fills the vacated slot
            mv[1 + 2 * (w + rank)] = d2;
            mv[2 + 2 * (w + rank)] = s2;
        }
        if (t == 0) mv[0] = w + __popc(vacmask);
        if (perm) {
            int* pb = perm + (long)b * perm_stride;
            const int o1 = act ? pb[ch] : 0;
            const int o2 = (s2 >= 0) ? pb[s2] : 0;
            __syncwarp();
            if (act) pb[c0 + t] = o1;
            __syncwarp();
            if (d2 >= 0) pb[d2] = o2;
        }
    }
}

// ------------------------------------------------------------------------------------------
// Warp-synchronous tournament round (block width 32).  One CTA = 4 warps = up to 256 candidate rows.
//   level 0: every warp runs Gaussian elimination with partial pivoting on ITS 64 rows (2 per lane, all 32 panel
//            columns of a row in the registers of one lane) and nominates 32 of them;
//   level 1: warps 0 / 1 merge two nominee lists each (64 rows, entries re-read from the matrix), nominate 32;
//   level 2: warp 0 merges the last two lists.
// Inside a warp a pivot step needs no CTA barrier and no cross-warp reduction: pivot search = REDUX + ballot, the
// pivot row goes through a 32-element per-warp buffer (one lane writes, all read with LDS.128), the update is a
// fully unrolled triangle of FMAs on registers.  Against k_tourn (128 rows per CTA, one CTA barrier per step, column
// groups spread over the warps) a round issues less than half the warp instructions per candidate row, the CTA
// reduces 256 rows instead of 128 (N = 1024: two launches per block instead of three), and the final round's
// latency chain is 3 x 32 warp steps + a warp-level in-place Gauss-Jordan inverse instead of 64 CTA-barrier steps.
// The final round (F64 on the original entries) fixes the pivot order, emits the pivot-block inverse, the net row
// moves and the permutation update exactly like k_tourn.
// ------------------------------------------------------------------------------------------
#ifdef TW_DEBUG
__device__ __forceinline__ double dbgv(float v) { return v; }
__device__ __forceinline__ double dbgv(double v) { return v; }
__device__ __forceinline__ double dbgv(float2 v) { return v.x; }
__device__ __forceinline__ double dbgv(double2 v) { return v.x; }
#endif
// 16-byte chunks of the per-warp pivot-row buffer: PER elements of type C per LDS.128 / STS.128 (no type punning
// through unions: explicit vector types)
template <typename C> struct TwChunk;
template <> struct TwChunk<float> {
    static constexpr int PER = 4;
    static __device__ __forceinline__ void st(float* p, const float (&e)[4]) { *reinterpret_cast<float4*>(p) = make_float4(e[0], e[1], e[2], e[3]); }
    static __device__ __forceinline__ void ld(const float* p, float (&e)[4]) {
        const float4 v = *reinterpret_cast<const float4*>(p);
        e[0] = v.x; e[1] = v.y; e[2] = v.z; e[3] = v.w;
    }
};
template <> struct TwChunk<float2> {
    static constexpr int PER = 2;
    static __device__ __forceinline__ void st(float2* p, const float2 (&e)[2]) { *reinterpret_cast<float4*>(p) = make_float4(e[0].x, e[0].y, e[1].x, e[1].y); }
    static __device__ __forceinline__ void ld(const float2* p, float2 (&e)[2]) {
        const float4 v = *reinterpret_cast<const float4*>(p);
        e[0] = make_float2(v.x, v.y); e[1] = make_float2(v.z, v.w);
    }
};
template <> struct TwChunk<double> {
    static constexpr int PER = 2;
    static __device__ __forceinline__ void st(double* p, const double (&e)[2]) { *reinterpret_cast<double2*>(p) = make_double2(e[0], e[1]); }
    static __device__ __forceinline__ void ld(const double* p, double (&e)[2]) {
        const double2 v = *reinterpret_cast<const double2*>(p);
        e[0] = v.x; e[1] = v.y;
    }
};
template <> struct TwChunk<double2> {
    static constexpr int PER = 1;
    static __device__ __forceinline__ void st(double2* p, const double2 (&e)[1]) { *p = e[0]; }
    static __device__ __forceinline__ void ld(const double2* p, double2 (&e)[1]) { e[0] = *p; }
};

template <typename T, int NR = 2>
__device__ __forceinline__ void tw_load(typename T::C (&a)[NR][32], const int (&rows)[NR], const cplx* __restrict__ Ab, int ld,
                                        int c0, int mixr) {
    // loads go out in groups of 8 columns (the compiler would otherwise keep all 32 raw 16-byte values of a row live)
#pragma unroll
    for (int rr = 0; rr < NR; rr++) {
        if (rows[rr] < 0) {
#pragma unroll
            for (int k = 0; k < 32; k++) a[rr][k] = T::zero();
        } else if (c0 < mixr) {                              // panel stored as real doubles (mixed layout)
            const double2* src = reinterpret_cast<const double2*>(gnb_real_view(Ab, mixr) + (long)rows[rr] * 2 * ld + c0);
#pragma unroll
            for (int q0 = 0; q0 < 16; q0 += 8) {
                double2 v[8];
#pragma unroll
                for (int q = 0; q < 8; q++) v[q] = src[q0 + q];
#pragma unroll
                for (int q = 0; q < 8; q++) {
                    a[rr][2 * (q0 + q)] = T::ld(cmake(v[q].x, 0.0));
                    a[rr][2 * (q0 + q) + 1] = T::ld(cmake(v[q].y, 0.0));
                }
                asm volatile("" ::: "memory");
            }
        } else {
            const cplx* src = Ab + (long)rows[rr] * ld + c0;
#pragma unroll
            for (int k0 = 0; k0 < 32; k0 += 8) {
                cplx v[8];
#pragma unroll
                for (int k = 0; k < 8; k++) v[k] = src[k0 + k];
#pragma unroll
                for (int k = 0; k < 8; k++) a[rr][k0 + k] = T::ld(v[k]);
                asm volatile("" ::: "memory");
            }
        }
    }
}

// WIDTH pivot steps [j0, j1) of the warp GEPP.  The rows are ROTATED left by one column per step (fused into the
// update: a[c - 1] = a[c] - l p[c]), so the pivot column is always register column 0, every register index is
// static and the step loop stays rolled: a fully unrolled triangle is faster per step on paper but ~9 k instructions
// per kernel, and a tournament CTA runs its code exactly once - it was instruction-fetch bound (85 us per final
// round against 51 us of the CTA-wide kernel).  WIDTH = live columns at j0 (32, then 16 for the second half).
template <typename T, bool F64, int WIDTH, int NR = 2>
__device__ __forceinline__ void tw_steps(typename T::C (&a)[NR][32], const int (&rows)[NR], unsigned& alive, int j0, int j1,
                                         typename T::C* s_prow, int* s_win, int lane, bool flag_singular, int* info) {
    typedef typename T::C C;
    typedef TwChunk<C> CH;
    constexpr int PER = CH::PER;
    const unsigned long long kzero = T::key(T::zero());
#pragma unroll 1
    for (int j = j0; j < j1; j++) {
        // first maximum wins (izamax): the rows of a lane are consecutive candidates, lower row first
        unsigned long long key = (alive & 1u) ? T::key(a[0][0]) : 0ull;
        int krr = 0;
#pragma unroll
        for (int rr = 1; rr < NR; rr++) {
            const unsigned long long kq = ((alive >> rr) & 1u) ? T::key(a[rr][0]) : 0ull;
            if (kq > key) { key = kq; krr = rr; }
        }
        unsigned bal;
        bool nonzero;
        if (!F64) {
            const unsigned k32 = (unsigned)key;
            const unsigned kmax = __reduce_max_sync(0xffffffffu, k32);
            bal = __ballot_sync(0xffffffffu, k32 == kmax);
            nonzero = kmax > (unsigned)kzero;
        } else {
            const unsigned hi = (unsigned)(key >> 32), lo = (unsigned)key;
            const unsigned hmax = __reduce_max_sync(0xffffffffu, hi);
            const unsigned lmax = __reduce_max_sync(0xffffffffu, hi == hmax ? lo : 0u);
            bal = __ballot_sync(0xffffffffu, hi == hmax && lo == lmax);
            nonzero = (((unsigned long long)hmax << 32) | lmax) > kzero;
        }
        const int src = __ffs(bal) - 1;
        if (lane == src) {                                   // the lane of the pivot row publishes its live columns
            alive &= ~(1u << krr);
#pragma unroll
            for (int rr = 0; rr < NR; rr++) {
                if (krr == rr) {
                    s_win[j] = rows[rr];
#pragma unroll
                    for (int cb = 0; cb < WIDTH; cb += PER) {
                        C u[PER];
#pragma unroll
                        for (int e = 0; e < PER; e++) u[e] = a[rr][cb + e];
                        CH::st(&s_prow[cb], u);
                    }
                }
            }
        }
        __syncwarp();
        if (flag_singular && !nonzero && lane == 0) *info = 1;              // exactly singular pivot (LAPACK info > 0)
        // LAPACK zgetf2 scales the column by the reciprocal of the pivot
        C l[NR];
        {
            C u0[PER];
            CH::ld(&s_prow[0], u0);
            const C rinv = nonzero ? T::rcp(u0[0]) : T::zero();
#pragma unroll
            for (int rr = 0; rr < NR; rr++) l[rr] = T::mul(a[rr][0], rinv);
#pragma unroll
            for (int e = 1; e < PER; e++) {
#pragma unroll
                for (int rr = 0; rr < NR; rr++) a[rr][e - 1] = T::fnma(a[rr][e], l[rr], u0[e]);
            }
        }
#pragma unroll
        for (int cb = PER; cb < WIDTH; cb += PER) {
            C u[PER];
            CH::ld(&s_prow[cb], u);
#pragma unroll
            for (int e = 0; e < PER; e++) {
#pragma unroll
                for (int rr = 0; rr < NR; rr++) a[rr][cb + e - 1] = T::fnma(a[rr][cb + e], l[rr], u[e]);
            }
        }
#pragma unroll
        for (int rr = 0; rr < NR; rr++) a[rr][WIDTH - 1] = T::zero();
        __syncwarp();                                        // s_prow is rewritten in the next step
    }
}

// GEPP of the (up to) 64 rows held by one warp; winners (global row numbers, pivot order) -> s_win[0 .. nsel).
// s_prow / s_win are written by one lane and read by the others: never pass them as __restrict__ (the compiler then
// keeps stale copies of the buffer in registers across __syncwarp).
template <typename T, bool F64, int NR = 2>
__device__ __forceinline__ void tw_gepp(typename T::C (&a)[NR][32], const int (&rows)[NR], int nsel, typename T::C* s_prow,
                                        int* s_win, int lane, bool flag_singular, int* info) {
    unsigned alive = 0;
#pragma unroll
    for (int rr = 0; rr < NR; rr++) alive |= (rows[rr] >= 0 ? 1u : 0u) << rr;
    tw_steps<T, F64, 32, NR>(a, rows, alive, 0, min(nsel, 16), s_prow, s_win, lane, flag_singular, info);
    if (nsel > 16) tw_steps<T, F64, 16, NR>(a, rows, alive, 16, nsel, s_prow, s_win, lane, flag_singular, info);
}

// Final round of a real panel, ONE warp: explicit inverse of the pivot block (chosen rows s_win[0 .. nfin), pivot order) by
// in-place Gauss-Jordan without further pivoting, then the net row moves and the permutation update (as in k_tourn).
__device__ __forceinline__ void tw_finish_real(const cplx* __restrict__ Ab, int ld, int c0, int mixr, const int* s_win, int nfin,
                                               double* s_gj, cplx* __restrict__ LUb, int* __restrict__ mvb, int* pb, int lane,
                                               int* info = nullptr) {
    constexpr int w = GNB_NB;
    // ---- final round, warp 0: explicit inverse of the pivot block (chosen rows, pivot order) by in-place Gauss-Jordan
    // without further pivoting (the row order IS the partial-pivoting order); lane r holds row r.  Real panels only
    // (the launcher keeps complex FP64 final rounds on k_tourn).
    {
        double m[32];
        const int row = lane < nfin ? s_win[lane] : -1;
        if (row >= 0) {
            if (c0 < mixr) {
                const double2* src = reinterpret_cast<const double2*>(gnb_real_view(Ab, mixr) + (long)row * 2 * ld + c0);
#pragma unroll
                for (int q = 0; q < 16; q++) { const double2 v = src[q]; m[2 * q] = v.x; m[2 * q + 1] = v.y; }
            } else {
                const cplx* src = Ab + (long)row * ld + c0;
#pragma unroll
                for (int k = 0; k < 32; k++) m[k] = src[k].x;
            }
        } else {
#pragma unroll
            for (int k = 0; k < 32; k++) m[k] = (k == lane) ? 1.0 : 0.0;
        }
        // the row is rotated left by one column per step WITH wrap-around (the finished inverse column goes to
        // register column 31): the pivot column is always register column 0 and after 32 steps the columns are back
        // in natural order
#pragma unroll 1
        for (int k = 0; k < 32; k++) {
            if (lane == k) {
#pragma unroll
                for (int q = 0; q < 16; q++) *reinterpret_cast<double2*>(&s_gj[2 * q]) = make_double2(m[2 * q], m[2 * q + 1]);
            }
            __syncwarp();
            const bool isp = lane == k;
            const double f = m[0];
            double rk;
            {
                const double2 pv = *reinterpret_cast<const double2*>(&s_gj[0]);
                rk = pv.x != 0.0 ? 1.0 / pv.x : 0.0;
                if (info && pv.x == 0.0 && lane == 0) *info = 1;       // exactly singular pivot (when the order came from FP32 rounds)
                const double p1 = pv.y * rk;                           // scaled pivot row
                m[0] = isp ? p1 : fma(-f, p1, m[1]);
            }
#pragma unroll
            for (int q = 1; q < 16; q++) {
                const double2 pv = *reinterpret_cast<const double2*>(&s_gj[2 * q]);
                const double p0 = pv.x * rk, p1 = pv.y * rk;
                m[2 * q - 1] = isp ? p0 : fma(-f, p0, m[2 * q]);
                m[2 * q] = isp ? p1 : fma(-f, p1, m[2 * q + 1]);
            }
            m[31] = isp ? rk : -f * rk;
            __syncwarp();
        }
        cplx* inv = LUb + (long)lane * GNB_NB;
#pragma unroll
        for (int k = 0; k < 32; k++) inv[k] = cmake(m[k], 0.0);
    }
    // ---- net row moves + permutation bookkeeping (same as k_tourn) ----------------------------------
    {
        const bool act = lane < w;
        const int ch = act ? s_win[lane] : -1;                 // chosen global row, pivot order
        const bool in_blk = act && ch < c0 + w;                // ch >= c0 always
        const unsigned chosen_pos = __reduce_or_sync(0xffffffffu, in_blk ? (1u << (ch - c0)) : 0u);
        const unsigned vacmask = __ballot_sync(0xffffffffu, act && !in_blk);
        const unsigned dismask = 0xffffffffu & ~chosen_pos;    // block rows that were not chosen
        int* mv = mvb;
        int d2 = -1, s2 = -1;
        if (act) { mv[1 + 2 * lane] = c0 + lane; mv[2 + 2 * lane] = ch; }
        if (act && !in_blk) {
            const int rank = __popc(vacmask & ((1u << lane) - 1u));
            const int p = __fns(dismask, 0, rank + 1);
            d2 = ch; s2 = c0 + p;                              // displaced block row fills the vacated slot
            mv[1 + 2 * (w + rank)] = d2;
            mv[2 + 2 * (w + rank)] = s2;
        }
        if (lane == 0) mv[0] = w + __popc(vacmask);
        if (pb) {
            const int o1 = act ? pb[ch] : 0;
            const int o2 = (s2 >= 0) ? pb[s2] : 0;
            __syncwarp();
            if (act) pb[c0 + lane] = o1;
            __syncwarp();
            if (d2 >= 0) pb[d2] = o2;
        }
    }
}

// Net row moves + permutation bookkeeping of a finished tournament (one warp; lanes 0..31 = pivot order)
__device__ __forceinline__ void tw_bookkeeping(int c0, const int* s_win, int* __restrict__ mvb, int* pb, int lane) {
    constexpr int w = GNB_NB;
    const bool act = lane < w;
    const int ch = act ? s_win[lane] : -1;                 // chosen global row, pivot order
    const bool in_blk = act && ch < c0 + w;                // ch >= c0 always
    const unsigned chosen_pos = __reduce_or_sync(0xffffffffu, in_blk ? (1u << (ch - c0)) : 0u);
    const unsigned vacmask = __ballot_sync(0xffffffffu, act && !in_blk);
    const unsigned dismask = 0xffffffffu & ~chosen_pos;    // block rows that were not chosen
    int d2 = -1, s2 = -1;
    if (act) { mvb[1 + 2 * lane] = c0 + lane; mvb[2 + 2 * lane] = ch; }
    if (act && !in_blk) {
        const int rank = __popc(vacmask & ((1u << lane) - 1u));
        const int p = __fns(dismask, 0, rank + 1);
        d2 = ch; s2 = c0 + p;                              // displaced block row fills the vacated slot
        mvb[1 + 2 * (w + rank)] = d2;
        mvb[2 + 2 * (w + rank)] = s2;
    }
    if (lane == 0) mvb[0] = w + __popc(vacmask);
    if (pb) {
        const int o1 = act ? pb[ch] : 0;
        const int o2 = (s2 >= 0) ? pb[s2] : 0;
        __syncwarp();
        if (act) pb[c0 + lane] = o1;
        __syncwarp();
        if (d2 >= 0) pb[d2] = o2;
    }
}

// Final round of a COMPLEX panel, one warp: as tw_finish_real, complex arithmetic (lane r holds row r of the pivot block as
// 32 complex doubles; s_gj: 32 complex).  The pivot order comes from the single-precision eliminations.
__device__ __forceinline__ void tw_finish_cplx(const cplx* __restrict__ Ab, int ld, int c0, const int* s_win, int nfin, cplx* s_gj,
                                               cplx* __restrict__ LUb, int* __restrict__ mvb, int* pb, int lane, int* info) {
    {
        cplx m[32];
        const int row = lane < nfin ? s_win[lane] : -1;
        if (row >= 0) {
            const cplx* src = Ab + (long)row * ld + c0;
#pragma unroll
            for (int k = 0; k < 32; k++) m[k] = src[k];
        } else {
#pragma unroll
            for (int k = 0; k < 32; k++) m[k] = cmake((k == lane) ? 1.0 : 0.0, 0.0);
        }
#pragma unroll 1
        for (int k = 0; k < 32; k++) {
            if (lane == k) {
#pragma unroll
                for (int q = 0; q < 32; q++) s_gj[q] = m[q];
            }
            __syncwarp();
            const bool isp = lane == k;
            const cplx f = m[0];
            const cplx p0 = s_gj[0];
            const bool nz = p0.x != 0.0 || p0.y != 0.0;
            if (info && !nz && lane == 0) *info = 1;                       // exactly singular pivot
            const cplx rk = nz ? crcp_fast(p0) : cmake(0.0, 0.0);
#pragma unroll
            for (int q = 1; q < 32; q++) {
                const cplx pq = cmul(s_gj[q], rk);                         // scaled pivot row
                m[q - 1] = isp ? pq : cfnma(m[q], f, pq);
            }
            m[31] = isp ? rk : cneg(cmul(f, rk));
            __syncwarp();
        }
        cplx* inv = LUb + (long)lane * GNB_NB;
#pragma unroll
        for (int k = 0; k < 32; k++) inv[k] = m[k];
    }
    tw_bookkeeping(c0, s_win, mvb, pb, lane);
}

template <typename T, bool F64, int MINB>
__global__ void __launch_bounds__(128, MINB)
k_tournw(const cplx* __restrict__ A, long strideA, int ld, int c0, int r0, int n_in, const int* __restrict__ cand_in,
         int cand_in_stride, int* __restrict__ cand_out, int cand_out_stride, int final_round, cplx* __restrict__ LU,
         int* __restrict__ moves, int* __restrict__ perm, int perm_stride, int* __restrict__ info, int mixr) {
    typedef typename T::C C;
    constexpr int w = GNB_NB;
    const int b = blockIdx.y, g = blockIdx.x, t = threadIdx.x, lane = t & 31, warp = t >> 5;
    __shared__ __align__(16) C s_prow[4][32];
    __shared__ int s_list[2][4][32];
    __shared__ int s_len[2][4];
    __shared__ __align__(16) double s_gj[F64 ? 32 : 1];
    const cplx* Ab = A + (long)b * strideA;
    const int base = g * 256;
    const int ncta = min(256, n_in - base);                  // candidate rows of this CTA
    int rows[2];
    C a[2][32];
    int nlists = (ncta + 63) / 64, cur = 0;                  // lists of nominees alive after the current level
    for (int level = 0;; level++) {                          // block-uniform loop: level 0 = 64 rows per warp, then merges
        int nsel = 0, dst = cur;
        bool flag = false;
        if (level == 0) {
            const int nw = max(0, min(64, ncta - 64 * warp));
#pragma unroll
            for (int rr = 0; rr < 2; rr++) {
                const int i = 64 * warp + 2 * lane + rr;
                rows[rr] = i < ncta ? (cand_in ? cand_in[(long)b * cand_in_stride + base + i] : r0 + base + i) : -1;
            }
            nsel = min(w, nw);
            flag = final_round && nlists == 1;
            if (lane == 0) s_len[cur][warp] = nsel;
        } else {
            const int nnext = (nlists + 1) / 2;
            dst = cur ^ 1;
            if (warp < nnext) {
                const int la = 2 * warp, lb = 2 * warp + 1;
                const int na = s_len[cur][la], nb_ = lb < nlists ? s_len[cur][lb] : 0;
                if (nb_ == 0) {                              // an unpaired list goes up unchanged (already in pivot order)
                    if (lane < na) s_list[dst][warp][lane] = s_list[cur][la][lane];
                } else {
#pragma unroll
                    for (int rr = 0; rr < 2; rr++) {
                        const int i = 2 * lane + rr;
                        rows[rr] = i < na ? s_list[cur][la][i] : (i < na + nb_ ? s_list[cur][lb][i - na] : -1);
                    }
                    nsel = min(w, na + nb_);
                    flag = final_round && nnext == 1;
                }
                if (lane == 0) s_len[dst][warp] = min(w, na + nb_);
            }
            cur = dst;
            nlists = nnext;
        }
        if (nsel > 0) {                                      // warp-uniform
            tw_load<T>(a, rows, Ab, ld, c0, mixr);
            tw_gepp<T, F64>(a, rows, nsel, s_prow[warp], s_list[dst][warp], lane, flag, info);
        }
        if (nlists == 1) break;
        __syncthreads();
    }
    __syncthreads();
    const int nfin = s_len[cur][0];
    const int* s_win = s_list[cur][0];
    if (!final_round || !F64) {
        if (t < nfin) cand_out[(long)b * cand_out_stride + g * w + t] = s_win[t];
        return;
    }
    if (warp != 0) return;
    tw_finish_real(Ab, ld, c0, mixr, s_win, nfin, s_gj, LU + (long)b * GNB_NB * GNB_NB, moves + (long)b * GNB_MOVES_STRIDE,
                   perm ? perm + (long)b * perm_stride : nullptr, lane);
}

// ------------------------------------------------------------------------------------------
// Warp-INDEPENDENT tournament round for real panels (block width 32).  One WARP = one group of up to 256 candidate
// rows: it runs the level-0 eliminations of its (up to four) 64-row lists one after the other, then the merges, with
// no CTA barrier anywhere.  k_tournw gives a 256-row group to a 4-warp CTA, but after level 0 two, then three of the
// four warps only wait at the barrier while they hold their registers (ncu, N = 1024 step: barrier = 37 % of the
// stall samples, 39 % issue utilisation, 23 % of the warp slots busy with 128 registers per thread); here every
// resident warp always has a pivot step to run, so an SM works on 16-20 groups at a time instead of on 4.
// The final round (F64) is one warp per matrix: eliminations, merges, the in-place Gauss-Jordan inverse of the pivot
// block and the bookkeeping stay in that warp (17 k warp instructions per matrix against 52 k of the CTA-wide k_tourn,
// whose shared-memory Gauss-Jordan runs complex arithmetic on real data).
// ------------------------------------------------------------------------------------------
template <typename T, bool F64, int NW, int MINB, int NR0>
__global__ void __launch_bounds__(32 * NW, MINB)
k_tournq(const cplx* __restrict__ A, long strideA, int ld, int c0, int r0, int n_in, int grp, int total_groups,
         const int* __restrict__ cand_in, int cand_in_stride, int* __restrict__ cand_out, int cand_out_stride,
         int final_round, cplx* __restrict__ LU, int* __restrict__ moves, int* __restrict__ perm, int perm_stride,
         int* __restrict__ info, int mixr) {
    // NR0 = candidate rows per lane at level 0 (lists of 32 NR0 rows): 4 halves the number of sequential eliminations
    // of a 256-row group (2 + 1 instead of 4 + 2 + 1) at 1.5 x the work per pivot step; the merges always hold 2 rows per lane
    typedef typename T::C C;
    constexpr bool CPLX = !std::is_floating_point<C>::value;          // float2 / double2 rows
    static_assert(!CPLX || (NR0 == 2 && !F64), "complex panels: two rows per lane, single-precision rounds");
    static_assert(NR0 == 2 || NR0 == 4, "rows per lane at level 0");
    constexpr int w = GNB_NB, L0 = 32 * NR0;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    __shared__ __align__(16) C s_prow_all[NW][32];
    __shared__ int s_list_all[NW][2][4][32];
    __shared__ __align__(16) double s_gj_all[NW][CPLX ? 64 : 32];
    const int gid = blockIdx.x * NW + warp;
    if (gid >= total_groups) return;                         // warp-uniform; no CTA barrier below
    const int b = gid / grp, g = gid - b * grp;
    C* s_prow = s_prow_all[warp];
    int (*s_list)[4][32] = s_list_all[warp];
    const cplx* Ab = A + (long)b * strideA;
    const int base = g * 256;
    const int ncand = min(256, n_in - base);                 // candidate rows of this warp
    int nl = (ncand + L0 - 1) / L0, cur = 0, level = 0, q = 0;
    unsigned lens = 0;                                       // 8 bits per list: nominees in list q of the current level
    unsigned lens_next = 0;
    if constexpr (NR0 == 4) {                                // level 0 with four rows per lane
        for (q = 0; q < nl; q++) {
            int rows4[4];
            C a4[4][32];
            const int nwr = min(L0, ncand - L0 * q);
#pragma unroll
            for (int rr = 0; rr < 4; rr++) {
                const int i = L0 * q + 4 * lane + rr;
                rows4[rr] = i < ncand ? (cand_in ? cand_in[(long)b * cand_in_stride + base + i] : r0 + base + i) : -1;
            }
            const int nsel = min(w, nwr);
            lens |= (unsigned)nsel << (8 * q);
            tw_load<T, 4>(a4, rows4, Ab, ld, c0, mixr);
            tw_gepp<T, F64, 4>(a4, rows4, nsel, s_prow, s_list[0][q], lane, F64 && final_round && nl == 1, info);
            __syncwarp();
        }
        level = 1; q = 0;
    }
    int rows[2];
    C a[2][32];
    while (nl > 1 || level == 0) {                           // warp-uniform loop over (level, q): ONE inlined elimination
        int nsel = 0, dst = level == 0 ? 0 : (cur ^ 1);
        bool flag = false;
        const int nn = (nl + 1) / 2;
        if (level == 0) {
            const int nwr = min(64, ncand - 64 * q);
#pragma unroll
            for (int rr = 0; rr < 2; rr++) {
                const int i = 64 * q + 2 * lane + rr;
                rows[rr] = i < ncand ? (cand_in ? cand_in[(long)b * cand_in_stride + base + i] : r0 + base + i) : -1;
            }
            nsel = min(w, nwr);
            flag = F64 && final_round && nl == 1;
            lens_next |= (unsigned)nsel << (8 * q);
        } else {
            const int la = 2 * q, lb = 2 * q + 1;
            const int na = (lens >> (8 * la)) & 255, nb_ = lb < nl ? (lens >> (8 * lb)) & 255 : 0;
            if (nb_ == 0) {                                  // an unpaired list goes up unchanged (already in pivot order)
                if (lane < na) s_list[dst][q][lane] = s_list[cur][la][lane];
            } else {
#pragma unroll
                for (int rr = 0; rr < 2; rr++) {
                    const int i = 2 * lane + rr;
                    rows[rr] = i < na ? s_list[cur][la][i] : (i < na + nb_ ? s_list[cur][lb][i - na] : -1);
                }
                nsel = min(w, na + nb_);
                flag = F64 && final_round && nn == 1;
            }
            lens_next |= (unsigned)min(w, na + nb_) << (8 * q);
        }
        if (nsel > 0) {
            tw_load<T>(a, rows, Ab, ld, c0, mixr);
            tw_gepp<T, F64>(a, rows, nsel, s_prow, s_list[dst][q], lane, flag, info);
        }
        __syncwarp();
        q++;
        if (q == (level == 0 ? nl : nn)) {                   // level finished
            if (level > 0) { cur ^= 1; nl = nn; }
            lens = lens_next; lens_next = 0;
            level++; q = 0;
        }
    }
    const int nfin = lens & 255;
    const int* s_win = s_list[cur][0];
    if (!final_round) {
        if (lane < nfin) cand_out[(long)b * cand_out_stride + g * w + lane] = s_win[lane];
        return;
    }
    // final round: the pivot block is inverted in FP64 on the original entries.  With T = float the pivot ORDER comes from
    // single-precision eliminations (a near-tie may resolve differently than in FP64: a threshold-pivoting order with
    // threshold 1 - 1e-7, as stable as partial pivoting); exact singularity is then detected by the FP64 Gauss-Jordan
    if constexpr (CPLX)
        tw_finish_cplx(Ab, ld, c0, s_win, nfin, reinterpret_cast<cplx*>(s_gj_all[warp]), LU + (long)b * GNB_NB * GNB_NB,
                       moves + (long)b * GNB_MOVES_STRIDE, perm ? perm + (long)b * perm_stride : nullptr, lane, info);
    else
        tw_finish_real(Ab, ld, c0, mixr, s_win, nfin, s_gj_all[warp], LU + (long)b * GNB_NB * GNB_NB,
                       moves + (long)b * GNB_MOVES_STRIDE, perm ? perm + (long)b * perm_stride : nullptr, lane, F64 ? nullptr : info);
}

// ------------------------------------------------------------------------------------------
// Row moves of one step applied to a 64-column tile, fused with the product that turns the pivot row
// block into  W = (L11 U11)^-1 A[k,:]  (explicit 32x32 inverse from the tournament's final round,
// broadcast from smem; 8 independent accumulators per thread).
// JORDAN: columns of the pivot block get the identity as right-hand side (in-place inverse trick).
// ------------------------------------------------------------------------------------------
#define PS_TC 64
// mode 0: FORWARD rule (form W only for columns right of the pivot block), 1: JORDAN rule (all
// columns; identity right-hand side in the pivot columns), 2: row moves only.
// pre_L != nullptr (two-level elimination, second inner block on the far columns): the pivot rows
// first receive the pending update of the first inner block, R -= pre_L * W_a, with pre_L the
// w x pre_k block left of the pivot block and W_a = rows [pre_row, pre_row + pre_k) of A.
__global__ void __launch_bounds__(256) k_permute_solve(cplx* __restrict__ A, long strideA, int ld,
                                                       int c0, int w, int col_lo, int col_hi,
                                                       const int* __restrict__ moves,
                                                       const cplx* __restrict__ LU, int mode,
                                                       const cplx* __restrict__ pre_L, long strideL, int ldL,
                                                       int pre_row, int pre_k) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cplx* tile = reinterpret_cast<cplx*>(smem_raw);                 // [2*NB][PS_TC]
    cplx* sLU = tile + 2 * GNB_NB * PS_TC;                          // [NB][NB]
    cplx* sL = sLU + GNB_NB * GNB_NB;                               // [NB][NB] (pre-update block)
    __shared__ int s_dst[2 * GNB_NB], s_src[2 * GNB_NB];
    const int b = blockIdx.y, t = threadIdx.x;
    const int cs = col_lo + blockIdx.x * PS_TC;
    cplx* Ab = A + (long)b * strideA;
    const int* mv = moves + (long)b * GNB_MOVES_STRIDE;
    const int nm = mv[0];
    if (t < nm) { s_dst[t] = mv[1 + 2 * t]; s_src[t] = mv[2 + 2 * t]; }
    if (mode != 2)
        for (int idx = t; idx < GNB_NB * GNB_NB; idx += 256) sLU[idx] = LU[(long)b * GNB_NB * GNB_NB + idx];
    if (pre_L)
        for (int idx = t; idx < GNB_NB * GNB_NB; idx += 256) {
            const int i = idx / GNB_NB, j = idx - i * GNB_NB;
            sL[idx] = (i < w && j < pre_k) ? pre_L[(long)b * strideL + (long)i * ldL + j] : cmake(0.0, 0.0);
        }
    __syncthreads();
    for (int idx = t; idx < nm * PS_TC; idx += 256) {
        const int m = idx / PS_TC, c = idx - m * PS_TC, col = cs + c;
        tile[idx] = (col < col_hi) ? Ab[(long)s_src[m] * ld + col] : cmake(0.0, 0.0);
    }
    __syncthreads();
    if (mode != 2) {
        // thread -> column c = t % 64, rows t/64 + 4 q (8 independent accumulators)
        const int c = t & (PS_TC - 1), rg = t >> 6, col = cs + c;
        const bool in_piv = (col >= c0 && col < c0 + w);
        const bool do_solve = (col < col_hi) && (mode == 1 ? true : (col >= c0 + w));
        cplx acc[8];
        if (pre_L) {                                             // R = rows - pre_L * W_a
            if (do_solve && !in_piv) {
#pragma unroll
                for (int q = 0; q < 8; q++) acc[q] = tile[(rg + 4 * q) * PS_TC + c];
                for (int j = 0; j < pre_k; j++) {
                    const cplx wa = Ab[(long)(pre_row + j) * ld + col];
#pragma unroll
                    for (int q = 0; q < 8; q++) acc[q] = cfnma(acc[q], sL[(rg + 4 * q) * GNB_NB + j], wa);
                }
            }
            __syncthreads();
            if (do_solve && !in_piv) {
#pragma unroll
                for (int q = 0; q < 8; q++)
                    if (rg + 4 * q < w) tile[(rg + 4 * q) * PS_TC + c] = acc[q];
            }
            __syncthreads();
        }
        if (do_solve) {                                          // W = (L11 U11)^-1 R
            if (in_piv) {                                        // identity right-hand side: W[:, K] = inverse
#pragma unroll
                for (int q = 0; q < 8; q++) acc[q] = sLU[(rg + 4 * q) * GNB_NB + (col - c0)];
            } else {
#pragma unroll
                for (int q = 0; q < 8; q++) acc[q] = cmake(0.0, 0.0);
                for (int j = 0; j < w; j++) {
                    const cplx r = tile[j * PS_TC + c];
#pragma unroll
                    for (int q = 0; q < 8; q++) acc[q] = cfma(acc[q], sLU[(rg + 4 * q) * GNB_NB + j], r);
                }
            }
        }
        __syncthreads();
        if (do_solve) {
#pragma unroll
            for (int q = 0; q < 8; q++)
                if (rg + 4 * q < w) tile[(rg + 4 * q) * PS_TC + c] = acc[q];
        }
    }
    __syncthreads();
    for (int idx = t; idx < nm * PS_TC; idx += 256) {
        const int m = idx / PS_TC, c = idx - m * PS_TC, col = cs + c;
        if (col < col_hi) Ab[(long)s_dst[m] * ld + col] = tile[idx];
    }
}

// JORDAN: save the (permuted) panel column as the left GEMM operand and clear it in place, so that
// the rank-NB update  A <- A - P W  writes  -P W[:,K]  there (in-place inverse).
__global__ void __launch_bounds__(256) k_save_panel(cplx* __restrict__ A, long strideA, int ld, int nrows,
                                                    int c0, int w, cplx* __restrict__ Pws, long stridePws,
                                                    int pld, int pcol) {
    const int b = blockIdx.y;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < nrows * GNB_NB; idx += gridDim.x * blockDim.x) {
        const int i = idx / GNB_NB, c = idx - i * GNB_NB;
        const bool piv_row = (i >= c0 && i < c0 + w);
        cplx v = cmake(0.0, 0.0);
        if (!piv_row && c < w) {
            cplx* a = A + (long)b * strideA + (long)i * ld + c0 + c;
            v = *a;
            *a = cmake(0.0, 0.0);
        }
        Pws[(long)b * stridePws + (long)i * pld + pcol + c] = v;
    }
}

// ------------------------------------------------------------------------------------------
// Complex rank-K update on the FP64 tensor pipe:   C (+/-)= (scale * P) * W
// CTA tile 64x64, 8 warps as 4(M) x 2(N), warp tile 16x32 = 2x4 DMMA.8x8x4 tiles, each complex
// tile product = 4 real DMMAs.  Shared-memory row strides are chosen == 64 B (P) / 32 B (W)
// modulo 128 B so that every LDS.128 fragment load is bank-conflict free.
// WT     : W is given as rows [n][k] and used conjugate-transposed (G Gamma G^dagger products).
// BATCHK : the batch is folded into the K loop (deterministic on-device reduction over energies),
//          P scaled by wscale[b] (quadrature weight).
// ------------------------------------------------------------------------------------------
#define GM_T 64
#define GM_KC 32
#define GM_PS (GM_KC + 4)      // P tile row stride (cplx): 36*16 = 576 B == 64 mod 128
#define GM_WS (GM_T + 2)       // W tile row stride (cplx): 66*16 = 1056 B == 32 mod 128

template <bool WT, bool BATCHK, int BM>
__global__ void __launch_bounds__(BM * 4, BM == 64 ? 2 : 4) k_gemm(GnbGemmArgs g) {
    constexpr int NT = BM * 4;                                      // 8 warps (BM=64) or 4 warps (BM=32)
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cplx* Ps = reinterpret_cast<cplx*>(smem_raw);                   // [BM][GM_PS]
    cplx* Ws = Ps + BM * GM_PS;                                     // [32][GM_WS] or [64][GM_PS]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int gid = lane >> 2, tig = lane & 3;
    const int wm = warp >> 1, wn = warp & 1;
    const int i0 = g.ilo + blockIdx.y * BM, j0 = g.jlo + blockIdx.x * GM_T;
    const int bz = blockIdx.z;
    cplx* Cb = g.C + (BATCHK ? 0 : (long)bz * g.strideC);

    double cre[2][4][2], cim[2][4][2];
#pragma unroll
    for (int mi = 0; mi < 2; mi++)
#pragma unroll
        for (int ni = 0; ni < 4; ni++) {
            const int r = i0 + wm * 16 + mi * 8 + gid;
            const int c = j0 + wn * 32 + ni * 8 + tig * 2;
            cplx v0 = cmake(0.0, 0.0), v1 = cmake(0.0, 0.0);
            if (!g.zero_init && r < g.ihi) {
                if (c < g.jhi) v0 = Cb[(long)r * g.ldc + c];
                if (c + 1 < g.jhi) v1 = Cb[(long)r * g.ldc + c + 1];
            }
            cre[mi][ni][0] = v0.x; cim[mi][ni][0] = v0.y;
            cre[mi][ni][1] = v1.x; cim[mi][ni][1] = v1.y;
        }

    const int nb = BATCHK ? g.nbatch_k : 1;
    for (int bb = 0; bb < nb; bb++) {
        const int b = BATCHK ? bb : bz;
        const cplx* Pb = g.P + (long)b * g.strideP;
        const cplx* Wb = g.W + (long)b * g.strideW;
        cplx sc = cmake(g.plus ? 1.0 : -1.0, 0.0);
        if (g.wscale) { const cplx ws = g.wscale[b]; sc = g.plus ? ws : cneg(ws); }
        for (int k0 = 0; k0 < g.kdim; k0 += GM_KC) {
            __syncthreads();
#pragma unroll
            for (int q = 0; q < (BM * GM_KC) / NT; q++) {
                const int idx = tid + q * NT;
                const int r = idx / GM_KC, k = idx - r * GM_KC;
                cplx v = cmake(0.0, 0.0);
                if (i0 + r < g.ihi && k0 + k < g.kdim) {
                    if (g.Pr) {                                  // real-stored operand (mixed layout)
                        const double pr = g.Pr[(long)b * g.stridePr + (long)(i0 + r) * g.ldpr + k0 + k];
                        v = cmake(sc.x * pr, sc.y * pr);
                    } else {
                        v = cmul(sc, Pb[(long)(i0 + r) * g.ldp + k0 + k]);
                    }
                }
                Ps[r * GM_PS + k] = v;
            }
            if (WT) {
#pragma unroll
                for (int q = 0; q < (GM_T * GM_KC) / NT; q++) {
                    const int idx = tid + q * NT;
                    const int n = idx / GM_KC, k = idx - n * GM_KC;
                    cplx v = cmake(0.0, 0.0);
                    if (j0 + n < g.jhi && k0 + k < g.kdim) v = cconj(Wb[(long)(j0 + n) * g.ldw + k0 + k]);
                    Ws[n * GM_PS + k] = v;
                }
            } else {
#pragma unroll
                for (int q = 0; q < (GM_T * GM_KC) / NT; q++) {
                    const int idx = tid + q * NT;
                    const int k = idx / GM_T, n = idx - k * GM_T;
                    cplx v = cmake(0.0, 0.0);
                    if (k0 + k < g.kdim && j0 + n < g.jhi) v = Wb[(long)(k0 + k) * g.ldw + j0 + n];
                    Ws[k * GM_WS + n] = v;
                }
            }
            __syncthreads();
            const int kend = min(GM_KC, ((g.kdim - k0) + 3) & ~3);
            for (int kk = 0; kk < kend; kk += 4) {
                cplx af[2], bf[4];
#pragma unroll
                for (int mi = 0; mi < 2; mi++) af[mi] = Ps[(wm * 16 + mi * 8 + gid) * GM_PS + kk + tig];
#pragma unroll
                for (int ni = 0; ni < 4; ni++)
                    bf[ni] = WT ? Ws[(wn * 32 + ni * 8 + gid) * GM_PS + kk + tig]
                                : Ws[(kk + tig) * GM_WS + wn * 32 + ni * 8 + gid];
                // 16 independent accumulators are touched between two uses of the same one
#pragma unroll
                for (int mi = 0; mi < 2; mi++) {
#pragma unroll
                    for (int ni = 0; ni < 4; ni++) dmma884(cre[mi][ni][0], cre[mi][ni][1], af[mi].x, bf[ni].x);
#pragma unroll
                    for (int ni = 0; ni < 4; ni++) dmma884(cim[mi][ni][0], cim[mi][ni][1], af[mi].x, bf[ni].y);
                }
#pragma unroll
                for (int mi = 0; mi < 2; mi++) {
                    const double nim = -af[mi].y;
#pragma unroll
                    for (int ni = 0; ni < 4; ni++) dmma884(cre[mi][ni][0], cre[mi][ni][1], nim, bf[ni].y);
#pragma unroll
                    for (int ni = 0; ni < 4; ni++) dmma884(cim[mi][ni][0], cim[mi][ni][1], af[mi].y, bf[ni].x);
                }
            }
        }
    }
#pragma unroll
    for (int mi = 0; mi < 2; mi++)
#pragma unroll
        for (int ni = 0; ni < 4; ni++) {
            const int r = i0 + wm * 16 + mi * 8 + gid;
            const int c = j0 + wn * 32 + ni * 8 + tig * 2;
            if (r < g.ihi && !(r >= g.skip_lo && r < g.skip_hi)) {
                if (c < g.jhi) Cb[(long)r * g.ldc + c] = cmake(cre[mi][ni][0], cim[mi][ni][0]);
                if (c + 1 < g.jhi) Cb[(long)r * g.ldc + c + 1] = cmake(cre[mi][ni][1], cim[mi][ni][1]);
            }
        }
}

// ------------------------------------------------------------------------------------------
// Pipelined variant of the rank-K (K <= 32) update used by the elimination hot path.
// Persistent CTAs (2 per SM) walk the (matrix, row tile, col tile) space; per CTA tile 64x32,
// 8 warps as 4(M) x 2(N), warp tile 16x16.  While a tile is being multiplied, the P / W operands of
// the CTA's next tile stream into the other shared-memory stage with cp.async (LDGSTS, L2-only) and
// its C fragment is prefetched into a second register set, so global-memory latency is overlapped
// with DMMA issue inside each CTA instead of relying on CTA-level interleaving.
// ------------------------------------------------------------------------------------------
#define GP_BM 64
#define GP_BN 32
#define GP_PS (GM_KC + 4)      // 36 cplx: 576 B == 64 mod 128
#define GP_WS (GP_BN + 2)      // 34 cplx: 544 B == 32 mod 128
#define GP_STAGE (GP_BM * GP_PS + GM_KC * GP_WS)

__device__ __forceinline__ double flipsign(double x, unsigned mask) {
    return __hiloint2double(__double2hiint(x) ^ (int)mask, __double2loint(x));
}
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, int src_bytes) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" :: "r"(d), "l"(gsrc), "r"(src_bytes) : "memory");
}

__global__ void __launch_bounds__(256, 2) k_gemm_pipe(GnbGemmArgs g, int nti, int ntj, int total) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cplx* sm = reinterpret_cast<cplx*>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int gid = lane >> 2, tig = lane & 3;
    const int wm = warp >> 1, wn = warp & 1;
    const int per_mat = nti * ntj;
    const int nch = (g.kdim + GM_KC - 1) / GM_KC;               // K chunks of 32 = pipeline items per tile

    auto issue_loads = [&](int tile, int ch, int stage) {
        const int b = tile / per_mat, rem = tile - b * per_mat;
        const int ti = rem / ntj, tj = rem - ti * ntj;
        const int i0 = g.ilo + ti * GP_BM, j0 = g.jlo + tj * GP_BN, k0 = ch * GM_KC;
        const cplx* Pb = g.P + (long)b * g.strideP;
        const cplx* Wb = g.W + (long)b * g.strideW;
        cplx* Ps = sm + stage * GP_STAGE;
        cplx* Ws = Ps + GP_BM * GP_PS;
#pragma unroll
        for (int q = 0; q < (GP_BM * GM_KC) / 256; q++) {
            const int idx = tid + q * 256;
            const int r = idx / GM_KC, k = idx - r * GM_KC;
            const bool ok = (i0 + r < g.ihi) && (k0 + k < g.kdim);
            cp_async16(&Ps[r * GP_PS + k], ok ? (const void*)(Pb + (long)(i0 + r) * g.ldp + k0 + k) : (const void*)g.P, ok ? 16 : 0);
        }
#pragma unroll
        for (int q = 0; q < (GM_KC * GP_BN) / 256; q++) {
            const int idx = tid + q * 256;
            const int k = idx / GP_BN, n = idx - k * GP_BN;
            const bool ok = (k0 + k < g.kdim) && (j0 + n < g.jhi);
            cp_async16(&Ws[k * GP_WS + n], ok ? (const void*)(Wb + (long)(k0 + k) * g.ldw + j0 + n) : (const void*)g.W, ok ? 16 : 0);
        }
    };
    auto load_c = [&](int tile, double (&cr)[2][2][2], double (&ci)[2][2][2]) {
        const int b = tile / per_mat, rem = tile - b * per_mat;
        const int ti = rem / ntj, tj = rem - ti * ntj;
        const int i0 = g.ilo + ti * GP_BM, j0 = g.jlo + tj * GP_BN;
        const cplx* Cb = g.C + (long)b * g.strideC;
#pragma unroll
        for (int mi = 0; mi < 2; mi++)
#pragma unroll
            for (int ni = 0; ni < 2; ni++) {
                const int r = i0 + wm * 16 + mi * 8 + gid;
                const int c = j0 + wn * 16 + ni * 8 + tig * 2;
                cplx v0 = cmake(0.0, 0.0), v1 = cmake(0.0, 0.0);
                if (!g.zero_init && r < g.ihi) {
                    if (c < g.jhi) v0 = Cb[(long)r * g.ldc + c];
                    if (c + 1 < g.jhi) v1 = Cb[(long)r * g.ldc + c + 1];
                }
                cr[mi][ni][0] = v0.x; ci[mi][ni][0] = v0.y;
                cr[mi][ni][1] = v1.x; ci[mi][ni][1] = v1.y;
            }
    };

    int tile = blockIdx.x;
    if (tile >= total) return;
    double cre[2][2][2], cim[2][2][2], pre[2][2][2], pim[2][2][2];
    issue_loads(tile, 0, 0);
    asm volatile("cp.async.commit_group;" ::: "memory");
    load_c(tile, cre, cim);
    const unsigned smask = g.plus ? 0u : 0x80000000u;     // C -= P W : flip the sign bit of the A fragments
    int stage = 0;
    for (;;) {
        const int next = tile + gridDim.x;
        const bool has_next = next < total;
        for (int ch = 0; ch < nch; ch++) {
            // item (tile, ch) has landed for this thread; the barrier also certifies that every warp
            // is done with the other stage, which is refilled right away with the next item.
            asm volatile("cp.async.wait_group 0;" ::: "memory");
            __syncthreads();
            if (ch + 1 < nch) issue_loads(tile, ch + 1, stage ^ 1);
            else if (has_next) issue_loads(next, 0, stage ^ 1);
            asm volatile("cp.async.commit_group;" ::: "memory");
            if (ch == 0 && has_next) load_c(next, pre, pim);          // C fragment of the next tile -> registers
            const cplx* Ps = sm + stage * GP_STAGE;
            const cplx* Ws = Ps + GP_BM * GP_PS;
            const int kend = min(GM_KC, (g.kdim - ch * GM_KC + 3) & ~3);
#ifndef GP_VARIANT
#define GP_VARIANT 0
#endif
#if GP_VARIANT == 2 || GP_VARIANT == 3
#pragma unroll
#else
#pragma unroll 2
#endif
            for (int kk = 0; kk < GM_KC; kk += 4) {
                if (kk < kend) {
                cplx af[2], bf[2];
#pragma unroll
                for (int mi = 0; mi < 2; mi++) af[mi] = Ps[(wm * 16 + mi * 8 + gid) * GP_PS + kk + tig];
#pragma unroll
                for (int ni = 0; ni < 2; ni++) bf[ni] = Ws[(kk + tig) * GP_WS + wn * 16 + ni * 8 + gid];
#if GP_VARIANT == 1 || GP_VARIANT == 3
                double ax[2], ay[2], nay[2];
#pragma unroll
                for (int mi = 0; mi < 2; mi++) {
                    ax[mi] = flipsign(af[mi].x, smask); ay[mi] = flipsign(af[mi].y, smask);
                    nay[mi] = flipsign(ay[mi], 0x80000000u);
                }
#pragma unroll
                for (int mi = 0; mi < 2; mi++)
#pragma unroll
                    for (int ni = 0; ni < 2; ni++) dmma884(cre[mi][ni][0], cre[mi][ni][1], ax[mi], bf[ni].x);
#pragma unroll
                for (int mi = 0; mi < 2; mi++)
#pragma unroll
                    for (int ni = 0; ni < 2; ni++) dmma884(cim[mi][ni][0], cim[mi][ni][1], ax[mi], bf[ni].y);
#pragma unroll
                for (int mi = 0; mi < 2; mi++)
#pragma unroll
                    for (int ni = 0; ni < 2; ni++) dmma884(cre[mi][ni][0], cre[mi][ni][1], nay[mi], bf[ni].y);
#pragma unroll
                for (int mi = 0; mi < 2; mi++)
#pragma unroll
                    for (int ni = 0; ni < 2; ni++) dmma884(cim[mi][ni][0], cim[mi][ni][1], ay[mi], bf[ni].x);
#else
#pragma unroll
                for (int mi = 0; mi < 2; mi++) {
                    const double ax = flipsign(af[mi].x, smask), ay = flipsign(af[mi].y, smask);
                    const double nay = flipsign(ay, 0x80000000u);
#pragma unroll
                    for (int ni = 0; ni < 2; ni++) dmma884(cre[mi][ni][0], cre[mi][ni][1], ax, bf[ni].x);
#pragma unroll
                    for (int ni = 0; ni < 2; ni++) dmma884(cim[mi][ni][0], cim[mi][ni][1], ax, bf[ni].y);
#pragma unroll
                    for (int ni = 0; ni < 2; ni++) dmma884(cre[mi][ni][0], cre[mi][ni][1], nay, bf[ni].y);
#pragma unroll
                    for (int ni = 0; ni < 2; ni++) dmma884(cim[mi][ni][0], cim[mi][ni][1], ay, bf[ni].x);
                }
#endif
                }
            }
            stage ^= 1;
        }
        {
            const int b = tile / per_mat, rem = tile - b * per_mat;
            const int ti = rem / ntj, tj = rem - ti * ntj;
            const int i0 = g.ilo + ti * GP_BM, j0 = g.jlo + tj * GP_BN;
            cplx* Cb = g.C + (long)b * g.strideC;
#pragma unroll
            for (int mi = 0; mi < 2; mi++)
#pragma unroll
                for (int ni = 0; ni < 2; ni++) {
                    const int r = i0 + wm * 16 + mi * 8 + gid;
                    const int c = j0 + wn * 16 + ni * 8 + tig * 2;
                    if (r < g.ihi && !(r >= g.skip_lo && r < g.skip_hi)) {
                        if (c < g.jhi) Cb[(long)r * g.ldc + c] = cmake(cre[mi][ni][0], cim[mi][ni][0]);
                        if (c + 1 < g.jhi) Cb[(long)r * g.ldc + c + 1] = cmake(cre[mi][ni][1], cim[mi][ni][1]);
                    }
                }
        }
        if (!has_next) break;
#pragma unroll
        for (int mi = 0; mi < 2; mi++)
#pragma unroll
            for (int ni = 0; ni < 2; ni++) {
                cre[mi][ni][0] = pre[mi][ni][0]; cre[mi][ni][1] = pre[mi][ni][1];
                cim[mi][ni][0] = pim[mi][ni][0]; cim[mi][ni][1] = pim[mi][ni][1];
            }
        tile = next;
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
}

// ------------------------------------------------------------------------------------------
// Host-side launchers
// ------------------------------------------------------------------------------------------
static inline int cdiv_i(long a, long b) { return (int)((a + b - 1) / b); }

static int g_gemm_pipe = 1;    // 1: pipelined persistent kernel for the K <= 32 hot path
static int g_num_sms = 148;
void gnb_set_gemm_pipe(int on) { g_gemm_pipe = on; }
static const size_t kPipeSmem = (size_t)2 * GP_STAGE * sizeof(cplx);
static int g_gemm_bm = 32;     // rows per CTA of the rank-K update (64: 8 warps x 2 CTAs/SM, 32: 4 warps x 4 CTAs/SM)
void gnb_set_gemm_bm(int bm) { g_gemm_bm = (bm == 32) ? 32 : 64; }
static size_t gemm_smem(bool wt, int bm) {
    return (size_t)(bm * GM_PS + (wt ? GM_T * GM_PS : GM_KC * GM_WS)) * sizeof(cplx);
}
static const size_t kPsSmem = (size_t)(2 * GNB_NB * PS_TC + 2 * GNB_NB * GNB_NB) * sizeof(cplx);

cudaError_t gnb_kernels_init() {
    cudaError_t e;
#define GNB_SET_SMEM(WT_, BK_, BM_)                                                                      \
    if ((e = cudaFuncSetAttribute(k_gemm<WT_, BK_, BM_>, cudaFuncAttributeMaxDynamicSharedMemorySize,    \
                                  (int)gemm_smem(WT_, BM_)))) return e;
    if ((e = cudaFuncSetAttribute(k_gemm_pipe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kPipeSmem))) return e;
    { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev); }
    GNB_SET_SMEM(false, false, 64) GNB_SET_SMEM(false, true, 64) GNB_SET_SMEM(true, false, 64) GNB_SET_SMEM(true, true, 64)
    GNB_SET_SMEM(false, false, 32) GNB_SET_SMEM(false, true, 32) GNB_SET_SMEM(true, false, 32) GNB_SET_SMEM(true, true, 32)
    if ((e = cudaFuncSetAttribute(k_permute_solve, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kPsSmem))) return e;
    return cudaSuccess;
}

void gnb_launch_assemble(cudaStream_t st, int M, cplx* A, long strideA, int ld, int N, const cplx* F,
                         const cplx* S, const cplx* Sig0, const cplx* SigB, long strideSigB, const cplx* E,
                         const int* pi, int mixr) {
    if (M <= 0) return;
    dim3 grid(N, cdiv_i(M, ASM_EB));
    k_assemble<<<grid, 256, 0, st>>>(A, strideA, ld, N, F, S, Sig0, SigB, strideSigB, E, pi, mixr, M);
}

void gnb_launch_scatter_sub(cudaStream_t st, int M, cplx* A, long strideA, int ld, const int* inds, int nc,
                            const cplx* blk, long strideBlk, const int* map) {
    if (M <= 0 || nc <= 0) return;
    dim3 grid(min(cdiv_i((long)nc * nc, 256), 1024), M);
    k_scatter_sub<<<grid, 256, 0, st>>>(A, strideA, ld, inds, nc, blk, strideBlk, map);
}

void gnb_launch_set_aug(cudaStream_t st, int M, cplx* A, long strideA, int ld, int N, int xoff, const int* cols, int m,
                        const int* map) {
    if (M <= 0 || m <= 0) return;
    dim3 grid(min(cdiv_i((long)N * m, 256), 1024), M);
    k_set_aug<<<grid, 256, 0, st>>>(A, strideA, ld, N, xoff, cols, m, map);
}
void gnb_launch_pad_diag(cudaStream_t st, int M, cplx* A, long strideA, int ld, int N, int Np) {
    if (M <= 0 || Np <= N) return;
    k_pad_diag<<<M, 32, 0, st>>>(A, strideA, ld, N, Np);
}

void gnb_launch_gemm(cudaStream_t st, const GnbGemmArgs& g, int nbatch, bool wt, bool batchk) {
    const int ni = g.ihi - g.ilo, nj = g.jhi - g.jlo;
    if (ni <= 0 || nj <= 0 || g.kdim <= 0 || nbatch <= 0) return;
    if (g_gemm_pipe && !wt && !batchk && g.kdim <= 4 * GM_KC && !g.wscale && !g.Pr) {
        const int nti = cdiv_i(ni, GP_BM), ntj = cdiv_i(nj, GP_BN);
        const long total = (long)nbatch * nti * ntj;
        if (total < (1L << 31)) {
            const int grid = (int)std::min<long>(total, 2L * g_num_sms);
            k_gemm_pipe<<<grid, 256, kPipeSmem, st>>>(g, nti, ntj, (int)total);
            return;
        }
    }
    const int bm = g_gemm_bm;
    dim3 grid(cdiv_i(nj, GM_T), cdiv_i(ni, bm), batchk ? 1 : nbatch);
    const size_t sm = gemm_smem(wt, bm);
#define GNB_GO(WT_, BK_)                                                         \
    do {                                                                         \
        if (bm == 64) k_gemm<WT_, BK_, 64><<<grid, 256, sm, st>>>(g);            \
        else k_gemm<WT_, BK_, 32><<<grid, 128, sm, st>>>(g);                     \
    } while (0)
    if (wt) { if (batchk) GNB_GO(true, true); else GNB_GO(true, false); }
    else { if (batchk) GNB_GO(false, true); else GNB_GO(false, false); }
}

static int g_tourn_fp32 = 1;      // nominating (non-final) tournament rounds in single precision
void gnb_set_tourn_group(int g) { g_tourn_fp32 = g != 0; }
// Tournament pivoting of the 32-wide panel at column c0 (candidate rows [c0, N)); used by the recursive engine.
static int g_tournq_cplx_min_m = 400;
static int g_tourn_warp = 249;     // bit mask: warp-synchronous tournament kernels (k_tournw / k_tournq) where they apply, see below
void gnb_set_tourn_warp(int on) { g_tourn_warp = on; }
void gnb_set_tournq_cplx_min_m(int m) { g_tournq_cplx_min_m = m; }
long gnb_launch_tournament(cudaStream_t st, int M, int N, const cplx* A, long strideA, int ld, int c0, int w,
                           int* cand0, int* cand1, int cand_stride, cplx* LU, int* moves, int* perm, int perm_stride,
                           int* info, int real_panel, int mixr) {
    int n = N - c0;
    const int* cin = nullptr;
    int* cout = cand0;
    long launches = 0;
    for (;;) {
        // k_tournw: 256 rows per CTA; every nominating round, and the FP64 final round of real panels
        const bool f32 = g_tourn_fp32 != 0;
        const bool warp_ok = g_tourn_warp && w == GNB_NB;
        const int G = warp_ok ? 256 : 128;
        const int groups = cdiv_i(n, G);
        const int fin = groups == 1;
        // g_tourn_warp bits: 1 = nominating rounds of real panels, 2 = FP64 final round of real panels, 4 = FP32
        // nominating rounds of complex panels.  Measured on the N = 1024 T(E) step (tools/sweep.py, profiles/r02_sweeps.txt):
        // 1 -> +1.7 %; 2 -> -1.8 % (44 us against 51 us alone, but its 224 registers keep it from sharing an SM with the
        // other sub-batch's rank-K CTAs); 4 -> -4 % (252 registers, 2 CTAs per SM); 8 = warp-independent k_tournq for every
        // round of a real panel (takes precedence over 1 and 2): +3 %; 16 = four candidate rows per lane at level 0 of the
        // FP32 nominating rounds of k_tournq (3 sequential eliminations per 256-row group instead of 7): +1.8 %; 32 = the
        // final round takes its pivot ORDER from FP32 eliminations too (4 rows per lane: one elimination for <= 128 nominees)
        // and inverts the pivot block in FP64 (2 phases instead of 4): +1.1 % (N = 512: +2.9 %); 64 / 128 = nominating / final
        // rounds of COMPLEX panels on k_tournq (float2 rows, FP32 order + FP64 complex inverse) for batches >= 400.  Default: 249.
        const bool use_w = warp_ok && (fin ? (real_panel && (g_tourn_warp & 2))
                                           : (real_panel ? (g_tourn_warp & 1) != 0 : (f32 && (g_tourn_warp & 4))));
        // complex panels on the warp-independent kernel: measured +0.9 % on the complex-F T(E) step (625 matrices per stream),
        // -2.5 % on GrInt (148 per stream: one warp per 256-row group leaves the SMs with 4 warps each), hence the batch limit
        if (warp_ok && !real_panel && f32 && (g_tourn_warp & 64) && M >= g_tournq_cplx_min_m) {
            const int grp = cdiv_i(n, 256);
            const int total = M * grp;
            if (grp > 1 || !(g_tourn_warp & 128)) {
                if (grp > 1) {
                    k_tournq<TT<float>, false, 2, 5, 2><<<cdiv_i(total, 2), 64, 0, st>>>(A, strideA, ld, c0, c0, n, grp, total, cin, cand_stride,
                                                                                        cout, cand_stride, 0, LU, moves, perm, perm_stride, info, mixr);
                    launches++;
                    n = (grp - 1) * w + min(w, n - (grp - 1) * 256);
                    cin = cout;
                    cout = (cout == cand0) ? cand1 : cand0;
                    continue;
                }
                // fall through: FP64 final round on k_tourn (<= 256 rows are reduced by it in two launches)
            } else {
                k_tournq<TT<float>, false, 2, 5, 2><<<cdiv_i(total, 2), 64, 0, st>>>(A, strideA, ld, c0, c0, n, grp, total, cin, cand_stride,
                                                                                    cout, cand_stride, 1, LU, moves, perm, perm_stride, info, mixr);
                launches++;
                break;
            }
        }
        if (warp_ok && real_panel && (g_tourn_warp & 8)) {    // warp-independent kernel: one warp per 256-row group / matrix
            const int grp = cdiv_i(n, 256);
            const int total = M * grp;
            if (grp > 1) {
                if (f32 && (g_tourn_warp & 16))
                    k_tournq<TTR<float>, false, 4, 3, 4><<<cdiv_i(total, 4), 128, 0, st>>>(A, strideA, ld, c0, c0, n, grp, total, cin, cand_stride,
                                                                                        cout, cand_stride, 0, LU, moves, perm, perm_stride, info, mixr);
                else if (f32)
                    k_tournq<TTR<float>, false, 4, 4, 2><<<cdiv_i(total, 4), 128, 0, st>>>(A, strideA, ld, c0, c0, n, grp, total, cin, cand_stride,
                                                                                        cout, cand_stride, 0, LU, moves, perm, perm_stride, info, mixr);
                else
                    k_tournq<TTR<double>, true, 2, 4, 2><<<cdiv_i(total, 2), 64, 0, st>>>(A, strideA, ld, c0, c0, n, grp, total, cin, cand_stride,
                                                                                       cout, cand_stride, 0, LU, moves, perm, perm_stride, info, mixr);
            } else if (f32 && (g_tourn_warp & 32)) {        // final round: FP32 pivot order (4 rows per lane), FP64 inverse
                k_tournq<TTR<float>, false, 2, 6, 4><<<cdiv_i(total, 2), 64, 0, st>>>(A, strideA, ld, c0, c0, n, grp, total, cin, cand_stride,
                                                                                     cout, cand_stride, 1, LU, moves, perm, perm_stride, info, mixr);
            } else {
                k_tournq<TTR<double>, true, 2, 4, 2><<<cdiv_i(total, 2), 64, 0, st>>>(A, strideA, ld, c0, c0, n, grp, total, cin, cand_stride,
                                                                                   cout, cand_stride, 1, LU, moves, perm, perm_stride, info, mixr);
            }
            launches++;
            if (grp == 1) break;
            n = (grp - 1) * w + min(w, n - (grp - 1) * 256);
            cin = cout;
            cout = (cout == cand0) ? cand1 : cand0;
            continue;
        }
        const int Gk = use_w ? 256 : 128;
        const int grp = cdiv_i(n, Gk);
        const int fink = grp == 1;
        dim3 grid(grp, M);
#define GNB_TOURN(T_, F64_, MINB_, FIN_)                                                                             \
    k_tourn<128, T_, F64_, MINB_><<<grid, 128, 0, st>>>(A, strideA, ld, c0, w, c0, n, cin, cand_stride, cout, cand_stride, \
                                                        FIN_, LU, moves, perm, perm_stride, info, mixr)
#define GNB_TOURNW(T_, F64_, MINB_, FIN_)                                                                            \
    k_tournw<T_, F64_, MINB_><<<grid, 128, 0, st>>>(A, strideA, ld, c0, c0, n, cin, cand_stride, cout, cand_stride, FIN_, LU, \
                                                    moves, perm, perm_stride, info, mixr)
        if (use_w) {
            if (!fink && f32) { if (real_panel) GNB_TOURNW(TTR<float>, false, 4, 0); else GNB_TOURNW(TT<float>, false, 2, 0); }
            else GNB_TOURNW(TTR<double>, true, 2, fink);      // real panels only (see use_w)
        } else if (!fink && f32) {
            if (real_panel) GNB_TOURN(TTR<float>, false, 8, 0);
            else GNB_TOURN(TT<float>, false, 5, 0);
        } else {
            if (real_panel) GNB_TOURN(TTR<double>, true, 5, fink);
            else GNB_TOURN(TT<double>, true, 3, fink);
        }
#undef GNB_TOURN
#undef GNB_TOURNW
        launches++;
        if (fink) break;
        n = (grp - 1) * w + min(w, n - (grp - 1) * Gk);
        cin = cout;
        cout = (cout == cand0) ? cand1 : cand0;
        (void)fin; (void)groups; (void)G;
    }
    return launches;
}
void gnb_launch_init_perm(cudaStream_t st, int M, int* perm, int stride, int N) {
    dim3 grid(cdiv_i(N, 256), M);
    k_init_perm<<<grid, 256, 0, st>>>(perm, stride, N);
}

// Block elimination of a batch of M matrices  [A | B]  (N x (N + naug), leading dimension ld).
//   jordan = 1 : A <- (P A)^-1 in place, perm[pos] = original row now at pos  (A^-1[:, perm[pos]] = stored[:, pos])
//   jordan = 0 : forward elimination + back-substitution; the naug augmented columns end up holding A^-1 B
// Two-level blocking: pivoting and the pivot-row products advance in 32-column inner blocks, but
// only the 64 "near" columns of an outer step are updated block by block; every other ("far")
// column receives ONE rank-64 update per outer step, which halves the C traffic of the dominant
// kernel (tools/proto_blockgj.py: two_level_jordan / two_level_forward are the numpy models).
long gnb_launch_tournament(cudaStream_t st, int M, int N, const cplx* A, long strideA, int ld, int c0, int w,
                           int* cand0, int* cand1, int cand_stride, cplx* LU, int* moves, int* perm, int perm_stride,
                           int* info, int real_panel, int mixr);
static int g_two_level = 1;
void gnb_set_two_level(int on) { g_two_level = on; }

namespace {
struct Elim {
    cudaStream_t st; int M, N, naug; cplx* A; long strideA; int ld; int jordan;
    const GnbElimWork& ws; long launches;

    cplx* lu(int slot) const { return ws.LU + (long)slot * M * GNB_NB * GNB_NB; }
    int* mv(int slot) const { return ws.moves + (long)slot * M * GNB_MOVES_STRIDE; }

    void tournament(int c0, int w, int slot) {
        launches += gnb_launch_tournament(st, M, N, A, strideA, ld, c0, w, ws.cand0, ws.cand1, ws.cand_stride, lu(slot),
                                          mv(slot), jordan ? ws.perm : nullptr, ws.perm_stride, ws.info, 0, 0);
    }
    void permute_solve(cplx* buf, long stride, int bld, int c0, int w, int lo, int hi, int slot, int mode,
                       const cplx* preL = nullptr, long strideL = 0, int ldL = 0, int pre_row = 0, int pre_k = 0) {
        if (hi <= lo) return;
        dim3 grid(cdiv_i(hi - lo, PS_TC), M);
        k_permute_solve<<<grid, 256, kPsSmem, st>>>(buf, stride, bld, c0, w, lo, hi, mv(slot), lu(slot), mode, preL,
                                                     strideL, ldL, pre_row, pre_k);
        launches++;
    }
    void save_panel(int c0, int w, int pld, int pcol) {
        dim3 grid(min(cdiv_i((long)N * GNB_NB, 256), 1024), M);
        k_save_panel<<<grid, 256, 0, st>>>(A, strideA, ld, N, c0, w, ws.Pws, (long)N * pld, pld, pcol);
        launches++;
    }
    void gemm(int ilo, int ihi, int jlo, int jhi, const cplx* P, long strideP, int ldp, const cplx* W, int kdim,
              int skip_lo, int skip_hi) {
        if (ihi <= ilo || jhi <= jlo || kdim <= 0) return;
        GnbGemmArgs g{};
        g.C = A; g.strideC = strideA; g.ldc = ld;
        g.P = P; g.strideP = strideP; g.ldp = ldp;
        g.W = W; g.strideW = strideA; g.ldw = ld;
        g.ilo = ilo; g.ihi = ihi; g.jlo = jlo; g.jhi = jhi; g.kdim = kdim;
        g.skip_lo = skip_lo; g.skip_hi = skip_hi; g.zero_init = 0; g.plus = 0; g.wscale = nullptr; g.nbatch_k = 0;
        if (ws.timer) ws.timer->begin(st);
        gnb_launch_gemm(st, g, M, false, false);
        if (ws.timer) ws.timer->end(st, 8.0 * (double)(ihi - ilo) * (double)(jhi - jlo) * kdim * M);
        launches++;
    }

    void single_step(int c0, int w) {
        tournament(c0, w, 0);
        if (jordan) {
            permute_solve(A, strideA, ld, c0, w, 0, N, 0, 1);
            save_panel(c0, w, GNB_NB, 0);
            gemm(0, N, 0, N, ws.Pws, (long)N * GNB_NB, GNB_NB, A + (long)c0 * ld, w, c0, c0 + w);
        } else {
            permute_solve(A, strideA, ld, c0, w, c0, N + naug, 0, 0);
            gemm(c0 + w, N, c0 + w, N + naug, A + c0, strideA, ld, A + (long)c0 * ld, w, -1, -1);
        }
    }

    void double_step(int c0, int wb) {
        const int wa = GNB_NB, cb = c0 + GNB_NB, near_hi = cb + wb;
        const cplx* Wa = A + (long)c0 * ld;                 // rows of inner block a (and b below it)
        const cplx* Wb = A + (long)cb * ld;
        if (!jordan) {
            tournament(c0, wa, 0);
            permute_solve(A, strideA, ld, c0, wa, c0, near_hi, 0, 0);
            gemm(cb, N, cb, near_hi, A + c0, strideA, ld, Wa, wa, -1, -1);
            tournament(cb, wb, 1);
            permute_solve(A, strideA, ld, cb, wb, c0, near_hi, 1, 2);
            const int far_lo = near_hi, far_hi = N + naug;
            permute_solve(A, strideA, ld, c0, wa, far_lo, far_hi, 0, 0);
            permute_solve(A, strideA, ld, cb, wb, far_lo, far_hi, 1, 0, A + (long)cb * ld + c0, strideA, ld, c0, wa);
            gemm(near_hi, N, far_lo, far_hi, A + c0, strideA, ld, Wa, wa + wb, -1, -1);
            return;
        }
        const long sP = (long)N * 64;
        tournament(c0, wa, 0);
        permute_solve(A, strideA, ld, c0, wa, c0, near_hi, 0, 1);
        save_panel(c0, wa, 64, 0);
        gemm(0, N, c0, near_hi, ws.Pws, sP, 64, Wa, wa, c0, cb);
        tournament(cb, wb, 1);
        permute_solve(A, strideA, ld, cb, wb, c0, near_hi, 1, 1);
        permute_solve(ws.Pws, sP, 64, 0, 0, 0, GNB_NB, 1, 2);         // the saved panel follows the row moves
        save_panel(cb, wb, 64, GNB_NB);
        gemm(0, N, c0, near_hi, ws.Pws + GNB_NB, sP, 64, Wb, wb, cb, near_hi);
        const int lo[2] = {0, near_hi}, hi[2] = {c0, N};
        for (int r = 0; r < 2; r++) {
            if (hi[r] <= lo[r]) continue;
            permute_solve(A, strideA, ld, c0, wa, lo[r], hi[r], 0, 1);
            permute_solve(A, strideA, ld, cb, wb, lo[r], hi[r], 1, 1, ws.Pws + (long)cb * 64, sP, 64, c0, wa);
            gemm(0, N, lo[r], hi[r], ws.Pws, sP, 64, Wa, wa + wb, c0, near_hi);           // rows outside both blocks
            gemm(c0, cb, lo[r], hi[r], ws.Pws + GNB_NB, sP, 64, Wb, wb, -1, -1);           // then W_a -= W_a[:,Kb] W_b
        }
    }
};
}  // namespace

long gnb_eliminate(cudaStream_t st, int M, int N, int naug, cplx* A, long strideA, int ld, int jordan,
                   const GnbElimWork& ws) {
    if (M <= 0) return 0;
    Elim e{st, M, N, naug, A, strideA, ld, jordan, ws, 0};
    if (jordan) {
        dim3 grid(cdiv_i(N, 256), M);
        k_init_perm<<<grid, 256, 0, st>>>(ws.perm, ws.perm_stride, N);
        e.launches++;
    }
    int c0 = 0;
    while (c0 < N) {
        const int rest = N - c0;
        if (g_two_level && rest > GNB_NB) {
            e.double_step(c0, min(GNB_NB, rest - GNB_NB));
            c0 += 2 * GNB_NB;
        } else {
            e.single_step(c0, min(GNB_NB, rest));
            c0 += GNB_NB;
        }
    }
    if (!jordan && naug > 0) {
        // X[0:c0,:] -= Wstored[0:c0, K] X[K,:]  from the last block upwards (unit block upper triangular)
        const int nblk = (N + GNB_NB - 1) / GNB_NB;
        for (int blk = nblk - 1; blk >= 1; blk--) {
            const int b0 = blk * GNB_NB, w = min(GNB_NB, N - b0);
            GnbGemmArgs g{};
            g.C = A + N; g.strideC = strideA; g.ldc = ld;
            g.P = A + b0; g.strideP = strideA; g.ldp = ld;
            g.W = A + (long)b0 * ld + N; g.strideW = strideA; g.ldw = ld;
            g.ilo = 0; g.ihi = b0; g.jlo = 0; g.jhi = naug; g.kdim = w;
            g.skip_lo = g.skip_hi = -1; g.zero_init = 0; g.plus = 0; g.wscale = nullptr;
            if (ws.timer) ws.timer->begin(st);
            gnb_launch_gemm(st, g, M, false, false);
            if (ws.timer) ws.timer->end(st, 8.0 * (double)b0 * (double)naug * g.kdim * M);
            e.launches++;
        }
    }
    return e.launches;
}
