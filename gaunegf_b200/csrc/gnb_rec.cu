// Recursive (multi-level) batched complex128 block elimination for sm_100a (B200).
//
// Same mathematics as gnb_elim.cu (tournament-pivoted block Gauss-Jordan / block Gaussian elimination of the
// reference's `solve(A, I)`, utils.py:52-54, integrate.py:67-82, transport.py:150-157) but organised as a
// binary recursion over column ranges so that the floating-point work sits in rank-K updates with K up to
// N/2 instead of K = 32/64:  at N = 1024 more than 85 % of the update flops run at K >= 128
// (tools/proto_recursive.py is the numpy model, kernel by kernel).
//
// The rank-K update  C -= P W  is a warp-specialised persistent kernel on the FP64 tensor pipe
// (mma.sync.m8n8k4.f64 -> SASS DMMA.8x8x4; tcgen05 has no f64 kind):
//   * operands are kept PACKED in global memory as ready-made shared-memory images
//       Ppk[b][kc][rb][32][RK_PPS]   saved panel columns  (kc = 16-wide K chunk, rb = 32-row block)
//       Wpk[b][kc][cb][16][RK_WPS]   normalised pivot rows (cb = 32-column block)
//     so ONE producer lane feeds a 4-deep ring with two cp.async.bulk (TMA) copies per K chunk, completion
//     on mbarriers (complete_tx); the row strides 20 / 34 make every LDS.128 fragment load conflict-free;
//   * 8 consumer warps (4 x 2, warp tile 16 x 32) run LDS + DMMA with no CTA-wide barrier, accumulate from
//     zero and read-modify-write C in the epilogue while the producer already streams the next tile;
//   * optional 3M arithmetic (3 real DMMAs per complex tile product instead of 4).
// One CTA per SM (persistent over the (matrix, row tile, column tile) space).
#include <algorithm>
#include <cstring>
#include "gnb_common.cuh"
#include "gnb_kernels.h"

static inline int cdiv_i(long a, long b) { return (int)((a + b - 1) / b); }

// Mixed layout (transmission path with the real-structure shortcut): the first mixr columns of every row are
// stored as real doubles, the others as complex128.  `A` is always the LOGICAL complex base, i.e.
// A[row * ld + col] is the element for col >= mixr; the real view of the same rows starts mixr * 8 bytes later
// and has a row stride of 2 * ld doubles.  mixr is a multiple of 32, so 32-column tiles never straddle.
__device__ __forceinline__ double* rk_real_view(cplx* A, int mixr) {
    return reinterpret_cast<double*>(reinterpret_cast<char*>(A) + (size_t)mixr * 8);
}
__device__ __forceinline__ const double* rk_real_view(const cplx* A, int mixr) {
    return reinterpret_cast<const double*>(reinterpret_cast<const char*>(A) + (size_t)mixr * 8);
}

// ------------------------------------------------------------------------------------------------
// mbarrier / bulk-copy helpers
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
        "RK_WAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra RK_DONE;\n\t"
        "bra RK_WAIT;\n\t"
        "RK_DONE:\n\t}" :: "r"(smem_u32(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, int bytes, unsigned long long* b) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(b)) : "memory");
}

// bulk copy with an L2 evict-last policy: the packed operands are what the tiles of a matrix share
__device__ __forceinline__ unsigned long long l2_evict_last_policy() {
    unsigned long long pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ void bulk_g2s_keep(void* dst, const void* src, int bytes, unsigned long long* b, unsigned long long pol) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                 :: "r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(b)), "l"(pol) : "memory");
}
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" :: "l"(p)); }
// C is read once and written once per launch and nothing re-reads it before it has left the L2 (the working set of a
// chunk is ~100 x the L2): streaming (evict-first) accesses keep it from displacing the packed operands, which every
// row / column tile of a matrix re-reads (ncu: DRAM reads of the mid-K launches were 1.6 x their algorithmic bytes)
__device__ __forceinline__ double2 rk_ldc(const double2* p, int cs) { return cs ? __ldcs(p) : *p; }
__device__ __forceinline__ void rk_stc(double2* p, double2 v, int cs) { if (cs) __stcs(p, v); else *p = v; }

// ------------------------------------------------------------------------------------------------
// Rank-K update on packed operands:  C[ilo:ihi, jlo:jhi] -= P[ilo:ihi, klo:khi] * W[klo:khi, jlo:jhi]
// All bounds are multiples of 32.  CTA tile 64 x 64 anchored at (ilo, jlo); a trailing half tile is masked.
// kskip != 0 (JORDAN far update): for row tiles inside [klo, khi) the panel is zero at and below the block
// diagonal, so the K loop of such a tile starts at the chunk after the tile's first row block.
// ------------------------------------------------------------------------------------------------
struct RkGemmArgs {
    cplx* C; long sC; int ldc;
    const cplx* P; long sP; int nrb;
    const cplx* W; long sW; int ncb;
    int ilo, ihi, jlo, jhi, klo, khi;
    int kskip;
    int preal;      // the panel operand P is real (imaginary parts exactly zero)
    int nreal;      // columns [0, nreal) of C / W are real (real F, S, E with the contact orbitals ordered last)
    int mixr;       // columns [0, mixr) of C are STORED as real doubles (launches never straddle mixr)
    int cs;         // streaming (evict-first) accesses to C
};

#define RK_ST 4
template <int RB, int CB> constexpr size_t rk_smem() {
    return (size_t)RK_ST * (RB * RK_PBLK + CB * RK_WBLK) * sizeof(cplx) + 2 * RK_ST * 8;
}

// CTA tile = (32 RB) x (32 CB), RB * CB = 4: 64 x 64 for the general case, 128 x 32 for 32-column strips.
template <int M3, int RB, int CB>
__global__ void __launch_bounds__(288, 1) k_rk_gemm(RkGemmArgs g, int nti, int ntj, int total, int g0, int tcap) {
    static_assert(RB * CB == 4, "8 consumer warps of 16 x 32");
    constexpr int NCW = 8, WN = CB, MI = 2, NI = 4, KC = 16;
    constexpr int TM = 32 * RB, TN = 32 * CB;
    constexpr int RK_PST = RB * RK_PBLK, RK_WST = CB * RK_WBLK, RK_STAGE = RK_PST + RK_WST;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    cplx* sm = reinterpret_cast<cplx*>(smem_raw);
    unsigned long long* full = reinterpret_cast<unsigned long long*>(sm + RK_ST * RK_STAGE);
    unsigned long long* empty = full + RK_ST;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int per_mat = nti * ntj;
    // CTA c works on tiles first, first + g0, ... (at most tcap of them): g0 = CTAs resident at a time, so a wave of CTAs
    // walks a contiguous window of tiles; tcap bounds the CTA's lifetime (short-lived CTAs let the high-priority panel
    // kernels of the other sub-batch in), tcap >= total / g0 is the persistent kernel
    const int first = ((int)blockIdx.x / g0) * g0 * tcap + (int)blockIdx.x % g0;
    const int my_tiles = first < total ? min(tcap, (total - first + g0 - 1) / g0) : 0;
    if (tid == 0) {
        for (int s = 0; s < RK_ST; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], NCW); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int kc0 = g.klo / KC, nch_all = (g.khi - g.klo) / KC;

    if (warp == NCW) {
        // ---------------- producer: one lane, two bulk copies per K chunk ----------------
        if (lane != 0) return;
        const unsigned long long keep = l2_evict_last_policy();
        int q = 0;
        for (int it = 0; it < my_tiles; it++) {
            const int tile = first + it * g0;
            const int b = tile / per_mat, rem = tile - b * per_mat;
            const int ti = rem / ntj, tj = rem - ti * ntj;
            const int i0 = g.ilo + ti * TM, j0 = g.jlo + tj * TN;
            int ch0 = 0;
            if (g.kskip && i0 >= g.klo && i0 + TM <= g.khi) ch0 = (i0 + 32 - g.klo) / KC;   // panel is zero up to the block diagonal
            const cplx* Pb = g.P + (long)b * g.sP + ((long)kc0 * g.nrb + i0 / 32) * RK_PBLK;
            const cplx* Wb = g.W + (long)b * g.sW + ((long)kc0 * g.ncb + j0 / 32) * RK_WBLK;
            for (int ch = ch0; ch < nch_all; ch++, q++) {
                const int s = q % RK_ST;
                if (q >= RK_ST) mbar_wait(&empty[s], ((q / RK_ST) - 1) & 1);
                cplx* Ps = sm + s * RK_STAGE;
                mbar_expect_tx(&full[s], RK_STAGE * 16);
                if (g.cs & 2) {
                    bulk_g2s_keep(Ps, Pb + (long)ch * g.nrb * RK_PBLK, RK_PST * 16, &full[s], keep);
                    bulk_g2s_keep(Ps + RK_PST, Wb + (long)ch * g.ncb * RK_WBLK, RK_WST * 16, &full[s], keep);
                } else {
                    bulk_g2s(Ps, Pb + (long)ch * g.nrb * RK_PBLK, RK_PST * 16, &full[s]);
                    bulk_g2s(Ps + RK_PST, Wb + (long)ch * g.ncb * RK_WBLK, RK_WST * 16, &full[s]);
                }
            }
        }
        return;
    }
    // ---------------- consumers ----------------
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
        const int tile = first + it * g0;
        const int b = tile / per_mat, rem = tile - b * per_mat;
        const int ti = rem / ntj, tj = rem - ti * ntj;
        const int i0 = g.ilo + ti * TM, j0 = g.jlo + tj * TN;
        int ch0 = 0;
        if (g.kskip && i0 >= g.klo && i0 + TM <= g.khi) ch0 = (i0 + 32 - g.klo) / KC;
        const bool active = (wm * 16 < g.ihi - i0) && (wn * 32 < g.jhi - j0);     // warp tile inside the range
        const int mode = !g.preal ? 3 : (j0 + TN <= g.nreal ? 1 : 2);              // 1: real x real, 2: real x complex
        if ((g.cs & 4) && active && ch0 < nch_all && j0 >= g.mixr) {               // C tile on its way to L2 while the K loop runs
            const cplx* Cp = g.C + (long)b * g.sC + (long)(i0 + wm * 16 + gid) * g.ldc + j0 + wn * 32 + tig * 2;
#pragma unroll
            for (int mi = 0; mi < MI; mi++)
#pragma unroll
                for (int ni = 0; ni < NI; ni++) prefetch_l2(Cp + (long)(mi * 8) * g.ldc + ni * 8);
        }
        for (int ch = ch0; ch < nch_all; ch++, q++) {
            const int s = q % RK_ST;
            mbar_wait(&full[s], (q / RK_ST) & 1);
            if (active) {
                const cplx* Ps = sm + s * RK_STAGE + (wm * 16 + gid) * RK_PPS + tig;
                const cplx* Ws = sm + s * RK_STAGE + RK_PST + wn * RK_WBLK + tig * RK_WPS + gid;
                if (mode == 1) {
                    // real P, real W: one real DMMA per tile product, the imaginary part of C stays exactly zero
#pragma unroll
                    for (int kk = 0; kk < KC; kk += 4) {
                        double af[MI], bf[NI];
#pragma unroll
                        for (int mi = 0; mi < MI; mi++) af[mi] = Ps[(mi * 8) * RK_PPS + kk].x;
#pragma unroll
                        for (int ni = 0; ni < NI; ni++) bf[ni] = Ws[kk * RK_WPS + ni * 8].x;
#pragma unroll
                        for (int mi = 0; mi < MI; mi++)
#pragma unroll
                            for (int ni = 0; ni < NI; ni++) dmma884(cre[mi][ni][0], cre[mi][ni][1], af[mi], bf[ni]);
                    }
                } else if (mode == 2) {
                    // real P, complex W: re += ar br, im += ar bi
#pragma unroll
                    for (int kk = 0; kk < KC; kk += 4) {
                        double af[MI];
                        cplx bf[NI];
#pragma unroll
                        for (int mi = 0; mi < MI; mi++) af[mi] = Ps[(mi * 8) * RK_PPS + kk].x;
#pragma unroll
                        for (int ni = 0; ni < NI; ni++) bf[ni] = Ws[kk * RK_WPS + ni * 8];
#pragma unroll
                        for (int mi = 0; mi < MI; mi++) {
#pragma unroll
                            for (int ni = 0; ni < NI; ni++) dmma884(cre[mi][ni][0], cre[mi][ni][1], af[mi], bf[ni].x);
#pragma unroll
                            for (int ni = 0; ni < NI; ni++) dmma884(cim[mi][ni][0], cim[mi][ni][1], af[mi], bf[ni].y);
                        }
                    }
                } else {
#pragma unroll
                for (int kk = 0; kk < KC; kk += 4) {
                    cplx af[MI], bf[NI];
#pragma unroll
                    for (int mi = 0; mi < MI; mi++) af[mi] = Ps[(mi * 8) * RK_PPS + kk];
#pragma unroll
                    for (int ni = 0; ni < NI; ni++) bf[ni] = Ws[kk * RK_WPS + ni * 8];
                    if (M3) {
                        // 3M: X = ar br, Y = ai bi, Z = (ar + ai)(br + bi);  re = X - Y, im = Z - X - Y
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
                }
            }
            // generic-proxy reads of this stage are ordered before the producer's next async-proxy write
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[s]);
        }
        if (active && ch0 < nch_all) {
            if (j0 < g.mixr) {
                // real-stored C (mode 1 by construction): 16-byte read-modify-write of two adjacent doubles
                double* Cr = rk_real_view(g.C + (long)b * g.sC, g.mixr) + (long)(i0 + wm * 16 + gid) * 2 * g.ldc + j0 + wn * 32 + tig * 2;
#pragma unroll
                for (int mi = 0; mi < MI; mi++) {
                    double2 v[NI];
#pragma unroll
                    for (int ni = 0; ni < NI; ni++) v[ni] = rk_ldc(reinterpret_cast<const double2*>(Cr + (long)(mi * 8) * 2 * g.ldc + ni * 8), g.cs);
#pragma unroll
                    for (int ni = 0; ni < NI; ni++) {
                        v[ni].x -= cre[mi][ni][0]; v[ni].y -= cre[mi][ni][1];
                        cre[mi][ni][0] = cre[mi][ni][1] = 0.0;
                        rk_stc(reinterpret_cast<double2*>(Cr + (long)(mi * 8) * 2 * g.ldc + ni * 8), v[ni], g.cs);
                    }
                }
            } else {
            cplx* Cb = g.C + (long)b * g.sC + (long)(i0 + wm * 16 + gid) * g.ldc + j0 + wn * 32 + tig * 2;
#pragma unroll
            for (int mi = 0; mi < MI; mi++) {
                cplx v[NI][2];
#pragma unroll
                for (int ni = 0; ni < NI; ni++) {
                    v[ni][0] = rk_ldc(&Cb[(long)(mi * 8) * g.ldc + ni * 8], g.cs);
                    v[ni][1] = rk_ldc(&Cb[(long)(mi * 8) * g.ldc + ni * 8 + 1], g.cs);
                }
#pragma unroll
                for (int ni = 0; ni < NI; ni++) {
#pragma unroll
                    for (int e = 0; e < 2; e++) {
                        double re, im;
                        if (M3 && mode == 3) { re = cre[mi][ni][e] - cim[mi][ni][e]; im = c3[mi][ni][e] - cre[mi][ni][e] - cim[mi][ni][e]; }
                        else { re = cre[mi][ni][e]; im = cim[mi][ni][e]; }
                        v[ni][e].x -= re; v[ni][e].y -= im;
                        cre[mi][ni][e] = 0.0; cim[mi][ni][e] = 0.0;
                        if (M3) c3[mi][ni][e] = 0.0;
                    }
                    rk_stc(&Cb[(long)(mi * 8) * g.ldc + ni * 8], v[ni][0], g.cs);
                    rk_stc(&Cb[(long)(mi * 8) * g.ldc + ni * 8 + 1], v[ni][1], g.cs);
                }
            }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Rank-K update with a REAL-PACKED panel (mixed layout, panels left of mixr):
//   PpkR[b][kc][rb][32][RK_PRS] doubles;   WR = 1: W real-packed WpkR[b][kc][cb][16][RK_WRS] doubles, C real-stored;
//   WR = 0: W complex-packed (columns right of mixr: contact orbitals and augmented columns), C complex.
// Same warp-specialised structure as k_rk_gemm; one real DMMA per tile product (WR = 1) or two (WR = 0).
// Row strides 20 / 36 doubles keep the LDS.64 fragment loads conflict-free.
// ------------------------------------------------------------------------------------------------
// Column-tile order of a launch whose right half of column tiles is cheap (real-valued augmented columns, see jre): cheap
// and expensive tiles alternate in the tile index, and the host makes the CTA stride odd, so every persistent CTA works
// on both kinds in turn (with the natural order and a stride that is a multiple of ntj a CTA sees ONE column tile only:
// half of the CTAs would finish in half the time).
__device__ __forceinline__ int rk_col_tile(int tjp, int ntj, bool interleave) {
    return interleave ? ((tjp & 1) ? (ntj >> 1) + (tjp >> 1) : (tjp >> 1)) : tjp;
}
struct RkGemmRpArgs {
    cplx* C; long sC; int ldc;
    const double* P; long sP; int nrb;
    const void* W; long sW; int ncb;        // doubles (WR = 1) or cplx (WR = 0); sW in elements of that type
    int ilo, ihi, jlo, jhi, klo, khi;
    int mixr;
    int cs;         // streaming (evict-first) accesses to C
    int jint;       // interleave cheap / expensive column tiles (rk_col_tile)
    int jre;        // WR = 0: columns >= jre hold REAL values in complex storage (the augmented unit columns stay real
                    // while the pivot blocks are real: real multipliers, real pivot-block inverses) - no imaginary DMMAs
};
template <int RB, int CB, int WR> constexpr size_t rk_rp_smem() {
    return (size_t)RK_ST * (RB * RK_PRBLK * 8 + CB * (WR ? RK_WRBLK * 8 : RK_WBLK * 16)) + 2 * RK_ST * 8;
}

template <int RB, int CB, int WR>
__global__ void __launch_bounds__(288, 2) k_rk_gemm_rp(RkGemmRpArgs g, int nti, int ntj, int total, int g0, int tcap) {
    static_assert(RB * CB == 4, "8 consumer warps of 16 x 32");
    constexpr int NCW = 8, WN = CB, MI = 2, NI = 4, KC = 16;
    constexpr int TM = 32 * RB, TN = 32 * CB;
    constexpr int PBYTES = RB * RK_PRBLK * 8, WBYTES = CB * (WR ? RK_WRBLK * 8 : RK_WBLK * 16), SBYTES = PBYTES + WBYTES;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned long long* full = reinterpret_cast<unsigned long long*>(smem_raw + RK_ST * SBYTES);
    unsigned long long* empty = full + RK_ST;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int per_mat = nti * ntj;
    // CTA c works on tiles first, first + g0, ... (at most tcap of them): g0 = CTAs resident at a time, so a wave of CTAs
    // walks a contiguous window of tiles; tcap bounds the CTA's lifetime (short-lived CTAs let the high-priority panel
    // kernels of the other sub-batch in), tcap >= total / g0 is the persistent kernel
    const int first = ((int)blockIdx.x / g0) * g0 * tcap + (int)blockIdx.x % g0;
    const int my_tiles = first < total ? min(tcap, (total - first + g0 - 1) / g0) : 0;
    if (tid == 0) {
        for (int s = 0; s < RK_ST; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], NCW); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int kc0 = g.klo / KC, nch_all = (g.khi - g.klo) / KC;

    if (warp == NCW) {
        if (lane != 0) return;
        const unsigned long long keep = l2_evict_last_policy();
        int q = 0;
        for (int it = 0; it < my_tiles; it++) {
            const int tile = first + it * g0;
            const int b = tile / per_mat, rem = tile - b * per_mat;
            const int ti = rem / ntj, tj = rk_col_tile(rem - ti * ntj, ntj, !WR && g.jint);
            const int i0 = g.ilo + ti * TM, j0 = g.jlo + tj * TN;
            const double* Pb = g.P + (long)b * g.sP + ((long)kc0 * g.nrb + i0 / 32) * RK_PRBLK;
            const char* Wb = WR ? reinterpret_cast<const char*>(reinterpret_cast<const double*>(g.W) + (long)b * g.sW +
                                                                ((long)kc0 * g.ncb + j0 / 32) * RK_WRBLK)
                                : reinterpret_cast<const char*>(reinterpret_cast<const cplx*>(g.W) + (long)b * g.sW +
                                                                ((long)kc0 * g.ncb + j0 / 32) * RK_WBLK);
            const long wstep = (long)g.ncb * (WR ? RK_WRBLK * 8 : RK_WBLK * 16);
            for (int ch = 0; ch < nch_all; ch++, q++) {
                const int s = q % RK_ST;
                if (q >= RK_ST) mbar_wait(&empty[s], ((q / RK_ST) - 1) & 1);
                unsigned char* st = smem_raw + s * SBYTES;
                mbar_expect_tx(&full[s], SBYTES);
                if (g.cs & 2) {
                    bulk_g2s_keep(st, Pb + (long)ch * g.nrb * RK_PRBLK, PBYTES, &full[s], keep);
                    bulk_g2s_keep(st + PBYTES, Wb + ch * wstep, WBYTES, &full[s], keep);
                } else {
                    bulk_g2s(st, Pb + (long)ch * g.nrb * RK_PRBLK, PBYTES, &full[s]);
                    bulk_g2s(st + PBYTES, Wb + ch * wstep, WBYTES, &full[s]);
                }
            }
        }
        return;
    }
    const int gid = lane >> 2, tig = lane & 3;
    const int wm = warp / WN, wn = warp % WN;
    double cre[MI][NI][2], cim[WR ? 1 : MI][WR ? 1 : NI][2];
#pragma unroll
    for (int mi = 0; mi < MI; mi++)
#pragma unroll
        for (int ni = 0; ni < NI; ni++) {
            cre[mi][ni][0] = cre[mi][ni][1] = 0.0;
            if (!WR) cim[mi][ni][0] = cim[mi][ni][1] = 0.0;
        }
    int q = 0;
    for (int it = 0; it < my_tiles; it++) {
        const int tile = first + it * g0;
        const int b = tile / per_mat, rem = tile - b * per_mat;
        const int ti = rem / ntj, tj = rk_col_tile(rem - ti * ntj, ntj, !WR && g.jint);
        const int i0 = g.ilo + ti * TM, j0 = g.jlo + tj * TN;
        const bool active = (wm * 16 < g.ihi - i0) && (wn * 32 < g.jhi - j0);
        const bool wre = !WR && (j0 + wn * 32 >= g.jre);      // warp-uniform: this warp's 32 columns are real-valued
        if ((g.cs & 4) && active) {                            // C tile on its way to L2 while the K loop runs
            if (WR) {
                const double* Cr = rk_real_view(g.C + (long)b * g.sC, g.mixr) + (long)(i0 + wm * 16 + gid) * 2 * g.ldc + j0 + wn * 32 + tig * 2;
#pragma unroll
                for (int mi = 0; mi < MI; mi++)
#pragma unroll
                    for (int ni = 0; ni < NI; ni++) prefetch_l2(Cr + (long)(mi * 8) * 2 * g.ldc + ni * 8);
            } else {
                const cplx* Cp = g.C + (long)b * g.sC + (long)(i0 + wm * 16 + gid) * g.ldc + j0 + wn * 32 + tig * 2;
#pragma unroll
                for (int mi = 0; mi < MI; mi++)
#pragma unroll
                    for (int ni = 0; ni < NI; ni++) prefetch_l2(Cp + (long)(mi * 8) * g.ldc + ni * 8);
            }
        }
        for (int ch = 0; ch < nch_all; ch++, q++) {
            const int s = q % RK_ST;
            mbar_wait(&full[s], (q / RK_ST) & 1);
            if (active) {
                const double* Ps = reinterpret_cast<const double*>(smem_raw + s * SBYTES) + (wm * 16 + gid) * RK_PRS + tig;
                if (WR) {
                    const double* Ws = reinterpret_cast<const double*>(smem_raw + s * SBYTES + PBYTES) + wn * RK_WRBLK + tig * RK_WRS + gid;
#pragma unroll
                    for (int kk = 0; kk < KC; kk += 4) {
                        double af[MI], bf[NI];
#pragma unroll
                        for (int mi = 0; mi < MI; mi++) af[mi] = Ps[(mi * 8) * RK_PRS + kk];
#pragma unroll
                        for (int ni = 0; ni < NI; ni++) bf[ni] = Ws[kk * RK_WRS + ni * 8];
#pragma unroll
                        for (int mi = 0; mi < MI; mi++)
#pragma unroll
                            for (int ni = 0; ni < NI; ni++) dmma884(cre[mi][ni][0], cre[mi][ni][1], af[mi], bf[ni]);
                    }
                } else if (wre) {
                    // real-valued W in complex storage: only the real parts are read, one DMMA per tile product.  A separate
                    // loop, not a predicate: a predicated-off DMMA still pays its 16 issue cycles
                    const double* Ws = reinterpret_cast<const double*>(reinterpret_cast<const cplx*>(smem_raw + s * SBYTES + PBYTES) +
                                                                       wn * RK_WBLK + tig * RK_WPS + gid);
#pragma unroll
                    for (int kk = 0; kk < KC; kk += 4) {
                        double af[MI], bf[NI];
#pragma unroll
                        for (int mi = 0; mi < MI; mi++) af[mi] = Ps[(mi * 8) * RK_PRS + kk];
#pragma unroll
                        for (int ni = 0; ni < NI; ni++) bf[ni] = Ws[2 * (kk * RK_WPS + ni * 8)];
#pragma unroll
                        for (int mi = 0; mi < MI; mi++)
#pragma unroll
                            for (int ni = 0; ni < NI; ni++) dmma884(cre[mi][ni][0], cre[mi][ni][1], af[mi], bf[ni]);
                    }
                } else {
                    const cplx* Ws = reinterpret_cast<const cplx*>(smem_raw + s * SBYTES + PBYTES) + wn * RK_WBLK + tig * RK_WPS + gid;
#pragma unroll
                    for (int kk = 0; kk < KC; kk += 4) {
                        double af[MI];
                        cplx bf[NI];
#pragma unroll
                        for (int mi = 0; mi < MI; mi++) af[mi] = Ps[(mi * 8) * RK_PRS + kk];
#pragma unroll
                        for (int ni = 0; ni < NI; ni++) bf[ni] = Ws[kk * RK_WPS + ni * 8];
#pragma unroll
                        for (int mi = 0; mi < MI; mi++) {
#pragma unroll
                            for (int ni = 0; ni < NI; ni++) dmma884(cre[mi][ni][0], cre[mi][ni][1], af[mi], bf[ni].x);
#pragma unroll
                            for (int ni = 0; ni < NI; ni++)
                                dmma884(cim[WR ? 0 : mi][WR ? 0 : ni][0], cim[WR ? 0 : mi][WR ? 0 : ni][1], af[mi], bf[ni].y);
                        }
                    }
                }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[s]);
        }
        if (active) {
            if (WR) {
                double* Cr = rk_real_view(g.C + (long)b * g.sC, g.mixr) + (long)(i0 + wm * 16 + gid) * 2 * g.ldc + j0 + wn * 32 + tig * 2;
#pragma unroll
                for (int mi = 0; mi < MI; mi++) {
                    double2 v[NI];
#pragma unroll
                    for (int ni = 0; ni < NI; ni++) v[ni] = rk_ldc(reinterpret_cast<const double2*>(Cr + (long)(mi * 8) * 2 * g.ldc + ni * 8), g.cs);
#pragma unroll
                    for (int ni = 0; ni < NI; ni++) {
                        v[ni].x -= cre[mi][ni][0]; v[ni].y -= cre[mi][ni][1];
                        cre[mi][ni][0] = cre[mi][ni][1] = 0.0;
                        rk_stc(reinterpret_cast<double2*>(Cr + (long)(mi * 8) * 2 * g.ldc + ni * 8), v[ni], g.cs);
                    }
                }
            } else {
                cplx* Cb = g.C + (long)b * g.sC + (long)(i0 + wm * 16 + gid) * g.ldc + j0 + wn * 32 + tig * 2;
#pragma unroll
                for (int mi = 0; mi < MI; mi++) {
                    cplx v[NI][2];
#pragma unroll
                    for (int ni = 0; ni < NI; ni++) {
                        v[ni][0] = rk_ldc(&Cb[(long)(mi * 8) * g.ldc + ni * 8], g.cs);
                        v[ni][1] = rk_ldc(&Cb[(long)(mi * 8) * g.ldc + ni * 8 + 1], g.cs);
                    }
#pragma unroll
                    for (int ni = 0; ni < NI; ni++) {
#pragma unroll
                        for (int e = 0; e < 2; e++) {
                            v[ni][e].x -= cre[mi][ni][e]; v[ni][e].y -= cim[WR ? 0 : mi][WR ? 0 : ni][e];
                            cre[mi][ni][e] = 0.0; cim[WR ? 0 : mi][WR ? 0 : ni][e] = 0.0;
                        }
                        rk_stc(&Cb[(long)(mi * 8) * g.ldc + ni * 8], v[ni][0], g.cs);
                        rk_stc(&Cb[(long)(mi * 8) * g.ldc + ni * 8 + 1], v[ni][1], g.cs);
                    }
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Panel kernels of the base step (block K = [c0, c0 + 32))
// ------------------------------------------------------------------------------------------------
// Phase 1: Ppk[r][K] = A[src(r)][K] for rows r in [rlo, N) (src = row move of this block), zero for the pivot
// rows.  Reads A only and writes Ppk only, so the row moves need no separate pass over the panel.
#define PS_ROWS 128
__global__ void __launch_bounds__(256) k_rk_panel_save(const cplx* __restrict__ A, long strideA, int ld, int N, int c0,
                                                       int rlo, const int* __restrict__ moves, cplx* __restrict__ Ppk,
                                                       long stridePk, int nrb, int mixr, double* __restrict__ PpkR,
                                                       long stridePkR) {
    // one CTA = PS_ROWS consecutive rows: their sources are resolved once into shared memory (identity, then the
    // moves that land in the range), so the copy loop is address arithmetic + one load + one store per element
    __shared__ int s_map[PS_ROWS];
    const int b = blockIdx.y, t = threadIdx.x, lane = t & 31, wrp = t >> 5;
    const int* mv = moves + (long)b * GNB_MOVES_STRIDE;
    const int r0 = rlo + blockIdx.x * PS_ROWS;
    if (t < PS_ROWS) s_map[t] = r0 + t;
    __syncthreads();
    const int nm = mv[0];
    if (t < nm) {
        const int d = mv[1 + 2 * t] - r0;
        if (d >= 0 && d < PS_ROWS) s_map[d] = mv[2 + 2 * t];
    }
    __syncthreads();
    const cplx* Ab = A + (long)b * strideA;
    cplx* Pb = Ppk + (long)b * stridePk;
    const int kc = c0 / 16 + (lane >> 4), kk = lane & 15;
    const bool realp = c0 < mixr;                            // real-stored panel -> real-packed copy
    const double* Arv = rk_real_view(Ab, mixr);
    double* PRb = PpkR + (long)b * stridePkR;
    // all 16 rows of this warp are loaded before the first store (16 independent loads in flight per lane: ncu showed the
    // 4-deep version at 83 % long_scoreboard and 3.6 TB/s)
    constexpr int RPW = PS_ROWS / 8;
    if (realp) {
        double v[RPW];
#pragma unroll
        for (int j = 0; j < RPW; j++) {
            const int q = wrp + 8 * j, r = r0 + q;
            const bool piv = (r >= c0 && r < c0 + GNB_NB);
            v[j] = (r < N && !piv) ? Arv[(long)s_map[q] * 2 * ld + c0 + lane] : 0.0;
        }
#pragma unroll
        for (int j = 0; j < RPW; j++) {
            const int r = r0 + wrp + 8 * j;
            if (r < N) PRb[((long)kc * nrb + (r >> 5)) * RK_PBLK + (r & 31) * RK_PPS + kk] = v[j];
        }
    } else {
#pragma unroll
        for (int h = 0; h < 2; h++) {                      // two passes of 8 rows: the same register footprint as the real case
            cplx v[RPW / 2];
#pragma unroll
            for (int j = 0; j < RPW / 2; j++) {
                const int q = wrp + 8 * (j + h * (RPW / 2)), r = r0 + q;
                const bool piv = (r >= c0 && r < c0 + GNB_NB);
                v[j] = (r < N && !piv) ? Ab[(long)s_map[q] * ld + c0 + lane] : cmake(0.0, 0.0);
            }
#pragma unroll
            for (int j = 0; j < RPW / 2; j++) {
                const int r = r0 + wrp + 8 * (j + h * (RPW / 2));
                if (r < N) Pb[((long)kc * nrb + (r >> 5)) * RK_PBLK + (r & 31) * RK_PPS + kk] = v[j];
            }
        }
    }
}

// Phase 2 (JORDAN): A[r][K] = -P[r][K] inv  (rows outside the pivot block), A[K][K] = inv.
// One CTA = 64 rows; P rows and the inverse are staged in shared memory, every thread forms a 4 x 4 block of
// the product (8 LDS.128 per 16 complex FMAs, FP64-FMA bound).
#define PF_ROWS 64
#define PF_PS 33
__global__ void __launch_bounds__(128) k_rk_panel_fin(cplx* __restrict__ A, long strideA, int ld, int N, int c0,
                                                      const cplx* __restrict__ inv, const cplx* __restrict__ Ppk,
                                                      long stridePk, int nrb) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    cplx* s_inv = reinterpret_cast<cplx*>(smem_raw);         // [32][32]
    cplx* s_P = s_inv + GNB_NB * GNB_NB;                     // [PF_ROWS][PF_PS]
    const int b = blockIdx.y, t = threadIdx.x;
    const int r0 = blockIdx.x * PF_ROWS;
    for (int i = t; i < GNB_NB * GNB_NB; i += 128) s_inv[i] = inv[(long)b * GNB_NB * GNB_NB + i];
    const cplx* Pb = Ppk + (long)b * stridePk;
    for (int e = t; e < PF_ROWS * GNB_NB; e += 128) {
        const int rr = e >> 5, k = e & 31, r = r0 + rr;
        cplx v = cmake(0.0, 0.0);
        if (r < N) v = Pb[((long)(c0 / 16 + (k >> 4)) * nrb + (r >> 5)) * RK_PBLK + (r & 31) * RK_PPS + (k & 15)];
        s_P[rr * PF_PS + k] = v;
    }
    __syncthreads();
    const int tr = t >> 3, tq = t & 7;                      // rows 4 tr .. 4 tr + 3, columns 4 tq .. 4 tq + 3
    cplx acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int c = 0; c < 4; c++) acc[i][c] = cmake(0.0, 0.0);
#pragma unroll 4
    for (int j = 0; j < GNB_NB; j++) {
        cplx p[4], v[4];
#pragma unroll
        for (int i = 0; i < 4; i++) p[i] = s_P[(4 * tr + i) * PF_PS + j];
#pragma unroll
        for (int c = 0; c < 4; c++) v[c] = s_inv[j * GNB_NB + 4 * tq + c];
#pragma unroll
        for (int i = 0; i < 4; i++)
#pragma unroll
            for (int c = 0; c < 4; c++) acc[i][c] = cfnma(acc[i][c], p[i], v[c]);
    }
    cplx* Ab = A + (long)b * strideA;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const int r = r0 + 4 * tr + i;
        if (r >= N) continue;
        const bool piv = (r >= c0 && r < c0 + GNB_NB);
#pragma unroll
        for (int c = 0; c < 4; c++)
            Ab[(long)r * ld + c0 + 4 * tq + c] = piv ? s_inv[(r - c0) * GNB_NB + 4 * tq + c] : acc[i][c];
    }
}

// Phase 2 on the tensor pipe (JORDAN): A[:, K] = 0 with inv in the pivot rows, and inv as the packed W operand of
// K's own rows / column block (a slot no far update reads).  The rank-32 strip update A[:, K] -= P[:, K] inv of the
// packed DMMA kernel then produces exactly what k_rk_panel_fin computes with FP64 FMAs (P has zero pivot rows).
__global__ void __launch_bounds__(256) k_rk_panel_prep(cplx* __restrict__ A, long strideA, int ld, int N, int c0,
                                                       const cplx* __restrict__ inv, cplx* __restrict__ Wpk, long strideWk,
                                                       int ncb) {
    const int b = blockIdx.y, t = threadIdx.x, col = t & 31;
    cplx* Ab = A + (long)b * strideA;
    const cplx* ib = inv + (long)b * GNB_NB * GNB_NB;
    for (int r = blockIdx.x * 8 + (t >> 5); r < N; r += gridDim.x * 8) {
        const bool piv = (r >= c0 && r < c0 + GNB_NB);
        const cplx v = piv ? ib[(r - c0) * GNB_NB + col] : cmake(0.0, 0.0);
        Ab[(long)r * ld + c0 + col] = v;
        if (piv)
            Wpk[(long)b * strideWk + ((long)(r >> 4) * ncb + (c0 >> 5)) * RK_WBLK + (r & 15) * RK_WPS + col] = v;
    }
}

// Row moves of blocks [blk_lo, blk_hi) applied, in order, to a 32-column tile of A.
template <typename ET>
__device__ __forceinline__ void rk_moves_tile(ET* __restrict__ Xb, long ldx, int col, bool colok, const int* __restrict__ moves,
                                              long moves_blk_stride, int b, int blk_lo, int blk_hi, int* s_dst, int* s_src) {
    const int t = threadIdx.x, m0 = t >> 5;                  // this thread handles moves m0, m0 + 8, ...
    for (int blk = blk_lo; blk < blk_hi; blk++) {
        const int* mv = moves + (long)blk * moves_blk_stride + (long)b * GNB_MOVES_STRIDE;
        const int nm = mv[0];
        __syncthreads();
        if (t < nm) { s_dst[t] = mv[1 + 2 * t]; s_src[t] = mv[2 + 2 * t]; }
        __syncthreads();
        ET v[8];
#pragma unroll
        for (int q = 0; q < 8; q++) {
            const int m = m0 + 8 * q;
            if (m < nm && colok) v[q] = Xb[(long)s_src[m] * ldx + col];
        }
        __syncthreads();
#pragma unroll
        for (int q = 0; q < 8; q++) {
            const int m = m0 + 8 * q;
            if (m < nm && colok) Xb[(long)s_dst[m] * ldx + col] = v[q];
        }
    }
}
__global__ void __launch_bounds__(256, 5) k_rk_moves_A(cplx* __restrict__ A, long strideA, int ld, int jlo, int jhi,
                                                    const int* __restrict__ moves, long moves_blk_stride, int blk_lo,
                                                    int blk_hi, int mixr) {
    __shared__ int s_dst[2 * GNB_NB], s_src[2 * GNB_NB];
    const int b = blockIdx.y;
    const int col = jlo + blockIdx.x * 32 + (threadIdx.x & 31);
    cplx* Ab = A + (long)b * strideA;
    if (jlo + (int)blockIdx.x * 32 < mixr)                   // block-uniform: this tile lives in the real-stored columns
        rk_moves_tile<double>(rk_real_view(Ab, mixr), 2L * ld, col, col < jhi, moves, moves_blk_stride, b, blk_lo, blk_hi, s_dst, s_src);
    else
        rk_moves_tile<cplx>(Ab, ld, col, col < jhi, moves, moves_blk_stride, b, blk_lo, blk_hi, s_dst, s_src);
}

// Row moves of block `blk` applied to the saved panels (K chunks [kc_lo, kc_hi) of Ppk).  JORDAN (Lpk != null):
// afterwards the rows of the new pivot block are moved out of Ppk into Lpk (they are the L_ab blocks of the
// forward W solve) and cleared, which leaves Ppk strictly block-upper on pivot rows.
template <typename ET>
__global__ void __launch_bounds__(256) k_rk_moves_P(ET* __restrict__ Ppk, ET* __restrict__ Lpk, long stridePk, int nrb,
                                                    int kc_lo, int kc_hi, int c0, const int* __restrict__ moves) {
    __shared__ int s_dst[2 * GNB_NB], s_src[2 * GNB_NB];
    const int b = blockIdx.y, t = threadIdx.x;
    const int* mv = moves + (long)b * GNB_MOVES_STRIDE;
    const int nm = mv[0];
    if (t < nm) { s_dst[t] = mv[1 + 2 * t]; s_src[t] = mv[2 + 2 * t]; }
    __syncthreads();
    const int kk = t & 15, m0 = t >> 4;                      // 16 moves per pass
    for (int kc = kc_lo + blockIdx.x; kc < kc_hi; kc += gridDim.x) {
        ET* Pc = Ppk + (long)b * stridePk + (long)kc * nrb * RK_PBLK;
        ET v[4];
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const int m = m0 + 16 * q;
            if (m < nm) { const int r = s_src[m]; v[q] = Pc[(long)(r >> 5) * RK_PBLK + (r & 31) * RK_PPS + kk]; }
        }
        __syncthreads();
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const int m = m0 + 16 * q;
            if (m < nm) { const int r = s_dst[m]; Pc[(long)(r >> 5) * RK_PBLK + (r & 31) * RK_PPS + kk] = v[q]; }
        }
        if (Lpk) {
            __syncthreads();
            ET* Lc = Lpk + (long)b * stridePk + (long)kc * nrb * RK_PBLK + (long)(c0 >> 5) * RK_PBLK;
            ET* Pr = Pc + (long)(c0 >> 5) * RK_PBLK;
            for (int e = t; e < 32 * 16; e += 256) {
                const int off = (e >> 4) * RK_PPS + (e & 15);
                Lc[off] = Pr[off];
                Pr[off] = ET();
            }
        }
        __syncthreads();
    }
}

// Forward W of a leaf of nb (1 or 2) adjacent pivot blocks on columns [jlo, jhi):
//     W_a = inv_a A[K_a, cols];   W_b = inv_b (A[K_b, cols] - L_ba W_a)
// written to A and to Wpk.  One CTA walks several 64-column tiles of one matrix, so the pivot-block inverses and
// L_ba are staged in shared memory once, and W_a never makes a round trip through global memory.
#define WS_TC 64
__global__ void __launch_bounds__(256, 2) k_rk_wsolve(cplx* __restrict__ A, long strideA, int ld, int c0, int nb, int jlo,
                                                      int jhi, int tiles_per_cta, const cplx* __restrict__ inv_a,
                                                      const cplx* __restrict__ inv_b, const cplx* __restrict__ Lsrc,
                                                      long stridePk, int nrb, cplx* __restrict__ Wpk, long strideWk,
                                                      int ncb) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    cplx* tile = reinterpret_cast<cplx*>(smem_raw);          // [64][WS_TC]: rows of block a, then rows of block b
    cplx* sInv = tile + 2 * GNB_NB * WS_TC;                  // [2][32][32]
    cplx* sL = sInv + 2 * GNB_NB * GNB_NB;                   // [32][32]  L_ba
    const int b = blockIdx.y, t = threadIdx.x;
    cplx* Ab = A + (long)b * strideA;
    for (int idx = t; idx < GNB_NB * GNB_NB; idx += 256) sInv[idx] = inv_a[(long)b * GNB_NB * GNB_NB + idx];
    if (nb == 2) {
        const int cb = c0 + GNB_NB;
        const cplx* Lb = Lsrc + (long)b * stridePk;
        for (int idx = t; idx < GNB_NB * GNB_NB; idx += 256) {
            sInv[GNB_NB * GNB_NB + idx] = inv_b[(long)b * GNB_NB * GNB_NB + idx];
            const int i = idx >> 5, j = idx & 31, k = c0 + j;
            sL[idx] = Lb[((long)(k >> 4) * nrb + (cb >> 5)) * RK_PBLK + i * RK_PPS + (k & 15)];
        }
    }
    const int nrows = nb * GNB_NB;
    const int c = t & (WS_TC - 1), rg = t >> 6;
    for (int tt = 0; tt < tiles_per_cta; tt++) {
        const int cs = jlo + (blockIdx.x * tiles_per_cta + tt) * WS_TC;
        if (cs >= jhi) break;                                 // block-uniform
        __syncthreads();                                      // previous tile fully consumed (and sInv/sL staged)
        for (int idx = t; idx < nrows * WS_TC; idx += 256) {
            const int i = idx / WS_TC, cc = idx - i * WS_TC, col = cs + cc;
            tile[idx] = (col < jhi) ? Ab[(long)(c0 + i) * ld + col] : cmake(0.0, 0.0);
        }
        __syncthreads();
        const int col = cs + c;
        const bool ok = col < jhi;
        cplx acc[8];
        // ---- W_a = inv_a R_a
#pragma unroll
        for (int q = 0; q < 8; q++) acc[q] = cmake(0.0, 0.0);
        for (int j = 0; j < GNB_NB; j++) {
            const cplx r = tile[j * WS_TC + c];
#pragma unroll
            for (int q = 0; q < 8; q++) acc[q] = cfma(acc[q], sInv[(rg + 4 * q) * GNB_NB + j], r);
        }
        cplx* Wb = Wpk + (long)b * strideWk + (long)(col >> 5) * RK_WBLK + (col & 31);
        if (ok) {
#pragma unroll
            for (int q = 0; q < 8; q++) {
                const int k = c0 + rg + 4 * q;
                Ab[(long)k * ld + col] = acc[q];
                Wb[(long)(k >> 4) * ncb * RK_WBLK + (k & 15) * RK_WPS] = acc[q];
            }
        }
        if (nb == 2) {
            __syncthreads();                                  // every thread is done reading R_a
#pragma unroll
            for (int q = 0; q < 8; q++) tile[(rg + 4 * q) * WS_TC + c] = acc[q];      // W_a replaces R_a
            __syncthreads();
            // ---- R_b -= L_ba W_a
#pragma unroll
            for (int q = 0; q < 8; q++) acc[q] = tile[(GNB_NB + rg + 4 * q) * WS_TC + c];
            for (int j = 0; j < GNB_NB; j++) {
                const cplx wa = tile[j * WS_TC + c];
#pragma unroll
                for (int q = 0; q < 8; q++) acc[q] = cfnma(acc[q], sL[(rg + 4 * q) * GNB_NB + j], wa);
            }
            __syncthreads();                                  // R_b rows are private per (row, col) but read by all below
#pragma unroll
            for (int q = 0; q < 8; q++) tile[(GNB_NB + rg + 4 * q) * WS_TC + c] = acc[q];
            __syncthreads();
            // ---- W_b = inv_b R_b
#pragma unroll
            for (int q = 0; q < 8; q++) acc[q] = cmake(0.0, 0.0);
            for (int j = 0; j < GNB_NB; j++) {
                const cplx r = tile[(GNB_NB + j) * WS_TC + c];
#pragma unroll
                for (int q = 0; q < 8; q++) acc[q] = cfma(acc[q], sInv[GNB_NB * GNB_NB + (rg + 4 * q) * GNB_NB + j], r);
            }
            if (ok) {
#pragma unroll
                for (int q = 0; q < 8; q++) {
                    const int k = c0 + GNB_NB + rg + 4 * q;
                    Ab[(long)k * ld + col] = acc[q];
                    Wb[(long)(k >> 4) * ncb * RK_WBLK + (k & 15) * RK_WPS] = acc[q];
                }
            }
        }
    }
}

// DMMA version of the leaf forward-W kernel (same contract as k_rk_wsolve; columns in 32-wide tiles):
// the three 32 x 32 x 32 products of a leaf pair run on the FP64 tensor pipe.  8 warps = 2 (rows) x 4 (columns),
// warp tile 16 x 8; A operands (inv_a, L_ba, inv_b) with row stride 36, the R / W tile with row stride 34
// (conflict-free LDS.128 fragment loads, as in k_rk_gemm).
#define WM_TC 32
#define WM_AS 36
#define WM_BS 34
// mode 1: A and B real (one DMMA per tile product), 2: A real, B complex (two), 3: both complex (four)
// AT = double: the A operands (inv_a, L_ba, inv_b) are known to be real and staged as doubles (modes 1 and 2 only)
template <typename AT>
__device__ __forceinline__ void ws_mma_product(double (&cre)[2][2], double (&cim)[2][2], const AT* __restrict__ sAm,
                                               const cplx* __restrict__ sBk, int wm, int wn, int gid, int tig, bool neg,
                                               int mode) {
    constexpr bool AR = sizeof(AT) == sizeof(double);
#pragma unroll
    for (int kk = 0; kk < GNB_NB; kk += 4) {
        double ax[2], ay[2];
#pragma unroll
        for (int mi = 0; mi < 2; mi++) {
            if (AR) {
                ax[mi] = reinterpret_cast<const double*>(sAm)[(wm * 16 + mi * 8 + gid) * WM_AS + kk + tig];
                ay[mi] = 0.0;
            } else {
                const cplx a = reinterpret_cast<const cplx*>(sAm)[(wm * 16 + mi * 8 + gid) * WM_AS + kk + tig];
                ax[mi] = a.x; ay[mi] = a.y;
            }
            if (neg) { ax[mi] = -ax[mi]; ay[mi] = -ay[mi]; }
        }
        const cplx bq = sBk[(kk + tig) * WM_BS + wn * 8 + gid];
#pragma unroll
        for (int mi = 0; mi < 2; mi++) {
            dmma884(cre[mi][0], cre[mi][1], ax[mi], bq.x);
            if (mode >= 2) dmma884(cim[mi][0], cim[mi][1], ax[mi], bq.y);
            if (!AR && mode == 3) {
                dmma884(cre[mi][0], cre[mi][1], -ay[mi], bq.y);
                dmma884(cim[mi][0], cim[mi][1], ay[mi], bq.x);
            }
        }
    }
}

template <typename AT>
__global__ void __launch_bounds__(256, sizeof(AT) == sizeof(double) ? 3 : 2) k_rk_wsolve_mma(cplx* __restrict__ A, long strideA, int ld, int c0, int nb,
                                                          int jlo, int jhi, int tiles_per_cta,
                                                          const cplx* __restrict__ inv_a, const cplx* __restrict__ inv_b,
                                                          const cplx* __restrict__ Lsrc, long stridePk, int nrb,
                                                          cplx* __restrict__ Wpk, long strideWk, int ncb, int nreal,
                                                          int mixr, const double* __restrict__ PpkR, long stridePkR,
                                                          double* __restrict__ WpkR, long strideWkR, int a_lo, int jre) {
    // jre: columns >= jre hold real values in complex storage (see RkGemmRpArgs::jre)
    // a_lo: W rows below a_lo are not written back to A (FORWARD mode: only the packed copy is ever read again; the
    // back-substitution of the contact-column path touches rows >= back_row_lo only)
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr bool AR = sizeof(AT) == sizeof(double);
    AT* sA = reinterpret_cast<AT*>(smem_raw);                // [3][32][WM_AS]: inv_a, L_ba, inv_b (doubles when real)
    cplx* sB = reinterpret_cast<cplx*>(sA + 3 * GNB_NB * WM_AS);   // [64][WM_BS]: R_a / W_a rows, then R_b rows
    auto put_a = [&](int idx, cplx v) { if constexpr (AR) reinterpret_cast<double*>(sA)[idx] = v.x; else reinterpret_cast<cplx*>(sA)[idx] = v; };
    const int b = blockIdx.y, t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int gid = lane >> 2, tig = lane & 3, wm = warp >> 2, wn = warp & 3;
    cplx* Ab = A + (long)b * strideA;
    for (int idx = t; idx < GNB_NB * GNB_NB; idx += 256) {
        const int i = idx >> 5, j = idx & 31;
        put_a(i * WM_AS + j, inv_a[(long)b * GNB_NB * GNB_NB + idx]);
        if (nb == 2) {
            const int cb = c0 + GNB_NB, k = c0 + j;
            put_a((2 * GNB_NB + i) * WM_AS + j, inv_b[(long)b * GNB_NB * GNB_NB + idx]);
            const long loff = ((long)(k >> 4) * nrb + (cb >> 5)) * RK_PBLK + i * RK_PPS + (k & 15);
            put_a((GNB_NB + i) * WM_AS + j, (c0 + GNB_NB <= mixr) ? cmake(PpkR[(long)b * stridePkR + loff], 0.0)   // panel a is real-packed
                                                                  : Lsrc[(long)b * stridePk + loff]);
        }
    }
    const int nrows = nb * GNB_NB;
    for (int tt = 0; tt < tiles_per_cta; tt++) {
        const int cs = jlo + (blockIdx.x * tiles_per_cta + tt) * WM_TC;
        if (cs >= jhi) break;                                 // block-uniform
        __syncthreads();                                      // previous tile consumed; operands staged
        const bool rstore = cs < mixr;                       // tile stored as real doubles (mixed layout)
        double* Abr = rk_real_view(Ab, mixr);
        for (int idx = t; idx < nrows * WM_TC; idx += 256) {
            const int i = idx >> 5, cc = idx & 31;
            sB[i * WM_BS + cc] = rstore ? cmake(Abr[(long)(c0 + i) * 2 * ld + cs + cc], 0.0) : Ab[(long)(c0 + i) * ld + cs + cc];
        }
        if (tt + 1 < tiles_per_cta && cs + WM_TC < jhi) {     // the next tile of this CTA on its way to L2
            const int csn = cs + WM_TC, i = t >> 2, part = t & 3;
            if (i < nrows) {
                if (csn < mixr) prefetch_l2(Abr + (long)(c0 + i) * 2 * ld + csn + part * 8);
                else { prefetch_l2(Ab + (long)(c0 + i) * ld + csn + part * 8); prefetch_l2(Ab + (long)(c0 + i) * ld + csn + part * 8 + 4); }
            }
        }
        __syncthreads();
        double cre[2][2], cim[2][2];
        // pivot blocks left of nreal are real; so are the columns left of nreal (see RkGemmArgs::nreal)
        const int mode = (c0 + nrows <= nreal) ? ((cs + WM_TC <= nreal || cs >= jre) ? 1 : 2) : 3;
        const int col = cs + wn * 8 + tig * 2;                // this thread's two adjacent output columns
        cplx* Wb = Wpk + (long)b * strideWk + (long)(col >> 5) * RK_WBLK + (col & 31);
        double* Wr = WpkR + (long)b * strideWkR + (long)(col >> 5) * RK_WRBLK + (col & 31);   // real-packed copy (rstore tiles)
        // ---- W_a = inv_a R_a
#pragma unroll
        for (int mi = 0; mi < 2; mi++) cre[mi][0] = cre[mi][1] = cim[mi][0] = cim[mi][1] = 0.0;
        ws_mma_product<AT>(cre, cim, sA, sB, wm, wn, gid, tig, false, mode);
#pragma unroll
        for (int mi = 0; mi < 2; mi++) {
            const int k = c0 + wm * 16 + mi * 8 + gid;
            const cplx v0 = cmake(cre[mi][0], cim[mi][0]), v1 = cmake(cre[mi][1], cim[mi][1]);
            if (k >= a_lo) {
                if (rstore) *reinterpret_cast<double2*>(Abr + (long)k * 2 * ld + col) = make_double2(v0.x, v1.x);
                else { Ab[(long)k * ld + col] = v0; Ab[(long)k * ld + col + 1] = v1; }
            }
            if (rstore) {
                *reinterpret_cast<double2*>(Wr + (long)(k >> 4) * ncb * RK_WRBLK + (k & 15) * RK_WRS) = make_double2(v0.x, v1.x);
            } else {
                cplx* wq = Wb + (long)(k >> 4) * ncb * RK_WBLK + (k & 15) * RK_WPS;
                wq[0] = v0; wq[1] = v1;
            }
        }
        if (nb == 2) {
            __syncthreads();                                  // every warp is done reading R_a
#pragma unroll
            for (int mi = 0; mi < 2; mi++) {
                const int r = wm * 16 + mi * 8 + gid, cc = wn * 8 + tig * 2;
                sB[r * WM_BS + cc] = cmake(cre[mi][0], cim[mi][0]);          // W_a replaces R_a
                sB[r * WM_BS + cc + 1] = cmake(cre[mi][1], cim[mi][1]);
                const cplx r0 = sB[(GNB_NB + r) * WM_BS + cc], r1 = sB[(GNB_NB + r) * WM_BS + cc + 1];
                cre[mi][0] = r0.x; cim[mi][0] = r0.y; cre[mi][1] = r1.x; cim[mi][1] = r1.y;   // accumulators <- R_b
            }
            __syncthreads();
            // ---- R_b -= L_ba W_a
            ws_mma_product<AT>(cre, cim, sA + GNB_NB * WM_AS, sB, wm, wn, gid, tig, true, mode);
            // R_b of this warp's rows x columns is read only through B fragments of the next product
#pragma unroll
            for (int mi = 0; mi < 2; mi++) {
                const int r = wm * 16 + mi * 8 + gid, cc = wn * 8 + tig * 2;
                sB[(GNB_NB + r) * WM_BS + cc] = cmake(cre[mi][0], cim[mi][0]);
                sB[(GNB_NB + r) * WM_BS + cc + 1] = cmake(cre[mi][1], cim[mi][1]);
            }
            __syncthreads();
            // ---- W_b = inv_b R_b
#pragma unroll
            for (int mi = 0; mi < 2; mi++) cre[mi][0] = cre[mi][1] = cim[mi][0] = cim[mi][1] = 0.0;
            ws_mma_product<AT>(cre, cim, sA + 2 * GNB_NB * WM_AS, sB + GNB_NB * WM_BS, wm, wn, gid, tig, false, mode);
#pragma unroll
            for (int mi = 0; mi < 2; mi++) {
                const int k = c0 + GNB_NB + wm * 16 + mi * 8 + gid;
                const cplx v0 = cmake(cre[mi][0], cim[mi][0]), v1 = cmake(cre[mi][1], cim[mi][1]);
                if (k >= a_lo) {
                    if (rstore) *reinterpret_cast<double2*>(Abr + (long)k * 2 * ld + col) = make_double2(v0.x, v1.x);
                    else { Ab[(long)k * ld + col] = v0; Ab[(long)k * ld + col + 1] = v1; }
                }
                if (rstore) {
                    *reinterpret_cast<double2*>(Wr + (long)(k >> 4) * ncb * RK_WRBLK + (k & 15) * RK_WRS) = make_double2(v0.x, v1.x);
                } else {
                    cplx* wq = Wb + (long)(k >> 4) * ncb * RK_WBLK + (k & 15) * RK_WPS;
                    wq[0] = v0; wq[1] = v1;
                }
            }
        }
    }
}

// Fully real forward-W kernel: real pivot blocks (inv_a, L_ba, inv_b real) AND real columns (left of nreal).  Same
// contract and warp layout as k_rk_wsolve_mma, every operand staged as doubles: 46 KB of shared memory and <= 64
// registers, 4 CTAs per SM - the kernel is bound by the latency of its tile loads, so residency is what it needs.
#define WR_BS 36
__device__ __forceinline__ void ws_rr_product(double (&c)[2][2], const double* __restrict__ sAm, const double* __restrict__ sBk,
                                              int wm, int wn, int gid, int tig, bool neg) {
#pragma unroll
    for (int kk = 0; kk < GNB_NB; kk += 4) {
        const double bq = sBk[(kk + tig) * WR_BS + wn * 8 + gid];
#pragma unroll
        for (int mi = 0; mi < 2; mi++) {
            const double ax = sAm[(wm * 16 + mi * 8 + gid) * WM_AS + kk + tig];
            dmma884(c[mi][0], c[mi][1], neg ? -ax : ax, bq);
        }
    }
}
__global__ void __launch_bounds__(256, 4) k_rk_wsolve_rr(cplx* __restrict__ A, long strideA, int ld, int c0, int nb, int jlo,
                                                         int jhi, int tiles_per_cta, const cplx* __restrict__ inv_a,
                                                         const cplx* __restrict__ inv_b, const cplx* __restrict__ Lsrc,
                                                         long stridePk, int nrb, cplx* __restrict__ Wpk, long strideWk, int ncb,
                                                         int mixr, const double* __restrict__ PpkR, long stridePkR,
                                                         double* __restrict__ WpkR, long strideWkR, int a_lo, int fused) {
    static_assert(WM_AS == WR_BS, "the staged A operands double as B operands of the prologue products");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* sA = reinterpret_cast<double*>(smem_raw);        // [3][32][WM_AS]: inv_a, L_ba (fused: -inv_b L_ba inv_a), inv_b
    double* sB = sA + 3 * GNB_NB * WM_AS;                    // [64][WR_BS]: R_a / W_a rows, then R_b rows
    const int b = blockIdx.y, t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int gid = lane >> 2, tig = lane & 3, wm = warp >> 2, wn = warp & 3;
    cplx* Ab = A + (long)b * strideA;
    double* Abr = rk_real_view(Ab, mixr);
    for (int idx = t; idx < GNB_NB * GNB_NB; idx += 256) {
        const int i = idx >> 5, j = idx & 31;
        sA[i * WM_AS + j] = inv_a[(long)b * GNB_NB * GNB_NB + idx].x;
        if (nb == 2) {
            const int cb = c0 + GNB_NB, k = c0 + j;
            sA[(2 * GNB_NB + i) * WM_AS + j] = inv_b[(long)b * GNB_NB * GNB_NB + idx].x;
            const long loff = ((long)(k >> 4) * nrb + (cb >> 5)) * RK_PBLK + i * RK_PPS + (k & 15);
            sA[(GNB_NB + i) * WM_AS + j] = (c0 + GNB_NB <= mixr) ? PpkR[(long)b * stridePkR + loff] : Lsrc[(long)b * stridePk + loff].x;
        }
    }
    const int nrows = nb * GNB_NB;
    // Fused leaf pair: with Y = -inv_b L_ba inv_a (two 32^3 products, once per CTA) the pair is ONE block product
    //     [W_a; W_b] = [[inv_a, 0], [Y, inv_b]] [R_a; R_b]
    // whose three 32 x 32 x columns products do not depend on each other: no barrier between them, where the
    // three-step form (W_a, R_b -= L W_a, W_b) needs four per tile.  Same flops.
    fused = fused && nb == 2;
    if (fused) {
        double c[2][2];
        __syncthreads();
#pragma unroll
        for (int mi = 0; mi < 2; mi++) c[mi][0] = c[mi][1] = 0.0;
        ws_rr_product(c, sA + GNB_NB * WM_AS, sA, wm, wn, gid, tig, false);          // X = L_ba inv_a  -> sB rows 0..31
#pragma unroll
        for (int mi = 0; mi < 2; mi++) {
            const int r = wm * 16 + mi * 8 + gid, cc = wn * 8 + tig * 2;
            sB[r * WR_BS + cc] = c[mi][0]; sB[r * WR_BS + cc + 1] = c[mi][1];
        }
        __syncthreads();
#pragma unroll
        for (int mi = 0; mi < 2; mi++) c[mi][0] = c[mi][1] = 0.0;
        ws_rr_product(c, sA + 2 * GNB_NB * WM_AS, sB, wm, wn, gid, tig, true);       // Y = -inv_b X    -> the L_ba slot
#pragma unroll
        for (int mi = 0; mi < 2; mi++) {
            const int r = wm * 16 + mi * 8 + gid, cc = wn * 8 + tig * 2;
            sA[(GNB_NB + r) * WM_AS + cc] = c[mi][0]; sA[(GNB_NB + r) * WM_AS + cc + 1] = c[mi][1];
        }
    }
    for (int tt = 0; tt < tiles_per_cta; tt++) {
        const int cs = jlo + (blockIdx.x * tiles_per_cta + tt) * WM_TC;
        if (cs >= jhi) break;                                 // block-uniform
        __syncthreads();                                      // previous tile consumed; operands staged
        const bool rstore = cs < mixr;                       // tile stored as real doubles (mixed layout)
        for (int idx = t; idx < nrows * WM_TC; idx += 256) {
            const int i = idx >> 5, cc = idx & 31;
            sB[i * WR_BS + cc] = rstore ? Abr[(long)(c0 + i) * 2 * ld + cs + cc] : Ab[(long)(c0 + i) * ld + cs + cc].x;
        }
        if (tt + 1 < tiles_per_cta && cs + WM_TC < jhi) {     // the next tile of this CTA on its way to L2 (one 32-byte sector per thread)
            const int csn = cs + WM_TC, i = t >> 2, part = t & 3;
            if (i < nrows) {
                if (csn < mixr) prefetch_l2(Abr + (long)(c0 + i) * 2 * ld + csn + part * 8);
                else { prefetch_l2(Ab + (long)(c0 + i) * ld + csn + part * 8); prefetch_l2(Ab + (long)(c0 + i) * ld + csn + part * 8 + 4); }
            }
        }
        __syncthreads();
        const int col = cs + wn * 8 + tig * 2;                // this thread's two adjacent output columns
        cplx* Wb = Wpk + (long)b * strideWk + (long)(col >> 5) * RK_WBLK + (col & 31);
        double* Wr = WpkR + (long)b * strideWkR + (long)(col >> 5) * RK_WRBLK + (col & 31);
        auto emit = [&](int k, double v0, double v1) {        // row k of W: into A and into the packed W operand
            if (rstore) {
                if (k >= a_lo) *reinterpret_cast<double2*>(Abr + (long)k * 2 * ld + col) = make_double2(v0, v1);
                *reinterpret_cast<double2*>(Wr + (long)(k >> 4) * ncb * RK_WRBLK + (k & 15) * RK_WRS) = make_double2(v0, v1);
            } else {
                if (k >= a_lo) { Ab[(long)k * ld + col] = cmake(v0, 0.0); Ab[(long)k * ld + col + 1] = cmake(v1, 0.0); }
                cplx* wq = Wb + (long)(k >> 4) * ncb * RK_WBLK + (k & 15) * RK_WPS;
                wq[0] = cmake(v0, 0.0); wq[1] = cmake(v1, 0.0);
            }
        };
        double c[2][2];
        // ---- W_a = inv_a R_a
#pragma unroll
        for (int mi = 0; mi < 2; mi++) c[mi][0] = c[mi][1] = 0.0;
        ws_rr_product(c, sA, sB, wm, wn, gid, tig, false);
#pragma unroll
        for (int mi = 0; mi < 2; mi++) emit(c0 + wm * 16 + mi * 8 + gid, c[mi][0], c[mi][1]);
        if (fused) {
#pragma unroll
            for (int mi = 0; mi < 2; mi++) c[mi][0] = c[mi][1] = 0.0;
            ws_rr_product(c, sA + GNB_NB * WM_AS, sB, wm, wn, gid, tig, false);                       // Y R_a
            ws_rr_product(c, sA + 2 * GNB_NB * WM_AS, sB + GNB_NB * WR_BS, wm, wn, gid, tig, false);   // + inv_b R_b
#pragma unroll
            for (int mi = 0; mi < 2; mi++) emit(c0 + GNB_NB + wm * 16 + mi * 8 + gid, c[mi][0], c[mi][1]);
        } else if (nb == 2) {
            __syncthreads();                                  // every warp is done reading R_a
#pragma unroll
            for (int mi = 0; mi < 2; mi++) {
                const int r = wm * 16 + mi * 8 + gid, cc = wn * 8 + tig * 2;
                sB[r * WR_BS + cc] = c[mi][0];                // W_a replaces R_a
                sB[r * WR_BS + cc + 1] = c[mi][1];
                c[mi][0] = sB[(GNB_NB + r) * WR_BS + cc];     // accumulators <- R_b
                c[mi][1] = sB[(GNB_NB + r) * WR_BS + cc + 1];
            }
            __syncthreads();
            ws_rr_product(c, sA + GNB_NB * WM_AS, sB, wm, wn, gid, tig, true);      // R_b -= L_ba W_a
#pragma unroll
            for (int mi = 0; mi < 2; mi++) {
                const int r = wm * 16 + mi * 8 + gid, cc = wn * 8 + tig * 2;
                sB[(GNB_NB + r) * WR_BS + cc] = c[mi][0];
                sB[(GNB_NB + r) * WR_BS + cc + 1] = c[mi][1];
            }
            __syncthreads();
#pragma unroll
            for (int mi = 0; mi < 2; mi++) c[mi][0] = c[mi][1] = 0.0;
            ws_rr_product(c, sA + 2 * GNB_NB * WM_AS, sB + GNB_NB * WR_BS, wm, wn, gid, tig, false);   // W_b = inv_b R_b
#pragma unroll
            for (int mi = 0; mi < 2; mi++) emit(c0 + GNB_NB + wm * 16 + mi * 8 + gid, c[mi][0], c[mi][1]);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Host side
// ------------------------------------------------------------------------------------------------
size_t gnb_rec_pk_elems(int N) { return (size_t)(N / 16) * (N / 32) * RK_PBLK + 4 * RK_PBLK; }
size_t gnb_rec_wk_elems(int N, int ld) { return (size_t)(N / 16) * (ld / 32) * RK_WBLK + 4 * RK_WBLK; }

static int g_rk_m3 = 1;          // 3M complex arithmetic in the rank-K update (3 real DMMAs per complex tile product)
static int g_rk_m3_mink = 64;    // ... for K >= this
static int g_rk_kskip = 1;
static int g_rk_real = 1;        // skip the imaginary DMMAs where the operands are known to be real
static int g_rk_strip = 1;       // 128 x 32 CTA tiles for 32-column strips
static int g_rk_sms = 148;
static int g_rk_wsolve_fused = 1;   // leaf pair of the fully real forward-W kernel as one block product (no barriers between its products)
static int g_rk_lookahead = 0;   // FORWARD: wide part of a far update on a second stream beside the next leaf pair's panel work
static int g_rk_la_mink = 64;    // ... for far updates with K >= this
static int g_rk_la_ctas = 1;     // ... with this many rank-K CTAs per SM
static int g_rk_wskip = 1;       // forward-W: skip the dead write of W into A (FORWARD mode, rows above back_row_lo)
static int g_rk_cs = 4;          // bit 2: L2 prefetch of the C tile at the start of its K loop; bit 0: streaming (evict-first) read-modify-write of C in the rank-K kernels, bit 1: L2 evict-last operand
                                 // copies.  Measured: no effect on the step (59.23 / 59.28 / 59.22 / 59.31 ms for 0 / 1 / 2 / 3), so off
static int g_rk_augreal = 1;     // FORWARD: real arithmetic on the augmented columns while the pivot blocks are real
static int g_rk_tcap_k = 0;      // > 0: a rank-K CTA works on at most max(1, tcap_k / K) tiles (short-lived CTAs), 0: persistent
static int g_rk_lowprio = 0;     // rank-K launches carry the lowest launch priority (the sub-batch streams are high priority)
static const size_t kPfSmem = (size_t)(GNB_NB * GNB_NB + PF_ROWS * PF_PS) * sizeof(cplx);
static const size_t kWmSmem = (size_t)(3 * GNB_NB * WM_AS + 2 * GNB_NB * WM_BS) * sizeof(cplx);
static const size_t kWmSmemR = (size_t)3 * GNB_NB * WM_AS * sizeof(double) + (size_t)2 * GNB_NB * WM_BS * sizeof(cplx);
static const size_t kWrSmem = (size_t)(3 * GNB_NB * WM_AS + 2 * GNB_NB * WR_BS) * sizeof(double);
static int g_rk_wsolve_areal = 2; // real pivot blocks: A operands of the forward-W kernel staged as doubles (3 CTAs per SM)
static int g_rk_rp2 = 1;         // real-panel x complex-W rank-K kernel at 2 CTAs per SM
static int g_rk_wsolve_mma = 1;  // leaf forward-W products on the FP64 tensor pipe
static int g_rk_fin_mma = 1;     // JORDAN pivot-column update A[:,K] = -P inv as a rank-32 strip update on the tensor pipe
static const size_t kWsSmem = (size_t)(2 * GNB_NB * WS_TC + 3 * GNB_NB * GNB_NB) * sizeof(cplx);

cudaError_t gnb_rec_init() {
    cudaError_t e;
#define RK_ATTR(M3_, RB_, CB_)                                                                                  \
    if ((e = cudaFuncSetAttribute(k_rk_gemm<M3_, RB_, CB_>, cudaFuncAttributeMaxDynamicSharedMemorySize,        \
                                  (int)rk_smem<RB_, CB_>()))) return e;
    RK_ATTR(0, 2, 2) RK_ATTR(1, 2, 2) RK_ATTR(0, 4, 1) RK_ATTR(1, 4, 1)
#define RK_RP_ATTR(RB_, CB_, WR_)                                                                                \
    if ((e = cudaFuncSetAttribute(k_rk_gemm_rp<RB_, CB_, WR_>, cudaFuncAttributeMaxDynamicSharedMemorySize,     \
                                  (int)rk_rp_smem<RB_, CB_, WR_>()))) return e;
    RK_RP_ATTR(2, 2, 1) RK_RP_ATTR(2, 2, 0) RK_RP_ATTR(4, 1, 1) RK_RP_ATTR(4, 1, 0)
    if ((e = cudaFuncSetAttribute(k_rk_wsolve, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kWsSmem))) return e;
    if ((e = cudaFuncSetAttribute(k_rk_wsolve_mma<cplx>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kWmSmem))) return e;
    if ((e = cudaFuncSetAttribute(k_rk_wsolve_mma<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kWmSmemR))) return e;
    if ((e = cudaFuncSetAttribute(k_rk_wsolve_rr, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kWrSmem))) return e;
    if ((e = cudaFuncSetAttribute(k_rk_panel_fin, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kPfSmem))) return e;
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_rk_sms, cudaDevAttrMultiProcessorCount, dev);
    return cudaSuccess;
}
int gnb_rec_real_enabled() { return g_rk_real; }
void gnb_rec_set_option(const char* name, int value) {
    if (!strcmp(name, "rk_m3")) g_rk_m3 = value;
    else if (!strcmp(name, "rk_m3_mink")) g_rk_m3_mink = value;
    else if (!strcmp(name, "rk_kskip")) g_rk_kskip = value;
    else if (!strcmp(name, "rk_strip")) g_rk_strip = value;
    else if (!strcmp(name, "rk_real")) g_rk_real = value;
    else if (!strcmp(name, "rk_wsolve_mma")) g_rk_wsolve_mma = value;
    else if (!strcmp(name, "rk_wsolve_areal")) g_rk_wsolve_areal = value;
    else if (!strcmp(name, "rk_rp2")) g_rk_rp2 = value;
    else if (!strcmp(name, "rk_fin_mma")) g_rk_fin_mma = value;
    else if (!strcmp(name, "rk_tcap_k")) g_rk_tcap_k = value;
    else if (!strcmp(name, "rk_wskip")) g_rk_wskip = value;
    else if (!strcmp(name, "rk_lookahead")) g_rk_lookahead = value;
    else if (!strcmp(name, "rk_la_mink")) g_rk_la_mink = value;
    else if (!strcmp(name, "rk_la_ctas")) g_rk_la_ctas = value;
    else if (!strcmp(name, "rk_wsolve_fused")) g_rk_wsolve_fused = value;
    else if (!strcmp(name, "rk_augreal")) g_rk_augreal = value;
    else if (!strcmp(name, "rk_cs")) g_rk_cs = value;
    else if (!strcmp(name, "rk_lowprio")) g_rk_lowprio = value;
    else if (!strcmp(name, "rk_sms") && value > 0) g_rk_sms = value;      // CTAs of the persistent rank-K kernels
}

// Developer trace: CUDA events around every launch of the engine, per stream (tools/trace_elim.py).
#include <vector>
#include <cstdio>
namespace {
struct TraceRec { const char* name; cudaStream_t st; cudaEvent_t e0, e1; int M; };
std::vector<TraceRec> g_trace;
cudaEvent_t g_trace_base = nullptr;
int g_trace_on = 0;
struct TraceScope {
    size_t idx = (size_t)-1;
    TraceScope(const char* name, cudaStream_t st, int M) {
        if (!g_trace_on) return;
        TraceRec r{name, st, nullptr, nullptr, M};
        cudaEventCreate(&r.e0); cudaEventCreate(&r.e1);
        cudaEventRecord(r.e0, st);
        idx = g_trace.size();
        g_trace.push_back(r);
    }
    ~TraceScope() { if (idx != (size_t)-1) cudaEventRecord(g_trace[idx].e1, g_trace[idx].st); }
};
}  // namespace
void gnb_rec_trace_start() {
    g_trace.clear();
    g_trace_on = 1;
    if (!g_trace_base) cudaEventCreate(&g_trace_base);
    cudaEventRecord(g_trace_base, 0);
}
int gnb_rec_trace_dump(const char* path) {
    g_trace_on = 0;
    cudaDeviceSynchronize();
    FILE* f = fopen(path, "w");
    if (!f) return 1;
    for (auto& r : g_trace) {
        float t0 = 0, t1 = 0;
        cudaEventElapsedTime(&t0, g_trace_base, r.e0);
        cudaEventElapsedTime(&t1, g_trace_base, r.e1);
        fprintf(f, "%s %p %d %.4f %.4f\n", r.name, (void*)r.st, r.M, t0, t1);
        cudaEventDestroy(r.e0); cudaEventDestroy(r.e1);
    }
    fclose(f);
    g_trace.clear();
    return 0;
}

// Launch of a rank-K kernel: g0 = CTAs resident at a time, tile cap per CTA from the K extent, optional low priority.
template <typename KT, typename AT>
static void rk_launch(KT kern, const AT& args, int nti, int ntj, long total, int resident, int K, size_t smem, cudaStream_t st) {
    const int g0 = (int)std::min<long>(total, (long)resident);
    const int persistent = cdiv_i(total, g0);
    const int tcap = g_rk_tcap_k > 0 ? std::min(persistent, std::max(1, g_rk_tcap_k / std::max(K, 1))) : persistent;
    const int grid = (int)((total + (long)g0 * tcap - 1) / ((long)g0 * tcap)) * g0;      // whole groups; surplus CTAs find no tile
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(288); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributePriority;
    at[0].val.priority = 0;                                  // numerically largest = lowest priority
    cfg.attrs = at; cfg.numAttrs = g_rk_lowprio ? 1 : 0;
    cudaLaunchKernelEx(&cfg, kern, args, nti, ntj, (int)total, g0, tcap);
}

namespace {
struct Rec {
    cudaStream_t st; int M, N, naug; cplx* A; long strideA; int ld; int jordan;
    const GnbRecWork& ws; long launches;
    int nrb, ncb;
    int mixr;                                                // real-stored leading columns (mixed layout), 0 = none
    // Look-ahead (FORWARD mode): after a far update has reached the 64 columns of the next leaf pair, the rest of it
    // (the WIDE part: forward-W + rank-K update of all other columns) runs on a second stream with one rank-K CTA per
    // SM, while the main stream already factors that pair - its tournaments, panel saves and leaf forward-W are
    // latency-bound kernels that fit beside the rank-K CTAs.  The row moves of the saved panels (k_rk_moves_P), which
    // the wide update is still reading, are deferred to the join.
    bool la_pending = false, la_defer = false;
    int la_jn = 0, la_ctas = 0;
    std::vector<std::pair<int, int>> la_deferred;            // (c0, live_lo) of postponed k_rk_moves_P calls

    cplx* inv(int c0) const { return ws.inv + (long)(c0 / GNB_NB) * ws.inv_blk_stride; }
    int* mv(int c0) const { return ws.moves + (long)(c0 / GNB_NB) * ws.moves_blk_stride; }

    void gemm(int ilo, int ihi, int jlo, int jhi, int klo, int khi, const cplx* P, int kskip) {
        if (ihi <= ilo || jhi <= jlo || khi <= klo) return;
        if (mixr > 0 && jlo < mixr && jhi > mixr) {          // a launch never straddles the real-stored / complex boundary
            gemm(ilo, ihi, jlo, mixr, klo, khi, P, kskip);
            gemm(ilo, ihi, mixr, jhi, klo, khi, P, kskip);
            return;
        }
        if (mixr > 0 && klo < mixr && khi > mixr) {          // panels left of mixr are real-packed, the others complex
            gemm(ilo, ihi, jlo, jhi, klo, mixr, P, kskip);
            gemm(ilo, ihi, jlo, jhi, mixr, khi, P, kskip);
            return;
        }
        if (mixr > 0 && khi <= mixr) {                       // real-packed panel (mixed layout)
            const bool wr = jhi <= mixr;                    // real-stored columns: real-packed W, real C
            RkGemmRpArgs r{};
            r.C = A; r.sC = strideA; r.ldc = ld;
            r.P = ws.PpkR; r.sP = ws.stridePkR; r.nrb = nrb;
            r.W = wr ? (const void*)ws.WpkR : (const void*)ws.Wpk; r.sW = wr ? ws.strideWkR : ws.strideWk; r.ncb = ncb;
            r.ilo = ilo; r.ihi = ihi; r.jlo = jlo; r.jhi = jhi; r.klo = klo; r.khi = khi; r.mixr = mixr;
            r.cs = g_rk_cs;
            // FORWARD: the augmented unit columns (from column N on) are real-valued as long as the pivot blocks are
            r.jre = (!jordan && g_rk_augreal && khi <= ws.nreal) ? N : (1 << 30);
            const bool strip = g_rk_strip && (jhi - jlo == 32) && (ihi - ilo >= 128);
            const int tm = strip ? 128 : 64, tn = strip ? 32 : 64;
            const int nti = cdiv_i(ihi - ilo, tm), ntj = cdiv_i(jhi - jlo, tn);
            const long total = (long)M * nti * ntj;
            int resident = g_rk_sms * (la_ctas > 0 ? la_ctas : ((wr || g_rk_rp2) ? 2 : 1));
            const int K = khi - klo;
            // exactly half of the column tiles are the cheap real-valued ones: interleave them, odd CTA stride
            r.jint = (!wr && !strip && r.jre > jlo && r.jre < jhi && (ntj % 2 == 0) && (r.jre - jlo) == (ntj / 2) * tn) ? 1 : 0;
            if (r.jint && resident % 2 == 0) resident--;
            const double cre_cols = wr ? 0.0 : (double)std::max(0, jhi - std::max(jlo, r.jre));     // real-valued complex-stored columns
            const double flops = (double)(ihi - ilo) * (double)(khi - klo) * M *
                                 (wr ? 2.0 * (jhi - jlo) : 4.0 * ((jhi - jlo) - cre_cols) + 2.0 * cre_cols);
            TraceScope ts(khi - klo >= 256 ? "gemm256+" : khi - klo >= 128 ? "gemm128" : khi - klo >= 64 ? "gemm64" : "gemm32", st, M);
            if (ws.timer) ws.timer->begin(st);
            if (strip) {
                if (wr) rk_launch(k_rk_gemm_rp<4, 1, 1>, r, nti, ntj, total, resident, K, rk_rp_smem<4, 1, 1>(), st);
                else rk_launch(k_rk_gemm_rp<4, 1, 0>, r, nti, ntj, total, resident, K, rk_rp_smem<4, 1, 0>(), st);
            } else {
                if (wr) rk_launch(k_rk_gemm_rp<2, 2, 1>, r, nti, ntj, total, resident, K, rk_rp_smem<2, 2, 1>(), st);
                else rk_launch(k_rk_gemm_rp<2, 2, 0>, r, nti, ntj, total, resident, K, rk_rp_smem<2, 2, 0>(), st);
            }
            if (ws.timer) ws.timer->end(st, flops);
            if (ws.flops_acc) *ws.flops_acc += flops;
            launches++;
            return;
        }
        RkGemmArgs g{};
        g.C = A; g.sC = strideA; g.ldc = ld;
        g.P = P; g.sP = ws.stridePk; g.nrb = nrb;
        g.W = ws.Wpk; g.sW = ws.strideWk; g.ncb = ncb;
        g.ilo = ilo; g.ihi = ihi; g.jlo = jlo; g.jhi = jhi; g.klo = klo; g.khi = khi;
        g.kskip = (kskip && g_rk_kskip) ? 1 : 0;
        g.nreal = g_rk_real ? ws.nreal : 0;
        g.preal = (g.nreal > 0 && khi <= g.nreal) ? 1 : 0;
        g.mixr = mixr;
        g.cs = g_rk_cs;
        const bool strip = g_rk_strip && (jhi - jlo == 32) && (ihi - ilo >= 128);     // 128 x 32 tiles for 32-column strips
        const int tm = strip ? 128 : 64, tn = strip ? 32 : 64;
        const int nti = cdiv_i(ihi - ilo, tm), ntj = cdiv_i(jhi - jlo, tn);
        const long total = (long)M * nti * ntj;
        // executed arithmetic in 4-multiplication-equivalent real flops: 8 per complex MAC, 4 where P is real and W
        // complex, 2 where both are real (tiles entirely left of nreal)
        int ncol1 = 0;
        if (g.preal) ncol1 = std::max(0, (std::min(jhi, g.nreal) - jlo) / tn * tn);
        const double colw = 2.0 * ncol1 + (g.preal ? 4.0 : 8.0) * (double)(jhi - jlo - ncol1);
        double flops = (double)(ihi - ilo) * colw * (double)(khi - klo) * M;
        if (g.kskip)            // rows inside [klo, khi) only meet the strictly block-upper part of the panel
            for (int i0 = ilo; i0 < ihi; i0 += tm)
                if (i0 >= klo && i0 + tm <= khi) flops -= tm * colw * (double)(i0 + 32 - klo) * M;
        TraceScope ts(khi - klo >= 256 ? "gemm256+" : khi - klo >= 128 ? "gemm128" : khi - klo >= 64 ? "gemm64" : "gemm32", st, M);
        if (ws.timer) ws.timer->begin(st);
        const bool m3 = g_rk_m3 && khi - klo >= g_rk_m3_mink;
        if (strip) {
            if (m3) rk_launch(k_rk_gemm<1, 4, 1>, g, nti, ntj, total, g_rk_sms, khi - klo, rk_smem<4, 1>(), st);
            else rk_launch(k_rk_gemm<0, 4, 1>, g, nti, ntj, total, g_rk_sms, khi - klo, rk_smem<4, 1>(), st);
        } else {
            if (m3) rk_launch(k_rk_gemm<1, 2, 2>, g, nti, ntj, total, g_rk_sms, khi - klo, rk_smem<2, 2>(), st);
            else rk_launch(k_rk_gemm<0, 2, 2>, g, nti, ntj, total, g_rk_sms, khi - klo, rk_smem<2, 2>(), st);
        }
        if (ws.timer) ws.timer->end(st, flops);
        if (ws.flops_acc) *ws.flops_acc += flops;
        launches++;
    }

    // row moves of block c0 on the saved panels [live_lo, c0)
    void moves_P(int c0, int live_lo) {
        const int rp = mixr > 0 ? 1 : 0;                   // panels left of mixr live real-packed in PpkR
        if (live_lo < c0) {
            const int split = rp ? std::min(c0, mixr) : live_lo;       // [live_lo, split): real-packed chunks
            if (split > live_lo) {
                dim3 grid(std::min((split - live_lo) / 16, 64), M);
                k_rk_moves_P<double><<<grid, 256, 0, st>>>(ws.PpkR, (double*)nullptr, ws.stridePkR, nrb, live_lo / 16, split / 16,
                                                           c0, mv(c0));
                launches++;
            }
            if (c0 > split) {
                dim3 grid(std::min((c0 - split) / 16, 64), M);
                k_rk_moves_P<cplx><<<grid, 256, 0, st>>>(ws.Ppk, jordan ? ws.Lpk : (cplx*)nullptr, ws.stridePk, nrb, split / 16,
                                                         c0 / 16, c0, mv(c0));
                launches++;
            }
        }
    }
    // end of a look-ahead window: the main stream waits for the wide update, then applies the postponed panel moves
    void la_join() {
        if (!la_pending) return;
        cudaStreamWaitEvent(st, ws.la_join, 0);
        la_pending = false; la_defer = false;
        for (auto& d : la_deferred) moves_P(d.first, d.second);
        la_deferred.clear();
    }

    void base_step(int c0, int live_lo) {
        {
            TraceScope ts("tourn", st, M);
            launches += gnb_launch_tournament(st, M, N, A, strideA, ld, c0, GNB_NB, ws.cand0, ws.cand1, ws.cand_stride,
                                              inv(c0), mv(c0), jordan ? ws.perm : nullptr, ws.perm_stride, ws.info,
                                              (g_rk_real && c0 + GNB_NB <= ws.nreal) ? 1 : 0, mixr);
        }
        TraceScope ts2("panel", st, M);
        if (la_defer) la_deferred.emplace_back(c0, live_lo);     // the wide update of the look-ahead still reads these panels
        else moves_P(c0, live_lo);
        const int rlo = jordan ? 0 : c0 + GNB_NB;
        if (rlo < N) {
            dim3 grid(cdiv_i(N - rlo, PS_ROWS), M);
            k_rk_panel_save<<<grid, 256, 0, st>>>(A, strideA, ld, N, c0, rlo, mv(c0), ws.Ppk, ws.stridePk, nrb, mixr, ws.PpkR,
                                                  ws.stridePkR);
            launches++;
        }
        if (jordan && g_rk_fin_mma && N >= 128) {
            dim3 grid(std::min(cdiv_i(N, 8), 64), M);
            k_rk_panel_prep<<<grid, 256, 0, st>>>(A, strideA, ld, N, c0, inv(c0), ws.Wpk, ws.strideWk, ncb);
            launches++;
            gemm(0, N, c0, c0 + GNB_NB, c0, c0 + GNB_NB, ws.Ppk, 0);
        } else if (jordan) {
            dim3 grid(cdiv_i(N, PF_ROWS), M);
            k_rk_panel_fin<<<grid, 128, kPfSmem, st>>>(A, strideA, ld, N, c0, inv(c0), ws.Ppk, ws.stridePk, nrb);
            launches++;
        }
    }

    // forward W of blocks [c0, c0 + w) on columns [jlo, jhi)  (row moves already applied)
    void trsm(int c0, int w, int jlo, int jhi) {
        const int nb = w / GNB_NB;
        const cplx* Lsrc = jordan ? ws.Lpk : ws.Ppk;
        if (nb <= 2) {
            TraceScope ts("wsolve", st, M);
            if (ws.flops_acc) {                             // 1 (nb = 1) or 3 (nb = 2) products 32 x 32 x columns
                const int nr = g_rk_real ? ws.nreal : 0;
                const double cr = (c0 + w <= nr) ? std::max(0, std::min(jhi, nr) - jlo) : 0;     // real x real columns
                const double per = (c0 + w <= nr) ? 2.0 * cr + 4.0 * ((jhi - jlo) - cr) : 8.0 * (jhi - jlo);
                *ws.flops_acc += (nb == 2 ? 3.0 : 1.0) * 32.0 * 32.0 * per * M;
            }
            if (g_rk_wsolve_mma || mixr > 0) {             // the FMA kernel does not know the mixed layout
                // FORWARD: W rows above the back-substitution's first row live on in the packed copy only
                const int a_lo = (!jordan && g_rk_wskip) ? std::max(0, ws.back_row_lo) / GNB_NB * GNB_NB : 0;
                const int jre = (!jordan && g_rk_augreal && c0 + w <= (g_rk_real ? ws.nreal : 0)) ? N : (1 << 30);
                const int ntile = (jhi - jlo) / WM_TC;
                const int nr = g_rk_real ? ws.nreal : 0;
                const bool areal = g_rk_wsolve_areal && c0 + w <= nr;      // real pivot blocks: inv_a, L_ba, inv_b are real
                // areal = 2: the real columns [jlo, min(jhi, nreal)) go to the fully real kernel (4 CTAs per SM)
                const int jreal = (areal && g_rk_wsolve_areal >= 2) ? std::max(jlo, std::min(jhi, nr / WM_TC * WM_TC)) : jlo;
                if (jreal > jlo) {
                    const int ntile = (jreal - jlo) / WM_TC;
                    const int split = std::max(1, std::min(ntile, cdiv_i(8 * g_rk_sms, M)));
                    const int per = cdiv_i(ntile, split);
                    dim3 grid(cdiv_i(ntile, per), M);
                    k_rk_wsolve_rr<<<grid, 256, kWrSmem, st>>>(A, strideA, ld, c0, nb, jlo, jreal, per, inv(c0),
                                                               nb == 2 ? inv(c0 + GNB_NB) : nullptr, Lsrc, ws.stridePk, nrb, ws.Wpk,
                                                               ws.strideWk, ncb, mixr, ws.PpkR, ws.stridePkR, ws.WpkR, ws.strideWkR, a_lo, g_rk_wsolve_fused);
                    launches++;
                }
                if (jhi > jreal) {
                    const int ntile = (jhi - jreal) / WM_TC;
                    const int split = std::max(1, std::min(ntile, cdiv_i((areal ? 6 : 4) * g_rk_sms, M)));      // CTAs per matrix
                    const int per = cdiv_i(ntile, split);
                    dim3 grid(cdiv_i(ntile, per), M);
                    if (areal)
                        k_rk_wsolve_mma<double><<<grid, 256, kWmSmemR, st>>>(A, strideA, ld, c0, nb, jreal, jhi, per, inv(c0),
                                                                             nb == 2 ? inv(c0 + GNB_NB) : nullptr, Lsrc, ws.stridePk,
                                                                             nrb, ws.Wpk, ws.strideWk, ncb, nr, mixr, ws.PpkR,
                                                                             ws.stridePkR, ws.WpkR, ws.strideWkR, a_lo, jre);
                    else
                        k_rk_wsolve_mma<cplx><<<grid, 256, kWmSmem, st>>>(A, strideA, ld, c0, nb, jreal, jhi, per, inv(c0),
                                                                          nb == 2 ? inv(c0 + GNB_NB) : nullptr, Lsrc, ws.stridePk, nrb,
                                                                          ws.Wpk, ws.strideWk, ncb, nr, mixr, ws.PpkR,
                                                                          ws.stridePkR, ws.WpkR, ws.strideWkR, a_lo, jre);
                    launches++;
                }
                return;
            } else {
                const int ntile = cdiv_i(jhi - jlo, WS_TC);
                const int split = std::max(1, std::min(ntile, cdiv_i(4 * g_rk_sms, M)));      // CTAs per matrix
                const int per = cdiv_i(ntile, split);
                dim3 grid(cdiv_i(ntile, per), M);
                k_rk_wsolve<<<grid, 256, kWsSmem, st>>>(A, strideA, ld, c0, nb, jlo, jhi, per, inv(c0),
                                                         nb == 2 ? inv(c0 + GNB_NB) : nullptr, Lsrc, ws.stridePk, nrb, ws.Wpk,
                                                         ws.strideWk, ncb);
            }
            launches++;
            return;
        }
        const int h = (nb + 1) / 2 * GNB_NB;
        trsm(c0, h, jlo, jhi);
        gemm(c0 + h, c0 + w, jlo, jhi, c0, c0 + h, Lsrc, 0);
        trsm(c0 + h, w - h, jlo, jhi);
    }

    void apply_far(int c0, int w, int jlo, int jhi) {
        if (jhi <= jlo) return;
        if (la_pending && jhi > la_jn) la_join();            // leaves the columns of the look-ahead window
        dim3 grid(cdiv_i(jhi - jlo, 32), M);
        TraceScope ts("movesA", st, M);
        k_rk_moves_A<<<grid, 256, 0, st>>>(A, strideA, ld, jlo, jhi, ws.moves, ws.moves_blk_stride, c0 / GNB_NB,
                                           (c0 + w) / GNB_NB, mixr);
        launches++;
        trsm(c0, w, jlo, jhi);
        if (jordan) gemm(0, N, jlo, jhi, c0, c0 + w, ws.Ppk, 1);
        else gemm(c0 + w, N, jlo, jhi, c0, c0 + w, ws.Ppk, 0);
    }

    // live: an ancestor will still apply this range's panels to other columns, so the saved panels from
    // column live_lo on must follow the row moves of every later block of the range.
    void factor(int c0, int w, bool live, int live_lo) {
        if (w == GNB_NB) { base_step(c0, live ? live_lo : c0); return; }
        const int h = (w / GNB_NB + 1) / 2 * GNB_NB;
        factor(c0, h, true, live ? live_lo : c0);
        const int hi = (!jordan && c0 + w == N) ? N + naug : c0 + w;     // augmented columns ride along
        if (!jordan && ws.side && g_rk_lookahead && !ws.timer && h >= g_rk_la_mink && w - h >= 2 * GNB_NB &&
            hi - (c0 + h) > 2 * GNB_NB) {
            const int jn = c0 + h + 2 * GNB_NB;                          // the next leaf pair's columns: narrow part
            apply_far(c0, h, c0 + h, jn);                                // (joins a previous window first)
            cudaEventRecord(ws.la_fork, st);
            cudaStreamWaitEvent(ws.side, ws.la_fork, 0);
            cudaStream_t main_st = st;
            st = ws.side; la_ctas = g_rk_la_ctas;
            apply_far(c0, h, jn, hi);                                    // wide part on the second stream
            cudaEventRecord(ws.la_join, st);
            st = main_st; la_ctas = 0;
            la_pending = true; la_defer = true; la_jn = jn;
        } else {
            apply_far(c0, h, c0 + h, hi);
        }
        factor(c0 + h, w - h, jordan ? true : live, live ? live_lo : c0 + h);
        if (jordan) apply_far(c0 + h, w - h, c0, c0 + h);
    }

    // X[c0 : c0 + w] of the unit-block-upper system; the blocks above the diagonal are the normalised rows in A.
    // Only rows >= row_lo of the solution are produced (contacts-last ordering of the transmission path).
    void backsub(int c0, int w, int row_lo) {
        if (w <= GNB_NB) return;
        const int h = (w / GNB_NB + 1) / 2 * GNB_NB;
        backsub(c0 + h, w - h, row_lo);
        const int ilo = std::max(c0, row_lo);
        if (ilo >= c0 + h) return;
        // K = columns [c0 + h, c0 + w): the part left of mixr is stored as real doubles (mixed layout), the rest complex
        const int ksplit = mixr > 0 ? std::min(std::max(c0 + h, mixr), c0 + w) : c0 + h;
        for (int part = 0; part < 2; part++) {
            const int klo_ = part == 0 ? c0 + h : ksplit, khi_ = part == 0 ? ksplit : c0 + w;
            if (khi_ <= klo_) continue;
            GnbGemmArgs g{};
            g.C = A + N; g.strideC = strideA; g.ldc = ld;
            g.P = A + klo_; g.strideP = strideA; g.ldp = ld;
            if (part == 0 && mixr > 0) {
                g.Pr = reinterpret_cast<const double*>(reinterpret_cast<const char*>(A) + (size_t)mixr * 8) + klo_;
                g.stridePr = 2 * strideA; g.ldpr = 2 * ld;
            }
            g.W = A + (long)klo_ * ld + N; g.strideW = strideA; g.ldw = ld;
            g.ilo = ilo; g.ihi = c0 + h; g.jlo = 0; g.jhi = naug; g.kdim = khi_ - klo_;
            g.skip_lo = g.skip_hi = -1; g.zero_init = 0; g.plus = 0; g.wscale = nullptr;
            TraceScope ts("backsub", st, M);
            if (ws.timer) ws.timer->begin(st);
            gnb_launch_gemm(st, g, M, false, false);
            // real P x complex W: 4 flops per (row, column, k)
            const double fl = (g.Pr ? 4.0 : 8.0) * (double)(c0 + h - ilo) * (double)naug * g.kdim * M;
            if (ws.timer) ws.timer->end(st, fl);
            if (ws.flops_acc) *ws.flops_acc += fl;
            launches++;
        }
        backsub(c0, h, row_lo);
    }
};
}  // namespace

long gnb_eliminate_rec(cudaStream_t st, int M, int N, int naug, cplx* A, long strideA, int ld, int jordan,
                       const GnbRecWork& ws) {
    if (M <= 0) return 0;
    Rec e{st, M, N, naug, A, strideA, ld, jordan, ws, 0, N / 32, (N + naug) / 32, (!jordan && g_rk_real) ? ws.mixr : 0};
    if (jordan) { gnb_launch_init_perm(st, M, ws.perm, ws.perm_stride, N); e.launches++; }
    e.factor(0, N, false, 0);
    e.la_join();
    if (!jordan && naug > 0) {
        e.apply_far(N - GNB_NB, GNB_NB, N, N + naug);      // the last block is nobody's left sibling
        e.backsub(0, N, std::max(0, ws.back_row_lo) / GNB_NB * GNB_NB);
    }
    return e.launches;
}
