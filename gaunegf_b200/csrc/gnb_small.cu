// Small-orbital-count path: ONE CTA per energy point, the whole matrix on chip (BASELINE cfg 1: N = 64).
//
// The kernels assemble A = E S - F - Sigma0 - Sigma_k(E), invert it in place with a pivoted Gauss-Jordan
// elimination and reduce G = A^-1 without leaving the SM:
//   mode GREEN : G written out (consumers that need the whole matrix, utils.inv / integrate.py:67-71)
//   mode DOS   : -Im diag(G)/pi per orbital and its sum (transport.py:183-190)
//   mode T     : Re Tr[Gamma1 G Gamma2 G^H] from the contact block G[C1, C2] only (transport.py:150-157)
// so a whole cohTrans / DOS call is ONE launch instead of assemble + (tournament, panel, update) x N/32 + reductions.
//
//   k_reg_gj   (N <= 96, default): matrix in REGISTERS, a column spread over the lanes of one warp, pivot search by
//              warp REDUX/shuffle in the owning warp, one CTA barrier per column, pure-FMA register-tile update.
//   k_small_gj (N <= 119, developer switch small_reg=0): matrix in SHARED MEMORY (N*(N|1)*16 B + bookkeeping fits
//              the 227 KB of one sm_100a CTA), every warp finds the pivot redundantly, two barriers per column.  Every
//              step reads and writes the whole matrix through shared memory, which bounds it at ~3.5 TFLOP/s; the
//              lock-step block engine is faster at those sizes (profiles/r01_small_probe.json), so it is not a default.
#include <algorithm>
#include <type_traits>
#include "gnb_common.cuh"
#include "gnb_kernels.h"

namespace {

constexpr double kInvPi = 0.31830988618379067154;

// 1 / v = conj(v) / |v|^2 with one hardware-seeded reciprocal (MUFU.RCP64H + Newton steps) instead of the three
// IEEE divisions of Smith's algorithm: the reciprocal pivot sits on the critical path of every elimination step.
__device__ __forceinline__ cplx crcp_fast(cplx v) {
    const double inv = __drcp_rn(fma(v.x, v.x, v.y * v.y));
    return cmake(v.x * inv, -v.y * inv);
}

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
        const cplx r = crcp_fast(Am[p * ld + k]);
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

// if-chain over a WARP-UNIFORM index with the index as a compile-time constant inside fn: register arrays are
// addressed statically without per-element selects
template <int I, int NMAX, class F>
__device__ __forceinline__ void static_switch(int i, F&& fn) {
    if constexpr (I < NMAX) {
        if (i == I) fn(std::integral_constant<int, I>{});
        else static_switch<I + 1, NMAX>(i, fn);
    }
}

template <int I, int NMAX, class F>
__device__ __forceinline__ void static_for(F&& fn) {
    if constexpr (I < NMAX) {
        fn(std::integral_constant<int, I>{});
        static_for<I + 1, NMAX>(fn);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Register-resident variant (N <= 96): the matrix lives in the register file, distributed so that a matrix COLUMN is
// spread over the 32 lanes of ONE warp:  thread (lane, warp) owns rows lane + 32 a (a < RA) and columns
// warp + NW b (b < CB).  Step k of the (implicitly pivoted, no row exchanges) Gauss-Jordan inverse:
//   1. the warp that owns column k finds the pivot among the rows not used yet with warp shuffles, publishes the
//      multiplier column, the pivot row index and the reciprocal pivot in shared memory;        -- barrier --
//   2. every thread reads its RA multipliers, gets the pivot-row entries of its CB columns by warp shuffle from
//      lane p % 32 of its own warp, and updates its RA x CB register tile with pure FMAs.
//   3. look-ahead: the owner of column k+1 updates that column first and publishes its pivot (bar.arrive on a named
//      barrier) before finishing its step-k update, so the search is hidden behind the other warps' updates.
// One barrier per column; shared-memory traffic per step is RA complex reads per thread instead of a read and a
// write of every matrix element (the shared-memory-resident kernel above is bound by exactly that).  Without row exchanges the in-place result is stored[i][m] = G[invp[i]][piv[m]] (piv[k] = pivot row
// of step k); the epilogues address G through these two maps.
template <int RA, int CB, int NW>
__global__ void __launch_bounds__(32 * NW, (RA * CB <= 8 && NW == 16) ? 2 : (RA * CB <= 16) ? 2 : 1) k_reg_gj(const GnbSmallArgs a) {
    extern __shared__ __align__(16) unsigned char sm_raw[];
    constexpr int NT = 32 * NW, NR = 32 * RA;
    const int N = a.N, ld = N | 1;
    cplx* colbuf = reinterpret_cast<cplx*>(sm_raw);                  // [2][NR]   multipliers of step k (parity k & 1)
    cplx* rinfo = colbuf + 2 * NR;                                   // [2]       reciprocal pivot
    int* piv = reinterpret_cast<int*>(rinfo + 2);                    // [NR] pivot row of step k
    int* invp = piv + NR;                                            // [NR] step at which row i was the pivot row
    int* pos1 = invp + NR;                                           // [NR] position of an orbital in contact ca / cb
    int* pos2 = pos1 + NR;
    double* red = reinterpret_cast<double*>(pos2 + NR);              // [NW]
    cplx* slab = reinterpret_cast<cplx*>(red + NW + (NW & 1));       // [32][ld] staging / G12 buffer
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int e = blockIdx.x;
    cplx A[RA][CB];

    // ---- assemble through a 32-row slab so that the global reads are coalesced
    {
        const cplx E = a.Araw ? cmake(0.0, 0.0) : a.E[e];
        const cplx* SB = a.SigB ? a.SigB + (size_t)e * a.strideSigB : nullptr;
        const cplx* Ar = a.Araw ? a.Araw + (size_t)e * N * N : nullptr;
#pragma unroll
        for (int ra = 0; ra < RA; ra++) {
            const int r0 = 32 * ra;
            for (int rr = warp; rr < 32 && r0 + rr < N; rr += NW)
                for (int j = lane; j < N; j += 32) {
                    const size_t g = (size_t)(r0 + rr) * N + j;
                    cplx v;
                    if (Ar) v = Ar[g];
                    else {
                        v = csub(cmul(E, a.S[g]), a.F[g]);
                        if (a.Sig0) v = csub(v, a.Sig0[g]);
                        if (SB) v = csub(v, SB[g]);
                    }
                    slab[rr * ld + j] = v;
                }
            __syncthreads();
            for (int cidx = 0; cidx < a.ncontacts; cidx++) {          // contacts may overlap: one at a time
                const GnbSmallContact& ct = a.ct[cidx];
                const cplx* blk = ct.blk + (size_t)e * ct.blk_stride;
                for (int idx = t; idx < ct.nc * ct.nc; idx += NT) {
                    const int r = idx / ct.nc, cc = idx - r * ct.nc;
                    const int row = ct.inds[r] - r0;
                    if (row >= 0 && row < 32) {
                        cplx* p = &slab[row * ld + ct.inds[cc]];
                        *p = csub(*p, blk[idx]);
                    }
                }
                __syncthreads();
            }
#pragma unroll
            for (int b = 0; b < CB; b++) {
                const int j = warp + NW * b;
                A[ra][b] = (r0 + lane < N && j < N) ? slab[lane * ld + j] : cmake(0.0, 0.0);
            }
            __syncthreads();
        }
    }

    unsigned used = 0;                                                // bit ra: row lane + 32 ra was a pivot row

    // ---- 1. pivot search of column kk, by the warp that owns it (the column is entirely in its registers)
    auto pivot_search = [&](int kk) {
        const int pr = kk & 1;
        cplx colv[RA];
        static_switch<0, CB>(kk / NW, [&](auto B) {
#pragma unroll
            for (int ra = 0; ra < RA; ra++) colv[ra] = A[ra][B.value];
        });
        // largest |a|^2 among my unused rows as an ordered int key (non-negative floats order like ints; single
        // precision is ample for CHOOSING a pivot), then two REDUX steps: max key, lowest row holding it
        int bk = -1, bi = 0x7fffffff;
#pragma unroll
        for (int ra = 0; ra < RA; ra++) {
            const int i = lane + 32 * ra;
            const int key = __float_as_int(__double2float_rn(fma(colv[ra].x, colv[ra].x, colv[ra].y * colv[ra].y))) &
                            0x7fffffff;                               // NaN -> largest
            if (!((used >> ra) & 1) && i < N && key > bk) { bk = key; bi = i; }
        }
        const int kmax = __reduce_max_sync(0xffffffffu, bk);
        const int p = __reduce_min_sync(0xffffffffu, bk == kmax ? bi : 0x7fffffff);
        cplx pv = colv[0];
#pragma unroll
        for (int ra = 1; ra < RA; ra++)
            if (ra == (p >> 5)) pv = colv[ra];
        pv.x = __shfl_sync(0xffffffffu, pv.x, p & 31);
        pv.y = __shfl_sync(0xffffffffu, pv.y, p & 31);
#pragma unroll
        for (int ra = 0; ra < RA; ra++) colbuf[pr * NR + lane + 32 * ra] = colv[ra];
        if (lane == 0) {
            if (pv.x == 0.0 && pv.y == 0.0) *a.info = 1;              // exactly singular column
            rinfo[pr] = crcp_fast(pv);
            piv[kk] = p;
            invp[p] = kk;
        }
        // column kk restarts from zero, with a 1 in the pivot row: the update then leaves -f r in it (and r in
        // the pivot row)
        static_switch<0, CB>(kk / NW, [&](auto B) {
#pragma unroll
            for (int ra = 0; ra < RA; ra++)
                A[ra][B.value] = cmake((lane == (p & 31) && ra == (p >> 5)) ? 1.0 : 0.0, 0.0);
        });
    };

    // Look-ahead: the warp that owns column k+1 updates that column first, searches its pivot and publishes it
    // (bar.arrive, named barrier 1 + parity) BEFORE finishing its own step-k update; the other warps meet it with
    // bar.sync at the top of step k+1.  The pivot search is therefore off the critical path of every step; all NT
    // threads take part in every barrier phase, so no warp runs more than one step ahead and the double-buffered
    // colbuf / rinfo are safe.
    if (warp == 0) pivot_search(0);
    __syncthreads();
    for (int k = 0; k < N; k++) {
        const int par = k & 1;
        if (k > 0 && warp != k % NW) {
            if (par) asm volatile("bar.sync 2, %0;" ::"n"(NT) : "memory");
            else asm volatile("bar.sync 1, %0;" ::"n"(NT) : "memory");
        }
        // ---- 2. every thread: multipliers f r of its rows, the pivot-row entries of its columns by shuffle from
        //         lane p % 32 (unscaled: the scaling rides on the multipliers), then a pure-FMA tile update.
        //         Pivot row: its entries restart from 0 with multiplier -r  ->  r v exactly.
        const int p = piv[k];
        const cplx r = rinfo[par];
        const int psrc = p & 31;
        const bool mine = lane == psrc;
        if (mine) used |= 1u << (p >> 5);
        cplx f[RA];
#pragma unroll
        for (int ra = 0; ra < RA; ra++) f[ra] = cmul(colbuf[par * NR + lane + 32 * ra], r);
        const bool next_owner = (k + 1 < N) && warp == (k + 1) % NW;
        const int kb1 = (k + 1) / NW;
        static_switch<0, RA>(p >> 5, [&](auto PA) {
            if (mine) f[PA.value] = cneg(r);
            auto column = [&](auto B) {
                cplx v = A[PA.value][B.value];
                v.x = __shfl_sync(0xffffffffu, v.x, psrc);
                v.y = __shfl_sync(0xffffffffu, v.y, psrc);
                if (mine) A[PA.value][B.value] = cmake(0.0, 0.0);
#pragma unroll
                for (int ra = 0; ra < RA; ra++) A[ra][B.value] = cfnma(A[ra][B.value], f[ra], v);
            };
            if (next_owner) {
                static_switch<0, CB>(kb1, column);
                pivot_search(k + 1);
                __threadfence_block();
                if ((k + 1) & 1) asm volatile("bar.arrive 2, %0;" ::"n"(NT) : "memory");
                else asm volatile("bar.arrive 1, %0;" ::"n"(NT) : "memory");
                static_for<0, CB>([&](auto B) {
                    if (B.value != kb1) column(B);
                });
            } else {
                static_for<0, CB>(column);
            }
        });
    }
    __syncthreads();

    // ---- epilogues: stored[i][m] = G[invp[i]][piv[m]]
    if (a.mode == GNB_SMALL_GREEN) {
        // 32 rows of G at a time through the slab, so that the global stores are whole rows (the register tiles hold
        // G scattered by the two pivot maps)
        cplx* G = a.G + (size_t)e * a.strideG;
        int grow[RA], gcol[CB];
#pragma unroll
        for (int ra = 0; ra < RA; ra++) grow[ra] = (lane + 32 * ra < N) ? invp[lane + 32 * ra] : -1;
#pragma unroll
        for (int b = 0; b < CB; b++) gcol[b] = (warp + NW * b < N) ? piv[warp + NW * b] : -1;
        for (int r0 = 0; r0 < N; r0 += 32) {
#pragma unroll
            for (int ra = 0; ra < RA; ra++) {
                const int rr = grow[ra] - r0;
                if (rr >= 0 && rr < 32) {
#pragma unroll
                    for (int b = 0; b < CB; b++)
                        if (gcol[b] >= 0) slab[rr * ld + gcol[b]] = A[ra][b];
                }
            }
            __syncthreads();
            for (int rr = warp; rr < 32 && r0 + rr < N; rr += NW)
                for (int j = lane; j < N; j += 32) G[(size_t)(r0 + rr) * a.ldg + j] = slab[rr * ld + j];
            __syncthreads();
        }
        return;
    }
    double acc = 0.0;
    if (a.mode == GNB_SMALL_DOS) {
#pragma unroll
        for (int ra = 0; ra < RA; ra++) {
            const int i = lane + 32 * ra;
            if (i >= N) continue;
            const int gr = invp[i];
#pragma unroll
            for (int b = 0; b < CB; b++) {
                const int m = warp + NW * b;
                if (m < N && piv[m] == gr) {
                    const double d = -A[ra][b].y * kInvPi;
                    if (a.dos_site) a.dos_site[(size_t)e * N + gr] = d;
                    acc += d;
                }
            }
        }
    } else {                                                          // GNB_SMALL_T: gather G12 = G[C1, C2], then reduce
        const GnbSmallContact& c1 = a.ct[a.ca];
        const GnbSmallContact& c2 = a.ct[a.cb];
        const int n1 = c1.nc, n2 = c2.nc;
        for (int i = t; i < N; i += NT) { pos1[i] = -1; pos2[i] = -1; }
        __syncthreads();
        for (int i = t; i < n1; i += NT) pos1[c1.inds[i]] = i;
        for (int i = t; i < n2; i += NT) pos2[c2.inds[i]] = i;
        __syncthreads();
        cplx* G12 = slab;
#pragma unroll
        for (int ra = 0; ra < RA; ra++) {
            const int i = lane + 32 * ra;
            if (i >= N) continue;
            const int b1 = pos1[invp[i]];
            if (b1 < 0) continue;
#pragma unroll
            for (int b = 0; b < CB; b++) {
                const int m = warp + NW * b;
                if (m < N) {
                    const int d2 = pos2[piv[m]];
                    if (d2 >= 0) G12[b1 * n2 + d2] = A[ra][b];
                }
            }
        }
        __syncthreads();
        const cplx* g1 = c1.gam + (size_t)e * c1.gam_stride;
        const cplx* g2 = c2.gam + (size_t)e * c2.gam_stride;
        for (int idx = t; idx < n1 * n2; idx += NT) {
            const int b = idx / n2, d = idx - b * n2;
            cplx Y = cmake(0.0, 0.0), W = cmake(0.0, 0.0);
            for (int cc = 0; cc < n2; cc++) Y = cfma(Y, G12[b * n2 + cc], g2[cc * n2 + d]);
            for (int aa = 0; aa < n1; aa++) W = cfma(W, cconj(g1[aa * n1 + b]), G12[aa * n2 + d]);
            acc += Y.x * W.x + Y.y * W.y;
        }
    }
#pragma unroll
    for (int off = 16; off; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
    if (lane == 0) red[warp] = acc;
    __syncthreads();
    if (t == 0) {
        double s = 0.0;
        for (int w = 0; w < NW; w++) s += red[w];
        if (a.mode == GNB_SMALL_DOS) a.dos_tot[e] = s;
        else a.T[e] = s;
    }
}

template <int RA, int CB, int NW>
size_t reg_smem(const GnbSmallArgs& a) {
    const int N = a.N, ld = N | 1;
    size_t slab = (size_t)32 * ld;
    if (a.mode == GNB_SMALL_T) slab = std::max(slab, (size_t)a.ct[a.ca].nc * a.ct[a.cb].nc);
    return (size_t)(2 * 32 * RA + 2 + slab) * sizeof(cplx) + (size_t)4 * 32 * RA * sizeof(int) +
           (size_t)(NW + (NW & 1)) * sizeof(double);
}

template <int RA, int CB, int NW>
void launch_reg(cudaStream_t st, const GnbSmallArgs& a) {
    k_reg_gj<RA, CB, NW><<<a.M, 32 * NW, reg_smem<RA, CB, NW>(a), st>>>(a);
}

// ---------------------------------------------------------------------------------------------------------------
// Cluster variant of the register-resident inverse (96 < N <= 128): the matrix columns are spread over the warps of
// a thread-block CLUSTER of CL CTAs (global warp gw = rank * NW + warp owns columns gw + CL NW b), rows over lanes as
// before.  The owning warp publishes the multiplier column, the pivot row index and the reciprocal pivot into the
// shared memory of EVERY CTA of the cluster (st.shared::cluster through mapa), and the per-column barrier is the
// cluster barrier, split into arrive (by the look-ahead owner, right after publishing) and wait (by everybody at the
// top of the next step).  GREEN mode only (inverse written out).  Measured (tools/cluster_probe.py,
// profiles/r01_cluster_probe.json): one launch instead of ~25 latency-bound ones makes utils.inv 1.3-1.6x faster for
// up to ~64 matrices of n = 128; at the 512-problem lock-step batches of the 1-D chain fixed point (BASELINE cfg 4)
// the block engine is 7 % faster (3.44 s vs 3.67 s), so only gnb_inverse_batch dispatches here.
__device__ __forceinline__ unsigned cluster_rank() {
    unsigned r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
// arrive without release semantics: for warps that published nothing since the last barrier (their reads of the step are
// complete by data dependence); the release variant makes every arriving warp drain its memory operations (ncu: membar =
// 32 % of the stall samples when all 32 warps of the cluster arrived with .release at every column)
__device__ __forceinline__ void cluster_arrive_relaxed() { asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ unsigned remote_addr(const void* p, unsigned rank) {
    unsigned l = (unsigned)__cvta_generic_to_shared(p), r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(l), "r"(rank));
    return r;
}
__device__ __forceinline__ void st_remote(unsigned addr, cplx v) {
    asm volatile("st.shared::cluster.v2.f64 [%0], {%1, %2};" ::"r"(addr), "d"(v.x), "d"(v.y) : "memory");
}
__device__ __forceinline__ void st_remote(unsigned addr, int v) {
    asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}

template <int RA, int CB, int NW, int CL>
__global__ void __cluster_dims__(CL, 1, 1) __launch_bounds__(32 * NW, 1) k_reg_gj_cl(const GnbSmallArgs a) {
    extern __shared__ __align__(16) unsigned char sm_raw[];
    constexpr int NT = 32 * NW, NR = 32 * RA, GW = CL * NW;
    const int N = a.N, ld = N | 1;
    cplx* colbuf = reinterpret_cast<cplx*>(sm_raw);                  // [2][NR]
    cplx* rinfo = colbuf + 2 * NR;                                   // [2]
    int* piv = reinterpret_cast<int*>(rinfo + 2);                    // [NR]
    int* invp = piv + NR;                                            // [NR]
    cplx* slab = reinterpret_cast<cplx*>(invp + NR);                 // [32][ld]
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const unsigned rank = cluster_rank();
    const int gw = (int)rank * NW + warp;
    const int e = blockIdx.x / CL;
    cplx A[RA][CB];

    {   // ---- assemble: every CTA stages whole 32-row slabs and keeps its own columns
        const cplx E = a.Araw ? cmake(0.0, 0.0) : a.E[e];
        const cplx* SB = a.SigB ? a.SigB + (size_t)e * a.strideSigB : nullptr;
        const cplx* Ar = a.Araw ? a.Araw + (size_t)e * N * N : nullptr;
#pragma unroll
        for (int ra = 0; ra < RA; ra++) {
            const int r0 = 32 * ra;
            for (int rr = warp; rr < 32 && r0 + rr < N; rr += NW)
                for (int j = lane; j < N; j += 32) {
                    const size_t g = (size_t)(r0 + rr) * N + j;
                    cplx v;
                    if (Ar) v = Ar[g];
                    else {
                        v = csub(cmul(E, a.S[g]), a.F[g]);
                        if (a.Sig0) v = csub(v, a.Sig0[g]);
                        if (SB) v = csub(v, SB[g]);
                    }
                    slab[rr * ld + j] = v;
                }
            __syncthreads();
            for (int cidx = 0; cidx < a.ncontacts; cidx++) {
                const GnbSmallContact& ct = a.ct[cidx];
                const cplx* blk = ct.blk + (size_t)e * ct.blk_stride;
                for (int idx = t; idx < ct.nc * ct.nc; idx += NT) {
                    const int r = idx / ct.nc, cc = idx - r * ct.nc;
                    const int row = ct.inds[r] - r0;
                    if (row >= 0 && row < 32) {
                        cplx* p = &slab[row * ld + ct.inds[cc]];
                        *p = csub(*p, blk[idx]);
                    }
                }
                __syncthreads();
            }
#pragma unroll
            for (int b = 0; b < CB; b++) {
                const int j = gw + GW * b;
                A[ra][b] = (r0 + lane < N && j < N) ? slab[lane * ld + j] : cmake(0.0, 0.0);
            }
            __syncthreads();
        }
    }

    unsigned used = 0;
    auto pivot_search = [&](int kk) {
        const int pr = kk & 1;
        cplx colv[RA];
        static_switch<0, CB>(kk / GW, [&](auto B) {
#pragma unroll
            for (int ra = 0; ra < RA; ra++) colv[ra] = A[ra][B.value];
        });
        int bk = -1, bi = 0x7fffffff;
#pragma unroll
        for (int ra = 0; ra < RA; ra++) {
            const int i = lane + 32 * ra;
            const int key = __float_as_int(__double2float_rn(fma(colv[ra].x, colv[ra].x, colv[ra].y * colv[ra].y))) &
                            0x7fffffff;
            if (!((used >> ra) & 1) && i < N && key > bk) { bk = key; bi = i; }
        }
        const int kmax = __reduce_max_sync(0xffffffffu, bk);
        const int p = __reduce_min_sync(0xffffffffu, bk == kmax ? bi : 0x7fffffff);
        cplx pv = colv[0];
#pragma unroll
        for (int ra = 1; ra < RA; ra++)
            if (ra == (p >> 5)) pv = colv[ra];
        pv.x = __shfl_sync(0xffffffffu, pv.x, p & 31);
        pv.y = __shfl_sync(0xffffffffu, pv.y, p & 31);
        const cplx rcp = crcp_fast(pv);
#pragma unroll
        for (unsigned rk = 0; rk < (unsigned)CL; rk++) {             // publish into every CTA of the cluster
#pragma unroll
            for (int ra = 0; ra < RA; ra++) st_remote(remote_addr(&colbuf[pr * NR + lane + 32 * ra], rk), colv[ra]);
            if (lane == 0) {
                st_remote(remote_addr(&rinfo[pr], rk), rcp);
                st_remote(remote_addr(&piv[kk], rk), p);
                st_remote(remote_addr(&invp[p], rk), kk);
            }
        }
        if (lane == 0 && pv.x == 0.0 && pv.y == 0.0) *a.info = 1;
        static_switch<0, CB>(kk / GW, [&](auto B) {
#pragma unroll
            for (int ra = 0; ra < RA; ra++)
                A[ra][B.value] = cmake((lane == (p & 31) && ra == (p >> 5)) ? 1.0 : 0.0, 0.0);
        });
    };

    cluster_arrive();                                                // every CTA of the cluster is resident before
    cluster_wait();                                                  // anybody addresses a peer's shared memory
    if (gw == 0) pivot_search(0);
    cluster_arrive();
    cluster_wait();
    for (int k = 0; k < N; k++) {
        const int par = k & 1;
        if (k > 0) {
            if (gw != k % GW) { if (a.cl_relaxed) cluster_arrive_relaxed(); else cluster_arrive(); }   // the owner of column k arrived (release) when it published
            cluster_wait();
        }
        const int p = piv[k];
        const cplx r = rinfo[par];
        const int psrc = p & 31;
        const bool mine = lane == psrc;
        if (mine) used |= 1u << (p >> 5);
        cplx f[RA];
#pragma unroll
        for (int ra = 0; ra < RA; ra++) f[ra] = cmul(colbuf[par * NR + lane + 32 * ra], r);
        const bool next_owner = (k + 1 < N) && gw == (k + 1) % GW;
        const int kb1 = (k + 1) / GW;
        static_switch<0, RA>(p >> 5, [&](auto PA) {
            if (mine) f[PA.value] = cneg(r);
            auto column = [&](auto B) {
                cplx v = A[PA.value][B.value];
                v.x = __shfl_sync(0xffffffffu, v.x, psrc);
                v.y = __shfl_sync(0xffffffffu, v.y, psrc);
                if (mine) A[PA.value][B.value] = cmake(0.0, 0.0);
#pragma unroll
                for (int ra = 0; ra < RA; ra++) A[ra][B.value] = cfnma(A[ra][B.value], f[ra], v);
            };
            if (next_owner) {
                static_switch<0, CB>(kb1, column);
                pivot_search(k + 1);
                cluster_arrive();
                static_for<0, CB>([&](auto B) {
                    if (B.value != kb1) column(B);
                });
            } else {
                static_for<0, CB>(column);
            }
        });
    }
    cluster_arrive();                                                // nobody leaves while a peer may still address it
    cluster_wait();

    // ---- G[invp[i]][piv[m]] = stored[i][m]
    cplx* G = a.G + (size_t)e * a.strideG;
#pragma unroll
    for (int ra = 0; ra < RA; ra++) {
        const int i = lane + 32 * ra;
        if (i >= N) continue;
        const size_t row = (size_t)invp[i] * a.ldg;
#pragma unroll
        for (int b = 0; b < CB; b++) {
            const int m = gw + GW * b;
            if (m < N) G[row + piv[m]] = A[ra][b];
        }
    }
}

static int g_cl_relaxed = 0;            // developer switch "small_cl_relaxed" (measured: within noise, off)
template <int RA, int CB, int NW, int CL>
void launch_reg_cl(cudaStream_t st, const GnbSmallArgs& a_) {
    GnbSmallArgs a = a_;
    a.cl_relaxed = g_cl_relaxed;
    const size_t smem = (size_t)(2 * 32 * RA + 2 + 32 * (a.N | 1)) * sizeof(cplx) + (size_t)2 * 32 * RA * sizeof(int);
    k_reg_gj_cl<RA, CB, NW, CL><<<a.M * CL, 32 * NW, smem, st>>>(a);
}

size_t small_smem(int N, int nt) {
    const int ld = N | 1;
    return (size_t)N * ld * sizeof(cplx) + (size_t)(2 * N) * sizeof(int) + (size_t)(nt / 32) * sizeof(double);
}

}  // namespace

void gnb_small_set_cl_relaxed(int on) { g_cl_relaxed = on; }
static int g_reg_resident = 1;          // developer switch "small_reg"
static int g_reg_wide = 0;              // developer switch "small_wide": N <= 64 on 16 warps x (2 x 4) tiles
void gnb_small_set_wide(int on) { g_reg_wide = on; }
void gnb_small_set_reg(int on) { g_reg_resident = on; }
// Largest N the one-CTA-per-energy path takes.  Measured on B200 (tools/small_probe.py, profiles/r01_small_probe.json):
// the register-resident kernel beats the lock-step block engine up to its limit of 96; the shared-memory-resident
// kernel (97..119) does not (it is bound by shared-memory bandwidth), so it only runs when asked for (small_reg=0).
int gnb_small_max_n() { return g_reg_resident ? GNB_SMALL_REG_MAX_N : GNB_SMALL_MAX_N; }

cudaError_t gnb_small_init() {
    cudaError_t e;
    if ((e = cudaFuncSetAttribute(k_small_gj<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024))) return e;
    if ((e = cudaFuncSetAttribute(k_small_gj<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024))) return e;
    if ((e = cudaFuncSetAttribute(k_small_gj<512>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024))) return e;
    if ((e = cudaFuncSetAttribute(k_reg_gj<1, 8, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024))) return e;
    if ((e = cudaFuncSetAttribute(k_reg_gj<2, 8, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024))) return e;
    if ((e = cudaFuncSetAttribute(k_reg_gj<2, 4, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024))) return e;
    if ((e = cudaFuncSetAttribute(k_reg_gj<3, 6, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024))) return e;
    if ((e = cudaFuncSetAttribute(k_reg_gj_cl<4, 4, 16, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024))) return e;
    return cudaFuncSetAttribute(k_reg_gj_cl<6, 6, 8, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 120 * 1024);
}

static int g_cluster = 1;               // developer switch "small_cluster"
static int g_cluster_max_m = 0;         // developer switch "small_cluster_maxm": > 0 overrides the measured batch limits
void gnb_small_set_cluster(int on) { g_cluster = on; }
void gnb_small_set_cluster_max_m(int m) { g_cluster_max_m = m; }
// The cluster kernels win on latency: one launch instead of the block engine's launch chain.  For large batches the block
// engine's tensor-pipe updates catch up, so energy-grid calls use the cluster kernels up to a measured batch size
// (profiles/r02_small_probe.json, GrInt / DOS through the C ABI: N = 112 / 128: 2 x faster up to 36 matrices, even at 324;
// N = 160 / 192: 2 x faster up to 12, 1.3 x at 36, slower at 108).
int gnb_small_cluster_max_m(int n) {
    if (!(g_reg_resident && g_cluster)) return 0;
    if (g_cluster_max_m > 0) return g_cluster_max_m;
    return n <= 128 ? 324 : 72;
}
// largest n for which a plain inverse (GREEN mode from given matrices) runs on chip
int gnb_small_inverse_max_n() { return (g_reg_resident && g_cluster) ? GNB_SMALL_CLUSTER_MAX_N : gnb_small_max_n(); }

void gnb_launch_small(cudaStream_t st, const GnbSmallArgs& a) {
    if (a.M <= 0) return;
    const int N = a.N;
    if (g_reg_resident && g_cluster && N > GNB_SMALL_REG_MAX_N && N <= GNB_SMALL_CLUSTER_MAX_N && a.mode == GNB_SMALL_GREEN) {
        if (N <= 128) launch_reg_cl<4, 4, 16, 2>(st, a);       // 2 CTAs x 16 warps, 4 x 4 complex per thread
        else launch_reg_cl<6, 6, 8, 4>(st, a);                 // 4 CTAs x 8 warps, 6 x 6 complex per thread (N <= 192)
        return;
    }
    if (g_reg_resident && N <= GNB_SMALL_REG_MAX_N) {
        if (N <= 32) launch_reg<1, 8, 4>(st, a);
        else if (N <= 64 && g_reg_wide) launch_reg<2, 4, 16>(st, a);
        else if (N <= 64) launch_reg<2, 8, 8>(st, a);
        else launch_reg<3, 6, 16>(st, a);
        return;
    }
    if (N <= 32) k_small_gj<128><<<a.M, 128, small_smem(N, 128), st>>>(a);
    else if (N <= 64) k_small_gj<256><<<a.M, 256, small_smem(N, 256), st>>>(a);
    else k_small_gj<512><<<a.M, 512, small_smem(N, 512), st>>>(a);
}
