// Energy-dependent self-energy providers on the device.
//
//  * 1-D chain (surfG1D.py:223-295, 344-373): the reference's DAMPED FIXED POINT (not Sancho-Rubio)
//        g0 = inv(A);  repeat  g_new = inv(A - B g B^H);  diff = max|g_new-g| / max(|g_new|,1e-12);
//        g = relax*g_new + (1-relax)*g;  until diff <= conv or max_iter
//    batched over all energies of a chunk in lock-step on the elimination engine (gnb_elim.cu); each
//    problem carries its own active flag / iteration count, so converged energies freeze exactly like
//    lanes of a vmapped jax.lax.while_loop.
//  * Bethe lattice (surfGBethe.py:958-1108, 479-542): 12-direction bulk sweep with frozen sigTot and
//    in-place (Gauss-Seidel) sigmaK, then the 6-direction surface sweep; one CTA per energy, all
//    9x9 blocks in shared memory, 6 warps = the 6 mutually independent directions of a half sweep.
#include <algorithm>
#include <string>

#include "../../include/gaunegf_b200.h"
#include "gnb_common.cuh"
#include "gnb_ctx.h"

static inline int cdiv_i(long a, long b) { return (int)((a + b - 1) / b); }

// ------------------------------------------------------------------------------------------
// 1-D chain
// ------------------------------------------------------------------------------------------
// out[b] = (E[b] + i*eta) * Sm - Hm      (surfG1D.py:260-261; eta = 0 gives t = E stau - tau, :370)
__global__ void __launch_bounds__(256) k_chain_pencil(int nn, const cplx* __restrict__ E, double eta,
                                                      const cplx* __restrict__ Sm, const cplx* __restrict__ Hm,
                                                      cplx* __restrict__ out) {
    const int b = blockIdx.y;
    const cplx z = cmake(E[b].x, E[b].y + eta);
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < nn; idx += gridDim.x * blockDim.x)
        out[(long)b * nn + idx] = csub(cmul(z, Sm[idx]), Hm[idx]);
}

struct ChainFlags { int active; int count; };

__global__ void k_chain_init_flags(ChainFlags* f, double* diffs, int M, int max_iter) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < M) {
        f[b].active = max_iter > 0 ? 1 : 0;
        f[b].count = 0;
        diffs[b] = __longlong_as_double(0x7ff0000000000000LL);   // +inf
    }
}

// One CTA per problem: convergence metric + relaxation mixing (surfG1D.py:278-283)
__global__ void __launch_bounds__(256) k_chain_mix(int nn, cplx* __restrict__ g, const cplx* __restrict__ gn,
                                                   ChainFlags* __restrict__ f, double* __restrict__ diffs,
                                                   double conv, double relax, int max_iter) {
    const int b = blockIdx.x, t = threadIdx.x;
    if (!f[b].active) return;
    __shared__ double red[256];
    cplx* gb = g + (long)b * nn;
    const cplx* nb = gn + (long)b * nn;
    double m = 0.0;
    bool isnan_any = false;
    for (int i = t; i < nn; i += 256) {
        const cplx a = nb[i], o = gb[i];
        const double d = cabs2(csub(a, o)) / fmax(cabs2(a), 1e-12);
        if (d != d) isnan_any = true;
        m = fmax(m, d);
    }
    red[t] = isnan_any ? __longlong_as_double(0x7ff8000000000000LL) : m;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if (t < s) {
            const double x = red[t], y = red[t + s];
            red[t] = (x != x || y != y) ? __longlong_as_double(0x7ff8000000000000LL) : fmax(x, y);
        }
        __syncthreads();
    }
    const double diff = red[0];
    const double keep = 1.0 - relax;
    for (int i = t; i < nn; i += 256) {
        const cplx a = nb[i], o = gb[i];
        gb[i] = cmake(a.x * relax + o.x * keep, a.y * relax + o.y * keep);
    }
    if (t == 0) {
        const int cnt = f[b].count + 1;
        f[b].count = cnt;
        diffs[b] = diff;
        if (!(diff > conv) || cnt >= max_iter) f[b].active = 0;     // NaN > conv is false, as in the reference
    }
}

__global__ void k_chain_count_active(const ChainFlags* f, int M, int* out) {
    __shared__ int s;
    if (threadIdx.x == 0) s = 0;
    __syncthreads();
    int local = 0;
    for (int b = threadIdx.x; b < M; b += blockDim.x) local += f[b].active;
    atomicAdd(&s, local);
    __syncthreads();
    if (threadIdx.x == 0) *out = s;
}

__global__ void k_chain_export(const ChainFlags* f, const double* diffs, int M, int* iters, double* dout) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < M) { iters[b] = f[b].count; dout[b] = diffs[b]; }
}

static int g_chain_compact = 1;        // developer switch "chain_compact"
void gnb_chain_set_compact(int on) { g_chain_compact = on; }

static int chain_invert(gnb_ctx* c, int M, int nc, cplx* Min, cplx* Gout) {
    int rc;
    // one CTA per matrix, on chip (gnb_small.cu).  The thread-block-cluster kernels (97..192) serve the TAIL of the fixed
    // point only: with few live problems the iteration is bound by the block engine's launch chain (~25 dependent launches
    // per inverse), which one cluster launch replaces; at the 512-problem batches of BASELINE cfg 4 the block engine is faster
    if (gnb_small_enabled() && (nc <= gnb_small_max_n() || (nc <= gnb_small_inverse_max_n() && M <= std::min(128, gnb_small_cluster_max_m(nc))))) {   // (here the matrices are given: no
        // assembly to fuse, so the cluster kernels pay off later than on the energy-grid calls: 2.76 s at 256, 2.0-2.3 s at 128)
        GnbSmallArgs sa{};
        sa.N = nc; sa.M = M; sa.mode = GNB_SMALL_GREEN; sa.Araw = Min; sa.info = c->info.as<int>();
        sa.G = Gout; sa.strideG = (long)nc * nc; sa.ldg = nc;
        gnb_launch_small(c->stream, sa);
        c->launches++;
        GNB_CK(cudaGetLastError());
        return GNB_OK;
    }
    GnbElimWork w = gnb_elim_work(c, M, nc, true, &rc);
    if (rc) return rc;
    const long nn = (long)nc * nc;
    c->launches += gnb_eliminate(c->stream, M, nc, 0, Min, nn, nc, 1, w);
    gnb_launch_invperm(c->stream, M, c->perm.as<int>(), c->invperm.as<int>(), nc, nc);
    gnb_launch_unpermute(c->stream, M, nc, Min, nn, nc, c->invperm.as<int>(), nc, Gout, nn);
    c->launches += 2;
    GNB_CK(cudaGetLastError());
    return GNB_OK;
}

// Surface Green's functions of K chain contacts (same block size and iteration parameters) for every energy of the
// chunk, as ONE lock-step batch of K * M fixed-point problems (problem k * M + e): the iteration is a chain of ~30
// short, latency-bound launches, so two contacts in one batch cost far less than two batches.
// g of contact k -> c->cg + k * M * nc * nc ; iteration counts -> cts[k]->iters / diffs
// Active-set compaction: every `check_every` iterations the host reads the per-problem flags; when enough problems
// have converged, the live ones (A, B, g, flags) are gathered into a dense prefix of a second buffer set and the
// lock-step batch shrinks (converged problems are frozen by the reference's while_loop anyway, so retiring them
// changes nothing).  c->cg / fin_flags / fin_diffs hold the results in the ORIGINAL problem order.
__global__ void __launch_bounds__(256) k_chain_gather(int nn, const int* __restrict__ src_slot, const cplx* __restrict__ A0,
                                                      const cplx* __restrict__ B0, const cplx* __restrict__ g0,
                                                      cplx* __restrict__ A1, cplx* __restrict__ B1, cplx* __restrict__ g1) {
    const long s = (long)src_slot[blockIdx.y] * nn, d = (long)blockIdx.y * nn;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nn; i += gridDim.x * blockDim.x) {
        A1[d + i] = A0[s + i]; B1[d + i] = B0[s + i]; g1[d + i] = g0[s + i];
    }
}
// results of the current slots -> original problem order
__global__ void __launch_bounds__(256) k_chain_scatter(int nn, const int* __restrict__ slot_prob, const cplx* __restrict__ g,
                                                       const ChainFlags* __restrict__ f, const double* __restrict__ diffs,
                                                       cplx* __restrict__ g_fin, ChainFlags* __restrict__ f_fin,
                                                       double* __restrict__ d_fin) {
    const int slot = blockIdx.y, prob = slot_prob[slot];
    const long s = (long)slot * nn, d = (long)prob * nn;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nn; i += gridDim.x * blockDim.x) g_fin[d + i] = g[s + i];
    if (blockIdx.x == 0 && threadIdx.x == 0) { f_fin[prob] = f[slot]; d_fin[prob] = diffs[slot]; }
}
__global__ void k_chain_gather_flags(int n, const int* __restrict__ src_slot, const ChainFlags* __restrict__ f0,
                                     const double* __restrict__ d0, ChainFlags* __restrict__ f1, double* __restrict__ d1) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { f1[i] = f0[src_slot[i]]; d1[i] = d0[src_slot[i]]; }
}

int gnb_chain1d_surface_g_multi(gnb_ctx* c, Contact* const* cts, int K, int M, const cplx* dE) {
    const Contact& ct = *cts[0];
    const int nc = ct.nc;
    const long nn = (long)nc * nc;
    const int MT = K * M;
    const size_t bytes = (size_t)MT * nn * sizeof(cplx);
    const size_t fbytes = (size_t)MT * (sizeof(ChainFlags) + sizeof(double)) + 64;
    GNB_CK(c->cA.ensure(bytes)); GNB_CK(c->cB.ensure(bytes)); GNB_CK(c->cg.ensure(bytes));
    GNB_CK(c->cgn.ensure(bytes)); GNB_CK(c->cT1.ensure(bytes)); GNB_CK(c->cM.ensure(bytes));
    GNB_CK(c->cgw.ensure(bytes));
    GNB_CK(c->cflags.ensure(3 * fbytes + 2 * (size_t)MT * sizeof(int)));
    // buffer set 0: cA, cB, cgw ; set 1 (allocated at the first compaction): cA2, cB2, cgw2
    cplx *A = c->cA.as<cplx>(), *B = c->cB.as<cplx>(), *g = c->cgw.as<cplx>(), *gn = c->cgn.as<cplx>(),
         *T1 = c->cT1.as<cplx>(), *Mx = c->cM.as<cplx>(), *g_fin = c->cg.as<cplx>();
    char* fb = c->cflags.as<char>();
    ChainFlags* flags = reinterpret_cast<ChainFlags*>(fb);
    double* diffs = reinterpret_cast<double*>(flags + MT);
    int* d_nact = reinterpret_cast<int*>(diffs + MT);
    ChainFlags* flags2 = reinterpret_cast<ChainFlags*>(fb + fbytes);
    double* diffs2 = reinterpret_cast<double*>(flags2 + MT);
    ChainFlags* fin_flags = reinterpret_cast<ChainFlags*>(fb + 2 * fbytes);
    double* fin_diffs = reinterpret_cast<double*>(fin_flags + MT);
    int* d_slot_prob = reinterpret_cast<int*>(fb + 3 * fbytes);
    int* d_src_slot = d_slot_prob + MT;
    cudaStream_t st = c->stream;

    dim3 pg(std::min(cdiv_i(nn, 256), 1024), M);
    for (int k = 0; k < K; k++) {
        Contact& ck = *cts[k];
        GNB_CK(ck.iters.ensure((size_t)M * sizeof(int)));
        GNB_CK(ck.diffs.ensure((size_t)M * sizeof(double)));
        const long off = (long)k * M * nn;
        k_chain_pencil<<<pg, 256, 0, st>>>((int)nn, dE, ck.eta, ck.Salpha.as<cplx>(), ck.alpha.as<cplx>(), A + off);
        k_chain_pencil<<<pg, 256, 0, st>>>((int)nn, dE, ck.eta, ck.Sbeta.as<cplx>(), ck.beta.as<cplx>(), B + off);
        c->launches += 2;
    }
    GNB_CK(cudaMemcpyAsync(Mx, A, bytes, cudaMemcpyDeviceToDevice, st));
    int rc = chain_invert(c, MT, nc, Mx, g);                        // g0 = inv(A)   (surfG1D.py:287)
    if (rc) return rc;
    k_chain_init_flags<<<cdiv_i(MT, 256), 256, 0, st>>>(flags, diffs, MT, ct.max_iter);
    c->launches++;

    std::vector<int> slot_prob(MT), src_slot;
    for (int i = 0; i < MT; i++) slot_prob[i] = i;
    std::vector<ChainFlags> hflags(MT);
    bool prob_uploaded = false;
    int Mcur = MT;
    const dim3 cp_grid_x(std::min(cdiv_i(nn, 256), 64));
    auto scatter_current = [&]() -> int {                           // current slots -> original order
        if (!prob_uploaded)
            GNB_CK(cudaMemcpyAsync(d_slot_prob, slot_prob.data(), (size_t)Mcur * sizeof(int), cudaMemcpyHostToDevice, st));
        prob_uploaded = true;
        k_chain_scatter<<<dim3(cp_grid_x.x, Mcur), 256, 0, st>>>((int)nn, d_slot_prob, g, flags, diffs, g_fin, fin_flags,
                                                                 fin_diffs);
        c->launches++;
        return GNB_OK;
    };

    GnbGemmArgs ga{};
    ga.ilo = 0; ga.ihi = nc; ga.jlo = 0; ga.jhi = nc; ga.kdim = nc; ga.skip_lo = ga.skip_hi = -1;
    ga.strideC = ga.strideP = ga.strideW = nn; ga.ldc = ga.ldp = ga.ldw = nc;
    const int check_every = 16;
    for (int it = 0; it < ct.max_iter && Mcur > 0; it++) {
        const size_t cur_bytes = (size_t)Mcur * nn * sizeof(cplx);
        ga.C = T1; ga.P = B; ga.W = g; ga.zero_init = 1; ga.plus = 1;        // T1 = B g
        gnb_launch_gemm(st, ga, Mcur, false, false);
        GNB_CK(cudaMemcpyAsync(Mx, A, cur_bytes, cudaMemcpyDeviceToDevice, st));
        ga.C = Mx; ga.P = T1; ga.W = B; ga.zero_init = 0; ga.plus = 0;       // Mx = A - T1 B^H
        gnb_launch_gemm(st, ga, Mcur, true, false);
        c->launches += 2;
        if ((rc = chain_invert(c, Mcur, nc, Mx, gn))) return rc;            // g_new = inv(A - B g B^H)
        k_chain_mix<<<Mcur, 256, 0, st>>>((int)nn, g, gn, flags, diffs, ct.conv, ct.relax, ct.max_iter);
        c->launches++;
        if ((it + 1) % check_every == 0 || it + 1 == ct.max_iter) {
            GNB_CK(cudaMemcpyAsync(hflags.data(), flags, (size_t)Mcur * sizeof(ChainFlags), cudaMemcpyDeviceToHost, st));
            GNB_CK(cudaStreamSynchronize(st));
            src_slot.clear();
            for (int i = 0; i < Mcur; i++)
                if (hflags[i].active) src_slot.push_back(i);
            const int nact = (int)src_slot.size();
            if (nact == 0) break;
            if (g_chain_compact && nact <= Mcur - std::max(8, Mcur / 8) && it + 1 < ct.max_iter) {
                if ((rc = scatter_current())) return rc;                     // retire: results of every current slot
                GNB_CK(c->cA2.ensure(bytes)); GNB_CK(c->cB2.ensure(bytes)); GNB_CK(c->cgw2.ensure(bytes));
                const bool on0 = (A == c->cA.as<cplx>());
                cplx *A1 = on0 ? c->cA2.as<cplx>() : c->cA.as<cplx>(), *B1 = on0 ? c->cB2.as<cplx>() : c->cB.as<cplx>(),
                     *g1 = on0 ? c->cgw2.as<cplx>() : c->cgw.as<cplx>();
                GNB_CK(cudaMemcpyAsync(d_src_slot, src_slot.data(), (size_t)nact * sizeof(int), cudaMemcpyHostToDevice, st));
                k_chain_gather<<<dim3(cp_grid_x.x, nact), 256, 0, st>>>((int)nn, d_src_slot, A, B, g, A1, B1, g1);
                k_chain_gather_flags<<<cdiv_i(nact, 256), 256, 0, st>>>(nact, d_src_slot, flags, diffs, flags2, diffs2);
                c->launches += 2;
                GNB_CK(cudaStreamSynchronize(st));                           // src_slot / slot_prob are reused on the host
                for (int i = 0; i < nact; i++) slot_prob[i] = slot_prob[src_slot[i]];
                A = A1; B = B1; g = g1;
                std::swap(flags, flags2); std::swap(diffs, diffs2);
                Mcur = nact;
                prob_uploaded = false;
            }
        }
    }
    (void)d_nact;
    if ((rc = scatter_current())) return rc;
    for (int k = 0; k < K; k++) {
        k_chain_export<<<cdiv_i(M, 256), 256, 0, st>>>(fin_flags + (long)k * M, fin_diffs + (long)k * M, M,
                                                       cts[k]->iters.as<int>(), cts[k]->diffs.as<double>());
        c->launches++;
    }
    GNB_CK(cudaGetLastError());
    return GNB_OK;
}

// one contact: leaves g in c->cg
int gnb_chain1d_surface_g(gnb_ctx* c, Contact& ct, int M, const cplx* dE) {
    Contact* one = &ct;
    return gnb_chain1d_surface_g_multi(c, &one, 1, M, dE);
}

// ------------------------------------------------------------------------------------------
// Bethe lattice
// ------------------------------------------------------------------------------------------
#define BD 9
#define BNN 12
#define BSZ 81
#define BWARPS 6

// warp-level 9x9 complex inverse (Gauss-Jordan, partial pivoting on |re|+|im|) in shared memory.
// M: 9x9 input (destroyed), Inv: 9x9 output.  aug is a [9][18] scratch.
__device__ void warp_inv9(const cplx* __restrict__ Min, cplx* __restrict__ aug, cplx* __restrict__ Inv, int lane) {
    for (int idx = lane; idx < BD * 2 * BD; idx += 32) {
        const int r = idx / (2 * BD), cc = idx - r * 2 * BD;
        aug[idx] = cc < BD ? Min[r * BD + cc] : cmake(cc - BD == r ? 1.0 : 0.0, 0.0);
    }
    __syncwarp();
    for (int col = 0; col < BD; col++) {
        // pivot search among rows col..8
        double m = (lane >= col && lane < BD) ? cabs1(aug[lane * 2 * BD + col]) : -1.0;
        int idx = lane;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            const double m2 = __shfl_xor_sync(0xffffffffu, m, off);
            const int i2 = __shfl_xor_sync(0xffffffffu, idx, off);
            if (m2 > m || (m2 == m && i2 < idx)) { m = m2; idx = i2; }
        }
        const int p = idx;
        if (p != col && lane < 2 * BD) {
            const cplx a = aug[col * 2 * BD + lane], b = aug[p * 2 * BD + lane];
            aug[col * 2 * BD + lane] = b;
            aug[p * 2 * BD + lane] = a;
        }
        __syncwarp();
        const cplx piv = aug[col * 2 * BD + col];
        __syncwarp();
        if (lane < 2 * BD) aug[col * 2 * BD + lane] = cdiv(aug[col * 2 * BD + lane], piv);
        __syncwarp();
        // eliminate column `col` from the other 8 rows: 8 x 18 elements over the warp
        cplx fv[5], pv[5];
#pragma unroll
        for (int q = 0; q < 5; q++) {
            const int e = lane + 32 * q;
            if (e < (BD - 1) * 2 * BD) {
                int r = e / (2 * BD);
                const int cc = e - r * 2 * BD;
                if (r >= col) r++;
                fv[q] = aug[r * 2 * BD + col];
                pv[q] = aug[col * 2 * BD + cc];
            }
        }
        __syncwarp();
#pragma unroll
        for (int q = 0; q < 5; q++) {
            const int e = lane + 32 * q;
            if (e < (BD - 1) * 2 * BD) {
                int r = e / (2 * BD);
                const int cc = e - r * 2 * BD;
                if (r >= col) r++;
                aug[r * 2 * BD + cc] = cfnma(aug[r * 2 * BD + cc], fv[q], pv[q]);
            }
        }
        __syncwarp();
    }
    for (int idx = lane; idx < BSZ; idx += 32) {
        const int r = idx / BD, cc = idx - r * BD;
        Inv[idx] = aug[r * 2 * BD + BD + cc];
    }
    __syncwarp();
}

// out = mix * (B g B^H) + (1-mix) * old      (surfGBethe.py:1013, 1093)
__device__ void warp_bgbh_mix(const cplx* __restrict__ B, const cplx* __restrict__ gm, cplx* __restrict__ tmp,
                              const cplx* __restrict__ old, cplx* __restrict__ out, double mix, int lane) {
    for (int idx = lane; idx < BSZ; idx += 32) {                  // tmp = B g
        const int r = idx / BD, cc = idx - r * BD;
        cplx acc = cmake(0.0, 0.0);
#pragma unroll
        for (int k = 0; k < BD; k++) acc = cfma(acc, B[r * BD + k], gm[k * BD + cc]);
        tmp[idx] = acc;
    }
    __syncwarp();
    for (int idx = lane; idx < BSZ; idx += 32) {                  // (tmp B^H)[r][cc] = sum_k tmp[r][k] conj(B[cc][k])
        const int r = idx / BD, cc = idx - r * BD;
        cplx acc = cmake(0.0, 0.0);
#pragma unroll
        for (int k = 0; k < BD; k++) acc = cfma(acc, tmp[r * BD + k], cconj(B[cc * BD + k]));
        const cplx o = old[idx];
        out[idx] = cmake(mix * acc.x + (1.0 - mix) * o.x, mix * acc.y + (1.0 - mix) * o.y);
    }
    __syncwarp();
}

struct BetheSmem {
    cplx sK[BNN * BSZ];
    cplx sOld[BNN * BSZ];
    cplx sB[BNN * BSZ];
    cplx A[BSZ];
    cplx tot[BSZ];
    cplx gsurf[BSZ];
    cplx wM[BWARPS][BSZ];
    cplx wInv[BWARPS][BSZ];
    cplx wTmp[BWARPS][BSZ];
    cplx wAug[BWARPS][2 * BSZ];
    double red[BWARPS * 32 * 2];
    double diff;
};

__device__ double bethe_diff(BetheSmem& s, int nblk, int t) {
    double num = 0.0, den = 0.0;
    for (int i = t; i < nblk * BSZ; i += BWARPS * 32) {
        num = fmax(num, cabs2(csub(s.sK[i], s.sOld[i])));
        den = fmax(den, cabs2(s.sOld[i]));
    }
    s.red[t] = num;
    s.red[BWARPS * 32 + t] = den;
    __syncthreads();
    if (t == 0) {
        double a = 0.0, b = 0.0;
        bool nanv = false;
        for (int i = 0; i < BWARPS * 32; i++) {
            a = fmax(a, s.red[i]);
            b = fmax(b, s.red[BWARPS * 32 + i]);
            if (s.red[i] != s.red[i]) nanv = true;
        }
        s.diff = nanv ? __longlong_as_double(0x7ff8000000000000LL) : a / b;
    }
    __syncthreads();
    return s.diff;
}

// One CTA (6 warps) per energy.  which: 1 -> write 12 bulk blocks, 2/0 -> write 9 surface blocks.
__global__ void __launch_bounds__(BWARPS * 32) k_bethe(const cplx* __restrict__ E, double eta,
                                                       const cplx* __restrict__ H, const cplx* __restrict__ Sl,
                                                       const cplx* __restrict__ Vl, double conv, double mix,
                                                       int max_iter, int which, cplx* __restrict__ out,
                                                       int* __restrict__ iters, double* __restrict__ diffs) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    BetheSmem& s = *reinterpret_cast<BetheSmem*>(smem_raw);
    const int b = blockIdx.x, t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const cplx z = cmake(E[b].x, E[b].y - eta);                    // E MINUS i*eta (surfGBethe.py:995)
    for (int i = t; i < BNN * BSZ; i += BWARPS * 32) {
        const int e = i % BSZ;
        s.sB[i] = csub(cmul(z, Sl[i]), Vl[i]);
        s.sK[i] = cmake(0.0, (e / BD == e % BD) ? -1.0 : 0.0);     // sigmaK0 = -i * I
    }
    for (int i = t; i < BSZ; i += BWARPS * 32) {
        const cplx zi = (i / BD == i % BD) ? z : cmake(0.0, 0.0);
        s.A[i] = csub(zi, H[i]);
    }
    __syncthreads();
    // ---- bulk: 12 directions (surfGBethe.py:1000-1022)
    int count = 0;
    double diff = __longlong_as_double(0x7ff0000000000000LL);
    while (diff > conv && count < max_iter) {
        for (int i = t; i < BNN * BSZ; i += BWARPS * 32) s.sOld[i] = s.sK[i];
        __syncthreads();
        for (int i = t; i < BSZ; i += BWARPS * 32) {               // sigTot, frozen for the sweep, summed k = 0..11
            cplx acc = cmake(0.0, 0.0);
            for (int k = 0; k < BNN; k++) acc = cadd(acc, s.sOld[k * BSZ + i]);
            s.tot[i] = acc;
        }
        __syncthreads();
        for (int half = 0; half < 2; half++) {
            const int k = half * 6 + warp, pair = (k + 6) % 12;
            // half 0 reads the not-yet-updated sK[k+6]; half 1 reads the sK[k-6] updated in half 0
            for (int i = lane; i < BSZ; i += 32) s.wM[warp][i] = cadd(csub(s.A[i], s.tot[i]), s.sK[pair * BSZ + i]);
            __syncwarp();
            warp_inv9(s.wM[warp], s.wAug[warp], s.wInv[warp], lane);
            __syncthreads();                                         // all reads of sK[pair] done before writes
            warp_bgbh_mix(&s.sB[k * BSZ], s.wInv[warp], s.wTmp[warp], &s.sOld[k * BSZ], &s.sK[k * BSZ], mix, lane);
            __syncthreads();
        }
        diff = bethe_diff(s, BNN, t);
        count++;
    }
    if (which == 1) {
        for (int i = t; i < BNN * BSZ; i += BWARPS * 32) out[(long)b * BNN * BSZ + i] = s.sK[i];
        if (t == 0) { if (iters) iters[b] = count; if (diffs) diffs[b] = diff; }
        return;
    }
    // ---- surface: first 9 directions, in-plane ones [0,1,2,6,7,8] relaxed (surfGBethe.py:1074-1102)
    int count2 = 0;
    diff = __longlong_as_double(0x7ff0000000000000LL);
    const int plane[6] = {0, 1, 2, 6, 7, 8};
    while (diff > conv && count2 < max_iter) {
        for (int i = t; i < 9 * BSZ; i += BWARPS * 32) s.sOld[i] = s.sK[i];
        __syncthreads();
        if (warp == 0) {
            for (int i = lane; i < BSZ; i += 32) {
                cplx acc = cmake(0.0, 0.0);
                for (int k = 0; k < 9; k++) acc = cadd(acc, s.sOld[k * BSZ + i]);
                s.wM[0][i] = csub(s.A[i], acc);
            }
            __syncwarp();
            warp_inv9(s.wM[0], s.wAug[0], s.gsurf, lane);
        }
        __syncthreads();
        {
            const int k = plane[warp];
            warp_bgbh_mix(&s.sB[k * BSZ], s.gsurf, s.wTmp[warp], &s.sOld[k * BSZ], &s.sK[k * BSZ], mix, lane);
        }
        __syncthreads();
        diff = bethe_diff(s, 9, t);
        count2++;
    }
    for (int i = t; i < 9 * BSZ; i += BWARPS * 32) out[(long)b * 9 * BSZ + i] = s.sK[i];
    if (t == 0) { if (iters) iters[b] = count * 10000 + count2; if (diffs) diffs[b] = diff; }
}

// contact block (natoms*9)^2: per atom  sum_{d<9} sigSurf[d] - sum_{n in connected} sigSurf[n]  on the
// atom's diagonal 9x9 block (surfGBethe.py:519-527)
__global__ void __launch_bounds__(128) k_bethe_block(const cplx* __restrict__ surf, int natoms,
                                                     const int* __restrict__ nb_off, const int* __restrict__ nb_dirs,
                                                     cplx* __restrict__ blk) {
    const int b = blockIdx.x, nc = natoms * BD;
    const cplx* sb = surf + (long)b * 9 * BSZ;
    cplx* ob = blk + (long)b * nc * nc;
    for (int idx = threadIdx.x; idx < nc * nc; idx += blockDim.x) {
        const int r = idx / nc, cc = idx - r * nc;
        const int a = r / BD;
        cplx v = cmake(0.0, 0.0);
        if (cc / BD == a) {
            const int e = (r - a * BD) * BD + (cc - a * BD);
            for (int d = 0; d < 9; d++) v = cadd(v, sb[d * BSZ + e]);
            for (int q = nb_off[a]; q < nb_off[a + 1]; q++) v = csub(v, sb[nb_dirs[q] * BSZ + e]);
        }
        ob[idx] = v;
    }
}

// per-context (per-device) kernel attributes; called from gnb_create after cudaSetDevice
cudaError_t gnb_sigma_init() {
    return cudaFuncSetAttribute(k_bethe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(BetheSmem));
}

int gnb_bethe_raw(gnb_ctx* c, Contact& ct, int M, const cplx* dE, int which, cplx* d_out) {
    GNB_CK(ct.iters.ensure((size_t)M * sizeof(int)));
    GNB_CK(ct.diffs.ensure((size_t)M * sizeof(double)));
    k_bethe<<<M, BWARPS * 32, sizeof(BetheSmem), c->stream>>>(dE, ct.eta, ct.H.as<cplx>(), ct.Slist.as<cplx>(),
                                                             ct.Vlist.as<cplx>(), ct.conv, ct.mix, ct.max_iter,
                                                             which, d_out, ct.iters.as<int>(), ct.diffs.as<double>());
    c->launches++;
    GNB_CK(cudaGetLastError());
    return GNB_OK;
}

// ------------------------------------------------------------------------------------------
// Contact self-energy blocks of a chunk (+ Gamma)
// ------------------------------------------------------------------------------------------
int gnb_contact_eval(gnb_ctx* c, Contact& ct, int M, const cplx* dE, int want_gamma, const cplx* g_ready) {
    const int nc = ct.nc;
    const long nn = (long)nc * nc;
    GNB_CK(ct.blk.ensure((size_t)M * nn * sizeof(cplx)));
    cudaStream_t st = c->stream;
    if (ct.kind == GNB_C_CHAIN1D) {
        if (!g_ready) {                                   // not part of a joint batch (gnb_chain1d_surface_g_multi)
            int rc = gnb_chain1d_surface_g(c, ct, M, dE);
            if (rc) return rc;
            g_ready = c->cg.as<cplx>();
        }
        GNB_CK(c->ct.ensure((size_t)M * nn * sizeof(cplx)));
        cplx* tmat = c->ct.as<cplx>();
        dim3 pg(std::min(cdiv_i(nn, 256), 1024), M);
        k_chain_pencil<<<pg, 256, 0, st>>>((int)nn, dE, 0.0, ct.stau.as<cplx>(), ct.tau.as<cplx>(), tmat);
        GnbGemmArgs ga{};
        ga.ilo = 0; ga.ihi = nc; ga.jlo = 0; ga.jhi = nc; ga.kdim = nc; ga.skip_lo = ga.skip_hi = -1;
        ga.strideC = ga.strideP = ga.strideW = nn; ga.ldc = ga.ldp = ga.ldw = nc;
        ga.zero_init = 1; ga.plus = 1;
        ga.C = c->cT1.as<cplx>(); ga.P = tmat; ga.W = g_ready;                    // T1 = t g
        gnb_launch_gemm(st, ga, M, false, false);
        ga.C = ct.blk.as<cplx>(); ga.P = c->cT1.as<cplx>(); ga.W = tmat;          // Sigma = T1 t^H
        gnb_launch_gemm(st, ga, M, true, false);
        c->launches += 3;
    } else if (ct.kind == GNB_C_BETHE) {
        GNB_CK(ct.surf.ensure((size_t)M * 9 * BSZ * sizeof(cplx)));
        int rc = gnb_bethe_raw(c, ct, M, dE, 2, ct.surf.as<cplx>());
        if (rc) return rc;
        k_bethe_block<<<M, 128, 0, st>>>(ct.surf.as<cplx>(), ct.natoms, ct.d_nb_off.as<int>(), ct.d_nb_dirs.as<int>(),
                                         ct.blk.as<cplx>());
        c->launches++;
    } else {
        return gnb_fail(c, GNB_ERR_ARG, "contact_eval: constant contact");
    }
    ct.blk_ptr = ct.blk.as<cplx>(); ct.blk_stride = nn;
    if (want_gamma) {
        GNB_CK(ct.gam.ensure((size_t)M * nn * sizeof(cplx)));
        gnb_launch_gamma_from_sigma(st, M, ct.blk.as<cplx>(), nn, nc, ct.gam.as<cplx>());
        c->launches++;
        ct.gam_ptr = ct.gam.as<cplx>(); ct.gam_stride = nn;
    }
    GNB_CK(cudaGetLastError());
    return GNB_OK;
}

// one chunk of gnb_sigma_eval (m energies already in c->dE)
static int sigma_eval_chunk(gnb_ctx* c, Contact& ct, int which, int m, double* out_blk, int32_t* iters, double* diffs) {
    const cplx* dE = c->dE.as<cplx>();
    const long nn = (long)ct.nc * ct.nc;
    const cplx* src = nullptr;
    size_t per = 0;
    int rc = GNB_OK;
    if (ct.kind == GNB_C_CHAIN1D) {
        if (which == 1) { rc = gnb_chain1d_surface_g(c, ct, m, dE); src = c->cg.as<cplx>(); }
        else { rc = gnb_contact_eval(c, ct, m, dE, 0); src = ct.blk.as<cplx>(); }
        per = nn;
    } else {
        if (which == 0) { rc = gnb_contact_eval(c, ct, m, dE, 0); src = ct.blk.as<cplx>(); per = nn; }
        else {
            per = (which == 1 ? 12 : 9) * BSZ;
            GNB_CK(ct.surf.ensure((size_t)m * 12 * BSZ * sizeof(cplx)));
            rc = gnb_bethe_raw(c, ct, m, dE, which, ct.surf.as<cplx>());
            src = ct.surf.as<cplx>();
        }
    }
    if (rc) return rc;
    GNB_CK(cudaMemcpyAsync(out_blk, src, (size_t)m * per * sizeof(cplx), cudaMemcpyDeviceToHost, c->stream));
    if (iters) GNB_CK(cudaMemcpyAsync(iters, ct.iters.p, (size_t)m * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    if (diffs) GNB_CK(cudaMemcpyAsync(diffs, ct.diffs.p, (size_t)m * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    GNB_CK(cudaStreamSynchronize(c->stream));
    return GNB_OK;
}

extern "C" int gnb_sigma_eval(gnb_ctx* c, int contact, int which, int M, const double* E, double* out_blk,
                              int32_t* iters, double* diffs) {
    if (!c || contact < 0 || contact >= (int)c->contacts.size() || M < 0 || (M > 0 && (!E || !out_blk)))
        return gnb_fail(c, GNB_ERR_ARG, "sigma_eval: bad arguments");
    cudaSetDevice(c->device);
    if (M == 0) return GNB_OK;
    Contact& ct = c->contacts[contact];
    GNB_CK(c->info.ensure(sizeof(int) * 4));
    GNB_CK(cudaMemsetAsync(c->info.p, 0, sizeof(int) * 4, c->stream));
    const long nn = (long)ct.nc * ct.nc;
    if (ct.kind == GNB_C_CONST) {
        std::vector<cplx> h(nn);
        GNB_CK(cudaMemcpyAsync(h.data(), ct.d_const.p, nn * sizeof(cplx), cudaMemcpyDeviceToHost, c->stream));
        GNB_CK(cudaStreamSynchronize(c->stream));
        for (int b = 0; b < M; b++) memcpy(out_blk + (size_t)b * nn * 2, h.data(), nn * sizeof(cplx));
        if (iters) std::fill(iters, iters + M, 0);
        if (diffs) std::fill(diffs, diffs + M, 0.0);
        return GNB_OK;
    }
    // energies in chunks: bounded workspace (the chain fixed point keeps ~10 n_c x n_c matrices per energy) and
    // gridDim.y = chunk <= 8192
    const size_t per_out = ct.kind == GNB_C_CHAIN1D ? (size_t)nn : (which == 0 ? (size_t)nn : (size_t)(which == 1 ? 12 : 9) * BSZ);
    const size_t per_ws = (ct.kind == GNB_C_CHAIN1D ? 12 * (size_t)nn : (size_t)nn + 12 * BSZ) * sizeof(cplx) + 256;
    size_t mc = std::max<size_t>(1, std::min<size_t>(c->ws_limit / per_ws, 8192));
    const size_t nchunks = ((size_t)M + mc - 1) / mc;
    mc = ((size_t)M + nchunks - 1) / nchunks;
    for (int k0 = 0; k0 < M; k0 += (int)mc) {
        const int m = std::min<int>((int)mc, M - k0);
        GNB_CK(c->dE.ensure((size_t)m * sizeof(cplx)));
        GNB_CK(cudaMemcpyAsync(c->dE.p, E + 2 * (size_t)k0, (size_t)m * sizeof(cplx), cudaMemcpyHostToDevice, c->stream));
        int rc = sigma_eval_chunk(c, ct, which, m, out_blk + (size_t)k0 * per_out * 2, iters ? iters + k0 : nullptr,
                                  diffs ? diffs + k0 : nullptr);
        if (rc) return rc;
    }
    int info = 0;
    GNB_CK(cudaMemcpyAsync(&info, c->info.p, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    GNB_CK(cudaStreamSynchronize(c->stream));
    if (info) return gnb_fail(c, GNB_ERR_SINGULAR, "Singular matrix");
    return GNB_OK;
}
