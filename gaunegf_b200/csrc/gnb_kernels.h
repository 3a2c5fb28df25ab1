// Internal launcher declarations shared by the .cu translation units of libgaunegf_b200.so.
#pragma once
#include <cuda_runtime.h>

typedef double2 cplx;

#ifndef GNB_NB
#define GNB_NB 32          // elimination block width (pivot block is NB x NB)
#endif
#define GNB_MOVES_STRIDE 132   // 1 + 2 * (2*NB) ints per matrix, padded

struct GnbGemmArgs {
    cplx* C; long strideC; int ldc;
    const cplx* P; long strideP; int ldp;
    const cplx* W; long strideW; int ldw;
    int ilo, ihi, jlo, jhi, kdim;
    int skip_lo, skip_hi;        // rows in [skip_lo, skip_hi) are computed but not written
    int zero_init, plus;         // C = 0 before accumulation; C += (plus) or C -= (minus)
    int nbatch_k;                // BATCHK: number of batches folded into K
    const cplx* wscale;          // optional per-batch complex scale applied to P
};

// Optional per-launch CUDA-event timing of the rank-K update kernel (bench.py's roofline leg).
struct GnbGemmTimer {
    virtual void begin(cudaStream_t st) = 0;
    virtual void end(cudaStream_t st, double flops) = 0;
    virtual ~GnbGemmTimer() {}
};

struct GnbElimWork {
    GnbGemmTimer* timer;                       // nullptr = no per-kernel timing
    int* cand0; int* cand1; int cand_stride;   // tournament candidate lists
    cplx* LU;                                  // [2][M][NB][NB] inverse of the pivot block (two slots)
    int* moves;                                // [2][M][GNB_MOVES_STRIDE]
    int* perm; int perm_stride;                // [M][N] running row permutation (JORDAN)
    cplx* Pws;                                 // [M][N][2*NB] saved panel columns (JORDAN)
    int* info;                                 // device flag: 1 = exactly singular pivot met
};

cudaError_t gnb_kernels_init();
void gnb_set_gemm_bm(int bm);
void gnb_set_gemm_pipe(int on);
void gnb_set_two_level(int on);
void gnb_launch_assemble(cudaStream_t st, int M, cplx* A, long strideA, int ld, int N, const cplx* F,
                         const cplx* S, const cplx* Sig0, const cplx* SigB, long strideSigB, const cplx* E);
void gnb_launch_scatter_sub(cudaStream_t st, int M, cplx* A, long strideA, int ld, const int* inds, int nc,
                            const cplx* blk, long strideBlk);
void gnb_launch_set_aug(cudaStream_t st, int M, cplx* A, long strideA, int ld, int N, const int* cols, int m);
void gnb_launch_gemm(cudaStream_t st, const GnbGemmArgs& g, int nbatch, bool wt, bool batchk);
long gnb_eliminate(cudaStream_t st, int M, int N, int naug, cplx* A, long strideA, int ld, int jordan,
                   const GnbElimWork& ws);

// gnb_reduce.cu
void gnb_launch_invperm(cudaStream_t st, int M, const int* perm, int* invperm, int stride, int N);
void gnb_launch_weighted_sum(cudaStream_t st, int M, int N, const cplx* A, long strideA, int ld,
                             const int* invperm, int pstride, const cplx* w, cplx* out, int accumulate);
void gnb_launch_unpermute(cudaStream_t st, int M, int N, const cplx* A, long strideA, int ld,
                          const int* invperm, int pstride, cplx* G, long strideG);
void gnb_launch_dos(cudaStream_t st, int M, int N, const cplx* A, long strideA, int ld,
                    const int* invperm, int pstride, double* tot, double* per_site);
void gnb_launch_gather_rows(cudaStream_t st, int M, const cplx* X, long strideX, int ldx, const int* rows,
                            int nr, int ncols, cplx* out, long strideOut);
void gnb_launch_trace_dot(cudaStream_t st, int M, const cplx* Z, const cplx* X, long stride, int n, double* T);
void gnb_launch_gamma_from_sigma(cudaStream_t st, int M, const cplx* sig, long stride, int n, cplx* gam);
void gnb_launch_scale_cols(cudaStream_t st, int M, cplx* X, long stride, int n, const cplx* w);
void gnb_launch_trace_dot_strided(cudaStream_t st, int M, const cplx* Z, long strideZ, int ldz, const cplx* X,
                                  long strideX, int ldx, int nr, int ncols, double* T, int tstride, int toff);
