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
    const double* Pr;            // optional: P given as REAL doubles (mixed layout of gnb_rec.cu), replaces P; k_gemm only
    long stridePr; int ldpr;     // batch stride / row stride of Pr in doubles
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
void gnb_set_tourn_group(int g);
void gnb_set_tourn_warp(int on);
void gnb_launch_assemble(cudaStream_t st, int M, cplx* A, long strideA, int ld, int N, const cplx* F,
                         const cplx* S, const cplx* Sig0, const cplx* SigB, long strideSigB, const cplx* E,
                         const int* pi = nullptr, int mixr = 0);
void gnb_launch_scatter_sub(cudaStream_t st, int M, cplx* A, long strideA, int ld, const int* inds, int nc,
                            const cplx* blk, long strideBlk, const int* map = nullptr);
void gnb_launch_set_aug(cudaStream_t st, int M, cplx* A, long strideA, int ld, int N, int xoff, const int* cols, int m,
                        const int* map = nullptr);
void gnb_launch_pad_diag(cudaStream_t st, int M, cplx* A, long strideA, int ld, int N, int Np);
void gnb_launch_gemm(cudaStream_t st, const GnbGemmArgs& g, int nbatch, bool wt, bool batchk);
long gnb_eliminate(cudaStream_t st, int M, int N, int naug, cplx* A, long strideA, int ld, int jordan,
                   const GnbElimWork& ws);

long gnb_launch_tournament(cudaStream_t st, int M, int N, const cplx* A, long strideA, int ld, int c0, int w,
                           int* cand0, int* cand1, int cand_stride, cplx* LU, int* moves, int* perm, int perm_stride,
                           int* info, int real_panel = 0, int mixr = 0);
void gnb_launch_init_perm(cudaStream_t st, int M, int* perm, int stride, int N);

// gnb_rec.cu : recursive (multi-level) elimination on a padded layout; the rank-K updates run on the
// warp-specialised packed-operand DMMA kernel.  N and naug must be multiples of 32, ld = N + naug.
#define RK_PPS 20                      // row stride (cplx) of a packed panel block [32 rows][16 k]
#define RK_WPS 34                      // row stride (cplx) of a packed W block     [16 k][32 cols]
#define RK_PBLK (32 * RK_PPS)
#define RK_WBLK (16 * RK_WPS)
#define RK_PRS 20                      // real-packed panel block  [32 rows][16 k] doubles, row stride 20
#define RK_WRS 36                      // real-packed W block      [16 k][32 cols] doubles, row stride 36
#define RK_PRBLK (32 * RK_PRS)
#define RK_WRBLK (16 * RK_WRS)
struct GnbRecWork {
    GnbGemmTimer* timer;
    int* cand0; int* cand1; int cand_stride;
    cplx* inv; long inv_blk_stride;    // [N/32][Mtot][32][32] inverse of every pivot block (block stride in cplx)
    int* moves; long moves_blk_stride; // [N/32][Mtot][GNB_MOVES_STRIDE]                (block stride in ints)
    int* perm; int perm_stride;        // JORDAN: running row permutation
    cplx* Ppk; cplx* Lpk; long stridePk;   // packed panels [M][N/16][N/32][32][RK_PPS]
    cplx* Wpk; long strideWk;          // packed pivot rows [M][N/16][ld/32][16][RK_WPS]
    double* PpkR; double* WpkR;        // real-packed panels / pivot rows of the real columns (mixed layout); strides below
    long stridePkR, strideWkR;
    int* info;
    int back_row_lo;                   // FORWARD: only rows >= back_row_lo of the solution are needed
    int nreal;                         // FORWARD: columns [0, nreal) of the matrices are real (0 = unknown / complex)
    int mixr;                          // mixed layout: columns [0, mixr) stored as real doubles (0 or == nreal)
    double* flops_acc;                 // host accumulator: executed real FP64 flops of the rank-K / forward-W /
                                       // back-substitution launches (4-multiplication equivalents; may be null)
    cudaStream_t side;                 // look-ahead: second stream of this (sub-)batch, null = none
    cudaEvent_t la_fork, la_join;      // look-ahead fork / join events
};
size_t gnb_rec_pk_elems(int N);               // cplx elements per matrix of Ppk / Lpk (incl. slack)
size_t gnb_rec_wk_elems(int N, int ld);       // cplx elements per matrix of Wpk (incl. slack)
cudaError_t gnb_rec_init();
void gnb_rec_set_option(const char* name, int value);
int gnb_rec_real_enabled();
long gnb_eliminate_rec(cudaStream_t st, int M, int N, int naug, cplx* A, long strideA, int ld, int jordan,
                       const GnbRecWork& ws);

// gnb_small.cu : one CTA per energy, matrix resident in shared memory (N <= GNB_SMALL_MAX_N)
#define GNB_SMALL_MAX_N 119            // N*(N|1)*16 B + bookkeeping <= 227 KB
#define GNB_SMALL_REG_MAX_N 96         // register-resident variant: 32*RA rows x NW*CB columns over 32*NW threads
#define GNB_SMALL_CLUSTER_MAX_N 192     // thread-block-cluster register-resident inverse: 2 CTAs to 128, 4 CTAs to 192 (GREEN mode)
#define GNB_SMALL_MAX_CONTACTS 6
enum { GNB_SMALL_GREEN = 0, GNB_SMALL_DOS = 1, GNB_SMALL_T = 2 };
struct GnbSmallContact {
    const int* inds; int nc;
    const cplx* blk; long blk_stride;      // Sigma block per energy (stride 0: energy independent)
    const cplx* gam; long gam_stride;      // Gamma block (mode T only)
};
struct GnbSmallArgs {
    int N, M, mode;
    const cplx *F, *S, *Sig0, *SigB; long strideSigB;   // Sig0: constant dense (or null); SigB: per energy (or null)
    const cplx* E;
    const cplx* Araw;                      // non-null: invert these [M][N][N] matrices instead of assembling
    int ncontacts; GnbSmallContact ct[GNB_SMALL_MAX_CONTACTS];
    cplx* G; long strideG; int ldg;        // GREEN
    double* dos_tot; double* dos_site;     // DOS (dos_site may be null)
    int ca, cb; double* T;                 // T
    int* info;
    int cl_relaxed;                        // cluster kernels: non-owner warps arrive at the column barrier without release semantics (set by the launcher)
};
cudaError_t gnb_small_init();
void gnb_small_set_reg(int on);        // developer switch "small_reg": register-resident (1) or shared-memory (0) kernel
int gnb_small_max_n();                 // 96 (register-resident kernel, default) or 119 (small_reg=0)
int gnb_small_inverse_max_n();         // largest n whose plain inverse runs on chip (192 with the cluster kernels)
int gnb_small_cluster_max_m(int n);    // largest batch for which the cluster kernels are preferred to the block engine
void gnb_small_set_wide(int on);       // developer switch "small_wide"
void gnb_small_set_cluster(int on);    // developer switch "small_cluster"
void gnb_small_set_cluster_max_m(int m);   // developer switch "small_cluster_maxm"
void gnb_small_set_cl_relaxed(int on);     // developer switch "small_cl_relaxed"
int gnb_small_enabled();               // developer switch "small_fused" (gnb_api.cu)
void gnb_launch_small(cudaStream_t st, const GnbSmallArgs& a);

// gnb_reduce.cu
void gnb_launch_invperm(cudaStream_t st, int M, const int* perm, int* invperm, int stride, int N);
void gnb_launch_weighted_sum(cudaStream_t st, int M, int N, const cplx* A, long strideA, int ld,
                             const int* invperm, int pstride, const cplx* w, cplx* out, int accumulate);
void gnb_launch_unpermute(cudaStream_t st, int M, int N, const cplx* A, long strideA, int ld,
                          const int* invperm, int pstride, cplx* G, long strideG);
void gnb_launch_dos(cudaStream_t st, int M, int N, const cplx* A, long strideA, int ld,
                    const int* invperm, int pstride, double* tot, double* per_site);
void gnb_launch_gather_rows(cudaStream_t st, int M, const cplx* X, long strideX, int ldx, const int* rows,
                            int nr, int ncols, cplx* out, long strideOut, const int* map = nullptr);
void gnb_launch_trace_dot(cudaStream_t st, int M, const cplx* Z, const cplx* X, long stride, int n, double* T);
void gnb_launch_gamma_from_sigma(cudaStream_t st, int M, const cplx* sig, long stride, int n, cplx* gam);
void gnb_launch_unpermute_sym(cudaStream_t st, int N, const cplx* in, const int* pi, cplx* out);
void gnb_launch_scale_cols(cudaStream_t st, int M, cplx* X, long stride, int n, const cplx* w);
void gnb_launch_xi_gather(cudaStream_t st, int n, const cplx* Xi, const int* inds, int nc, cplx* U, cplx* V);
void gnb_launch_scatter_add(cudaStream_t st, int M, cplx* dense, long strideD, int n, const int* inds, int nc,
                            const cplx* blk, long strideBlk);
void gnb_launch_kron_expand(cudaStream_t st, int M, int n, int mode, const cplx* in, cplx* out);
void gnb_launch_trace_dot_strided(cudaStream_t st, int M, const cplx* Z, long strideZ, int ldz, const cplx* X,
                                  long strideX, int ldx, int nr, int ncols, double* T, int tstride, int toff);
