/* libgaunegf_b200 — C ABI of the B200-native energy-grid Green's-function path of GauNEGF.
 *
 * The reference (wliverno/GauNEGF) is pure Python and defines no FFI; its boundary for this path is
 * the Python module API + the duck-typed surfG protocol.  These are the entry points
 * a maintainer would bind from those Python functions (ctypes stubs: INTEGRATION.md).  Each entry
 * cites the reference code it replaces (paths under gauNEGF/).
 *
 * Conventions
 *   - complex128 arrays are interleaved (re, im) doubles, matrices row-major (numpy default).
 *   - every function returns 0 on success, a GNB_ERR_* code otherwise; gnb_last_error() gives text.
 *     Nothing throws across the boundary.
 *   - `loc` arguments: GNB_HOST (pointer to host memory; copies are made inside the call) or
 *     GNB_DEVICE (pointer to memory of the context's GPU; used in place on the context's stream).
 *   - the caller owns every buffer it passes; the library owns its workspace; one context per host
 *     thread and GPU.  Energies / weights / index lists are always host pointers.
 *   - no CPU fallback exists: without a CUDA device gnb_create() fails.
 */
#ifndef GAUNEGF_B200_H
#define GAUNEGF_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct gnb_ctx gnb_ctx;

enum { GNB_HOST = 0, GNB_DEVICE = 1 };

enum {
    GNB_OK = 0,
    GNB_ERR_CUDA = 1,        /* CUDA runtime error (text in gnb_last_error) */
    GNB_ERR_ARG = 2,         /* invalid argument / call order */
    GNB_ERR_SINGULAR = 3,    /* an exactly singular pivot was met (numpy raises LinAlgError here) */
    GNB_ERR_NOMEM = 4
};

/* ---- lifetime ------------------------------------------------------------------------------ */
int gnb_create(gnb_ctx** out, int device);
int gnb_destroy(gnb_ctx* ctx);
const char* gnb_last_error(const gnb_ctx* ctx);
const char* gnb_version(void);
/* cudaStream_t to launch on (e.g. torch.cuda.current_stream().cuda_stream); default: stream 0 */
int gnb_set_stream(gnb_ctx* ctx, void* cuda_stream);
/* upper bound for the per-call energy-chunk workspace in bytes (default: 60 % of the GPU's memory) */
int gnb_set_workspace_limit(gnb_ctx* ctx, size_t bytes);
/* kernels launched by this context since creation (bench.py's gpu_launches) */
int64_t gnb_launch_count(const gnb_ctx* ctx);
/* device-side milliseconds spent inside the elimination launches of the last compute call
 * (CUDA events on the context's stream; bench.py's roofline leg) */
double gnb_last_elim_ms(const gnb_ctx* ctx);
/* executed real FP64 flops (4-multiplication equivalents: 8 per complex multiply-add, 4 / 2 where one / both
 * operands are known to be real) of the rank-K update, forward-W and back-substitution launches of the last
 * compute call (recursive engine); counted on the host at launch time, no timing needed (bench.py's step_frac) */
double gnb_last_elim_flops(const gnb_ctx* ctx);
/* enable/disable the event timing above (adds one event synchronisation per energy chunk) */
int gnb_set_timing(gnb_ctx* ctx, int on);
/* with timing on: accumulated device time (ms, CUDA events around each launch), algorithmic real
 * flops (8 per complex multiply-add) and launch count of the rank-K update kernel (the dominant
 * kernel) inside eliminations since the last reset */
int gnb_gemm_stats(gnb_ctx* ctx, double* ms, double* flops, int64_t* launches, int reset);

/* ---- system: F, S  (jnp.asarray(F), jnp.asarray(S): integrate.py:92-93, transport.py:418-419) -- */
int gnb_set_system(gnb_ctx* ctx, int N, const double* F, const double* S, int loc);
/* The same for HOST arrays that usually repeat between calls (the reference passes F and S to every integrator call,
 * integrate.py:92-95; an SCF step changes F but not S): both are compared with pinned shadow copies kept by the context
 * (one multi-threaded pass that also detects real-valued input) and only the matrix that changed goes over PCIe.  F and S
 * are fully consumed when the call returns.  real_input: bit 0 = F holds N*N real doubles, bit 1 = S does (restricted-spin
 * Gaussian output; otherwise complex128).  uploaded (may be NULL): bit 0 = F was sent, bit 1 = S was sent. */
int gnb_set_system_cached(gnb_ctx* ctx, int N, const double* F, const double* S, int real_input, int* uploaded);
/* Read-only test whether (F, S) differ from the resident pair of gnb_set_system_cached.  full = 1: every element; full = 0:
 * size and a strided sample only.  With one process per GPU every rank receives the same F and S: rank 0 runs the full
 * comparison, the others the sample, and one flag is agreed on (gaunegf_b200/parallel.py) instead of N full passes over host
 * memory that all ranks of a box share.  differs: bit 0 = F differs, bit 1 = S differs (3 when there is nothing resident). */
int gnb_system_differs(gnb_ctx* ctx, int N, const double* F, const double* S, int real_input, int full, int* differs);
/* gnb_set_system_cached without the comparison, for a caller that knows which matrices changed (changed: bit 0 = F,
 * bit 1 = S, e.g. the flags the ranks agreed on): those are copied into the shadows and uploaded, the others stay. */
int gnb_set_system_known(gnb_ctx* ctx, int N, const double* F, const double* S, int real_input, int changed, int* uploaded);

/* ---- self-energy description ------------------------------------------------------------------
 * Sigma_tot(E) = Sigma0 + sum_c scatter(inds_c, blk_c(E)).  Contacts are numbered 0..nc-1; the
 * reference's contact index -1 is the last one.
 */
int gnb_sigma_clear(gnb_ctx* ctx);
/* energy-independent dense N x N term added to Sigma_tot (surfGTester.py:113-132; transport.py:83-89) */
int gnb_sigma_set_dense0(gnb_ctx* ctx, const double* sig0, int loc);
/* energy-independent contact: nc x nc block on orbitals inds (transport.py:86-87 np.diag(vector),
 * matTools.py:69-72 formSigma blocks).  Gamma_c = i (blk - blk^H) is kept compact. */
int gnb_sigma_add_const_block(gnb_ctx* ctx, int nc, const int32_t* inds, const double* blk);
/* 1-D chain contact (surfG1D.py:223-295, 344-373): g <- relax*inv(A - B g B^H) + (1-relax) g from
 * g0 = inv(A), A = (E+i eta) Salpha - alpha, B = (E+i eta) Sbeta - beta, stop when
 * max|dg|/max(|g_new|,1e-12) <= conv or after max_iter; Sigma = t g t^H, t = E stau - tau.
 * alpha, Salpha, beta, Sbeta: nc x nc; tau, stau: nc x nc (rows: coupling orbitals) */
int gnb_sigma_add_chain1d(gnb_ctx* ctx, int nc, const int32_t* inds, const double* alpha,
                          const double* Salpha, const double* beta, const double* Sbeta,
                          const double* tau, const double* stau, double eta, double conv,
                          double relax, int max_iter);
/* Bethe-lattice contact (surfGBethe.py:479-542, 958-1108): natoms atoms of 9 orbitals each
 * (inds: natoms*9), per atom the list of connected directions (nb_dirs, offsets nb_off[natoms+1]);
 * H 9x9, Slist/Vlist 12 x 9x9 (complex interleaved); bulk+surface fixed points with mixing `mix`. */
int gnb_sigma_add_bethe(gnb_ctx* ctx, int natoms, const int32_t* inds, const int32_t* nb_off,
                        const int32_t* nb_dirs, const double* H, const double* Slist,
                        const double* Vlist, double eta, double conv, double mix, int max_iter);

/* De-orthonormalisation and spin expansion of the contact self-energies (surfGBethe.py:529-539):
 *   Sigma_tot(E) = expand( Xi [ sum_c scatter(inds_c, blk_c(E)) ] Xi ),
 * Xi: n x n (NULL = identity; the reference uses S^(1/2) when the Bethe parameters are orthonormal), spin_mode 0: none
 * (n = N), 1: kron(I2, .) ('u', 'ro'), 2: kron(., I2) ('g') with n = N/2.  Call after gnb_sigma_clear and before
 * adding contacts, whose orbital indices then refer to the n-dimensional space.  The drivers build the dense
 * Sigma(E_k) / Gamma(E_k) on the device and run the full-inverse algorithms. */
int gnb_sigma_set_transform(gnb_ctx* ctx, int n, const double* Xi, int spin_mode, int loc);

/* ---- Sigma(E) providers on their own (surfG.g / sigma, surfGBAt.sigmaK / sigma) ----------------
 * out_blk: M x nc x nc compact contact blocks (host); iters/diffs: per energy iteration count and
 * last convergence measure (may be NULL).  which: 0 = contact self-energy block,
 * 1 = surface Green's function g (chain1d) / 12 bulk sigmaK blocks (bethe: out is M x 12 x 81),
 * 2 = bethe 9 surface blocks (M x 9 x 81). */
int gnb_sigma_eval(gnb_ctx* ctx, int contact, int which, int M, const double* E, double* out_blk,
                   int32_t* iters, double* diffs);

/* ---- per-energy reductions (E: M complex energies, host) --------------------------------------- */
/* full G(E_k) = (E_k S - F - Sigma_tot)^-1, out: M x N x N   (utils.py:52-54, integrate.py:67-71) */
int gnb_green(gnb_ctx* ctx, int M, const double* E, double* G, int loc);
/* T(E_k) = Re Tr[Gamma_a G Gamma_b G^H] between contacts ca and cb (transport.py:150-157), using
 * only the contact columns of G (low-rank Gamma).  T: M doubles (host). */
int gnb_transmission(gnb_ctx* ctx, int M, const double* E, int ca, int cb, double* T);
/* -Im diag(G)/pi and its sum (transport.py:183-190; density.py:49-54). per_site may be NULL. */
int gnb_dos(gnb_ctx* ctx, int M, const double* E, double* dos_total, double* dos_per_site);
/* out = sum_k w_k G(E_k)   (integrate.GrInt, integrate.py:146-173).  out: N x N complex. */
int gnb_gr_int(gnb_ctx* ctx, int M, const double* E, const double* w, double* out, int loc);
/* nseg such sums over consecutive energy ranges in ONE batch: out[s] = sum of w_k G(E_k) for seg_end[s-1] <= k <
 * seg_end[s] (seg_end non-decreasing, seg_end[nseg-1] = M).  out: nseg x N x N complex.  Serves the nested adaptive
 * quadratures (density.integratePointsAdaptiveANT, density.py:211-273: the nodes of the next levels are known before
 * the convergence test of the current one, so several levels share one launch chain). */
int gnb_gr_int_seg(gnb_ctx* ctx, int M, const double* E, const double* w, int nseg, const int32_t* seg_end,
                   double* out, int loc);
/* out = sum_k w_k G Gamma G^H  (integrate.GrLessInt, integrate.py:177-208);
 * contact >= 0: Gamma of that contact; contact = -1: Gamma of Sigma_tot (ind=None). */
int gnb_gless_int(gnb_ctx* ctx, int M, const double* E, const double* w, int contact, double* out, int loc);

/* ---- generic dense variants: any Python sigma callable evaluated per energy by the caller ------
 * sig: Sigma_tot, gam*: Gamma matrices, each either one N x N matrix (stride 0) or M of them
 * (stride N*N complex elements); all host pointers.  These run the full inverse. */
int gnb_green_dense(gnb_ctx* ctx, int M, const double* E, const double* sig, long sig_stride, double* G, int loc);
int gnb_transmission_dense(gnb_ctx* ctx, int M, const double* E, const double* sig, long sig_stride,
                           const double* gam1, long g1_stride, const double* gam2, long g2_stride, double* T);
/* spin-resolved transmission (transport.py:159-181): the system is 2N x 2N in block order; T4: M x 4
 * = [T_uu, T_ud, T_du, T_dd] with T_i = Re Tr[Gamma1[r,r] G[r,c] Gamma2[c,c] Ga[r,c]].
 * sig = gam1 = gam2 = NULL: use the described self-energies (needs gnb_sigma_set_transform with a spin mode),
 * contacts 0 and -1. */
int gnb_transmission_spin(gnb_ctx* ctx, int M, const double* E, const double* sig, long sig_stride,
                          const double* gam1, long g1_stride, const double* gam2, long g2_stride, double* T4);
int gnb_dos_dense(gnb_ctx* ctx, int M, const double* E, const double* sig, long sig_stride,
                  double* dos_total, double* dos_per_site);
int gnb_gr_int_dense(gnb_ctx* ctx, int M, const double* E, const double* w, const double* sig,
                     long sig_stride, double* out, int loc);
int gnb_gr_int_seg_dense(gnb_ctx* ctx, int M, const double* E, const double* w, int nseg, const int32_t* seg_end,
                         const double* sig, long sig_stride, double* out, int loc);
int gnb_gless_int_dense(gnb_ctx* ctx, int M, const double* E, const double* w, const double* sig,
                        long sig_stride, const double* gam, long gam_stride, double* out, int loc);

/* ---- utils.inv drop-in: batched inverse of M arbitrary n x n complex matrices (utils.py:52-54) -- */
int gnb_inverse_batch(gnb_ctx* ctx, int n, int M, const double* A, double* Ainv, int loc);

#ifdef __cplusplus
}
#endif
#endif /* GAUNEGF_B200_H */
