/* Developer switches and probes of libgaunegf_b200 — NOT part of the drop-in ABI (include/gaunegf_b200.h).
 *
 * They exist for A/B measurements of kernel variants (tools/, tests/test_gpu_engine.py) and are PROCESS-WIDE:
 * a switch set here applies to every context of the process, so they must not be flipped while another host thread
 * is inside a library call.  Product code never calls them (GNB_DEV_OPTS in the environment is the only entry).
 */
#ifndef GAUNEGF_B200_DEV_H
#define GAUNEGF_B200_DEV_H

#include "gaunegf_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* name = value; names: engine_rec, rec_streams, rec_stream_min_m, rec_stagger_us, small_fused, small_reg, small_cluster,
 * small_cluster_maxm, small_cl_relaxed, small_wide, chain_joint, chain_compact, contacts_last, mixed_layout, gless_mixed,
 * tourn_fp32, tourn_warp (bit mask of the tournament kernels, gnb_elim.cu), tournq_cplx_min_m, two_level, gemm_pipe, gemm_bm,
 * rk_* (gnb_rec.cu: rk_m3, rk_m3_mink, rk_kskip, rk_strip, rk_real, rk_wsolve_mma, rk_wsolve_areal, rk_wsolve_fused, rk_rp2,
 * rk_fin_mma, rk_sms, rk_wskip, rk_augreal, rk_cs, rk_tcap_k, rk_lowprio, rk_lookahead, rk_la_mink, rk_la_ctas).  Defaults are
 * the measured best; DESIGN.md section 5 lists what each one measured. */
int gnb_dev_set_option(const char* name, int value);
/* CUDA-event trace of every launch of the recursive engine (tools/trace_elim.py) */
int gnb_dev_trace_start(void);
int gnb_dev_trace_dump(const char* path);
/* time the unpacked-operand DMMA GEMM alone on zero data (tools/gemm_bench.py) */
int gnb_dev_gemm_bench(gnb_ctx* ctx, int M, int n, int k, int bm, int iters, double* ms_out);
/* FP64 tensor-pipe ceiling of this GPU: a register-resident DMMA.8x8x4 issue loop on every SM for about `ms_target`
 * milliseconds; returns TFLOP/s (2 * 8 * 8 * 4 flops per warp-wide DMMA) in *tflops (bench.py's roofline peak) */
int gnb_dev_fp64_peak(gnb_ctx* ctx, double ms_target, double* tflops);

#ifdef __cplusplus
}
#endif
#endif /* GAUNEGF_B200_DEV_H */
