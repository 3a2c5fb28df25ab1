// Context object behind the C ABI (include/gaunegf_b200.h).  Internal to the library.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <vector>
#include "gnb_kernels.h"

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) { cudaFree(p); p = nullptr; cap = 0; }
        size_t want = bytes + bytes / 8;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) { cudaGetLastError(); e = cudaMalloc(&p, bytes); want = bytes; }
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};

enum { GNB_C_CONST = 0, GNB_C_CHAIN1D = 1, GNB_C_BETHE = 2 };

struct Contact {
    int kind = GNB_C_CONST;
    int nc = 0;
    std::vector<int> h_inds;
    DevBuf d_inds;
    DevBuf d_const;                 // nc*nc block (CONST)
    // CHAIN1D parameters (device, nc*nc each)
    DevBuf alpha, Salpha, beta, Sbeta, tau, stau;
    double eta = 0, conv = 0, relax = 0, mix = 0;
    int max_iter = 0;
    // BETHE parameters
    int natoms = 0;
    std::vector<int> nb_off, nb_dirs;
    DevBuf d_nb_off, d_nb_dirs, H, Slist, Vlist;
    // per-chunk products
    DevBuf blk, gam, iters, diffs, surf;
    DevBuf xiU, xiV;                // Xi[:, inds] (n x nc) and Xi[inds, :] (nc x n) when a transform is set
    const cplx* blk_ptr = nullptr; long blk_stride = 0;   // valid after prepare
    const cplx* gam_ptr = nullptr; long gam_stride = 0;
};

struct EventTimer : GnbGemmTimer {
    std::vector<cudaEvent_t> pool;
    std::vector<double> flops;
    size_t used = 0;
    double total_ms = 0.0, total_flops = 0.0;
    long count = 0;
    cudaEvent_t get() {
        if (used == pool.size()) { cudaEvent_t e; cudaEventCreate(&e); pool.push_back(e); }
        return pool[used++];
    }
    void begin(cudaStream_t st) override { cudaEventRecord(get(), st); }
    void end(cudaStream_t st, double fl) override { cudaEventRecord(get(), st); flops.push_back(fl); }
    void resolve() {          // call after the stream is synchronised
        for (size_t i = 0; i + 1 < used; i += 2) {
            float ms = 0;
            if (cudaEventElapsedTime(&ms, pool[i], pool[i + 1]) == cudaSuccess) {
                total_ms += ms; total_flops += flops[i / 2]; count++;
            }
        }
        used = 0; flops.clear();
    }
    void reset() { used = 0; flops.clear(); total_ms = total_flops = 0.0; count = 0; }
    ~EventTimer() { for (auto e : pool) cudaEventDestroy(e); }
};

#define GNB_MAX_SUBSTREAMS 8
struct gnb_ctx {
    EventTimer gemm_timer;
    cudaStream_t sub[GNB_MAX_SUBSTREAMS] = {};
    cudaEvent_t sub_ev[GNB_MAX_SUBSTREAMS] = {};
    cudaStream_t side[GNB_MAX_SUBSTREAMS] = {};      // look-ahead streams of the sub-batches (gnb_rec.cu)
    cudaEvent_t la_fork[GNB_MAX_SUBSTREAMS] = {}, la_join[GNB_MAX_SUBSTREAMS] = {};
    cudaEvent_t fork_ev = nullptr;
    int device = 0;
    cudaStream_t stream = 0;
    size_t ws_limit = (size_t)48 << 30;
    std::string err;
    int64_t launches = 0;
    int N = 0;
    DevBuf dF, dS, dSig0;
    bool has_sig0 = false;
    bool real_FS = false;          // F and S given with zero imaginary parts (host arrays only)
    // Sigma transform (gnb_sigma_set_transform): contact indices refer to an n-dimensional space,
    // Sigma_tot = expand(Xi Sigma Xi); sig_n = 0 means none
    int sig_n = 0, sig_spin = 0;
    bool has_xi = false;
    DevBuf dXi, sigN, xiY;
    std::vector<Contact> contacts;
    std::vector<Contact> contact_pool;  // retired descriptions whose device buffers are reused (gnb_api.cu)
    // workspaces
    DevBuf A, Pws, LU, moves, cand0, cand1, perm, invperm, info, dE, dW, G, Y, Z, Xr, out, dT, dDosT, dDosP,
        sigB, gam1B, gam2B, cols, rows, in_stage, Ppk, Lpk, Wpk, PpkR, WpkR;
    // chain1d fixed-point workspaces
    DevBuf cA, cB, cg, cgn, cT1, cM, cflags, ct, cgw, cA2, cB2, cgw2;
    // pinned landing buffer for N x N results going to (pageable) caller memory: one DMA at PCIe rate + a host
    // memcpy after the synchronisation in end_call, instead of the driver's chunked pageable staging
    void* h_pin = nullptr; size_t h_pin_cap = 0;
    void* pend_dst = nullptr; size_t pend_bytes = 0;
    // pinned host shadows of the resident F and S (gnb_set_system_cached): shadow_N = size they mirror, 0 = none
    void* hF = nullptr; void* hS = nullptr; size_t shadow_cap = 0; int shadow_N = 0; cudaEvent_t shadow_ev = nullptr;
    std::vector<char> realF_slice, realS_slice;      // per host-thread slice: imaginary parts all zero
    // segmented GrInt (gnb_gr_int_seg): host array of cumulative segment ends, valid during the call only
    int seg_n = 0; const int* seg_end = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    double elim_ms = 0.0;
    double elim_flops = 0.0;        // executed FP64 flops of the elimination launches of the last compute call
    bool timing = false;
};

// gnb_sigma.cu
cudaError_t gnb_sigma_init();
int gnb_contact_eval(gnb_ctx* c, Contact& ct, int M, const cplx* dE, int want_gamma, const cplx* g_ready = nullptr);
void gnb_chain_set_compact(int on);
int gnb_chain1d_surface_g_multi(gnb_ctx* c, Contact* const* cts, int K, int M, const cplx* dE);   // g_k in c->cg + k*M*nc*nc
int gnb_chain1d_surface_g(gnb_ctx* c, Contact& ct, int M, const cplx* dE);     // leaves g in c->cg
int gnb_bethe_raw(gnb_ctx* c, Contact& ct, int M, const cplx* dE, int which, cplx* d_out);
GnbElimWork gnb_elim_work(gnb_ctx* c, int M, int N, bool jordan, int* rc);
GnbRecWork gnb_rec_work(gnb_ctx* c, int M, int Np, int ld, bool jordan, int* rc, int mixr = 0);
int gnb_fail(gnb_ctx* c, int code, const std::string& msg);
int gnb_cuda_fail(gnb_ctx* c, cudaError_t e, const char* where);

#define GNB_CK(expr)                                                              \
    do {                                                                          \
        cudaError_t _e = (expr);                                                  \
        if (_e != cudaSuccess) return gnb_cuda_fail(c, _e, #expr);                \
    } while (0)
