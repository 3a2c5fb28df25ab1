// Fused reductions downstream of the batched elimination (all HBM-bound streaming kernels):
//   sum_k w_k G_k            (integrate.py:104-105,116-121)  -> k_weighted_sum, deterministic order
//   -Im diag(G)/pi, its sum  (transport.py:188-189, density.py:54) -> k_dos
//   Re Tr[Z G12^dagger]      (transport.py:156-157)          -> k_trace_dot
// After a JORDAN elimination the stored matrix is (P A)^-1; the column gather through invperm
// (G[:, j] = stored[:, invperm[j]]) is folded into each consumer instead of a separate pass.
#include "gnb_common.cuh"
#include "gnb_kernels.h"

static inline int cdiv_i(long a, long b) { return (int)((a + b - 1) / b); }

__global__ void k_invperm(const int* __restrict__ perm, int* __restrict__ inv, int stride, int N) {
    const int b = blockIdx.y;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += gridDim.x * blockDim.x)
        inv[(long)b * stride + perm[(long)b * stride + i]] = i;
}

__global__ void __launch_bounds__(256) k_weighted_sum(int M, int N, const cplx* __restrict__ A, long strideA, int ld,
                                                      const int* __restrict__ inv, int pstride,
                                                      const cplx* __restrict__ w, cplx* __restrict__ out,
                                                      int accumulate) {
    const long total = (long)N * N;
    for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
        const int i = (int)(idx / N), j = (int)(idx - (long)i * N);
        cplx acc = accumulate ? out[idx] : cmake(0.0, 0.0);
        for (int b = 0; b < M; b++) {
            const int jj = inv ? inv[(long)b * pstride + j] : j;
            acc = cfma(acc, w[b], A[(long)b * strideA + (long)i * ld + jj]);
        }
        out[idx] = acc;
    }
}

__global__ void __launch_bounds__(256) k_unpermute(int N, const cplx* __restrict__ A, long strideA, int ld,
                                                   const int* __restrict__ inv, int pstride,
                                                   cplx* __restrict__ G, long strideG) {
    const int b = blockIdx.y;
    const long total = (long)N * N;
    for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
        const int i = (int)(idx / N), j = (int)(idx - (long)i * N);
        const int jj = inv ? inv[(long)b * pstride + j] : j;
        G[(long)b * strideG + idx] = A[(long)b * strideA + (long)i * ld + jj];
    }
}

__global__ void __launch_bounds__(256) k_dos(int N, const cplx* __restrict__ A, long strideA, int ld,
                                             const int* __restrict__ inv, int pstride,
                                             double* __restrict__ tot, double* __restrict__ per_site) {
    const int b = blockIdx.x, t = threadIdx.x;
    __shared__ double red[256];
    double acc = 0.0;
    for (int i = t; i < N; i += 256) {
        const int jj = inv ? inv[(long)b * pstride + i] : i;
        const double d = -A[(long)b * strideA + (long)i * ld + jj].y / 3.141592653589793;
        if (per_site) per_site[(long)b * N + i] = d;
        acc += d;
    }
    red[t] = acc;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if (t < s) red[t] += red[t + s];
        __syncthreads();
    }
    if (t == 0) tot[b] = red[0];
}

__global__ void __launch_bounds__(256) k_gather_rows(const cplx* __restrict__ X, long strideX, int ldx,
                                                     const int* __restrict__ rows, int nr, int ncols,
                                                     cplx* __restrict__ out, long strideOut,
                                                     const int* __restrict__ map) {
    const int b = blockIdx.y;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < nr * ncols; idx += gridDim.x * blockDim.x) {
        const int r = idx / ncols, c = idx - r * ncols;
        const int row = map ? map[rows[r]] : rows[r];
        out[(long)b * strideOut + idx] = X[(long)b * strideX + (long)row * ldx + c];
    }
}

__global__ void __launch_bounds__(256) k_trace_dot(const cplx* __restrict__ Z, const cplx* __restrict__ X,
                                                   long stride, int n, double* __restrict__ T) {
    const int b = blockIdx.x, t = threadIdx.x;
    __shared__ double red[256];
    double acc = 0.0;
    for (int i = t; i < n; i += 256) {
        const cplx z = Z[(long)b * stride + i], x = X[(long)b * stride + i];
        acc += z.x * x.x + z.y * x.y;          // Re(z * conj(x))
    }
    red[t] = acc;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if (t < s) red[t] += red[t + s];
        __syncthreads();
    }
    if (t == 0) T[b] = red[0];
}

// strided variant for spin-block traces: T[b*tstride + toff] = Re sum_ij Z[i][j] conj(X[i][j])
__global__ void __launch_bounds__(256) k_trace_dot_strided(const cplx* __restrict__ Z, long strideZ, int ldz,
                                                           const cplx* __restrict__ X, long strideX, int ldx,
                                                           int nr, int ncols, double* __restrict__ T, int tstride,
                                                           int toff) {
    const int b = blockIdx.x, t = threadIdx.x;
    __shared__ double red[256];
    double acc = 0.0;
    for (int idx = t; idx < nr * ncols; idx += 256) {
        const int i = idx / ncols, j = idx - i * ncols;
        const cplx z = Z[(long)b * strideZ + (long)i * ldz + j], x = X[(long)b * strideX + (long)i * ldx + j];
        acc += z.x * x.x + z.y * x.y;
    }
    red[t] = acc;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if (t < s) red[t] += red[t + s];
        __syncthreads();
    }
    if (t == 0) T[(long)b * tstride + toff] = red[0];
}

// Gamma = i (Sigma - Sigma^dagger) on compact n x n contact blocks (transport.py:143-146, integrate.py:80)
__global__ void __launch_bounds__(256) k_gamma(const cplx* __restrict__ sig, long stride, int n, cplx* __restrict__ gam) {
    const int b = blockIdx.y;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < n * n; idx += gridDim.x * blockDim.x) {
        const int p = idx / n, q = idx - p * n;
        const cplx s = sig[(long)b * stride + idx], sd = sig[(long)b * stride + (long)q * n + p];
        // i * ((s.x - sd.x) + i (s.y + sd.y)) = -(s.y + sd.y) + i (s.x - sd.x)
        gam[(long)b * stride + idx] = cmake(-(s.y + sd.y), s.x - sd.x);
    }
}

void gnb_launch_invperm(cudaStream_t st, int M, const int* perm, int* invperm, int stride, int N) {
    if (M <= 0) return;
    dim3 grid(cdiv_i(N, 256), M);
    k_invperm<<<grid, 256, 0, st>>>(perm, invperm, stride, N);
}
void gnb_launch_weighted_sum(cudaStream_t st, int M, int N, const cplx* A, long strideA, int ld,
                             const int* invperm, int pstride, const cplx* w, cplx* out, int accumulate) {
    const int grid = min(cdiv_i((long)N * N, 256), 148 * 16);
    k_weighted_sum<<<grid, 256, 0, st>>>(M, N, A, strideA, ld, invperm, pstride, w, out, accumulate);
}
void gnb_launch_unpermute(cudaStream_t st, int M, int N, const cplx* A, long strideA, int ld,
                          const int* invperm, int pstride, cplx* G, long strideG) {
    if (M <= 0) return;
    dim3 grid(min(cdiv_i((long)N * N, 256 * 4), 4096), M);
    k_unpermute<<<grid, 256, 0, st>>>(N, A, strideA, ld, invperm, pstride, G, strideG);
}
void gnb_launch_dos(cudaStream_t st, int M, int N, const cplx* A, long strideA, int ld,
                    const int* invperm, int pstride, double* tot, double* per_site) {
    if (M <= 0) return;
    k_dos<<<M, 256, 0, st>>>(N, A, strideA, ld, invperm, pstride, tot, per_site);
}
void gnb_launch_gather_rows(cudaStream_t st, int M, const cplx* X, long strideX, int ldx, const int* rows,
                            int nr, int ncols, cplx* out, long strideOut, const int* map) {
    if (M <= 0 || nr <= 0 || ncols <= 0) return;
    dim3 grid(min(cdiv_i((long)nr * ncols, 256), 1024), M);
    k_gather_rows<<<grid, 256, 0, st>>>(X, strideX, ldx, rows, nr, ncols, out, strideOut, map);
}
void gnb_launch_trace_dot(cudaStream_t st, int M, const cplx* Z, const cplx* X, long stride, int n, double* T) {
    if (M <= 0) return;
    k_trace_dot<<<M, 256, 0, st>>>(Z, X, stride, n, T);
}
void gnb_launch_gamma_from_sigma(cudaStream_t st, int M, const cplx* sig, long stride, int n, cplx* gam) {
    if (M <= 0 || n <= 0) return;
    dim3 grid(min(cdiv_i((long)n * n, 256), 1024), M);
    k_gamma<<<grid, 256, 0, st>>>(sig, stride, n, gam);
}
void gnb_launch_trace_dot_strided(cudaStream_t st, int M, const cplx* Z, long strideZ, int ldz, const cplx* X,
                                  long strideX, int ldx, int nr, int ncols, double* T, int tstride, int toff) {
    if (M <= 0) return;
    k_trace_dot_strided<<<M, 256, 0, st>>>(Z, strideZ, ldz, X, strideX, ldx, nr, ncols, T, tstride, toff);
}

// out[pi[i]][pi[j]] = in[i][j]  (undo the symmetric contacts-last reordering of a result matrix)
__global__ void __launch_bounds__(256) k_unpermute_sym(int N, const cplx* __restrict__ in, const int* __restrict__ pi,
                                                       cplx* __restrict__ out) {
    const long total = (long)N * N;
    for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
        const int i = (int)(idx / N), j = (int)(idx - (long)i * N);
        out[(long)pi[i] * N + pi[j]] = in[idx];
    }
}
void gnb_launch_unpermute_sym(cudaStream_t st, int N, const cplx* in, const int* pi, cplx* out) {
    k_unpermute_sym<<<min(cdiv_i((long)N * N, 256), 4096), 256, 0, st>>>(N, in, pi, out);
}

// ------------------------------------------------------------------------------------------
// De-orthonormalisation / spin expansion of contact self-energies (surfGBethe.py:529-539):
//   Sigma_c -> Xi Sigma_c Xi = Xi[:, C] blk_c Xi[C, :]  and  kron(I2, Sigma) ('u', 'ro') / kron(Sigma, I2) ('g').
// ------------------------------------------------------------------------------------------
// U[i][p] = Xi[i][inds[p]] (n x nc),  V[p][j] = Xi[inds[p]][j] (nc x n)
__global__ void __launch_bounds__(256) k_xi_gather(int n, const cplx* __restrict__ Xi, const int* __restrict__ inds, int nc,
                                                   cplx* __restrict__ U, cplx* __restrict__ V) {
    const long total = (long)n * nc;
    for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
        const int i = (int)(idx / nc), p = (int)(idx - (long)i * nc);
        U[idx] = Xi[(long)i * n + inds[p]];
        const int q = (int)(idx / n), j = (int)(idx - (long)q * n);
        V[idx] = Xi[(long)inds[q] * n + j];
    }
}
void gnb_launch_xi_gather(cudaStream_t st, int n, const cplx* Xi, const int* inds, int nc, cplx* U, cplx* V) {
    k_xi_gather<<<min(cdiv_i((long)n * nc, 256), 2048), 256, 0, st>>>(n, Xi, inds, nc, U, V);
}

// dense[b][inds[p]][inds[q]] += blk[b][p][q]
__global__ void __launch_bounds__(256) k_scatter_add(cplx* __restrict__ dense, long strideD, int n, const int* __restrict__ inds,
                                                     int nc, const cplx* __restrict__ blk, long strideBlk) {
    const int b = blockIdx.y;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < nc * nc; idx += gridDim.x * blockDim.x) {
        const int p = idx / nc, q = idx - p * nc;
        cplx* d = dense + (long)b * strideD + (long)inds[p] * n + inds[q];
        *d = cadd(*d, blk[(long)b * strideBlk + idx]);
    }
}
void gnb_launch_scatter_add(cudaStream_t st, int M, cplx* dense, long strideD, int n, const int* inds, int nc,
                            const cplx* blk, long strideBlk) {
    if (M <= 0 || nc <= 0) return;
    dim3 grid(min(cdiv_i((long)nc * nc, 256), 256), M);
    k_scatter_add<<<grid, 256, 0, st>>>(dense, strideD, n, inds, nc, blk, strideBlk);
}

// out (2n x 2n) = kron(I2, in) (mode 1) or kron(in, I2) (mode 2); in is n x n
__global__ void __launch_bounds__(256) k_kron_expand(int n, int mode, const cplx* __restrict__ in, cplx* __restrict__ out) {
    const int b = blockIdx.y, n2 = 2 * n;
    const long total = (long)n2 * n2;
    for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
        const int r = (int)(idx / n2), c = (int)(idx - (long)r * n2);
        int i, j, sr, sc;
        if (mode == 1) { sr = r / n; i = r - sr * n; sc = c / n; j = c - sc * n; }
        else { i = r >> 1; sr = r & 1; j = c >> 1; sc = c & 1; }
        out[(long)b * total + idx] = (sr == sc) ? in[(long)b * n * n + (long)i * n + j] : cmake(0.0, 0.0);
    }
}
void gnb_launch_kron_expand(cudaStream_t st, int M, int n, int mode, const cplx* in, cplx* out) {
    if (M <= 0) return;
    dim3 grid(min(cdiv_i(4L * n * n, 256), 1024), M);
    k_kron_expand<<<grid, 256, 0, st>>>(n, mode, in, out);
}
