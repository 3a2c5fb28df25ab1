// dev lab: old vs warp-synchronous tournament final round on one random real panel
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/tourn_lab tools/tourn_lab.cu && tools/tourn_lab [n]
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cmath>
#define TW_DEBUG 1
#include "../gaunegf_b200/csrc/gnb_elim.cu"

int main(int argc, char** argv) {
    const int n = argc > 1 ? atoi(argv[1]) : 128, ld = 64, c0 = 0;
    std::vector<cplx> hA((size_t)n * ld);
    srand(1);
    for (auto& v : hA) v = make_double2((rand() / (double)RAND_MAX) - 0.5, 0.0);
    cplx *dA, *dLU; int *dmv, *dinfo, *dc0, *dc1;
    cudaMalloc(&dA, hA.size() * 16); cudaMalloc(&dLU, 1024 * 16); cudaMalloc(&dmv, 4 * GNB_MOVES_STRIDE);
    cudaMalloc(&dinfo, 16); cudaMalloc(&dc0, 4096); cudaMalloc(&dc1, 4096);
    cudaMemcpy(dA, hA.data(), hA.size() * 16, cudaMemcpyHostToDevice);
    {   // host GEPP (izamax, first maximum wins) for reference
        std::vector<double> P((size_t)n * 32);
        for (int r = 0; r < n; r++) for (int c = 0; c < 32; c++) P[r * 32 + c] = hA[(size_t)r * ld + c0 + c].x;
        std::vector<char> alive(n, 1);
        printf("host GEPP winners:");
        for (int j = 0; j < 32 && j < n; j++) {
            int p = -1; double best = -1;
            for (int r = 0; r < n; r++) if (alive[r] && fabs(P[r * 32 + j]) > best) { best = fabs(P[r * 32 + j]); p = r; }
            alive[p] = 0; printf(" %d", p);
            if (j < 6) printf("[piv %.6f; row3: %.6f row2: %.6f]", P[p * 32 + j], P[3 * 32 + j], P[2 * 32 + j]);
            for (int r = 0; r < n; r++) if (alive[r]) { double l = P[r * 32 + j] / P[p * 32 + j]; for (int c = j + 1; c < 32; c++) P[r * 32 + c] -= l * P[p * 32 + c]; }
        }
        printf("\n");
    }
    std::vector<cplx> inv[2]; std::vector<int> mv[2];
    for (int which = 0; which < 2; which++) {
        cudaMemset(dinfo, 0, 16); cudaMemset(dLU, 0, 1024 * 16); cudaMemset(dmv, 0, 4 * GNB_MOVES_STRIDE);
        if (which == 0) {
            if (n > 128) { printf("old final handles <= 128 rows\n"); }
            k_tourn<128, TTR<double>, true, 5><<<dim3(1, 1), 128>>>(dA, 0, ld, c0, 32, c0, std::min(n, 128), nullptr, 0, dc0, 256, 1, dLU, dmv,
                                                                    nullptr, 0, dinfo, 0);
        } else {
            k_tournw<TTR<double>, true, 2><<<dim3(1, 1), 128>>>(dA, 0, ld, c0, c0, n, nullptr, 0, dc0, 256, 1, dLU, dmv, nullptr, 0,
                                                               dinfo, 0);
        }
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("kernel %d failed: %s\n", which, cudaGetErrorString(e)); return 1; }
        inv[which].resize(1024); mv[which].resize(GNB_MOVES_STRIDE);
        cudaMemcpy(inv[which].data(), dLU, 1024 * 16, cudaMemcpyDeviceToHost);
        cudaMemcpy(mv[which].data(), dmv, 4 * GNB_MOVES_STRIDE, cudaMemcpyDeviceToHost);
        int info; cudaMemcpy(&info, dinfo, 4, cudaMemcpyDeviceToHost);
        printf("%s: info=%d nmoves=%d winners:", which ? "new" : "old", info, mv[which][0]);
        for (int t = 0; t < 32; t++) printf(" %d", mv[which][2 + 2 * t]);
        printf("\n");
        // check inv * B = I with B = rows winners
        double worst = 0, growth = 0;
        for (int r = 0; r < 32; r++)
            for (int c = 0; c < 32; c++) {
                double s = 0;
                for (int k = 0; k < 32; k++) s += inv[which][r * 32 + k].x * hA[(size_t)mv[which][2 + 2 * k] * ld + c0 + c].x;
                worst = fmax(worst, fabs(s - (r == c)));
                growth = fmax(growth, fabs(inv[which][r * 32 + c].x));
            }
        printf("   max |inv*B - I| = %.3e, max |inv| = %.3e\n", worst, growth);
    }
    return 0;
}
