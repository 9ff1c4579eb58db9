// Phase timing + correctness of the 128x128 diagonal-block kernel (the serial spine of the blocked Cholesky).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -I../../pygpso_b200/csrc diag_probe.cu -o diag_probe
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>
__device__ long long g_stamps[32];
#define DIAG_STAMP(i) do { if (threadIdx.x == 0) g_stamps[i] = clock64(); } while (0)
#define DIAG_STAMP_W(w, i) do { if (threadIdx.x == 32 * (w)) g_stamps[i] = clock64(); } while (0)
#include "kern_dense.cuh"
using namespace gpso;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__global__ void rsqrt_test(const double* x, double* y, double* z, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) rsqrt_rcp_fast(x[i], y[i], z[i]);
}

int main() {
    const int n = TB, Np = 256;  // block p=1 of a 256x256 matrix
    std::vector<double> A((size_t)Np * Np, 0.0), B((size_t)n * n);
    srand(1);
    for (auto& v : B) v = rand() / (double)RAND_MAX - 0.5;
    std::vector<double> Ad((size_t)n * n);
    for (int i = 0; i < n; i++)
        for (int j = 0; j < n; j++) {
            double s = (i == j) ? 0.05 : 0.0;
            for (int k = 0; k < n; k++) s += B[i * n + k] * B[j * n + k] / n;
            Ad[i * n + j] = s;
        }
    const int p = 1;
    for (int i = 0; i < n; i++)
        for (int j = 0; j < n; j++) A[(size_t)(p * n + i) * Np + p * n + j] = Ad[i * n + j];
    double *dK, *dK0, *dLi, *dLiT, *dlog;
    int* dinfo;
    size_t bytes = (size_t)Np * Np * sizeof(double);
    CK(cudaMalloc(&dK, bytes)); CK(cudaMalloc(&dK0, bytes)); CK(cudaMalloc(&dLi, bytes)); CK(cudaMalloc(&dLiT, bytes)); CK(cudaMemset(dLi, 0, bytes)); CK(cudaMemset(dLiT, 0, bytes));
    CK(cudaMalloc(&dlog, 64)); CK(cudaMalloc(&dinfo, 4));
    CK(cudaMemcpy(dK0, A.data(), bytes, cudaMemcpyHostToDevice));
    CK(cudaMemset(dinfo, 0, 4));
    CK(cudaFuncSetAttribute(diag_factor_inverse_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DIAG_SMEM_BYTES));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9, tot = 0;
    const int reps = 50;
    for (int it = 0; it < reps; it++) {
        CK(cudaMemcpy(dK, dK0, bytes, cudaMemcpyDeviceToDevice));
        CK(cudaDeviceSynchronize());
        cudaEventRecord(e0);
        diag_factor_inverse_kernel<<<1, DIAG_THREADS, DIAG_SMEM_BYTES>>>(dK, dLi, Np, p, Np, dlog, dinfo);
        cudaEventRecord(e1);
        CK(cudaDeviceSynchronize());
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (it >= 5) { tot += ms; best = fminf(best, ms); }
    }
    diag_transpose_kernel<<<Np / TB, 256>>>(dLi, dLiT, Np);
    CK(cudaDeviceSynchronize());
    printf("diag_factor_inverse_kernel: avg %.2f us  best %.2f us\n", tot / (reps - 5) * 1e3, best * 1e3);
    long long st[32];
    CK(cudaMemcpyFromSymbol(st, g_stamps, sizeof st));
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    printf("phase stamps (cycles since stamp 1 = block loaded; SM clock attr %d kHz):\n", clk);
    for (int i = 1; i <= 18; i++) printf("  %2d: %8lld  (+%lld)\n", i, st[i] - st[1], i > 1 ? st[i] - st[i - 1] : 0);
    printf("chol_inv_32 (last call): load %lld  columns %lld  store %lld cycles\n", st[21] - st[20], st[22] - st[21], st[23] - st[22]);
    printf("  load: %lld cycles; warp4 tail: issue done %lld, wait done %lld (since stamp 1)\n", st[1] - st[0], st[30] - st[1], st[31] - st[1]);
    // correctness
    std::vector<double> L((size_t)Np * Np), Li((size_t)Np * Np), LiT((size_t)Np * Np);
    CK(cudaMemcpy(L.data(), dK, bytes, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(Li.data(), dLi, bytes, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(LiT.data(), dLiT, bytes, cudaMemcpyDeviceToHost));
    double e_llt = 0, e_inv = 0, e_t = 0, ld = 0, dl;
    auto Lat = [&](int i, int j) { return j <= i ? L[(size_t)(p * n + i) * Np + p * n + j] : 0.0; };
    auto Iat = [&](int i, int j) { return Li[(size_t)(p * n + i) * Np + p * n + j]; };
    for (int i = 0; i < n; i++)
        for (int j = 0; j < n; j++) {
            double s = 0, t = 0;
            for (int k = 0; k < n; k++) { s += Lat(i, k) * Lat(j, k); t += Iat(i, k) * Lat(k, j); }
            e_llt = fmax(e_llt, fabs(s - Ad[i * n + j]));
            e_inv = fmax(e_inv, fabs(t - (i == j)));
            e_t = fmax(e_t, fabs(Iat(i, j) - LiT[(size_t)(p * n + j) * Np + p * n + i]));
        }
    for (int i = 0; i < n; i++) ld += log(Lat(i, i));
    CK(cudaMemcpy(&dl, dlog + p, 8, cudaMemcpyDeviceToHost));
    int info; CK(cudaMemcpy(&info, dinfo, 4, cudaMemcpyDeviceToHost));
    printf("max|LL^T-A| %.3e  max|Linv L - I| %.3e  max|LinvT-Linv^T| %.3e  logdet %.15g vs %.15g  info %d\n", e_llt, e_inv, e_t, dl, ld, info);
    {
        const int n = 1 << 20;
        std::vector<double> x(n), y(n), z(n);
        for (int i = 0; i < n; i++) x[i] = ldexp(1.0 + rand() / (double)RAND_MAX * 3.0, (rand() % 120) - 60);
        double *dx, *dy, *dz;
        CK(cudaMalloc(&dx, n * 8)); CK(cudaMalloc(&dy, n * 8)); CK(cudaMalloc(&dz, n * 8));
        CK(cudaMemcpy(dx, x.data(), n * 8, cudaMemcpyHostToDevice));
        rsqrt_test<<<n / 256, 256>>>(dx, dy, dz, n);
        CK(cudaMemcpy(y.data(), dy, n * 8, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(z.data(), dz, n * 8, cudaMemcpyDeviceToHost));
        double worst = 0, worst2 = 0;
        for (int i = 0; i < n; i++) {
            long double ref = 1.0L / sqrtl((long double)x[i]);
            worst = fmax(worst, (double)fabsl((y[i] - ref) / ref));
            long double ref2 = 1.0L / (long double)x[i];
            worst2 = fmax(worst2, (double)fabsl((z[i] - ref2) / ref2));
        }
        printf("rsqrt_rcp_fast max rel err: rsqrt %.3e  rcp %.3e (eps = 1.1e-16)\n", worst, worst2);
    }
    return 0;
}
