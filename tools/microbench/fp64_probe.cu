// FP64 hardware probe for B200 (sm_100a): decides the design of the predict / Cholesky kernels.
//   1. DMMA.8x8x4 issue-bound peak (mma.sync.m8n8k4.f64), by warps/SM
//   2. DFMA issue-bound peak
//   3. DMMA + DFMA mixed in the same warp / different warps: do the pipes add or share?
//   4. exp()/sqrt() fp64 throughput (cost of generating one cross-covariance element)
//   5. cuBLAS DGEMM / DSYRK / DTRSM and cuSOLVER DPOTRF / DPOTRI (the library bars)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo fp64_probe.cu -lcublas -lcusolver -o fp64_probe
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cmath>
#include <cuda_runtime.h>
#include <cublas_v2.h>
#include <cusolverDn.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int NACC>
__global__ void k_dmma(double* out, int iters, double a0, double b0) {
    double c[NACC][2];
#pragma unroll
    for (int i = 0; i < NACC; i++) { c[i][0] = 0; c[i][1] = 0; }
    double a = a0 + threadIdx.x * 1e-9, b = b0;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < NACC; i++) dmma884(c[i][0], c[i][1], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; i++) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NACC>
__global__ void k_dfma(double* out, int iters, double a0, double b0) {
    double c[NACC];
#pragma unroll
    for (int i = 0; i < NACC; i++) c[i] = i;
    double a = a0 + threadIdx.x * 1e-9, b = b0;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < NACC; i++) c[i] = fma(c[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; i++) s += c[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// same warp issues NM dmma + NF dfma per iteration
template <int NM, int NF>
__global__ void k_mixed(double* out, int iters, double a0, double b0) {
    double c[NM][2];
    double f[NF > 0 ? NF : 1];
#pragma unroll
    for (int i = 0; i < NM; i++) { c[i][0] = 0; c[i][1] = 0; }
#pragma unroll
    for (int i = 0; i < NF; i++) f[i] = i;
    double a = a0 + threadIdx.x * 1e-9, b = b0;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < NM; i++) dmma884(c[i][0], c[i][1], a, b);
#pragma unroll
        for (int i = 0; i < NF; i++) f[i] = fma(f[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NM; i++) s += c[i][0] + c[i][1];
#pragma unroll
    for (int i = 0; i < NF; i++) s += f[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// warp-specialised: even warps DMMA, odd warps DFMA
__global__ void k_split(double* out, int iters, double a0, double b0, int fma_mult) {
    int w = threadIdx.x >> 5;
    double a = a0 + threadIdx.x * 1e-9, b = b0;
    double s = 0;
    if (w & 1) {
        double f[16];
#pragma unroll
        for (int i = 0; i < 16; i++) f[i] = i;
        for (int it = 0; it < iters * fma_mult; it++) {
#pragma unroll
            for (int i = 0; i < 16; i++) f[i] = fma(f[i], a, b);
        }
#pragma unroll
        for (int i = 0; i < 16; i++) s += f[i];
    } else {
        double c[16][2];
#pragma unroll
        for (int i = 0; i < 16; i++) { c[i][0] = 0; c[i][1] = 0; }
        for (int it = 0; it < iters; it++) {
#pragma unroll
            for (int i = 0; i < 16; i++) dmma884(c[i][0], c[i][1], a, b);
        }
#pragma unroll
        for (int i = 0; i < 16; i++) s += c[i][0] + c[i][1];
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_exp(double* out, int iters, double x0) {
    double x = x0 + threadIdx.x * 1e-3;
    double s = 0;
    for (int it = 0; it < iters; it++) {
        double r0 = exp(-x), r1 = exp(-x - 0.1), r2 = exp(-x - 0.2), r3 = exp(-x - 0.3);
        s += r0 + r1 + r2 + r3;
        x += 1e-7;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_sqrt(double* out, int iters, double x0) {
    double x = x0 + threadIdx.x * 1e-3;
    double s = 0;
    for (int it = 0; it < iters; it++) {
        double r0 = sqrt(x), r1 = sqrt(x + 0.1), r2 = sqrt(x + 0.2), r3 = sqrt(x + 0.3);
        s += r0 + r1 + r2 + r3;
        x += 1e-7;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
float time_ms(F f, int reps = 3) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    f();
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < reps; r++) {
        CK(cudaEventRecord(e0));
        f();
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    return best;
}

int main() {
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    int nsm = prop.multiProcessorCount;
    printf("device %s  SMs %d  clock %d kHz  smem/blk optin %zu  L2 %d MB\n", prop.name, nsm, prop.clockRate,
           prop.sharedMemPerBlockOptin, prop.l2CacheSize >> 20);
    double* out; CK(cudaMalloc(&out, sizeof(double) * nsm * 8 * 1024));
    const int iters = 20000;

    printf("\n== DMMA.8x8x4 (256 FMA / warp-instr), 16 independent accumulators/warp ==\n");
    for (int warps : {1, 2, 4, 8, 16, 32}) {
        for (int bps : {1, 2}) {
            if (warps * bps > 64) continue;
            float ms = time_ms([&] { k_dmma<16><<<nsm * bps, warps * 32>>>(out, iters, 1.0, 1e-3); });
            double flop = 2.0 * 256 * 16 * (double)iters * warps * bps * nsm;
            printf("  warps/blk %2d blk/SM %d : %8.3f ms  %7.2f TFLOP/s\n", warps, bps, ms, flop / ms * 1e-9);
        }
    }
    printf("== DMMA with 4 / 8 / 32 accumulators per warp, 8 warps/SM ==\n");
    {
        float ms = time_ms([&] { k_dmma<4><<<nsm, 256>>>(out, iters, 1.0, 1e-3); });
        printf("  nacc 4 : %7.2f TFLOP/s\n", 2.0 * 256 * 4 * (double)iters * 8 * nsm / ms * 1e-9);
        ms = time_ms([&] { k_dmma<8><<<nsm, 256>>>(out, iters, 1.0, 1e-3); });
        printf("  nacc 8 : %7.2f TFLOP/s\n", 2.0 * 256 * 8 * (double)iters * 8 * nsm / ms * 1e-9);
        ms = time_ms([&] { k_dmma<32><<<nsm, 256>>>(out, iters, 1.0, 1e-3); });
        printf("  nacc 32: %7.2f TFLOP/s\n", 2.0 * 256 * 32 * (double)iters * 8 * nsm / ms * 1e-9);
    }

    printf("\n== DFMA (32 FMA / warp-instr), 16 independent chains/thread ==\n");
    for (int warps : {4, 8, 16, 32}) {
        float ms = time_ms([&] { k_dfma<16><<<nsm * 2, warps * 32>>>(out, iters, 1.0000001, 1e-3); });
        double flop = 2.0 * 32 * 16 * (double)iters * warps * 2 * nsm;
        printf("  warps/blk %2d blk/SM 2 : %8.3f ms  %7.2f TFLOP/s\n", warps, ms, flop / ms * 1e-9);
    }

    printf("\n== mixed in one warp: (NM dmma + NF dfma)/iter, 16 warps/SM ==\n");
    {
        float ms;
        ms = time_ms([&] { k_mixed<16, 0><<<nsm, 512>>>(out, iters, 1.0000001, 1e-3); });
        printf("  16 dmma +  0 dfma: %8.3f ms\n", ms);
        ms = time_ms([&] { k_mixed<16, 16><<<nsm, 512>>>(out, iters, 1.0000001, 1e-3); });
        printf("  16 dmma + 16 dfma: %8.3f ms  (dfma adds %.1f%% flops)\n", ms, 100.0 * 16 * 32 / (16 * 256));
        ms = time_ms([&] { k_mixed<16, 64><<<nsm, 512>>>(out, iters, 1.0000001, 1e-3); });
        printf("  16 dmma + 64 dfma: %8.3f ms  (dfma adds %.1f%% flops)\n", ms, 100.0 * 64 * 32 / (16 * 256));
        ms = time_ms([&] { k_mixed<16, 128><<<nsm, 512>>>(out, iters, 1.0000001, 1e-3); });
        printf("  16 dmma +128 dfma: %8.3f ms  (dfma adds %.1f%% flops)\n", ms, 100.0 * 128 * 32 / (16 * 256));
        ms = time_ms([&] { k_mixed<1, 128><<<nsm, 512>>>(out, iters, 1.0000001, 1e-3); });
        printf("   1 dmma +128 dfma: %8.3f ms\n", ms);
    }
    printf("== warp-specialised: even warps 16 dmma/iter, odd warps 16*mult dfma/iter, 16 warps/SM ==\n");
    for (int mult : {0, 1, 4, 8}) {
        float ms = time_ms([&] { k_split<<<nsm, 512>>>(out, iters, 1.0000001, 1e-3, mult); });
        printf("  fma_mult %d : %8.3f ms\n", mult, ms);
    }

    printf("\n== fp64 exp / sqrt throughput ==\n");
    {
        int it2 = 20000;
        float ms = time_ms([&] { k_exp<<<nsm * 4, 256>>>(out, it2, 0.5); });
        double n = 4.0 * it2 * 256 * 4 * nsm;
        printf("  exp : %8.3f ms  %.3e /s   (= %.1f DFMA-equivalents at 37 TF/s-DFMA-rate 18.5e12 FMA/s)\n", ms, n / ms * 1e3,
               18.5e12 / (n / ms * 1e3));
        ms = time_ms([&] { k_sqrt<<<nsm * 4, 256>>>(out, it2, 0.5); });
        printf("  sqrt: %8.3f ms  %.3e /s   (= %.1f DFMA-equivalents)\n", ms, n / ms * 1e3, 18.5e12 / (n / ms * 1e3));
    }

    printf("\n== cuBLAS / cuSOLVER fp64 bars ==\n");
    cublasHandle_t hb; cublasCreate(&hb);
    cusolverDnHandle_t hs; cusolverDnCreate(&hs);
    for (int n : {2048, 4096, 8192}) {
        double *A, *B, *C;
        size_t bytes = sizeof(double) * (size_t)n * n;
        CK(cudaMalloc(&A, bytes)); CK(cudaMalloc(&B, bytes)); CK(cudaMalloc(&C, bytes));
        std::vector<double> h((size_t)n * n);
        for (size_t i = 0; i < h.size(); i++) h[i] = (double)rand() / RAND_MAX - 0.5;
        CK(cudaMemcpy(A, h.data(), bytes, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(B, h.data(), bytes, cudaMemcpyHostToDevice));
        double one = 1.0, zero = 0.0;
        float ms = time_ms([&] { cublasDgemm(hb, CUBLAS_OP_N, CUBLAS_OP_N, n, n, n, &one, A, n, B, n, &zero, C, n); });
        printf("  N=%5d dgemm NN : %9.3f ms  %7.2f TFLOP/s\n", n, ms, 2.0 * n * n * n / ms * 1e-9);
        ms = time_ms([&] { cublasDgemm(hb, CUBLAS_OP_T, CUBLAS_OP_N, n, n, n, &one, A, n, B, n, &zero, C, n); });
        printf("  N=%5d dgemm TN : %9.3f ms  %7.2f TFLOP/s\n", n, ms, 2.0 * n * n * n / ms * 1e-9);
        ms = time_ms([&] { cublasDsyrk(hb, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_N, n, n, &one, A, n, &zero, C, n); });
        printf("  N=%5d dsyrk    : %9.3f ms  %7.2f TFLOP/s\n", n, ms, 1.0 * n * n * n / ms * 1e-9);
        ms = time_ms([&] { cublasDtrmm(hb, CUBLAS_SIDE_LEFT, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_N, CUBLAS_DIAG_NON_UNIT, n, n, &one, A, n, B, n, C, n); });
        printf("  N=%5d dtrmm    : %9.3f ms  %7.2f TFLOP/s (n^3 flops)\n", n, ms, 1.0 * n * n * n / ms * 1e-9);
        // SPD matrix for potrf: C = A*A^T + n*I
        cublasDsyrk(hb, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_N, n, n, &one, A, n, &zero, C, n);
        CK(cudaDeviceSynchronize());
        {
            std::vector<double> d((size_t)n, (double)n);
            // add n to diagonal
            std::vector<double> hc((size_t)n * n);
            CK(cudaMemcpy(hc.data(), C, bytes, cudaMemcpyDeviceToHost));
            for (int i = 0; i < n; i++) hc[(size_t)i * n + i] += n;
            CK(cudaMemcpy(C, hc.data(), bytes, cudaMemcpyHostToDevice));
        }
        int lwork = 0, lwork2 = 0; int* info; CK(cudaMalloc(&info, 4));
        cusolverDnDpotrf_bufferSize(hs, CUBLAS_FILL_MODE_LOWER, n, B, n, &lwork);
        cusolverDnDpotri_bufferSize(hs, CUBLAS_FILL_MODE_LOWER, n, B, n, &lwork2);
        if (lwork2 > lwork) lwork = lwork2;
        double* work; CK(cudaMalloc(&work, sizeof(double) * (size_t)lwork));
        float best = 1e30f, besti = 1e30f, bestt = 1e30f;
        for (int r = 0; r < 3; r++) {
            CK(cudaMemcpy(B, C, bytes, cudaMemcpyDeviceToDevice));
            cudaEvent_t e0, e1, e2; cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventCreate(&e2);
            cudaEventRecord(e0);
            cusolverDnDpotrf(hs, CUBLAS_FILL_MODE_LOWER, n, B, n, work, lwork, info);
            cudaEventRecord(e1);
            cusolverDnDpotri(hs, CUBLAS_FILL_MODE_LOWER, n, B, n, work, lwork, info);
            cudaEventRecord(e2);
            cudaEventSynchronize(e2);
            float m1, m2; cudaEventElapsedTime(&m1, e0, e1); cudaEventElapsedTime(&m2, e1, e2);
            if (m1 < best) best = m1;
            if (m2 < besti) besti = m2;
            // trsm with n rhs for comparison
            CK(cudaMemcpy(B, C, bytes, cudaMemcpyDeviceToDevice));
            cusolverDnDpotrf(hs, CUBLAS_FILL_MODE_LOWER, n, B, n, work, lwork, info);
            cudaEventRecord(e0);
            cublasDtrsm(hb, CUBLAS_SIDE_LEFT, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_N, CUBLAS_DIAG_NON_UNIT, n, n, &one, B, n, A, n);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            cudaEventElapsedTime(&m1, e0, e1);
            if (m1 < bestt) bestt = m1;
            CK(cudaMemcpy(A, h.data(), bytes, cudaMemcpyHostToDevice));
        }
        int hinfo; CK(cudaMemcpy(&hinfo, info, 4, cudaMemcpyDeviceToHost));
        printf("  N=%5d dpotrf   : %9.3f ms  %7.2f TFLOP/s (n^3/3)  info=%d\n", n, best, n / 3.0 * n * n / best * 1e-9, hinfo);
        printf("  N=%5d dpotri   : %9.3f ms  %7.2f TFLOP/s (2n^3/3)\n", n, besti, 2.0 * n / 3.0 * n * n / besti * 1e-9);
        printf("  N=%5d dtrsm nxn: %9.3f ms  %7.2f TFLOP/s (n^3)\n", n, bestt, 1.0 * n * n * n / bestt * 1e-9);
        cudaFree(A); cudaFree(B); cudaFree(C); cudaFree(work); cudaFree(info);
    }
    // sustained dgemm: 3 s back to back to see the power-capped rate
    {
        int n = 8192; size_t bytes = sizeof(double) * (size_t)n * n;
        double *A, *B, *C; CK(cudaMalloc(&A, bytes)); CK(cudaMalloc(&B, bytes)); CK(cudaMalloc(&C, bytes));
        CK(cudaMemset(A, 0, bytes)); CK(cudaMemset(B, 0, bytes));
        std::vector<double> h((size_t)n * n);
        for (size_t i = 0; i < h.size(); i++) h[i] = (double)rand() / RAND_MAX - 0.5;
        CK(cudaMemcpy(A, h.data(), bytes, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(B, h.data(), bytes, cudaMemcpyHostToDevice));
        double one = 1.0, zero = 0.0;
        int reps = 100;
        float ms = time_ms([&] { for (int r = 0; r < reps; r++) cublasDgemm(hb, CUBLAS_OP_N, CUBLAS_OP_N, n, n, n, &one, A, n, B, n, &zero, C, n); }, 1);
        printf("  sustained dgemm 8192^3 x%d: %9.3f ms total  %7.2f TFLOP/s\n", reps, ms, 2.0 * n * n * n * reps / ms * 1e-9);
    }
    return 0;
}
