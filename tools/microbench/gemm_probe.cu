// Mainloop rate of the FP64 DMMA tile core (gemm_core.cuh) in isolation: 148 CTAs, each one 128x128 tile with a long K
// (operands L2-resident), no epilogue traffic.  Build with -DGPSO_GK=.. -DGPSO_GSTAGES=.. to compare slab shapes.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "gemm_core.cuh"
using namespace gpso;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__global__ void __launch_bounds__(GTHREADS, 1) probe(const double* A, const double* B, int ld, int K, double* out) {
    extern __shared__ double smem[];
    TileOperands w;
    w.A = A + (size_t)(blockIdx.x % 8) * GM * ld;
    w.B = B + (size_t)(blockIdx.x % 8) * GN * ld;
    w.lda = w.ldb = ld;
    w.kbeg = 0;
    w.kend = K;
    w.tri_off = TRI_DENSE;
    TileAcc acc;
    gemm_tile_mainloop(w, acc, smem);
    double s = 0;
    for (int i = 0; i < 8; i++)
        for (int j = 0; j < 4; j++) s += acc.v[i][j][0] + acc.v[i][j][1];
    out[blockIdx.x * GTHREADS + threadIdx.x] = s;
}

int main() {
    const int K = 8192, ld = K, rows = 8 * 128;
    double *A, *B, *out;
    CK(cudaMalloc(&A, (size_t)rows * ld * 8)); CK(cudaMalloc(&B, (size_t)rows * ld * 8)); CK(cudaMalloc(&out, 148 * GTHREADS * 8));
    CK(cudaMemset(A, 0, (size_t)rows * ld * 8)); CK(cudaMemset(B, 0, (size_t)rows * ld * 8));
    CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM_BYTES));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int kk : {128, 512, 2048, 8192}) {
        probe<<<148, GTHREADS, GEMM_SMEM_BYTES>>>(A, B, ld, kk, out);
        CK(cudaDeviceSynchronize());
        cudaEventRecord(e0);
        const int reps = kk >= 2048 ? 5 : 50;
        for (int r = 0; r < reps; r++) probe<<<148, GTHREADS, GEMM_SMEM_BYTES>>>(A, B, ld, kk, out);
        cudaEventRecord(e1);
        CK(cudaDeviceSynchronize());
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        ms /= reps;
        printf("GK=%d stages=%d smem=%d  K=%5d: %.3f ms/launch  %.2f TFLOP/s (%.1f%% of 37.03)\n", GK, GSTAGES, GEMM_SMEM_BYTES, kk, ms,
               148.0 * 2.0 * 128 * 128 * kk / ms * 1e-9, 148.0 * 2.0 * 128 * 128 * kk / ms * 1e-9 / 37.03 * 100);
    }
    return 0;
}
