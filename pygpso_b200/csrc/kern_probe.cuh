// Pipe-peak probes run by bench.py inside its own process, so that the roofline denominators are measured on the same
// GPU, in the same run, as the numbers they divide (the driver's MEASURED_PEAKS.json has HBM copy bandwidth and bf16 cuBLAS
// only; this path runs on the int8 tensor pipe, the FP64 pipe and L2 -> shared-memory bulk copies).
//   probe_i8_kernel     back-to-back tcgen05.mma.cta_group::1.kind::i8, M = 128, N = 256, K = 32 from one thread per SM
//   probe_dmma_kernel   back-to-back mma.sync.m8n8k4.f64 (DMMA.8x8x4), 16 independent accumulator pairs per warp
//   probe_bulk_kernel   cp.async.bulk global -> shared, 8 chunks of 24 KB in flight per SM, working set resident in L2
#pragma once
#include "kern_ozaki.cuh"

namespace gpso {

__global__ void __launch_bounds__(128) probe_i8_kernel(int iters, int* sink) {
    extern __shared__ __align__(1024) uint8_t probe_smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base;
    constexpr int N = 256;
    for (int e = threadIdx.x; e < 2 * (4096 + N * 32) / 4; e += 128) ((uint32_t*)probe_smem)[e] = 0x01010101u * (e & 3);
    if (threadIdx.x == 0) oz_mbar_init(&bar, 1);
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(oz_smem(&tmem_base)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    oz_fence_before();
    __syncthreads();
    oz_fence_after();
    const uint32_t tbase = tmem_base;
    if (threadIdx.x == 0) {
        const uint32_t idesc = oz_idesc(N);
        const uint64_t a0 = oz_desc(oz_smem(probe_smem)), a1 = oz_desc(oz_smem(probe_smem + 4096));
        const uint64_t b0 = oz_desc(oz_smem(probe_smem + 8192)), b1 = oz_desc(oz_smem(probe_smem + 8192 + N * 32));
        oz_mma(tbase, a0, b0, idesc, 0);
        oz_mma(tbase + N, a1, b1, idesc, 0);
        for (int it = 2; it + 2 <= iters; it += 2) {
            oz_mma(tbase, a1, b0, idesc, 1);
            oz_mma(tbase + N, a0, b1, idesc, 1);
        }
        oz_commit(&bar);
    }
    oz_mbar_wait(&bar, 0);
    oz_fence_after();
    uint32_t v;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(v) : "r"(tbase + ((uint32_t)((threadIdx.x >> 5) * 32) << 16)));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    if (v == 0x12345678u) sink[0] = 1;
    oz_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase), "r"(512));
}

__global__ void __launch_bounds__(256) probe_dmma_kernel(double* out, int iters, double a0, double b0) {
    double c[16][2];
#pragma unroll
    for (int i = 0; i < 16; i++) c[i][0] = c[i][1] = 0.0;
    const double a = a0 + threadIdx.x * 1e-9, b = b0;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 16; i++) dmma884(c[i][0], c[i][1], a, b);
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 16; i++) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void __launch_bounds__(128) probe_bulk_kernel(const uint8_t* src, size_t nchunks, int chunk, int per_cta) {
    extern __shared__ __align__(1024) uint8_t probe_smem[];
    __shared__ uint64_t bars[8];
    if (threadIdx.x == 0)
        for (int i = 0; i < 8; i++) oz_mbar_init(&bars[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int i = 0; i < per_cta; i++) {
            const int slot = i & 7;
            if (i >= 8) oz_mbar_wait(&bars[slot], ((i >> 3) - 1) & 1);
            const size_t c = ((size_t)i * gridDim.x + blockIdx.x) % nchunks;
            oz_mbar_expect_tx(&bars[slot], chunk);
            oz_bulk_g2s(probe_smem + (size_t)slot * chunk, src + c * chunk, chunk, &bars[slot]);
        }
        for (int i = per_cta; i < per_cta + 8; i++) {
            const int slot = i & 7;
            if (i >= 8) oz_mbar_wait(&bars[slot], ((i >> 3) - 1) & 1);
        }
    }
    __syncthreads();
}

}  // namespace gpso
