// tcgen05.mma kind::i8 probe for B200 (sm_100a): (1) one-instruction correctness of the no-swizzle K-major shared
// memory layout + TMEM read-back, (2) issue-rate of M=128 MMAs for several N, 1-CTA and 2-CTA groups.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o i8_probe i8_probe.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#define CK(x)                                                                                   \
    do {                                                                                        \
        cudaError_t e = (x);                                                                    \
        if (e != cudaSuccess) {                                                                 \
            printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__);      \
            exit(1);                                                                            \
        }                                                                                       \
    } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// K-major, no swizzle: core matrix = 8 rows x 16 bytes, contiguous (128 B).  lbo = byte stride between the two 16-byte
// K chunks of one K=32 instruction, sbo = byte stride between 8-row groups.
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3fff);
    d |= (uint64_t)((lbo >> 4) & 0x3fff) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3fff) << 32;
    d |= (uint64_t)1 << 46;  // version = 1 (Blackwell)
    return d;
}

__host__ __device__ constexpr uint32_t make_idesc_i8(int M, int N) {
    return (2u << 4) /* D = S32 */ | (1u << 7) /* A = INT8 */ | (1u << 10) /* B = INT8 */ | ((uint32_t)(N >> 3) << 17) |
           ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ void mma_i8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void mma_i8_2cta(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred P1;\n\tWAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
        "@P1 bra DONE;\n\tbra WAIT_LOOP;\n\tDONE:\n\t}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void umma_commit_2cta(uint64_t* bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
                 "h"(mask)
                 : "memory");
}

// ---- (1) correctness: D[128 x N] = A[128 x K] * B[N x K]^T with K = 32*ksteps, int8 in, int32 out ------------------
// global A, B are plain row-major; the kernel re-tiles them into the canonical layout in shared memory.
template <int N>
__global__ void __launch_bounds__(128) check_kernel(const int8_t* A, const int8_t* B, int ksteps, int32_t* D) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base;
    uint8_t* sA = smem;                      // per k-step: 128 rows * 32 B = 4096 B
    uint8_t* sB = smem + ksteps * 4096;      // per k-step: N rows * 32 B
    const int K = 32 * ksteps;
    // canonical: element (row r, k byte kb) of k-step s -> s*tile + (r/8)*256 + ((kb%32)/16)*128 + (r%8)*16 + kb%16
    for (int e = threadIdx.x; e < 128 * K; e += 128) {
        int r = e / K, kb = e % K, s = kb / 32, kk = kb % 32;
        sA[s * 4096 + (r / 8) * 256 + (kk / 16) * 128 + (r % 8) * 16 + kk % 16] = (uint8_t)A[e];
    }
    for (int e = threadIdx.x; e < N * K; e += 128) {
        int r = e / K, kb = e % K, s = kb / 32, kk = kb % 32;
        sB[s * (N * 32) + (r / 8) * 256 + (kk / 16) * 128 + (r % 8) * 16 + kk % 16] = (uint8_t)B[e];
    }
    if (threadIdx.x == 0) mbar_init(&bar, 1);
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    // make the generic-proxy smem writes visible to the async proxy (UMMA reads smem through it)
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tbase = tmem_base;
    if (threadIdx.x == 0) {
        const uint32_t idesc = make_idesc_i8(128, N);
        for (int s = 0; s < ksteps; s++) {
            uint64_t ad = make_desc(smem_u32(sA + s * 4096), 128, 256);
            uint64_t bd = make_desc(smem_u32(sB + s * (N * 32)), 128, 256);
            mma_i8(tbase, ad, bd, idesc, s > 0);
        }
        umma_commit(&bar);
    }
    mbar_wait(&bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int warp = threadIdx.x >> 5;
    for (int c0 = 0; c0 < N; c0 += 8) {
        uint32_t v[8];
        uint32_t taddr = tbase + ((uint32_t)(warp * 32) << 16) + c0;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                     : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int j = 0; j < 8; j++) D[(size_t)threadIdx.x * N + c0 + j] = (int32_t)v[j];
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase), "r"(512));
}

// ---- (2) issue rate, 1 CTA per SM ----------------------------------------------------------------------------------
template <int N>
__global__ void __launch_bounds__(128) rate_kernel(int iters, int nslots, int* sink) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base;
    for (int e = threadIdx.x; e < nslots * (4096 + N * 32) / 4; e += 128) ((uint32_t*)smem)[e] = 0x01010101u * (e & 3);
    if (threadIdx.x == 0) mbar_init(&bar, 1);
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tbase = tmem_base;
    if (threadIdx.x == 0) {
        const uint32_t idesc = make_idesc_i8(128, N);
        uint8_t* sA = smem;
        uint8_t* sB = smem + nslots * 4096;
        uint64_t ad[4], bd[4];
        for (int q = 0; q < 4; q++) {
            ad[q] = make_desc(smem_u32(sA + (q % nslots) * 4096), 128, 256);
            bd[q] = make_desc(smem_u32(sB + (q % nslots) * (N * 32)), 128, 256);
        }
        const uint32_t col1 = (2 * N <= 512) ? (uint32_t)N : 0u;
        mma_i8(tbase, ad[0], bd[0], idesc, 0);
        mma_i8(tbase + col1, ad[1], bd[1], idesc, 0);
        for (int it = 2; it + 4 <= iters; it += 4) {
            mma_i8(tbase, ad[2], bd[2], idesc, 1);
            mma_i8(tbase + col1, ad[3], bd[3], idesc, 1);
            mma_i8(tbase, ad[0], bd[0], idesc, 1);
            mma_i8(tbase + col1, ad[1], bd[1], idesc, 1);
        }
        umma_commit(&bar);
    }
    mbar_wait(&bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t v;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(v) : "r"(tbase + ((uint32_t)((threadIdx.x >> 5) * 32) << 16)));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    if (v == 0x12345678u) sink[0] = 1;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase), "r"(512));
}

// ---- (3) issue rate, CTA pairs (cta_group::2): D = 256 x N, each CTA holds 128 rows of A and N/2 rows of B ----------
template <int N>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128) rate2_kernel(int iters, int nslots, int* sink) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base;
    uint32_t rank;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    for (int e = threadIdx.x; e < nslots * (4096 + N * 16) / 4; e += 128) ((uint32_t*)smem)[e] = 0x01010101u * (e & 3);
    if (threadIdx.x == 0) mbar_init(&bar, 1);
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tbase = tmem_base;
    if (rank == 0 && threadIdx.x == 0) {
        const uint32_t idesc = make_idesc_i8(256, N);
        uint8_t* sA = smem;
        uint8_t* sB = smem + nslots * 4096;
        uint64_t ad[4], bd[4];
        for (int q = 0; q < 4; q++) {
            ad[q] = make_desc(smem_u32(sA + (q % nslots) * 4096), 128, 256);
            bd[q] = make_desc(smem_u32(sB + (q % nslots) * (N * 16)), 128, 256);
        }
        const uint32_t col1 = (2 * N <= 512) ? (uint32_t)N : 0u;
        mma_i8_2cta(tbase, ad[0], bd[0], idesc, 0);
        mma_i8_2cta(tbase + col1, ad[1], bd[1], idesc, 0);
        for (int it = 2; it + 4 <= iters; it += 4) {
            mma_i8_2cta(tbase, ad[2], bd[2], idesc, 1);
            mma_i8_2cta(tbase + col1, ad[3], bd[3], idesc, 1);
            mma_i8_2cta(tbase, ad[0], bd[0], idesc, 1);
            mma_i8_2cta(tbase + col1, ad[1], bd[1], idesc, 1);
        }
        umma_commit_2cta(&bar, 3);
    }
    mbar_wait(&bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t v;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(v) : "r"(tbase + ((uint32_t)((threadIdx.x >> 5) * 32) << 16)));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    if (v == 0x12345678u) sink[0] = 1;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tbase), "r"(512));
}


// ---- (4) L2 -> shared memory bandwidth through cp.async.bulk (UBLKCP), one CTA per SM, 8 x CHUNK in flight per SM ------
__global__ void __launch_bounds__(128) bulk_bw_kernel(const uint8_t* src, size_t nchunks, int chunk, int per_cta) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bars[8];
    if (threadIdx.x == 0)
        for (int i = 0; i < 8; i++) mbar_init(&bars[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int i = 0; i < per_cta; i++) {
            int slot = i & 7;
            if (i >= 8) mbar_wait(&bars[slot], ((i >> 3) - 1) & 1);
            size_t c = ((size_t)i * gridDim.x + blockIdx.x) % nchunks;
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bars[slot])), "r"(chunk) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                             smem_u32(smem + (size_t)slot * chunk)),
                         "l"(src + c * chunk), "r"(chunk), "r"(smem_u32(&bars[slot]))
                         : "memory");
        }
        for (int i = per_cta; i < per_cta + 8; i++) {
            int slot = i & 7;
            if (i >= 8) mbar_wait(&bars[slot], ((i >> 3) - 1) & 1);
        }
    }
    __syncthreads();
}

static void run_bulk_bw(int nsm, size_t total_mb, int chunk) {
    uint8_t* src;
    size_t bytes = total_mb << 20;
    CK(cudaMalloc(&src, bytes));
    CK(cudaMemset(src, 1, bytes));
    size_t nchunks = bytes / chunk;
    int per_cta = 4096;
    size_t sm = (size_t)8 * chunk;
    CK(cudaFuncSetAttribute(bulk_bw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    bulk_bw_kernel<<<nsm, 128, sm>>>(src, nchunks, chunk, per_cta);
    CK(cudaDeviceSynchronize());
    cudaEventRecord(e0);
    bulk_bw_kernel<<<nsm, 128, sm>>>(src, nchunks, chunk, per_cta);
    cudaEventRecord(e1);
    CK(cudaDeviceSynchronize());
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    double gb = (double)per_cta * nsm * chunk / 1e9;
    printf("bulk copy global->smem, working set %4zu MB, chunk %5d B : %8.3f ms  %8.1f GB/s  (%.1f B/clk/SM at 1.965 GHz)\n", total_mb, chunk,
           ms, gb / (ms * 1e-3), gb * 1e9 / (ms * 1e-3) / nsm / 1.965e9);
    cudaFree(src);
}

template <int N>
static void run_check(int ksteps) {
    const int K = 32 * ksteps;
    std::vector<int8_t> A(128 * K), B(N * K);
    srand(1234 + N + ksteps);
    for (auto& x : A) x = (int8_t)(rand() % 129 - 64);
    for (auto& x : B) x = (int8_t)(rand() % 129 - 64);
    int8_t *dA, *dB;
    int32_t* dD;
    CK(cudaMalloc(&dA, A.size()));
    CK(cudaMalloc(&dB, B.size()));
    CK(cudaMalloc(&dD, 128 * N * 4));
    CK(cudaMemcpy(dA, A.data(), A.size(), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, B.data(), B.size(), cudaMemcpyHostToDevice));
    CK(cudaMemset(dD, 0xff, 128 * N * 4));
    size_t sm = (size_t)ksteps * (4096 + N * 32);
    CK(cudaFuncSetAttribute(check_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    check_kernel<N><<<1, 128, sm>>>(dA, dB, ksteps, dD);
    CK(cudaDeviceSynchronize());
    std::vector<int32_t> D(128 * N);
    CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
    long bad = 0;
    for (int i = 0; i < 128; i++)
        for (int j = 0; j < N; j++) {
            int32_t ref = 0;
            for (int k = 0; k < K; k++) ref += (int32_t)A[i * K + k] * (int32_t)B[j * K + k];
            if (ref != D[i * N + j]) {
                if (bad < 5) printf("   mismatch (%d,%d): got %d want %d\n", i, j, D[i * N + j], ref);
                bad++;
            }
        }
    printf("check M=128 N=%d K=%d : %s (%ld mismatches)\n", N, K, bad ? "FAIL" : "ok", bad);
    cudaFree(dA);
    cudaFree(dB);
    cudaFree(dD);
}

template <int N>
static void run_rate(int nsm) {
    int* sink;
    CK(cudaMalloc(&sink, 4));
    const int iters = 8192, nslots = 4;
    size_t sm = (size_t)nslots * (4096 + N * 32);
    CK(cudaFuncSetAttribute(rate_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    rate_kernel<N><<<nsm, 128, sm>>>(iters, nslots, sink);
    CK(cudaDeviceSynchronize());
    cudaEventRecord(e0);
    rate_kernel<N><<<nsm, 128, sm>>>(iters, nslots, sink);
    cudaEventRecord(e1);
    CK(cudaDeviceSynchronize());
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    double ops = 2.0 * 128 * N * 32 * (double)iters * nsm;
    printf("rate  1-CTA M=128 N=%3d : %8.3f ms  %8.1f TOPS  (%.1f clk/MMA at 1.965 GHz)\n", N, ms, ops / ms * 1e-9,
           ms * 1e-3 * 1.965e9 / iters);
    cudaFree(sink);
}

template <int N>
static void run_rate2(int nsm) {
    int* sink;
    CK(cudaMalloc(&sink, 4));
    const int iters = 8192, nslots = 4;
    size_t sm = (size_t)nslots * (4096 + N * 16);
    CK(cudaFuncSetAttribute(rate2_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    rate2_kernel<N><<<nsm, 128, sm>>>(iters, nslots, sink);
    CK(cudaDeviceSynchronize());
    cudaEventRecord(e0);
    rate2_kernel<N><<<nsm, 128, sm>>>(iters, nslots, sink);
    cudaEventRecord(e1);
    CK(cudaDeviceSynchronize());
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    double ops = 2.0 * 256 * N * 32 * (double)iters * (nsm / 2);
    printf("rate  2-CTA M=256 N=%3d : %8.3f ms  %8.1f TOPS  (%.1f clk/MMA at 1.965 GHz)\n", N, ms, ops / ms * 1e-9,
           ms * 1e-3 * 1.965e9 / iters);
    cudaFree(sink);
}

int main() {
    cudaDeviceProp p;
    CK(cudaGetDeviceProperties(&p, 0));
    printf("device %s  SMs %d\n", p.name, p.multiProcessorCount);
    run_check<64>(1);
    run_check<64>(4);
    run_check<256>(2);
    run_check<128>(3);
    int nsm = p.multiProcessorCount;
    run_rate<64>(nsm);
    run_rate<128>(nsm);
    run_rate<192>(nsm);
    run_rate<256>(nsm);
    run_rate2<64>(nsm);
    run_rate2<128>(nsm);
    run_rate2<256>(nsm);
    run_bulk_bw(nsm, 32, 16384);
    run_bulk_bw(nsm, 64, 16384);
    run_bulk_bw(nsm, 96, 16384);
    run_bulk_bw(nsm, 64, 8192);
    run_bulk_bw(nsm, 64, 24576);
    run_bulk_bw(nsm, 1024, 16384);
    return 0;
}
