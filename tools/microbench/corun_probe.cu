// Which SM pipes can run beside tcgen05.mma kind::i8?  A persistent int8 tensor-core kernel (M=128, N=256 MMAs back to
// back, one CTA per SM) runs on one stream; on a second stream a one-block-per-SM side kernel that exercises ONE resource
// (FP64 FMA, FP32 FMA, integer MAD, shared-memory loads, F2I/I2F conversions, MUFU) is timed alone and beside it.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o corun_probe corun_probe.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x)                                                                              \
    do {                                                                                   \
        cudaError_t e = (x);                                                               \
        if (e != cudaSuccess) {                                                            \
            printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); \
            exit(1);                                                                       \
        }                                                                                  \
    } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((saddr >> 4) & 0x3fff) | ((uint64_t)((lbo >> 4) & 0x3fff) << 16) | ((uint64_t)((sbo >> 4) & 0x3fff) << 32) |
           ((uint64_t)1 << 46);
}
__device__ __forceinline__ void mma_i8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}

constexpr int MMA_N = 256;
__global__ void __launch_bounds__(128) mma_kernel(int iters, int* sink) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base;
    constexpr int nslots = 4;
    for (int e = threadIdx.x; e < nslots * (4096 + MMA_N * 32) / 4; e += 128) ((uint32_t*)smem)[e] = 0x01010101u * (e & 3);
    if (threadIdx.x == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
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
        const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(MMA_N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        uint64_t ad[4], bd[4];
        for (int q = 0; q < 4; q++) {
            ad[q] = make_desc(smem_u32(smem + q * 4096), 128, 256);
            bd[q] = make_desc(smem_u32(smem + nslots * 4096 + q * (MMA_N * 32)), 128, 256);
        }
        mma_i8(tbase, ad[0], bd[0], idesc, 0);
        mma_i8(tbase + MMA_N, ad[1], bd[1], idesc, 0);
        for (int it = 2; it + 4 <= iters; it += 4) {
            mma_i8(tbase, ad[2], bd[2], idesc, 1);
            mma_i8(tbase + MMA_N, ad[3], bd[3], idesc, 1);
            mma_i8(tbase, ad[0], bd[0], idesc, 1);
            mma_i8(tbase + MMA_N, ad[1], bd[1], idesc, 1);
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    asm volatile(
        "{\n\t.reg .pred P1;\n\tW:\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], 0;\n\t@P1 bra D;\n\tbra W;\n\tD:\n\t}\n" ::"r"(
            smem_u32(&bar))
        : "memory");
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t v;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(v) : "r"(tbase + ((uint32_t)((threadIdx.x >> 5) * 32) << 16)));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    if (v == 0x12345678u) sink[0] = 1;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase), "r"(512));
}

// ---- side kernels: 256 threads, 8 independent chains per thread --------------------------------------------------------
__global__ void __launch_bounds__(256) dfma_kernel(int iters, double* out) {
    double a[8];
    for (int i = 0; i < 8; i++) a[i] = threadIdx.x * 1e-3 + i;
    const double m = 1.0000001, c = 1e-9;
    for (int it = 0; it < iters; it++)
#pragma unroll
        for (int i = 0; i < 8; i++) a[i] = fma(a[i], m, c);
    double s = 0;
    for (int i = 0; i < 8; i++) s += a[i];
    if (s == 12345.0) out[0] = s;
}
__global__ void __launch_bounds__(256) ffma_kernel(int iters, float* out) {
    float a[8];
    for (int i = 0; i < 8; i++) a[i] = threadIdx.x * 1e-3f + i;
    const float m = 1.0000001f, c = 1e-9f;
    for (int it = 0; it < iters; it++)
#pragma unroll
        for (int i = 0; i < 8; i++) a[i] = fmaf(a[i], m, c);
    float s = 0;
    for (int i = 0; i < 8; i++) s += a[i];
    if (s == 12345.0f) out[0] = s;
}
__global__ void __launch_bounds__(256) imad_kernel(int iters, int* out) {
    int a[8];
    for (int i = 0; i < 8; i++) a[i] = threadIdx.x + i;
    int m = 3 + (int)blockIdx.x, c = 7;
    for (int it = 0; it < iters; it++)
#pragma unroll
        for (int i = 0; i < 8; i++) a[i] = a[i] * m + c;
    int s = 0;
    for (int i = 0; i < 8; i++) s += a[i];
    if (s == 12345) out[0] = s;
}
__global__ void __launch_bounds__(256) lds_kernel(int iters, double* out) {
    __shared__ double buf[2048];
    for (int e = threadIdx.x; e < 2048; e += 256) buf[e] = e;
    __syncthreads();
    double s = 0;
    int idx = threadIdx.x;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) s += buf[(idx + i * 256) & 2047];
        idx = (idx + 17) & 2047;
    }
    if (s == 12345.0) out[0] = s;
}
__global__ void __launch_bounds__(256) cvt_kernel(int iters, double* out) {  // F2I.S64.F64 + I2F.F64.S64 round trips
    double a[8];
    for (int i = 0; i < 8; i++) a[i] = threadIdx.x * 3.7 + i;
    for (int it = 0; it < iters; it++)
#pragma unroll
        for (int i = 0; i < 8; i++) a[i] = (double)(__double2ll_rn(a[i]) ^ 1LL);
    double s = 0;
    for (int i = 0; i < 8; i++) s += a[i];
    if (s == 12345.0) out[0] = s;
}
__global__ void __launch_bounds__(256) rsq_kernel(int iters, double* out) {  // MUFU.RSQ64H
    double a[8];
    for (int i = 0; i < 8; i++) a[i] = threadIdx.x * 3.7 + i + 1.0;
    for (int it = 0; it < iters; it++)
#pragma unroll
        for (int i = 0; i < 8; i++) asm volatile("rsqrt.approx.ftz.f64 %0, %0;" : "+d"(a[i]));
    double s = 0;
    for (int i = 0; i < 8; i++) s += a[i];
    if (s == 12345.0) out[0] = s;
}

template <class F>
static void corun(const char* name, F side, int nsm, int mma_iters, double unit_ops, const char* unit) {
    cudaStream_t sa, sb;
    CK(cudaStreamCreateWithFlags(&sa, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&sb, cudaStreamNonBlocking));
    cudaEvent_t a0, a1, b0, b1;
    cudaEventCreate(&a0); cudaEventCreate(&a1); cudaEventCreate(&b0); cudaEventCreate(&b1);
    int* sink;
    CK(cudaMalloc(&sink, 64));
    const size_t sm = 4 * (4096 + MMA_N * 32);
    float alone_side, alone_mma, co_side, co_mma;
    // alone
    side(sb); CK(cudaDeviceSynchronize());
    cudaEventRecord(b0, sb); side(sb); cudaEventRecord(b1, sb); CK(cudaDeviceSynchronize());
    cudaEventElapsedTime(&alone_side, b0, b1);
    mma_kernel<<<nsm, 128, sm, sa>>>(mma_iters, sink); CK(cudaDeviceSynchronize());
    cudaEventRecord(a0, sa); mma_kernel<<<nsm, 128, sm, sa>>>(mma_iters, sink); cudaEventRecord(a1, sa); CK(cudaDeviceSynchronize());
    cudaEventElapsedTime(&alone_mma, a0, a1);
    // together: the tensor kernel first (it is the longer one), the side kernel 1 block per SM beside it
    cudaEventRecord(a0, sa); mma_kernel<<<nsm, 128, sm, sa>>>(mma_iters, sink); cudaEventRecord(a1, sa);
    cudaEventRecord(b0, sb); side(sb); cudaEventRecord(b1, sb);
    CK(cudaDeviceSynchronize());
    cudaEventElapsedTime(&co_side, b0, b1);
    cudaEventElapsedTime(&co_mma, a0, a1);
    printf("%-6s side alone %7.3f ms (%8.2f %s)  beside MMA %7.3f ms (x%.2f) | MMA alone %7.3f ms  beside side %7.3f ms (x%.2f)\n", name,
           alone_side, unit_ops / alone_side * 1e-9, unit, co_side, co_side / alone_side, alone_mma, co_mma, co_mma / alone_mma);
    cudaFree(sink);
}

int main() {
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int nsm = prop.multiProcessorCount;
    printf("device: %s, %d SMs\n", prop.name, nsm);
    CK(cudaFuncSetAttribute(mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * (4096 + MMA_N * 32)));
    double* dout;
    CK(cudaMalloc(&dout, 64));
    const int mma_iters = 65536;  // x 134.7 clk = 8.8e6 clk ~ 4.5 ms
    const int it64 = 40000, it32 = 160000;
    const double th = 256.0 * nsm * 8;
    corun("DFMA", [&](cudaStream_t s) { dfma_kernel<<<nsm, 256, 0, s>>>(it64, dout); }, nsm, mma_iters, th * it64 * 2, "TFLOP/s");
    corun("FFMA", [&](cudaStream_t s) { ffma_kernel<<<nsm, 256, 0, s>>>(it32, (float*)dout); }, nsm, mma_iters, th * it32 * 2, "TFLOP/s");
    corun("IMAD", [&](cudaStream_t s) { imad_kernel<<<nsm, 256, 0, s>>>(it32, (int*)dout); }, nsm, mma_iters, th * it32, "TIOP/s");
    corun("LDS", [&](cudaStream_t s) { lds_kernel<<<nsm, 256, 0, s>>>(it32 / 4, dout); }, nsm, mma_iters, th * (it32 / 4) * 8, "TB/s");
    corun("CVT64", [&](cudaStream_t s) { cvt_kernel<<<nsm, 256, 0, s>>>(it64 / 8, dout); }, nsm, mma_iters, th * (it64 / 8) * 2, "Tcvt/s");
    corun("RSQ64", [&](cudaStream_t s) { rsq_kernel<<<nsm, 256, 0, s>>>(it64 / 8, dout); }, nsm, mma_iters, th * (it64 / 8), "Tmufu/s");
    // 2 and 3 side blocks per SM (more warps): does the FP64 side kernel scale beside the MMA kernel?
    corun("DFMAx2", [&](cudaStream_t s) { dfma_kernel<<<2 * nsm, 256, 0, s>>>(it64, dout); }, nsm, mma_iters, 2 * th * it64 * 2, "TFLOP/s");
    return 0;
}
