// FP64 latency / single-warp issue probe (B200): dependent DFMA chain, independent DFMA stream from one warp,
// MUFU.RSQ64H + correction chain, DMMA dependent chain, shared-memory store->load hop.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 lat_probe.cu -o lat_probe
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__global__ void k(double* out, long long* cyc, double a, double b) {
    __shared__ double sh[64];
    double x = a + threadIdx.x;
    long long t0, t1;
    // 1. dependent DFMA chain
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < 256; i++) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(x) : "d"(b), "d"(a));
    t1 = clock64();
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
    // 2. 16 independent chains, one warp
    double y[16];
#pragma unroll
    for (int j = 0; j < 16; j++) y[j] = x + j;
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < 64; i++)
#pragma unroll
        for (int j = 0; j < 16; j++) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(y[j]) : "d"(b), "d"(a));
    t1 = clock64();
    if (threadIdx.x == 0) cyc[1] = t1 - t0;
#pragma unroll
    for (int j = 0; j < 16; j++) x += y[j];
    // 3. rsqrt seed + dependent multiply chain
    double z = fabs(x) + 1.0;
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < 64; i++) {
        double r;
        asm volatile("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(z));
        z = r + 1.5;
    }
    t1 = clock64();
    if (threadIdx.x == 0) cyc[2] = t1 - t0;  // 64 x (MUFU + DADD)
    x += z;
    // 4. dependent DMMA chain
    double c0 = x, c1 = a;
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < 64; i++) dmma884(c0, c1, a, b);
    t1 = clock64();
    if (threadIdx.x == 0) cyc[3] = t1 - t0;
    x += c0 + c1;
    // 5. smem hop: STS -> bar.warp.sync -> LDS (dependent)
    double v = x;
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < 64; i++) {
        sh[threadIdx.x] = v;
        __syncwarp();
        v = sh[(threadIdx.x + 1) & 31];
        __syncwarp();
    }
    t1 = clock64();
    if (threadIdx.x == 0) cyc[4] = t1 - t0;
    // 6. 4 independent DMMA accumulators
    double d[4][2];
#pragma unroll
    for (int j = 0; j < 4; j++) d[j][0] = d[j][1] = v + j;
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < 32; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) dmma884(d[j][0], d[j][1], a, b);
    t1 = clock64();
    if (threadIdx.x == 0) cyc[5] = t1 - t0;
#pragma unroll
    for (int j = 0; j < 4; j++) v += d[j][0] + d[j][1];
    // 7. shuffle dependent chain
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < 64; i++) v = __shfl_sync(0xffffffffu, v, (threadIdx.x + 1) & 31);
    t1 = clock64();
    if (threadIdx.x == 0) cyc[6] = t1 - t0;
    out[threadIdx.x] = x + v;
}
int main() {
    double* out; long long* cyc;
    cudaMalloc(&out, 32 * 8 * 8); cudaMalloc(&cyc, 64);
    for (int warps = 1; warps <= 8; warps *= 2) {
        k<<<1, 32 * warps>>>(out, cyc, 0.999, 1.0000001);
        cudaDeviceSynchronize();
        long long h[8]; cudaMemcpy(h, cyc, 56, cudaMemcpyDeviceToHost);
        printf("warps=%d: DFMA dependent %.1f cyc | DFMA 16-indep %.2f cyc/instr | MUFU.RSQ64H+DADD %.1f cyc | DMMA dependent %.1f | smem hop(STS,sync,LDS,sync) %.1f | DMMA 4-indep %.1f cyc/instr | SHFL(64-bit) dep %.1f\n",
               warps, h[0] / 256.0, h[1] / 1024.0, h[2] / 64.0, h[3] / 64.0, h[4] / 64.0, h[5] / 128.0, h[6] / 64.0);
    }
    return 0;
}
