// Common device helpers for the gpso_b200 kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

namespace gpso {

constexpr int KERNEL_MATERN12 = 0;
constexpr int KERNEL_MATERN32 = 1;
constexpr int KERNEL_MATERN52 = 2;
constexpr int KERNEL_SE = 3;

constexpr int TB = 128;           // tile / panel edge: every dense matrix is padded to a multiple of TB
constexpr double R2_CLIP = 1e-36; // GPflow clips the scaled squared distance before the sqrt (Matern kernels)

// tf.maximum(r2, 1e-36) propagates NaN (a NaN input must reach the Cholesky and be reported, not be clipped away);
// fmax() would return the non-NaN operand
__device__ __forceinline__ double clip_r2(double r2) { return (r2 < R2_CLIP) ? R2_CLIP : r2; }

// ---- FP64 tensor op: D(8x8) = A(8x4) * B(4x8) + C.  SASS: DMMA.8x8x4 -------------------------------------------
// fragment layout (g = lane>>2, t = lane&3):  a = A[g][t],  b = B[t][g],  c0 = C[g][2t], c1 = C[g][2t+1]
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

// ---- cp.async (LDGSTS) 16-byte global->shared copies ------------------------------------------------------------
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}

// ---- covariance functions (gpflow.kernels.stationaries) ----------------------------------------------------------
// r2 = scaled squared distance (inputs already divided by the lengthscale), var = kernel variance
template <int KID>
__device__ __forceinline__ double cov_from_r2(double r2, double var) {
    if (KID == KERNEL_SE) {
        return var * exp(-0.5 * r2);
    } else {
        double r = sqrt(clip_r2(r2));
        if (KID == KERNEL_MATERN52) {
            const double s5 = 2.23606797749978969641;
            double sr = s5 * r;
            // variance * (1 + sqrt5 r + 5/3 r^2) * exp(-sqrt5 r), GPflow's evaluation order
            return var * (1.0 + sr + (5.0 / 3.0) * (r * r)) * exp(-sr);
        } else if (KID == KERNEL_MATERN32) {
            const double s3 = 1.73205080756887729353;
            double sr = s3 * r;
            return var * (1.0 + sr) * exp(-sr);
        } else {
            return var * exp(-r);
        }
    }
}

// covariance value and the radial factor g with  dK/d(ls_j) = g * Delta_j^2 / ls_j^3  (Delta in UNscaled units),
// i.e. for a scalar lengthscale dK/d(ls) = g * r2 / ls.  g = 0 where r2 was clipped (GPflow: zero gradient there).
template <int KID>
__device__ __forceinline__ void cov_and_radial(double r2, double var, double& k, double& g) {
    if (KID == KERNEL_SE) {
        k = var * exp(-0.5 * r2);
        g = k;
    } else {
        bool live = r2 > R2_CLIP;
        double r = sqrt(clip_r2(r2));
        if (KID == KERNEL_MATERN52) {
            const double s5 = 2.23606797749978969641;
            double sr = s5 * r;
            double e = exp(-sr);
            k = var * (1.0 + sr + (5.0 / 3.0) * (r * r)) * e;
            g = live ? (5.0 / 3.0) * var * (1.0 + sr) * e : 0.0;
        } else if (KID == KERNEL_MATERN32) {
            const double s3 = 1.73205080756887729353;
            double sr = s3 * r;
            double e = exp(-sr);
            k = var * (1.0 + sr) * e;
            g = live ? 3.0 * var * e : 0.0;
        } else {
            k = var * exp(-r);
            g = live ? k / r : 0.0;
        }
    }
}

// ---- deterministic block reductions -----------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// sum over a block of NT threads (NT multiple of 32, <= 1024); result valid in thread 0; fixed order
template <int NT>
__device__ __forceinline__ double block_sum(double v, double* scratch /* >= NT/32 doubles */) {
    v = warp_sum(v);
    int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    __syncthreads();
    if (l == 0) scratch[w] = v;
    __syncthreads();
    double s = 0.0;
    if (threadIdx.x == 0) {
#pragma unroll
        for (int i = 0; i < NT / 32; i++) s += scratch[i];
    }
    return s;
}

// arg-max record with numpy semantics: first NaN wins, otherwise largest value, lowest index on ties
struct Best {
    double val;
    long long idx;
};
__device__ __forceinline__ bool best_better(double av, long long ai, double bv, long long bi) {
    bool an = isnan(av), bn = isnan(bv);
    if (an || bn) {
        if (an && bn) return ai < bi;
        return an;
    }
    if (av > bv) return true;
    if (av < bv) return false;
    return ai < bi;
}

}  // namespace gpso
