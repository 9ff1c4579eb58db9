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

// ---- exp(-s) for s >= 0 -----------------------------------------------------------------------------------------------
// Every covariance function on the path evaluates exp of a non-positive argument.  Branch-free Cody-Waite reduction
// (n = rint(-s log2 e) by the 1.5*2^52 trick, r = -s - n ln2 in two FMAs) and a degree-13 Horner polynomial; the result
// is scaled by adding n to the exponent field.  19 FP64-pipe operations against ~26 DFMA-equivalents of exp()
// (profiles/r01_fp64_probe.txt) and no per-call constant materialisation; max relative error 2.2e-16 on [0, 690]
// (the same as libm's exp against long double), exactly 1 at s = 0, 0 beyond s = 700 (true value < 1e-304), NaN -> NaN.
__device__ __forceinline__ double exp_neg(double s) {
    const double MAGIC = 6755399441055744.0;
    double t = fma(-s, 1.4426950408889634074, MAGIC);
    double nf = t - MAGIC;
    int n = __double2loint(t);
    double r = fma(-nf, 6.93147180369123816490e-01, -s);
    r = fma(-nf, 1.90821492927058770002e-10, r);
    double p = 1.0 / 6227020800.0;
    p = fma(p, r, 1.0 / 479001600.0);
    p = fma(p, r, 1.0 / 39916800.0);
    p = fma(p, r, 1.0 / 3628800.0);
    p = fma(p, r, 1.0 / 362880.0);
    p = fma(p, r, 1.0 / 40320.0);
    p = fma(p, r, 1.0 / 5040.0);
    p = fma(p, r, 1.0 / 720.0);
    p = fma(p, r, 1.0 / 120.0);
    p = fma(p, r, 1.0 / 24.0);
    p = fma(p, r, 1.0 / 6.0);
    p = fma(p, r, 0.5);
    p = fma(p, r, 1.0);
    p = fma(p, r, 1.0);
    double v = __longlong_as_double(__double_as_longlong(p) + ((long long)n << 52));
    v = (s > 700.0) ? 0.0 : v;
    return (s != s) ? s : v;
}

// ---- sqrt of a positive finite number ---------------------------------------------------------------------------------
// sqrt(x) = x / sqrt(x):  hardware seed y ~ 1/sqrt(x) (MUFU.RSQ64H, ~2^-22) and one third-order correction
// y(1 + e/2 + 3e^2/8), e = 1 - x y^2 (remaining error < 2^-64 before the final rounding; result within 1.5 ulp, measured
// 1.4e-16 in tools/microbench/diag_probe.cu).  7 FP64-pipe operations instead of the ~17 DFMA-equivalents of the IEEE
// sqrt() (profiles/r01_fp64_probe.txt).  The argument is always >= 1e-36 here (GPflow's clip).  NaN -> NaN.
__device__ __forceinline__ double sqrt_pos(double x) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double t = x * y;          // ~ sqrt(x)
    const double e = fma(-t, y, 1.0);
    const double pl = fma(0.375, e, 0.5);
    const double q = t * e;
    return fma(q, pl, t);
}

// ---- covariance functions (gpflow.kernels.stationaries) ----------------------------------------------------------
// r2 = scaled squared distance (inputs already divided by the lengthscale), var = kernel variance
template <int KID>
__device__ __forceinline__ double cov_from_r2(double r2, double var) {
    if (KID == KERNEL_SE) {
        return var * exp_neg(0.5 * r2);
    } else {
        double r = sqrt_pos(clip_r2(r2));
        if (KID == KERNEL_MATERN52) {
            const double s5 = 2.23606797749978969641;
            double sr = s5 * r;
            // variance * (1 + sqrt5 r + 5/3 r^2) * exp(-sqrt5 r), GPflow's evaluation order
            return var * (1.0 + sr + (5.0 / 3.0) * (r * r)) * exp_neg(sr);
        } else if (KID == KERNEL_MATERN32) {
            const double s3 = 1.73205080756887729353;
            double sr = s3 * r;
            return var * (1.0 + sr) * exp_neg(sr);
        } else {
            return var * exp_neg(r);
        }
    }
}

// ---- W-wide variants for the hot cross-covariance loop ---------------------------------------------------------------
// The same operation sequences as exp_neg / sqrt_pos / cov_from_r2 (bit-identical results), written over W independent
// arguments so that the W dependent chains interleave in the instruction stream (the scalar forms compile to one serial
// chain per element, ~8 clk between dependent DFMAs), with the 64-bit polynomial coefficients read from the constant bank
// as DFMA operands instead of being rebuilt by two UMOV / IMAD.MOV per use (28 of 140 instructions per element before).
__constant__ double EXPNEG_C[16] = {
    1.4426950408889634074,        // 0: log2(e)
    6.93147180369123816490e-01,   // 1: ln2 high
    1.90821492927058770002e-10,   // 2: ln2 low
    1.0 / 6227020800.0, 1.0 / 479001600.0, 1.0 / 39916800.0, 1.0 / 3628800.0, 1.0 / 362880.0, 1.0 / 40320.0, 1.0 / 5040.0,
    1.0 / 720.0, 1.0 / 120.0, 1.0 / 24.0, 1.0 / 6.0,  // 3..13: 1/13! .. 1/3!
    2.23606797749978969641,       // 14: sqrt(5)
    1.73205080756887729353,       // 15: sqrt(3)
};

template <int W>
__device__ __forceinline__ void exp_neg_v(const double (&s)[W], double (&v)[W]) {
    const double MAGIC = 6755399441055744.0;
    double nf[W], r[W], p[W];
    int n[W];
#pragma unroll
    for (int i = 0; i < W; i++) {
        double t = fma(-s[i], EXPNEG_C[0], MAGIC);
        nf[i] = t - MAGIC;
        n[i] = __double2loint(t);
    }
#pragma unroll
    for (int i = 0; i < W; i++) r[i] = fma(-nf[i], EXPNEG_C[1], -s[i]);
#pragma unroll
    for (int i = 0; i < W; i++) r[i] = fma(-nf[i], EXPNEG_C[2], r[i]);
#pragma unroll
    for (int i = 0; i < W; i++) p[i] = fma(EXPNEG_C[3], r[i], EXPNEG_C[4]);
#pragma unroll
    for (int c = 5; c <= 13; c++) {
#pragma unroll
        for (int i = 0; i < W; i++) p[i] = fma(p[i], r[i], EXPNEG_C[c]);
    }
#pragma unroll
    for (int i = 0; i < W; i++) p[i] = fma(p[i], r[i], 0.5);
#pragma unroll
    for (int i = 0; i < W; i++) p[i] = fma(p[i], r[i], 1.0);
#pragma unroll
    for (int i = 0; i < W; i++) p[i] = fma(p[i], r[i], 1.0);
#pragma unroll
    for (int i = 0; i < W; i++) {
        double x = __longlong_as_double(__double_as_longlong(p[i]) + ((long long)n[i] << 52));
        x = (s[i] > 700.0) ? 0.0 : x;
        v[i] = (s[i] != s[i]) ? s[i] : x;
    }
}

template <int W>
__device__ __forceinline__ void sqrt_pos_v(const double (&x)[W], double (&out)[W]) {
    double y[W], t[W], e[W];
#pragma unroll
    for (int i = 0; i < W; i++) asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y[i]) : "d"(x[i]));
#pragma unroll
    for (int i = 0; i < W; i++) t[i] = x[i] * y[i];
#pragma unroll
    for (int i = 0; i < W; i++) e[i] = fma(-t[i], y[i], 1.0);
#pragma unroll
    for (int i = 0; i < W; i++) {
        const double pl = fma(0.375, e[i], 0.5);
        const double q = t[i] * e[i];
        out[i] = fma(q, pl, t[i]);
    }
}

template <int KID, int W>
__device__ __forceinline__ void cov_from_r2_v(const double (&r2)[W], double var, double (&k)[W]) {
    double s[W], e[W];
    if (KID == KERNEL_SE) {
#pragma unroll
        for (int i = 0; i < W; i++) s[i] = 0.5 * r2[i];
        exp_neg_v<W>(s, e);
#pragma unroll
        for (int i = 0; i < W; i++) k[i] = var * e[i];
    } else {
        double c[W], r[W];
#pragma unroll
        for (int i = 0; i < W; i++) c[i] = clip_r2(r2[i]);
        sqrt_pos_v<W>(c, r);
        if (KID == KERNEL_MATERN52) {
#pragma unroll
            for (int i = 0; i < W; i++) s[i] = EXPNEG_C[14] * r[i];
            exp_neg_v<W>(s, e);
#pragma unroll
            for (int i = 0; i < W; i++) k[i] = var * (1.0 + s[i] + (5.0 / 3.0) * (r[i] * r[i])) * e[i];
        } else if (KID == KERNEL_MATERN32) {
#pragma unroll
            for (int i = 0; i < W; i++) s[i] = EXPNEG_C[15] * r[i];
            exp_neg_v<W>(s, e);
#pragma unroll
            for (int i = 0; i < W; i++) k[i] = var * (1.0 + s[i]) * e[i];
        } else {
            exp_neg_v<W>(r, e);
#pragma unroll
            for (int i = 0; i < W; i++) k[i] = var * e[i];
        }
    }
}

// covariance value and the radial factor g with  dK/d(ls_j) = g * Delta_j^2 / ls_j^3  (Delta in UNscaled units),
// i.e. for a scalar lengthscale dK/d(ls) = g * r2 / ls.  g = 0 where r2 was clipped (GPflow: zero gradient there).
template <int KID>
__device__ __forceinline__ void cov_and_radial(double r2, double var, double& k, double& g) {
    if (KID == KERNEL_SE) {
        k = var * exp_neg(0.5 * r2);
        g = k;
    } else {
        bool live = r2 > R2_CLIP;
        double r = sqrt_pos(clip_r2(r2));
        if (KID == KERNEL_MATERN52) {
            const double s5 = 2.23606797749978969641;
            double sr = s5 * r;
            double e = exp_neg(sr);
            k = var * (1.0 + sr + (5.0 / 3.0) * (r * r)) * e;
            g = live ? (5.0 / 3.0) * var * (1.0 + sr) * e : 0.0;
        } else if (KID == KERNEL_MATERN32) {
            const double s3 = 1.73205080756887729353;
            double sr = s3 * r;
            double e = exp_neg(sr);
            k = var * (1.0 + sr) * e;
            g = live ? 3.0 * var * e : 0.0;
        } else {
            k = var * exp_neg(r);
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
