// Covariance-function kernels: input scaling, Gram build, cross-covariance window (+ posterior mean), and the
// log-marginal-likelihood gradient reduction.  All HBM access is coalesced along the training-point index j:
// the scaled inputs are stored dimension-major, Xs[dim][Np], so consecutive lanes read consecutive doubles.
#pragma once
#include "common.cuh"

namespace gpso {

// Xs[dim*Np + j] = X[j*d + dim] / ls[dim]   (GPflow divides by the lengthscale, Stationary.scale); pad j>=N with 0
__global__ void scale_inputs_kernel(const double* __restrict__ X, const double* __restrict__ ls, int n_ls, int N, int d,
                                    int Np, double* __restrict__ Xs) {
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= Np) return;
    for (int dim = 0; dim < d; dim++) {
        double l = ls[n_ls > 1 ? dim : 0];
        Xs[(size_t)dim * Np + j] = (j < N) ? X[(size_t)j * d + dim] / l : 0.0;
    }
}

// ---- Gram matrix K + noise*I, lower 64x64 tiles of the padded [Np,Np] matrix; padding = identity ----------------
constexpr int CT = 64;    // covariance tile edge
constexpr int CDCH = 32;  // dimensions staged in shared memory per pass

template <int KID>
__global__ void __launch_bounds__(256) gram_kernel(const double* __restrict__ Xs, int N, int d, int Np, double var,
                                                   double noise, double* __restrict__ K) {
    const int bj = blockIdx.x, bi = blockIdx.y;
    if (bj > bi) return;
    __shared__ double sXi[CDCH][CT];
    const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;
    const int j = bj * CT + tx;
    double r2[16];
#pragma unroll
    for (int r = 0; r < 16; r++) r2[r] = 0.0;
    for (int d0 = 0; d0 < d; d0 += CDCH) {
        int dc = min(CDCH, d - d0);
        __syncthreads();
        for (int e = threadIdx.x; e < dc * CT; e += 256) sXi[e / CT][e % CT] = Xs[(size_t)(d0 + e / CT) * Np + bi * CT + e % CT];
        __syncthreads();
        for (int dd = 0; dd < dc; dd++) {
            double xj = Xs[(size_t)(d0 + dd) * Np + j];
#pragma unroll
            for (int r = 0; r < 16; r++) {
                double df = sXi[dd][ty + 4 * r] - xj;
                r2[r] = fma(df, df, r2[r]);
            }
        }
    }
#pragma unroll
    for (int r = 0; r < 16; r++) {
        int i = bi * CT + ty + 4 * r;
        double v;
        if (i < N && j < N) {
            v = cov_from_r2<KID>(r2[r], var);
            if (i == j) v += noise;
        } else {
            v = (i == j) ? 1.0 : 0.0;
        }
        K[(size_t)i * Np + j] = v;
    }
}

// ---- cross-covariance window -------------------------------------------------------------------------------------
// KsT[c][j] = k(x_j, xc_c) for the window's candidates c (row-major [Mw_pad, Np], j contiguous -> it is the "B"
// operand of the triangular product), zero for j >= N and for padding candidates c >= Mw; and the posterior mean
// mean[c] = sum_j KsT[c][j] * alpha[j] + c0.  One block = XG candidates x all training points: each scaled training
// coordinate is loaded once per block and reused for XG candidates.  Every candidate sees exactly the same sequence
// of operations (position-independent results: duplicated candidates give bit-identical outputs).
constexpr int XG = 8;

template <int KID>
__global__ void __launch_bounds__(256) crosscov_kernel(const double* __restrict__ Xc, long long Mw, int d,
                                                       const double* __restrict__ ls, int n_ls,
                                                       const double* __restrict__ Xs, const double* __restrict__ alpha,
                                                       int N, int Np, double var, double c0, double* __restrict__ KsT,
                                                       double* __restrict__ mean) {
    extern __shared__ double sm[];  // [XG][d] scaled candidate coords, then [8 warps][XG] reduction scratch
    double* sC = sm;
    double* sR = sm + XG * d;
    const long long cbase = (long long)blockIdx.x * XG;
    for (int e = threadIdx.x; e < XG * d; e += 256) {
        long long c = cbase + e / d;
        int dim = e % d;
        sC[e] = (c < Mw) ? Xc[c * d + dim] / ls[n_ls > 1 ? dim : 0] : 0.0;
    }
    __syncthreads();
    double macc[XG];
#pragma unroll
    for (int c = 0; c < XG; c++) macc[c] = 0.0;
    for (int j = threadIdx.x; j < Np; j += 256) {
        double r2[XG];
#pragma unroll
        for (int c = 0; c < XG; c++) r2[c] = 0.0;
        for (int dim = 0; dim < d; dim++) {
            double xj = Xs[(size_t)dim * Np + j];
#pragma unroll
            for (int c = 0; c < XG; c++) {
                double df = xj - sC[c * d + dim];
                r2[c] = fma(df, df, r2[c]);
            }
        }
        double a = alpha[j];  // zero for j >= N
#pragma unroll
        for (int c = 0; c < XG; c++) {
            double k = (j < N && cbase + c < Mw) ? cov_from_r2<KID>(r2[c], var) : 0.0;
            KsT[(size_t)(cbase + c) * Np + j] = k;
            macc[c] = fma(k, a, macc[c]);
        }
    }
#pragma unroll
    for (int c = 0; c < XG; c++) macc[c] = warp_sum(macc[c]);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) {
#pragma unroll
        for (int c = 0; c < XG; c++) sR[w * XG + c] = macc[c];
    }
    __syncthreads();
    if (threadIdx.x < XG) {
        double s = 0.0;
#pragma unroll
        for (int ww = 0; ww < 8; ww++) s += sR[ww * XG + threadIdx.x];
        mean[cbase + threadIdx.x] = s + c0;
    }
}

// ---- LML gradient: sum_ij W_ij dK_ij/dtheta with W = alpha alpha^T - K_y^-1, never materialising dK -------------
// One block per lower 64x64 tile (bi >= bj).  Strictly-lower tiles count twice (symmetry).  Partial sums per block:
//   part[blk][0] = sum W*k            (-> d/d variance after / variance)
//   part[blk][1] = sum_i W_ii         (-> d/d noise)
//   part[blk][2 + q] = sum W*g*r2     (scalar lengthscale, q = 0)  or  sum W*g*(xs_i - xs_j)_dim^2 for dim = dim0+q
constexpr int GRAD_DCH = 4;  // ARD dimensions reduced per launch

template <int KID, bool ARD>
__global__ void __launch_bounds__(256) lml_grad_kernel(const double* __restrict__ Xs, const double* __restrict__ alpha,
                                                       const double* __restrict__ Kinv, int N, int d, int Np, double var,
                                                       int dim0, int ndim, double* __restrict__ part, int part_stride) {
    // triangular block index -> (bi, bj)
    int blk = blockIdx.x;
    int bi = (int)((sqrt(8.0 * blk + 1.0) - 1.0) * 0.5);
    while ((bi + 1) * (bi + 2) / 2 <= blk) bi++;
    while (bi * (bi + 1) / 2 > blk) bi--;
    int bj = blk - bi * (bi + 1) / 2;
    __shared__ double sXi[CDCH][CT];
    __shared__ double sRed[8];
    const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;
    const int j = bj * CT + tx;
    double r2[16];
    double gd[16][ARD ? GRAD_DCH : 1];
#pragma unroll
    for (int r = 0; r < 16; r++) {
        r2[r] = 0.0;
#pragma unroll
        for (int q = 0; q < (ARD ? GRAD_DCH : 1); q++) gd[r][q] = 0.0;
    }
    for (int d0 = 0; d0 < d; d0 += CDCH) {
        int dc = min(CDCH, d - d0);
        __syncthreads();
        for (int e = threadIdx.x; e < dc * CT; e += 256) sXi[e / CT][e % CT] = Xs[(size_t)(d0 + e / CT) * Np + bi * CT + e % CT];
        __syncthreads();
        for (int dd = 0; dd < dc; dd++) {
            double xj = Xs[(size_t)(d0 + dd) * Np + j];
            int q = d0 + dd - dim0;
#pragma unroll
            for (int r = 0; r < 16; r++) {
                double df = sXi[dd][ty + 4 * r] - xj;
                double sq = df * df;
                r2[r] += sq;
                if (ARD) {
#pragma unroll
                    for (int qq = 0; qq < GRAD_DCH; qq++)
                        if (qq == q) gd[r][qq] = sq;
                }
            }
        }
    }
    double s_k = 0.0, s_tr = 0.0, s_l[ARD ? GRAD_DCH : 1];
#pragma unroll
    for (int q = 0; q < (ARD ? GRAD_DCH : 1); q++) s_l[q] = 0.0;
    const double aj = alpha[j];
    const double wsym = (bi == bj) ? 1.0 : 2.0;
#pragma unroll
    for (int r = 0; r < 16; r++) {
        int i = bi * CT + ty + 4 * r;
        // diagonal tiles: the tile holds both triangles of a symmetric quantity, take every element once
        if (i < N && j < N) {
            double kin = (j <= i) ? Kinv[(size_t)i * Np + j] : Kinv[(size_t)j * Np + i];
            double W = fma(alpha[i], aj, -kin) * wsym;
            double k, g;
            cov_and_radial<KID>(r2[r], var, k, g);
            s_k = fma(W, k, s_k);
            if (i == j) s_tr += W;
            double wg = W * g;
            if (ARD) {
#pragma unroll
                for (int q = 0; q < GRAD_DCH; q++) s_l[q] = fma(wg, gd[r][q], s_l[q]);
            } else {
                s_l[0] = fma(wg, fmax(r2[r], 0.0), s_l[0]);
            }
        }
    }
    double v;
    v = block_sum<256>(s_k, sRed);
    if (threadIdx.x == 0 && dim0 == 0) part[(size_t)blk * part_stride + 0] = v;
    v = block_sum<256>(s_tr, sRed);
    if (threadIdx.x == 0 && dim0 == 0) part[(size_t)blk * part_stride + 1] = v;
#pragma unroll
    for (int q = 0; q < (ARD ? GRAD_DCH : 1); q++) {
        v = block_sum<256>(s_l[q], sRed);
        if (threadIdx.x == 0 && q < ndim) part[(size_t)blk * part_stride + 2 + dim0 + q] = v;
    }
}

// out[q] = sum_blk part[blk][q], fixed order; one block per q
__global__ void __launch_bounds__(256) reduce_partials_kernel(const double* __restrict__ part, int nblk, int stride,
                                                              double* __restrict__ out) {
    __shared__ double sRed[8];
    int q = blockIdx.x;
    double s = 0.0;
    for (int b = threadIdx.x; b < nblk; b += 256) s += part[(size_t)b * stride + q];
    double v = block_sum<256>(s, sRed);
    if (threadIdx.x == 0) out[q] = v;
}

}  // namespace gpso
