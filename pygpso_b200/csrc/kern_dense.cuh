// Dense fp64 linear algebra of the fit: blocked right-looking Cholesky, triangular inverse by recursive doubling,
// K_y^-1 = L^-T L^-1, and the triangular matrix-vector products for a = L^-1 (y - c), alpha = L^-T a.
// Every O(N^3) step is a batch of 128x128 tiles of the DMMA core in gemm_core.cuh; only the 128x128 diagonal blocks
// are factorised / inverted by a single CTA in shared memory.
//
// Storage: all matrices row-major [Np, Np], Np = N padded to a multiple of 128.
//   K     : Gram + noise (lower) -> overwritten in place by L (lower)
//   Linv  : L^-1 (lower, explicit zeros above the diagonal inside diagonal blocks)
//   LinvT : (L^-1)^T (upper) -- kept as well because the tile core wants both operands K-contiguous
//   T     : scratch for the inverse recursion,  Kinv : K_y^-1 (lower)
#pragma once
#include "gemm_core.cuh"

namespace gpso {

// ---- diagonal block: Cholesky of the 128x128 block p, its inverse, and sum(log(diag)) ----------------------------
// In:  K[p-block] (lower triangle used).  Out: L_pp -> K (lower), L_pp^-1 -> Linv (lower, zeros above),
// (L_pp^-1)^T -> LinvT, logdet[p] = sum_j log(L_jj) over the real (un-padded) rows, info = first non-positive pivot.
//
// This kernel is the serial spine of the factorisation (one launch per 128-column panel), so it is organised to keep
// the dependent chain short: the block is a 4x4 grid of 32x32 sub-blocks held in shared memory.
//   * a 32x32 diagonal sub-block is factorised and inverted by ONE warp with its rows in registers (pivot broadcast by
//     shuffle, no block barrier inside the 32 column steps);
//   * everything else is 32x32x32 sub-block products done by four 64-thread groups in parallel (4x4 register tiles):
//     panel  L_ip = A_ip D_pp^-T,  trailing  A_ik -= L_ip L_kp^T,  and the in-place block inverse, column by column
//     from the right:  X_ij = -( sum_{k=j+1..i} X_ik L_kj ) X_jj.
constexpr int DB_PITCH = TB + 1;
constexpr int SB = 32;                 // sub-block edge
constexpr int SB_PITCH = SB + 1;
constexpr int DIAG_SMEM_DOUBLES = TB * DB_PITCH + 4 * SB * SB_PITCH /* D^-1 */ + 3 * SB * SB_PITCH /* temporaries */ + 8;
constexpr int DIAG_SMEM_BYTES = DIAG_SMEM_DOUBLES * (int)sizeof(double);

// C(32x32) = alpha * A(32x32) * op(B) (+ C);  op(B)(k,n) = BT ? B[n][k] : B[k][n].  Executed by one 64-thread group; `bar`
// is the group's named barrier, used when C aliases an operand (all reads of the group happen before any write).
template <bool BT>
__device__ __forceinline__ void sub_gemm(const double* A, int pa, const double* B, int pb, double* C, int pc, double alpha,
                                         bool accumulate, bool in_place, int gt, int bar) {
    const int r0 = (gt >> 3) * 4, c0 = (gt & 7) * 4;
    double acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j] = 0.0;
#pragma unroll 4
    for (int k = 0; k < SB; k++) {
        double a[4], b[4];
#pragma unroll
        for (int i = 0; i < 4; i++) a[i] = A[(r0 + i) * pa + k];
#pragma unroll
        for (int j = 0; j < 4; j++) b[j] = BT ? B[(c0 + j) * pb + k] : B[k * pb + c0 + j];
#pragma unroll
        for (int i = 0; i < 4; i++)
#pragma unroll
            for (int j = 0; j < 4; j++) acc[i][j] = fma(a[i], b[j], acc[i][j]);
    }
    if (in_place) asm volatile("bar.sync %0, 64;" ::"r"(bar) : "memory");
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) {
            double* c = C + (r0 + i) * pc + c0 + j;
            *c = accumulate ? fma(alpha, acc[i][j], *c) : alpha * acc[i][j];
        }
}

// One warp: Cholesky of the 32x32 sub-block D (lower, in shared memory, pitch pd) in place, and its inverse -> Dinv
// (pitch SB_PITCH, explicit zeros above the diagonal).  pivot_base = global 1-based index of the block's first pivot.
// Rows / columns live in registers; the column loops are unrolled by template recursion so that every register index is
// a compile-time constant (a plain `#pragma unroll` of the 32x32 nest is refused and spills the row to local memory).
template <int J>
struct CholCol {
    static __device__ __forceinline__ void run(double (&a)[SB], int lane, int pivot_base, int* info) {
        const double djj = __shfl_sync(0xffffffffu, a[J], J);
        if (lane == 0 && !(djj > 0.0)) atomicCAS(info, 0, pivot_base + J);
        // reciprocal square root + one multiply instead of sqrt + divide on the dependent chain (<= 1 ulp apart)
        const double il = rsqrt(djj);
        const double lrj = (lane > J) ? a[J] * il : (lane == J ? djj * il : 0.0);
        a[J] = lrj;
#pragma unroll
        for (int k = J + 1; k < SB; k++) {
            const double lkj = __shfl_sync(0xffffffffu, lrj, k);
            if (lane >= k) a[k] = fma(-lrj, lkj, a[k]);
        }
        CholCol<J + 1>::run(a, lane, pivot_base, info);
    }
};
template <>
struct CholCol<SB> {
    static __device__ __forceinline__ void run(double (&)[SB], int, int, int*) {}
};

template <int R>
struct InvRow {
    static __device__ __forceinline__ void run(double (&x)[SB], const double* D, int pd, int lane, double rdiag) {
        double s0 = (R == lane) ? 1.0 : 0.0, s1 = 0.0;
#pragma unroll
        for (int k = 0; k + 1 < R; k += 2) {
            s0 = fma(-D[R * pd + k], x[k], s0);
            s1 = fma(-D[R * pd + k + 1], x[k + 1], s1);
        }
        if (R & 1) s0 = fma(-D[R * pd + R - 1], x[R - 1], s0);
        const double rd = __shfl_sync(0xffffffffu, rdiag, R);  // executed by all lanes (never inside the select below)
        x[R] = (R >= lane) ? (s0 + s1) * rd : 0.0;
        InvRow<R + 1>::run(x, D, pd, lane, rdiag);
    }
};
template <>
struct InvRow<SB> {
    static __device__ __forceinline__ void run(double (&)[SB], const double*, int, int, double) {}
};

__device__ __forceinline__ void chol_inv_32(double* D, int pd, double* Dinv, int lane, int pivot_base, int* info) {
    double a[SB];
#pragma unroll
    for (int k = 0; k < SB; k++) a[k] = (k <= lane) ? D[lane * pd + k] : 0.0;
    CholCol<0>::run(a, lane, pivot_base, info);
#pragma unroll
    for (int k = 0; k < SB; k++)
        if (k <= lane) D[lane * pd + k] = a[k];
    __syncwarp();
    // inverse: lane c owns column c of X = D^-1;  X[r][c] = (delta_rc - sum_{k<r} D[r][k] X[k][c]) / D[r][r]
    double x[SB];
    const double rdiag = 1.0 / D[lane * pd + lane];  // all 32 reciprocals at once, broadcast by shuffle when needed
    InvRow<0>::run(x, D, pd, lane, rdiag);
#pragma unroll
    for (int r = 0; r < SB; r++) Dinv[r * SB_PITCH + lane] = x[r];
}

__global__ void __launch_bounds__(256) diag_factor_inverse_kernel(double* __restrict__ K, double* __restrict__ Linv,
                                                                  double* __restrict__ LinvT, int Np, int p, int N,
                                                                  double* __restrict__ logdet, int* __restrict__ info) {
    extern __shared__ double sm[];
    double* S = sm;                                  // [TB][DB_PITCH]
    double* Dinv = sm + TB * DB_PITCH;               // [4][SB][SB_PITCH] inverses of the diagonal sub-blocks
    double* Tmp = Dinv + 4 * SB * SB_PITCH;          // [3][SB][SB_PITCH]
    double* red = Tmp + 3 * SB * SB_PITCH;           // [8]
    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int grp = tid >> 6, gt = tid & 63;         // 64-thread group and index inside it
    const int bar = 1 + grp;
    const size_t base = (size_t)p * TB * Np + (size_t)p * TB;
#pragma unroll 16
    for (int e = tid; e < TB * TB; e += 256) {
        int r = e >> 7, c = e & 127;
        S[r * DB_PITCH + c] = K[base + (size_t)r * Np + c];
    }
    __syncthreads();
#define SBLK(i, k) (S + (i) * SB * DB_PITCH + (k) * SB)
    // ---- blocked right-looking Cholesky over the 4 sub-block columns ----
    for (int pb = 0; pb < 4; pb++) {
        if (warp == 0) chol_inv_32(SBLK(pb, pb), DB_PITCH, Dinv + pb * SB * SB_PITCH, lane, p * TB + pb * SB + 1, info);
        __syncthreads();
        // panel: L_ip = A_ip * D^-T, in place
        if (pb + 1 + grp < 4)
            sub_gemm<true>(SBLK(pb + 1 + grp, pb), DB_PITCH, Dinv + pb * SB * SB_PITCH, SB_PITCH, SBLK(pb + 1 + grp, pb), DB_PITCH, 1.0,
                           false, true, gt, bar);
        __syncthreads();
        // trailing: A_ik -= L_ip L_kp^T for pb < k <= i
        int q = 0;
        for (int i = pb + 1; i < 4; i++)
            for (int k = pb + 1; k <= i; k++, q++)
                if ((q & 3) == grp)
                    sub_gemm<true>(SBLK(i, pb), DB_PITCH, SBLK(k, pb), DB_PITCH, SBLK(i, k), DB_PITCH, -1.0, true, false, gt, bar);
        __syncthreads();
    }
    // log-determinant contribution (fixed order) and write L back
    {
        double v = 0.0;
        if (tid < TB && p * TB + tid < N) v = log(S[tid * DB_PITCH + tid]);
        double s = block_sum<256>(v, red);
        if (tid == 0) logdet[p] = s;
    }
    for (int e = tid; e < TB * TB; e += 256) {
        int r = e >> 7, c = e & 127;
        if (c <= r) K[base + (size_t)r * Np + c] = S[r * DB_PITCH + c];
    }
    __syncthreads();
    // ---- in-place block inverse, block columns from the right; diagonal blocks live in Dinv ----
    for (int j = 2; j >= 0; j--) {
        // T_i = sum_{k=j+1..i} X_ik L_kj   (X_ii = Dinv[i], X_ik = S block (i,k) already inverted)
        const int i = j + 1 + grp;
        if (i < 4) {
            double* T = Tmp + grp * SB * SB_PITCH;
            for (int k = j + 1; k <= i; k++) {
                const double* X = (k == i) ? Dinv + i * SB * SB_PITCH : SBLK(i, k);
                const int px = (k == i) ? SB_PITCH : DB_PITCH;
                sub_gemm<false>(X, px, SBLK(k, j), DB_PITCH, T, SB_PITCH, 1.0, k > j + 1, false, gt, bar);
            }
        }
        __syncthreads();
        // X_ij = -T_i X_jj
        if (i < 4)
            sub_gemm<false>(Tmp + grp * SB * SB_PITCH, SB_PITCH, Dinv + j * SB * SB_PITCH, SB_PITCH, SBLK(i, j), DB_PITCH, -1.0, false,
                            false, gt, bar);
        __syncthreads();
    }
#undef SBLK
    for (int e = tid; e < TB * TB; e += 256) {
        int r = e >> 7, c = e & 127;
        double v = 0.0;
        if (c <= r) v = ((r >> 5) == (c >> 5)) ? Dinv[(r >> 5) * SB * SB_PITCH + (r & 31) * SB_PITCH + (c & 31)] : S[r * DB_PITCH + c];
        Linv[base + (size_t)r * Np + c] = v;
    }
    for (int e = tid; e < TB * TB; e += 256) {
        int r = e >> 7, c = e & 127;  // LinvT[r][c] = Linv[c][r]
        double v = 0.0;
        if (r <= c) v = ((r >> 5) == (c >> 5)) ? Dinv[(r >> 5) * SB * SB_PITCH + (c & 31) * SB_PITCH + (r & 31)] : S[c * DB_PITCH + r];
        LinvT[base + (size_t)r * Np + c] = v;
    }
}

// ---- tile GEMM with store epilogues -------------------------------------------------------------------------------
enum GemmMode {
    MODE_CHOL_PANEL = 0,  // L[i,p] = A[i,p] * Linv_pp^T                       tiles: i = p+1 .. nb-1
    MODE_CHOL_TRAIL = 1,  // A[i,j] -= L[i,p] * L[j,p]^T                        tiles: p < j <= i
    MODE_TRTRI_XT = 2,    // T[u-blk, v-blk] = LinvT11 * L21^T  (= X^T)        per pair of half-blocks of size s
    MODE_TRTRI_Y = 3,     // Linv21 = -(Linv22 * X), LinvT12 = Linv21^T
    MODE_LAUUM = 4        // Kinv[i,j] = sum_{k >= i} LinvT[i][k] LinvT[j][k]   tiles: j <= i
};

struct DenseParams {
    double* K;
    double* Linv;
    double* LinvT;
    double* T;
    double* Kinv;
    int Np, nb;
    int p;  // Cholesky panel
    int s;  // half-block size (in tiles) of the inverse recursion level
};

__device__ __forceinline__ void tri_index(int t, int& i, int& j) {  // t -> (i >= j), row-major lower enumeration
    i = (int)((sqrt(8.0 * t + 1.0) - 1.0) * 0.5);
    while ((i + 1) * (i + 2) / 2 <= t) i++;
    while (i * (i + 1) / 2 > t) i--;
    j = t - i * (i + 1) / 2;
}

// number of tiles of a recursion level: pairs q = 0.. ; pair q covers tile rows [2qs, 2qs+2s) clipped to nb
__host__ __device__ inline int trtri_pair_vtiles(int nb, int s, int q) {
    int a = 2 * q * s;
    int rem = nb - (a + s);
    return rem <= 0 ? 0 : (rem < s ? rem : s);
}

template <int MODE>
__global__ void __launch_bounds__(GTHREADS, 1) dense_gemm_kernel(DenseParams P) {
    extern __shared__ double smem[];
    const int Np = P.Np;
    TileOperands w;
    w.lda = w.ldb = Np;
    w.tri_off = TRI_DENSE;
    int orow = 0, ocol = 0;  // output tile origin (rows, cols)
    if (MODE == MODE_CHOL_PANEL) {
        int i = P.p + 1 + blockIdx.x;
        w.A = P.K + (size_t)i * TB * Np + (size_t)P.p * TB;
        w.B = P.Linv + (size_t)P.p * TB * Np + (size_t)P.p * TB;
        w.kbeg = 0;
        w.kend = TB;
        orow = i * TB;
        ocol = P.p * TB;
    } else if (MODE == MODE_CHOL_TRAIL) {
        int ti, tj;
        tri_index(blockIdx.x, ti, tj);
        int i = P.p + 1 + ti, j = P.p + 1 + tj;
        w.A = P.K + (size_t)i * TB * Np + (size_t)P.p * TB;
        w.B = P.K + (size_t)j * TB * Np + (size_t)P.p * TB;
        w.kbeg = 0;
        w.kend = TB;
        orow = i * TB;
        ocol = j * TB;
    } else if (MODE == MODE_TRTRI_XT || MODE == MODE_TRTRI_Y) {
        // decode (pair q, u tile, v tile): tiles are enumerated pair by pair, u-major
        int s = P.s, t = blockIdx.x, q = 0;
        for (;; q++) {
            int cnt = s * trtri_pair_vtiles(P.nb, s, q);
            if (t < cnt) break;
            t -= cnt;
        }
        int nv = trtri_pair_vtiles(P.nb, s, q);
        int ut = t / nv, vt = t % nv;
        int a = 2 * q * s;  // first tile of the pair
        if (MODE == MODE_TRTRI_XT) {
            // X^T[u][v] = sum_k LinvT11[u][k] * L21[v][k],  k in [u-tile start, s*TB) (LinvT11 is upper triangular)
            w.A = P.LinvT + (size_t)(a + ut) * TB * Np + (size_t)a * TB;
            w.B = P.K + (size_t)(a + s + vt) * TB * Np + (size_t)a * TB;
            w.kbeg = ut * TB;
            w.kend = s * TB;
            orow = (a + ut) * TB;
            ocol = (a + s + vt) * TB;
        } else {
            // Y[v][u] = -sum_k Linv22[v][k] * X^T[u][k],  k in [0, (vt+1)*TB) (Linv22 is lower triangular)
            w.A = P.Linv + (size_t)(a + s + vt) * TB * Np + (size_t)(a + s) * TB;
            w.B = P.T + (size_t)(a + ut) * TB * Np + (size_t)(a + s) * TB;
            w.kbeg = 0;
            w.kend = (vt + 1) * TB;
            w.tri_off = vt * TB;
            orow = (a + s + vt) * TB;
            ocol = (a + ut) * TB;
        }
    } else {  // MODE_LAUUM
        int i, j;
        tri_index(blockIdx.x, i, j);
        w.A = P.LinvT + (size_t)i * TB * Np;
        w.B = P.LinvT + (size_t)j * TB * Np;
        w.kbeg = i * TB;
        w.kend = Np;
        orow = i * TB;
        ocol = j * TB;
    }

    TileAcc acc;
    gemm_tile_mainloop(w, acc, smem);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int wm = warp >> 2, wn = warp & 3, g = lane >> 2, t = lane & 3;
#pragma unroll
    for (int i = 0; i < 8; i++) {
#pragma unroll
        for (int j = 0; j < 4; j++) {
            int r = orow + wm * 64 + i * 8 + g;
            int c = ocol + wn * 32 + j * 8 + 2 * t;
            double v0 = acc.v[i][j][0], v1 = acc.v[i][j][1];
            if (MODE == MODE_CHOL_PANEL) {
                double2* dst = reinterpret_cast<double2*>(P.K + (size_t)r * Np + c);
                *dst = make_double2(v0, v1);
            } else if (MODE == MODE_CHOL_TRAIL) {
                double2* dst = reinterpret_cast<double2*>(P.K + (size_t)r * Np + c);
                double2 old = *dst;
                *dst = make_double2(old.x - v0, old.y - v1);
            } else if (MODE == MODE_TRTRI_XT) {
                double2* dst = reinterpret_cast<double2*>(P.T + (size_t)r * Np + c);
                *dst = make_double2(v0, v1);
            } else if (MODE == MODE_TRTRI_Y) {
                double2* dst = reinterpret_cast<double2*>(P.Linv + (size_t)r * Np + c);
                *dst = make_double2(-v0, -v1);
                P.LinvT[(size_t)c * Np + r] = -v0;
                P.LinvT[(size_t)(c + 1) * Np + r] = -v1;
            } else {
                double2* dst = reinterpret_cast<double2*>(P.Kinv + (size_t)r * Np + c);
                *dst = make_double2(v0, v1);
            }
        }
    }
}

// ---- triangular matrix-vector products: one warp per row, fixed summation order ----------------------------------
// out[i] = sum_{k in [kbeg(i), kend(i))} M[i][k] * x[k],  LOWER: k <= i ; UPPER: k >= i
template <bool LOWER>
__global__ void __launch_bounds__(256) tri_matvec_kernel(const double* __restrict__ M, const double* __restrict__ x,
                                                         int Np, double* __restrict__ out) {
    int row = blockIdx.x * 8 + (threadIdx.x >> 5);
    int lane = threadIdx.x & 31;
    if (row >= Np) return;
    int k0 = LOWER ? 0 : (row & ~31);
    int k1 = LOWER ? row + 1 : Np;
    const double* mr = M + (size_t)row * Np;
    double s = 0.0;
    for (int k = k0 + lane; k < k1; k += 32) {
        if (LOWER || k >= row) s = fma(mr[k], x[k], s);
    }
    s = warp_sum(s);
    if (lane == 0) out[row] = s;
}

// resid[j] = y[j] - c (j < N), 0 for padding
__global__ void residual_kernel(const double* __restrict__ y, double c, int N, int Np, double* __restrict__ resid) {
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < Np) resid[j] = (j < N) ? y[j] - c : 0.0;
}

// scalars[0] = a^T a, scalars[1] = sum_p logdet[p], scalars[2] = sum_j alpha[j]      (single block)
__global__ void __launch_bounds__(256) lml_scalars_kernel(const double* __restrict__ a, const double* __restrict__ alpha,
                                                          const double* __restrict__ logdet, int N, int nb,
                                                          double* __restrict__ scalars) {
    __shared__ double red[8];
    double s0 = 0.0, s2 = 0.0, s1 = 0.0;
    for (int j = threadIdx.x; j < N; j += 256) {
        s0 = fma(a[j], a[j], s0);
        s2 += alpha[j];
    }
    for (int p = threadIdx.x; p < nb; p += 256) s1 += logdet[p];
    double v0 = block_sum<256>(s0, red);
    double v1 = block_sum<256>(s1, red);
    double v2 = block_sum<256>(s2, red);
    if (threadIdx.x == 0) {
        scalars[0] = v0;
        scalars[1] = v1;
        scalars[2] = v2;
    }
}

}  // namespace gpso
