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
// In:  K[p-block] (lower triangle used).  Out: L_pp -> K (lower; the entries above the diagonal of the four diagonal
// sub-blocks receive unused values), L_pp^-1 -> Linv (lower; the zero sub-blocks above the diagonal are never written, the
// buffer is cleared when it is (re)shaped), logdet[p] = sum_j log(L_jj) over the real (un-padded) rows, info = first
// non-positive pivot.  (L_pp^-1)^T -> LinvT is produced for all panels at once by diag_transpose_kernel.
//
// This kernel is the serial spine of the factorisation (one launch per 128-column panel), so it is organised around its
// dependent chain.  The block is a 4x4 grid of 32x32 sub-blocks in shared memory (row pitch 132: every DMMA fragment load
// is bank-conflict free).  Per sub-block column pb:
//   P1  warp 0 factorises the 32x32 diagonal sub-block AND inverts it in the same 32 column steps (rows of A and the
//       running sums of the inverse live in registers; each column is broadcast through shared memory -- no shuffles, one
//       __syncwarp per column; the inverse chain fills the latency bubbles of the Cholesky chain).  Meanwhile warps 1-7
//       accumulate T_pb,j = sum_k L_pb,k X_kj, the row pb of the block inverse, from finished blocks (DMMA).
//   P2  panel  L_i,pb = A_i,pb X_pb,pb^T (i > pb)  and  X_pb,j = -X_pb,pb T_pb,j (j < pb), 8x32 DMMA strips over all warps
//   P3  trailing update  A_ik -= L_i,pb L_k,pb^T  (pb < k <= i), DMMA strips over all warps
// Measured on B200 (tools/microbench/diag_probe.cu): the previous shuffle-based version spent 28-38k cycles in each of the
// four warp-level factorisations (every __shfl_sync compiled to WARPSYNC + SHFL + ENDCOLLECTIVE) and 113 us in total.
#ifndef DIAG_STAMP
#define DIAG_STAMP(i)
#define DIAG_STAMP_W(w, i)
#endif
constexpr int DIAG_THREADS = 256;
constexpr int DP = TB + 4;             // row pitch of the 128x128 block (== 4 mod 16)
constexpr int SB = 32;                 // sub-block edge
constexpr int SP = SB + 4;             // row pitch of stand-alone 32x32 tiles (== 4 mod 16)
// the inverse is kept as a "staircase": sub-block row k holds k+1 sub-blocks with row pitch XP(k) (== 4 mod 16)
__host__ __device__ constexpr int XP(int k) { return SB * (k + 1) + 4; }
__host__ __device__ constexpr int XO(int k) { return SB * (16 * k * (k + 1) + 4 * k); }
constexpr int DIAG_SMEM_DOUBLES = TB * DP + XO(4) /* inverse */ + 6 * SB /* column + diagonal broadcast, x3 */;
constexpr int DIAG_SMEM_BYTES = DIAG_SMEM_DOUBLES * (int)sizeof(double);

// 1/sqrt(d) and 1/d together: hardware seed y (MUFU.RSQ64H, ~2^-22) and e = 1 - d y^2, then
//   1/sqrt(d) = y (1 + e/2 + 3e^2/8),   1/d = y^2 (1 + e + e^2)        (neglected terms ~ e^3 < 2^-64)
// Four dependent FP64 operations after the seed for either result.  d <= 0 / NaN -> inf / NaN.
__device__ __forceinline__ double rsqrt_seed(double d) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
    return y;
}
__device__ __forceinline__ void rsqrt_rcp_finish(double d, double y, double& il, double& ild) {
    const double t = d * y;
    const double y2 = y * y;
    const double e = fma(-t, y, 1.0);
    const double q = y * e;
    const double pl = fma(0.375, e, 0.5);
    const double ee = fma(e, e, e);
    il = fma(q, pl, y);
    ild = fma(y2, ee, y2);
}
__device__ __forceinline__ void rsqrt_rcp_fast(double d, double& il, double& ild) { rsqrt_rcp_finish(d, rsqrt_seed(d), il, ild); }
__device__ __forceinline__ double rsqrt_fast(double d) {
    double il, ild;
    rsqrt_rcp_fast(d, il, ild);
    return il;
}

// shared-memory accesses of the column broadcast: volatile asm keeps them ordered around the warp barrier while the
// compiler stays free to schedule the register arithmetic of one column into the latency gaps of the next
__device__ __forceinline__ double lds_f64(unsigned addr) {
    double v;
    asm volatile("ld.volatile.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void lds_v2f64(unsigned addr, double& x, double& y) {
    asm volatile("ld.volatile.shared.v2.f64 {%0,%1}, [%2];" : "=d"(x), "=d"(y) : "r"(addr));
}
__device__ __forceinline__ void sts_f64(unsigned addr, double v) { asm volatile("st.volatile.shared.f64 [%0], %1;" ::"r"(addr), "d"(v)); }
__device__ __forceinline__ void warp_bar() { asm volatile("bar.warp.sync 0xffffffff;"); }

// One warp, 32 column steps (compile-time unrolled by template recursion so that every register index is a constant and
// the whole factorisation is one basic block).
//   a[k] = row `lane` of the sub-block (lower part), becomes row `lane` of L;  dg = the lane's own diagonal entry
//   s[k] = sum_{j<k} L[k][j] X[j][lane]  (lane owns column `lane` of X = L^-1)
// Column J of the partially reduced block (cb) and the partially reduced diagonal (dgb) are published in shared memory
// (three buffers in rotation).  Dependent chain per column:  d_J -> (1/sqrt(d_J), 1/d_J) -> d_J+1 = dgb[J+1] - cb[J+1]^2 / d_J,
// i.e. one seed + five FP64 operations; the shared-memory hop (a[J+1] update -> publish -> read back) runs beside it.
template <int J>
struct CholInvCol {
    static __device__ __forceinline__ void run(double (&a)[SB], double (&s)[SB], double& dg, double d, double y, unsigned cb_addr,
                                               double* D, double* X, int xp, int lane, int& bad, double& my_il) {
        // three buffers in rotation: column J still reads its buffer after the barrier below (the tail of its update loop), and
        // the next stores into that buffer come from column J+2, i.e. behind column J+1's barrier
        const unsigned cb = cb_addr + (J % 3) * 2 * SB * 8;          // this column:   cb[0..31], dgb[32..63]
        const unsigned cbn = cb_addr + ((J + 1) % 3) * 2 * SB * 8;   // next column
        bad = (bad == 0 && !(d > 0.0)) ? J + 1 : bad;
        double il, ild;
        rsqrt_rcp_finish(d, y, il, ild);
        const double w = a[J] * ild;  // a_rJ / d
        double d_next = 0.0, cn = 0.0, y_next = 0.0;
        if (J + 1 < SB) {
            cn = lds_f64(cb + (J + 1) * 8);
            const double dn = lds_f64(cb + (SB + J + 1) * 8);
            d_next = fma(-(cn * cn), ild, dn);
            y_next = rsqrt_seed(d_next);
            a[J + 1] = fma(-w, cn, a[J + 1]);
            dg = fma(-w, a[J], dg);
            sts_f64(cbn + lane * 8, a[J + 1]);
            sts_f64(cbn + (SB + lane) * 8, dg);
            warp_bar();
        }
        const double xj = ((lane == J ? 1.0 : 0.0) - s[J]) * il;
        const double xs = xj * il;
        if (lane == J) my_il = il;
        X[J * xp + lane] = (lane <= J) ? xj : 0.0;
        if (J + 1 < SB) s[J + 1] = fma(cn, xs, s[J + 1]);
        constexpr int K0 = J + 2;
        if (K0 < SB && (K0 & 1)) {
            const double c = lds_f64(cb + K0 * 8);
            a[K0] = fma(-w, c, a[K0]);
            s[K0] = fma(c, xs, s[K0]);
        }
#pragma unroll
        for (int k = K0 + (K0 & 1); k + 1 < SB; k += 2) {
            double c0, c1;
            lds_v2f64(cb + k * 8, c0, c1);
            a[k] = fma(-w, c0, a[k]);
            s[k] = fma(c0, xs, s[k]);
            a[k + 1] = fma(-w, c1, a[k + 1]);
            s[k + 1] = fma(c1, xs, s[k + 1]);
        }
        D[lane * DP + J] = a[J] * il;  // L[lane][J] (rows above the diagonal receive unused values)
        CholInvCol<J + 1>::run(a, s, dg, d_next, y_next, cb_addr, D, X, xp, lane, bad, my_il);
    }
};
template <>
struct CholInvCol<SB> {
    static __device__ __forceinline__ void run(double (&)[SB], double (&)[SB], double&, double, double, unsigned, double*, double*, int, int,
                                               int&, double&) {}
};

// D: 32x32 sub-block in the big array (pitch DP), overwritten by L (lower); X: its inverse (pitch xp, zeros above);
// colbuf: 6*SB doubles, 16-byte aligned.  Returns 1/L_jj of row `lane`.
__device__ __noinline__ double chol_inv_32(double* D, double* X, int xp, double* colbuf, int lane, int pivot_base, int* info) {
    double a[SB], s[SB];
    DIAG_STAMP(20);
#pragma unroll
    for (int k = 0; k < SB; k += 2) {
        const double2 v = *reinterpret_cast<const double2*>(D + lane * DP + k);
        a[k] = (k <= lane) ? v.x : 0.0;
        a[k + 1] = (k + 1 <= lane) ? v.y : 0.0;
        s[k] = s[k + 1] = 0.0;
    }
    double dg = D[lane * DP + lane];
    const unsigned cb_addr = (unsigned)__cvta_generic_to_shared(colbuf);
    sts_f64(cb_addr + lane * 8, a[0]);
    sts_f64(cb_addr + (SB + lane) * 8, dg);
    warp_bar();
    const double d0 = lds_f64(cb_addr);
    double my_il = 1.0;
    int bad = 0;
    DIAG_STAMP(21);
    CholInvCol<0>::run(a, s, dg, d0, rsqrt_seed(d0), cb_addr, D, X, xp, lane, bad, my_il);
    DIAG_STAMP(22);
    if (bad != 0 && lane == 0) atomicCAS(info, 0, pivot_base + bad - 1);
    __syncwarp();
    DIAG_STAMP(23);
    return my_il;
}

// 8x16 half-strip of a 32x32x32 sub-block product on DMMA:  acc[j][e] += sum_k A[g][k] * op(B)[k][col0 + 8j + 2t + e]
//   A: first row of the strip (row pitch pa, K-contiguous);  BT: op(B)[k][n] = B[n][k]  else B[k][n]
// 16 DMMA per call: at the 16 cycles an SMSP needs per DMMA (tools/microbench/lat_probe.cu) a job is 256 cycles of pipe time.
template <bool BT>
__device__ __forceinline__ void half_mma(const double* A, int pa, const double* B, int pb, int col0, double (&acc)[2][2], int lane) {
    const int g = lane >> 2, t = lane & 3;
#pragma unroll
    for (int kk = 0; kk < SB / 4; kk++) {
        const double af = A[g * pa + kk * 4 + t];
#pragma unroll
        for (int j = 0; j < 2; j++) {
            const double bf = BT ? B[(col0 + j * 8 + g) * pb + kk * 4 + t] : B[(kk * 4 + t) * pb + col0 + j * 8 + g];
            dmma884(acc[j][0], acc[j][1], af, bf);
        }
    }
}

// C[g][col0 + 8j + 2t + e] = sign * acc (+ C), C = first row of the strip
__device__ __forceinline__ void half_store(double* C, int pc, int col0, const double (&acc)[2][2], double sign, bool accumulate, int lane) {
    const int g = lane >> 2, t = lane & 3;
#pragma unroll
    for (int j = 0; j < 2; j++) {
        double2* dst = reinterpret_cast<double2*>(C + g * pc + col0 + j * 8 + 2 * t);
        double2 v = make_double2(sign * acc[j][0], sign * acc[j][1]);
        if (accumulate) {
            const double2 old = *dst;
            v.x += old.x;
            v.y += old.y;
        }
        *dst = v;
    }
}

// lower-triangle enumeration q -> (ii >= kk)
__device__ __forceinline__ void tri_small(int q, int& ii, int& kk) {
    ii = 0;
    while ((ii + 1) * (ii + 2) / 2 <= q) ii++;
    kk = q - ii * (ii + 1) / 2;
}

// ---- outputs by bulk async copies (shared -> global through the copy engine; a single SM sustains only ~12-16 B/clk
// with ordinary STG, measured in tools/microbench/diag_probe.cu, which made the stores the longest phase of this kernel).
__device__ __forceinline__ void bulk_store(double* gdst, const double* ssrc, unsigned bytes) {
    const unsigned src = (unsigned)__cvta_generic_to_shared(ssrc);
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// (L_pp^-1)^T for every diagonal block, after the factorisation loop (grid = nb): LinvT_pp[r][c] = Linv_pp[c][r]
// (L_pp^-1)^T -> LinvT: the lower 32x32 sub-blocks of the diagonal tile go through shared memory in ONE round (all global loads
// in flight together, one barrier, coalesced stores).  Sub-block by sub-block with two barriers each, this was 24 us of pure
// latency -- a quarter of an N = 100 evaluation and the tail of every small factorisation.
constexpr int TRP = TB + 1;                                            // odd pitch: conflict-free transposed reads
constexpr int DIAG_TRANSPOSE_SMEM = TB * TRP * (int)sizeof(double);    // 132 096 B
__device__ __forceinline__ void diag_transpose_device(const double* __restrict__ Linv, double* __restrict__ LinvT, int Np, int p,
                                                      double* tile /* TB x TRP doubles of shared memory */) {
    const size_t base = (size_t)p * TB * Np + (size_t)p * TB;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    __syncthreads();  // previous user of the shared memory is done
#pragma unroll
    for (int bi = 0; bi < 4; bi++)
#pragma unroll
        for (int bj = 0; bj <= bi; bj++)
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const int r = bi * 32 + ty + 8 * k, c = bj * 32 + tx;
                tile[r * TRP + c] = __ldcg(Linv + base + (size_t)r * Np + c);  // written by another SM: through L2
            }
    __syncthreads();
#pragma unroll
    for (int bi = 0; bi < 4; bi++)
#pragma unroll
        for (int bj = 0; bj <= bi; bj++)
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const int c = bj * 32 + ty + 8 * k, r = bi * 32 + tx;
                LinvT[base + (size_t)c * Np + r] = tile[r * TRP + c];
            }
}
__global__ void __launch_bounds__(256) diag_transpose_kernel(const double* __restrict__ Linv, double* __restrict__ LinvT, int Np) {
    extern __shared__ __align__(16) double transpose_tile[];
    diag_transpose_device(Linv, LinvT, Np, blockIdx.x, transpose_tile);
}

// Lprev != nullptr: the tile still lacks its last update; it is applied here, in shared memory, before the factorisation:
// A_pp -= L_p,q L_p,q^T over the nprev panels q whose tiles start at Lprev (row pitch Np; nprev = 1: last narrow update,
// nprev = W: the wide update of the previous block), streamed through the not yet used inverse area in halves of 64
// columns.  ~10 us per panel here instead of a separate 30-85 us task plus a hand-over on the critical chain.
constexpr int LHP = 64 + 4;  // row pitch of a 128x64 half of the left neighbour tile
__device__ __forceinline__ void diag_block_device(double* __restrict__ K, double* __restrict__ Linv, int Np, int p, int N,
                                                  double* __restrict__ logdet, int* __restrict__ info, double* sm,
                                                  const double* __restrict__ Lprev = nullptr, int nprev = 0, int pivot_off = 0) {
    double* S = sm;                  // [TB][DP]: lower sub-blocks A -> L in place; upper sub-blocks (0,1..3) = T_pb,j scratch
    double* Xs = sm + TB * DP;       // staircase inverse: sub-block (k,j), j <= k, at Xs + XO(k) + j*SB, row pitch XP(k)
    double* colbuf = Xs + XO(4);     // [3][2][SB]
    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const size_t base = (size_t)p * TB * Np + (size_t)p * TB;
    DIAG_STAMP(0);
    {
        const int c2 = (tid & 63) * 2;
#pragma unroll
        for (int r = tid >> 6; r < TB; r += DIAG_THREADS / 64)
            if (c2 < ((r >> 5) + 1) * SB) cp_async16(S + r * DP + c2, K + base + (size_t)r * Np + c2);
        cp_async_commit();
        cp_async_wait<0>();
    }
    __syncthreads();
#define SBLK(i, k) (S + (i) * SB * DP + (k) * SB)
    if (Lprev != nullptr) {
        double* Lh = Xs;  // [TB][LHP]
        for (int half = 0; half < 2 * nprev; half++) {
            const int c2 = (tid & 31) * 2;
#pragma unroll
            for (int r = tid >> 5; r < TB; r += DIAG_THREADS / 32) cp_async16(Lh + r * LHP + c2, Lprev + (size_t)r * Np + half * 64 + c2);
            cp_async_commit();
            cp_async_wait<0>();
            __syncthreads();
            for (int job = warp; job < 80; job += 8) {
                int ii, kk;
                tri_small(job >> 3, ii, kk);
                const int strip = (job >> 1) & 3, col0 = (job & 1) * 16;
                double acc[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
                half_mma<true>(Lh + (ii * SB + strip * 8) * LHP, LHP, Lh + kk * SB * LHP, LHP, col0, acc, lane);
                half_mma<true>(Lh + (ii * SB + strip * 8) * LHP + SB, LHP, Lh + kk * SB * LHP + SB, LHP, col0, acc, lane);
                half_store(SBLK(ii, kk) + strip * 8 * DP, DP, col0, acc, -1.0, true, lane);
            }
            __syncthreads();
        }
    }
    DIAG_STAMP(1);
#define TBLK(j) SBLK(0, 1 + (j))
#define XBLK(k, j) (Xs + XO(k) + (j) * SB)
    double logacc = 0.0;
    for (int pb = 0; pb < 4; pb++) {
        // ---- P1: warp 0 factors + inverts the diagonal sub-block; the other warps work in its shadow
        if (warp == 0) {
            const double il = chol_inv_32(SBLK(pb, pb), XBLK(pb, pb), XP(pb), colbuf, lane, pivot_off + p * TB + pb * SB + 1, info);
            if (p * TB + pb * SB + lane < N) logacc -= log(il);
        } else if (pb > 0) {
            if (warp == 4) {
                // shares its SM sub-partition with warp 0: no FP64 work here, only the copies of the finished row pb-1
                // (one bulk copy per matrix row: rows of Linv from the staircase, rows of L from S)
                const int rb = pb - 1;
                fence_async_smem();
                const size_t grow = base + (size_t)(rb * SB + lane) * Np;
                bulk_store(Linv + grow, Xs + XO(rb) + lane * XP(rb), (rb + 1) * SB * 8);
                bulk_store(K + grow, S + (rb * SB + lane) * DP, (rb + 1) * SB * 8);
                bulk_commit();
            } else {
                const int sw = warp < 4 ? warp - 1 : warp - 2;  // 0..5
                // (a) deferred trailing update of step pb-1: A_ik -= L_i,pb-1 L_k,pb-1^T for pb+1 <= k <= i
                const int m = 3 - pb;
                const int n_def = m * (m + 1) / 2 * 8;
                // (b) T_pb,j = sum_{k=j}^{pb-1} L_pb,k X_kj  (j < pb)
                const int n_t = pb * 8;
                for (int job = sw; job < n_def + n_t; job += 6) {
                    double acc[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
                    if (job < n_def) {
                        int ii, kk;
                        tri_small(job >> 3, ii, kk);
                        const int i = pb + 1 + ii, k = pb + 1 + kk, strip = (job >> 1) & 3, col0 = (job & 1) * 16;
                        half_mma<true>(SBLK(i, pb - 1) + strip * 8 * DP, DP, SBLK(k, pb - 1), DP, col0, acc, lane);
                        half_store(SBLK(i, k) + strip * 8 * DP, DP, col0, acc, -1.0, true, lane);
                    } else {
                        const int q = job - n_def;
                        const int j = q >> 3, strip = (q >> 1) & 3, col0 = (q & 1) * 16;
                        for (int k = j; k < pb; k++)
                            half_mma<false>(SBLK(pb, k) + strip * 8 * DP, DP, XBLK(k, j), XP(k), col0, acc, lane);
                        half_store(TBLK(j) + strip * 8 * DP, DP, col0, acc, 1.0, false, lane);
                    }
                }
            }
        }
        fence_async_smem();  // by the writers: the rows written here are read by bulk copies issued after the barrier
        __syncthreads();
        DIAG_STAMP(2 + 3 * pb);
        // ---- P2: panel L_i,pb = A_i,pb X_pb,pb^T (i > pb, in place) and X_pb,j = -X_pb,pb T_pb,j (j < pb).
        // Round = sub-block; warps (2s, 2s+1) share strip s and take one half each.  A panel half reads the whole strip it
        // partly overwrites, so the two warps of a pair meet at a named barrier between their loads and their stores.
        for (int blk = 0; blk < 3; blk++) {
            double acc[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
            const int strip = warp >> 1, col0 = (warp & 1) * 16;
            if (blk < 3 - pb) {
                const int i = pb + 1 + blk;
                half_mma<true>(SBLK(i, pb) + strip * 8 * DP, DP, XBLK(pb, pb), XP(pb), col0, acc, lane);
                asm volatile("bar.sync %0, 64;" ::"r"(1 + strip) : "memory");
                half_store(SBLK(i, pb) + strip * 8 * DP, DP, col0, acc, 1.0, false, lane);
            } else {
                const int j = blk - (3 - pb);
                half_mma<false>(XBLK(pb, pb) + strip * 8 * XP(pb), XP(pb), TBLK(j), DP, col0, acc, lane);
                half_store(XBLK(pb, j) + strip * 8 * XP(pb), XP(pb), col0, acc, -1.0, false, lane);
            }
        }
        fence_async_smem();
        __syncthreads();
        DIAG_STAMP(3 + 3 * pb);
        // ---- P3: the part of the trailing update the next step waits for: A_i,pb+1 -= L_i,pb L_pb+1,pb^T (i > pb)
        for (int job = warp; job < (3 - pb) * 8; job += 8) {
            double acc[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
            const int i = pb + 1 + (job >> 3), strip = (job >> 1) & 3, col0 = (job & 1) * 16;
            half_mma<true>(SBLK(i, pb) + strip * 8 * DP, DP, SBLK(pb + 1, pb), DP, col0, acc, lane);
            half_store(SBLK(i, pb + 1) + strip * 8 * DP, DP, col0, acc, -1.0, true, lane);
        }
        __syncthreads();
        DIAG_STAMP(4 + 3 * pb);
    }
    // log-determinant contribution (fixed order: the 32 rows of each sub-block in lane order, sub-blocks in sequence)
    if (warp == 0) {
        const double v = warp_sum(logacc);
        if (lane == 0) logdet[p] = v;
    }
    DIAG_STAMP(14);
    // ---- remaining outputs: row 3 of the inverse and of L, four matrix rows per warp; every bulk copy must have landed
    // before the kernel ends
    if (lane < 8) {
        fence_async_smem();
        const int rr = warp * 4 + (lane & 3);
        const size_t grow = base + (size_t)(3 * SB + rr) * Np;
        if (lane < 4) bulk_store(Linv + grow, Xs + XO(3) + rr * XP(3), 4 * SB * 8);
        else bulk_store(K + grow, S + (3 * SB + rr) * DP, 4 * SB * 8);
        bulk_commit();
    }
    DIAG_STAMP_W(4, 30);
    if (lane < 8 || warp == 4) bulk_wait_all();
    DIAG_STAMP_W(4, 31);
#undef SBLK
#undef TBLK
#undef XBLK
    DIAG_STAMP(18);
}

__global__ void __launch_bounds__(DIAG_THREADS) diag_factor_inverse_kernel(double* __restrict__ K, double* __restrict__ Linv,
                                                                           int Np, int p, int N, double* __restrict__ logdet,
                                                                           int* __restrict__ info, int pivot_off = 0) {
    extern __shared__ __align__(16) double sm[];
    diag_block_device(K, Linv, Np, p, N, logdet, info, sm, nullptr, 0, pivot_off);
}

// ---- tile GEMM with store epilogues -------------------------------------------------------------------------------
enum GemmMode {
    MODE_CHOL_PANEL = 0,  // L[i,p] = A[i,p] * Linv_pp^T                       tiles: i = p+1 .. nb-1
    MODE_CHOL_TRAIL = 1,  // A[i,j] -= L[i,p] * L[j,p]^T                        tiles: p < j <= i
    MODE_TRTRI_XT = 2,    // T[u-blk, v-blk] = LinvT11 * L21^T  (= X^T)        per pair of half-blocks of size s
    MODE_TRTRI_Y = 3,     // Linv21 = -(Linv22 * X), LinvT12 = Linv21^T
    MODE_LAUUM = 4,       // Kinv[i,j] = sum_{k >= i} LinvT[i][k] LinvT[j][k]   tiles: j <= i
    MODE_CHOL_UPD = 5,    // A[ui,uj] -= L[ui, uk0..uk0+ukn) * L[uj, same panels]^T   one tile, K = ukn * 128
    MODE_LAUUM_PART = 6   // one 128-wide k-chunk of a LAUUM tile into the partial buffer (small matrices: see lauum_reduce_kernel)
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
    int ui, uj, uk0, ukn;  // MODE_CHOL_UPD: output tile and range of panels contracted over
    int pivot_off = 0;     // row index of the sub-matrix's first row in the whole matrix (LAPACK info of a hybrid leaf)
    double* part = nullptr;  // MODE_LAUUM_PART: [chunk][128][128] partial products
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
__device__ __forceinline__ void gemm_tile_device(const DenseParams& P, const int tile, double* smem) {
    const int Np = P.Np;
    TileOperands w;
    w.lda = w.ldb = Np;
    w.tri_off = TRI_DENSE;
    int orow = 0, ocol = 0;  // output tile origin (rows, cols)
    if (MODE == MODE_CHOL_PANEL) {
        int i = P.p + 1 + tile;
        w.A = P.K + (size_t)i * TB * Np + (size_t)P.p * TB;
        w.B = P.Linv + (size_t)P.p * TB * Np + (size_t)P.p * TB;
        w.kbeg = 0;
        w.kend = TB;
        orow = i * TB;
        ocol = P.p * TB;
    } else if (MODE == MODE_CHOL_TRAIL) {
        int ti, tj;
        tri_index(tile, ti, tj);
        int i = P.p + 1 + ti, j = P.p + 1 + tj;
        w.A = P.K + (size_t)i * TB * Np + (size_t)P.p * TB;
        w.B = P.K + (size_t)j * TB * Np + (size_t)P.p * TB;
        w.kbeg = 0;
        w.kend = TB;
        orow = i * TB;
        ocol = j * TB;
    } else if (MODE == MODE_CHOL_UPD) {
        w.A = P.K + (size_t)P.ui * TB * Np + (size_t)P.uk0 * TB;
        w.B = P.K + (size_t)P.uj * TB * Np + (size_t)P.uk0 * TB;
        w.kbeg = 0;
        w.kend = P.ukn * TB;
        orow = P.ui * TB;
        ocol = P.uj * TB;
    } else if (MODE == MODE_TRTRI_XT || MODE == MODE_TRTRI_Y) {
        // decode (pair q, u tile, v tile): tiles are enumerated pair by pair, u-major
        int s = P.s, t = tile, q = 0;
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
    } else if (MODE == MODE_LAUUM_PART) {
        // chunk enumeration: tiles (i, j <= i) row-major, per tile the nb - i chunks k in [(i + c) TB, (i + c + 1) TB)
        int i = 0, t = tile;
        for (;; i++) {
            const int cnt = (i + 1) * (P.nb - i);
            if (t < cnt) break;
            t -= cnt;
        }
        const int j = t / (P.nb - i), c = t - j * (P.nb - i);
        w.A = P.LinvT + (size_t)i * TB * Np;
        w.B = P.LinvT + (size_t)j * TB * Np;
        w.kbeg = (i + c) * TB;
        w.kend = w.kbeg + TB;
    } else {  // MODE_LAUUM
        int i, j;
        tri_index(tile, i, j);
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
            } else if (MODE == MODE_CHOL_TRAIL || MODE == MODE_CHOL_UPD) {
                double2* dst = reinterpret_cast<double2*>(P.K + (size_t)r * Np + c);
                double2 old = __ldcg(dst);  // L2: another SM may have updated this tile (persistent scheduler)
                *dst = make_double2(old.x - v0, old.y - v1);
            } else if (MODE == MODE_TRTRI_XT) {
                double2* dst = reinterpret_cast<double2*>(P.T + (size_t)r * Np + c);
                *dst = make_double2(v0, v1);
            } else if (MODE == MODE_TRTRI_Y) {
                double2* dst = reinterpret_cast<double2*>(P.Linv + (size_t)r * Np + c);
                *dst = make_double2(-v0, -v1);
                P.LinvT[(size_t)c * Np + r] = -v0;
                P.LinvT[(size_t)(c + 1) * Np + r] = -v1;
            } else if (MODE == MODE_LAUUM_PART) {
                double2* dst = reinterpret_cast<double2*>(P.part + (size_t)tile * TB * TB + (size_t)(r - orow) * TB + (c - ocol));
                *dst = make_double2(v0, v1);
            } else {
                double2* dst = reinterpret_cast<double2*>(P.Kinv + (size_t)r * Np + c);
                *dst = make_double2(v0, v1);
            }
        }
    }
}

template <int MODE>
__global__ void __launch_bounds__(GTHREADS, 1) dense_gemm_kernel(DenseParams P) {
    extern __shared__ double smem[];
    gemm_tile_device<MODE>(P, blockIdx.x, smem);
}

// K_y^-1 of a matrix of two to four tiles: the longest LAUUM tile contracts over the whole matrix on ONE SM (63 us of a 300 us
// evaluation at N = 300).  dense_gemm_kernel<MODE_LAUUM_PART> gives every 128-wide k-chunk of every tile its own SM; this kernel
// adds the chunks of a tile in ascending k (fixed order: the result does not depend on scheduling).  grid = tiles (i, j <= i).
__global__ void __launch_bounds__(256) lauum_reduce_kernel(const double* __restrict__ part, int nb, int Np, double* __restrict__ Kinv) {
    int i, j;
    tri_index(blockIdx.x, i, j);
    int first = 0;
    for (int ii = 0; ii < i; ii++) first += (ii + 1) * (nb - ii);
    first += j * (nb - i);
    const int chunks = nb - i;
    for (int e = threadIdx.x * 2; e < TB * TB; e += 512) {
        double2 acc = *reinterpret_cast<const double2*>(part + (size_t)first * TB * TB + e);
        for (int c = 1; c < chunks; c++) {
            const double2 v = *reinterpret_cast<const double2*>(part + (size_t)(first + c) * TB * TB + e);
            acc.x += v.x;
            acc.y += v.y;
        }
        const int r = e >> 7, col = e & 127;
        *reinterpret_cast<double2*>(Kinv + (size_t)(i * TB + r) * Np + j * TB + col) = acc;
    }
}

// ---- the panel tile the next diagonal block waits for, in four row strips -------------------------------------------------
// L[p+1,p] = A[p+1,p] L_pp^-T sits on the critical chain DIAG(p) -> PANEL(p+1,p) -> DIAG(p+1): 25 us as one 128x128x128 tile on
// one SM.  Cut into PANEL_STRIPS strips of 32 rows, each on its own SM (whole operands in shared memory: 32 x 128 of A, the
// lower triangle of L_pp^-1; warp w owns 16 output columns and only multiplies up to its last column), the chain waits ~8 us.
constexpr int PANEL_STRIPS = 4;
constexpr int PANEL_STRIP_ROWS = TB / PANEL_STRIPS;
__device__ __forceinline__ void panel_strip_device(const DenseParams& P, int p, int strip, double* smem) {
    static_assert(PANEL_STRIP_ROWS == 32, "warp tiling below assumes 32-row strips");
    double* sA = smem;                          // [32][DP]
    double* sB = smem + PANEL_STRIP_ROWS * DP;  // [128][DP]
    const int Np = P.Np, tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    double* Atile = P.K + ((size_t)(p + 1) * TB + strip * PANEL_STRIP_ROWS) * Np + (size_t)p * TB;
    const double* Btile = P.Linv + (size_t)p * TB * Np + (size_t)p * TB;
    __syncthreads();  // previous user of the shared memory is done
    for (int c = tid; c < PANEL_STRIP_ROWS * (TB / 2); c += GTHREADS) {
        const int r = c >> 6, k2 = (c & 63) * 2;
        cp_async16(sA + r * DP + k2, Atile + (size_t)r * Np + k2);
    }
    for (int c = tid; c < TB * (TB / 2); c += GTHREADS) {
        const int r = c >> 6, k2 = (c & 63) * 2;
        if (k2 < ((r >> 4) + 1) * 16) cp_async16(sB + r * DP + k2, Btile + (size_t)r * Np + k2);  // columns a warp reads for row r
    }
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();
    // column block of this warp: the two warps of an SM sub-partition (w, w + 4) take blocks (w, 7 - w): equal DMMA counts
    const int cb = warp < 4 ? warp : 11 - warp;
    double acc[4][2][2];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 2; j++) acc[i][j][0] = acc[i][j][1] = 0.0;
    const double* pa = sA + g * DP + t;
    const double* pb = sB + (cb * 16 + g) * DP + t;
    const int nkk = (cb + 1) * 4;  // k <= last column of the block (L_pp^-1 is lower triangular)
    for (int kk = 0; kk < nkk; kk++) {
        double af[4], bf[2];
#pragma unroll
        for (int i = 0; i < 4; i++) af[i] = pa[i * 8 * DP + kk * 4];
#pragma unroll
        for (int j = 0; j < 2; j++) bf[j] = pb[j * 8 * DP + kk * 4];
#pragma unroll
        for (int i = 0; i < 4; i++)
#pragma unroll
            for (int j = 0; j < 2; j++) dmma884(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
    }
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 2; j++)
            *reinterpret_cast<double2*>(Atile + (size_t)(i * 8 + g) * Np + cb * 16 + j * 8 + 2 * t) = make_double2(acc[i][j][0], acc[i][j][1]);
}

// ---- persistent factorisation: one CTA per SM, tile tasks with explicit dependencies ----------------------------------
// One launch runs the blocked Cholesky AND the inverse factor.  The host (gpso_capi.cu: build_factor_tasks) writes every
// task with the counters it waits for and the counter it signals; the kernel is an interpreter.
//
// Cholesky.  Two-level blocking: panels are grouped in blocks of W (4 from 48 panels up, else 1 = unblocked).  A tile (i, j) in block
// column bj = j / W receives, in this order, bj "wide" updates (one per earlier block, K = W * 128: the read-modify-write of
// the tile -- 8 us at the ~14 B/clk an SM can store -- is paid once per W panels), then j - bj*W "narrow" updates from the
// earlier panels of its own block (K = 128), then its final operation (factor+invert if i == j, panel solve otherwise).
// A counter per tile counts the operations completed on it (ops(j) = bj + j - bj*W updates, final at ops(j) + 1).  The
// last update of a diagonal tile (narrow, or wide for the first panel of a block) is fused into its DIAG task.
//   DIAG(p)       waits tile (p,p) at ops(p)-1 and tile (p,p-1) final       PANEL(i,p)  waits (p,p) final, (i,p) at ops(p)
//   The tile (p+1,p) is solved in PANEL_STRIPS row strips (tasks PANEL with s = strip + 1) that count up their own counter
//   STRIP(p); whoever reads that tile waits for STRIP(p) = PANEL_STRIPS instead of the tile counter.
//   UPD(i,j,p)    waits (i,p), (j,p) final, (i,j) at bj + p - bj*W          WIDE(i,j,b) waits (i,q), (j,q) final (q = last
//                                                                            panel of b), (i,j) at b
// Inverse factor by recursive doubling (level s merges the inverses of tile ranges [a, a+s) and [a+s, a+2s)):
//   TRANSPOSE(p)  (L_pp^-1)^T -> LinvT                                      waits DIAG(p)
//   XT(s,q,u,v)   T = LinvT11 L21^T            waits all transposes, the level-s/2 merge of the first half, row a+s+v of L
//   Y(s,q,u,v)    Linv21 = -Linv22 T (+ LinvT12)  waits all XT of the pair and the inverse of the second half
// Scheduling: tickets by atomicAdd from one queue in a topological order chosen by the host (claiming the queue head with
// compare-and-swap after a readiness check was tried and serialises at one claim per ~3 us).  A CTA that draws a task
// whose inputs are not complete polls their counters.  Every dependency of a task sits earlier in the queue and a CTA
// holds one task at a time, so the CTA that owns the oldest unfinished task can always run: no deadlock as long as tasks
// are only taken by resident CTAs (grid <= #SMs).  The order threads the factorisation chain of block b+1 through the wide
// updates of block b (look-ahead).  Tiles cross SMs through L2 only (cp.async.cg / ld.cg), counters with release / acquire
// at gpu scope.
constexpr int CT_DIAG = 0, CT_PANEL = 1, CT_UPD = 2, CT_TRANSPOSE = 3, CT_XT = 4, CT_Y = 5;
constexpr int TASK_WORDS = 16;
// task words: 0 op | 1 p | 2 i | 3 j | 4 s | 5 tile | 6-8 dependency counter (-1: none) | 9-11 value it must reach |
//             12 counter to signal | 13 value to store with release (0: add 1) | 14-15 unused
__host__ __device__ inline int chol_ops(int j, int W) { return j / W + j % W; }
constexpr long long CHOL_SPIN_LIMIT = 1LL << 24;  // polls (each >= 100 ns) before a waiter gives up and reports an error

__device__ __forceinline__ int ld_acquire(const int* p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release(int* p, int v) { asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }

__device__ __forceinline__ bool chol_wait(const int* c, int need, int* err) {
    long long spins = 0;
    while (ld_acquire(c) < need) {
        __nanosleep(100);
        if (++spins > CHOL_SPIN_LIMIT || ld_acquire(err) != 0) {
            atomicExch(err, 1);
            return false;
        }
    }
    return true;
}

// state: [0] next ticket, [1] error flag, [2...] counters
__global__ void __launch_bounds__(GTHREADS, 1) factor_persistent_kernel(DenseParams P, int N, const int* __restrict__ tasks, int ntasks,
                                                                        int* __restrict__ state, double* __restrict__ logdet,
                                                                        int* __restrict__ info) {
    extern __shared__ __align__(16) double smem[];
    __shared__ int s_t[TASK_WORDS];
    __shared__ int s_task, s_ok;
    int* err = state + 1;
    int* ctr = state + 2;
    if (threadIdx.x == 0) s_ok = 1;
    for (;;) {
        if (threadIdx.x == 0) s_task = atomicAdd(state, 1);
        __syncthreads();
        const int t = s_task;
        if (t >= ntasks) break;
        if (threadIdx.x < TASK_WORDS) s_t[threadIdx.x] = tasks[(size_t)t * TASK_WORDS + threadIdx.x];
        __syncthreads();
        if (threadIdx.x == 0) {
            bool ok = true;
            for (int k = 0; k < 3 && ok; k++)
                if (s_t[6 + k] >= 0) ok = chol_wait(ctr + s_t[6 + k], s_t[9 + k], err);
            s_ok = ok;
        }
        __syncthreads();
        if (!s_ok) break;
        const int op = s_t[0], p = s_t[1], i = s_t[2], j = s_t[3], sx = s_t[4], tile = s_t[5], done_idx = s_t[12], done_val = s_t[13];
        if (op == CT_DIAG) {
            const double* Lprev = sx ? P.K + (size_t)p * TB * P.Np + (size_t)(p - sx) * TB : nullptr;
            diag_block_device(P.K, P.Linv, P.Np, p, N, logdet, info, smem, Lprev, sx, P.pivot_off);
        } else if (op == CT_PANEL && sx > 0) {
            panel_strip_device(P, p, sx - 1, smem);  // strip sx - 1 of the tile (p + 1, p)
        } else if (op == CT_PANEL) {
            DenseParams Q = P;
            Q.p = p;
            gemm_tile_device<MODE_CHOL_PANEL>(Q, i - p - 1, smem);
        } else if (op == CT_UPD) {
            DenseParams Q = P;
            Q.ui = i;
            Q.uj = j;
            Q.uk0 = p;
            Q.ukn = sx;
            gemm_tile_device<MODE_CHOL_UPD>(Q, 0, smem);
        } else if (op == CT_TRANSPOSE) {
            diag_transpose_device(P.Linv, P.LinvT, P.Np, p, smem);
        } else if (op == CT_XT) {
            DenseParams Q = P;
            Q.s = sx;
            gemm_tile_device<MODE_TRTRI_XT>(Q, tile, smem);
        } else {
            DenseParams Q = P;
            Q.s = sx;
            gemm_tile_device<MODE_TRTRI_Y>(Q, tile, smem);
        }
        __threadfence();
        __syncthreads();
        if (threadIdx.x == 0) {
            if (done_val > 0) st_release(ctr + done_idx, done_val);
            else atomicAdd(ctr + done_idx, 1);  // after the fence above: a release increment
        }
    }
    if (threadIdx.x == 0 && !s_ok) atomicExch(info, -1);  // a dependency never arrived: reported as a scheduler failure
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
