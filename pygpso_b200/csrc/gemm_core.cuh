// FP64 tile GEMM core on DMMA (mma.sync.m8n8k4.f64 -> SASS DMMA.8x8x4), shared by every dense step of the path:
// Cholesky trailing update and panel solve, the L^-1 recursion, K_y^-1 = L^-T L^-1, and the predict product
// V = L^-1 K*.  One CTA (256 threads, 8 warps as 2x4) produces one 128x128 tile
//
//        acc[m][n] = sum_{k in [kbeg,kend)} A[m][k] * B[n][k]          ("TN": both operands are K-contiguous rows)
//
// Operands stream global -> shared through a 4-stage cp.async (LDGSTS) ring of [128 x 16] slabs stored with a row
// pitch of 20 doubles, which makes every DMMA fragment load (8 rows x 4 k) hit 16 distinct 8-byte bank pairs per
// half-warp (conflict-free LDS.64).  B200 measurements (profiles/r01_fp64_probe.txt): DMMA peak 37.0 TFLOP/s and it
// shares the FP64 pipe with DFMA, so the only way to the roofline is to issue nothing but DMMA in the main loop:
// per k-step of 4 a warp issues 12 LDS.64 for 32 DMMA.
#pragma once
#include "common.cuh"

namespace gpso {

constexpr int GM = 128;                    // tile rows    (m)
constexpr int GN = 128;                    // tile columns (n)
#ifndef GPSO_GK
#define GPSO_GK 32
#define GPSO_GSTAGES 3
#endif
constexpr int GK = GPSO_GK;                // k-slab per pipeline stage
constexpr int GLD = GK + 4;                // shared row pitch in doubles (== 4 mod 16: 16 B aligned, conflict-free)
constexpr int GSTAGES = GPSO_GSTAGES;
constexpr int GTHREADS = 256;
constexpr int GSLAB = GM * GLD;            // doubles per operand per stage
constexpr int GEMM_SMEM_BYTES = GSTAGES * 2 * GSLAB * (int)sizeof(double);  // 163840

struct TileOperands {
    const double* A;  // first row of the 128-row A block, column 0 of the k axis
    const double* B;  // first row of the 128-row B block
    int lda, ldb;     // row pitches (doubles, even)
    int kbeg, kend;   // k range, multiples of GK
    // triangular skip: rows m of this tile only have non-zeros for k <= tri_off + m (tri_off = INT_MAX/2: dense).
    // Used for diagonal blocks of lower-triangular A so that no DMMA is spent on structural zeros.
    int tri_off;
};

constexpr int TRI_DENSE = 1 << 29;

// accumulators of one thread: acc[i][j][e] = C[wm*64 + i*8 + g][wn*32 + j*8 + 2t + e]
struct TileAcc {
    double v[8][4][2];
};

__device__ __forceinline__ void gemm_load_slab(double* sA, double* sB, const TileOperands& w, int k, int tid) {
#pragma unroll
    for (int i = 0; i < (GM * GK / 2) / GTHREADS; i++) {  // GK/4 x 16-byte chunks per thread per operand
        int id = tid + i * GTHREADS;
        int r = id / (GK / 2), c = (id % (GK / 2)) * 2;
        cp_async16(sA + r * GLD + c, w.A + (size_t)r * w.lda + k + c);
        cp_async16(sB + r * GLD + c, w.B + (size_t)r * w.ldb + k + c);
    }
}

// Main loop.  `smem` must hold GEMM_SMEM_BYTES.  All 256 threads must call it (contains __syncthreads).
__device__ __forceinline__ void gemm_tile_mainloop(const TileOperands& w, TileAcc& acc, double* smem) {
    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int wm = warp >> 2, wn = warp & 3;
    const int g = lane >> 2, t = lane & 3;
    double* sA = smem;
    double* sB = smem + GSTAGES * GSLAB;

#pragma unroll
    for (int i = 0; i < 8; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc.v[i][j][0] = acc.v[i][j][1] = 0.0;

    const int nk = (w.kend - w.kbeg) / GK;
    __syncthreads();  // previous user of the shared ring (earlier tile / epilogue scratch) is done
#pragma unroll
    for (int s = 0; s < GSTAGES - 1; s++) {
        if (s < nk) gemm_load_slab(sA + s * GSLAB, sB + s * GSLAB, w, w.kbeg + s * GK, tid);
        cp_async_commit();
    }
    const int a_off = (wm * 64 + g) * GLD + t;
    const int b_off = (wn * 32 + g) * GLD + t;
    // last row of the warp's rows that is non-zero for a given k: row >= k - tri_off
    for (int kt = 0; kt < nk; kt++) {
        cp_async_wait<GSTAGES - 2>();
        __syncthreads();
        {
            int nx = kt + GSTAGES - 1;
            if (nx < nk) {
                int s = nx % GSTAGES;
                gemm_load_slab(sA + s * GSLAB, sB + s * GSLAB, w, w.kbeg + nx * GK, tid);
            }
            cp_async_commit();
        }
        const int s = kt % GSTAGES;
        const double* pa = sA + s * GSLAB + a_off;
        const double* pb = sB + s * GSLAB + b_off;
        const int kglob = w.kbeg + kt * GK;
        // first 8-row group of this warp that still has non-zeros in this slab (warp-uniform)
        // rows m of group i: wm*64 + i*8 .. +7 ; non-zero iff k <= tri_off + m  -> need tri_off + m_max >= k
        int ifirst_slab = 0;
        {
            long need = (long)kglob - (long)w.tri_off - (long)(wm * 64) - 7;  // smallest i*8 with i*8 >= need
            if (need > 0) ifirst_slab = (int)((need + 7) >> 3);
        }
        if (ifirst_slab >= 8) continue;  // whole warp tile is structurally zero for this slab
#pragma unroll
        for (int kk = 0; kk < GK / 4; kk++) {
            double af[8], bf[4];
#pragma unroll
            for (int i = 0; i < 8; i++) af[i] = pa[i * 8 * GLD + kk * 4];
#pragma unroll
            for (int j = 0; j < 4; j++) bf[j] = pb[j * 8 * GLD + kk * 4];
            if (ifirst_slab == 0) {
#pragma unroll
                for (int i = 0; i < 8; i++)
#pragma unroll
                    for (int j = 0; j < 4; j++) dmma884(acc.v[i][j][0], acc.v[i][j][1], af[i], bf[j]);
            } else {
                int need = kglob + kk * 4 - w.tri_off - wm * 64 - 7;
                int ifirst = need > 0 ? ((need + 7) >> 3) : 0;
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    if (i >= ifirst) {
#pragma unroll
                        for (int j = 0; j < 4; j++) dmma884(acc.v[i][j][0], acc.v[i][j][1], af[i], bf[j]);
                    }
                }
            }
        }
    }
    cp_async_wait<0>();
}

}  // namespace gpso
