// Bit-exact batched replacement of LeafNode.grow(depth) (reference gpso/param_space.py:175-200, with
// ternary_split :257-307 and get_center_as_list :202-217).
//
// Output row r is the centre of the r-th node of the throw-away subtree in the reference's order: level 0 (the leaf
// itself), then every level's nodes ordered by parent and, per parent, children l, c, r.  The index of a node inside its
// level written in base 3 (most significant digit first) is therefore its path from the subtree root, and one thread
// replays that path with exactly the reference's floating-point operations:
//     w_j = hi_j - lo_j ; k = first arg-max_j w_j ; delta = w_k / 3 ; cut_i = lo_k + i * delta  (product and sum rounded
//     separately -- Python has no FMA) ; child i in {0,1,2} gets (cut_i, cut_{i+1}) ; centre_j = (lo_j + hi_j) / 2.
// The split dimension depends on last-ulp differences between sibling widths, so it is decided per node, not per level.
#pragma once
#include "common.cuh"

namespace gpso {

constexpr int LEAF_MAXD = 64;

// Rows [row0, row0 + nrows) of the batch go to out[0 .. nrows) (a rank of a candidate-sharded run generates only its shard).
// The box of a thread lives in shared memory, dimension-major (lo[j][thread], hi[j][thread]: the split dimension is a run-time
// index, which in registers means local memory -- 1.9e6 local loads per depth-12 batch, 66 us); the centres leave through the
// same tile so that a block writes its 128 x d doubles as one contiguous range.  Dynamic shared memory: 2 * d * 128 doubles.
constexpr int LEAF_THREADS = 128;
__host__ __device__ constexpr int leaf_smem_bytes(int d) { return 2 * d * LEAF_THREADS * (int)sizeof(double); }

__global__ void __launch_bounds__(LEAF_THREADS) grow_leaves_kernel(const double* __restrict__ bounds /* [d][2] */, int d, int depth,
                                                                   long long nrows, double* __restrict__ out /* [nrows][d] */,
                                                                   long long row0 = 0) {
    extern __shared__ double leaf_sm[];
    const int tid = threadIdx.x;
    double* lo = leaf_sm + tid;                       // lo[j * LEAF_THREADS]
    double* hi = leaf_sm + d * LEAF_THREADS + tid;    // hi[j * LEAF_THREADS]
    const long long block0 = (long long)blockIdx.x * LEAF_THREADS;
    const long long local = block0 + tid;
    if (local < nrows) {
        const long long r = row0 + local;
        // level of row r: offsets (3^l - 1)/2
        int level = 0;
        long long off = 0, width = 1;  // width = 3^level
        while (off + width <= r) {
            off += width;
            width *= 3;
            level++;
        }
        // path from the subtree root = the base-3 digits of the index inside the level, most significant first.  The index is
        // below 3^20 < 2^32 for every supported depth: 32-bit divisions by the constant 3, digits packed two bits each
        unsigned idx32 = (unsigned)(r - off);
        unsigned long long path = 0;
        for (int s = 0; s < level; s++) {
            const unsigned q = idx32 / 3u;
            path |= (unsigned long long)(idx32 - 3u * q) << (2 * s);
            idx32 = q;
        }
        for (int j = 0; j < d; j++) {
            lo[j * LEAF_THREADS] = bounds[2 * j];
            hi[j * LEAF_THREADS] = bounds[2 * j + 1];
        }
        for (int s = 0; s < level; s++) {
            const int digit = (int)((path >> (2 * (level - 1 - s))) & 3ULL);
            int k = 0;
            double wk = __dsub_rn(hi[0], lo[0]);
            for (int j = 1; j < d; j++) {
                const double wj = __dsub_rn(hi[j * LEAF_THREADS], lo[j * LEAF_THREADS]);
                if (wj > wk) {
                    wk = wj;
                    k = j;
                }
            }
            const double delta = __ddiv_rn(wk, 3.0);
            const double base = lo[k * LEAF_THREADS];
            const double c0 = __dadd_rn(base, __dmul_rn((double)digit, delta));
            const double c1 = __dadd_rn(base, __dmul_rn((double)(digit + 1), delta));
            lo[k * LEAF_THREADS] = c0;
            hi[k * LEAF_THREADS] = c1;
        }
        for (int j = 0; j < d; j++) lo[j * LEAF_THREADS] = __dmul_rn(__dadd_rn(lo[j * LEAF_THREADS], hi[j * LEAF_THREADS]), 0.5);
    }
    __syncthreads();
    // centres of the block's rows, row-major and contiguous in out
    const long long rows_here = nrows - block0 < LEAF_THREADS ? nrows - block0 : LEAF_THREADS;
    const int total = (int)rows_here * d;
    double* dst = out + block0 * d;
    for (int e = tid; e < total; e += LEAF_THREADS) {
        const int rr = e / d, j = e - rr * d;
        dst[e] = leaf_sm[j * LEAF_THREADS + rr];
    }
}

}  // namespace gpso
