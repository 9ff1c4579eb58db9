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
__global__ void __launch_bounds__(128) grow_leaves_kernel(const double* __restrict__ bounds /* [d][2] */, int d, int depth,
                                                          long long nrows, double* __restrict__ out /* [nrows][d] */,
                                                          long long row0 = 0) {
    const long long local = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (local >= nrows) return;
    const long long r = row0 + local;
    // level of row r: offsets (3^l - 1)/2
    int level = 0;
    long long off = 0, width = 1;  // width = 3^level
    while (off + width <= r) {
        off += width;
        width *= 3;
        level++;
    }
    long long idx = r - off;
    double lo[LEAF_MAXD], hi[LEAF_MAXD];
    for (int j = 0; j < d; j++) {
        lo[j] = bounds[2 * j];
        hi[j] = bounds[2 * j + 1];
    }
    long long pw = width;  // 3^level
    for (int s = 0; s < level; s++) {
        pw /= 3;
        int digit = (int)((idx / pw) % 3);
        int k = 0;
        double wk = __dsub_rn(hi[0], lo[0]);
        for (int j = 1; j < d; j++) {
            double wj = __dsub_rn(hi[j], lo[j]);
            if (wj > wk) {
                wk = wj;
                k = j;
            }
        }
        double delta = __ddiv_rn(wk, 3.0);
        double base = lo[k];
        double c0 = __dadd_rn(base, __dmul_rn((double)digit, delta));
        double c1 = __dadd_rn(base, __dmul_rn((double)(digit + 1), delta));
        lo[k] = c0;
        hi[k] = c1;
    }
    for (int j = 0; j < d; j++) out[local * d + j] = __dmul_rn(__dadd_rn(lo[j], hi[j]), 0.5);
}

}  // namespace gpso
