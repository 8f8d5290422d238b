// cbs.cuh -- "column-blocked compressed slices": the device layout the row-wise formats (ELL, JDS, SS) switch to when
// x does not fit in L2 and the matrix gathers from all over it (BASELINE.json config 2: uniform random, x = 134 MB).
//
// Why: every x[col] of such a matrix misses L2 and pulls a 64-byte DRAM burst for 8 useful bytes -- ncu on round 1's
// ELL kernel: 37.6 GB of DRAM traffic for 6.7 GB of matrix (profiles/r1_ncu_kernels.md).  The reference's own answer
// is column blocking (CSS, src/opt_css.cpp:33-45): process the entries of one column block for ALL rows before moving
// on, so that the block's slice of x stays cache-resident.  CSS pays for it with one row_ptr per block and separate
// block sums.  This layout keeps the row-wise formats' properties instead:
//   * slices of 32 rows (one warp, lane = row) like sliced ELL, but with NO padding: inside (column block, slice) the
//     entries are stored "jagged-diagonal compressed" -- for j = 0, 1, ...: the j-th entry of every row that has one, in
//     lane order.  A lane finds its slot with one ballot + popc; each step of the warp reads one contiguous run.
//   * per row and block one byte (the entry count); per (block, slice) one offset.
//   * pass b continues the running sum of pass b-1 (acc = y[r]; acc += ...; y[r] = acc), so every row is still summed
//     strictly in ascending column order with unfused mul/add: bit-identical to the reference CRS / ELL / JDS result.
// Block width: the largest multiple of 32 columns whose slice of x is <= CBS_SLICE_BYTES (45 MB; round 1 measured 2 / 3 /
// 4 / 8 blocks on config 2: 3 is the optimum, profiles/r1_experiments.md).
#pragma once
#include "common.cuh"

namespace b2 {

constexpr long long CBS_SLICE_BYTES = 45LL << 20;
constexpr int CBS_MAX_BLOCKS = 32;

struct ColBlockSell {
    int nRow = 0, nCol = 0, nnz = 0, nBlock = 0, B = 0, nSlices = 0;
    bool active = false;
    DevBuf<unsigned char> cnt;      // [nBlock][nSlices * 32]  entries of row r in block b
    DevBuf<int> base;               // [nBlock * nSlices + 1]  first entry of (block, slice) in ccol / cval
    DevBuf<int> ccol;               // [nnz]
    DevBuf<double> cval;            // [nnz]

    // want: 0 = decide (x larger than L2's useful share AND rows spread over the blocks), > 0 = that many column blocks,
    // < 0 = never.  ptr/col/val: the sorted entries with their CRS row pointer.  Leaves active = false when not used.
    int build(const int *ptr, const int *col, const double *val, int nRow_, int nCol_, int nnz_, int want, cudaStream_t s);
    // y[rb..re) = A x   (all column blocks, one launch each)
    int run(const double *x, double *y, int rb, int re, cudaStream_t s) const;
    long long meta_bytes() const { return (long long)cnt.bytes() + (long long)base.bytes(); }
    void release()
    {
        cnt.release(); base.release(); ccol.release(); cval.release();
        active = false;
    }
};

}  // namespace b2
