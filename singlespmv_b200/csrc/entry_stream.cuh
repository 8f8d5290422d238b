// entry_stream.cuh -- the CRS "entry stream": the row-length-agnostic multiply for rows too long for the row-chunk stream
// (BASELINE.json: "warp- and vector-per-row CRS with an adaptive row-bin scheduler" -- here the scheduler disappears: every
// warp moves the same number of entries whatever the rows look like, 27 entries per row (config 4) or 0 ... 10^5 (R-MAT)).
//
// It is the COO entry stream (coo.cu) without the 4 B/entry row ids.  At conversion the row pointers are turned into
//   bits[]     one bit per entry: set where a (non-empty) row begins                        (nnz / 8 bytes)
//   grpRun[g]  number of row starts in front of entry 128 g                                 (4 bytes per 128 entries)
//   nzrow[k]   the k-th non-empty row -- only when the matrix has empty rows (else the identity)
//   empty[]    the empty rows (zero-filled by a tiny kernel: beta = 0)
// and the multiply never reads row_ptr: a lane takes 4 consecutive entries (idx, val: 12 B/entry, by TMA bulk copies on banded
// matrices, by its own 128-bit loads on gather-bound ones), its 4 start bits, sums its runs in registers; a segmented warp
// scan per 128 entries closes the runs that cross lanes; ballots over the start bits give every lane the index of the run it
// continues.  Chunk pieces are stitched per tile, tile pieces by a fix-up kernel in tile order: deterministic, no atomics.
// Sums are re-associated across lanes: within the 1e-12 tolerance, not the bits of the sequential order (options.crs_path = 1
// keeps the tile-stream, which preserves them for rows of up to 64 entries).
#pragma once
#include <map>

#include "common.cuh"

namespace b2 {

struct EntryStream {
    // borrowed
    const int *ptr = nullptr, *idx = nullptr;
    const double *val = nullptr;
    int nRow = 0, nnz = 0;
    // owned
    bool ok = false, tma = false;
    int E = 1024, nTiles = 0, nEmpty = 0;
    DevBuf<unsigned> bits;
    DevBuf<int> grpRun, nzrow, empty;
    DevBuf<double> carry;
    std::map<std::pair<int, int>, std::pair<int, int>> range_cache;

    // gather_bound: the x gathers need L2 residency (primitives.cu gathers_need_l2) -> load-fed, else TMA-fed
    int build(const int *ptr_d, const int *idx_d, const double *val_d, int nRow_, int nnz_, bool gather_bound, cudaStream_t s);
    int prepare(int rb, int re);
    int run_rows(const double *x, double *y, int rb, int re, cudaStream_t s);
};

}  // namespace b2
