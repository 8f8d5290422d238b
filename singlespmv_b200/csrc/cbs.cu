// cbs.cu -- column-blocked compressed slices (see cbs.cuh): conversion on the device + the per-block multiply kernel.
#include <algorithm>
#include <cstdlib>

#include "cbs.cuh"

namespace b2 {

__device__ __forceinline__ int warp_sum_i(int v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ int warp_max_i(int v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// lane = row: entries of the row per column block (the row's columns ascend, so one walk), totals per (block, slice),
// and the statistics the layout decision is taken from: stats[0] = sum over rows of blocks touched, stats[1] = non-empty
// rows, stats[2] = a row has more than 255 entries in one block (the one-byte counts do not apply)
__global__ void cbs_count_kernel(const int *__restrict__ ptr, const int *__restrict__ col, int nRow, int nSlices, int B,
                                 int nBlock, unsigned char *__restrict__ cnt, int *__restrict__ total,
                                 unsigned long long *__restrict__ stats)
{
    const int s = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (s >= nSlices) return;
    const int r = s * 32 + lane;
    const int q = r < nRow ? ptr[r + 1] : 0;
    int j = r < nRow ? ptr[r] : 0;
    const bool nonEmpty = q > j;
    const size_t rowsPad = (size_t)nSlices * 32;
    int touched = 0, over = 0;
    for (int b = 0; b < nBlock; b++) {
        const long long hi = (long long)(b + 1) * B;
        int c = 0;
        while (j < q && col[j] < hi) { c++; j++; }
        over |= c > 255;
        cnt[(size_t)b * rowsPad + r] = (unsigned char)min(c, 255);
        touched += c > 0;
        const int t = warp_sum_i(c);
        if (lane == 0) total[(size_t)b * nSlices + s] = t;
    }
    touched = warp_sum_i(touched);
    const unsigned ne = __ballot_sync(0xffffffffu, nonEmpty), ov = __ballot_sync(0xffffffffu, over != 0);
    if (lane == 0) {
        if (touched) atomicAdd(&stats[0], (unsigned long long)touched);
        if (ne) atomicAdd(&stats[1], (unsigned long long)__popc(ne));
        if (ov) atomicAdd(&stats[2], 1ull);
    }
}

// warp = slice: writes the entries of every (block, slice) in jagged-diagonal compressed order
__global__ void cbs_fill_kernel(const int *__restrict__ ptr, const int *__restrict__ col, const double *__restrict__ val,
                                int nRow, int nSlices, int nBlock, const unsigned char *__restrict__ cnt,
                                const int *__restrict__ base, int *__restrict__ ccol, double *__restrict__ cval)
{
    const int s = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (s >= nSlices) return;
    const int r = s * 32 + lane;
    const size_t rowsPad = (size_t)nSlices * 32;
    const unsigned lt = (1u << lane) - 1u;
    int j = r < nRow ? ptr[r] : 0;
    for (int b = 0; b < nBlock; b++) {
        const int c = cnt[(size_t)b * rowsPad + r];
        int pos = base[(size_t)b * nSlices + s];
        const int maxc = warp_max_i(c);
        for (int k = 0; k < maxc; k++) {
            const unsigned m = __ballot_sync(0xffffffffu, k < c);
            if (k < c) {
                const int at = pos + __popc(m & lt);
                ccol[at] = col[j + k];
                cval[at] = val[j + k];
            }
            pos += __popc(m);
        }
        j += c;
    }
}

// One column block, all slices: lane = row.  Step j of the warp reads the j-th entry (within this block) of every row
// that has one -- a contiguous run of col / val -- and each lane continues its row's running sum in ascending column
// order with unfused mul/add (pass b picks the sum up from y where pass b-1 left it): the reference's order.
template <bool FIRST, int U>
__global__ void __launch_bounds__(256)
cbs_spmv_kernel(const unsigned char *__restrict__ cnt_b, const int *__restrict__ base_b, const int *__restrict__ ccol,
                const double *__restrict__ cval, const double *__restrict__ x, double *__restrict__ y, int rowBegin,
                int rowEnd, int sliceBegin, int sliceEnd)
{
    const int s = sliceBegin + ((blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
    if (s >= sliceEnd) return;
    const uint64_t pol_stream = policy_evict_first(), pol_x = policy_evict_last();
    const int r = s * 32 + lane;
    const int c = cnt_b[r];
    int pos = base_b[s];
    const bool mine = r >= rowBegin && r < rowEnd;
    double acc = (FIRST || !mine) ? 0.0 : y[r];
    const int maxc = warp_max_i(c);
    const unsigned lt = (1u << lane) - 1u;
    for (int j0 = 0; j0 < maxc; j0 += U) {
        int ci[U];
        double v[U], xs[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            const bool on = j0 + u < c;
            const unsigned m = __ballot_sync(0xffffffffu, on);
            if (on) {
                const int at = pos + __popc(m & lt);
                ci[u] = ld_stream_i1(ccol + at, pol_stream);
                v[u] = ld_stream_d1(cval + at, pol_stream);
            }
            pos += __popc(m);
        }
#pragma unroll
        for (int u = 0; u < U; u++)
            if (j0 + u < c) xs[u] = ld_x(x + ci[u], pol_x);
#pragma unroll
        for (int u = 0; u < U; u++)
            if (j0 + u < c) acc = __dadd_rn(acc, __dmul_rn(v[u], xs[u]));
    }
    if (mine) y[r] = acc;
}

int ColBlockSell::build(const int *ptr, const int *col, const double *val, int nRow_, int nCol_, int nnz_, int want,
                        cudaStream_t s)
{
    release();
    nRow = nRow_; nCol = nCol_; nnz = nnz_;
    static const char *env = getenv("B200SPMV_CBS");                 // experiments: "0" never, "n" = n blocks
    if (env) want = atoi(env) == 0 ? -1 : atoi(env);
    if (want < 0 || nnz == 0 || nRow == 0 || nCol == 0) return B200SPMV_OK;
    int nb;
    if (want > 0) nb = want;
    else {
        if ((long long)nCol * 8 <= 64LL << 20) return B200SPMV_OK;   // x fits in L2 next to the matrix stream
        nb = (int)(((long long)nCol * 8 + CBS_SLICE_BYTES - 1) / CBS_SLICE_BYTES);
    }
    nb = std::max(1, std::min(nb, CBS_MAX_BLOCKS));
    B = (int)(((((long long)nCol + nb - 1) / nb) + 31) & ~31LL);
    nBlock = (int)(((long long)nCol + B - 1) / B);
    nSlices = ceil_div(nRow, 32);
    if ((long long)nBlock * nSlices + 1 > 0x7fffffffLL) return B200SPMV_OK;
    DevBuf<int> total;
    DevBuf<unsigned long long> stats;
    B2_TRY(cnt.alloc((size_t)nBlock * nSlices * 32));
    B2_TRY(total.alloc((size_t)nBlock * nSlices + 1));
    B2_TRY(stats.alloc(3));
    B2_CUDA(cudaMemsetAsync(stats.p, 0, stats.bytes(), s));
    B2_CUDA(cudaMemsetAsync(total.p + (size_t)nBlock * nSlices, 0, sizeof(int), s));
    const int grid = ceil_div((long long)nSlices * 32, 256);
    cbs_count_kernel<<<grid, 256, 0, s>>>(ptr, col, nRow, nSlices, B, nBlock, cnt.p, total.p, stats.p);
    B2_KERNEL_CHECK();
    unsigned long long st[3] = {0, 0, 0};
    B2_CUDA(cudaMemcpyAsync(st, stats.p, sizeof st, cudaMemcpyDeviceToHost, s));
    B2_CUDA(cudaStreamSynchronize(s));
    // a row with more than 255 entries in one block (power-law heads) is not for this layout; nor is a matrix whose rows
    // each live in one block (stencils, banded matrices: x is already reused through L1/L2 by neighbouring rows)
    if (st[2] || (want == 0 && (double)st[0] < 1.5 * (double)st[1])) {
        release();
        return B200SPMV_OK;
    }
    B2_TRY(base.alloc((size_t)nBlock * nSlices + 1));
    B2_TRY(exclusive_scan_i32(total.p, base.p, nBlock * nSlices + 1, s));
    B2_TRY(ccol.alloc((size_t)nnz));
    B2_TRY(cval.alloc((size_t)nnz));
    cbs_fill_kernel<<<grid, 256, 0, s>>>(ptr, col, val, nRow, nSlices, nBlock, cnt.p, base.p, ccol.p, cval.p);
    B2_KERNEL_CHECK();
    B2_CUDA(cudaStreamSynchronize(s));
    active = true;
    return B200SPMV_OK;
}

int ColBlockSell::run(const double *x, double *y, int rb, int re, cudaStream_t s) const
{
    if (rb < 0 || re > nRow || rb > re) { set_error("multiply_rows: bad row range [%d,%d)", rb, re); return B200SPMV_ERR_INVALID; }
    if (rb == re) return B200SPMV_OK;
    const int sb = rb / 32, se = ceil_div(re, 32);
    const int grid = ceil_div((long long)(se - sb) * 32, 256);
    const size_t rowsPad = (size_t)nSlices * 32;
    for (int b = 0; b < nBlock; b++) {
        const unsigned char *cb = cnt.p + (size_t)b * rowsPad;
        const int *bb = base.p + (size_t)b * nSlices;
        if (b == 0) cbs_spmv_kernel<true, 8><<<grid, 256, 0, s>>>(cb, bb, ccol.p, cval.p, x, y, rb, re, sb, se);
        else cbs_spmv_kernel<false, 8><<<grid, 256, 0, s>>>(cb, bb, ccol.p, cval.p, x, y, rb, re, sb, se);
    }
    B2_KERNEL_CHECK();
    return B200SPMV_OK;
}

}  // namespace b2
