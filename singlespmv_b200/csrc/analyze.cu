// analyze.cu -- matrix statistics on the device and a format recommendation from them.
// The statistics are those of the reference's matrix/script/counter.cpp:19-42 (row/column non-zero counts:
// max, min, variance of the row counts) plus the number of non-empty diagonals (the quantity the
// reference's DIA plugin derives, src/opt_dia.cpp:29-34) and the number of empty rows.
// The recommendation encodes what was measured on B200 (profiles/): DIA when a few dense diagonals hold the
// matrix, column-blocked CSS when x is too large to stay in L2 under random gathers, CSR5 for heavily skewed
// rows, sliced ELL for near-uniform rows, CRS otherwise.
#include <algorithm>
#include <cmath>

#include "common.cuh"

using namespace b2;

namespace {

__global__ void an_counts_kernel(const int *__restrict__ row, const int *__restrict__ col, long long nnz, int shift,
                                 int *__restrict__ rowCnt, int *__restrict__ colCnt, int *__restrict__ diagFlag)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nnz) return;
    const int r = row[i], c = col[i];
    atomicAdd(&rowCnt[r], 1);
    atomicAdd(&colCnt[c], 1);
    diagFlag[c - r + shift] = 1;
}

struct MinMaxSum {
    int mn, mx;
    unsigned long long sum, sumsq, zeros;
};

__global__ void an_reduce_kernel(const int *__restrict__ cnt, int n, MinMaxSum *__restrict__ out)
{
    int mn = 0x7fffffff, mx = 0;
    unsigned long long sum = 0, sumsq = 0, zeros = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int v = cnt[i];
        mn = min(mn, v);
        mx = max(mx, v);
        sum += (unsigned long long)v;
        sumsq += (unsigned long long)v * (unsigned long long)v;
        zeros += v == 0;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        sum += __shfl_xor_sync(0xffffffffu, sum, o);
        sumsq += __shfl_xor_sync(0xffffffffu, sumsq, o);
        zeros += __shfl_xor_sync(0xffffffffu, zeros, o);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMin(&out->mn, mn);
        atomicMax(&out->mx, mx);
        atomicAdd(&out->sum, sum);
        atomicAdd(&out->sumsq, sumsq);
        atomicAdd(&out->zeros, zeros);
    }
}

static int reduce_counts(const int *cnt, int n, MinMaxSum *out, cudaStream_t s)
{
    MinMaxSum init;
    init.mn = 0x7fffffff; init.mx = 0; init.sum = init.sumsq = init.zeros = 0;
    if (n <= 0) { *out = init; out->mn = 0; return B200SPMV_OK; }
    DevBuf<MinMaxSum> d;
    B2_TRY(d.alloc(1));
    B2_CUDA(cudaMemcpyAsync(d.p, &init, sizeof init, cudaMemcpyHostToDevice, s));
    an_reduce_kernel<<<std::min(ceil_div(n, 256), 148 * 8), 256, 0, s>>>(cnt, n, d.p);
    B2_KERNEL_CHECK();
    B2_CUDA(cudaMemcpyAsync(out, d.p, sizeof(MinMaxSum), cudaMemcpyDeviceToHost, s));
    B2_CUDA(cudaStreamSynchronize(s));
    return B200SPMV_OK;
}

}  // namespace

extern "C" int b200spmv_analyze(const b200spmv_coo *coo, b200spmv_stats *out, void *stream)
{
    clear_error();
    cudaStream_t s = (cudaStream_t)stream;
    if (!coo || !out) { set_error("analyze: NULL argument"); return B200SPMV_ERR_INVALID; }
    memset(out, 0, sizeof *out);
    const int nRow = coo->rowEnd > coo->rowBegin ? coo->nRow : coo->nRow, nCol = coo->nCol;
    const long long nnz = coo->nnz;
    const long long N = (long long)nRow + nCol - 1;
    if (coo->rowBegin != 0 || (coo->rowEnd != 0 && coo->rowEnd != coo->nRow)) {
        set_error("analyze: pass the whole matrix (rows [0,nRow)), not a row slice");
        return B200SPMV_ERR_INVALID;
    }
    DevBuf<int> rowCnt, colCnt, diag;
    B2_TRY(rowCnt.alloc((size_t)nRow));
    B2_TRY(colCnt.alloc((size_t)nCol));
    B2_TRY(diag.alloc((size_t)(N > 0 ? N : 0)));
    B2_CUDA(cudaMemsetAsync(rowCnt.p, 0, rowCnt.bytes() ? rowCnt.bytes() : 4, s));
    B2_CUDA(cudaMemsetAsync(colCnt.p, 0, colCnt.bytes() ? colCnt.bytes() : 4, s));
    B2_CUDA(cudaMemsetAsync(diag.p, 0, diag.bytes() ? diag.bytes() : 4, s));
    if (nnz) {
        an_counts_kernel<<<ceil_div(nnz, 256), 256, 0, s>>>(coo->row_d, coo->col_d, nnz, nRow - 1, rowCnt.p, colCnt.p, diag.p);
        B2_KERNEL_CHECK();
    }
    MinMaxSum r, c, d;
    B2_TRY(reduce_counts(rowCnt.p, nRow, &r, s));
    B2_TRY(reduce_counts(colCnt.p, nCol, &c, s));
    B2_TRY(reduce_counts(diag.p, (int)(N > 0 ? N : 0), &d, s));
    out->nRow = nRow; out->nCol = nCol; out->nnz = nnz;
    out->rowMax = r.mx; out->rowMin = r.mn; out->colMax = c.mx; out->colMin = c.mn;
    out->nEmptyRows = (long long)r.zeros;
    out->nDiag = (long long)d.sum;
    const double ave = nRow > 0 ? (double)r.sum / nRow : 0.0;
    out->rowMean = ave;
    out->rowVar = nRow > 0 ? (double)r.sumsq / nRow - ave * ave : 0.0;     // = sum (cnt-ave)^2 / N of counter.cpp:31-34
    if (out->rowVar < 0) out->rowVar = 0;
    return B200SPMV_OK;
}

extern "C" int b200spmv_recommend_format(const b200spmv_stats *st, b200spmv_options *opts)
{
    if (!st) { set_error("recommend_format: NULL stats"); return B200SPMV_ERR_INVALID; }
    if (opts) memset(opts, 0, sizeof *opts);
    if (st->nnz == 0 || st->nRow == 0) return B200SPMV_CRS;
    const double diaFill = st->nDiag > 0 ? (double)st->nnz / ((double)st->nDiag * st->nRow) : 0.0;
    const double ellFill = st->rowMax > 0 ? (double)st->nnz / ((double)st->rowMax * st->nRow) : 0.0;
    const double cv = st->rowMean > 0 ? sqrt(st->rowVar) / st->rowMean : 0.0;
    const double xBytes = 8.0 * st->nCol;
    // a few well-filled diagonals: 8 B per stored slot and no index stream at all (c4: 0.88, c5: 0.99 of the copy peak)
    if (st->nDiag <= 64 && diaFill >= 0.5) return B200SPMV_DIA;
    // heavy skew: tiles of equal non-zero count with per-lane segments (c3: +11 % over cuSPARSE, +23 % over CRS)
    if (cv > 1.0) return B200SPMV_CSR5;
    // x larger than what L2 keeps under random gathers: column blocks of ~45 MB (c2: CSS(3) is 2x ELL/CRS)
    if (xBytes > 64e6 && st->rowMean >= 8.0 && diaFill < 0.05) {
        if (opts) opts->n_block = (int)((xBytes + 45e6 - 1) / 45e6);
        return B200SPMV_CSS;
    }
    // near-uniform rows: padded slices cost little and the kernel is a pure stream (c4: above the copy peak)
    if (ellFill >= 0.9 && st->rowMax <= st->nCol) return B200SPMV_ELL;
    // mostly uniform rows with a few long ones (ELL alone would store 10-100 % padding): ELL part + COO tail
    if (ellFill > 0.5 && cv <= 0.5) return B200SPMV_HYB;
    return B200SPMV_CRS;
}
