// synth.cu -- deterministic synthetic matrices generated directly in HBM.
// The reference ingests Matrix-Market text (/root/reference/src/util.cpp:30-66) and ships small
// pattern generators (/root/reference/matrix/artificial/generator.cpp:12-79); BASELINE.json's
// shapes reach 938 M non-zeros, so the same kinds of patterns are produced on the device, as
// COO sorted by (row, col) without duplicates -- the plugins' input contract.
// Definitions are shared with the test-side statement in oracle/synth_oracle.c (bit-identical).
#include <cub/cub.cuh>

#include "common.cuh"

namespace b2 {

__host__ __device__ __forceinline__ uint64_t mix64(uint64_t z)      // splitmix64 output function
{
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
__host__ __device__ __forceinline__ double u01(uint64_t h) { return (double)(h >> 11) * (1.0 / 9007199254740992.0); }
__host__ __device__ __forceinline__ double entry_value(uint64_t seed, int r, int c)
{
    return u01(mix64(mix64(seed ^ 0xA5A5A5A5A5A5A5A5ull) + (((uint64_t)(uint32_t)r << 32) | (uint32_t)c)));
}

// ---------------------------------------------------------------- stencils
__device__ __forceinline__ int span(int i, int n) { return 1 + (i > 0) + (i < n - 1); }

template <typename OT>
__global__ void stencil_count_kernel(int kind, int n, int rowBegin, int rows, OT *__restrict__ cnt)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i > rows) return;
    if (i == rows) { cnt[i] = 0; return; }
    const int r = rowBegin + i;
    if (kind == B200SPMV_SYNTH_LAP2D5) {
        cnt[i] = span(r / n, n) + span(r % n, n) - 1;
    } else {
        const int k = r % n, j = (r / n) % n, ii = r / (n * n);
        cnt[i] = kind == B200SPMV_SYNTH_LAP3D7 ? span(ii, n) + span(j, n) + span(k, n) - 2
                                               : span(ii, n) * span(j, n) * span(k, n);
    }
}

template <typename OT>
__global__ void stencil_fill_kernel(int kind, int n, int rowBegin, int rows, const OT *__restrict__ off,
                                    int *__restrict__ row, int *__restrict__ col, double *__restrict__ val)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows) return;
    const int r = rowBegin + i;
    OT at = off[i];
    if (kind == B200SPMV_SYNTH_LAP2D5) {
        const int gi = r / n, gj = r % n;
        for (int di = -1; di <= 1; di++)
            for (int dj = -1; dj <= 1; dj++) {
                if (di != 0 && dj != 0) continue;
                const int a = gi + di, b = gj + dj;
                if (a < 0 || a >= n || b < 0 || b >= n) continue;
                row[at] = r; col[at] = a * n + b; val[at] = (di == 0 && dj == 0) ? 4.0 : -1.0; at++;
            }
        return;
    }
    const int gk = r % n, gj = (r / n) % n, gi = r / (n * n);
    for (int di = -1; di <= 1; di++)
        for (int dj = -1; dj <= 1; dj++)
            for (int dk = -1; dk <= 1; dk++) {
                const int taxi = abs(di) + abs(dj) + abs(dk);
                if (kind == B200SPMV_SYNTH_LAP3D7 && taxi > 1) continue;
                const int a = gi + di, b = gj + dj, c = gk + dk;
                if (a < 0 || a >= n || b < 0 || b >= n || c < 0 || c >= n) continue;
                row[at] = r; col[at] = (a * n + b) * n + c;
                val[at] = taxi == 0 ? (kind == B200SPMV_SYNTH_LAP3D7 ? 6.0 : 26.0) : -1.0; at++;
            }
}

// ---------------------------------------------------------------- uniform random, K distinct columns per row
constexpr int UNIFORM_MAX_K = 64;

__global__ void uniform_kernel(uint64_t seed, int nCol, int K, int rowBegin, int rows, int *__restrict__ row,
                               int *__restrict__ col, double *__restrict__ val)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows) return;
    const int r = rowBegin + i;
    const uint64_t rowkey = mix64(mix64(seed) ^ (uint64_t)(uint32_t)r);
    int pick[UNIFORM_MAX_K];
    int have = 0;
    for (uint64_t j = 0; have < K; j++) {
        const int c = (int)(mix64(rowkey + j) % (uint64_t)nCol);
        bool dup = false;
        for (int t = 0; t < have; t++) dup |= pick[t] == c;
        if (!dup) pick[have++] = c;
    }
    for (int a = 1; a < K; a++) {
        const int c = pick[a];
        int b = a - 1;
        while (b >= 0 && pick[b] > c) { pick[b + 1] = pick[b]; b--; }
        pick[b + 1] = c;
    }
    const size_t base = (size_t)i * K;
    for (int t = 0; t < K; t++) {
        row[base + t] = r; col[base + t] = pick[t]; val[base + t] = entry_value(seed, r, pick[t]);
    }
}

// ---------------------------------------------------------------- R-MAT
#define RMAT_A 2448131358u    // floor(0.57 * 2^32)
#define RMAT_AB 3264175144u   // floor(0.76 * 2^32)
#define RMAT_ABC 4080218931u  // floor(0.95 * 2^32)

__global__ void rmat_edges_kernel(uint64_t seed, int scale, long long nEdges, uint64_t *__restrict__ key)
{
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= nEdges) return;
    const uint64_t edgekey = mix64(mix64(seed) + (uint64_t)e);
    uint32_t r = 0, c = 0;
    for (int l = 0; l < scale; l++) {
        const uint32_t t = (uint32_t)(mix64(edgekey + (uint64_t)l) >> 32);
        const uint32_t rb = t >= RMAT_AB, cb = (t >= RMAT_A && t < RMAT_AB) || t >= RMAT_ABC;
        r = (r << 1) | rb;
        c = (c << 1) | cb;
    }
    key[e] = ((uint64_t)r << 32) | c;
}

__global__ void rmat_unpack_kernel(uint64_t seed, const uint64_t *__restrict__ key, long long n,
                                   int *__restrict__ row, int *__restrict__ col, double *__restrict__ val)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int r = (int)(key[i] >> 32), c = (int)(key[i] & 0xFFFFFFFFu);
    row[i] = r; col[i] = c; val[i] = entry_value(seed, r, c);
}

static int alloc_coo(b200spmv_coo *out, long long nnz)
{
    size_t n = nnz > 0 ? (size_t)nnz : 1;
    out->row_d = out->col_d = nullptr;
    out->val_d = nullptr;
    B2_CUDA(cudaMalloc((void **)&out->row_d, n * sizeof(int)));
    B2_CUDA(cudaMalloc((void **)&out->col_d, n * sizeof(int)));
    B2_CUDA(cudaMalloc((void **)&out->val_d, n * sizeof(double)));
    out->nnz = nnz;
    return B200SPMV_OK;
}

static int synth_rmat(long long scale, long long nEdges, uint64_t seed, b200spmv_coo *out, cudaStream_t s)
{
    if (scale < 1 || scale > 30 || nEdges < 0 || nEdges > 0x7fffffffLL) {
        set_error("synth RMAT: scale=%lld edges=%lld out of range", scale, nEdges);
        return B200SPMV_ERR_INVALID;
    }
    out->nRow = out->nCol = 1 << scale;
    out->rowBegin = 0;
    out->rowEnd = out->nRow;
    DevBuf<uint64_t> a, b, uniq;
    DevBuf<long long> nsel;
    B2_TRY(a.alloc((size_t)nEdges));
    B2_TRY(b.alloc((size_t)nEdges));
    B2_TRY(nsel.alloc(1));
    if (nEdges == 0) return alloc_coo(out, 0);
    rmat_edges_kernel<<<ceil_div(nEdges, 256), 256, 0, s>>>(seed, (int)scale, nEdges, a.p);
    B2_KERNEL_CHECK();
    cub::DoubleBuffer<uint64_t> keys(a.p, b.p);
    size_t tmp = 0;
    B2_CUDA(cub::DeviceRadixSort::SortKeys(nullptr, tmp, keys, (int)nEdges, 0, 32 + (int)scale, s));
    {
        DevBuf<char> t;
        B2_TRY(t.alloc(tmp));
        B2_CUDA(cub::DeviceRadixSort::SortKeys(t.p, tmp, keys, (int)nEdges, 0, 32 + (int)scale, s));
        B2_CUDA(cudaStreamSynchronize(s));
    }
    uint64_t *sorted = keys.Current(), *other = keys.Alternate();
    B2_CUDA(cub::DeviceSelect::Unique(nullptr, tmp, sorted, other, nsel.p, (int)nEdges, s));
    {
        DevBuf<char> t;
        B2_TRY(t.alloc(tmp));
        B2_CUDA(cub::DeviceSelect::Unique(t.p, tmp, sorted, other, nsel.p, (int)nEdges, s));
        B2_CUDA(cudaStreamSynchronize(s));
    }
    long long nnz = 0;
    B2_CUDA(cudaMemcpy(&nnz, nsel.p, sizeof(long long), cudaMemcpyDeviceToHost));
    B2_TRY(alloc_coo(out, nnz));
    rmat_unpack_kernel<<<ceil_div(nnz, 256), 256, 0, s>>>(seed, other, nnz, out->row_d, out->col_d, out->val_d);
    B2_KERNEL_CHECK();
    B2_CUDA(cudaStreamSynchronize(s));
    return B200SPMV_OK;
}

}  // namespace b2
using namespace b2;

extern "C" int b200spmv_synth(int kind, long long p0, long long p1, unsigned long long seed, int rowBegin,
                              int rowEnd, b200spmv_coo *out, void *stream)
{
    clear_error();
    cudaStream_t s = (cudaStream_t)stream;
    if (!out) { set_error("synth: out is NULL"); return B200SPMV_ERR_INVALID; }
    memset(out, 0, sizeof *out);
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        set_error("synth: no CUDA device available; libb200spmv has no CPU fallback");
        return B200SPMV_ERR_CUDA;
    }
    if (kind == B200SPMV_SYNTH_RMAT) {
        if (rowBegin != 0 || rowEnd > 0) { set_error("synth RMAT: row ranges are not supported"); return B200SPMV_ERR_UNSUPPORTED; }
        return synth_rmat(p0, p1, seed, out, s);
    }
    long long nRowLL;
    if (kind == B200SPMV_SYNTH_LAP2D5) nRowLL = p0 * p0;
    else if (kind == B200SPMV_SYNTH_LAP3D7 || kind == B200SPMV_SYNTH_BOX3D27) nRowLL = p0 * p0 * p0;
    else if (kind == B200SPMV_SYNTH_UNIFORM) nRowLL = p0;
    else { set_error("synth: unknown kind %d", kind); return B200SPMV_ERR_INVALID; }
    if (p0 < 1 || nRowLL > 0x7fffffffLL) { set_error("synth: size parameter %lld out of range", p0); return B200SPMV_ERR_INVALID; }
    const int nRow = (int)nRowLL;
    if (rowEnd <= 0) rowEnd = nRow;
    if (rowBegin < 0 || rowBegin > rowEnd || rowEnd > nRow) { set_error("synth: bad row range [%d,%d)", rowBegin, rowEnd); return B200SPMV_ERR_INVALID; }
    const int rows = rowEnd - rowBegin;
    out->nRow = out->nCol = nRow;
    out->rowBegin = rowBegin;
    out->rowEnd = rowEnd;
    if (kind == B200SPMV_SYNTH_UNIFORM) {
        if (p1 < 1 || p1 > UNIFORM_MAX_K || p1 > nRow) { set_error("synth UNIFORM: K=%lld must be in [1,%d] and <= nCol", p1, UNIFORM_MAX_K); return B200SPMV_ERR_INVALID; }
        const long long nnz = (long long)rows * p1;
        B2_TRY(alloc_coo(out, nnz));
        if (rows) uniform_kernel<<<ceil_div(rows, 128), 128, 0, s>>>(seed, nRow, (int)p1, rowBegin, rows, out->row_d, out->col_d, out->val_d);
        B2_KERNEL_CHECK();
        B2_CUDA(cudaStreamSynchronize(s));
        return B200SPMV_OK;
    }
    // int32 offsets unless the requested rows can hold more than 2^31-1 entries (rows x longest row): then 64-bit
    // offsets -- one GPU holds such a matrix easily (180 GB); the formats take it as row blocks (csrc/blocked.cu)
    const long long perRow = kind == B200SPMV_SYNTH_LAP2D5 ? 5 : kind == B200SPMV_SYNTH_LAP3D7 ? 7 : 27;
    if ((long long)rows * perRow > 0x7fffffffLL) {
        DevBuf<long long> cnt;
        B2_TRY(cnt.alloc((size_t)rows + 1));
        stencil_count_kernel<long long><<<ceil_div((long long)rows + 1, 256), 256, 0, s>>>(kind, (int)p0, rowBegin, rows, cnt.p);
        B2_KERNEL_CHECK();
        B2_TRY(exclusive_scan_i64(cnt.p, cnt.p, rows + 1, s));
        long long nnz = 0;
        B2_CUDA(cudaMemcpy(&nnz, cnt.p + rows, sizeof(long long), cudaMemcpyDeviceToHost));
        B2_TRY(alloc_coo(out, nnz));
        stencil_fill_kernel<long long><<<ceil_div(rows, 256), 256, 0, s>>>(kind, (int)p0, rowBegin, rows, cnt.p, out->row_d, out->col_d, out->val_d);
        B2_KERNEL_CHECK();
        B2_CUDA(cudaStreamSynchronize(s));
        return B200SPMV_OK;
    }
    DevBuf<int> cnt;
    B2_TRY(cnt.alloc((size_t)rows + 1));
    stencil_count_kernel<int><<<ceil_div((long long)rows + 1, 256), 256, 0, s>>>(kind, (int)p0, rowBegin, rows, cnt.p);
    B2_KERNEL_CHECK();
    B2_TRY(exclusive_scan_i32(cnt.p, cnt.p, rows + 1, s));
    int nnz = 0;
    B2_CUDA(cudaMemcpy(&nnz, cnt.p + rows, sizeof(int), cudaMemcpyDeviceToHost));
    B2_TRY(alloc_coo(out, nnz));
    if (rows) stencil_fill_kernel<int><<<ceil_div(rows, 256), 256, 0, s>>>(kind, (int)p0, rowBegin, rows, cnt.p, out->row_d, out->col_d, out->val_d);
    B2_KERNEL_CHECK();
    B2_CUDA(cudaStreamSynchronize(s));
    return B200SPMV_OK;
}

// rows [rowBegin, rowEnd) of a host COO (global row ids, sorted) -> device arrays owned by *out
extern "C" int b200spmv_coo_upload(int nRow, int nCol, int rowBegin, int rowEnd, long long nnz, const int *row_h,
                                   const int *col_h, const double *val_h, b200spmv_coo *out)
{
    clear_error();
    if (!out || nnz < 0 || (nnz > 0 && (!row_h || !col_h || !val_h)) || rowBegin < 0 || rowEnd < rowBegin || rowEnd > nRow) {
        set_error("coo_upload: bad argument");
        return B200SPMV_ERR_INVALID;
    }
    memset(out, 0, sizeof *out);
    out->nRow = nRow; out->nCol = nCol; out->rowBegin = rowBegin; out->rowEnd = rowEnd;
    B2_TRY(alloc_coo(out, nnz));
    if (nnz) {
        B2_CUDA(cudaMemcpy(out->row_d, row_h, (size_t)nnz * sizeof(int), cudaMemcpyHostToDevice));
        B2_CUDA(cudaMemcpy(out->col_d, col_h, (size_t)nnz * sizeof(int), cudaMemcpyHostToDevice));
        B2_CUDA(cudaMemcpy(out->val_d, val_h, (size_t)nnz * sizeof(double), cudaMemcpyHostToDevice));
    }
    return B200SPMV_OK;
}

extern "C" int b200spmv_coo_free(b200spmv_coo *coo)
{
    if (!coo) return B200SPMV_OK;
    if (coo->row_d) cudaFree(coo->row_d);
    if (coo->col_d) cudaFree(coo->col_d);
    if (coo->val_d) cudaFree(coo->val_d);
    coo->row_d = coo->col_d = nullptr;
    coo->val_d = nullptr;
    coo->nnz = 0;
    return B200SPMV_OK;
}

extern "C" int b200spmv_coo_download(const b200spmv_coo *coo, int *row_h, int *col_h, double *val_h)
{
    clear_error();
    if (!coo) { set_error("coo_download: NULL coo"); return B200SPMV_ERR_INVALID; }
    const size_t n = (size_t)coo->nnz;
    if (n == 0) return B200SPMV_OK;
    if (!row_h || !col_h || !val_h) { set_error("coo_download: NULL destination"); return B200SPMV_ERR_INVALID; }
    B2_CUDA(cudaMemcpy(row_h, coo->row_d, n * sizeof(int), cudaMemcpyDeviceToHost));
    B2_CUDA(cudaMemcpy(col_h, coo->col_d, n * sizeof(int), cudaMemcpyDeviceToHost));
    B2_CUDA(cudaMemcpy(val_h, coo->val_d, n * sizeof(double), cudaMemcpyDeviceToHost));
    return B200SPMV_OK;
}
