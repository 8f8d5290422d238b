// halo.cu -- device side of the row-partitioned multi-GPU multiply.
// The reference is single-node OpenMP (no distributed code at all); this is new design: contiguous
// row blocks balanced by non-zero count, x distributed like the rows, and per multiply only the x
// entries a block's columns actually reference outside its own range ("halo") cross NVLink.
//
// Local column numbering is MONOTONE in the global column id:
//     [ left halo (global < colBegin) | owned slice [colBegin,colEnd) | right halo (global >= colEnd) ]
// so every row stays sorted and is summed in the same order as on one GPU -> bit-identical y.
#include <cub/cub.cuh>

#include "common.cuh"

using namespace b2;

struct b200spmv_halo {
    int rowBegin = 0, rowEnd = 0, colBegin = 0, colEnd = 0;
    int nLocal = 0, nLeft = 0, nRight = 0;
    int interiorBegin = 0, interiorEnd = 0;       // local rows [interiorBegin, interiorEnd) touch no halo column
    DevBuf<int> halo_cols;                        // global ids, ascending, nLeft + nRight
    DevBuf<int> send_idx;                         // indices into the owned x slice, in peer order
    long long nSend = 0;
};

namespace {

struct IsRemote {
    int lo, hi;
    __host__ __device__ bool operator()(const int &c) const { return c < lo || c >= hi; }
};

__device__ __forceinline__ int lower_bound_dev(const int *a, int n, int key)
{
    int lo = 0, hi = n;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (a[mid] < key) lo = mid + 1;
        else hi = mid;
    }
    return lo;
}

// in place: rows become block-local, columns get the monotone local numbering; rows that touch a
// halo column are recorded so that the largest all-interior middle range can be found
__global__ void halo_remap_kernel(int *__restrict__ row, int *__restrict__ col, long long nnz, int rowBegin,
                                  int colBegin, int colEnd, const int *__restrict__ halo, int nHalo, int nLeft,
                                  int nLocalRows, int *__restrict__ edge /* [2]: max boundary row in first half, min in second */)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nnz) return;
    const int r = row[i] - rowBegin, c = col[i];
    row[i] = r;
    if (c >= colBegin && c < colEnd) {
        col[i] = nLeft + (c - colBegin);
        return;
    }
    const int k = lower_bound_dev(halo, nHalo, c);
    col[i] = c < colBegin ? k : (colEnd - colBegin) + k;
    if (r < nLocalRows / 2) atomicMax(&edge[0], r);
    else atomicMin(&edge[1], r);
}

__global__ void halo_pack_kernel(const double *__restrict__ x_local, const int *__restrict__ idx, long long n,
                                 double *__restrict__ out)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = x_local[idx[i]];
}

__global__ void send_idx_kernel(const int *__restrict__ cols, long long n, int colBegin, int colEnd, int *__restrict__ idx,
                                int *__restrict__ bad)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int c = cols[i];
    if (c < colBegin || c >= colEnd) atomicExch(bad, 1);
    idx[i] = c - colBegin;
}

// first row r with ptr-equivalent >= target, read off the sorted COO row array
__global__ void partition_kernel(const int *__restrict__ row, long long nnz, int nRow, int nParts, int *__restrict__ bounds)
{
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g > nParts) return;
    if (g == 0) { bounds[0] = 0; return; }
    if (g == nParts) { bounds[g] = nRow; return; }
    const long long e = nnz * g / nParts;
    if (e >= nnz) { bounds[g] = nRow; return; }
    const int r = row[e];
    const bool starts_row = e == 0 || row[e - 1] != r;
    bounds[g] = starts_row ? r : r + 1;
}

__global__ void count_search_kernel(const long long *__restrict__ prefix /* [nRow+1] */, int nRow, int nParts,
                                    int *__restrict__ bounds)
{
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g > nParts) return;
    const long long target = prefix[nRow] * g / nParts;
    int lo = 0, hi = nRow;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (prefix[mid] < target) lo = mid + 1;
        else hi = mid;
    }
    bounds[g] = g == nParts ? nRow : lo;
}

__device__ __forceinline__ int span1(int i, int n) { return 1 + (i > 0) + (i < n - 1); }
__global__ void stencil_len_kernel(int kind, int n, int nRow, long long *__restrict__ cnt)
{
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r > nRow) return;
    if (r == nRow) { cnt[r] = 0; return; }
    if (kind == B200SPMV_SYNTH_LAP2D5) { cnt[r] = span1(r / n, n) + span1(r % n, n) - 1; return; }
    const int k = r % n, j = (r / n) % n, i = r / (n * n);
    cnt[r] = kind == B200SPMV_SYNTH_LAP3D7 ? span1(i, n) + span1(j, n) + span1(k, n) - 2
                                            : span1(i, n) * span1(j, n) * span1(k, n);
}

}  // namespace

extern "C" {

int b200spmv_partition_rows(const int *row_d, long long nnz, int nRow, int nParts, int *bounds_h)
{
    clear_error();
    if (nParts < 1 || !bounds_h || nRow < 0 || nnz < 0) { set_error("partition_rows: bad argument"); return B200SPMV_ERR_INVALID; }
    DevBuf<int> b;
    B2_TRY(b.alloc((size_t)nParts + 1));
    partition_kernel<<<ceil_div(nParts + 1, 64), 64>>>(row_d, nnz, nRow, nParts, b.p);
    B2_KERNEL_CHECK();
    B2_CUDA(cudaMemcpy(bounds_h, b.p, sizeof(int) * ((size_t)nParts + 1), cudaMemcpyDeviceToHost));
    for (int g = 1; g <= nParts; g++)
        if (bounds_h[g] < bounds_h[g - 1]) bounds_h[g] = bounds_h[g - 1];
    return B200SPMV_OK;
}

int b200spmv_partition_synth(int kind, long long p0, long long p1, int nParts, int *bounds_h)
{
    clear_error();
    if (nParts < 1 || !bounds_h) { set_error("partition_synth: bad argument"); return B200SPMV_ERR_INVALID; }
    long long nRowLL;
    if (kind == B200SPMV_SYNTH_LAP2D5) nRowLL = p0 * p0;
    else if (kind == B200SPMV_SYNTH_LAP3D7 || kind == B200SPMV_SYNTH_BOX3D27) nRowLL = p0 * p0 * p0;
    else if (kind == B200SPMV_SYNTH_UNIFORM) nRowLL = p0;
    else { set_error("partition_synth: kind %d has no closed-form row lengths (partition the generated COO with b200spmv_partition_rows)", kind); return B200SPMV_ERR_UNSUPPORTED; }
    if (p0 < 1 || nRowLL > 0x7fffffffLL) { set_error("partition_synth: size out of range"); return B200SPMV_ERR_INVALID; }
    const int nRow = (int)nRowLL;
    if (kind == B200SPMV_SYNTH_UNIFORM) {          // every row has p1 entries
        for (int g = 0; g <= nParts; g++) bounds_h[g] = (int)((long long)nRow * g / nParts);
        return B200SPMV_OK;
    }
    DevBuf<long long> cnt;
    DevBuf<int> b;
    B2_TRY(cnt.alloc((size_t)nRow + 1));
    B2_TRY(b.alloc((size_t)nParts + 1));
    stencil_len_kernel<<<ceil_div((long long)nRow + 1, 256), 256>>>(kind, (int)p0, nRow, cnt.p);
    B2_KERNEL_CHECK();
    B2_TRY(exclusive_scan_i64(cnt.p, cnt.p, nRow + 1, nullptr));
    count_search_kernel<<<ceil_div(nParts + 1, 64), 64>>>(cnt.p, nRow, nParts, b.p);
    B2_KERNEL_CHECK();
    B2_CUDA(cudaMemcpy(bounds_h, b.p, sizeof(int) * ((size_t)nParts + 1), cudaMemcpyDeviceToHost));
    bounds_h[0] = 0;
    return B200SPMV_OK;
}

int b200spmv_halo_plan(b200spmv_coo *coo, int colBegin, int colEnd, b200spmv_halo **out, void *stream)
{
    clear_error();
    cudaStream_t s = (cudaStream_t)stream;
    if (!coo || !out || colBegin < 0 || colEnd < colBegin || colEnd > coo->nCol) { set_error("halo_plan: bad argument"); return B200SPMV_ERR_INVALID; }
    *out = nullptr;
    std::unique_ptr<b200spmv_halo> h(new b200spmv_halo());
    h->rowBegin = coo->rowBegin; h->rowEnd = coo->rowEnd;
    h->colBegin = colBegin; h->colEnd = colEnd;
    h->nLocal = colEnd - colBegin;
    const long long nnz = coo->nnz;
    const int nLocalRows = coo->rowEnd - coo->rowBegin;
    if (nnz > 0x7fffffffLL) { set_error("halo_plan: block has more than 2^31-1 entries"); return B200SPMV_ERR_INVALID; }

    // 1. remote columns -> sorted unique list
    DevBuf<int> remote, sorted;
    DevBuf<long long> nsel;
    B2_TRY(remote.alloc((size_t)nnz));
    B2_TRY(nsel.alloc(1));
    long long nRemote = 0;
    if (nnz) {
        size_t tmp = 0;
        IsRemote pred{colBegin, colEnd};
        B2_CUDA(cub::DeviceSelect::If(nullptr, tmp, coo->col_d, remote.p, nsel.p, (int)nnz, pred, s));
        DevBuf<char> t;
        B2_TRY(t.alloc(tmp));
        B2_CUDA(cub::DeviceSelect::If(t.p, tmp, coo->col_d, remote.p, nsel.p, (int)nnz, pred, s));
        B2_CUDA(cudaMemcpyAsync(&nRemote, nsel.p, sizeof(long long), cudaMemcpyDeviceToHost, s));
        B2_CUDA(cudaStreamSynchronize(s));
    }
    long long nHalo = 0;
    if (nRemote) {
        B2_TRY(sorted.alloc((size_t)nRemote));
        size_t tmp = 0;
        B2_CUDA(cub::DeviceRadixSort::SortKeys(nullptr, tmp, remote.p, sorted.p, (int)nRemote, 0, 32, s));
        {
            DevBuf<char> t;
            B2_TRY(t.alloc(tmp));
            B2_CUDA(cub::DeviceRadixSort::SortKeys(t.p, tmp, remote.p, sorted.p, (int)nRemote, 0, 32, s));
            B2_CUDA(cudaStreamSynchronize(s));
        }
        B2_CUDA(cub::DeviceSelect::Unique(nullptr, tmp, sorted.p, remote.p, nsel.p, (int)nRemote, s));
        {
            DevBuf<char> t;
            B2_TRY(t.alloc(tmp));
            B2_CUDA(cub::DeviceSelect::Unique(t.p, tmp, sorted.p, remote.p, nsel.p, (int)nRemote, s));
            B2_CUDA(cudaMemcpyAsync(&nHalo, nsel.p, sizeof(long long), cudaMemcpyDeviceToHost, s));
            B2_CUDA(cudaStreamSynchronize(s));
        }
    }
    B2_TRY(h->halo_cols.alloc((size_t)nHalo));
    if (nHalo) B2_CUDA(cudaMemcpyAsync(h->halo_cols.p, remote.p, sizeof(int) * (size_t)nHalo, cudaMemcpyDeviceToDevice, s));
    // nLeft = halo columns below the owned range
    {
        std::vector<int> hc((size_t)nHalo);
        if (nHalo) B2_CUDA(cudaMemcpyAsync(hc.data(), h->halo_cols.p, sizeof(int) * (size_t)nHalo, cudaMemcpyDeviceToHost, s));
        B2_CUDA(cudaStreamSynchronize(s));
        int nl = 0;
        while (nl < (int)nHalo && hc[(size_t)nl] < colBegin) nl++;
        h->nLeft = nl;
        h->nRight = (int)nHalo - nl;
    }
    // 2. remap in place + interior range
    DevBuf<int> edge;
    B2_TRY(edge.alloc(2));
    int init[2] = {-1, nLocalRows};
    B2_CUDA(cudaMemcpyAsync(edge.p, init, sizeof init, cudaMemcpyHostToDevice, s));
    if (nnz) {
        halo_remap_kernel<<<ceil_div(nnz, 256), 256, 0, s>>>(coo->row_d, coo->col_d, nnz, coo->rowBegin, colBegin, colEnd,
                                                            h->halo_cols.p, (int)nHalo, h->nLeft, nLocalRows, edge.p);
        B2_KERNEL_CHECK();
    }
    int got[2];
    B2_CUDA(cudaMemcpyAsync(got, edge.p, sizeof got, cudaMemcpyDeviceToHost, s));
    B2_CUDA(cudaStreamSynchronize(s));
    h->interiorBegin = got[0] + 1;
    h->interiorEnd = got[1];
    if (h->interiorEnd < h->interiorBegin) h->interiorEnd = h->interiorBegin;
    coo->nRow = nLocalRows;
    coo->nCol = h->nLeft + h->nLocal + h->nRight;
    coo->rowBegin = 0;
    coo->rowEnd = nLocalRows;
    *out = h.release();
    return B200SPMV_OK;
}

int b200spmv_halo_info(const b200spmv_halo *h, long long *info8)
{
    if (!h || !info8) { set_error("halo_info: NULL argument"); return B200SPMV_ERR_INVALID; }
    info8[0] = h->nLocal; info8[1] = h->nLeft; info8[2] = h->nRight;
    info8[3] = h->interiorBegin; info8[4] = h->interiorEnd;
    info8[5] = h->nSend; info8[6] = h->rowBegin; info8[7] = h->rowEnd;
    return B200SPMV_OK;
}

long long b200spmv_halo_cols(const b200spmv_halo *h, int *cols_h, long long cap_bytes)
{
    if (!h) { set_error("halo_cols: NULL handle"); return B200SPMV_ERR_INVALID; }
    return export_device(h->halo_cols.p, h->halo_cols.bytes(), cols_h, cap_bytes);
}

int b200spmv_halo_set_send(b200spmv_halo *h, const int *send_cols_h, long long n)
{
    clear_error();
    if (!h || n < 0 || (n > 0 && !send_cols_h)) { set_error("halo_set_send: bad argument"); return B200SPMV_ERR_INVALID; }
    DevBuf<int> cols, bad;
    B2_TRY(cols.alloc((size_t)n));
    B2_TRY(bad.alloc(1));
    B2_TRY(h->send_idx.alloc((size_t)n));
    B2_CUDA(cudaMemset(bad.p, 0, sizeof(int)));
    if (n) {
        B2_CUDA(cudaMemcpy(cols.p, send_cols_h, sizeof(int) * (size_t)n, cudaMemcpyHostToDevice));
        send_idx_kernel<<<ceil_div(n, 256), 256>>>(cols.p, n, h->colBegin, h->colEnd, h->send_idx.p, bad.p);
        B2_KERNEL_CHECK();
    }
    int b = 0;
    B2_CUDA(cudaMemcpy(&b, bad.p, sizeof(int), cudaMemcpyDeviceToHost));
    if (b) { set_error("halo_set_send: a requested column is not owned by this block [%d,%d)", h->colBegin, h->colEnd); return B200SPMV_ERR_INVALID; }
    h->nSend = n;
    return B200SPMV_OK;
}

int b200spmv_halo_pack(const b200spmv_halo *h, const double *x_local_d, double *sendbuf_d, void *stream)
{
    if (!h) { set_error("halo_pack: NULL handle"); return B200SPMV_ERR_INVALID; }
    if (h->nSend == 0) return B200SPMV_OK;
    halo_pack_kernel<<<ceil_div(h->nSend, 256), 256, 0, (cudaStream_t)stream>>>(x_local_d, h->send_idx.p, h->nSend, sendbuf_d);
    B2_KERNEL_CHECK();
    return B200SPMV_OK;
}

int b200spmv_halo_free(b200spmv_halo *h)
{
    delete h;
    return B200SPMV_OK;
}

}  // extern "C"
