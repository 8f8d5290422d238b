// hyb.cu -- HYB = ELL + COO tail.  The reference tree only names the format: ANONYMOUSLIB_FORMAT_HYB5
// is declared and never implemented (opt/Benchmark_SpMV_using_CSR5/CSR5_cuda/detail/common.h:22).  It is the classic
// hybrid of Bell & Garland: the first K entries of every row go into an ELL part of width K, whatever is left of the
// longer rows into a COO tail -- ELL's pure streaming for the bulk of a matrix whose rows are nearly uniform, without
// ELL's padding blow-up on the few long rows.
//   K            the largest width that at least max(4096, nRow / 3) rows still fill (Bell & Garland's rule), or
//                options.hyb_k
//   ELL part     sliced ELL exactly as ell.cu stores it (slice-local width, 128-bit loads); exported as [nRow][K]
//                with the reference's padding rule col = k, val = 0 (src/opt_ell.cpp:46-52)
//   COO tail     entries K, K+1, ... of the rows longer than K, sorted by (row, col) like the input
// Multiply: the ELL kernel writes y, the COO kernel continues every row's running sum with its tail entries
// (coo_accumulate): ascending column order throughout, so rows whose tail has at most 64 entries are bit-identical to
// the reference CRS result; longer tails are reduced by a warp (within the 1e-12 tolerance).
#include "common.cuh"

namespace b2 {

int coo_accumulate(const int *row, const int *col, const double *val, int nnz, int nRow, const double *x, double *y,
                   double *carry, cudaStream_t s);
int coo_tile_entries();

// hist[k] += rows of length exactly k (k clipped to cap)
__global__ void hyb_hist_kernel(const int *__restrict__ ptr, int nRow, int cap, unsigned *__restrict__ hist)
{
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r < nRow) atomicAdd(&hist[min(ptr[r + 1] - ptr[r], cap)], 1u);
}
// tail offsets: tcnt[r] = max(0, len - K)
__global__ void hyb_tail_count_kernel(const int *__restrict__ ptr, int nRow, int K, int *__restrict__ tcnt)
{
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r > nRow) return;
    tcnt[r] = r < nRow ? max(0, ptr[r + 1] - ptr[r] - K) : 0;
}
// head (first K entries of every row) and tail triplets
__global__ void hyb_split_kernel(const int *__restrict__ ptr, const int *__restrict__ col, const double *__restrict__ val,
                                 int nRow, int K, const int *__restrict__ hptr, const int *__restrict__ tptr,
                                 int *__restrict__ hrow, int *__restrict__ hcol, double *__restrict__ hval,
                                 int *__restrict__ trow, int *__restrict__ tcol, double *__restrict__ tval)
{
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nRow) return;
    const int b = ptr[r], len = ptr[r + 1] - b, nh = min(len, K);
    for (int k = 0; k < nh; k++) {
        hrow[hptr[r] + k] = r;
        hcol[hptr[r] + k] = col[b + k];
        hval[hptr[r] + k] = val[b + k];
    }
    for (int k = nh; k < len; k++) {
        trow[tptr[r] + k - nh] = r;
        tcol[tptr[r] + k - nh] = col[b + k];
        tval[tptr[r] + k - nh] = val[b + k];
    }
}
__global__ void hyb_head_count_kernel(const int *__restrict__ ptr, int nRow, int K, int *__restrict__ hcnt)
{
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r > nRow) return;
    hcnt[r] = r < nRow ? min(K, ptr[r + 1] - ptr[r]) : 0;
}

struct HybFormat : Format {
    int K = 0, tailNnz = 0, k_opt;
    std::unique_ptr<Format> ell;          // the ELL part (ell.cu), built from the head triplets
    DevBuf<int> trow, tcol;
    DevBuf<double> tval, carry;
    b200spmv_options ell_opt;
    explicit HybFormat(const b200spmv_options &o) : k_opt(o.hyb_k), ell_opt(o) { ell_opt.col_blocks = -1; }

    int convert(const CooView &A, cudaStream_t s) override
    {
        nRow = A.nRow; nCol = A.nCol; nnz = A.nnz;
        B2_TRY(validate_sorted_coo(A, s));
        DevBuf<int> ptr, hptr, tptr;
        B2_TRY(ptr.alloc((size_t)nRow + 1));
        B2_TRY(hptr.alloc((size_t)nRow + 1));
        B2_TRY(tptr.alloc((size_t)nRow + 1));
        B2_TRY(build_row_ptr(A.row, nnz, nRow, ptr.p, s));
        int maxLen = 0;
        B2_TRY(max_row_length(ptr.p, nRow, &maxLen, s));
        K = k_opt > 0 ? std::min(k_opt, maxLen) : 0;
        if (k_opt <= 0 && nRow > 0 && maxLen > 0) {
            // rowsWith(k) = rows with at least k entries; K = largest k with rowsWith(k) >= max(4096, nRow / 3)
            const int cap = std::min(maxLen, 4096);
            DevBuf<unsigned> hist;
            B2_TRY(hist.alloc((size_t)cap + 1));
            B2_CUDA(cudaMemsetAsync(hist.p, 0, hist.bytes(), s));
            hyb_hist_kernel<<<ceil_div(nRow, 256), 256, 0, s>>>(ptr.p, nRow, cap, hist.p);
            B2_KERNEL_CHECK();
            std::vector<unsigned> h((size_t)cap + 1);
            B2_CUDA(cudaMemcpyAsync(h.data(), hist.p, hist.bytes(), cudaMemcpyDeviceToHost, s));
            B2_CUDA(cudaStreamSynchronize(s));
            const long long need = std::min<long long>(nRow, std::max<long long>(4096, nRow / 3));
            long long rowsWith = 0;
            K = 0;
            for (int k = cap; k >= 1; k--) {
                rowsWith += h[(size_t)k];
                if (rowsWith >= need) { K = k; break; }
            }
            if (K == 0) K = 1;
        }
        const int gb = ceil_div((long long)nRow + 1, 256);
        hyb_head_count_kernel<<<gb, 256, 0, s>>>(ptr.p, nRow, K, hptr.p);
        hyb_tail_count_kernel<<<gb, 256, 0, s>>>(ptr.p, nRow, K, tptr.p);
        B2_KERNEL_CHECK();
        B2_TRY(exclusive_scan_i32(hptr.p, hptr.p, nRow + 1, s));
        B2_TRY(exclusive_scan_i32(tptr.p, tptr.p, nRow + 1, s));
        int headNnz = 0;
        B2_CUDA(cudaMemcpy(&headNnz, hptr.p + nRow, sizeof(int), cudaMemcpyDeviceToHost));
        B2_CUDA(cudaMemcpy(&tailNnz, tptr.p + nRow, sizeof(int), cudaMemcpyDeviceToHost));
        DevBuf<int> hrow, hcol;
        DevBuf<double> hval;
        B2_TRY(hrow.alloc((size_t)headNnz));
        B2_TRY(hcol.alloc((size_t)headNnz));
        B2_TRY(hval.alloc((size_t)headNnz));
        B2_TRY(trow.alloc((size_t)tailNnz));
        B2_TRY(tcol.alloc((size_t)tailNnz));
        B2_TRY(tval.alloc((size_t)tailNnz));
        B2_TRY(carry.alloc((size_t)ceil_div(tailNnz, coo_tile_entries()) + 1));
        if (nRow) {
            hyb_split_kernel<<<ceil_div(nRow, 256), 256, 0, s>>>(ptr.p, A.col, A.val, nRow, K, hptr.p, tptr.p, hrow.p, hcol.p,
                                                                hval.p, trow.p, tcol.p, tval.p);
            B2_KERNEL_CHECK();
        }
        ell.reset(make_ell(ell_opt));
        CooView H{nRow, nCol, headNnz, hrow.p, hcol.p, hval.p};
        B2_TRY(ell->convert(H, s));
        B2_CUDA(cudaStreamSynchronize(s));
        return B200SPMV_OK;
    }

    int multiply(const double *x, double *y, cudaStream_t s) override
    {
        if (nRow == 0) return B200SPMV_OK;
        B2_TRY(ell->multiply(x, y, s));
        return coo_accumulate(trow.p, tcol.p, tval.p, tailNnz, nRow, x, y, carry.p, s);
    }

    bool scalar(const std::string &n, long long *out) override
    {
        if (n == "K") { *out = K; return true; }
        if (n == "tail_nnz") { *out = tailNnz; return true; }
        if (n == "alg_bytes") {   // 12 B per stored ELL slot + 16 B per tail entry + x + y
            long long slots = 0, nsl = 0;
            ell->scalar("slots", &slots);
            nsl = (nRow + 31) / 32;
            *out = 12LL * slots + 8LL * (nsl + 1) + 16LL * tailNnz + 8LL * nCol + 8LL * nRow;
            return true;
        }
        if (n == "launches") { *out = 1 + (tailNnz > 0 ? (tailNnz > coo_tile_entries() ? 2 : 1) : 0); return true; }
        return false;
    }

    long long array(const std::string &n, void *dst, long long cap) override
    {
        if (n == "ell_col_idx") return ell->array("col_idx", dst, cap);
        if (n == "ell_val") return ell->array("val", dst, cap);
        if (n == "coo_row_idx") return export_device(trow.p, trow.bytes(), dst, cap);
        if (n == "coo_col_idx") return export_device(tcol.p, tcol.bytes(), dst, cap);
        if (n == "coo_val") return export_device(tval.p, tval.bytes(), dst, cap);
        return -1000;
    }
};

Format *make_hyb(const b200spmv_options &o) { return new HybFormat(o); }

}  // namespace b2
