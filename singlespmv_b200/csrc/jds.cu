// jds.cu -- JDS plugin (/root/reference/src/opt_jds.{h,cpp}): rows permuted by non-increasing length
// (opt_jds.cpp:41-46), jagged diagonal c = c-th entry of every row that has one, stored one after the
// other (ptr[c], :47-59); multiply walks down the diagonals of permuted row r (:91-103).
//
// The reference's layout is already what a GPU wants: position r of every diagonal is handled by
// thread r, so each warp request is one contiguous run.  perm comes from a stable device radix sort
// (ties by ascending row); the reference's std::sort is unstable, so its tie order is whatever
// libstdc++ does -- b200spmv_jds_set_perm_host() imposes a given order for bit-exact array parity.
#include <cub/cub.cuh>

#include "colblocks.cuh"
#include "common.cuh"

namespace b2 {

constexpr int JDS_LONG = 2048;   // rows longer than this get a whole warp

__global__ void jds_length_kernel(const int *__restrict__ ptr, int nRow, int *__restrict__ length, int *__restrict__ iota)
{
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nRow) return;
    length[r] = ptr[r + 1] - ptr[r];          // opt_jds.cpp:37
    iota[r] = r;
}

__global__ void jds_gather_len_kernel(const int *__restrict__ length, const int *__restrict__ perm, int nRow,
                                      int *__restrict__ slen, int *__restrict__ bad)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nRow) return;
    const int r = perm[i];
    if (r < 0 || r >= nRow) { atomicExch(bad, 1); slen[i] = 0; return; }
    slen[i] = length[r];
}

// seen[perm[i]]++ and order check for a caller-supplied permutation
__global__ void jds_check_perm_kernel(const int *__restrict__ perm, const int *__restrict__ slen, int nRow,
                                      int *__restrict__ seen, int *__restrict__ bad)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nRow) return;
    const int r = perm[i];
    if (r < 0 || r >= nRow) { atomicExch(bad, 1); return; }
    if (atomicAdd(&seen[r], 1) != 0) atomicExch(bad, 1);
    if (i > 0 && slen[i - 1] < slen[i]) atomicExch(bad, 1);
}

// cnt[c] = number of rows longer than c.  slen is non-increasing, so position i closes the counts
// c in [slen[i+1], slen[i]): every c is written exactly once.
__global__ void jds_diag_count_kernel(const int *__restrict__ slen, int nRow, int maxLength, int *__restrict__ cnt)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i > nRow) return;
    if (i == nRow) { cnt[maxLength] = 0; return; }
    const int hi = slen[i], lo = i + 1 < nRow ? slen[i + 1] : 0;
    for (int c = lo; c < hi; c++) cnt[c] = i + 1;
}

// jcol/jval[jptr[c] + i] = c-th entry of row perm[i]      (opt_jds.cpp:48-58)
__global__ void jds_fill_kernel(const int *__restrict__ ptr, const int *__restrict__ col, const double *__restrict__ val,
                                const int *__restrict__ perm, const int *__restrict__ jptr, int nRow,
                                int *__restrict__ jcol, double *__restrict__ jval)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nRow) return;
    const int r = perm[i], b = ptr[r], len = ptr[r + 1] - b;
    for (int c = 0; c < len; c++) {
        const int at = jptr[c] + i;
        jcol[at] = col[b + c];
        jval[at] = val[b + c];
    }
}

// One thread per position; the length of position i is implied by the diagonal sizes
// (i < jptr[c+1]-jptr[c]), so neither length[] nor a sorted copy of it is read.  Ascending c with
// unfused mul/add = the reference's order = opt_crs.cpp's order -> bit-identical y.
__global__ void __launch_bounds__(256)
jds_spmv_kernel(const int *__restrict__ jptr, const int *__restrict__ jcol, const double *__restrict__ jval,
                const int *__restrict__ perm, const double *__restrict__ x, double *__restrict__ y, int nRow,
                int maxLength, int firstPos)
{
    const int i = firstPos + blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nRow) return;
    const uint64_t pol_stream = policy_evict_first(), pol_x = policy_evict_last();
    double acc = 0.0;
    int c = 0;
    // four diagonals in flight while position i is inside all of them (diagonals only get shorter)
    for (; c + 4 <= maxLength; c += 4) {
        const int b0 = jptr[c], b1 = jptr[c + 1], b2 = jptr[c + 2], b3 = jptr[c + 3], b4 = jptr[c + 4];
        if (i >= b4 - b3) break;
        const int c0 = ld_stream_i1(jcol + b0 + i, pol_stream), c1 = ld_stream_i1(jcol + b1 + i, pol_stream);
        const int c2 = ld_stream_i1(jcol + b2 + i, pol_stream), c3 = ld_stream_i1(jcol + b3 + i, pol_stream);
        const double v0 = ld_stream_d1(jval + b0 + i, pol_stream), v1 = ld_stream_d1(jval + b1 + i, pol_stream);
        const double v2 = ld_stream_d1(jval + b2 + i, pol_stream), v3 = ld_stream_d1(jval + b3 + i, pol_stream);
        const double x0 = ld_x(x + c0, pol_x), x1 = ld_x(x + c1, pol_x), x2 = ld_x(x + c2, pol_x), x3 = ld_x(x + c3, pol_x);
        acc = __dadd_rn(acc, __dmul_rn(v0, x0));
        acc = __dadd_rn(acc, __dmul_rn(v1, x1));
        acc = __dadd_rn(acc, __dmul_rn(v2, x2));
        acc = __dadd_rn(acc, __dmul_rn(v3, x3));
    }
    for (; c < maxLength; c++) {
        const int b = jptr[c];
        if (i >= jptr[c + 1] - b) break;
        acc = __dadd_rn(acc, __dmul_rn(ld_stream_d1(jval + b + i, pol_stream), ld_x(x + ld_stream_i1(jcol + b + i, pol_stream), pol_x)));
    }
    y[perm[i]] = acc;
}

// positions [0, nLong): very long rows (power-law heads), one warp each, lanes stride the diagonals
__global__ void __launch_bounds__(256)
jds_spmv_long_kernel(const int *__restrict__ jptr, const int *__restrict__ jcol, const double *__restrict__ jval,
                     const int *__restrict__ perm, const double *__restrict__ x, double *__restrict__ y, int nLong,
                     int maxLength)
{
    const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (i >= nLong) return;
    double acc = 0.0;
    for (int c = lane; c < maxLength; c += 32) {
        const int b = jptr[c];
        if (i >= jptr[c + 1] - b) break;
        acc += jval[b + i] * x[jcol[b + i]];
    }
    acc = warp_sum(acc);
    if (lane == 0) y[perm[i]] = acc;
}

struct JdsFormat : Format {
    int maxLength = 0, nLong = 0;
    DevBuf<int> perm, length, jptr, jcol;
    DevBuf<double> jval;
    std::vector<int> user_perm;
    bool have_user_perm = false;
    std::unique_ptr<ColBlockEngine> cb;   // column-blocked multiply layout when x does not fit L2 (colblocks.cuh)
    int cbs_want = 0;
    explicit JdsFormat(const b200spmv_options &o) : cbs_want(o.col_blocks) {}

    int set_perm(const int *perm_h, int n) override
    {
        user_perm.assign(perm_h, perm_h + n);
        have_user_perm = true;
        return B200SPMV_OK;
    }

    int convert(const CooView &A, cudaStream_t s) override
    {
        nRow = A.nRow; nCol = A.nCol; nnz = A.nnz;
        B2_TRY(validate_sorted_coo(A, s));
        if (have_user_perm && (int)user_perm.size() != nRow) {
            set_error("JDS: the imposed permutation has %zu entries, the matrix %d rows", user_perm.size(), nRow);
            return B200SPMV_ERR_INVALID;
        }
        DevBuf<int> ptr, iota, slen, cnt, flag;
        B2_TRY(ptr.alloc((size_t)nRow + 1));
        B2_TRY(build_row_ptr(A.row, nnz, nRow, ptr.p, s));
        B2_TRY(max_row_length(ptr.p, nRow, &maxLength, s));
        B2_TRY(perm.alloc((size_t)nRow));
        B2_TRY(length.alloc((size_t)nRow));
        B2_TRY(iota.alloc((size_t)nRow));
        B2_TRY(slen.alloc((size_t)nRow));
        B2_TRY(jptr.alloc((size_t)maxLength + 1));
        B2_TRY(cnt.alloc((size_t)maxLength + 1));
        B2_TRY(jcol.alloc((size_t)nnz));
        B2_TRY(jval.alloc((size_t)nnz));
        B2_TRY(flag.alloc(1));
        B2_CUDA(cudaMemsetAsync(flag.p, 0, sizeof(int), s));
        const int gb = ceil_div(nRow, 256);
        if (nRow) {
            jds_length_kernel<<<gb, 256, 0, s>>>(ptr.p, nRow, length.p, iota.p);
            B2_KERNEL_CHECK();
        }
        if (have_user_perm) {
            DevBuf<int> seen;
            B2_TRY(seen.alloc((size_t)nRow));
            B2_CUDA(cudaMemsetAsync(seen.p, 0, seen.bytes(), s));
            if (nRow) {
                B2_CUDA(cudaMemcpyAsync(perm.p, user_perm.data(), sizeof(int) * (size_t)nRow, cudaMemcpyHostToDevice, s));
                jds_gather_len_kernel<<<gb, 256, 0, s>>>(length.p, perm.p, nRow, slen.p, flag.p);
                jds_check_perm_kernel<<<gb, 256, 0, s>>>(perm.p, slen.p, nRow, seen.p, flag.p);
                B2_KERNEL_CHECK();
            }
            int bad = 0;
            B2_CUDA(cudaMemcpyAsync(&bad, flag.p, sizeof(int), cudaMemcpyDeviceToHost, s));
            B2_CUDA(cudaStreamSynchronize(s));
            if (bad) {
                set_error("JDS: the imposed permutation is not a permutation ordering rows by non-increasing length");
                return B200SPMV_ERR_INVALID;
            }
        } else if (nRow) {
            // stable descending radix sort on the row length: ties keep ascending row order
            int bits = 1;
            while (bits < 31 && (1 << bits) <= maxLength) bits++;
            size_t tmp = 0;
            B2_CUDA(cub::DeviceRadixSort::SortPairsDescending(nullptr, tmp, length.p, slen.p, iota.p, perm.p, nRow, 0, bits, s));
            DevBuf<char> t;
            B2_TRY(t.alloc(tmp));
            B2_CUDA(cub::DeviceRadixSort::SortPairsDescending(t.p, tmp, length.p, slen.p, iota.p, perm.p, nRow, 0, bits, s));
            B2_CUDA(cudaStreamSynchronize(s));
        }
        jds_diag_count_kernel<<<ceil_div((long long)nRow + 1, 256), 256, 0, s>>>(slen.p, nRow, maxLength, cnt.p);
        B2_KERNEL_CHECK();
        B2_TRY(exclusive_scan_i32(cnt.p, jptr.p, maxLength + 1, s));       // opt_jds.cpp:48-59
        if (nRow) {
            jds_fill_kernel<<<gb, 256, 0, s>>>(ptr.p, A.col, A.val, perm.p, jptr.p, nRow, jcol.p, jval.p);
            B2_KERNEL_CHECK();
        }
        B2_TRY(make_col_block_engine(A, ptr.p, cbs_want, s, &cb));
        nLong = 0;
        if (maxLength > JDS_LONG)
            B2_CUDA(cudaMemcpy(&nLong, cnt.p + JDS_LONG, sizeof(int), cudaMemcpyDeviceToHost));
        B2_CUDA(cudaStreamSynchronize(s));
        return B200SPMV_OK;
    }

    int multiply(const double *x, double *y, cudaStream_t s) override
    {
        if (nRow == 0) return B200SPMV_OK;
        // gather-bound matrices: same rows, same ascending-column sums, column block by column block (colblocks.cuh); the
        // permutation only orders the reference arrays, y[row] does not depend on it
        if ((cb != nullptr)) return cb->run(x, y, 0, nRow, s);
        if (nLong > 0) {
            jds_spmv_long_kernel<<<ceil_div((long long)nLong * 32, 256), 256, 0, s>>>(jptr.p, jcol.p, jval.p, perm.p, x, y, nLong, maxLength);
            B2_KERNEL_CHECK();
        }
        if (nRow > nLong) {
            jds_spmv_kernel<<<ceil_div(nRow - nLong, 256), 256, 0, s>>>(jptr.p, jcol.p, jval.p, perm.p, x, y, nRow, maxLength, nLong);
            B2_KERNEL_CHECK();
        }
        return B200SPMV_OK;
    }

    bool scalar(const std::string &n, long long *out) override
    {
        if (n == "maxLength") { *out = maxLength; return true; }
        if (n == "nLong") { *out = nLong; return true; }
        if (n == "alg_bytes") {   // 12 nnz + 4 (maxLength+1) + 4 nRow (perm) + 8 nCol + 8 nRow
            *out = 12LL * nnz + 4LL * (maxLength + 1) + 4LL * nRow + 8LL * nCol + 8LL * nRow;
            return true;
        }
        if (n == "launches") { *out = (cb != nullptr) ? cb->n_blocks() : (nLong > 0) + (nRow > nLong); return true; }
        if (n == "col_blocks") { *out = (cb != nullptr) ? cb->n_blocks() : 0; return true; }
        if (n == "col_block_engine") { *out = (cb == nullptr) ? 0 : (cb->name()[0] == 'e' ? 1 : 2); return true; }   // 1 = sliced ELL per block, 2 = tile-stream
        return false;
    }

    long long array(const std::string &n, void *dst, long long cap) override
    {
        if (n == "perm") return export_device(perm.p, perm.bytes(), dst, cap);
        if (n == "length") return export_device(length.p, length.bytes(), dst, cap);
        if (n == "ptr") return export_device(jptr.p, jptr.bytes(), dst, cap);
        if (n == "col_idx") return export_device(jcol.p, jcol.bytes(), dst, cap);
        if (n == "val") return export_device(jval.p, jval.bytes(), dst, cap);
        return -1000;
    }
};

Format *make_jds(const b200spmv_options &o) { return new JdsFormat(o); }

}  // namespace b2
