// crs.cu -- CRS plugin: device conversion + adaptive tile-stream multiply.
// Reference: /root/reference/src/opt_crs.{h,cpp} (SpMatOpt{ptr,idx,val}; OptimizeProblem :10-42; SpMV :44-70).
#include "tile_stream.cuh"

namespace b2 {

// fp32 storage of the matrix values (options.value_f32): val is rounded to nearest once, at conversion, and
// widened back (exactly) in the kernel -- x, y and every product and sum stay fp64.  Cuts the stream from
// 12 to 8 B/nnz; y equals the fp64 path run on the rounded matrix bit for bit.
__global__ void crs_to_f32_kernel(const double *__restrict__ in, int n, float *__restrict__ out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = __double2float_rn(in[i]);
}
__global__ void crs_to_f64_kernel(const float *__restrict__ in, int n, double *__restrict__ out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (double)in[i];
}

struct CrsFormat : Format {
    DevBuf<int> ptr, idx;
    DevBuf<double> val;
    DevBuf<float> val32;
    bool f32;
    TileStream ts;
    explicit CrsFormat(const b200spmv_options &o) : f32(o.value_f32 != 0) {}

    int convert(const CooView &A, cudaStream_t s) override
    {
        nRow = A.nRow; nCol = A.nCol; nnz = A.nnz;
        B2_TRY(validate_sorted_coo(A, s));
        B2_TRY(ptr.alloc((size_t)nRow + 1));
        B2_TRY(idx.alloc((size_t)nnz));
        B2_TRY(build_row_ptr(A.row, nnz, nRow, ptr.p, s));                      // opt_crs.cpp:27-33
        B2_CUDA(cudaMemcpyAsync(idx.p, A.col, idx.bytes(), cudaMemcpyDeviceToDevice, s));   // :29
        if (f32) {
            val.release();
            B2_TRY(val32.alloc((size_t)nnz));
            if (nnz) crs_to_f32_kernel<<<ceil_div(nnz, 256), 256, 0, s>>>(A.val, nnz, val32.p);
            B2_KERNEL_CHECK();
            B2_TRY(ts.build(ptr.p, idx.p, val32.p, true, nRow, nnz, s));
        } else {
            B2_TRY(val.alloc((size_t)nnz));
            B2_CUDA(cudaMemcpyAsync(val.p, A.val, val.bytes(), cudaMemcpyDeviceToDevice, s));   // :30
            B2_TRY(ts.build(ptr.p, idx.p, val.p, false, nRow, nnz, s));
        }
        B2_CUDA(cudaStreamSynchronize(s));
        return B200SPMV_OK;
    }

    int multiply(const double *x, double *y, cudaStream_t s) override { return ts.run_all(x, y, false, s); }

    int multiply_rows(int rb, int re, const double *x, double *y, cudaStream_t s) override { return ts.run_rows(x, y, false, rb, re, s); }
    bool has_rows() const override { return true; }

    bool scalar(const std::string &n, long long *out) override
    {
        if (n == "alg_bytes") {   // 12 nnz + 4 (nRow+1) + 8 nCol + 8 nRow
            *out = (f32 ? 8LL : 12LL) * nnz + 4LL * (nRow + 1) + 8LL * nCol + 8LL * nRow;
            return true;
        }
        if (n == "launches") { *out = ts.nTiles > 1 ? 2 : 1; return true; }
        if (n == "nTiles") { *out = ts.nTiles; return true; }
        if (n == "value_f32") { *out = f32 ? 1 : 0; return true; }
        return false;
    }

    long long array(const std::string &n, void *dst, long long cap) override
    {
        if (n == "ptr") return export_device(ptr.p, ptr.bytes(), dst, cap);
        if (n == "idx") return export_device(idx.p, idx.bytes(), dst, cap);
        if (n == "val") {
            if (!f32) return export_device(val.p, val.bytes(), dst, cap);
            if (!dst) return (long long)(sizeof(double) * (size_t)nnz);
            DevBuf<double> w;
            if (w.alloc((size_t)nnz)) return B200SPMV_ERR_NOMEM;
            if (nnz) crs_to_f64_kernel<<<ceil_div(nnz, 256), 256>>>(val32.p, nnz, w.p);
            return export_device(w.p, w.bytes(), dst, cap);
        }
        return -1000;
    }
};

Format *make_crs(const b200spmv_options &o) { return new CrsFormat(o); }

}  // namespace b2
