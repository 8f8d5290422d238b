// crs.cu -- CRS plugin: device conversion + adaptive tile-stream multiply.
// Reference: /root/reference/src/opt_crs.{h,cpp} (SpMatOpt{ptr,idx,val}; OptimizeProblem :10-42; SpMV :44-70).
#include "tile_stream.cuh"

namespace b2 {

struct CrsFormat : Format {
    DevBuf<int> ptr, idx;
    DevBuf<double> val;
    TileStream ts;

    int convert(const CooView &A, cudaStream_t s) override
    {
        nRow = A.nRow; nCol = A.nCol; nnz = A.nnz;
        B2_TRY(validate_sorted_coo(A, s));
        B2_TRY(ptr.alloc((size_t)nRow + 1));
        B2_TRY(idx.alloc((size_t)nnz));
        B2_TRY(val.alloc((size_t)nnz));
        B2_TRY(build_row_ptr(A.row, nnz, nRow, ptr.p, s));                      // opt_crs.cpp:27-33
        B2_CUDA(cudaMemcpyAsync(idx.p, A.col, idx.bytes(), cudaMemcpyDeviceToDevice, s));   // :29
        B2_CUDA(cudaMemcpyAsync(val.p, A.val, val.bytes(), cudaMemcpyDeviceToDevice, s));   // :30
        B2_TRY(ts.build(ptr.p, idx.p, val.p, nRow, nnz, s));
        B2_CUDA(cudaStreamSynchronize(s));
        return B200SPMV_OK;
    }

    int multiply(const double *x, double *y, cudaStream_t s) override { return ts.run_all(x, y, false, s); }

    int multiply_rows(int rb, int re, const double *x, double *y, cudaStream_t s) override { return ts.run_rows(x, y, false, rb, re, s); }
    bool has_rows() const override { return true; }

    bool scalar(const std::string &n, long long *out) override
    {
        if (n == "alg_bytes") {   // 12 nnz + 4 (nRow+1) + 8 nCol + 8 nRow
            *out = 12LL * nnz + 4LL * (nRow + 1) + 8LL * nCol + 8LL * nRow;
            return true;
        }
        if (n == "launches") { *out = ts.nTiles > 1 ? 2 : 1; return true; }
        if (n == "nTiles") { *out = ts.nTiles; return true; }
        return false;
    }

    long long array(const std::string &n, void *dst, long long cap) override
    {
        if (n == "ptr") return export_device(ptr.p, ptr.bytes(), dst, cap);
        if (n == "idx") return export_device(idx.p, idx.bytes(), dst, cap);
        if (n == "val") return export_device(val.p, val.bytes(), dst, cap);
        return -1000;
    }
};

Format *make_crs(const b200spmv_options &) { return new CrsFormat(); }

}  // namespace b2
