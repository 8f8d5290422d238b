// cusparse_cmp.cu -- COMPARISON POINT ONLY, never on the product path (libb200spmv.so does not link it).
// The reference's cuSPARSE plugin (/root/reference/src/opt_cusparse.cpp:57-83) calls cusparseDcsrmv, which
// cuSPARSE 12 no longer has; the same operation (general, zero-based, alpha = 1, beta = 0, non-transpose,
// opt_cusparse.cpp:75-81) is cusparseSpMV on a CSR descriptor.  Device-resident x and y, no per-call
// handle creation or PCIe copies (the reference leaks a handle and copies both vectors every call).
#include <cuda_runtime.h>
#include <cusparse.h>
#include <cstdio>

#define CMP_CUDA(e) do { cudaError_t e_ = (e); if (e_ != cudaSuccess) { snprintf(g_err, sizeof g_err, "%s: %s", #e, cudaGetErrorString(e_)); return -2; } } while (0)
#define CMP_SP(e) do { cusparseStatus_t s_ = (e); if (s_ != CUSPARSE_STATUS_SUCCESS) { snprintf(g_err, sizeof g_err, "%s: %s", #e, cusparseGetErrorString(s_)); return -3; } } while (0)
static char g_err[512];

__global__ void cmp_row_ptr_kernel(const int *__restrict__ row, int nnz, int nRow, int *__restrict__ ptr)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i > nnz) return;
    int prev = i == 0 ? -1 : row[i - 1];
    int cur = i == nnz ? nRow : row[i];
    for (int r = prev + 1; r <= cur; r++) ptr[r] = i;
}

extern "C" {
__attribute__((visibility("default"))) const char *b200cmp_last_error(void) { return g_err; }

// alg: 0 = CUSPARSE_SPMV_ALG_DEFAULT, 1 = CSR_ALG1, 2 = CSR_ALG2.  Times `iters` multiplies with CUDA events
// after `warmup` (and cusparseSpMV_preprocess); y_d receives the last result.  ms_out = mean per multiply.
__attribute__((visibility("default")))
int b200cmp_cusparse_csr(int nRow, int nCol, long long nnz, const int *row_d, const int *col_d, const double *val_d,
                         const double *x_d, double *y_d, int alg, int warmup, int iters, float *ms_out, void *stream)
{
    cudaStream_t s = (cudaStream_t)stream;
    int *ptr = nullptr;
    CMP_CUDA(cudaMalloc((void **)&ptr, sizeof(int) * ((size_t)nRow + 1)));
    cmp_row_ptr_kernel<<<(unsigned)((nnz + 1 + 255) / 256), 256, 0, s>>>(row_d, (int)nnz, nRow, ptr);
    CMP_CUDA(cudaGetLastError());
    cusparseHandle_t h;
    CMP_SP(cusparseCreate(&h));
    CMP_SP(cusparseSetStream(h, s));
    cusparseSpMatDescr_t A;
    cusparseDnVecDescr_t vx, vy;
    CMP_SP(cusparseCreateCsr(&A, nRow, nCol, nnz, ptr, (void *)col_d, (void *)val_d, CUSPARSE_INDEX_32I, CUSPARSE_INDEX_32I,
                             CUSPARSE_INDEX_BASE_ZERO, CUDA_R_64F));
    CMP_SP(cusparseCreateDnVec(&vx, nCol, (void *)x_d, CUDA_R_64F));
    CMP_SP(cusparseCreateDnVec(&vy, nRow, y_d, CUDA_R_64F));
    const double alpha = 1.0, beta = 0.0;
    const cusparseSpMVAlg_t a = alg == 1 ? CUSPARSE_SPMV_CSR_ALG1 : alg == 2 ? CUSPARSE_SPMV_CSR_ALG2 : CUSPARSE_SPMV_ALG_DEFAULT;
    size_t ws = 0;
    CMP_SP(cusparseSpMV_bufferSize(h, CUSPARSE_OPERATION_NON_TRANSPOSE, &alpha, A, vx, &beta, vy, CUDA_R_64F, a, &ws));
    void *buf = nullptr;
    CMP_CUDA(cudaMalloc(&buf, ws ? ws : 4));
    CMP_SP(cusparseSpMV_preprocess(h, CUSPARSE_OPERATION_NON_TRANSPOSE, &alpha, A, vx, &beta, vy, CUDA_R_64F, a, buf));
    for (int i = 0; i < warmup; i++)
        CMP_SP(cusparseSpMV(h, CUSPARSE_OPERATION_NON_TRANSPOSE, &alpha, A, vx, &beta, vy, CUDA_R_64F, a, buf));
    cudaEvent_t e0, e1;
    CMP_CUDA(cudaEventCreate(&e0));
    CMP_CUDA(cudaEventCreate(&e1));
    CMP_CUDA(cudaEventRecord(e0, s));
    for (int i = 0; i < iters; i++)
        CMP_SP(cusparseSpMV(h, CUSPARSE_OPERATION_NON_TRANSPOSE, &alpha, A, vx, &beta, vy, CUDA_R_64F, a, buf));
    CMP_CUDA(cudaEventRecord(e1, s));
    CMP_CUDA(cudaEventSynchronize(e1));
    float ms = 0;
    CMP_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    *ms_out = ms / (iters > 0 ? iters : 1);
    cusparseDestroySpMat(A);
    cusparseDestroyDnVec(vx);
    cusparseDestroyDnVec(vy);
    cusparseDestroy(h);
    cudaFree(buf);
    cudaFree(ptr);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return 0;
}
}
