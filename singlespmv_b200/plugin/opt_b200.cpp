// opt_b200.cpp -- OptimizeProblem / SpMV of the B200 plugin (see opt_b200.h).  Host code only: every
// device operation goes through the C-ABI of libb200spmv.so; cudart is used for the vector buffers.
// Error behaviour follows the reference's plugins: print and exit (src/util.h:48-55 CUDA_SAFE_CALL,
// src/util.cpp:32-35); the C-ABI underneath returns status codes and never exits.
#include "opt_b200.h"
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime_api.h>

#ifndef B200_FORMAT
#define B200_FORMAT CRS
#endif
#define B200_CAT_(a, b) a##b
#define B200_CAT(a, b) B200_CAT_(a, b)
#define B200_STR_(a) #a
#define B200_STR(a) B200_STR_(a)
#define B200_FORMAT_ENUM B200_CAT(B200SPMV_, B200_FORMAT)
#ifndef SEGMENT_WIDTH
#define SEGMENT_WIDTH 0          /* 0 = library default = ALIGNMENT/8 with ALIGNMENT=32 (src/param.h:9-11) */
#endif
#ifndef N_BLOCK
#define N_BLOCK 0                /* 0 = library default 1 (src/param.h:18-20) */
#endif
#ifndef B200_SIGMA
#define B200_SIGMA 0
#endif

static void b200_check (int status, const char *what) {
    if (status == B200SPMV_OK) return;
    fprintf(stderr, "b200spmv: %s failed (%d): %s\n", what, status, b200spmv_last_error());
    exit(1);
}
static void b200_cuda (cudaError_t e, const char *what) {
    if (e == cudaSuccess) return;
    fprintf(stderr, "b200spmv: %s: %s\n", what, cudaGetErrorString(e));
    exit(1);
}
const char *B200FormatName () { return B200_STR(B200_FORMAT); }

void OptimizeProblem (const SpMat &A, const Vec &x, SpMatOpt &A_opt, VecOpt &x_opt) {
    x_opt.size = x.size;
    x_opt.val = x.val;
    A_opt.nRow = A.nRow;
    A_opt.nCol = A.nCol;
    A_opt.nNnz = A.nNnz;
    A_opt.x_dev = A_opt.y_dev = NULL;
    A_opt.stream = NULL;
    b200spmv_options opt = b200spmv_options();
    opt.segment_width = SEGMENT_WIDTH;
    opt.n_block = N_BLOCK;
    opt.csr5_sigma = B200_SIGMA;
#ifdef B200_VALUE_F32
    opt.value_f32 = 1;              // CRS only: fp32 storage of the matrix values, fp64 arithmetic
#endif
    b200_check(b200spmv_create(B200_FORMAT_ENUM, &opt, &A_opt.handle), "create");
    b200_check(b200spmv_convert_coo_host(A_opt.handle, A.nRow, A.nCol, A.nNnz, A.row_idx, A.col_idx, A.val), "convert");
#ifdef B200_DEVICE_RESIDENT
    b200_cuda(cudaMalloc((void **)&A_opt.x_dev, sizeof(double) * (A.nCol > 0 ? A.nCol : 1)), "cudaMalloc x");
    b200_cuda(cudaMalloc((void **)&A_opt.y_dev, sizeof(double) * (A.nRow > 0 ? A.nRow : 1)), "cudaMalloc y");
    B200UploadVector(A_opt, x_opt);
#endif
}

extern "C" {
void SpMV (const SpMatOpt &A, const VecOpt &x, Vec &y) {
#ifdef B200_DEVICE_RESIDENT
    (void)x; (void)y;
    b200_check(b200spmv_multiply(A.handle, A.x_dev, A.y_dev, A.stream), "multiply");
#else
    b200_check(b200spmv_multiply_host(A.handle, x.val, y.val), "multiply");
#endif
}
}

void B200UploadVector (const SpMatOpt &A, const VecOpt &x) {
    if (!A.x_dev) return;
    b200_cuda(cudaMemcpy(A.x_dev, x.val, sizeof(double) * A.nCol, cudaMemcpyHostToDevice), "H2D x");
}
void B200FetchResult (const SpMatOpt &A, Vec &y) {
    if (!A.y_dev) return;
    b200_cuda(cudaStreamSynchronize((cudaStream_t)A.stream), "sync");
    b200_cuda(cudaMemcpy(y.val, A.y_dev, sizeof(double) * A.nRow, cudaMemcpyDeviceToHost), "D2H y");
}
void B200Synchronize (const SpMatOpt &A) { b200_cuda(cudaStreamSynchronize((cudaStream_t)A.stream), "sync"); }
long long B200Scalar (const SpMatOpt &A, const char *name) {
    long long v = 0;
    b200_check(b200spmv_get_scalar(A.handle, name, &v), name);
    return v;
}
