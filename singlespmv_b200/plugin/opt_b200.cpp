// opt_b200.cpp -- OptimizeProblem / SpMV of the B200 plugin (see opt_b200.h).  Host code only: every
// device operation goes through the C-ABI of libb200spmv.so; cudart is used for the vector buffers.
// Error behaviour follows the reference's plugins: print and exit (src/util.h:48-55 CUDA_SAFE_CALL,
// src/util.cpp:32-35); the C-ABI underneath returns status codes and never exits.
#include "opt_b200.h"
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime_api.h>

// report globals the SS / CSS plugins fill for the driver (defined in the host project's util.cpp:16-18,
// read at src/main.cpp:157-174 and :83-98)
extern std::vector<int> g_step_count;
extern std::vector<double> g_step_time;
extern std::vector<double> g_profile;

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
    A_opt.handle = NULL;
    A_opt.mg = NULL;
    A_opt.nGPU = 1;
    b200spmv_options opt = b200spmv_options();
    opt.segment_width = SEGMENT_WIDTH;
    opt.n_block = N_BLOCK;
    opt.csr5_sigma = B200_SIGMA;
#ifdef B200_VALUE_F32
    opt.value_f32 = 1;              // CRS only: fp32 storage of the matrix values, fp64 arithmetic
#endif
#ifdef B200_SS_FAITHFUL
    opt.ss_faithful = 1;            // SS / CSS: the reference's three-phase Mul / fold / gather schedule
#ifdef PROFILING
    opt.profile = 1;                // per-phase times -> g_profile[0] (Mul), g_profile[1] (Sum), src/opt_ss.cpp:225-304
#endif
#endif
#ifdef B200_NGPU
    {   // one process, several GPUs: row blocks by non-zero balance, x halo pulled over NVLink (include/b200spmv.h, mg)
        int n = B200_NGPU;
        if (getenv("B200_NGPU")) n = atoi(getenv("B200_NGPU"));
        if (n <= 0) b200_check(b200spmv_device_count(&n), "device_count");
        A_opt.nGPU = n;
        b200_check(b200spmv_mg_create(n, B200_FORMAT_ENUM, &opt, &A_opt.mg), "mg_create");
        b200_check(b200spmv_mg_convert_coo_host(A_opt.mg, A.nRow, A.nCol, A.nNnz, A.row_idx, A.col_idx, A.val), "mg_convert");
#ifdef B200_DEVICE_RESIDENT
        B200UploadVector(A_opt, x_opt);
#else
        if (!getenv("B200_NO_HOST_REGISTER") && x.size > 0) b200spmv_host_register(x.val, sizeof(double) * (unsigned long long)x.size);
#endif
        return;
    }
#endif
    b200_check(b200spmv_create(B200_FORMAT_ENUM, &opt, &A_opt.handle), "create");
    b200_check(b200spmv_convert_coo_host(A_opt.handle, A.nRow, A.nCol, A.nNnz, A.row_idx, A.col_idx, A.val), "convert");
    if (B200_FORMAT_ENUM == B200SPMV_SS && A_opt.handle) {          // src/opt_ss.cpp:143-147: one count per fold step for the report
        const long long bytes = b200spmv_get_array(A_opt.handle, "sum_segs_count", NULL, 0);
        g_step_count.assign(bytes > 0 ? (size_t)bytes / sizeof(int) : 0, 0);
        if (bytes > 0) b200spmv_get_array(A_opt.handle, "sum_segs_count", g_step_count.data(), bytes);
        g_step_time.assign(g_step_count.size(), 0.0);
    }
#ifndef B200_DEVICE_RESIDENT
    // the driver's vectors are pageable (_mm_malloc, src/util.cpp:92-102): page-lock x now and y on the first SpMV so
    // that the copies inside every SpMV are asynchronous DMA that overlaps the multiply.  Never fatal.
    if (!getenv("B200_NO_HOST_REGISTER") && x.size > 0 &&
        b200spmv_host_register(x.val, sizeof(double) * (unsigned long long)x.size) != B200SPMV_OK)
        fprintf(stderr, "b200spmv: x stays pageable (%s)\n", b200spmv_last_error());
#endif
#ifdef B200_DEVICE_RESIDENT
    b200_cuda(cudaMalloc((void **)&A_opt.x_dev, sizeof(double) * (A.nCol > 0 ? A.nCol : 1)), "cudaMalloc x");
    b200_cuda(cudaMalloc((void **)&A_opt.y_dev, sizeof(double) * (A.nRow > 0 ? A.nRow : 1)), "cudaMalloc y");
    B200UploadVector(A_opt, x_opt);
#endif
}

extern "C" {
void SpMV (const SpMatOpt &A, const VecOpt &x, Vec &y) {
    if (A.mg) {
#ifdef B200_DEVICE_RESIDENT
        (void)x; (void)y;
        b200_check(b200spmv_mg_multiply(A.mg), "mg_multiply");
#else
        static const double *registered_y_mg = NULL;          // page-lock the driver's y once: asynchronous D2H from every GPU
        if (registered_y_mg != y.val && y.size > 0 && !getenv("B200_NO_HOST_REGISTER")) {
            b200spmv_host_register(y.val, sizeof(double) * (unsigned long long)y.size);
            registered_y_mg = y.val;
        }
        b200_check(b200spmv_mg_multiply_host(A.mg, x.val, y.val), "mg_multiply_host");
#endif
        return;
    }
#ifdef B200_DEVICE_RESIDENT
    (void)x; (void)y;
    b200_check(b200spmv_multiply(A.handle, A.x_dev, A.y_dev, A.stream), "multiply");
#else
    static const double *registered_y = NULL;
    if (registered_y != y.val && y.size > 0 && !getenv("B200_NO_HOST_REGISTER")) {
        if (b200spmv_host_register(y.val, sizeof(double) * (unsigned long long)y.size) != B200SPMV_OK)
            fprintf(stderr, "b200spmv: y stays pageable (%s)\n", b200spmv_last_error());
        registered_y = y.val;
    }
    b200_check(b200spmv_multiply_host(A.handle, x.val, y.val), "multiply");
#endif
#if defined(PROFILING) && defined(B200_SS_FAITHFUL)
    if (g_profile.size() >= 2) {                      // PROF_BEGIN / PROF_END accumulate seconds (src/util.h:59-65)
        long long mul = 0, sum = 0;
        b200spmv_get_scalar(A.handle, "MulTime_ns", &mul);
        b200spmv_get_scalar(A.handle, "SumTime_ns", &sum);
        g_profile[0] += mul * 1e-9;
        g_profile[1] += sum * 1e-9;
    }
#endif
}
}

void B200UploadVector (const SpMatOpt &A, const VecOpt &x) {
    if (A.mg) { b200_check(b200spmv_mg_upload_x(A.mg, x.val), "mg_upload_x"); return; }
    if (!A.x_dev) return;
    b200_cuda(cudaMemcpy(A.x_dev, x.val, sizeof(double) * A.nCol, cudaMemcpyHostToDevice), "H2D x");
}
void B200FetchResult (const SpMatOpt &A, Vec &y) {
    if (A.mg) {
#ifdef B200_DEVICE_RESIDENT
        b200_check(b200spmv_mg_download_y(A.mg, y.val), "mg_download_y");
#endif
        return;
    }
    if (!A.y_dev) return;
    b200_cuda(cudaStreamSynchronize((cudaStream_t)A.stream), "sync");
    b200_cuda(cudaMemcpy(y.val, A.y_dev, sizeof(double) * A.nRow, cudaMemcpyDeviceToHost), "D2H y");
}
void B200Synchronize (const SpMatOpt &A) {
    if (A.mg) { b200_check(b200spmv_mg_synchronize(A.mg), "mg_synchronize"); return; }
    b200_cuda(cudaStreamSynchronize((cudaStream_t)A.stream), "sync");
}
long long B200Scalar (const SpMatOpt &A, const char *name) {
    long long v = 0;
    if (A.mg) b200_check(b200spmv_mg_get_scalar(A.mg, name, &v), name);
    else b200_check(b200spmv_get_scalar(A.handle, name, &v), name);
    return v;
}
