// opt_b200.h -- drop-in plugin header for hir0shim/singleSpMV's format registry.
//
// The reference selects ONE format per binary: src/opt.h:1-28 includes opt_<fmt>.h (types) and
// src/opt.cpp:5-33 includes opt_<fmt>.cpp (code); every plugin defines the same names SpMatOpt, VecOpt,
// OptimizeProblem, SpMV (e.g. src/opt_crs.h:3-18).  This header + opt_b200.cpp are one more such pair,
// backed by libb200spmv.so (include/b200spmv.h).  Compile-time selection, like the reference:
//
//     -DOPT_B200 -DB200_FORMAT=CRS|COO|ELL|JDS|DIA|SS|CSS|CSR5
//     [-DSEGMENT_WIDTH=W] [-DN_BLOCK=N]   (same macros as src/param.h:9-20, used by SS / CSS)
//     [-DB200_SIGMA=s]                    (CSR5 sigma, 0 = auto)
//     [-DB200_VALUE_F32]                  (CRS: matrix values stored as fp32, arithmetic in fp64)
//     [-DB200_SS_FAITHFUL [-DPROFILING]]  (SS / CSS: the reference's three-phase schedule; with PROFILING the Mul and
//                                          Sum phase times go to g_profile[0] / [1] like src/opt_ss.cpp:225-304)
//     [-DB200_NGPU=N]                     (row-partition the matrix over N GPUs of the box, one process: b200spmv_mg_*;
//                                          N = 0: all visible GPUs; the environment variable B200_NGPU overrides)
//     [-DB200_DEVICE_RESIDENT]            (x stays in HBM, y is copied back only by B200FetchResult();
//                                          default = host semantics like src/opt_cusparse.cpp:72-82)
//
// Inside the reference tree: add `#elif defined(OPT_B200) / #include "opt_b200.h"` to src/opt.h and the
// matching `#include "opt_b200.cpp"` to src/opt.cpp (README.md:5-8 procedure); see INTEGRATION.md.
#pragma once
#include "util.h"            // SpMat, Vec: the host project's util.h -- the reference's src/util.h:7-28 inside its tree,
                             // plugin/standalone/util.h (same names and layout) for the driver built here
#include "b200spmv.h"

struct SpMatOpt {
    int nRow;
    int nCol;
    int nNnz;
    b200spmv_matrix *handle;    // the converted matrix lives in HBM behind the C-ABI
    b200spmv_mg *mg;            // -DB200_NGPU: the same matrix row-partitioned over several GPUs (handle is NULL then)
    int nGPU;
    double *x_dev;              // device-resident mode only
    double *y_dev;
    void *stream;               // cudaStream_t used by the multiply (NULL = default stream)
};
struct VecOpt {
    int size;
    double *val;                // aliases the caller's x, like every reference plugin (src/opt_crs.cpp:11-12)
};
void OptimizeProblem (const SpMat &A, const Vec &x, SpMatOpt &A_opt, VecOpt &x_opt);
extern "C" {
void SpMV (const SpMatOpt &A, const VecOpt &x, Vec &y);
}
// device-resident mode helpers (no counterpart in the reference: its vectors never leave the host)
void B200UploadVector (const SpMatOpt &A, const VecOpt &x);     // call again whenever x changes
void B200FetchResult (const SpMatOpt &A, Vec &y);               // D2H of y + synchronisation
void B200Synchronize (const SpMatOpt &A);
const char *B200FormatName ();
long long B200Scalar (const SpMatOpt &A, const char *name);     // alg_bytes, launches, K, nDiag, ...
