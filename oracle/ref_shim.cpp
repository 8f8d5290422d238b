// ref_shim.cpp -- flat C entry points around the UNMODIFIED reference plugins.
//
// TEST INFRASTRUCTURE ONLY (see oracle/spmv_oracle.c header).  This file is ours; the
// reference sources are compiled where they lie (-I/root/reference/src plus
// /root/reference/src/{util,opt}.cpp on the command line, see oracle/Makefile) and the
// result goes to oracle/_ref/libref_<variant>.so, which is git-ignored.
//
// One library per compile-time variant, exactly like the reference builds one binary per
// (-DOPT_<FMT>, SEGMENT_WIDTH, N_BLOCK) choice (reference Makefile:10-21, src/opt.h:1-28).
//
// Exposed: ref_convert() = OptimizeProblem (e.g. src/opt_crs.cpp:10), ref_spmv() = SpMV
// (e.g. src/opt_crs.cpp:45), ref_scalar()/ref_array() = read-back of the SpMatOpt fields.
#include <cstring>
#include <cstdlib>
#include <string>
#include <vector>
#include "opt.h"
#include "util.h"
#include "param.h"

static SpMat g_A;
static Vec g_x;
static SpMatOpt g_opt;
static VecOpt g_xopt;

template <typename T> static T *dup(const T *src, size_t n)
{
    T *p = (T *)malloc(sizeof(T) * (n ? n : 1));
    if (n) memcpy(p, src, sizeof(T) * n);
    return p;
}

static long put(void *dst, const void *src, size_t bytes)
{
    if (dst && bytes) memcpy(dst, src, bytes);
    return (long)bytes;
}

extern "C" {

__attribute__((visibility("default")))
int ref_convert(int nRow, int nCol, int nnz, const int *row, const int *col, const double *val,
                const double *x)
{
    g_A.nRow = nRow; g_A.nCol = nCol; g_A.nNnz = nnz;
    g_A.row_idx = dup(row, (size_t)nnz);
    g_A.col_idx = dup(col, (size_t)nnz);
    g_A.val = dup(val, (size_t)nnz);
    g_x.size = nCol;
    g_x.val = dup(x, (size_t)nCol);
    OptimizeProblem(g_A, g_x, g_opt, g_xopt);
    return 0;
}

__attribute__((visibility("default")))
void ref_set_x(const double *x) { memcpy(g_x.val, x, sizeof(double) * (size_t)g_x.size); }

__attribute__((visibility("default")))
void ref_spmv(double *y)
{
    Vec out; out.size = g_A.nRow; out.val = y;
    SpMV(g_opt, g_xopt, out);
}

// Returns 1 and stores the value if `name` is a scalar of this variant's SpMatOpt.
__attribute__((visibility("default")))
int ref_scalar(const char *name, long *out)
{
    std::string n(name);
    if (n == "nRow") { *out = g_opt.nRow; return 1; }
    if (n == "nCol") { *out = g_opt.nCol; return 1; }
    if (n == "nNnz") { *out = g_opt.nNnz; return 1; }
#if defined(OPT_ELL)
    if (n == "K") { *out = g_opt.K; return 1; }
#elif defined(OPT_JDS)
    if (n == "maxLength") { *out = g_opt.maxLength; return 1; }
#elif defined(OPT_DIA)
    if (n == "nDiag") { *out = g_opt.nDiag; return 1; }
#elif defined(OPT_SS)
    if (n == "H") { *out = g_opt.H; return 1; }
    if (n == "nStep") { *out = g_opt.nStep; return 1; }
    if (n == "W") { *out = (long)(SEGMENT_WIDTH); return 1; }
#elif defined(OPT_CSS)
    if (n == "B") { *out = g_opt.B; return 1; }
    if (n == "nBlock") { *out = g_opt.nBlock; return 1; }
    if (n == "totalH") { *out = g_opt.totalH; return 1; }
    if (n == "W") { *out = (long)(SEGMENT_WIDTH); return 1; }
#endif
    return 0;
}

// Copies the named array into dst (dst == NULL: size query).  Returns bytes, or -1 if unknown.
__attribute__((visibility("default")))
long ref_array(const char *name, void *dst)
{
    std::string n(name);
    const size_t nRow = (size_t)g_opt.nRow, nnz = (size_t)g_opt.nNnz;
    (void)nRow; (void)nnz;
#if defined(OPT_CRS)
    if (n == "ptr") return put(dst, g_opt.ptr, sizeof(int) * (nRow + 1));
    if (n == "idx") return put(dst, g_opt.idx, sizeof(int) * nnz);
    if (n == "val") return put(dst, g_opt.val, sizeof(double) * nnz);
#elif defined(OPT_COO)
    if (n == "row_idx") return put(dst, g_opt.row_idx, sizeof(int) * nnz);
    if (n == "col_idx") return put(dst, g_opt.col_idx, sizeof(int) * nnz);
    if (n == "val") return put(dst, g_opt.val, sizeof(double) * nnz);
#elif defined(OPT_ELL)
    const size_t K = (size_t)g_opt.K;
    if (n == "col_idx" || n == "val") {       // logical [nRow][K] from the array of row pointers
        const bool isCol = (n == "col_idx");
        const size_t el = isCol ? sizeof(int) : sizeof(double);
        if (dst)
            for (size_t r = 0; r < nRow; r++)
                memcpy((char *)dst + r * K * el,
                       isCol ? (const void *)g_opt.col_idx[r] : (const void *)g_opt.val[r], K * el);
        return (long)(nRow * K * el);
    }
#elif defined(OPT_JDS)
    const size_t L = (size_t)g_opt.maxLength;
    if (n == "perm") return put(dst, g_opt.perm, sizeof(int) * nRow);
    if (n == "length") return put(dst, g_opt.length, sizeof(int) * nRow);
    if (n == "ptr") return put(dst, g_opt.ptr, sizeof(int) * (L + 1));
    if (n == "col_idx") return put(dst, g_opt.col_idx, sizeof(int) * nnz);
    if (n == "val") return put(dst, g_opt.val, sizeof(double) * nnz);
#elif defined(OPT_DIA)
    const size_t D = (size_t)g_opt.nDiag, nCol = (size_t)g_opt.nCol;
    if (n == "ioff") return put(dst, g_opt.ioff, sizeof(int) * D);
    if (n == "diag") {                          // logical [nDiag][nCol]
        if (dst)
            for (size_t d = 0; d < D; d++)
                memcpy((double *)dst + d * nCol, g_opt.diag[d], sizeof(double) * nCol);
        return (long)(D * nCol * sizeof(double));
    }
#elif defined(OPT_SS)
    const size_t H = (size_t)g_opt.H, W = (size_t)(SEGMENT_WIDTH);
    if (n == "row_ptr") return put(dst, g_opt.row_ptr, sizeof(int) * (nRow + 1));
    if (n == "row_idx") return put(dst, H ? g_opt.row_idx[0] : NULL, sizeof(int) * H * W);
    if (n == "col_idx") return put(dst, H ? g_opt.col_idx[0] : NULL, sizeof(idx_t) * H * W);
    if (n == "val") return put(dst, H ? g_opt.val[0] : NULL, sizeof(double) * H * W);
    if (n == "segment_index") return put(dst, g_opt.segment_index, sizeof(int) * H);
    if (n == "sum_segs_count") return put(dst, g_opt.sum_segs_count, sizeof(int) * (size_t)g_opt.nStep);
    if (n == "sum_segs") {                      // levels concatenated
        size_t total = 0;
        for (int s = 0; s < g_opt.nStep; s++) {
            size_t c = (size_t)g_opt.sum_segs_count[s];
            if (dst) memcpy((int *)dst + total, g_opt.sum_segs[s], sizeof(int) * c);
            total += c;
        }
        return (long)(total * sizeof(int));
    }
#elif defined(OPT_CSS)
    const size_t W = (size_t)(SEGMENT_WIDTH), nB = (size_t)g_opt.nBlock, tH = (size_t)g_opt.totalH;
    if (n == "H") return put(dst, g_opt.H, sizeof(int) * nB);
    if (n == "nStep") return put(dst, g_opt.nStep, sizeof(int) * nB);
    if (n == "col_idx") return put(dst, tH ? g_opt.col_idx[0][0] : NULL, sizeof(idx_t) * tH * W);
    if (n == "val") return put(dst, tH ? g_opt.val[0][0] : NULL, sizeof(double) * tH * W);
    if (n == "row_ptr") {                       // [nBlock][nRow+1]
        if (dst)
            for (size_t b = 0; b < nB; b++)
                memcpy((int *)dst + b * (nRow + 1), g_opt.row_ptr[b], sizeof(int) * (nRow + 1));
        return (long)(nB * (nRow + 1) * sizeof(int));
    }
    if (n == "sum_segs_count") {                // blocks concatenated, nStep[b] entries each
        size_t total = 0;
        for (size_t b = 0; b < nB; b++) {
            size_t c = (size_t)g_opt.nStep[b];
            if (dst && c) memcpy((int *)dst + total, g_opt.sum_segs_count[b], sizeof(int) * c);
            total += c;
        }
        return (long)(total * sizeof(int));
    }
    if (n == "sum_segs") {                      // blocks, then levels, concatenated
        size_t total = 0;
        for (size_t b = 0; b < nB; b++)
            for (int s = 0; s < g_opt.nStep[b]; s++) {
                size_t c = (size_t)g_opt.sum_segs_count[b][s];
                if (dst && c) memcpy((int *)dst + total, g_opt.sum_segs[b][s], sizeof(int) * c);
                total += c;
            }
        return (long)(total * sizeof(int));
    }
#endif
    return -1;
}

}  // extern "C"
