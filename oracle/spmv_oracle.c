/*
 * spmv_oracle.c -- CPU restatement of hir0shim/singleSpMV's convert + multiply path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product (singlespmv_b200/, include/) may
 * link, import or execute this file.  It is used by tests/, __graft_entry__.smoke()
 * and the cpu_baseline / --impl reference legs of bench.py, and only as the checker.
 *
 * Parity status: PINNED.  Every function below is validated bit-for-bit against the
 * reference's own plugins compiled unmodified from /root/reference/src (oracle/_ref,
 * built by oracle/Makefile) in tests/test_oracle_vs_ref.py, and against the
 * known-answer vectors of SURVEY.md Appendix B + tests/golden/ (generated from the
 * reference by tests/golden/make_golden.py).
 *
 * Everything is plain C over flat arrays: fp64 values, int32 indices, input = COO sorted
 * by (row, col) without duplicate coordinates (reference src/util.cpp:51).
 * Build:  gcc -O2 -ffp-contract=off -fopenmp -shared -fPIC  (no -march: keeps mul and
 * add unfused, the summation order of reference src/opt_crs.cpp:61-67).
 *
 * Citations are file:line under /root/reference/.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>

#define ORC_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------ CRS ---- */

/* src/opt_crs.cpp:27-33 -- ptr[r] = index of first COO entry with row >= r. */
ORC_API void orc_crs_row_ptr(int nRow, int nnz, const int *row, int *ptr)
{
    int next = 0;
    for (int i = 0; i < nnz; i++)
        for (; next <= row[i]; next++) ptr[next] = i;
    for (; next <= nRow; next++) ptr[next] = nnz;
}

/* src/opt_crs.cpp:57-69 -- row-parallel dot products, ascending column order, unfused. */
ORC_API void orc_crs_spmv(int nRow, const int *ptr, const int *idx, const double *val,
                          const double *x, double *y)
{
#pragma omp parallel for
    for (int r = 0; r < nRow; r++) {
        double acc = 0;
        for (int j = ptr[r]; j < ptr[r + 1]; j++) {
            double prod = val[j] * x[idx[j]];
            acc += prod;
        }
        y[r] = acc;
    }
}

/* ------------------------------------------------------------------ COO ---- */

/* src/opt_coo.cpp:34-46 (and the serial verifier src/util.cpp:67-72): y=0 then scatter-add.
 * Serial here, so the per-row order is the ascending-column order. */
ORC_API void orc_coo_spmv(int nRow, int nnz, const int *row, const int *col,
                          const double *val, const double *x, double *y)
{
    for (int r = 0; r < nRow; r++) y[r] = 0;
    for (int i = 0; i < nnz; i++) {
        double prod = val[i] * x[col[i]];
        y[row[i]] += prod;
    }
}

/* ------------------------------------------------------------------ ELL ---- */

/* src/opt_ell.cpp:28-31 -- K = longest row. */
ORC_API int orc_ell_width(int nRow, int nnz, const int *row)
{
    int *len = (int *)calloc((size_t)(nRow > 0 ? nRow : 1), sizeof(int));
    int K = 0;
    for (int i = 0; i < nnz; i++) len[row[i]]++;
    for (int r = 0; r < nRow; r++) if (len[r] > K) K = len[r];
    free(len);
    return K;
}

/* src/opt_ell.cpp:38-52 -- logical [nRow][K]; slot k<len: COO order; k>=len: col=k, val=0. */
ORC_API void orc_ell_convert(int nRow, int nnz, int K, const int *row, const int *col,
                             const double *val, int *ecol, double *eval)
{
    int *fill = (int *)calloc((size_t)(nRow > 0 ? nRow : 1), sizeof(int));
    for (int i = 0; i < nnz; i++) {
        int r = row[i];
        size_t at = (size_t)r * K + fill[r]++;
        ecol[at] = col[i];
        eval[at] = val[i];
    }
    for (int r = 0; r < nRow; r++)
        for (int k = fill[r]; k < K; k++) {
            ecol[(size_t)r * K + k] = k;
            eval[(size_t)r * K + k] = 0;
        }
    free(fill);
}

/* src/opt_ell.cpp:75-89 -- y[r] += x[col]*val over all K slots, padding included. */
ORC_API void orc_ell_spmv(int nRow, int K, const int *ecol, const double *eval,
                          const double *x, double *y)
{
    for (int r = 0; r < nRow; r++) {
        y[r] = 0;
        for (int k = 0; k < K; k++) {
            double prod = x[ecol[(size_t)r * K + k]] * eval[(size_t)r * K + k];
            y[r] += prod;
        }
    }
}

/* ------------------------------------------------------------------ JDS ---- */

/* src/opt_jds.cpp:35-60.  perm_in == NULL: rows ordered by length descending, ties by
 * ascending row (the documented stable convention; the reference's unstable std::sort
 * leaves tie order to libstdc++).  perm_in != NULL: take that
 * permutation (e.g. the one the reference produced) and build everything else from it.
 * jptr has maxLength+1 entries (caller sizes it with orc_ell_width()+1). */
ORC_API int orc_jds_convert(int nRow, int nnz, const int *row, const int *col, const double *val,
                            const int *perm_in, int *perm, int *length, int *jptr,
                            int *jcol, double *jval)
{
    int *start = (int *)malloc(sizeof(int) * ((size_t)nRow + 1));
    orc_crs_row_ptr(nRow, nnz, row, start);
    int maxLength = 0;
    for (int r = 0; r < nRow; r++) {
        length[r] = start[r + 1] - start[r];
        if (length[r] > maxLength) maxLength = length[r];
    }
    if (perm_in) {
        memcpy(perm, perm_in, sizeof(int) * (size_t)nRow);
    } else {
        /* counting sort by length, descending, stable in row id */
        int *bucket = (int *)calloc((size_t)maxLength + 2, sizeof(int));
        for (int r = 0; r < nRow; r++) bucket[maxLength - length[r] + 1]++;
        for (int l = 0; l <= maxLength; l++) bucket[l + 1] += bucket[l];
        for (int r = 0; r < nRow; r++) perm[bucket[maxLength - length[r]]++] = r;
        free(bucket);
    }
    int out = 0;
    for (int c = 0; c < maxLength; c++) {
        jptr[c] = out;
        for (int i = 0; i < nRow; i++) {
            int r = perm[i];
            if (c >= length[r]) break;
            jcol[out] = col[start[r] + c];
            jval[out] = val[start[r] + c];
            out++;
        }
    }
    jptr[maxLength] = out;
    free(start);
    return maxLength;
}

/* src/opt_jds.cpp:91-103 -- position r handles row perm[r], walking down the diagonals. */
ORC_API void orc_jds_spmv(int nRow, const int *perm, const int *length, const int *jptr,
                          const int *jcol, const double *jval, const double *x, double *y)
{
    for (int r = 0; r < nRow; r++) y[r] = 0;
    for (int pos = 0; pos < nRow; pos++) {
        int r = perm[pos];
        for (int c = 0; c < length[r]; c++) {
            double prod = jval[jptr[c] + pos] * x[jcol[jptr[c] + pos]];
            y[r] += prod;
        }
    }
}

/* ------------------------------------------------------------------ DIA ---- */

/* src/opt_dia.cpp:21-45 -- diagonal id = col - row + (nRow-1); ioff = ascending distinct ids.
 * Returns nDiag; ioff may be NULL for a counting call. */
ORC_API int orc_dia_offsets(int nRow, int nCol, int nnz, const int *row, const int *col, int *ioff)
{
    size_t N = (size_t)nRow + (size_t)nCol - 1;
    unsigned char *seen = (unsigned char *)calloc(N ? N : 1, 1);
    for (int i = 0; i < nnz; i++) seen[col[i] - row[i] + (nRow - 1)] = 1;
    int nDiag = 0;
    for (size_t d = 0; d < N; d++)
        if (seen[d]) { if (ioff) ioff[nDiag] = (int)d; nDiag++; }
    free(seen);
    return nDiag;
}

/* src/opt_dia.cpp:47-56 -- diag[p][col] = val (last write wins), zero elsewhere; flat [nDiag][nCol]. */
ORC_API void orc_dia_convert(int nRow, int nCol, int nnz, const int *row, const int *col,
                             const double *val, int nDiag, const int *ioff, double *diag)
{
    size_t N = (size_t)nRow + (size_t)nCol - 1;
    int *slot = (int *)malloc(sizeof(int) * (N ? N : 1));
    for (size_t d = 0; d < N; d++) slot[d] = -1;
    for (int p = 0; p < nDiag; p++) slot[ioff[p]] = p;
    memset(diag, 0, sizeof(double) * (size_t)nDiag * (size_t)nCol);
    for (int i = 0; i < nnz; i++)
        diag[(size_t)slot[col[i] - row[i] + (nRow - 1)] * nCol + col[i]] = val[i];
    free(slot);
}

/* src/opt_dia.cpp:83-92 -- serial, diagonal-major accumulation. */
ORC_API void orc_dia_spmv(int nRow, int nCol, int nDiag, const int *ioff, const double *diag,
                          const double *x, double *y)
{
    for (int r = 0; r < nRow; r++) y[r] = 0;
    for (int p = 0; p < nDiag; p++)
        for (int c = 0; c < nCol; c++) {
            int r = c + (nRow - 1) - ioff[p];
            if (r < 0 || r >= nRow) continue;
            double prod = diag[(size_t)p * nCol + c] * x[c];
            y[r] += prod;
        }
}

/* ------------------------------------------------------------------- SS ---- */

/* Shared by SS and CSS: chain position of each W-wide segment (src/opt_ss.cpp:91-107,
 * src/opt_css.cpp:118-134).  A segment continues a chain when its first row id equals the
 * previous segment's first row id and all its W row ids are equal. */
static void segment_chain(int H, int W, const int *row2d, int *segment_index)
{
    if (H > 0) segment_index[0] = 0;
    for (int s = 1; s < H; s++) {
        const int *cur = row2d + (size_t)s * W;
        int chained = (row2d[(size_t)(s - 1) * W] == cur[0]);
        for (int j = 1; j < W && chained; j++)
            if (cur[j - 1] != cur[j]) chained = 0;
        segment_index[s] = chained ? segment_index[s - 1] + 1 : 0;
    }
}

static int chain_steps(int H, const int *segment_index)
{
    int deepest = 0;
    for (int s = 0; s < H; s++) if (segment_index[s] > deepest) deepest = segment_index[s];
    return (int)ceil(log2((double)deepest + 1.0));   /* src/opt_ss.cpp:121 */
}

/* src/opt_ss.cpp:122-142 -- level s lists, ascending, the segments whose chain position lies
 * in [2^(nStep-1-s), 2^(nStep-s)).  Writes counts[nStep]; if segs != NULL also the
 * concatenated lists.  Returns the total number of listed segments. */
static int chain_schedule(int H, int nStep, const int *segment_index, int *counts, int *segs)
{
    int total = 0;
    for (int s = 0; s < nStep; s++) {
        int lo = 1 << (nStep - 1 - s), n = 0;
        for (int h = 0; h < H; h++)
            if (lo <= segment_index[h] && segment_index[h] < 2 * lo) {
                if (segs) segs[total + n] = h;
                n++;
            }
        counts[s] = n;
        total += n;
    }
    return total;
}

ORC_API int orc_ss_height(int nnz, int W) { return nnz / W + (nnz % W != 0); }  /* opt_ss.cpp:33 */

/* src/opt_ss.cpp:64-107,121 -- slabs [H][W] with padding (row=nRow, col=0, val=0), CRS-style
 * row_ptr, segment_index.  Returns nStep. */
ORC_API int orc_ss_convert(int nRow, int nnz, int W, const int *row, const int *col,
                           const double *val, int *row_ptr, int *row2d, int *col2d,
                           double *val2d, int *segment_index)
{
    int H = orc_ss_height(nnz, W);
    size_t slots = (size_t)H * W;
    for (size_t p = 0; p < slots; p++) {
        if (p < (size_t)nnz) { row2d[p] = row[p]; col2d[p] = col[p]; val2d[p] = val[p]; }
        else                 { row2d[p] = nRow;   col2d[p] = 0;      val2d[p] = 0; }
    }
    orc_crs_row_ptr(nRow, nnz, row, row_ptr);
    segment_chain(H, W, row2d, segment_index);
    return chain_steps(H, segment_index);
}

ORC_API int orc_ss_schedule(int H, int nStep, const int *segment_index, int *counts, int *segs)
{
    return chain_schedule(H, nStep, segment_index, counts, segs);
}

/* src/opt_ss.cpp:188-221 (SIMPLE): products into val_buf, then per-row sequential gather. */
ORC_API void orc_ss_spmv_simple(int nRow, int H, int W, const int *row_ptr, const int *col2d,
                                const double *val2d, double *val_buf, const double *x, double *y)
{
    size_t slots = (size_t)H * W;
    for (size_t p = 0; p < slots; p++) val_buf[p] = val2d[p] * x[col2d[p]];
    for (int r = 0; r < nRow; r++) {
        y[r] = 0;
        for (int j = row_ptr[r]; j < row_ptr[r + 1]; j++) y[r] += val_buf[j];
    }
}

/* Chain tree-add (src/opt_ss.cpp:242-260): level s folds segment h onto h - 2^(nStep-1-s). */
static void chain_fold(int W, int nStep, const int *counts, const int *segs, double *val_buf)
{
    int base = 0;
    for (int s = 0; s < nStep; s++) {
        int dist = 1 << (nStep - 1 - s);
        for (int i = 0; i < counts[s]; i++) {
            int h = segs[base + i];
            for (int j = 0; j < W; j++)
                val_buf[(size_t)(h - dist) * W + j] += val_buf[(size_t)h * W + j];
        }
        base += counts[s];
    }
}

/* Row gather after the fold (src/opt_ss.cpp:264-303): head partial ascending, tail partial
 * descending, then the W column sums of the chain head.  W must be a power of two. */
static double folded_row_sum(int W, int begin, int end, const double *val_buf)
{
    double acc = 0;
    if (begin / W == end / W) {
        for (int j = begin; j < end; j++) acc += val_buf[j];
        return acc;
    }
    if (begin & (W - 1)) {
        int stop = (begin & ~(W - 1)) + W;
        for (int j = begin; j < stop; j++) acc += val_buf[j];
        begin = stop;
    }
    if (end & (W - 1)) {
        int stop = end & ~(W - 1);
        for (int j = end; j > stop; j--) acc += val_buf[j - 1];
        end = stop;
    }
    if (begin != end)
        for (int j = 0; j < W; j++) acc += val_buf[begin + j];
    return acc;
}

/* src/opt_ss.cpp:222-303 (OPTIMIZED, no PADDING). */
ORC_API void orc_ss_spmv_optimized(int nRow, int H, int W, const int *row_ptr, const int *col2d,
                                   const double *val2d, int nStep, const int *counts,
                                   const int *segs, double *val_buf, const double *x, double *y)
{
    size_t slots = (size_t)H * W;
    for (size_t p = 0; p < slots; p++) val_buf[p] = val2d[p] * x[col2d[p]];
    chain_fold(W, nStep, counts, segs, val_buf);
    for (int r = 0; r < nRow; r++) y[r] = folded_row_sum(W, row_ptr[r], row_ptr[r + 1], val_buf);
}

/* ------------------------------------------------------------------ CSS ---- */

/* src/opt_css.cpp:33-36 -- column block width B and block count. */
ORC_API int orc_css_block_width(int nCol, int nBlockWanted)
{
    return (int)ceil((double)nCol / nBlockWanted);
}
ORC_API int orc_css_num_blocks(int nCol, int B) { return nCol / B + (nCol % B ? 1 : 0); }

/* src/opt_css.cpp:37-57 -- per-block nnz counts -> H[b]; returns totalH. */
ORC_API int orc_css_heights(int nnz, int W, int B, int nBlock, const int *col, int *Hb, int *blockNnz)
{
    for (int b = 0; b < nBlock; b++) blockNnz[b] = 0;
    for (int i = 0; i < nnz; i++) blockNnz[col[i] / B]++;
    int totalH = 0;
    for (int b = 0; b < nBlock; b++) {
        Hb[b] = blockNnz[b] / W + (blockNnz[b] % W ? 1 : 0);
        totalH += Hb[b];
    }
    return totalH;
}

/* src/opt_css.cpp:60-111,113-137 -- one slab of totalH*W for all blocks (block-major, COO order
 * kept inside a block), padding row=0/col=0/val=0, per-block row_ptr[nBlock][nRow+1],
 * per-block segment chains and nStep[b]. */
ORC_API void orc_css_convert(int nRow, int nnz, int W, int B, int nBlock, const int *row,
                             const int *col, const double *val, const int *Hb,
                             int *row_ptr, int *row2d, int *col2d, double *val2d,
                             int *segment_index, int *nStep)
{
    size_t *cursor = (size_t *)malloc(sizeof(size_t) * (size_t)nBlock);
    size_t *base = (size_t *)malloc(sizeof(size_t) * (size_t)nBlock);
    size_t at = 0;
    for (int b = 0; b < nBlock; b++) { base[b] = cursor[b] = at; at += (size_t)Hb[b] * W; }
    for (size_t p = 0; p < at; p++) { row2d[p] = 0; col2d[p] = 0; val2d[p] = 0; }
    int *next = (int *)calloc((size_t)nBlock, sizeof(int));
    for (int i = 0; i < nnz; i++) {
        int b = col[i] / B;
        size_t p = cursor[b]++;
        row2d[p] = row[i]; col2d[p] = col[i]; val2d[p] = val[i];
        int local = (int)(p - base[b]);
        int *rp = row_ptr + (size_t)b * (nRow + 1);
        for (; next[b] <= row[i]; next[b]++) rp[next[b]] = local;
    }
    size_t seg0 = 0;
    for (int b = 0; b < nBlock; b++) {
        int *rp = row_ptr + (size_t)b * (nRow + 1);
        int cnt = (int)(cursor[b] - base[b]);
        for (; next[b] <= nRow; next[b]++) rp[next[b]] = cnt;
        segment_chain(Hb[b], W, row2d + base[b], segment_index + seg0);
        nStep[b] = chain_steps(Hb[b], segment_index + seg0);
        seg0 += (size_t)Hb[b];
    }
    free(next); free(cursor); free(base);
}

/* src/opt_css.cpp:138-156 for one block: counts[nStep[b]] and the concatenated lists. */
ORC_API int orc_css_schedule(int Hblock, int nStepBlock, const int *segment_index_block,
                             int *counts, int *segs)
{
    return chain_schedule(Hblock, nStepBlock, segment_index_block, counts, segs);
}

/* src/opt_css.cpp:226-303 (OPTIMIZED): products for the whole slab, per-block fold, then
 * y[r] = sum over blocks (block order) of the block's folded row sum.
 * counts/segs are the per-block schedules concatenated in block order. */
ORC_API void orc_css_spmv_optimized(int nRow, int W, int nBlock, const int *Hb, const int *row_ptr,
                                    const int *col2d, const double *val2d, const int *nStep,
                                    const int *counts, const int *segs, double *val_buf,
                                    const double *x, double *y)
{
    size_t slots = 0;
    for (int b = 0; b < nBlock; b++) slots += (size_t)Hb[b] * W;
    for (size_t p = 0; p < slots; p++) val_buf[p] = val2d[p] * x[col2d[p]];
    size_t base = 0; int cbase = 0, sbase = 0;
    for (int r = 0; r < nRow; r++) y[r] = 0;
    for (int b = 0; b < nBlock; b++) {
        double *buf = val_buf + base;
        chain_fold(W, nStep[b], counts + cbase, segs + sbase, buf);
        const int *rp = row_ptr + (size_t)b * (nRow + 1);
        for (int r = 0; r < nRow; r++) y[r] += folded_row_sum(W, rp[r], rp[r + 1], buf);
        for (int s = 0; s < nStep[b]; s++) sbase += counts[cbase + s];
        cbase += nStep[b];
        base += (size_t)Hb[b] * W;
    }
}

/* ------------------------------------------------------------- verifier ---- */

/* src/util.cpp:67-83 -- the reference's in-binary check: fail iff abs>1e-6 AND rel>1e-6. */
ORC_API int orc_verify(int nRow, int nnz, const int *row, const int *col, const double *val,
                       const double *x, const double *y)
{
    double *res = (double *)calloc((size_t)(nRow > 0 ? nRow : 1), sizeof(double));
    for (int i = 0; i < nnz; i++) res[row[i]] += val[i] * x[col[i]];
    int ok = 1;
    for (int r = 0; r < nRow && ok; r++) {
        double abs_err = fabs(res[r] - y[r]);
        double rel_err = fabs(abs_err / res[r]);
        if (abs_err > 1e-6 && rel_err > 1e-6) ok = 0;
    }
    free(res);
    return ok;
}

/* src/util.cpp:92-102 + src/main.cpp:18,31-32 -- glibc rand() stream: x first, then y. */
ORC_API void orc_reference_vectors(unsigned seed, int nCol, int nRow, double *x, double *y)
{
    srand(seed);
    for (int i = 0; i < nCol; i++) x[i] = (double)rand() / RAND_MAX;
    for (int i = 0; i < nRow; i++) { double v = (double)rand() / RAND_MAX; if (y) y[i] = v; }
}

/* ------------------------------------------------------------ statistics ---- */

/* matrix/script/counter.cpp:19-42 -- non-zeros per row and per column: max/min of each, and the variance of
 * the row counts, accumulated term by term like the reference does.  out5 = rowMax,rowMin,colMax,colMin,nDiag'
 * where nDiag' (not in counter.cpp) is the number of distinct col-row values (src/opt_dia.cpp:29-34). */
ORC_API double orc_counter(int nRow, int nCol, int nnz, const int *row, const int *col, long long *out5)
{
    int *cr = (int *)calloc((size_t)(nRow > 0 ? nRow : 1), sizeof(int));
    int *cc = (int *)calloc((size_t)(nCol > 0 ? nCol : 1), sizeof(int));
    size_t N = (size_t)nRow + (size_t)nCol;
    unsigned char *seen = (unsigned char *)calloc(N ? N : 1, 1);
    for (int i = 0; i < nnz; i++) { cr[row[i]]++; cc[col[i]]++; seen[col[i] - row[i] + (nRow - 1)] = 1; }
    int rmax = 0, rmin = nRow ? cr[0] : 0, cmax = 0, cmin = nCol ? cc[0] : 0;
    double ave = nRow ? (double)nnz / nRow : 0, var = 0;
    for (int r = 0; r < nRow; r++) {
        if (cr[r] > rmax) rmax = cr[r];
        if (cr[r] < rmin) rmin = cr[r];
        var += (cr[r] - ave) * (cr[r] - ave) / nRow;
    }
    for (int c = 0; c < nCol; c++) {
        if (cc[c] > cmax) cmax = cc[c];
        if (cc[c] < cmin) cmin = cc[c];
    }
    long long nd = 0;
    for (size_t d = 0; d < N; d++) nd += seen[d];
    out5[0] = rmax; out5[1] = rmin; out5[2] = cmax; out5[3] = cmin; out5[4] = nd;
    free(cr); free(cc); free(seen);
    return var;
}
