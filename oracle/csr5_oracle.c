/*
 * csr5_oracle.c -- CPU restatement of the CSR -> CSR5 conversion (Liu & Vinter, ICS'15) as vendored in
 * the reference under opt/Benchmark_SpMV_using_CSR5/, with omega = 32 (the GPU tile width,
 * CSR5_cuda/detail/cuda/common_cuda.h:11), plus a straightforward CSR5 multiply over those arrays.
 *
 * TEST INFRASTRUCTURE ONLY (see spmv_oracle.c).
 *
 * Parity status: PINNED.  The executable reference is the AVX2 twin of the conversion
 * (CSR5_avx2/detail/avx2/format_avx2.h:8-458), compiled unmodified with omega overridden to 32 by
 * oracle/ref_csr5_shim.cpp into oracle/_ref/libref_csr5.so; tests/test_oracle.py compares every array
 * bit-for-bit on the fixtures and on randomized inputs, and tests/golden/csr5_*.npz hold its outputs.
 * (The CUDA twin, CSR5_cuda/detail/cuda/format_cuda.h, cannot be built here: CUDA-samples headers and
 * pre-Volta shuffles.  Where the two differ -- the empty-row test of the tile pointer covers rows
 * [start, stop] here but [start, stop) there, format_avx2.h:48-55 vs format_cuda.h:74-83 -- this file
 * follows the AVX2 form, the one that can be executed.)
 *
 * Citations: file:line under /root/reference/opt/Benchmark_SpMV_using_CSR5/.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_API __attribute__((visibility("default")))
#define OMEGA 32

/* CSR5_cuda/anonymouslib_cuda.h:293-317 -- auto-tuned sigma from the mean row length. */
ORC_API int orc_csr5_auto_sigma(int m, int nnz)
{
    int per_row = m > 0 ? nnz / m : 0;
    if (per_row <= 4) return 4;
    if (per_row <= 32) return per_row;
    if (per_row <= 256) return 32;
    return 6;
}

/* CSR5_avx2/anonymouslib_avx2.h:121-137 -- field widths, packets per lane, number of tiles. */
ORC_API void orc_csr5_shape(int nnz, int sigma, int *bit_y_offset, int *bit_scansum_offset, int *num_packet, int *p)
{
    int base = 2, by = 1, bs = 1;
    while (base < OMEGA * sigma) { base *= 2; by++; }
    base = 2;
    while (base < OMEGA) { base *= 2; bs++; }
    *bit_y_offset = by;
    *bit_scansum_offset = bs;
    *num_packet = (by + bs + sigma + 31) / 32;
    *p = (int)(((long long)nnz + (long long)OMEGA * sigma - 1) / ((long long)OMEGA * sigma));
}

/* CSR5_avx2/detail/avx2/utils_avx2.h:24-46 -- number of entries <= key. */
static int count_le(const int *a, int key, int size)
{
    int lo = 0, hi = size - 1;
    while (hi >= lo) {
        int mid = (hi + lo) / 2;
        if (key >= a[mid]) lo = mid + 1;
        else hi = mid - 1;
    }
    return lo;
}

/* format_avx2.h:8-61 -- tile_ptr[t] = row holding non-zero min(t*omega*sigma, nnz); MSB = the tile's
 * row span [start, stop] contains an empty row (row index m is never tested: the reference reads one
 * element past row_ptr there, ref_csr5_shim.cpp pads it with -1). */
ORC_API void orc_csr5_tile_ptr(int m, int nnz, int sigma, int p, const int *row_ptr, uint32_t *tile_ptr)
{
    for (int t = 0; t <= p; t++) {
        long long b = (long long)t * sigma * OMEGA;
        int boundary = b > nnz ? nnz : (int)b;
        tile_ptr[t] = (uint32_t)(count_le(row_ptr, boundary, m + 1) - 1);
    }
    for (int t = 0; t < p; t++) {
        uint32_t start = tile_ptr[t] & 0x7FFFFFFFu, stop = tile_ptr[t + 1] & 0x7FFFFFFFu;
        if (start == stop) continue;
        for (uint32_t r = start; r <= stop && r < (uint32_t)m; r++)
            if (row_ptr[r] == row_ptr[r + 1]) { tile_ptr[t] = start | 0x80000000u; break; }
    }
}

static inline uint32_t lane_flags(const uint32_t *desc_tile, int lane, int num_packet, int bit_all, int sigma)
{
    /* the sigma bit flags of one lane, MSB-first across its packets; returned with step i at bit (31-i) */
    uint64_t w = (uint64_t)desc_tile[lane] << 32;
    if (num_packet > 1) w |= desc_tile[OMEGA + lane];
    w <<= bit_all;
    uint32_t f = (uint32_t)(w >> 32);
    if (sigma < 32) f &= ~((1u << (32 - sigma)) - 1u);
    return f;
}

/* format_avx2.h:88-277 -- bit flags (s1), y_offset / scansum_offset (s2), offsets of empty-row tiles.
 * desc: p*omega*num_packet words, zero-initialised here.  offset_ptr: p+1 entries.
 * Returns num_offsets; call again with offset != NULL to fill the y-index table (format_avx2.h:279-364). */
ORC_API int orc_csr5_descriptor(int m, int nnz, int sigma, int p, int bit_y_offset, int bit_scansum_offset,
                                int num_packet, const int *row_ptr, const uint32_t *tile_ptr, uint32_t *desc,
                                int *offset_ptr, int *offset)
{
    const int bit_all = bit_y_offset + bit_scansum_offset;
    const int T = OMEGA * sigma;
    if (!offset) {
        memset(desc, 0, sizeof(uint32_t) * (size_t)p * OMEGA * num_packet);
        memset(offset_ptr, 0, sizeof(int) * ((size_t)p + 1));
        /* s1: one flag per row start that falls into a full tile */
        for (int r = 0; r <= m; r++) {
            int ptr = row_ptr[r], pid = ptr / T;
            if (pid >= p - 1) continue;
            int lx = (ptr / sigma) % OMEGA, glid = ptr % sigma + bit_all;
            desc[(size_t)pid * OMEGA * num_packet + (size_t)(glid / 32) * OMEGA + lx] |= 1u << (31 - glid % 32);
        }
        /* s2 */
        for (int t = 0; t < p - 1; t++) {
            uint32_t start = tile_ptr[t] & 0x7FFFFFFFu, stop = tile_ptr[t + 1] & 0x7FFFFFFFu;
            if (start == stop) continue;
            uint32_t *dt = desc + (size_t)t * OMEGA * num_packet;
            int segn[OMEGA + 1], present[OMEGA + 1];
            for (int l = 0; l < OMEGA; l++) {
                uint32_t f = lane_flags(dt, l, num_packet, bit_all, sigma);
                if (l == 0) f |= 0x80000000u;
                segn[l] = __builtin_popcount(f);
                present[l] = f != 0;
            }
            present[OMEGA] = 1;
            int run = 0;
            for (int l = 0; l <= OMEGA; l++) { int c = l < OMEGA ? segn[l] : 0; segn[l] = run; run += c; }
            if (tile_ptr[t] >> 31) offset_ptr[t] = segn[OMEGA];
            for (int l = 0; l < OMEGA; l++) {
                int y_offset = l ? segn[l] - 1 : 0, scansum = 0;
                if (present[l])
                    for (int k = l + 1; k < OMEGA && !present[k]; k++) scansum++;
                dt[l] |= (uint32_t)y_offset << (32 - bit_y_offset);
                dt[l] |= (uint32_t)scansum << (32 - bit_all);
            }
        }
        int run = 0;
        for (int t = 0; t <= p; t++) { int c = t < p ? offset_ptr[t] : 0; offset_ptr[t] = run; run += c; }
        return offset_ptr[p];
    }
    for (int t = 0; t < p - 1; t++) {
        if (!(tile_ptr[t] >> 31)) continue;
        int start = (int)(tile_ptr[t] & 0x7FFFFFFFu), stop = (int)(tile_ptr[t + 1] & 0x7FFFFFFFu);
        const uint32_t *dt = desc + (size_t)t * OMEGA * num_packet;
        for (int l = 0; l < OMEGA; l++) {
            int y_offset = (int)(dt[l] >> (32 - bit_y_offset));
            uint32_t f = lane_flags(dt, l, num_packet, bit_all, sigma);
            for (int i = 0; i < sigma; i++) {
                if (!((f >> (31 - i)) & 1u) || (l == 0 && i == 0)) continue;   /* lane 0 / step 0 is never listed */
                int idx = t * T + l * sigma + i;
                offset[offset_ptr[t] + y_offset] = count_le(row_ptr + start + 1, idx, stop - start) - 1;
                y_offset++;
            }
        }
    }
    return offset_ptr[p];
}

/* format_avx2.h:366-437 -- inside every full tile whose RAW tile_ptr differs from its successor's,
 * entry (lane l, step i) moves from l*sigma+i to i*omega+l.  In place. */
ORC_API void orc_csr5_transpose(int nnz, int sigma, int p, const uint32_t *tile_ptr, int *col, double *val)
{
    const int T = OMEGA * sigma;
    int *ci = (int *)malloc(sizeof(int) * (size_t)T);
    double *cv = (double *)malloc(sizeof(double) * (size_t)T);
    (void)nnz;
    for (int t = 0; t < p - 1; t++) {
        if (tile_ptr[t] == tile_ptr[t + 1]) continue;
        int *c = col + (size_t)t * T;
        double *v = val + (size_t)t * T;
        for (int idx = 0; idx < T; idx++) {
            int l = idx / sigma, i = idx % sigma;
            ci[i * OMEGA + l] = c[idx];
            cv[i * OMEGA + l] = v[idx];
        }
        memcpy(c, ci, sizeof(int) * (size_t)T);
        memcpy(v, cv, sizeof(double) * (size_t)T);
    }
    free(ci); free(cv);
}

/* A plain multiply over the CSR5 arrays, following the structure of
 * CSR5_cuda/detail/cuda/csr5_spmv_cuda.h:59-200,313-419 (segments closed at bit flags, first partial of a
 * tile added to y[tile_ptr[t]], tail tile as CSR), but with true overwrite semantics (y is zeroed
 * first; upstream relies on a one-time memset, CSR5_cuda/main.cu:57).  Serial, ascending order. */
ORC_API void orc_csr5_spmv(int m, int nnz, int sigma, int p, int bit_y_offset, int bit_scansum_offset,
                           int num_packet, const int *row_ptr, const uint32_t *tile_ptr, const uint32_t *desc,
                           const int *offset_ptr, const int *offset, const int *col, const double *val,
                           const double *x, double *y)
{
    const int bit_all = bit_y_offset + bit_scansum_offset, T = OMEGA * sigma;
    for (int r = 0; r < m; r++) y[r] = 0;
    for (int t = 0; t < p - 1; t++) {
        int start = (int)(tile_ptr[t] & 0x7FFFFFFFu), stop = (int)(tile_ptr[t + 1] & 0x7FFFFFFFu);
        const int *c = col + (size_t)t * T;
        const double *v = val + (size_t)t * T;
        int transposed = tile_ptr[t] != tile_ptr[t + 1];
        if (start == stop) {
            double s = 0;
            for (int k = 0; k < T; k++) s += v[k] * x[c[k]];
            y[start] += s;
            continue;
        }
        const uint32_t *dt = desc + (size_t)t * OMEGA * num_packet;
        int dirty = (int)(tile_ptr[t] >> 31), seg = -1;      /* seg = index of the open segment; -1 = carried-in */
        double s = 0;
        for (int l = 0; l < OMEGA; l++) {
            uint32_t f = lane_flags(dt, l, num_packet, bit_all, sigma);
            for (int i = 0; i < sigma; i++) {
                if (((f >> (31 - i)) & 1u) && !(l == 0 && i == 0)) {
                    int row = seg < 0 ? start : start + 1 + (dirty ? offset[offset_ptr[t] + seg] : seg);
                    y[row] += s;
                    s = 0;
                    seg++;
                }
                int k = transposed ? i * OMEGA + l : l * sigma + i;
                s += v[k] * x[c[k]];
            }
        }
        y[seg < 0 ? start : start + 1 + (dirty ? offset[offset_ptr[t] + seg] : seg)] += s;
    }
    if (p > 0) {
        int r0 = (int)(tile_ptr[p - 1] & 0x7FFFFFFFu);
        for (int r = r0; r < m; r++) {
            int b = r == r0 ? (p - 1) * T : row_ptr[r];
            double s = 0;
            for (int j = b; j < row_ptr[r + 1]; j++) s += val[j] * x[col[j]];
            y[r] += s;
        }
    }
}
