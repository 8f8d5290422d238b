/*
 * synth_oracle.c -- CPU statement of the deterministic synthetic inputs.
 *
 * TEST INFRASTRUCTURE ONLY (see spmv_oracle.c).  The product generates the same matrices on
 * the device (singlespmv_b200/csrc/synth.cu); tests compare the two bit-for-bit at mini
 * sizes, and bench.py's CPU legs use this file to build the bounded host-side sample.
 *
 * All matrices: square unless stated, COO sorted by (row, col), no duplicate coordinates,
 * int32 indices, fp64 values -- the input contract of the reference's plugins
 * (/root/reference/src/util.cpp:51 sorts; the plugins assume it).
 * Shapes follow /root/reference/matrix/artificial/generator.cpp:12-79 in spirit (band /
 * random patterns) but are defined here, not there: BASELINE.json names the five configs.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_API __attribute__((visibility("default")))

static inline uint64_t mix64(uint64_t z)          /* splitmix64 output function */
{
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
static inline double u01(uint64_t h) { return (double)(h >> 11) * (1.0 / 9007199254740992.0); }

/* value of a hashed-value matrix entry: a pure function of (seed, row, col) */
static inline double entry_value(uint64_t seed, int r, int c)
{
    return u01(mix64(mix64(seed ^ 0xA5A5A5A5A5A5A5A5ull) + (((uint64_t)(uint32_t)r << 32) | (uint32_t)c)));
}

/* ---- stencils: kind 0 = 2-D 5-point (n*n rows), 1 = 3-D 7-point, 2 = 3-D 27-point (n^3 rows).
 * Dirichlet (no wrap).  Diagonal = (#points-1), off-diagonals = -1.  Row-major grid index. */
ORC_API long long synth_stencil_rows(int kind, int n)
{
    return kind == 0 ? (long long)n * n : (long long)n * n * n;
}
ORC_API long long synth_stencil_nnz(int kind, int n)
{
    long long N = n;
    if (kind == 0) return 5 * N * N - 4 * N;
    if (kind == 1) return 7 * N * N * N - 6 * N * N;
    return (3 * N - 2) * (3 * N - 2) * (3 * N - 2);
}
ORC_API long long synth_stencil(int kind, int n, int *row, int *col, double *val)
{
    long long at = 0;
    if (kind == 0) {
        for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) {
            int r = i * n + j;
            for (int di = -1; di <= 1; di++) for (int dj = -1; dj <= 1; dj++) {
                if (di != 0 && dj != 0) continue;
                int ii = i + di, jj = j + dj;
                if (ii < 0 || ii >= n || jj < 0 || jj >= n) continue;
                row[at] = r; col[at] = ii * n + jj; val[at] = (di == 0 && dj == 0) ? 4.0 : -1.0; at++;
            }
        }
        return at;
    }
    for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) for (int k = 0; k < n; k++) {
        int r = (i * n + j) * n + k;
        for (int di = -1; di <= 1; di++) for (int dj = -1; dj <= 1; dj++) for (int dk = -1; dk <= 1; dk++) {
            int taxi = abs(di) + abs(dj) + abs(dk);
            if (kind == 1 && taxi > 1) continue;
            int ii = i + di, jj = j + dj, kk = k + dk;
            if (ii < 0 || ii >= n || jj < 0 || jj >= n || kk < 0 || kk >= n) continue;
            row[at] = r; col[at] = (ii * n + jj) * n + kk;
            val[at] = taxi == 0 ? (kind == 1 ? 6.0 : 26.0) : -1.0; at++;
        }
    }
    return at;
}

/* Rows [rowBegin, rowEnd) of the same stencil matrices (row ids stay global): what one rank of the
 * row-partitioned multi-GPU path owns, and the bounded CPU-baseline sample of bench.py.
 * The caller sizes the arrays for (rowEnd-rowBegin) * {5,7,27} entries.  Returns the count. */
ORC_API long long synth_stencil_range(int kind, int n, int rowBegin, int rowEnd, int *row, int *col, double *val)
{
    long long at = 0;
    for (int r = rowBegin; r < rowEnd; r++) {
        if (kind == 0) {
            int i = r / n, j = r % n;
            for (int di = -1; di <= 1; di++) for (int dj = -1; dj <= 1; dj++) {
                if (di != 0 && dj != 0) continue;
                int ii = i + di, jj = j + dj;
                if (ii < 0 || ii >= n || jj < 0 || jj >= n) continue;
                row[at] = r; col[at] = ii * n + jj; val[at] = (di == 0 && dj == 0) ? 4.0 : -1.0; at++;
            }
            continue;
        }
        int k = r % n, j = (r / n) % n, i = r / (n * n);
        for (int di = -1; di <= 1; di++) for (int dj = -1; dj <= 1; dj++) for (int dk = -1; dk <= 1; dk++) {
            int taxi = abs(di) + abs(dj) + abs(dk);
            if (kind == 1 && taxi > 1) continue;
            int ii = i + di, jj = j + dj, kk = k + dk;
            if (ii < 0 || ii >= n || jj < 0 || jj >= n || kk < 0 || kk >= n) continue;
            row[at] = r; col[at] = (ii * n + jj) * n + kk;
            val[at] = taxi == 0 ? (kind == 1 ? 6.0 : 26.0) : -1.0; at++;
        }
    }
    return at;
}

/* ---- uniform random: every row has exactly K distinct columns.  Row r draws candidates
 * c_j = mix64(rowkey + j) mod nCol, j = 0,1,2,..., keeping the first K distinct ones, then
 * sorts them.  Rows [rowBegin, rowEnd) are written (row ids stay global). */
ORC_API void synth_uniform(uint64_t seed, int nCol, int K, int rowBegin, int rowEnd,
                           int *row, int *col, double *val)
{
    int *pick = (int *)malloc(sizeof(int) * (size_t)K);
    long long at = 0;
    for (int r = rowBegin; r < rowEnd; r++) {
        uint64_t rowkey = mix64(mix64(seed) ^ (uint64_t)(uint32_t)r);
        int have = 0;
        for (uint64_t j = 0; have < K; j++) {
            int c = (int)(mix64(rowkey + j) % (uint64_t)nCol);
            int dup = 0;
            for (int t = 0; t < have; t++) if (pick[t] == c) { dup = 1; break; }
            if (!dup) pick[have++] = c;
        }
        for (int a = 1; a < K; a++) {                 /* insertion sort */
            int c = pick[a], b = a - 1;
            while (b >= 0 && pick[b] > c) { pick[b + 1] = pick[b]; b--; }
            pick[b + 1] = c;
        }
        for (int t = 0; t < K; t++) {
            row[at] = r; col[at] = pick[t]; val[at] = entry_value(seed, r, pick[t]); at++;
        }
    }
    free(pick);
}

/* ---- R-MAT (a,b,c,d) = (0.57,0.19,0.19,0.05) on 2^scale vertices, nEdges draws, duplicates
 * removed.  Edge e, level l: t = high 32 bits of mix64(edgekey + l); quadrant by integer
 * thresholds so host and device agree exactly. */
#define RMAT_A   2448131358u   /* floor(0.57 * 2^32) */
#define RMAT_AB  3264175144u   /* floor(0.76 * 2^32) */
#define RMAT_ABC 4080218931u   /* floor(0.95 * 2^32) */

static int cmp_u64(const void *a, const void *b)
{
    uint64_t x = *(const uint64_t *)a, y = *(const uint64_t *)b;
    return x < y ? -1 : x > y;
}

ORC_API long long synth_rmat(uint64_t seed, int scale, long long nEdges, int *row, int *col, double *val)
{
    uint64_t *key = (uint64_t *)malloc(sizeof(uint64_t) * (size_t)(nEdges ? nEdges : 1));
    for (long long e = 0; e < nEdges; e++) {
        uint64_t edgekey = mix64(mix64(seed) + (uint64_t)e);
        uint32_t r = 0, c = 0;
        for (int l = 0; l < scale; l++) {
            uint32_t t = (uint32_t)(mix64(edgekey + (uint64_t)l) >> 32);
            uint32_t rb = t >= RMAT_AB, cb = (t >= RMAT_A && t < RMAT_AB) || t >= RMAT_ABC;
            r = (r << 1) | rb; c = (c << 1) | cb;
        }
        key[e] = ((uint64_t)r << 32) | c;
    }
    qsort(key, (size_t)nEdges, sizeof(uint64_t), cmp_u64);
    long long nnz = 0;
    for (long long e = 0; e < nEdges; e++) {
        if (e && key[e] == key[e - 1]) continue;
        int r = (int)(key[e] >> 32), c = (int)(key[e] & 0xFFFFFFFFu);
        row[nnz] = r; col[nnz] = c; val[nnz] = entry_value(seed, r, c); nnz++;
    }
    free(key);
    return nnz;
}
