// coo.cu -- COO plugin (/root/reference/src/opt_coo.{h,cpp}): the converted matrix IS the sorted
// triplet list (the reference aliases the input arrays, opt_coo.cpp:14-19); multiply is y = 0 followed
// by an `omp atomic` scatter-add per entry (:34-46).
//
// B200 multiply: no atomics, no zero-fill pass.  The entry stream is cut into tiles of COO_TILE
// entries (one CTA each, equal bytes per CTA whatever the row lengths).  A CTA streams row/col/val
// with 128-bit loads, parks the products and the row ids in shared memory and reduces every run of
// equal row ids that STARTS in the tile: short runs one thread each, sequentially in storage order
// with unfused mul/add (bit-identical to opt_crs.cpp:61-67), long runs one warp each (shuffle tree).
// The thread that finds a run start also zero-fills the empty rows in front of it (beta = 0).  Runs
// crossing a tile boundary are finished by a second tiny kernel (short: recomputed sequentially;
// long: per-tile carries added in tile order -> deterministic).
#include "common.cuh"

namespace b2 {

constexpr int COO_THREADS = 256;
constexpr int COO_IPT = 8;
constexpr int COO_TILE = COO_THREADS * COO_IPT;
constexpr int COO_LONG = 64;
constexpr int COO_MAXLONG = COO_TILE / COO_LONG + 2;

// ACC (the COO tail of HYB, hyb.cu): y already holds the first part of every row's sum; the runs CONTINUE it
// (acc = y[r]; acc += ...; y[r] = acc), nothing is zero-filled, and a short run that crosses into the next tile is left
// entirely to the fix-up kernel (it restarts from the untouched y[r]).
template <bool ACC>
__global__ void __launch_bounds__(COO_THREADS)
coo_tile_kernel(const int *__restrict__ row, const int *__restrict__ col, const double *__restrict__ val,
                const double *__restrict__ x, double *__restrict__ y, double *__restrict__ carry, int nnz, int nRow,
                int vec_ok)
{
    __shared__ __align__(16) double prod[COO_TILE];
    __shared__ __align__(16) int srow[COO_TILE];
    __shared__ unsigned head[COO_TILE / 32 + 1];     // bit i = entry i starts a run of equal row ids
    __shared__ int long_start[COO_MAXLONG];
    __shared__ int n_long;

    const int tid = threadIdx.x, lane = tid & 31;
    const int t = blockIdx.x;
    const int t0 = t * COO_TILE;
    const int n = min(COO_TILE, nnz - t0);
    const uint64_t pol_stream = policy_evict_first(), pol_x = policy_evict_last();
    if (tid == 0) n_long = 0;
    const bool fast = n == COO_TILE && vec_ok;

    if (fast) {
#pragma unroll
        for (int k = 0; k < COO_IPT / 4; k++) {
            const int o = 4 * (tid + k * COO_THREADS);
            const int4 r = ld_stream_i4(row + t0 + o, pol_stream);
            const int4 c = ld_stream_i4(col + t0 + o, pol_stream);
            const double2 v0 = ld_stream_d2(val + t0 + o, pol_stream), v1 = ld_stream_d2(val + t0 + o + 2, pol_stream);
            const double x0 = ld_x(x + c.x, pol_x), x1 = ld_x(x + c.y, pol_x), x2 = ld_x(x + c.z, pol_x), x3 = ld_x(x + c.w, pol_x);
            *reinterpret_cast<int4 *>(srow + o) = r;
            double2 *dst = reinterpret_cast<double2 *>(prod + o);
            dst[0] = make_double2(__dmul_rn(v0.x, x0), __dmul_rn(v0.y, x1));
            dst[1] = make_double2(__dmul_rn(v1.x, x2), __dmul_rn(v1.y, x3));
            // run starts, in registers: the row id in front of this lane's four entries comes from the lane below
            int prev = __shfl_up_sync(0xffffffffu, r.w, 1);
            if (lane == 0) prev = (t0 + o) > 0 ? row[t0 + o - 1] : -1;
            unsigned bits = (unsigned)(r.x != prev) | ((unsigned)(r.y != r.x) << 1) | ((unsigned)(r.z != r.y) << 2) |
                            ((unsigned)(r.w != r.z) << 3);
            bits <<= 4 * (lane & 7);                           // eight lanes share one 32-entry word
            bits |= __shfl_xor_sync(0xffffffffu, bits, 1);
            bits |= __shfl_xor_sync(0xffffffffu, bits, 2);
            bits |= __shfl_xor_sync(0xffffffffu, bits, 4);
            if ((lane & 7) == 0) head[o >> 5] = bits;
        }
    } else {
        for (int i = tid; i < COO_TILE / 32 + 1; i += COO_THREADS) head[i] = 0u;
        for (int i = tid; i < n; i += COO_THREADS) {
            srow[i] = ld_stream_i1(row + t0 + i, pol_stream);
            prod[i] = __dmul_rn(ld_stream_d1(val + t0 + i, pol_stream), ld_x(x + ld_stream_i1(col + t0 + i, pol_stream), pol_x));
        }
    }
    const int prev_row = t0 > 0 ? row[t0 - 1] : -1;
    const bool last_tile = t0 + n == nnz;
    const int next_row = (ACC && !last_tile) ? row[t0 + n] : -1;
    __syncthreads();
    if (!fast) {
        for (int i = tid; i < n; i += COO_THREADS)
            if (srow[i] != (i ? srow[i - 1] : prev_row)) atomicOr(&head[i >> 5], 1u << (i & 31));
        __syncthreads();
    }

    // ---- one thread per run start among its eight consecutive entries (bits come from one shared word)
    const int base = tid * COO_IPT;
    if (base < n) {
        unsigned mine = (head[base >> 5] >> (base & 31)) & 0xFFu;
        while (mine) {
            const int i = base + __ffs(mine) - 1;
            mine &= mine - 1;
            // end of the run = next run start (or the end of the tile)
            int end;
            {
                int w = i >> 5;
                unsigned rest = head[w] & ~((2u << (i & 31)) - 1u);      // bits above i in its word
                while (!rest && (w + 1) * 32 < n) rest = head[++w];
                end = rest ? min(n, w * 32 + __ffs(rest) - 1) : n;
            }
            const int r = srow[i];
            if (!ACC) {
                const int before = i ? srow[i - 1] : prev_row;
                for (int e = before + 1; e < r; e++) y[e] = 0.0;    // empty rows in front of this run (beta = 0)
            }
            if (end - i > COO_LONG) {
                long_start[atomicAdd(&n_long, 1)] = i;
                continue;
            }
            if (ACC && end == n && next_row == r) {
                // the run goes on in the next tile.  At most COO_LONG entries in all: the fix-up recomputes it from the
                // untouched y[r]; longer: this piece is added now and the fix-up adds the carries of the other tiles
                const int probe = t0 + n + (COO_LONG - (n - i));
                if (!(probe < nnz && row[probe] == r)) continue;
            }
            double acc = ACC ? y[r] : 0.0;
            for (int j = i; j < end; j++) acc = __dadd_rn(acc, prod[j]);
            y[r] = acc;       // complete unless the run continues in the next tile (then the fix-up finishes it)
        }
    }
    if (tid == 0 && n > 0 && srow[0] == prev_row) long_start[atomicAdd(&n_long, 1)] = -1;   // carried-in piece
    if (!ACC && last_tile && n > 0)
        for (int e = srow[n - 1] + 1 + tid; e < nRow; e += COO_THREADS) y[e] = 0.0;          // trailing empty rows
    __syncthreads();

    // ---- one warp per long run / carried-in piece
    const int warp = tid >> 5;
    for (int q = warp; q < n_long; q += COO_THREADS / 32) {
        const int s = long_start[q];
        const int b = s < 0 ? 0 : s;
        const int r = srow[b];
        double acc = 0.0;
        for (int k = b + lane; k < n && srow[k] == r; k += 32) acc += prod[k];
        acc = warp_sum(acc);
        if (lane == 0) {
            if (s < 0) carry[t] = acc;
            else y[r] = ACC ? y[r] + acc : acc;
        }
    }
}

// one thread per tile: finishes the row whose entries continue from the previous tile (first carrying tile only)
template <bool ACC>
__global__ void coo_fixup_kernel(const int *__restrict__ row, const int *__restrict__ col, const double *__restrict__ val,
                                 const double *__restrict__ x, double *__restrict__ y, const double *__restrict__ carry,
                                 int nnz, int nTiles)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < 1 || t >= nTiles) return;
    const int t0 = t * COO_TILE;
    const int rc = row[t0];
    if (row[t0 - 1] != rc) return;                              // a run starts exactly at the tile start
    const int p0 = t0 - COO_TILE;                               // previous tile: does the run merely pass through it?
    if (row[p0] == rc && (p0 == 0 ? false : row[p0 - 1] == rc)) return;
    int b = t0, e = t0;
    while (b > 0 && t0 - b <= COO_LONG && row[b - 1] == rc) b--;
    while (e < nnz && e - t0 <= COO_LONG && row[e] == rc) e++;
    const bool whole = (b == 0 || row[b - 1] != rc) && (e == nnz || row[e] != rc);
    if (whole && e - b <= COO_LONG) {
        double acc = ACC ? y[rc] : 0.0;
        for (int j = b; j < e; j++) acc = __dadd_rn(acc, __dmul_rn(val[j], x[col[j]]));
        y[rc] = acc;
    } else {
        double sum = 0.0;
        for (int u = t; u < nTiles && row[u * COO_TILE] == rc; u++) sum += carry[u];
        y[rc] += sum;
    }
}

struct CooFormat : Format {
    DevBuf<int> row, col;
    DevBuf<double> val, carry;
    int nTiles = 0;

    int convert(const CooView &A, cudaStream_t s) override
    {
        nRow = A.nRow; nCol = A.nCol; nnz = A.nnz;
        B2_TRY(validate_sorted_coo(A, s));
        B2_TRY(row.alloc((size_t)nnz));
        B2_TRY(col.alloc((size_t)nnz));
        B2_TRY(val.alloc((size_t)nnz));
        B2_CUDA(cudaMemcpyAsync(row.p, A.row, row.bytes(), cudaMemcpyDeviceToDevice, s));   // opt_coo.cpp:14-19
        B2_CUDA(cudaMemcpyAsync(col.p, A.col, col.bytes(), cudaMemcpyDeviceToDevice, s));
        B2_CUDA(cudaMemcpyAsync(val.p, A.val, val.bytes(), cudaMemcpyDeviceToDevice, s));
        nTiles = ceil_div(nnz, COO_TILE);
        B2_TRY(carry.alloc((size_t)nTiles));
        B2_CUDA(cudaStreamSynchronize(s));
        return B200SPMV_OK;
    }

    int multiply(const double *x, double *y, cudaStream_t s) override
    {
        if (nRow == 0) return B200SPMV_OK;
        if (nTiles == 0) {
            B2_CUDA(cudaMemsetAsync(y, 0, sizeof(double) * (size_t)nRow, s));
            return B200SPMV_OK;
        }
        coo_tile_kernel<false><<<nTiles, COO_THREADS, 0, s>>>(row.p, col.p, val.p, x, y, carry.p, nnz, nRow, 1);
        B2_KERNEL_CHECK();
        if (nTiles > 1) {
            coo_fixup_kernel<false><<<ceil_div(nTiles, 256), 256, 0, s>>>(row.p, col.p, val.p, x, y, carry.p, nnz, nTiles);
            B2_KERNEL_CHECK();
        }
        return B200SPMV_OK;
    }

    bool scalar(const std::string &n, long long *out) override
    {
        if (n == "alg_bytes") {   // 16 nnz + 8 nCol + 8 nRow
            *out = 16LL * nnz + 8LL * nCol + 8LL * nRow;
            return true;
        }
        if (n == "launches") { *out = nTiles > 1 ? 2 : 1; return true; }
        if (n == "nTiles") { *out = nTiles; return true; }
        return false;
    }

    long long array(const std::string &n, void *dst, long long cap) override
    {
        if (n == "row_idx") return export_device(row.p, row.bytes(), dst, cap);
        if (n == "col_idx") return export_device(col.p, col.bytes(), dst, cap);
        if (n == "val") return export_device(val.p, val.bytes(), dst, cap);
        return -1000;
    }
};

Format *make_coo(const b200spmv_options &) { return new CooFormat(); }

// y[r] continues with the sorted triplets' products, run by run (HYB's COO tail, hyb.cu); carry: ceil(nnz / COO_TILE) doubles
int coo_accumulate(const int *row, const int *col, const double *val, int nnz, int nRow, const double *x, double *y,
                   double *carry, cudaStream_t s)
{
    if (nnz == 0) return B200SPMV_OK;
    const int nTiles = ceil_div(nnz, COO_TILE);
    const int vec_ok = ((reinterpret_cast<uintptr_t>(row) | reinterpret_cast<uintptr_t>(col) | reinterpret_cast<uintptr_t>(val)) & 15) == 0;
    coo_tile_kernel<true><<<nTiles, COO_THREADS, 0, s>>>(row, col, val, x, y, carry, nnz, nRow, vec_ok);
    B2_KERNEL_CHECK();
    if (nTiles > 1) {
        coo_fixup_kernel<true><<<ceil_div(nTiles, 256), 256, 0, s>>>(row, col, val, x, y, carry, nnz, nTiles);
        B2_KERNEL_CHECK();
    }
    return B200SPMV_OK;
}
int coo_tile_entries() { return COO_TILE; }

}  // namespace b2
