// tile_stream.cu -- kernels of the tile-stream multiply (see tile_stream.cuh).
#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "tile_stream.cuh"

namespace b2 {

// tile_row[t] = first row r with row_ptr[r] >= t*TS_TILE (lower bound over row_ptr[0..nRow]).
// Rows [tile_row[t], tile_row[t+1]) are OWNED by tile t: their first entry lies in it.  Empty
// rows are owned by the tile that contains their (shared) position; trailing rows by the last.
__global__ void tile_row_kernel(const int *__restrict__ ptr, int nRow, int nTiles, int tile, int *__restrict__ tile_row)
{
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t > nTiles) return;
    if (t == nTiles) {
        tile_row[t] = nRow;
        return;
    }
    int key = t * tile, lo = 0, hi = nRow + 1;
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (ptr[mid] < key) lo = mid + 1;
        else hi = mid;
    }
    tile_row[t] = lo;
}

// four consecutive matrix values as doubles: fp64 storage = two 128-bit loads, fp32 storage = one
__device__ __forceinline__ void ld_val4(const double *p, uint64_t pol, double &a, double &b, double &c, double &d)
{
    const double2 lo = ld_stream_d2(p, pol), hi = ld_stream_d2(p + 2, pol);
    a = lo.x; b = lo.y; c = hi.x; d = hi.y;
}
__device__ __forceinline__ void ld_val4(const float *p, uint64_t pol, double &a, double &b, double &c, double &d)
{
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p), "l"(pol));
    a = r.x; b = r.y; c = r.z; d = r.w;                 // fp32 -> fp64 is exact
}
__device__ __forceinline__ double ld_val1(const double *p, uint64_t pol) { return ld_stream_d1(p, pol); }
__device__ __forceinline__ double ld_val1(const float *p, uint64_t pol)
{
    float r;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(r) : "l"(p), "l"(pol));
    return (double)r;
}

// TMA = true: the tile's col / val slices are brought into shared memory by two 1-D bulk copies issued by one thread
// (cp.async.bulk, SASS UBLKCP) instead of through every thread's load pipeline.  An SM sustains about one L1-missing
// sector per clock (profiles/r2_gather_ceiling.md); on gather-bound matrices the x gathers need all of that, so the
// 12 B/nnz matrix stream is taken off it.  The products overwrite the staged values in place.
template <typename VT, int TH, bool TMA>
__global__ void __launch_bounds__(TH)
tile_stream_kernel(const int *__restrict__ row_ptr, const int *__restrict__ col,
                   const VT *__restrict__ val, const int *__restrict__ tile_row,
                   const double *__restrict__ x, double *__restrict__ y, double *__restrict__ carry,
                   int nnz, int tileLo, int rowLo, int rowHi, int accumulate, int vec_ok)
{
    constexpr int TILE = TH * TS_IPT;
    __shared__ __align__(16) double prod[TILE];
    __shared__ __align__(16) int scol[TMA ? TILE : 4];
    __shared__ __align__(16) VT sval32[(TMA && sizeof(VT) == 4) ? TILE : 4];   // fp32 values cannot share prod's slots
    __shared__ __align__(8) uint64_t bar;
    __shared__ int long_row[TILE / TS_LONG + 2];
    __shared__ int n_long;

    const int tid = threadIdx.x;
    const int t = tileLo + blockIdx.x;
    const int t0 = t * TILE;
    const int t1 = min(t0 + TILE, nnz);
    const uint64_t pol_stream = policy_evict_first();
    const uint64_t pol_x = policy_evict_last();
    if (tid == 0) n_long = 0;

    // ---- phase 1: stream the tile, gather x, park products in shared memory
    if (TMA) {
        VT *sval = sizeof(VT) == 8 ? reinterpret_cast<VT *>(prod) : sval32;
        const int n = t1 - t0;
        if (tid == 0) mbar_init(&bar, 1);
        __syncthreads();
        if (tid == 0) {
            const uint32_t bytesI = (uint32_t)((n * 4 + 15) & ~15), bytesV = (uint32_t)((n * (int)sizeof(VT) + 15) & ~15);
            mbar_expect_tx(&bar, bytesI + bytesV);
            tma_load_1d(scol, col + t0, bytesI, &bar, pol_stream);
            tma_load_1d(sval, val + t0, bytesV, &bar, pol_stream);
        }
        mbar_wait(&bar, 0);
        if (n == TILE) {
            double xs[TS_IPT];
            int4 c[TS_IPT / 4];
#pragma unroll
            for (int k = 0; k < TS_IPT / 4; k++) c[k] = *reinterpret_cast<const int4 *>(scol + 4 * (tid + k * TH));
#pragma unroll
            for (int k = 0; k < TS_IPT / 4; k++) {
                xs[4 * k + 0] = ld_x(x + c[k].x, pol_x);
                xs[4 * k + 1] = ld_x(x + c[k].y, pol_x);
                xs[4 * k + 2] = ld_x(x + c[k].z, pol_x);
                xs[4 * k + 3] = ld_x(x + c[k].w, pol_x);
            }
#pragma unroll
            for (int k = 0; k < TS_IPT / 4; k++) {
                const int o = 4 * (tid + k * TH);
                const double v0 = (double)sval[o], v1 = (double)sval[o + 1], v2 = (double)sval[o + 2], v3 = (double)sval[o + 3];
                double2 *dst = reinterpret_cast<double2 *>(prod + o);
                dst[0] = make_double2(__dmul_rn(v0, xs[4 * k]), __dmul_rn(v1, xs[4 * k + 1]));
                dst[1] = make_double2(__dmul_rn(v2, xs[4 * k + 2]), __dmul_rn(v3, xs[4 * k + 3]));
            }
        } else {
            for (int i = tid; i < n; i += TH) {
                const double v = (double)sval[i];
                prod[i] = __dmul_rn(v, ld_x(x + scol[i], pol_x));
            }
        }
    } else if (t1 - t0 == TILE && vec_ok) {
        int4 c[TS_IPT / 4];
        double2 v[TS_IPT / 2];
#pragma unroll
        for (int k = 0; k < TS_IPT / 4; k++) {
            const int e = t0 + 4 * (tid + k * TH);
            c[k] = ld_stream_i4(col + e, pol_stream);
            ld_val4(val + e, pol_stream, v[2 * k].x, v[2 * k].y, v[2 * k + 1].x, v[2 * k + 1].y);
        }
        double xs[TS_IPT];
#pragma unroll
        for (int k = 0; k < TS_IPT / 4; k++) {
            xs[4 * k + 0] = ld_x(x + c[k].x, pol_x);
            xs[4 * k + 1] = ld_x(x + c[k].y, pol_x);
            xs[4 * k + 2] = ld_x(x + c[k].z, pol_x);
            xs[4 * k + 3] = ld_x(x + c[k].w, pol_x);
        }
#pragma unroll
        for (int k = 0; k < TS_IPT / 4; k++) {
            double2 *dst = reinterpret_cast<double2 *>(prod + 4 * (tid + k * TH));
            dst[0] = make_double2(__dmul_rn(v[2 * k].x, xs[4 * k]), __dmul_rn(v[2 * k].y, xs[4 * k + 1]));
            dst[1] = make_double2(__dmul_rn(v[2 * k + 1].x, xs[4 * k + 2]),
                                  __dmul_rn(v[2 * k + 1].y, xs[4 * k + 3]));
        }
    } else {
        for (int i = tid; i < t1 - t0; i += TH)
            prod[i] = __dmul_rn(ld_val1(val + t0 + i, pol_stream),
                                ld_x(x + ld_stream_i1(col + t0 + i, pol_stream), pol_x));
    }

    const int r_lo = tile_row[t], r_hi = tile_row[t + 1];
    const int first = row_ptr[r_lo];          // r_lo <= nRow, row_ptr[nRow] = nnz
    const int cin_end = min(first, t1);       // [t0, cin_end) belongs to row r_lo-1
    __syncthreads();

    // ---- phase 2a: one thread per owned row, in-tile row-bin scheduler
    for (int r = r_lo + tid; r < r_hi; r += TH) {
        if (r < rowLo || r >= rowHi) continue;
        const int b = row_ptr[r], e_full = row_ptr[r + 1];
        const int e = min(e_full, t1);
        if (e_full > t1 && e_full - b <= TS_LONG) continue;   // short row crossing: fix-up recomputes it
        if (e - b > TS_LONG) {
            long_row[atomicAdd(&n_long, 1)] = r;
            continue;
        }
        double acc = accumulate == CS_CONTINUE ? y[r] : 0.0;          // CONTINUE: the running sum of the earlier column blocks
        for (int j = b; j < e; j++) acc = __dadd_rn(acc, prod[j - t0]);
        y[r] = accumulate == CS_ADD ? __dadd_rn(y[r], acc) : acc;
    }
    if (tid == 0 && cin_end > t0) {
        const int rc = r_lo - 1;
        if (first - row_ptr[rc] > TS_LONG && rc >= rowLo && rc < rowHi)
            long_row[atomicAdd(&n_long, 1)] = -1;
    }
    __syncthreads();

    // ---- phase 2b: one warp per long row / carried-in piece
    const int lane = tid & 31, warp = tid >> 5;
    const int nl = n_long;
    for (int i = warp; i < nl; i += TH / 32) {
        const int r = long_row[i];
        int b, e;
        if (r < 0) {
            b = t0;
            e = cin_end;
        } else {
            b = row_ptr[r];
            e = min(row_ptr[r + 1], t1);
        }
        double acc = 0.0;
        for (int j = b + lane; j < e; j += 32) acc += prod[j - t0];
        acc = warp_sum(acc);
        if (lane == 0) {
            if (r < 0) carry[t] = acc;
            else y[r] = accumulate ? y[r] + acc : acc;
        }
    }
}

// Finishes the rows that cross tile boundaries; one thread per tile, only the FIRST carrying
// tile of a row acts.  Short rows are recomputed from global memory in the reference's order.
template <typename VT>
__global__ void tile_fixup_kernel(const int *__restrict__ row_ptr, const int *__restrict__ col,
                                  const VT *__restrict__ val, const int *__restrict__ tile_row,
                                  const double *__restrict__ x, double *__restrict__ y,
                                  const double *__restrict__ carry, int tileLo, int tileHi, int rowLo,
                                  int rowHi, int accumulate, int tile)
{
    const int t = tileLo + blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= tileHi || t == 0) return;
    const int t0 = t * tile;
    const int r0 = tile_row[t];
    const int e = row_ptr[r0];
    if (e <= t0) return;                       // a row starts exactly at the tile start
    const int rc = r0 - 1;
    if (rc < rowLo || rc >= rowHi) return;
    const int b = row_ptr[rc];
    if (t != b / tile + 1) return;
    if (e - b <= TS_LONG) {
        double acc = accumulate == CS_CONTINUE ? y[rc] : 0.0;
        for (int j = b; j < e; j++) acc = __dadd_rn(acc, __dmul_rn((double)val[j], x[col[j]]));
        y[rc] = accumulate == CS_ADD ? __dadd_rn(y[rc], acc) : acc;
    } else {
        const int last = (e - 1) / tile;
        double sum = 0.0;
        for (int u = t; u <= last; u++) sum += carry[u];
        y[rc] += sum;
    }
}

int TileStream::build(const int *row_ptr_d, const int *col_d, const void *val_d, bool val_is_f32, int nRow_, int nnz_,
                      cudaStream_t s)
{
    f32 = val_is_f32;
    row_ptr = row_ptr_d;
    col = col_d;
    val = val_d;
    nRow = nRow_;
    nnz = nnz_;
    static const int env_threads = getenv("B200SPMV_TS_THREADS") ? atoi(getenv("B200SPMV_TS_THREADS")) : 256;
    threads = (env_threads == 128 || env_threads == 512) ? env_threads : 256;
    static const char *env_load = getenv("B200SPMV_TS_LOAD");           // "ldg" / "tma": how the tile's slices are read
    if (env_load) tma = !strcmp(env_load, "tma");
    tile = threads * TS_IPT;
    nTiles = ceil_div(nnz, tile);
    B2_TRY(tile_row.alloc((size_t)nTiles + 1));
    B2_TRY(carry.alloc((size_t)nTiles));
    tile_row_kernel<<<ceil_div(nTiles + 1, 256), 256, 0, s>>>(row_ptr, nRow, nTiles, tile, tile_row.p);
    B2_KERNEL_CHECK();
    range_cache.clear();
    return B200SPMV_OK;
}

int TileStream::run_rows(const double *x, double *y, int accumulate, int rb, int re, cudaStream_t s)
{
    if (rb < 0 || re > nRow || rb > re) {
        set_error("multiply_rows: bad row range [%d,%d) for %d rows", rb, re, nRow);
        return B200SPMV_ERR_INVALID;
    }
    if (rb == re) return B200SPMV_OK;
    if (rb == 0 && re == nRow) return run_all(x, y, accumulate, s);
    const auto key = std::make_pair(rb, re);
    auto it = range_cache.find(key);
    if (it == range_cache.end()) {
        // first use of this row range: two row pointers have to be read back.  b200spmv_prepare_rows() does this ahead
        // of time; here it is refused during stream capture (a synchronous copy would invalidate the capture)
        cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
        if (cudaStreamIsCapturing(s, &cap) == cudaSuccess && cap != cudaStreamCaptureStatusNone) {
            set_error("multiply_rows: row range [%d,%d) is used for the first time inside a stream capture; call "
                      "b200spmv_prepare_rows() (or run one eager multiply_rows) first", rb, re);
            return B200SPMV_ERR_STATE;
        }
        B2_TRY(prepare(rb, re));
        it = range_cache.find(key);
    }
    return run(x, y, accumulate, rb, re, it->second.first, it->second.second, s);
}

// tiles that hold any entry (or the shared position of empty rows) of rows [rb, re); synchronous, cached
int TileStream::prepare(int rb, int re)
{
    if (rb < 0 || re > nRow || rb > re) {
        set_error("prepare_rows: bad row range [%d,%d) for %d rows", rb, re, nRow);
        return B200SPMV_ERR_INVALID;
    }
    const auto key = std::make_pair(rb, re);
    if (range_cache.count(key)) return B200SPMV_OK;
    int pb = 0, pe = 0;
    B2_CUDA(cudaMemcpy(&pb, row_ptr + rb, sizeof(int), cudaMemcpyDeviceToHost));
    B2_CUDA(cudaMemcpy(&pe, row_ptr + re, sizeof(int), cudaMemcpyDeviceToHost));
    const int lo = nTiles ? std::min(pb / tile, nTiles - 1) : 0;
    const int hi = std::min(nTiles, pe / tile + 1);
    range_cache.emplace(key, std::make_pair(lo, hi));
    return B200SPMV_OK;
}

int TileStream::run(const double *x, double *y, int accumulate, int rowLo, int rowHi, int tileLo,
                    int tileHi, cudaStream_t s) const
{
    if (rowHi <= rowLo) return B200SPMV_OK;
    if (nTiles == 0) {   // no non-zeros at all: y = 0 (beta = 0)
        if (!accumulate) B2_CUDA(cudaMemsetAsync(y + rowLo, 0, sizeof(double) * (size_t)(rowHi - rowLo), s));
        return B200SPMV_OK;
    }
    const int vec_ok = ((reinterpret_cast<uintptr_t>(col) | reinterpret_cast<uintptr_t>(val)) & 15) == 0;
    const int nT = tileHi - tileLo, acc = accumulate;
    const bool fix = nT > 1 || tileLo > 0;
    // bulk copies need 16-byte aligned slices and may read up to 15 bytes past the last entry (callers keep CS_SLACK)
    const bool use_tma = tma && vec_ok;
    static bool carveout_set = false;
    if (use_tma && !carveout_set) {       // 8 resident CTAs x 24.7 KB: ask for the large shared-memory split once
        carveout_set = true;
        cudaFuncSetAttribute(tile_stream_kernel<double, 256, true>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
        cudaFuncSetAttribute(tile_stream_kernel<float, 256, true>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
        cudaFuncSetAttribute(tile_stream_kernel<double, 128, true>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
        cudaFuncSetAttribute(tile_stream_kernel<float, 128, true>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
        cudaGetLastError();
    }
#define TS_LAUNCH(VT, TH, v)                                                                                                      \
    do {                                                                                                                          \
        if (use_tma && TH <= 256) tile_stream_kernel<VT, (TH <= 256 ? TH : 256), true><<<nT, TH, 0, s>>>(row_ptr, col, v, tile_row.p, x, y, carry.p, nnz, tileLo, rowLo, rowHi, acc, vec_ok); \
        else tile_stream_kernel<VT, TH, false><<<nT, TH, 0, s>>>(row_ptr, col, v, tile_row.p, x, y, carry.p, nnz, tileLo, rowLo, rowHi, acc, vec_ok); \
        if (fix) tile_fixup_kernel<VT><<<ceil_div(nT, 256), 256, 0, s>>>(row_ptr, col, v, tile_row.p, x, y, carry.p, tileLo, tileHi, rowLo, rowHi, acc, tile); \
    } while (0)
    if (f32) {
        const float *v = static_cast<const float *>(val);
        if (threads == 128) TS_LAUNCH(float, 128, v);
        else if (threads == 512) TS_LAUNCH(float, 512, v);
        else TS_LAUNCH(float, 256, v);
    } else {
        const double *v = static_cast<const double *>(val);
        if (threads == 128) TS_LAUNCH(double, 128, v);
        else if (threads == 512) TS_LAUNCH(double, 512, v);
        else TS_LAUNCH(double, 256, v);
    }
    B2_KERNEL_CHECK();
    return B200SPMV_OK;
}

}  // namespace b2
