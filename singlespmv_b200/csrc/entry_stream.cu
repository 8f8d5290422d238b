// entry_stream.cu -- kernels and launcher of the CRS entry stream (see entry_stream.cuh).
#include <algorithm>
#include <cstdlib>

#include "entry_stream.cuh"

namespace b2 {

constexpr int ES_THREADS = 256;
constexpr int ES_WARPS = ES_THREADS / 32;

// ---------------------------------------------------------------- conversion: row_ptr -> start bits, run counts, row lists
__global__ void es_bits_kernel(const int *__restrict__ ptr, int nRow, unsigned *__restrict__ bits, int *__restrict__ empty,
                               int *__restrict__ nEmpty)
{
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nRow) return;
    const int b = ptr[r], e = ptr[r + 1];
    if (e > b) atomicOr(&bits[b >> 5], 1u << (b & 31));
    else if (empty) empty[atomicAdd(nEmpty, 1)] = r;
}
__global__ void es_empty_count_kernel(const int *__restrict__ ptr, int nRow, int *__restrict__ nEmpty)
{
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    int c = (r < nRow && ptr[r + 1] == ptr[r]) ? 1 : 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(nEmpty, c);
}
// cnt[g] = row starts among entries [128 g, 128 g + 128); cnt[nGroups] = 0 (scan sentinel)
__global__ void es_group_count_kernel(const unsigned *__restrict__ bits, int nGroups, int *__restrict__ cnt)
{
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g > nGroups) return;
    int c = 0;
    if (g < nGroups) {
        const uint4 w = *reinterpret_cast<const uint4 *>(bits + 4 * (size_t)g);
        c = __popc(w.x) + __popc(w.y) + __popc(w.z) + __popc(w.w);
    }
    cnt[g] = c;
}
// nzrow[k] = the k-th non-empty row (k = number of row starts in front of its first entry)
__global__ void es_nzrow_kernel(const int *__restrict__ ptr, int nRow, const unsigned *__restrict__ bits,
                                const int *__restrict__ grpRun, int *__restrict__ nzrow)
{
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nRow) return;
    const int p = ptr[r];
    if (ptr[r + 1] == p) return;
    const int g = p >> 7;
    int k = grpRun[g];
    for (int w = 4 * g; w < (p >> 5); w++) k += __popc(bits[w]);
    k += __popc(bits[p >> 5] & ((1u << (p & 31)) - 1u));
    nzrow[k] = r;
}
__global__ void es_zero_rows_kernel(const int *__restrict__ rows, int n, int rowLo, int rowHi, double *__restrict__ y)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int r = rows[i];
    if (r >= rowLo && r < rowHi) y[r] = 0.0;
}

// ---------------------------------------------------------------- multiply
struct EsChunkRec {
    double piece, tail;         // sum in front of the chunk's first row start (whole chunk if none); sum of the run open at its end
    int tailRun, pad;           // that run's index, -1 if no row starts in the chunk
};

// TMA = true: persistent CTAs, (idx, val) of tile k+1 arriving by two bulk copies while tile k is reduced.
// TMA = false: one tile per CTA, the lanes load their entries themselves (gather-bound matrices: x must stay in L2).
template <int E, bool TMA>
__global__ void __launch_bounds__(ES_THREADS)
entry_stream_kernel(const int *__restrict__ idx, const double *__restrict__ val, const unsigned *__restrict__ bits,
                    const int *__restrict__ grpRun, const int *__restrict__ nzrow, const double *__restrict__ x,
                    double *__restrict__ y, double *__restrict__ carry, int nnz, int tileLo, int tileHi, int rowLo, int rowHi)
{
    constexpr int CHUNK = E / ES_WARPS;
    constexpr int GROUPS = CHUNK / 128;
    static_assert(GROUPS >= 1, "a warp pass covers 128 entries");
    extern __shared__ __align__(128) unsigned char es_smem[];
    __shared__ __align__(8) uint64_t full[2];
    __shared__ EsChunkRec rec[2][ES_WARPS];
    int *scol = reinterpret_cast<int *>(es_smem);                       // [2][E]
    double *sval = reinterpret_cast<double *>(es_smem + 2 * E * sizeof(int));   // [2][E]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint64_t pol_stream = policy_evict_first(), pol_x = policy_evict_last();
    const unsigned lt = (1u << lane) - 1u;

    auto store = [&](int run, double v) {                   // y of the run-th non-empty row, if the caller asked for that row
        const int r = nzrow ? nzrow[run] : run;
        if (r >= rowLo && r < rowHi) y[r] = v;
    };
    auto issue = [&](int s, int t) {                        // thread 0 only
        const long long t0 = (long long)t * E;
        const int n = (int)min((long long)E, (long long)nnz - t0);
        const uint32_t b4 = (uint32_t)((n * 4 + 15) & ~15), b8 = (uint32_t)(n * 8 + 15) & ~15u;   // the arrays carry slack
        mbar_expect_tx(&full[s], b4 + b8);
        tma_load_1d(scol + s * E, idx + t0, b4, &full[s], pol_stream);
        tma_load_1d(sval + s * E, val + t0, b8, &full[s], pol_stream);
    };
    const int first = tileLo + blockIdx.x;
    if (TMA) {
        if (tid == 0) { mbar_init(&full[0], 1); mbar_init(&full[1], 1); }
        __syncthreads();
        if (tid == 0) {
            if (first < tileHi) issue(0, first);
            if (first + (int)gridDim.x < tileHi) issue(1, first + gridDim.x);
        }
    }
    int k = 0;
    for (int t = first; t < tileHi; t += gridDim.x, k++) {
        const int s = k & 1;
        const long long t0 = (long long)t * E;
        const int n = (int)min((long long)E, (long long)nnz - t0);
        const int c0 = warp * CHUNK;
        const int nGroups = c0 < n ? min(GROUPS, (n - c0 + 127) >> 7) : 0;
        const int nw = min(ES_WARPS, (n + CHUNK - 1) / CHUNK);
        if (TMA) mbar_wait(&full[s], (uint32_t)(k >> 1) & 1u);
        const int *Cc = TMA ? scol + s * E : idx + t0;
        const double *V = TMA ? sval + s * E : val + t0;
        if (nGroups > 0) {
            bool started = false;                           // a row has started in this chunk
            double cin = 0.0, piece = 0.0;
            int lastRun = -1;
#pragma unroll
            for (int g = 0; g < GROUPS; g++) {
                if (g >= nGroups) break;
                const int e = c0 + g * 128 + 4 * lane;
                const long long ge = t0 + e;                // global entry index of the lane's first entry
                int4 c;
                double2 v0, v1;
                if (c0 + g * 128 + 128 <= n) {
                    if (TMA) {
                        c = *reinterpret_cast<const int4 *>(Cc + e);
                        v0 = *reinterpret_cast<const double2 *>(V + e);
                        v1 = *reinterpret_cast<const double2 *>(V + e + 2);
                    } else {
                        c = ld_stream_i4(Cc + e, pol_stream);
                        v0 = ld_stream_d2(V + e, pol_stream);
                        v1 = ld_stream_d2(V + e + 2, pol_stream);
                    }
                } else {                                    // ragged end of the last tile: missing entries are zeros of the open row
                    c.x = e < n ? Cc[e] : 0; c.y = e + 1 < n ? Cc[e + 1] : 0; c.z = e + 2 < n ? Cc[e + 2] : 0; c.w = e + 3 < n ? Cc[e + 3] : 0;
                    v0.x = e < n ? V[e] : 0.0; v0.y = e + 1 < n ? V[e + 1] : 0.0; v1.x = e + 2 < n ? V[e + 2] : 0.0; v1.y = e + 3 < n ? V[e + 3] : 0.0;
                }
                const unsigned word = __ldg(bits + (ge >> 5));                  // beyond nnz the bits are zero
                const unsigned b = (word >> (unsigned)(ge & 31)) & 0xFu;
                const int base = __ldg(grpRun + (ge >> 7));                     // row starts in front of the group
                const double x0 = ld_x(x + c.x, pol_x), x1 = ld_x(x + c.y, pol_x), x2 = ld_x(x + c.z, pol_x), x3 = ld_x(x + c.w, pol_x);
                const double p0 = __dmul_rn(v0.x, x0), p1 = __dmul_rn(v0.y, x1), p2 = __dmul_rn(v1.x, x2), p3 = __dmul_rn(v1.y, x3);
                const bool b0 = b & 1u, b1 = b & 2u, b2 = b & 4u, b3 = b & 8u;
                const bool has = b != 0u;
                const int nStarts = __popc(b);
                // starts in the lower lanes of the group -> index of the run that is open in front of this lane's entries
                const unsigned m = __ballot_sync(0xffffffffu, has);             // lanes that hold a start
                int runOpen = base - 1 + __popc(m & lt);                        // right when no lane holds two starts
                // head = sum of the entries before the lane's first start, acc = sum from its last start on
                double head, acc;
                if (__any_sync(0xffffffffu, nStarts > 1)) {
                    const unsigned m0 = __ballot_sync(0xffffffffu, b0), m1 = __ballot_sync(0xffffffffu, b1),
                                   m2 = __ballot_sync(0xffffffffu, b2), m3 = __ballot_sync(0xffffffffu, b3);
                    runOpen = base - 1 + __popc(m0 & lt) + __popc(m1 & lt) + __popc(m2 & lt) + __popc(m3 & lt);
                    head = 0.0; acc = 0.0;
                    bool seen = false;
                    int run = runOpen;
                    auto step = [&](bool st, double p) {
                        if (st) {
                            if (!seen) head = acc;
                            else store(run, acc);           // a row that begins and ends inside this lane
                            seen = true;
                            run++;
                            acc = p;
                        } else acc = __dadd_rn(acc, p);
                    };
                    step(b0, p0); step(b1, p1); step(b2, p2); step(b3, p3);
                } else {
                    const bool f0 = b0, f1 = f0 | b1, f2 = f1 | b2;             // "a start at or before this entry"
                    head = __dadd_rn(__dadd_rn((has && !f0) ? p0 : 0.0, (has && !f1) ? p1 : 0.0), (has && !f2) ? p2 : 0.0);
                    acc = __dadd_rn(__dadd_rn(__dadd_rn((!has || f0) ? p0 : 0.0, (!has || f1) ? p1 : 0.0), (!has || f2) ? p2 : 0.0), p3);
                }
                // segmented inclusive scan over the lanes: segments begin at lanes that hold a start
                const unsigned below = m & (0xffffffffu >> (31 - lane));
                const int seg = below ? 31 - __clz(below) : 0;
                double v = acc;
                // a segment of k lanes needs the steps d < k only: find the longest gap between starts (with a virtual start at
                // lane 0) from the warp-uniform mask -- 7-entry rows (c5) need ONE step, 27-entry rows (c4) three
                unsigned cov = m | 1u;
                cov |= cov << 1;
                int steps = 1;
                if (cov != 0xffffffffu) {
                    cov |= cov << 2; steps = 2;
                    if (cov != 0xffffffffu) {
                        cov |= cov << 4; steps = 3;
                        if (cov != 0xffffffffu) { cov |= cov << 8; steps = cov != 0xffffffffu ? 5 : 4; }
                    }
                }
#pragma unroll
                for (int i = 0; i < 5; i++) {
                    if (i < steps) {
                        const int d = 1 << i;
                        const double u = __shfl_up_sync(0xffffffffu, v, d);
                        if (lane - d >= seg) v = __dadd_rn(v, u);
                    }
                }
                if (!below) v = __dadd_rn(v, cin);
                double excl = __shfl_up_sync(0xffffffffu, v, 1);
                if (lane == 0) excl = cin;
                if (has) {
                    const double total = __dadd_rn(excl, head);                 // the run that was open in front of this lane ends here
                    const bool firstInChunk = !started && !(m & lt);
                    if (!firstInChunk) store(runOpen, total);
                    else piece = total;                     // continues a row of an earlier chunk (0 if the chunk begins with a start)
                }
                if (!started && m) piece = __shfl_sync(0xffffffffu, piece, __ffs(m) - 1);
                cin = __shfl_sync(0xffffffffu, v, 31);
                lastRun = __shfl_sync(0xffffffffu, runOpen + nStarts, 31);
                started = started || m != 0u;
            }
            if (lane == 0) {
                EsChunkRec q;
                q.piece = started ? piece : cin;
                q.tail = cin;
                q.tailRun = started ? lastRun : -1;
                q.pad = 0;
                rec[s][warp] = q;
            }
        }
        __syncthreads();                                    // every warp is done with stage s, the chunk records are in place
        if (tid == 32) {
            double tilePiece = 0.0, a2 = 0.0;
            int accRun = -1;
            for (int w = 0; w < nw; w++) {
                const EsChunkRec *o = &rec[s][w];
                const double oPiece = o->piece, oTail = o->tail;
                const int oRun = o->tailRun;
                if (accRun >= 0) a2 = __dadd_rn(a2, oPiece);
                else tilePiece = __dadd_rn(tilePiece, oPiece);
                if (oRun >= 0) {
                    if (accRun >= 0) store(accRun, a2);
                    accRun = oRun;
                    a2 = oTail;
                }
            }
            if (accRun >= 0) store(accRun, a2);
            carry[t] = tilePiece;
        }
        if (TMA && tid == 0 && t + 2 * (int)gridDim.x < tileHi) issue(s, t + 2 * gridDim.x);
    }
}

// one thread per tile: a row that runs through several tiles gets their leading pieces added, in tile order
__global__ void entry_stream_fixup_kernel(const unsigned *__restrict__ bits, const int *__restrict__ grpRun,
                                          const int *__restrict__ nzrow, double *__restrict__ y, const double *__restrict__ carry,
                                          int E, int tileLo, int tileHi, int rowLo, int rowHi)
{
    const int t = tileLo + 1 + blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= tileHi) return;
    auto first_bit = [&](int u) { const long long e = (long long)u * E; return (bits[e >> 5] >> (unsigned)(e & 31)) & 1u; };
    auto run_at = [&](int u) { return grpRun[((long long)u * E) >> 7] - 1 + (int)first_bit(u); };   // run of the tile's first entry
    if (first_bit(t)) return;                               // a row begins exactly at the tile start
    const int run = run_at(t);
    if (!first_bit(t - 1) && run_at(t - 1) == run) return;  // the tile in front merely passes the row through: not the first piece
    double sum = 0.0;
    for (int u = t; u < tileHi && !first_bit(u) && run_at(u) == run; u++) sum = __dadd_rn(sum, carry[u]);
    const int r = nzrow ? nzrow[run] : run;
    if (r >= rowLo && r < rowHi) y[r] = __dadd_rn(y[r], sum);
}

int EntryStream::build(const int *ptr_d, const int *idx_d, const double *val_d, int nRow_, int nnz_, bool gather_bound,
                       cudaStream_t s)
{
    ptr = ptr_d; idx = idx_d; val = val_d; nRow = nRow_; nnz = nnz_;
    ok = false;
    range_cache.clear();
    if (nnz <= 0 || nRow <= 0) return B200SPMV_OK;
    tma = !gather_bound;
    static const int env_e = getenv("B200SPMV_ES_E") ? atoi(getenv("B200SPMV_ES_E")) : 0;
    E = env_e == 1024 || env_e == 2048 ? env_e : 2048;        // c4 (TMA-fed) 659 against 620 GFLOP/s with 1024; c3 (load-fed) 418 against 411
    nTiles = ceil_div(nnz, E);
    const int nGroups = ceil_div((long long)nTiles * E, 128);
    B2_TRY(bits.alloc((size_t)nGroups * 4 + 4));
    B2_TRY(grpRun.alloc((size_t)nGroups + 1));
    B2_TRY(carry.alloc((size_t)nTiles));
    B2_CUDA(cudaMemsetAsync(bits.p, 0, bits.bytes(), s));
    DevBuf<int> cnt;
    B2_TRY(cnt.alloc(1));
    B2_CUDA(cudaMemsetAsync(cnt.p, 0, sizeof(int), s));
    es_empty_count_kernel<<<ceil_div(nRow, 256), 256, 0, s>>>(ptr, nRow, cnt.p);
    B2_KERNEL_CHECK();
    B2_CUDA(cudaMemcpyAsync(&nEmpty, cnt.p, sizeof(int), cudaMemcpyDeviceToHost, s));
    B2_CUDA(cudaStreamSynchronize(s));
    B2_TRY(empty.alloc((size_t)nEmpty));
    B2_CUDA(cudaMemsetAsync(cnt.p, 0, sizeof(int), s));
    es_bits_kernel<<<ceil_div(nRow, 256), 256, 0, s>>>(ptr, nRow, bits.p, nEmpty ? empty.p : nullptr, cnt.p);
    B2_KERNEL_CHECK();
    es_group_count_kernel<<<ceil_div(nGroups + 1, 256), 256, 0, s>>>(bits.p, nGroups, grpRun.p);
    B2_KERNEL_CHECK();
    B2_TRY(exclusive_scan_i32(grpRun.p, grpRun.p, nGroups + 1, s));
    if (nEmpty) {
        B2_TRY(nzrow.alloc((size_t)(nRow - nEmpty)));
        es_nzrow_kernel<<<ceil_div(nRow, 256), 256, 0, s>>>(ptr, nRow, bits.p, grpRun.p, nzrow.p);
        B2_KERNEL_CHECK();
    } else {
        nzrow.release();
    }
    B2_CUDA(cudaStreamSynchronize(s));
    ok = true;
    return B200SPMV_OK;
}

int EntryStream::prepare(int rb, int re)
{
    if (rb < 0 || re > nRow || rb > re) { set_error("prepare_rows: bad row range [%d,%d) for %d rows", rb, re, nRow); return B200SPMV_ERR_INVALID; }
    const auto key = std::make_pair(rb, re);
    if (range_cache.count(key)) return B200SPMV_OK;
    int pb = 0, pe = 0;
    B2_CUDA(cudaMemcpy(&pb, ptr + rb, sizeof(int), cudaMemcpyDeviceToHost));
    B2_CUDA(cudaMemcpy(&pe, ptr + re, sizeof(int), cudaMemcpyDeviceToHost));
    const int lo = pb / E, hi = pe > pb ? (pe - 1) / E + 1 : lo;           // tiles that hold an entry of the rows
    range_cache.emplace(key, std::make_pair(lo, hi));
    return B200SPMV_OK;
}

template <int E, bool TMA>
static int es_launch(EntryStream &T, const double *x, double *y, int tileLo, int tileHi, int rb, int re, cudaStream_t s)
{
    constexpr size_t smem = TMA ? 24 * (size_t)E : 0;
    auto kern = entry_stream_kernel<E, TMA>;
    static std::map<int, int> per_sm;
    static int sms = 0;
    static const int env_b = getenv("B200SPMV_ES_CTAS") ? atoi(getenv("B200SPMV_ES_CTAS")) : 0;
    int dev = 0;
    B2_CUDA(cudaGetDevice(&dev));
    if (!sms) B2_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    auto it = per_sm.find(dev);
    if (it == per_sm.end()) {
        if (smem) B2_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int n = 0;
        B2_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kern, ES_THREADS, smem));
        if (n < 1) { set_error("entry stream: %zu bytes of shared memory do not fit", smem); return B200SPMV_ERR_UNSUPPORTED; }
        it = per_sm.emplace(dev, n).first;
    }
    const int nT = tileHi - tileLo;
    const int perSm = env_b > 0 ? std::min(env_b, it->second) : std::min(it->second, 6);
    const int grid = TMA ? std::min(nT, sms * perSm) : nT;
    kern<<<grid, ES_THREADS, smem, s>>>(T.idx, T.val, T.bits.p, T.grpRun.p, T.nEmpty ? T.nzrow.p : nullptr, x, y, T.carry.p, T.nnz, tileLo,
                                        tileHi, rb, re);
    B2_KERNEL_CHECK();
    if (nT > 1) {
        entry_stream_fixup_kernel<<<ceil_div(nT - 1, 256), 256, 0, s>>>(T.bits.p, T.grpRun.p, T.nEmpty ? T.nzrow.p : nullptr, y, T.carry.p, E,
                                                                       tileLo, tileHi, rb, re);
        B2_KERNEL_CHECK();
    }
    return B200SPMV_OK;
}

int EntryStream::run_rows(const double *x, double *y, int rb, int re, cudaStream_t s)
{
    if (rb < 0 || re > nRow || rb > re) { set_error("multiply_rows: bad row range [%d,%d) for %d rows", rb, re, nRow); return B200SPMV_ERR_INVALID; }
    if (rb == re) return B200SPMV_OK;
    int lo = 0, hi = nTiles;
    if (!(rb == 0 && re == nRow)) {
        auto it = range_cache.find(std::make_pair(rb, re));
        if (it == range_cache.end()) {
            cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
            if (cudaStreamIsCapturing(s, &cap) == cudaSuccess && cap != cudaStreamCaptureStatusNone) {
                set_error("multiply_rows: row range [%d,%d) is used for the first time inside a stream capture; call b200spmv_prepare_rows() first", rb, re);
                return B200SPMV_ERR_STATE;
            }
            B2_TRY(prepare(rb, re));
            it = range_cache.find(std::make_pair(rb, re));
        }
        lo = it->second.first;
        hi = it->second.second;
    }
    if (nEmpty) {
        es_zero_rows_kernel<<<ceil_div(nEmpty, 256), 256, 0, s>>>(empty.p, nEmpty, rb, re, y);
        B2_KERNEL_CHECK();
    }
    if (hi <= lo) return B200SPMV_OK;
    if (tma) return E == 1024 ? es_launch<1024, true>(*this, x, y, lo, hi, rb, re, s) : es_launch<2048, true>(*this, x, y, lo, hi, rb, re, s);
    return E == 1024 ? es_launch<1024, false>(*this, x, y, lo, hi, rb, re, s) : es_launch<2048, false>(*this, x, y, lo, hi, rb, re, s);
}

}  // namespace b2
