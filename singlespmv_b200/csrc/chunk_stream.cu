// chunk_stream.cu -- kernel and launcher of the TMA-fed row-chunk stream (see chunk_stream.cuh).
#include <algorithm>
#include <cstdlib>
#include <map>

#include "chunk_stream.cuh"

namespace b2 {

// VT = stored value type, XT = type of x and y, AT = type the row sums are accumulated in
template <typename VT, typename XT, typename AT, int MAXL, int TH, bool ACC>
__global__ void __launch_bounds__(TH)
chunk_stream_kernel(const int *__restrict__ ptr, const int *__restrict__ idx, const VT *__restrict__ val,
                    const XT *__restrict__ x, XT *__restrict__ y, int rowBegin, int rowEnd, int chunk0,
                    int nChunks, int capI, int capV, int S, int acc_mode)
{
    extern __shared__ __align__(128) unsigned char cs_smem[];
    __shared__ __align__(8) uint64_t bar[CS_MAXSTAGES];
    __shared__ int sbaseI[CS_MAXSTAGES], sbaseV[CS_MAXSTAGES];    // entry index that sits at slot 0 of the stage
    int *sidx = reinterpret_cast<int *>(cs_smem);
    VT *sval = reinterpret_cast<VT *>(cs_smem + (size_t)S * capI * sizeof(int));
    const int tid = threadIdx.x;
    const uint64_t pol_stream = policy_evict_first(), pol_x = policy_evict_last();

    // thread 0 only: bulk copies of entries [b, e) into stage s; sources rounded down / sizes rounded up to 16 bytes
    auto issue = [&](int s, int b, int e) {
        const int *srcI = idx + b;
        const VT *srcV = val + b;
        const int misI = (int)((reinterpret_cast<uintptr_t>(srcI) & 15) / sizeof(int));
        const int misV = (int)((reinterpret_cast<uintptr_t>(srcV) & 15) / sizeof(VT));
        const int n = e - b;
        const uint32_t bytesI = n ? (uint32_t)(((n + misI) * (int)sizeof(int) + 15) & ~15) : 0u;
        const uint32_t bytesV = n ? (uint32_t)(((n + misV) * (int)sizeof(VT) + 15) & ~15) : 0u;
        sbaseI[s] = b - misI;
        sbaseV[s] = b - misV;
        mbar_expect_tx(&bar[s], bytesI + bytesV);               // release: sbase*[s] visible to whoever passes the wait
        if (n) {
            tma_load_1d(sidx + (size_t)s * capI, srcI - misI, bytesI, &bar[s], pol_stream);
            tma_load_1d(sval + (size_t)s * capV, srcV - misV, bytesV, &bar[s], pol_stream);
        }
    };
    // chunk g covers rows [g TH, (g+1) TH) clipped to [rowBegin, rowEnd)
    auto lo_of = [&](long long g) { return (int)min((long long)rowEnd, max((long long)rowBegin, g * TH)); };

    if (tid == 0)
        for (int s = 0; s < S; s++) mbar_init(&bar[s], 1);
    __syncthreads();
    if (tid == 0) {
        for (int s = 0; s < S; s++) {
            const long long c = (long long)blockIdx.x + (long long)s * gridDim.x;
            if (c < nChunks) issue(s, ptr[lo_of(chunk0 + c)], ptr[lo_of(chunk0 + c + 1)]);
        }
    }
    long long g = (long long)chunk0 + blockIdx.x;
    int r = (int)min(g * TH + tid, (long long)0x7fffffff);
    bool valid = r >= rowBegin && r < rowEnd;
    int p = valid ? ptr[r] : 0, q = valid ? ptr[r + 1] : 0;
    int k = 0;
    for (long long c = blockIdx.x; c < nChunks; c += gridDim.x, k++) {
        const int s = k % S;
        // prefetches that land while this chunk is being reduced: the bounds of the chunk issued at the end of this
        // iteration (thread 0), this thread's row pointers in the CTA's next chunk, and y for the accumulating modes
        const long long cIssue = c + (long long)S * gridDim.x, cNext = c + gridDim.x;
        int nb = 0, ne = 0, pn = 0, qn = 0;
        if (tid == 0 && cIssue < nChunks) {
            nb = ptr[lo_of(chunk0 + cIssue)];
            ne = ptr[lo_of(chunk0 + cIssue + 1)];
        }
        const long long rnl = (chunk0 + cNext) * TH + tid;
        const int rn = (int)min(rnl, (long long)0x7fffffff);
        const bool validn = cNext < nChunks && rn >= rowBegin && rn < rowEnd;
        if (validn) {
            pn = ptr[rn];
            qn = ptr[rn + 1];
        }
        const AT y0 = (ACC && valid) ? (AT)y[r] : (AT)0;
        mbar_wait(&bar[s], (uint32_t)(k / S) & 1u);
        const int io = s * capI + (p - sbaseI[s]), vo = s * capV + (p - sbaseV[s]);
        const int len = q - p;
        AT acc = (ACC && acc_mode == CS_CONTINUE) ? y0 : (AT)0;
        XT xs[MAXL];
#pragma unroll
        for (int j = 0; j < MAXL; j++)
            if (j < len) xs[j] = ld_x(x + sidx[io + j], pol_x);
#pragma unroll
        for (int j = 0; j < MAXL; j++)
            if (j < len) acc = Arith<AT>::add(acc, Arith<AT>::mul((AT)sval[vo + j], (AT)xs[j]));
        if (valid) y[r] = (XT)((ACC && acc_mode == CS_ADD) ? Arith<AT>::add(y0, acc) : acc);
        __syncthreads();                                       // every thread is done with stage s
        if (tid == 0 && cIssue < nChunks) issue(s, nb, ne);
        r = rn; p = pn; q = qn; valid = validn;
    }
}

__global__ void chunk_cap_kernel(const int *__restrict__ ptr, int nRow, int th, int nChunks, int *__restrict__ out)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    int n = 0;
    if (c < nChunks) n = ptr[(int)min((long long)nRow, ((long long)c + 1) * th)] - ptr[(long long)c * th];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) n = max(n, __shfl_xor_sync(0xffffffffu, n, o));
    if ((threadIdx.x & 31) == 0 && n > 0) atomicMax(out, n);
}

int ChunkStream::build(const int *ptr_d, const int *idx_d, const void *val_d, bool val_is_f32, int nRow_, int nnz,
                       int maxLen_, cudaStream_t s)
{
    ptr = ptr_d; idx = idx_d; val = val_d; f32 = val_is_f32; nRow = nRow_; maxLen = maxLen_;
    static const int env_max = getenv("B200SPMV_TMA_MAXLEN") ? atoi(getenv("B200SPMV_TMA_MAXLEN")) : CS_MAXLEN;
    static const int env_th = getenv("B200SPMV_TMA_R") ? atoi(getenv("B200SPMV_TMA_R")) : 0;
    ok = nnz > 0 && nRow > 0 && maxLen > 0 && maxLen <= std::min(env_max, CS_MAXLEN);
    if (!ok) return B200SPMV_OK;
    // rows per chunk = threads per CTA (c5: 256 and 512 within 1.5 % of each other, 128: -15 %)
    th = env_th == 512 ? 512 : 256;
    const int nChunks = ceil_div(nRow, th);
    DevBuf<int> m;
    B2_TRY(m.alloc(1));
    B2_CUDA(cudaMemsetAsync(m.p, 0, sizeof(int), s));
    chunk_cap_kernel<<<ceil_div(nChunks, 256), 256, 0, s>>>(ptr, nRow, th, nChunks, m.p);
    B2_KERNEL_CHECK();
    B2_CUDA(cudaMemcpyAsync(&cap, m.p, sizeof(int), cudaMemcpyDeviceToHost, s));
    B2_CUDA(cudaStreamSynchronize(s));
    // a stage must fit twice into one SM's shared memory next to nothing else
    const size_t stage = (size_t)(cap + 16) * (4 + (f32 ? 4 : 8));
    if (2 * stage > 200 * 1024) ok = false;
    return B200SPMV_OK;
}

template <typename VT, typename XT, typename AT, int MAXL, int TH, bool ACC>
static int cs_launch(const ChunkStream &c, const XT *x, XT *y, int rb, int re, int acc, cudaStream_t s)
{
    constexpr int VA = 16 / (int)sizeof(VT);
    static const int env_s = getenv("B200SPMV_TMA_S") ? atoi(getenv("B200SPMV_TMA_S")) : 0;
    static const int env_b = getenv("B200SPMV_TMA_CTAS") ? atoi(getenv("B200SPMV_TMA_CTAS")) : 0;
    const int S = (env_s >= 1 && env_s <= CS_MAXSTAGES) ? env_s : 2;
    const int capI = (c.cap + 8) & ~3, capV = (c.cap + 2 * VA) & ~(VA - 1);
    const size_t smem = (size_t)S * ((size_t)capI * 4 + (size_t)capV * sizeof(VT));
    auto kern = chunk_stream_kernel<VT, XT, AT, MAXL, TH, ACC>;
    static int sms = 0;
    // occupancy of THIS instantiation by (device, shared-memory size): the opt-in shared-memory attribute is per device
    static std::map<std::pair<int, size_t>, int> per_sm;
    int dev = 0;
    B2_CUDA(cudaGetDevice(&dev));
    if (!sms) B2_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const auto key = std::make_pair(dev, smem);
    auto it = per_sm.find(key);
    if (it == per_sm.end()) {
        B2_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max<size_t>(smem, 200 * 1024)));
        int n = 0;
        B2_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kern, TH, smem));
        if (n < 1) { set_error("row-chunk stream: %zu bytes of shared memory do not fit", smem); return B200SPMV_ERR_UNSUPPORTED; }
        it = per_sm.emplace(key, n).first;
    }
    const int perSm = env_b > 0 ? std::min(it->second, env_b) : it->second;
    const int chunk0 = rb / TH, nChunks = ceil_div(re, TH) - chunk0;
    const int grid = std::min(nChunks, sms * perSm);
    kern<<<grid, TH, smem, s>>>(c.ptr, c.idx, static_cast<const VT *>(c.val), x, y, rb, re, chunk0, nChunks, capI, capV, S, acc);
    B2_KERNEL_CHECK();
    return B200SPMV_OK;
}

template <typename VT, typename XT, typename AT>
static int cs_dispatch(const ChunkStream &c, const XT *x, XT *y, int rb, int re, int acc, cudaStream_t s)
{
#define CS_GO(MAXL)                                                                                       \
    do {                                                                                                  \
        if (c.th == 512 && acc) return cs_launch<VT, XT, AT, MAXL, 512, true>(c, x, y, rb, re, acc, s);   \
        if (c.th == 512) return cs_launch<VT, XT, AT, MAXL, 512, false>(c, x, y, rb, re, acc, s);         \
        if (acc) return cs_launch<VT, XT, AT, MAXL, 256, true>(c, x, y, rb, re, acc, s);                  \
        return cs_launch<VT, XT, AT, MAXL, 256, false>(c, x, y, rb, re, acc, s);                          \
    } while (0)
    if (c.maxLen <= 8) CS_GO(8);
    else CS_GO(16);
#undef CS_GO
}

int ChunkStream::run(const double *x, double *y, int rb, int re, int acc, cudaStream_t s) const
{
    if (rb < 0 || re > nRow || rb > re) { set_error("multiply_rows: bad row range [%d,%d) for %d rows", rb, re, nRow); return B200SPMV_ERR_INVALID; }
    if (rb == re) return B200SPMV_OK;
    if (f32) return cs_dispatch<float, double, double>(*this, x, y, rb, re, acc, s);
    return cs_dispatch<double, double, double>(*this, x, y, rb, re, acc, s);
}

// fp32 vectors (values must be stored as fp32): accumulation in fp32 or fp64
int ChunkStream::run_f32(const float *x, float *y, int rb, int re, int acc, bool acc64, cudaStream_t s) const
{
    if (rb < 0 || re > nRow || rb > re) { set_error("multiply_rows: bad row range [%d,%d) for %d rows", rb, re, nRow); return B200SPMV_ERR_INVALID; }
    if (!f32) { set_error("row-chunk stream: fp32 vectors need fp32 value storage"); return B200SPMV_ERR_STATE; }
    if (rb == re) return B200SPMV_OK;
    if (acc64) return cs_dispatch<float, float, double>(*this, x, y, rb, re, acc, s);
    return cs_dispatch<float, float, float>(*this, x, y, rb, re, acc, s);
}

}  // namespace b2
