// xwin.cu -- the x halo exchange of the one-process-per-GPU path as ONE kernel over NVLink peer memory.
// The reference has no distributed code; singlespmv_b200/dist.py is the host side.  Round 1 moved the halo with a pack
// kernel + grouped NCCL send/recv (two NCCL kernels, ~0.07 ms of the 0.34 ms step exposed at 8 GPUs).  Here every rank
// keeps its extended x vector [left halo | owned | right halo] in an "x window": a cudaMalloc'ed block exported to the
// other processes of the node through CUDA IPC, so that each rank maps every peer's window (NVLink / NVSwitch peer
// access) and one kernel per step does the whole exchange:
//     1. tell my readers that my owned slice is in place       (release store of the step number into THEIR flag words)
//     2. wait until the owners of my halo columns said the same (acquire loads of MY flag words)
//     3. pull exactly the halo entries out of the owners' slices (peer loads that bypass L1) into my halo slots
//     4. tell the owners I am done, and wait until my readers are (so that when the step ends x may be overwritten)
// It runs on the communication stream next to the interior rows; no pack kernel, no staging buffer, no NCCL kernels.
// The single-process twin is mg_pull_kernel (mg.cu), which needs no flags because one host thread orders all GPUs.
#include <algorithm>

#include "common.cuh"

using namespace b2;

namespace {

constexpr int XW_MAXPEERS = 64;
constexpr size_t XW_HEADER = 4096;                 // flag words in front of x_ext (same allocation = one IPC handle)
constexpr int XW_THREADS = 64;                     // small CTAs: they must fit into what the interior rows' persistent CTAs leave of an SM
constexpr long long XW_SPIN_LIMIT = 20000000000LL; // ~10 s of SM clocks: a rank that never shows up ends the kernel with an error flag

struct XwHeader {                                  // lives at the start of every window
    unsigned long long ready[XW_MAXPEERS];         // ready[p] = last step for which rank p's owned slice is in place
    unsigned long long done[XW_MAXPEERS];          // done[p]  = last step whose pull rank p has finished
    unsigned long long step;                       // steps finished by the owner of this window
    unsigned int ticket, timeout;                  // last-block election; set when a spin ran into XW_SPIN_LIMIT
};
static_assert(sizeof(XwHeader) <= XW_HEADER, "window header");

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ double ld_peer(const double *p)        // never served from this SM's L1
{
    double v;
    asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ bool spin_until(const unsigned long long *flag, unsigned long long s, unsigned int *timeout)
{
    const long long t0 = clock64();
    while (ld_acquire_sys(flag) < s) {
        if (clock64() - t0 > XW_SPIN_LIMIT) { atomicExch(timeout, 1u); return false; }
        __nanosleep(64);
    }
    return true;
}

struct XwArgs {
    XwHeader *mine;                                // this rank's header
    char *const *peer_base;                        // [world] mapped window of every rank (own entry = own window)
    const int *sources, *readers;                  // ranks this one pulls from / that pull from this one
    int nSources, nReaders, me;
    const double *const *src;                      // per halo entry: its address inside the owner's (mapped) slice
    int nHalo, nLeft, nLocal;
    double *x_ext;
};

__global__ void __launch_bounds__(XW_THREADS) xwin_exchange_kernel(XwArgs a)
{
    __shared__ int is_last;
    const int tid = threadIdx.x;
    const unsigned long long s = *(volatile unsigned long long *)&a.mine->step + 1ULL;
    // 1. my owned slice was written by earlier work of this stream (or of the stream this one waited on)
    if (blockIdx.x == 0 && tid < a.nReaders) {
        __threadfence_system();
        XwHeader *h = reinterpret_cast<XwHeader *>(a.peer_base[a.readers[tid]]);
        st_release_sys(&h->ready[a.me], s);
    }
    // 2. every block waits for the owners itself (no inter-block dependency)
    if (tid < a.nSources) spin_until(&a.mine->ready[a.sources[tid]], s, &a.mine->timeout);
    __syncthreads();
    // 3. pull: x_ext halo slot i <- owner's slice
    const int stride = gridDim.x * XW_THREADS;
    for (int i0 = blockIdx.x * XW_THREADS + tid; i0 < a.nHalo; i0 += 4 * stride) {
        double v[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int i = i0 + u * stride;
            if (i < a.nHalo) v[u] = ld_peer(a.src[i]);
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int i = i0 + u * stride;
            if (i < a.nHalo) a.x_ext[i < a.nLeft ? i : a.nLocal + i] = v[u];
        }
    }
    // 4. the last block to finish closes the step
    __syncthreads();
    if (tid == 0) {
        __threadfence();
        is_last = atomicAdd(&a.mine->ticket, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (!is_last) return;
    if (tid < a.nSources) {
        XwHeader *h = reinterpret_cast<XwHeader *>(a.peer_base[a.sources[tid]]);
        st_release_sys(&h->done[a.me], s);
    }
    if (tid < a.nReaders) spin_until(&a.mine->done[a.readers[tid]], s, &a.mine->timeout);
    __syncthreads();
    if (tid == 0) {
        a.mine->ticket = 0u;
        *(volatile unsigned long long *)&a.mine->step = s;
        __threadfence();
    }
}

}  // namespace

struct b200spmv_xwin {
    int rank = 0, world = 0, dev = 0;
    long long nExt = 0, ownedOff = 0;              // doubles in x_ext; offset (doubles) of the owned slice
    char *base = nullptr;                          // header + x_ext
    std::vector<char *> peer;                      // mapped peer windows (own entry = base)
    std::vector<long long> peerOwnedOff;           // bytes, from the peer's window base
    std::vector<bool> opened;
    DevBuf<char *> peer_d;
    DevBuf<const double *> src_d;
    DevBuf<int> sources_d, readers_d;
    int sms = 0;
    int nSources = 0, nReaders = 0, nHalo = 0, nLeft = 0, nLocal = 0;
    bool planned = false;
};

extern "C" {

int b200spmv_xwin_create(int rank, int world, long long nExt, long long ownedOff, b200spmv_xwin **out)
{
    clear_error();
    if (!out || world < 1 || world > XW_MAXPEERS || rank < 0 || rank >= world || nExt < 0 || ownedOff < 0 || ownedOff > nExt) {
        set_error("xwin_create: bad argument (at most %d ranks)", XW_MAXPEERS);
        return B200SPMV_ERR_INVALID;
    }
    *out = nullptr;
    std::unique_ptr<b200spmv_xwin> w(new b200spmv_xwin());
    w->rank = rank; w->world = world; w->nExt = nExt; w->ownedOff = ownedOff;
    B2_CUDA(cudaGetDevice(&w->dev));
    const size_t bytes = XW_HEADER + sizeof(double) * (size_t)std::max<long long>(nExt, 1);
    B2_CUDA(cudaMalloc((void **)&w->base, bytes));
    B2_CUDA(cudaMemset(w->base, 0, bytes));
    B2_CUDA(cudaDeviceSynchronize());
    w->peer.assign((size_t)world, nullptr);
    w->peerOwnedOff.assign((size_t)world, 0);
    w->opened.assign((size_t)world, false);
    w->peer[(size_t)rank] = w->base;
    w->peerOwnedOff[(size_t)rank] = (long long)XW_HEADER + 8 * ownedOff;
    *out = w.release();
    return B200SPMV_OK;
}

void *b200spmv_xwin_x_ext(b200spmv_xwin *w) { return w ? (void *)(w->base + XW_HEADER) : nullptr; }

int b200spmv_xwin_export(b200spmv_xwin *w, void *blob80)
{
    clear_error();
    if (!w || !blob80) { set_error("xwin_export: NULL argument"); return B200SPMV_ERR_INVALID; }
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    cudaIpcMemHandle_t h;
    B2_CUDA(cudaIpcGetMemHandle(&h, w->base));
    char *b = static_cast<char *>(blob80);
    memcpy(b, &h, 64);
    const long long off = w->peerOwnedOff[(size_t)w->rank];
    const long long dev = w->dev;
    memcpy(b + 64, &off, 8);
    memcpy(b + 72, &dev, 8);
    return B200SPMV_OK;
}

int b200spmv_xwin_import(b200spmv_xwin *w, int peer, const void *blob80)
{
    clear_error();
    if (!w || !blob80 || peer < 0 || peer >= w->world) { set_error("xwin_import: bad argument"); return B200SPMV_ERR_INVALID; }
    if (peer == w->rank) return B200SPMV_OK;
    const char *b = static_cast<const char *>(blob80);
    cudaIpcMemHandle_t h;
    memcpy(&h, b, 64);
    long long off = 0;
    memcpy(&off, b + 64, 8);
    void *p = nullptr;
    B2_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    w->peer[(size_t)peer] = static_cast<char *>(p);
    w->peerOwnedOff[(size_t)peer] = off;
    w->opened[(size_t)peer] = true;
    return B200SPMV_OK;
}

// same-process twin of export + import (tests, several blocks driven by one process): map `other` as rank `peer`
int b200spmv_xwin_attach(b200spmv_xwin *w, int peer, b200spmv_xwin *other)
{
    clear_error();
    if (!w || !other || peer < 0 || peer >= w->world || peer == w->rank || other->rank != peer) { set_error("xwin_attach: bad argument"); return B200SPMV_ERR_INVALID; }
    if (other->dev != w->dev) {
        int can = 0;
        B2_CUDA(cudaDeviceCanAccessPeer(&can, w->dev, other->dev));
        if (!can) { set_error("xwin_attach: GPU %d cannot map GPU %d's memory", w->dev, other->dev); return B200SPMV_ERR_UNSUPPORTED; }
        int cur = 0;
        B2_CUDA(cudaGetDevice(&cur));
        B2_CUDA(cudaSetDevice(w->dev));
        const cudaError_t e = cudaDeviceEnablePeerAccess(other->dev, 0);
        cudaSetDevice(cur);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { set_error("cudaDeviceEnablePeerAccess: %s", cudaGetErrorString(e)); return B200SPMV_ERR_CUDA; }
        cudaGetLastError();
    }
    w->peer[(size_t)peer] = other->base;
    w->peerOwnedOff[(size_t)peer] = other->peerOwnedOff[(size_t)other->rank];
    return B200SPMV_OK;
}

// halo_cols: the block's halo columns (global ids, ascending; b200spmv_halo_cols), bounds: the row/column split of all ranks,
// readers: the ranks that asked this one for columns (dist.plan_requests)
int b200spmv_xwin_plan(b200spmv_xwin *w, const int *halo_cols_h, int nHalo, int nLeft, int nLocal, const long long *bounds_h,
                       const int *readers_h, int nReaders)
{
    clear_error();
    if (!w || nHalo < 0 || (nHalo && !halo_cols_h) || !bounds_h || nReaders < 0 || (nReaders && !readers_h)) { set_error("xwin_plan: bad argument"); return B200SPMV_ERR_INVALID; }
    if ((long long)nLeft + nLocal + (nHalo - nLeft) != w->nExt || w->ownedOff != nLeft) { set_error("xwin_plan: the halo does not match the window layout"); return B200SPMV_ERR_INVALID; }
    std::vector<const double *> src((size_t)nHalo);
    std::vector<int> sources;
    for (int i = 0; i < nHalo; i++) {
        const long long c = halo_cols_h[i];
        const int p = (int)(std::upper_bound(bounds_h, bounds_h + w->world + 1, c) - bounds_h) - 1;
        if (p < 0 || p >= w->world || p == w->rank) { set_error("xwin_plan: halo column %lld has no remote owner", c); return B200SPMV_ERR_INVALID; }
        if (!w->peer[(size_t)p]) { set_error("xwin_plan: the window of rank %d was not imported", p); return B200SPMV_ERR_STATE; }
        src[(size_t)i] = reinterpret_cast<const double *>(w->peer[(size_t)p] + w->peerOwnedOff[(size_t)p]) + (c - bounds_h[p]);
        if (sources.empty() || sources.back() != p) {
            if (std::find(sources.begin(), sources.end(), p) == sources.end()) sources.push_back(p);
        }
    }
    for (int i = 0; i < nReaders; i++)
        if (readers_h[i] < 0 || readers_h[i] >= w->world || readers_h[i] == w->rank || !w->peer[(size_t)readers_h[i]]) { set_error("xwin_plan: bad reader rank"); return B200SPMV_ERR_INVALID; }
    if ((int)sources.size() > XW_THREADS || nReaders > XW_THREADS) { set_error("xwin_plan: too many peers"); return B200SPMV_ERR_UNSUPPORTED; }
    w->nHalo = nHalo; w->nLeft = nLeft; w->nLocal = nLocal;
    w->nSources = (int)sources.size(); w->nReaders = nReaders;
    B2_TRY(w->peer_d.alloc((size_t)w->world));
    B2_TRY(w->sources_d.alloc(sources.size()));
    B2_TRY(w->readers_d.alloc((size_t)nReaders));
    B2_TRY(w->src_d.alloc((size_t)nHalo));
    B2_CUDA(cudaDeviceGetAttribute(&w->sms, cudaDevAttrMultiProcessorCount, w->dev));
    B2_CUDA(cudaMemcpy(w->peer_d.p, w->peer.data(), sizeof(char *) * (size_t)w->world, cudaMemcpyHostToDevice));
    if (!sources.empty()) B2_CUDA(cudaMemcpy(w->sources_d.p, sources.data(), sizeof(int) * sources.size(), cudaMemcpyHostToDevice));
    if (nReaders) B2_CUDA(cudaMemcpy(w->readers_d.p, readers_h, sizeof(int) * (size_t)nReaders, cudaMemcpyHostToDevice));
    if (nHalo) B2_CUDA(cudaMemcpy(w->src_d.p, src.data(), sizeof(const double *) * (size_t)nHalo, cudaMemcpyHostToDevice));
    w->planned = true;
    return B200SPMV_OK;
}

// one step's exchange, asynchronous on `stream` (capturable).  Every rank must call it once per step.
int b200spmv_xwin_exchange(b200spmv_xwin *w, void *stream)
{
    if (!w || !w->planned) { set_error("xwin_exchange: window not planned"); return B200SPMV_ERR_STATE; }
    if (w->nSources == 0 && w->nReaders == 0) return B200SPMV_OK;
    XwArgs a;
    a.mine = reinterpret_cast<XwHeader *>(w->base);
    a.peer_base = w->peer_d.p;
    a.sources = w->sources_d.p; a.readers = w->readers_d.p;
    a.nSources = w->nSources; a.nReaders = w->nReaders; a.me = w->rank;
    a.src = w->src_d.p;
    a.nHalo = w->nHalo; a.nLeft = w->nLeft; a.nLocal = w->nLocal;
    a.x_ext = reinterpret_cast<double *>(w->base + XW_HEADER);
    // The interior rows run as persistent CTAs that fill every SM up to a few thousand registers (chunk_stream: 5 x 256
    // threads x 48 registers of 65536): CTAs of 64 threads x 32 registers still fit beside them, two per SM, whichever
    // kernel the hardware (or a graph replay) starts first.  Measured with 256-thread CTAs: the exchange only ran after the
    // interior rows had drained whenever it was launched second (2 GPUs, graph replay: 1.26 ms against 1.05).
    const int grid = std::max(1, std::min(2 * w->sms, ceil_div(w->nHalo, 4 * XW_THREADS)));
    xwin_exchange_kernel<<<grid, XW_THREADS, 0, (cudaStream_t)stream>>>(a);
    B2_KERNEL_CHECK();
    return B200SPMV_OK;
}

// steps finished; *timed_out != 0 if a flag wait ever gave up (a peer that did not take part in a step)
int b200spmv_xwin_status(b200spmv_xwin *w, long long *steps, int *timed_out)
{
    if (!w) { set_error("xwin_status: NULL handle"); return B200SPMV_ERR_INVALID; }
    XwHeader h;
    B2_CUDA(cudaMemcpy(&h, w->base, sizeof h, cudaMemcpyDeviceToHost));
    if (steps) *steps = (long long)h.step;
    if (timed_out) *timed_out = (int)h.timeout;
    return B200SPMV_OK;
}

int b200spmv_xwin_free(b200spmv_xwin *w)
{
    if (!w) return B200SPMV_OK;
    for (int p = 0; p < w->world; p++)
        if (w->opened[(size_t)p] && w->peer[(size_t)p]) cudaIpcCloseMemHandle(w->peer[(size_t)p]);
    if (w->base) cudaFree(w->base);
    cudaGetLastError();
    delete w;
    return B200SPMV_OK;
}

}  // extern "C"
