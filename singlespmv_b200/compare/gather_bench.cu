// gather_bench.cu -- MEASUREMENT TOOL ONLY (libb200cmp.so), never on the product path.
// What is the ceiling of random 8-byte gathers on this GPU?  Every x[col] of a uniform random matrix (BASELINE
// config 2) is one 32-byte L2 sector request for 8 useful bytes; when x does not fit in L2 it is also a 64-byte
// DRAM burst.  The kernel below issues nothing but such gathers (8 independent ones in flight per thread, indices
// from a counter hash: no index stream) into a table of a chosen size, so the number it prints is the rate the
// memory system sustains for the access pattern itself: table <= L2 -> L2 sector ceiling, table >> L2 -> DRAM
// random-access ceiling.  profiles/r2_gather_ceiling.md records the results the c2 kernels are held against.
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

static char g_gerr[256];
#define GB_CUDA(e) do { cudaError_t e_ = (e); if (e_ != cudaSuccess) { snprintf(g_gerr, sizeof g_gerr, "%s: %s", #e, cudaGetErrorString(e_)); return -2; } } while (0)

__device__ __forceinline__ uint64_t mix64(uint64_t z)
{
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

template <int U>
__global__ void __launch_bounds__(256) gather_kernel(const double *__restrict__ table, uint32_t n, int rounds, double *__restrict__ out)
{
    const uint64_t gid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    double acc = 0.0;
    for (int r = 0; r < rounds; r++) {
        double v[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            const uint64_t h = mix64(gid * 0x100000001B3ull + (uint64_t)(r * U + u));
            v[u] = __ldg(table + (uint32_t)(((h >> 32) * (uint64_t)n) >> 32));
        }
#pragma unroll
        for (int u = 0; u < U; u++) acc += v[u];
    }
    if (acc == 123.456) out[gid & 1023] = acc;          // keeps the loads alive
}

// The same gathers next to a coalesced 12 B/gather stream read through the threads' own 128-bit loads (what a CRS / ELL kernel
// does with idx + val): does the stream share the limit of the gathers (sectors on the SM's miss path) or ride along (the
// limit counts L1 wavefronts, and a coalesced 128-bit load of a warp is only 4 of them for 128 entries)?
__global__ void __launch_bounds__(256) gather_stream_kernel(const double *__restrict__ table, uint32_t n, int rounds,
                                                            const int4 *__restrict__ sidx, const double2 *__restrict__ sval,
                                                            double *__restrict__ out)
{
    const uint64_t gid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t total = (uint64_t)gridDim.x * blockDim.x;
    double acc = 0.0;
    for (int r = 0; r < rounds; r++) {
        const uint64_t e = (uint64_t)r * total + gid;                     // this thread's 4 entries of the round
        int4 c;
        double2 v0, v1;
        asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(c.x), "=r"(c.y), "=r"(c.z), "=r"(c.w) : "l"(sidx + e));
        asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0,%1}, [%2];" : "=d"(v0.x), "=d"(v0.y) : "l"(sval + 2 * e));
        asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0,%1}, [%2];" : "=d"(v1.x), "=d"(v1.y) : "l"(sval + 2 * e + 1));
        double x[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const uint64_t h = mix64(gid * 0x100000001B3ull + (uint64_t)(r * 4 + u));
            x[u] = __ldg(table + (uint32_t)(((h >> 32) * (uint64_t)n) >> 32));
        }
        acc += v0.x * x[0] + v0.y * x[1] + v1.x * x[2] + v1.y * x[3] + (double)(c.x ^ c.y ^ c.z ^ c.w);
    }
    if (acc == 123.456) out[gid & 1023] = acc;
}

extern "C" __attribute__((visibility("default"))) const char *b200cmp_gather_last_error(void) { return g_gerr; }

// gathers + the 12 B/gather stream; returns million gathers per launch, mean ms in *ms_out
extern "C" __attribute__((visibility("default")))
int b200cmp_gather_stream(long long table_bytes, long long gathers, int warmup, int iters, float *ms_out)
{
    const uint32_t n = (uint32_t)(table_bytes / 8);
    const int threads = 256, blocks = 148 * 8 * 4;
    int rounds = (int)(gathers / ((long long)threads * blocks * 4));
    if (rounds < 1) rounds = 1;
    const size_t entries = (size_t)rounds * threads * blocks * 4;
    double *table = nullptr, *out = nullptr;
    int4 *sidx = nullptr;
    double2 *sval = nullptr;
    GB_CUDA(cudaMalloc((void **)&table, (size_t)n * 8));
    GB_CUDA(cudaMalloc((void **)&out, 1024 * 8));
    GB_CUDA(cudaMalloc((void **)&sidx, entries * 4));
    GB_CUDA(cudaMalloc((void **)&sval, entries * 8));
    GB_CUDA(cudaMemset(table, 0, (size_t)n * 8));
    GB_CUDA(cudaMemset(sidx, 0, entries * 4));
    GB_CUDA(cudaMemset(sval, 0, entries * 8));
    cudaEvent_t a, b;
    GB_CUDA(cudaEventCreate(&a));
    GB_CUDA(cudaEventCreate(&b));
    for (int i = 0; i < warmup; i++) gather_stream_kernel<<<blocks, threads>>>(table, n, rounds, sidx, sval, out);
    GB_CUDA(cudaEventRecord(a));
    for (int i = 0; i < iters; i++) gather_stream_kernel<<<blocks, threads>>>(table, n, rounds, sidx, sval, out);
    GB_CUDA(cudaEventRecord(b));
    GB_CUDA(cudaEventSynchronize(b));
    GB_CUDA(cudaGetLastError());
    float ms = 0;
    GB_CUDA(cudaEventElapsedTime(&ms, a, b));
    *ms_out = ms / iters;
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    cudaFree(table); cudaFree(out); cudaFree(sidx); cudaFree(sval);
    return (int)(entries / 1000000);
}

// table_bytes: size of the gathered table; gathers: total number of 8-byte loads; returns mean ms of `iters` launches
extern "C" __attribute__((visibility("default")))
int b200cmp_gather(long long table_bytes, long long gathers, int warmup, int iters, float *ms_out)
{
    const uint32_t n = (uint32_t)(table_bytes / 8);
    double *table = nullptr, *out = nullptr;
    GB_CUDA(cudaMalloc((void **)&table, (size_t)n * 8));
    GB_CUDA(cudaMalloc((void **)&out, 1024 * 8));
    GB_CUDA(cudaMemset(table, 0, (size_t)n * 8));
    constexpr int U = 8;
    const int threads = 256, blocks = 148 * 8 * 4;
    int rounds = (int)(gathers / ((long long)threads * blocks * U));
    if (rounds < 1) rounds = 1;
    cudaEvent_t a, b;
    GB_CUDA(cudaEventCreate(&a));
    GB_CUDA(cudaEventCreate(&b));
    for (int i = 0; i < warmup; i++) gather_kernel<U><<<blocks, threads>>>(table, n, rounds, out);
    GB_CUDA(cudaEventRecord(a));
    for (int i = 0; i < iters; i++) gather_kernel<U><<<blocks, threads>>>(table, n, rounds, out);
    GB_CUDA(cudaEventRecord(b));
    GB_CUDA(cudaEventSynchronize(b));
    GB_CUDA(cudaGetLastError());
    float ms = 0;
    GB_CUDA(cudaEventElapsedTime(&ms, a, b));
    *ms_out = ms / iters;
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    cudaFree(table);
    cudaFree(out);
    return (int)((long long)rounds * threads * blocks * U / 1000000);   // million gathers actually issued per launch
}
