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

extern "C" __attribute__((visibility("default"))) const char *b200cmp_gather_last_error(void) { return g_gerr; }

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
