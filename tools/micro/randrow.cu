// Micro-benchmark: random 128-byte row reads (8 lanes x 16 B, like the source-batched SSSP kernel) over working
// sets of growing size.  Tells TLB / page-walk limits apart from DRAM limits.  Build: nvcc -arch=sm_100a -O3.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__global__ void randrow(const ulonglong2 *buf, uint64_t rows, int iters, uint64_t seed, unsigned long long *sink, int rows_per_cta_window)
{
    const uint32_t lane8 = threadIdx.x & 7u;
    uint64_t x = seed + (uint64_t)(blockIdx.x * blockDim.x + threadIdx.x) / 8 * 0x9E3779B97F4A7C15ull;
    unsigned long long acc = 0;
    // window: each CTA confines itself to a contiguous window of the buffer (0 = whole buffer)
    uint64_t base = 0, span = rows;
    if (rows_per_cta_window > 0) { span = (uint64_t)rows_per_cta_window; base = ((uint64_t)blockIdx.x * span) % (rows - span + 1); }
    for (int i = 0; i < iters; ++i) {
        x ^= x << 13; x ^= x >> 7; x ^= x << 17;
        uint64_t r0 = base + (x % span), r1 = base + ((x >> 20) % span), r2 = base + ((x >> 9) % span), r3 = base + ((x >> 29) % span);
        ulonglong2 a = __ldcg(buf + r0 * 8 + lane8), b = __ldcg(buf + r1 * 8 + lane8), c = __ldcg(buf + r2 * 8 + lane8), d = __ldcg(buf + r3 * 8 + lane8);
        acc += a.x + b.y + c.x + d.y;
    }
    if (acc == 0x1234567ull) *sink = acc;
}
int main()
{
    size_t maxb = (size_t)96 << 30;
    ulonglong2 *buf; unsigned long long *sink;
    if (cudaMalloc(&buf, maxb) != cudaSuccess) { printf("alloc failed\n"); return 1; }
    cudaMalloc(&sink, 8);
    cudaMemset(buf, 1, maxb);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 2000;
    for (int window = 0; window < 3; ++window)
    for (double gb : {0.125, 0.5, 2.0, 8.0, 32.0, 96.0}) {
        uint64_t rows = (uint64_t)(gb * (1ull << 30)) / 128;
        int win_rows = window == 0 ? 0 : window == 1 ? (int)((128ull << 20) / 128) : (int)((16ull << 20) / 128);   // 128 MB or 16 MB per CTA
        if (win_rows && (uint64_t)win_rows > rows) continue;
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(e0);
            randrow<<<148, 1024>>>(buf, rows, iters, 12345 + rep, sink, win_rows);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
        }
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double rows_read = 148.0 * 1024 / 8 * iters * 4;
        printf("window=%s ws=%.3f GB: %.3f ms, %.2f G rows/s, %.1f GB/s\n", window == 0 ? "all" : window == 1 ? "128MB/CTA" : "16MB/CTA", gb, ms, rows_read / ms / 1e6, rows_read * 128 / ms / 1e6);
    }
    return 0;
}
