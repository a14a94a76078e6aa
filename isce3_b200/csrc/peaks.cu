// Roofline denominators measured on the device the bench runs on: sustained FP32 FMA,
// FP64 FMA and SFU (MUFU sin/cos) throughput plus the SM clock seen by the kernels
// (/root/repo/MEASURED_PEAKS.json carries HBM and bf16-GEMM peaks only; SURVEY.md 8d asks
// for an FMA / MUFU microbenchmark on the same box).
#include <cuda_runtime.h>

#include <algorithm>
#include <cstring>

#include "launch.h"

namespace i3b {

constexpr int PEAK_ITERS = 4096;
constexpr int PEAK_CHAINS = 8;

__global__ void __launch_bounds__(256) peak_ffma_kernel(float* out, float a, float b)
{
    float x[PEAK_CHAINS];
#pragma unroll
    for (int i = 0; i < PEAK_CHAINS; ++i) x[i] = threadIdx.x * 1e-3f + i;
    for (int it = 0; it < PEAK_ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < PEAK_CHAINS; ++i) x[i] = fmaf(x[i], a, b);
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < PEAK_CHAINS; ++i) s += x[i];
    if (s == 12345.678f) out[0] = s;
}

// packed FP32x2 FMA (Blackwell fma.rn.f32x2): same flops per lane-cycle, half the issue slots
__global__ void __launch_bounds__(256) peak_ffma2_kernel(float* out, float a, float b)
{
    unsigned long long x[PEAK_CHAINS];
    unsigned long long av, bv;
    asm("mov.b64 %0, {%1, %1};" : "=l"(av) : "f"(a));
    asm("mov.b64 %0, {%1, %1};" : "=l"(bv) : "f"(b));
#pragma unroll
    for (int i = 0; i < PEAK_CHAINS; ++i) {
        const float v = threadIdx.x * 1e-3f + i;
        asm("mov.b64 %0, {%1, %1};" : "=l"(x[i]) : "f"(v));
    }
    for (int it = 0; it < PEAK_ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < PEAK_CHAINS; ++i)
            asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(x[i]) : "l"(av), "l"(bv));
    }
    unsigned long long s = 0;
#pragma unroll
    for (int i = 0; i < PEAK_CHAINS; ++i) s ^= x[i];
    if (s == 0x1234567812345678ull) out[0] = 1.f;
}

__global__ void __launch_bounds__(256) peak_dfma_kernel(float* out, double a, double b)
{
    double x[PEAK_CHAINS];
#pragma unroll
    for (int i = 0; i < PEAK_CHAINS; ++i) x[i] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < PEAK_ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < PEAK_CHAINS; ++i) x[i] = fma(x[i], a, b);
    }
    double s = 0.;
#pragma unroll
    for (int i = 0; i < PEAK_CHAINS; ++i) s += x[i];
    if (s == 12345.678) out[0] = (float) s;
}

__global__ void __launch_bounds__(256) peak_mufu_kernel(float* out, float a, long long* clocks)
{
    float x[PEAK_CHAINS];
#pragma unroll
    for (int i = 0; i < PEAK_CHAINS; ++i) x[i] = threadIdx.x * 1e-3f + i * 0.1f;
    unsigned long long t0, t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    const long long c0 = clock64();
    for (int it = 0; it < PEAK_ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < PEAK_CHAINS; i += 2) {
            x[i] = __sinf(x[i] + a);
            x[i + 1] = __cosf(x[i + 1] + a);
        }
    }
    const long long c1 = clock64();
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < PEAK_CHAINS; ++i) s += x[i];
    if (s == 12345.678f) out[0] = s;
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        clocks[0] = c1 - c0;              // SM cycles ...
        clocks[1] = (long long) (t1 - t0); // ... over this many nanoseconds
    }
}

template<class L>
static int time_kernel(L launch, float* ms_best)
{
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(e0);
        launch();
        cudaEventRecord(e1);
        cudaError_t e = cudaEventSynchronize(e1);
        if (e != cudaSuccess) return (int) e;
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0) best = std::min(best, ms);
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    *ms_best = best;
    return (int) cudaGetLastError();
}

int measure_peaks(int device, I3B_Peaks* out)
{
    std::memset(out, 0, sizeof *out);
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) return (int) e;
    cudaDeviceProp prop;
    e = cudaGetDeviceProperties(&prop, device);
    if (e != cudaSuccess) return (int) e;
    out->sm_count = prop.multiProcessorCount;
    float* d_out = nullptr;
    long long* d_clk = nullptr;
    if ((e = cudaMalloc(&d_out, 64)) != cudaSuccess) return (int) e;
    if ((e = cudaMalloc(&d_clk, 64)) != cudaSuccess) return (int) e;
    const int grid = prop.multiProcessorCount * 16, block = 256;
    const double nthreads = (double) grid * block;
    const double ops = nthreads * PEAK_ITERS * PEAK_CHAINS;
    float ms = 0.f;
    int rc;
    rc = time_kernel([&]() { peak_ffma_kernel<<<grid, block>>>(d_out, 0.999f, 0.001f); }, &ms);
    if (rc) return rc;
    const double ffma = 2.0 * ops / (ms * 1e-3) / 1e12;
    rc = time_kernel([&]() { peak_ffma2_kernel<<<grid, block>>>(d_out, 0.999f, 0.001f); }, &ms);
    if (rc) return rc;
    const double ffma2 = 4.0 * ops / (ms * 1e-3) / 1e12;
    out->fp32_tflops = std::max(ffma, ffma2);
    rc = time_kernel([&]() { peak_dfma_kernel<<<grid, block>>>(d_out, 0.999, 0.001); }, &ms);
    if (rc) return rc;
    out->fp64_tflops = 2.0 * ops / (ms * 1e-3) / 1e12;
    rc = time_kernel([&]() { peak_mufu_kernel<<<grid, block>>>(d_out, 0.001f, d_clk); }, &ms);
    if (rc) return rc;
    out->sfu_gops = ops / (ms * 1e-3) / 1e9;
    long long clk[2] = {0, 0};
    cudaMemcpy(clk, d_clk, sizeof clk, cudaMemcpyDeviceToHost);
    out->sm_mhz = clk[1] > 0 ? (double) clk[0] / (double) clk[1] * 1e3 : 0.0; // cycles per ns -> MHz
    out->_pad = (int) (ffma2 > ffma); // 1 when the packed form was the faster one
    cudaFree(d_out);
    cudaFree(d_clk);
    return 0;
}

} // namespace i3b
