// fp64 FMA peak of the device (vector pipe): 8 independent DFMA chains per thread, full occupancy.
// Also an LDS.64 random-gather microbenchmark is NOT here; this only prints DFMA/s and the SM clock seen.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(256) dfma(double* out, int iters, double a, double b) {
    double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < iters; ++i) {
        x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
        x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}
int main() {
    int sms = 0, clk = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const int blocks = sms * 8, threads = 256, iters = 20000;
    double* out;
    cudaMalloc(&out, sizeof(double) * blocks * threads);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0);
        dfma<<<blocks, threads>>>(out, iters, 0.999999, 1e-6);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        double n = (double)blocks * threads * iters * 8;
        printf("{\"sms\": %d, \"max_clock_mhz\": %d, \"dfma_per_s\": %.4e, \"fp64_tflops\": %.2f, \"ms\": %.3f, \"dfma_per_clk_per_sm_at_max_clock\": %.1f}\n",
               sms, clk / 1000, n / (ms * 1e-3), 2 * n / (ms * 1e-3) / 1e12, ms, n / (ms * 1e-3) / sms / (clk * 1e3));
    }
    return 0;
}
