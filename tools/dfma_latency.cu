// DFMA throughput vs. (independent chains per thread) x (warps per SM sub-partition): how much ILP x TLP the
// fp64 pipe needs.  Prints DFMA per clock per SM for each combination (64 = peak).
#include <cstdio>
#include <cuda_runtime.h>
template <int C>
__global__ void k(double* out, int iters, double a, double b) {
    double x[C];
#pragma unroll
    for (int i = 0; i < C; ++i) x[i] = threadIdx.x + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < C; ++i) x[i] = fma(x[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < C; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int C>
void run(double* out, int warps_per_smsp) {
    int sms = 148;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 40000 / C;
    int threads = 32 * 4 * warps_per_smsp;  // one CTA per SM
    k<C><<<sms, threads>>>(out, 100, 0.999999, 1e-6);
    cudaEventRecord(e0);
    k<C><<<sms, threads>>>(out, iters, 0.999999, 1e-6);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double n = (double)sms * threads * iters * C;
    printf("chains %d warps/smsp %d : %.1f DFMA/clk/SM (at 1.965 GHz)  cycles per dependent DFMA per warp: %.1f\n", C, warps_per_smsp,
           n / (ms * 1e-3) / sms / 1.965e9, (ms * 1e-3) * 1.965e9 / iters);
}
int main() {
    double* out; cudaMalloc(&out, 8 * 148 * 1024);
    for (int w : {1, 2, 4, 8}) { run<1>(out, w); run<2>(out, w); run<4>(out, w); run<8>(out, w); }
    return 0;
}
