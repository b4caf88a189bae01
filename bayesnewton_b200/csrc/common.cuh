// Shared host-side helpers of the C ABI: error slot, argument checks, chunk planning,
// deterministic reductions.
#pragma once
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdarg>
#include "../../include/bn_b200.h"
#include "gen.cuh"

namespace bn {
void set_error(const char* fmt, ...);
void ktimer_begin(const char* name, cudaStream_t st);
void ktimer_end(cudaStream_t st);
}  // namespace bn

namespace BN_NS {
#ifdef BN_REAL32
using ::bn::set_error;
#endif

#define BN_REQUIRE(cond, ...)                     \
    do {                                          \
        if (!(cond)) {                            \
            ::bn::set_error(__VA_ARGS__);         \
            return -1;                            \
        }                                         \
    } while (0)

#define BN_CUDA(expr)                                                            \
    do {                                                                         \
        cudaError_t err__ = (expr);                                              \
        if (err__ != cudaSuccess) {                                              \
            ::bn::set_error("%s: %s", #expr, cudaGetErrorString(err__));         \
            return (int)err__;                                                   \
        }                                                                        \
    } while (0)

// Optional per-kernel device timing (bn_timing_enable / bn_timing_report): when enabled, every
// launch made through BN_LAUNCH is bracketed by CUDA events on its own stream.
#define BN_LAUNCH(name, st, ...)          \
    do {                                  \
        ::bn::ktimer_begin(name, st);     \
        __VA_ARGS__;                      \
        ::bn::ktimer_end(st);             \
    } while (0)

constexpr int kChunkThreads = 128;
constexpr long long kTargetChunks = 148LL * 512;  // one resident wave of 512 threads per SM

struct ChunkPlan {
    int L;               // steps per chunk (one thread each)
    long long nchunks;
};

inline ChunkPlan plan_chunks(long long N) {
    long long L = (N + kTargetChunks - 1) / kTargetChunks;
    if (L < 4) L = 4;
    if (L > 128) L = 128;
    ChunkPlan p;
    p.L = (int)L;
    p.nchunks = (N + L - 1) / L;
    return p;
}

inline int family_dim(int family) {
    switch (family) {
        case BN_MATERN12: return 1;
        case BN_MATERN32: return 2;
        case BN_MATERN52: return 3;
        case BN_MATERN72: return 4;
    }
    return 0;
}

enum { PHASE_ALL = 0, PHASE_REDUCE = 1, PHASE_APPLY = 2 };
constexpr int kNotHandled = -1000;

// Instantiation groups: one translation unit each, so the build parallelises across cores.
// Stationary kernels (family, components):
#define BN_GROUP_M_A(X) X(BN_MATERN12, 1) X(BN_MATERN12, 2) X(BN_MATERN32, 1)
#define BN_GROUP_M_B(X) X(BN_MATERN32, 2) X(BN_MATERN32, 3)
#define BN_GROUP_M_C(X) X(BN_MATERN52, 1) X(BN_MATERN52, 2)
#define BN_GROUP_M_D(X) X(BN_MATERN72, 1)
#define BN_FOR_EACH_MATERN(X) BN_GROUP_M_A(X) BN_GROUP_M_B(X) BN_GROUP_M_C(X) BN_GROUP_M_D(X)
// Array entry (state dim, observation dim):
#define BN_GROUP_A_A(X) X(1, 1) X(2, 1) X(3, 1) X(4, 1)
#define BN_GROUP_A_B(X) X(2, 2) X(3, 2) X(4, 2) X(3, 3)
#define BN_GROUP_A_C(X) X(6, 2)
// pairs filter of the sparse Markov model (ops.py:383-426): state [u_-; u_+], H = I
#define BN_GROUP_A_D(X) X(4, 4) X(6, 6)

// sum of n doubles in a fixed order (strided partials, then a shared-memory tree): run-to-run
// bit-stable, unlike atomics.  NANSUM skips NaNs (np.nansum, inference.py:218).
template <bool NANSUM>
__global__ void __launch_bounds__(1024) sum_kernel(const real* x, long long n, real* out, real scale) {
    __shared__ double sh[1024];  // accumulated in fp64 whatever `real` is
    double s = 0.0;
    for (long long i = threadIdx.x; i < n; i += 1024) {
        double v = (double)x[i];
        if (NANSUM && isnan(v)) v = 0.0;
        s += v;
    }
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int off = 512; off > 0; off >>= 1) {
        if ((int)threadIdx.x < off) sh[threadIdx.x] += sh[threadIdx.x + off];
        __syncthreads();
    }
    if (threadIdx.x == 0) *out = real(sh[0] * (double)scale);
}

// block-level deterministic partial sum: every thread contributes v; thread 0 writes the block total
template <int THREADS>
__device__ __forceinline__ void block_sum_store(real v, real* partials) {
    __shared__ double sh_bs[THREADS];
    sh_bs[threadIdx.x] = (double)v;
    __syncthreads();
#pragma unroll
    for (int off = THREADS / 2; off > 0; off >>= 1) {
        if ((int)threadIdx.x < off) sh_bs[threadIdx.x] += sh_bs[threadIdx.x + off];
        __syncthreads();
    }
    if (threadIdx.x == 0) partials[blockIdx.x] = real(sh_bs[0]);
    __syncthreads();  // the shared buffer is reused by the next call
}

}  // namespace BN_NS
