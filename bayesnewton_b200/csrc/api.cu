// Library-level pieces of the C ABI: error slot, sizing queries and the stand-alone
// discretisation kernel (As, Qs materialised for callers that want the reference's arrays:
// vmap(kernel.state_transition)(dt), vmap(process_noise_covariance), ops.py:274-278).
#include <cstring>
#include <vector>
#include <string>
#include <map>
#include "common.cuh"
#include "core.cuh"
#include "scan.cuh"

namespace bn {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

// ---- per-kernel timing ---------------------------------------------------------------------------
struct KRec { const char* name; cudaEvent_t e0, e1; };
static bool g_timing = false;
static std::vector<KRec> g_recs;

void ktimer_begin(const char* name, cudaStream_t st) {
    if (!g_timing) return;
    KRec r{name, nullptr, nullptr};
    cudaEventCreate(&r.e0);
    cudaEventCreate(&r.e1);
    cudaEventRecord(r.e0, st);
    g_recs.push_back(r);
}
void ktimer_end(cudaStream_t st) {
    if (!g_timing) return;
    cudaEventRecord(g_recs.back().e1, st);
}

template <class Gen>
__global__ void discretise_kernel(Gen gen, long long N, double* As, double* Qs) {
    constexpr int d = Gen::d;
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= N) return;
    double A[d * d], Q[symn(d)];
    gen.step(k, A, Q);
    double* a = As + k * (d * d);
    double* q = Qs + k * (d * d);
#pragma unroll
    for (int i = 0; i < d; ++i)
#pragma unroll
        for (int j = 0; j < d; ++j) {
            a[i * d + j] = A[i * d + j];
            q[i * d + j] = Q[sidx(i, j)];
        }
}

}  // namespace bn

using namespace bn;

extern "C" const char* bn_last_error(void) { return g_err; }
extern "C" int bn_version(void) { return 100; }

extern "C" int bn_timing_enable(int on) {
    for (auto& r : g_recs) { cudaEventDestroy(r.e0); cudaEventDestroy(r.e1); }
    g_recs.clear();
    g_timing = on != 0;
    return 0;
}

// "name count total_ms" per line, aggregated over the launches recorded since bn_timing_enable(1).
// Synchronises the recorded events.  Returns the number of bytes the full report needs.
extern "C" int bn_timing_report(char* buf, size_t len) {
    std::map<std::string, std::pair<long, double>> agg;
    std::vector<std::string> order;
    for (auto& r : g_recs) {
        cudaEventSynchronize(r.e1);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, r.e0, r.e1);
        if (!agg.count(r.name)) order.push_back(r.name);
        agg[r.name].first += 1;
        agg[r.name].second += ms;
    }
    std::string out;
    for (auto& n : order) {
        char line[160];
        snprintf(line, sizeof(line), "%s %ld %.6f\n", n.c_str(), agg[n].first, agg[n].second);
        out += line;
    }
    if (buf && len > 0) {
        size_t k = out.size() < len - 1 ? out.size() : len - 1;
        memcpy(buf, out.data(), k);
        buf[k] = 0;
    }
    return (int)out.size() + 1;
}

// fp64 FMA rate of the current device: 8 independent DFMA chains per thread at full occupancy (measurement
// aid for bench.py's fp64 roofline; ~3 ms).  Synchronises the device.
__global__ void __launch_bounds__(256) dfma_peak_kernel(double* out, int iters, double a, double b) {
    double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < iters; ++i) {
        x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
        x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
    }
    if (out) out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

extern "C" int bn_measure_dfma_peak(double* scratch, size_t scratch_doubles, double* dfma_per_s_host) {
    BN_REQUIRE(dfma_per_s_host != nullptr, "output is null");
    int dev = 0, sms = 0;
    BN_CUDA(cudaGetDevice(&dev));
    BN_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int blocks = sms * 8, threads = 256, iters = 20000;
    BN_REQUIRE(scratch && scratch_doubles >= (size_t)blocks * threads, "scratch of %d doubles needed", blocks * threads);
    cudaEvent_t e0, e1;
    BN_CUDA(cudaEventCreate(&e0));
    BN_CUDA(cudaEventCreate(&e1));
    double best = 0.0;
    for (int rep = 0; rep < 3; ++rep) {
        BN_CUDA(cudaEventRecord(e0, 0));
        dfma_peak_kernel<<<blocks, threads>>>(scratch, iters, 0.999999, 1e-6);
        BN_CUDA(cudaEventRecord(e1, 0));
        BN_CUDA(cudaEventSynchronize(e1));
        float ms = 0.f;
        BN_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        double r = (double)blocks * threads * iters * 8 / (ms * 1e-3);
        if (r > best) best = r;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    *dfma_per_s_host = best;
    return 0;
}

// fp64 tensor-pipe (DMMA) rate of the current device: 8 independent mma.sync.m8n8k4.f64 accumulator chains per warp at full
// occupancy; reported as fused multiply-adds per second (one DMMA.8x8x4 = 256 of them).  The roofline denominator of
// the dense spatio-temporal kernels.  Synchronises the device.
__global__ void __launch_bounds__(256) dmma_peak_kernel(double* out, int iters, double a, double b) {
    double c[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) c[i] = threadIdx.x + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[2 * i]), "+d"(c[2 * i + 1]) : "d"(a), "d"(b));
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += c[i];
    if (out) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

extern "C" int bn_measure_dmma_peak(double* scratch, size_t scratch_doubles, double* fma_per_s_host) {
    BN_REQUIRE(fma_per_s_host != nullptr, "output is null");
    int dev = 0, sms = 0;
    BN_CUDA(cudaGetDevice(&dev));
    BN_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int blocks = sms * 8, threads = 256, iters = 4000;
    BN_REQUIRE(scratch && scratch_doubles >= (size_t)blocks * threads, "scratch of %d doubles needed", blocks * threads);
    cudaEvent_t e0, e1;
    BN_CUDA(cudaEventCreate(&e0));
    BN_CUDA(cudaEventCreate(&e1));
    double best = 0.0;
    for (int rep = 0; rep < 3; ++rep) {
        BN_CUDA(cudaEventRecord(e0, 0));
        dmma_peak_kernel<<<blocks, threads>>>(scratch, iters, 1e-3, 1e-3);
        BN_CUDA(cudaEventRecord(e1, 0));
        BN_CUDA(cudaEventSynchronize(e1));
        float ms = 0.f;
        BN_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        const double r = (double)blocks * (threads / 32) * iters * 8 * 256.0 / (ms * 1e-3);
        if (r > best) best = r;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    *fma_per_s_host = best;
    return 0;
}

extern "C" int bn_state_dim(const bn_kernel_spec* k) {
    if (!k || k->n_components < 1 || k->n_components > BN_MAX_COMPONENTS) return -1;
    int n = family_dim(k->family);
    return n ? n * k->n_components : -1;
}

extern "C" int bn_rts_carry_len(int d) { return 2 * d * d + d; }

namespace bn { size_t gd_workspace_doubles(long long N, int d); }

extern "C" size_t bn_workspace_bytes(int64_t N, int d, int D) {
    (void)D;
    if (N < 0 || d < 1) return 0;
    ChunkPlan cp = plan_chunks(N > 0 ? N : 1);
    long long elem = (long long)d * d + 2 * d + 2 * symn(d);
    long long doubles = 64 + scan_plan_doubles(cp.nchunks, (int)elem) + cp.nchunks;
    long long site_partials = 4 * ((N + 127) / 128 + 1024);
    if (site_partials > doubles) doubles = site_partials;
    if (d <= 16) {  // the warp-cooperative path for the (d, D) pairs without a register-resident instantiation (gd.cu)
        const long long gd = (long long)gd_workspace_doubles(N, d);
        if (gd > doubles) doubles = gd;
    }
    return (size_t)(doubles + 1024) * sizeof(double);
}

extern "C" int bn_discretise(const bn_kernel_spec* k, int64_t N, const double* dt, double* As, double* Qs,
                             void* stream) {
    BN_REQUIRE(k != nullptr, "kernel spec is null");
    BN_REQUIRE(N >= 0, "N must be non-negative");
    if (N == 0) return 0;
    BN_REQUIRE(dt && As && Qs, "null array");
    unsigned grid = (unsigned)((N + 255) / 256);
#define X(FAM, NC)                                                                      \
    if (k->family == FAM && k->n_components == NC) {                                    \
        MaternGen<FAM, NC> gen;                                                         \
        gen.spec = *k;                                                                  \
        gen.dt = dt;                                                                    \
        discretise_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(gen, N, As, Qs);      \
        BN_CUDA(cudaGetLastError());                                                    \
        return 0;                                                                       \
    }
    BN_FOR_EACH_MATERN(X)
#undef X
    set_error("unsupported kernel spec: family %d with %d components", k->family, k->n_components);
    return -1;
}
