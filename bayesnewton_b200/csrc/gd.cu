// Warp-cooperative filter / smoother for state dimensions up to 16 (gd_impl.cuh): kernels and host drivers behind
// bn_kf_arrays / bn_rts_arrays for the (d, D) pairs the register-resident instantiations do not cover.
#include "gd_impl.cuh"

namespace bn {

static __device__ __forceinline__ GdW gd_lane() { return GdW{(int)(threadIdx.x & 31), 32}; }
static __device__ __forceinline__ GdPool gd_warp_pool(double* smem, int d) {
    return GdPool{smem + (threadIdx.x >> 5) * gd_pool_doubles(d)};
}

__global__ void __launch_bounds__(32) gd_kf_seq_kernel(GdKf a, double* ell_out, int want_ell) {
    extern __shared__ double gd_smem[];
    GdPool pool = gd_warp_pool(gd_smem, a.d);
    const GdW w = gd_lane();
    double* m = pool.take(a.d);
    double* P = pool.take(a.d * a.d);
    gd_copy(w, m, a.m0, a.d);
    gd_copy(w, P, a.P0, a.d * a.d);
    const double ell = gd_kf_run(w, a, 0, a.N, m, P, want_ell != 0, pool);
    if (want_ell && ell_out && w.lane == 0) *ell_out = ell;
}

__global__ void __launch_bounds__(32 * kGdWarps) gd_kf_reduce_kernel(GdKf a, int L, long long nchunks, double* agg) {
    extern __shared__ double gd_smem[];
    const long long c = (long long)blockIdx.x * kGdWarps + (threadIdx.x >> 5);
    if (c >= nchunks) return;
    gd_kf_reduce_chunk(gd_lane(), a, L, c, agg, gd_warp_pool(gd_smem, a.d));
}

template <class K>
static int gd_allow_smem(K kern, size_t bytes) {
    BN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    return 0;
}

// phase 2 (gd_impl.cuh: gd_scan_group / gd_scan_down), one warp per group resp. per element
template <bool FILTER>
__global__ void __launch_bounds__(32 * kGdWarps)
gd_scan_group_kernel(int d, long long n, const double* in, double* prefix, double* totals) {
    extern __shared__ double gd_smem[];
    const long long g = (long long)blockIdx.x * kGdWarps + (threadIdx.x >> 5);
    if (g * kGdScanGroup >= n) return;
    gd_scan_group<FILTER>(gd_lane(), d, n, in, prefix, totals, g, gd_warp_pool(gd_smem, d));
}
template <bool FILTER>
__global__ void __launch_bounds__(32 * kGdWarps) gd_scan_down_kernel(int d, long long n, double* prefix, const double* upper) {
    extern __shared__ double gd_smem[];
    const long long i = (long long)blockIdx.x * kGdWarps + (threadIdx.x >> 5) + kGdScanGroup;  // group 0 needs nothing
    if (i >= n) return;
    gd_scan_down<FILTER>(gd_lane(), d, prefix, upper, i, gd_warp_pool(gd_smem, d));
}

// inclusive scan of the n chunk elements agg[] into prefix[]; `upper` = gd_scan_upper_elems(n) elements of scratch
template <bool FILTER>
static int gd_run_scan(int d, long long n, const double* agg, double* prefix, double* upper, size_t warp_smem, cudaStream_t st) {
    const int ne = gd_elem<FILTER>(d);
    if (int rc = gd_allow_smem(gd_scan_group_kernel<FILTER>, kGdWarps * warp_smem)) return rc;
    if (int rc = gd_allow_smem(gd_scan_down_kernel<FILTER>, kGdWarps * warp_smem)) return rc;
    // up: level l scanned in groups; its totals are level l + 1 (scanned in place)
    constexpr int kMaxLevels = 8;
    long long cnt[kMaxLevels];
    double* arr[kMaxLevels];
    int levels = 0;
    cnt[0] = n; arr[0] = prefix;
    const double* in = agg;
    while (true) {
        const long long m = cnt[levels], groups = (m + kGdScanGroup - 1) / kGdScanGroup;
        double* totals = groups > 1 ? upper : nullptr;
        BN_LAUNCH("gd_scan", st, (gd_scan_group_kernel<FILTER><<<(unsigned)((groups + kGdWarps - 1) / kGdWarps), 32 * kGdWarps,
                                                              kGdWarps * warp_smem, st>>>(d, m, in, arr[levels], totals)));
        if (groups <= 1) break;
        BN_REQUIRE(levels + 1 < kMaxLevels, "scan too deep");
        ++levels;
        cnt[levels] = groups; arr[levels] = upper; in = upper;
        upper += groups * ne;
    }
    for (int l = levels - 1; l >= 0; --l) {
        const long long m = cnt[l] - kGdScanGroup;  // elements outside group 0
        if (m <= 0) continue;
        BN_LAUNCH("gd_scan_down", st, (gd_scan_down_kernel<FILTER><<<(unsigned)((m + kGdWarps - 1) / kGdWarps), 32 * kGdWarps,
                                                                  kGdWarps * warp_smem, st>>>(d, cnt[l], arr[l], arr[l + 1])));
    }
    BN_CUDA(cudaGetLastError());
    return 0;
}

__global__ void __launch_bounds__(32 * kGdWarps)
gd_kf_apply_kernel(GdKf a, int L, long long nchunks, const double* prefix, double* ell_partials, int want_ell) {
    extern __shared__ double gd_smem[];
    const long long c = (long long)blockIdx.x * kGdWarps + (threadIdx.x >> 5);
    if (c >= nchunks) return;
    const GdW w = gd_lane();
    const double ell = gd_kf_apply_chunk(w, a, L, c, prefix, want_ell != 0, gd_warp_pool(gd_smem, a.d));
    if (want_ell && w.lane == 0) ell_partials[c] = ell;
}

__global__ void __launch_bounds__(32) gd_rts_seq_kernel(GdRts a) {
    extern __shared__ double gd_smem[];
    GdPool pool = gd_warp_pool(gd_smem, a.d);
    const GdW w = gd_lane();
    double* sm = pool.take(a.d);
    double* sP = pool.take(a.d * a.d);
    gd_fill(w, sm, 0.0, a.d);
    gd_fill(w, sP, 0.0, a.d * a.d);
    gd_rts_run(w, a, 0, a.N, sm, sP, true, pool);
}

__global__ void __launch_bounds__(32 * kGdWarps) gd_rts_reduce_kernel(GdRts a, int L, long long nchunks, double* agg) {
    extern __shared__ double gd_smem[];
    const long long c = (long long)blockIdx.x * kGdWarps + (threadIdx.x >> 5);
    if (c >= nchunks) return;
    gd_rts_reduce_chunk(gd_lane(), a, L, nchunks, c, agg, gd_warp_pool(gd_smem, a.d));
}


__global__ void __launch_bounds__(32 * kGdWarps)
gd_rts_apply_kernel(GdRts a, int L, long long nchunks, const double* prefix) {
    extern __shared__ double gd_smem[];
    const long long c = (long long)blockIdx.x * kGdWarps + (threadIdx.x >> 5);
    if (c >= nchunks) return;
    gd_rts_apply_chunk(gd_lane(), a, L, nchunks, c, prefix, gd_warp_pool(gd_smem, a.d));
}

size_t gd_workspace_doubles(long long N, int d) {
    const GdPlan p = gd_plan(N > 0 ? N : 1, d);
    return ((size_t)p.nchunks * 2 + (size_t)gd_scan_upper_elems(p.nchunks)) * gd_felem(d) + (size_t)p.nchunks + 64;
}

int gd_kf_arrays(int form, const GdKf& a, double* ell, void* ws, size_t ws_bytes, cudaStream_t st) {
    BN_REQUIRE(a.d >= 1 && a.d <= kGdMaxD && a.D >= 1 && a.D <= a.d,
               "the warp-cooperative filter covers 1 <= D <= d <= %d, got (d, D) = (%d, %d)", kGdMaxD, a.d, a.D);
    if (a.N == 0) return 0;
    const size_t warp_smem = (size_t)gd_pool_doubles(a.d) * sizeof(double);
    if (form == BN_SEQUENTIAL) {
        if (int rc = gd_allow_smem(gd_kf_seq_kernel, warp_smem)) return rc;
        BN_LAUNCH("gd_kf_seq", st, (gd_kf_seq_kernel<<<1, 32, warp_smem, st>>>(a, ell, ell != nullptr)));
        BN_CUDA(cudaGetLastError());
        return 0;
    }
    const GdPlan p = gd_plan(a.N, a.d);
    const size_t need = gd_workspace_doubles(a.N, a.d) * sizeof(double);
    BN_REQUIRE(ws != nullptr && ws_bytes >= need, "workspace too small: need %zu bytes, got %zu", need, ws_bytes);
    double* agg = (double*)ws;
    double* prefix = agg + (size_t)p.nchunks * gd_felem(a.d);
    double* upper = prefix + (size_t)p.nchunks * gd_felem(a.d);
    double* partials = upper + (size_t)gd_scan_upper_elems(p.nchunks) * gd_felem(a.d);
    const unsigned grid = (unsigned)((p.nchunks + kGdWarps - 1) / kGdWarps);
    if (int rc = gd_allow_smem(gd_kf_reduce_kernel, kGdWarps * warp_smem)) return rc;
    if (int rc = gd_allow_smem(gd_kf_apply_kernel, kGdWarps * warp_smem)) return rc;
    BN_LAUNCH("gd_kf_reduce", st, (gd_kf_reduce_kernel<<<grid, 32 * kGdWarps, kGdWarps * warp_smem, st>>>(a, p.L, p.nchunks, agg)));
    if (int rc = gd_run_scan<true>(a.d, p.nchunks, agg, prefix, upper, warp_smem, st)) return rc;
    BN_LAUNCH("gd_kf_apply", st, (gd_kf_apply_kernel<<<grid, 32 * kGdWarps, kGdWarps * warp_smem, st>>>(a, p.L, p.nchunks, prefix, partials,
                                                                                                     ell != nullptr)));
    BN_CUDA(cudaGetLastError());
    if (ell) {
        BN_LAUNCH("sum", st, (sum_kernel<false><<<1, 1024, 0, st>>>(partials, p.nchunks, ell, 1.0)));
        BN_CUDA(cudaGetLastError());
    }
    return 0;
}

int gd_rts_arrays(int form, const GdRts& a, void* ws, size_t ws_bytes, cudaStream_t st) {
    BN_REQUIRE(a.d >= 1 && a.d <= kGdMaxD && a.Df >= 1 && a.Df <= a.d,
               "the warp-cooperative smoother covers 1 <= D <= d <= %d, got (d, D) = (%d, %d)", kGdMaxD, a.d, a.Df);
    if (a.N == 0) return 0;
    const size_t warp_smem = (size_t)gd_pool_doubles(a.d) * sizeof(double);
    if (form == BN_SEQUENTIAL) {
        if (int rc = gd_allow_smem(gd_rts_seq_kernel, warp_smem)) return rc;
        BN_LAUNCH("gd_rts_seq", st, (gd_rts_seq_kernel<<<1, 32, warp_smem, st>>>(a)));
        BN_CUDA(cudaGetLastError());
        return 0;
    }
    const GdPlan p = gd_plan(a.N, a.d);
    const size_t need = gd_workspace_doubles(a.N, a.d) * sizeof(double);
    BN_REQUIRE(ws != nullptr && ws_bytes >= need, "workspace too small: need %zu bytes, got %zu", need, ws_bytes);
    double* agg = (double*)ws;
    double* prefix = agg + (size_t)p.nchunks * gd_selem(a.d);
    const unsigned grid = (unsigned)((p.nchunks + kGdWarps - 1) / kGdWarps);
    if (int rc = gd_allow_smem(gd_rts_reduce_kernel, kGdWarps * warp_smem)) return rc;
    if (int rc = gd_allow_smem(gd_rts_apply_kernel, kGdWarps * warp_smem)) return rc;
    BN_LAUNCH("gd_rts_reduce", st, (gd_rts_reduce_kernel<<<grid, 32 * kGdWarps, kGdWarps * warp_smem, st>>>(a, p.L, p.nchunks, agg)));
    double* supper = prefix + (size_t)p.nchunks * gd_selem(a.d);
    if (int rc = gd_run_scan<false>(a.d, p.nchunks, agg, prefix, supper, warp_smem, st)) return rc;
    BN_LAUNCH("gd_rts_apply", st, (gd_rts_apply_kernel<<<grid, 32 * kGdWarps, kGdWarps * warp_smem, st>>>(a, p.L, p.nchunks, prefix)));
    BN_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace bn
