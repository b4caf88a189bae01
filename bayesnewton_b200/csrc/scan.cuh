// Multi-level inclusive scan of block aggregates under an associative algebra (FilterAlg or
// SmootherAlg).  This is the middle phase of the three-phase temporally parallel filter /
// smoother (reduce chunk -> scan aggregates -> re-run chunk from its incoming state), the GPU
// counterpart of lax.associative_scan in bayesnewton/ops.py:241,348.
//
// Level kernel: one CTA scans kScanGroup consecutive elements: Kogge-Stone inside each warp
// through register shuffles (5 combines), Kogge-Stone over the warp totals (3 combines), one
// fix-up combine.  The within-group inclusive prefixes stay in HBM (SoA, coalesced); the group
// totals feed the next level.  Going back down costs ONE combine per element.
#pragma once
#include "common.cuh"
#include "core.cuh"

namespace bn {

constexpr int kScanGroup = 288;  // 9 warps: 75 776 chunk elements (one resident wave) scan in two levels (264 groups <= 288)

// in: n elements (SoA stride n_stride).  out_prefix: within-group inclusive prefixes (same
// indexing).  totals: one element per group (SoA stride t_stride), nullable on the top level.
template <class Alg>
__global__ void __launch_bounds__(kScanGroup)
scan_level_kernel(const double* in, long long n, long long in_stride,
                  double* out_prefix, long long out_stride,
                  double* totals, long long t_stride) {
    using Elem = typename Alg::Elem;
    __shared__ double sh[(kScanGroup / 32) * Alg::kElem];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long i = (long long)blockIdx.x * kScanGroup + threadIdx.x;
    Elem mine;
    if (i < n) Alg::load(in, in_stride, i, mine); else Alg::identity(mine);
#pragma unroll 1
    for (int off = 1; off < 32; off <<= 1) {
        Elem other = mine;
        Alg::shfl_up(other, off);
        if (lane >= off) { Elem r; Alg::combine(other, mine, r); mine = r; }
    }
    double* mp = reinterpret_cast<double*>(&mine);
    if (lane == 31) {
#pragma unroll
        for (int k = 0; k < Alg::kElem; ++k) sh[warp * Alg::kElem + k] = mp[k];
    }
    __syncthreads();
    if (warp == 0) {
        Elem w;
        double* wp = reinterpret_cast<double*>(&w);
        if (lane < kScanGroup / 32) {
#pragma unroll
            for (int k = 0; k < Alg::kElem; ++k) wp[k] = sh[lane * Alg::kElem + k];
        } else {
            Alg::identity(w);
        }
#pragma unroll 1
        for (int off = 1; off < kScanGroup / 32; off <<= 1) {
            Elem other = w;
            Alg::shfl_up(other, off);
            if (lane >= off) { Elem r; Alg::combine(other, w, r); w = r; }
        }
        if (lane < kScanGroup / 32) {
#pragma unroll
            for (int k = 0; k < Alg::kElem; ++k) sh[lane * Alg::kElem + k] = wp[k];
        }
    }
    __syncthreads();
    if (warp > 0) {
        Elem prev, r;
        double* pp = reinterpret_cast<double*>(&prev);
#pragma unroll
        for (int k = 0; k < Alg::kElem; ++k) pp[k] = sh[(warp - 1) * Alg::kElem + k];
        Alg::combine(prev, mine, r);
        mine = r;
    }
    if (i < n) Alg::store(out_prefix, out_stride, i, mine);
    if (totals && threadIdx.x == kScanGroup - 1) Alg::store(totals, t_stride, blockIdx.x, mine);
}

// prefix[i] (within-group inclusive) <- combine(group_prefix[group(i) - 1], prefix[i]) : after this
// prefix[] holds inclusive prefixes over the whole level.
template <class Alg>
__global__ void __launch_bounds__(kScanGroup)
scan_down_kernel(double* prefix, long long n, long long stride,
                 const double* group_prefix, long long g_stride) {
    using Elem = typename Alg::Elem;
    const long long i = (long long)blockIdx.x * kScanGroup + threadIdx.x;
    if (blockIdx.x == 0 || i >= n) return;
    Elem prev, mine, r;
    Alg::load(group_prefix, g_stride, (long long)blockIdx.x - 1, prev);
    Alg::load(prefix, stride, i, mine);
    Alg::combine(prev, mine, r);
    Alg::store(prefix, stride, i, r);
}

// Host-side plan of the level buffers inside the workspace.
struct ScanPlan {
    static constexpr int kMaxLevels = 6;
    int levels = 0;
    long long count[kMaxLevels];
    double* prefix[kMaxLevels];  // inclusive prefixes of level l (count[l] elements)
    double* input0 = nullptr;    // level-0 input aggregates
};

inline long long scan_plan_doubles(long long n0, int elem) {
    long long tot = n0 * elem;  // input0
    long long n = n0;
    while (true) {
        tot += n * elem;
        if (n <= kScanGroup) break;
        n = (n + kScanGroup - 1) / kScanGroup;
    }
    return tot;
}

inline ScanPlan make_scan_plan(double* ws, long long n0, int elem) {
    ScanPlan p;
    p.input0 = ws;
    ws += n0 * elem;
    long long n = n0;
    while (true) {
        p.count[p.levels] = n;
        p.prefix[p.levels] = ws;
        ws += n * elem;
        ++p.levels;
        if (n <= kScanGroup) break;
        n = (n + kScanGroup - 1) / kScanGroup;
    }
    return p;
}

// Runs the full scan: after return plan.prefix[0] holds the inclusive prefixes of input0 and
// plan.prefix[levels-1][count-1] is the total.  Level l>=1 input = totals of level l-1, stored
// temporarily in prefix[l] (scanned in place).
template <class Alg>
inline cudaError_t run_scan(const ScanPlan& p, cudaStream_t st) {
    for (int l = 0; l < p.levels; ++l) {
        long long n = p.count[l];
        unsigned grid = (unsigned)((n + kScanGroup - 1) / kScanGroup);
        const double* in = (l == 0) ? p.input0 : p.prefix[l];
        bool top = (l == p.levels - 1);
        BN_LAUNCH("scan_level", st,
                  scan_level_kernel<Alg><<<grid, kScanGroup, 0, st>>>(
                      in, n, n, p.prefix[l], n, top ? nullptr : p.prefix[l + 1], top ? 0 : p.count[l + 1]));
    }
    for (int l = p.levels - 2; l >= 0; --l) {
        long long n = p.count[l];
        unsigned grid = (unsigned)((n + kScanGroup - 1) / kScanGroup);
        BN_LAUNCH("scan_down", st,
                  scan_down_kernel<Alg><<<grid, kScanGroup, 0, st>>>(p.prefix[l], n, n, p.prefix[l + 1],
                                                                      p.count[l + 1]));
    }
    return cudaGetLastError();
}

// ---- carries exchanged between time shards (multi-GPU two-level scan)
template <class Alg>
BN_DEV void export_carry_body(const double* top_prefix, long long n_top, double* carry) {
    typename Alg::Elem e;
    Alg::load(top_prefix, n_top, n_top - 1, e);
    Alg::to_carry(e, carry);
}

template <class Alg>
__global__ void export_carry_kernel(const double* top_prefix, long long n_top, double* carry) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    export_carry_body<Alg>(top_prefix, n_top, carry);
}

// s0 <- fold of the carries of the ranks that precede this one in scan order
// (filter: ranks 0..rank-1 ascending; smoother: ranks world-1..rank+1 descending)
template <class Alg>
BN_DEV void fold_carries_body(const double* carries, int first, int last_excl, int step, double* s0) {
    typename Alg::State s;
    Alg::zero_state(s);
    for (int r = first; r != last_excl; r += step) {
        typename Alg::Elem e;
        Alg::from_carry(carries + (long long)r * Alg::kCarry, e);
        typename Alg::State t;
        Alg::apply(e, s, t);
        s = t;
    }
    Alg::store_state(s0, 1, 0, s);
}

template <class Alg>
__global__ void fold_carries_kernel(const double* carries, int first, int last_excl, int step, double* s0) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    fold_carries_body<Alg>(carries, first, last_excl, step, s0);
}

}  // namespace bn
