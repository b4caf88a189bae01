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

namespace BN_NS {

constexpr int kScanGroup = 288;  // 9 warps: 75 776 chunk elements (one resident wave) scan in two levels (264 groups <= 288)

// in: n elements (SoA stride n_stride).  out_prefix: within-group inclusive prefixes (same
// indexing).  totals: one element per group (SoA stride t_stride), nullable on the top level.
// load that bypasses L1 (data another CTA of the same launch has just written)
__device__ __forceinline__ real ld_cg(const real* p) {
#ifdef BN_REAL32
    return real(__ldcg(&p->v));
#else
    return __ldcg(p);
#endif
}

// CG: the inputs were written by other CTAs of this launch -- read them from L2.  (An element is kElem reals in the
// field order of Alg::load.)
template <class Alg, bool CG = false>
__device__ __forceinline__ void scan_group(const real* in, long long n, long long in_stride, real* out_prefix, long long out_stride,
                                           real* totals, long long t_stride, long long group) {
    using Elem = typename Alg::Elem;
    __shared__ real sh[(kScanGroup / 32) * Alg::kElem];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long i = group * kScanGroup + threadIdx.x;
    Elem mine;
    if (i >= n) {
        Alg::identity(mine);
    } else if constexpr (CG) {
        real* q = reinterpret_cast<real*>(&mine);
#pragma unroll
        for (int k = 0; k < Alg::kElem; ++k) q[k] = ld_cg(in + (long long)k * in_stride + i);
    } else {
        Alg::load(in, in_stride, i, mine);
    }
    // elements and warps of this group that hold data (CTA-uniform): a short top level skips the rounds it cannot need
    const long long left = n - group * kScanGroup;
    const int nin = left < kScanGroup ? (int)left : kScanGroup, nwarps = (nin + 31) >> 5;
#pragma unroll 1
    for (int off = 1; off < 32 && off < nin; off <<= 1) {
        Elem other = mine;
        Alg::shfl_up(other, off);
        if (lane >= off) { Elem r; Alg::combine(other, mine, r); mine = r; }
    }
    if (nwarps == 1) {
        if (i < n) Alg::store(out_prefix, out_stride, i, mine);
        if (totals && threadIdx.x == nin - 1) Alg::store(totals, t_stride, group, mine);
        return;
    }
    real* mp = reinterpret_cast<real*>(&mine);
    if (lane == 31) {
#pragma unroll
        for (int k = 0; k < Alg::kElem; ++k) sh[warp * Alg::kElem + k] = mp[k];
    }
    __syncthreads();
    if (warp == 0) {
        Elem w;
        real* wp = reinterpret_cast<real*>(&w);
        if (lane < kScanGroup / 32) {
#pragma unroll
            for (int k = 0; k < Alg::kElem; ++k) wp[k] = sh[lane * Alg::kElem + k];
        } else {
            Alg::identity(w);
        }
#pragma unroll 1
        for (int off = 1; off < nwarps; off <<= 1) {
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
        real* pp = reinterpret_cast<real*>(&prev);
#pragma unroll
        for (int k = 0; k < Alg::kElem; ++k) pp[k] = sh[(warp - 1) * Alg::kElem + k];
        Alg::combine(prev, mine, r);
        mine = r;
    }
    if (i < n) Alg::store(out_prefix, out_stride, i, mine);
    if (totals && threadIdx.x == nin - 1) Alg::store(totals, t_stride, group, mine);
}

template <class Alg>
__global__ void __launch_bounds__(kScanGroup)
scan_level_kernel(const real* in, long long n, long long in_stride,
                  real* out_prefix, long long out_stride,
                  real* totals, long long t_stride) {
    scan_group<Alg>(in, n, in_stride, out_prefix, out_stride, totals, t_stride, blockIdx.x);
}

// The last two levels in ONE launch: every CTA scans its group of level l (within-group prefixes stay as they are) and
// leaves its total in level l + 1; the CTA that arrives last (ticket) scans those totals in place.  The consumer applies
// group prefix, within-group prefix and within-warp prefix one after the other (apply_prefix), so no pass goes back down.
// `ticket` must be zero at launch (the producer kernel clears it); it is left at zero.
template <class Alg>
__global__ void __launch_bounds__(kScanGroup)
scan_last_levels_kernel(real* level, long long n, real* top, long long n_top, unsigned int* ticket) {
    __shared__ unsigned int last;
    scan_group<Alg>(level, n, n, level, n, top, n_top, blockIdx.x);
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) last = (atomicAdd(ticket, 1u) == gridDim.x - 1) ? 1u : 0u;
    __syncthreads();
    if (!last) return;
    __threadfence();
    scan_group<Alg, true>(top, n_top, n_top, top, n_top, nullptr, 0, 0);
    if (threadIdx.x == 0) *ticket = 0u;
}

// prefix[i] (within-group inclusive) <- combine(group_prefix[group(i) - 1], prefix[i]) : after this
// prefix[] holds inclusive prefixes over the whole level.
template <class Alg>
__global__ void __launch_bounds__(kScanGroup)
scan_down_kernel(real* prefix, long long n, long long stride,
                 const real* group_prefix, long long g_stride) {
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
    real* prefix[kMaxLevels];  // inclusive prefixes of level l (count[l] elements)
    real* input0 = nullptr;    // level-0 input aggregates
    unsigned int* ticket = nullptr;  // arrival counter of the one-launch upper levels (scan_last_levels_kernel), or null
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

inline ScanPlan make_scan_plan(real* ws, long long n0, int elem) {
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
        const real* in = (l == 0) ? p.input0 : p.prefix[l];
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

// ---- level 0 scanned by the producer of the elements, one warp per group of 32 consecutive elements -------------
// The kernel that forms the chunk elements holds 32 consecutive ones in the lanes of a warp: it scans them in registers
// (5 shuffle rounds) before they leave the SM, stores the within-warp inclusive prefixes and the warp total, and only the
// totals (n / 32 elements) go through the level kernels.  The consumer applies the two parts one after the other
// (apply_prefix2): no pass over the level-0 elements is left in the scan itself.
constexpr int kWarpGroup = 32;

inline long long scan_upper_doubles(long long n0, int elem) {
    long long tot = 0, n = (n0 + kWarpGroup - 1) / kWarpGroup;
    while (true) {
        tot += n * elem;
        if (n <= kScanGroup) break;
        n = (n + kScanGroup - 1) / kScanGroup;
    }
    return tot;
}

// levels 0 (input0 / prefix0 given) and >= 1 (laid out in `upper`, scan_upper_doubles elements)
inline ScanPlan make_scan_plan_warp(real* input0, real* prefix0, real* upper, long long n0, int elem) {
    ScanPlan p;
    p.input0 = input0;
    p.count[0] = n0;
    p.prefix[0] = prefix0;
    p.levels = 1;
    if (n0 <= kWarpGroup) return p;
    long long n = (n0 + kWarpGroup - 1) / kWarpGroup;
    while (true) {
        p.count[p.levels] = n;
        p.prefix[p.levels] = upper;
        upper += n * elem;
        ++p.levels;
        if (n <= kScanGroup) break;
        n = (n + kScanGroup - 1) / kScanGroup;
    }
    return p;
}

// after the producer's warp_prescan: prefix[1] <- inclusive prefixes of the warp totals
// levels == 3 (the wave-sized plans): ONE launch, and the consumer gets the three parts (upper_parts)
template <class Alg>
inline cudaError_t run_scan_upper(const ScanPlan& p, cudaStream_t st) {
    unsigned int* ticket = p.ticket;
    if (p.levels == 3 && ticket) {
        const unsigned grid = (unsigned)p.count[2];
        BN_LAUNCH("scan_level", st,
                  scan_last_levels_kernel<Alg><<<grid, kScanGroup, 0, st>>>(p.prefix[1], p.count[1], p.prefix[2], p.count[2], ticket));
        return cudaGetLastError();
    }
    for (int l = 1; l < p.levels; ++l) {
        const long long n = p.count[l];
        const unsigned grid = (unsigned)((n + kScanGroup - 1) / kScanGroup);
        const bool top = (l == p.levels - 1);
        BN_LAUNCH("scan_level", st,
                  scan_level_kernel<Alg><<<grid, kScanGroup, 0, st>>>(p.prefix[l], n, n, p.prefix[l], n,
                                                                      top ? nullptr : p.prefix[l + 1], top ? 0 : p.count[l + 1]));
    }
    for (int l = p.levels - 2; l >= 1; --l) {
        const long long n = p.count[l];
        const unsigned grid = (unsigned)((n + kScanGroup - 1) / kScanGroup);
        BN_LAUNCH("scan_down", st,
                  scan_down_kernel<Alg><<<grid, kScanGroup, 0, st>>>(p.prefix[l], n, n, p.prefix[l + 1], p.count[l + 1]));
    }
    return cudaGetLastError();
}

// The upper parts of a prefix: warp = inclusive prefixes of the warp totals -- over the whole level (group == null) or
// within groups of `gsize` warps, in which case group = inclusive prefixes of the group totals.  All null: prefix0
// already holds the global prefixes.
struct PrefixParts {
    const real* warp = nullptr;
    long long n_warp = 0;
    const real* group = nullptr;
    long long n_group = 0;
    int gsize = 0;
};

// what the consumer of a plan scanned by run_scan_upper applies (fused = the plan went through the one-launch path)
inline PrefixParts upper_parts(const ScanPlan& p) {
    PrefixParts u;
    if (p.levels > 1) { u.warp = p.prefix[1]; u.n_warp = p.count[1]; }
    if (p.levels == 3 && p.ticket) { u.group = p.prefix[2]; u.n_group = p.count[2]; u.gsize = kScanGroup; }
    return u;
}

#ifdef __CUDACC__
// every lane of the warp calls this (i = element index of the lane, consecutive across the lanes and warp-aligned);
// lanes past the end contribute the identity
template <class Alg>
__device__ __forceinline__ void warp_prescan(typename Alg::Elem& mine, long long i, const ScanPlan& p) {
    const int lane = threadIdx.x & 31;
    const long long n = p.count[0];
    if (i >= n) Alg::identity(mine);
#pragma unroll 1
    for (int off = 1; off < 32; off <<= 1) {
        typename Alg::Elem other = mine;
        Alg::shfl_up(other, off);
        if (lane >= off) { typename Alg::Elem r; Alg::combine(other, mine, r); mine = r; }
    }
    if (i < n) Alg::store(p.prefix[0], n, i, mine);
    if (lane == 31 && p.levels > 1) Alg::store(p.prefix[1], p.count[1], i >> 5, mine);
    if (i == 0 && p.ticket) *p.ticket = 0u;  // the upper levels run after this kernel and count their CTAs from zero
}
#endif

// state after the elements 0..q of the scan order, from the incoming state s (prefix0 = within-warp inclusive prefixes)
template <class Alg>
BN_DEV void apply_prefix(const real* prefix0, long long n0, const PrefixParts& up, long long q, typename Alg::State& s) {
    typename Alg::Elem e;
    typename Alg::State t;
    if (up.warp && (q >> 5) > 0) {
        const long long w1 = (q >> 5) - 1;  // the last whole warp below q
        if (up.group && w1 / up.gsize > 0) {
            Alg::load(up.group, up.n_group, w1 / up.gsize - 1, e);
            Alg::apply(e, s, t);
            s = t;
        }
        Alg::load(up.warp, up.n_warp, w1, e);
        Alg::apply(e, s, t);
        s = t;
    }
    Alg::load(prefix0, n0, q, e);
    Alg::apply(e, s, t);
    s = t;
}

// ---- carries exchanged between time shards (multi-GPU two-level scan)
template <class Alg>
BN_DEV void export_carry_body(const real* top_prefix, long long n_top, real* carry) {
    typename Alg::Elem e;
    Alg::load(top_prefix, n_top, n_top - 1, e);
    Alg::to_carry(e, carry);
}

template <class Alg>
__global__ void export_carry_kernel(const real* top_prefix, long long n_top, real* carry) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    export_carry_body<Alg>(top_prefix, n_top, carry);
}

// s0 <- fold of the carries of the ranks that precede this one in scan order
// (filter: ranks 0..rank-1 ascending; smoother: ranks world-1..rank+1 descending)
template <class Alg>
BN_DEV void fold_carries_body(const real* carries, int first, int last_excl, int step, real* s0) {
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
__global__ void fold_carries_kernel(const real* carries, int first, int last_excl, int step, real* s0) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    fold_carries_body<Alg>(carries, first, last_excl, step, s0);
}

}  // namespace BN_NS
