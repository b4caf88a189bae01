// Fused posterior update  (MarkovGaussianProcess.update_posterior, basemodels.py:689-706:
// kalman_filter -> rauch_tung_striebel_smoother, ops.py:256-285, 357-380, `parallel=True` form)
// for stationary Matern stacks.  Five launches:
//   up_reduce    one thread per chunk of L steps folds its steps into a filtering element      (phase 1)
//   run_scan     prefix "product" of the chunk elements (scan.cuh)                            (phase 2)
//   up_filter    one thread per chunk re-runs the plain filter from its incoming state, keeps the
//                filtered states in a warp-tiled scratch layout and sums the log-likelihood    (phase 3)
//   up_selem     one thread per chunk turns its filtering element + the filtered states at the
//                chunk ends into the chunk's smoothing element (fast_core.cuh); run_scan again
//   up_smooth    one thread per chunk runs the RTS recursion down its chunk and writes H sm, H sP H^T
// The second per-step reduction pass a stand-alone parallel smoother needs (ops.py:318-325 over
// all N) does not exist here: the smoothing elements come from the filter's own chunk elements.
//
// HBM traffic per step (d = 3, D = 1, fp64): up_reduce 24 B, up_filter 24 + 72 B, up_smooth 72 + 8 + 16 B.
// Inputs and outputs move through per-warp shared-memory tiles (coalesced row segments of kUpTJ
// steps per chunk); the filtered-state scratch is laid out [warp tile][step][field][lane] so every
// access is a fully coalesced 256-byte warp transaction with no staging.
#pragma once
#include <cstdlib>
#include "common.cuh"
#include "fast_core.cuh"
#include "scan.cuh"

namespace BN_NS {

#ifndef BN_UP_TJ
#define BN_UP_TJ 8
#endif
#ifndef BN_UP_BLOCKS
#define BN_UP_BLOCKS 4
#endif
constexpr int kUpTJ = BN_UP_TJ;          // steps per staged sub-block
constexpr int kUpWarps = 4;              // warps per CTA
constexpr int kUpThreads = 32 * kUpWarps;
constexpr int kUpBlocksPerSM = BN_UP_BLOCKS;  // resident CTAs per SM the single-latent kernels are compiled for
// states with d > 3 need up to 255 registers per thread: two 128-thread CTAs per SM is what the register file holds
constexpr int kUpBlocksPerSMWide = 2;
constexpr long long kUpTargetChunks = 148LL * kUpBlocksPerSM * kUpThreads;  // one resident wave
#ifndef BN_UP_GRAD_BLOCKS
#define BN_UP_GRAD_BLOCKS 3
#endif
constexpr int kUpGradBlocksPerSM = BN_UP_GRAD_BLOCKS;  // the gradient-carrying smoother keeps more registers live

struct UpIO {
    long long N;
    const real* dt;           // [N]
    const real* y;            // [N,D]     pseudo observations (site means)
    const real* R;            // [N,D,D]   site covariances
    const unsigned char* mask;  // [N,D] or null
    real* post_mean;          // [N,D]     H sm
    real* post_cov;           // [N,D,D]   H sP H^T
};

// one resident wave of chunk threads: the kernels of states with d > 3 are compiled for one CTA per SM (registers), so
// their wave -- and with it the number of chunk elements the scan has to combine -- is four times smaller
// BN_B200_CHUNK_TARGET=<n> overrides the wave size of the d <= 3 plan (tuning aid; read once per process)
inline long long up_env_chunk_target() {
    static const long long v = [] {
        const char* e = getenv("BN_B200_CHUNK_TARGET");
        return e ? atoll(e) : 0LL;
    }();
    return v;
}
inline ChunkPlan up_plan_chunks(long long N, bool grad = false, int d = 3) {
    long long target = (d > 3) ? 148LL * kUpBlocksPerSMWide * kUpThreads : (grad ? 148LL * kUpGradBlocksPerSM * kUpThreads : kUpTargetChunks);
    if (d <= 3 && !grad && up_env_chunk_target() > 0) target = up_env_chunk_target();
    long long L = (N + target - 1) / target;
    L = (L + kUpTJ - 1) / kUpTJ * kUpTJ;
    if (L < kUpTJ) L = kUpTJ;  // no upper bound: for large N the chunk count stays at one resident wave
    ChunkPlan p;
    p.L = (int)L;
    p.nchunks = (N + L - 1) / L;
    return p;
}

// filtered-state scratch: field f of step j of chunk c
BN_DEV long long fs_index(long long c, int L, int j, int f, int nfields) {
    return ((((c >> 5) * L + j) * nfields + f) << 5) + (c & 31);
}
inline long long fs_doubles(long long nchunks, int L, int nfields) {
    return ((nchunks + 31) / 32) * 32 * (long long)L * nfields;
}

template <int d>
BN_DEV void fs_store(real* fs, long long c, int L, int j, const real* m, const real* P) {
    constexpr int nf = d + symn(d);
#pragma unroll
    for (int f = 0; f < d; ++f) fs[fs_index(c, L, j, f, nf)] = m[f];
#pragma unroll
    for (int f = 0; f < symn(d); ++f) fs[fs_index(c, L, j, d + f, nf)] = P[f];
}
template <int d>
BN_DEV void fs_load(const real* fs, long long c, int L, int j, real* m, real* P) {
    constexpr int nf = d + symn(d);
#pragma unroll
    for (int f = 0; f < d; ++f) m[f] = fs[fs_index(c, L, j, f, nf)];
#pragma unroll
    for (int f = 0; f < symn(d); ++f) P[f] = fs[fs_index(c, L, j, d + f, nf)];
}

// ------------------------------------------------------------------------------------------ IO contexts
// A context feeds one lane (= one chunk) its per-step inputs and takes its per-step outputs.
// DirectCtx touches global memory per step (host emulation harness; also valid on the device).
template <int D>
struct DirectCtx {
    UpIO io;
    BN_DEV real dt(long long k, int) const { return k < io.N ? io.dt[k] : 0.0; }
    BN_DEV void obs(long long k, int, real* y, real* R) const {
#pragma unroll
        for (int i = 0; i < D; ++i) y[i] = io.y[k * D + i];
#pragma unroll
        for (int i = 0; i < D * D; ++i) R[i] = io.R[k * (D * D) + i];
    }
    BN_DEV void put(long long k, int, const real* pm, const real* pc) {
#pragma unroll
        for (int i = 0; i < D; ++i) io.post_mean[k * D + i] = pm[i];
#pragma unroll
        for (int i = 0; i < D * D; ++i) io.post_cov[k * (D * D) + i] = pc[i];
    }
};

#ifdef __CUDACC__
__device__ __forceinline__ void cp_async_8(real* dst_smem, const real* src) {
    const unsigned a = (unsigned)__cvta_generic_to_shared(dst_smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(a), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int NPENDING>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(NPENDING) : "memory"); }

// WarpCtx: the 32 chunks of a warp stage kUpTJ steps at a time through shared memory.  Global
// accesses are row segments of kUpTJ * W consecutive doubles per chunk, read and written by
// consecutive lanes; each lane then walks its own row (odd row pitch: no bank conflicts).
// For D = 1 the input tiles are real-buffered and filled with cp.async (LDGSTS): the tile of
// the next sub-block is in flight while the warp computes the current one, so the HBM latency
// of the staging loads is off the critical path.
template <int W>
struct Tile {
    static constexpr int RW = kUpTJ * W;
    static constexpr int P = RW | 1;
    static constexpr int kDoubles = 32 * P;
    real* sm;
    __device__ void load(const real* X, long long N, long long cbase, int L, int j0, int lane) {
        for (int i = lane; i < 32 * RW; i += 32) {
            const int r = i / RW, col = i - r * RW;
            const long long e0 = (cbase + r) * L + j0;
            sm[r * P + col] = (e0 + col / W < N) ? X[e0 * W + col] : 0.0;
        }
    }
    __device__ void load_async(const real* X, long long N, long long cbase, int L, int j0, int lane) {
#pragma unroll 4
        for (int i = lane; i < 32 * RW; i += 32) {
            const int r = i / RW, col = i - r * RW;
            const long long e0 = (cbase + r) * L + j0;
            if (e0 + col / W < N) cp_async_8(sm + r * P + col, X + e0 * W + col);
        }
    }
    __device__ void store(real* X, long long N, long long cbase, int L, int j0, int lane) const {
        for (int i = lane; i < 32 * RW; i += 32) {
            const int r = i / RW, col = i - r * RW;
            const long long e0 = (cbase + r) * L + j0;
            if (e0 + col / W < N) X[e0 * W + col] = sm[r * P + col];
        }
    }
    __device__ real& at(int lane, int jj, int w) { return sm[lane * P + jj * W + w]; }
};

template <int D>
struct WarpCtx {
    static constexpr bool kAsync = (D == 1);
    static constexpr int kBuf = kAsync ? 2 : 1;
    static constexpr int kInDoubles = Tile<1>::kDoubles + Tile<D>::kDoubles + Tile<D * D>::kDoubles;
    static constexpr int kOutDoubles = Tile<D>::kDoubles + Tile<D * D>::kDoubles;
    static constexpr int kSmootherDoubles = kBuf * Tile<1>::kDoubles + kOutDoubles;
    static constexpr int kDoublesPerWarp = (kBuf * kInDoubles > kSmootherDoubles) ? kBuf * kInDoubles : kSmootherDoubles;

    UpIO io;
    real* base;
    Tile<1> tdt;
    Tile<D> ty;       // observations in (filter) / posterior means out (smoother)
    Tile<D * D> tR;   // site covariances in / posterior covariances out
    int lane, cur;
    bool smoother, primed;

    // filter layout: [buf0: dt | y | R][buf1: dt | y | R];  smoother layout: [dt buf0][dt buf1][out mean | out cov]
    __device__ WarpCtx(const UpIO& io_, real* smem_warp, bool smoother_) : io(io_), base(smem_warp) {
        lane = threadIdx.x & 31;
        cur = 0;
        smoother = smoother_;
        primed = false;
        point_in(0);
        if (smoother) {
            ty.sm = base + kBuf * Tile<1>::kDoubles;
            tR.sm = ty.sm + Tile<D>::kDoubles;
        }
    }
    __device__ void point_in(int b) {
        if (smoother) {
            tdt.sm = base + b * Tile<1>::kDoubles;
        } else {
            tdt.sm = base + b * kInDoubles;
            ty.sm = tdt.sm + Tile<1>::kDoubles;
            tR.sm = ty.sm + Tile<D>::kDoubles;
        }
    }
    __device__ void issue(int b, long long cbase, int L, int j0) {
        point_in(b);
        tdt.load_async(io.dt, io.N, cbase, L, j0, lane);
        if (!smoother) {
            ty.load_async(io.y, io.N, cbase, L, j0, lane);
            tR.load_async(io.R, io.N, cbase, L, j0, lane);
        }
        cp_async_commit();
    }
    // cbase: first chunk of this warp; dir: +1 when the sub-blocks are walked forward, -1 backward
    __device__ void begin(long long cbase, int L, int j0, int dir) {
        __syncwarp();  // every lane is done with the tile that is about to be overwritten
        if constexpr (kAsync) {
            if (!primed) {
                issue(cur, cbase, L, j0);
                primed = true;
            }
            const int jn = j0 + dir * kUpTJ;
            if (jn >= 0 && jn < L && cbase * L + jn < io.N) {
                issue(cur ^ 1, cbase, L, jn);
                cp_async_wait<1>();
            } else {
                cp_async_wait<0>();
            }
            point_in(cur);
            cur ^= 1;
        } else {
            tdt.load(io.dt, io.N, cbase, L, j0, lane);
            if (!smoother) {
                ty.load(io.y, io.N, cbase, L, j0, lane);
                tR.load(io.R, io.N, cbase, L, j0, lane);
            }
        }
        __syncwarp();
    }
    __device__ void end(long long cbase, int L, int j0) {
        __syncwarp();
        ty.store(io.post_mean, io.N, cbase, L, j0, lane);
        tR.store(io.post_cov, io.N, cbase, L, j0, lane);
    }
    __device__ real dt(long long, int jj) { return tdt.at(lane, jj, 0); }
    __device__ void obs(long long, int jj, real* y, real* R) {
#pragma unroll
        for (int i = 0; i < D; ++i) y[i] = ty.at(lane, jj, i);
#pragma unroll
        for (int i = 0; i < D * D; ++i) R[i] = tR.at(lane, jj, i);
    }
    __device__ void put(long long, int jj, const real* pm, const real* pc) {
#pragma unroll
        for (int i = 0; i < D; ++i) ty.at(lane, jj, i) = pm[i];
#pragma unroll
        for (int i = 0; i < D * D; ++i) tR.at(lane, jj, i) = pc[i];
    }
};
#endif

// DirectCtx spelling of the block hooks used by the shared chunk bodies
template <int D> BN_DEV void ctx_begin(DirectCtx<D>&, long long, int, int, int) {}
template <int D> BN_DEV void ctx_end(DirectCtx<D>&, long long, int, int) {}
#ifdef __CUDACC__
template <int D> __device__ __forceinline__ void ctx_begin(WarpCtx<D>& cx, long long cbase, int L, int j0, int dir) {
    cx.begin(cbase, L, j0, dir);
}
template <int D> __device__ __forceinline__ void ctx_end(WarpCtx<D>& cx, long long cbase, int L, int j0) {
    cx.end(cbase, L, j0);
}
#endif

// ------------------------------------------------------------------------------------------ chunk bodies
// All bodies are called by every lane of a warp (`active` = this lane owns a real chunk) so the
// staging hooks stay warp-uniform.
#pragma nv_exec_check_disable
template <class G, class Ctx>
BN_DEV void up_reduce_chunk(const G& g, Ctx& cx, long long N, int L, long long nchunks, int is_first, real* agg,
                            long long c, bool active) {
    constexpr int D = G::D;
    using Alg = FilterAlg<G::d>;
    typename Alg::Elem el;
    Alg::identity(el);
    const long long k0 = c * L, cbase = c & ~31LL;
    for (int j0 = 0; j0 < L; j0 += kUpTJ) {
        if (cbase * L + j0 >= N) break;  // warp-uniform: nothing left for any lane
        ctx_begin(cx, cbase, L, j0, 1);
        // the transition of step k+1 does not depend on the recursion: it is formed while step k runs,
        // so its exp() chain fills the dependency stalls of the update (slot kUpTJ of a tile row is padding)
        real Abn[G::kBlockA];
        g.trans(cx.dt(k0 + j0, 0), Abn);
#pragma unroll 1
        for (int jj = 0; jj < kUpTJ; ++jj) {
            const long long k = k0 + j0 + jj;
            if (active && k < N) {
                real y[D], R[D * D], Ab[G::kBlockA];
                cx.obs(k, jj, y, R);
#pragma unroll
                for (int i = 0; i < G::kBlockA; ++i) Ab[i] = Abn[i];
                g.trans(cx.dt(k + 1, jj + 1), Abn);
                fkf_absorb<G>(g, el, Ab, y, R, is_first && k == 0);
            }
        }
    }
    if (active) Alg::store(agg, nchunks, c, el);
}

#pragma nv_exec_check_disable
template <class G, bool WANT_ELL, class Ctx>
BN_DEV void up_filter_chunk(const G& g, Ctx& cx, long long N, int L, long long nchunks, int is_first,
                            const real* prefix, const real* s0, real* fs, real* ell_partials, long long c,
                            bool active) {
    constexpr int d = G::d, D = G::D;
    using Alg = FilterAlg<d>;
    typename Alg::State s;
    Alg::zero_state(s);
    if (active) {
        Alg::load_state(s0, 1, 0, s);
        if (c > 0) {
            typename Alg::Elem e;
            Alg::load(prefix, nchunks, c - 1, e);
            typename Alg::State t;
            Alg::apply(e, s, t);
            s = t;
        } else if (is_first) {  // global step 0 starts from the stationary prior (m0 = 0, P0 = Pinf)
            Alg::zero_state(s);
            g.pinf_full(s.P);
        }
    }
    real ell = 0.0;
    const long long k0 = c * L, cbase = c & ~31LL;
    for (int j0 = 0; j0 < L; j0 += kUpTJ) {
        if (cbase * L + j0 >= N) break;
        ctx_begin(cx, cbase, L, j0, 1);
        real Abn[G::kBlockA];  // transition of the next step, formed one step ahead (see up_reduce_chunk)
        g.trans(cx.dt(k0 + j0, 0), Abn);
#pragma unroll 1
        for (int jj = 0; jj < kUpTJ; ++jj) {
            const long long k = k0 + j0 + jj;
            if (active && k < N) {
                real y[D], R[D * D], Ab[G::kBlockA], mp[d], Pp[symn(d)];
                unsigned char mk[D];
                cx.obs(k, jj, y, R);
                if (cx.io.mask) {
#pragma unroll
                    for (int i = 0; i < D; ++i) mk[i] = cx.io.mask[k * D + i];
                }
#pragma unroll
                for (int i = 0; i < G::kBlockA; ++i) Ab[i] = Abn[i];
                g.trans(cx.dt(k + 1, jj + 1), Abn);
                ell += fkf_step<G, WANT_ELL>(g, s.m, s.P, Ab, y, R, cx.io.mask ? mk : nullptr, mp, Pp);
                fs_store<d>(fs, c, L, j0 + jj, s.m, s.P);
            }
        }
    }
    if (WANT_ELL && active) ell_partials[c] = ell;
}

// smoothing element of chunk c, stored at scan position nchunks-1-c (the smoother scans right to left)
template <class G>
BN_DEV void up_selem_chunk(long long N, int L, long long nchunks, int need_first, const real* agg,
                           const real* s0, const real* fs, real* selems, long long c) {
    constexpr int d = G::d;
    using FA = FilterAlg<d>;
    using SA = SmootherAlg<d>;
    typename SA::Elem se;
    if (c == 0 && !need_first) {
        SA::identity(se);
    } else {
        typename FA::Elem fe;
        FA::load(agg, nchunks, c, fe);
        real ma[d], Pa[symn(d)], mb[d], Pb[symn(d)];
        if (c == 0) {
            typename FA::State s;
            FA::load_state(s0, 1, 0, s);
#pragma unroll
            for (int i = 0; i < d; ++i) ma[i] = s.m[i];
#pragma unroll
            for (int i = 0; i < symn(d); ++i) Pa[i] = s.P[i];
        } else {
            fs_load<d>(fs, c - 1, L, L - 1, ma, Pa);
        }
        const long long kend = ((c + 1) * L < N) ? (c + 1) * L : N;
        fs_load<d>(fs, c, L, (int)(kend - c * L) - 1, mb, Pb);
        chunk_smoothing_element<d>(fe, ma, Pa, mb, Pb, se);
    }
    SA::store(selems, nchunks, nchunks - 1 - c, se);
}

// GRAD: the sweep also accumulates this chunk's share of d ell / d hyper-parameters (fast_core.cuh) for the
// transitions INTO each of its steps -- inside the chunk as part of frts_step, and for its first step from the
// last filtered state of the chunk on its left (the prior when it is global step 0, s0 on a later shard) --
// and leaves the GradAcc fields in gpart[field * nchunks + c].
#pragma nv_exec_check_disable
template <class G, bool GRAD, class Ctx>
BN_DEV void up_smooth_chunk(const G& g, Ctx& cx, long long N, int L, long long nchunks, const real* sprefix,
                            const real* sinit, const real* fs, long long c, bool active, int is_first = 0,
                            const real* s0 = nullptr, real* gpart = nullptr) {
    constexpr int d = G::d, D = G::D;
    using Alg = SmootherAlg<d>;
    typename Alg::State s;
    Alg::zero_state(s);
    GradAcc<G> acc;
    real h_next = 0.0;  // length of the step out of the state being processed (GRAD only)
    if constexpr (GRAD) acc.zero();
    const long long p = nchunks - 1 - c;
    if (active) {
        Alg::load_state(sinit, 1, 0, s);
        if (p > 0) {
            typename Alg::Elem e;
            Alg::load(sprefix, nchunks, p - 1, e);
            typename Alg::State t;
            Alg::apply(e, s, t);
            s = t;
        }
    }
    const long long k0 = c * L, cbase = c & ~31LL;
    const long long kend = ((c + 1) * L < N) ? (c + 1) * L : N;
    const int j_last = (int)(kend - k0) - 1;  // s is the smoothed state of this step
    real nfm[d], nfP[symn(d)];              // filtered state of the next step to process, loaded one step ahead
    real Abn[G::kBlockA], Qbn[G::kBlockS];  // discretisation of the next step to process, formed one step ahead
#pragma unroll
    for (int i = 0; i < G::kBlockA; ++i) Abn[i] = 0.0;
#pragma unroll
    for (int i = 0; i < G::kBlockS; ++i) Qbn[i] = 0.0;
    if (active && j_last >= 1) fs_load<d>(fs, c, L, j_last - 1, nfm, nfP);
    for (int j0 = L - kUpTJ; j0 >= 0; j0 -= kUpTJ) {
        if (cbase * L + j0 >= N) continue;
        ctx_begin(cx, cbase, L, j0, -1);
#pragma unroll 1
        for (int jj = kUpTJ - 1; jj >= 0; --jj) {
            const int j = j0 + jj;
            const long long k = k0 + j;
            if (active && j <= j_last) {
                const real h_k = cx.dt(k, jj);
                real Ab[G::kBlockA], Qb[G::kBlockS];
#pragma unroll
                for (int i = 0; i < G::kBlockA; ++i) Ab[i] = Abn[i];
#pragma unroll
                for (int i = 0; i < G::kBlockS; ++i) Qb[i] = Qbn[i];
                // the discretisation of the step below (k-1 -> k, length h_k) is formed while this step's
                // dependent chain runs
                const real h_out = h_next;
                h_next = h_k;
                g.trans(h_k, Abn);
                g.noise(Abn, Qbn);
                if (j < j_last) {
                    real fm[d], fP[symn(d)];
#pragma unroll
                    for (int i = 0; i < d; ++i) fm[i] = nfm[i];
#pragma unroll
                    for (int i = 0; i < symn(d); ++i) fP[i] = nfP[i];
                    if (j >= 1) fs_load<d>(fs, c, L, j - 1, nfm, nfP);  // in flight during this step's arithmetic
                    if constexpr (GRAD) frts_step<G, true>(Ab, Qb, fm, fP, s.m, s.P, &g, h_out, &acc);
                    else frts_step<G>(Ab, Qb, fm, fP, s.m, s.P);
                }
                real pm[D], pc[D * D];
#pragma unroll
                for (int a = 0; a < D; ++a) {
                    pm[a] = s.m[G::sel(a)];
#pragma unroll
                    for (int b = 0; b < D; ++b) pc[a * D + b] = s.P[sidx(G::sel(a), G::sel(b))];
                }
                cx.put(k, jj, pm, pc);
            }
        }
        ctx_end(cx, cbase, L, j0);
    }
    if constexpr (GRAD) {
        if (active) {
            // transition into the chunk's first step k0: s is its smoothed state, (Abn, Qbn) its discretisation, h_next its length
            real pm_[d], pP_[symn(d)], mp[d], Pp[symn(d)], dm[d], dP[symn(d)];
            const bool prior = (c == 0 && is_first);
            if (prior) {
#pragma unroll
                for (int i = 0; i < d; ++i) pm_[i] = mp[i] = 0.0;
                g.pinf_full(pP_);
#pragma unroll
                for (int i = 0; i < symn(d); ++i) Pp[i] = pP_[i];
            } else {
                if (c == 0) {
                    typename FilterAlg<d>::State t;
                    FilterAlg<d>::load_state(s0, 1, 0, t);
#pragma unroll
                    for (int i = 0; i < d; ++i) pm_[i] = t.m[i];
#pragma unroll
                    for (int i = 0; i < symn(d); ++i) pP_[i] = t.P[i];
                } else {
                    fs_load<d>(fs, c - 1, L, L - 1, pm_, pP_);
                }
                real X[d * d];
                bd_matvec<G>(Abn, pm_, mp);
                bd_mat_sym<G>(Abn, pP_, X);
                bd_abt_sym<G>(X, Abn, Qbn, Pp);
            }
#pragma unroll
            for (int i = 0; i < d; ++i) dm[i] = s.m[i] - mp[i];
#pragma unroll
            for (int i = 0; i < symn(d); ++i) dP[i] = s.P[i] - Pp[i];
            ldlt<d>(Pp);
            grad_accumulate<G>(g, Abn, h_next, Pp, dm, dP, pm_, pP_, prior, acc);
#pragma unroll
            for (int f = 0; f < G::NC * symn(G::n); ++f) gpart[f * nchunks + c] = acc.Gam[f];
#pragma unroll
            for (int f = 0; f < G::NC; ++f) gpart[(G::NC * symn(G::n) + f) * nchunks + c] = acc.gl[f];
        }
    }
}

// state of the last step (tiled scratch -> flat state buffer): the smoother's start on the last rank
template <int d>
BN_DEV void up_last_state(long long N, int L, const real* fs, real* sinit) {
    const long long c = (N - 1) / L;
    typename SmootherAlg<d>::State s;
    fs_load<d>(fs, c, L, (int)(N - 1 - c * L), s.m, s.P);
    SmootherAlg<d>::store_state(sinit, 1, 0, s);
}

// carry of this rank for the smoother exchange: composition of all its chunk elements, closed on
// the last rank by the terminal element (0, fm_N-1, fP_N-1) of ops.py:314-315
template <int d>
BN_DEV void up_export_scarry(const real* top_prefix, long long n_top, int is_last, long long N, int L,
                             const real* fs, real* carry) {
    using SA = SmootherAlg<d>;
    typename SA::Elem tot;
    SA::load(top_prefix, n_top, n_top - 1, tot);
    if (is_last) {
        typename SA::Elem term, r;
#pragma unroll
        for (int i = 0; i < d * d; ++i) term.E[i] = 0.0;
        const long long c = (N - 1) / L;
        fs_load<d>(fs, c, L, (int)(N - 1 - c * L), term.g, term.L);
        SA::combine(term, tot, r);
        tot = r;
    }
    SA::to_carry(tot, carry);
}

#ifdef __CUDACC__
// ------------------------------------------------------------------------------------------ kernels
template <class G>
__global__ void __launch_bounds__(kUpThreads, (G::d <= 3 ? kUpBlocksPerSM : 1))
up_reduce_kernel(G g, UpIO io, int L, long long nchunks, int is_first, real* agg) {
    extern __shared__ real up_smem[];
    WarpCtx<G::D> cx(io, up_smem + (threadIdx.x >> 5) * WarpCtx<G::D>::kDoublesPerWarp, false);
    const long long c = (long long)blockIdx.x * kUpThreads + threadIdx.x;
    up_reduce_chunk(g, cx, io.N, L, nchunks, is_first, agg, c, c < nchunks);
}

template <class G, bool WANT_ELL>
__global__ void __launch_bounds__(kUpThreads, (G::d <= 3 ? kUpBlocksPerSM : 1))
up_filter_kernel(G g, UpIO io, int L, long long nchunks, int is_first, const real* prefix, const real* s0,
                 real* fs, real* ell_partials) {
    extern __shared__ real up_smem[];
    WarpCtx<G::D> cx(io, up_smem + (threadIdx.x >> 5) * WarpCtx<G::D>::kDoublesPerWarp, false);
    const long long c = (long long)blockIdx.x * kUpThreads + threadIdx.x;
    up_filter_chunk<G, WANT_ELL>(g, cx, io.N, L, nchunks, is_first, prefix, s0, fs, ell_partials, c, c < nchunks);
}

template <class G>
__global__ void __launch_bounds__(128)
up_selem_kernel(long long N, int L, long long nchunks, int need_first, const real* agg, const real* s0,
                const real* fs, real* selems) {
    const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c < nchunks) up_selem_chunk<G>(N, L, nchunks, need_first, agg, s0, fs, selems, c);
}

template <class G>
__global__ void __launch_bounds__(kUpThreads, (G::d <= 3 ? kUpBlocksPerSM : 1))
up_smooth_kernel(G g, UpIO io, int L, long long nchunks, const real* sprefix, const real* sinit,
                 const real* fs) {
    extern __shared__ real up_smem[];
    WarpCtx<G::D> cx(io, up_smem + (threadIdx.x >> 5) * WarpCtx<G::D>::kDoublesPerWarp, true);
    const long long c = (long long)blockIdx.x * kUpThreads + threadIdx.x;
    up_smooth_chunk<G, false>(g, cx, io.N, L, nchunks, sprefix, sinit, fs, c, c < nchunks);
}

// the same sweep with the hyper-gradient accumulation (more live registers: its own occupancy bound)
template <class G>
__global__ void __launch_bounds__(kUpThreads, (G::d <= 3 ? kUpGradBlocksPerSM : 1))
up_smooth_grad_kernel(G g, UpIO io, int L, long long nchunks, const real* sprefix, const real* sinit,
                      const real* fs, int is_first, const real* s0, real* gpart) {
    extern __shared__ real up_smem[];
    WarpCtx<G::D> cx(io, up_smem + (threadIdx.x >> 5) * WarpCtx<G::D>::kDoublesPerWarp, true);
    const long long c = (long long)blockIdx.x * kUpThreads + threadIdx.x;
    up_smooth_chunk<G, true>(g, cx, io.N, L, nchunks, sprefix, sinit, fs, c, c < nchunks, is_first, s0, gpart);
}

// deterministic sums of the per-chunk GradAcc fields (one block per field), then the chain to (variance, lengthscale)
static __global__ void __launch_bounds__(1024) up_grad_sum_kernel(const real* gpart, long long nchunks, real* fields) {
    __shared__ real sh[1024];
    const real* x = gpart + (long long)blockIdx.x * nchunks;
    real s = 0.0;
    for (long long i = threadIdx.x; i < nchunks; i += 1024) s += x[i];
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int off = 512; off > 0; off >>= 1) {
        if ((int)threadIdx.x < off) sh[threadIdx.x] += sh[threadIdx.x + off];
        __syncthreads();
    }
    if (threadIdx.x == 0) fields[blockIdx.x] = sh[0];
}
template <class G>
__global__ void up_grad_finish_kernel(bn_kernel_spec spec, const real* fields, real* dvar, real* dlen) {
    if (blockIdx.x == 0 && threadIdx.x == 0) grad_finish<G>(spec, fields, dvar, dlen);
}

template <int d>
__global__ void up_last_state_kernel(long long N, int L, const real* fs, real* sinit) {
    if (blockIdx.x == 0 && threadIdx.x == 0) up_last_state<d>(N, L, fs, sinit);
}

template <int d>
__global__ void up_export_scarry_kernel(const real* top_prefix, long long n_top, int is_last, long long N, int L,
                                        const real* fs, real* carry) {
    if (blockIdx.x == 0 && threadIdx.x == 0) up_export_scarry<d>(top_prefix, n_top, is_last, N, L, fs, carry);
}

// ------------------------------------------------------------------------------------------ host driver
enum { UP_ALL = 0, UP_REDUCE = 1, UP_FILTER = 2, UP_SMOOTH = 3 };

constexpr int kUpMaxGradFields = 16;  // >= NC * (symn(n) + 1) for every instantiated stack

struct UpWs {
    real *s0, *sinit, *partials, *gpart, *gfields, *fs;
    ScanPlan fplan, splan;
};

template <int d>
inline size_t up_ws_doubles(long long N) {
    // sized for the larger of the two chunk plans (with / without the hyper-gradient accumulation)
    size_t need = 0;
    for (int grad = 0; grad < 2; ++grad) {
        ChunkPlan cp = up_plan_chunks(N > 0 ? N : 1, grad != 0, d);
        size_t n = 128 + 64 + scan_plan_doubles(cp.nchunks, FilterAlg<d>::kElem) +
                   scan_plan_doubles(cp.nchunks, SmootherAlg<d>::kElem) + cp.nchunks +
                   (size_t)cp.nchunks * kUpMaxGradFields + fs_doubles(cp.nchunks, cp.L, d + symn(d));
        if (n > need) need = n;
    }
    return need;
}

template <int d>
inline UpWs up_ws(void* ws, const ChunkPlan& cp) {
    UpWs w;
    real* p = (real*)ws;
    w.s0 = p; p += 64;
    w.sinit = p; p += 64;
    w.fplan = make_scan_plan(p, cp.nchunks, FilterAlg<d>::kElem);
    p += scan_plan_doubles(cp.nchunks, FilterAlg<d>::kElem);
    w.splan = make_scan_plan(p, cp.nchunks, SmootherAlg<d>::kElem);
    p += scan_plan_doubles(cp.nchunks, SmootherAlg<d>::kElem);
    w.partials = p; p += cp.nchunks;
    w.gfields = p; p += 64;
    w.gpart = p; p += (size_t)cp.nchunks * kUpMaxGradFields;
    w.fs = p;
    return w;
}

struct UpCall {
    const bn_kernel_spec* spec;
    UpIO io;
    real* ell;
    void* ws;
    size_t ws_bytes;
    cudaStream_t st;
    int phase, rank, world;
    real* carry_out;       // UP_REDUCE: filter carry; UP_FILTER: smoother carry
    const real* carries;   // UP_FILTER: filter carries [world]; UP_SMOOTH: smoother carries [world]
    int grad;                // 1: hyper-gradient plan; every phase of one update must agree on it
    real* dvar;            // UP_ALL / UP_SMOOTH with grad: d ell / d variance[NC] (this shard's share)
    real* dlen;            //                               d ell / d lengthscale[NC]
};

template <class G>
inline int up_run(const UpCall& c) {
    constexpr int d = G::d;
    using FA = FilterAlg<d>;
    using SA = SmootherAlg<d>;
    cudaStream_t st = c.st;
    const UpIO& io = c.io;
    if (io.N == 0) {
        if (c.ell) BN_CUDA(cudaMemsetAsync(c.ell, 0, sizeof(real), st));
        return 0;
    }
    G g;
    g.prepare(*c.spec);
    static_assert(GradAcc<G>::kFields <= kUpMaxGradFields, "raise kUpMaxGradFields");
    ChunkPlan cp = up_plan_chunks(io.N, c.grad != 0, d);
    size_t need = up_ws_doubles<d>(io.N) * sizeof(real);
    BN_REQUIRE(c.ws != nullptr && c.ws_bytes >= need, "workspace too small: need %zu bytes, got %zu", need, c.ws_bytes);
    UpWs w = up_ws<d>(c.ws, cp);
    const unsigned grid = (unsigned)((cp.nchunks + kUpThreads - 1) / kUpThreads);
    const size_t smem = (size_t)kUpWarps * WarpCtx<G::D>::kDoublesPerWarp * sizeof(real);
    const int is_first = (c.rank == 0), is_last = (c.rank == c.world - 1);
    const bool sharded = c.phase != UP_ALL;
    if (smem > 48 * 1024) {  // multi-latent tiles exceed the default dynamic shared-memory window
        BN_CUDA(cudaFuncSetAttribute(up_reduce_kernel<G>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        BN_CUDA(cudaFuncSetAttribute(up_filter_kernel<G, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        BN_CUDA(cudaFuncSetAttribute(up_filter_kernel<G, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        BN_CUDA(cudaFuncSetAttribute(up_smooth_kernel<G>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        BN_CUDA(cudaFuncSetAttribute(up_smooth_grad_kernel<G>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }

    if (c.phase == UP_ALL || c.phase == UP_REDUCE) {
        BN_LAUNCH("up_reduce", st,
                  (up_reduce_kernel<G><<<grid, kUpThreads, smem, st>>>(g, io, cp.L, cp.nchunks, is_first, w.fplan.input0)));
        BN_CUDA(cudaGetLastError());
        BN_CUDA(run_scan<FA>(w.fplan, st));
        if (c.carry_out && c.phase == UP_REDUCE) {
            int top = w.fplan.levels - 1;
            export_carry_kernel<FA><<<1, 1, 0, st>>>(w.fplan.prefix[top], w.fplan.count[top], c.carry_out);
            BN_CUDA(cudaGetLastError());
        }
    }
    if (c.phase == UP_ALL || c.phase == UP_FILTER) {
        if (sharded) {
            fold_carries_kernel<FA><<<1, 1, 0, st>>>(c.carries, 0, c.rank, 1, w.s0);
            BN_CUDA(cudaGetLastError());
        } else {
            BN_CUDA(cudaMemsetAsync(w.s0, 0, FA::kState * sizeof(real), st));
        }
        if (c.ell) {
            BN_LAUNCH("up_filter", st,
                      (up_filter_kernel<G, true><<<grid, kUpThreads, smem, st>>>(
                          g, io, cp.L, cp.nchunks, is_first, w.fplan.prefix[0], w.s0, w.fs, w.partials)));
            BN_CUDA(cudaGetLastError());
            BN_LAUNCH("sum", st, (sum_kernel<false><<<1, 1024, 0, st>>>(w.partials, cp.nchunks, c.ell, 1.0)));
        } else {
            BN_LAUNCH("up_filter", st,
                      (up_filter_kernel<G, false><<<grid, kUpThreads, smem, st>>>(
                          g, io, cp.L, cp.nchunks, is_first, w.fplan.prefix[0], w.s0, w.fs, nullptr)));
        }
        BN_CUDA(cudaGetLastError());
        const unsigned g2 = (unsigned)((cp.nchunks + 127) / 128);
        BN_LAUNCH("up_selem", st,
                  (up_selem_kernel<G><<<g2, 128, 0, st>>>(io.N, cp.L, cp.nchunks, !is_first, w.fplan.input0, w.s0,
                                                          w.fs, w.splan.input0)));
        BN_CUDA(cudaGetLastError());
        BN_CUDA(run_scan<SA>(w.splan, st));
        if (c.carry_out && c.phase == UP_FILTER) {
            int top = w.splan.levels - 1;
            up_export_scarry_kernel<d><<<1, 1, 0, st>>>(w.splan.prefix[top], w.splan.count[top], is_last, io.N, cp.L,
                                                        w.fs, c.carry_out);
            BN_CUDA(cudaGetLastError());
        }
    }
    if (c.phase == UP_ALL || c.phase == UP_SMOOTH) {
        if (sharded && !is_last) {
            fold_carries_kernel<SA><<<1, 1, 0, st>>>(c.carries, c.world - 1, c.rank, -1, w.sinit);
        } else {
            up_last_state_kernel<d><<<1, 1, 0, st>>>(io.N, cp.L, w.fs, w.sinit);
        }
        BN_CUDA(cudaGetLastError());
        if (c.grad && c.dvar) {
            BN_LAUNCH("up_smooth_grad", st,
                      (up_smooth_grad_kernel<G><<<grid, kUpThreads, smem, st>>>(
                          g, io, cp.L, cp.nchunks, w.splan.prefix[0], w.sinit, w.fs, is_first, w.s0, w.gpart)));
            BN_CUDA(cudaGetLastError());
            BN_LAUNCH("up_grad_sum", st,
                      (up_grad_sum_kernel<<<GradAcc<G>::kFields, 1024, 0, st>>>(w.gpart, cp.nchunks, w.gfields)));
            up_grad_finish_kernel<G><<<1, 1, 0, st>>>(*c.spec, w.gfields, c.dvar, c.dlen);
        } else {
            BN_LAUNCH("up_smooth", st,
                      (up_smooth_kernel<G><<<grid, kUpThreads, smem, st>>>(g, io, cp.L, cp.nchunks, w.splan.prefix[0],
                                                                           w.sinit, w.fs)));
        }
        BN_CUDA(cudaGetLastError());
    }
    return 0;
}

#define BN_UP_SPEC_CASE(FAM, NC)                                           \
    if (c.spec->family == FAM && c.spec->n_components == NC) return up_run<FastGen<FAM, NC>>(c);
#endif  // __CUDACC__

}  // namespace BN_NS
