// Fused inference iteration on chunk-tiled resident state (single latent, one site per step).
//
// One iteration of a temporal Markov GP (inference.py:65-90 followed by energy(), inference.py:197-222) is
//     update_posterior (F, S)  ->  site update (U)  ->  update_posterior (F, S)  ->  energy terms (V, X)
// and every array that only travels between those stages (dt, the observations, the sites, the filtered states, the
// posterior marginals) is private to the path.  Here they live in ONE layout, chosen for the thread that consumes
// them: a step series x[0..N) is cut into the chunks of the temporally parallel filter (one thread per chunk of L
// steps, 32 consecutive chunks per warp) and stored
//     x_t[((c >> 5) * L + j) * 32 + (c & 31)]        (chunk c, step j of the chunk)
// so the lane that owns chunk c reads / writes step j of all 32 chunks of its warp as ONE 256-byte transaction: no
// shared-memory staging, no transposition, no index arithmetic beyond a pointer bump.  The two smoother sweeps then
// run the per-step site work in their epilogue, on the smoothed marginal that is still in registers:
//     sweep 1 (+ BN_ITER_SITES):  variational_expectation / log_likelihood_gradients, ensure_psd, newton_update, damping,
//                                 reparametrise (inference.py:72-86, 105-128, 170-195) -> new sites, in place
//     sweep 2 (+ BN_ITER_ENERGY): E_q[log p(y|f)] and E_q[log N(pseudo_y | f, pseudo_var)] (inference.py:208-218,
//                                 utils.py:510-531) summed per chunk, posterior marginals written once
// so (H sm, H sP H^T) never round-trip through HBM before the cubature, and the shared-memory-bound table gathers of
// the probit cubature overlap the fp64-bound RTS recursion of the other warps on the same SM.
// HBM traffic per step of an iteration (d = 3): 2 x (24 reduce + 24 + 72 filter + 72 + 8 smoother) + 16 + 16 (U: y, old
// sites in; sites out) + 8 + 16 + 16 (V, X in; marginals out) = 472 B against 636 B on the reference-interface layouts.
#pragma once
#include "up_impl.cuh"
#include "site_math.cuh"

namespace BN_NS {

#ifndef BN_IT_SMOOTH_UNROLL
#define BN_IT_SMOOTH_UNROLL 1
#endif
constexpr int kItSmoothUnroll = BN_IT_SMOOTH_UNROLL;  // steps of the RTS recursion unrolled together

// ------------------------------------------------------------------------------------------ tiled layout
constexpr int kTlPadRows = 2;  // rows past the last step a prefetch may touch
BN_DEV long long tl_base(long long c, int L) { return (((c >> 5) * (long long)L) << 5) + (c & 31); }
inline long long tl_len(long long nchunks, int L) { return ((nchunks + 31) / 32) * 32LL * L + 32LL * kTlPadRows; }

struct ItIO {
    long long N;
    const real* dt;           // tiled
    const real* y;            // tiled observations (the data; site pass / energy pass only)
    real* sy;                 // tiled pseudo observations (site means)
    real* sR;                 // tiled pseudo variances (site covariances)
    const unsigned char* mask;  // tiled, 1 = the pseudo observation of this step is missing; nullable
    real* pm;                 // tiled posterior marginal means (plain / energy pass)
    real* pc;                 // tiled posterior marginal variances
    real* pm_lin;             // nullable: the marginals go to these [N] arrays in time order instead (the layout the
    real* pc_lin;             // reference's posterior_mean / posterior_variance have); one 8-byte store per lane and step
};

// ------------------------------------------------------------------------------------------ chunk bodies
// phase 1: fold the chunk's steps into one filtering element (ops.py:183-219)
template <class G>
BN_DEV void it_reduce_chunk(const G& g, const ItIO& io, int L, long long nchunks, int is_first, real* agg, long long c) {
    static_assert(G::D == 1, "the tiled path carries one site per step");
    using Alg = FilterAlg<G::d>;
    typename Alg::Elem el;
    Alg::identity(el);
    const long long k0 = c * L, rem = io.N - k0;
    const int cnt = rem < L ? (int)rem : L;
    const long long b = tl_base(c, L);
    const real* pdt = io.dt + b;
    const real* py = io.sy + b;
    const real* pR = io.sR + b;
    real Abn[G::kBlockA];
    g.trans(pdt[0], Abn);
    real yn = py[0], Rn = pR[0], hn = pdt[32];
#pragma unroll 1
    for (int j = 0; j < cnt; ++j) {
        real y[1] = {yn}, R[1] = {Rn}, Ab[G::kBlockA];
#pragma unroll
        for (int i = 0; i < G::kBlockA; ++i) Ab[i] = Abn[i];
        const real h1 = hn;
        // the next step's inputs are in flight while this step's dependent chain runs (rows past the chunk are padding)
        py += 32; pR += 32; pdt += 32;
        yn = py[0]; Rn = pR[0]; hn = pdt[32];
        g.trans(h1, Abn);
        fkf_absorb<G>(g, el, Ab, y, R, is_first && k0 + j == 0);
    }
    Alg::store(agg, nchunks, c, el);
}

// phase 1 with speculation.  The sensitivity A of a chunk aggregate to the state entering the chunk decays
// geometrically (the filter forgets); once every |A_ij| (in units of the stationary standard deviations) is below 2^-100
// the incoming state can no longer change anything at fp64 resolution: J and eta stop moving, and (b, C) IS the
// filtered state of every later step whatever the chunk started from.  From there on the pass runs the plain filter
// step on (b, C), writes the filtered states and sums the log-likelihood itself, and phase 3 only has to redo the
// steps before the switch.  For chunks much longer than the forgetting time (N = 1e8: 1320 steps against ~90) that
// removes most of the second pass over the inputs; for short chunks the switch never happens and phase 3 does what it
// always did.  The switch is taken by all 32 lanes of a warp together (the tiled layout needs them on the same step).
// "Forgotten" = every entry of the sensitivity, scaled by the stationary standard deviations of the states it links, is
// below 2^-56: what the incoming state could still add to the filtered state is under an eighth of an fp64 ulp of the
// state's own scale (measured: the energy of C5 agrees to 16 digits with the 2^-100 the path started with; each factor
// 2^-8 shortens the wait by about 12 steps at C5's lengthscale).  BN_B200_SPEC_LOG2 / BN_B200_SPEC_MIN_CHUNK override.
// (fp32 build: 2^-30, the same eighth of an ulp in float)
constexpr real kSpecThreshold = kReal32 ? 9.313225746154785e-10 : 1.3877787807814457e-17;  // 2^-30 : 2^-56

template <class G, bool WANT_ELL>
struct SpecReduce {
    static constexpr int d = G::d, nf = G::d + symn(G::d);
    using Alg = FilterAlg<G::d>;
    typename Alg::Elem el;
    const real *pdt, *py, *pR;
    const unsigned char* pk;
    real* pf;
    real Abn[G::kBlockA], yn, Rn, hn, ell, isd[G::d], sd[G::d];
    int cnt;

    BN_DEV void init(const G& g, const ItIO& io, int L, real* fs, long long c, bool active, real thr = kSpecThreshold) {
        Alg::identity(el);
        ell = 0.0;
        const long long k0 = c * L, rem = io.N - k0;
        cnt = active ? (rem < L ? (int)rem : L) : 0;
        const long long b = tl_base(c, L);
        pdt = io.dt + b;
        py = io.sy + b;
        pR = io.sR + b;
        pk = io.mask ? io.mask + b : nullptr;
        pf = fs + fs_index(c, L, 0, 0, nf);
        real P[symn(d)];
        g.pinf_full(P);
#pragma unroll
        for (int i = 0; i < d; ++i) {
            sd[i] = sqrt(P[sidx(i, i)]);
            isd[i] = thr / sd[i];
        }
        if (active) {
            g.trans(pdt[0], Abn);
            yn = py[0]; Rn = pR[0]; hn = pdt[32];
        } else {
#pragma unroll
            for (int i = 0; i < G::kBlockA; ++i) Abn[i] = 0.0;
            yn = Rn = hn = 0.0;
        }
    }
    BN_DEV void advance(const G& g, real* y, real* R, real* Ab) {
        y[0] = yn; R[0] = Rn;
#pragma unroll
        for (int i = 0; i < G::kBlockA; ++i) Ab[i] = Abn[i];
        const real h1 = hn;
        py += 32; pR += 32; pdt += 32;
        yn = py[0]; Rn = pR[0]; hn = pdt[32];
        g.trans(h1, Abn);
    }
    // folds one step into the aggregate; true when the aggregate no longer depends on the incoming state
    BN_DEV bool absorb(const G& g, bool first) {
        real y[1], R[1], Ab[G::kBlockA];
        advance(g, y, R, Ab);
        fkf_absorb<G>(g, el, Ab, y, R, first);
        if (pk) pk += 32;
        pf += nf * 32;
        bool dec = true;
#pragma unroll
        for (int i = 0; i < d; ++i)
#pragma unroll
            for (int j = 0; j < d; ++j) dec = dec && (fabs(el.A[i * d + j]) * sd[j] <= isd[i] * (sd[i] * sd[i]));
        return dec;
    }
    // (b, C) of the aggregate after the filter phase (same SoA layout as FilterAlg::store)
    BN_DEV void store_state(real* agg, long long stride, long long c) const {
        real* p = agg + c + (long long)(d * d) * stride;
#pragma unroll
        for (int k = 0; k < d; ++k) p[k * stride] = el.b[k];
#pragma unroll
        for (int k = 0; k < symn(d); ++k) p[(d + k) * stride] = el.C[k];
    }
    // one plain filter step on (b, C): the filtered state of this step, stored; log-likelihood increment
    BN_DEV void filter(const G& g) {
        real y[1], R[1], Ab[G::kBlockA], mp[d], Pp[symn(d)];
        unsigned char mk[1] = {0};
        if (pk) { mk[0] = pk[0]; pk += 32; }
        advance(g, y, R, Ab);
        ell += fkf_step<G, WANT_ELL>(g, el.b, el.C, Ab, y, R, pk ? mk : nullptr, mp, Pp);
#pragma unroll
        for (int f = 0; f < d; ++f) pf[f * 32] = el.b[f];
#pragma unroll
        for (int f = 0; f < symn(d); ++f) pf[(d + f) * 32] = el.C[f];
        pf += nf * 32;
    }
};

// phase 3: plain filter from the chunk's incoming state; filtered states -> scratch, log-likelihood partial
template <class G, bool WANT_ELL>
BN_DEV void it_filter_chunk(const G& g, const ItIO& io, int L, long long nchunks, int is_first, const real* prefix,
                            const real* s0, real* fs, real* ell_partials, long long c, const int* jst = nullptr,
                            const PrefixParts& up = PrefixParts()) {
    constexpr int d = G::d;
    using Alg = FilterAlg<d>;
    typename Alg::State s;
    Alg::load_state(s0, 1, 0, s);
    if (c > 0) {
        apply_prefix<Alg>(prefix, nchunks, up, c - 1, s);
    } else if (is_first) {  // global step 0 starts from the stationary prior (m0 = 0, P0 = Pinf)
        Alg::zero_state(s);
        g.pinf_full(s.P);
    }
    real ell = 0.0;
    const long long k0 = c * L, rem = io.N - k0;
    // with a speculative phase 1 only the steps before its switch are left (their states depend on the incoming one)
    const int cnt = jst ? jst[c] : (rem < L ? (int)rem : L);
    const long long b = tl_base(c, L);
    const real* pdt = io.dt + b;
    const real* py = io.sy + b;
    const real* pR = io.sR + b;
    const unsigned char* pk = io.mask ? io.mask + b : nullptr;
    real* pf = fs + fs_index(c, L, 0, 0, d + symn(d));
    real Abn[G::kBlockA];
    g.trans(pdt[0], Abn);
    real yn = py[0], Rn = pR[0], hn = pdt[32];
#pragma unroll 1
    for (int j = 0; j < cnt; ++j) {
        real y[1] = {yn}, R[1] = {Rn}, Ab[G::kBlockA], mp[d], Pp[symn(d)];
        unsigned char mk[1] = {0};
        if (pk) mk[0] = pk[(long long)j * 32];
#pragma unroll
        for (int i = 0; i < G::kBlockA; ++i) Ab[i] = Abn[i];
        const real h1 = hn;
        py += 32; pR += 32; pdt += 32;
        yn = py[0]; Rn = pR[0]; hn = pdt[32];
        g.trans(h1, Abn);
        ell += fkf_step<G, WANT_ELL>(g, s.m, s.P, Ab, y, R, pk ? mk : nullptr, mp, Pp);
#pragma unroll
        for (int f = 0; f < d; ++f) pf[f * 32] = s.m[f];
#pragma unroll
        for (int f = 0; f < symn(d); ++f) pf[(d + f) * 32] = s.P[f];
        pf += (d + symn(d)) * 32;
    }
    if (WANT_ELL) ell_partials[c] = ell;
}

// RTS recursion down the chunk (ops.py:290-301); the epilogue takes the smoothed marginal of every step
template <class G, class Epi>
BN_DEV void it_smooth_chunk(const G& g, const ItIO& io, int L, long long nchunks, const real* sprefix,
                            const real* sinit, const real* fs, long long c, Epi& epi, const PrefixParts& sup = PrefixParts()) {
    constexpr int d = G::d, nf = d + symn(d);
    using Alg = SmootherAlg<d>;
    typename Alg::State s;
    const long long p = nchunks - 1 - c;
    Alg::load_state(sinit, 1, 0, s);
    if (p > 0) apply_prefix<Alg>(sprefix, nchunks, sup, p - 1, s);
    const long long k0 = c * L, rem = io.N - k0;
    const int j_last = (rem < L ? (int)rem : L) - 1;  // s is the smoothed state of this step
    const long long b = tl_base(c, L);
    const real* pdt = io.dt + b + (long long)j_last * 32;
    const real* pf = fs + fs_index(c, L, j_last, 0, nf);
    long long ti = b + (long long)j_last * 32;
    real nfm[d], nfP[symn(d)];  // filtered state of the next step to process, loaded one step ahead
    real Abn[G::kBlockA];       // transition of the next step to process, formed one step ahead (its exp() chain is
                                  // independent of the recursion); the noise blocks are formed at their use: keeping
                                  // them a step ahead as well costs 12 registers the fused epilogues need
#pragma unroll
    for (int i = 0; i < G::kBlockA; ++i) Abn[i] = 0.0;
    real hn = pdt[0];
    if (j_last >= 1) {
#pragma unroll
        for (int f = 0; f < d; ++f) nfm[f] = pf[(f - nf) * 32];
#pragma unroll
        for (int f = 0; f < symn(d); ++f) nfP[f] = pf[(d + f - nf) * 32];
    }
#pragma unroll kItSmoothUnroll
    for (int j = j_last; j >= 0; --j) {
        real Ab[G::kBlockA];
#pragma unroll
        for (int i = 0; i < G::kBlockA; ++i) Ab[i] = Abn[i];
        const real h_k = hn;
        pdt -= 32;
        if (j >= 1) hn = pdt[0];
        epi.prefetch(ti);
        // the transition of the step below (k-1 -> k, length h_k) is formed while this step's dependent chain runs
        g.trans(h_k, Abn);
        if (j < j_last) {
            real fm[d], fP[symn(d)], Qb[G::kBlockS];
#pragma unroll
            for (int i = 0; i < d; ++i) fm[i] = nfm[i];
#pragma unroll
            for (int i = 0; i < symn(d); ++i) fP[i] = nfP[i];
            pf -= nf * 32;
            if (j >= 1) {  // in flight during this step's arithmetic
#pragma unroll
                for (int f = 0; f < d; ++f) nfm[f] = pf[(f - nf) * 32];
#pragma unroll
                for (int f = 0; f < symn(d); ++f) nfP[f] = pf[(d + f - nf) * 32];
            }
            g.noise(Ab, Qb);
            frts_step<G>(Ab, Qb, fm, fP, s.m, s.P);
        }
        epi.step(ti, k0 + j, s.m[G::sel(0)], s.P[sidx(G::sel(0), G::sel(0))]);
        ti -= 32;
    }
    epi.finish(c);
}

#ifndef BN_IT_TAB_UNROLL
#define BN_IT_TAB_UNROLL 4  // table evaluations in flight per thread inside the fused sweeps (measured: 4 beats 2 by 3 % of the iteration)
#endif

// ------------------------------------------------------------------------------------------ smoother epilogues
// plain: the posterior marginals, tiled
struct EpiStore {
    real* pm;
    real* pc;
    real* pm_lin;
    real* pc_lin;
    BN_DEV void prefetch(long long) {}
    BN_DEV void step(long long ti, long long k, real m, real v) {
        if (pm_lin) {
            pm_lin[k] = m;
            pc_lin[k] = v;
        } else {
            pm[ti] = m;
            pc[ti] = v;
        }
    }
    BN_DEV void finish(long long) {}
};

// what the site epilogues need besides the tiled arrays (kernel parameter)
struct ItSiteArgs {
    real lik_param, lr, power;
    int ensure_psd, pad_;
    real* part1;  // [nchunks] per-chunk partial sums: SITES |delta nat1|, ENERGY the likelihood term
    real* part2;  // [nchunks]                         SITES |delta nat2|, ENERGY E_q[log N(pseudo_y | f, pseudo_var)]
};

// site update on the smoothed marginal (the body of inference.py:72-86 for one step); sites rewritten in place
template <int LIK, int METHOD, bool TAB>
struct EpiSites {
    ItIO io;
    ItSiteArgs a;
    Lik1<LIK, TAB> lik;
    const Cub1* cub;
    real d1, d2, yq, oy, oR;
    BN_DEV EpiSites(const ItIO& io_, const ItSiteArgs& a_, const Cub1* cub_, const double* tab)
        : io(io_), a(a_), lik{a_.lik_param, tab}, cub(cub_), d1(0.0), d2(0.0), yq(0.0), oy(0.0), oR(1.0) {}
    BN_DEV void prefetch(long long ti) { yq = io.y[ti]; }
    BN_DEV void step(long long ti, long long, real m, real v) {
        // the old site is needed after the cubature loop only: its loads are issued here and land while the loop runs
        oy = io.sy[ti];
        oR = io.sR[ti];
        // natural parameters of the site as reparametrise leaves them (basemodels.py:85-100): nat2 = 1 / cov, nat1 = nat2 mean
        const real o2 = 1.0 / oR, o1 = oy * o2;
        SiteStats1 s;
        real h, r1, r2, e1, e2;
        site_update_scalar<LIK, METHOD, TAB, BN_IT_TAB_UNROLL>(lik, *cub, yq, m, v, o1, o2, a.lr, a.power, a.ensure_psd, s, h, r1, r2, e1, e2);
        d1 += e1;
        d2 += e2;
        const real Lc = sqrt(r2);
        io.sy[ti] = (r1 / Lc) / Lc;
        io.sR[ti] = (1.0 / Lc) / Lc;
    }
    BN_DEV void finish(long long c) {
        a.part1[c] = d1;
        a.part2[c] = d2;
    }
};

// energy terms on the smoothed marginal + the marginals themselves
template <int LIK, int METHOD, bool TAB>
struct EpiEnergy {
    ItIO io;
    ItSiteArgs a;
    Lik1<LIK, TAB> lik;
    const Cub1* cub;
    real accV, accX, yq, oy, oR;
    unsigned char mk;
    BN_DEV EpiEnergy(const ItIO& io_, const ItSiteArgs& a_, const Cub1* cub_, const double* tab)
        : io(io_), a(a_), lik{a_.lik_param, tab}, cub(cub_), accV(0.0), accX(0.0), yq(0.0), oy(0.0), oR(1.0), mk(0) {}
    BN_DEV void prefetch(long long ti) { yq = io.y[ti]; }
    BN_DEV void step(long long ti, long long k, real m, real v) {
        // the site of this step enters after the cubature loop only: its loads land while the loop runs
        oy = io.sy[ti];
        oR = io.sR[ti];
        if (io.mask) mk = io.mask[ti];
        if (io.pm_lin) {
            io.pm_lin[k] = m;
            io.pc_lin[k] = v;
        } else {
            io.pm[ti] = m;
            io.pc[ti] = v;
        }
        if constexpr (METHOD == BN_METHOD_EP) {
            // EP energy (inference.py:286-325): log Z of the tilted distribution at the cavity, and the same for the site
            // (basemodels.py:247-262) -- the cavity needs the site's natural parameters
            const real o2 = 1.0 / oR, o1 = oy * o2;
            const SiteStats1 s = site_stats_1<LIK, METHOD, false, TAB, BN_IT_TAB_UNROLL>(lik, yq, m, v, o1, o2, a.power, *cub);
            if (!isnan(s.val)) accV += s.val;
            accX += ep_pseudo_step<1>(a.power, 1, &oy, &oR, &m, &v, &o1, &o2, io.mask ? &mk : nullptr, 0);
        } else {
            const SiteStats1 s = site_stats_1<LIK, METHOD, false, TAB, BN_IT_TAB_UNROLL>(lik, yq, m, v, 0.0, 0.0, a.power, *cub);
            if (!isnan(s.val)) accV += s.val;  // nansum (inference.py:218)
            accX += gaussian_ell_step<1>(&oy, &m, &v, &oR, io.mask ? &mk : nullptr, 0);
        }
    }
    BN_DEV void finish(long long c) {
        a.part1[c] = accV;
        a.part2[c] = accX;
    }
};

#ifdef __CUDACC__
}  // namespace BN_NS
#include "tma_stage.cuh"
namespace BN_NS {
// ------------------------------------------------------------------------------------------ kernels
// the table-gathering sweeps own a whole SM: one CTA with the table in shared memory, 16 warps (d <= 3) or the 4 warps the
// registers of a larger state leave room for
// (fp32 build: the table is 8 KB, every CTA of the ordinary configuration stages its own copy)
template <class G> constexpr int kItTabThreads = kReal32 ? kUpThreads : (G::d <= 3 ? 512 : kUpBlocksPerSMWide * kUpThreads);
#ifdef BN_REAL32
constexpr unsigned kItTabBytes = kPt32Bytes;
#else
constexpr unsigned kItTabBytes = sizeof(double) * kPtDoubles;
#endif

// level 0 of the scan, done by the warp that produced the 32 elements (scan.cuh: warp_prescan); the element is read
// back from where the chunk body left it so that the body's registers are dead by now
template <class Alg>
__device__ __forceinline__ void it_prescan_tail(const real* elems, long long n, long long i, const ScanPlan& plan) {
    __syncwarp();
    typename Alg::Elem e;
    if (i < n) Alg::load(elems, n, i, e);
    warp_prescan<Alg>(e, i, plan);
}

// smoothing elements of the chunks in scan order (position p = nchunks - 1 - c), with level 0 of their scan
template <class G>
__global__ void __launch_bounds__(128)
it_selem_kernel(long long N, int L, long long nchunks, int need_first, const real* agg, const real* s0, const real* fs,
                real* selems, ScanPlan plan) {
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if ((p & ~31LL) >= nchunks) return;
    if (p < nchunks) up_selem_chunk<G>(N, L, nchunks, need_first, agg, s0, fs, selems, nchunks - 1 - p);
    it_prescan_tail<SmootherAlg<G::d>>(selems, nchunks, p, plan);
}

template <class G>
__global__ void __launch_bounds__(kUpThreads, (G::d <= 3 ? kUpBlocksPerSM : kUpBlocksPerSMWide))
it_reduce_kernel(G g, ItIO io, int L, long long nchunks, int is_first, real* agg, ScanPlan plan) {
    const long long c = (long long)blockIdx.x * kUpThreads + threadIdx.x;
    if ((c & ~31LL) >= nchunks) return;
    if (c < nchunks) it_reduce_chunk(g, io, L, nchunks, is_first, agg, c);
    if (c == 0 && plan.ticket) plan.ticket[2] = 0u;  // the arrival counter of this pass's sums (it_sum_kernel)
    it_prescan_tail<FilterAlg<G::d>>(agg, nchunks, c, plan);
}

template <class G, bool WANT_ELL>
__global__ void __launch_bounds__(kUpThreads, (G::d <= 3 ? kUpBlocksPerSM : kUpBlocksPerSMWide))
it_filter_kernel(G g, ItIO io, int L, long long nchunks, int is_first, const real* prefix, PrefixParts wprefix,
                 const real* s0, real* fs, real* ell_partials, const int* jst) {
    const long long c = (long long)blockIdx.x * kUpThreads + threadIdx.x;
    if (c < nchunks) it_filter_chunk<G, WANT_ELL>(g, io, L, nchunks, is_first, prefix, s0, fs, ell_partials, c, jst, wprefix);
}

// phase 1 with speculation (SpecReduce): every lane of a warp takes part in the vote, chunk or not
template <class G, bool WANT_ELL>
__global__ void __launch_bounds__(kUpThreads, (G::d <= 3 ? kUpBlocksPerSM : kUpBlocksPerSMWide))
it_reduce_spec_kernel(G g, ItIO io, int L, long long nchunks, int is_first, real* agg, real* fs, real* ell_partials,
                      int* jst, ScanPlan plan, real thr) {
    const long long c = (long long)blockIdx.x * kUpThreads + threadIdx.x;
    if ((c & ~31LL) >= nchunks) return;
    const bool active = c < nchunks;
    SpecReduce<G, WANT_ELL> sr;
    sr.init(g, io, L, fs, active ? c : 0, active, thr);
    int j = 0;
    bool all_dec = false;
#pragma unroll 1
    while (!all_dec) {
        bool dec = true;
        if (j < sr.cnt) dec = sr.absorb(g, is_first && c == 0 && j == 0);
        ++j;
        all_dec = __all_sync(0xffffffffu, dec);
    }
    const int jstar = j < sr.cnt ? j : sr.cnt;
    // A, J, eta are final here: the whole aggregate goes out now, and the filter phase keeps only (b, C) in registers
    if (active) FilterAlg<G::d>::store(agg, nchunks, c, sr.el);
#pragma unroll 1
    for (; j < sr.cnt; ++j) sr.filter(g);
    if (active) {
        if (jstar < sr.cnt) sr.store_state(agg, nchunks, c);
        jst[c] = jstar;
        if (WANT_ELL) ell_partials[c] = sr.ell;
    }
    if (c == 0 && plan.ticket) plan.ticket[2] = 0u;
    it_prescan_tail<FilterAlg<G::d>>(agg, nchunks, c, plan);
}

// Fixed-order sums of the per-chunk partials (run-to-run bit-stable), accumulated in fp64 whatever `real` is, in ONE
// launch: kSumCtas CTAs reduce a fixed slice each; the CTA that arrives last (ticket, zero at launch and left at zero) adds
// the CTA partials in index order.
//   two = 0: out[0] = sum a (+ sum b when b is given)        two = 1: out[0] = sum a, out[1] = sum b
constexpr int kSumCtas = 74;
static __global__ void __launch_bounds__(256) it_sum_kernel(const real* a, const real* b, long long n, real* out, int two,
                                                            double* scratch, unsigned int* ticket) {
    __shared__ double sh[2][256];
    __shared__ unsigned int last;
    const long long per = (n + kSumCtas - 1) / kSumCtas, lo = blockIdx.x * per, hi = (lo + per < n) ? lo + per : n;
    double sa = 0.0, sb = 0.0;
    for (long long i = lo + threadIdx.x; i < hi; i += 256) {
        sa += (double)a[i];
        if (b) sb += (double)b[i];
    }
    sh[0][threadIdx.x] = sa;
    sh[1][threadIdx.x] = sb;
    __syncthreads();
    for (int off = 128; off > 0; off >>= 1) {
        if ((int)threadIdx.x < off) {
            sh[0][threadIdx.x] += sh[0][threadIdx.x + off];
            sh[1][threadIdx.x] += sh[1][threadIdx.x + off];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        scratch[blockIdx.x] = sh[0][0];
        scratch[kSumCtas + blockIdx.x] = sh[1][0];
        __threadfence();
        last = (atomicAdd(ticket, 1u) == gridDim.x - 1) ? 1u : 0u;
    }
    __syncthreads();
    if (!last || threadIdx.x >= 32) return;
    __threadfence();
    sa = sb = 0.0;
    for (int i = threadIdx.x; i < kSumCtas; i += 32) {
        sa += __ldcg(scratch + i);
        sb += __ldcg(scratch + kSumCtas + i);
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        sa += __shfl_down_sync(0xffffffffu, sa, off);
        sb += __shfl_down_sync(0xffffffffu, sb, off);
    }
    if (threadIdx.x == 0) {
        if (two) { out[0] = real(sa); out[1] = real(sb); }
        else out[0] = real(sa + sb);
        *ticket = 0u;
    }
}
inline cudaError_t it_sum(const real* a, const real* b, long long n, real* out, int two, double* scratch, unsigned int* ticket,
                          cudaStream_t st) {
    it_sum_kernel<<<kSumCtas, 256, 0, st>>>(a, b, n, out, two, scratch, ticket);
    return cudaGetLastError();
}

template <class G>
__global__ void __launch_bounds__(kUpThreads, (G::d <= 3 ? kUpBlocksPerSM : kUpBlocksPerSMWide))
it_smooth_plain_kernel(G g, ItIO io, int L, long long nchunks, const real* sprefix, PrefixParts swprefix, const real* sinit,
                       const real* fs) {
    const long long c = (long long)blockIdx.x * kUpThreads + threadIdx.x;
    if (c >= nchunks) return;
    EpiStore epi{io.pm, io.pc, io.pm_lin, io.pc_lin};
    it_smooth_chunk(g, io, L, nchunks, sprefix, sinit, fs, c, epi, swprefix);
}

// the probit log-density table in device memory (sites.cu fills it once per device and hands out its address)
int probit_table_device(cudaStream_t st, const double** tab);

template <class G, template <int, int, bool> class Epi, int LIK, int METHOD, bool TAB>
__global__ void __launch_bounds__(TAB ? kItTabThreads<G> : kUpThreads, (TAB && !kReal32) ? 1 : (G::d <= 3 ? kUpBlocksPerSM : kUpBlocksPerSMWide))
it_smooth_site_kernel(G g, ItIO io, const __grid_constant__ Cub1 cub, ItSiteArgs a, int L, long long nchunks,
                      const real* sprefix, PrefixParts swprefix, const real* sinit, const real* fs, const double* gtab) {
    extern __shared__ __align__(16) double it_smem[];
    const double* tab = nullptr;
    if constexpr (TAB) {
        tma_stage_to_smem(it_smem, gtab, kItTabBytes);  // cp.async.bulk + mbarrier
        tab = it_smem;
    }
    const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nchunks) return;
    Epi<LIK, METHOD, TAB> epi(io, a, &cub, tab);
    it_smooth_chunk(g, io, L, nchunks, sprefix, sinit, fs, c, epi, swprefix);
}

// linear [N] <-> tiled, 32 chunks x 32 steps per CTA through a padded shared-memory tile (both sides coalesced)
template <typename T, bool TO_TILED>
__global__ void __launch_bounds__(256) it_transpose_kernel(long long N, int L, long long nchunks, const T* in, T* out, T fill) {
    __shared__ T sm[32][33];
    const long long tile = blockIdx.x;
    const int j0 = blockIdx.y * 32, tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    if constexpr (TO_TILED) {
#pragma unroll
        for (int r = ty; r < 32; r += 8) {  // r: chunk of the tile, tx: step
            const long long c = tile * 32 + r, k = c * L + j0 + tx;
            sm[r][tx] = (j0 + tx < L && c < nchunks && k < N) ? in[k] : fill;
        }
        __syncthreads();
#pragma unroll
        for (int r = ty; r < 32; r += 8)  // r: step, tx: chunk
            if (j0 + r < L) out[((tile * L + j0 + r) << 5) + tx] = sm[tx][r];
    } else {
#pragma unroll
        for (int r = ty; r < 32; r += 8)
            if (j0 + r < L) sm[r][tx] = in[((tile * L + j0 + r) << 5) + tx];
        __syncthreads();
#pragma unroll
        for (int r = ty; r < 32; r += 8) {
            const long long c = tile * 32 + r, k = c * L + j0 + tx;
            if (j0 + tx < L && c < nchunks && k < N) out[k] = sm[tx][r];
        }
    }
}

template <typename T>
__global__ void it_fill_pad_kernel(T* x, long long from, long long n, T fill) {
    const long long i = from + (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) x[i] = fill;
}

// ------------------------------------------------------------------------------------------ host driver
enum { IT_PLAIN = 0, IT_SITES = 1, IT_ENERGY = 2 };

struct ItCall {
    const bn_kernel_spec* spec;
    ItIO io;
    int mode;                 // smoother epilogue
    int method, likelihood, use_table;
    ItSiteArgs sa;            // part1 / part2 are set by the driver
    const Cub1* cub;
    real* ell;              // nullable
    real* sums;             // [2], SITES: sum |delta nat1|, sum |delta nat2|; ENERGY: likelihood term, pseudo term
    void* ws;
    size_t ws_bytes;
    cudaStream_t st;
    int phase, rank, world;
    real* carry_out;
    const real* carries;
    int spec_filter;          // phase 1 may switch to the plain filter once the chunk has forgotten its start (SpecReduce)
    int spec_min_chunk;       // ... for chunks at least this long
    double spec_thr;          // ... "forgotten" = scaled sensitivity to the incoming state below this
    int want_ell;             // the pass produces the log-likelihood (every phase of one pass must agree on it)
};

template <int d>
inline size_t it_ws_doubles(long long N) {
    return up_ws_doubles<d>(N) + 4 * (size_t)up_plan_chunks(N > 0 ? N : 1, false, d).nchunks + 64 + 4 * kSumCtas + 16;
}

// speculative phase 1 pays off when the chunks are much longer than the filter's forgetting time
constexpr int kSpecMinChunk = 160;  // measured break-even: on at L = 165 (-2.3 % per iteration), off at L = 132 (+1 %)

// fused (likelihood, method) pairs of the site / energy epilogues
#define BN_FOR_EACH_ITER_SITE(X)                                                                                       \
    X(BN_LIK_GAUSSIAN, BN_METHOD_VI) X(BN_LIK_GAUSSIAN, BN_METHOD_NEWTON) X(BN_LIK_GAUSSIAN, BN_METHOD_EP)               \
    X(BN_LIK_BERNOULLI_PROBIT, BN_METHOD_VI) X(BN_LIK_BERNOULLI_PROBIT, BN_METHOD_NEWTON) X(BN_LIK_BERNOULLI_PROBIT, BN_METHOD_EP) \
    X(BN_LIK_BERNOULLI_LOGIT, BN_METHOD_VI) X(BN_LIK_BERNOULLI_LOGIT, BN_METHOD_NEWTON) X(BN_LIK_BERNOULLI_LOGIT, BN_METHOD_EP) \
    X(BN_LIK_POISSON_EXP, BN_METHOD_VI) X(BN_LIK_POISSON_EXP, BN_METHOD_NEWTON) X(BN_LIK_POISSON_EXP, BN_METHOD_EP)

template <class G, template <int, int, bool> class Epi>
inline int it_launch_site_sweep(const ItCall& c, const G& g, const ChunkPlan& cp, const UpWs& w, const ItSiteArgs& sa) {
    cudaStream_t st = c.st;
    const char* name = (c.mode == IT_SITES) ? "it_smooth_sites" : "it_smooth_energy";
    const PrefixParts swp = upper_parts(w.splan);
#define X(LK, M)                                                                                                      \
    if (c.likelihood == LK && c.method == M) {                                                                         \
        if constexpr (LK == BN_LIK_BERNOULLI_PROBIT && (M == BN_METHOD_VI || M == BN_METHOD_EP)) {                     \
            if (c.use_table) {                                                                                         \
                auto kfn = it_smooth_site_kernel<G, Epi, LK, M, true>;                                                 \
                const size_t smem = kItTabBytes;                                                                       \
                const double* gtab = nullptr;                                                                          \
                if (int rc = probit_table_device(st, &gtab)) return rc;                                                \
                BN_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));            \
                const unsigned grid = (unsigned)((cp.nchunks + kItTabThreads<G> - 1) / kItTabThreads<G>);              \
                BN_LAUNCH(name, st, (kfn<<<grid, kItTabThreads<G>, smem, st>>>(g, c.io, *c.cub, sa, cp.L, cp.nchunks,  \
                                                                            w.splan.prefix[0], swp, w.sinit, w.fs, gtab))); \
                BN_CUDA(cudaGetLastError());                                                                           \
                return 0;                                                                                              \
            }                                                                                                          \
        }                                                                                                              \
        const unsigned grid = (unsigned)((cp.nchunks + kUpThreads - 1) / kUpThreads);                                  \
        BN_LAUNCH(name, st, (it_smooth_site_kernel<G, Epi, LK, M, false><<<grid, kUpThreads, 0, st>>>(                 \
                                g, c.io, *c.cub, sa, cp.L, cp.nchunks, w.splan.prefix[0], swp, w.sinit, w.fs, nullptr))); \
        BN_CUDA(cudaGetLastError());                                                                                   \
        return 0;                                                                                                      \
    }
    BN_FOR_EACH_ITER_SITE(X)
#undef X
    set_error("the fused iteration has no epilogue for (likelihood, method) = (%d, %d)", c.likelihood, c.method);
    return -1;
}

template <class G>
inline int it_run(const ItCall& c) {
    constexpr int d = G::d;
    using FA = FilterAlg<d>;
    using SA = SmootherAlg<d>;
    cudaStream_t st = c.st;
    const ItIO& io = c.io;
    G g;
    g.prepare(*c.spec);
    ChunkPlan cp = up_plan_chunks(io.N, false, d);
    const size_t need = it_ws_doubles<d>(io.N) * sizeof(real);
    BN_REQUIRE(c.ws != nullptr && c.ws_bytes >= need, "workspace too small: need %zu bytes, got %zu", need, c.ws_bytes);
    UpWs w = up_ws<d>(c.ws, cp);
    // level 0 of both scans is done by the kernels that produce the elements (warp_prescan); the upper levels live in
    // the region the hyper-gradient partials of the unfused update would use
    w.fplan = make_scan_plan_warp(w.fplan.input0, w.fplan.prefix[0], w.gpart, cp.nchunks, FA::kElem);
    w.splan = make_scan_plan_warp(w.splan.input0, w.splan.prefix[0], w.gpart + scan_upper_doubles(cp.nchunks, FA::kElem),
                                  cp.nchunks, SA::kElem);
    static_assert(2 * (FA::kElem + SA::kElem) <= 32 * kUpMaxGradFields, "upper scan levels do not fit the gradient region");
    real* part = (real*)c.ws + up_ws_doubles<d>(io.N);
    real* ell1 = part + 2 * cp.nchunks;              // log-likelihood partials of a speculative phase 1
    int* jst = (int*)(part + 3 * cp.nchunks);          // its switch step per chunk
    double* sum_scratch = (double*)(((uintptr_t)(part + 4 * cp.nchunks) + 15) & ~(uintptr_t)15);  // 2 kSumCtas fp64 partials
    // arrival counters of the one-launch upper scan levels (cleared by the kernels that produce the elements)
    w.fplan.ticket = (unsigned int*)(sum_scratch + 2 * kSumCtas);
    w.splan.ticket = w.fplan.ticket + 1;
    unsigned int* sum_ticket = w.fplan.ticket + 2;  // arrival counter of the sums (cleared by the reduce kernels as well)
    const PrefixParts fwp = upper_parts(w.fplan), swp = upper_parts(w.splan);
    const unsigned grid = (unsigned)((cp.nchunks + kUpThreads - 1) / kUpThreads);
    const int is_first = (c.rank == 0), is_last = (c.rank == c.world - 1);
    const bool sharded = c.phase != UP_ALL;
    const bool spec = c.spec_filter && cp.L >= (c.spec_min_chunk > 0 ? c.spec_min_chunk : kSpecMinChunk);
    const real spec_thr = c.spec_thr > 0.0 ? c.spec_thr : (double)kSpecThreshold;

    if (c.phase == UP_ALL || c.phase == UP_REDUCE) {
        if (spec) {
            if (c.want_ell) BN_LAUNCH("it_reduce_spec", st, (it_reduce_spec_kernel<G, true><<<grid, kUpThreads, 0, st>>>(
                                                           g, io, cp.L, cp.nchunks, is_first, w.fplan.input0, w.fs, ell1, jst, w.fplan, spec_thr)));
            else BN_LAUNCH("it_reduce_spec", st, (it_reduce_spec_kernel<G, false><<<grid, kUpThreads, 0, st>>>(
                                                     g, io, cp.L, cp.nchunks, is_first, w.fplan.input0, w.fs, ell1, jst, w.fplan, spec_thr)));
        } else {
            BN_LAUNCH("it_reduce", st, (it_reduce_kernel<G><<<grid, kUpThreads, 0, st>>>(g, io, cp.L, cp.nchunks, is_first, w.fplan.input0, w.fplan)));
        }
        BN_CUDA(cudaGetLastError());
        BN_CUDA(run_scan_upper<FA>(w.fplan, st));
        if (c.carry_out && c.phase == UP_REDUCE) {
            const int top = w.fplan.levels - 1;
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
            BN_LAUNCH("it_filter", st, (it_filter_kernel<G, true><<<grid, kUpThreads, 0, st>>>(
                                           g, io, cp.L, cp.nchunks, is_first, w.fplan.prefix[0], fwp, w.s0, w.fs, w.partials,
                                           spec ? jst : nullptr)));
            BN_CUDA(cudaGetLastError());
            BN_LAUNCH("sum", st, (it_sum(w.partials, spec ? ell1 : nullptr, cp.nchunks, c.ell, 0, sum_scratch, sum_ticket, st)));
        } else {
            BN_LAUNCH("it_filter", st, (it_filter_kernel<G, false><<<grid, kUpThreads, 0, st>>>(
                                           g, io, cp.L, cp.nchunks, is_first, w.fplan.prefix[0], fwp, w.s0, w.fs, nullptr,
                                           spec ? jst : nullptr)));
        }
        BN_CUDA(cudaGetLastError());
        const unsigned g2 = (unsigned)((cp.nchunks + 127) / 128);
        BN_LAUNCH("it_selem", st, (it_selem_kernel<G><<<g2, 128, 0, st>>>(io.N, cp.L, cp.nchunks, !is_first, w.fplan.input0,
                                                                          w.s0, w.fs, w.splan.input0, w.splan)));
        BN_CUDA(cudaGetLastError());
        BN_CUDA(run_scan_upper<SA>(w.splan, st));
        if (c.carry_out && c.phase == UP_FILTER) {
            const int top = w.splan.levels - 1;
            up_export_scarry_kernel<d><<<1, 1, 0, st>>>(w.splan.prefix[top], w.splan.count[top], is_last, io.N, cp.L, w.fs,
                                                        c.carry_out);
            BN_CUDA(cudaGetLastError());
        }
    }
    if (c.phase == UP_ALL || c.phase == UP_SMOOTH) {
        if (sharded && !is_last) fold_carries_kernel<SA><<<1, 1, 0, st>>>(c.carries, c.world - 1, c.rank, -1, w.sinit);
        else up_last_state_kernel<d><<<1, 1, 0, st>>>(io.N, cp.L, w.fs, w.sinit);
        BN_CUDA(cudaGetLastError());
        if (c.mode == IT_PLAIN) {
            BN_LAUNCH("it_smooth", st, (it_smooth_plain_kernel<G><<<grid, kUpThreads, 0, st>>>(g, io, cp.L, cp.nchunks,
                                                                                               w.splan.prefix[0], swp, w.sinit, w.fs)));
            BN_CUDA(cudaGetLastError());
        } else {
            ItSiteArgs sa = c.sa;
            sa.part1 = part;
            sa.part2 = part + cp.nchunks;
            int rc = (c.mode == IT_SITES) ? it_launch_site_sweep<G, EpiSites>(c, g, cp, w, sa)
                                          : it_launch_site_sweep<G, EpiEnergy>(c, g, cp, w, sa);
            if (rc) return rc;
            if (c.sums) {
                BN_LAUNCH("sum", st, (it_sum(sa.part1, sa.part2, cp.nchunks, c.sums, 1, sum_scratch, sum_ticket, st)));
                BN_CUDA(cudaGetLastError());
            }
        }
    }
    return 0;
}
#endif  // __CUDACC__

}  // namespace BN_NS
