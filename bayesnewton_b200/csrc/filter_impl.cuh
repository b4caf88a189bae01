// Kalman filter kernels: the sequential recursion (one thread walks time in order) and the
// three-phase temporally parallel form
//     kf_reduce  : one thread per chunk of L steps folds its steps into a filtering element
//     run_scan   : CTA/warp-shuffle scan over the chunk elements (scan.cuh)
//     kf_apply   : one thread per chunk re-runs the plain filter from its now-known incoming state
// Reference: bayesnewton/ops.py:154-180 (_sequential_kf), :183-253 (_parallel_kf), :256-285.
#pragma once
#include "common.cuh"
#include "core.cuh"
#include "scan.cuh"

namespace bn {

struct KfIO {
    long long N;
    const double* y;            // [N,D]
    const double* R;            // [N,D,D]
    const unsigned char* mask;  // [N,D] or null
    double* fms;                // [N,d] or null
    double* fPs;                // [N,d,d] or null
    int return_predict;
};

template <int d>
BN_DEV void write_state(double* fms, double* fPs, long long k, const double* m, const double* P) {
    double* pm = fms + k * d;
#pragma unroll
    for (int i = 0; i < d; ++i) pm[i] = m[i];
    double* pP = fPs + k * (d * d);
#pragma unroll
    for (int i = 0; i < d; ++i)
#pragma unroll
        for (int j = 0; j < d; ++j) pP[i * d + j] = P[sidx(i, j)];
}

template <int D>
BN_DEV void load_obs(const KfIO& io, long long k, double* y, double* R, unsigned char* mk) {
#pragma unroll
    for (int i = 0; i < D; ++i) y[i] = io.y[k * D + i];
#pragma unroll
    for (int i = 0; i < D * D; ++i) R[i] = io.R[k * (D * D) + i];
#pragma unroll
    for (int i = 0; i < D; ++i) mk[i] = io.mask ? io.mask[k * D + i] : (unsigned char)0;
}

// ---------------------------------------------------------------------------- sequential form
template <class Gen, bool WANT_ELL>
BN_DEV void kf_seq_body(const Gen& gen, const KfIO& io, double* ell_out) {
    constexpr int d = Gen::d, D = Gen::D;
    double m[d], P[symn(d)], H[D * d];
    gen.m0(m);
    gen.pinf(P);
    gen.H(H);
    double ell = 0.0;
    for (long long k = 0; k < io.N; ++k) {
        double A[d * d], Q[symn(d)], y[D], R[D * D], mp[d], Pp[symn(d)];
        unsigned char mk[D];
        gen.step(k, A, Q);
        load_obs<D>(io, k, y, R, mk);
        ell += kf_step<d, D, WANT_ELL>(m, P, A, Q, H, y, R, io.mask ? mk : nullptr, mp, Pp);
        if (io.fms) {
            if (io.return_predict) write_state<d>(io.fms, io.fPs, k, mp, Pp);
            else write_state<d>(io.fms, io.fPs, k, m, P);
        }
    }
    if (WANT_ELL && ell_out) *ell_out = ell;
}

template <class Gen, bool WANT_ELL>
__global__ void kf_seq_kernel(Gen gen, KfIO io, double* ell_out) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    kf_seq_body<Gen, WANT_ELL>(gen, io, ell_out);
}

// ---------------------------------------------------------------------------- scan form, phase 1
template <class Gen>
BN_DEV void kf_reduce_chunk(const Gen& gen, const KfIO& io, int L, long long nchunks, int is_first, double* agg,
                            long long c) {
    constexpr int d = Gen::d, D = Gen::D;
    using Alg = FilterAlg<d>;
    double H[D * d];
    gen.H(H);
    typename Alg::Elem g;
    Alg::identity(g);
    const long long k0 = c * L, k1 = (k0 + L < io.N) ? k0 + L : io.N;
    for (long long k = k0; k < k1; ++k) {
        double A[d * d], Q[symn(d)], y[D], R[D * D];
        unsigned char mk[D];
        gen.step(k, A, Q);
        load_obs<D>(io, k, y, R, mk);
        if (k == 0 && is_first) {
            double m0[d], P0[symn(d)];
            gen.m0(m0);
            gen.pinf(P0);
            filter_absorb<d, D>(g, A, P0, H, y, R, true, m0);
        } else {
            filter_absorb<d, D>(g, A, Q, H, y, R, false, (const double*)nullptr);
        }
    }
    Alg::store(agg, nchunks, c, g);
}

template <class Gen>
__global__ void __launch_bounds__(kChunkThreads)
kf_reduce_kernel(Gen gen, KfIO io, int L, long long nchunks, int is_first, double* agg) {
    const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nchunks) return;
    kf_reduce_chunk(gen, io, L, nchunks, is_first, agg, c);
}

// ---------------------------------------------------------------------------- scan form, phase 3
// prefix: inclusive prefixes of the chunk elements; s0: the state entering this shard
// (zero on a single GPU / rank 0, where global step 0 starts from (m0, P0) by the first-step rule).
template <class Gen, bool WANT_ELL>
BN_DEV void kf_apply_chunk(const Gen& gen, const KfIO& io, int L, long long nchunks, int is_first,
                           const double* prefix, const double* s0, double* ell_partials, long long c) {
    constexpr int d = Gen::d, D = Gen::D;
    using Alg = FilterAlg<d>;
    double H[D * d];
    gen.H(H);
    typename Alg::State s;
    Alg::load_state(s0, 1, 0, s);
    if (c > 0) {
        typename Alg::Elem e;
        Alg::load(prefix, nchunks, c - 1, e);
        typename Alg::State t;
        Alg::apply(e, s, t);
        s = t;
    }
    double ell = 0.0;
    const long long k0 = c * L, k1 = (k0 + L < io.N) ? k0 + L : io.N;
    for (long long k = k0; k < k1; ++k) {
        double A[d * d], Q[symn(d)], y[D], R[D * D], mp[d], Pp[symn(d)];
        unsigned char mk[D];
        gen.step(k, A, Q);
        load_obs<D>(io, k, y, R, mk);
        const unsigned char* mkp = io.mask ? mk : nullptr;
        if (k == 0 && is_first) {
            // scan-form first step (ops.py:222-229, 245-248): the update starts from (m0, P0)
            // itself; the log-likelihood / predicted outputs use A_0 m0, A_0 P0 A_0^T + Q_0.
            double m0[d], P0[symn(d)];
            gen.m0(m0);
            gen.pinf(P0);
            if (WANT_ELL || io.return_predict) {
                double mt[d], Pt[symn(d)];
#pragma unroll
                for (int i = 0; i < d; ++i) mt[i] = m0[i];
#pragma unroll
                for (int i = 0; i < symn(d); ++i) Pt[i] = P0[i];
                ell += kf_step<d, D, WANT_ELL>(mt, Pt, A, Q, H, y, R, mkp, mp, Pp);
            }
            double I[d * d], Z[symn(d)], mq[d], Pq[symn(d)];
#pragma unroll
            for (int i = 0; i < d; ++i)
#pragma unroll
                for (int j = 0; j < d; ++j) I[i * d + j] = (i == j) ? 1.0 : 0.0;
#pragma unroll
            for (int i = 0; i < symn(d); ++i) Z[i] = 0.0;
#pragma unroll
            for (int i = 0; i < d; ++i) s.m[i] = m0[i];
#pragma unroll
            for (int i = 0; i < symn(d); ++i) s.P[i] = P0[i];
            kf_step<d, D, false>(s.m, s.P, I, Z, H, y, R, mkp, mq, Pq);
        } else {
            ell += kf_step<d, D, WANT_ELL>(s.m, s.P, A, Q, H, y, R, mkp, mp, Pp);
        }
        if (io.fms) {
            if (io.return_predict) write_state<d>(io.fms, io.fPs, k, mp, Pp);
            else write_state<d>(io.fms, io.fPs, k, s.m, s.P);
        }
    }
    if (WANT_ELL) ell_partials[c] = ell;
}

template <class Gen, bool WANT_ELL>
__global__ void __launch_bounds__(kChunkThreads)
kf_apply_kernel(Gen gen, KfIO io, int L, long long nchunks, int is_first,
                const double* prefix, const double* s0, double* ell_partials) {
    const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nchunks) return;
    kf_apply_chunk<Gen, WANT_ELL>(gen, io, L, nchunks, is_first, prefix, s0, ell_partials, c);
}

// ---------------------------------------------------------------------------- host drivers
struct KfWs {
    double* s0;
    ScanPlan plan;
    double* partials;
};

template <int d>
inline size_t kf_ws_doubles(long long nchunks) {
    return 64 + scan_plan_doubles(nchunks, FilterAlg<d>::kElem) + nchunks;
}

template <int d>
inline KfWs kf_ws(void* ws, long long nchunks) {
    KfWs w;
    double* p = (double*)ws;
    w.s0 = p;
    p += 64;
    w.plan = make_scan_plan(p, nchunks, FilterAlg<d>::kElem);
    p += scan_plan_doubles(nchunks, FilterAlg<d>::kElem);
    w.partials = p;
    return w;
}



template <class Gen>
inline int kf_run(const Gen& gen, int form, KfIO io, double* ell, void* ws, size_t ws_bytes, cudaStream_t st,
                  int phase, int is_first, double* carry_out, const double* carries, int rank) {
    constexpr int d = Gen::d;
    using Alg = FilterAlg<d>;
    if (io.N == 0) {
        if (ell) BN_CUDA(cudaMemsetAsync(ell, 0, sizeof(double), st));
        return 0;
    }
    if (form == BN_SEQUENTIAL) {
        if (ell) BN_LAUNCH("kf_seq", st, kf_seq_kernel<Gen, true><<<1, 1, 0, st>>>(gen, io, ell));
        else BN_LAUNCH("kf_seq", st, kf_seq_kernel<Gen, false><<<1, 1, 0, st>>>(gen, io, nullptr));
        BN_CUDA(cudaGetLastError());
        return 0;
    }
    ChunkPlan cp = plan_chunks(io.N);
    size_t need = kf_ws_doubles<d>(cp.nchunks) * sizeof(double);
    BN_REQUIRE(ws != nullptr && ws_bytes >= need, "workspace too small: need %zu bytes, got %zu", need, ws_bytes);
    KfWs w = kf_ws<d>(ws, cp.nchunks);
    unsigned grid = (unsigned)((cp.nchunks + kChunkThreads - 1) / kChunkThreads);
    if (phase == PHASE_ALL || phase == PHASE_REDUCE) {
        BN_LAUNCH("kf_reduce", st,
                  kf_reduce_kernel<Gen><<<grid, kChunkThreads, 0, st>>>(gen, io, cp.L, cp.nchunks, is_first,
                                                                         w.plan.input0));
        BN_CUDA(cudaGetLastError());
        BN_CUDA(run_scan<Alg>(w.plan, st));
        if (carry_out) {
            int top = w.plan.levels - 1;
            export_carry_kernel<Alg><<<1, 1, 0, st>>>(w.plan.prefix[top], w.plan.count[top], carry_out);
            BN_CUDA(cudaGetLastError());
        }
    }
    if (phase == PHASE_ALL || phase == PHASE_APPLY) {
        if (carries) {
            fold_carries_kernel<Alg><<<1, 1, 0, st>>>(carries, 0, rank, 1, w.s0);
            BN_CUDA(cudaGetLastError());
        } else {
            BN_CUDA(cudaMemsetAsync(w.s0, 0, Alg::kState * sizeof(double), st));
        }
        if (ell) {
            BN_LAUNCH("kf_apply", st,
                      kf_apply_kernel<Gen, true><<<grid, kChunkThreads, 0, st>>>(
                          gen, io, cp.L, cp.nchunks, is_first, w.plan.prefix[0], w.s0, w.partials));
            BN_CUDA(cudaGetLastError());
            BN_LAUNCH("sum", st, sum_kernel<false><<<1, 1024, 0, st>>>(w.partials, cp.nchunks, ell, 1.0));
        } else {
            BN_LAUNCH("kf_apply", st,
                      kf_apply_kernel<Gen, false><<<grid, kChunkThreads, 0, st>>>(
                          gen, io, cp.L, cp.nchunks, is_first, w.plan.prefix[0], w.s0, nullptr));
        }
        BN_CUDA(cudaGetLastError());
    }
    return 0;
}


// one call, type-erased over the generator: what the per-group translation units dispatch on
struct KfCall {
    int form;
    KfIO io;
    double* ell;
    void* ws;
    size_t ws_bytes;
    cudaStream_t st;
    int phase, is_first;
    double* carry_out;
    const double* carries;
    int rank;
    // stationary-kernel entry
    const bn_kernel_spec* spec;
    const double* dt;
    // array entry
    int d, D;
    const double *As, *Qs, *H, *m0, *P0;
};


#define BN_KF_SPEC_CASE(FAM, NC)                                                                        \
    if (c.spec->family == FAM && c.spec->n_components == NC) {                                           \
        MaternGen<FAM, NC> gen;                                                                          \
        gen.spec = *c.spec;                                                                              \
        gen.dt = c.dt;                                                                                   \
        return kf_run(gen, c.form, c.io, c.ell, c.ws, c.ws_bytes, c.st, c.phase, c.is_first, c.carry_out, \
                      c.carries, c.rank);                                                                \
    }

#define BN_KF_ARR_CASE(DD, OD)                                                                           \
    if (c.d == DD && c.D == OD) {                                                                        \
        ArrayGen<DD, OD> gen{c.As, c.Qs, c.H, c.m0, c.P0};                                               \
        return kf_run(gen, c.form, c.io, c.ell, c.ws, c.ws_bytes, c.st, c.phase, c.is_first, c.carry_out, \
                      c.carries, c.rank);                                                                \
    }

}  // namespace bn
