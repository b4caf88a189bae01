// Rauch-Tung-Striebel smoother kernels: the sequential backward recursion and the three-phase
// temporally parallel form (reduce chunk -> scan chunk elements right-to-left -> re-run chunk).
// Reference: bayesnewton/ops.py:288-311 (_sequential_rts), :314-354 (_parallel_rts), :357-380.
// A_k, Q_k here belong to the step OUT OF k (the caller passes dt shifted by one,
// basemodels.py:700), so the last step has A = I, Q = 0.
#pragma once
#include "common.cuh"
#include "core.cuh"
#include "scan.cuh"

namespace bn {

struct RtsIO {
    long long N;
    const double* fms;  // [N,d]
    const double* fPs;  // [N,d,d]
    double* sms;        // [N,Df] or [N,d]
    double* sPs;        // [N,Df,Df] or [N,d,d]
    double* gains;      // [N,d,d] or null
    int return_full;
};

template <int d>
BN_DEV void load_filtered(const RtsIO& io, long long k, double* fm, double* fP) {
    const double* pm = io.fms + k * d;
#pragma unroll
    for (int i = 0; i < d; ++i) fm[i] = pm[i];
    const double* pP = io.fPs + k * (d * d);
#pragma unroll
    for (int i = 0; i < d; ++i)
#pragma unroll
        for (int j = 0; j <= i; ++j) fP[sidx(i, j)] = pP[i * d + j];
}

template <int d, int Df>
BN_DEV void write_smoothed(const RtsIO& io, long long k, const double* H, const double* sm, const double* sP,
                           const double* G) {
    if (io.return_full) {
        double* pm = io.sms + k * d;
#pragma unroll
        for (int i = 0; i < d; ++i) pm[i] = sm[i];
        double* pP = io.sPs + k * (d * d);
#pragma unroll
        for (int i = 0; i < d; ++i)
#pragma unroll
            for (int j = 0; j < d; ++j) pP[i * d + j] = sP[sidx(i, j)];
    } else {
        double hm[Df], HP[Df * d];
        matvec<Df, d>(H, sm, hm);
        mat_sym<Df, d>(H, sP, HP);
        double* pm = io.sms + k * Df;
#pragma unroll
        for (int i = 0; i < Df; ++i) pm[i] = hm[i];
        double* pP = io.sPs + k * (Df * Df);
#pragma unroll
        for (int i = 0; i < Df; ++i)
#pragma unroll
            for (int j = 0; j < Df; ++j) {
                double s = 0.0;
#pragma unroll
                for (int l = 0; l < d; ++l) s = fma(HP[i * d + l], H[j * d + l], s);
                pP[i * Df + j] = s;
            }
    }
    if (io.gains) {
        double* pg = io.gains + k * (d * d);
#pragma unroll
        for (int i = 0; i < d * d; ++i) pg[i] = G[i];
    }
}

// ---------------------------------------------------------------------------- sequential form
template <class Gen>
BN_DEV void rts_seq_body(const Gen& gen, const RtsIO& io) {
    constexpr int d = Gen::d, Df = Gen::D;
    double H[Df * d], sm[d], sP[symn(d)];
    gen.H(H);
    load_filtered<d>(io, io.N - 1, sm, sP);
    for (long long k = io.N - 1; k >= 0; --k) {
        double A[d * d], Q[symn(d)], fm[d], fP[symn(d)], G[d * d], pm[d], pP[symn(d)];
        gen.step(k, A, Q);
        load_filtered<d>(io, k, fm, fP);
        rts_gain<d>(fm, fP, A, Q, G, pm, pP);
        rts_step<d>(sm, sP, fm, fP, G, pm, pP);
        write_smoothed<d, Df>(io, k, H, sm, sP, G);
    }
}

template <class Gen>
__global__ void rts_seq_kernel(Gen gen, RtsIO io) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    rts_seq_body(gen, io);
}

// ---------------------------------------------------------------------------- scan form, phase 1
// chunk c covers steps [cL, (c+1)L); its scan position is p = nchunks-1-c (right-to-left).
template <class Gen>
BN_DEV void rts_reduce_chunk(const Gen& gen, const RtsIO& io, int L, long long nchunks, int is_last, double* agg,
                             long long c) {
    constexpr int d = Gen::d;
    using Alg = SmootherAlg<d>;
    typename Alg::Elem acc;
    Alg::identity(acc);
    const long long k0 = c * L, k1 = (k0 + L < io.N) ? k0 + L : io.N;
    for (long long k = k1 - 1; k >= k0; --k) {
        double A[d * d], Q[symn(d)], fm[d], fP[symn(d)];
        load_filtered<d>(io, k, fm, fP);
        typename Alg::Elem e, r;
        if (k == io.N - 1 && is_last) {  // last_parallel_smoothing_element, ops.py:314-315
#pragma unroll
            for (int i = 0; i < d * d; ++i) e.E[i] = 0.0;
#pragma unroll
            for (int i = 0; i < d; ++i) e.g[i] = fm[i];
#pragma unroll
            for (int i = 0; i < symn(d); ++i) e.L[i] = fP[i];
        } else {
            gen.step(k, A, Q);
            rts_element<d>(fm, fP, A, Q, e);
        }
        Alg::combine(acc, e, r);
        acc = r;
    }
    Alg::store(agg, nchunks, nchunks - 1 - c, acc);
}

template <class Gen>
__global__ void __launch_bounds__(kChunkThreads)
rts_reduce_kernel(Gen gen, RtsIO io, int L, long long nchunks, int is_last, double* agg) {
    const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nchunks) return;
    rts_reduce_chunk(gen, io, L, nchunks, is_last, agg, c);
}

// ---------------------------------------------------------------------------- scan form, phase 3
template <class Gen>
BN_DEV void rts_apply_chunk(const Gen& gen, const RtsIO& io, int L, long long nchunks, int is_last,
                            const double* prefix, const double* s0, long long c) {
    constexpr int d = Gen::d, Df = Gen::D;
    using Alg = SmootherAlg<d>;
    double H[Df * d];
    gen.H(H);
    const long long p = nchunks - 1 - c;
    typename Alg::State s;
    Alg::load_state(s0, 1, 0, s);
    if (p > 0) {
        typename Alg::Elem e;
        Alg::load(prefix, nchunks, p - 1, e);
        typename Alg::State t;
        Alg::apply(e, s, t);
        s = t;
    }
    const long long k0 = c * L, k1 = (k0 + L < io.N) ? k0 + L : io.N;
    for (long long k = k1 - 1; k >= k0; --k) {
        double A[d * d], Q[symn(d)], fm[d], fP[symn(d)], G[d * d], pm[d], pP[symn(d)];
        load_filtered<d>(io, k, fm, fP);
        gen.step(k, A, Q);
        rts_gain<d>(fm, fP, A, Q, G, pm, pP);
        if (k == io.N - 1 && is_last) {
#pragma unroll
            for (int i = 0; i < d; ++i) s.m[i] = fm[i];
#pragma unroll
            for (int i = 0; i < symn(d); ++i) s.P[i] = fP[i];
        } else {
            rts_step<d>(s.m, s.P, fm, fP, G, pm, pP);
        }
        write_smoothed<d, Df>(io, k, H, s.m, s.P, G);
    }
}

template <class Gen>
__global__ void __launch_bounds__(kChunkThreads)
rts_apply_kernel(Gen gen, RtsIO io, int L, long long nchunks, int is_last, const double* prefix, const double* s0) {
    const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nchunks) return;
    rts_apply_chunk(gen, io, L, nchunks, is_last, prefix, s0, c);
}

// ---------------------------------------------------------------------------- host drivers

template <class Gen>
inline int rts_run(const Gen& gen, int form, RtsIO io, void* ws, size_t ws_bytes, cudaStream_t st, int phase,
                   int is_last, double* carry_out, const double* carries, int rank, int world) {
    constexpr int d = Gen::d;
    using Alg = SmootherAlg<d>;
    if (io.N == 0) return 0;
    if (form == BN_SEQUENTIAL) {
        BN_LAUNCH("rts_seq", st, rts_seq_kernel<Gen><<<1, 1, 0, st>>>(gen, io));
        BN_CUDA(cudaGetLastError());
        return 0;
    }
    ChunkPlan cp = plan_chunks(io.N);
    size_t need = (64 + scan_plan_doubles(cp.nchunks, Alg::kElem)) * sizeof(double);
    BN_REQUIRE(ws != nullptr && ws_bytes >= need, "workspace too small: need %zu bytes, got %zu", need, ws_bytes);
    double* s0 = (double*)ws;
    ScanPlan plan = make_scan_plan(s0 + 64, cp.nchunks, Alg::kElem);
    unsigned grid = (unsigned)((cp.nchunks + kChunkThreads - 1) / kChunkThreads);
    if (phase == PHASE_ALL || phase == PHASE_REDUCE) {
        BN_LAUNCH("rts_reduce", st,
                  rts_reduce_kernel<Gen><<<grid, kChunkThreads, 0, st>>>(gen, io, cp.L, cp.nchunks, is_last,
                                                                          plan.input0));
        BN_CUDA(cudaGetLastError());
        BN_CUDA(run_scan<Alg>(plan, st));
        if (carry_out) {
            int top = plan.levels - 1;
            export_carry_kernel<Alg><<<1, 1, 0, st>>>(plan.prefix[top], plan.count[top], carry_out);
            BN_CUDA(cudaGetLastError());
        }
    }
    if (phase == PHASE_ALL || phase == PHASE_APPLY) {
        if (carries) {
            fold_carries_kernel<Alg><<<1, 1, 0, st>>>(carries, world - 1, rank, -1, s0);
            BN_CUDA(cudaGetLastError());
        } else {
            BN_CUDA(cudaMemsetAsync(s0, 0, Alg::kState * sizeof(double), st));
        }
        BN_LAUNCH("rts_apply", st,
                  rts_apply_kernel<Gen><<<grid, kChunkThreads, 0, st>>>(gen, io, cp.L, cp.nchunks, is_last,
                                                                         plan.prefix[0], s0));
        BN_CUDA(cudaGetLastError());
    }
    return 0;
}


struct RtsCall {
    int form;
    RtsIO io;
    void* ws;
    size_t ws_bytes;
    cudaStream_t st;
    int phase, is_last;
    double* carry_out;
    const double* carries;
    int rank, world;
    const bn_kernel_spec* spec;
    const double* dt;
    int d, Df;
    const double *As, *Qs, *H;
};


#define BN_RTS_SPEC_CASE(FAM, NC)                                                                          \
    if (c.spec->family == FAM && c.spec->n_components == NC) {                                             \
        MaternGen<FAM, NC> gen;                                                                            \
        gen.spec = *c.spec;                                                                                \
        gen.dt = c.dt;                                                                                     \
        return rts_run(gen, c.form, c.io, c.ws, c.ws_bytes, c.st, c.phase, c.is_last, c.carry_out,         \
                       c.carries, c.rank, c.world);                                                        \
    }

#define BN_RTS_ARR_CASE(DD, OD)                                                                            \
    if (c.d == DD && c.Df == OD) {                                                                         \
        ArrayGen<DD, OD> gen{c.As, c.Qs, c.H, nullptr, nullptr};                                           \
        return rts_run(gen, c.form, c.io, c.ws, c.ws_bytes, c.st, c.phase, c.is_last, c.carry_out,         \
                       c.carries, c.rank, c.world);                                                        \
    }

}  // namespace bn
