// C ABI of the RTS smoother (include/bn_b200.h).  Kernels: smoother_impl.cuh; instantiated per
// generator group in smoother_m_*.cu / smoother_a_*.cu.
#include "smoother_impl.cuh"
#include "gd_impl.cuh"

namespace bn {
int rts_group_m_a(const RtsCall&);
int rts_group_m_b(const RtsCall&);
int rts_group_m_c(const RtsCall&);
int rts_group_m_d(const RtsCall&);
int rts_group_a_a(const RtsCall&);
int rts_group_a_b(const RtsCall&);
int rts_group_a_c(const RtsCall&);
int gd_rts_arrays(int form, const GdRts& a, void* ws, size_t ws_bytes, cudaStream_t st);

static int rts_dispatch(const RtsCall& c) {
    int r;
    if (c.spec) {
        if ((r = rts_group_m_a(c)) != kNotHandled) return r;
        if ((r = rts_group_m_b(c)) != kNotHandled) return r;
        if ((r = rts_group_m_c(c)) != kNotHandled) return r;
        if ((r = rts_group_m_d(c)) != kNotHandled) return r;
        set_error("unsupported kernel spec: family %d with %d components (use the array-level entry)",
                  c.spec->family, c.spec->n_components);
        return -1;
    }
    if ((r = rts_group_a_a(c)) != kNotHandled) return r;
    if ((r = rts_group_a_b(c)) != kNotHandled) return r;
    if ((r = rts_group_a_c(c)) != kNotHandled) return r;
    if (c.phase != PHASE_ALL) {
        set_error("unsupported (state dim, latent dim) = (%d, %d) for the time-sharded smoother", c.d, c.Df);
        return -1;
    }
    GdRts a{c.io.N, c.d, c.Df, c.io.fms, c.io.fPs, c.As, c.Qs, c.H, c.io.return_full, c.io.sms, c.io.sPs, c.io.gains};
    return gd_rts_arrays(c.form, a, c.ws, c.ws_bytes, c.st);
}
}  // namespace bn

using namespace bn;

static RtsCall make_call(int form, RtsIO io, void* ws, size_t ws_bytes, void* stream) {
    RtsCall c{};
    c.form = form;
    c.io = io;
    c.ws = ws;
    c.ws_bytes = ws_bytes;
    c.st = (cudaStream_t)stream;
    c.phase = PHASE_ALL;
    c.is_last = 1;
    c.world = 1;
    return c;
}

extern "C" int bn_rts_arrays(int form, int64_t N, int d, int Df, const double* fms, const double* fPs,
                             const double* As, const double* Qs, const double* H, int return_full, double* sms,
                             double* sPs, double* gains, void* workspace, size_t workspace_bytes, void* stream) {
    BN_REQUIRE(N >= 0, "N must be non-negative");
    BN_REQUIRE(form == BN_SEQUENTIAL || form == BN_SCAN, "unknown form %d", form);
    BN_REQUIRE(N == 0 || (fms && fPs && As && Qs && H && sms && sPs), "null array");
    RtsCall c = make_call(form, RtsIO{N, fms, fPs, sms, sPs, gains, return_full}, workspace, workspace_bytes, stream);
    c.d = d; c.Df = Df; c.As = As; c.Qs = Qs; c.H = H;
    return rts_dispatch(c);
}

extern "C" int bn_rts_smoother(const bn_kernel_spec* k, int form, int64_t N, const double* dt,
                               const double* filter_mean, const double* filter_cov, int return_full, double* means,
                               double* covs, double* gains, void* workspace, size_t workspace_bytes, void* stream) {
    BN_REQUIRE(k != nullptr, "kernel spec is null");
    BN_REQUIRE(N >= 0, "N must be non-negative");
    BN_REQUIRE(form == BN_SEQUENTIAL || form == BN_SCAN, "unknown form %d", form);
    BN_REQUIRE(N == 0 || (dt && filter_mean && filter_cov && means && covs), "null array");
    RtsCall c = make_call(form, RtsIO{N, filter_mean, filter_cov, means, covs, gains, return_full}, workspace,
                          workspace_bytes, stream);
    c.spec = k; c.dt = dt;
    return rts_dispatch(c);
}

extern "C" int bn_rts_shard_reduce(const bn_kernel_spec* k, int64_t N, int is_last, const double* dt,
                                   const double* filter_mean, const double* filter_cov, double* carry,
                                   void* workspace, size_t workspace_bytes, void* stream) {
    BN_REQUIRE(k != nullptr, "kernel spec is null");
    BN_REQUIRE(N > 0, "a time shard must hold at least one step");
    BN_REQUIRE(dt && filter_mean && filter_cov && carry, "null array");
    RtsCall c = make_call(BN_SCAN, RtsIO{N, filter_mean, filter_cov, nullptr, nullptr, nullptr, 0}, workspace,
                          workspace_bytes, stream);
    c.spec = k; c.dt = dt; c.phase = PHASE_REDUCE; c.is_last = is_last; c.carry_out = carry;
    return rts_dispatch(c);
}

extern "C" int bn_rts_shard_apply(const bn_kernel_spec* k, int64_t N, int rank, int world, const double* carries,
                                  const double* dt, const double* filter_mean, const double* filter_cov,
                                  int return_full, double* means, double* covs, double* gains, void* workspace,
                                  size_t workspace_bytes, void* stream) {
    BN_REQUIRE(k != nullptr, "kernel spec is null");
    BN_REQUIRE(N > 0, "a time shard must hold at least one step");
    BN_REQUIRE(rank >= 0 && rank < world, "rank %d outside world %d", rank, world);
    BN_REQUIRE(dt && filter_mean && filter_cov && carries && means && covs, "null array");
    RtsCall c = make_call(BN_SCAN, RtsIO{N, filter_mean, filter_cov, means, covs, gains, return_full}, workspace,
                          workspace_bytes, stream);
    c.spec = k; c.dt = dt; c.phase = PHASE_APPLY; c.is_last = (rank == world - 1); c.carries = carries;
    c.rank = rank; c.world = world;
    return rts_dispatch(c);
}
