// C ABI of the Kalman filter (include/bn_b200.h).  The kernels live in filter_impl.cuh and are
// instantiated per generator group in filter_m_*.cu / filter_a_*.cu.
#include <cstdlib>
#include "filter_impl.cuh"
#include "gd_impl.cuh"

namespace bn {
int kf_group_m_a(const KfCall&);
int kf_group_m_b(const KfCall&);
int kf_group_m_c(const KfCall&);
int kf_group_m_d(const KfCall&);
int kf_group_a_a(const KfCall&);
int kf_group_a_b(const KfCall&);
int kf_group_a_c(const KfCall&);
int kf_group_a_d(const KfCall&);
int gd_kf_arrays(int form, const GdKf& a, double* ell, void* ws, size_t ws_bytes, cudaStream_t st);

static int kf_dispatch(const KfCall& c) {
    int r;
    if (c.spec) {
        if ((r = kf_group_m_a(c)) != kNotHandled) return r;
        if ((r = kf_group_m_b(c)) != kNotHandled) return r;
        if ((r = kf_group_m_c(c)) != kNotHandled) return r;
        if ((r = kf_group_m_d(c)) != kNotHandled) return r;
        set_error("unsupported kernel spec: family %d with %d components (use the array-level entry)",
                  c.spec->family, c.spec->n_components);
        return -1;
    }
    // (6, 6) -- the pairs filter of the sparse Markov model at Matern-5/2 -- has a register-resident instantiation, but its
    // 90-double scan element spills (1 ms per scan level at 25 000 chunks): the warp-cooperative path keeps the element in
    // shared memory and is ~7x faster end to end.  BN_B200_PAIRS_REGISTERS=1 selects the old instantiation (A/B aid).
    if (c.d == 6 && c.D == 6 && c.phase == PHASE_ALL && !(getenv("BN_B200_PAIRS_REGISTERS") && getenv("BN_B200_PAIRS_REGISTERS")[0] == '1')) {
        GdKf a{c.io.N, c.d, c.D, c.As, c.Qs, c.H, c.io.y, c.io.R, c.m0, c.P0, c.io.mask, c.io.return_predict, c.io.fms, c.io.fPs};
        return gd_kf_arrays(c.form, a, c.ell, c.ws, c.ws_bytes, c.st);
    }
    if ((r = kf_group_a_a(c)) != kNotHandled) return r;
    if ((r = kf_group_a_b(c)) != kNotHandled) return r;
    if ((r = kf_group_a_c(c)) != kNotHandled) return r;
    if ((r = kf_group_a_d(c)) != kNotHandled) return r;
    // any other (d, D) with D <= d <= 16: the warp-cooperative path (gd.cu), single shard
    if (c.phase != PHASE_ALL) {
        set_error("unsupported (state dim, obs dim) = (%d, %d) for the time-sharded filter", c.d, c.D);
        return -1;
    }
    GdKf a{c.io.N, c.d, c.D, c.As, c.Qs, c.H, c.io.y, c.io.R, c.m0, c.P0, c.io.mask, c.io.return_predict, c.io.fms, c.io.fPs};
    return gd_kf_arrays(c.form, a, c.ell, c.ws, c.ws_bytes, c.st);
}
}  // namespace bn

using namespace bn;

static KfCall make_call(int form, KfIO io, double* ell, void* ws, size_t ws_bytes, void* stream) {
    KfCall c{};
    c.form = form;
    c.io = io;
    c.ell = ell;
    c.ws = ws;
    c.ws_bytes = ws_bytes;
    c.st = (cudaStream_t)stream;
    c.phase = PHASE_ALL;
    c.is_first = 1;
    return c;
}

extern "C" int bn_kf_arrays(int form, int64_t N, int d, int D, const double* As, const double* Qs, const double* H,
                            const double* ys, const double* Rs, const double* m0, const double* P0,
                            const uint8_t* masks, int return_predict, double* ell, double* fms, double* fPs,
                            void* workspace, size_t workspace_bytes, void* stream) {
    BN_REQUIRE(N >= 0, "N must be non-negative");
    BN_REQUIRE(form == BN_SEQUENTIAL || form == BN_SCAN, "unknown form %d", form);
    BN_REQUIRE((fms == nullptr) == (fPs == nullptr), "fms and fPs must both be given or both be null");
    BN_REQUIRE(N == 0 || (As && Qs && H && ys && Rs && m0 && P0), "null input array");
    KfCall c = make_call(form, KfIO{N, ys, Rs, masks, fms, fPs, return_predict}, ell, workspace, workspace_bytes,
                         stream);
    c.d = d; c.D = D; c.As = As; c.Qs = Qs; c.H = H; c.m0 = m0; c.P0 = P0;
    return kf_dispatch(c);
}

extern "C" int bn_kalman_filter(const bn_kernel_spec* k, int form, int64_t N, const double* dt, const double* y,
                                const double* noise_cov, const uint8_t* mask, int return_predict, double* ell,
                                double* means, double* covs, void* workspace, size_t workspace_bytes, void* stream) {
    BN_REQUIRE(k != nullptr, "kernel spec is null");
    BN_REQUIRE(N >= 0, "N must be non-negative");
    BN_REQUIRE(form == BN_SEQUENTIAL || form == BN_SCAN, "unknown form %d", form);
    BN_REQUIRE((means == nullptr) == (covs == nullptr), "means and covs must both be given or both be null");
    BN_REQUIRE(N == 0 || (dt && y && noise_cov), "null input array");
    KfCall c = make_call(form, KfIO{N, y, noise_cov, mask, means, covs, return_predict}, ell, workspace,
                         workspace_bytes, stream);
    c.spec = k; c.dt = dt;
    return kf_dispatch(c);
}

extern "C" int bn_kf_shard_reduce(const bn_kernel_spec* k, int64_t N, int is_first, const double* dt, const double* y,
                                  const double* noise_cov, double* carry, void* workspace, size_t workspace_bytes,
                                  void* stream) {
    BN_REQUIRE(k != nullptr, "kernel spec is null");
    BN_REQUIRE(N > 0, "a time shard must hold at least one step");
    BN_REQUIRE(dt && y && noise_cov && carry, "null array");
    KfCall c = make_call(BN_SCAN, KfIO{N, y, noise_cov, nullptr, nullptr, nullptr, 0}, nullptr, workspace,
                         workspace_bytes, stream);
    c.spec = k; c.dt = dt; c.phase = PHASE_REDUCE; c.is_first = is_first; c.carry_out = carry;
    return kf_dispatch(c);
}

extern "C" int bn_kf_shard_apply(const bn_kernel_spec* k, int64_t N, int rank, int world, const double* carries,
                                 const double* dt, const double* y, const double* noise_cov, const uint8_t* mask,
                                 int return_predict, double* ell, double* means, double* covs, void* workspace,
                                 size_t workspace_bytes, void* stream) {
    BN_REQUIRE(k != nullptr, "kernel spec is null");
    BN_REQUIRE(N > 0, "a time shard must hold at least one step");
    BN_REQUIRE(rank >= 0 && rank < world, "rank %d outside world %d", rank, world);
    BN_REQUIRE(dt && y && noise_cov && carries, "null array");
    BN_REQUIRE((means == nullptr) == (covs == nullptr), "means and covs must both be given or both be null");
    KfCall c = make_call(BN_SCAN, KfIO{N, y, noise_cov, mask, means, covs, return_predict}, ell, workspace,
                         workspace_bytes, stream);
    c.spec = k; c.dt = dt; c.phase = PHASE_APPLY; c.is_first = (rank == 0); c.carries = carries; c.rank = rank;
    return kf_dispatch(c);
}

extern "C" int bn_kf_carry_len(int d) { return 3 * d * d + 2 * d; }
