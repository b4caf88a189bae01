// Legacy XLA GPU custom-call wrappers (include/bn_b200_xla.h): unpack the opaque descriptor, forward to the
// C ABI on XLA's stream.  jax 0.4.14 (the reference's pin) registers these with
// xla_client.register_custom_call_target(name, capsule, platform="CUDA").
#include <atomic>
#include <cstring>
#include "common.cuh"
#include "../../include/bn_b200_xla.h"

namespace bn {
static std::atomic<long> g_xla_errors{0};

template <class Desc>
static bool unpack(const char* opaque, size_t len, Desc& d, const char* who) {
    if (opaque == nullptr || len != sizeof(Desc)) {
        set_error("%s: opaque descriptor of %zu bytes, expected %zu", who, len, sizeof(Desc));
        ++g_xla_errors;
        return false;
    }
    memcpy(&d, opaque, sizeof(Desc));
    return true;
}
static void note(int rc) {
    if (rc != 0) ++g_xla_errors;
}
static void fill_site(const bn_xla_site_desc& d, bn_site_args& a) {
    memset(&a, 0, sizeof(a));
    a.method = d.method; a.likelihood = d.likelihood; a.lik_param = d.lik_param;
    a.lik_param2 = d.lik_param2;
    a.N = d.N; a.D = d.D; a.Q = d.Q;
    a.cub_x = d.cub_x; a.cub_w = d.cub_w;
    a.lr = d.lr; a.power = d.power; a.ensure_psd = d.ensure_psd;
}
}  // namespace bn

using namespace bn;

extern "C" void bn_xla_update_posterior(void* stream, void** b, const char* opaque, size_t len) {
    bn_xla_markov_desc d;
    if (!unpack(opaque, len, d, "bn_xla_update_posterior")) return;
    int i = 0;
    const double* dt = (const double*)b[i++];
    const double* y = (const double*)b[i++];
    const double* R = (const double*)b[i++];
    const uint8_t* mask = d.has_mask ? (const uint8_t*)b[i++] : nullptr;
    double* ell = (double*)b[i++];
    double* pm = (double*)b[i++];
    double* pc = (double*)b[i++];
    if (d.want_grad) {
        double* dv = (double*)b[i++];
        double* dl = (double*)b[i++];
        void* ws = b[i++];
        if (mask) {
            set_error("bn_xla_update_posterior: the hyper-gradient is not available with a mask");
            ++g_xla_errors;
            return;
        }
        note(bn_update_posterior_grad(&d.spec, d.N, dt, y, R, ell, pm, pc, dv, dl, ws, d.workspace_bytes, stream));
    } else {
        void* ws = b[i++];
        note(bn_update_posterior(&d.spec, d.N, dt, y, R, mask, ell, pm, pc, ws, d.workspace_bytes, stream));
    }
}

extern "C" void bn_xla_kalman_filter(void* stream, void** b, const char* opaque, size_t len) {
    bn_xla_markov_desc d;
    if (!unpack(opaque, len, d, "bn_xla_kalman_filter")) return;
    int i = 0;
    const double* dt = (const double*)b[i++];
    const double* y = (const double*)b[i++];
    const double* R = (const double*)b[i++];
    const uint8_t* mask = d.has_mask ? (const uint8_t*)b[i++] : nullptr;
    double* ell = (double*)b[i++];
    double* means = (double*)b[i++];
    double* covs = (double*)b[i++];
    void* ws = b[i++];
    note(bn_kalman_filter(&d.spec, d.form, d.N, dt, y, R, mask, d.return_predict, ell, means, covs, ws,
                          d.workspace_bytes, stream));
}

extern "C" void bn_xla_rts_smoother(void* stream, void** b, const char* opaque, size_t len) {
    bn_xla_markov_desc d;
    if (!unpack(opaque, len, d, "bn_xla_rts_smoother")) return;
    note(bn_rts_smoother(&d.spec, d.form, d.N, (const double*)b[0], (const double*)b[1], (const double*)b[2],
                         d.return_full, (double*)b[3], (double*)b[4], (double*)b[5], b[6], d.workspace_bytes, stream));
}

extern "C" void bn_xla_site_update(void* stream, void** b, const char* opaque, size_t len) {
    static thread_local bn_xla_site_desc d;  // 13 kB: keep it off the stack of XLA's dispatch thread
    if (!unpack(opaque, len, d, "bn_xla_site_update")) return;
    bn_site_args a;
    fill_site(d, a);
    a.y = (const double*)b[0];
    a.post_mean = (const double*)b[1];
    a.post_cov = (const double*)b[2];
    const size_t n1 = (size_t)d.N * d.D * sizeof(double), n2 = n1 * d.D;
    a.nat1 = (double*)b[5];
    a.nat2 = (double*)b[6];
    if (cudaMemcpyAsync(a.nat1, b[3], n1, cudaMemcpyDeviceToDevice, (cudaStream_t)stream) != cudaSuccess ||
        cudaMemcpyAsync(a.nat2, b[4], n2, cudaMemcpyDeviceToDevice, (cudaStream_t)stream) != cudaSuccess) {
        set_error("bn_xla_site_update: copying the natural parameters failed");
        ++g_xla_errors;
        return;
    }
    a.site_mean = (double*)b[7];
    a.site_cov = (double*)b[8];
    a.diffs = (double*)b[9];
    note(bn_site_update(&a, b[10], d.workspace_bytes, stream));
}

extern "C" void bn_xla_expected_density(void* stream, void** b, const char* opaque, size_t len) {
    static thread_local bn_xla_site_desc d;
    if (!unpack(opaque, len, d, "bn_xla_expected_density")) return;
    bn_site_args a;
    fill_site(d, a);
    a.y = (const double*)b[0];
    a.post_mean = (const double*)b[1];
    a.post_cov = (const double*)b[2];
    a.nat1 = (double*)b[3];   // read only on this path
    a.nat2 = (double*)b[4];
    note(bn_expected_density(&a, nullptr, (double*)b[5], b[6], d.workspace_bytes, stream));
}

extern "C" void bn_xla_gaussian_expected_log_lik(void* stream, void** b, const char* opaque, size_t len) {
    static thread_local bn_xla_site_desc d;
    if (!unpack(opaque, len, d, "bn_xla_gaussian_expected_log_lik")) return;
    int i = 4;
    const uint8_t* mask = d.has_mask ? (const uint8_t*)b[i++] : nullptr;
    double* sum = (double*)b[i++];
    void* ws = b[i++];
    note(bn_gaussian_expected_log_lik(d.N, d.D, (const double*)b[0], (const double*)b[1], (const double*)b[2],
                                      (const double*)b[3], mask, nullptr, sum, ws, d.workspace_bytes, stream));
}

extern "C" int bn_xla_targets(const char** names, void** targets, int max) {
    static const char* kNames[] = {"bn_xla_update_posterior", "bn_xla_kalman_filter", "bn_xla_rts_smoother",
                                   "bn_xla_site_update", "bn_xla_expected_density", "bn_xla_gaussian_expected_log_lik"};
    void* kTargets[] = {(void*)bn_xla_update_posterior, (void*)bn_xla_kalman_filter, (void*)bn_xla_rts_smoother,
                        (void*)bn_xla_site_update, (void*)bn_xla_expected_density,
                        (void*)bn_xla_gaussian_expected_log_lik};
    const int n = 6;
    for (int i = 0; i < n && i < max; ++i) {
        if (names) names[i] = kNames[i];
        if (targets) targets[i] = kTargets[i];
    }
    return n;
}

extern "C" long bn_xla_error_count(void) { return g_xla_errors.load(); }
