/*
 * bn_b200_xla.h -- XLA custom-call entry points of libbn_b200.so.
 *
 * The reference pins jax==0.4.14 (requirements.txt:1-2, setup.py:11-12), which has no jax.ffi; its GPU
 * custom-call ABI is the header-free legacy signature
 *
 *     void target(cudaStream_t stream, void** buffers, const char* opaque, size_t opaque_len);
 *
 * with `buffers` = the operands followed by the results of the HLO custom-call, all device pointers
 * owned by XLA, and `opaque` = a byte string fixed at lowering time.  Each wrapper below unpacks a POD
 * descriptor from `opaque` and forwards to the C ABI of bn_b200.h on XLA's stream; the LAST result of every
 * call is a byte workspace XLA allocates (its size is the descriptor's workspace_bytes, obtained at lowering
 * time from bn_update_posterior_workspace_bytes / bn_workspace_bytes).  Nothing is allocated or synchronised.
 * The legacy ABI has no status channel: on a bad descriptor the wrapper records the message for
 * bn_last_error(), leaves the results untouched and bumps bn_xla_error_count().
 *
 * Binding on the reference side: wrap each pointer in a PyCapsule named
 * "xla._CUSTOM_CALL_TARGET" and pass it to jax.lib.xla_client.register_custom_call_target(name, capsule,
 * platform="CUDA"); the ops.py entry points (kalman_filter ops.py:256, rauch_tung_striebel_smoother :357)
 * and MarkovGaussianProcess.update_posterior (basemodels.py:689-706) then lower to one custom-call each.
 */
#ifndef BN_B200_XLA_H
#define BN_B200_XLA_H

#include "bn_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

#define BN_XLA_MAX_CUBATURE 1200   /* doubles of cubature nodes the site descriptor can carry: [D,Q], Q <= 400 at D = 2 */

/* descriptor of the filter / smoother / fused-update calls */
typedef struct {
    bn_kernel_spec spec;
    int64_t N;
    uint64_t workspace_bytes;
    int32_t form;            /* BN_SEQUENTIAL / BN_SCAN (ignored by the fused update: scan form) */
    int32_t has_mask;        /* operand list carries a mask [N,D] bytes */
    int32_t want_grad;       /* update_posterior: results carry dell_dvariance, dell_dlengthscale */
    int32_t return_predict;  /* kalman_filter (ops.py:175-178) */
    int32_t return_full;     /* rauch_tung_striebel_smoother (ops.py:303-309) */
    int32_t pad_;
} bn_xla_markov_desc;

/* descriptor of the site calls: the scalar fields of bn_site_args plus the cubature rule by value */
typedef struct {
    int32_t method, likelihood;
    double lik_param;
    int64_t N;
    int32_t D, Q;
    double lr, power;
    int32_t ensure_psd, has_mask;
    uint64_t workspace_bytes;
    double cub_w[400];
    double cub_x[BN_XLA_MAX_CUBATURE];
    double lik_param2;       /* second likelihood parameter (bn_site_args.lik_param2) */
} bn_xla_site_desc;

/* operands: dt[N], pseudo_y[N,D,1], pseudo_var[N,D,D], (mask[N,D,1] u8)
 * results : ell[], post_mean[N,D,1], post_cov[N,D,D], (dell_dvariance[NC], dell_dlengthscale[NC]), workspace */
void bn_xla_update_posterior(void* stream, void** buffers, const char* opaque, size_t opaque_len);
/* operands: dt[N], y[N,D,1], noise_cov[N,D,D], (mask)      results: ell[], means[N,d,1], covs[N,d,d], workspace */
void bn_xla_kalman_filter(void* stream, void** buffers, const char* opaque, size_t opaque_len);
/* operands: dt[N], filter_mean[N,d,1], filter_cov[N,d,d]    results: means, covs, gains[N,d,d], workspace */
void bn_xla_rts_smoother(void* stream, void** buffers, const char* opaque, size_t opaque_len);
/* operands: y[N], post_mean[N,D,1], post_cov[N,D,D], nat1[N,D,1], nat2[N,D,D]
 * results : nat1_new, nat2_new, site_mean[N,D,1], site_cov[N,D,D], diffs[2], workspace   (XLA is functional: the
 *           old naturals are copied to the new buffers, then updated in place) */
void bn_xla_site_update(void* stream, void** buffers, const char* opaque, size_t opaque_len);
/* operands: y[N], post_mean, post_cov, nat1, nat2           results: sum[], workspace */
void bn_xla_expected_density(void* stream, void** buffers, const char* opaque, size_t opaque_len);
/* operands: pseudo_y, post_mean, post_cov, pseudo_var, (mask)   results: sum[], workspace   (desc: site desc, N/D/has_mask used) */
void bn_xla_gaussian_expected_log_lik(void* stream, void** buffers, const char* opaque, size_t opaque_len);

/* registration table: names[i] / targets[i] for i < return value (the Python side builds the capsules) */
int bn_xla_targets(const char** names, void** targets, int max);
long bn_xla_error_count(void);

#ifdef __cplusplus
}
#endif
#endif
