/*
 * bn_b200.h -- C ABI of the B200-native Markov-GP inference hot path.
 *
 * Drop-in boundary for the one path of AaltoML/BayesNewton this library
 * replaces.  The reference has no FFI of its own (it is pure
 * Python on JAX); each entry point below names the reference function whose
 * array-level contract it reproduces (paths relative to the reference root):
 *
 *   bn_discretise          vmap(kernel.state_transition)(dt) + process_noise_covariance
 *                          bayesnewton/ops.py:149-151,274-278; kernels.py:158-165,216-224,273-286,344-365
 *   bn_kf_arrays           _sequential_kf / _parallel_kf          ops.py:154-180, 237-253
 *   bn_rts_arrays          _sequential_rts / _parallel_rts        ops.py:288-311, 338-354
 *   bn_kalman_filter       kalman_filter                          ops.py:256-285
 *   bn_rts_smoother        rauch_tung_striebel_smoother           ops.py:357-380
 *   bn_kf_shard_*          the same filter, split in the three phases a time-sharded
 *   bn_rts_shard_*         (multi-GPU) two-level scan needs       ops.py:203-219, 328-335
 *   bn_update_posterior    MarkovGaussianProcess.update_posterior  basemodels.py:689-706 (filter + smoother fused)
 *   bn_up_shard_*          the same, on one time shard of a multi-GPU run
 *   bn_update_posterior_grad   + d compute_log_lik / d kernel hyper-parameters (the reverse-mode pass of
 *                          objax.GradValues(model.energy, ...), README.md:56-70; basemodels.py:726-741)
 *   bn_site_update         update_variational_params + newton_update + damped update_nat_params
 *                          inference.py:21-39,65-90,105-128,170-195,238-284,339-371; basemodels.py:85-100
 *   bn_expected_density    the value-only likelihood term of energy()
 *                          inference.py:130-154,197-222,286-325,373-428
 *   bn_gaussian_expected_log_lik   vmap(gaussian_expected_log_lik) utils.py:510-531, basemodels.py:715-721
 *   bn_st_kalman_filter / bn_st_rts_smoother   the same two ops for SpatioTemporalKernel (dense d = M n state,
 *                          kernels.py:385-586), and bn_st_pseudo_lik / bn_st_posterior_to_data /
 *                          bn_st_gaussian_expected_log_lik / bn_spd_inverse_batched: the projection steps either
 *                          side of them (basemodels.py:676-687, 743-764, 708-724; utils.py:30-35)
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in _host or its comment says HOST; the caller owns
 *     all buffers including outputs and the workspace; nothing is allocated, freed or
 *     retained; calls are asynchronous on `stream` (a cudaStream_t passed as void*).
 *   - arrays are contiguous row-major with the time axis leading, exactly as in the
 *     reference: means [N,d,1] -> N*d doubles, covariances [N,d,d] -> N*d*d doubles,
 *     masks [N,D,1] bool -> N*D bytes (non-zero = missing).
 *   - dtype is IEEE double (the reference runs with jax_enable_x64, basemodels.py:46-47).
 *   - return value: 0 ok, <0 bad argument (see bn_last_error), >0 a cudaError_t.
 *     Numerical failure is not an error: a non-PD Cholesky yields NaN as in JAX.
 *   - form: BN_SEQUENTIAL is the reference's lax.scan recursion executed in time order
 *     by a single GPU thread (bit-for-bit the sequential rounding order);
 *     BN_SCAN is the temporally-parallel form (`parallel=True`).
 */
#ifndef BN_B200_H
#define BN_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BN_MAX_COMPONENTS 4

enum { BN_SEQUENTIAL = 0, BN_SCAN = 1 };

/* kernel families with a closed-form discretisation (kernels.py:123-382) */
enum { BN_MATERN12 = 1, BN_MATERN32 = 2, BN_MATERN52 = 3, BN_MATERN72 = 4 };

/* A stationary prior = `Independent` stack (kernels.py:1499-1616) of n_components Matern
 * kernels of ONE family (n_components = 1 is the plain kernel).  State dim d =
 * n_components * family order dim, latent dim D = n_components, H = blockdiag([1,0,..]). */
typedef struct {
    int32_t family;
    int32_t n_components;
    double variance[BN_MAX_COMPONENTS];
    double lengthscale[BN_MAX_COMPONENTS];
} bn_kernel_spec;

/* likelihoods of the site kernels (likelihoods.py:684-860, 891-1008, 1011-1189, 1244-1281).
 *   lik_param  = Gaussian variance | Poisson bin size | Student-t scale | Gamma shape | NegBin alpha | Beta scale
 *   lik_param2 = Student-t degrees of freedom | NegBin scale                         (0 otherwise) */
enum { BN_LIK_GAUSSIAN = 1, BN_LIK_BERNOULLI_PROBIT = 2, BN_LIK_BERNOULLI_LOGIT = 3,
       BN_LIK_HETEROSCEDASTIC_SOFTPLUS = 4, BN_LIK_HETEROSCEDASTIC_EXP = 5, BN_LIK_POISSON_EXP = 6,
       BN_LIK_STUDENTS_T = 7, BN_LIK_GAMMA_EXP = 8, BN_LIK_NEGBIN_EXP = 9, BN_LIK_BETA_PROBIT = 10 };

/* inference schemes (inference.py:99-428) */
enum { BN_METHOD_VI = 1, BN_METHOD_EP = 2, BN_METHOD_NEWTON = 3, BN_METHOD_PL = 4 };

const char* bn_last_error(void);
int bn_version(void);

/* ---- optional per-kernel device timing (measurement aid, used by bench.py) --------------------- */
/* bn_timing_enable(1) clears the log and brackets every kernel launch of this library with CUDA
 * events on its stream; bn_timing_enable(0) stops.  bn_timing_report writes "name count total_ms"
 * lines (one per kernel name) into buf and returns the size the full report needs. */
int bn_timing_enable(int on);
int bn_timing_report(char* buf_host, size_t len);
/* fp64 FMA rate of the current device (DFMA per second, 8 chains/thread, full occupancy): the
 * denominator of the fp64 roofline bench.py reports beside the HBM one.  scratch: >= 8*SMs*256 doubles. */
int bn_measure_dfma_peak(double* scratch, size_t scratch_doubles, double* dfma_per_s_host);
/* the same for the fp64 tensor pipe (mma.sync.m8n8k4.f64 = DMMA.8x8x4): fused multiply-adds per second; the roofline
 * denominator of the dense spatio-temporal kernels (bench.py --workload C4) */
int bn_measure_dmma_peak(double* scratch, size_t scratch_doubles, double* fma_per_s_host);

/* ---- discretisation: As[N,d,d], Qs[N,d,d] from dt[N] -------------------------------------- */
int bn_state_dim(const bn_kernel_spec* k);
int bn_discretise(const bn_kernel_spec* k, int64_t N, const double* dt, double* As, double* Qs, void* stream);

/* ---- workspace ------------------------------------------------------------------------- */
/* bytes of caller-provided scratch any filter/smoother call below needs for (N, d, D) */
size_t bn_workspace_bytes(int64_t N, int d, int D);

/* ---- array-level filter / smoother ----------------------------------------------------- */
/* H[D,d]; ys[N,D,1]; Rs[N,D,D]; m0[d,1]; P0[d,d]; masks[N,D,1] (nullable = nothing missing).
 * ell: one double (nullable: skip the log-likelihood); fms[N,d,1], fPs[N,d,d] (both nullable
 * together: log-likelihood only).  return_predict as ops.py:175-178. */
int bn_kf_arrays(int form, int64_t N, int d, int D,
                 const double* As, const double* Qs, const double* H,
                 const double* ys, const double* Rs, const double* m0, const double* P0,
                 const uint8_t* masks, int return_predict,
                 double* ell, double* fms, double* fPs,
                 void* workspace, size_t workspace_bytes, void* stream);

/* return_full=0: sms[N,Df,1] = H sm, sPs[N,Df,Df] = H sP H^T; return_full=1: sms[N,d,1], sPs[N,d,d].
 * gains[N,d,d] nullable (update_posterior discards them, basemodels.py:701). */
int bn_rts_arrays(int form, int64_t N, int d, int Df,
                  const double* fms, const double* fPs,
                  const double* As, const double* Qs, const double* H, int return_full,
                  double* sms, double* sPs, double* gains,
                  void* workspace, size_t workspace_bytes, void* stream);

/* ---- kernel-level filter / smoother: A_k, Q_k generated in-kernel from dt --------------- */
int bn_kalman_filter(const bn_kernel_spec* k, int form, int64_t N,
                     const double* dt, const double* y, const double* noise_cov, const uint8_t* mask,
                     int return_predict, double* ell, double* means, double* covs,
                     void* workspace, size_t workspace_bytes, void* stream);

int bn_rts_smoother(const bn_kernel_spec* k, int form, int64_t N,
                    const double* dt, const double* filter_mean, const double* filter_cov, int return_full,
                    double* means, double* covs, double* gains,
                    void* workspace, size_t workspace_bytes, void* stream);

/* ---- fused posterior update: filter + smoother in one call --------------------------------- */
/* MarkovGaussianProcess.update_posterior (basemodels.py:689-706): kalman_filter(dt, kernel, pseudo_y,
 * pseudo_var, mask, parallel=True) followed by rauch_tung_striebel_smoother(dt', kernel, fm, fP,
 * parallel=True); dt' = [dt[1:], 0] is formed internally.  post_mean[N,D,1] = H sm, post_cov[N,D,D] =
 * H sP H^T.  ell (nullable) receives the filter log-likelihood, the quantity compute_log_lik()
 * (basemodels.py:726-741) recomputes from the same inputs.  The filtered states never leave the
 * workspace and the smoother's per-chunk elements are derived from the filter's chunk elements, so
 * one call moves 216 B/step at d = 3 instead of 368 B/step for the two stand-alone calls. */
size_t bn_update_posterior_workspace_bytes(const bn_kernel_spec* k, int64_t N);
int bn_update_posterior(const bn_kernel_spec* k, int64_t N, const double* dt,
                        const double* pseudo_y, const double* pseudo_var, const uint8_t* mask,
                        double* ell, double* post_mean, double* post_cov,
                        void* workspace, size_t workspace_bytes, void* stream);
/* The same update plus the hyper-parameter gradient of the filter log-likelihood:
 *   dell_dvariance[n_components], dell_dlengthscale[n_components] = d ell / d (variance_c, lengthscale_c)
 * of the UNTRANSFORMED hyper-parameters (the host chains the softplus of kernels.py:80-95).  This is the
 * only route from the kernel hyper-parameters to energy() in a temporal model, i.e. what
 * objax.GradValues(model.energy, model.vars()) (README.md:56-70, demos/regression.py:63-70) back-propagates
 * through compute_log_lik (basemodels.py:726-741) and the lax.scan of ops.py:154-180: d energy = -d ell.
 * The adjoint of the predicted state of each step is formed in closed form from the smoothed state inside the
 * smoother sweep (csrc/fast_core.cuh), so the gradient adds arithmetic but no HBM traffic.  No mask: the
 * reference's mask rule (utils.py:376-396) drops masked densities from ell but keeps their updates, which this
 * identity does not cover. */
int bn_update_posterior_grad(const bn_kernel_spec* k, int64_t N, const double* dt,
                             const double* pseudo_y, const double* pseudo_var,
                             double* ell, double* post_mean, double* post_cov,
                             double* dell_dvariance, double* dell_dlengthscale,
                             void* workspace, size_t workspace_bytes, void* stream);
/* The same update on a time shard (one rank of a multi-GPU run), in three phases around two carry
 * all-gathers.  All three calls of one update must be given the SAME workspace (it keeps the chunk
 * elements and the filtered states alive between phases).  dt here is the shard's slice of the
 * global dt (no shifted copy, no halo: the step across a shard boundary belongs to the right shard).
 *   reduce : local steps -> one filtering carry [bn_kf_carry_len(d)]
 *   filter : kf_carries[world] -> local filter pass, local log-likelihood partial (nullable),
 *            one smoothing carry [bn_rts_carry_len(d)] mapping the state at this shard's last step
 *            to the state at the previous shard's last step (closed by the terminal element on the
 *            last rank, ops.py:314-315)
 *   smooth : rts_carries[world] -> local smoother pass, post_mean / post_cov of the local steps
 * want_grad (reduce, filter) and non-null dell_* (smooth: this shard's share of the hyper-gradient, to be
 * summed over ranks) select the gradient-carrying chunk plan and must agree across the three calls. */
int bn_up_shard_reduce(const bn_kernel_spec* k, int64_t N, int rank, int world, const double* dt,
                       const double* pseudo_y, const double* pseudo_var, double* kf_carry, int want_grad,
                       void* workspace, size_t workspace_bytes, void* stream);
int bn_up_shard_filter(const bn_kernel_spec* k, int64_t N, int rank, int world, const double* kf_carries,
                       const double* dt, const double* pseudo_y, const double* pseudo_var, const uint8_t* mask,
                       double* ell, double* rts_carry, int want_grad,
                       void* workspace, size_t workspace_bytes, void* stream);
int bn_up_shard_smooth(const bn_kernel_spec* k, int64_t N, int rank, int world, const double* rts_carries,
                       const double* dt, double* post_mean, double* post_cov,
                       double* dell_dvariance, double* dell_dlengthscale,
                       void* workspace, size_t workspace_bytes, void* stream);

/* ---- time-sharded (multi-GPU) scan: reduce -> exchange carries -> apply ------------------ */
/* number of doubles in one filtering carry (A,b,C,J,eta full storage: 3d^2+2d) / smoothing carry
 * (E,g,L: 2d^2+d) -- the O(d^2) messages ranks all-gather. */
int bn_kf_carry_len(int d);
int bn_rts_carry_len(int d);
/* phase 1: reduce this rank's N local steps to one carry (device, bn_kf_carry_len doubles).
 * is_first != 0 on the rank that owns global step 0 (the Q_0 := P_0 rule of ops.py:222-229). */
int bn_kf_shard_reduce(const bn_kernel_spec* k, int64_t N, int is_first,
                       const double* dt, const double* y, const double* noise_cov,
                       double* carry, void* workspace, size_t workspace_bytes, void* stream);
/* phase 2+3: given the carries of all ranks (carries[world, carry_len], device) and this rank's
 * index, fold the lower ranks' carries into the incoming state and run the local filter.
 * ell receives this rank's partial log-likelihood (sum over its steps). */
int bn_kf_shard_apply(const bn_kernel_spec* k, int64_t N, int rank, int world, const double* carries,
                      const double* dt, const double* y, const double* noise_cov, const uint8_t* mask,
                      int return_predict, double* ell, double* means, double* covs,
                      void* workspace, size_t workspace_bytes, void* stream);
/* smoother: dt is the shifted step array of THIS shard (dt'[k] = step out of k; the last entry of
 * the last rank is 0, every other rank's last entry is its right neighbour's first dt). */
int bn_rts_shard_reduce(const bn_kernel_spec* k, int64_t N, int is_last,
                        const double* dt, const double* filter_mean, const double* filter_cov,
                        double* carry, void* workspace, size_t workspace_bytes, void* stream);
int bn_rts_shard_apply(const bn_kernel_spec* k, int64_t N, int rank, int world, const double* carries,
                       const double* dt, const double* filter_mean, const double* filter_cov, int return_full,
                       double* means, double* covs, double* gains,
                       void* workspace, size_t workspace_bytes, void* stream);

/* ---- sites ------------------------------------------------------------------------------ */
typedef struct {
    int32_t method;        /* BN_METHOD_* */
    int32_t likelihood;    /* BN_LIK_* */
    double lik_param;      /* Gaussian: observation variance */
    int64_t N;
    int32_t D;             /* latent dim per step (1, or 2 for the heteroscedastic likelihood) */
    int32_t Q;             /* number of cubature points */
    const double* cub_x;   /* HOST pointer: [D,Q] sigma points of the unit Gaussian (cubature.py:76-84) */
    const double* cub_w;   /* HOST pointer: [Q] weights.  The rule is tiny and built on the host (numpy in the
                              reference too); it travels as a kernel parameter.  D = 1: Q <= 64. */
    const double* y;       /* [N] observations; NaN = missing */
    const double* post_mean;  /* [N,D,1] */
    const double* post_cov;   /* [N,D,D] */
    double lr;             /* damping (inference.py:83-86) */
    double power;          /* EP power */
    int32_t ensure_psd;    /* utils.py:89-96 */
    int32_t pad_;
    /* site natural parameters, read (old) and overwritten (new) in place */
    double* nat1;          /* [N,D,1] */
    double* nat2;          /* [N,D,D] */
    /* outputs (each nullable) */
    double* site_mean;     /* [N,D,1]  reparametrised new sites = the next filter's pseudo_y */
    double* site_cov;      /* [N,D,D]                                          pseudo_var */
    double* out_mean;      /* [N,D,1]  the (mean, jacobian, hessian) triple inference() returns */
    double* out_jac;       /* [N,D,1] */
    double* out_hess;      /* [N,D,D] */
    double* diffs;         /* [2] mean |delta nat1|, mean |delta nat2| (inference.py:79-80) */
    double lik_param2;     /* second likelihood parameter (see BN_LIK_*); appended: older callers leave it 0 */
} bn_site_args;

int bn_site_update(const bn_site_args* a, void* workspace, size_t workspace_bytes, void* stream);

/* Likelihood-level statistics evaluated AT the given (post_mean, post_cov), nothing around them --
 * what inference.py vmaps over N (likelihoods.py:336-355,363-383,401-412,613-675; cubature.py):
 *   VI     variational_expectation       -> val = E_q[log p],  d1 = dE/dm,   d2 = d2E/dm2
 *   EP     moment_match (mean/cov = cavity, a->power) -> val = log Z, d1 = dlZ/dm, d2 = d2lZ/dm2
 *   Newton log_likelihood_gradients at f = post_mean  -> val = log p, d1 = J, d2 = H
 *   PL     statistical_linear_regression (y unused)   -> val = mu,  d1 = dmu/dm, d2 = omega
 * val[N], d1[N,D,1], d2[N,D,D]; each nullable.  nat1/nat2/lr/ensure_psd are ignored. */
int bn_likelihood_stats(const bn_site_args* a, double* val, double* d1, double* d2,
                        void* workspace, size_t workspace_bytes, void* stream);

/* per-step value of the likelihood term of energy(): VI E_q[log p], Newton log p(y|m),
 * EP/PL log Z at the cavity (computed in-kernel from post + nat).  values[N] nullable;
 * sum: one double = nansum of the values (inference.py:218,321). */
int bn_expected_density(const bn_site_args* a, double* values, double* sum,
                        void* workspace, size_t workspace_bytes, void* stream);

/* sum_n d(likelihood term of energy())_n / d(lik_param) for the Gaussian likelihood (lik_param = its variance): the route
 * from the likelihood hyper-parameter to the energy in objax.GradValues(model.energy, model.vars()) (README.md:56-70;
 * posterior and sites are StateVars).  VI: E_q[log N(y | f, s2)] (likelihoods.py:727-753); Newton: log N(y | m, s2);
 * EP: log N(y | m_cav, s2 / power + v_cav) + pep_constant (likelihoods.py:755-782), cavity from post + nat.
 * Missing observations (NaN) contribute 0. */
int bn_likelihood_param_grad(const bn_site_args* a, double* sum, void* workspace, size_t workspace_bytes, void* stream);

/* The two per-step sums of a single-latent VI / Newton energy in ONE pass over the posterior marginals
 * (inference.py:130-154, 197-222): sums[0] = nansum_n of the scheme's likelihood term (as bn_expected_density),
 * sums[1] = sum_n gaussian_expected_log_lik(site_mean_n, post_mean_n, post_cov_n, site_cov_n, mask_n)
 * (as bn_gaussian_expected_log_lik; a->site_mean / a->site_cov are the pseudo observations). */
int bn_energy_terms(const bn_site_args* a, const uint8_t* mask, double* sums,
                    void* workspace, size_t workspace_bytes, void* stream);

/* sum_n gaussian_expected_log_lik(pseudo_y_n, post_mean_n, post_cov_n, pseudo_var_n, mask_n) */
int bn_gaussian_expected_log_lik(int64_t N, int D, const double* pseudo_y, const double* post_mean,
                                 const double* post_cov, const double* pseudo_var, const uint8_t* mask,
                                 double* values, double* sum,
                                 void* workspace, size_t workspace_bytes, void* stream);

/* EP energy helper: sum_n [ log N(pseudo_y | cav_mean, pseudo_var/power + cav_cov) + pep_constant ]
 * (basemodels.py:247-262; with_pep_constant=0 gives the PL variant, inference.py:406-411) */
int bn_ep_pseudo_density(int64_t N, int D, double power, int with_pep_constant,
                         const double* pseudo_y, const double* pseudo_var,
                         const double* post_mean, const double* post_cov,
                         const double* nat1, const double* nat2, const uint8_t* mask,
                         double* sum, void* workspace, size_t workspace_bytes, void* stream);

/* ---- dense spatio-temporal path ------------------------------------------
 * SpatioTemporalKernel (kernels.py:385-586): state x = [u_1; ..; u_M] with u_i the temporal state of spatial
 * inducing point i;  A_k = I_M (x) A_t(dt_k) (:543-551), Pinf = I_M (x) Pinf_t (:517-524), H = I_M (x) H_t (:534-541).
 * d = M * dim(temporal family), D = M.  `temporal` has n_components = 1 and family Matern-1/2, -3/2 or -5/2.
 * Time is sequential (lax.scan, ops.py:154-180, 288-311); each pass is ONE persistent kernel whose phases
 * (predict, blocked Cholesky with the right-hand sides stacked under it, A B^T tiles) are separated by grid barriers. */
size_t bn_st_workspace_bytes(const bn_kernel_spec* temporal, int M, int64_t N, int Ns);

/* kalman_filter(dt, kernel, y, noise_cov, mask=None, parallel=False, return_predict) for a SpatioTemporalKernel
 * (ops.py:256-285 -> _sequential_kf :154-180).  y[N,M,1], noise_cov[N,M,M] dense SPD.  mask[N,M,1] nullable: as in
 * the reference it only enters the log-likelihood (mvn_logpdf, utils.py:376-396), the update is not masked; the
 * reference passes one when M equals the number of observations per step (basemodels.py:136-137, 652-653).  ell nullable.
 * means[N,d,1], covs[N,d,d]. */
int bn_st_kalman_filter(const bn_kernel_spec* temporal, int M, int64_t N,
                        const double* dt, const double* y, const double* noise_cov, const uint8_t* mask,
                        int return_predict, double* ell, double* means, double* covs,
                        void* workspace, size_t workspace_bytes, void* stream);

/* rauch_tung_striebel_smoother(dt, kernel, filter_mean, filter_cov, return_full, parallel=False)
 * (ops.py:357-380 -> _sequential_rts :288-311); dt is the step OUT OF n (basemodels.py:700).
 * return_full=0: means[N,M,1] = H sm, covs[N,M,M] = H sP H^T;  =1: means[N,d,1], covs[N,d,d].  gains[N,d,d] nullable. */
int bn_st_rts_smoother(const bn_kernel_spec* temporal, int M, int64_t N,
                       const double* dt, const double* filter_mean, const double* filter_cov, int return_full,
                       double* means, double* covs, double* gains,
                       void* workspace, size_t workspace_bytes, void* stream);

/* measurement aid: SM cycles per phase of the LAST bn_st_kalman_filter / bn_st_rts_smoother launch, summed over time
 * steps (slots: 0 assemble, 1 its barrier, 2 Cholesky sweep, 3 tile phases, 4 their barriers, 5 panel solve, 6 look-ahead,
 * 7 barrier wait of the last row block's owner, 8 factor+invert of the diagonal blocks).  Synchronises the device. */
int bn_st_profile(int64_t* cycles_host, int n);

/* inv_vmap (utils.py:30-35): inv[k] = (A[k] + jitter I)^-1 through the Cholesky factor, one CTA per matrix.
 * Optional: sol[N,n] = inv[k] rhs[k] (rhs[N,n]); logdet[N] = log det (A[k] + jitter I). */
int bn_spd_inverse_batched(int64_t N, int n, const double* A, const double* rhs, double jitter,
                           double* inv, double* sol, double* logdet,
                           void* workspace, size_t workspace_bytes, void* stream);

/* MarkovGaussianProcess.compute_full_pseudo_lik, spatio-temporal branch (basemodels.py:676-687):
 *   nat1_full = B^T nat1, nat2_full = B^T nat2 B, pseudo_var = inv(nat2_full + jitter I), pseudo_y = pseudo_var nat1_full
 * for a time-invariant projection B (Bt = B^T, [M,Ns] row-major; the reference recomputes the same B at every
 * step, kernels.py:494-501) and factorising sites (nat2 diagonal: nat2_diag[N,Ns]).  jitter = 1e-12 in the reference.
 * pseudo_y[N,M,1], pseudo_var[N,M,M]; nat2_full[N,M,M] and logdet[N] = log det(nat2_full + jitter I) nullable. */
int bn_st_pseudo_lik(int64_t N, int Ns, int M, const double* Bt, const double* nat1, const double* nat2_diag,
                     double jitter, double* pseudo_y, double* pseudo_var, double* nat2_full, double* logdet,
                     void* workspace, size_t workspace_bytes, void* stream);

/* conditional_posterior_to_data (basemodels.py:743-764), the marginals the factorising likelihoods read
 * (likelihoods.py:371-373 takes diag(cov_f)):  mean_f[N,Ns] = B m,  var_f[N,Ns] = diag(B V B^T) + cdiag.
 * B[Ns,M] and cdiag[Ns] (nullable) time-invariant. */
int bn_st_posterior_to_data(int64_t N, int Ns, int M, const double* B, const double* cdiag,
                            const double* post_mean, const double* post_cov, double* mean_f, double* var_f,
                            void* stream);

/* vmap(gaussian_expected_log_lik)(pseudo_y, post_mean, post_cov, pseudo_var, mask) with full M x M blocks
 * (utils.py:510-531 as called from compute_kl, basemodels.py:715-721): mask[N,M] nullable, values[N] (required),
 * sum nullable. */
int bn_st_gaussian_expected_log_lik(int64_t N, int M, const double* pseudo_y, const double* post_mean,
                                    const double* post_cov, const double* pseudo_var, const uint8_t* mask,
                                    double* values, double* sum,
                                    void* workspace, size_t workspace_bytes, void* stream);

/* ---- prediction at test inputs ---------------------------------------------------
 * temporal_conditional(X_aug, X_test, mean, cov, gain, kernel) (utils.py:122-136 -> predict_from_state :99-120 ->
 * compute_conditional_statistics :173-215) as MarkovGaussianProcess.predict calls it (basemodels.py:766-816): the
 * dummy states at -1e10 / +1e10 with (minf, Pinf) are implied, `x` are the N sorted training inputs, mean/cov/gain
 * the smoother output with return_full=1.  return_full=0 applies H: out_mean[Ns,Df,1], out_cov[Ns,Df,Df];
 * =1 gives the state: [Ns,d,1], [Ns,d,d].  One thread per test point; state dimension <= 4. */
int bn_temporal_conditional(const bn_kernel_spec* k, int64_t N, const double* x, int64_t N_test, const double* x_test,
                            const double* mean, const double* cov, const double* gain, int return_full,
                            double* out_mean, double* out_cov, void* stream);

/* Likelihood.predict (likelihoods.py:493-506; Gaussian :802-803; predict_cubature cubature.py:438-465) for scalar
 * latents: mean_y[N], var_y[N] from mean_f[N], var_f[N].  cub_x[Q], cub_w[Q]: DEVICE arrays (unused for Gaussian). */
int bn_likelihood_predict(int likelihood, double lik_param, int64_t N, const double* mean_f, const double* var_f,
                          int Q, const double* cub_x, const double* cub_w, double* mean_y, double* var_y, void* stream);
/* the same with the second likelihood parameter (Student-t, negative binomial; likelihoods.py:1043-1044, 1184-1189) */
int bn_likelihood_predict2(int likelihood, double lik_param, double lik_param2, int64_t N, const double* mean_f,
                           const double* var_f, int Q, const double* cub_x, const double* cub_w, double* mean_y,
                           double* var_y, void* stream);

/* ---- sparse Markov GP ---------------------------------------------------------
 * kalman_filter_pairs (ops.py:383-426) = bn_pairs_discretise (construct_pair, :411-419: A_pair = [[0,I],[0,A]],
 * Q_pair = [[1e-32 I,0],[0,Q]] per transition, [Mt,2n,2n] each) followed by bn_kf_arrays with d = D = 2n, H = I,
 * m0 = 0, P0 = blockdiag(Pinf, Pinf); the caller drops the first step and keeps the leading n x n blocks (:426).
 * Single-component Matern-1/2, -3/2, -5/2 kernels (n = 1, 2, 3). */
int bn_pairs_discretise(const bn_kernel_spec* k, int64_t Mt, const double* dz, double* Apairs, double* Qpairs, void* stream);

/* vmap(build_joint) over the Mt transitions (utils.py:544-553 as called at basemodels.py:996-1006): mean[Mt-1,n,1],
 * cov[Mt-1,n,n], gain[Mt-1,n,n] = the full-state smoother output at the inducing points; the dummy states (minf, Pinf)
 * at both ends and the leading zero gain are implied.  joint_mean[Mt,2n,1], joint_cov[Mt,2n,2n]. */
int bn_build_joint(const bn_kernel_spec* k, int64_t Mt, const double* mean, const double* cov, const double* gain,
                   double* joint_mean, double* joint_cov, void* stream);

size_t bn_sparse_workspace_bytes(int64_t Mz);

/* The site pass of one VI iteration of SparseMarkovGaussianProcess, full batch, single-latent likelihood: for every data
 * point conditional_posterior_to_data (basemodels.py:1071-1104, compute_conditional_statistics utils.py:173-215), the
 * variational expectation (likelihoods.py:363-383), ensure_psd, conditional_data_to_posterior (:1106-1112), newton_update
 * (inference.py:21-39); per transition group_natural_params (:1114-1138), the damped update (inference.py:83-86) and
 * reparametrise (basemodels.py:85-100).  x[N] sorted data inputs, y[N] (NaN = missing), z[Mz] sorted inducing inputs,
 * start[Mz+2]: data of transition m are [start[m], start[m+1]) (ind = searchsorted(Z_aug, x) - 1, utils.py:556-559).
 * post_mean[Mt,2n,1], post_cov[Mt,2n,2n]: joint posterior (bn_build_joint).  cub_x[Q], cub_w[Q]: HOST arrays.
 * nat1[Mt,2n,1], nat2[Mt,2n,2n] updated in place; site_mean, site_cov written; diffs[2] (nullable) = mean |d nat|. */
int bn_sparse_site_update(const bn_kernel_spec* k, int likelihood, double lik_param, int64_t N, int64_t Mz,
                          const double* x, const double* y, const double* z, const int64_t* start,
                          const double* post_mean, const double* post_cov,
                          int Q, const double* cub_x_host, const double* cub_w_host, double lr, int ensure_psd,
                          double* nat1, double* nat2, double* site_mean, double* site_cov, double* diffs,
                          void* workspace, size_t workspace_bytes, void* stream);

/* nansum_n E_q[log p(y_n | f_n)] through the same conditional (the likelihood term of energy(), inference.py:197-222) */
int bn_sparse_expected_density(const bn_kernel_spec* k, int likelihood, double lik_param, int64_t N, int64_t Mz,
                               const double* x, const double* y, const double* z, const int64_t* start,
                               const double* post_mean, const double* post_cov,
                               int Q, const double* cub_x_host, const double* cub_w_host, double* sum,
                               void* workspace, size_t workspace_bytes, void* stream);

/* ---- mean-field spatio-temporal filter / smoother -------------------------------
 * kalman_filter_meanfield (ops.py:581-611 -> _sequential_kf_mf :429-467): the state covariance keeps only its M
 * diagonal n x n blocks.  means[N,M,n,1], covs[N,M,n,n] (the reference's block layout); y, noise_cov, mask, ell as
 * bn_st_kalman_filter.  One persistent kernel; a step costs one M x M stacked Cholesky sweep + O(M^2). */
int bn_st_kalman_filter_meanfield(const bn_kernel_spec* temporal, int M, int64_t N,
                                  const double* dt, const double* y, const double* noise_cov, const uint8_t* mask,
                                  double* ell, double* means, double* covs,
                                  void* workspace, size_t workspace_bytes, void* stream);

/* rauch_tung_striebel_smoother_meanfield (ops.py:681-706 -> _sequential_rts_mf :614-650): an independent RTS pass per
 * block.  return_full=0: means[N,M,1] = H sm, covs[N,M,M] = H blockdiag(sP) H^T (diagonal, zeros written);
 * return_full=1: means[N,M n,1], covs as BLOCKS [N,M,n,n] (the reference scatters them into a dense [N,d,d]).
 * gains (nullable) as blocks [N,M,n,n]. */
int bn_st_rts_smoother_meanfield(const bn_kernel_spec* temporal, int M, int64_t N,
                                 const double* dt, const double* filter_mean, const double* filter_cov, int return_full,
                                 double* means, double* covs, double* gains, void* stream);

/* ---- carry exchange of the time-sharded two-level scan over NVLink peer memory ----------
 * One kernel per exchange does the collective: every rank stores its carry[len] into a slot of every peer's inbox
 * (symmetric, peer-mapped buffers of bn_carry_exchange_bytes(world) bytes, zero-initialised once;
 * peer_buffers_dev[world] = their device addresses as seen from THIS rank, a device array), publishes `seq` with a
 * system-scope release, waits for all peers' `seq` and writes out[world,len].  seq = 1, 2, 3, ... identical on all
 * ranks; len <= 256.  Replaces the all-gather of bayesnewton_b200.distributed (ops.py:203-219, 328-335 carries). */
size_t bn_carry_exchange_bytes(int world);
int bn_carry_exchange(const uint64_t* peer_buffers_dev, int world, int rank, const double* carry, int len,
                      uint64_t seq, double* out, void* stream);

/* MarkovGaussianProcess.predict for a SpatioTemporalKernel (basemodels.py:766-816): the temporal_conditional step
 * (utils.py:99-136, 173-215) on the Kronecker state followed by H (.) H^T.  x[N]: training times; x_test[Nq];
 * mean[N,d,1], cov[N,d,d], gain[N,d,d] from bn_st_rts_smoother(return_full=1).  f_mean[Nq,M,1], f_cov[Nq,M,M]: the
 * latent at the inducing points at the test times; bn_st_posterior_to_data then maps them to any spatial inputs. */
size_t bn_st_predict_workspace_bytes(const bn_kernel_spec* temporal, int M, int64_t N_test);
int bn_st_predict_state(const bn_kernel_spec* temporal, int M, int64_t N, const double* x, int64_t N_test,
                        const double* x_test, const double* mean, const double* cov, const double* gain,
                        double* f_mean, double* f_cov, void* workspace, size_t workspace_bytes, void* stream);

/* ---- fused inference iteration on chunk-tiled resident state (single latent, one site per step) ------------
 * One iteration of a temporal Markov GP is  inference(): update_posterior -> update_variational_params + newton_update +
 * damped update_nat_params -> update_posterior  (inference.py:65-90), then energy(): E_q[log p(y|f)] summed, the filter
 * log-likelihood, and E_q[log N(pseudo_y | f, pseudo_var)] summed (inference.py:197-222; basemodels.py:708-741).  The
 * arrays that only travel between those stages -- dt, Y, the sites (pseudo_y, pseudo_var), the posterior marginals --
 * are kept here in the layout of the thread that consumes them ("tiled": the series is cut into the chunks of the
 * temporally parallel filter, chunk c step j at [((c >> 5) * L + j) * 32 + (c & 31)], L = bn_iter_chunk_len), and the
 * two smoother sweeps run the per-step site work in their epilogue on the marginal they hold in registers:
 *   BN_ITER_PLAIN   update_posterior()                                     basemodels.py:689-706
 *   BN_ITER_SITES   update_posterior(); the site update of inference.py:72-86 for every step, sites rewritten IN PLACE
 *                   sums[0] = sum_n |nat1_new - nat1|, sums[1] = sum_n |nat2_new - nat2| (before damping; the `diff`
 *                   terms of inference.py:78-79 times N)
 *   BN_ITER_ENERGY  update_posterior(); sums[0] = nansum_n likelihood term (VI: E_q[log p]; Newton: log p(y | m);
 *                   EP: log Z_n of the tilted distribution at the cavity, inference.py:297-305),
 *                   sums[1] = sum_n gaussian_expected_log_lik(pseudo_y_n, m_n, v_n, pseudo_var_n)  (utils.py:510-531)
 *                   (EP: sum_n log Z of the site at the cavity with the power-EP constant, basemodels.py:247-262)
 * ell (nullable) = the filter log-likelihood of the pass = compute_log_lik() (basemodels.py:726-741).
 * Supported: n_components = 1 of any Matern family; likelihood in {Gaussian, Bernoulli probit / logit, Poisson};
 * method VI, Newton or EP (power in bn_iter_args).  cub_x / cub_w: HOST arrays (the 1-D rule, Q <= 64; ignored by Newton
 * and the closed forms).
 * Ranks of a time-sharded run call the three phases with their carries exchanged in between (as bn_up_shard_*). */
enum { BN_ITER_PLAIN = 0, BN_ITER_SITES = 1, BN_ITER_ENERGY = 2 };

typedef struct {
    int64_t N;                 /* steps of this time shard */
    int32_t rank, world;       /* position of the shard (0, 1 for a single GPU) */
    /* the arrays below hold fp64 values for the bn_iter_* entry points and fp32 values for bn_iter_*_f32 */
    const void* dt_t;          /* tiled dt (dt[0] of rank 0 is ignored: the prior is stationary) */
    const void* y_t;           /* tiled observations Y (SITES / ENERGY) */
    void* site_mean_t;         /* tiled pseudo observations */
    void* site_cov_t;          /* tiled pseudo variances */
    const uint8_t* mask_t;     /* tiled mask of missing pseudo observations, nullable */
    void* post_mean_t;         /* tiled posterior marginals, written by PLAIN / ENERGY */
    void* post_cov_t;
    int32_t method, likelihood;
    double lik_param;          /* Gaussian variance / Poisson bin size */
    int32_t Q, ensure_psd;
    const double* cub_x_host;  /* [Q] */
    const double* cub_w_host;  /* [Q] */
    double lr, power;
    int32_t want_ell;          /* shard phases: the pass will produce the filter log-likelihood (all phases must agree) */
    int32_t reserved_;
    void* post_mean;           /* nullable pair: PLAIN / ENERGY write the marginals to these [N] arrays in time order */
    void* post_cov;            /* (the reference's posterior_mean / posterior_variance layout) instead of the tiled ones */
} bn_iter_args;

int bn_iter_chunk_len(const bn_kernel_spec* k, int64_t N);       /* L */
int64_t bn_iter_tiled_len(const bn_kernel_spec* k, int64_t N);   /* elements of a tiled array (padding included) */
size_t bn_iter_workspace_bytes(const bn_kernel_spec* k, int64_t N);
/* layout conversion, both sides coalesced; padding of the tiled side is filled with `fill` */
int bn_iter_to_tiled(const bn_kernel_spec* k, int64_t N, const double* x, double* x_t, double fill, void* stream);
int bn_iter_from_tiled(const bn_kernel_spec* k, int64_t N, const double* x_t, double* x, void* stream);
int bn_iter_to_tiled_u8(const bn_kernel_spec* k, int64_t N, const uint8_t* x, uint8_t* x_t, void* stream);
/* one pass on a single GPU (world = 1) */
int bn_iter_pass(const bn_kernel_spec* k, const bn_iter_args* a, int mode, double* ell, double* sums,
                 void* workspace, size_t workspace_bytes, void* stream);
/* the same pass in the three phases of a time-sharded run; the workspace carries state from phase to phase */
int bn_iter_shard_reduce(const bn_kernel_spec* k, const bn_iter_args* a, double* kf_carry,
                         void* workspace, size_t workspace_bytes, void* stream);
int bn_iter_shard_filter(const bn_kernel_spec* k, const bn_iter_args* a, const double* kf_carries, double* ell,
                         double* rts_carry, void* workspace, size_t workspace_bytes, void* stream);
int bn_iter_shard_smooth(const bn_kernel_spec* k, const bn_iter_args* a, int mode, const double* rts_carries,
                         double* sums, void* workspace, size_t workspace_bytes, void* stream);

/* fp32 mode of the fused iteration: the same kernels compiled with a float scalar type (storage AND arithmetic in fp32;
 * sums accumulated in fp64; log Phi through erff / logf instead of the packed fp64 table).  Every array of
 * bn_iter_args, ell, sums and the carries are float; the chunk plan (bn_iter_chunk_len, bn_iter_tiled_len) is shared
 * with the fp64 entry points.  Parity bar: 1e-4 relative against the fp64 results (tests/test_fp32_mode.py). */
size_t bn_iter_workspace_bytes_f32(const bn_kernel_spec* k, int64_t N);
int bn_iter_to_tiled_f32(const bn_kernel_spec* k, int64_t N, const float* x, float* x_t, float fill, void* stream);
int bn_iter_from_tiled_f32(const bn_kernel_spec* k, int64_t N, const float* x_t, float* x, void* stream);
int bn_iter_pass_f32(const bn_kernel_spec* k, const bn_iter_args* a, int mode, float* ell, float* sums,
                     void* workspace, size_t workspace_bytes, void* stream);
int bn_iter_shard_reduce_f32(const bn_kernel_spec* k, const bn_iter_args* a, float* kf_carry,
                             void* workspace, size_t workspace_bytes, void* stream);
int bn_iter_shard_filter_f32(const bn_kernel_spec* k, const bn_iter_args* a, const float* kf_carries, float* ell,
                             float* rts_carry, void* workspace, size_t workspace_bytes, void* stream);
int bn_iter_shard_smooth_f32(const bn_kernel_spec* k, const bn_iter_args* a, int mode, const float* rts_carries,
                             float* sums, void* workspace, size_t workspace_bytes, void* stream);

/* ---- infinite-horizon (steady-state) filter / smoother ---------------------------------
 * kalman_filter_infinite_horizon (ops.py:881-952) and rauch_tung_striebel_smoother_infinite_horizon (:1018-1068) for one
 * latent with one site per step (H = e_0^T, d <= 4).  The d x d algebra that does not depend on N -- the Riccati fixed
 * point Pdare (ops.py:796-824), the stationary gain, the smoother's fixed point -- is formed by the caller on the host
 * (A_host, Pdare_host, gain_host: row-major [d,d] HOST arrays); these entries run the O(N) affine recursions of the mean
 *     m_k = (A - K_k H A) m_{k-1} + K_k y_k,  K_k = Pdare H^T / (H Pdare H^T + R_k)     (_sequential_kf_ih / _parallel_kf_ih)
 *     sm_k = fm_k + G (sm_{k+1} - A fm_k)                                               (_sequential_rts_ih / _parallel_rts_ih)
 * form BN_SEQUENTIAL: one thread in time order; BN_SCAN: three-phase scan over (M, v) pairs.
 * noise_var[N] (noise_is_scalar = 0: heteroscedastic) or noise_var[1] (= 1: the tied variance for every step);
 * mask[N] nullable: only enters ell (utils.py:376-396); ell nullable; means[N,d,1].
 * bn_ih_smoother: filter_mean[N,d,1] -> means[N,1,1] = H sm (return_full = 0) or [N,d,1]. */
size_t bn_ih_workspace_bytes(int d, int64_t N);
int bn_ih_filter(int form, int d, int64_t N, const double* A_host, const double* Pdare_host, const double* y,
                 const double* noise_var, int noise_is_scalar, const uint8_t* mask, double* ell, double* means,
                 void* workspace, size_t workspace_bytes, void* stream);
int bn_ih_smoother(int form, int d, int64_t N, const double* A_host, const double* gain_host, const double* filter_mean,
                   int return_full, double* means, void* workspace, size_t workspace_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* BN_B200_H */
