/* bobe_b200 -- C-ABI of the B200-native GP surrogate hot path (drop-in under BOBE/gp.py, BOBE/acquisition.py).
 *
 * The reference (Ameek94/BOBE) has no FFI boundary of its own: the boundary is the Python surface of
 * `class GP` (BOBE/gp.py:199-772) whose arithmetic is JAX library calls.  Each entry point below replaces
 * one group of those call sites; the file:line each one stands in for is cited on the declaration.
 *
 * Conventions (what an XLA FFI handler lives under, so that each maps 1:1 onto a jax.ffi custom call):
 *   - `stream` is a cudaStream_t passed as void*; every call is asynchronous on it, never synchronises
 *     the host, never allocates: scratch comes from the caller (`ws`, sized by the *_workspace_bytes query).
 *   - all arrays are DEVICE pointers to float64, C-order; sizes are int64; scalars are host doubles.
 *   - return value: 0 on success, a negative BOBE_E_* code otherwise; bobe_last_error_string() gives the
 *     text (thread-local).  Numerical failure (non-PD K) is NOT an error: NaN results + info != 0,
 *     exactly like jnp.linalg.cholesky in the reference.
 *   - no global mutable state except a per-device cache of function attributes.
 *
 * Matrices that live in caller memory between calls (`Linv`, `L`) use the padded size
 *   npad = bobe_npad(n)   (n rounded up to a multiple of 64)
 * as both dimension and leading dimension; rows/cols >= n hold the identity.
 */
#ifndef BOBE_B200_H
#define BOBE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BOBE_KERNEL_RBF 0      /* BOBE/gp.py:124-154 */
#define BOBE_KERNEL_MATERN52 1 /* BOBE/gp.py:156-168 */

#define BOBE_MAX_DIM 144 /* largest input dimension d (shared-memory staging of one 64-row tile per operand) */

#define BOBE_OK 0
#define BOBE_E_ARG (-1)       /* bad argument (shape, null pointer, misaligned) */
#define BOBE_E_WORKSPACE (-2) /* workspace too small */
#define BOBE_E_CUDA (-3)      /* CUDA runtime error (launch failure etc.) */

/* predict modes (bit mask) */
#define BOBE_PREDICT_MEAN 1
#define BOBE_PREDICT_VAR 2
#define BOBE_PREDICT_STANDARDISED 4 /* predict_single semantics: no un-standardise, NaN -> 1e-12 */

/* fantasy-variance reductions */
#define BOBE_REDUCE_NONE 0      /* out[C, n_mc] */
#define BOBE_REDUCE_MEAN 1      /* WIPV   (BOBE/acquisition.py:438-440): out[C] */
#define BOBE_REDUCE_MEAN_SQRT 2 /* WIPStd (BOBE/acquisition.py:463-465): out[C] */

#define BOBE_ACQ_EI 0    /* BOBE/acquisition.py:226-253 */
#define BOBE_ACQ_LOGEI 1 /* BOBE/acquisition.py:318-330 */

const char* bobe_last_error_string(void);
int32_t bobe_abi_version(void);
int64_t bobe_npad(int64_t n);

/* K(xa, xb) [+ noise*I]  -- rbf_kernel / matern_kernel, BOBE/gp.py:124-168 (dist_sq :80-96 fused).
 * out is (n1, ldo) row-major, ldo >= n2.  `ls` is a device vector of d lengthscales. */
int32_t bobe_kernel_matrix(void* stream, int32_t kind, const double* xa, int64_t n1, const double* xb, int64_t n2,
                           int64_t d, const double* ls, double kv, double noise, int32_t add_noise, double* out,
                           int64_t ldo);

/* K -> L -> L^-1 -> alpha for `batch` hyper-parameter settings at once
 * -- GP.__init__ / recompute_cholesky, BOBE/gp.py:258-260,544-550 (jnp.linalg.cholesky + cho_solve).
 * ls is (batch, d) device, kv is (batch) device.  Outputs (any may be NULL):
 *   L (batch, npad, npad) lower with zero upper; Linv (batch, npad, npad) = L^-1, lower;
 *   alpha (batch, npad); logdet (batch) = sum log L_ii; quad (batch) = y^T K^-1 y; info (batch) int32,
 *   0 = PD, 1 = not PD (outputs NaN). */
int64_t bobe_factorize_workspace_bytes(int64_t n, int64_t d, int64_t batch);
int32_t bobe_factorize(void* stream, int32_t kind, const double* X, const double* y, int64_t n, int64_t d,
                       const double* ls, const double* kv, double noise, int64_t batch, double* L, double* Linv,
                       double* alpha, double* logdet, double* quad, int32_t* info, void* ws, int64_t ws_bytes);

/* The same for caller-supplied kernel matrices K (batch, n, ldk) (lower triangle read, noise already on the diagonal)
 * -- gp_mll(k, train_y, num_points), BOBE/gp.py:170-178: jnp.linalg.cholesky(k) + cho_solve((L, True), train_y).
 * y is (n), shared by the batch.  Outputs as bobe_factorize; gp_mll = -quad/2 - logdet - n/2 log(2 pi). */
int64_t bobe_cholesky_workspace_bytes(int64_t n, int64_t batch);
int32_t bobe_cholesky_batched(void* stream, const double* K, int64_t n, int64_t ldk, int64_t batch, const double* y,
                              double* L, double* Linv, double* alpha, double* logdet, double* quad, int32_t* info,
                              void* ws, int64_t ws_bytes);

/* out[i][j] = sum_k (xa[i][k] - xb[j][k])^2 by direct differences -- dist_sq, BOBE/gp.py:80-96.  out is (n1, ldo). */
int32_t bobe_dist_sq(void* stream, const double* xa, int64_t n1, const double* xb, int64_t n2, int64_t d, double* out,
                     int64_t ldo);

/* Rank-b append: the factorisation of the first n_old points, held in buffers already sized for the padded
 * npad(n_old + b) (identity beyond n_old), is extended IN PLACE by the points n_old .. n_old+b-1 of X, and alpha is
 * re-solved for the targets y of all n_old + b points -- GP.update, BOBE/gp.py:495-541, whose hyper-parameters are
 * unchanged and whose O(n^3) re-factorisation (:541) this replaces by O(b n^2) work; also the kriging-believer
 * updates of BOBE/acquisition.py:182-194.  info (device int32): 1 if an appended pivot is not positive. */
int64_t bobe_factor_append_workspace_bytes(int64_t n_new, int64_t d);
int32_t bobe_factor_append(void* stream, int32_t kind, const double* X, const double* y, int64_t n_old, int64_t b,
                           int64_t d, const double* ls, double kv, double noise, double* L, double* Linv, double* alpha,
                           int32_t* info, void* ws, int64_t ws_bytes);

/* log marginal likelihood and its gradient w.r.t. the log-parameters, for R restarts at once
 * -- value_and_grad of the data term of GP.neg_mll, BOBE/gp.py:385-398 + gp_mll :170-178, as called at
 * BOBE/optim.py:118,211,309.  log_params is (R, P) device, layout [log l_1..log l_d, log kv?, log tausq?]
 * (BOBE/gp.py:368-383); has_kv = 0 means kernel variance is fixed to `fixed_kv`.  Outputs: val (R) = log p(y)
 * (no prior, not negated), grad (R, P) = d log p / d log_params (tausq column 0), info (R).
 * Priors are O(d) and stay on the host. */
int64_t bobe_mll_grad_workspace_bytes(int64_t n, int64_t d, int64_t R);
int32_t bobe_mll_grad_batched(void* stream, int32_t kind, const double* X, const double* y, int64_t n, int64_t d,
                              const double* log_params, int64_t R, int64_t P, int32_t has_kv, double fixed_kv,
                              double noise, double* val, double* grad, int32_t* info, void* ws, int64_t ws_bytes);

/* posterior mean / variance at M query points
 * -- predict_mean_single/_batched BOBE/gp.py:450-457,468-470; predict_var_* :459-466,472-474;
 *    predict_single/_batched :476-493.  Linv is (npad, npad) from bobe_factorize, alpha (npad).
 * mean_out / var_out are (M); either may be NULL if its mode bit is clear. */
int64_t bobe_predict_workspace_bytes(int64_t n, int64_t d, int64_t M, int32_t mode);
int32_t bobe_predict(void* stream, int32_t kind, const double* X, int64_t n, int64_t d, const double* ls, double kv,
                     double noise, const double* Linv, const double* alpha, const double* Xq, int64_t M,
                     double y_mean, double y_std, int32_t mode, double* mean_out, double* var_out, void* ws,
                     int64_t ws_bytes);

/* LinvT = Linv^T (npad, npad) -- the second orientation of the inverse factor, which the input gradient of the
 * variance needs for w = K^-1 k* = Linv^T (Linv k*); built once per factorisation and cached by the caller. */
int32_t bobe_linv_transpose(void* stream, const double* Linv, int64_t n, double* LinvT);

/* posterior mean / variance AND their gradients with respect to the query point
 * -- what jax.grad / jax.value_and_grad of predict_mean_single / predict_var_single / predict_single
 *    (BOBE/gp.py:450-489) yield where the reference differentiates the surrogate: NUTS BOBE/samplers.py:268-285,
 *    EI / LogEI optimisation BOBE/acquisition.py:281-290 via BOBE/optim.py:118,309.
 * Same mode bits and value semantics as bobe_predict; dmean_out / dvar_out are (M, d).  The gradient of the
 * variance is zero where the floor / clip of BOBE/gp.py:465,487-488 is active. */
int64_t bobe_predict_grad_workspace_bytes(int64_t n, int64_t d, int64_t M);
int32_t bobe_predict_grad(void* stream, int32_t kind, const double* X, int64_t n, int64_t d, const double* ls, double kv,
                          double noise, const double* Linv, const double* LinvT, const double* alpha, const double* Xq,
                          int64_t M, double y_mean, double y_std, int32_t mode, double* mean_out, double* var_out,
                          double* dmean_out, double* dvar_out, void* ws, int64_t ws_bytes);

/* fantasy variance at n_mc Monte-Carlo points for C candidate points
 * -- GP.fantasy_var BOBE/gp.py:552-576 (fast_update_cholesky :181-197 folded in algebraically) and the
 *    WIPV / WIPStd reductions BOBE/acquisition.py:438-440,463-465, candidate sweep :390-397.
 * If Xcand == NULL the MC points themselves are the candidates (C must equal n_mc). */
int64_t bobe_fantasy_var_workspace_bytes(int64_t n, int64_t d, int64_t n_mc, int64_t C);
int32_t bobe_fantasy_var(void* stream, int32_t kind, const double* X, int64_t n, int64_t d, const double* ls,
                         double kv, double noise, const double* Linv, double y_std, const double* Xmc,
                         int64_t n_mc, const double* Xcand, int64_t C, int32_t reduce, double* out, void* ws,
                         int64_t ws_bytes);

/* WIPV / WIPStd values AND their gradients with respect to each candidate point
 * -- jax.value_and_grad of WIPV.fun / WIPStd.fun (BOBE/acquisition.py:438-440,463-465, i.e. of GP.fantasy_var
 *    BOBE/gp.py:552-576 in new_x) as the n <= 500 polish step takes it: BOBE/acquisition.py:400-412 through
 *    BOBE/optim.py:118,309.  reduce must be BOBE_REDUCE_MEAN or BOBE_REDUCE_MEAN_SQRT; out is (C), dout is (C, d).
 *    The gradient of an MC term is zero where the NaN / 1e-12 floor of BOBE/gp.py:574-575 is active.
 *    LinvT from bobe_linv_transpose. */
int64_t bobe_fantasy_var_grad_workspace_bytes(int64_t n, int64_t d, int64_t n_mc, int64_t C);
int32_t bobe_fantasy_var_grad(void* stream, int32_t kind, const double* X, int64_t n, int64_t d, const double* ls,
                              double kv, double noise, const double* Linv, const double* LinvT, double y_std,
                              const double* Xmc, int64_t n_mc, const double* Xcand, int64_t C, int32_t reduce,
                              double* out, double* dout, void* ws, int64_t ws_bytes);

/* rank-1 append to a lower Cholesky factor -- fast_update_cholesky BOBE/gp.py:181-197.
 * L (n, ldl) lower; k (n); L_out (n+1, ldo) fully written (zero upper). */
int32_t bobe_chol_append(void* stream, const double* L, int64_t n, int64_t ldl, const double* k, double k_self,
                         double* L_out, int64_t ldo);

/* negated EI / LogEI from standardised (mean, var) -- BOBE/acquisition.py:21-75,226-253,318-330. */
int32_t bobe_acq_ei(void* stream, int32_t which, const double* mean, const double* var, int64_t M, double best_y,
                    double zeta, double* out);

/* SVM feasibility mask of GPwithClassifier -- BOBE/clf_gp.py:173-205 (predict_*_single: jnp.where(clf_probs >=
 * threshold, value, fill)) with the RBF-SVM decision function of BOBE/clf.py:188-214:
 *   decision_q = sum_j dual_coef_j exp(-gamma |sv_j - x_q|^2) + intercept;   infeasible (decision < 0):
 *   mean_inout[q] = minus_inf, var_inout[q] = var_fill.   Any of mean_inout / var_inout / decision_out may be NULL. */
int64_t bobe_svm_mask_workspace_bytes(int64_t d, int64_t M);
int32_t bobe_svm_mask(void* stream, const double* sv, int64_t n_sv, int64_t d, const double* dual_coef, double intercept,
                      double gamma, const double* Xq, int64_t M, double minus_inf, double var_fill, double* mean_inout,
                      double* var_inout, double* decision_out, void* ws, int64_t ws_bytes);

/* ---- measurement hook (bench.py only; not part of the drop-in surface) -------------------------------------
 * Launches exactly the dominant kernel of bobe_predict (the fused triangular multiply + column sum of squares)
 * once over `rows_pad` queries whose K* panel (rows_pad, npad) is already in `kstar`, so that bench.py can time
 * that kernel alone with CUDA events for the roofline figure.  rows_pad must be a multiple of 128. */
int32_t bobe_bench_trmm_sumsq(void* stream, const double* Linv, int64_t n, const double* kstar, int64_t rows_pad,
                              double kk, double* var_out);

#ifdef __cplusplus
}
#endif
#endif /* BOBE_B200_H */
