// Input gradients of the posterior mean and variance: what jax.grad / jax.value_and_grad produce
// when the reference differentiates GP.predict_mean_single / predict_var_single / predict_single with respect to the
// query point (NUTS on the surrogate, BOBE/samplers.py:268-285; EI / LogEI optimisation, BOBE/acquisition.py:281-290
// through BOBE/optim.py:118,309).
//
//   dk(x, x_j)/dx_k = -G_j (x_k - X_jk) / l_k^2,   G = k (RBF),  G = kv 5/3 (1 + sqrt5 r) exp(-sqrt5 r) (Matern-5/2,
//                                                   0 where the 1e-30 clamp of BOBE/gp.py:162 is active)
//   dmean/dx_k = sum_j alpha_j dk_j/dx_k
//   dvar/dx_k  = -2 sum_j w_j dk_j/dx_k,  w = K^-1 k* = Linv^T (Linv k*)   (0 where the variance floor / clip is active)
//
// With the coefficient panels Cm[q][j] = alpha_j G_qj and Cv[q][j] = w_qj G_qj both gradients are
//   -(+2) [ (x_qk / l_k) rowsum(C)[q] - (C Xs^T)[q][k] ] / l_k,      Xs[k][j] = X_jk / l_k,
// i.e. two skinny NT GEMMs against the pre-scaled training inputs extended by a row of ones (which yields the row
// sums), after the two triangular products V = K* Linv^T, W = V Linv on the FP64 tensor pipe (as many flops as one
// product with an explicit K^-1, but with errors ~cond(L) eps instead of ~cond(K) eps).  Values (mean, var) come from
// the same kernels as bobe_predict, so they are bitwise those of the value-only call.
#include "gemm_nt.cuh"
#include "kernels.cuh"

namespace bobe {

namespace {

constexpr int64_t GCHUNK = 4096;  // queries per chunk (four npad-wide panels live at once)

inline double* align256(void* p) { return (double*)(((uintptr_t)p + 255) & ~(uintptr_t)255); }

// U = Linv^T (npad x npad), 32 x 32 tiles through shared memory
__global__ void __launch_bounds__(256) transpose_kernel(const double* __restrict__ in, double* __restrict__ out, int n) {
    __shared__ double tile[32][33];
    const int bx = blockIdx.x * 32, by = blockIdx.y * 32, tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int r = ty; r < 32; r += 8) tile[r][tx] = in[(int64_t)(by + r) * n + bx + tx];
    __syncthreads();
    for (int r = ty; r < 32; r += 8) out[(int64_t)(bx + r) * n + by + tx] = tile[tx][r];
}

// xs_ext row d = 1 for j < n (row sums through the GEMM), rows d+1 .. dpad-1 = 0
__global__ void ones_row_kernel(double* xs, int64_t n, int64_t npad, int64_t d, int64_t dpad) {
    int64_t j = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (j >= npad) return;
    xs[d * npad + j] = j < n ? 1.0 : 0.0;
    for (int64_t k = d + 1; k < dpad; ++k) xs[k * npad + j] = 0.0;
}

// Cm[q][j] = alpha_j G_qj,  Cv[q][j] = W[q][j] G_qj     (one thread per element; zero outside (M, n))
template <int KIND>
__global__ void __launch_bounds__(256) grad_coef_kernel(const double* __restrict__ Xq, int64_t M, int64_t rows_pad,
                                                        const double* __restrict__ xs, int64_t n, int64_t npad, int d,
                                                        const double* __restrict__ ls, double kv,
                                                        const double* __restrict__ alpha, const double* __restrict__ W,
                                                        double* __restrict__ Cm, double* __restrict__ Cv) {
    const int64_t j = (int64_t)blockIdx.x * 256 + threadIdx.x, q = blockIdx.y;
    if (j >= npad) return;
    double cm = 0.0, cv = 0.0;
    if (q < M && j < n) {
        double qq = 0.0;
        for (int k = 0; k < d; ++k) {
            double df = Xq[q * d + k] / ls[k] - xs[(int64_t)k * npad + j];
            qq = fma(df, df, qq);
        }
        double G;
        if (KIND == BOBE_KERNEL_RBF) {
            G = kv * exp_nonpos(-0.5 * qq);
        } else {
            const bool clamped = qq < 1e-30;
            const double r = sqrt_pos(clamped ? 1e-30 : qq);
            G = clamped ? 0.0 : kv * (5.0 / 3.0) * (1.0 + SQRT5 * r) * exp_nonpos(-SQRT5 * r);
        }
        cm = alpha ? alpha[j] * G : 0.0;
        cv = W ? W[q * npad + j] * G : 0.0;
    }
    if (Cm) Cm[q * npad + j] = cm;
    if (Cv) Cv[q * npad + j] = cv;
}

// d/dx_k from the GEMM results P[q][0..d-1] = sum_j C_qj X_jk / l_k and P[q][d] = sum_j C_qj
__global__ void __launch_bounds__(256) grad_combine_kernel(const double* __restrict__ Xq, int64_t M, int d,
                                                           const double* __restrict__ ls, const double* __restrict__ Pm,
                                                           const double* __restrict__ Pv, int64_t ldp,
                                                           const double* __restrict__ var, double mean_scale,
                                                           double var_scale, double var_floor,
                                                           double* __restrict__ dmean, double* __restrict__ dvar) {
    const int64_t idx = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (idx >= M * d) return;
    const int64_t q = idx / d;
    const int k = (int)(idx - q * d);
    const double l = ls[k], xk = Xq[q * d + k] / l;
    if (dmean) dmean[idx] = -mean_scale * (xk * Pm[q * ldp + d] - Pm[q * ldp + k]) / l;
    if (dvar) {
        // the reference's clip / where has zero gradient where the floor is active (BOBE/gp.py:465,487-488); NaN
        // variances were replaced by the floor as well in the standardised flavour
        const double v = var[q];
        const bool floored = !(v > var_floor);
        dvar[idx] = floored ? 0.0 : 2.0 * var_scale * (xk * Pv[q * ldp + d] - Pv[q * ldp + k]) / l;
    }
}

struct GradLayout {
    double *xs, *kstar, *W, *Cm, *Cv, *Pm, *Pv;
    int64_t rows, dpad, bytes;
};
GradLayout grad_layout(void* ws, int64_t n, int64_t d, int64_t M) {
    const int64_t npad = npad_of(n);
    GradLayout l{};
    l.rows = round_up(M < GCHUNK ? M : GCHUNK, 128);
    l.dpad = round_up(d + 1, 2);
    double* base = ws ? align256(ws) : nullptr;
    int64_t off = 0;
    auto take = [&](int64_t doubles) {
        double* p = base ? base + off : nullptr;
        off += round_up(doubles, 32);
        return p;
    };
    l.xs = take(l.dpad * npad);
    l.kstar = take(l.rows * npad);
    l.W = take(l.rows * npad);
    l.Cm = take(l.rows * npad);
    l.Cv = take(l.rows * npad);
    l.Pm = take(l.rows * l.dpad);
    l.Pv = take(l.rows * l.dpad);
    l.bytes = off * 8 + 256;
    return l;
}

}  // namespace
}  // namespace bobe

using namespace bobe;

extern "C" int32_t bobe_linv_transpose(void* stream_, const double* Linv, int64_t n, double* LinvT) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!Linv || !LinvT || n <= 0 || Linv == LinvT) {
        set_error("linv_transpose: bad arguments");
        return BOBE_E_ARG;
    }
    const int npad = (int)npad_of(n);
    transpose_kernel<<<dim3(npad / 32, npad / 32), 256, 0, stream>>>(Linv, LinvT, npad);
    return check_launch("transpose_kernel");
}

extern "C" int64_t bobe_predict_grad_workspace_bytes(int64_t n, int64_t d, int64_t M) {
    if (n <= 0 || d <= 0 || M <= 0) return 256;
    return grad_layout(nullptr, n, d, M).bytes;
}

extern "C" int32_t bobe_predict_grad(void* stream_, int32_t kind, const double* X, int64_t n, int64_t d, const double* ls,
                                     double kv, double noise, const double* Linv, const double* LinvT, const double* alpha,
                                     const double* Xq, int64_t M, double y_mean, double y_std, int32_t mode,
                                     double* mean_out, double* var_out, double* dmean_out, double* dvar_out, void* ws,
                                     int64_t ws_bytes) {
    cudaStream_t stream = (cudaStream_t)stream_;
    const bool want_mean = mode & BOBE_PREDICT_MEAN, want_var = mode & BOBE_PREDICT_VAR;
    const int standardised = (mode & BOBE_PREDICT_STANDARDISED) ? 1 : 0;
    if (M == 0) return BOBE_OK;  // empty query set
    if (!X || !ls || !Xq || n <= 0 || d <= 0 || d > BOBE_MAX_DIM || M < 0 || !(want_mean || want_var) ||
        (want_mean && (!alpha || !mean_out || !dmean_out)) || (want_var && (!Linv || !LinvT || !var_out || !dvar_out))) {
        set_error("predict_grad: bad arguments");
        return BOBE_E_ARG;
    }
    if (M == 0) return BOBE_OK;
    if (!ws || ws_bytes < bobe_predict_grad_workspace_bytes(n, d, M)) {
        set_error("predict_grad: workspace too small (%lld < %lld)", (long long)ws_bytes,
                  (long long)bobe_predict_grad_workspace_bytes(n, d, M));
        return BOBE_E_WORKSPACE;
    }
    if (want_var && ((((uintptr_t)Linv) | ((uintptr_t)LinvT)) & 15)) {
        set_error("predict_grad: Linv / LinvT must be 16-byte aligned");
        return BOBE_E_ARG;
    }
    const int64_t npad = npad_of(n);
    GradLayout l = grad_layout(ws, n, d, M);
    if (int32_t rc = launch_prescale(stream, X, n, d, ls, 0, l.xs, npad, 0, 1)) return rc;
    ones_row_kernel<<<(unsigned)((npad + 255) / 256), 256, 0, stream>>>(l.xs, n, npad, d, l.dpad);
    if (int32_t rc = check_launch("ones_row_kernel")) return rc;
    const double mean_scale = standardised ? 1.0 : y_std, var_scale = standardised ? 1.0 : y_std * y_std;
    const double var_floor = SAFE_FLOOR * var_scale;
    for (int64_t q0 = 0; q0 < M; q0 += l.rows) {
        const int64_t rows = (M - q0 < l.rows) ? M - q0 : l.rows;
        const int64_t rows_pad = round_up(rows, 128);
        const double* xq = Xq + q0 * d;
        {   // K* panel, mean (same launch as bobe_predict)
            KmatArgs a{};
            a.xa = xq; a.xb = X; a.ls = ls; a.kv = kv; a.noise = noise;
            a.xbs = l.xs; a.xbs_ld = npad;
            a.alpha = want_mean ? alpha : nullptr;
            a.mean_out = want_mean ? mean_out + q0 : nullptr;
            a.out = want_var ? l.kstar : nullptr;
            a.n1 = rows; a.n2 = n; a.d = d; a.ldo = npad; a.rows_pad = rows_pad; a.cols_pad = npad;
            a.store_rows = rows_pad; a.store_cols = npad; a.vec_ok = 1;
            a.y_mean = y_mean; a.y_std = y_std; a.mean_standardised = standardised;
            if (want_mean || want_var)
                if (int32_t rc = launch_kmat(stream, kind, a, 1)) return rc;
        }
        if (want_var) {
            // (W is free until the triangular products below: it lends its first rows to the row-split partial sums)
            if (int32_t rc = launch_trmm_sumsq(stream, Linv, (int)n, (int)npad, l.kstar, npad, rows_pad, q0, M, kv + noise,
                                               y_std * y_std, standardised, var_out, npad >= TRMM_MAX_SPLIT ? l.W : nullptr))
                return rc;
            // w = K^-1 k* as two triangular products (triangular operand = row operand, transposed store):
            //   V[q][i] = sum_k Linv[i][k] K*[q][k]  (into Cm's buffer, free until the coefficient pass)
            //   W[q][j] = sum_i LinvT[j][i] V[q][i]
            GemmArgs g{};
            g.A = Linv; g.Bt = l.kstar; g.C = nullptr; g.Ct = l.Cm; g.lda = g.ldb = g.ldct = npad;
            g.M = (int)npad; g.N = (int)rows_pad; g.K = (int)npad; g.alpha = 1.0; g.flags = GEMM_A_LOWER;
            if (int32_t rc = launch_gemm_nt(stream, g, 1)) return rc;
            g.A = LinvT; g.Bt = l.Cm; g.Ct = l.W; g.flags = GEMM_A_UPPER;
            if (int32_t rc = launch_gemm_nt(stream, g, 1)) return rc;
        }
        dim3 grid((unsigned)((npad + 255) / 256), (unsigned)rows_pad);
        double* Cm = want_mean ? l.Cm : nullptr;
        double* Cv = want_var ? l.Cv : nullptr;
        if (kind == BOBE_KERNEL_RBF)
            grad_coef_kernel<BOBE_KERNEL_RBF><<<grid, 256, 0, stream>>>(xq, rows, rows_pad, l.xs, n, npad, (int)d, ls, kv,
                                                                       alpha, want_var ? l.W : nullptr, Cm, Cv);
        else
            grad_coef_kernel<BOBE_KERNEL_MATERN52><<<grid, 256, 0, stream>>>(xq, rows, rows_pad, l.xs, n, npad, (int)d, ls,
                                                                            kv, alpha, want_var ? l.W : nullptr, Cm, Cv);
        if (int32_t rc = check_launch("grad_coef_kernel")) return rc;
        for (int which = 0; which < 2; ++which) {
            const double* Cc = which == 0 ? Cm : Cv;
            if (!Cc) continue;
            GemmArgs g{};  // P[q][k] = sum_j C[q][j] xs_ext[k][j]
            g.A = Cc; g.Bt = l.xs; g.C = which == 0 ? l.Pm : l.Pv; g.lda = g.ldb = npad; g.ldc = l.dpad;
            g.M = (int)rows_pad; g.N = (int)l.dpad; g.K = (int)npad; g.alpha = 1.0;
            if (int32_t rc = launch_gemm_nt(stream, g, 1)) return rc;
        }
        grad_combine_kernel<<<(unsigned)((rows * d + 255) / 256), 256, 0, stream>>>(
            xq, rows, (int)d, ls, l.Pm, l.Pv, l.dpad, want_var ? var_out + q0 : nullptr, mean_scale, var_scale, var_floor,
            want_mean ? dmean_out + q0 * d : nullptr, want_var ? dvar_out + q0 * d : nullptr);
        if (int32_t rc = check_launch("grad_combine_kernel")) return rc;
    }
    return BOBE_OK;
}
