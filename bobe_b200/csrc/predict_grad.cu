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

// ---- gradient of the fantasy-variance acquisitions with respect to the candidate point ---------------------------
// WIPV / WIPStd (BOBE/acquisition.py:438-440,463-465) are mean_j phi(s_j(x)) with
//   v = Linv k(X,x), delta2 = k** - v.v, V_j = Linv k(X,mc_j), t_j = k(x,mc_j) - v.V_j, s_j = k** - |V_j|^2 - t_j^2/delta2.
// jax.value_and_grad of that in x (the n <= 500 polish, BOBE/acquisition.py:400-412 through BOBE/optim.py:118,309):
//   ds_j/dx = -2 t_j t_j'/delta2 + t_j^2 delta2'/delta2^2,
//   t_j' = dk(x,mc_j)/dx - sum_i dk(x,X_i)/dx (K^-1 k(X,mc_j))_i,    delta2' = -2 sum_i (K^-1 k(X,x))_i dk(x,X_i)/dx.
// With c_j = phi'(s_j) (0 where the NaN / 1e-12 floor of BOBE/gp.py:574-575 is active), a_j = -2 c_j t_j/delta2,
// b = sum_j c_j t_j^2/delta2^2, z = sum_j a_j V_j:
//   n_mc * d out/dx = sum_j a_j dk(x,mc_j)/dx + sum_i e_i dk(x,X_i)/dx,    e = -Linv^T (z + 2 b v),
// i.e. per MC chunk one extra skinny GEMM (Z += A V^T) beside the products bobe_fantasy_var already does, then ONE
// triangular product for all candidates and the two "coefficient x G" reductions of bobe_predict_grad.
namespace bobe {
namespace {

constexpr int64_t FG_MCCHUNK = 16384;

// one warp per row: out[r] = c0 - sum_k M[r][k]^2
__global__ void __launch_bounds__(256) fg_row_sumsq_kernel(const double* __restrict__ Mtx, int64_t ld, int64_t rows,
                                                           int64_t cols, double c0, double* __restrict__ out) {
    const int64_t r = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (r >= rows) return;
    const double* m = Mtx + r * ld;
    double s = 0.0;
    for (int64_t k = lane; k < cols; k += 32) s = fma(m[k], m[k], s);
    s = warp_sum(s);
    if (lane == 0) out[r] = c0 - s;
}

__device__ __forceinline__ double block_sum_256(double v, double* red) {  // fixed order; result valid in thread 0
    v = warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0.0;
    if (threadIdx.x == 0)
        for (int w = 0; w < 8; ++w) t += red[w];
    return t;
}

// One CTA per candidate c, over the MC points of one chunk:
//   value accumulation (as fantasy_combine_kernel), A[c][j] = a_j, Cmc[c][j] = a_j G(x_c, mc_j), bacc[c] += b.
template <int KIND>
__global__ void __launch_bounds__(256) fantasy_grad_coef_kernel(
    const double* __restrict__ base, const double* __restrict__ kc, const double* __restrict__ G, int64_t ldg,
    const double* __restrict__ delta2, int64_t nj, int64_t nj_pad, double scale, int reduce,
    const double* __restrict__ Xcand, const double* __restrict__ xms, int64_t ldx, int d, const double* __restrict__ ls,
    double kv, double* __restrict__ A, double* __restrict__ Cmc, double* __restrict__ acc, double* __restrict__ bacc) {
    __shared__ double red[8];
    __shared__ double xc[BOBE_MAX_DIM];
    const int64_t c = blockIdx.x;
    for (int k = threadIdx.x; k < d; k += 256) xc[k] = Xcand[c * d + k] / ls[k];
    __syncthreads();
    double d2 = delta2[c];
    if (d2 < 0.0) d2 = nan("");  // sqrt of a negative pivot in fast_update_cholesky (BOBE/gp.py:187)
    double s_val = 0.0, s_b = 0.0;
    for (int64_t j = threadIdx.x; j < nj_pad; j += 256) {
        double a = 0.0, cm = 0.0;
        if (j < nj) {
            const double kcj = kc[c * ldg + j];
            const double t = kcj - G[c * ldg + j];
            double s = base[j] - t * t / d2;
            const bool floored = !(s >= SAFE_FLOOR);  // NaN or below the floor (BOBE/gp.py:574-575): zero gradient
            if (floored) s = SAFE_FLOOR;
            const double val = s * scale;
            double cj;
            if (reduce == BOBE_REDUCE_MEAN_SQRT) {
                const double sq = sqrt(val);
                s_val += sq;
                cj = scale / (2.0 * sq);
            } else {
                s_val += val;
                cj = scale;
            }
            if (!floored) {
                a = -2.0 * cj * t / d2;
                s_b += cj * t * t / (d2 * d2);
                double Gk;
                if (KIND == BOBE_KERNEL_RBF) {
                    Gk = kcj;
                } else {
                    double qq = 0.0;
                    for (int k = 0; k < d; ++k) {
                        const double df = xc[k] - xms[(int64_t)k * ldx + j];
                        qq = fma(df, df, qq);
                    }
                    const bool clamped = qq < 1e-30;
                    const double r = sqrt_pos(clamped ? 1e-30 : qq);
                    Gk = clamped ? 0.0 : kv * (5.0 / 3.0) * (1.0 + SQRT5 * r) * exp_nonpos(-SQRT5 * r);
                }
                cm = a * Gk;
            }
        }
        A[c * ldg + j] = a;
        Cmc[c * ldg + j] = cm;
    }
    const double tv = block_sum_256(s_val, red);
    const double tb = block_sum_256(s_b, red);
    if (threadIdx.x == 0) {  // chunks arrive in stream order: deterministic
        acc[c] += tv;
        bacc[c] += tb;
    }
}

// Y[c][k] = Z[c][k] + 2 b_c VcT[c][k]
__global__ void __launch_bounds__(256) fantasy_grad_y_kernel(const double* __restrict__ Z, const double* __restrict__ VcT,
                                                             const double* __restrict__ bacc, int64_t npad,
                                                             double* __restrict__ Y) {
    const int64_t k = (int64_t)blockIdx.x * 256 + threadIdx.x, c = blockIdx.y;
    if (k < npad) Y[c * npad + k] = fma(2.0 * bacc[c], VcT[c * npad + k], Z[c * npad + k]);
}

__global__ void __launch_bounds__(256) fantasy_grad_combine_kernel(const double* __restrict__ Xcand, int64_t C, int d,
                                                                   const double* __restrict__ ls,
                                                                   const double* __restrict__ Pmc,
                                                                   const double* __restrict__ Pe, int64_t ldp,
                                                                   const double* __restrict__ acc, double inv_nmc,
                                                                   double* __restrict__ out, double* __restrict__ dout) {
    const int64_t idx = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (idx >= C * d) return;
    const int64_t c = idx / d;
    const int k = (int)(idx - c * d);
    const double l = ls[k], xk = Xcand[c * d + k] / l;
    const double gm = xk * Pmc[c * ldp + d] - Pmc[c * ldp + k];
    const double ge = xk * Pe[c * ldp + d] - Pe[c * ldp + k];
    dout[idx] = -(gm + ge) / l * inv_nmc;
    if (k == 0) out[c] = acc[c] * inv_nmc;
}

struct FGradLayout {
    double *xs, *xms, *Kmc, *VT, *V, *base, *Kc, *VcT, *delta2, *G, *kc, *A, *Cmc, *acc, *bacc, *Z, *Y, *E, *Ce, *Pmc, *Pe;
    int64_t chunk, cpad, dpad, bytes;
};
FGradLayout fgrad_layout(void* ws, int64_t n, int64_t d, int64_t n_mc, int64_t C) {
    const int64_t npad = npad_of(n);
    FGradLayout l{};
    l.chunk = round_up(n_mc < FG_MCCHUNK ? n_mc : FG_MCCHUNK, 64);
    l.cpad = round_up(C, 64);
    l.dpad = round_up(d + 1, 2);
    double* b = ws ? align256(ws) : nullptr;
    int64_t off = 0;
    auto take = [&](int64_t doubles) {
        double* p = b ? b + off : nullptr;
        off += round_up(doubles, 32);
        return p;
    };
    l.xs = take(l.dpad * npad);
    l.xms = take(l.dpad * l.chunk);
    l.Kmc = take(l.chunk * npad);
    l.VT = take(l.chunk * npad);
    l.V = take(npad * l.chunk);
    l.base = take(l.chunk);
    l.Kc = take(l.cpad * npad);
    l.VcT = take(l.cpad * npad);
    l.delta2 = take(l.cpad);
    l.G = take(l.cpad * l.chunk);
    l.kc = take(l.cpad * l.chunk);
    // zero-initialised block (one memset): A, Cmc, acc, bacc, Z, Pmc
    l.A = take(l.cpad * l.chunk);
    l.Cmc = take(l.cpad * l.chunk);
    l.acc = take(l.cpad);
    l.bacc = take(l.cpad);
    l.Z = take(l.cpad * npad);
    l.Pmc = take(l.cpad * l.dpad);
    l.Y = take(l.cpad * npad);
    l.E = take(l.cpad * npad);
    l.Ce = take(l.cpad * npad);
    l.Pe = take(l.cpad * l.dpad);
    l.bytes = off * 8 + 256;
    return l;
}

}  // namespace
}  // namespace bobe

extern "C" int64_t bobe_fantasy_var_grad_workspace_bytes(int64_t n, int64_t d, int64_t n_mc, int64_t C) {
    if (n <= 0 || d <= 0 || n_mc <= 0 || C <= 0) return 256;
    return fgrad_layout(nullptr, n, d, n_mc, C).bytes;
}

extern "C" int32_t bobe_fantasy_var_grad(void* stream_, int32_t kind, const double* X, int64_t n, int64_t d,
                                         const double* ls, double kv, double noise, const double* Linv,
                                         const double* LinvT, double y_std, const double* Xmc, int64_t n_mc,
                                         const double* Xcand, int64_t C, int32_t reduce, double* out, double* dout,
                                         void* ws, int64_t ws_bytes) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (C == 0) return BOBE_OK;  // no candidates
    if (!X || !ls || !Linv || !LinvT || !Xmc || !Xcand || !out || !dout || n <= 0 || d <= 0 || d > BOBE_MAX_DIM ||
        n_mc <= 0 || C < 0 || (reduce != BOBE_REDUCE_MEAN && reduce != BOBE_REDUCE_MEAN_SQRT) ||
        (kind != BOBE_KERNEL_RBF && kind != BOBE_KERNEL_MATERN52)) {
        set_error("fantasy_var_grad: bad arguments");
        return BOBE_E_ARG;
    }
    if (!ws || ws_bytes < bobe_fantasy_var_grad_workspace_bytes(n, d, n_mc, C)) {
        set_error("fantasy_var_grad: workspace too small (%lld < %lld)", (long long)ws_bytes,
                  (long long)bobe_fantasy_var_grad_workspace_bytes(n, d, n_mc, C));
        return BOBE_E_WORKSPACE;
    }
    if ((((uintptr_t)Linv) | ((uintptr_t)LinvT)) & 15) {
        set_error("fantasy_var_grad: Linv / LinvT must be 16-byte aligned");
        return BOBE_E_ARG;
    }
    const int64_t npad = npad_of(n);
    const double kk = kv + noise;  // kernel_diag(..., include_noise=True), BOBE/gp.py:561,570
    FGradLayout l = fgrad_layout(ws, n, d, n_mc, C);
    if (int32_t rc = launch_prescale(stream, X, n, d, ls, 0, l.xs, npad, 0, 1)) return rc;
    ones_row_kernel<<<(unsigned)((npad + 255) / 256), 256, 0, stream>>>(l.xs, n, npad, d, l.dpad);
    if (int32_t rc = check_launch("ones_row_kernel")) return rc;
    // A .. Pmc are contiguous in the layout: the accumulators start from zero, pad rows / columns stay zero
    if (cudaMemsetAsync(l.A, 0, (size_t)((l.Pmc + round_up(l.cpad * l.dpad, 32)) - l.A) * 8, stream) != cudaSuccess) {
        set_error("fantasy_var_grad: memset failed");
        return BOBE_E_CUDA;
    }
    auto kstar_panel = [&](const double* pts, int64_t rows, int64_t rows_pad, double* Kout) {
        KmatArgs a{};
        a.xa = pts; a.xb = X; a.ls = ls; a.kv = kv; a.noise = noise; a.out = Kout;
        a.xbs = l.xs; a.xbs_ld = npad;
        a.n1 = rows; a.n2 = n; a.d = d; a.ldo = npad; a.rows_pad = rows_pad; a.cols_pad = npad;
        a.store_rows = rows_pad; a.store_cols = npad; a.vec_ok = 1;
        return launch_kmat(stream, kind, a, 1);
    };
    auto apply_linv = [&](const double* Kin, int64_t rows_pad, double* VTout, double* Vout, int64_t ldv) {
        GemmArgs g{};  // V = Linv Kin^T; VTout[j][i] (transposed store) and optionally Vout[i][j]
        g.A = Linv; g.Bt = Kin; g.C = Vout; g.Ct = VTout; g.lda = g.ldb = g.ldct = npad; g.ldc = ldv;
        g.M = (int)npad; g.N = (int)rows_pad; g.K = (int)npad; g.alpha = 1.0; g.flags = GEMM_A_LOWER;
        return launch_gemm_nt(stream, g, 1);
    };
    auto rows_sumsq = [&](const double* Vin, int64_t rows, double* o) {
        fg_row_sumsq_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, stream>>>(Vin, npad, rows, npad, kk, o);
        return check_launch("fg_row_sumsq_kernel");
    };
    if (int32_t rc = kstar_panel(Xcand, C, l.cpad, l.Kc)) return rc;
    if (int32_t rc = apply_linv(l.Kc, l.cpad, l.VcT, nullptr, 0)) return rc;
    if (int32_t rc = rows_sumsq(l.VcT, C, l.delta2)) return rc;
    for (int64_t j0 = 0; j0 < n_mc; j0 += l.chunk) {
        const int64_t nj = (n_mc - j0 < l.chunk) ? n_mc - j0 : l.chunk;
        const int64_t nj_pad = round_up(nj, 64);
        const double* xmc = Xmc + j0 * d;
        if (int32_t rc = kstar_panel(xmc, nj, nj_pad, l.Kmc)) return rc;
        if (int32_t rc = apply_linv(l.Kmc, nj_pad, l.VT, l.V, l.chunk)) return rc;
        if (int32_t rc = rows_sumsq(l.VT, nj, l.base)) return rc;
        {  // G[c][j] = sum_i VcT[c][i] VT[j][i]
            GemmArgs g{};
            g.A = l.VcT; g.Bt = l.VT; g.C = l.G; g.lda = g.ldb = npad; g.ldc = l.chunk;
            g.M = (int)l.cpad; g.N = (int)nj_pad; g.K = (int)npad; g.alpha = 1.0;
            if (int32_t rc = launch_gemm_nt(stream, g, 1)) return rc;
        }
        {  // kc[c][j] = k(x_c, mc_j)   (BOBE/gp.py:565-568)
            KmatArgs a{};
            a.xa = Xcand; a.xb = xmc; a.ls = ls; a.kv = kv; a.noise = noise; a.out = l.kc;
            a.n1 = C; a.n2 = nj; a.d = d; a.ldo = l.chunk; a.rows_pad = l.cpad; a.cols_pad = nj_pad;
            a.store_rows = l.cpad; a.store_cols = nj_pad; a.vec_ok = 1;
            if (int32_t rc = launch_kmat(stream, kind, a, 1)) return rc;
        }
        if (int32_t rc = launch_prescale(stream, xmc, nj, d, ls, 0, l.xms, l.chunk, 0, 1)) return rc;
        ones_row_kernel<<<(unsigned)((l.chunk + 255) / 256), 256, 0, stream>>>(l.xms, nj, l.chunk, d, l.dpad);
        if (int32_t rc = check_launch("ones_row_kernel")) return rc;
        if (kind == BOBE_KERNEL_RBF)
            fantasy_grad_coef_kernel<BOBE_KERNEL_RBF><<<(unsigned)C, 256, 0, stream>>>(
                l.base, l.kc, l.G, l.chunk, l.delta2, nj, nj_pad, y_std * y_std, reduce, Xcand, l.xms, l.chunk, (int)d, ls, kv,
                l.A, l.Cmc, l.acc, l.bacc);
        else
            fantasy_grad_coef_kernel<BOBE_KERNEL_MATERN52><<<(unsigned)C, 256, 0, stream>>>(
                l.base, l.kc, l.G, l.chunk, l.delta2, nj, nj_pad, y_std * y_std, reduce, Xcand, l.xms, l.chunk, (int)d, ls, kv,
                l.A, l.Cmc, l.acc, l.bacc);
        if (int32_t rc = check_launch("fantasy_grad_coef_kernel")) return rc;
        {  // Z[c][i] += sum_j A[c][j] V[i][j]
            GemmArgs g{};
            g.A = l.A; g.Bt = l.V; g.C = l.Z; g.D = l.Z; g.lda = g.ldb = l.chunk; g.ldc = g.ldd = npad;
            g.M = (int)l.cpad; g.N = (int)npad; g.K = (int)nj_pad; g.alpha = 1.0;
            if (int32_t rc = launch_gemm_nt(stream, g, 1)) return rc;
        }
        {  // Pmc[c][k] += sum_j Cmc[c][j] xms_ext[k][j]
            GemmArgs g{};
            g.A = l.Cmc; g.Bt = l.xms; g.C = l.Pmc; g.D = l.Pmc; g.lda = g.ldb = l.chunk; g.ldc = g.ldd = l.dpad;
            g.M = (int)l.cpad; g.N = (int)l.dpad; g.K = (int)nj_pad; g.alpha = 1.0;
            if (int32_t rc = launch_gemm_nt(stream, g, 1)) return rc;
        }
    }
    fantasy_grad_y_kernel<<<dim3((unsigned)((npad + 255) / 256), (unsigned)l.cpad), 256, 0, stream>>>(l.Z, l.VcT, l.bacc, npad,
                                                                                                    l.Y);
    if (int32_t rc = check_launch("fantasy_grad_y_kernel")) return rc;
    {  // E[c][i] = -sum_k LinvT[i][k] Y[c][k]
        GemmArgs g{};
        g.A = LinvT; g.Bt = l.Y; g.C = nullptr; g.Ct = l.E; g.lda = g.ldb = g.ldct = npad;
        g.M = (int)npad; g.N = (int)l.cpad; g.K = (int)npad; g.alpha = -1.0; g.flags = GEMM_A_UPPER;
        if (int32_t rc = launch_gemm_nt(stream, g, 1)) return rc;
    }
    {  // Ce[c][i] = E[c][i] G(x_c, X_i)
        dim3 grid((unsigned)((npad + 255) / 256), (unsigned)l.cpad);
        if (kind == BOBE_KERNEL_RBF)
            grad_coef_kernel<BOBE_KERNEL_RBF><<<grid, 256, 0, stream>>>(Xcand, C, l.cpad, l.xs, n, npad, (int)d, ls, kv, nullptr,
                                                                       l.E, nullptr, l.Ce);
        else
            grad_coef_kernel<BOBE_KERNEL_MATERN52><<<grid, 256, 0, stream>>>(Xcand, C, l.cpad, l.xs, n, npad, (int)d, ls, kv,
                                                                            nullptr, l.E, nullptr, l.Ce);
        if (int32_t rc = check_launch("grad_coef_kernel")) return rc;
    }
    {  // Pe[c][k] = sum_i Ce[c][i] xs_ext[k][i]
        GemmArgs g{};
        g.A = l.Ce; g.Bt = l.xs; g.C = l.Pe; g.lda = g.ldb = npad; g.ldc = l.dpad;
        g.M = (int)l.cpad; g.N = (int)l.dpad; g.K = (int)npad; g.alpha = 1.0;
        if (int32_t rc = launch_gemm_nt(stream, g, 1)) return rc;
    }
    fantasy_grad_combine_kernel<<<(unsigned)((C * d + 255) / 256), 256, 0, stream>>>(Xcand, C, (int)d, ls, l.Pmc, l.Pe, l.dpad,
                                                                                   l.acc, 1.0 / (double)n_mc, out, dout);
    return check_launch("fantasy_grad_combine_kernel");
}
