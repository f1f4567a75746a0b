// extern "C" entry points declared in include/bobe_b200.h (except bobe_mll_grad_batched, see mll_grad.cu).
#include <algorithm>
#include <cmath>

#include "gemm_nt.cuh"
#include "kernels.cuh"

using namespace bobe;

namespace {

constexpr int64_t QCHUNK = 148 * 128;  // queries per trmm_sumsq launch: one 128-query tile per SM
constexpr int64_t KCHUNKS = 3;         // chunks per kernel-matrix launch: 888 CTAs = two full waves at 3 CTAs/SM

inline double* align256(void* p) { return (double*)(((uintptr_t)p + 255) & ~(uintptr_t)255); }
inline bool aligned16(const void* p) { return (((uintptr_t)p) & 15) == 0; }

struct Carver {
    double* base;
    int64_t off = 0;
    explicit Carver(void* ws) : base(ws ? align256(ws) : nullptr) {}
    double* take(int64_t doubles) {
        double* p = base ? base + off : nullptr;
        off += round_up(doubles, 32);
        return p;
    }
    int64_t bytes() const { return off * 8 + 256; }
    bool fits(void* ws, int64_t ws_bytes) const { return (char*)(base + off) <= (char*)ws + ws_bytes; }
};

// one warp per row: out[r] = c0 - sum_k M[r][k]^2
__global__ void __launch_bounds__(256) row_sumsq_kernel(const double* __restrict__ Mtx, int64_t ld, int64_t rows,
                                                        int64_t cols, double c0, double* __restrict__ out) {
    int64_t r = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    int lane = threadIdx.x & 31;
    if (r >= rows) return;
    const double* m = Mtx + r * ld;
    double s = 0.0;
    for (int64_t k = lane; k < cols; k += 32) s = fma(m[k], m[k], s);
    s = warp_sum(s);
    if (lane == 0) out[r] = c0 - s;
}

// var[c][j] = base[j] - (kc[c][j] - G[c][j])^2 / delta2[c]  with the reference's NaN / floor handling
// (BOBE/gp.py:572-576); optional mean / mean-sqrt reduction over j accumulated into acc[c] chunk by chunk.
__global__ void __launch_bounds__(256) fantasy_combine_kernel(const double* __restrict__ base, const double* __restrict__ kc,
                                                              const double* __restrict__ G, int64_t ldg,
                                                              const double* __restrict__ delta2, int64_t nj,
                                                              double scale, int reduce, double* __restrict__ out,
                                                              int64_t ldo, int64_t j_begin, double* __restrict__ acc) {
    __shared__ double red[8];
    const int64_t c = blockIdx.x;
    double d2 = delta2[c];
    if (d2 < 0.0) d2 = nan("");  // sqrt of a negative pivot in fast_update_cholesky (BOBE/gp.py:187)
    double s = 0.0;
    for (int64_t j = threadIdx.x; j < nj; j += 256) {
        double t = kc[c * ldg + j] - G[c * ldg + j];
        double var = base[j] - t * t / d2;
        if (isnan(var)) var = SAFE_FLOOR;
        if (var < SAFE_FLOOR) var = SAFE_FLOOR;
        var *= scale;
        if (reduce == BOBE_REDUCE_NONE)
            out[c * ldo + j_begin + j] = var;
        else
            s += (reduce == BOBE_REDUCE_MEAN_SQRT) ? sqrt(var) : var;
    }
    if (reduce != BOBE_REDUCE_NONE) {
        s = warp_sum(s);
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
        __syncthreads();
        if (threadIdx.x == 0) {
            double t = 0.0;
            for (int w = 0; w < 8; ++w) t += red[w];
            acc[c] += t;  // chunks arrive in stream order: deterministic
        }
    }
}

__global__ void scale_out_kernel(const double* acc, int64_t C, double inv, double* out) {
    int64_t c = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (c < C) out[c] = acc[c] * inv;
}

// ---- rank-1 append ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) chol_append_solve_kernel(const double* __restrict__ L, int64_t n, int64_t ldl,
                                                                 const double* __restrict__ k, double k_self,
                                                                 double* __restrict__ last_row) {
    extern __shared__ double v[];  // n
    __shared__ double part[32];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int64_t b = 0; b < n; b += 32) {
        int64_t r = b + warp;
        double s = 0.0;
        if (r < n)
            for (int64_t j = lane; j < b; j += 32) s = fma(L[r * ldl + j], v[j], s);
        s = warp_sum(s);
        if (lane == 0) part[warp] = s;
        __syncthreads();
        if (warp == 0) {
            int64_t row = b + lane;
            double rhs = (row < n) ? k[row] - part[lane] : 0.0;
            double vl = 0.0;
            for (int j = 0; j < 32 && b + j < n; ++j) {
                double ljj = L[(b + j) * ldl + b + j];
                double vj = __shfl_sync(0xffffffffu, rhs, j) / ljj;
                if (lane == j) vl = vj;
                if (lane > j && row < n) rhs = fma(-L[row * ldl + b + j], vj, rhs);
            }
            if (row < n) v[row] = vl;
        }
        __syncthreads();
    }
    double s = 0.0;
    for (int64_t j = threadIdx.x; j < n; j += 1024) {
        last_row[j] = v[j];
        s = fma(v[j], v[j], s);
    }
    s = warp_sum(s);
    __syncthreads();
    if (lane == 0) part[warp] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < 32; ++w) t += part[w];
        last_row[n] = sqrt(k_self - t);  // NaN if the fantasy point is numerically inside the span (BOBE/gp.py:187)
    }
}

__global__ void chol_append_copy_kernel(const double* __restrict__ L, int64_t n, int64_t ldl, double* __restrict__ out,
                                        int64_t ldo) {
    int64_t c = (int64_t)blockIdx.x * 256 + threadIdx.x, r = blockIdx.y;
    if (c <= n) out[r * ldo + c] = (c <= r && c < n) ? L[r * ldl + c] : 0.0;
}

// ---- EI / LogEI epilogue ---------------------------------------------------------------------------------------
__device__ double log1mexp_dev(double x) {  // tfp.math.log1mexp: log(1 - exp(-|x|))
    x = fabs(x);
    return x < 0.6931471805599453 ? log(-expm1(-x)) : log1p(-exp(-x));
}

__global__ void acq_ei_kernel(int which, const double* __restrict__ mean, const double* __restrict__ var, int64_t M,
                              double best_y, double zeta, double* __restrict__ out) {
    int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= M) return;
    const double inv_sqrt_2pi = 0.39894228040143267794, log_2pi = 1.8378770664093454836;
    double v = var[i];
    if (which == BOBE_ACQ_EI) {
        v = fmax(v, 1e-20);  // jnp.clip(var, a_min=1e-20), BOBE/acquisition.py:247
        double sigma = sqrt(v);
        double u = ((mean[i] - zeta) - best_y) / sigma;
        double ei = (exp(-0.5 * u * u) * inv_sqrt_2pi + u * normcdf(u)) * sigma;
        out[i] = -ei;
    } else {
        v = fmax(v, 1e-18);  // BOBE/acquisition.py:324
        double sigma = sqrt(v);
        double u = ((mean[i] - zeta) - best_y) / sigma;
        double r;
        if (u > -1.0) {  // BOBE/acquisition.py:57-59
            r = log(exp(-0.5 * u * u) * inv_sqrt_2pi + u * normcdf(u));
        } else {  // BOBE/acquisition.py:61-73
            double ue = u < -1e6 ? -1e6 : u;
            double w = log(fabs(ue) * erfcx(-0.70710678118654752440 * ue)) + 0.22579135264472743236;  // 1/2 log(pi/2)
            double second = (u > -1e6) ? log1mexp_dev(w) : -2.0 * log(fabs(u));
            r = -0.5 * (u * u + log_2pi) + second;
        }
        out[i] = -(r + log(sigma));
    }
}

}  // namespace

extern "C" const char* bobe_last_error_string(void) { return bobe::last_error(); }
extern "C" int32_t bobe_abi_version(void) { return 1; }
extern "C" int64_t bobe_npad(int64_t n) { return npad_of(n); }

extern "C" int32_t bobe_kernel_matrix(void* stream, int32_t kind, const double* xa, int64_t n1, const double* xb,
                                      int64_t n2, int64_t d, const double* ls, double kv, double noise,
                                      int32_t add_noise, double* out, int64_t ldo) {
    if (!xa || !xb || !ls || !out || n1 < 0 || n2 < 0 || d <= 0 || ldo < n2) {
        set_error("kernel_matrix: bad arguments");
        return BOBE_E_ARG;
    }
    if (add_noise && n1 != n2) {  // noise * eye(n1) only broadcasts for a square matrix (BOBE/gp.py:153)
        set_error("kernel_matrix: add_noise needs a square matrix");
        return BOBE_E_ARG;
    }
    if (n1 == 0 || n2 == 0) return BOBE_OK;
    // rows are chunked so that grid.y stays small
    const int64_t RCH = 64 * 32768;
    for (int64_t r0 = 0; r0 < n1; r0 += RCH) {
        int64_t rows = (n1 - r0 < RCH) ? n1 - r0 : RCH;
        KmatArgs a{};
        a.xa = xa + r0 * d; a.xb = xb; a.ls = ls; a.out = out + r0 * ldo;
        a.n1 = rows; a.n2 = n2; a.d = d; a.ldo = ldo;
        a.rows_pad = round_up(rows, 64); a.cols_pad = round_up(n2, 64);
        a.store_rows = rows; a.store_cols = n2;
        a.vec_ok = aligned16(out) && (ldo % 2 == 0);
        a.kv = kv; a.noise = noise; a.add_noise = add_noise && r0 == 0 && rows == n1;
        if (add_noise && rows != n1) {
            set_error("kernel_matrix: add_noise with more than %lld rows unsupported", (long long)RCH);
            return BOBE_E_ARG;
        }
        if (int32_t rc = launch_kmat((cudaStream_t)stream, kind, a, 1)) return rc;
    }
    return BOBE_OK;
}

// ---- factorize ------------------------------------------------------------------------------------------
namespace {
struct FactorLayout {
    double *KB, *L, *Lt, *Linv, *U, *Q, *diag, *stat, *z, *alpha, *xs;
    int64_t bytes;
    bool fits;
};
FactorLayout factor_layout(void* ws, int64_t ws_bytes, int64_t n, int64_t d, int64_t batch, bool need_L,
                           bool need_Linv) {
    int64_t npad = npad_of(n), m2 = npad * npad;
    Carver c(ws);
    FactorLayout l{};
    l.KB = c.take(batch * m2);
    l.Lt = c.take(batch * m2);
    l.U = c.take(batch * m2);
    l.L = c.take(need_L ? batch * m2 : 0);
    l.Linv = c.take(need_Linv ? batch * m2 : 0);
    l.Q = c.take(batch * factor_q_elems(npad));
    l.diag = c.take(batch * npad);
    l.stat = c.take(2 * batch + (factor_gate_rows(npad) * batch + 1) / 2);  // min / max pivot + the gate rows (ints)
    l.z = c.take(solve_ws_doubles(npad, batch));
    l.alpha = c.take(batch * npad);
    l.xs = c.take(batch * d * npad);
    l.bytes = c.bytes();
    l.fits = ws ? c.fits(ws, ws_bytes) : false;
    return l;
}
}  // namespace

extern "C" int64_t bobe_factorize_workspace_bytes(int64_t n, int64_t d, int64_t batch) {
    if (n <= 0 || d <= 0 || batch <= 0) return 0;
    return factor_layout(nullptr, 0, n, d, batch, true, true).bytes;
}

extern "C" int32_t bobe_factorize(void* stream_, int32_t kind, const double* X, const double* y, int64_t n, int64_t d,
                                  const double* ls, const double* kv, double noise, int64_t batch, double* L,
                                  double* Linv, double* alpha, double* logdet, double* quad, int32_t* info, void* ws,
                                  int64_t ws_bytes) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!X || !y || !ls || !kv || !ws || n <= 0 || d <= 0 || batch <= 0) {
        set_error("factorize: bad arguments");
        return BOBE_E_ARG;
    }
    FactorLayout l = factor_layout(ws, ws_bytes, n, d, batch, L == nullptr, Linv == nullptr);
    if (!l.fits) {
        set_error("factorize: workspace too small (%lld < %lld)", (long long)ws_bytes, (long long)l.bytes);
        return BOBE_E_WORKSPACE;
    }
    if ((L && !aligned16(L)) || (Linv && !aligned16(Linv))) {
        set_error("factorize: L / Linv must be 16-byte aligned");
        return BOBE_E_ARG;
    }
    const int npad = (int)npad_of(n);
    FactorBuffers fb{l.KB, L ? L : l.L, l.Lt, Linv ? Linv : l.Linv, l.U, l.Q, l.diag, l.stat, (int*)(l.stat + 2 * batch), 0, 0, factor_live_rows(n)};
    if (int32_t rc = launch_prescale(stream, X, n, d, ls, d, l.xs, npad, d * (int64_t)npad, (int)batch)) return rc;
    KmatArgs ka{};
    ka.xa = X; ka.xb = X; ka.ls = ls; ka.kv_ptr = kv; ka.out = fb.KB;
    ka.xbs = l.xs; ka.xbs_ld = npad; ka.xbs_stride = d * (int64_t)npad;
    ka.n1 = n; ka.n2 = n; ka.d = d; ka.ldo = npad; ka.rows_pad = npad; ka.cols_pad = npad;
    ka.store_rows = npad; ka.store_cols = npad; ka.vec_ok = 1;
    ka.ls_stride = d; ka.out_stride = (int64_t)npad * npad; ka.noise = noise; ka.add_noise = 1; ka.pad_identity = 1;
    ka.lower_only = 1;  // the factorisation reads the lower triangle only
    if (int32_t rc = launch_kmat(stream, kind, ka, (int)batch)) return rc;
    {
        StreamPool* pool = stream_pool();
        if (!pool) return BOBE_E_CUDA;
        std::lock_guard<std::mutex> pool_lock(pool->enqueue_mu);
        if (int32_t rc = factor_on_pool(stream, pool, fb, npad, (int)batch)) return rc;
    }
    SolveArgs sa{kind, X, ls, kv, d, noise, l.xs};
    return launch_solve_vectors(stream, fb, sa, y, n, npad, (int)batch, l.z, alpha ? alpha : l.alpha, logdet, quad,
                                info);
}

// ---- factorise a caller-supplied K (gp_mll) ---------------------------------------------------------------------
extern "C" int64_t bobe_cholesky_workspace_bytes(int64_t n, int64_t batch) {
    if (n <= 0 || batch <= 0) return 0;
    return factor_layout(nullptr, 0, n, 1, batch, true, true).bytes;
}

extern "C" int32_t bobe_cholesky_batched(void* stream_, const double* K, int64_t n, int64_t ldk, int64_t batch,
                                         const double* y, double* L, double* Linv, double* alpha, double* logdet,
                                         double* quad, int32_t* info, void* ws, int64_t ws_bytes) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!K || !y || !ws || n <= 0 || ldk < n || batch <= 0) {
        set_error("cholesky_batched: bad arguments");
        return BOBE_E_ARG;
    }
    FactorLayout l = factor_layout(ws, ws_bytes, n, 1, batch, L == nullptr, Linv == nullptr);
    if (!l.fits) {
        set_error("cholesky_batched: workspace too small (%lld < %lld)", (long long)ws_bytes, (long long)l.bytes);
        return BOBE_E_WORKSPACE;
    }
    if ((L && !aligned16(L)) || (Linv && !aligned16(Linv))) {
        set_error("cholesky_batched: L / Linv must be 16-byte aligned");
        return BOBE_E_ARG;
    }
    const int npad = (int)npad_of(n);
    FactorBuffers fb{l.KB, L ? L : l.L, l.Lt, Linv ? Linv : l.Linv, l.U, l.Q, l.diag, l.stat, (int*)(l.stat + 2 * batch), 0, 0, factor_live_rows(n)};
    if (int32_t rc = launch_pad_k(stream, K, ldk, n * ldk, (int)n, npad, (int)batch, fb.KB)) return rc;
    {
        StreamPool* pool = stream_pool();
        if (!pool) return BOBE_E_CUDA;
        std::lock_guard<std::mutex> pool_lock(pool->enqueue_mu);
        if (int32_t rc = factor_on_pool(stream, pool, fb, npad, (int)batch)) return rc;
    }
    SolveArgs sa{};
    sa.Kin = K; sa.ldk = ldk; sa.kstride = n * ldk;
    return launch_solve_vectors(stream, fb, sa, y, n, npad, (int)batch, l.z, alpha ? alpha : l.alpha, logdet, quad, info);
}

// ---- squared distances ---------------------------------------------------------------------------------------------
namespace {
__global__ void __launch_bounds__(256) dist_sq_kernel(const double* __restrict__ xa, int64_t n1, const double* __restrict__ xb,
                                                      int64_t n2, int d, double* __restrict__ out, int64_t ldo) {
    const int64_t i = (int64_t)blockIdx.y * 16 + (threadIdx.x >> 4), j = (int64_t)blockIdx.x * 16 + (threadIdx.x & 15);
    if (i >= n1 || j >= n2) return;
    double s = 0.0;
    for (int k = 0; k < d; ++k) {  // direct differences, summed in index order (BOBE/gp.py:94-96)
        const double df = xa[i * d + k] - xb[j * d + k];
        s = fma(df, df, s);
    }
    out[i * ldo + j] = s;
}
}  // namespace

extern "C" int32_t bobe_dist_sq(void* stream, const double* xa, int64_t n1, const double* xb, int64_t n2, int64_t d,
                                double* out, int64_t ldo) {
    if (!xa || !xb || !out || n1 < 0 || n2 < 0 || d <= 0 || ldo < n2 || n1 > 65535 * 16) {
        set_error("dist_sq: bad arguments");
        return BOBE_E_ARG;
    }
    if (n1 == 0 || n2 == 0) return BOBE_OK;
    dist_sq_kernel<<<dim3((unsigned)((n2 + 15) / 16), (unsigned)((n1 + 15) / 16)), 256, 0, (cudaStream_t)stream>>>(
        xa, n1, xb, n2, (int)d, out, ldo);
    return check_launch("dist_sq_kernel");
}

// ---- rank-b append ------------------------------------------------------------------------------------------
extern "C" int64_t bobe_factor_append_workspace_bytes(int64_t n_new, int64_t d) {
    (void)d;
    if (n_new <= 0) return 0;
    return append_ws_doubles(npad_of(n_new)) * 8 + 256;
}

extern "C" int32_t bobe_factor_append(void* stream_, int32_t kind, const double* X, const double* y, int64_t n_old,
                                      int64_t b, int64_t d, const double* ls, double kv, double noise, double* L,
                                      double* Linv, double* alpha, int32_t* info, void* ws, int64_t ws_bytes) {
    if (!X || !y || !ls || !L || !Linv || !alpha || !info || !ws || n_old < 0 || b <= 0 || d <= 0 || d > BOBE_MAX_DIM) {
        set_error("factor_append: bad arguments");
        return BOBE_E_ARG;
    }
    if (ws_bytes < bobe_factor_append_workspace_bytes(n_old + b, d)) {
        set_error("factor_append: workspace too small (%lld < %lld)", (long long)ws_bytes,
                  (long long)bobe_factor_append_workspace_bytes(n_old + b, d));
        return BOBE_E_WORKSPACE;
    }
    if (!aligned16(L) || !aligned16(Linv)) {
        set_error("factor_append: L / Linv must be 16-byte aligned");
        return BOBE_E_ARG;
    }
    return factor_append((cudaStream_t)stream_, kind, X, y, n_old, b, d, ls, kv, noise, L, Linv, alpha, info, align256(ws));
}

// ---- predict --------------------------------------------------------------------------------------------
namespace {
constexpr int SMALL_M = 16;  // up to this many queries the variance goes through the matrix-vector path below

// V[q][i] = sum_{k <= i} Linv[i][k] K*[q][k] for a handful of queries: one warp per row of Linv, the row is read once
// and dotted with every query's K* row.  A single 128-query tile of trmm_sumsq would run the whole triangular product
// on ONE SM (1.5 ms at n = 1500); this spreads the n^2/2 multiply-adds of a single-point call over the machine.
template <int MQ>
__global__ void __launch_bounds__(256) linv_apply_small_kernel(const double* __restrict__ Linv, int n, int npad,
                                                               const double* __restrict__ Kstar, int64_t ldk, int M,
                                                               double* __restrict__ V) {
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= n) return;
    const double* l = Linv + (int64_t)row * npad;
    double acc[MQ];
#pragma unroll
    for (int q = 0; q < MQ; ++q) acc[q] = 0.0;
    for (int k = lane; k <= row; k += 32) {
        const double lv = l[k];
#pragma unroll
        for (int q = 0; q < MQ; ++q)
            if (q < M) acc[q] = fma(lv, Kstar[q * ldk + k], acc[q]);
    }
#pragma unroll
    for (int q = 0; q < MQ; ++q) {
        const double s = warp_sum(acc[q]);
        if (lane == 0 && q < M) V[(int64_t)q * npad + row] = s;
    }
}

// var[q] = kk - sum_i V[q][i]^2 with the floor / scale semantics of trmm_sumsq_kernel (one CTA per query, fixed order)
__global__ void __launch_bounds__(256) small_var_kernel(const double* __restrict__ V, int n, int npad, double kk,
                                                        double scale, int standardised, double* __restrict__ var_out) {
    __shared__ double red[8];
    const int q = blockIdx.x;
    double s = 0.0;
    for (int i = threadIdx.x; i < n; i += 256) {
        const double v = V[(int64_t)q * npad + i];
        s = fma(v, v, s);
    }
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < 8; ++w) t += red[w];
        double var = kk - t;
        if (standardised) {  // BOBE/gp.py:487-488
            if (isnan(var)) var = SAFE_FLOOR;
            if (var < SAFE_FLOOR) var = SAFE_FLOOR;
        } else {  // BOBE/gp.py:465-466
            if (var < SAFE_FLOOR) var = SAFE_FLOOR;
            var *= scale;
        }
        var_out[q] = var;
    }
}
}  // namespace

extern "C" int64_t bobe_predict_workspace_bytes(int64_t n, int64_t d, int64_t M, int32_t mode) {
    if (M <= 0 || n <= 0 || d <= 0) return 256;
    int64_t bytes = round_up(d * npad_of(n), 32) * 8 + 512;  // scaled, transposed training inputs
    if (mode & BOBE_PREDICT_VAR) {  // K* panel of up to KCHUNKS chunks + row-split partial sums of one chunk
        bytes += round_up(M < KCHUNKS * QCHUNK ? M : KCHUNKS * QCHUNK, 128) * npad_of(n) * 8;
        bytes += TRMM_MAX_SPLIT * round_up(M < QCHUNK ? M : QCHUNK, 128) * 8;
        bytes += TRMM_COUNTERS * 4;  // tile counters of the shared-panel schedule
    }
    return bytes;
}

extern "C" int32_t bobe_predict(void* stream_, int32_t kind, const double* X, int64_t n, int64_t d, const double* ls,
                                double kv, double noise, const double* Linv, const double* alpha, const double* Xq,
                                int64_t M, double y_mean, double y_std, int32_t mode, double* mean_out,
                                double* var_out, void* ws, int64_t ws_bytes) {
    cudaStream_t stream = (cudaStream_t)stream_;
    const bool want_mean = mode & BOBE_PREDICT_MEAN, want_var = mode & BOBE_PREDICT_VAR;
    const int standardised = (mode & BOBE_PREDICT_STANDARDISED) ? 1 : 0;
    if (M == 0) return BOBE_OK;  // empty query set: nothing to do (the output pointers may be null then)
    if (!X || !ls || !Xq || n <= 0 || d <= 0 || M < 0 || (want_mean && (!alpha || !mean_out)) ||
        (want_var && (!Linv || !var_out)) || !(want_mean || want_var)) {
        set_error("predict: bad arguments");
        return BOBE_E_ARG;
    }
    const int64_t npad = npad_of(n);
    if (!ws || ws_bytes < bobe_predict_workspace_bytes(n, d, M, mode)) {
        set_error("predict: workspace too small (%lld < %lld)", (long long)ws_bytes,
                  (long long)bobe_predict_workspace_bytes(n, d, M, mode));
        return BOBE_E_WORKSPACE;
    }
    if (want_var && !aligned16(Linv)) {
        set_error("predict: Linv must be 16-byte aligned");
        return BOBE_E_ARG;
    }
    double* xs = align256(ws);  // (d, npad) = X^T / l, built once per call
    double* kstar = want_var ? xs + round_up(d * npad, 32) : nullptr;
    double* partial = want_var ? kstar + round_up(M < KCHUNKS * QCHUNK ? M : KCHUNKS * QCHUNK, 128) * npad : nullptr;
    int* counters = want_var ? reinterpret_cast<int*>(partial + TRMM_MAX_SPLIT * round_up(M < QCHUNK ? M : QCHUNK, 128)) : nullptr;
    if (counters && cudaMemsetAsync(counters, 0, TRMM_COUNTERS * sizeof(int), stream) != cudaSuccess) {
        set_error("predict: cudaMemsetAsync failed");
        return BOBE_E_CUDA;
    }
    if (int32_t rc = launch_prescale(stream, X, n, d, ls, 0, xs, npad, 0, 1)) return rc;
    // mean only: no K* panel to bound, so the rows go out in launches as large as the grid allows (more CTAs per SM
    // for the kernel-matrix kernel); with the variance, the K* panel of KCHUNKS chunks is built by one launch and
    // consumed by one trmm_sumsq launch per 148 x 128-query chunk
    // BOBE_TRMM_SPLIT = S > 1 (experiment knob, see launch_trmm_sumsq): chunks of 148 / S query tiles, one chunk per
    // kernel-matrix launch, so that the K* panels written by one launch are still in L2 when the next one reads them
    static const int64_t tsplit = std::max<int64_t>(1, std::min<int64_t>(TRMM_MAX_SPLIT, env_int("BOBE_TRMM_SPLIT", 1)));
    static const int64_t kchunks = std::max<int64_t>(1, std::min<int64_t>(KCHUNKS, env_int("BOBE_KCHUNKS", tsplit > 1 ? 1 : KCHUNKS)));
    const int64_t qchunk = tsplit > 1 ? (148 / tsplit) * 128 : (int64_t)trmm_chunk_tiles((int)n, (int)npad) * 128;
    const int64_t step = want_var ? kchunks * qchunk : (int64_t)64 * 32768;
    for (int64_t q0 = 0; q0 < M; q0 += step) {
        int64_t rows = (M - q0 < step) ? M - q0 : step;
        int64_t rows_pad = round_up(rows, 128);
        KmatArgs a{};
        a.xa = Xq + q0 * d; a.xb = X; a.ls = ls; a.kv = kv; a.noise = noise;
        a.xbs = xs; a.xbs_ld = npad;
        a.alpha = want_mean ? alpha : nullptr;
        a.mean_out = want_mean ? mean_out + q0 : nullptr;
        a.out = kstar;
        a.n1 = rows; a.n2 = n; a.d = d; a.ldo = npad; a.rows_pad = rows_pad; a.cols_pad = npad;
        a.store_rows = rows_pad; a.store_cols = npad; a.vec_ok = 1;
        a.y_mean = y_mean; a.y_std = y_std; a.mean_standardised = standardised;
        if (int32_t rc = launch_kmat(stream, kind, a, 1)) return rc;
        if (want_var && M <= SMALL_M) {  // a handful of queries: matrix-vector path (V in rows 64.. of the K* panel)
            double* V = kstar + 64 * npad;
            linv_apply_small_kernel<SMALL_M><<<(unsigned)((n + 7) / 8), 256, 0, stream>>>(Linv, (int)n, (int)npad, kstar, npad,
                                                                                       (int)M, V);
            small_var_kernel<<<(unsigned)M, 256, 0, stream>>>(V, (int)n, (int)npad, kv + noise, y_std * y_std, standardised,
                                                            var_out);
            if (int32_t rc = check_launch("small-M variance")) return rc;
        } else if (want_var) {
            for (int64_t c0 = 0; c0 < rows_pad; c0 += qchunk) {
                const int64_t crows = (rows_pad - c0 < qchunk) ? rows_pad - c0 : qchunk;
                if (int32_t rc = launch_trmm_sumsq(stream, Linv, (int)n, (int)npad, kstar + c0 * npad, npad, crows, q0 + c0, M,
                                                   kv + noise, y_std * y_std, standardised, var_out, partial, counters))
                    return rc;
            }
        }
    }
    return BOBE_OK;
}

// ---- fantasy variance ---------------------------------------------------------------------------------------
namespace {
constexpr int64_t MCCHUNK = 148 * 128;  // MC columns per chunk: one 128-column tile per SM for the TMA trmm route
struct FantasyLayout {
    double *Kmc, *VT, *base, *Kc, *VcT, *delta2, *G, *kc, *acc, *xs;
    int64_t chunk, cpad, bytes;
    bool fits;
};
FantasyLayout fantasy_layout(void* ws, int64_t ws_bytes, int64_t n, int64_t d, int64_t n_mc, int64_t C, bool self) {
    int64_t npad = npad_of(n);
    FantasyLayout l{};
    l.chunk = self ? round_up(n_mc, 64) : round_up(n_mc < MCCHUNK ? n_mc : MCCHUNK, 64);
    l.cpad = self ? l.chunk : round_up(C, 64);
    Carver c(ws);
    l.Kmc = c.take(l.chunk * npad);
    l.VT = c.take(l.chunk * npad);
    l.base = c.take(l.chunk);
    l.Kc = c.take(self ? 0 : l.cpad * npad);
    l.VcT = c.take(self ? 0 : l.cpad * npad);
    l.delta2 = c.take(self ? 0 : l.cpad);
    l.G = c.take(l.cpad * l.chunk);
    l.kc = c.take(l.cpad * l.chunk);
    l.acc = c.take(l.cpad);
    l.xs = c.take(d * npad);
    l.bytes = c.bytes();
    l.fits = ws ? c.fits(ws, ws_bytes) : false;
    return l;
}
}  // namespace

extern "C" int64_t bobe_fantasy_var_workspace_bytes(int64_t n, int64_t d, int64_t n_mc, int64_t C) {
    if (n <= 0 || n_mc <= 0 || d <= 0) return 256;
    // C <= 0 means "the MC points are the candidates"
    return fantasy_layout(nullptr, 0, n, d, n_mc, C, C <= 0).bytes;
}

extern "C" int32_t bobe_fantasy_var(void* stream_, int32_t kind, const double* X, int64_t n, int64_t d,
                                    const double* ls, double kv, double noise, const double* Linv, double y_std,
                                    const double* Xmc, int64_t n_mc, const double* Xcand, int64_t C, int32_t reduce,
                                    double* out, void* ws, int64_t ws_bytes) {
    cudaStream_t stream = (cudaStream_t)stream_;
    const bool self = (Xcand == nullptr);
    if (!X || !ls || !Linv || !Xmc || !out || !ws || n <= 0 || d <= 0 || n_mc <= 0 || (self ? C != n_mc : C <= 0)) {
        set_error("fantasy_var: bad arguments");
        return BOBE_E_ARG;
    }
    FantasyLayout l = fantasy_layout(ws, ws_bytes, n, d, n_mc, C, self);
    if (!l.fits) {
        set_error("fantasy_var: workspace too small (%lld < %lld)", (long long)ws_bytes, (long long)l.bytes);
        return BOBE_E_WORKSPACE;
    }
    const int64_t npad = npad_of(n);
    const double kk = kv + noise;  // kernel_diag(..., include_noise=True), BOBE/gp.py:561,570
    if (int32_t rc = launch_prescale(stream, X, n, d, ls, 0, l.xs, npad, 0, 1)) return rc;
    auto kstar_panel = [&](const double* pts, int64_t rows, int64_t rows_pad, double* Kout) {
        KmatArgs a{};
        a.xa = pts; a.xb = X; a.ls = ls; a.kv = kv; a.noise = noise; a.out = Kout;
        a.xbs = l.xs; a.xbs_ld = npad;
        a.n1 = rows; a.n2 = n; a.d = d; a.ldo = npad; a.rows_pad = rows_pad; a.cols_pad = npad;
        a.store_rows = rows_pad; a.store_cols = npad; a.vec_ok = 1;
        return launch_kmat(stream, kind, a, 1);
    };
    auto apply_linv = [&](const double* Kin, int64_t rows_pad, double* Vout) {  // Vout[j][i] = sum_k Kin[j][k] Linv[i][k]
        // full chunks: the TMA trmm pipeline (one 128-column tile per SM, V stored instead of squared); bitwise the same V
        const int32_t rt = launch_trmm_store(stream, Linv, (int)n, (int)npad, Kin, npad, rows_pad, Vout, npad);
        if (rt <= 0) return rt;
        GemmArgs g{};  // computed as V = Linv * Kin^T with only the transposed store (triangular operand = row operand)
        g.A = Linv; g.Bt = Kin; g.C = nullptr; g.Ct = Vout; g.lda = g.ldb = g.ldct = npad;
        g.M = (int)npad; g.N = (int)rows_pad; g.K = (int)npad; g.alpha = 1.0; g.flags = GEMM_A_LOWER;
        return launch_gemm_nt(stream, g, 1);
    };
    auto rows_sumsq = [&](const double* Vin, int64_t rows, double* o) {
        row_sumsq_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, stream>>>(Vin, npad, rows, npad, kk, o);
        return check_launch("row_sumsq_kernel");
    };
    if (cudaMemsetAsync(l.acc, 0, l.cpad * 8, stream) != cudaSuccess) {
        set_error("fantasy_var: memset failed");
        return BOBE_E_CUDA;
    }
    const double* VcT = l.VcT;
    const double* delta2 = l.delta2;
    if (!self) {
        if (int32_t rc = kstar_panel(Xcand, C, l.cpad, l.Kc)) return rc;
        if (int32_t rc = apply_linv(l.Kc, l.cpad, l.VcT)) return rc;
        if (int32_t rc = rows_sumsq(l.VcT, C, l.delta2)) return rc;
    }
    for (int64_t j0 = 0; j0 < n_mc; j0 += l.chunk) {
        int64_t nj = (n_mc - j0 < l.chunk) ? n_mc - j0 : l.chunk;
        int64_t nj_pad = round_up(nj, 64);
        if (int32_t rc = kstar_panel(Xmc + j0 * d, nj, nj_pad, l.Kmc)) return rc;
        if (int32_t rc = apply_linv(l.Kmc, nj_pad, l.VT)) return rc;
        if (int32_t rc = rows_sumsq(l.VT, nj, l.base)) return rc;
        if (self) {
            VcT = l.VT;
            delta2 = l.base;
        }
        {  // G[c][j] = sum_i VcT[c][i] VT[j][i]
            GemmArgs g{};
            g.A = VcT; g.Bt = l.VT; g.C = l.G; g.lda = g.ldb = npad; g.ldc = l.chunk;
            g.M = (int)l.cpad; g.N = (int)nj_pad; g.K = (int)npad; g.alpha = 1.0;
            if (int32_t rc = launch_gemm_nt(stream, g, 1)) return rc;
        }
        {  // kc[c][j] = k(x_c, mc_j)   (BOBE/gp.py:565-568)
            KmatArgs a{};
            a.xa = self ? Xmc : Xcand; a.xb = Xmc + j0 * d; a.ls = ls; a.kv = kv; a.noise = noise; a.out = l.kc;
            a.n1 = C; a.n2 = nj; a.d = d; a.ldo = l.chunk; a.rows_pad = l.cpad; a.cols_pad = nj_pad;
            a.store_rows = l.cpad; a.store_cols = nj_pad; a.vec_ok = 1;
            if (int32_t rc = launch_kmat(stream, kind, a, 1)) return rc;
        }
        fantasy_combine_kernel<<<(unsigned)C, 256, 0, stream>>>(l.base, l.kc, l.G, l.chunk, delta2, nj, y_std * y_std,
                                                              reduce, out, n_mc, j0, l.acc);
        if (int32_t rc = check_launch("fantasy_combine_kernel")) return rc;
    }
    if (reduce != BOBE_REDUCE_NONE) {
        scale_out_kernel<<<(unsigned)((C + 255) / 256), 256, 0, stream>>>(l.acc, C, 1.0 / (double)n_mc, out);
        return check_launch("scale_out_kernel");
    }
    return BOBE_OK;
}

// ---- rank-1 append --------------------------------------------------------------------------------------
extern "C" int32_t bobe_chol_append(void* stream_, const double* L, int64_t n, int64_t ldl, const double* k,
                                    double k_self, double* L_out, int64_t ldo) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!L_out || n < 0 || ldo < n + 1 || (n > 0 && (!L || !k || ldl < n))) {
        set_error("chol_append: bad arguments");
        return BOBE_E_ARG;
    }
    if (n * 8 > 200 * 1024) {
        set_error("chol_append: n=%lld too large for the single-CTA solve", (long long)n);
        return BOBE_E_ARG;
    }
    if (n > 0) {
        dim3 grid((unsigned)((n + 1 + 255) / 256), (unsigned)n);
        chol_append_copy_kernel<<<grid, 256, 0, stream>>>(L, n, ldl, L_out, ldo);
        if (int32_t rc = check_launch("chol_append_copy_kernel")) return rc;
    }
    int smem = (int)(n * 8 + 16);
    if (int32_t rc = ensure_smem<chol_append_solve_kernel>(smem)) return rc;
    chol_append_solve_kernel<<<1, 1024, smem, stream>>>(L, n, ldl, k, k_self, L_out + n * ldo);
    return check_launch("chol_append_solve_kernel");
}

extern "C" int32_t bobe_acq_ei(void* stream, int32_t which, const double* mean, const double* var, int64_t M,
                               double best_y, double zeta, double* out) {
    if (!mean || !var || !out || M < 0 || (which != BOBE_ACQ_EI && which != BOBE_ACQ_LOGEI)) {
        set_error("acq_ei: bad arguments");
        return BOBE_E_ARG;
    }
    if (M == 0) return BOBE_OK;
    acq_ei_kernel<<<(unsigned)((M + 255) / 256), 256, 0, (cudaStream_t)stream>>>(which, mean, var, M, best_y, zeta, out);
    return check_launch("acq_ei_kernel");
}

// ---- SVM feasibility mask (GPwithClassifier) ------------------------------------------------------------------
namespace {
__global__ void fill_kernel(double* p, int64_t n, double v) {
    int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (i < n) p[i] = v;
}
// jnp.where(clf_probs >= threshold, value, fill) with clf_probs = (decision >= 0), BOBE/clf_gp.py:173-205
__global__ void svm_mask_kernel(const double* __restrict__ dec, int64_t M, double minus_inf, double var_fill,
                                double* __restrict__ mean, double* __restrict__ var) {
    int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= M) return;
    const bool feasible = dec[i] >= 0.0;
    if (mean && !feasible) mean[i] = minus_inf;
    if (var && !feasible) var[i] = var_fill;
}
}  // namespace

extern "C" int64_t bobe_svm_mask_workspace_bytes(int64_t d, int64_t M) {
    if (d <= 0 || M <= 0) return 256;
    return (round_up(d, 32) + round_up(M, 32)) * 8 + 256;
}

extern "C" int32_t bobe_svm_mask(void* stream_, const double* sv, int64_t n_sv, int64_t d, const double* dual_coef,
                                 double intercept, double gamma, const double* Xq, int64_t M, double minus_inf,
                                 double var_fill, double* mean_inout, double* var_inout, double* decision_out, void* ws,
                                 int64_t ws_bytes) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!sv || !dual_coef || !Xq || !ws || n_sv <= 0 || d <= 0 || d > BOBE_MAX_DIM || M < 0 || !(gamma > 0.0)) {
        set_error("svm_mask: bad arguments");
        return BOBE_E_ARG;
    }
    if (M == 0) return BOBE_OK;
    if (ws_bytes < bobe_svm_mask_workspace_bytes(d, M)) {
        set_error("svm_mask: workspace too small");
        return BOBE_E_WORKSPACE;
    }
    double* ls = align256(ws);
    double* dec = decision_out ? decision_out : ls + round_up(d, 32);
    // exp(-gamma |x - s|^2) = exp(-1/2 |x - s|^2 / l^2) with l = 1 / sqrt(2 gamma): the SVM decision function
    // (BOBE/clf.py:188-209) is an isotropic RBF kernel-row product, i.e. the fused mean pass of the kernel-matrix kernel
    fill_kernel<<<(unsigned)((d + 255) / 256), 256, 0, stream>>>(ls, d, 1.0 / sqrt(2.0 * gamma));
    if (int32_t rc = check_launch("fill_kernel")) return rc;
    const int64_t RCH = 64 * 32768;
    for (int64_t r0 = 0; r0 < M; r0 += RCH) {
        const int64_t rows = (M - r0 < RCH) ? M - r0 : RCH;
        KmatArgs a{};
        a.xa = Xq + r0 * d; a.xb = sv; a.ls = ls; a.kv = 1.0; a.alpha = dual_coef; a.mean_out = dec + r0;
        a.n1 = rows; a.n2 = n_sv; a.d = d; a.rows_pad = round_up(rows, 64); a.cols_pad = round_up(n_sv, 64);
        a.y_mean = intercept; a.y_std = 1.0; a.mean_standardised = 0;
        if (int32_t rc = launch_kmat(stream, BOBE_KERNEL_RBF, a, 1)) return rc;
    }
    svm_mask_kernel<<<(unsigned)((M + 255) / 256), 256, 0, stream>>>(dec, M, minus_inf, var_fill, mean_inout, var_inout);
    return check_launch("svm_mask_kernel");
}

extern "C" int32_t bobe_bench_trmm_sumsq(void* stream, const double* Linv, int64_t n, const double* kstar,
                                         int64_t rows_pad, double kk, double* var_out) {
    if (!Linv || !kstar || !var_out || n <= 0 || rows_pad <= 0) {
        set_error("bench_trmm_sumsq: bad arguments");
        return BOBE_E_ARG;
    }
    const int64_t npad = npad_of(n);
    // the scratch bobe_predict takes from its workspace (benchmark entry: allocated once, never freed), so that the launch
    // timed here is the one the product makes
    static double* scratch = nullptr;
    static int* counters = nullptr;
    if (!scratch && rows_pad <= QCHUNK) {
        if (cudaMalloc(&scratch, 8 * QCHUNK * sizeof(double)) != cudaSuccess ||
            cudaMalloc(&counters, TRMM_COUNTERS * sizeof(int)) != cudaSuccess ||
            cudaMemset(counters, 0, TRMM_COUNTERS * sizeof(int)) != cudaSuccess) {
            set_error("bench_trmm_sumsq: scratch allocation failed");
            return BOBE_E_CUDA;
        }
    }
    const bool fits = scratch && rows_pad <= QCHUNK && rows_pad >= 74 * 128;  // (below that the row split would engage)
    return launch_trmm_sumsq((cudaStream_t)stream, Linv, (int)n, (int)npad, kstar, npad, rows_pad, 0, rows_pad, kk, 1.0, 0,
                             var_out, fits ? scratch : nullptr, fits ? counters : nullptr);
}
