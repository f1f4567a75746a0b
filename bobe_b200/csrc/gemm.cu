// Launchers for the DMMA NT-GEMM and the fused "triangular multiply + column sum of squares" kernel that
// produces the predictive variance term ||L^-1 k*||^2 without ever writing V = L^-1 K* to memory.
#include <cstdarg>
#include <mutex>

#include "gemm_nt.cuh"
#include "kernels.cuh"

namespace bobe {

// ---- error plumbing ------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
const char* last_error() { return g_err; }
int32_t check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("%s: %s", what, cudaGetErrorString(e));
        return BOBE_E_CUDA;
    }
    return BOBE_OK;
}

template <class K>
static int32_t ensure_smem(K kernel, int bytes) {
    // idempotent and cheap; set on every call so that it holds for whichever device is current
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e != cudaSuccess) {
        set_error("cudaFuncSetAttribute(smem=%d): %s", bytes, cudaGetErrorString(e));
        return BOBE_E_CUDA;
    }
    return BOBE_OK;
}

int32_t launch_gemm_nt(cudaStream_t stream, const GemmArgs& a, int batch) {
    if (a.M <= 0 || a.N <= 0 || batch <= 0) return BOBE_OK;
    if ((a.K % 16) || (a.N % 2) || (a.lda % 2) || (a.ldb % 2) || (a.ldc % 2)) {
        set_error("gemm_nt: K must be a multiple of 16 and N/ld even (M=%d N=%d K=%d)", a.M, a.N, a.K);
        return BOBE_E_ARG;
    }
    bool small = (a.M <= 64 || a.N <= 64);
    if (small) {
        using Cfg = CfgSmall;
        if (int32_t rc = ensure_smem(gemm_nt_kernel<Cfg>, Cfg::SMEM_BYTES)) return rc;
        dim3 grid((a.N + Cfg::BN - 1) / Cfg::BN, (a.M + Cfg::BM - 1) / Cfg::BM, batch);
        gemm_nt_kernel<Cfg><<<grid, Cfg::THREADS, Cfg::SMEM_BYTES, stream>>>(a);
    } else {
        using Cfg = CfgBig;
        if (int32_t rc = ensure_smem(gemm_nt_kernel<Cfg>, Cfg::SMEM_BYTES)) return rc;
        dim3 grid((a.N + Cfg::BN - 1) / Cfg::BN, (a.M + Cfg::BM - 1) / Cfg::BM, batch);
        gemm_nt_kernel<Cfg><<<grid, Cfg::THREADS, Cfg::SMEM_BYTES, stream>>>(a);
    }
    return check_launch("gemm_nt_kernel");
}

// ---- predictive variance: var_j = kk - sum_i ( sum_{k<=i} Linv[i][k] Kstar[j][k] )^2 -----------------------
// One CTA owns BN queries and sweeps all row blocks of Linv (a lower-triangular NT product), squaring and
// summing each finished BM x BN block of V into per-query registers.  V never leaves the SM.
template <class Cfg>
__global__ void __launch_bounds__(Cfg::THREADS, 1)
    trmm_sumsq_kernel(const double* __restrict__ Linv, int npad, const double* __restrict__ Kstar, int64_t ldk,
                      int64_t q_begin, int64_t M, double kk, double scale, int standardised,
                      double* __restrict__ var_out) {
    extern __shared__ __align__(16) double smem[];
    __shared__ double red[Cfg::WM][Cfg::BN];
    const int j0 = blockIdx.x * Cfg::BN;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int wm = warp / Cfg::WN, wn = warp % Cfg::WN;
    const double* Bt = Kstar + (int64_t)j0 * ldk;

    double colsum[Cfg::NF][2];
#pragma unroll
    for (int nf = 0; nf < Cfg::NF; ++nf) colsum[nf][0] = colsum[nf][1] = 0.0;

    for (int i0 = 0; i0 < npad; i0 += Cfg::BM) {
        double acc[Cfg::MF][Cfg::NF][2];
#pragma unroll
        for (int mf = 0; mf < Cfg::MF; ++mf)
#pragma unroll
            for (int nf = 0; nf < Cfg::NF; ++nf) acc[mf][nf][0] = acc[mf][nf][1] = 0.0;
        int ke = min(npad, i0 + Cfg::BM);
        Mainloop<Cfg>::run(acc, Linv + (int64_t)i0 * npad, npad, min(Cfg::BM, npad - i0), Bt, ldk, Cfg::BN, 0, ke, smem);
#pragma unroll
        for (int mf = 0; mf < Cfg::MF; ++mf)
#pragma unroll
            for (int nf = 0; nf < Cfg::NF; ++nf) {
                colsum[nf][0] = fma(acc[mf][nf][0], acc[mf][nf][0], colsum[nf][0]);
                colsum[nf][1] = fma(acc[mf][nf][1], acc[mf][nf][1], colsum[nf][1]);
            }
    }
    // reduce over the 8 row groups of the warp (lanes differing in g), then over the WM warps along M
#pragma unroll
    for (int nf = 0; nf < Cfg::NF; ++nf)
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            double v = colsum[nf][c];
            v += __shfl_xor_sync(0xffffffffu, v, 4);
            v += __shfl_xor_sync(0xffffffffu, v, 8);
            v += __shfl_xor_sync(0xffffffffu, v, 16);
            if (g == 0) red[wm][wn * Cfg::WTN + nf * 8 + 2 * t + c] = v;
        }
    __syncthreads();
    for (int c = threadIdx.x; c < Cfg::BN; c += Cfg::THREADS) {
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < Cfg::WM; ++w) s += red[w][c];
        int64_t q = q_begin + j0 + c;
        if (q < M) {
            double var = kk - s;
            if (standardised) {  // predict_single, BOBE/gp.py:487-488: NaN -> floor, then < floor -> floor
                if (isnan(var)) var = SAFE_FLOOR;
                if (var < SAFE_FLOOR) var = SAFE_FLOOR;
            } else {  // predict_var_single, BOBE/gp.py:465-466: clip (NaN propagates), times y_std^2
                if (var < SAFE_FLOOR) var = SAFE_FLOOR;
                var *= scale;
            }
            var_out[q] = var;
        }
    }
}

int32_t launch_trmm_sumsq(cudaStream_t stream, const double* Linv, int npad, const double* Kstar, int64_t ldk,
                          int64_t rows_pad, int64_t q_begin, int64_t M, double kk, double scale, int standardised,
                          double* var_out) {
    using Cfg = CfgBig;
    if (rows_pad % Cfg::BN) {
        set_error("trmm_sumsq: query chunk must be padded to %d", Cfg::BN);
        return BOBE_E_ARG;
    }
    if (int32_t rc = ensure_smem(trmm_sumsq_kernel<Cfg>, Cfg::SMEM_BYTES)) return rc;
    dim3 grid(rows_pad / Cfg::BN);
    trmm_sumsq_kernel<Cfg><<<grid, Cfg::THREADS, Cfg::SMEM_BYTES, stream>>>(Linv, npad, Kstar, ldk, q_begin, M, kk,
                                                                         scale, standardised, var_out);
    return check_launch("trmm_sumsq_kernel");
}

}  // namespace bobe
