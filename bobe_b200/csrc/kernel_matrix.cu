// Fused kernel-matrix builds: K(X,X)+noise*I for the factorisation (batched over hyper-parameter sets) and
// K(X*,X) panels for prediction, with the predictive mean (alpha-dot) folded into the same pass.
// Restates BOBE/gp.py:80-96 (dist_sq by direct differences), :124-154 (RBF), :156-168 (Matern-5/2).
#include "kernels.cuh"

namespace bobe {

constexpr int KT = 64;    // tile edge
constexpr int KLD = 66;   // smem row stride (doubles): even, so 16-byte vector reads stay aligned

template <int KIND>
__global__ void __launch_bounds__(256) kmat_kernel(KmatArgs p) {
    extern __shared__ __align__(16) double sm[];
    const int d = (int)p.d;
    double* sa = sm;                 // [d][KLD] scaled rows of xa
    double* sb = sm + d * KLD;       // [d][KLD] scaled rows of xb
    double* sal = sb + d * KLD;      // [KT] alpha tile
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int64_t z = blockIdx.z;
    const double* ls = p.ls + z * p.ls_stride;
    const double kv = p.kv_ptr ? p.kv_ptr[z] : p.kv;
    double* out = p.out ? p.out + z * p.out_stride : nullptr;
    if (p.gate && p.gate[z] == 0) return;
    const double* alpha = p.alpha ? p.alpha + z * p.alpha_stride : nullptr;
    const int64_t i0 = (int64_t)blockIdx.y * KT;

    for (int idx = tid; idx < KT * d; idx += 256) {
        int r = idx / d, k = idx - r * d;
        int64_t row = i0 + r;
        sa[k * KLD + r] = row < p.n1 ? p.xa[row * d + k] / ls[k] : 0.0;
    }
    double macc[4] = {0.0, 0.0, 0.0, 0.0};
    const int64_t col_tiles = p.cols_pad / KT;
    for (int64_t ct = blockIdx.x; ct < col_tiles; ct += gridDim.x) {
        const int64_t j0 = ct * KT;
        __syncthreads();  // previous tile's readers are done with sb / sal
        for (int idx = tid; idx < KT * d; idx += 256) {
            int r = idx / d, k = idx - r * d;
            int64_t col = j0 + r;
            sb[k * KLD + r] = col < p.n2 ? p.xb[col * d + k] / ls[k] : 0.0;
        }
        if (alpha && tid < KT) sal[tid] = (j0 + tid < p.n2) ? alpha[j0 + tid] : 0.0;
        __syncthreads();

        double q[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) q[i][j] = 0.0;
#pragma unroll 4
        for (int k = 0; k < d; ++k) {
            double2 a01 = *reinterpret_cast<const double2*>(sa + k * KLD + ty * 4);
            double2 a23 = *reinterpret_cast<const double2*>(sa + k * KLD + ty * 4 + 2);
            double2 b01 = *reinterpret_cast<const double2*>(sb + k * KLD + tx * 4);
            double2 b23 = *reinterpret_cast<const double2*>(sb + k * KLD + tx * 4 + 2);
            double a[4] = {a01.x, a01.y, a23.x, a23.y}, b[4] = {b01.x, b01.y, b23.x, b23.y};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    double df = a[i] - b[j];
                    q[i][j] = fma(df, df, q[i][j]);
                }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            int64_t row = i0 + ty * 4 + i;
            double v[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                int64_t col = j0 + tx * 4 + j;
                if (row < p.n1 && col < p.n2) {
                    v[j] = kernel_from_q<KIND>(q[i][j], kv);
                    if (p.add_noise && row == col) v[j] += p.noise;
                } else {
                    v[j] = (p.pad_identity && row == col) ? 1.0 : 0.0;
                }
            }
            if (p.alpha) {
#pragma unroll
                for (int j = 0; j < 4; ++j) macc[i] = fma(sal[tx * 4 + j], v[j], macc[i]);
            }
            if (out && row < p.store_rows) {
                int64_t c0 = j0 + tx * 4;
                double* dstp = out + row * p.ldo + c0;
                if (c0 + 3 < p.store_cols && p.vec_ok) {
                    double2* dst = reinterpret_cast<double2*>(dstp);
                    dst[0] = make_double2(v[0], v[1]);
                    dst[1] = make_double2(v[2], v[3]);
                } else {
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        if (c0 + j < p.store_cols) dstp[j] = v[j];
                }
            }
        }
    }
    if (p.alpha && p.mean_out) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            double m = macc[i];
            m += __shfl_xor_sync(0xffffffffu, m, 8);
            m += __shfl_xor_sync(0xffffffffu, m, 4);
            m += __shfl_xor_sync(0xffffffffu, m, 2);
            m += __shfl_xor_sync(0xffffffffu, m, 1);
            int64_t row = i0 + ty * 4 + i;
            if (tx == 0 && row < p.n1)  // BOBE/gp.py:456 (un-standardised) / :483 (standardised)
                p.mean_out[z * p.mean_stride + row] = p.mean_standardised ? m : m * p.y_std + p.y_mean;
        }
    }
}

int32_t launch_kmat(cudaStream_t stream, int kind, const KmatArgs& a, int batch) {
    if (a.rows_pad <= 0 || a.cols_pad <= 0 || batch <= 0) return BOBE_OK;
    if (a.rows_pad % KT || a.cols_pad % KT) {
        set_error("kmat: padded extents must be multiples of %d", KT);
        return BOBE_E_ARG;
    }
    if (a.d < 1 || a.d > 200) {
        set_error("kmat: d=%lld unsupported (1..200)", (long long)a.d);
        return BOBE_E_ARG;
    }
    int smem = (int)((2 * a.d * KLD + KT) * sizeof(double));
    int64_t row_tiles = a.rows_pad / KT, col_tiles = a.cols_pad / KT;
    int64_t splits = 1;
    if (!a.alpha) {  // spread columns over CTAs until the grid covers the machine a few times over
        while (splits < col_tiles && row_tiles * splits * batch < 148 * 4) splits *= 2;
        if (splits > col_tiles) splits = col_tiles;
    }
    if (row_tiles > 65535) {
        set_error("kmat: too many row tiles (%lld); chunk the call", (long long)row_tiles);
        return BOBE_E_ARG;
    }
    dim3 grid((unsigned)splits, (unsigned)row_tiles, (unsigned)batch);
    if (kind == BOBE_KERNEL_RBF) {
        if (int32_t rc = ensure_smem<kmat_kernel<BOBE_KERNEL_RBF>>(smem)) return rc;
        kmat_kernel<BOBE_KERNEL_RBF><<<grid, 256, smem, stream>>>(a);
    } else {
        if (int32_t rc = ensure_smem<kmat_kernel<BOBE_KERNEL_MATERN52>>(smem)) return rc;
        kmat_kernel<BOBE_KERNEL_MATERN52><<<grid, 256, smem, stream>>>(a);
    }
    return check_launch("kmat_kernel");
}

}  // namespace bobe
