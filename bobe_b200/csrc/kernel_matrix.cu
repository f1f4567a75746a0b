// Fused kernel-matrix builds: K(X,X)+noise*I for the factorisation (batched over hyper-parameter sets) and
// K(X*,X) panels for prediction, with the predictive mean (alpha-dot) folded into the same pass.
// Restates BOBE/gp.py:80-96 (dist_sq by direct differences), :124-154 (RBF), :156-168 (Matern-5/2).
#include <type_traits>

#include "kernels.cuh"

namespace bobe {

constexpr int KT = 64;    // tile edge (rows and columns)
constexpr int KLD = 66;   // smem row stride (doubles): even, so 16-byte vector reads stay aligned
constexpr int KTHREADS = 128;

// One CTA (128 threads as 8 x 16) computes 64 x 64 tiles of K: thread (ty, tx) owns rows 8 ty .. 8 ty + 7 and the
// columns {2 tx, 2 tx + 1, 32 + 2 tx, 33 + 2 tx} -- 32 independent distance accumulators per thread, fed per
// input dimension by six LDS.128 (four of them warp-wide broadcasts of the row coordinates, the two column loads
// read 256 contiguous bytes per half-warp: conflict-free).  The FP64 pipe does 2 instructions per (element, dim)
// plus the exp / sqrt polynomials; everything else is kept off it:
//   * the scaled, transposed row coordinates x/l are staged once per CTA, the column coordinates once per tile;
//   * the raw rows of the NEXT column tile stream into shared memory with cp.async while the current tile is
//     computed, so the global-load latency never sits between two tiles;
//   * lower_only (symmetric K for the factorisation): tiles strictly above the diagonal are skipped.
//   * PRE: the column operand was scaled and transposed once by prescale_kernel (xbs[k][col] = xb[col][k] / l_k),
//     so a column tile is d rows of 512 contiguous bytes that cp.async drops straight into the (double-buffered)
//     compute layout: no per-tile divisions, one __syncthreads per tile.  Without PRE (the generic entry point,
//     which has no workspace) the raw rows are staged and divided in the kernel.
//   * KROWS = rows of a thread's 8 x 4 tile whose kernel values are evaluated together.  KROWS = 4 (16 values in lock
//     step, ~240 registers, 2 CTAs/SM) is the faster one when the grid cannot put more than two CTAs on an SM anyway
//     (a predict chunk is exactly 296 row tiles); KROWS = 2 fits 168 registers = 3 CTAs/SM and wins on large grids
//     (mean-only sweeps, the batched K(X,X) builds).
template <int KIND, bool PRE, int KROWS>
__global__ void __launch_bounds__(KTHREADS, KROWS == 4 ? 2 : 3) kmat_kernel(KmatArgs p) {
    extern __shared__ __align__(16) double sm[];
    const int d = (int)p.d;
    double* sa = sm;                 // [d][KLD] scaled rows of xa
    double* sb = sa + d * KLD;       // [d][KLD] scaled rows of xb  (PRE: two buffers)
    double* raw = sb + d * KLD;      // !PRE: [KT * d] raw rows of xb for the next tile;  PRE: second sb buffer
    double* sal = raw + (PRE ? d * KLD : KT * d);  // [2][KT] alpha tile
    double* sls = sal + 2 * KT;      // [d] lengthscales
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int64_t z = blockIdx.z;
    if (p.gate && p.gate[z] == 0) return;
    const double* ls = p.ls + z * p.ls_stride;
    const double kv = p.kv_ptr ? p.kv_ptr[z] : p.kv;
    double* out = p.out ? p.out + z * p.out_stride : nullptr;
    const double* alpha = p.alpha ? p.alpha + z * p.alpha_stride : nullptr;
    const int64_t i0 = (int64_t)blockIdx.y * KT;
    int64_t col_tiles = p.cols_pad / KT;
    if (p.lower_only && (int64_t)blockIdx.y + 1 < col_tiles) col_tiles = (int64_t)blockIdx.y + 1;
    if ((int64_t)blockIdx.x >= col_tiles) return;

    const double* xbs = PRE ? p.xbs + z * p.xbs_stride : nullptr;
    // !PRE: raw rows j0 .. j0+63 of xb are 64*d contiguous doubles; rows >= n2 are zero-filled.
    // PRE: d rows of 64 scaled coordinates into compute buffer `buf`, plus that tile's alpha values.
    auto prefetch = [&](int64_t j0, int buf) {
        if (PRE) {
            double* dst = buf ? raw : sb;
            for (int idx = tid; idx < d * (KT / 2); idx += KTHREADS) {
                const int k = idx >> 5, c = idx & 31;
                cp_async16(dst + k * KLD + 2 * c, xbs + (int64_t)k * p.xbs_ld + j0 + 2 * c, true);
            }
            if (alpha && tid < KT) {
                uint32_t s = static_cast<uint32_t>(__cvta_generic_to_shared(sal + buf * KT + tid));
                bool ok = j0 + tid < p.n2;
                int sz = ok ? 8 : 0;
                asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(s), "l"(ok ? alpha + j0 + tid : alpha),
                             "r"(sz));
            }
        } else {
            const int64_t valid = (p.n2 - j0 < KT ? p.n2 - j0 : (int64_t)KT) * d;  // doubles that exist
            const double* src = p.xb + j0 * d;
            for (int idx = tid; idx < KT * d; idx += KTHREADS) {
                uint32_t s = static_cast<uint32_t>(__cvta_generic_to_shared(raw + idx));
                bool ok = idx < valid;
                int sz = ok ? 8 : 0;
                asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(s), "l"(ok ? src + idx : p.xb), "r"(sz));
            }
        }
        cp_async_commit();
    };
    prefetch((int64_t)blockIdx.x * KT, 0);
    for (int k = tid; k < d; k += KTHREADS) sls[k] = ls[k];
    __syncthreads();
    for (int r = ty; r < KT; r += KTHREADS / 16)
        for (int k = tx; k < d; k += 16) {
            int64_t row = i0 + r;
            sa[k * KLD + r] = row < p.n1 ? p.xa[row * d + k] / sls[k] : 0.0;
        }

    double macc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) macc[i] = 0.0;
    int buf = 0;
    for (int64_t ct = blockIdx.x; ct < col_tiles; ct += gridDim.x, buf ^= PRE ? 1 : 0) {
        const int64_t j0 = ct * KT;
        cp_async_wait<0>();
        __syncthreads();  // this tile has landed; everyone is done reading the previous tile's buffers
        if (!PRE) {
            for (int r = ty; r < KT; r += KTHREADS / 16)
                for (int k = tx; k < d; k += 16) sb[k * KLD + r] = raw[r * d + k] / sls[k];
            if (alpha && tid < KT) sal[tid] = (j0 + tid < p.n2) ? alpha[j0 + tid] : 0.0;
            __syncthreads();
        }
        if (ct + gridDim.x < col_tiles) prefetch((ct + gridDim.x) * KT, buf ^ 1);
        const double* sbt = (PRE && buf) ? raw : sb;
        const double* salt = sal + (PRE ? buf * KT : 0);

        double q[8][4];
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) q[i][j] = 0.0;
        // register double buffer: the six LDS.128 of dimension k+1 are in flight while dimension k is consumed
        const double* ap = sa + ty * 8;
        const double* bp = sbt + 2 * tx;
        double2 a01 = *reinterpret_cast<const double2*>(ap), a23 = *reinterpret_cast<const double2*>(ap + 2);
        double2 a45 = *reinterpret_cast<const double2*>(ap + 4), a67 = *reinterpret_cast<const double2*>(ap + 6);
        double2 b01 = *reinterpret_cast<const double2*>(bp), b23 = *reinterpret_cast<const double2*>(bp + 32);
#pragma unroll 2
        for (int k = 0; k < d; ++k) {
            const int kn = (k + 1 < d ? k + 1 : k) * KLD;
            const double2 n01 = *reinterpret_cast<const double2*>(ap + kn);
            const double2 n23 = *reinterpret_cast<const double2*>(ap + kn + 2);
            const double2 n45 = *reinterpret_cast<const double2*>(ap + kn + 4);
            const double2 n67 = *reinterpret_cast<const double2*>(ap + kn + 6);
            const double2 m01 = *reinterpret_cast<const double2*>(bp + kn);
            const double2 m23 = *reinterpret_cast<const double2*>(bp + kn + 32);
            double a[8] = {a01.x, a01.y, a23.x, a23.y, a45.x, a45.y, a67.x, a67.y};
            double b[4] = {b01.x, b01.y, b23.x, b23.y};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    double df = a[i] - b[j];
                    q[i][j] = fma(df, df, q[i][j]);
                }
            a01 = n01; a23 = n23; a45 = n45; a67 = n67; b01 = m01; b23 = m23;
        }
        double al[4] = {0.0, 0.0, 0.0, 0.0};
        if (alpha) {
            al[0] = salt[2 * tx]; al[1] = salt[2 * tx + 1]; al[2] = salt[32 + 2 * tx]; al[3] = salt[33 + 2 * tx];
        }
        const int64_t c0 = j0 + 2 * tx, c1 = c0 + 32;
        // interior tiles (all 64 x 64 elements inside (n1, n2), no diagonal term) take the check-free path
        const bool interior = i0 + KT <= p.n1 && j0 + KT <= p.n2 && !(p.add_noise && i0 == j0);
#pragma unroll
        for (int ip = 0; ip < 8; ip += KROWS) {  // KROWS rows (4 KROWS values) per lock-step evaluation
            double qq[4 * KROWS], v[4 * KROWS];
#pragma unroll
            for (int h = 0; h < KROWS; ++h)
#pragma unroll
                for (int j = 0; j < 4; ++j) qq[4 * h + j] = q[ip + h][j];
            kernel_from_q_n<KIND, 4 * KROWS>(qq, kv, v);
#pragma unroll
            for (int h = 0; h < KROWS; ++h) {
                const int i = ip + h;
                const int64_t row = i0 + ty * 8 + i;
                double* vr = v + 4 * h;
                if (!interior) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int64_t col = (j < 2 ? c0 : c1 - 2) + j;
                        const bool inb = row < p.n1 && col < p.n2;
                        const double pad = (p.pad_identity && row == col) ? 1.0 : 0.0;
                        const double diag = (p.add_noise && row == col) ? p.noise : 0.0;
                        vr[j] = inb ? vr[j] + diag : pad;
                    }
                }
                if (alpha) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) macc[i] = fma(al[j], vr[j], macc[i]);
                }
                if (out && row < p.store_rows) {
                    double* dstp = out + row * p.ldo;
                    if (p.vec_ok && c1 + 1 < p.store_cols) {
                        *reinterpret_cast<double2*>(dstp + c0) = make_double2(vr[0], vr[1]);
                        *reinterpret_cast<double2*>(dstp + c1) = make_double2(vr[2], vr[3]);
                    } else {
                        if (c0 < p.store_cols) dstp[c0] = vr[0];
                        if (c0 + 1 < p.store_cols) dstp[c0 + 1] = vr[1];
                        if (c1 < p.store_cols) dstp[c1] = vr[2];
                        if (c1 + 1 < p.store_cols) dstp[c1 + 1] = vr[3];
                    }
                }
            }
        }
    }
    if (alpha && p.mean_out) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            double m = macc[i];
            m += __shfl_xor_sync(0xffffffffu, m, 8);
            m += __shfl_xor_sync(0xffffffffu, m, 4);
            m += __shfl_xor_sync(0xffffffffu, m, 2);
            m += __shfl_xor_sync(0xffffffffu, m, 1);
            int64_t row = i0 + ty * 8 + i;
            if (tx == 0 && row < p.n1)  // BOBE/gp.py:456 (un-standardised) / :483 (standardised)
                p.mean_out[z * p.mean_stride + row] = p.mean_standardised ? m : m * p.y_std + p.y_mean;
        }
    }
}

// xs[z][k][j] = x[j][k] / ls[z][k] for j < n, 0 for n <= j < ld   (the division the reference does at BOBE/gp.py:149,
// 160, done once per hyper-parameter set instead of once per tile)
__global__ void __launch_bounds__(256) prescale_kernel(const double* __restrict__ x, int64_t n, int d,
                                                       const double* __restrict__ ls, int64_t ls_stride,
                                                       double* __restrict__ xs, int64_t ld, int64_t xs_stride) {
    const int64_t j = (int64_t)blockIdx.x * 256 + threadIdx.x;
    const int64_t z = blockIdx.y;
    if (j >= ld) return;
    const double* l = ls + z * ls_stride;
    double* o = xs + z * xs_stride + j;
    for (int k = 0; k < d; ++k) o[(int64_t)k * ld] = j < n ? x[j * d + k] / l[k] : 0.0;
}

int32_t launch_prescale(cudaStream_t stream, const double* x, int64_t n, int64_t d, const double* ls, int64_t ls_stride,
                        double* xs, int64_t ld, int64_t xs_stride, int batch) {
    if (ld <= 0 || batch <= 0) return BOBE_OK;
    dim3 grid((unsigned)((ld + 255) / 256), (unsigned)batch);
    prescale_kernel<<<grid, 256, 0, stream>>>(x, n, (int)d, ls, ls_stride, xs, ld, xs_stride);
    return check_launch("prescale_kernel");
}

int32_t launch_kmat(cudaStream_t stream, int kind, const KmatArgs& a, int batch) {
    if (a.rows_pad <= 0 || a.cols_pad <= 0 || batch <= 0) return BOBE_OK;
    if (a.rows_pad % KT || a.cols_pad % KT) {
        set_error("kmat: padded extents must be multiples of %d", KT);
        return BOBE_E_ARG;
    }
    if (a.d < 1 || a.d > BOBE_MAX_DIM) {
        set_error("kmat: d=%lld unsupported (1..%d)", (long long)a.d, BOBE_MAX_DIM);
        return BOBE_E_ARG;
    }
    const bool pre = a.xbs != nullptr;
    if (pre && ((a.xbs_ld % 2) || a.xbs_ld < a.cols_pad || (((uintptr_t)a.xbs) & 15) || (a.xbs_stride % 2))) {
        set_error("kmat: prescaled operand must be 16-byte aligned with an even leading dimension >= cols_pad");
        return BOBE_E_ARG;
    }
    int smem = (int)((2 * a.d * KLD + (pre ? a.d * KLD : KT * a.d) + 2 * KT + a.d) * sizeof(double));
    int64_t row_tiles = a.rows_pad / KT, col_tiles = a.cols_pad / KT;
    int64_t splits = 1;
    if (!a.alpha) {  // spread columns over CTAs until the grid covers the machine a few times over
        while (splits < col_tiles && row_tiles * splits * batch < 148 * 4) splits *= 2;
        if (a.lower_only && splits < col_tiles / 4) splits = col_tiles / 4;  // row tile i only has i+1 tiles: keep CTAs short
        if (splits > col_tiles) splits = col_tiles;
    }
    if (a.lower_only && (a.rows_pad != a.cols_pad || a.alpha)) {
        set_error("kmat: lower_only needs a square build without the mean epilogue");
        return BOBE_E_ARG;
    }
    if (row_tiles > 65535) {
        set_error("kmat: too many row tiles (%lld); chunk the call", (long long)row_tiles);
        return BOBE_E_ARG;
    }
    dim3 grid((unsigned)splits, (unsigned)row_tiles, (unsigned)batch);
    // three CTAs per SM only materialise when the grid offers them
    const bool wide = (int64_t)grid.x * grid.y * grid.z >= 3 * 148;
    auto go = [&](auto kind_c, auto pre_c, auto kr_c) -> int32_t {
        constexpr int K_ = decltype(kind_c)::value;
        constexpr bool P_ = decltype(pre_c)::value;
        constexpr int R_ = decltype(kr_c)::value;
        if (int32_t rc = ensure_smem<kmat_kernel<K_, P_, R_>>(smem)) return rc;
        kmat_kernel<K_, P_, R_><<<grid, KTHREADS, smem, stream>>>(a);
        return BOBE_OK;
    };
    using std::integral_constant;
    auto by_rows = [&](auto kind_c, auto pre_c) -> int32_t {
        return wide ? go(kind_c, pre_c, integral_constant<int, 2>{}) : go(kind_c, pre_c, integral_constant<int, 4>{});
    };
    auto by_pre = [&](auto kind_c) -> int32_t {
        return pre ? by_rows(kind_c, integral_constant<bool, true>{}) : by_rows(kind_c, integral_constant<bool, false>{});
    };
    int32_t rc = kind == BOBE_KERNEL_RBF ? by_pre(integral_constant<int, BOBE_KERNEL_RBF>{})
                                         : by_pre(integral_constant<int, BOBE_KERNEL_MATERN52>{});
    if (rc) return rc;
    return check_launch("kmat_kernel");
}

}  // namespace bobe
