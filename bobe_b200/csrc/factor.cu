// Recursive Cholesky + triangular inverse, batched over hyper-parameter sets (grid.z).
//
//   chol_inv([A11 . ; A21 A22]):  chol_inv(A11);  L21 = A21 Linv11^T;  A22 -= L21 L21^T;  chol_inv(A22);
//                                 Linv21 = -(Linv22 L21) Linv11
//
// Every product is an NT GEMM on the FP64 tensor pipe (gemm_nt.cuh); each result is stored in both
// orientations (L and Lt, Linv and U) so that the next product is again NT.  Leaves are NB x NB blocks
// factorised and inverted by one CTA in shared memory.  Replaces jnp.linalg.cholesky + cho_solve +
// solve_triangular at BOBE/gp.py:175-176,259-260,549-550 and supplies the L^-1 the predictive variance
// (BOBE/gp.py:462,484) and the fantasy variance (:571) are built on.  A non-PD matrix yields NaN outputs
// and info = 1 for that batch entry only, never an error.
#include "gemm_nt.cuh"
#include "kernels.cuh"
#include "leaf.cuh"

namespace bobe {

__global__ void init_stat_kernel(double* dstat, int* gate, int batch, int force) {
    int z = blockIdx.x * blockDim.x + threadIdx.x;
    if (z < batch) {
        dstat[2 * z] = 1e300;
        dstat[2 * z + 1] = 0.0;
        gate[z] = force;
    }
}

int factor_live_rows(int64_t n) {
    static const bool live = env_int("BOBE_FACTOR_LIVE", 1) != 0;  // 0: every product on the padded size (round-1 behaviour)
    static const int gran = (env_int("BOBE_TILE", 0) > 2 || env_int("BOBE_GEMM_TMA", 0) != 0) ? 32 : 16;
    const int64_t npad = npad_of(n);
    if (!live) return (int)npad;
    const int64_t nl = round_up(n, gran);
    return (int)(nl < npad ? nl : npad);
}

int64_t factor_q_elems(int64_t npad) {  // >= npad^2 / 4 (inverse tree, phase 2) and >= 128 npad (panel correction)
    int64_t m = ((npad / NB + 1) / 2) * NB;
    return m * m > 128 * npad ? m * m : 128 * npad;
}

namespace {
struct Rec {
    cudaStream_t stream;
    const FactorBuffers& fb;
    int npad, batch;
    int64_t mstride, qstride;
    int32_t rc = BOBE_OK;

    double* at(double* base, int rb, int cb) const { return base + (int64_t)rb * NB * npad + (int64_t)cb * NB; }

    void gemm(const GemmArgs& a) {
        if (rc == BOBE_OK) rc = launch_gemm_nt(stream, a, batch);
    }

    void run(int b0, int b1) {
        if (rc != BOBE_OK) return;
        int nb = b1 - b0;
        if (nb <= 2) {  // leaf: one CTA per matrix does the whole 64- or 128-block in shared memory
            static const int panel4 = (int)env_int("BOBE_LEAF_PANEL4", 1);
            LeafIO io{fb.KB, fb.L, fb.Lt, fb.Linv, fb.U, fb.diag, fb.dstat, fb.gate, npad, b0 * NB, panel4, fb.gate, fb.gate, 3};
            if (nb == 1) {
                if ((rc = ensure_smem<leaf64_kernel>(LEAF64_SMEM)) != BOBE_OK) return;
                launch_pdl(leaf64_kernel, dim3(1, 1, batch), dim3(LEAF_THREADS), LEAF64_SMEM, stream, io);
            } else {
                if ((rc = ensure_smem<leaf128_kernel>(LEAF128_SMEM)) != BOBE_OK) return;
                launch_pdl(leaf128_kernel, dim3(1, 1, batch), dim3(LEAF_THREADS), LEAF128_SMEM, stream, io);
            }
            rc = check_launch("leaf kernel");
            return;
        }
        int mid = b0 + (nb + 1) / 2;
        int m1 = (mid - b0) * NB, m2 = (b1 - mid) * NB;
        run(b0, mid);
        auto base = [&]() {
            GemmArgs g{};
            g.lda = g.ldb = g.ldc = g.ldct = g.ldd = npad;
            g.strideA = g.strideB = g.strideC = g.strideCt = g.strideD = mstride;
            g.alpha = 1.0;
            return g;
        };
        {   // L21 = A21 * Linv11^T, computed as its transpose L21^T = Linv11 * A21^T so that the triangular operand
            // is the row operand (fragment-level skipping of its zeros); both orientations are stored either way
            GemmArgs g = base();
            g.A = at(fb.Linv, b0, b0); g.Bt = at(fb.KB, mid, b0); g.C = at(fb.Lt, b0, mid); g.Ct = at(fb.L, mid, b0);
            g.M = m1; g.N = m2; g.K = m1; g.flags = GEMM_A_LOWER;
            gemm(g);
        }
        // Correction of the panel solve, only for ill-conditioned matrices (gate set by the leaves):
        // a product with the explicit inverse is not backward stable (error ~ cond(L11) eps); one step of
        //   R = A21 - L21 L11^T,   L21 += R Linv11^T
        // restores an O(eps) residual (measured: profiles/r01/accuracy_probe.txt).
        {
            GemmArgs g = base();
            g.A = at(fb.L, mid, b0); g.Bt = at(fb.L, b0, b0); g.C = fb.Q; g.D = at(fb.KB, mid, b0);
            g.ldc = m1; g.strideC = qstride; g.gate = fb.gate;
            g.M = m2; g.N = m1; g.K = m1; g.alpha = -1.0; g.flags = GEMM_B_LOWER;
            gemm(g);
            g = base();
            g.A = fb.Q; g.lda = m1; g.strideA = qstride; g.Bt = at(fb.Linv, b0, b0);
            g.C = at(fb.L, mid, b0); g.D = g.C; g.Ct = at(fb.Lt, b0, mid); g.gate = fb.gate;
            g.M = m2; g.N = m1; g.K = m1; g.flags = GEMM_B_LOWER;
            gemm(g);
        }
        {   // A22 -= L21 * L21^T   (lower tiles only)
            GemmArgs g = base();
            g.A = at(fb.L, mid, b0); g.Bt = at(fb.L, mid, b0); g.C = at(fb.KB, mid, mid); g.D = g.C;
            g.M = m2; g.N = m2; g.K = m1; g.alpha = -1.0; g.flags = GEMM_C_LOWER;
            gemm(g);
        }
        run(mid, b1);
        {   // Q = Linv22 * L21
            GemmArgs g = base();
            g.A = at(fb.Linv, mid, mid); g.Bt = at(fb.Lt, b0, mid); g.C = fb.Q;
            g.ldc = m1; g.strideC = qstride;
            g.M = m2; g.N = m1; g.K = m2; g.flags = GEMM_A_LOWER;
            gemm(g);
        }
        {   // Linv21 = -Q * Linv11, computed as U12 = Linv21^T = -U11 * Q^T (triangular operand as the row operand)
            GemmArgs g = base();
            g.A = at(fb.U, b0, b0); g.Bt = fb.Q; g.ldb = m1; g.strideB = qstride;
            g.C = at(fb.U, b0, mid); g.Ct = at(fb.Linv, mid, b0);
            g.M = m1; g.N = m2; g.K = m1; g.alpha = -1.0; g.flags = GEMM_A_UPPER;
            gemm(g);
        }
    }
};
}  // namespace

// zero the off-diagonal upper blocks of L/Linv and lower blocks of Lt/U (the GEMMs never write them).
// band > 0: only the blocks within `band` block columns of the diagonal.  A tile of a triangular operand is read up
// to the end of the 128-wide k-tile that holds its diagonal, i.e. at most 127 columns beyond it (two 64-blocks);
// blocks further out are never touched, so the internal (log-ML) path skips them.
__global__ void zero_other_triangle_kernel(double* L, double* Lt, double* Linv, double* U, int npad, int band) {
    const int64_t zoff = (int64_t)blockIdx.z * npad * npad;
    const int rb = blockIdx.y;
    int cb = blockIdx.x;
    if (band > 0) {  // grid.x = 2 * band: block columns rb - band .. rb - 1 and rb + 1 .. rb + band
        cb = cb < band ? rb - band + cb : rb + 1 + (cb - band);
        if (cb < 0 || cb >= npad / NB) return;
    }
    if (rb == cb) return;
    double* lo0 = (cb > rb) ? L : Lt;     // L, Linv: zero where col block > row block
    double* lo1 = (cb > rb) ? Linv : U;   // Lt, U: zero where col block < row block   (null: buffer not kept)
    for (int idx = threadIdx.x; idx < NB * NB; idx += blockDim.x) {
        int r = idx >> 6, c = idx & 63;
        int64_t off = zoff + (int64_t)(rb * NB + r) * npad + cb * NB + c;
        if (lo0) lo0[off] = 0.0;
        if (lo1) lo1[off] = 0.0;
    }
}

}  // namespace bobe
#include "factor_tiled.cuh"
namespace bobe {

int32_t factor_recursive(cudaStream_t stream, const FactorBuffers& fb, int npad, int batch) {
    if (npad % NB) {
        set_error("factor: npad=%d not a multiple of %d", npad, NB);
        return BOBE_E_ARG;
    }
    int nbk = npad / NB;
    init_stat_kernel<<<(batch + 127) / 128, 128, 0, stream>>>(fb.dstat, fb.gate, batch, fb.force_refine);
    if (int32_t rc = check_launch("init_stat_kernel")) return rc;
    if (nbk > 1) {
        const int band = fb.zero_band > 0 && 2 * fb.zero_band < nbk ? fb.zero_band : 0;
        zero_other_triangle_kernel<<<dim3(band ? 2 * band : nbk, nbk, batch), 256, 0, stream>>>(fb.L, fb.Lt, fb.Linv, fb.U,
                                                                                              npad, band);
        if (int32_t rc = check_launch("zero_other_triangle_kernel")) return rc;
    }
    Rec r{stream, fb, npad, batch, (int64_t)npad * npad, factor_q_elems(npad)};
    r.run(0, nbk);
    return r.rc;
}

// Which scheme factorises: BOBE_FACTOR = 1 (default) the tile-column scheme of factor_tiled.cuh, 0 the recursive one above
// (kept for A/B measurements; needs fb.Lt).  Look-ahead on a second stream is used for sub-batches of at most
// BOBE_LOOKAHEAD_MAX matrices: a larger sub-batch fills the machine by itself and the sub-batches overlap each other.
int32_t factor_any(cudaStream_t stream, StreamPool* pool, int lane, const FactorBuffers& fb, int npad, int batch) {
    static const int64_t scheme = env_int("BOBE_FACTOR", 1);
    if (scheme == 0) {
        if (!fb.Lt) {
            set_error("factor: the recursive scheme needs the Lt buffer");
            return BOBE_E_ARG;
        }
        return factor_recursive(stream, fb, npad, batch);
    }
    return factor_tiled(factor_exec(stream, pool, lane, batch, batch), fb, npad, batch);
}

int32_t factor_on_pool(cudaStream_t stream, StreamPool* pool, const FactorBuffers& fb, int npad, int batch) {
    return factor_any(stream, pool, 0, fb, npad, batch);  // (the tiled scheme forks / joins its own streams)
}

FactorExec factor_exec(cudaStream_t stream, StreamPool* pool, int lane, int batch, int total_batch) {
    static const int64_t la_max = env_int("BOBE_LOOKAHEAD_MAX", 8);
    static const int64_t pw = env_int("BOBE_FACTOR_PW", 4);  // tile columns per outer panel (the same for every batch size)
    static const int64_t green_max = env_int("BOBE_GREEN_MAX", 16);  // most matrices in flight for the chain partition
    const bool la = pool && batch <= la_max && lane >= 0 && POOL_LANE_STREAMS * lane + 3 < POOL_STREAMS;
    if (!la) return FactorExec{stream, stream, nullptr, nullptr, nullptr, pool, lane < 0 ? 0 : lane, (int)pw};
    if (pool->green && total_batch <= green_max && total_batch <= pool->green_sms) {
        cudaStream_t* g = pool->gstreams + POOL_LANE_STREAMS * lane;
        return FactorExec{stream, g[0], g[1], g[2], g[3], pool, lane, (int)pw};
    }
    cudaStream_t* ps = pool->streams + POOL_LANE_STREAMS * lane;
    return FactorExec{stream, stream, ps[1], ps[2], ps[3], pool, lane, (int)pw};
}

int32_t launch_kinv(cudaStream_t stream, const FactorBuffers& fb, int npad, int batch) {
    GemmArgs g{};  // only the lower 128-tiles (diagonal tiles in full) are written: all the gradient kernel reads
    g.A = fb.U; g.Bt = fb.U; g.C = fb.KB; g.Ct = nullptr;
    g.lda = g.ldb = g.ldc = g.ldct = npad;
    g.strideA = g.strideB = g.strideC = g.strideCt = (int64_t)npad * npad;
    static const int64_t scheme = env_int("BOBE_FACTOR", 1);
    const int nl = (scheme != 0 && fb.n_live > 0 && fb.n_live <= npad) ? fb.n_live : npad;  // rows >= n_live of K^-1 are never read
    g.M = g.N = g.K = nl; g.alpha = 1.0;
    g.flags = GEMM_A_UPPER | GEMM_B_UPPER | GEMM_C_LOWER;
    return launch_gemm_nt(stream, g, batch);
}

// ---- vectors: alpha = U (Linv y) with one gated step of iterative refinement, logdet, quad, info -----------
// one warp per row; fixed summation order (deterministic)
__global__ void __launch_bounds__(256) matvec_tri_kernel(const double* __restrict__ Mtx, const double* __restrict__ x,
                                                         int64_t xstride, double* __restrict__ out, int64_t ostride,
                                                         int npad, int upper, int accumulate,
                                                         const int* __restrict__ gate) {
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    const int64_t z = blockIdx.y;
    if (row >= npad || (gate && gate[z] == 0)) return;
    const double* m = Mtx + z * (int64_t)npad * npad + (int64_t)row * npad;
    const double* xv = x + z * xstride;
    int kb = upper ? row : 0, ke = upper ? npad : row + 1;
    double s = 0.0;
    for (int k = kb + lane; k < ke; k += 32) s = fma(m[k], xv[k], s);
    s = warp_sum(s);
    if (lane == 0) {
        if (accumulate) s += out[z * ostride + row];
        out[z * ostride + row] = s;
    }
}

__global__ void __launch_bounds__(256) pad_y_kernel(const double* __restrict__ y, int64_t n, int npad, double* ypad) {
    int i = blockIdx.x * 256 + threadIdx.x;
    if (i < npad) ypad[i] = i < n ? y[i] : 0.0;
}

// r = y - (K0 alpha) - noise * alpha   (K0 alpha comes from the kernel-matrix pass; rows >= n stay 0)
__global__ void __launch_bounds__(256) residual_kernel(const double* __restrict__ ypad, const double* __restrict__ k0a,
                                                       const double* __restrict__ alpha, double noise, int64_t n,
                                                       int npad, double* __restrict__ r, const int* __restrict__ gate) {
    int i = blockIdx.x * 256 + threadIdx.x;
    const int64_t z = blockIdx.y;
    if (i >= npad || (gate && gate[z] == 0)) return;
    int64_t o = z * npad + i;
    r[o] = i < n ? (ypad[i] - k0a[o]) - noise * alpha[o] : 0.0;
}

// out[z][i] = sum_j Ksym[i][j] x[z][j] for a symmetric matrix of which only the LOWER triangle is stored (one warp per
// row; the j > i part walks down column i).  Only runs for gated (ill-conditioned) matrices.
__global__ void __launch_bounds__(256) symv_lower_kernel(const double* __restrict__ K, int64_t ldk, int64_t kstride, int n,
                                                         const double* __restrict__ x, int64_t xstride,
                                                         double* __restrict__ out, int64_t ostride,
                                                         const int* __restrict__ gate) {
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    const int64_t z = blockIdx.y;
    if (row >= n || (gate && gate[z] == 0)) return;
    const double* Kz = K + z * kstride;
    const double* xv = x + z * xstride;
    double s = 0.0;
    for (int j = lane; j <= row; j += 32) s = fma(Kz[(int64_t)row * ldk + j], xv[j], s);
    for (int j = row + 1 + lane; j < n; j += 32) s = fma(Kz[(int64_t)j * ldk + row], xv[j], s);
    s = warp_sum(s);
    if (lane == 0) out[z * ostride + row] = s;
}

// KB[z] = [K[z] lower 0; 0 I] padded to npad (the factorisation reads the lower triangle only)
__global__ void __launch_bounds__(256) pad_k_kernel(const double* __restrict__ K, int64_t ldk, int64_t kstride, int n,
                                                    int npad, double* __restrict__ KB) {
    const int64_t z = blockIdx.z;
    const int i = blockIdx.y, j = blockIdx.x * 256 + threadIdx.x;
    if (j >= npad) return;
    double v = (i == j) ? 1.0 : 0.0;
    if (i < n && j <= i) v = K[z * kstride + (int64_t)i * ldk + j];
    KB[z * (int64_t)npad * npad + (int64_t)i * npad + j] = v;
}
int32_t launch_pad_k(cudaStream_t stream, const double* K, int64_t ldk, int64_t kstride, int n, int npad, int batch, double* KB) {
    pad_k_kernel<<<dim3((npad + 255) / 256, npad, batch), 256, 0, stream>>>(K, ldk, kstride, n, npad, KB);
    return check_launch("pad_k_kernel");
}

__global__ void __launch_bounds__(256) factor_scalars_kernel(const double* __restrict__ diag,
                                                             const double* __restrict__ zv, int npad,
                                                             double* logdet, double* quad, int32_t* info) {
    __shared__ double s1[8], s2[8];
    __shared__ int sbad[8];
    const int64_t z = blockIdx.x;
    double a = 0.0, b = 0.0;
    int bad = 0;
    for (int i = threadIdx.x; i < npad; i += 256) {
        double d = diag[z * npad + i];
        if (!(d > 0.0) || isinf(d)) bad = 1;
        a += log(d);
        double zz = zv[z * npad + i];
        b = fma(zz, zz, b);
    }
    a = warp_sum(a);
    b = warp_sum(b);
    bad = __any_sync(0xffffffffu, bad);
    if ((threadIdx.x & 31) == 0) {
        s1[threadIdx.x >> 5] = a;
        s2[threadIdx.x >> 5] = b;
        sbad[threadIdx.x >> 5] = bad;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double ta = 0.0, tb = 0.0;
        int tbad = 0;
        for (int w = 0; w < 8; ++w) {
            ta += s1[w];
            tb += s2[w];
            tbad |= sbad[w];
        }
        if (tbad) ta = tb = nan("");
        if (logdet) logdet[z] = ta;
        if (quad) quad[z] = tb;
        if (info) info[z] = tbad;
    }
}

// scratch: padded y (npad) + three (batch, npad) vectors
int64_t solve_ws_doubles(int64_t npad, int64_t batch) { return (3 * batch + 1) * npad; }

// alpha must be (batch, npad).  For ill-conditioned matrices (gate set during the factorisation) alpha gets
// one step of iterative refinement against a freshly rebuilt K:  alpha += K^-1 (y - K alpha).
int32_t launch_solve_vectors(cudaStream_t stream, const FactorBuffers& fb, const SolveArgs& sa, const double* y,
                             int64_t n, int npad, int batch, double* z_ws, double* alpha, double* logdet,
                             double* quad, int32_t* info) {
    double* ypad = z_ws;                               // (npad) shared by all batch entries
    double* zv = z_ws + npad;                          // (batch, npad)
    double* k0a = zv + (int64_t)batch * npad;          // (batch, npad)
    double* rv = k0a + (int64_t)batch * npad;          // (batch, npad)
    pad_y_kernel<<<(npad + 255) / 256, 256, 0, stream>>>(y, n, npad, ypad);
    dim3 grid((npad + 7) / 8, batch);
    matvec_tri_kernel<<<grid, 256, 0, stream>>>(fb.Linv, ypad, 0, zv, npad, npad, 0, 0, nullptr);
    matvec_tri_kernel<<<grid, 256, 0, stream>>>(fb.U, zv, npad, alpha, npad, npad, 1, 0, nullptr);
    if (int32_t rc = check_launch("solve_vectors")) return rc;
    if (sa.Kin) {  // caller-supplied K: residual against K itself (its diagonal already holds the noise)
        symv_lower_kernel<<<grid, 256, 0, stream>>>(sa.Kin, sa.ldk, sa.kstride, (int)n, alpha, npad, k0a, npad, fb.gate);
        residual_kernel<<<dim3((npad + 255) / 256, batch), 256, 0, stream>>>(ypad, k0a, alpha, 0.0, n, npad, rv, fb.gate);
        matvec_tri_kernel<<<grid, 256, 0, stream>>>(fb.Linv, rv, npad, k0a, npad, npad, 0, 0, fb.gate);
        matvec_tri_kernel<<<grid, 256, 0, stream>>>(fb.U, k0a, npad, alpha, npad, npad, 1, 1, fb.gate);
    } else {
        KmatArgs ka{};
        ka.xa = sa.X; ka.xb = sa.X; ka.ls = sa.ls; ka.kv_ptr = sa.kv; ka.alpha = alpha; ka.mean_out = k0a;
        ka.xbs = sa.xs; ka.xbs_ld = npad; ka.xbs_stride = sa.d * (int64_t)npad;
        ka.n1 = n; ka.n2 = n; ka.d = sa.d; ka.rows_pad = npad; ka.cols_pad = npad;
        ka.ls_stride = sa.d; ka.alpha_stride = npad; ka.mean_stride = npad; ka.mean_standardised = 1;
        ka.gate = fb.gate;
        if (int32_t rc = launch_kmat(stream, sa.kind, ka, batch)) return rc;
        residual_kernel<<<dim3((npad + 255) / 256, batch), 256, 0, stream>>>(ypad, k0a, alpha, sa.noise, n, npad, rv,
                                                                           fb.gate);
        // z of the first solve is kept: quad = z^T z is a sum of squares (no cancellation), unlike y^T alpha
        matvec_tri_kernel<<<grid, 256, 0, stream>>>(fb.Linv, rv, npad, k0a, npad, npad, 0, 0, fb.gate);
        matvec_tri_kernel<<<grid, 256, 0, stream>>>(fb.U, k0a, npad, alpha, npad, npad, 1, 1, fb.gate);
    }
    factor_scalars_kernel<<<batch, 256, 0, stream>>>(fb.diag, zv, npad, logdet, quad, info);
    return check_launch("solve_vectors");
}

// ---- rank-b append: GP.update without the O(n^3) rebuild -------------------------------------
// With the hyper-parameters unchanged (BOBE/gp.py:541 re-factors the whole matrix), appending point m to a factor of
// the first m points is exactly one more step of the recursion above with a 1-row second block:
//   v = Linv k  (+ the panel correction  v += Linv (k - L v));   delta^2 = k** - v.v;
//   L[m] = [v^T, delta];   Linv[m] = [-(v^T Linv) / delta, 1 / delta]
// i.e. O(m^2) per point; alpha is then re-solved for the (re-standardised) targets of all points, with the same
// refinement step the batched factorisation applies to ill-conditioned matrices.
__global__ void __launch_bounds__(256) vec_sub_kernel(const double* a, const double* b, double* out, int n) {
    int i = blockIdx.x * 256 + threadIdx.x;
    if (i < n) out[i] = a[i] - b[i];
}

// delta from v, row m of L;  scal[0] = delta, scal[1] = 1 / delta;  *info |= 1 if the pivot is not positive
__global__ void __launch_bounds__(256) append_row_L_kernel(const double* __restrict__ v, int m, int npad, double kk,
                                                           double* __restrict__ L, double* __restrict__ scal,
                                                           int32_t* __restrict__ info) {
    __shared__ double red[8];
    double s = 0.0;
    for (int i = threadIdx.x; i < m; i += 256) {
        double vi = v[i];
        s = fma(vi, vi, s);
        L[(int64_t)m * npad + i] = vi;
    }
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < 8; ++w) t += red[w];
        const double d2 = kk - t;
        const double dl = sqrt(d2);  // NaN for a negative pivot, like the reference's jnp.sqrt (BOBE/gp.py:187)
        if (!(d2 > 0.0) || isinf(d2)) *info = 1;
        L[(int64_t)m * npad + m] = dl;
        scal[0] = dl;
        scal[1] = 1.0 / dl;
    }
}

// out[j] = scale * sum_{i=j}^{m-1} x[i] Mtx[i][j]   (x^T times a lower-triangular matrix; one thread per column,
// consecutive threads read consecutive addresses of each row).  scale_ptr (device) overrides scale when given.
__global__ void __launch_bounds__(128) vecmat_lower_kernel(const double* __restrict__ Mtx, const double* __restrict__ x,
                                                           int m, int npad, double scale,
                                                           const double* __restrict__ scale_ptr, int accumulate,
                                                           double* __restrict__ out) {
    const int j = blockIdx.x * 128 + threadIdx.x;
    if (j >= m) return;
    double s0 = 0.0, s1 = 0.0;
    int i = j;
    for (; i + 1 < m; i += 2) {
        s0 = fma(x[i], Mtx[(int64_t)i * npad + j], s0);
        s1 = fma(x[i + 1], Mtx[(int64_t)(i + 1) * npad + j], s1);
    }
    if (i < m) s0 = fma(x[i], Mtx[(int64_t)i * npad + j], s0);
    const double sc = scale_ptr ? -scale_ptr[1] : scale;
    double r = sc * (s0 + s1);
    if (accumulate) r += out[j];
    out[j] = r;
}

__global__ void append_diag_Linv_kernel(double* Linv, int m, int npad, const double* scal) {
    Linv[(int64_t)m * npad + m] = scal[1];
}

int64_t append_ws_doubles(int64_t npad) { return 8 * npad + 64 * npad + 64; }

int32_t factor_append(cudaStream_t stream, int kind, const double* X, const double* y, int64_t n_old, int64_t b, int64_t d,
                      const double* ls, double kv, double noise, double* L, double* Linv, double* alpha, int32_t* info,
                      double* ws) {
    const int npad = (int)npad_of(n_old + b);
    double* kbuf = ws;                    // (64, npad): row 0 is k(x_m, X); the tile kernel needs 64 padded rows
    double* v = kbuf + 64 * (int64_t)npad;
    double* tmp = v + npad;
    double* r = tmp + npad;
    double* ypad = r + npad;
    double* z = ypad + npad;
    double* k0a = z + npad;
    double* scal = k0a + npad;            // delta, 1/delta
    const double kk = kv + noise;
    if (cudaMemsetAsync(info, 0, sizeof(int32_t), stream) != cudaSuccess) {
        set_error("factor_append: memset failed");
        return BOBE_E_CUDA;
    }
    const dim3 mv_grid((npad + 7) / 8, 1);
    for (int64_t p = 0; p < b; ++p) {
        const int m = (int)(n_old + p);
        if (cudaMemsetAsync(kbuf, 0, (size_t)npad * 8, stream) != cudaSuccess) {
            set_error("factor_append: memset failed");
            return BOBE_E_CUDA;
        }
        KmatArgs ka{};
        ka.xa = X + (int64_t)m * d; ka.xb = X; ka.ls = ls; ka.kv = kv; ka.noise = noise; ka.out = kbuf;
        ka.n1 = 1; ka.n2 = m; ka.d = d; ka.ldo = npad; ka.rows_pad = 64; ka.cols_pad = round_up(m, 64);
        ka.store_rows = 1; ka.store_cols = ka.cols_pad; ka.vec_ok = 1;
        if (m > 0)
            if (int32_t rc = launch_kmat(stream, kind, ka, 1)) return rc;
        // v = Linv k;  tmp = L v;  r = k - tmp;  v += Linv r     (rows >= m of the padded factors are identity rows and
        // k is zero there, so the padded part of every vector stays zero)
        matvec_tri_kernel<<<mv_grid, 256, 0, stream>>>(Linv, kbuf, 0, v, 0, npad, 0, 0, nullptr);
        matvec_tri_kernel<<<mv_grid, 256, 0, stream>>>(L, v, 0, tmp, 0, npad, 0, 0, nullptr);
        vec_sub_kernel<<<(npad + 255) / 256, 256, 0, stream>>>(kbuf, tmp, r, npad);
        matvec_tri_kernel<<<mv_grid, 256, 0, stream>>>(Linv, r, 0, v, 0, npad, 0, 1, nullptr);
        append_row_L_kernel<<<1, 256, 0, stream>>>(v, m, npad, kk, L, scal, info);
        if (m > 0) vecmat_lower_kernel<<<(m + 127) / 128, 128, 0, stream>>>(Linv, v, m, npad, 0.0, scal, 0, Linv + (int64_t)m * npad);
        append_diag_Linv_kernel<<<1, 1, 0, stream>>>(Linv, m, npad, scal);
        if (int32_t rc = check_launch("factor_append row")) return rc;
    }
    // alpha = Linv^T (Linv y), then one refinement step against the rebuilt K:  alpha += K^-1 (y - K alpha)
    const int n = (int)(n_old + b);
    pad_y_kernel<<<(npad + 255) / 256, 256, 0, stream>>>(y, n, npad, ypad);
    matvec_tri_kernel<<<mv_grid, 256, 0, stream>>>(Linv, ypad, 0, z, 0, npad, 0, 0, nullptr);
    if (cudaMemsetAsync(alpha, 0, (size_t)npad * 8, stream) != cudaSuccess) {
        set_error("factor_append: memset failed");
        return BOBE_E_CUDA;
    }
    vecmat_lower_kernel<<<(n + 127) / 128, 128, 0, stream>>>(Linv, z, n, npad, 1.0, nullptr, 0, alpha);
    {
        KmatArgs ka{};
        ka.xa = X; ka.xb = X; ka.ls = ls; ka.kv = kv; ka.alpha = alpha; ka.mean_out = k0a;
        ka.n1 = n; ka.n2 = n; ka.d = d; ka.rows_pad = npad; ka.cols_pad = npad; ka.mean_standardised = 1;
        if (cudaMemsetAsync(k0a, 0, (size_t)npad * 8, stream) != cudaSuccess) {
            set_error("factor_append: memset failed");
            return BOBE_E_CUDA;
        }
        if (int32_t rc = launch_kmat(stream, kind, ka, 1)) return rc;
        residual_kernel<<<dim3((npad + 255) / 256, 1), 256, 0, stream>>>(ypad, k0a, alpha, noise, n, npad, r, nullptr);
        matvec_tri_kernel<<<mv_grid, 256, 0, stream>>>(Linv, r, 0, z, 0, npad, 0, 0, nullptr);
        vecmat_lower_kernel<<<(n + 127) / 128, 128, 0, stream>>>(Linv, z, n, npad, 1.0, nullptr, 1, alpha);
    }
    return check_launch("factor_append alpha");
}

}  // namespace bobe
