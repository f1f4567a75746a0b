// FP64 tensor-core (DMMA.8x8x4) "NT" GEMM building block:  C[i][j] (+)= alpha * sum_k A[i][k] * Bt[j][k]
// Both operands are row-major with the contraction index contiguous, which is the natural row.col layout
// of mma.sync.m8n8k4.f64.  All products on the hot path are brought into this form by keeping every
// triangular matrix in both orientations ("NT-only formulation").
//
// Shared-memory layout per stage and operand: two "k8 panels", each `rows x 8` doubles dense (64-byte
// rows).  Lane (g = lane>>2, t = lane&3) reads 16 bytes at (row g, doubles 2t..2t+1) with ONE LDS.128 and
// uses .x for the MMA that contracts k = {0,2,4,6} of the panel and .y for the MMA that contracts
// {1,3,5,7}: the k labelling inside an MMA is free as long as A and B agree.  A quarter-warp (rows g, g+1)
// therefore reads 128 contiguous bytes: conflict-free without padding or swizzle.
#pragma once
#include "common.cuh"

namespace bobe {

template <int BM_, int BN_, int WM_, int WN_, int STAGES_>
struct TileCfg {
    static constexpr int BM = BM_, BN = BN_, WM = WM_, WN = WN_, STAGES = STAGES_;
    static constexpr int BK = 16;
    static constexpr int THREADS = 32 * WM * WN;
    static constexpr int WTM = BM / WM, WTN = BN / WN;  // warp tile
    static constexpr int MF = WTM / 8, NF = WTN / 8;    // 8x8 fragments per warp tile
    static constexpr int STAGE_DOUBLES = (BM + BN) * BK;
    static constexpr int SMEM_BYTES = STAGES * STAGE_DOUBLES * 8;
    static_assert(WTM % 8 == 0 && WTN % 8 == 0, "warp tile must be a multiple of 8x8");
    static_assert((BM * 8) % THREADS == 0 && (BN * 8) % THREADS == 0, "loader mapping");
};

// flags describing known-zero structure of the operands (only used to shorten the k loop; the zeros are
// physically present in memory, so no element masking is ever needed)
enum : int {
    GEMM_A_LOWER = 1,   // A[i][k] == 0 for k > i
    GEMM_A_UPPER = 2,   // A[i][k] == 0 for k < i
    GEMM_B_LOWER = 4,   // Bt[j][k] == 0 for k > j
    GEMM_B_UPPER = 8,   // Bt[j][k] == 0 for k < j
    GEMM_C_LOWER = 16,  // only tiles touching the lower triangle (j0 <= i0 + BM - 1) are computed
};

template <class Cfg>
struct Mainloop {
    // one operand tile of one stage: rows x 16 doubles, global row stride ld
    template <int ROWS>
    __device__ static __forceinline__ void load_tile(double* s, const double* g, int64_t ld, int rows_valid, int k0) {
        constexpr int CHUNKS = ROWS * 8;
#pragma unroll
        for (int it = 0; it < CHUNKS / Cfg::THREADS; ++it) {
            int id = threadIdx.x + it * Cfg::THREADS;
            int c4 = id & 3, row = ((id >> 4) << 1) | ((id >> 2) & 1), panel = (id >> 3) & 1;
            bool ok = row < rows_valid;
            const double* src = g + (int64_t)(ok ? row : 0) * ld + k0 + panel * 8 + c4 * 2;
            cp_async16(s + ((panel * ROWS + row) * 8 + c4 * 2), src, ok);
        }
    }

    __device__ static __forceinline__ void load_stage(double* smem, int stage, const double* A, int64_t lda,
                                                      int rowsA, const double* Bt, int64_t ldb, int rowsB, int k0) {
        double* sA = smem + stage * Cfg::STAGE_DOUBLES;
        double* sB = sA + Cfg::BM * Cfg::BK;
        load_tile<Cfg::BM>(sA, A, lda, rowsA, k0);
        load_tile<Cfg::BN>(sB, Bt, ldb, rowsB, k0);
    }

    // acc += A[0:BM, kb:ke] * Bt[0:BN, kb:ke]^T   (kb, ke multiples of 16; A/Bt point at the tile's first row)
    __device__ static __forceinline__ void run(double (&acc)[Cfg::MF][Cfg::NF][2], const double* A, int64_t lda,
                                               int rowsA, const double* Bt, int64_t ldb, int rowsB, int kb, int ke,
                                               double* smem) {
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        const int g = lane >> 2, t = lane & 3;
        const int wm = warp / Cfg::WN, wn = warp % Cfg::WN;
        const int ktiles = (ke - kb) / Cfg::BK;

#pragma unroll
        for (int s = 0; s < Cfg::STAGES - 1; ++s) {
            if (s < ktiles) load_stage(smem, s, A, lda, rowsA, Bt, ldb, rowsB, kb + s * Cfg::BK);
            cp_async_commit();
        }
        for (int kt = 0; kt < ktiles; ++kt) {
            cp_async_wait<Cfg::STAGES - 2>();
            __syncthreads();
            {
                int nk = kt + Cfg::STAGES - 1;
                if (nk < ktiles) load_stage(smem, nk % Cfg::STAGES, A, lda, rowsA, Bt, ldb, rowsB, kb + nk * Cfg::BK);
                cp_async_commit();
            }
            const double* sA = smem + (kt % Cfg::STAGES) * Cfg::STAGE_DOUBLES;
            const double* sB = sA + Cfg::BM * Cfg::BK;
#pragma unroll
            for (int p = 0; p < 2; ++p) {
                double2 a[Cfg::MF], b[Cfg::NF];
#pragma unroll
                for (int mf = 0; mf < Cfg::MF; ++mf)
                    a[mf] = *reinterpret_cast<const double2*>(sA + ((p * Cfg::BM + wm * Cfg::WTM + mf * 8 + g) * 8 + 2 * t));
#pragma unroll
                for (int nf = 0; nf < Cfg::NF; ++nf)
                    b[nf] = *reinterpret_cast<const double2*>(sB + ((p * Cfg::BN + wn * Cfg::WTN + nf * 8 + g) * 8 + 2 * t));
#pragma unroll
                for (int mf = 0; mf < Cfg::MF; ++mf)
#pragma unroll
                    for (int nf = 0; nf < Cfg::NF; ++nf) {
                        dmma884(acc[mf][nf][0], acc[mf][nf][1], a[mf].x, b[nf].x);
                        dmma884(acc[mf][nf][0], acc[mf][nf][1], a[mf].y, b[nf].y);
                    }
            }
        }
        cp_async_wait<0>();
        __syncthreads();  // smem may be reused by the caller (next tile / epilogue)
    }
};

// k-range of a tile from the structure flags
__device__ __forceinline__ void tile_k_range(int flags, int i0, int j0, int BM, int BN, int K, int& kb, int& ke) {
    kb = 0;
    ke = K;
    if (flags & GEMM_A_LOWER) ke = min(ke, i0 + BM);
    if (flags & GEMM_B_LOWER) ke = min(ke, j0 + BN);
    if (flags & GEMM_A_UPPER) kb = max(kb, i0);
    if (flags & GEMM_B_UPPER) kb = max(kb, j0);
    kb = (kb / 16) * 16;
    ke = min(K, ((ke + 15) / 16) * 16);
    if (ke < kb) ke = kb;
}

struct GemmArgs {
    const double* A;
    const double* Bt;
    double* C;
    double* Ct;       // optional transposed copy of the result (may be null)
    const double* D;  // optional addend: C = alpha*A*Bt^T + D  (may alias C; null -> 0)
    const int* gate;  // optional per-batch switch: the launch is a no-op for batch entries with gate[z] == 0
    int64_t lda, ldb, ldc, ldct, ldd;
    int64_t strideA, strideB, strideC, strideCt, strideD;  // batch strides (grid.z)
    int M, N, K;
    double alpha;
    int flags;
};

// Batched NT GEMM with optional dual (normal + transposed) store.  grid = (tiles_n, tiles_m, batch).
template <class Cfg>
__global__ void __launch_bounds__(Cfg::THREADS, 1) gemm_nt_kernel(GemmArgs p) {
    extern __shared__ __align__(16) double smem[];
    const int i0 = blockIdx.y * Cfg::BM, j0 = blockIdx.x * Cfg::BN;
    if ((p.flags & GEMM_C_LOWER) && j0 > i0 + Cfg::BM - 1) return;
    const int64_t z = blockIdx.z;
    if (p.gate && p.gate[z] == 0) return;
    const double* A = p.A + z * p.strideA + (int64_t)i0 * p.lda;
    const double* Bt = p.Bt + z * p.strideB + (int64_t)j0 * p.ldb;
    int kb, ke;
    tile_k_range(p.flags, i0, j0, Cfg::BM, Cfg::BN, p.K, kb, ke);

    double acc[Cfg::MF][Cfg::NF][2];
#pragma unroll
    for (int mf = 0; mf < Cfg::MF; ++mf)
#pragma unroll
        for (int nf = 0; nf < Cfg::NF; ++nf) acc[mf][nf][0] = acc[mf][nf][1] = 0.0;

    Mainloop<Cfg>::run(acc, A, p.lda, min(Cfg::BM, p.M - i0), Bt, p.ldb, min(Cfg::BN, p.N - j0), kb, ke, smem);

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int wm = warp / Cfg::WN, wn = warp % Cfg::WN;
    double* C = p.C + z * p.strideC;
    double* Ct = p.Ct ? p.Ct + z * p.strideCt : nullptr;
    const double* D = p.D ? p.D + z * p.strideD : nullptr;
#pragma unroll
    for (int mf = 0; mf < Cfg::MF; ++mf) {
        int row = i0 + wm * Cfg::WTM + mf * 8 + g;
        if (row >= p.M) continue;
#pragma unroll
        for (int nf = 0; nf < Cfg::NF; ++nf) {
            int col = j0 + wn * Cfg::WTN + nf * 8 + 2 * t;
            if (col >= p.N) continue;  // N is even, so col+1 < N too
            double v0 = p.alpha * acc[mf][nf][0], v1 = p.alpha * acc[mf][nf][1];
            double2* dst = reinterpret_cast<double2*>(C + (int64_t)row * p.ldc + col);
            if (D) {
                double2 old = *reinterpret_cast<const double2*>(D + (int64_t)row * p.ldd + col);
                v0 += old.x;
                v1 += old.y;
            }
            *dst = make_double2(v0, v1);
            if (Ct) {
                Ct[(int64_t)col * p.ldct + row] = v0;
                Ct[(int64_t)(col + 1) * p.ldct + row] = v1;
            }
        }
    }
}

using CfgBig = TileCfg<128, 128, 2, 4, 4>;   // 256 threads, warp tile 64x32, 128 KB smem
using CfgSmall = TileCfg<64, 64, 2, 2, 4>;   // 128 threads, warp tile 32x32, 64 KB smem

int32_t launch_gemm_nt(cudaStream_t stream, const GemmArgs& args, int batch);

}  // namespace bobe
