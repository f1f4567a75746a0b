// FP64 tensor-core (DMMA.8x8x4) "NT" GEMM building block:  C[i][j] (+)= alpha * sum_k A[i][k] * Bt[j][k]
// Both operands are row-major with the contraction index contiguous, which is the natural row.col layout
// of mma.sync.m8n8k4.f64.  All products on the hot path are brought into this form by keeping every
// triangular matrix in both orientations ("NT-only formulation").
//
// Shared-memory layout per stage and operand: BK/8 "k8 panels", each `rows x 8` doubles dense (64-byte
// rows).  Lane (g = lane>>2, t = lane&3) reads 16 bytes at (row g, doubles 2t..2t+1) with ONE LDS.128 and
// uses .x for the MMA that contracts k = {0,2,4,6} of the panel and .y for the MMA that contracts
// {1,3,5,7}: the k labelling inside an MMA is free as long as A and B agree.  A quarter-warp (rows g, g+1)
// therefore reads 128 contiguous bytes: conflict-free without padding or swizzle.
#pragma once
#include "common.cuh"

namespace bobe {

// ILV: warp wm owns the 8-row groups {wm, wm + WM, wm + 2 WM, ...} of the tile instead of WTM consecutive rows, so
// that a triangular operand leaves every warp the same number of live fragments (see Mainloop::run, tri0).
template <int BM_, int BN_, int WM_, int WN_, int STAGES_, int BK_ = 16, bool ILV_ = false, int MINB_ = 1>
struct TileCfg {
    static constexpr int BM = BM_, BN = BN_, WM = WM_, WN = WN_, STAGES = STAGES_;
    static constexpr int MINB = MINB_;  // CTAs per SM the register allocation must allow (__launch_bounds__)
    static constexpr bool ILV = ILV_;
    // first row (within the tile) of fragment mf of warp-row wm
    __host__ __device__ static constexpr int frag_row(int wm, int mf) {
        return ILV_ ? mf * 8 * WM_ + wm * 8 : wm * (BM_ / WM_) + mf * 8;
    }
    static constexpr int BK = BK_;
    static constexpr int PANELS = BK / 8;  // k8 panels per stage
    static constexpr int THREADS = 32 * WM * WN;
    static constexpr int WTM = BM / WM, WTN = BN / WN;  // warp tile
    static constexpr int MF = WTM / 8, NF = WTN / 8;    // 8x8 fragments per warp tile
    static constexpr int STAGE_DOUBLES = (BM + BN) * BK;
    static constexpr int SMEM_BYTES = STAGES * STAGE_DOUBLES * 8;
    static_assert(WTM % 8 == 0 && WTN % 8 == 0, "warp tile must be a multiple of 8x8");
    static_assert(BK % 16 == 0, "BK must be a multiple of 16 (k ranges are 16-aligned)");
    static_assert((BM * BK / 2) % THREADS == 0 && (BN * BK / 2) % THREADS == 0, "loader mapping");
};

// flags describing known-zero structure of the operands (only used to shorten the k loop; the zeros are
// physically present in memory, so no element masking is ever needed)
enum : int {
    GEMM_A_LOWER = 1,   // A[i][k] == 0 for k > i
    GEMM_A_UPPER = 2,   // A[i][k] == 0 for k < i
    GEMM_B_LOWER = 4,   // Bt[j][k] == 0 for k > j
    GEMM_B_UPPER = 8,   // Bt[j][k] == 0 for k < j
    GEMM_C_LOWER = 16,  // only tiles touching the lower triangle (j0 <= i0 + BM - 1) are computed
    // launch-order hint (set by the launcher for small batches): grid = (batch, tiles_n, tiles_m) instead of
    // (tiles_n, tiles_m, batch), i.e. the batch index varies fastest and the tile rows come heaviest first (reversed for a
    // lower-triangular A, whose last rows have the longest k range) -- all matrices start with their long tiles and the
    // last wave consists of short ones.  With the default order the heavy tiles of the LAST matrix start at 7/8 of the run.
    GEMM_ORDER_BATCH_FIRST = 32,
};

enum : int { TRI_NONE = 0, TRI_LOWER = 1, TRI_UPPER = 2 };  // what Mainloop::run knows about A inside a tile

template <class Cfg>
struct Mainloop {
    static constexpr int ROWS_PER_SLOT = Cfg::THREADS / (8 * Cfg::PANELS) * 2;  // rows between a thread's chunks
    static constexpr int SLOTS_A = Cfg::BM / ROWS_PER_SLOT, SLOTS_B = Cfg::BN / ROWS_PER_SLOT;

    // Per-thread addressing of the global->shared copies, computed ONCE per tile (the k loop only adds BK).
    // Thread mapping: 8 consecutive threads fill two adjacent 64-byte panel rows (128 contiguous smem bytes),
    // the next 8 the same rows of the next k8 panel; slot s of a thread is ROWS_PER_SLOT rows further down.
    struct Loader {
        const double* gA;   // this thread's first 16-byte chunk of A at k = kb
        const double* gB;
        int64_t stepA, stepB;  // ROWS_PER_SLOT * ld
        uint32_t sA, sB;       // shared-memory byte addresses of the chunk in stage 0
        uint32_t okA, okB;     // bit s set: slot s is a valid row (others are zero-filled)

        __device__ __forceinline__ void init(const double* A, int64_t lda, int rowsA, const double* Bt, int64_t ldb,
                                             int rowsB, int kb, double* smem) {
            const int tid = threadIdx.x;
            const int c4 = tid & 3, panel = (tid >> 3) % Cfg::PANELS;
            const int row = ((tid / (8 * Cfg::PANELS)) << 1) | ((tid >> 2) & 1);
            gA = A + (int64_t)row * lda + kb + panel * 8 + c4 * 2;
            gB = Bt + (int64_t)row * ldb + kb + panel * 8 + c4 * 2;
            stepA = (int64_t)ROWS_PER_SLOT * lda;
            stepB = (int64_t)ROWS_PER_SLOT * ldb;
            uint32_t base = static_cast<uint32_t>(__cvta_generic_to_shared(smem));
            sA = base + ((panel * Cfg::BM + row) * 8 + c4 * 2) * 8;
            sB = base + (Cfg::BM * Cfg::BK + (panel * Cfg::BN + row) * 8 + c4 * 2) * 8;
            okA = okB = 0;
#pragma unroll
            for (int sl = 0; sl < SLOTS_A; ++sl) okA |= (row + sl * ROWS_PER_SLOT < rowsA) ? (1u << sl) : 0u;
#pragma unroll
            for (int sl = 0; sl < SLOTS_B; ++sl) okB |= (row + sl * ROWS_PER_SLOT < rowsB) ? (1u << sl) : 0u;
        }
        // copy k-tile `kt` (relative to kb) into pipeline stage `stage`
        __device__ __forceinline__ void issue(int kt, int stage) const {
            const uint32_t so = stage * (Cfg::STAGE_DOUBLES * 8);
            const double* a = gA + kt * Cfg::BK;
            const double* b = gB + kt * Cfg::BK;
#pragma unroll
            for (int sl = 0; sl < SLOTS_A; ++sl) {
                bool ok = (okA >> sl) & 1u;
                const double* src = ok ? a + sl * stepA : gA;
                int sz = ok ? 16 : 0;
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(sA + so + sl * (ROWS_PER_SLOT * 64)),
                             "l"(src), "r"(sz));
            }
#pragma unroll
            for (int sl = 0; sl < SLOTS_B; ++sl) {
                bool ok = (okB >> sl) & 1u;
                const double* src = ok ? b + sl * stepB : gB;
                int sz = ok ? 16 : 0;
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(sB + so + sl * (ROWS_PER_SLOT * 64)),
                             "l"(src), "r"(sz));
            }
        }
    };

    __device__ static __forceinline__ void mma_panel(double (&acc)[Cfg::MF][Cfg::NF][2], const double* sA,
                                                     const double* sB, int p, int wm, int wn, int g, int t) {
        double2 a[Cfg::MF], b[Cfg::NF];
#pragma unroll
        for (int mf = 0; mf < Cfg::MF; ++mf)
            a[mf] = *reinterpret_cast<const double2*>(sA + ((p * Cfg::BM + Cfg::frag_row(wm, mf) + g) * 8 + 2 * t));
#pragma unroll
        for (int nf = 0; nf < Cfg::NF; ++nf)
            b[nf] = *reinterpret_cast<const double2*>(sB + ((p * Cfg::BN + wn * Cfg::WTN + nf * 8 + g) * 8 + 2 * t));
        // two passes (k = {0,2,4,6} then {1,3,5,7}): MF*NF independent MMAs between the two that update the
        // same accumulator, so no MMA ever waits on the latency of the previous one
#pragma unroll
        for (int mf = 0; mf < Cfg::MF; ++mf)
#pragma unroll
            for (int nf = 0; nf < Cfg::NF; ++nf) dmma884(acc[mf][nf][0], acc[mf][nf][1], a[mf].x, b[nf].x);
#pragma unroll
        for (int mf = 0; mf < Cfg::MF; ++mf)
#pragma unroll
            for (int nf = 0; nf < Cfg::NF; ++nf) dmma884(acc[mf][nf][0], acc[mf][nf][1], a[mf].y, b[nf].y);
    }

    // same, restricted to the row fragments lo <= mf < hi (warp-uniform bounds): the others are known to
    // multiply zeros (above the diagonal of a lower-triangular A, or rows beyond the matrix)
    __device__ static __forceinline__ void mma_panel_range(double (&acc)[Cfg::MF][Cfg::NF][2], const double* sA,
                                                           const double* sB, int p, int wm, int wn, int g, int t,
                                                           int lo, int hi) {
        double2 a[Cfg::MF], b[Cfg::NF];
#pragma unroll
        for (int mf = 0; mf < Cfg::MF; ++mf)
            a[mf] = *reinterpret_cast<const double2*>(sA + ((p * Cfg::BM + Cfg::frag_row(wm, mf) + g) * 8 + 2 * t));
#pragma unroll
        for (int nf = 0; nf < Cfg::NF; ++nf)
            b[nf] = *reinterpret_cast<const double2*>(sB + ((p * Cfg::BN + wn * Cfg::WTN + nf * 8 + g) * 8 + 2 * t));
#pragma unroll
        for (int mf = 0; mf < Cfg::MF; ++mf)
            if (mf >= lo && mf < hi) {
#pragma unroll
                for (int nf = 0; nf < Cfg::NF; ++nf) dmma884(acc[mf][nf][0], acc[mf][nf][1], a[mf].x, b[nf].x);
            }
#pragma unroll
        for (int mf = 0; mf < Cfg::MF; ++mf)
            if (mf >= lo && mf < hi) {
#pragma unroll
                for (int nf = 0; nf < Cfg::NF; ++nf) dmma884(acc[mf][nf][0], acc[mf][nf][1], a[mf].y, b[nf].y);
            }
    }

    // same, for the fragments mf >= lo (FROM) or mf < hi (UPTO) where the bound is uniform over the CTA: a jump into
    // straight-line, unpredicated MMA code (a warp-dependent bound would cost a predicate and a WARPSYNC per fragment)
    template <bool UPTO>
    __device__ static __forceinline__ void mma_panel_jump(double (&acc)[Cfg::MF][Cfg::NF][2], const double* sA,
                                                          const double* sB, int p, int wm, int wn, int g, int t, int bound) {
        static_assert(Cfg::MF <= 8, "fall-through tables below cover 8 fragments");
        double2 a[Cfg::MF], b[Cfg::NF];
#pragma unroll
        for (int mf = 0; mf < Cfg::MF; ++mf)
            a[mf] = *reinterpret_cast<const double2*>(sA + ((p * Cfg::BM + Cfg::frag_row(wm, mf) + g) * 8 + 2 * t));
#pragma unroll
        for (int nf = 0; nf < Cfg::NF; ++nf)
            b[nf] = *reinterpret_cast<const double2*>(sB + ((p * Cfg::BN + wn * Cfg::WTN + nf * 8 + g) * 8 + 2 * t));
#define BOBE_FRAG(MFI, XY)                                                                                   \
    if (MFI < Cfg::MF) {                                                                                     \
        _Pragma("unroll") for (int nf = 0; nf < Cfg::NF; ++nf)                                               \
            dmma884(acc[MFI < Cfg::MF ? MFI : 0][nf][0], acc[MFI < Cfg::MF ? MFI : 0][nf][1],                \
                    a[MFI < Cfg::MF ? MFI : 0].XY, b[nf].XY);                                                \
    }
#define BOBE_PASS_FROM(XY)                             \
    switch (bound) {                                   \
        case 0: BOBE_FRAG(0, XY) [[fallthrough]];      \
        case 1: BOBE_FRAG(1, XY) [[fallthrough]];      \
        case 2: BOBE_FRAG(2, XY) [[fallthrough]];      \
        case 3: BOBE_FRAG(3, XY) [[fallthrough]];      \
        case 4: BOBE_FRAG(4, XY) [[fallthrough]];      \
        case 5: BOBE_FRAG(5, XY) [[fallthrough]];      \
        case 6: BOBE_FRAG(6, XY) [[fallthrough]];      \
        case 7: BOBE_FRAG(7, XY) break;                \
        default: break;                                \
    }
#define BOBE_PASS_UPTO(XY)                             \
    switch (bound) {                                   \
        case 8: BOBE_FRAG(7, XY) [[fallthrough]];      \
        case 7: BOBE_FRAG(6, XY) [[fallthrough]];      \
        case 6: BOBE_FRAG(5, XY) [[fallthrough]];      \
        case 5: BOBE_FRAG(4, XY) [[fallthrough]];      \
        case 4: BOBE_FRAG(3, XY) [[fallthrough]];      \
        case 3: BOBE_FRAG(2, XY) [[fallthrough]];      \
        case 2: BOBE_FRAG(1, XY) [[fallthrough]];      \
        case 1: BOBE_FRAG(0, XY) break;                \
        default: break;                                \
    }
        if (UPTO) {
            BOBE_PASS_UPTO(x)
            BOBE_PASS_UPTO(y)
        } else {
            BOBE_PASS_FROM(x)
            BOBE_PASS_FROM(y)
        }
#undef BOBE_PASS_FROM
#undef BOBE_PASS_UPTO
#undef BOBE_FRAG
    }

    // acc += A[0:BM, kb:ke] * Bt[0:BN, kb:ke]^T   (kb, ke multiples of BK; A/Bt point at the tile's first row)
    // MODE tells what is known about A beyond the tile-level k range (tri0 = the tile's first row in operand
    // coordinates):
    //   TRI_LOWER  A[r][k] == 0 for k > tri0 + r: the LAST k-tiles reach above the diagonal, fragment mf of warp row
    //              wm is dead from panel kp > tri0 + frag_row(wm, mf) + 7 on;
    //   TRI_UPPER  A[r][k] == 0 for k < tri0 + r: the FIRST k-tiles start below the diagonal, the fragment is dead
    //              while kp + 7 < tri0 + frag_row(wm, mf).
    // Those k-tiles only issue the MMAs of the fragments that can be non-zero; all others run the branch-free body.
    // rows_live < BM (TRI_LOWER only): rows beyond it contribute nothing at all (partial first block of trmm_sumsq).
    template <int MODE = TRI_NONE>
    __device__ static __forceinline__ void run(double (&acc)[Cfg::MF][Cfg::NF][2], const double* A, int64_t lda,
                                               int rowsA, const double* Bt, int64_t ldb, int rowsB, int kb, int ke,
                                               double* smem, int tri0 = 0, int rows_live = Cfg::BM) {
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        const int g = lane >> 2, t = lane & 3;
        const int wm = warp / Cfg::WN, wn = warp % Cfg::WN;
        const int ktiles = (ke - kb) / Cfg::BK;
        Loader ld;
        ld.init(A, lda, rowsA, Bt, ldb, rowsB, kb, smem);
        constexpr int FSTEP = Cfg::ILV ? 8 * Cfg::WM : 8;  // row distance between consecutive fragments of a warp
        constexpr int FLAST = Cfg::frag_row(Cfg::WM - 1, 0);  // first row of the last warp row
        // per-warp live range for the partial-block case: frag_row(wm, mf) < rows_live  <=>  mf < hi
        const int fbase = Cfg::frag_row(wm, 0);
        int hi = (rows_live - fbase + FSTEP - 1) / FSTEP;
        hi = hi < 0 ? 0 : (hi > Cfg::MF ? Cfg::MF : hi);

        // The jump bounds are CTA-uniform: TRI_LOWER uses the bound of the LAST warp row, TRI_UPPER that of the FIRST
        // (with interleaved rows the other warp rows then run at most one dead fragment).
        auto panel_lower = [&](const double* sA, const double* sB, int p, int kp) {
            if (rows_live >= Cfg::BM) {
                const int rel = kp - tri0 - FLAST - 7;
                const int lo = rel <= 0 ? 0 : (rel + FSTEP - 1) / FSTEP;
                if (lo == 0)
                    mma_panel(acc, sA, sB, p, wm, wn, g, t);
                else
                    mma_panel_jump<false>(acc, sA, sB, p, wm, wn, g, t, lo);
            } else {  // partial block: exact per-warp bounds, predicated
                const int rel = kp - tri0 - fbase - 7;
                const int lo = rel <= 0 ? 0 : (rel + FSTEP - 1) / FSTEP;
                if (lo < hi) mma_panel_range(acc, sA, sB, p, wm, wn, g, t, lo, hi);
            }
        };
        auto panel_upper = [&](const double* sA, const double* sB, int p, int kp) {
            const int rel = kp + 7 - tri0;  // fragment live iff mf * FSTEP <= rel (first warp row)
            int up = rel < 0 ? 0 : rel / FSTEP + 1;
            if (up >= Cfg::MF)
                mma_panel(acc, sA, sB, p, wm, wn, g, t);
            else
                mma_panel_jump<true>(acc, sA, sB, p, wm, wn, g, t, up);
        };
        auto panel_full = [&](const double* sA, const double* sB, int p, int) { mma_panel(acc, sA, sB, p, wm, wn, g, t); };

#pragma unroll
        for (int s = 0; s < Cfg::STAGES - 1; ++s) {
            if (s < ktiles) ld.issue(s, s);
            cp_async_commit();
        }
        int stage = 0;  // stage holding k-tile kt
        int kt = 0;
        auto ktile = [&](auto&& panel) {
            cp_async_wait<Cfg::STAGES - 2>();
            __syncthreads();  // k-tile kt has landed for everyone; everyone is done reading k-tile kt-1
            const double* sA = smem + stage * Cfg::STAGE_DOUBLES;
            const double* sB = sA + Cfg::BM * Cfg::BK;
            const int kp0 = kb + kt * Cfg::BK;
            panel(sA, sB, 0, kp0);
            {   // refill the stage freed by k-tile kt-1, issued BEHIND the first panel's MMAs so that the address
                // arithmetic and the LDGSTS issue overlap with tensor work instead of idling the pipe
                int nk = kt + Cfg::STAGES - 1;
                int nstage = stage == 0 ? Cfg::STAGES - 1 : stage - 1;
                if (nk < ktiles) ld.issue(nk, nstage);
                cp_async_commit();
            }
#pragma unroll
            for (int p = 1; p < Cfg::PANELS; ++p) panel(sA, sB, p, kp0 + 8 * p);
            stage = stage + 1 == Cfg::STAGES ? 0 : stage + 1;
        };
        int kt_full = ktiles;  // end of the branch-free k-tiles
        if (MODE == TRI_UPPER) {
            // k-tiles whose first panel still has a dead fragment: kp0 + 7 - tri0 < (MF - 1) * FSTEP
            int kt_tri = (tri0 + (Cfg::MF - 1) * FSTEP - 7 - kb + Cfg::BK - 1) / Cfg::BK;
            kt_tri = kt_tri < 0 ? 0 : (kt_tri > ktiles ? ktiles : kt_tri);
            for (; kt < kt_tri; ++kt) ktile(panel_upper);
        }
        if (MODE == TRI_LOWER) {
            // k-tile kt is full iff its last panel kb + kt BK + 8 (PANELS - 1) <= tri0 + FLAST + 7
            const int num = tri0 + FLAST + 7 - 8 * (Cfg::PANELS - 1) - kb;
            kt_full = rows_live >= Cfg::BM ? (num < 0 ? 0 : num / Cfg::BK + 1) : 0;
            kt_full = kt_full > ktiles ? ktiles : kt_full;
        }
        for (; kt < kt_full; ++kt) ktile(panel_full);
        if (MODE == TRI_LOWER)
            for (; kt < ktiles; ++kt) ktile(panel_lower);
        cp_async_wait<0>();
        __syncthreads();  // smem may be reused by the caller (next tile / epilogue)
    }
};

// k-range of a tile from the structure flags
__device__ __forceinline__ void tile_k_range(int flags, int i0, int j0, int BM, int BN, int BK, int K, int& kb,
                                             int& ke) {
    kb = 0;
    ke = K;
    if (flags & GEMM_A_LOWER) ke = min(ke, i0 + BM);
    if (flags & GEMM_B_LOWER) ke = min(ke, j0 + BN);
    if (flags & GEMM_A_UPPER) kb = max(kb, i0);
    if (flags & GEMM_B_UPPER) kb = max(kb, j0);
    kb = (kb / BK) * BK;
    ke = min(K, ((ke + BK - 1) / BK) * BK);
    if (ke < kb) ke = kb;
}

struct GemmArgs {
    const double* A;
    const double* Bt;
    double* C;        // result (may be null if Ct is given)
    double* Ct;       // optional transposed copy of the result (may be null)
    const double* D;  // optional addend: C = alpha*A*Bt^T + D  (may alias C; null -> 0)
    const int* gate;  // optional per-batch switch: the launch is a no-op for batch entries with gate[z] == 0
    int64_t lda, ldb, ldc, ldct, ldd;
    int64_t strideA, strideB, strideC, strideCt, strideD;  // batch strides (grid.z)
    int M, N, K;
    double alpha;
    int flags;
    // "Node mode" (node_count > 0): one launch serves the same step of ALL nodes of one level of the triangular-inverse
    // tree (factor.cu, phase 2).  grid.z = batch * node_count; node t covers rows/columns [2 m t, 2 m t + 2 m) of the
    // npad x npad matrices, m = node_m, its second half clipped to m2 = min(m, node_total - 2 m t - m) rows (nodes
    // with m2 <= 0 do nothing).  The operands are located from the matrix BASE pointers:
    //   node_kind 1:  P^T = U11 L21^T      A = U (upper),  Bt = L,  C = scratch [t m^2 ..), row length m2
    //   node_kind 2:  X21 = -X22 P         A = Linv (lower), Bt = scratch,  C = Linv,  Ct = U
    int node_count, node_m, node_total, node_kind;
};

// Batched NT GEMM with optional dual (normal + transposed) store.  grid = (tiles_n, tiles_m, batch).
// MODE: TRI_LOWER needs GEMM_A_LOWER, TRI_UPPER needs GEMM_A_UPPER (fragment-level skipping inside the diagonal
// k-tiles, see Mainloop::run).  A triangular B operand is brought into this form by the caller computing the
// transposed product (the kernel stores both orientations anyway).
template <class Cfg, int MODE>
__global__ void __launch_bounds__(Cfg::THREADS, Cfg::MINB) gemm_nt_kernel(GemmArgs p) {
    extern __shared__ __align__(16) double smem[];
    pdl_wait();     // launched with the PDL attribute: everything below reads what the previous launch wrote
    pdl_trigger();  // the next launch of the chain may become resident as soon as all CTAs of this one have started
    int by = blockIdx.y, bx = blockIdx.x;
    int64_t z = blockIdx.z;
    if (p.flags & GEMM_ORDER_BATCH_FIRST) {
        z = blockIdx.x;
        bx = blockIdx.y;
        by = (p.flags & GEMM_A_LOWER) ? (int)(gridDim.z - 1 - blockIdx.z) : (int)blockIdx.z;
    }
    const int i0 = by * Cfg::BM, j0 = bx * Cfg::BN;
    if ((p.flags & GEMM_C_LOWER) && j0 > i0 + Cfg::BM - 1) return;
    const double* A = p.A;
    const double* Bt = p.Bt;
    double* C = p.C;
    double* Ct = p.Ct;
    int M = p.M, N = p.N, K = p.K;
    int64_t ldb = p.ldb, ldc = p.ldc;
    if (p.node_count > 0) {
        const int t = (int)(z % p.node_count);
        z /= p.node_count;
        const int m = p.node_m;
        const int64_t o = 2 * (int64_t)m * t;
        const int64_t rem = (int64_t)p.node_total - o - m;
        const int m2 = rem < m ? (int)rem : m;
        if (m2 <= 0) return;
        if (p.node_kind == 1) {
            A += o * p.lda + o;
            Bt += (o + m) * p.ldb + o;
            C += (int64_t)t * m * m;
            ldc = m2;
            M = m; N = m2; K = m;
        } else {
            A += (o + m) * (p.lda + 1);
            Bt += (int64_t)t * m * m;
            ldb = m2;
            C += (o + m) * p.ldc + o;
            Ct += o * p.ldct + (o + m);
            M = m2; N = m; K = m2;
        }
        if (i0 >= M || j0 >= N) return;
    }
    if (p.gate && p.gate[z] == 0) return;
    A += z * p.strideA + (int64_t)i0 * p.lda;
    Bt += z * p.strideB + (int64_t)j0 * ldb;
    int kb, ke;
    tile_k_range(p.flags, i0, j0, Cfg::BM, Cfg::BN, Cfg::BK, K, kb, ke);

    double acc[Cfg::MF][Cfg::NF][2];
#pragma unroll
    for (int mf = 0; mf < Cfg::MF; ++mf)
#pragma unroll
        for (int nf = 0; nf < Cfg::NF; ++nf) acc[mf][nf][0] = acc[mf][nf][1] = 0.0;

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int wm = warp / Cfg::WN, wn = warp % Cfg::WN;
    C = C ? C + z * p.strideC : nullptr;
    Ct = Ct ? Ct + z * p.strideCt : nullptr;
    const double* D = p.D ? p.D + z * p.strideD : nullptr;
    if (D && t == 0) {
        // The addend tile is read once, in the epilogue, straight from HBM when many matrices are in flight: ask L2 for
        // its 64-byte row segments now, so that the k loop hides the DRAM latency instead of the epilogue exposing it
#pragma unroll
        for (int mf = 0; mf < Cfg::MF; ++mf) {
            const int row = i0 + Cfg::frag_row(wm, mf) + g;
            if (row >= M) continue;
#pragma unroll
            for (int nf = 0; nf < Cfg::NF; ++nf) {
                const int col = j0 + wn * Cfg::WTN + nf * 8;
                if (col < N) asm volatile("prefetch.global.L2 [%0];" ::"l"(D + (int64_t)row * p.ldd + col));
            }
        }
    }

    Mainloop<Cfg>::template run<MODE>(acc, A, p.lda, min(Cfg::BM, M - i0), Bt, ldb, min(Cfg::BN, N - j0), kb, ke, smem, i0);

#pragma unroll
    for (int mf = 0; mf < Cfg::MF; ++mf) {
        int row = i0 + Cfg::frag_row(wm, mf) + g;
        if (row >= M) continue;
#pragma unroll
        for (int nf = 0; nf < Cfg::NF; ++nf) {
            int col = j0 + wn * Cfg::WTN + nf * 8 + 2 * t;
            if (col >= N) continue;  // N is even, so col+1 < N too
            double v0 = p.alpha * acc[mf][nf][0], v1 = p.alpha * acc[mf][nf][1];
            if (D) {
                double2 old = *reinterpret_cast<const double2*>(D + (int64_t)row * p.ldd + col);
                v0 += old.x;
                v1 += old.y;
            }
            if (C) *reinterpret_cast<double2*>(C + (int64_t)row * ldc + col) = make_double2(v0, v1);
            if (Ct) {
                Ct[(int64_t)col * p.ldct + row] = v0;
                Ct[(int64_t)(col + 1) * p.ldct + row] = v1;
            }
        }
    }
}

// ---- predictive variance: var_j = kk - sum_i ( sum_{k<=i} Linv[i][k] Kstar[j][k] )^2 -----------------------
// One CTA owns BN queries and sweeps all row blocks of Linv (a lower-triangular NT product), squaring and
// summing each finished BM x BN block of V into per-query registers.  V never leaves the SM.
template <class Cfg, bool SPLIT = false>
__global__ void __launch_bounds__(Cfg::THREADS, Cfg::MINB)
    trmm_sumsq_kernel(const double* __restrict__ Linv, int n, int npad, const double* __restrict__ Kstar, int64_t ldk,
                      int64_t q_begin, int64_t M, double kk, double scale, int standardised,
                      double* __restrict__ var_out, double* __restrict__ partial, int64_t partial_ld) {
    // SPLIT ("row split", gridDim.y > 1, used when there are too few query tiles to fill the machine): CTA (x, y) only sweeps
    // the row blocks y, y + gridDim.y, ... and writes its partial column sums to partial[y][query]; a finishing kernel
    // adds them in fixed order.
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

    const int kmax = ((n + Cfg::BK - 1) / Cfg::BK) * Cfg::BK;  // K* columns >= n are zero
    // Row blocks are aligned to the END of the matrix: a partial block (n mod BM rows) is the FIRST one, where the
    // k range is shortest, instead of the last one, where it is longest.
    // (n is first rounded up to the 8-row fragment granularity: the extra rows are identity rows of the padded
    // Linv, which only meet zero columns of K*.)
    const int n8 = (n + 7) & ~7;
    const int r_first = n8 % Cfg::BM;
    const int first = r_first ? r_first : Cfg::BM;       // rows of block 0
    const int nblk = (n8 + Cfg::BM - 1) / Cfg::BM;
    for (int b = 0; b < nblk; ++b) {
        if (SPLIT) {
            // Row block b costs ~(b + 1) k-blocks.  Dealing the blocks out boustrophedon-wise (0 1 2 3 3 2 1 0 0 1 ...)
            // gives every CTA of a query tile the same triangular work (a plain stride would leave the last CTA with
            // up to 40 % more than the first); counted from the LAST block, so that an incomplete final round consists
            // of the shortest blocks (4 blocks over 3 CTAs: loads 4 / 3 / 2+1 instead of 1 / 2 / 3+4).
            const int S = gridDim.y, r = (nblk - 1 - b) % (2 * S);
            if ((r < S ? r : 2 * S - 1 - r) != (int)blockIdx.y) continue;
        }
        const int i0 = b == 0 ? 0 : first + (b - 1) * Cfg::BM;
        double acc[Cfg::MF][Cfg::NF][2];
#pragma unroll
        for (int mf = 0; mf < Cfg::MF; ++mf)
#pragma unroll
            for (int nf = 0; nf < Cfg::NF; ++nf) acc[mf][nf][0] = acc[mf][nf][1] = 0.0;
        const int rows_live = b == 0 ? first : Cfg::BM;
        int ke = min(kmax, ((i0 + rows_live + Cfg::BK - 1) / Cfg::BK) * Cfg::BK);
        // Linv is lower triangular: row i0 + r is zero beyond column i0 + r
        Mainloop<Cfg>::template run<TRI_LOWER>(acc, Linv + (int64_t)i0 * npad, npad, min(Cfg::BM, npad - i0), Bt, ldk, Cfg::BN, 0,
                                          ke, smem, i0, rows_live);
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
        if (SPLIT) {
            partial[(int64_t)blockIdx.y * partial_ld + j0 + c] = s;
        } else if (q < M) {
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

// rows are interleaved between the two warp rows in every configuration (harmless for unstructured products)
// tools/gemm_tune.cu sweep (profiles/r01/gemm_tune.txt): 16 warps (4 per scheduler) with 32-wide k-tiles beat the
// 8-warp / 16-wide configuration by 5 % (2.40 -> 2.28 ms per 18,944-query chunk): twice the warps to cover the
// MMA and LDS latencies, half the barriers per flop.
using CfgBig = TileCfg<128, 128, 4, 4, 3, 32, true>;   // 512 threads, warp tile 32x32, 3 stages x 64 KB = 192 KB smem
using CfgTrmm = CfgBig;
using CfgSmall = TileCfg<64, 64, 2, 2, 3, 16, true, 4>;  // 128 threads, warp tile 32x32, 48 KB smem: four CTAs per SM
using CfgMed = TileCfg<128, 64, 4, 2, 4, 16, true, 2>; // 256 threads, warp tile 32x32, 96 KB smem: two CTAs per SM
// Launches whose 64x64 tiles cannot give every SM a tile (the k = 128 products on the critical path of the tile-column
// factorisation: one 64x64x128 tile is 4.2 us of ONE SM's tensor pipe): half-height tiles, twice the CTAs
using CfgTiny = TileCfg<32, 64, 1, 4, 3, 16, true, 6>;  // 128 threads, warp tile 32x16, 36 KB smem: six CTAs per SM

int32_t launch_gemm_nt(cudaStream_t stream, const GemmArgs& args, int batch);

}  // namespace bobe
