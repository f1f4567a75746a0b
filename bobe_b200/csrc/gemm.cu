// Launchers for the DMMA NT-GEMM and the fused "triangular multiply + column sum of squares" kernel that
// produces the predictive variance term ||L^-1 k*||^2 without ever writing V = L^-1 K* to memory.
#include <algorithm>
#include <cstdarg>
#include <cstdlib>
#include <mutex>
#include <type_traits>

#include <cudaTypedefs.h>

#include "gemm_nt.cuh"
#include "kernels.cuh"
#include "trmm_tma.cuh"

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

int64_t env_int(const char* name, int64_t dflt) {
    const char* v = getenv(name);
    return (v && *v) ? atoll(v) : dflt;
}

// ---- TMA tensor maps (driver entry point fetched through the runtime: no link dependency on libcuda) ----------------
static PFN_cuTensorMapEncodeTiled_v12000 tensor_map_encoder() {
    static PFN_cuTensorMapEncodeTiled_v12000 fn = []() -> PFN_cuTensorMapEncodeTiled_v12000 {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            return nullptr;
        return (PFN_cuTensorMapEncodeTiled_v12000)p;
    }();
    return fn;
}
// row-major (rows, cols) float64 matrix with leading dimension ld; box = one k8 panel of `box_rows` rows
static bool make_panel_map(CUtensorMap* map, const double* base, int64_t rows, int64_t cols, int64_t ld, int box_rows) {
    auto enc = tensor_map_encoder();
    if (!enc) return false;
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 8};
    cuuint32_t box[2] = {8, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, (void*)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// batched variant: (batch, rows, cols) with row stride ld and matrix stride `stride` (in doubles); box = one k8 panel
static bool make_panel_map3(CUtensorMap* map, const double* base, int64_t batch, int64_t rows, int64_t cols, int64_t ld,
                            int64_t stride, int box_rows) {
    auto enc = tensor_map_encoder();
    if (!enc) return false;
    if (batch <= 1 || stride <= 0) stride = rows * ld;
    cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)rows, (cuuint64_t)(batch < 1 ? 1 : batch)};
    cuuint64_t strides[2] = {(cuuint64_t)ld * 8, (cuuint64_t)stride * 8};
    cuuint32_t box[3] = {8, (cuuint32_t)box_rows, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, (void*)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

static int sm_count() {
    static const int n = []() {
        int dev = 0, v = 148;
        if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
        return v > 0 ? v : 148;
    }();
    return n;
}

bool pdl_enabled(int64_t ctas) {
    static const bool on = env_int("BOBE_PDL", 1) != 0;
    static const int64_t max_ctas = env_int("BOBE_PDL_MAX_CTAS", 592);
    return on && ctas <= max_ctas;
}

int32_t launch_gemm_nt(cudaStream_t stream, const GemmArgs& a, int batch) {
    if (a.M <= 0 || a.N <= 0 || batch <= 0) return BOBE_OK;
    static const int64_t forced_tile = env_int("BOBE_TILE", 0);
    const int kgran = (forced_tile > 2 || env_int("BOBE_GEMM_TMA", 0) != 0) ? 32 : 16;  // k-tile width of the tiles in use
    if ((a.K % kgran) || (a.N % 2) || (a.lda % 2) || (a.ldb % 2) || (a.ldc % 2)) {
        set_error("gemm_nt: K must be a multiple of %d and N/ld even (M=%d N=%d K=%d)", kgran, a.M, a.N, a.K);
        return BOBE_E_ARG;
    }
    // Tile choice (measured, profiles/r01/README.md "tile choice"): 64x64 tiles with FOUR CTAs per SM beat the 128x128 /
    // one-CTA-per-SM configuration for every product issued through this launcher -- 21.6 vs 23.2 ms for the 64-restart
    // factorisation, 51.5 vs 56.5 ms for the WIPV solve.  The k loops here are short (<= n/2 for the recursion), so a
    // lone CTA leaves the tensor pipe idle during its pipeline fill and its dual-store epilogue; with four co-resident
    // CTAs another one is always in its main loop, and small CTAs also fit beside the leaf kernels of other streams.
    // The 128x128 configuration stays the right one for trmm_sumsq (k loops up to n, no epilogue stores).
    static const int64_t forced = env_int("BOBE_TILE", 0);  // experiment knob: 1 small, 2 medium, 3 big
    int tile = 1;
    if (forced && !(a.M <= 64 || a.N <= 64)) tile = (int)forced;
    if (!a.C && !a.Ct) {
        set_error("gemm_nt: no output");
        return BOBE_E_ARG;
    }
    const int mode = (a.flags & GEMM_A_LOWER) ? TRI_LOWER : ((a.flags & GEMM_A_UPPER) ? TRI_UPPER : TRI_NONE);
    // Products whose 128 x 128 tiles can fill the machine: the persistent TMA kernel (trmm_tma.cuh)
    static const int64_t tma_gemm = env_int("BOBE_GEMM_TMA", 0);
    static const int64_t tma_min_tiles = env_int("BOBE_GEMM_TMA_MIN_TILES", 2);  // live tiles per SM
    // BOBE_GEMM_TMA: 1 = every large product, 2 = only products without a triangular row operand (experiment knobs)
    if (tma_gemm && !forced && a.node_count == 0 && a.M >= 256 && a.N >= (tma_gemm == 2 ? 128 : 256) && a.K >= 128 &&
        (tma_gemm != 2 || mode == TRI_NONE) && ((((uintptr_t)a.A) | ((uintptr_t)a.Bt)) & 15) == 0 &&
        (a.strideA % 2) == 0 && (a.strideB % 2) == 0) {
        using Cfg = CfgBig;
        const int tiles_m = (a.M + Cfg::BM - 1) / Cfg::BM, tiles_n = (a.N + Cfg::BN - 1) / Cfg::BN;
        const int64_t total = (int64_t)tiles_m * tiles_n * batch;
        const int64_t live = (a.flags & GEMM_C_LOWER) ? total / 2 : total;
        CUtensorMap mapA, mapB;
        if (live >= tma_min_tiles * sm_count() && total < (int64_t)1 << 30 &&
            make_panel_map3(&mapA, a.A, batch, a.M, a.K, a.lda, a.strideA, Cfg::BM) &&
            make_panel_map3(&mapB, a.Bt, batch, a.N, a.K, a.ldb, a.strideB, Cfg::BN)) {
            constexpr int TMA_SMEM = Cfg::SMEM_BYTES + 128;
            const unsigned grid = (unsigned)std::min<int64_t>(total, sm_count());
            auto go_tma = [&](auto mode_c) -> int32_t {
                constexpr int MODE = decltype(mode_c)::value;
                if (int32_t rc = ensure_smem<gemm_nt_tma_kernel<Cfg, MODE>>(TMA_SMEM)) return rc;
                gemm_nt_tma_kernel<Cfg, MODE><<<grid, Cfg::THREADS, TMA_SMEM, stream>>>(mapA, mapB, a, tiles_m, tiles_n, (int)total);
                return check_launch("gemm_nt_tma_kernel");
            };
            if (mode == TRI_LOWER) return go_tma(std::integral_constant<int, TRI_LOWER>{});
            if (mode == TRI_UPPER) return go_tma(std::integral_constant<int, TRI_UPPER>{});
            return go_tma(std::integral_constant<int, TRI_NONE>{});
        }
    }
    // small batches of matrices with triangular structure: batch index fastest, heavy tile rows first (GEMM_ORDER_BATCH_FIRST)
    static const int64_t bf_max = env_int("BOBE_BATCH_FIRST_MAX", 8);
    GemmArgs ao = a;
    const int64_t zdim = (int64_t)batch * (a.node_count > 0 ? a.node_count : 1);
    if (zdim > 1 && zdim <= bf_max && zdim <= 65535 && (a.flags & (GEMM_A_LOWER | GEMM_A_UPPER)) && a.M >= 256)
        ao.flags |= GEMM_ORDER_BATCH_FIRST;
    auto go = [&](auto cfg, auto mode_c) -> int32_t {
        using Cfg = decltype(cfg);
        constexpr int MODE = decltype(mode_c)::value;
        if (int32_t rc = ensure_smem<gemm_nt_kernel<Cfg, MODE>>(Cfg::SMEM_BYTES)) return rc;
        const unsigned tn = (a.N + Cfg::BN - 1) / Cfg::BN, tm = (a.M + Cfg::BM - 1) / Cfg::BM;
        dim3 grid = (ao.flags & GEMM_ORDER_BATCH_FIRST) ? dim3((unsigned)zdim, tn, tm) : dim3(tn, tm, (unsigned)zdim);
        if (launch_pdl(gemm_nt_kernel<Cfg, MODE>, grid, dim3(Cfg::THREADS), Cfg::SMEM_BYTES, stream, ao) != cudaSuccess) {
            set_error("gemm_nt: launch failed: %s", cudaGetErrorString(cudaGetLastError()));
            return BOBE_E_CUDA;
        }
        return BOBE_OK;
    };
    using M0 = std::integral_constant<int, TRI_NONE>;
    using M1 = std::integral_constant<int, TRI_LOWER>;
    using M2 = std::integral_constant<int, TRI_UPPER>;
    int32_t rc;
    // under-filled launch: fewer 64x64 tiles than SMs (BOBE_TINY_MAX_TILES, default two per SM) -> 32x64 tiles
    static const int64_t tiny_max = env_int("BOBE_TINY_MAX_TILES", 2 * sm_count());
    const int64_t nodes = a.node_count > 0 ? a.node_count : 1;
    int64_t tiles64 = (int64_t)((a.M + 63) / 64) * ((a.N + 63) / 64) * batch * nodes;
    if (a.flags & GEMM_C_LOWER) tiles64 = tiles64 / 2 + ((a.M < a.N ? a.M : a.N) + 63) / 64 * batch;
    // ... and every 128-column product (measured at batches of 8 / 16 / 64 matrices: profiles/r02/gemm_bench*.txt)
    static const int64_t tiny_narrow = env_int("BOBE_TINY_NARROW", 1);
    if (tile == 1 && !forced && (tiles64 <= tiny_max || (tiny_narrow && a.N <= 128)) && a.M > 32)
        rc = mode == TRI_LOWER ? go(CfgTiny{}, M1{}) : (mode == TRI_UPPER ? go(CfgTiny{}, M2{}) : go(CfgTiny{}, M0{}));
    else if (tile == 1)
        rc = mode == TRI_LOWER ? go(CfgSmall{}, M1{}) : (mode == TRI_UPPER ? go(CfgSmall{}, M2{}) : go(CfgSmall{}, M0{}));
    else if (tile == 2)
        rc = mode == TRI_LOWER ? go(CfgMed{}, M1{}) : (mode == TRI_UPPER ? go(CfgMed{}, M2{}) : go(CfgMed{}, M0{}));
    else
        rc = mode == TRI_LOWER ? go(CfgBig{}, M1{}) : (mode == TRI_UPPER ? go(CfgBig{}, M2{}) : go(CfgBig{}, M0{}));
    if (rc) return rc;
    return check_launch("gemm_nt_kernel");
}

// VT[q][i] = sum_k Linv[i][k] Kin[q][k] for rows_pad (multiple of 128) rows of Kin, through the TMA trmm pipeline (STORE).
// Returns 1 if this route is not available (the caller then uses launch_gemm_nt), 0 on success, < 0 on error.
int32_t launch_trmm_store(cudaStream_t stream, const double* Linv, int n, int npad, const double* Kin, int64_t ldk,
                          int64_t rows_pad, double* VT, int64_t ldv) {
    using Cfg = CfgTrmm;
    static const int64_t use_tma = env_int("BOBE_TRMM_TMA", 1);
    const int64_t qtiles = rows_pad / Cfg::BN;
    // one 128-query tile per CTA sweeps the whole triangular product: only worth it when the tiles fill the machine
    if (!use_tma || rows_pad % Cfg::BN || qtiles < (sm_count() * 3) / 4 || ((((uintptr_t)Kin) | ((uintptr_t)Linv)) & 15) || (ldk % 2))
        return 1;
    CUtensorMap mapA, mapB;
    if (!make_panel_map(&mapA, Linv, npad, npad, npad, Cfg::BM) || !make_panel_map(&mapB, Kin, rows_pad, npad, ldk, Cfg::BN))
        return 1;
    constexpr int TMA_SMEM = Cfg::SMEM_BYTES + 128;
    if (int32_t rc = ensure_smem<trmm_sumsq_tma_kernel<Cfg, true>>(TMA_SMEM)) return rc;
    trmm_sumsq_tma_kernel<Cfg, true><<<(unsigned)qtiles, Cfg::THREADS, TMA_SMEM, stream>>>(mapA, mapB, n, npad, 0, rows_pad, 0.0, 1.0, 0,
                                                                                         nullptr, VT, ldv);
    return check_launch("trmm_sumsq_tma_kernel<STORE>");
}

// var = kk - sum over the row-split partial sums, with the floor / scale semantics of the fused path
__global__ void __launch_bounds__(256) trmm_finish_kernel(const double* __restrict__ partial, int64_t ld, int split,
                                                          int64_t rows, int64_t q_begin, int64_t M, double kk, double scale,
                                                          int standardised, double* __restrict__ var_out) {
    const int64_t c = (int64_t)blockIdx.x * 256 + threadIdx.x;
    const int64_t q = q_begin + c;
    if (c >= rows || q >= M) return;
    double s = 0.0;
    for (int g = 0; g < split; ++g) s += partial[(int64_t)g * ld + c];
    double var = kk - s;
    if (standardised) {
        if (isnan(var)) var = SAFE_FLOOR;
        if (var < SAFE_FLOOR) var = SAFE_FLOOR;
    } else {
        if (var < SAFE_FLOOR) var = SAFE_FLOOR;
        var *= scale;
    }
    var_out[q] = var;
}

// Shared-panel schedule of trmm_sumsq_tma_kernel: S CTAs per K* panel.  The smallest S <= 4 for which the panels in flight
// (148 / S panels of 128 x npad doubles) fit the L2 budget AND the row blocks deal out evenly; 1 = every CTA its own panel.
// BOBE_TRMM_SHARE: 1 = off (default: the schedule is 2 % slower -- see profiles/r02/trmm_shared_panels.txt), 0 = auto,
// S = forced (if it deals out evenly).
static int trmm_share(int n, int npad) {
    using Cfg = CfgTrmm;
    static const int64_t forced = env_int("BOBE_TRMM_SHARE", 1);
    static const int64_t budget_mb = env_int("BOBE_TRMM_L2_MB", 80);
    if (forced == 1) return 1;
    const int n8 = (n + 7) & ~7, nblk = (n8 + Cfg::BM - 1) / Cfg::BM;
    const int r_first = n8 % Cfg::BM, first = r_first ? r_first : Cfg::BM;
    const int kmax = ((n + Cfg::BK - 1) / Cfg::BK) * Cfg::BK;
    auto weight = [&](int b) {  // k-tiles x live rows of row block b (the kernel's own sweep)
        const int row0 = b == 0 ? 0 : first + (b - 1) * Cfg::BM, live = b == 0 ? first : Cfg::BM;
        const int ke = std::min(kmax, ((row0 + live + Cfg::BK - 1) / Cfg::BK) * Cfg::BK);
        return (double)(ke / Cfg::BK) * live;
    };
    const double panel_mb = 128.0 * npad * 8.0 / (1024.0 * 1024.0);
    int s_min = 1;
    while (s_min < 4 && (148 / s_min) * panel_mb > (double)budget_mb) ++s_min;
    if (forced > 1) s_min = (int)std::min<int64_t>(forced, 8);
    for (int S = s_min; S <= (forced > 1 ? s_min : 4); ++S) {
        if (S == 1) return 1;
        double w[8] = {0, 0, 0, 0, 0, 0, 0, 0}, total = 0.0, worst = 0.0;
        for (int m = 0; m < S; ++m)
            for (int i = 0;; ++i) {
                const int b = nblk - 1 - ((i >> 1) * 2 * S + ((i & 1) ? 2 * S - 1 - m : m));
                if (b < 0) break;
                w[m] += weight(b);
            }
        for (int m = 0; m < S; ++m) {
            total += w[m];
            worst = std::max(worst, w[m]);
        }
        if (worst * S <= 1.03 * total) return S;
    }
    return 1;
}

int trmm_chunk_tiles(int n, int npad) {
    const int S = trmm_share(n, npad);
    return (148 / S) * S;
}

int32_t launch_trmm_sumsq(cudaStream_t stream, const double* Linv, int n, int npad, const double* Kstar, int64_t ldk,
                          int64_t rows_pad, int64_t q_begin, int64_t M, double kk, double scale, int standardised,
                          double* var_out, double* partial, int* counters) {
    using Cfg = CfgTrmm;
    if (rows_pad % Cfg::BN) {
        set_error("trmm_sumsq: query chunk must be padded to %d", Cfg::BN);
        return BOBE_E_ARG;
    }
    if (int32_t rc = ensure_smem<trmm_sumsq_kernel<Cfg, false>>(Cfg::SMEM_BYTES)) return rc;
    if (int32_t rc = ensure_smem<trmm_sumsq_kernel<Cfg, true>>(Cfg::SMEM_BYTES)) return rc;
    const int qtiles = (int)(rows_pad / Cfg::BN);
    const int nblk = (((n + 7) & ~7) + Cfg::BM - 1) / Cfg::BM;
    // too few query tiles for 148 SMs: split the row blocks of Linv over several CTAs per tile (TRMM_MAX_SPLIT rows of
    // `partial`, each rows_pad long)
    int split = 1;
    if (partial && qtiles < 74 && nblk > 1) {
        split = 148 / qtiles;
        split = split > nblk ? nblk : split;
        split = split > TRMM_MAX_SPLIT ? TRMM_MAX_SPLIT : split;
    }
    // L2-resident schedule (bobe_predict with BOBE_TRMM_SPLIT = S > 1 sends chunks of 148 / S query tiles): S CTAs share
    // one K* panel, so the panels in flight (148 / S x 128 x npad x 8 bytes) stay in L2 between the row-block sweeps
    static const int64_t forced = env_int("BOBE_TRMM_SPLIT", 1);
    if (partial && forced > 1 && nblk >= forced && qtiles * forced <= 148 && split < forced) split = (int)forced;
    dim3 grid(qtiles, split);
    static const int64_t use_tma = env_int("BOBE_TRMM_TMA", 1);  // 0: the cp.async kernel (also the fallback below)
    if (split == 1 && use_tma && (((uintptr_t)Kstar | (uintptr_t)Linv) & 15) == 0 && (ldk % 2) == 0) {
        CUtensorMap mapA, mapB;
        if (make_panel_map(&mapA, Linv, npad, npad, npad, Cfg::BM) && make_panel_map(&mapB, Kstar, rows_pad, npad, ldk, Cfg::BN)) {
            constexpr int TMA_SMEM = Cfg::SMEM_BYTES + 128;  // room to align the stages to 128 bytes
            if (int32_t rc = ensure_smem<trmm_sumsq_tma_kernel<Cfg>>(TMA_SMEM)) return rc;
            const int share = (partial && counters && qtiles <= TRMM_COUNTERS) ? trmm_share(n, npad) : 1;
            if (share > 1) {
                const int groups = std::min(148 / share, qtiles);
                if (int32_t rc = ensure_smem<trmm_sumsq_tma_kernel<Cfg, false, true>>(TMA_SMEM)) return rc;
                trmm_sumsq_tma_kernel<Cfg, false, true><<<groups * share, Cfg::THREADS, TMA_SMEM, stream>>>(
                    mapA, mapB, n, npad, q_begin, M, kk, scale, standardised, var_out, nullptr, 0, share, qtiles, partial,
                    rows_pad, counters);
            } else {
                trmm_sumsq_tma_kernel<Cfg><<<grid, Cfg::THREADS, TMA_SMEM, stream>>>(mapA, mapB, n, npad, q_begin, M, kk,
                                                                                       scale, standardised, var_out);
            }
            return check_launch("trmm_sumsq_tma_kernel");
        }
    }
    if (split > 1)
        trmm_sumsq_kernel<Cfg, true><<<grid, Cfg::THREADS, Cfg::SMEM_BYTES, stream>>>(Linv, n, npad, Kstar, ldk, q_begin, M, kk,
                                                                                   scale, standardised, var_out, partial,
                                                                                   rows_pad);
    else
        trmm_sumsq_kernel<Cfg, false><<<grid, Cfg::THREADS, Cfg::SMEM_BYTES, stream>>>(Linv, n, npad, Kstar, ldk, q_begin, M, kk,
                                                                                    scale, standardised, var_out, nullptr, 0);
    if (split > 1)
        trmm_finish_kernel<<<(unsigned)((rows_pad + 255) / 256), 256, 0, stream>>>(partial, rows_pad, split, rows_pad, q_begin, M,
                                                                                 kk, scale, standardised, var_out);
    return check_launch("trmm_sumsq_kernel");
}

}  // namespace bobe
