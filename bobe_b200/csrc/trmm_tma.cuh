// TMA-fed variant of trmm_sumsq_kernel (gemm_nt.cuh): same tiles, same fragment ownership, same order of every sum (the
// results are bitwise those of the cp.async kernel), but
//   * the k-tiles of Linv and K* are brought in by cp.async.bulk.tensor (TMA) issued by ONE lane and land on a
//     transaction mbarrier per pipeline stage, instead of 8 LDGSTS per thread and k-tile;
//   * there is no block-wide barrier in the k loop: a warp waits on the "full" mbarrier of its stage, computes, and counts
//     itself out of the stage; the LAST warp to leave a stage issues its refill.  Warps drift apart by up to two k-tiles
//     instead of meeting at every k-tile;
//   * the (row block, k-tile) items of the whole triangular sweep form ONE pipeline: the first k-tiles of the next row
//     block are in flight while the current one finishes (the cp.async kernel drains and refills at each of the 16 row
//     blocks).
// Shared-memory layout per stage and operand is the one of gemm_nt.cuh (BK/8 "k8 panels" of rows x 64 bytes): one 2-D
// TMA box {8 doubles, 128 rows} is exactly one panel, so no swizzle is involved.
#pragma once
#include <cuda.h>

#include "gemm_nt.cuh"

namespace bobe {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// one 2-D box (c0 = first element along the contiguous dimension, c1 = first row) -> dense shared memory
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
            smem_u32(dst)),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}

// STORE = true: the same sweep, but instead of squaring V = Linv K*^T into per-query sums the finished row blocks are written
// out transposed, VT[q][i] (leading dimension ldv, rows i in [n8, npad) zeroed): the shared solve of the fantasy variance
// (bobe_fantasy_var), whose K* panel is K(MC, X).  The values are bitwise those of gemm_nt_kernel for the same product.
template <class Cfg, bool STORE = false, bool SHARED = false>
__global__ void __launch_bounds__(Cfg::THREADS, 1)
    trmm_sumsq_tma_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, int n, int npad,
                          int64_t q_begin, int64_t M, double kk, double scale, int standardised,
                          double* __restrict__ var_out, double* __restrict__ vt_out = nullptr, int64_t ldv = 0,
                          int share = 1, int qtiles = 0, double* __restrict__ partial = nullptr, int64_t pstride = 0,
                          int* __restrict__ counters = nullptr) {
    using ML = Mainloop<Cfg>;
    constexpr int STAGES = Cfg::STAGES, NWARPS = Cfg::THREADS / 32;
    constexpr uint32_t STAGE_BYTES = Cfg::STAGE_DOUBLES * 8;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // TMA destinations must be 128-byte aligned; the dynamic window starts behind the static arrays below
    double* smem = reinterpret_cast<double*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~(uintptr_t)127);
    __shared__ __align__(8) uint64_t full[STAGES];
    __shared__ int left[STAGES];  // warps that have not yet left the stage
    __shared__ double red[Cfg::WM][Cfg::BN];
    // SHARED (sum-of-squares mode only; share = S > 1): S consecutive CTAs form a group that walks the query tiles
    // group, group + G, ... (G groups) TOGETHER, each member taking the row blocks m, 2S-1-m, 2S+m, 4S-1-m, ... (counted from the last
    // one) of every tile (dealt out boustrophedon-wise: equal triangular work), so only G K* panels are in flight at a time and their re-reads
    // (once per row block) hit L2 instead of DRAM.  The members' partial column sums meet in `partial`; the last member to
    // finish a tile adds them in member order (fixed order: deterministic) and writes the variances.
    constexpr bool shared_mode = SHARED && !STORE;
    const int S = shared_mode ? share : 1;
    const int member = shared_mode ? (int)blockIdx.x % S : 0;
    const int group = shared_mode ? (int)blockIdx.x / S : (int)blockIdx.x;
    const int G = shared_mode ? (int)gridDim.x / S : 1;
    const int ntile_mine = shared_mode ? (group < qtiles ? (qtiles - group + G - 1) / G : 0) : 1;
    __shared__ int last_flag;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int wm = warp / Cfg::WN, wn = warp % Cfg::WN;

    // the sweep: row blocks aligned to the END of the matrix (see trmm_sumsq_kernel), k range of block b = [0, ke_b)
    const int kmax = ((n + Cfg::BK - 1) / Cfg::BK) * Cfg::BK;
    const int n8 = (n + 7) & ~7;
    const int r_first = n8 % Cfg::BM;
    const int first = r_first ? r_first : Cfg::BM;
    const int nblk = (n8 + Cfg::BM - 1) / Cfg::BM;
    auto row0 = [&](int b) { return b == 0 ? 0 : first + (b - 1) * Cfg::BM; };
    auto ktiles_of = [&](int b) {
        const int live = b == 0 ? first : Cfg::BM;
        const int ke = min(kmax, ((row0(b) + live + Cfg::BK - 1) / Cfg::BK) * Cfg::BK);
        return ke / Cfg::BK;
    };

    // the i-th row block of this CTA and how many it has (shared mode: dealt from the LAST, longest block downwards, so the
    // incomplete final round consists of the shortest blocks and every member gets the same work for any n)
    auto blk = [&](int i) {
        return shared_mode ? nblk - 1 - ((i >> 1) * 2 * S + ((i & 1) ? 2 * S - 1 - member : member)) : i;
    };
    int nbm = nblk;
    if (shared_mode) {
        nbm = 0;
        while (blk(nbm) >= 0) ++nbm;
    }
    const int tiles_run = nbm > 0 ? ntile_mine : 0;  // (a member without row blocks only joins the final sums)

    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full[s], 1);
            left[s] = NWARPS;
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    // issue the TMA loads of item (tile u of this CTA, its row block i, k-tile kt) into `stage` (one lane)
    auto issue = [&](int u, int i, int kt, int stage) {
        double* sA = smem + stage * Cfg::STAGE_DOUBLES;
        double* sB = sA + Cfg::BM * Cfg::BK;
        mbar_expect_tx(&full[stage], STAGE_BYTES);
        const int i0 = row0(blk(i)), k0 = kt * Cfg::BK, jq = (group + u * G) * Cfg::BN;
#pragma unroll
        for (int p = 0; p < Cfg::PANELS; ++p) {
            tma_load_2d(sA + p * Cfg::BM * 8, &mapA, &full[stage], k0 + 8 * p, i0);
            tma_load_2d(sB + p * Cfg::BN * 8, &mapB, &full[stage], k0 + 8 * p, jq);
        }
    };

    // look-ahead iterator (item + STAGES), kept by every warp's lane 0
    int lu = 0, lb = 0, lkt = 0;
    auto advance = [&](int& u, int& i, int& kt) {
        if (u >= tiles_run) return;
        if (++kt == ktiles_of(blk(i))) {
            kt = 0;
            if (++i == nbm) {
                i = 0;
                ++u;
            }
        }
    };
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            if (lu < tiles_run) issue(lu, lb, lkt, s);
            advance(lu, lb, lkt);
        }
    } else {
        for (int s = 0; s < STAGES; ++s) advance(lu, lb, lkt);
    }

    double colsum[Cfg::NF][2];
#pragma unroll
    for (int nf = 0; nf < Cfg::NF; ++nf) colsum[nf][0] = colsum[nf][1] = 0.0;

    constexpr int FSTEP = Cfg::ILV ? 8 * Cfg::WM : 8;
    constexpr int FLAST = Cfg::frag_row(Cfg::WM - 1, 0);
    const int fbase = Cfg::frag_row(wm, 0);
    int stage = 0;
    uint32_t parity = 0;
    for (int u = 0; u < ntile_mine; ++u) {
    const int j0 = (group + u * G) * Cfg::BN;
    for (int bi = 0; bi < nbm; ++bi) {
        const int b = blk(bi);
        const int i0 = row0(b);
        const int rows_live = b == 0 ? first : Cfg::BM;
        const int ktiles = ktiles_of(b);
        int hi = (rows_live - fbase + FSTEP - 1) / FSTEP;
        hi = hi < 0 ? 0 : (hi > Cfg::MF ? Cfg::MF : hi);
        // k-tiles without a dead fragment (Mainloop::run, TRI_LOWER)
        const int num = i0 + FLAST + 7 - 8 * (Cfg::PANELS - 1);
        int kt_full = rows_live >= Cfg::BM ? (num < 0 ? 0 : num / Cfg::BK + 1) : 0;
        kt_full = kt_full > ktiles ? ktiles : kt_full;

        double acc[Cfg::MF][Cfg::NF][2];
#pragma unroll
        for (int mf = 0; mf < Cfg::MF; ++mf)
#pragma unroll
            for (int nf = 0; nf < Cfg::NF; ++nf) acc[mf][nf][0] = acc[mf][nf][1] = 0.0;

        for (int kt = 0; kt < ktiles; ++kt) {
            mbar_wait(&full[stage], parity);
            const double* sA = smem + stage * Cfg::STAGE_DOUBLES;
            const double* sB = sA + Cfg::BM * Cfg::BK;
            if (kt < kt_full) {
#pragma unroll
                for (int p = 0; p < Cfg::PANELS; ++p) ML::mma_panel(acc, sA, sB, p, wm, wn, g, t);
            } else {
                const int kp0 = kt * Cfg::BK;
#pragma unroll
                for (int p = 0; p < Cfg::PANELS; ++p) {
                    const int kp = kp0 + 8 * p;
                    if (rows_live >= Cfg::BM) {
                        const int rel = kp - i0 - FLAST - 7;
                        const int lo = rel <= 0 ? 0 : (rel + FSTEP - 1) / FSTEP;
                        if (lo == 0)
                            ML::mma_panel(acc, sA, sB, p, wm, wn, g, t);
                        else
                            ML::template mma_panel_jump<false>(acc, sA, sB, p, wm, wn, g, t, lo);
                    } else {
                        const int rel = kp - i0 - fbase - 7;
                        const int lo = rel <= 0 ? 0 : (rel + FSTEP - 1) / FSTEP;
                        if (lo < hi) ML::mma_panel_range(acc, sA, sB, p, wm, wn, g, t, lo, hi);
                    }
                }
            }
            // leave the stage; the last warp out refills it with item + STAGES
            __syncwarp();
            if (lane == 0) {
                __threadfence_block();
                const int before = atomicSub(&left[stage], 1);
                if (before == 1) {
                    left[stage] = NWARPS;  // nobody touches it again before the refill has landed
                    __threadfence_block();
                    if (lu < tiles_run) {
                        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                        issue(lu, lb, lkt, stage);
                    }
                }
                advance(lu, lb, lkt);
            }
            if (++stage == STAGES) {
                stage = 0;
                parity ^= 1u;
            }
        }
        if (STORE) {  // VT[j0 + col][i0 + row] = acc (rows of a partial first block beyond rows_live belong to block 1)
#pragma unroll
            for (int mf = 0; mf < Cfg::MF; ++mf) {
                const int r = Cfg::frag_row(wm, mf) + g;
                if (r >= rows_live) continue;
#pragma unroll
                for (int nf = 0; nf < Cfg::NF; ++nf) {
                    const int64_t col = j0 + wn * Cfg::WTN + nf * 8 + 2 * t;
                    vt_out[col * ldv + i0 + r] = acc[mf][nf][0];
                    vt_out[(col + 1) * ldv + i0 + r] = acc[mf][nf][1];
                }
            }
        } else {
#pragma unroll
            for (int mf = 0; mf < Cfg::MF; ++mf)
#pragma unroll
                for (int nf = 0; nf < Cfg::NF; ++nf) {
                    colsum[nf][0] = fma(acc[mf][nf][0], acc[mf][nf][0], colsum[nf][0]);
                    colsum[nf][1] = fma(acc[mf][nf][1], acc[mf][nf][1], colsum[nf][1]);
                }
        }
    }  // row blocks of this CTA
    if (STORE) {  // identity rows of the padded Linv meet zero columns of K*: V is zero there
        const int tail = npad - n8;
        for (int idx = threadIdx.x; idx < tail * Cfg::BN; idx += Cfg::THREADS) {
            const int c = idx / tail, r = idx - c * tail;
            vt_out[(int64_t)(j0 + c) * ldv + n8 + r] = 0.0;
        }
    } else {
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
        if (shared_mode) {  // deposit this member's partial sums; the last member of the tile finishes it
            for (int c = threadIdx.x; c < Cfg::BN; c += Cfg::THREADS) {
                double s = 0.0;
#pragma unroll
                for (int w = 0; w < Cfg::WM; ++w) s += red[w][c];
                partial[member * pstride + j0 + c] = s;
                __threadfence();
            }
            __syncthreads();
            if (threadIdx.x == 0) {
                const int tile = group + u * G;
                const int prev = atomicAdd(&counters[tile], 1);
                last_flag = prev == S - 1;
                if (prev == S - 1) counters[tile] = 0;  // ready for the next launch
            }
            __syncthreads();
            if (!last_flag) {
#pragma unroll
                for (int nf = 0; nf < Cfg::NF; ++nf) colsum[nf][0] = colsum[nf][1] = 0.0;
                continue;
            }
            __threadfence();
        }
        for (int c = threadIdx.x; c < Cfg::BN; c += Cfg::THREADS) {
            double s = 0.0;
            if (shared_mode) {
                for (int m = 0; m < S; ++m) s += __ldcg(partial + m * pstride + j0 + c);
            } else {
#pragma unroll
                for (int w = 0; w < Cfg::WM; ++w) s += red[w][c];
            }
            const int64_t q = q_begin + j0 + c;
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
#pragma unroll
        for (int nf = 0; nf < Cfg::NF; ++nf) colsum[nf][0] = colsum[nf][1] = 0.0;
    }
    }  // tiles of this CTA
}

// ---- persistent TMA-fed NT GEMM ------------------------------------------------------------------------------------------
// gemm_nt_kernel (gemm_nt.cuh) with the operand feed of trmm_sumsq_tma_kernel: 128 x 128 tiles, one CTA per SM that walks the
// tiles t = blockIdx.x, blockIdx.x + gridDim.x, ... of the (batch, tile row, tile column) grid; the (tile, k-tile) items of
// a CTA form ONE pipeline, so the first k-tiles of the next tile are in flight while the warps store the current one.  Same
// arguments, flags, dual store and addend as gemm_nt_kernel; per element the same sum in the same order.  Used for the
// products whose tile count can fill the machine (launch_gemm_nt); the 64 x 64 cp.async kernel keeps the small ones.
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
            smem_u32(dst)),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}

struct TileWalk {  // position in a CTA's tile sequence
    int t;         // tile index in the (z, ti, tj) enumeration; >= total: finished
    int z, i0, j0, kb, ktiles;
};

template <class Cfg>
__device__ __forceinline__ void tile_seek(TileWalk& w, const GemmArgs& p, int total, int tiles_m, int tiles_n,
                                          bool skip_empty) {
    const int per = tiles_m * tiles_n;
    while (w.t < total) {
        const int z = w.t / per, r = w.t - z * per, tj = r % tiles_n;
        int ti = r / tiles_n;
        // longest k ranges first, so that the tail of the static round-robin consists of the cheapest tiles: a lower-
        // triangular A has its longest rows at the bottom (an upper-triangular one at the top: natural order)
        if (p.flags & GEMM_A_LOWER) ti = tiles_m - 1 - ti;
        const int i0 = ti * Cfg::BM, j0 = tj * Cfg::BN;
        const bool live = !((p.flags & GEMM_C_LOWER) && j0 > i0 + Cfg::BM - 1) && !(p.gate && p.gate[z] == 0);
        if (live) {
            int kb, ke;
            tile_k_range(p.flags, i0, j0, Cfg::BM, Cfg::BN, Cfg::BK, p.K, kb, ke);
            const int kts = (ke - kb) / Cfg::BK;
            if (kts > 0 || !skip_empty) {
                w.z = z; w.i0 = i0; w.j0 = j0; w.kb = kb; w.ktiles = kts;
                return;
            }
        }
        w.t += gridDim.x;
    }
}

template <class Cfg, int MODE>
__global__ void __launch_bounds__(Cfg::THREADS, 1)
    gemm_nt_tma_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, GemmArgs p,
                       int tiles_m, int tiles_n, int total) {
    using ML = Mainloop<Cfg>;
    constexpr int STAGES = Cfg::STAGES, NWARPS = Cfg::THREADS / 32;
    constexpr uint32_t STAGE_BYTES = Cfg::STAGE_DOUBLES * 8;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* smem = reinterpret_cast<double*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~(uintptr_t)127);
    __shared__ __align__(8) uint64_t full[STAGES];
    __shared__ int left[STAGES];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int wm = warp / Cfg::WN, wn = warp % Cfg::WN;

    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full[s], 1);
            left[s] = NWARPS;
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    // look-ahead walk over the items (tile, k-tile) that are STAGES ahead of the one being consumed
    TileWalk la{(int)blockIdx.x, 0, 0, 0, 0, 0};
    int lkt = 0;
    tile_seek<Cfg>(la, p, total, tiles_m, tiles_n, true);
    auto issue = [&](int stage) {  // item (la, lkt) -> stage
        double* sA = smem + stage * Cfg::STAGE_DOUBLES;
        double* sB = sA + Cfg::BM * Cfg::BK;
        mbar_expect_tx(&full[stage], STAGE_BYTES);
        const int k0 = la.kb + lkt * Cfg::BK;
#pragma unroll
        for (int pn = 0; pn < Cfg::PANELS; ++pn) {
            tma_load_3d(sA + pn * Cfg::BM * 8, &mapA, &full[stage], k0 + 8 * pn, la.i0, la.z);
            tma_load_3d(sB + pn * Cfg::BN * 8, &mapB, &full[stage], k0 + 8 * pn, la.j0, la.z);
        }
    };
    auto la_advance = [&]() {
        if (la.t >= total) return;
        if (++lkt == la.ktiles) {
            lkt = 0;
            la.t += gridDim.x;
            tile_seek<Cfg>(la, p, total, tiles_m, tiles_n, true);
        }
    };
    for (int s = 0; s < STAGES; ++s) {
        if (threadIdx.x == 0 && la.t < total) issue(s);
        la_advance();
    }

    constexpr int FSTEP = Cfg::ILV ? 8 * Cfg::WM : 8;
    constexpr int FLAST = Cfg::frag_row(Cfg::WM - 1, 0);
    int stage = 0;
    uint32_t parity = 0;
    TileWalk cur{(int)blockIdx.x, 0, 0, 0, 0, 0};
    for (tile_seek<Cfg>(cur, p, total, tiles_m, tiles_n, false); cur.t < total;
         cur.t += gridDim.x, tile_seek<Cfg>(cur, p, total, tiles_m, tiles_n, false)) {
        const int i0 = cur.i0, j0 = cur.j0, ktiles = cur.ktiles;
        double acc[Cfg::MF][Cfg::NF][2];
#pragma unroll
        for (int mf = 0; mf < Cfg::MF; ++mf)
#pragma unroll
            for (int nf = 0; nf < Cfg::NF; ++nf) acc[mf][nf][0] = acc[mf][nf][1] = 0.0;

        // which k-tiles carry dead fragments (Mainloop::run): TRI_UPPER the first kt_tri, TRI_LOWER those from kt_full on
        int kt_tri = 0, kt_full = ktiles;
        if (MODE == TRI_UPPER) {
            kt_tri = (i0 + (Cfg::MF - 1) * FSTEP - 7 - cur.kb + Cfg::BK - 1) / Cfg::BK;
            kt_tri = kt_tri < 0 ? 0 : (kt_tri > ktiles ? ktiles : kt_tri);
        }
        if (MODE == TRI_LOWER) {
            const int num = i0 + FLAST + 7 - 8 * (Cfg::PANELS - 1) - cur.kb;
            kt_full = num < 0 ? 0 : num / Cfg::BK + 1;
            kt_full = kt_full > ktiles ? ktiles : kt_full;
        }
        for (int kt = 0; kt < ktiles; ++kt) {
            mbar_wait(&full[stage], parity);
            const double* sA = smem + stage * Cfg::STAGE_DOUBLES;
            const double* sB = sA + Cfg::BM * Cfg::BK;
            const int kp0 = cur.kb + kt * Cfg::BK;
            if (MODE == TRI_UPPER && kt < kt_tri) {
#pragma unroll
                for (int pn = 0; pn < Cfg::PANELS; ++pn) {
                    const int rel = kp0 + 8 * pn + 7 - i0;  // fragment live iff mf * FSTEP <= rel (first warp row)
                    const int up = rel < 0 ? 0 : rel / FSTEP + 1;
                    if (up >= Cfg::MF)
                        ML::mma_panel(acc, sA, sB, pn, wm, wn, g, t);
                    else
                        ML::template mma_panel_jump<true>(acc, sA, sB, pn, wm, wn, g, t, up);
                }
            } else if (MODE == TRI_LOWER && kt >= kt_full) {
#pragma unroll
                for (int pn = 0; pn < Cfg::PANELS; ++pn) {
                    const int rel = kp0 + 8 * pn - i0 - FLAST - 7;
                    const int lo = rel <= 0 ? 0 : (rel + FSTEP - 1) / FSTEP;
                    if (lo == 0)
                        ML::mma_panel(acc, sA, sB, pn, wm, wn, g, t);
                    else
                        ML::template mma_panel_jump<false>(acc, sA, sB, pn, wm, wn, g, t, lo);
                }
            } else {
#pragma unroll
                for (int pn = 0; pn < Cfg::PANELS; ++pn) ML::mma_panel(acc, sA, sB, pn, wm, wn, g, t);
            }
            __syncwarp();
            if (lane == 0) {
                __threadfence_block();
                const int before = atomicSub(&left[stage], 1);
                if (before == 1) {
                    left[stage] = NWARPS;
                    __threadfence_block();
                    if (la.t < total) {
                        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                        issue(stage);
                    }
                }
            }
            la_advance();
            if (++stage == STAGES) {
                stage = 0;
                parity ^= 1u;
            }
        }

        // epilogue of gemm_nt_kernel
        const int64_t z = cur.z;
        double* C = p.C ? p.C + z * p.strideC : nullptr;
        double* Ct = p.Ct ? p.Ct + z * p.strideCt : nullptr;
        const double* D = p.D ? p.D + z * p.strideD : nullptr;
#pragma unroll
        for (int mf = 0; mf < Cfg::MF; ++mf) {
            const int row = i0 + Cfg::frag_row(wm, mf) + g;
            if (row >= p.M) continue;
#pragma unroll
            for (int nf = 0; nf < Cfg::NF; ++nf) {
                const int col = j0 + wn * Cfg::WTN + nf * 8 + 2 * t;
                if (col >= p.N) continue;
                double v0 = p.alpha * acc[mf][nf][0], v1 = p.alpha * acc[mf][nf][1];
                if (D) {
                    const double2 old = *reinterpret_cast<const double2*>(D + (int64_t)row * p.ldd + col);
                    v0 += old.x;
                    v1 += old.y;
                }
                if (C) *reinterpret_cast<double2*>(C + (int64_t)row * p.ldc + col) = make_double2(v0, v1);
                if (Ct) {
                    Ct[(int64_t)col * p.ldct + row] = v0;
                    Ct[(int64_t)(col + 1) * p.ldct + row] = v1;
                }
            }
        }
    }
}

}  // namespace bobe
