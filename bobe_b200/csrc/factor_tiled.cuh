// Tile-column Cholesky with look-ahead + level-parallel triangular inverse, batched over hyper-parameter sets (grid.z).
//
// Replaces jnp.linalg.cholesky + cho_solve + solve_triangular at BOBE/gp.py:175-176,259-260,549-550 and supplies the
// L^-1 the predictive variance (BOBE/gp.py:462,484) and the fantasy variance (:571) are built on.
//
// Phase 1 -- factor.  The matrix is cut into 128-wide tile columns j = 0 .. T-1 (the last one may be 64 wide) that are
// grouped into outer panels of PW tile columns.  Per tile column, on the CRITICAL stream:
//     [update]  C_j = A[o_j:, j] - L[o_j:, k0:o_j] L[j, k0:o_j]^T     (the part of the k range not yet applied, see below)
//     [leaf]    (L_jj, X_jj = L_jj^-1) = chol_inv(C_j[0:128])          one CTA per matrix, shared memory (leaf.cuh)
//     [panel]   L[o_j+128:, j] = C_j[128:] X_jj^T                      (+ the gated correction step for ill-conditioned K)
// and on the BULK stream, overlapping the leaf of the next column(s):
//     inside a panel (left-looking):  column j+2 gets the products of the columns [panel start, j] as soon as column j
//         is final ("a1"); the critical stream then only applies the single last tile column ("a2", k = 128);
//     at the end of a panel (right-looking):  every column behind the next one gets the whole panel (k = 128 PW) in one
//         trailing update; the next column itself is updated by the critical stream.
// PW = 1 is the classic right-looking algorithm (largest parallelism, k = 128 products), PW = T the left-looking one
// (long k loops, every tile written once); without a bulk stream the SAME operations run in program order on one
// stream (throughput mode for many restarts, where sub-batches on different streams overlap each other instead), so
// that a matrix is factorised by bitwise the same arithmetic whatever batch it is part of.
// The critical path of one matrix is T x (small update + leaf + panel) instead of the ~230 dependent launches and 16
// serial 128-leaves of the recursive scheme this replaces; the O(n^3) products run beside it.
//
// Phase 2 -- inverse.  X = L^-1 by recursive doubling over the tiles: the nodes of one level of the tree are independent,
// so each level is TWO launches for all its nodes and all matrices (GemmArgs node mode):
//     P^T = U11 L21^T   then   X21 = -X22 P      (U = X^T is kept alongside X: every product stays "NT")
//
// A non-PD matrix yields NaN outputs and info = 1 for that batch entry only, never an error.
// (included by factor.cu after the helper kernels it shares with the vector solves)
#pragma once
#include <mutex>
#include <vector>

namespace bobe {

// ---- internal side streams + dependency events, one pool per device ---------------------------------------------------
// They carry no state between calls: every call forks them from the caller's stream and joins them back before returning.
StreamPool* stream_pool() {
    static std::mutex mu;
    static StreamPool* pools[64] = {nullptr};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) {
        set_error("stream_pool: cudaGetDevice failed");
        return nullptr;
    }
    std::lock_guard<std::mutex> lock(mu);
    if (!pools[dev]) {
        StreamPool* p = new StreamPool();
        bool ok = cudaEventCreateWithFlags(&p->fork, cudaEventDisableTiming) == cudaSuccess;
        for (int i = 0; i < POOL_STREAMS && ok; ++i)
            ok = cudaStreamCreateWithFlags(&p->streams[i], cudaStreamNonBlocking) == cudaSuccess &&
                 cudaEventCreateWithFlags(&p->join[i], cudaEventDisableTiming) == cudaSuccess;
        if (!ok) {
            set_error("stream_pool: could not create internal streams: %s", cudaGetErrorString(cudaGetLastError()));
            delete p;
            return nullptr;
        }
        pools[dev] = p;
    }
    return pools[dev];
}

cudaEvent_t StreamPool::event(int lane, int idx) {
    std::vector<cudaEvent_t>& r = ring[lane];
    while ((int)r.size() <= idx) {
        cudaEvent_t e = nullptr;
        if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) return nullptr;
        r.push_back(e);
    }
    return r[idx];
}

namespace {

constexpr int TW = 128;  // tile-column width

struct Tiled {
    const FactorExec& ex;
    const FactorBuffers& fb;
    int npad, batch;
    int64_t mstride, qstride;
    int32_t rc = BOBE_OK;
    int T;
    bool two;                        // look-ahead on a second stream
    std::vector<int> last_writer;    // per tile column: index of the bulk event of the last bulk op that wrote it (-1: none)
    int n_bulk = 0;                  // bulk ops issued so far (their events are ring[lane][T + idx])
    int crit_waited = -1;            // highest bulk event index the critical stream already waited for
    int bulk_waited = -1;            // highest column event the bulk stream already waited for

    int off(int j) const { return j * TW; }
    int width(int j) const { return npad - off(j) < TW ? npad - off(j) : TW; }

    GemmArgs base() const {
        GemmArgs g{};
        g.lda = g.ldb = g.ldc = g.ldct = g.ldd = npad;
        g.strideA = g.strideB = g.strideC = g.strideCt = g.strideD = mstride;
        g.alpha = 1.0;
        return g;
    }
    void gemm(cudaStream_t st, const GemmArgs& a) {
        if (rc == BOBE_OK) rc = launch_gemm_nt(st, a, batch);
    }
    bool ok(cudaError_t e, const char* what) {
        if (e != cudaSuccess && rc == BOBE_OK) {
            set_error("factor_tiled: %s: %s", what, cudaGetErrorString(e));
            rc = BOBE_E_CUDA;
        }
        return e == cudaSuccess;
    }

    // KB[oc:, oc:oc+ncols] -= L[oc:, k0:k1] L[oc:oc+ncols, k0:k1]^T   (tiles above the diagonal skipped)
    void update(cudaStream_t st, int oc, int ncols, int k0, int k1) {
        if (k1 <= k0 || ncols <= 0) return;
        GemmArgs g = base();
        g.A = fb.L + (int64_t)oc * npad + k0;
        g.Bt = g.A;
        g.C = fb.KB + (int64_t)oc * npad + oc;
        g.D = g.C;
        g.M = npad - oc; g.N = ncols; g.K = k1 - k0; g.alpha = -1.0; g.flags = GEMM_C_LOWER;
        gemm(st, g);
    }

    // column events of the critical stream: ring[lane][j];  bulk events: ring[lane][T + idx]
    void crit_done(int j) {
        if (two) ok(cudaEventRecord(ex.pool->event(ex.lane, j), ex.crit), "event record");
    }
    void bulk_needs_col(int j) {  // the next bulk op reads L columns <= j
        if (two && j > bulk_waited) {
            ok(cudaStreamWaitEvent(ex.bulk, ex.pool->event(ex.lane, j), 0), "event wait");
            bulk_waited = j;
        }
    }
    void bulk_wrote(int c0, int c1) {  // the bulk op just issued wrote the tile columns [c0, c1)
        if (!two) return;
        ok(cudaEventRecord(ex.pool->event(ex.lane, T + n_bulk), ex.bulk), "event record");
        for (int c = c0; c < c1 && c < T; ++c) last_writer[c] = n_bulk;
        ++n_bulk;
    }
    void crit_needs_col(int j) {  // the critical stream is about to touch KB column j
        if (two && last_writer[j] > crit_waited) {
            ok(cudaStreamWaitEvent(ex.crit, ex.pool->event(ex.lane, T + last_writer[j]), 0), "event wait");
            crit_waited = last_writer[j];
        }
    }

    void leaf(int j) {
        if (rc != BOBE_OK) return;
        // BOBE_LEAF: 2 (default) recursive 32-base elimination, 1 four columns per barrier on the 64-block, 0 one column
        static const int mode = (int)env_int("BOBE_LEAF", 2);
        LeafIO io{fb.KB, fb.L, nullptr, fb.Linv, fb.U, fb.diag, fb.dstat, fb.gate, npad, off(j), mode == 0 ? 0 : 1};
        auto go = [&](auto kernel, int smem) {
            if ((rc = ensure_smem_fn(kernel, smem)) != BOBE_OK) return;
            launch_pdl(kernel, dim3(1, 1, batch), dim3(LEAF_THREADS), smem, ex.crit, io);
            rc = check_launch("tile leaf kernel");
        };
        if (width(j) == 64) {
            if (mode == 2) go(tile_leaf64_kernel<true>, LEAF64_SMEM); else go(tile_leaf64_kernel<false>, LEAF64_SMEM);
        } else {
            if (mode == 2) go(tile_leaf128_kernel<true>, LEAF128_SMEM); else go(tile_leaf128_kernel<false>, LEAF128_SMEM);
        }
    }

    // L[o+w:, o:o+w] = KB[o+w:, o:o+w] X_jj^T, then for gated (ill-conditioned) matrices one correction step
    //   R = C - L_col L_jj^T,  L_col += R X_jj^T
    // (a product with an explicit inverse is not backward stable; measured in profiles/r01/accuracy_probe.txt)
    void panel(int j) {
        const int o = off(j), w = width(j), below = npad - o - w;
        if (below <= 0) return;
        const int64_t col = (int64_t)(o + w) * npad + o, dg = (int64_t)o * npad + o;
        GemmArgs g = base();
        g.A = fb.KB + col; g.Bt = fb.Linv + dg; g.C = fb.L + col;
        g.M = below; g.N = w; g.K = w; g.flags = GEMM_B_LOWER;
        gemm(ex.crit, g);
        g = base();
        g.A = fb.L + col; g.Bt = fb.L + dg; g.D = fb.KB + col; g.C = fb.Q; g.ldc = w; g.strideC = qstride;
        g.M = below; g.N = w; g.K = w; g.alpha = -1.0; g.flags = GEMM_B_LOWER; g.gate = fb.gate;
        gemm(ex.crit, g);
        g = base();
        g.A = fb.Q; g.lda = w; g.strideA = qstride; g.Bt = fb.Linv + dg; g.D = fb.L + col; g.C = fb.L + col;
        g.M = below; g.N = w; g.K = w; g.flags = GEMM_B_LOWER; g.gate = fb.gate;
        gemm(ex.crit, g);
    }

    void phase1() {
        const int PW = ex.pw < 1 ? 1 : (ex.pw > T ? T : ex.pw);
        cudaStream_t bulk = two ? ex.bulk : ex.crit;
        for (int j = 0; j < T && rc == BOBE_OK; ++j) {
            const int s = (j / PW) * PW;  // first tile column of this outer panel
            crit_needs_col(j);
            if (j > 0) {
                if (j == s) {
                    update(ex.crit, off(j), width(j), off(s - PW), off(s));  // the previous panel, for this column only
                } else {
                    update(ex.crit, off(j), width(j), off(j - 1), off(j));   // a2: the column just finished
                }
            }
            leaf(j);
            panel(j);
            crit_done(j);
            if (j + 1 >= T) break;
            const bool panel_end = (j + 1) % PW == 0;
            if (panel_end) {
                // trailing update of everything behind the next column with this panel's k range
                if (j + 2 < T) {
                    bulk_needs_col(j);
                    update(bulk, off(j + 2), npad - off(j + 2), off(s), off(j + 1));
                    bulk_wrote(j + 2, T);
                }
            } else if (j + 2 < T && j + 2 < s + PW) {
                // a1 of column j+2: the in-panel columns s .. j (column j+1 follows on the critical stream as a2)
                bulk_needs_col(j);
                update(bulk, off(j + 2), width(j + 2), off(s), off(j + 1));
                bulk_wrote(j + 2, j + 3);
            }
        }
        if (two && n_bulk > 0 && crit_waited < n_bulk - 1)  // join
            ok(cudaStreamWaitEvent(ex.crit, ex.pool->event(ex.lane, T + n_bulk - 1), 0), "event wait");
    }

    void phase2() {
        for (int m = TW; m < npad && rc == BOBE_OK; m *= 2) {
            const int nodes = (npad + 2 * m - 1) / (2 * m);
            GemmArgs g = base();
            g.A = fb.U; g.Bt = fb.L; g.C = fb.Q; g.strideC = qstride;
            g.M = m; g.N = m; g.K = m; g.flags = GEMM_A_UPPER;
            g.node_count = nodes; g.node_m = m; g.node_total = npad; g.node_kind = 1;
            gemm(ex.crit, g);
            g = base();
            g.A = fb.Linv; g.Bt = fb.Q; g.strideB = qstride; g.C = fb.Linv; g.Ct = fb.U;
            g.M = m; g.N = m; g.K = m; g.alpha = -1.0; g.flags = GEMM_A_LOWER;
            g.node_count = nodes; g.node_m = m; g.node_total = npad; g.node_kind = 2;
            gemm(ex.crit, g);
        }
    }
};

}  // namespace

int32_t factor_tiled(const FactorExec& ex, const FactorBuffers& fb, int npad, int batch) {
    if (npad % NB) {
        set_error("factor: npad=%d not a multiple of %d", npad, NB);
        return BOBE_E_ARG;
    }
    init_stat_kernel<<<(batch + 127) / 128, 128, 0, ex.crit>>>(fb.dstat, fb.gate, batch, fb.force_refine);
    if (int32_t rc = check_launch("init_stat_kernel")) return rc;
    if (fb.zero_band == 0 && npad > NB) {  // buffers handed to the caller: the whole other triangle must read as zero
        zero_other_triangle_kernel<<<dim3(npad / NB, npad / NB, batch), 256, 0, ex.crit>>>(fb.L, nullptr, fb.Linv, fb.U, npad, 0);
        if (int32_t rc = check_launch("zero_other_triangle_kernel")) return rc;
    }
    Tiled t{ex, fb, npad, batch, (int64_t)npad * npad, factor_q_elems(npad)};
    t.T = (npad + TW - 1) / TW;
    t.two = ex.bulk != nullptr && ex.pool != nullptr && t.T > 2;
    t.last_writer.assign(t.T, -1);
    if (t.two) {  // make sure every event exists before the first record (creation failure -> single-stream fallback)
        if (!ex.pool->event(ex.lane, 2 * t.T + 2)) t.two = false;
    }
    t.phase1();
    if (t.rc == BOBE_OK) t.phase2();
    return t.rc;
}

}  // namespace bobe
