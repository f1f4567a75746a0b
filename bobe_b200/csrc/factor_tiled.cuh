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
// Phase 2 -- inverse.  X = L^-1 by recursive doubling over the tiles; per node of the tree (left half X11, right half X22)
//     P^T = U11 L21^T   then   X21 = -X22 P      (U = X^T is kept alongside X: every product stays "NT")
// Single-stream mode: the nodes of one level are independent, so each level is TWO launches for all its nodes and all
// matrices (GemmArgs node mode), after phase 1.  Look-ahead mode: every product is issued on a third stream as soon as its
// inputs exist -- P^T of a node when its left half is complete (half-way through the node's tile columns), X21 when its
// right half is -- so that phase 2 runs in the shadow of the phase-1 chain and only the last X21 of each level (the
// right spine of the tree) is left when the last leaf finishes.  Same products, same results either way.
//
// A non-PD matrix yields NaN outputs and info = 1 for that batch entry only, never an error.
// (included by factor.cu after the helper kernels it shares with the vector solves)
#pragma once
#include <mutex>
#include <vector>

#include <cuda.h>
#include <cudaTypedefs.h>

namespace bobe {

// ---- internal side streams + dependency events, one pool per device ---------------------------------------------------
// They carry no state between calls: every call forks them from the caller's stream and joins them back before returning.
// Two green contexts: BOBE_GREEN_SMS SMs for the chains, the rest for everything else (experiment knob, default off).  Entry points are
// fetched through the runtime (no link dependency on libcuda); any failure leaves pool->green false.
static void make_green_streams(StreamPool* p, int dev, int prio_lo, int prio_hi) {
    for (int i = 0; i < POOL_STREAMS; ++i) p->gstreams[i] = nullptr;
    const int64_t want = env_int("BOBE_GREEN_SMS", 0);  // off by default, see profiles/r02/README.md
    if (want <= 0) return;
    auto entry = [](const char* name, unsigned ver) -> void* {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPointByVersion(name, &fn, ver, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
            return nullptr;
        return fn;
    };
    auto get_res = (PFN_cuDeviceGetDevResource_v12040)entry("cuDeviceGetDevResource", 12040);
    auto split = (PFN_cuDevSmResourceSplitByCount_v12040)entry("cuDevSmResourceSplitByCount", 12040);
    auto gen_desc = (PFN_cuDevResourceGenerateDesc_v12040)entry("cuDevResourceGenerateDesc", 12040);
    auto ctx_create = (PFN_cuGreenCtxCreate_v12040)entry("cuGreenCtxCreate", 12040);
    auto stream_create = (PFN_cuGreenCtxStreamCreate_v12050)entry("cuGreenCtxStreamCreate", 12050);
    if (!get_res || !split || !gen_desc || !ctx_create || !stream_create) return;
    CUdevResource all, chain, rest;
    unsigned groups = 1;
    if (get_res(dev, &all, CU_DEV_RESOURCE_TYPE_SM) != CUDA_SUCCESS || all.sm.smCount < 2 * (unsigned)want) return;
    if (split(&chain, &groups, &all, &rest, 0, (unsigned)want) != CUDA_SUCCESS || groups != 1 || rest.sm.smCount == 0) return;
    CUdevResourceDesc d_chain, d_rest;
    CUgreenCtx g_chain, g_rest;
    if (gen_desc(&d_chain, &chain, 1) != CUDA_SUCCESS || gen_desc(&d_rest, &rest, 1) != CUDA_SUCCESS) return;
    if (ctx_create(&g_chain, d_chain, dev, CU_GREEN_CTX_DEFAULT_STREAM) != CUDA_SUCCESS ||
        ctx_create(&g_rest, d_rest, dev, CU_GREEN_CTX_DEFAULT_STREAM) != CUDA_SUCCESS)
        return;  // (contexts live as long as the process, like the pool)
    for (int i = 0; i < POOL_STREAMS; ++i) {
        const int r = i % POOL_LANE_STREAMS;
        const int prio = r == 0 ? prio_hi : (r == 1 ? prio_hi + (prio_lo - prio_hi) / 3 : (r == 2 ? prio_hi + 2 * (prio_lo - prio_hi) / 3 : prio_lo));
        CUstream st = nullptr;
        if (stream_create(&st, r == 0 ? g_chain : g_rest, CU_STREAM_NON_BLOCKING, prio) != CUDA_SUCCESS) return;
        p->gstreams[i] = (cudaStream_t)st;
    }
    p->green = true;
    p->green_sms = (int)chain.sm.smCount;
}

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
        // four streams per lane, in falling priority: the dependent chain, the rest of the current column, the trailing
        // products, the inverse tree -- when an SM slot frees up, a waiting CTA of the chain goes first
        int lo = 0, hi = 0;
        cudaDeviceGetStreamPriorityRange(&lo, &hi);  // lo: least priority (numerically largest), hi: greatest
        for (int i = 0; i < POOL_STREAMS && ok; ++i) {
            const int r = i % POOL_LANE_STREAMS;  // 0 chain, 1 rest of column, 2 trailing products, 3 inverse tree
            const int prio = r == 0 ? hi : (r == 1 ? hi + (lo - hi) / 3 : (r == 2 ? hi + 2 * (lo - hi) / 3 : lo));
            ok = cudaStreamCreateWithPriority(&p->streams[i], cudaStreamNonBlocking, prio) == cudaSuccess &&
                 cudaEventCreateWithFlags(&p->join[i], cudaEventDisableTiming) == cudaSuccess;
        }
        if (!ok) {
            set_error("stream_pool: could not create internal streams: %s", cudaGetErrorString(cudaGetLastError()));
            delete p;
            return nullptr;
        }
        make_green_streams(p, dev, lo, hi);
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
inline int off_last(int npad) { return ((npad + TW - 1) / TW - 1) * TW; }  // first column of the last tile column

// L[r][c] = Linv[r][c] = 0 and U[c][r] = 0 for the padded rows r in [nl, npad) and the columns c < c_end left of the last
// diagonal tile (which the leaf writes in full, identity part included).  grid = (ceil(c_end / 256), npad - nl, batch).
__global__ void __launch_bounds__(256) pad_rows_zero_kernel(double* L, double* Linv, double* U, int npad, int nl, int c_end) {
    const int c = blockIdx.x * 256 + threadIdx.x, r = nl + blockIdx.y;
    if (c >= c_end) return;
    const int64_t z = (int64_t)blockIdx.z * npad * npad;
    L[z + (int64_t)r * npad + c] = 0.0;
    Linv[z + (int64_t)r * npad + c] = 0.0;
    U[z + (int64_t)c * npad + r] = 0.0;
}

struct Tiled {
    const FactorExec& ex;
    const FactorBuffers& fb;
    int npad, batch;
    int64_t mstride, qstride;
    int32_t rc = BOBE_OK;
    int T;
    int nl = 0;                      // live rows / columns (fb.n_live): products stop there, the padding is filled, not computed
    bool two;                        // look-ahead mode: chain / rest-of-column / trailing products / inverse on four streams
    int inv_waited = -1;             // highest column the inverse stream already waited for
    bool inv_used = false;
    std::vector<int> last_writer;    // per tile column: index of the last bulk op that wrote it (-1: none)
    int n_bulk = 0;                  // bulk ops issued so far
    int crit_waited = -1, mid_waited = -1;  // highest bulk op the critical / mid stream already waited for
    int bulk_waited = -1;            // highest column the bulk stream already waited for

    int off(int j) const { return j * TW; }
    int width(int j) const { return npad - off(j) < TW ? npad - off(j) : TW; }
    int live(int o) const { return nl > o ? nl - o : 0; }                      // live rows / columns from offset o on
    int wlive(int j) const { return live(off(j)) < width(j) ? live(off(j)) : width(j); }  // live width of tile column j

    // events of lane ex.lane:  [j] top tile of panel j done (critical stream);  [T + j] leaf j done;
    // [2T + j] column j complete (mid stream);  [3T + j] rows of tile j+1 of column j updated (mid stream);
    // [4T + i] bulk op i done (fewer than 3T of them);  [7T + 1] inverse stream done
    cudaEvent_t ev_top(int j) const { return ex.pool->event(ex.lane, j); }
    cudaEvent_t ev_leaf(int j) const { return ex.pool->event(ex.lane, T + j); }
    cudaEvent_t ev_col(int j) const { return ex.pool->event(ex.lane, 2 * T + j); }
    cudaEvent_t ev_u1(int j) const { return ex.pool->event(ex.lane, 3 * T + j); }
    cudaEvent_t ev_bulk(int i) const { return ex.pool->event(ex.lane, 4 * T + i); }
    cudaEvent_t ev_inv() const { return ex.pool->event(ex.lane, 7 * T + 1); }
    cudaStream_t s_mid() const { return two ? ex.mid : ex.crit; }
    cudaStream_t s_bulk() const { return two ? ex.bulk : ex.crit; }

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
    void record(cudaEvent_t e, cudaStream_t st) {
        if (two) ok(cudaEventRecord(e, st), "event record");
    }
    void wait(cudaStream_t st, cudaEvent_t e) {
        if (two) ok(cudaStreamWaitEvent(st, e, 0), "event wait");
    }

    // KB[r0 : r0+nrows, oc : oc+ncols] -= L[r0.., k0:k1] L[oc : oc+ncols, k0:k1]^T.  r0 == oc: the block starts on the
    // diagonal and tiles above it are skipped; r0 >= oc + ncols: a block strictly below the diagonal.
    void update(cudaStream_t st, int r0, int nrows, int oc, int ncols, int k0, int k1) {
        if (k1 <= k0 || ncols <= 0 || nrows <= 0) return;
        GemmArgs g = base();
        g.A = fb.L + (int64_t)r0 * npad + k0;
        g.Bt = fb.L + (int64_t)oc * npad + k0;
        g.C = fb.KB + (int64_t)r0 * npad + oc;
        g.D = g.C;
        g.M = nrows; g.N = ncols; g.K = k1 - k0; g.alpha = -1.0; g.flags = (r0 == oc) ? GEMM_C_LOWER : 0;
        gemm(st, g);
    }

    void bulk_needs_col(int j) {  // the next bulk op reads L columns <= j
        if (two && j > bulk_waited) {
            wait(ex.bulk, ev_col(j));
            bulk_waited = j;
        }
    }
    void bulk_wrote(int c0, int c1) {  // the bulk op just issued wrote the tile columns [c0, c1)
        if (!two) return;
        record(ev_bulk(n_bulk), ex.bulk);
        for (int c = c0; c < c1 && c < T; ++c) last_writer[c] = n_bulk;
        ++n_bulk;
    }
    void needs_col(cudaStream_t st, int& waited, int j) {  // `st` is about to touch KB column j
        if (two && last_writer[j] > waited) {
            wait(st, ev_bulk(last_writer[j]));
            waited = last_writer[j];
        }
    }

    void leaf(int j) {
        if (rc != BOBE_OK) return;
        // BOBE_LEAF: 2 (default) recursive 32-base elimination, 1 four columns per barrier on the 64-block, 0 one column
        static const int mode = (int)env_int("BOBE_LEAF", 2);
        // products with 128-row tiles read whole diagonal tiles: then (and only then) the leaf also zeroes the 64-blocks
        // on the other side of the diagonal (the external path has zeroed the whole triangle beforehand anyway)
        static const bool wide_tiles = env_int("BOBE_TILE", 0) > 1 || env_int("BOBE_GEMM_TMA", 0) != 0;
        // gate rows: [0] latest state (initial value before leaf 0), [j + 1] snapshot after leaf j
        LeafIO io{fb.KB, fb.L, nullptr, fb.Linv, fb.U, fb.diag, fb.dstat, fb.gate + (int64_t)(j == 0 ? 0 : j) * batch, npad,
                  off(j), mode == 0 ? 0 : 1, fb.gate + (int64_t)(j + 1) * batch, fb.gate, (wide_tiles ? 1 : 0) | (two ? 0 : 2)};
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

    // L[r0 : r0+nrows, o:o+w] = KB[r0.., o:o+w] X_jj^T, then for gated (ill-conditioned) matrices one correction step
    //   R = C - L_col L_jj^T,  L_col += R X_jj^T
    // (a product with an explicit inverse is not backward stable; measured in profiles/r01/accuracy_probe.txt).
    // q: scratch of nrows x w doubles per matrix for R.
    void panel(cudaStream_t st, int j, int r0, int nrows, double* q) {
        if (nrows <= 0) return;
        const int o = off(j), w = width(j);
        const int64_t col = (int64_t)r0 * npad + o, dg = (int64_t)o * npad + o;
        const int* gate = fb.gate + (int64_t)(j + 1) * batch;  // the state right after leaf j (see leaf())
        GemmArgs g = base();
        g.A = fb.KB + col; g.Bt = fb.Linv + dg; g.C = fb.L + col;
        g.M = nrows; g.N = w; g.K = w; g.flags = GEMM_B_LOWER;
        gemm(st, g);
        g = base();
        g.A = fb.L + col; g.Bt = fb.L + dg; g.D = fb.KB + col; g.C = q; g.ldc = w; g.strideC = qstride;
        g.M = nrows; g.N = w; g.K = w; g.alpha = -1.0; g.flags = GEMM_B_LOWER; g.gate = gate;
        gemm(st, g);
        g = base();
        g.A = q; g.lda = w; g.strideA = qstride; g.Bt = fb.Linv + dg; g.D = fb.L + col; g.C = fb.L + col;
        g.M = nrows; g.N = w; g.K = w; g.flags = GEMM_B_LOWER; g.gate = gate;
        gemm(st, g);
    }

    // ---- inverse tree, one node at a time (look-ahead mode) ----------------------------------------------------
    // scratch of level m inside fb.Lt (not used as L^T by this scheme): levels are laid out one after the other
    double* pbuf(int m) const {
        int64_t o = 0;
        for (int mm = TW; mm < m; mm *= 2) o += (int64_t)mm * mm;
        return fb.Lt + o;
    }
    void node_p1(cudaStream_t st, int m, int o) {  // P^T = U11 L21^T
        const int m2 = live(o + m) < m ? live(o + m) : m;
        if (m2 <= 0) return;
        GemmArgs g = base();
        g.A = fb.U + (int64_t)o * npad + o; g.Bt = fb.L + (int64_t)(o + m) * npad + o; g.C = pbuf(m); g.ldc = m2;
        g.M = m; g.N = m2; g.K = m; g.flags = GEMM_A_UPPER;
        gemm(st, g);
    }
    void node_p2(cudaStream_t st, int m, int o) {  // X21 = -X22 P
        const int m2 = live(o + m) < m ? live(o + m) : m;
        if (m2 <= 0) return;
        GemmArgs g = base();
        g.A = fb.Linv + (int64_t)(o + m) * (npad + 1); g.Bt = pbuf(m); g.ldb = m2;
        g.C = fb.Linv + (int64_t)(o + m) * npad + o; g.Ct = fb.U + (int64_t)o * npad + (o + m);
        g.M = m2; g.N = m; g.K = m2; g.alpha = -1.0; g.flags = GEMM_A_LOWER;
        gemm(st, g);
    }
    // everything of the inverse tree that tile column j (just finished on the critical stream) makes possible
    void eager_inverse(int j) {
        const int done = j + 1;  // tile columns complete
        bool first = true;
        auto stream = [&]() {
            if (first) {
                first = false;
                inv_used = true;
                if (j > inv_waited) {
                    wait(ex.inv, ev_col(j));
                    inv_waited = j;
                }
            }
            return ex.inv;
        };
        // X21 of every node whose right half ends here (at a node boundary, or at the end of the matrix), small to large
        for (int m = TW; m < nl; m *= 2) {
            const int w = m / TW;  // tiles per half
            int t;
            if (done == T)
                t = (nl - 1) / (2 * m);
            else if (done % (2 * w) == 0)
                t = done / (2 * w) - 1;
            else
                continue;
            const int o = 2 * m * t;
            if (o + m < nl && o + m < done * TW) node_p2(stream(), m, o);
        }
        // P^T of the node whose left half ends here
        if (done < T) {
            int w = 1;
            while (done % (2 * w) == 0) w *= 2;  // largest power of two dividing `done`
            const int m = w * TW, t = (done / w - 1) / 2, o = 2 * m * t;
            if ((done / w) % 2 == 1 && o + m == done * TW && o + m < nl) node_p1(stream(), m, o);
        }
    }

    // One tile column.  Only ONE tile of each product is on the critical path: the next leaf needs tile (j+1, j+1) of the
    // working matrix, which needs the rows of tile j+1 of this column's panel ("top").  Everything below follows on the mid
    // stream while the next leaf runs -- for a batch of matrices those products are throughput-bound and would otherwise
    // double the length of the chain.  Per column:
    //   critical:  update of the diagonal tile;  leaf;  [after the mid stream's U1]  panel rows of tile j+1
    //   mid:       U1 = update of the rows of tile j+1 (first, and signalled: the critical stream's panel reads them),
    //              update of the rows below;  [after the leaf]  panel rows below tile j+1
    void step(int j, int PW) {
        const int s = (j / PW) * PW;  // first tile column of this outer panel
        const int o = off(j), w = width(j), wl = wlive(j), below = live(o + w);
        const int top = below < TW ? below : TW;  // (live) rows of tile j+1
        needs_col(ex.crit, crit_waited, j);
        if (below > 0) needs_col(s_mid(), mid_waited, j);
        if (j > 0) {
            // the products of this column not applied yet: the previous outer panel (j == s) or the previous column
            const int k0 = (j == s) ? off(s - PW) : off(j - 1), k1 = off(j);
            // diagonal tile: its operand rows L[tile j, k0:k1) are the top of panel j-1 (this stream) and, at the start
            // of an outer panel, lower rows of the columns before it (mid stream)
            if (j == s && j >= 2) wait(ex.crit, ev_col(j - 2));
            if (!two) {
                // single stream: the three row ranges in ONE launch (same arithmetic per element; fewer, larger launches)
                update(ex.crit, o, wl + below, o, wl, k0, k1);
            } else {
                update(ex.crit, o, wl, o, wl, k0, k1);
            }
            if (two && below > 0) {
                wait(s_mid(), ev_top(j - 1));  // the Bt operand L[tile j, k0:k1) includes the top of panel j-1
                update(s_mid(), o + w, top, o, wl, k0, k1);
                record(ev_u1(j), s_mid());
                update(s_mid(), o + w + top, below - top, o, wl, k0, k1);
            }
        }
        // (the critical stream's panel below reads rows the mid stream's U1 updated; that finished about when the diagonal
        // update above did, so the wait goes in FRONT of the leaf, where it costs nothing: behind it, it would break the
        // programmatic chaining leaf -> panel and expose a full launch latency on every step)
        if (below > 0 && j > 0) wait(ex.crit, ev_u1(j));
        leaf(j);
        record(ev_leaf(j), ex.crit);
        if (two) {  // U tile = (X tile)^T, off the critical path (single-stream mode: the leaf writes it itself)
            wait(ex.inv, ev_leaf(j));
            inv_used = true;
            if (rc == BOBE_OK) {
                launch_pdl(tile_transpose_kernel, dim3(w / 32, w / 32, batch), dim3(256), 0, ex.inv, (const double*)fb.Linv, fb.U, npad, o);
                rc = check_launch("tile_transpose_kernel");
            }
        }
        if (below > 0 && !two) {
            panel(ex.crit, j, o + w, below, fb.Q);  // single stream: all rows in one launch
        } else if (below > 0) {
            panel(ex.crit, j, o + w, top, fb.Q);                                   // rows of tile j+1
            wait(s_mid(), ev_leaf(j));
            panel(s_mid(), j, o + w + top, below - top, fb.Q + (int64_t)TW * TW);  // the rest
        }
        record(ev_top(j), ex.crit);
        if (two) {  // column j complete = both parts
            wait(s_mid(), ev_top(j));
            record(ev_col(j), s_mid());
            eager_inverse(j);
        }
        if (j + 1 >= T) return;
        const bool panel_end = (j + 1) % PW == 0;
        if (panel_end) {
            // trailing update of everything behind the next column with this panel's k range.  (Issuing the next panel's
            // columns as separate, earlier-finishing launches was measured and lost: for a batch of matrices the chain is
            // bound by the throughput of these products, and the narrow pieces run less efficiently than the square.)
            if (j + 2 < T) {
                bulk_needs_col(j);
                update(s_bulk(), off(j + 2), live(off(j + 2)), off(j + 2), live(off(j + 2)), off(s), off(j + 1));
                bulk_wrote(j + 2, T);
            }
        } else if (j + 2 < T && j + 2 < s + PW) {
            // a1 of column j+2: the in-panel columns s .. j (column j+1 follows as a2)
            bulk_needs_col(j);
            update(s_bulk(), off(j + 2), live(off(j + 2)), off(j + 2), wlive(j + 2), off(s), off(j + 1));
            bulk_wrote(j + 2, j + 3);
        }
    }

    void finish_phase1() {
        if (!two) return;
        if (n_bulk > 0 && crit_waited < n_bulk - 1) wait(ex.crit, ev_bulk(n_bulk - 1));  // join the bulk stream,
        wait(ex.crit, ev_col(T - 1));                                                    // the mid stream
        if (inv_used) {                                                                  // and the inverse stream
            record(ev_inv(), ex.inv);
            wait(ex.crit, ev_inv());
        }
    }

    void phase2() {
        if (two) return;  // done eagerly
        for (int m = TW; m < nl && rc == BOBE_OK; m *= 2) {
            const int nodes = (nl + 2 * m - 1) / (2 * m);
            GemmArgs g = base();
            g.A = fb.U; g.Bt = fb.L; g.C = fb.Q; g.strideC = qstride;
            g.M = m; g.N = m; g.K = m; g.flags = GEMM_A_UPPER;
            g.node_count = nodes; g.node_m = m; g.node_total = nl; g.node_kind = 1;
            gemm(ex.crit, g);
            g = base();
            g.A = fb.Linv; g.Bt = fb.Q; g.strideB = qstride; g.C = fb.Linv; g.Ct = fb.U;
            g.M = m; g.N = m; g.K = m; g.alpha = -1.0; g.flags = GEMM_A_LOWER;
            g.node_count = nodes; g.node_m = m; g.node_total = nl; g.node_kind = 2;
            gemm(ex.crit, g);
        }
    }
};

}  // namespace

// Stepwise interface: several sub-batches (each with its own streams) are advanced in lock step by the caller, so that the
// host enqueues step j of every chain before step j + 1 of any (a chain is ~100 launches: enqueued one after the other,
// the second chain would start half a millisecond late).
struct TiledFactor {
    FactorExec ex;
    FactorBuffers fb;
    Tiled t;
    int pw;
    TiledFactor(const FactorExec& e, const FactorBuffers& f, int npad, int batch)
        : ex(e), fb(f), t{ex, fb, npad, batch, (int64_t)npad * npad, factor_q_elems(npad)} {}
};

TiledFactor* tiled_begin(const FactorExec& ex, const FactorBuffers& fb, int npad, int batch, int32_t* rc_out) {
    *rc_out = BOBE_OK;
    if (npad % NB) {
        set_error("factor: npad=%d not a multiple of %d", npad, NB);
        *rc_out = BOBE_E_ARG;
        return nullptr;
    }
    const int T_ = (npad + TW - 1) / TW;
    if (ex.crit != ex.home) {  // the chain runs on a stream of its own: fork it from the caller's
        cudaEvent_t e = ex.pool->event(ex.lane, 7 * T_ + 4);
        if (!e || cudaEventRecord(e, ex.home) != cudaSuccess || cudaStreamWaitEvent(ex.crit, e, 0) != cudaSuccess) {
            set_error("factor: fork failed");
            *rc_out = BOBE_E_CUDA;
            return nullptr;
        }
    }
    init_stat_kernel<<<(batch + 127) / 128, 128, 0, ex.crit>>>(fb.dstat, fb.gate, batch, fb.force_refine);
    if ((*rc_out = check_launch("init_stat_kernel")) != BOBE_OK) return nullptr;
    TiledFactor* f = new TiledFactor(ex, fb, npad, batch);
    Tiled& t = f->t;
    t.T = (npad + TW - 1) / TW;
    t.nl = (fb.n_live > 0 && fb.n_live <= npad && fb.n_live > npad - NB && fb.n_live % 16 == 0) ? fb.n_live : npad;
    t.two = ex.bulk != nullptr && ex.mid != nullptr && ex.inv != nullptr && ex.pool != nullptr && t.T > 2 && fb.Lt != nullptr;
    t.last_writer.assign(t.T, -1);
    if (t.two) {  // make sure every event exists before the first record (creation failure -> single-stream fallback)
        if (!ex.pool->event(ex.lane, 7 * t.T + 5)) t.two = false;
    }
    if (t.nl < npad && t.T > 1) {
        // the padded rows of L / Linv (columns of U) left of the last diagonal tile: no product computes them any more
        pad_rows_zero_kernel<<<dim3((unsigned)((off_last(npad) + 255) / 256), npad - t.nl, batch), 256, 0, ex.crit>>>(
            fb.L, fb.Linv, fb.U, npad, t.nl, off_last(npad));
        if ((*rc_out = check_launch("pad_rows_zero_kernel")) != BOBE_OK) {
            delete f;
            return nullptr;
        }
    }
    if (fb.zero_band == 0 && npad > NB) {
        // buffers handed to the caller: the whole other triangle must read as zero.  Nothing in the factorisation reads
        // or writes those blocks, so in look-ahead mode the fill runs on the inverse stream, beside the chain
        cudaStream_t st = ex.crit;
        if (t.two) {
            cudaEvent_t e = ex.pool->event(ex.lane, 7 * t.T + 2);
            t.ok(cudaEventRecord(e, ex.crit), "event record");
            t.ok(cudaStreamWaitEvent(ex.inv, e, 0), "event wait");
            t.inv_used = true;
            st = ex.inv;
        }
        zero_other_triangle_kernel<<<dim3(npad / NB, npad / NB, batch), 256, 0, st>>>(fb.L, nullptr, fb.Linv, fb.U, npad, 0);
        if ((*rc_out = check_launch("zero_other_triangle_kernel")) != BOBE_OK) {
            delete f;
            return nullptr;
        }
    }
    f->pw = ex.pw < 1 ? 1 : (ex.pw > t.T ? t.T : ex.pw);
    return f;
}
int tiled_steps(const TiledFactor* f) { return f->t.T; }
void tiled_step(TiledFactor* f, int j) {
    if (f->t.rc == BOBE_OK && j < f->t.T) f->t.step(j, f->pw);
}
int32_t tiled_finish(TiledFactor* f) {
    if (f->t.rc == BOBE_OK) f->t.finish_phase1();
    if (f->t.rc == BOBE_OK) f->t.phase2();
    if (f->ex.crit != f->ex.home) {  // join (also after an error: the caller's stream must not run ahead of stray work)
        cudaEvent_t e = f->ex.pool->event(f->ex.lane, 7 * f->t.T + 3);
        if (!e || cudaEventRecord(e, f->ex.crit) != cudaSuccess || cudaStreamWaitEvent(f->ex.home, e, 0) != cudaSuccess) {
            if (f->t.rc == BOBE_OK) {
                set_error("factor: join failed");
                f->t.rc = BOBE_E_CUDA;
            }
        }
    }
    const int32_t rc = f->t.rc;
    delete f;
    return rc;
}

int32_t factor_tiled(const FactorExec& ex, const FactorBuffers& fb, int npad, int batch) {
    int32_t rc;
    TiledFactor* f = tiled_begin(ex, fb, npad, batch, &rc);
    if (!f) return rc;
    for (int j = 0; j < tiled_steps(f); ++j) tiled_step(f, j);
    return tiled_finish(f);
}

}  // namespace bobe
