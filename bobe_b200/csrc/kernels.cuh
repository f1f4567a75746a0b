// Internal launcher declarations shared between translation units.
#pragma once
#include <mutex>
#include <vector>

#include "common.cuh"

namespace bobe {

const char* last_error();
int64_t env_int(const char* name, int64_t dflt);  // tuning knobs (read once)

// kernel-matrix family (kernel_matrix.cu) ---------------------------------------------------------------
struct KmatArgs {
    const double* xa;      // (n1, d) rows of the output
    const double* xb;      // (n2, d) columns of the output
    const double* xbs;     // optional: xb already scaled and transposed, (batch, d, xbs_ld) from launch_prescale
    int64_t xbs_ld, xbs_stride;
    const double* ls;      // (batch, d) lengthscales, stride ls_stride
    const double* kv_ptr;  // (batch) kernel variances or null -> kv
    const double* alpha;   // (cols_pad) optional: mean_out[i] = sum_j alpha[j] K[i][j]   (needs col_splits == 1)
    double* out;           // (rows_pad, ldo) or null (mean only)
    double* mean_out;      // (n1) optional
    int64_t n1, n2, d, ldo;
    int64_t rows_pad, cols_pad;  // extent computed (multiples of 64); outside (n1, n2): identity if pad_identity else 0
    int64_t store_rows, store_cols;  // extent actually stored to `out`
    int vec_ok;                  // out base 16-byte aligned, ldo and out_stride even -> 16-byte stores
    int64_t ls_stride, out_stride, alpha_stride, mean_stride;
    const int* gate;             // optional per-batch switch (no-op where gate[z] == 0)
    double kv, noise, y_mean, y_std;
    int add_noise;     // + noise on the diagonal (square case)
    int pad_identity;  // 1: padded part is the identity (factorisation input), 0: zeros (K* panels)
    int lower_only;    // 1: square symmetric build, tiles strictly above the diagonal are skipped (left unwritten)
    int mean_standardised;
};
int32_t launch_kmat(cudaStream_t stream, int kind, const KmatArgs& a, int batch);
// xs[z][k][j] = x[j][k] / ls[z][k] (zero for n <= j < ld): the column operand of launch_kmat in its compute layout
int32_t launch_prescale(cudaStream_t stream, const double* x, int64_t n, int64_t d, const double* ls, int64_t ls_stride,
                        double* xs, int64_t ld, int64_t xs_stride, int batch);

// trmm + sumsq (gemm.cu) ---------------------------------------------------------------------------------
constexpr int TRMM_MAX_SPLIT = 32;  // rows of the `partial` scratch (each as long as the query chunk)
// partial: optional scratch of TRMM_MAX_SPLIT * rows_pad doubles; enables the row split for under-filled launches
int32_t launch_trmm_sumsq(cudaStream_t stream, const double* Linv, int n, int npad, const double* Kstar, int64_t ldk,
                          int64_t rows_pad, int64_t q_begin, int64_t M, double kk, double scale, int standardised,
                          double* var_out, double* partial = nullptr, int* counters = nullptr);
// counters: optional TRMM_COUNTERS zeroed ints (with `partial`): enables the shared-panel schedule of the TMA kernel, in which
// groups of S CTAs walk the query tiles together so that the K* panels in flight fit in L2 (see trmm_sumsq_tma_kernel).
constexpr int TRMM_COUNTERS = 512;
// query tiles one full-size launch should carry for this n (a multiple of the group count of the shared-panel schedule)
int trmm_chunk_tiles(int n, int npad);

// VT = (Linv Kin^T)^T through the TMA trmm pipeline; 1 = route not available (use launch_gemm_nt), 0 = done, < 0 = error
int32_t launch_trmm_store(cudaStream_t stream, const double* Linv, int n, int npad, const double* Kin, int64_t ldk,
                          int64_t rows_pad, double* VT, int64_t ldv);

// factorisation (factor.cu) -----------------------------------------------------------------------------
struct FactorBuffers {
    double* KB;    // (batch, npad, npad) in: K (lower used); out: scratch / K^-1 if requested
    double* L;     // (batch, npad, npad) lower factor, zero upper
    double* Lt;    // (batch, npad, npad) its transpose
    double* Linv;  // (batch, npad, npad) L^-1 lower
    double* U;     // (batch, npad, npad) (L^-1)^T upper
    double* Q;     // (batch, npad/2+NB, npad/2+NB) scratch
    double* diag;  // (batch, npad) diagonal of L
    double* dstat; // (batch, 2) running min / max pivot
    int* gate;     // (factor_gate_rows(npad), batch) ints: row 0 = latest state (1 once max/min pivot exceeds the refinement
                   // ratio), row j + 1 = snapshot after tile column j (factor_tiled.cuh)
    int force_refine;
    int zero_band;  // 0: zero the whole other triangle (L / Linv handed to the caller); 2: internal use only
    // rows / columns >= n_live (a multiple of 16, n <= n_live <= npad; 0: npad) are pure padding -- identity on the diagonal,
    // zero elsewhere: the tile-column scheme restricts every product to the live part ((npad / n)^3 = 7.4 % of the work at
    // n = 2000) and fills the padded rows of L / Linv / U with zeros instead of computing them
    int n_live = 0;
};
int factor_live_rows(int64_t n);  // n rounded up to the k granularity of the GEMM tiles in use (16; 32 with 128-wide tiles)
int64_t factor_q_elems(int64_t npad);
inline int64_t factor_gate_rows(int64_t npad) { return (npad + 127) / 128 + 2; }

// internal side streams + dependency events (one pool per device, created on first use)
constexpr int POOL_LANE_STREAMS = 4;  // per lane: chain, rest of the current column, trailing products, inverse tree
constexpr int POOL_STREAMS = 8 * POOL_LANE_STREAMS;
struct StreamPool {
    cudaStream_t streams[POOL_STREAMS];
    // the same roles on two GREEN CONTEXTS with disjoint SM sets (driver API, CUDA >= 12.4): role 0 (the dependent chain:
    // leaves + the small products between them) on a partition of a few SMs of its own, roles 1-3 on the rest.  A leaf
    // needs a whole SM's shared memory; on a machine filled with the look-ahead products of a batch of matrices it waited
    // for an SM to drain completely, several hundred microseconds at a time (stream priorities do not reserve anything).
    // Null when the driver refuses (then `streams` is used).
    cudaStream_t gstreams[POOL_STREAMS];
    bool green = false;
    int green_sms = 0;
    cudaEvent_t fork, join[POOL_STREAMS];
    std::vector<cudaEvent_t> ring[POOL_STREAMS];  // per-lane dependency events, created on demand
    std::mutex enqueue_mu;  // the events are shared: one host thread enqueues on the pool at a time
    cudaEvent_t event(int lane, int idx);
};
StreamPool* stream_pool();

// where one factorisation chain runs: `crit` carries the dependent chain; `bulk` / `inv` (optional, with pool / lane for
// their events) the look-ahead products and the inverse tree.  pw = tile columns per outer panel (factor_tiled.cuh).
struct FactorExec {
    cudaStream_t home;  // the stream the caller's stages before / after the factorisation run on (crit is forked from it)
    cudaStream_t crit;
    cudaStream_t mid;
    cudaStream_t bulk;
    cudaStream_t inv;
    StreamPool* pool;
    int lane;
    int pw;
};
int32_t factor_tiled(const FactorExec& ex, const FactorBuffers& fb, int npad, int batch);
// the same, one tile column at a time (sub-batches advanced in lock step); tiled_finish frees the handle
struct TiledFactor;
TiledFactor* tiled_begin(const FactorExec& ex, const FactorBuffers& fb, int npad, int batch, int32_t* rc_out);
int tiled_steps(const TiledFactor* f);
void tiled_step(TiledFactor* f, int j);
int32_t tiled_finish(TiledFactor* f);
// streams / knobs of lane `lane` (BOBE_FACTOR_PW, BOBE_LOOKAHEAD_MAX); bulk / inv are null when look-ahead is off
// total_batch: matrices in flight over ALL lanes of this call (decides whether the chain gets its own SM partition)
FactorExec factor_exec(cudaStream_t stream, StreamPool* pool, int lane, int batch, int total_batch);
// scheme dispatch (BOBE_FACTOR knob); pool may be null (no look-ahead), lane selects the pool streams / events used
int32_t factor_any(cudaStream_t stream, StreamPool* pool, int lane, const FactorBuffers& fb, int npad, int batch);
// factor_any on lane 0 of the pool (entry points that factorise outside bobe_mll_grad_batched)
int32_t factor_on_pool(cudaStream_t stream, StreamPool* pool, const FactorBuffers& fb, int npad, int batch);
int32_t factor_recursive(cudaStream_t stream, const FactorBuffers& fb, int npad, int batch);
int32_t launch_kinv(cudaStream_t stream, const FactorBuffers& fb, int npad, int batch);  // KB <- U U^T (lower 128-tiles)
struct SolveArgs {  // what the alpha refinement needs to rebuild K alpha
    int kind;
    const double* X;
    const double* ls;  // (batch, d)
    const double* kv;  // (batch)
    int64_t d;
    double noise;
    const double* xs;  // (batch, d, npad) scaled, transposed X from launch_prescale (same ls)
    // caller-supplied K instead of (X, ls, kv): (batch, n, ldk) row-major, lower triangle used, noise already inside
    // (bobe_cholesky_batched); null otherwise
    const double* Kin = nullptr;
    int64_t ldk = 0, kstride = 0;
};
int64_t solve_ws_doubles(int64_t npad, int64_t batch);
int32_t launch_pad_k(cudaStream_t stream, const double* K, int64_t ldk, int64_t kstride, int n, int npad, int batch, double* KB);
int32_t launch_solve_vectors(cudaStream_t stream, const FactorBuffers& fb, const SolveArgs& sa, const double* y,
                             int64_t n, int npad, int batch, double* z_ws, double* alpha, double* logdet,
                             double* quad, int32_t* info);

// rank-b append to padded (npad(n_old + b)) factors L / Linv, alpha re-solved for all n_old + b targets (factor.cu)
int64_t append_ws_doubles(int64_t npad);
int32_t factor_append(cudaStream_t stream, int kind, const double* X, const double* y, int64_t n_old, int64_t b, int64_t d,
                      const double* ls, double kv, double noise, double* L, double* Linv, double* alpha, int32_t* info,
                      double* ws);

}  // namespace bobe
