// Log marginal likelihood + analytic gradient for R hyper-parameter restarts in lock step.
// Replaces jax.value_and_grad(GP.neg_mll) (BOBE/gp.py:385-398, gp_mll :170-178; called from
// BOBE/optim.py:118,211,309) for the data term; priors are O(d) and stay on the host.
//
//   log p = -1/2 y^T K^-1 y - sum log L_ii - n/2 log 2 pi
//   d log p / d theta = 1/2 sum_ik W_ik dK_ik/d theta,   W = alpha alpha^T - K^-1
//   dK_ik/d log l_j = G_ik s_ikj,  s_ikj = ((x_ij - x_kj)/l_j)^2,
//       G = K0 (RBF);  G = kv 5/3 (1 + sqrt5 r) exp(-sqrt5 r), 0 where the 1e-30 clamp is active (Matern-5/2)
//   dK/d log kv = K0 (the noise-free kernel);  noise is never optimised;  tausq has no kernel gradient.
#include <algorithm>
#include <cmath>
#include <mutex>
#include <vector>

#include "gemm_nt.cuh"
#include "kernels.cuh"

namespace bobe {

__global__ void exp_params_kernel(const double* __restrict__ lp, int64_t R, int64_t P, int64_t d, int has_kv,
                                  double fixed_kv, double* __restrict__ ls, double* __restrict__ kv) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < R * d) {
        int64_t r = i / d, k = i - r * d;
        ls[i] = exp(lp[r * P + k]);
    }
    if (i < R) kv[i] = has_kv ? exp(lp[i * P + d]) : fixed_kv;
}

constexpr int GT = 64, GLD = 66;

// One CTA (256 threads as 16 x 16) per lower-triangular 64 x 64 tile pair: thread (ty, tx) owns rows 4 ty .. 4 ty + 3 and
// the columns {2 tx, 2 tx + 1, 32 + 2 tx, 33 + 2 tx} (conflict-free LDS.128 of the column operand, as in kmat_kernel).
// Both operands come from the pre-scaled, transposed inputs xs[k][j] = X_jk / l_k of this restart (plain copies, no
// divisions), the 16 kernel values of a thread are evaluated in lock step (kernel_and_g_from_q_n), and the weights
// W = alpha alpha^T - K^-1 are read as 16-byte vectors.
template <int KIND>
__global__ void __launch_bounds__(256, 2) mll_grad_tile_kernel(const double* __restrict__ xs_all, int64_t n, int d,
                                                            const double* __restrict__ kv_all,
                                                            const double* __restrict__ Kinv, int npad,
                                                            const double* __restrict__ alpha_all, int has_kv,
                                                            int P, double* __restrict__ partial, int ntiles) {
    extern __shared__ __align__(16) double sm[];
    double* sa = sm;                // [d][GLD]
    double* sb = sa + d * GLD;      // [d][GLD]
    double* sai = sb + d * GLD;     // [GT] alpha rows
    double* sak = sai + GT;         // [GT] alpha cols
    double* wred = sak + GT;        // [8][d+1] per-warp partial sums
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4, lane = tid & 31, warp = tid >> 5;
    const int64_t z = blockIdx.y;
    // lower-triangular tile index -> (ti, tj), ti >= tj
    int x = blockIdx.x;
    int ti = (int)((sqrt(8.0 * x + 1.0) - 1.0) * 0.5);
    while ((ti + 1) * (ti + 2) / 2 <= x) ++ti;
    while (ti * (ti + 1) / 2 > x) --ti;
    const int tj = x - ti * (ti + 1) / 2;
    const int64_t i0 = (int64_t)ti * GT, k0 = (int64_t)tj * GT;
    const double* xs = xs_all + z * (int64_t)d * npad;
    const double kv = kv_all[z];
    const double* alpha = alpha_all + z * npad;
    const double* Ki = Kinv + z * (int64_t)npad * npad;

    for (int idx = tid; idx < d * (GT / 2); idx += 256) {  // 16-byte copies; columns >= n of xs are zero
        const int k = idx >> 5, c = (idx & 31) * 2;
        *reinterpret_cast<double2*>(sa + k * GLD + c) = *reinterpret_cast<const double2*>(xs + (int64_t)k * npad + i0 + c);
        *reinterpret_cast<double2*>(sb + k * GLD + c) = *reinterpret_cast<const double2*>(xs + (int64_t)k * npad + k0 + c);
    }
    if (tid < GT) {
        sai[tid] = (i0 + tid < n) ? alpha[i0 + tid] : 0.0;
        sak[tid] = (k0 + tid < n) ? alpha[k0 + tid] : 0.0;
    }
    __syncthreads();

    const double* ap = sa + ty * 4;
    const double* bp = sb + 2 * tx;
    double q[16];
#pragma unroll
    for (int e = 0; e < 16; ++e) q[e] = 0.0;
#pragma unroll 2
    for (int k = 0; k < d; ++k) {
        const double2 a01 = *reinterpret_cast<const double2*>(ap + k * GLD);
        const double2 a23 = *reinterpret_cast<const double2*>(ap + k * GLD + 2);
        const double2 b01 = *reinterpret_cast<const double2*>(bp + k * GLD);
        const double2 b23 = *reinterpret_cast<const double2*>(bp + k * GLD + 32);
        const double a[4] = {a01.x, a01.y, a23.x, a23.y}, b[4] = {b01.x, b01.y, b23.x, b23.y};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const double df = a[i] - b[j];
                q[4 * i + j] = fma(df, df, q[4 * i + j]);
            }
    }
    double kval[16], G[16];
#pragma unroll
    for (int h = 0; h < 2; ++h) {  // two lock-step groups of 8: keeps the register footprint at two CTAs per SM
        double qq[8], kk8[8], gg8[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) qq[e] = q[8 * h + e];
        kernel_and_g_from_q_n<KIND, 8>(qq, kv, kk8, gg8);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            kval[8 * h + e] = kk8[e];
            G[8 * h + e] = gg8[e];
        }
    }

    const double wsym = (ti == tj) ? 1.0 : 2.0;  // off-diagonal tiles stand for both (i,k) and (k,i)
    const int64_t c0 = k0 + 2 * tx, c1 = c0 + 32;
    const double ak[4] = {sak[2 * tx], sak[2 * tx + 1], sak[32 + 2 * tx], sak[33 + 2 * tx]};
    double gkv = 0.0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int64_t row = i0 + ty * 4 + i;
        const double ai = sai[ty * 4 + i];
        // K^-1 row segment: rows < npad always exist; columns c0, c0+1, c1, c1+1 < npad
        const double2 k01 = *reinterpret_cast<const double2*>(Ki + row * npad + c0);
        const double2 k23 = *reinterpret_cast<const double2*>(Ki + row * npad + c1);
        const double kin[4] = {k01.x, k01.y, k23.x, k23.y};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int64_t col = (j < 2 ? c0 : c1 - 2) + j;
            const bool ok = row < n && col < n;
            const double w = ok ? wsym * (ai * ak[j] - kin[j]) : 0.0;
            gkv = fma(w, kval[4 * i + j], gkv);
            G[4 * i + j] = w * G[4 * i + j];  // wg
        }
    }
    const int np1 = d + 1;
    // Eight dimensions per trip: their eight warp reductions run interleaved (eight independent shuffle / add chains instead
    // of one 5-deep dependent chain per dimension inside the loop); the order of every sum is that of warp_sum().
    for (int kb = 0; kb < d; kb += 8) {
        double s[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int k = kb + u;
            double s0 = 0.0, s1 = 0.0;  // two chains: the 16 terms of one dimension do not serialise on one accumulator
            if (k < d) {
                const double2 a01 = *reinterpret_cast<const double2*>(ap + k * GLD);
                const double2 a23 = *reinterpret_cast<const double2*>(ap + k * GLD + 2);
                const double2 b01 = *reinterpret_cast<const double2*>(bp + k * GLD);
                const double2 b23 = *reinterpret_cast<const double2*>(bp + k * GLD + 32);
                const double a[4] = {a01.x, a01.y, a23.x, a23.y}, b[4] = {b01.x, b01.y, b23.x, b23.y};
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; j += 2) {
                        const double d0 = a[i] - b[j], d1 = a[i] - b[j + 1];
                        s0 = fma(G[4 * i + j], d0 * d0, s0);
                        s1 = fma(G[4 * i + j + 1], d1 * d1, s1);
                    }
            }
            s[u] = s0 + s1;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
#pragma unroll
            for (int u = 0; u < 8; ++u) s[u] += __shfl_xor_sync(0xffffffffu, s[u], o);
        if (lane == 0)
#pragma unroll
            for (int u = 0; u < 8; ++u)
                if (kb + u < d) wred[warp * np1 + kb + u] = s[u];
    }
    gkv = warp_sum(gkv);
    if (lane == 0) wred[warp * np1 + d] = gkv;
    __syncthreads();
    if (tid < np1) {
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < 8; ++w) s += wred[w * np1 + tid];
        partial[(z * ntiles + blockIdx.x) * np1 + tid] = s;
    }
}

// one warp per parameter: lanes stride over the tile partials (fixed order per lane, fixed shuffle tree: deterministic)
__global__ void __launch_bounds__(256) mll_finish_kernel(const double* __restrict__ partial, int ntiles, int d, int P,
                                                         int has_kv, int64_t n, const double* __restrict__ logdet,
                                                         const double* __restrict__ quad,
                                                         const int32_t* __restrict__ info, double* __restrict__ val,
                                                         double* __restrict__ grad) {
    const int64_t z = blockIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int np1 = d + 1;
    for (int j = warp; j < P; j += 8) {
        double g = 0.0;
        if (j < d || (has_kv && j == d)) {
            double s = 0.0;
            for (int t = lane; t < ntiles; t += 32) s += partial[(z * ntiles + t) * np1 + j];
            g = 0.5 * warp_sum(s);
        }
        if (info[z]) g = nan("");
        if (lane == 0) grad[z * P + j] = g;
    }
    if (threadIdx.x == 0) val[z] = -0.5 * quad[z] - logdet[z] - 0.5 * (double)n * 1.8378770664093454835606594728112;
}

constexpr int MLL_MAX_STREAMS = POOL_STREAMS / POOL_LANE_STREAMS;  // one lane of pool streams per sub-batch

struct MllLayout {
    int64_t npad, tiles, ntile_pairs;
    int64_t off_ls, off_kv, off_KB, off_L, off_Lt, off_Linv, off_U, off_Q, off_diag, off_stat, off_z, off_alpha, off_logdet,
        off_quad, off_partial, off_xs, total;
};

static MllLayout mll_layout(int64_t n, int64_t d, int64_t R) {
    MllLayout l{};
    l.npad = npad_of(n);
    l.tiles = (n + GT - 1) / GT;
    l.ntile_pairs = l.tiles * (l.tiles + 1) / 2;
    int64_t m2 = l.npad * l.npad, o = 0;
    auto take = [&](int64_t doubles) {
        int64_t at = o;
        o += round_up(doubles, 32);
        return at;
    };
    l.off_ls = take(R * d);
    l.off_kv = take(R);
    l.off_KB = take(R * m2);
    l.off_L = take(R * m2);
    l.off_Lt = take(R * m2);
    l.off_Linv = take(R * m2);
    l.off_U = take(R * m2);
    l.off_Q = take(R * factor_q_elems(l.npad));
    l.off_diag = take(R * l.npad);
    l.off_stat = take(2 * R + (factor_gate_rows(l.npad) * R + 1) / 2);  // min / max pivot + the gate rows (ints)
    l.off_z = take((3 * R + MLL_MAX_STREAMS) * l.npad);
    l.off_alpha = take(R * l.npad);
    l.off_logdet = take(R);
    l.off_quad = take(R);
    l.off_partial = take(R * l.ntile_pairs * (d + 1));
    l.off_xs = take(R * d * l.npad);
    l.total = o;
    return l;
}

}  // namespace bobe

using namespace bobe;

extern "C" int64_t bobe_mll_grad_workspace_bytes(int64_t n, int64_t d, int64_t R) {
    if (n <= 0 || d <= 0 || R <= 0) return 0;
    return mll_layout(n, d, R).total * 8 + 256;
}

// Optional (BOBE_MLL_GRAPH=1, default off): a call whose arguments (all pointers, sizes and scalars) repeat is captured
// into a CUDA graph at its third sighting -- the same launches, dependencies and programmatic-launch edges -- and
// replayed from then on: one graph launch (~0.05 ms of host time) instead of 1 - 2 ms of enqueueing several hundred
// launches and event operations on up to twelve streams.  Measured (profiles/r02/README.md): the enqueue already overlaps
// the device work, so the round does not get shorter (8 restarts: 3.47 vs 3.49 ms call + fetch), while capture +
// instantiation cost ~10 ms per key -- a loss for optimisers whose batch shrinks as restarts retire.  It pays only where
// the host thread is the scarce resource.  A caller that is itself capturing is left alone.
namespace {
struct GraphKey {
    const void *X, *y, *lp, *val, *grad, *info, *ws;
    int64_t n, d, R, P, ws_bytes;
    int32_t kind, has_kv;
    double fixed_kv, noise;
    int dev;
    bool operator==(const GraphKey& o) const {
        return X == o.X && y == o.y && lp == o.lp && val == o.val && grad == o.grad && info == o.info && ws == o.ws && n == o.n &&
               d == o.d && R == o.R && P == o.P && ws_bytes == o.ws_bytes && kind == o.kind && has_kv == o.has_kv &&
               fixed_kv == o.fixed_kv && noise == o.noise && dev == o.dev;
    }
};
struct GraphEntry {
    GraphKey key;
    cudaGraphExec_t exec;  // null: not captured yet
    uint64_t stamp;
    int sightings;
};
constexpr int GRAPH_CAPTURE_AFTER = 3;  // capture + instantiate costs several plain enqueues: only for keys that keep coming
constexpr size_t GRAPH_CACHE = 16;
std::mutex g_graph_mu;
std::vector<GraphEntry> g_graphs;
uint64_t g_graph_clock = 0;

int32_t mll_grad_enqueue(cudaStream_t stream, int32_t kind, const double* X, const double* y, int64_t n, int64_t d,
                         const double* log_params, int64_t R, int64_t P, int32_t has_kv, double fixed_kv, double noise,
                         double* val, double* grad, int32_t* info, void* ws, int64_t ws_bytes);
}  // namespace

extern "C" int32_t bobe_mll_grad_batched(void* stream_, int32_t kind, const double* X, const double* y, int64_t n,
                                         int64_t d, const double* log_params, int64_t R, int64_t P, int32_t has_kv,
                                         double fixed_kv, double noise, double* val, double* grad, int32_t* info,
                                         void* ws, int64_t ws_bytes) {
    cudaStream_t stream = (cudaStream_t)stream_;
    static const bool graphs = env_int("BOBE_MLL_GRAPH", 0) != 0;
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    int dev = 0;
    if (!graphs || cudaStreamIsCapturing(stream, &cap) != cudaSuccess || cap != cudaStreamCaptureStatusNone ||
        cudaGetDevice(&dev) != cudaSuccess)
        return mll_grad_enqueue(stream, kind, X, y, n, d, log_params, R, P, has_kv, fixed_kv, noise, val, grad, info, ws, ws_bytes);
    const GraphKey key{X, y, log_params, val, grad, info, ws, n, d, R, P, ws_bytes, kind, has_kv, fixed_kv, noise, dev};
    std::lock_guard<std::mutex> lock(g_graph_mu);
    GraphEntry* hit = nullptr;
    for (auto& e : g_graphs)
        if (e.key == key) hit = &e;
    if (hit && hit->exec) {
        hit->stamp = ++g_graph_clock;
        if (cudaGraphLaunch(hit->exec, stream) == cudaSuccess) return BOBE_OK;
        cudaGetLastError();
        cudaGraphExecDestroy(hit->exec);  // should not happen; fall back to the plain path for good
        hit->exec = nullptr;
        return mll_grad_enqueue(stream, kind, X, y, n, d, log_params, R, P, has_kv, fixed_kv, noise, val, grad, info, ws, ws_bytes);
    }
    if (!hit) {  // first sighting: remember the key, run plainly (also warms every function attribute / pool object)
        if (g_graphs.size() >= GRAPH_CACHE) {
            size_t old = 0;
            for (size_t i = 1; i < g_graphs.size(); ++i)
                if (g_graphs[i].stamp < g_graphs[old].stamp) old = i;
            if (g_graphs[old].exec) cudaGraphExecDestroy(g_graphs[old].exec);
            g_graphs.erase(g_graphs.begin() + old);
        }
        g_graphs.push_back(GraphEntry{key, nullptr, ++g_graph_clock, 1});
        return mll_grad_enqueue(stream, kind, X, y, n, d, log_params, R, P, has_kv, fixed_kv, noise, val, grad, info, ws, ws_bytes);
    }
    hit->stamp = ++g_graph_clock;
    if (++hit->sightings < GRAPH_CAPTURE_AFTER)
        return mll_grad_enqueue(stream, kind, X, y, n, d, log_params, R, P, has_kv, fixed_kv, noise, val, grad, info, ws, ws_bytes);
    // the key keeps coming: capture -- on a stream of our own (the caller's may be the legacy default stream, which cannot
    // be captured; torch's default stream is), the graph is then launched into the caller's stream
    static cudaStream_t cap_streams[64] = {nullptr};
    if (dev < 0 || dev >= 64 ||
        (!cap_streams[dev] && cudaStreamCreateWithFlags(&cap_streams[dev], cudaStreamNonBlocking) != cudaSuccess) ||
        cudaStreamBeginCapture(cap_streams[dev], cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
        cudaGetLastError();
        hit->key.dev = -1;
        return mll_grad_enqueue(stream, kind, X, y, n, d, log_params, R, P, has_kv, fixed_kv, noise, val, grad, info, ws, ws_bytes);
    }
    cudaStream_t cs = cap_streams[dev];
    const int32_t rc = mll_grad_enqueue(cs, kind, X, y, n, d, log_params, R, P, has_kv, fixed_kv, noise, val, grad, info, ws, ws_bytes);
    cudaGraph_t graph = nullptr;
    const cudaError_t ce = cudaStreamEndCapture(cs, &graph);
    cudaGraphExec_t exec = nullptr;
    if (rc == BOBE_OK && ce == cudaSuccess && graph && cudaGraphInstantiate(&exec, graph, 0) == cudaSuccess &&
        cudaGraphLaunch(exec, stream) == cudaSuccess) {
        cudaGraphDestroy(graph);
        hit->exec = exec;
        if (env_int("BOBE_MLL_GRAPH_DEBUG", 0)) fprintf(stderr, "bobe: mll_grad graph captured (R=%lld)\n", (long long)R);
        return BOBE_OK;
    }
    if (env_int("BOBE_MLL_GRAPH_DEBUG", 0))
        fprintf(stderr, "bobe: mll_grad graph capture FAILED rc=%d end=%s last=%s\n", rc, cudaGetErrorString(ce),
                cudaGetErrorString(cudaPeekAtLastError()));
    cudaGetLastError();  // capture did not work out here: nothing was executed, so run the call plainly (and stop trying)
    if (exec) cudaGraphExecDestroy(exec);
    if (graph) cudaGraphDestroy(graph);
    hit->key.dev = -1;  // never matches again
    if (rc != BOBE_OK) return rc;
    return mll_grad_enqueue(stream, kind, X, y, n, d, log_params, R, P, has_kv, fixed_kv, noise, val, grad, info, ws, ws_bytes);
}

namespace {
int32_t mll_grad_enqueue(cudaStream_t stream, int32_t kind, const double* X, const double* y, int64_t n, int64_t d,
                         const double* log_params, int64_t R, int64_t P, int32_t has_kv, double fixed_kv, double noise,
                         double* val, double* grad, int32_t* info, void* ws, int64_t ws_bytes) {
    if (!X || !y || !log_params || !val || !grad || !info || !ws) {
        set_error("mll_grad: null pointer");
        return BOBE_E_ARG;
    }
    if (n <= 0 || d <= 0 || R <= 0 || P < d + (has_kv ? 1 : 0) || P > 256 || d > BOBE_MAX_DIM) {
        set_error("mll_grad: bad sizes n=%lld d=%lld R=%lld P=%lld", (long long)n, (long long)d, (long long)R,
                  (long long)P);
        return BOBE_E_ARG;
    }
    MllLayout l = mll_layout(n, d, R);
    double* w = (double*)(((uintptr_t)ws + 255) & ~(uintptr_t)255);
    if ((char*)(w + l.total) > (char*)ws + ws_bytes) {
        set_error("mll_grad: workspace too small (%lld < %lld)", (long long)ws_bytes,
                  (long long)bobe_mll_grad_workspace_bytes(n, d, R));
        return BOBE_E_WORKSPACE;
    }
    double *ls = w + l.off_ls, *kv = w + l.off_kv;
    const int npad = (int)l.npad;
    const int64_t m2 = (int64_t)npad * npad, qel = factor_q_elems(npad);

    int64_t tot = R * d > R ? R * d : R;
    exp_params_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, stream>>>(log_params, R, P, d, has_kv, fixed_kv, ls, kv);
    if (int32_t rc = check_launch("exp_params_kernel")) return rc;

    // The factorisation is a chain of dependent launches whose small products cannot fill 148 SMs.  The restarts are
    // therefore cut into sub-batches that run the whole chain on separate internal streams (forked from and joined to
    // the caller's stream with events): one sub-batch's latency-bound leaves and small products overlap with another's
    // large tensor-core products.  The host enqueues the sub-batches STAGE BY STAGE (and the factorisation tile column
    // by tile column): enqueued one chain after the other, the second would start half a millisecond behind the first.
    // Measured policy (profiles/r02/sub_batch_sweep.log, n = 2000): two sub-batches from 8 restarts on, three from 32 on.
    // BOBE_MLL_STREAMS / BOBE_MLL_MIN_PER_STREAM force S = min(streams, R / min_per_stream) instead (experiments).
    static const int64_t forced_streams = env_int("BOBE_MLL_STREAMS", 0);
    static const int64_t max_streams = std::min<int64_t>(MLL_MAX_STREAMS, forced_streams > 0 ? forced_streams : 4);
    static const int64_t min_per_stream = env_int("BOBE_MLL_MIN_PER_STREAM", forced_streams > 0 ? 4 : 0);  // 0: not forced
    static const int64_t scheme = env_int("BOBE_FACTOR", 1);
    int S;
    if (forced_streams > 0 || min_per_stream > 0)
        S = (int)std::min<int64_t>(max_streams, std::max<int64_t>(1, R / std::max<int64_t>(1, min_per_stream)));
    else
        S = R < 8 ? 1 : (R < 32 ? 2 : 3);
    // XLA may call handlers from several host threads: the record / wait pairs below must not interleave with those
    // of another caller of the same device (stream order then keeps the shared side streams correct)
    StreamPool* pool = stream_pool();
    if (!pool) return BOBE_E_CUDA;
    std::unique_lock<std::mutex> pool_lock(pool->enqueue_mu);
    // (also for a single sub-batch: the pool's chain streams have a higher priority than its look-ahead streams)
    static const bool own_stream = env_int("BOBE_MLL_OWN_STREAM", 1) != 0;
    const bool forked = S > 1 || own_stream;
    if (forked) {
        if (cudaEventRecord(pool->fork, stream) != cudaSuccess) {
            set_error("mll_grad: event record failed");
            return BOBE_E_CUDA;
        }
    }
    struct Sub {
        int64_t r0, Rs;
        cudaStream_t st;
        FactorBuffers fb;
        double *zws, *alpha, *logdet, *quad, *partial, *xs;
        const double *ls_s, *kv_s;
        TiledFactor* tf = nullptr;
        int32_t rc = BOBE_OK;
    };
    std::vector<Sub> subs(S);
    for (int si = 0; si < S; ++si) {
        Sub& u = subs[si];
        u.r0 = si * R / S;
        u.Rs = (si + 1) * R / S - u.r0;
        const int64_t r0 = u.r0;
        u.st = forked ? pool->streams[POOL_LANE_STREAMS * si] : stream;
        if (forked) cudaStreamWaitEvent(u.st, pool->fork, 0);
        u.fb = FactorBuffers{w + l.off_KB + r0 * m2, w + l.off_L + r0 * m2, w + l.off_Lt + r0 * m2,
                             w + l.off_Linv + r0 * m2, w + l.off_U + r0 * m2, w + l.off_Q + r0 * qel,
                             w + l.off_diag + r0 * npad, w + l.off_stat + 2 * r0, (int*)(w + l.off_stat + 2 * R) + r0 * factor_gate_rows(npad), 0, 2, factor_live_rows(n)};
        u.zws = w + l.off_z + (3 * r0 + si) * npad;  // each sub-batch: own padded y + 3 vectors per restart
        u.alpha = w + l.off_alpha + r0 * npad;
        u.logdet = w + l.off_logdet + r0;
        u.quad = w + l.off_quad + r0;
        u.partial = w + l.off_partial + r0 * l.ntile_pairs * (d + 1);
        u.ls_s = ls + r0 * d;
        u.kv_s = kv + r0;
        u.xs = w + l.off_xs + r0 * d * npad;
    }
    auto each = [&](auto&& stage) {  // one stage for every sub-batch that is still healthy
        for (int si = 0; si < S; ++si)
            if (subs[si].rc == BOBE_OK) subs[si].rc = stage(subs[si], si);
    };
    each([&](Sub& u, int) -> int32_t {
        if (int32_t rc = launch_prescale(u.st, X, n, d, u.ls_s, d, u.xs, npad, d * (int64_t)npad, (int)u.Rs)) return rc;
        KmatArgs ka{};
        ka.xa = X; ka.xb = X; ka.ls = u.ls_s; ka.kv_ptr = u.kv_s; ka.out = u.fb.KB;
        ka.xbs = u.xs; ka.xbs_ld = npad; ka.xbs_stride = d * (int64_t)npad;
        ka.n1 = n; ka.n2 = n; ka.d = d; ka.ldo = npad; ka.rows_pad = npad; ka.cols_pad = npad;
        ka.store_rows = npad; ka.store_cols = npad; ka.vec_ok = 1;
        ka.ls_stride = d; ka.out_stride = m2; ka.noise = noise; ka.add_noise = 1; ka.pad_identity = 1;
        ka.lower_only = 1;  // the factorisation reads the lower triangle only
        return launch_kmat(u.st, kind, ka, (int)u.Rs);
    });
    if (scheme == 0) {
        each([&](Sub& u, int) -> int32_t { return factor_recursive(u.st, u.fb, npad, (int)u.Rs); });
    } else {
        int steps = 0;
        each([&](Sub& u, int si) -> int32_t {
            int32_t rc;
            u.tf = tiled_begin(factor_exec(u.st, pool, si, (int)u.Rs, (int)R), u.fb, npad, (int)u.Rs, &rc);
            if (u.tf) steps = std::max(steps, tiled_steps(u.tf));
            return rc;
        });
        for (int j = 0; j < steps; ++j)
            for (int si = 0; si < S; ++si)
                if (subs[si].tf) tiled_step(subs[si].tf, j);
        for (int si = 0; si < S; ++si)
            if (subs[si].tf) {
                const int32_t rc = tiled_finish(subs[si].tf);
                subs[si].tf = nullptr;
                if (subs[si].rc == BOBE_OK) subs[si].rc = rc;
            }
    }
    // alpha / log-det / quad (a dozen small, partly gated launches) and K^-1 = U U^T (the largest product of the call) only
    // share their INPUTS (Linv, U): the vector work runs on a side stream of the lane beside the product, joined before
    // the gradient kernel, which needs both
    each([&](Sub& u, int si) -> int32_t {
        static const bool use_side = env_int("BOBE_MLL_SIDE_VECTORS", 1) != 0;
        if (!use_side) {
            SolveArgs sa0{kind, X, u.ls_s, u.kv_s, d, noise, u.xs};
            if (int32_t rc0 = launch_solve_vectors(u.st, u.fb, sa0, y, n, npad, (int)u.Rs, u.zws, u.alpha, u.logdet, u.quad, info + u.r0))
                return rc0;
            return launch_kinv(u.st, u.fb, npad, (int)u.Rs);
        }
        cudaStream_t side = pool->streams[POOL_LANE_STREAMS * si + 1];
        cudaEvent_t e0 = pool->event(si, 0), e1 = pool->event(si, 1);  // (the factorisation's events of this lane are done with)
        if (!e0 || !e1 || cudaEventRecord(e0, u.st) != cudaSuccess || cudaStreamWaitEvent(side, e0, 0) != cudaSuccess) {
            set_error("mll_grad: fork failed");
            return BOBE_E_CUDA;
        }
        SolveArgs sa{kind, X, u.ls_s, u.kv_s, d, noise, u.xs};
        const int32_t rc = launch_solve_vectors(side, u.fb, sa, y, n, npad, (int)u.Rs, u.zws, u.alpha, u.logdet, u.quad, info + u.r0);
        const int32_t rk = launch_kinv(u.st, u.fb, npad, (int)u.Rs);
        if (cudaEventRecord(e1, side) != cudaSuccess || cudaStreamWaitEvent(u.st, e1, 0) != cudaSuccess) {
            set_error("mll_grad: join failed");
            return BOBE_E_CUDA;
        }
        return rc != BOBE_OK ? rc : rk;
    });
    each([&](Sub& u, int) -> int32_t {
        int smem = (int)((2 * d * GLD + 2 * GT + 8 * (d + 1)) * sizeof(double));
        dim3 grid((unsigned)l.ntile_pairs, (unsigned)u.Rs);
        if (kind == BOBE_KERNEL_RBF) {
            if (int32_t rc = ensure_smem<mll_grad_tile_kernel<BOBE_KERNEL_RBF>>(smem)) return rc;
            mll_grad_tile_kernel<BOBE_KERNEL_RBF><<<grid, 256, smem, u.st>>>(
                u.xs, n, (int)d, u.kv_s, u.fb.KB, npad, u.alpha, has_kv, (int)P, u.partial, (int)l.ntile_pairs);
        } else {
            if (int32_t rc = ensure_smem<mll_grad_tile_kernel<BOBE_KERNEL_MATERN52>>(smem)) return rc;
            mll_grad_tile_kernel<BOBE_KERNEL_MATERN52><<<grid, 256, smem, u.st>>>(
                u.xs, n, (int)d, u.kv_s, u.fb.KB, npad, u.alpha, has_kv, (int)P, u.partial, (int)l.ntile_pairs);
        }
        if (int32_t rc = check_launch("mll_grad_tile_kernel")) return rc;
        mll_finish_kernel<<<(unsigned)u.Rs, 256, 0, u.st>>>(u.partial, (int)l.ntile_pairs, (int)d, (int)P, has_kv, n, u.logdet,
                                                           u.quad, info + u.r0, val + u.r0, grad + u.r0 * P);
        return check_launch("mll_finish_kernel");
    });
    int32_t rc_all = BOBE_OK;
    for (int si = 0; si < S; ++si) {
        if (subs[si].rc != BOBE_OK && rc_all == BOBE_OK) rc_all = subs[si].rc;
        if (forked) {  // join (also on failure, so that the caller's stream never runs ahead of stray work)
            cudaEventRecord(pool->join[si], subs[si].st);
            cudaStreamWaitEvent(stream, pool->join[si], 0);
        }
    }
    return rc_all;
}
}  // namespace
