// Shared device/host helpers for the bobe_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <cstdio>
#include <utility>

#include "../../include/bobe_b200.h"

namespace bobe {

constexpr int NB = 64;             // leaf block of the recursive factorisation; npad is a multiple of it
constexpr double SAFE_FLOOR = 1e-12;  // BOBE/gp.py:16
constexpr double SQRT5 = 2.23606797749978969641;

inline int64_t npad_of(int64_t n) { return ((n + NB - 1) / NB) * NB; }
inline int64_t round_up(int64_t a, int64_t b) { return ((a + b - 1) / b) * b; }

void set_error(const char* fmt, ...);
int32_t check_launch(const char* what);

// Raise a kernel's dynamic shared-memory limit.  cudaFuncSetAttribute costs ~10 us of host time, which dominates a
// chain of hundreds of small dependent launches, so it is issued at most once per (kernel, device, size).
template <auto Kernel>
inline int32_t ensure_smem(int bytes) {
    static std::atomic<int> have[64];  // zero-initialised; largest limit already set per device
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) dev = 0;
    if (have[dev].load(std::memory_order_relaxed) >= bytes) return BOBE_OK;
    cudaError_t e = cudaFuncSetAttribute(Kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e != cudaSuccess) {
        set_error("cudaFuncSetAttribute(smem=%d): %s", bytes, cudaGetErrorString(e));
        return BOBE_E_CUDA;
    }
    int cur = have[dev].load(std::memory_order_relaxed);
    while (cur < bytes && !have[dev].compare_exchange_weak(cur, bytes)) {
    }
    return BOBE_OK;
}

// same, for a kernel given as a function-pointer VALUE (one static slot per distinct kernel type + first pointer seen;
// used where the kernel is picked at run time between a few instantiations of one signature)
template <class... KArgs>
inline int32_t ensure_smem_fn(void (*kernel)(KArgs...), int bytes) {
    struct Slot { void* fn; std::atomic<int> have[64]; };
    static Slot slots[8];
    static std::atomic<int> nslots{0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) dev = 0;
    const int n = nslots.load(std::memory_order_acquire);
    for (int i = 0; i < n; ++i)
        if (slots[i].fn == (void*)kernel && slots[i].have[dev].load(std::memory_order_relaxed) >= bytes) return BOBE_OK;
    cudaError_t e = cudaFuncSetAttribute((const void*)kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e != cudaSuccess) {
        set_error("cudaFuncSetAttribute(smem=%d): %s", bytes, cudaGetErrorString(e));
        return BOBE_E_CUDA;
    }
    for (int i = 0; i < n; ++i)
        if (slots[i].fn == (void*)kernel) {
            slots[i].have[dev].store(bytes, std::memory_order_relaxed);
            return BOBE_OK;
        }
    const int idx = nslots.fetch_add(1);
    if (idx < 8) {  // a racing duplicate entry is harmless (the attribute call is idempotent)
        slots[idx].fn = (void*)kernel;
        slots[idx].have[dev].store(bytes, std::memory_order_relaxed);
    }
    return BOBE_OK;
}

// ---- programmatic dependent launch (PDL) --------------------------------------------------------------------
// The factorisation is a chain of ~230 dependent launches; with the launch attribute below the next kernel of the chain
// is made resident while the previous one drains and parks at pdl_wait() until that one has completed and flushed.
// Every kernel launched through launch_pdl MUST call pdl_wait() before its first global-memory access.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
// Only launches that cannot fill the machine anyway are made programmatic (BOBE_PDL_MAX_CTAS, default 592 = 148 SMs x 4
// CTAs): there the chain is latency-bound and the early residency is free; for large grids the parked CTAs of the next
// launch would take slots from the productive CTAs of the other sub-batch streams (measured: -3 % at R = 64).
bool pdl_enabled(int64_t ctas);  // BOBE_PDL / BOBE_PDL_MAX_CTAS knobs (gemm.cu)
template <class... KArgs, class... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                              Args&&... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed =
        pdl_enabled((int64_t)grid.x * grid.y * grid.z) ? 1 : 0;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(std::forward<Args>(args))...);
}

// ---- FP64 tensor-core MMA: D(8x8) += A(8x4, row) * B(4x8, col).  SASS: DMMA.8x8x4 -------------------
// lane = 4*g + t :  a = A[g][t],  b = B[t][g],  c0/c1 = C[g][2t], C[g][2t+1]
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

// ---- cp.async (LDGSTS) 16-byte copies with zero-fill predicate ---------------------------------------
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src, bool valid) {
    uint32_t s = static_cast<uint32_t>(__cvta_generic_to_shared(smem_dst));
    int sz = valid ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(s), "l"(gmem_src), "r"(sz));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N));
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ---- branch-free FP64 elementary functions for the kernel builds -----------------------------------------------
// The CUDA math library's exp()/sqrt() carry special-case paths (denormals, overflow, NaN) behind branches and
// CALLs; the kernel builds only ever see x <= 0 resp. q >= 1e-30, so the straight-line cores suffice.  Both are
// accurate to ~1 ulp (Cody-Waite reduction + degree-13 Taylor on |r| <= ln2/2; rsqrt seed + 1 Goldschmidt step +
// a final residual correction).

// exp(x) for x <= 0.  Results below 2^-1021 (x < -707.7) are flushed to 0.
__device__ __forceinline__ double exp_nonpos(double x) {
    const double SHIFT = 6755399441055744.0;  // 1.5 * 2^52: rounds x*log2(e) to an integer in the low word
    double t = fma(x, 1.4426950408889634074, SHIFT);
    int n = __double2loint(t);
    double fn = t - SHIFT;
    double r = fma(fn, -6.93147180559945286227e-01, x);   // ln2 (hi)
    r = fma(fn, -2.31904681384629955842e-17, r);          // ln2 (lo)
    double p = 1.6059043836821613e-10;                     // 1/13!
    p = fma(p, r, 2.0876756987868099e-09);                 // 1/12!
    p = fma(p, r, 2.5052108385441719e-08);                 // 1/11!
    p = fma(p, r, 2.7557319223985891e-07);                 // 1/10!
    p = fma(p, r, 2.7557319223985893e-06);                 // 1/9!
    p = fma(p, r, 2.4801587301587302e-05);                 // 1/8!
    p = fma(p, r, 1.9841269841269841e-04);                 // 1/7!
    p = fma(p, r, 1.3888888888888889e-03);                 // 1/6!
    p = fma(p, r, 8.3333333333333332e-03);                 // 1/5!
    p = fma(p, r, 4.1666666666666664e-02);                 // 1/4!
    p = fma(p, r, 1.6666666666666666e-01);                 // 1/3!
    p = fma(p, r, 0.5);
    p = fma(p, r, 1.0);
    p = fma(p, r, 1.0);
    double res = __hiloint2double(__double2hiint(p) + (n << 20), __double2loint(p));  // p * 2^n, n in [-1021, 0]
    return x < -707.7 ? 0.0 : res;
}

// sqrt(q) for normal positive q: MUFU.RSQ64H seed (2^-20 relative), ONE Goldschmidt step (-> ~2^-39) and a residual
// correction (-> ~2^-78 before the final rounding).  tools/sqrt_probe.cu: identical to the correctly rounded
// __dsqrt_rn on all 2^30 sampled arguments over 2^-100 .. 2^20 (profiles/r01/sqrt_probe.txt).
__device__ __forceinline__ double sqrt_pos(double q) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(q));
    double g = q * y, h = 0.5 * y;
    double r = fma(-g, h, 0.5);
    g = fma(g, r, g);
    h = fma(h, r, h);
    return fma(fma(-g, g, q), h, g);  // residual correction: g + (q - g^2) / (2 g)
}

// kernel value from the squared scaled distance q (BOBE/gp.py:149-151 and :161-165)
template <int KIND>
__device__ __forceinline__ double kernel_from_q(double q, double kv) {
    if (KIND == BOBE_KERNEL_RBF) {
        return kv * exp_nonpos(-0.5 * q);
    } else {
        double r = sqrt_pos(q < 1e-30 ? 1e-30 : q);
        double e = exp_nonpos(-SQRT5 * r);
        double poly = 1.0 + r * (SQRT5 + r * (5.0 / 3.0));
        return kv * poly * e;
    }
}

// Taylor coefficients 1/12! .. 1/0! of exp (Horner order, after the leading 1/13!)
__constant__ double EXP_TAYLOR[14] = {2.0876756987868099e-09, 2.5052108385441719e-08, 2.7557319223985891e-07,
                                      2.7557319223985893e-06, 2.4801587301587302e-05, 1.9841269841269841e-04,
                                      1.3888888888888889e-03, 8.3333333333333332e-03, 4.1666666666666664e-02,
                                      1.6666666666666666e-01, 0.5, 1.0, 1.0, 0.0};

// ---- the same functions for N values in lock step ----------------------------------------------------------------
// One value at a time the polynomial is a chain of ~30 dependent FP64 instructions, each waiting out the pipe
// latency, and every 64-bit coefficient costs two UMOVs.  Evaluating N independent values per Horner step keeps
// N instructions in flight per warp and amortises each coefficient load over N uses.  Arithmetic per value is
// bit-identical to exp_nonpos / sqrt_pos / kernel_from_q above.
template <int N>
__device__ __forceinline__ void exp_nonpos_n(const double (&x)[N], double (&res)[N]) {
    const double SHIFT = 6755399441055744.0;
    double r[N], p[N];
    int n[N];
#pragma unroll
    for (int j = 0; j < N; ++j) {
        double t = fma(x[j], 1.4426950408889634074, SHIFT);
        n[j] = __double2loint(t);
        double fn = t - SHIFT;
        double rr = fma(fn, -6.93147180559945286227e-01, x[j]);
        r[j] = fma(fn, -2.31904681384629955842e-17, rr);
        p[j] = 1.6059043836821613e-10;
    }
    // A REAL loop over the Horner steps (not unrolled): ptxas otherwise re-serialises the unrolled code into one
    // 13-deep dependent chain per value, which leaves the FP64 pipe idle for most of each instruction's latency.
    // Two steps per trip; the next pair of coefficients is fetched (LDCU) before the current pair is used, so the
    // constant-bank latency hides behind 2 N DFMAs.
    double c0 = EXP_TAYLOR[0], c1 = EXP_TAYLOR[1];
#pragma unroll 1
    for (int c = 2; c <= 12; c += 2) {
        const double n0 = EXP_TAYLOR[c], n1 = EXP_TAYLOR[c + 1];
#pragma unroll
        for (int j = 0; j < N; ++j) p[j] = fma(p[j], r[j], c0);
#pragma unroll
        for (int j = 0; j < N; ++j) p[j] = fma(p[j], r[j], c1);
        c0 = n0;
        c1 = n1;
    }
#pragma unroll
    for (int j = 0; j < N; ++j) p[j] = fma(p[j], r[j], c0);  // c0 = EXP_TAYLOR[12] = 1/0!
#pragma unroll
    for (int j = 0; j < N; ++j) {
        // x < -707.7 decided on the high word (x <= 0, so a larger unsigned high word is a larger magnitude);
        // 0xC0861D99 is the high word of -707.7 and the low word's 2^-20 relative slack is immaterial (both sides of
        // the threshold give a normal, correctly scaled result down to x = -708.3)
        // (a NaN argument has a high word >= 0xFFF00000 or, positive, < 0x80000000: it must NOT read as "under" -- it falls
        // through to the polynomial, whose result is NaN, as exp_nonpos and the reference's jnp.exp give)
        const unsigned hx = (unsigned)__double2hiint(x[j]);
        const bool under = hx > 0xC0861D99u && hx <= 0xFFF00000u && !(hx == 0xFFF00000u && __double2loint(x[j]) != 0);
        const int hi = __double2hiint(p[j]) + (n[j] << 20);
        res[j] = __hiloint2double(under ? 0 : hi, under ? 0 : __double2loint(p[j]));
    }
}

template <int KIND, int N>
__device__ __forceinline__ void kernel_from_q_n(const double (&q)[N], double kv, double (&out)[N]) {
    if (KIND == BOBE_KERNEL_RBF) {
        double x[N], e[N];
#pragma unroll
        for (int j = 0; j < N; ++j) x[j] = -0.5 * q[j];
        exp_nonpos_n<N>(x, e);
#pragma unroll
        for (int j = 0; j < N; ++j) out[j] = kv * e[j];
    } else {
        double r[N], x[N], e[N];
#pragma unroll
        for (int j = 0; j < N; ++j) {
            // q >= 0, so the IEEE order is the integer order: an exact `q < 1e-30` without touching the FP64 pipe
            const double qc = (__double_as_longlong(q[j]) < __double_as_longlong(1e-30)) ? 1e-30 : q[j];
            r[j] = sqrt_pos(qc);
            x[j] = -SQRT5 * r[j];
        }
        exp_nonpos_n<N>(x, e);
#pragma unroll
        for (int j = 0; j < N; ++j) {
            double poly = 1.0 + r[j] * (SQRT5 + r[j] * (5.0 / 3.0));
            out[j] = kv * poly * e[j];
        }
    }
}

// kernel value K0 AND the factor G of its lengthscale / input derivatives (dK/dlog l_j = G s_j), N values in lock step:
//   RBF: G = K0;   Matern-5/2: G = kv 5/3 (1 + sqrt5 r) exp(-sqrt5 r), 0 where the 1e-30 clamp is active
template <int KIND, int N>
__device__ __forceinline__ void kernel_and_g_from_q_n(const double (&q)[N], double kv, double (&k0)[N], double (&G)[N]) {
    if (KIND == BOBE_KERNEL_RBF) {
        double x[N], e[N];
#pragma unroll
        for (int j = 0; j < N; ++j) x[j] = -0.5 * q[j];
        exp_nonpos_n<N>(x, e);
#pragma unroll
        for (int j = 0; j < N; ++j) {
            k0[j] = kv * e[j];
            G[j] = k0[j];
        }
    } else {
        double r[N], x[N], e[N];
        bool clamped[N];
#pragma unroll
        for (int j = 0; j < N; ++j) {
            clamped[j] = __double_as_longlong(q[j]) < __double_as_longlong(1e-30);
            r[j] = sqrt_pos(clamped[j] ? 1e-30 : q[j]);
            x[j] = -SQRT5 * r[j];
        }
        exp_nonpos_n<N>(x, e);
#pragma unroll
        for (int j = 0; j < N; ++j) {
            k0[j] = kv * (1.0 + r[j] * (SQRT5 + r[j] * (5.0 / 3.0))) * e[j];
            G[j] = clamped[j] ? 0.0 : kv * (5.0 / 3.0) * (1.0 + SQRT5 * r[j]) * e[j];
        }
    }
}

}  // namespace bobe
