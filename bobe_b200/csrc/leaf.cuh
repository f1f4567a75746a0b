// Leaves of the recursive factorisation: one CTA factorises AND inverts a 64x64 or a 128x128 diagonal block
// entirely in shared memory.  The lower levels of the recursion are chains of tiny dependent steps; doing a
// whole 128-block (two 64-leaves + the six 64^3 products between them) in one launch removes ~10 launches per
// block from the critical path (see profiles/r01/README.md, "leaf fusion").
#pragma once
#include "common.cuh"

namespace bobe {

constexpr int SLD = 72;  // smem row stride (doubles) of every 64x64 operand: 16-byte aligned rows, and rows g, g+1 land
                         // in different halves of the 32 banks, so LDS.128 fragment loads are conflict-free
constexpr int LEAF_THREADS = 256;
constexpr double REFINE_RATIO = 1e3;

// 1/x on the critical path of every elimination step: MUFU.RCP64H seed (~2^-20) and two Newton steps (4 dependent DFMAs,
// ~1 ulp), instead of __drcp_rn's five DFMAs plus a denormal / overflow fix-up branch.  A zero, negative, huge or tiny pivot
// (|x| outside ~[1e-300, 1e300]) gives inf / NaN / a less accurate value: such a matrix is not a valid K anyway and ends as
// NaN with info = 1 downstream.
__device__ __forceinline__ double rcp_newton(double x) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    double e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    e = fma(-x, r, 1.0);
    return fma(r, e, r);
}

// Same seed, ONE third-order step: r = r0 (1 + e + e^2), e = 1 - x r0 (|e| ~ 2^-20, so the truncation error e^3 is far
// below the rounding): three dependent FP64 operations after the MUFU instead of four.  Used on the pivot chain of the
// recursive leaf, where each FP64 latency is paid 128 times per tile.
__device__ __forceinline__ double rcp_cubic(double x) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    const double e = fma(-x, r, 1.0);
    const double p = fma(e, e, e);
    return fma(r, p, r);
}

// ---- 64x64 Cholesky + inverse, register-blocked ----------------------------------------------------------
// Thread (ty, tx) of a 16x16 grid owns A[ty+16a][tx+16b] and W[ty+16a][tx+16b] (a, b < 4) in REGISTERS for the
// whole sweep.  Step j: the owners publish column j of A and row j of W to a double-buffered smem line (one
// __syncthreads per step), every thread derives 1/d_j from the pivot and applies the rank-1 updates
//   A[i][k] -= A[i][j] A[k][j] / a_jj   (k > j)          W[i][c] -= A[i][j] W[j][c] / a_jj   (i > j)
// (W starts as I: Gauss-Jordan on L X = I).  Columns of A / rows of W stay unscaled until the end:
// L[i][j] = A[i][j] / d_j, X[j][c] = W[j][c] / d_j.  A negative pivot gives NaN (rsqrt), never a trap.
// In: A lower triangle (upper ignored).  Out: A = L (zero upper), W = L^-1 (zero upper), invd, dd (64 each).
__device__ __forceinline__ void chol_inv_64(double* A, double* W, double* colb, double* rowb, double* invd,
                                            double* dd) {
    const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
    double ra[4][4], rw[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            int i = ty + 16 * a, k = tx + 16 * b;
            ra[a][b] = A[i * SLD + k];
            rw[a][b] = (i == k) ? 1.0 : 0.0;
        }
    for (int j = 0; j < 64; ++j) {
        const int ja = j >> 4, jx = j & 15;
        double* cb = colb + (j & 1) * 64;
        double* rb = rowb + (j & 1) * 64;
        if (tx == jx) {
#pragma unroll
            for (int a = 0; a < 4; ++a) {
                double v = ra[a][0];
                v = (ja == 1) ? ra[a][1] : v;
                v = (ja == 2) ? ra[a][2] : v;
                v = (ja == 3) ? ra[a][3] : v;
                cb[ty + 16 * a] = v;
            }
        }
        if (ty == jx) {
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                double v = rw[0][b];
                v = (ja == 1) ? rw[1][b] : v;
                v = (ja == 2) ? rw[2][b] : v;
                v = (ja == 3) ? rw[3][b] : v;
                rb[tx + 16 * b] = v;
            }
        }
        __syncthreads();
        // Only 1/a_jj sits on the critical path of the sweep (the updates need A[i][j] A[k][j] / a_jj); the square
        // roots are taken for all 64 pivots at once after the loop.
        const double ajj = cb[j];
        if (tid == 0) dd[j] = ajj;  // pivot, turned into d_j below
        const double w = rcp_newton(ajj);
        double ai[4], ak[4], wj[4];
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            int i = ty + 16 * a;
            ai[a] = (i > j) ? cb[i] * w : 0.0;
            int k = tx + 16 * a;
            ak[a] = (k > j) ? cb[k] : 0.0;
            wj[a] = rb[k];  // zero beyond column j (W stays lower triangular)
        }
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                ra[a][b] = fma(-ai[a], ak[b], ra[a][b]);
                rw[a][b] = fma(-ai[a], wj[b], rw[a][b]);
            }
    }
    __syncthreads();
    if (tid < 64) {  // d = sqrt(pivot) and 1/d to (near) correct rounding: one rsqrt + one Newton step each; a
        const double piv = dd[tid];  // negative pivot gives NaN, a zero pivot inf -> NaN downstream, never a trap
        double r0 = rsqrt(piv);
        double d0 = piv * r0;
        double d = fma(fma(-d0, d0, piv), 0.5 * r0, d0);
        invd[tid] = fma(fma(-d, r0, 1.0), r0, r0);
        dd[tid] = d;
    }
    __syncthreads();
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            int i = ty + 16 * a, k = tx + 16 * b;
            A[i * SLD + k] = (k < i) ? ra[a][b] * invd[k] : (k == i ? dd[i] : 0.0);
            W[i * SLD + k] = (k <= i) ? rw[a][b] * invd[i] : 0.0;
        }
    __syncthreads();
}

// ---- the same elimination, FOUR columns per barrier ----------------------------------------------------------------------
// chol_inv_64 pays a publish -> barrier -> reciprocal -> update chain of ~850 cycles for every column.  Here the owners
// publish the four columns j0 .. j0+3 of A (and the four rows of W) as they are BEFORE the panel, and every thread replays
// the three intra-panel eliminations on the few published entries it needs (the 4 x 4 diagonal block, the panel entries of
// its four rows and of its four columns, the panel rows of W at its four columns) before applying the four rank-1 updates
// to its own 4 x 4 blocks.  Every element sees exactly the operations of chol_inv_64 in the same order, so the results are
// bitwise identical; the number of block-wide barriers drops from 64 to 16 and the reciprocals of a panel no longer wait
// for a round trip through shared memory.
__device__ __forceinline__ void chol_inv_64_panel4(double* A, double* W, double* colb, double* rowb, double* invd,
                                                   double* dd) {
    const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
    double ra[4][4], rw[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            int i = ty + 16 * a, k = tx + 16 * b;
            ra[a][b] = A[i * SLD + k];
            rw[a][b] = (i == k) ? 1.0 : 0.0;
        }
    // colb / rowb: [2][4][64] (double-buffered by panel parity)
    for (int j0 = 0; j0 < 64; j0 += 4) {
        const int ja = j0 >> 4, jx0 = j0 & 15;
        double* cb = colb + ((j0 >> 2) & 1) * 256;
        double* rb = rowb + ((j0 >> 2) & 1) * 256;
        if (tx >= jx0 && tx < jx0 + 4) {  // my column tx + 16 ja is panel column tx - jx0
            double* dst = cb + (tx - jx0) * 64;
#pragma unroll
            for (int a = 0; a < 4; ++a) {
                double v = ra[a][0];
                v = (ja == 1) ? ra[a][1] : v;
                v = (ja == 2) ? ra[a][2] : v;
                v = (ja == 3) ? ra[a][3] : v;
                dst[ty + 16 * a] = v;
            }
        }
        if (ty >= jx0 && ty < jx0 + 4) {  // my row ty + 16 ja is panel row ty - jx0
            double* dst = rb + (ty - jx0) * 64;
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                double v = rw[0][b];
                v = (ja == 1) ? rw[1][b] : v;
                v = (ja == 2) ? rw[2][b] : v;
                v = (ja == 3) ? rw[3][b] : v;
                dst[tx + 16 * b] = v;
            }
        }
        __syncthreads();
        double D[4][4], Ri[4][4], Ck[4][4], Wp[4][4];  // [.][jj]: panel column jj;  Wp[jj][b]: panel row jj at my column b
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
#pragma unroll
            for (int m = 0; m < 4; ++m) {
                D[m][jj] = cb[jj * 64 + j0 + m];
                Ri[m][jj] = cb[jj * 64 + ty + 16 * m];
                Ck[m][jj] = cb[jj * 64 + tx + 16 * m];
                Wp[jj][m] = rb[jj * 64 + tx + 16 * m];
            }
        }
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
            const int j = j0 + jj;
            const double ajj = D[jj][jj];
            if (tid == 0) dd[j] = ajj;
            const double w = rcp_newton(ajj);
            double ai[4], ak[4], wj[4], dm[4];
#pragma unroll
            for (int m = 0; m < 4; ++m) {
                ai[m] = (ty + 16 * m > j) ? Ri[m][jj] * w : 0.0;
                ak[m] = (tx + 16 * m > j) ? Ck[m][jj] : 0.0;
                wj[m] = Wp[jj][m];
                dm[m] = (m > jj) ? D[m][jj] * w : 0.0;  // multiplier of panel row j0 + m
            }
            // replay column j on the rest of the panel (the entries other threads will publish as columns j+1 ..)
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
                if (kk <= jj) continue;
                const double dk = D[kk][jj];  // A[j0 + kk][j], k > j
#pragma unroll
                for (int m = 0; m < 4; ++m) {
                    Ri[m][kk] = fma(-ai[m], dk, Ri[m][kk]);
                    const double ci = (tx + 16 * m > j) ? Ck[m][jj] * w : 0.0;  // as a ROW entry of the element (k_b, j0+kk)
                    Ck[m][kk] = fma(-ci, dk, Ck[m][kk]);
                    if (m > jj) D[m][kk] = fma(-dm[m], dk, D[m][kk]);
                }
            }
#pragma unroll
            for (int mm = 0; mm < 4; ++mm) {
                if (mm <= jj) continue;
#pragma unroll
                for (int b = 0; b < 4; ++b) Wp[mm][b] = fma(-dm[mm], wj[b], Wp[mm][b]);
            }
            // my own blocks
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                    ra[a][b] = fma(-ai[a], ak[b], ra[a][b]);
                    rw[a][b] = fma(-ai[a], wj[b], rw[a][b]);
                }
        }
    }
    __syncthreads();
    if (tid < 64) {
        const double piv = dd[tid];
        double r0 = rsqrt(piv);
        double d0 = piv * r0;
        double d = fma(fma(-d0, d0, piv), 0.5 * r0, d0);
        invd[tid] = fma(fma(-d, r0, 1.0), r0, r0);
        dd[tid] = d;
    }
    __syncthreads();
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            int i = ty + 16 * a, k = tx + 16 * b;
            A[i * SLD + k] = (k < i) ? ra[a][b] * invd[k] : (k == i ? dd[i] : 0.0);
            W[i * SLD + k] = (k <= i) ? rw[a][b] * invd[i] : 0.0;
        }
    __syncthreads();
}

// ---- 64x64x64 product on smem operands with DMMA: C = alpha * A * op(B) + D -------------------------------
// op(B)[k][j] = B[j][k] (BT, "NT") or B[k][j] (NN).  D may be null or alias C (each element is read and
// written by the same thread).  8 warps: warp w owns rows 16*(w&3).., columns 32*(w>>2)...
template <bool BT>
__device__ __forceinline__ void mma64(const double* A, const double* B, double alpha, double* C, const double* D) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
    const int r0 = (warp & 3) * 16, c0 = (warp >> 2) * 32;
    double acc[2][4][2];
#pragma unroll
    for (int mf = 0; mf < 2; ++mf)
#pragma unroll
        for (int nf = 0; nf < 4; ++nf) acc[mf][nf][0] = acc[mf][nf][1] = 0.0;
#pragma unroll 2
    for (int k0 = 0; k0 < 64; k0 += 8) {
        double2 a[2], b[4];
#pragma unroll
        for (int mf = 0; mf < 2; ++mf)
            a[mf] = *reinterpret_cast<const double2*>(A + (r0 + mf * 8 + g) * SLD + k0 + 2 * t);
#pragma unroll
        for (int nf = 0; nf < 4; ++nf) {
            if (BT) {
                b[nf] = *reinterpret_cast<const double2*>(B + (c0 + nf * 8 + g) * SLD + k0 + 2 * t);
            } else {
                b[nf].x = B[(k0 + 2 * t) * SLD + c0 + nf * 8 + g];
                b[nf].y = B[(k0 + 2 * t + 1) * SLD + c0 + nf * 8 + g];
            }
        }
#pragma unroll
        for (int mf = 0; mf < 2; ++mf)
#pragma unroll
            for (int nf = 0; nf < 4; ++nf) dmma884(acc[mf][nf][0], acc[mf][nf][1], a[mf].x, b[nf].x);
#pragma unroll
        for (int mf = 0; mf < 2; ++mf)
#pragma unroll
            for (int nf = 0; nf < 4; ++nf) dmma884(acc[mf][nf][0], acc[mf][nf][1], a[mf].y, b[nf].y);
    }
#pragma unroll
    for (int mf = 0; mf < 2; ++mf)
#pragma unroll
        for (int nf = 0; nf < 4; ++nf) {
            int off = (r0 + mf * 8 + g) * SLD + c0 + nf * 8 + 2 * t;
            double v0 = alpha * acc[mf][nf][0], v1 = alpha * acc[mf][nf][1];
            if (D) {
                double2 old = *reinterpret_cast<const double2*>(D + off);
                v0 += old.x;
                v1 += old.y;
            }
            *reinterpret_cast<double2*>(C + off) = make_double2(v0, v1);
        }
}

// development aid (tools/leaf_bench.cu): per-phase clock stamps of CTA 0, compiled in only with -DBOBE_LEAF_TIMING
#ifdef BOBE_LEAF_TIMING
#define LEAF_STAMP(io, i)                                                       \
    do {                                                                        \
        if ((io).stamps && threadIdx.x == 0 && blockIdx.z == 0) (io).stamps[i] = clock64(); \
    } while (0)
#else
#define LEAF_STAMP(io, i) \
    do {                  \
    } while (0)
#endif

struct LeafIO {
    const double* KB;
    double *L, *Lt, *Linv, *U, *diag, *dstat;
    int* gate;
    int npad, o;  // o = first row/column of the block
    int panel4;   // 1: chol_inv_64_panel4 (four columns per barrier), 0: chol_inv_64
    // gate: state BEFORE this leaf; gate_out: snapshot after it (what the products of THIS tile column read -- they may
    // run while later leaves already update the state); gate_final: latest state.  All three may be the same array.
    int* gate_out;
    int* gate_final;
    // tile leaves only.  bit 0: also write the zero 64-blocks on the other side of the diagonal (needed when a product may
    // read a whole 128 x 128 diagonal tile, i.e. with the 128-row tile configurations); bit 1: also write U = X^T
    // (otherwise tile_transpose_kernel does it off the critical path)
    int flags;
#ifdef BOBE_LEAF_TIMING
    long long* stamps;
#endif
};
__device__ __forceinline__ void chol_inv_64_any(int panel4, double* A, double* W, double* colb, double* rowb, double* invd,
                                                double* dd) {
    if (panel4)
        chol_inv_64_panel4(A, W, colb, rowb, invd, dd);
    else
        chol_inv_64(A, W, colb, rowb, invd, dd);
}

// global -> smem: 64x64 block at (row r0, col c0) of KB, as 16-byte cp.async copies (all blocks of a leaf are issued
// back to back and waited for once, so the leaf pays ONE global-load latency instead of one per element batch).
// The strict upper part of a diagonal block is copied as it is: chol_inv_64 never reads it.
__device__ __forceinline__ void leaf_load_async(double* S, const double* Kz, int npad, int r0, int c0) {
    for (int idx = threadIdx.x; idx < 64 * 32; idx += LEAF_THREADS) {
        int r = idx >> 5, c = (idx & 31) * 2;
        cp_async16(S + r * SLD + c, Kz + (int64_t)(r0 + r) * npad + c0 + c, true);
    }
}
// smem block -> global at (r0, c0), plus its transpose into Gt at (c0, r0).  The transpose goes through a staging
// buffer T with an odd row stride (65): reading S down a column directly would put all 32 lanes on two banks.
constexpr int TLD = 65;
__device__ __forceinline__ void leaf_store(const double* S, double* T, double* G, double* Gt, int npad, int r0, int c0) {
    for (int idx = threadIdx.x; idx < 64 * 32; idx += LEAF_THREADS) {  // direct copy, 16 bytes per lane
        int r = idx >> 5, c = (idx & 31) * 2;
        double2 v = *reinterpret_cast<const double2*>(S + r * SLD + c);
        *reinterpret_cast<double2*>(G + (int64_t)(r0 + r) * npad + c0 + c) = v;
    }
    for (int idx = threadIdx.x; idx < 64 * 64; idx += LEAF_THREADS) {  // T[c][r] = S[r][c], lanes along c
        int r = idx >> 6, c = idx & 63;
        T[c * TLD + r] = S[r * SLD + c];
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < 64 * 64; idx += LEAF_THREADS) {
        int r = idx >> 6, c = idx & 63;
        Gt[(int64_t)(c0 + r) * npad + r0 + c] = T[r * TLD + c];
    }
    __syncthreads();  // T is reused by the next block
}

// running extreme pivots of this matrix (leaves of one matrix run in stream order: no atomics needed).
// max/min pivot is a lower bound on cond(L); beyond REFINE_RATIO the panel solves get a correction step.
__device__ __forceinline__ bool leaf_update_stats(const LeafIO& io, const double* dd, int count, int64_t z,
                                                  bool second = false) {
    __shared__ int s_gate;
    if (threadIdx.x < 32) {  // warp 0: min / max of the pivots by shuffles (fmin / fmax skip NaN pivots, as a serial scan would)
        double lo = 1e300, hi = 0.0;
        for (int i = threadIdx.x; i < count; i += 32) {
            lo = fmin(lo, dd[i]);
            hi = fmax(hi, dd[i]);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo = fmin(lo, __shfl_xor_sync(0xffffffffu, lo, o));
            hi = fmax(hi, __shfl_xor_sync(0xffffffffu, hi, o));
        }
        if (threadIdx.x == 0) {
            lo = fmin(lo, io.dstat[z * 2]);
            hi = fmax(hi, io.dstat[z * 2 + 1]);
            io.dstat[z * 2] = lo;
            io.dstat[z * 2 + 1] = hi;
            int gte = second ? io.gate_out[z] : io.gate[z];
            if (hi > REFINE_RATIO * lo) gte = 1;
            io.gate_out[z] = gte;
            io.gate_final[z] = gte;
            s_gate = gte;
        }
    }
    __syncthreads();
    return s_gate != 0;
}

__global__ void __launch_bounds__(LEAF_THREADS) leaf64_kernel(LeafIO io) {
    extern __shared__ __align__(16) double sm[];
    pdl_wait();
    pdl_trigger();
    double* A = sm;
    double* W = A + 64 * SLD;
    double* colb = W + 64 * SLD;  // [2][4][64]
    double* rowb = colb + 512;    // [2][4][64]
    double* invd = rowb + 512;    // [64]
    double* dd = invd + 64;       // [64]
    const int64_t z = blockIdx.z, zoff = z * (int64_t)io.npad * io.npad;
    double* T = dd + 64;          // [64][TLD] transpose staging
    leaf_load_async(A, io.KB + zoff, io.npad, io.o, io.o);
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();
    chol_inv_64_any(io.panel4, A, W, colb, rowb, invd, dd);
    leaf_store(A, T, io.L + zoff, io.Lt + zoff, io.npad, io.o, io.o);
    leaf_store(W, T, io.Linv + zoff, io.U + zoff, io.npad, io.o, io.o);
    if (threadIdx.x < 64) io.diag[z * io.npad + io.o + threadIdx.x] = dd[threadIdx.x];
    leaf_update_stats(io, dd, 64, z);
}
constexpr int LEAF64_SMEM = (2 * 64 * SLD + 2 * 512 + 2 * 64 + 64 * TLD) * 8;

// 128x128 block = [A11 .; A21 A22]:
//   (L11, X11) = chol_inv(A11);  L21 = A21 X11^T [+ gated correction];  A22 -= L21 L21^T;
//   (L22, X22) = chol_inv(A22);  X21 = -(X22 L21) X11
__global__ void __launch_bounds__(LEAF_THREADS) leaf128_kernel(LeafIO io) {
    extern __shared__ __align__(16) double sm[];
    pdl_wait();
    pdl_trigger();
    double* B0 = sm;              // A11 -> L11
    double* B1 = B0 + 64 * SLD;   // X11
    double* B2 = B1 + 64 * SLD;   // A21 -> residual -> T
    double* B3 = B2 + 64 * SLD;   // A22 -> L22
    double* B4 = B3 + 64 * SLD;   // X22
    double* B5 = B4 + 64 * SLD;   // L21 -> X21
    double* colb = B5 + 64 * SLD;  // [2][4][64]
    double* rowb = colb + 512;     // [2][4][64]
    double* invd = rowb + 512;     // [128]
    double* dd = invd + 128;      // [128]
    const int64_t z = blockIdx.z, zoff = z * (int64_t)io.npad * io.npad;
    const int o = io.o, npad = io.npad;
    const double* Kz = io.KB + zoff;
    leaf_load_async(B0, Kz, npad, o, o);
    leaf_load_async(B2, Kz, npad, o + 64, o);
    leaf_load_async(B3, Kz, npad, o + 64, o + 64);
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();
    chol_inv_64_any(io.panel4, B0, B1, colb, rowb, invd, dd);
    const bool refine = leaf_update_stats(io, dd, 64, z);
    mma64<true>(B2, B1, 1.0, B5, nullptr);  // L21 = A21 X11^T
    __syncthreads();
    if (refine) {  // same correction as the recursion applies to its panel solves (factor.cu)
        mma64<true>(B5, B0, -1.0, B2, B2);  // R = A21 - L21 L11^T
        __syncthreads();
        mma64<true>(B2, B1, 1.0, B5, B5);   // L21 += R X11^T
        __syncthreads();
    }
    mma64<true>(B5, B5, -1.0, B3, B3);  // A22 -= L21 L21^T
    __syncthreads();
    leaf_store(B5, B2, io.L + zoff, io.Lt + zoff, npad, o + 64, o);  // L21 (B2 = A21 / residual is dead: staging)
    chol_inv_64_any(io.panel4, B3, B4, colb, rowb, invd + 64, dd + 64);
    leaf_update_stats(io, dd + 64, 64, z, true);
    mma64<false>(B4, B5, 1.0, B2, nullptr);  // T = X22 L21
    __syncthreads();
    mma64<false>(B2, B1, -1.0, B5, nullptr);  // X21 = -T X11
    __syncthreads();
    // B2 (T) is dead from here on and serves as the transpose staging buffer
    leaf_store(B0, B2, io.L + zoff, io.Lt + zoff, npad, o, o);
    leaf_store(B3, B2, io.L + zoff, io.Lt + zoff, npad, o + 64, o + 64);
    leaf_store(B1, B2, io.Linv + zoff, io.U + zoff, npad, o, o);
    leaf_store(B4, B2, io.Linv + zoff, io.U + zoff, npad, o + 64, o + 64);
    leaf_store(B5, B2, io.Linv + zoff, io.U + zoff, npad, o + 64, o);
    if (threadIdx.x < 128) io.diag[z * npad + o + threadIdx.x] = dd[threadIdx.x];
}
constexpr int LEAF128_SMEM = (6 * 64 * SLD + 2 * 512 + 2 * 128) * 8;

// ==== recursive leaf: 32 x 32 elimination blocks joined by DMMA products =================================================
// chol_inv_64_panel4 runs 64 rank-1 updates over the WHOLE 64 x 64 block (both triangles, finished rows included):
// ~4x the useful FP64 work, all of it on one SM, which makes the leaf throughput-bound (~20 us per 64-block).  Here the
// elimination only runs on 32 x 32 diagonal blocks (a quarter of the elements, half the steps each) and the blocks are
// joined by small tensor-core products that skip the zero halves of their triangular operands:
//   (L11, X11) = chol_inv_32(A11);  L21 = A21 X11^T (+ one correction step, always);  A22 -= L21 L21^T;
//   (L22, X22) = chol_inv_32(A22);  X21 = -(X22 L21) X11

// C(N x N) = alpha * A * op(B) + D on smem operands of row stride SLD, N = 32 or 64, 8 warps.
// op(B) = B^T (BT) or B.  TRI tells which operand is lower triangular so that whole k8 steps / fragments are skipped:
enum : int { MM_FULL = 0, MM_BT_BLOWER = 1, MM_NN_ALOWER = 2, MM_NN_BLOWER = 3, MM_SYRK_LOWER = 4 };
//   MM_BT_BLOWER   C = A B^T, B lower:  sum over k <= column          MM_NN_ALOWER  C = A B, A lower: k <= row
//   MM_NN_BLOWER   C = A B,   B lower:  sum over k >= column          MM_SYRK_LOWER C = A A^T-like, only fragments that touch
//                                                                      the lower triangle are computed (others untouched)
template <int N, bool BT, int TRI>
__device__ __forceinline__ void mma_blk(const double* A, const double* B, double alpha, double* C, const double* D) {
    constexpr int MF = N / 32, NF = N / 16;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
    const int r0 = (warp & 3) * (N / 4), c0 = (warp >> 2) * (N / 2);
    double acc[MF][NF][2];
#pragma unroll
    for (int mf = 0; mf < MF; ++mf)
#pragma unroll
        for (int nf = 0; nf < NF; ++nf) acc[mf][nf][0] = acc[mf][nf][1] = 0.0;
    // warp-uniform k range over all of this warp's fragments; per-fragment bounds inside
    int kbeg = 0, kend = N;
    if (TRI == MM_BT_BLOWER) kend = min(N, c0 + N / 2);
    if (TRI == MM_NN_ALOWER) kend = min(N, r0 + N / 4);
    if (TRI == MM_NN_BLOWER) kbeg = c0;
    if (TRI == MM_SYRK_LOWER && c0 > r0 + N / 4 - 1) kend = 0;
    for (int k0 = kbeg; k0 < kend; k0 += 8) {
        double2 a[MF], b[NF];
#pragma unroll
        for (int mf = 0; mf < MF; ++mf)
            a[mf] = *reinterpret_cast<const double2*>(A + (r0 + mf * 8 + g) * SLD + k0 + 2 * t);
#pragma unroll
        for (int nf = 0; nf < NF; ++nf) {
            if (BT) {
                b[nf] = *reinterpret_cast<const double2*>(B + (c0 + nf * 8 + g) * SLD + k0 + 2 * t);
            } else {
                b[nf].x = B[(k0 + 2 * t) * SLD + c0 + nf * 8 + g];
                b[nf].y = B[(k0 + 2 * t + 1) * SLD + c0 + nf * 8 + g];
            }
        }
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
            for (int mf = 0; mf < MF; ++mf)
#pragma unroll
                for (int nf = 0; nf < NF; ++nf) {
                    bool live = true;  // warp-uniform
                    if (TRI == MM_BT_BLOWER) live = k0 <= c0 + nf * 8 + 7;
                    if (TRI == MM_NN_ALOWER) live = k0 <= r0 + mf * 8 + 7;
                    if (TRI == MM_NN_BLOWER) live = k0 + 7 >= c0 + nf * 8;
                    if (TRI == MM_SYRK_LOWER) live = c0 + nf * 8 <= r0 + mf * 8 + 7;
                    if (live) dmma884(acc[mf][nf][0], acc[mf][nf][1], h ? a[mf].y : a[mf].x, h ? b[nf].y : b[nf].x);
                }
    }
#pragma unroll
    for (int mf = 0; mf < MF; ++mf)
#pragma unroll
        for (int nf = 0; nf < NF; ++nf) {
            if (TRI == MM_SYRK_LOWER && c0 + nf * 8 > r0 + mf * 8 + 7) continue;
            int off = (r0 + mf * 8 + g) * SLD + c0 + nf * 8 + 2 * t;
            double v0 = alpha * acc[mf][nf][0], v1 = alpha * acc[mf][nf][1];
            if (D) {
                double2 old = *reinterpret_cast<const double2*>(D + off);
                v0 += old.x;
                v1 += old.y;
            }
            *reinterpret_cast<double2*>(C + off) = make_double2(v0, v1);
        }
}

// chol_inv_64_panel4 generalised to an NBk x NBk block (NBk = 32 or 64; identical operations for 64): thread (ty, tx) of
// the 16 x 16 grid owns the E x E elements (ty + 16 a, tx + 16 b), E = NBk / 16.  A and W have row stride SLD.
// colb / rowb: [2][4][NBk] each.  invd / dd: [NBk].
template <int NBk>
__device__ __forceinline__ void chol_inv_blk(double* A, double* W, double* colb, double* rowb, double* invd, double* dd) {
    constexpr int E = NBk / 16;
    const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
    double ra[E][E], rw[E][E];
#pragma unroll
    for (int a = 0; a < E; ++a)
#pragma unroll
        for (int b = 0; b < E; ++b) {
            int i = ty + 16 * a, k = tx + 16 * b;
            ra[a][b] = A[i * SLD + k];
            rw[a][b] = (i == k) ? 1.0 : 0.0;
        }
    for (int j0 = 0; j0 < NBk; j0 += 4) {
        const int ja = j0 >> 4, jx0 = j0 & 15;
        double* cb = colb + ((j0 >> 2) & 1) * (4 * NBk);
        double* rb = rowb + ((j0 >> 2) & 1) * (4 * NBk);
        if (tx >= jx0 && tx < jx0 + 4) {
            double* dst = cb + (tx - jx0) * NBk;
#pragma unroll
            for (int a = 0; a < E; ++a) {
                double v = ra[a][0];
#pragma unroll
                for (int q = 1; q < E; ++q) v = (ja == q) ? ra[a][q] : v;
                dst[ty + 16 * a] = v;
            }
        }
        if (ty >= jx0 && ty < jx0 + 4) {
            double* dst = rb + (ty - jx0) * NBk;
#pragma unroll
            for (int b = 0; b < E; ++b) {
                double v = rw[0][b];
#pragma unroll
                for (int q = 1; q < E; ++q) v = (ja == q) ? rw[q][b] : v;
                dst[tx + 16 * b] = v;
            }
        }
        __syncthreads();
        double D[4][4], Ri[E][4], Ck[E][4], Wp[4][E];
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
#pragma unroll
            for (int p = 0; p < 4; ++p) D[p][jj] = cb[jj * NBk + j0 + p];
#pragma unroll
            for (int e = 0; e < E; ++e) {
                Ri[e][jj] = cb[jj * NBk + ty + 16 * e];
                Ck[e][jj] = cb[jj * NBk + tx + 16 * e];
                Wp[jj][e] = rb[jj * NBk + tx + 16 * e];
            }
        }
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
            const int j = j0 + jj;
            const double ajj = D[jj][jj];
            if (tid == 0) dd[j] = ajj;
            const double w = rcp_cubic(ajj);
            double ai[E], ak[E], wj[E], dm[4];
#pragma unroll
            for (int e = 0; e < E; ++e) {
                ai[e] = (ty + 16 * e > j) ? Ri[e][jj] * w : 0.0;
                ak[e] = (tx + 16 * e > j) ? Ck[e][jj] : 0.0;
                wj[e] = Wp[jj][e];
            }
#pragma unroll
            for (int p = 0; p < 4; ++p) dm[p] = (p > jj) ? D[p][jj] * w : 0.0;
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
                if (kk <= jj) continue;
                const double dk = D[kk][jj];
#pragma unroll
                for (int e = 0; e < E; ++e) {
                    Ri[e][kk] = fma(-ai[e], dk, Ri[e][kk]);
                    const double ci = (tx + 16 * e > j) ? Ck[e][jj] * w : 0.0;
                    Ck[e][kk] = fma(-ci, dk, Ck[e][kk]);
                }
                // pivot block: the product of the two column entries does not depend on w, so only ONE operation (the fma
                // with w) sits between the reciprocal and the next pivot
#pragma unroll
                for (int p = 0; p < 4; ++p)
                    if (p > jj) D[p][kk] = fma(-(D[p][jj] * dk), w, D[p][kk]);
            }
#pragma unroll
            for (int pp = 0; pp < 4; ++pp) {
                if (pp <= jj) continue;
#pragma unroll
                for (int e = 0; e < E; ++e) Wp[pp][e] = fma(-dm[pp], wj[e], Wp[pp][e]);
            }
#pragma unroll
            for (int a = 0; a < E; ++a)
#pragma unroll
                for (int b = 0; b < E; ++b) {
                    ra[a][b] = fma(-ai[a], ak[b], ra[a][b]);
                    rw[a][b] = fma(-ai[a], wj[b], rw[a][b]);
                }
        }
    }
    __syncthreads();
    if (tid < NBk) {
        const double piv = dd[tid];
        double r0 = rsqrt(piv);
        double d0 = piv * r0;
        double d = fma(fma(-d0, d0, piv), 0.5 * r0, d0);
        invd[tid] = fma(fma(-d, r0, 1.0), r0, r0);
        dd[tid] = d;
    }
    __syncthreads();
#pragma unroll
    for (int a = 0; a < E; ++a)
#pragma unroll
        for (int b = 0; b < E; ++b) {
            int i = ty + 16 * a, k = tx + 16 * b;
            A[i * SLD + k] = (k < i) ? ra[a][b] * invd[k] : (k == i ? dd[i] : 0.0);
            W[i * SLD + k] = (k <= i) ? rw[a][b] * invd[i] : 0.0;
        }
    __syncthreads();
}

// 64 x 64 block by two 32 x 32 eliminations.  A -> L (zero upper), W -> L^-1 (zero upper); invd / dd: [64].
__device__ __forceinline__ void chol_inv_64_rec(double* A, double* W, double* colb, double* rowb, double* invd, double* dd) {
    double* A10 = A + 32 * SLD;        // rows 32.., cols 0..31
    double* A11 = A + 32 * SLD + 32;
    double* W10 = W + 32 * SLD;
    double* W11 = W + 32 * SLD + 32;
    double* S1 = W + 32;               // rows 0..31, cols 32..63 of W: scratch, zero again at the end
    chol_inv_blk<32>(A, W, colb, rowb, invd, dd);
    mma_blk<32, true, MM_BT_BLOWER>(A10, W, 1.0, S1, nullptr);   // L21 = A21 X11^T
    __syncthreads();
    mma_blk<32, true, MM_BT_BLOWER>(S1, A, -1.0, A10, A10);      // R = A21 - L21 L11^T   (in place of A21)
    __syncthreads();
    mma_blk<32, true, MM_BT_BLOWER>(A10, W, 1.0, S1, S1);        // L21 += R X11^T
    __syncthreads();
    mma_blk<32, true, MM_SYRK_LOWER>(S1, S1, -1.0, A11, A11);    // A22 -= L21 L21^T (lower fragments)
    for (int idx = threadIdx.x; idx < 32 * 32; idx += LEAF_THREADS) {  // L21 to its place (R is dead)
        int r = idx >> 5, c = idx & 31;
        A10[r * SLD + c] = S1[r * SLD + c];
    }
    __syncthreads();
    chol_inv_blk<32>(A11, W11, colb, rowb, invd + 32, dd + 32);
    mma_blk<32, false, MM_NN_ALOWER>(W11, A10, 1.0, S1, nullptr);  // T = X22 L21
    __syncthreads();
    mma_blk<32, false, MM_NN_BLOWER>(S1, W, -1.0, W10, nullptr);   // X21 = -T X11
    __syncthreads();
    for (int idx = threadIdx.x; idx < 32 * 32; idx += LEAF_THREADS) {
        int r = idx >> 5, c = idx & 31;
        S1[r * SLD + c] = 0.0;
        A[r * SLD + 32 + c] = 0.0;  // upper-right block of L (the eliminations only clear inside their own blocks)
    }
    __syncthreads();
}

// ==== leaves of the tile-column factorisation (factor_tiled.cu) =====================================================
// Same arithmetic as leaf64_kernel / leaf128_kernel; the outputs differ: only L (no transpose of L is kept by the tiled
// scheme), Linv and U = Linv^T, and ALL four 64-blocks of the tile are written (zeros in the block on the other side of
// the diagonal), so that no separate zero-fill pass is needed for anything a product may read inside a diagonal tile.
__device__ __forceinline__ void tile_store_direct(const double* S, double* G, int npad, int r0, int c0) {
    for (int idx = threadIdx.x; idx < 64 * 32; idx += LEAF_THREADS) {
        int r = idx >> 5, c = (idx & 31) * 2;
        *reinterpret_cast<double2*>(G + (int64_t)(r0 + r) * npad + c0 + c) = *reinterpret_cast<const double2*>(S + r * SLD + c);
    }
}
__device__ __forceinline__ void tile_store_zero(double* G, int npad, int r0, int c0) {
    for (int idx = threadIdx.x; idx < 64 * 32; idx += LEAF_THREADS) {
        int r = idx >> 5, c = (idx & 31) * 2;
        *reinterpret_cast<double2*>(G + (int64_t)(r0 + r) * npad + c0 + c) = make_double2(0.0, 0.0);
    }
}
// Gt[c0 + c][r0 + r] = S[r][c] through the odd-stride staging buffer T (see leaf_store)
__device__ __forceinline__ void tile_store_transposed(const double* S, double* T, double* Gt, int npad, int r0, int c0) {
    for (int idx = threadIdx.x; idx < 64 * 64; idx += LEAF_THREADS) {
        int r = idx >> 6, c = idx & 63;
        T[c * TLD + r] = S[r * SLD + c];
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < 64 * 64; idx += LEAF_THREADS) {
        int r = idx >> 6, c = idx & 63;
        Gt[(int64_t)(c0 + r) * npad + r0 + c] = T[r * TLD + c];
    }
    __syncthreads();
}

template <bool REC>
__device__ __forceinline__ void chol_inv_64_sel(int panel4, double* A, double* W, double* colb, double* rowb, double* invd,
                                                double* dd) {
    if (REC)
        chol_inv_64_rec(A, W, colb, rowb, invd, dd);
    else
        chol_inv_64_any(panel4, A, W, colb, rowb, invd, dd);
}

// REC: recursive 32-base elimination + triangular-aware products (default); otherwise the round-1 arithmetic
template <bool REC>
__global__ void __launch_bounds__(LEAF_THREADS) tile_leaf64_kernel(LeafIO io) {
    extern __shared__ __align__(16) double sm[];
    pdl_wait();
    pdl_trigger();
    double* A = sm;
    double* W = A + 64 * SLD;
    double* colb = W + 64 * SLD;
    double* rowb = colb + 512;
    double* invd = rowb + 512;
    double* dd = invd + 64;
    double* T = dd + 64;
    const int64_t z = blockIdx.z, zoff = z * (int64_t)io.npad * io.npad;
    leaf_load_async(A, io.KB + zoff, io.npad, io.o, io.o);
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();
    chol_inv_64_sel<REC>(io.panel4, A, W, colb, rowb, invd, dd);
    tile_store_direct(A, io.L + zoff, io.npad, io.o, io.o);
    tile_store_direct(W, io.Linv + zoff, io.npad, io.o, io.o);
    if (io.flags & 2) tile_store_transposed(W, T, io.U + zoff, io.npad, io.o, io.o);
    if (threadIdx.x < 64) io.diag[z * io.npad + io.o + threadIdx.x] = dd[threadIdx.x];
    leaf_update_stats(io, dd, 64, z);
}

template <bool REC>
__global__ void __launch_bounds__(LEAF_THREADS) tile_leaf128_kernel(LeafIO io) {
    extern __shared__ __align__(16) double sm[];
    pdl_wait();
    pdl_trigger();
    double* B0 = sm;              // A11 -> L11
    double* B1 = B0 + 64 * SLD;   // X11
    double* B2 = B1 + 64 * SLD;   // A21 -> residual -> T
    double* B3 = B2 + 64 * SLD;   // A22 -> L22
    double* B4 = B3 + 64 * SLD;   // X22
    double* B5 = B4 + 64 * SLD;   // L21 -> X21
    double* colb = B5 + 64 * SLD;
    double* rowb = colb + 512;
    double* invd = rowb + 512;
    double* dd = invd + 128;
    const int64_t z = blockIdx.z, zoff = z * (int64_t)io.npad * io.npad;
    const int o = io.o, npad = io.npad;
    const double* Kz = io.KB + zoff;
    double *Lz = io.L + zoff, *Xz = io.Linv + zoff, *Uz = io.U + zoff;
    LEAF_STAMP(io, 0);
    leaf_load_async(B0, Kz, npad, o, o);
    leaf_load_async(B2, Kz, npad, o + 64, o);
    leaf_load_async(B3, Kz, npad, o + 64, o + 64);
    cp_async_commit();
    if (io.flags & 1) {  // the blocks on the other side of the diagonal, while the loads are in flight
        tile_store_zero(Lz, npad, o, o + 64);
        tile_store_zero(Xz, npad, o, o + 64);
        tile_store_zero(Uz, npad, o + 64, o);
    }
    cp_async_wait<0>();
    __syncthreads();
    LEAF_STAMP(io, 1);
    chol_inv_64_sel<REC>(io.panel4, B0, B1, colb, rowb, invd, dd);
    LEAF_STAMP(io, 2);
    const bool refine = leaf_update_stats(io, dd, 64, z);
    if (REC) {
        mma_blk<64, true, MM_BT_BLOWER>(B2, B1, 1.0, B5, nullptr);  // L21 = A21 X11^T
        __syncthreads();
        if (refine) {
            mma_blk<64, true, MM_BT_BLOWER>(B5, B0, -1.0, B2, B2);  // R = A21 - L21 L11^T
            __syncthreads();
            mma_blk<64, true, MM_BT_BLOWER>(B2, B1, 1.0, B5, B5);   // L21 += R X11^T
            __syncthreads();
        }
        mma_blk<64, true, MM_SYRK_LOWER>(B5, B5, -1.0, B3, B3);     // A22 -= L21 L21^T
    } else {
        mma64<true>(B2, B1, 1.0, B5, nullptr);
        __syncthreads();
        if (refine) {
            mma64<true>(B5, B0, -1.0, B2, B2);
            __syncthreads();
            mma64<true>(B2, B1, 1.0, B5, B5);
            __syncthreads();
        }
        mma64<true>(B5, B5, -1.0, B3, B3);
    }
    __syncthreads();
    LEAF_STAMP(io, 3);
    tile_store_direct(B5, Lz, npad, o + 64, o);  // L21
    tile_store_direct(B0, Lz, npad, o, o);       // L11
    LEAF_STAMP(io, 4);
    chol_inv_64_sel<REC>(io.panel4, B3, B4, colb, rowb, invd + 64, dd + 64);
    LEAF_STAMP(io, 5);
    leaf_update_stats(io, dd + 64, 64, z, true);
    if (REC) {
        mma_blk<64, false, MM_NN_ALOWER>(B4, B5, 1.0, B2, nullptr);   // T = X22 L21
        __syncthreads();
        mma_blk<64, false, MM_NN_BLOWER>(B2, B1, -1.0, B5, nullptr);  // X21 = -T X11
    } else {
        mma64<false>(B4, B5, 1.0, B2, nullptr);
        __syncthreads();
        mma64<false>(B2, B1, -1.0, B5, nullptr);
    }
    __syncthreads();
    LEAF_STAMP(io, 6);
    tile_store_direct(B3, Lz, npad, o + 64, o + 64);
    tile_store_direct(B1, Xz, npad, o, o);
    tile_store_direct(B4, Xz, npad, o + 64, o + 64);
    tile_store_direct(B5, Xz, npad, o + 64, o);
    LEAF_STAMP(io, 7);
    if (io.flags & 2) {  // B2 (T) is dead: transpose staging
        tile_store_transposed(B1, B2, Uz, npad, o, o);
        tile_store_transposed(B4, B2, Uz, npad, o + 64, o + 64);
        tile_store_transposed(B5, B2, Uz, npad, o + 64, o);
    }
    if (threadIdx.x < 128) io.diag[z * npad + o + threadIdx.x] = dd[threadIdx.x];
    LEAF_STAMP(io, 8);
}

// U tile = (X tile)^T for the 64-blocks on and below the diagonal of tile column `o` (w = 64 or 128 wide): what the
// leaves leave out when LeafIO::flags bit 1 is clear.  grid = (w / 32, w / 32, batch), 256 threads (32 x 8).
__global__ void __launch_bounds__(256) tile_transpose_kernel(const double* __restrict__ X, double* __restrict__ U, int npad,
                                                             int o) {
    __shared__ double t[32][33];
    pdl_wait();
    pdl_trigger();
    const int br = blockIdx.y, bc = blockIdx.x;  // 32-blocks: rows br, columns bc of the tile
    if ((bc >> 1) > (br >> 1)) return;           // 64-block above the diagonal: not part of X
    const int64_t zoff = (int64_t)blockIdx.z * npad * npad;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int r = ty; r < 32; r += 8) t[r][tx] = X[zoff + (int64_t)(o + br * 32 + r) * npad + o + bc * 32 + tx];
    __syncthreads();
    for (int r = ty; r < 32; r += 8) U[zoff + (int64_t)(o + bc * 32 + r) * npad + o + br * 32 + tx] = t[tx][r];
}

}  // namespace bobe
