// Microbenchmark of the tile leaf (development aid): phase stamps of one CTA and wall time for 1 / 8 / 148 matrices.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -DBOBE_LEAF_TIMING -o tools/leaf_bench tools/leaf_bench.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cmath>
#include "../bobe_b200/csrc/leaf.cuh"
namespace bobe {
void set_error(const char*, ...) {}
int32_t check_launch(const char*) { return 0; }
bool pdl_enabled(int64_t) { return false; }
}
using namespace bobe;

template <bool REC>
void run(int batch, int npad, const std::vector<double>& K, int panel4, const char* label, int flags = 3) {
    size_t m2 = (size_t)npad * npad;
    double *dK, *dL, *dX, *dU, *ddiag, *dstat; int* gate; long long* stamps;
    cudaMalloc(&dK, batch * m2 * 8); cudaMalloc(&dL, batch * m2 * 8); cudaMalloc(&dX, batch * m2 * 8); cudaMalloc(&dU, batch * m2 * 8);
    cudaMalloc(&ddiag, batch * npad * 8); cudaMalloc(&dstat, batch * 16); cudaMalloc(&gate, batch * 4); cudaMalloc(&stamps, 64 * 8);
    for (int b = 0; b < batch; ++b) cudaMemcpy(dK + b * m2, K.data(), m2 * 8, cudaMemcpyHostToDevice);
    std::vector<double> st(2 * batch); for (int b = 0; b < batch; ++b) { st[2*b] = 1e300; st[2*b+1] = 0; }
    cudaMemcpy(dstat, st.data(), batch * 16, cudaMemcpyHostToDevice); cudaMemset(gate, 0, batch * 4);
    LeafIO io{dK, dL, nullptr, dX, dU, ddiag, dstat, gate, npad, 0, panel4, gate, gate, flags, stamps};
    cudaFuncSetAttribute(tile_leaf128_kernel<REC>, cudaFuncAttributeMaxDynamicSharedMemorySize, LEAF128_SMEM);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int i = 0; i < 3; ++i) tile_leaf128_kernel<REC><<<dim3(1, 1, batch), LEAF_THREADS, LEAF128_SMEM>>>(io);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    const int iters = 20;
    for (int i = 0; i < iters; ++i) tile_leaf128_kernel<REC><<<dim3(1, 1, batch), LEAF_THREADS, LEAF128_SMEM>>>(io);
    cudaEventRecord(e1); cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long h[16]; cudaMemcpy(h, stamps, 9 * 8, cudaMemcpyDeviceToHost);
    printf("%-28s batch %3d: %.2f us per launch (back to back) | clocks: load %lld, chol1 %lld, stats+L21+syrk %lld, storeL %lld, chol2 %lld, "
           "T+X21 %lld, storeLX %lld, storeU %lld, total %lld\n", label, batch, ms * 1e3 / iters, h[1]-h[0], h[2]-h[1], h[3]-h[2], h[4]-h[3],
           h[5]-h[4], h[6]-h[5], h[7]-h[6], h[8]-h[7], h[8]-h[0]);
    // residual check of matrix 0: |L L^T - K| and |X L - I|
    std::vector<double> L(m2), X(m2), U(m2); cudaMemcpy(L.data(), dL, m2 * 8, cudaMemcpyDeviceToHost);
    cudaMemcpy(X.data(), dX, m2 * 8, cudaMemcpyDeviceToHost); cudaMemcpy(U.data(), dU, m2 * 8, cudaMemcpyDeviceToHost);
    double e1m = 0, e2m = 0, e3m = 0;
    for (int i = 0; i < 128; ++i) for (int j = 0; j < 128; ++j) {
        double s = 0, t = 0; for (int k = 0; k < 128; ++k) { s += L[i*npad+k] * L[j*npad+k]; t += X[i*npad+k] * L[k*npad+j]; }
        if (j <= i) e1m = fmax(e1m, fabs(s - K[i*npad+j]));
        e2m = fmax(e2m, fabs(t - (i == j)));
        e3m = fmax(e3m, fabs(U[j*npad+i] - X[i*npad+j]));
    }
    printf("    residuals: |LL^T-K| %.2e  |XL-I| %.2e  |U-X^T| %.2e\n", e1m, e2m, e3m);
    cudaFree(dK); cudaFree(dL); cudaFree(dX); cudaFree(dU); cudaFree(ddiag); cudaFree(dstat); cudaFree(gate); cudaFree(stamps);
}

int main() {
    const int npad = 128;
    std::vector<double> K((size_t)npad * npad);
    srand(1);
    std::vector<double> pts(npad * 3); for (auto& p : pts) p = rand() / (double)RAND_MAX;
    for (int i = 0; i < npad; ++i) for (int j = 0; j < npad; ++j) {
        double q = 0; for (int k = 0; k < 3; ++k) { double d = (pts[i*3+k] - pts[j*3+k]) / 0.7; q += d * d; }
        double r = sqrt(q); K[i*npad+j] = (1 + r * (2.2360679775 + r * 5.0 / 3.0)) * exp(-2.2360679775 * r) + (i == j ? 1e-8 : 0.0);
    }
    void pieces(const std::vector<double>& K, int npad);
    pieces(K, npad);
    for (int batch : {1}) {
        run<true>(batch, npad, K, 1, "recursive 32-base leaf");
        run<true>(batch, npad, K, 1, "recursive, no U / zero stores", 0);
        run<false>(batch, npad, K, 1, "panel4 leaf (round 1)");
    }
    return 0;
}

// ---- pieces of chol_inv_64_rec timed separately (CTA 0) ----
__global__ void __launch_bounds__(LEAF_THREADS) pieces_kernel(const double* K, int npad, long long* st) {
    extern __shared__ __align__(16) double sm[];
    double* A = sm; double* W = A + 64 * SLD; double* colb = W + 64 * SLD; double* rowb = colb + 512;
    double* invd = rowb + 512; double* dd = invd + 64;
    leaf_load_async(A, K, npad, 0, 0); cp_async_commit(); cp_async_wait<0>(); __syncthreads();
    long long t[12]; int i = 0;
    t[i++] = clock64();
    chol_inv_blk<32>(A, W, colb, rowb, invd, dd);
    t[i++] = clock64();
    mma_blk<32, true, MM_BT_BLOWER>(A + 32 * SLD, W, 1.0, W + 32, nullptr); __syncthreads();
    t[i++] = clock64();
    mma_blk<32, true, MM_SYRK_LOWER>(W + 32, W + 32, -1.0, A + 32 * SLD + 32, A + 32 * SLD + 32); __syncthreads();
    t[i++] = clock64();
    mma_blk<32, false, MM_NN_ALOWER>(W, A + 32 * SLD, 1.0, W + 32, nullptr); __syncthreads();
    t[i++] = clock64();
    mma_blk<32, false, MM_NN_BLOWER>(W + 32, W, -1.0, W + 32 * SLD, nullptr); __syncthreads();
    t[i++] = clock64();
    __syncthreads();
    t[i++] = clock64();
    mma_blk<64, true, MM_BT_BLOWER>(A, W, 1.0, colb + 2000, nullptr); __syncthreads();  // (writes garbage region: timing only)
    t[i++] = clock64();
    if (threadIdx.x == 0) for (int k = 0; k < i; ++k) st[k] = t[k];
}
void pieces(const std::vector<double>& K, int npad) {
    double* dK; long long* st; cudaMalloc(&dK, K.size() * 8); cudaMalloc(&st, 128);
    cudaMemcpy(dK, K.data(), K.size() * 8, cudaMemcpyHostToDevice);
    int smem = (4 * 64 * SLD + 2 * 512 + 2 * 64) * 8;
    cudaFuncSetAttribute(pieces_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    pieces_kernel<<<1, LEAF_THREADS, smem>>>(dK, npad, st);
    cudaDeviceSynchronize();
    pieces_kernel<<<1, LEAF_THREADS, smem>>>(dK, npad, st);
    cudaDeviceSynchronize();
    long long h[12]; cudaMemcpy(h, st, 8 * 8, cudaMemcpyDeviceToHost);
    printf("pieces (clocks): base32 %lld | mma32 BT_BLOWER %lld | mma32 SYRK %lld | mma32 NN_ALOWER %lld | mma32 NN_BLOWER %lld | sync %lld | mma64 BT_BLOWER %lld\n",
           h[1]-h[0], h[2]-h[1], h[3]-h[2], h[4]-h[3], h[5]-h[4], h[6]-h[5], h[7]-h[6]);
    printf("err: %s\n", cudaGetErrorString(cudaGetLastError()));
}
