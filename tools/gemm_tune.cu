// Tile-configuration sweep for the fused triangular-multiply + sum-of-squares kernel (development aid).
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../bobe_b200/csrc/gemm_nt.cuh"
namespace bobe { void set_error(const char*, ...) {} int32_t check_launch(const char*) { return 0; } }
using namespace bobe;

template <class Cfg>
void run(const char* name, const double* Linv, int npad, const double* K, double* out, int n) {
    int rows = 148 * 128;  // same chunk for every configuration (grid = rows / BN CTAs)
    cudaFuncSetAttribute(trmm_sumsq_kernel<Cfg>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int it = 0; it < 6; ++it) {
        cudaEventRecord(e0);
        trmm_sumsq_kernel<Cfg><<<rows / Cfg::BN, Cfg::THREADS, Cfg::SMEM_BYTES>>>(Linv, n, npad, K, npad, 0, rows, 1.0, 1.0, 0, out, nullptr, 0);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (it && ms < best) best = ms;
    }
    cudaError_t e = cudaGetLastError();
    double fl = double(rows) * n * n;
    printf("%-28s rows %6d smem %3d KB: %7.3f ms  %6.2f TF algorithmic  (%s)\n", name, rows, Cfg::SMEM_BYTES / 1024, best,
           fl / best / 1e9, cudaGetErrorString(e));
}

int main(int argc, char** argv) {
    const int n = argc > 1 ? atoi(argv[1]) : 2000, npad = ((n + 63) / 64) * 64;
    const int maxrows = 148 * 256;
    std::vector<double> h((size_t)npad * npad, 0.0);
    for (int i = 0; i < npad; ++i) for (int j = 0; j <= i; ++j) h[(size_t)i * npad + j] = (i == j) ? 1.0 : 1e-3 * ((i * 131 + j * 7) % 97 - 48);
    double *Linv, *K, *out;
    cudaMalloc(&Linv, sizeof(double) * npad * npad); cudaMalloc(&K, sizeof(double) * (size_t)maxrows * npad); cudaMalloc(&out, sizeof(double) * maxrows);
    cudaMemcpy(Linv, h.data(), sizeof(double) * npad * npad, cudaMemcpyHostToDevice);
    std::vector<double> hk((size_t)maxrows * npad);
    for (size_t i = 0; i < hk.size(); ++i) hk[i] = ((i * 2654435761u) % 1000) * 1e-3;
    cudaMemcpy(K, hk.data(), sizeof(double) * hk.size(), cudaMemcpyHostToDevice);
    run<TileCfg<128, 128, 4, 4, 3, 32, true>>("128x128 w4x4 s3 bk32 ilv", Linv, npad, K, out, n);
    run<TileCfg<128, 64, 4, 2, 4, 16, true, 2>>("128x64 w4x2 s4 bk16 ilv x2", Linv, npad, K, out, n);
    run<TileCfg<128, 64, 4, 2, 3, 16, true, 2>>("128x64 w4x2 s3 bk16 ilv x2", Linv, npad, K, out, n);
    run<TileCfg<128, 64, 2, 2, 3, 16, true, 3>>("128x64 w2x2 s3 bk16 ilv x3", Linv, npad, K, out, n);
    run<TileCfg<64, 64, 2, 2, 3, 16, true, 4>>("64x64 w2x2 s3 bk16 ilv x4", Linv, npad, K, out, n);
    run<TileCfg<128, 128, 4, 4, 2, 32, true>>("128x128 w4x4 s2 bk32 ilv", Linv, npad, K, out, n);
    run<TileCfg<64, 128, 2, 4, 3, 16, true, 2>>("64x128 w2x4 s3 bk16 ilv x2", Linv, npad, K, out, n);
    run<TileCfg<64, 64, 2, 2, 4, 16, true, 3>>("64x64 w2x2 s4 bk16 ilv x3", Linv, npad, K, out, n);
    return 0;
}
