// Efficiency of the batched NT GEMM on the product shapes of the tile-column factorisation (development aid).
// Links against the built library:  nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -o tools/gemm_bench \
//     tools/gemm_bench.cu -Lbobe_b200/lib -lbobe_b200 -Xlinker -rpath -Xlinker $PWD/bobe_b200/lib
#include <cstdio>
#include <vector>
#include "../bobe_b200/csrc/gemm_nt.cuh"
using namespace bobe;
struct Shape { const char* name; int M, N, K, flags; bool inplace; double live; };  // live: fraction of the M*N*K box that is algorithmic
int main(int argc, char** argv) {
    const int npad = 2048;
    const int64_t m2 = (int64_t)npad * npad;
    std::vector<Shape> shapes = {
        {"update k=128  (rows x 128)", 1664, 128, 128, 0, true, 1.0},
        {"panel k=128 B lower", 1664, 128, 128, GEMM_B_LOWER, false, 0.75},
        {"a1 k=256", 1664, 128, 256, 0, true, 1.0},
        {"a1 k=384", 1408, 128, 384, 0, true, 1.0},
        {"column k=512", 1408, 128, 512, 0, true, 1.0},
        {"trailing 1408^2 k=512 C lower", 1408, 1408, 512, GEMM_C_LOWER, true, 0.5},
        {"trailing 896^2 k=512 C lower", 896, 896, 512, GEMM_C_LOWER, true, 0.5},
        {"P1 top 1024^3 A upper", 1024, 1024, 1024, GEMM_A_UPPER, false, 0.5},
        {"P2 top 1024^3 A lower dual store", 1024, 1024, 1024, GEMM_A_LOWER, false, 0.5},
        {"K^-1 2048^3 upper x upper, C lower", 2048, 2048, 2048, GEMM_A_UPPER | GEMM_B_UPPER | GEMM_C_LOWER, false, 1.0 / 3.0},
    };
    for (int batch : {64, 16, 8}) {
        double *A, *B, *C, *Ct;
        cudaMalloc(&A, batch * m2 * 8); cudaMalloc(&B, batch * m2 * 8); cudaMalloc(&C, batch * m2 * 8); cudaMalloc(&Ct, batch * m2 * 8);
        cudaMemset(A, 0, batch * m2 * 8); cudaMemset(B, 0, batch * m2 * 8); cudaMemset(C, 0, batch * m2 * 8);
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        for (auto& s : shapes) {
            GemmArgs g{};
            g.A = A; g.Bt = B; g.C = C; g.D = s.inplace ? C : nullptr; g.Ct = (s.flags & GEMM_A_LOWER) ? Ct : nullptr;
            g.lda = g.ldb = g.ldc = g.ldct = g.ldd = npad;
            g.strideA = g.strideB = g.strideC = g.strideCt = g.strideD = m2;
            g.M = s.M; g.N = s.N; g.K = s.K; g.alpha = -1.0; g.flags = s.flags;
            for (int i = 0; i < 2; ++i) launch_gemm_nt(0, g, batch);
            cudaDeviceSynchronize();
            const int iters = 5;
            cudaEventRecord(e0);
            for (int i = 0; i < iters; ++i) launch_gemm_nt(0, g, batch);
            cudaEventRecord(e1); cudaDeviceSynchronize();
            float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= iters;
            double fl = 2.0 * s.M * s.N * s.K * s.live * batch;
            printf("batch %2d  %-36s %8.1f us  %6.2f TF/s algorithmic (%4.1f %% of 35.46)\n", batch, s.name, ms * 1e3, fl / ms / 1e9, fl / ms / 1e9 / 35.46 * 100);
        }
        cudaFree(A); cudaFree(B); cudaFree(C); cudaFree(Ct);
        printf("\n");
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
