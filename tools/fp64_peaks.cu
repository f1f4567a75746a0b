// Microbenchmark: FP64 DFMA vs DMMA (mma.sync m8n8k4 f64) issue throughput on sm_100a.
// Used once to establish the FP64 roofline denominator (MEASURED_PEAKS.json has none).
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

template <int NACC>
__global__ void k_dfma(double* out, int iters, double s) {
    double acc[NACC];
#pragma unroll
    for (int i = 0; i < NACC; i++) acc[i] = threadIdx.x * 1e-3 + i;
    double a = s, b = 1.0 - s;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < NACC; i++) acc[i] = fma(acc[i], a, b);
    }
    double r = 0;
#pragma unroll
    for (int i = 0; i < NACC; i++) r += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <int NACC>
__global__ void k_dmma(double* out, int iters, double s) {
    double c0[NACC], c1[NACC];
#pragma unroll
    for (int i = 0; i < NACC; i++) { c0[i] = threadIdx.x * 1e-3 + i; c1[i] = i; }
    double a = s, b = 1.0 - s;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < NACC; i++) dmma(c0[i], c1[i], a, b);
    }
    double r = 0;
#pragma unroll
    for (int i = 0; i < NACC; i++) r += c0[i] + c1[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

// mixed: NACC dmma + NF dfma per iteration, to see whether the pipes are shared
template <int NACC, int NF>
__global__ void k_mix(double* out, int iters, double s) {
    double c0[NACC], c1[NACC], f[NF];
#pragma unroll
    for (int i = 0; i < NACC; i++) { c0[i] = threadIdx.x * 1e-3 + i; c1[i] = i; }
#pragma unroll
    for (int i = 0; i < NF; i++) f[i] = threadIdx.x + i;
    double a = s, b = 1.0 - s;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < NACC; i++) dmma(c0[i], c1[i], a, b);
#pragma unroll
        for (int i = 0; i < NF; i++) f[i] = fma(f[i], a, b);
    }
    double r = 0;
#pragma unroll
    for (int i = 0; i < NACC; i++) r += c0[i] + c1[i];
#pragma unroll
    for (int i = 0; i < NF; i++) r += f[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <typename F>
float timeit(F f) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 5; r++) {
        cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    return best;
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount;
    printf("device %s sms %d clock %d kHz\n", p.name, sms, p.clockRate);
    double* out; cudaMalloc(&out, sizeof(double) * sms * 8 * 1024);
    const int iters = 20000;
    for (int warps : {4, 8, 16, 32}) {
        int threads = warps * 32; int blocks = sms * (warps >= 32 ? 1 : 2);
        double nthreads = double(threads) * blocks;
        float ms = timeit([&] { k_dfma<8><<<blocks, threads>>>(out, iters, 0.5); });
        printf("DFMA  warps/blk %2d blocks %d: %.2f TFLOP/s\n", warps, blocks, 2.0 * 8 * iters * nthreads / ms / 1e9);
        ms = timeit([&] { k_dmma<8><<<blocks, threads>>>(out, iters, 0.5); });
        printf("DMMA  warps/blk %2d blocks %d: %.2f TFLOP/s\n", warps, blocks, 2.0 * 8 * 8 * iters * nthreads / ms / 1e9);
        ms = timeit([&] { k_mix<8, 8><<<blocks, threads>>>(out, iters, 0.5); });
        printf("MIX8+8 warps/blk %2d: dmma-part %.2f TF + dfma-part %.2f TF (%.3f ms)\n", warps,
               2.0 * 8 * 8 * iters * nthreads / ms / 1e9, 2.0 * 8 * iters * nthreads / ms / 1e9, ms);
        ms = timeit([&] { k_mix<8, 32><<<blocks, threads>>>(out, iters, 0.5); });
        printf("MIX8+32 warps/blk %2d: dmma-part %.2f TF + dfma-part %.2f TF (%.3f ms)\n", warps,
               2.0 * 8 * 8 * iters * nthreads / ms / 1e9, 2.0 * 32 * iters * nthreads / ms / 1e9, ms);
    }
    cudaError_t e = cudaDeviceSynchronize();
    printf("status %s\n", cudaGetErrorString(e));
    return e != cudaSuccess;
}
