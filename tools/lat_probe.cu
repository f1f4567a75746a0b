// Dependent-issue latencies that shape the leaf kernel (development aid): DFMA, DMUL, MUFU.RCP64H, DMMA chain, LDS, bar.sync
#include <cstdio>
__global__ void probe(double* out, long long* t, double x) {
    __shared__ double sm[1024];
    long long c0, c1; double a = x, b = x * 0.5 + 1e-3, r;
    sm[threadIdx.x] = x; __syncthreads();
    c0 = clock64();
#pragma unroll
    for (int i = 0; i < 64; ++i) a = fma(a, b, b);
    c1 = clock64(); if (threadIdx.x == 0) t[0] = (c1 - c0); out[0] = a;
    c0 = clock64();
#pragma unroll
    for (int i = 0; i < 64; ++i) a = a * b;
    c1 = clock64(); if (threadIdx.x == 0) t[1] = (c1 - c0); out[1] = a;
    a = x + 1.0;
    c0 = clock64();
#pragma unroll
    for (int i = 0; i < 64; ++i) { asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(a)); a = r; }
    c1 = clock64(); if (threadIdx.x == 0) t[2] = (c1 - c0); out[2] = a;
    double d0 = 0, d1 = 0; a = x; 
    c0 = clock64();
#pragma unroll
    for (int i = 0; i < 64; ++i) asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
    c1 = clock64(); if (threadIdx.x == 0) t[3] = (c1 - c0); out[3] = d0 + d1;
    int idx = threadIdx.x;
    c0 = clock64();
#pragma unroll
    for (int i = 0; i < 64; ++i) idx = (int)sm[idx & 1023] + threadIdx.x;
    c1 = clock64(); if (threadIdx.x == 0) t[4] = (c1 - c0); out[4] = idx;
    c0 = clock64();
#pragma unroll
    for (int i = 0; i < 64; ++i) __syncthreads();
    c1 = clock64(); if (threadIdx.x == 0) t[5] = (c1 - c0);
    // 8 independent DFMA chains (throughput per warp)
    double v[8]; for (int k = 0; k < 8; ++k) v[k] = x + k;
    c0 = clock64();
#pragma unroll
    for (int i = 0; i < 32; ++i)
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = fma(v[k], b, b);
    c1 = clock64(); if (threadIdx.x == 0) t[6] = (c1 - c0); for (int k = 0; k < 8; ++k) out[5] += v[k];
    // shuffle of a double (2 x SHFL)
    a = x;
    c0 = clock64();
#pragma unroll
    for (int i = 0; i < 64; ++i) a = __shfl_xor_sync(0xffffffffu, a, 1) + 1.0;
    c1 = clock64(); if (threadIdx.x == 0) t[7] = (c1 - c0); out[6] = a;
}
int main() {
    double* out; long long* t; cudaMalloc(&out, 64); cudaMalloc(&t, 64);
    for (int threads : {32, 256}) {
        probe<<<1, threads>>>(out, t, 0.0); probe<<<1, threads>>>(out, t, 0.0); cudaDeviceSynchronize();
        long long h[8]; cudaMemcpy(h, t, 64, cudaMemcpyDeviceToHost);
        printf("%3d threads: DFMA %.1f  DMUL %.1f  MUFU.RCP64H(+cvt) %.1f  DMMA dep %.1f  LDS dep(+cvt+add) %.1f  bar.sync %.1f  8-indep DFMA per instr %.2f  shfl64+dadd %.1f clk\n",
               threads, h[0] / 64.0, h[1] / 64.0, h[2] / 64.0, h[3] / 64.0, h[4] / 64.0, h[5] / 64.0, h[6] / 256.0, h[7] / 64.0);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}
