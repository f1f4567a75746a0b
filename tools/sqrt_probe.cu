// Accuracy probe for the branch-free sqrt used by the Matern kernel builds: rsqrt seed (MUFU.RSQ64H) followed by one or
// two Goldschmidt steps and a residual correction, against the correctly rounded __dsqrt_rn.  Prints the mismatch
// count and the largest error in ulps.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/sqrt_probe tools/sqrt_probe.cu
#include <cstdio>
#include <cstdint>
#include <cmath>
__device__ __forceinline__ double sqrt_iter(double q, int iters) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(q));
    double g = q * y, h = 0.5 * y;
    for (int i = 0; i < iters; ++i) {
        double r = fma(-g, h, 0.5);
        g = fma(g, r, g);
        h = fma(h, r, h);
    }
    return fma(fma(-g, g, q), h, g);
}
__global__ void probe(uint64_t seed, int iters, double lo_exp, double hi_exp, unsigned long long* mism, unsigned long long* maxulp,
                      double* seed_err) {
    uint64_t s = seed + (uint64_t)(blockIdx.x * blockDim.x + threadIdx.x) * 0x9E3779B97F4A7C15ull;
    unsigned long long mm = 0, mu = 0;
    double se = 0.0;
    for (int it = 0; it < 4096; ++it) {
        s ^= s << 13; s ^= s >> 7; s ^= s << 17;
        double u = (double)(s >> 11) * (1.0 / 9007199254740992.0);
        s ^= s << 13; s ^= s >> 7; s ^= s << 17;
        double m = 1.0 + (double)(s >> 11) * (1.0 / 9007199254740992.0);
        double q = m * exp2(floor(lo_exp + u * (hi_exp - lo_exp)));
        double ref = __dsqrt_rn(q), got = sqrt_iter(q, iters);
        long long d = llabs(__double_as_longlong(ref) - __double_as_longlong(got));
        if (d) ++mm;
        if ((unsigned long long)d > mu) mu = d;
        double y;
        asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(q));
        se = fmax(se, fabs(y * ref - 1.0));
    }
    atomicAdd(mism, mm);
    atomicMax(maxulp, mu);
    atomicMax((unsigned long long*)seed_err, (unsigned long long)__double_as_longlong(se));
}
int main() {
    unsigned long long *d, h[2];
    double* dse; double hse;
    cudaMalloc(&d, 16); cudaMalloc(&dse, 8);
    for (int iters = 0; iters <= 2; ++iters) {
        cudaMemset(d, 0, 16); cudaMemset(dse, 0, 8);
        probe<<<1024, 256>>>(12345, iters, -100.0, 20.0, d, d + 1, dse);
        cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost); cudaMemcpy(&hse, dse, 8, cudaMemcpyDeviceToHost);
        printf("goldschmidt steps %d: %llu mismatches of %llu vs __dsqrt_rn, max %llu ulp; rsqrt seed max rel err %.3e\n", iters, h[0],
               1024ull * 256 * 4096, h[1], hse);
    }
    return 0;
}
