// Probe (development aid): can runtime-API kernels be launched into streams of two green contexts with disjoint SM sets,
// do events order work across them, and which SMs does each use?
#include <cstdio>
#include <cuda.h>
#include <cuda_runtime.h>
#include <cudaTypedefs.h>
#include <set>
#include <vector>
#define GET(name, ver, T)                                                                                     \
    T name = nullptr;                                                                                           \
    {                                                                                                           \
        void* p_ = nullptr; cudaDriverEntryPointQueryResult q_;                                                 \
        if (cudaGetDriverEntryPointByVersion(#name, &p_, ver, cudaEnableDefault, &q_) != cudaSuccess || q_ != cudaDriverEntryPointSuccess) { \
            printf("no entry point %s\n", #name); return 1; }                                                   \
        name = (T)p_;                                                                                           \
    }
__global__ void who(int* smid, int iters, double* sink) {
    unsigned id; asm volatile("mov.u32 %0, %%smid;" : "=r"(id));
    double a = threadIdx.x;
    for (int i = 0; i < iters; ++i) a = fma(a, 1.0000001, 0.5);
    if (threadIdx.x == 0) smid[blockIdx.x] = (int)id;
    if (a == 12345.678) *sink = a;
}
int main() {
    cudaFree(0);
    GET(cuDeviceGetDevResource, 12040, PFN_cuDeviceGetDevResource_v12040)
    GET(cuDevSmResourceSplitByCount, 12040, PFN_cuDevSmResourceSplitByCount_v12040)
    GET(cuDevResourceGenerateDesc, 12040, PFN_cuDevResourceGenerateDesc_v12040)
    GET(cuGreenCtxCreate, 12040, PFN_cuGreenCtxCreate_v12040)
    GET(cuGreenCtxStreamCreate, 12050, PFN_cuGreenCtxStreamCreate_v12050)
    CUdevResource all, grp, rest;
    CUresult r = cuDeviceGetDevResource(0, &all, CU_DEV_RESOURCE_TYPE_SM);
    printf("get resource: %d, sm count %u\n", (int)r, all.sm.smCount);
    unsigned n = 1;
    r = cuDevSmResourceSplitByCount(&grp, &n, &all, &rest, 0, 16);
    printf("split: %d groups %u: group %u SMs, remaining %u SMs\n", (int)r, n, grp.sm.smCount, rest.sm.smCount);
    CUdevResourceDesc dA, dB;
    printf("desc: %d %d\n", (int)cuDevResourceGenerateDesc(&dA, &grp, 1), (int)cuDevResourceGenerateDesc(&dB, &rest, 1));
    CUgreenCtx gA, gB;
    printf("green ctx: %d %d\n", (int)cuGreenCtxCreate(&gA, dA, 0, CU_GREEN_CTX_DEFAULT_STREAM), (int)cuGreenCtxCreate(&gB, dB, 0, CU_GREEN_CTX_DEFAULT_STREAM));
    CUstream sA, sB;
    printf("streams: %d %d\n", (int)cuGreenCtxStreamCreate(&sA, gA, CU_STREAM_NON_BLOCKING, -5), (int)cuGreenCtxStreamCreate(&sB, gB, CU_STREAM_NON_BLOCKING, 0));
    int *da, *db; double* sink; cudaMalloc(&da, 4096 * 4); cudaMalloc(&db, 4096 * 4); cudaMalloc(&sink, 8);
    cudaEvent_t e0, e1, eA; cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventCreateWithFlags(&eA, cudaEventDisableTiming);
    cudaStream_t plain; cudaStreamCreateWithFlags(&plain, cudaStreamNonBlocking);
    // B: a long low-priority grid filling its partition; A: a short grid meanwhile; plain stream waits for both via events
    cudaEventRecord(e0, plain);
    cudaEventRecord(eA, plain);
    cudaStreamWaitEvent((cudaStream_t)sA, eA, 0); cudaStreamWaitEvent((cudaStream_t)sB, eA, 0);
    who<<<2000, 256, 100 * 1024, (cudaStream_t)sB>>>(db, 200000, sink);
    who<<<64, 256, 200 * 1024, (cudaStream_t)sA>>>(da, 20000, sink);
    printf("launch: %s\n", cudaGetErrorString(cudaGetLastError()));
    cudaEvent_t fa, fb; cudaEventCreate(&fa); cudaEventCreate(&fb);
    cudaEventRecord(fa, (cudaStream_t)sA); cudaEventRecord(fb, (cudaStream_t)sB);
    cudaStreamWaitEvent(plain, fa, 0); cudaStreamWaitEvent(plain, fb, 0);
    cudaEventRecord(e1, plain);
    printf("sync: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    float tA, tB, tAll; cudaEventElapsedTime(&tA, e0, fa); cudaEventElapsedTime(&tB, e0, fb); cudaEventElapsedTime(&tAll, e0, e1);
    std::vector<int> ha(64), hb(2000); cudaMemcpy(ha.data(), da, 64 * 4, cudaMemcpyDeviceToHost); cudaMemcpy(hb.data(), db, 2000 * 4, cudaMemcpyDeviceToHost);
    std::set<int> sa(ha.begin(), ha.end()), sb(hb.begin(), hb.end());
    int common = 0; for (int x : sa) common += sb.count(x);
    printf("A used %zu SMs (done at %.3f ms), B used %zu SMs (done at %.3f ms), common %d, all %.3f ms\n", sa.size(), tA, sb.size(), tB, common, tAll);
    // func attribute check inside green stream: dynamic smem > 48 KB was requested above without cudaFuncSetAttribute -> expect launch error unless set
    cudaFuncSetAttribute(who, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    who<<<64, 256, 200 * 1024, (cudaStream_t)sA>>>(da, 20000, sink);
    who<<<2000, 256, 100 * 1024, (cudaStream_t)sB>>>(db, 200000, sink);
    printf("launch 2: %s, sync %s\n", cudaGetErrorString(cudaGetLastError()), cudaGetErrorString(cudaDeviceSynchronize()));
    cudaMemcpy(ha.data(), da, 64 * 4, cudaMemcpyDeviceToHost); cudaMemcpy(hb.data(), db, 2000 * 4, cudaMemcpyDeviceToHost);
    sa = std::set<int>(ha.begin(), ha.end()); sb = std::set<int>(hb.begin(), hb.end());
    common = 0; for (int x : sa) common += sb.count(x);
    printf("after attr: A used %zu SMs, B used %zu SMs, common %d\n", sa.size(), sb.size(), common);
    return 0;
}
