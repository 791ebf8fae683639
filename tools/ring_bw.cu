// Micro-benchmark: how fast can a thread block stream a (L2-resident) weight buffer through a ring of
// cp.async.bulk stages?  Prints GB/s per block and aggregate for several ring shapes / grid sizes.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <vector>
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mb_init(uint32_t b, int c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(b), "r"(c)); }
__device__ __forceinline__ void mb_tx(uint32_t b, uint32_t n) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(n) : "memory"); }
__device__ __forceinline__ void mb_arrive(uint32_t b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(b) : "memory"); }
__device__ __forceinline__ bool mb_try(uint32_t b, uint32_t ph) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(b), "r"(ph) : "memory");
    return ok;
}
__device__ __forceinline__ void bulk(uint32_t dst, const void* src, uint32_t n, uint32_t b) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(n), "r"(b) : "memory");
}
// warp 0 lane 0 = producer (runs ahead), 8 consumer warps read every float4 of each stage once
__global__ void ring_kernel(const float* __restrict__ w, size_t total_floats, int stage_floats, int nstage, int reps, float* sink) {
    extern __shared__ __align__(128) float sm[];
    uint64_t* bars = (uint64_t*)sm;                 // full[nstage], empty[nstage]
    float* ring = sm + 64;
    const int nchunk = (int)(total_floats / stage_floats);
    if (threadIdx.x == 0) {
        for (int i = 0; i < nstage; ++i) { mb_init(s32(bars + i), 1); mb_init(s32(bars + nstage + i), 8); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int total = nchunk * reps;
    if (threadIdx.x >= 256) {                       // producer warp
        if (threadIdx.x == 256)
            for (int j = 0; j < total; ++j) {
                const int st = j % nstage;
                while (!mb_try(s32(bars + nstage + st), ((j / nstage) & 1) ^ 1)) {}
                mb_tx(s32(bars + st), stage_floats * 4);
                // different blocks start at different chunks so that they do not read in lock-step
                const size_t c = (size_t)((j + blockIdx.x * 7) % nchunk);
                bulk(s32(ring + (size_t)st * stage_floats), w + c * stage_floats, stage_floats * 4, s32(bars + st));
            }
        return;
    }
    float acc = 0.f;
    for (int j = 0; j < total; ++j) {
        const int st = j % nstage;
        while (!mb_try(s32(bars + st), (j / nstage) & 1)) {}
        const float4* p = (const float4*)(ring + (size_t)st * stage_floats);
        for (int i = threadIdx.x; i < stage_floats / 4; i += 256) { float4 v = p[i]; acc += v.x + v.y + v.z + v.w; }
        __syncwarp();
        if ((threadIdx.x & 31) == 0) mb_arrive(s32(bars + nstage + st));
    }
    if (acc == 123.456f) sink[0] = acc;
}
int main() {
    const size_t total = 3 * 1024 * 1024;           // 12 MB of "weights" (L2 resident)
    float *w, *sink;
    cudaMalloc(&w, total * 4); cudaMalloc(&sink, 4);
    cudaMemset(w, 0, total * 4);
    cudaFuncSetAttribute(ring_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    const int grids[] = {1, 32, 64, 128, 148};
    const int shapes[][2] = {{4096, 3}, {4096, 6}, {4096, 12}, {8192, 6}, {8192, 12}, {16384, 6}, {2048, 24}};
    for (auto& sh : shapes)
        for (int g : grids) {
            const int sf = sh[0], ns = sh[1], reps = 4;
            const size_t smem = 256 + (size_t)sf * ns * 4;
            cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
            ring_kernel<<<g, 288, smem>>>(w, total, sf, ns, 1, sink);
            cudaEventRecord(a);
            ring_kernel<<<g, 288, smem>>>(w, total, sf, ns, reps, sink);
            cudaEventRecord(b); cudaEventSynchronize(b);
            float ms; cudaEventElapsedTime(&ms, a, b);
            cudaError_t e = cudaGetLastError();
            double gb = (double)total * 4 * reps / 1e9;
            printf("stage %5d KB x %2d (%3zu KB ring) grid %3d: %7.3f ms  %7.1f GB/s per block  %8.1f GB/s aggregate %s\n", sf * 4 / 1024, ns,
                   smem / 1024, g, ms, gb / (ms * 1e-3), gb * g / (ms * 1e-3), e == cudaSuccess ? "" : cudaGetErrorString(e));
        }
    return 0;
}
