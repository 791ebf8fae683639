// Why does fragment streaming by LDG.128 reach only ~24 GB/s per SM?  Variants: load flavour, with / without MMA work.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
__device__ __forceinline__ void mma(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
template <int MODE>
__device__ __forceinline__ float4 ld(const float4* p) {
    float4 v;
    if (MODE == 0) asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    else if (MODE == 1) asm volatile("ld.global.nc.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    else if (MODE == 2) asm volatile("ld.global.cg.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    else asm volatile("ld.global.ca.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}
// every warp streams `ksteps` consecutive 512-byte blocks per layer, U loads in flight; WORK: 0 = sum only, 1 = 3 MMAs per block
template <int MODE, int U, int WORK>
__global__ void k(const float4* __restrict__ w, size_t total_f4, int ksteps, int layers, float* sink, long long* cyc, int share) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    float c[4] = {0.f, 0.f, 0.f, 0.f};
    // share = number of blocks that read the SAME addresses (the sequence kernel: all clusters read the same panels)
    size_t base = ((size_t)(blockIdx.x / share) * 7919 * 32) % (total_f4 / 2);
    long long t0 = clock64();
    for (int L = 0; L < layers; ++L) {
        const float4* p = w + base + (size_t)warp * ksteps * 32 + lane;
        float4 buf[U];
#pragma unroll
        for (int j = 0; j < U; ++j) buf[j] = ld<MODE>(p + j * 32);
        for (int kk = 0; kk < ksteps; kk += U) {
#pragma unroll
            for (int j = 0; j < U; ++j) {
                const float4 a = buf[j];
                if (WORK) {
                    uint32_t af[4] = {__float_as_uint(a.x), __float_as_uint(a.y), __float_as_uint(a.z), __float_as_uint(a.w)};
                    mma(c, af, 0x3f800000u, 0x3f800000u); mma(c, af, 0x3f000000u, 0x3f800000u); mma(c, af, 0x3f800000u, 0x3f000000u);
                } else {
                    c[0] += a.x; c[1] += a.y; c[2] += a.z; c[3] += a.w;
                }
                if (kk + j + U < ksteps) buf[j] = ld<MODE>(p + (size_t)(kk + j + U) * 32);
            }
        }
        base = (base + (size_t)nw * ksteps * 32 + 32 * 1031) % (total_f4 / 2);
        __syncthreads();
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[blockIdx.x] = (t1 - t0) / layers;
    if (c[0] + c[1] + c[2] + c[3] == 123.456f) sink[0] = c[0];
}
template <int MODE, int U, int WORK>
void run(const char* name, const float4* w, size_t total_f4, int grid, int nt, int ksteps, float* sink, long long* cyc, int share = 1) {
    const int layers = 200;
    k<MODE, U, WORK><<<grid, nt>>>(w, total_f4, ksteps, 10, sink, cyc, share);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a);
    k<MODE, U, WORK><<<grid, nt>>>(w, total_f4, ksteps, layers, sink, cyc, share);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    const double bytes = (double)(nt / 32) * ksteps * 512;
    printf("share %3d %-16s U=%2d work=%d grid %3d thr %3d ksteps/warp %3d: %7.0f cyc/layer  %6.1f cyc/kstep/warp  %6.1f GB/s per SM %8.1f GB/s total %s\n", share, name, U, WORK,
           grid, nt, ksteps, (double)cyc[0], (double)cyc[0] / ksteps, bytes * layers / (ms * 1e-3) / 1e9, bytes * layers * grid / (ms * 1e-3) / 1e9,
           cudaGetLastError() == cudaSuccess ? "" : "ERR");
}
int main() {
    const size_t total_f4 = 6 * 1024 * 1024;        // 96 MB
    float4* w; float* sink; long long* cyc;
    cudaMalloc(&w, total_f4 * 16); cudaMalloc(&sink, 4); cudaMallocManaged(&cyc, 8 * 1024);
    cudaMemset(w, 0, total_f4 * 16);
    for (int share : {1, 4, 32, 128}) {
        run<1, 8, 0>("nc", w, total_f4, 128, 384, 40, sink, cyc, share);
        run<1, 4, 1>("nc", w, total_f4, 128, 384, 40, sink, cyc, share);
        run<1, 4, 1>("nc", w, total_f4, 128, 384, 12, sink, cyc, share);
        run<1, 4, 0>("nc", w, total_f4, 128, 384, 4, sink, cyc, share);
    }
    return 0;
}
