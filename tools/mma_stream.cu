// Micro-benchmark for the v3 dense layer: every warp streams its weight fragments (mma.m16n8k8 A operands,
// pre-arranged in fragment order) straight from L2 with LDG.128 and multiplies them with activations held in
// shared memory, 3xTF32 (hi*hi + hi*lo + lo*hi).  Reports cycles per "layer" and the implied L2->SM rate.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
__device__ __forceinline__ uint32_t tf32(float x) { uint32_t r; asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x)); return r; }
__device__ __forceinline__ void mma(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ float4 ldg_stream(const float4* p) {
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}
template <int U>
__global__ void k(const float4* __restrict__ w, size_t total_f4, int nmt, int ksteps, int ksplit, int layers, float* sink, long long* cyc) {
    extern __shared__ float x[];                         // [K][8]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int i = threadIdx.x; i < ksteps * 8 * 8; i += blockDim.x) x[i] = 0.001f * (float)(i % 97);
    __syncthreads();
    const size_t layer_f4 = (size_t)nmt * ksteps * 32;
    size_t base = ((size_t)blockIdx.x * 7919 * 32) % (total_f4 - layer_f4 * 2);
    float tot = 0.f;
    long long t0 = clock64();
    for (int L = 0; L < layers; ++L) {
        for (int u = warp; u < nmt * ksplit; u += nw) {
            const int mt = u % nmt, sl = u / nmt;
            const int per = (ksteps + ksplit - 1) / ksplit, k0 = sl * per, k1 = min(ksteps, k0 + per);
            const float4* p = w + base + ((size_t)mt * ksteps + k0) * 32 + lane;
            float c[4] = {0.f, 0.f, 0.f, 0.f};
            float4 buf[U];
#pragma unroll
            for (int j = 0; j < U; ++j) if (k0 + j < k1) buf[j] = ldg_stream(p + j * 32);
            for (int kk = k0; kk < k1; kk += U) {
#pragma unroll
                for (int j = 0; j < U; ++j) {
                    if (kk + j < k1) {
                        const float4 a = buf[j];
                        if (kk + j + U < k1) buf[j] = ldg_stream(p + (size_t)(kk - k0 + j + U) * 32);
                        const float b0f = x[((kk + j) * 8 + (lane & 3)) * 8 + (lane >> 2)];
                        const float b1f = x[((kk + j) * 8 + (lane & 3) + 4) * 8 + (lane >> 2)];
                        uint32_t ah[4], al[4];
                        const float av[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
                        for (int q = 0; q < 4; ++q) { ah[q] = tf32(av[q]); al[q] = tf32(av[q] - __uint_as_float(ah[q])); }
                        const uint32_t b0h = tf32(b0f), b1h = tf32(b1f);
                        const uint32_t b0l = tf32(b0f - __uint_as_float(b0h)), b1l = tf32(b1f - __uint_as_float(b1h));
                        mma(c, al, b0h, b1h);
                        mma(c, ah, b0l, b1l);
                        mma(c, ah, b0h, b1h);
                    }
                }
            }
            tot += c[0] + c[1] + c[2] + c[3];
        }
        base = (base + layer_f4 + 32 * 1031) % (total_f4 - layer_f4 * 2);
        __syncthreads();
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[blockIdx.x] = (t1 - t0) / layers;
    if (tot == 123.456f) sink[0] = tot;
}
template <int U>
void run(const float4* w, size_t total_f4, int grid, int nt, int nmt, int ksteps, int ksplit, float* sink, long long* cyc) {
    const int layers = 200;
    const int smem = ksteps * 64 * 4;
    cudaFuncSetAttribute(k<U>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    k<U><<<grid, nt, smem>>>(w, total_f4, nmt, ksteps, ksplit, 20, sink, cyc);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a);
    k<U><<<grid, nt, smem>>>(w, total_f4, nmt, ksteps, ksplit, layers, sink, cyc);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    cudaError_t e = cudaGetLastError();
    const double bytes = (double)nmt * ksteps * 512;
    printf("U=%2d grid %3d threads %3d  layer %3d cols x %4d k (%6.1f KB) ksplit %d: %7.0f cyc/layer (blk0)  %6.1f GB/s per SM  %7.1f GB/s aggregate %s\n", U, grid, nt,
           nmt * 16, ksteps * 8, bytes / 1024, ksplit, (double)cyc[0], bytes * layers / (ms * 1e-3) / 1e9, bytes * layers * grid / (ms * 1e-3) / 1e9,
           e == cudaSuccess ? "" : cudaGetErrorString(e));
}
int main() {
    const size_t total_f4 = 3 * 1024 * 1024;        // 48 MB of weights (L2 resident)
    float4* w; float* sink; long long* cyc;
    cudaMalloc(&w, total_f4 * 16); cudaMalloc(&sink, 4); cudaMallocManaged(&cyc, 8 * 1024);
    cudaMemset(w, 0, total_f4 * 16);
    for (int grid : {1, 120, 148}) {
        for (int nt : {256, 384, 512}) {
            const int nwarp = nt / 32;
            // big layer 96 x 672, mid layer 96 x 256, small 16 x 128
            run<4>(w, total_f4, grid, nt, 6, 84, nwarp / 6 > 0 ? nwarp / 6 : 1, sink, cyc);
            run<8>(w, total_f4, grid, nt, 6, 84, nwarp / 6 > 0 ? nwarp / 6 : 1, sink, cyc);
            run<8>(w, total_f4, grid, nt, 6, 84, (nwarp + 5) / 6, sink, cyc);
            run<8>(w, total_f4, grid, nt, 6, 32, (nwarp + 5) / 6, sink, cyc);
            run<4>(w, total_f4, grid, nt, 6, 32, (nwarp + 5) / 6, sink, cyc);
            run<8>(w, total_f4, grid, nt, 1, 16, 4, sink, cyc);
            run<4>(w, total_f4, grid, nt, 1, 16, nwarp, sink, cyc);
        }
    }
    return 0;
}
