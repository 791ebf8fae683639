// Dense-call inner loop candidates for R = 5 rows per block, measured under the product kernel's conditions (128 blocks x 12 warps,
// 190 KB of dynamic shared memory so that the L1 is as small as in the sequence kernel, 32 blocks reading the same weights):
//   MMA : the product's k-step -- 2 LDG.128 per lane (tf32 hi | lo fragments, 1 KB per warp) + 2 LDS + split + 4 mma.m16n8k8 + FADDs
//         per 128 weights, U k-steps in flight
//   FFMA: 1 LDG.128 per lane (4 consecutive k of one output column, 512 B per warp, unsplit fp32) + 5 broadcast LDS.128 (x[k..k+3][0..4])
//         + 20 FFMA per 128 weights, U loads in flight.  Exact fp32, half the weight bytes.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
__device__ __forceinline__ void mma(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void mma0(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%10,%10,%10};"
                 : "=f"(d[0]), "=f"(d[1]), "=f"(d[2]), "=f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1), "f"(0.f));
}
__device__ __forceinline__ float4 ld(const float4* p) {
    float4 v;
    asm volatile("ld.global.nc.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ void split(float x, uint32_t& hi, uint32_t& lo) {
    hi = __float_as_uint(x) & 0xffffe000u;
    lo = __float_as_uint(x - __uint_as_float(hi)) + 0x1000u;
}
extern __shared__ __align__(16) float smem[];

// steps = 128-weight steps per warp and layer
template <int KIND, int U>
__global__ void __launch_bounds__(384, 1) k(const float4* __restrict__ w, size_t total_f4, int steps, int layers, float* sink, long long* cyc, int share) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int i = threadIdx.x; i < 8192; i += blockDim.x) smem[i] = 1.f + (float)(i & 7);
    __syncthreads();
    size_t base = ((size_t)(blockIdx.x / share) * 7919 * 64) % (total_f4 / 2);
    float out = 0.f;
    long long t0 = clock64();
    for (int L = 0; L < layers; ++L) {
        if (KIND == 0) {                                   // MMA, pre-split fragments
            const float4* p = w + base + (size_t)warp * steps * 64 + lane;
            const int g = lane >> 2, t = lane & 3, gr = g < 5 ? g : 4;
            const float* xp = smem + t * 5 + gr;
            float4 bh[U], bl[U];
#pragma unroll
            for (int j = 0; j < U; ++j) { bh[j] = ld(p + j * 64); bl[j] = ld(p + j * 64 + 32); }
            float acc[4] = {0.f, 0.f, 0.f, 0.f};
            for (int kk = 0; kk < steps; kk += U) {
#pragma unroll
                for (int j = 0; j < U; ++j) {
                    const float4 h = bh[j], l = bl[j];
                    const uint32_t ah[4] = {__float_as_uint(h.x), __float_as_uint(h.y), __float_as_uint(h.z), __float_as_uint(h.w)};
                    const uint32_t al[4] = {__float_as_uint(l.x), __float_as_uint(l.y), __float_as_uint(l.z), __float_as_uint(l.w)};
                    uint32_t b0h, b0l, b1h, b1l;
                    split(xp[0], b0h, b0l); split(xp[20], b1h, b1l);
                    xp += 40; if (xp > smem + 7000) xp -= 6000;
                    float d[4], e[4];
                    mma0(d, al, b0l, b1l); mma0(e, ah, b0l, b1l); mma(d, al, b0h, b1h); mma(e, ah, b0h, b1h);
                    acc[0] += d[0] + e[0]; acc[1] += d[1] + e[1]; acc[2] += d[2] + e[2]; acc[3] += d[3] + e[3];
                    if (kk + j + U < steps) { bh[j] = ld(p + (size_t)(kk + j + U) * 64); bl[j] = ld(p + (size_t)(kk + j + U) * 64 + 32); }
                }
            }
            out += acc[0] + acc[1] + acc[2] + acc[3];
            base = (base + (size_t)nw * steps * 64 + 64 * 1031) % (total_f4 / 2);
        } else {                                           // FFMA, unsplit weights
            const float4* p = w + base + (size_t)warp * steps * 32 + lane;
            const float4* xp = reinterpret_cast<const float4*>(smem);
            float4 buf[U];
#pragma unroll
            for (int j = 0; j < U; ++j) buf[j] = ld(p + j * 32);
            float acc[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
            for (int kk = 0; kk < steps; kk += U) {
#pragma unroll
                for (int j = 0; j < U; ++j) {
                    const float4 a = buf[j];
                    if (kk + j + U < steps) buf[j] = ld(p + (size_t)(kk + j + U) * 32);
                    const float4 x0 = xp[0], x1 = xp[1], x2 = xp[2], x3 = xp[3], x4 = xp[4];     // x[k..k+3][0..4], 20 floats
                    xp += 5; if (xp > reinterpret_cast<const float4*>(smem) + 1700) xp -= 1500;
                    acc[0] += a.x * x0.x; acc[1] += a.x * x0.y; acc[2] += a.x * x0.z; acc[3] += a.x * x0.w; acc[4] += a.x * x1.x;
                    acc[0] += a.y * x1.y; acc[1] += a.y * x1.z; acc[2] += a.y * x1.w; acc[3] += a.y * x2.x; acc[4] += a.y * x2.y;
                    acc[0] += a.z * x2.z; acc[1] += a.z * x2.w; acc[2] += a.z * x3.x; acc[3] += a.z * x3.y; acc[4] += a.z * x3.z;
                    acc[0] += a.w * x3.w; acc[1] += a.w * x4.x; acc[2] += a.w * x4.y; acc[3] += a.w * x4.z; acc[4] += a.w * x4.w;
                }
            }
            out += acc[0] + acc[1] + acc[2] + acc[3] + acc[4];
            base = (base + (size_t)nw * steps * 32 + 32 * 1031) % (total_f4 / 2);
        }
        __syncthreads();
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[blockIdx.x] = (t1 - t0) / layers;
    if (out == 123.456f) sink[0] = out;
}
template <int KIND, int U>
void run(const float4* w, size_t total_f4, int steps, float* sink, long long* cyc, int share, int smem_bytes) {
    const int layers = 300, grid = 128, nt = 384;
    cudaFuncSetAttribute(k<KIND, U>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
    k<KIND, U><<<grid, nt, smem_bytes>>>(w, total_f4, steps, 10, sink, cyc, share);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a);
    k<KIND, U><<<grid, nt, smem_bytes>>>(w, total_f4, steps, layers, sink, cyc, share);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    const double wbytes = (double)(nt / 32) * steps * 512;          // unique fp32 weight bytes per block and layer
    printf("%-5s U=%d smem %3d KB share %3d steps/warp %3d: %7.0f cyc/layer  %6.1f cyc per 128 weights per warp  %6.1f GB/s of fp32 weights per SM  %s\n",
           KIND == 0 ? "MMA" : "FFMA", U, smem_bytes / 1024, share, steps, (double)cyc[0], (double)cyc[0] / steps, wbytes * layers / (ms * 1e-3) / 1e9,
           cudaGetLastError() == cudaSuccess ? "" : "ERR");
}
int main() {
    const size_t total_f4 = 6 * 1024 * 1024;        // 96 MB
    float4* w; float* sink; long long* cyc;
    cudaMalloc(&w, total_f4 * 16); cudaMalloc(&sink, 4); cudaMallocManaged(&cyc, 8 * 1024);
    cudaMemset(w, 0, total_f4 * 16);
    for (int smem_kb : {190, 64})
        for (int steps : {24, 8}) {
            const int sb = smem_kb * 1024;
            run<0, 2>(w, total_f4, steps, sink, cyc, 32, sb);
            run<0, 4>(w, total_f4, steps, sink, cyc, 32, sb);
            run<1, 2>(w, total_f4, steps, sink, cyc, 32, sb);
            run<1, 4>(w, total_f4, steps, sink, cyc, 32, sb);
            run<1, 8>(w, total_f4, steps, sink, cyc, 32, sb);
        }
    return 0;
}
