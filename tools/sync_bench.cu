// Micro-benchmark of the synchronisation primitives the sequence kernel uses per dense layer.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>
__global__ void k_syncthreads(int iters, long long* out) {
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) __syncthreads();
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = (clock64() - t0) / iters;
}
__global__ void k_cluster(int iters, long long* out) {
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
        asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
        asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
    }
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = (clock64() - t0) / iters;
}
__global__ void k_cluster_relaxed(int iters, long long* out) {
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
        asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory");
        asm volatile("barrier.cluster.wait.aligned;" ::: "memory");
    }
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = (clock64() - t0) / iters;
}
__global__ void k_chase(const int* p, int iters, long long* out) {
    int j = threadIdx.x;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) j = __ldg(p + j);
    if (threadIdx.x == 0) { out[0] = (clock64() - t0) / iters; out[1] = j; }
}
int main() {
    long long* out; cudaMallocManaged(&out, 64);
    const int N = 1 << 22;   // 16 MB pointer-chase table (L2 resident, larger than L1)
    int* p; cudaMalloc(&p, N * 4);
    int* h = (int*)malloc(N * 4);
    for (int i = 0; i < N; ++i) h[i] = (int)(((long long)i * 104729 + 12345) % N);
    cudaMemcpy(p, h, N * 4, cudaMemcpyHostToDevice);
    for (int nt : {256, 288}) {
        k_syncthreads<<<120, nt>>>(10000, out); cudaDeviceSynchronize();
        printf("__syncthreads (%d threads): %lld cycles\n", nt, out[0]);
    }
    for (int C : {2, 3, 4, 8}) {
        for (int variant = 0; variant < 2; ++variant) {
            cudaLaunchConfig_t cfg = {}; cfg.gridDim = dim3(C * (120 / C)); cfg.blockDim = dim3(288);
            cudaLaunchAttribute a[1]; a[0].id = cudaLaunchAttributeClusterDimension; a[0].val.clusterDim = {(unsigned)C, 1, 1};
            cfg.attrs = a; cfg.numAttrs = 1;
            if (variant == 0) cudaLaunchKernelEx(&cfg, k_cluster, 10000, out); else cudaLaunchKernelEx(&cfg, k_cluster_relaxed, 10000, out);
            cudaError_t e = cudaDeviceSynchronize();
            printf("cluster barrier %s C=%d (288 threads): %lld cycles %s\n", variant ? "relaxed" : "release/acquire", C, out[0], e ? cudaGetErrorString(e) : "");
        }
    }
    k_chase<<<1, 32>>>(p, 2000, out); cudaDeviceSynchronize();
    k_chase<<<1, 32>>>(p, 2000, out); cudaDeviceSynchronize();
    printf("dependent __ldg from L2 (16 MB table): %lld cycles\n", out[0]);
    int dev_clock; cudaDeviceGetAttribute(&dev_clock, cudaDevAttrClockRate, 0);
    printf("clock rate attr: %d kHz\n", dev_clock);
    return 0;
}
