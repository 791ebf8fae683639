// Issue-rate micro-benchmark: legacy mma.sync shapes vs FFMA on sm_100a (cycles per warp-instruction per SM sub-partition).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
template <int MODE, int NACC>
__global__ void k(int iters, float* sink, long long* cyc) {
    float c[NACC][4];
    for (int i = 0; i < NACC; ++i) for (int j = 0; j < 4; ++j) c[i][j] = (float)(threadIdx.x + i + j);
    uint32_t a[4] = {threadIdx.x, threadIdx.x * 3u, 7u, 9u}, b0 = threadIdx.x * 5u, b1 = 11u;
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) {
            if (MODE == 0)
                asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
            else if (MODE == 1)
                asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
            else if (MODE == 2)
                asm volatile("mma.sync.aligned.m16n8k4.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                             : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3]) : "r"(a[0]), "r"(a[1]), "r"(b0));
            else {
#pragma unroll
                for (int j = 0; j < 4; ++j) asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(c[i][j]) : "f"(__uint_as_float(a[j])), "f"(__uint_as_float(b0)));
            }
        }
    }
    long long t1 = clock64();
    float s = 0.f;
    for (int i = 0; i < NACC; ++i) for (int j = 0; j < 4; ++j) s += c[i][j];
    if (s == 123.456f) sink[0] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
template <int MODE, int NACC>
void run(const char* name, int nt, float* sink, long long* cyc) {
    const int iters = 2000;
    k<MODE, NACC><<<148, nt>>>(iters, sink, cyc); cudaDeviceSynchronize();
    k<MODE, NACC><<<148, nt>>>(iters, sink, cyc); cudaDeviceSynchronize();
    const double per_warp_instr = (double)cyc[0] / ((double)iters * NACC * (MODE == 3 ? 4 : 1));
    const int warps_per_smsp = nt / 32 / 4 > 0 ? nt / 32 / 4 : 1;
    printf("%-28s threads %4d acc-chains %d: %7.2f cyc per instr per warp -> %6.2f cyc per instr per SMSP\n", name, nt, NACC, per_warp_instr,
           per_warp_instr / warps_per_smsp);
}
int main() {
    float* sink; long long* cyc; cudaMalloc(&sink, 4); cudaMallocManaged(&cyc, 64);
    for (int nt : {128, 384, 512}) {
        run<0, 1>("mma m16n8k8 tf32 (dep chain)", nt, sink, cyc);
        run<0, 4>("mma m16n8k8 tf32", nt, sink, cyc);
        run<0, 8>("mma m16n8k8 tf32", nt, sink, cyc);
        run<2, 8>("mma m16n8k4 tf32", nt, sink, cyc);
        run<1, 1>("mma m16n8k16 bf16 (dep chain)", nt, sink, cyc);
        run<1, 8>("mma m16n8k16 bf16", nt, sink, cyc);
        run<3, 1>("ffma (dep chain)", nt, sink, cyc);
        run<3, 8>("ffma", nt, sink, cyc);
    }
    return 0;
}
