// Prototype measurement: ONE dense layer of the forward chain as a TMA-fed tcgen05 pipeline, in a
// chain of dependent calls, against the register-fed mma.sync unit of the product kernel at the same shape.
//
//   y[f, r] = elu( sum_k W[k, f] x[k, r] )        f < 128 out-features of this block (UMMA M = 128), r < N rows (16 / 32),
//                                                  k < K in-features (256 / 672), fp32 accuracy via 3xTF32
//
// Per call: warp 0 streams the layer's weight panel W[K][128] (plain fp32, NOT pre-split) through a ring of TMA stages
// (MN-major, SWIZZLE_128B_ATOM_32B) -- the producer runs ahead across calls because weights do not depend on
// activations; warps 6-9 add the lo = w - trunc(w) copy of every stage; warp 1 issues lo.hi + hi.lo + hi.hi per k-step
// against the activations x (K-major SWIZZLE_128B tiles x_hi / x_lo in shared memory) into a TMEM accumulator; warps
// 2-5 read the accumulator (tcgen05.ld), apply the activation and write the NEXT call's x_hi / x_lo.  The dependent
// chain per call is therefore: last MMAs -> commit -> TMEM load -> activation -> shared-memory stores -> barrier.
// Build + run on the GPU box:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o /tmp/fwd_dense_tc tools/fwd_dense_tc.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

constexpr int MO = 128;                 // out-features per block
constexpr int RK = 32;                  // reduction rows per stage
constexpr int THREADS = 320;            // warp 0 TMA, warp 1 MMA, warps 2-5 epilogue, warps 6-9 split
constexpr int MAX_STAGES = 6;

struct Args {
    int K, N, layers, stages, panel_rows;     // panel_rows: rows of the weight buffer [rows][128]
    float* out;                               // [grid][128][N] final activations (check)
    long long* cycles;                        // [grid] cycles per call
    int verify;                               // 1: a single layer, x = test pattern
    int nodep;                                // 1: the MMA warp does not wait for the previous call's activations (streaming bound)
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t n) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(n)); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tWL:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra WD;\n\tbra WL;\n\tWD:\n\t}\n" ::"r"(smem_u32(b)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst),
                 "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
}
__device__ __forceinline__ uint64_t desc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)((lbo >> 4) & 0x3FFFu) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFFu) << 32) | (1ull << 46) |
           ((uint64_t)layout << 61);
}
__device__ __forceinline__ void umma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d), "l"(a), "l"(b),
                 "r"(idesc), "r"(acc)
                 : "memory");
}
__device__ __forceinline__ void commit(uint64_t* b) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(b)) : "memory");
}
// x tile: K-major, SWIZZLE_128B: k-blocks of 32 in-features; per k-block N rows of 128 bytes, 16-byte chunks XOR (row & 7)
__device__ __forceinline__ uint32_t x_off(int r, int k, int N) {
    return (uint32_t)((k >> 5) * (N * 128) + (r >> 3) * 1024 + (r & 7) * 128 + ((((k & 31) >> 2) ^ (r & 7)) << 4) + (k & 3) * 4);
}
__device__ __forceinline__ void split(float v, float& hi, float& lo) {
    hi = __uint_as_float(__float_as_uint(v) & 0xffffe000u);
    lo = __uint_as_float(__float_as_uint(v - hi) + 0x1000u);
}

__global__ void __launch_bounds__(THREADS, 1) dense_chain(const __grid_constant__ CUtensorMap tmw, const __grid_constant__ Args A) {
    extern __shared__ __align__(1024) uint8_t raw[];
    __shared__ uint64_t bar_full[MAX_STAGES], bar_conv[MAX_STAGES], bar_empty[MAX_STAGES], bar_accum, bar_act;
    __shared__ uint32_t tmem_s;
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~(uintptr_t)1023);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int K = A.K, N = A.N, nst = K / RK;                          // stages per call
    const uint32_t blk = RK * 128u, hi_bytes = 4u * blk, stage_bytes = 2u * hi_bytes;
    const uint32_t xt_bytes = (uint32_t)(K / 32) * (uint32_t)N * 128u;  // one activation tile (hi or lo)
    uint8_t* ring = smem;
    uint8_t* xhi = smem + (size_t)A.stages * stage_bytes;
    uint8_t* xlo = xhi + xt_bytes;
    if (threadIdx.x == 0) {
        for (int s = 0; s < A.stages; ++s) { mbar_init(&bar_full[s], 1); mbar_init(&bar_conv[s], 128); mbar_init(&bar_empty[s], 1); }
        mbar_init(&bar_accum, 1);
        mbar_init(&bar_act, 128);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_s)), "r"(32u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // initial activations
    for (int i = threadIdx.x; i < K * N; i += THREADS) {
        const int k = i / N, r = i % N;
        const float v = A.verify ? 0.01f * (float)((k * 7 + r * 13) % 31 - 15) : 0.05f * (float)((k + 3 * r) % 17 - 8);
        float h, l;
        split(v, h, l);
        *reinterpret_cast<float*>(xhi + x_off(r, k, N)) = h;
        *reinterpret_cast<float*>(xlo + x_off(r, k, N)) = l;
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_s;
    const int total = A.layers * nst;
    const int panel0 = (blockIdx.x & 3) * (A.panel_rows / 4);             // four blocks (a cluster) read different panels
    long long t0 = 0;

    if (warp == 0) {
        if (lane == 0) {
            for (int it = 0; it < total; ++it) {
                const int s = it % A.stages;
                const uint32_t ph = (uint32_t)(it / A.stages) & 1u;
                mbar_wait(&bar_empty[s], ph ^ 1u);
                mbar_expect_tx(&bar_full[s], hi_bytes);
                const int layer = it / nst, st = it % nst;
                const int row = A.verify ? st * RK : (panel0 + (layer * K + st * RK) % (A.panel_rows / 4 - K));
                const uint32_t base = smem_u32(ring) + (uint32_t)s * stage_bytes;
                for (int b = 0; b < 4; ++b) tma_load_3d(base + (uint32_t)b * blk, &tmw, 32 * b, 0, row, &bar_full[s]);
            }
        }
    } else if (warp == 1) {
        // A = W: MN-major (layout 1 = SWIZZLE_128B_BASE32B), B = x: K-major (layout 2 = SWIZZLE_128B); D f32; M = 128, N rows
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(MO >> 4) << 24);
        for (int layer = 0; layer < A.layers; ++layer) {
            if (!A.nodep) mbar_wait(&bar_act, (uint32_t)(layer & 1) ^ 1u);   // activations of this call are in place (phase -1 passes)
            for (int st = 0; st < nst; ++st) {
                const int it = layer * nst + st, s = it % A.stages;
                const uint32_t ph = (uint32_t)(it / A.stages) & 1u;
                mbar_wait(&bar_full[s], ph);
                mbar_wait(&bar_conv[s], ph);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                if (lane == 0) {
                    // descriptors: everything but the start address is constant; the 14-bit address field takes the k-step offset
                    const uint32_t a_hi = smem_u32(ring) + (uint32_t)s * stage_bytes;
                    const int k0 = st * RK;
                    const uint32_t xo = (uint32_t)((k0 >> 5) * (N * 128));                  // RK = 32: one k-block of x per stage
                    uint64_t dah = desc(a_hi, blk, 512u, 1u), dal = desc(a_hi + hi_bytes, blk, 512u, 1u);
                    uint64_t dbh = desc(smem_u32(xhi) + xo, 16u, 1024u, 2u), dbl = desc(smem_u32(xlo) + xo, 16u, 1024u, 2u);
#pragma unroll
                    for (int j = 0; j < RK / 8; ++j) {
                        umma(tmem, dal, dbh, idesc, (st > 0 || j > 0) ? 1u : 0u);
                        umma(tmem, dah, dbl, idesc, 1u);
                        umma(tmem, dah, dbh, idesc, 1u);
                        dah += 1024u >> 4; dal += 1024u >> 4;                               // next 8 reduction rows of W
                        dbh += 32u >> 4; dbl += 32u >> 4;                                   // next 8 in-features of x (32 bytes inside the swizzle span)
                    }
                    commit(&bar_empty[s]);
                    if (st == nst - 1) commit(&bar_accum);
                }
                __syncwarp();
            }
        }
    } else if (warp < 6) {
        // epilogue: TMEM -> registers -> activation -> next call's x_hi / x_lo (in-features 0 .. 127 of the next layer)
        const int q = warp & 3, f = 32 * q + lane;
        if (threadIdx.x == 64) t0 = clock64();
        for (int layer = 0; layer < A.layers; ++layer) {
            mbar_wait(&bar_accum, (uint32_t)(layer & 1));
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            uint32_t v[32];
            const uint32_t taddr = tmem + ((uint32_t)(32 * q) << 16);
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                         : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                           "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                         : "r"(taddr)
                         : "memory");
            if (N > 16)
                asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                             : "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
                               "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                             : "r"(taddr + 16u)
                             : "memory");
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            const bool last = layer == A.layers - 1;
#pragma unroll
            for (int r = 0; r < 32; ++r) {
                if (r < N) {
                    float y = __uint_as_float(v[r]);
                    y = y > 0.f ? y : expm1f(y);                       // ELU
                    if (last) A.out[((size_t)blockIdx.x * MO + f) * N + r] = y;
                    float h, l;
                    split(y * (A.verify ? 1.f : 0.25f), h, l);
                    *reinterpret_cast<float*>(xhi + x_off(r, f, N)) = h;
                    *reinterpret_cast<float*>(xlo + x_off(r, f, N)) = l;
                }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_arrive(&bar_act);
        }
        if (threadIdx.x == 64) A.cycles[blockIdx.x] = (clock64() - t0) / A.layers;
    } else {
        const int ct = threadIdx.x - 192;
        for (int it = 0; it < total; ++it) {
            const int s = it % A.stages;
            const uint32_t ph = (uint32_t)(it / A.stages) & 1u;
            mbar_wait(&bar_full[s], ph);
            const uint4* hi = reinterpret_cast<const uint4*>(ring + (size_t)s * stage_bytes);
            uint4* lo = reinterpret_cast<uint4*>(ring + (size_t)s * stage_bytes + hi_bytes);
#pragma unroll 4
            for (int c = ct; c < (int)(hi_bytes >> 4); c += 128) {
                const uint4 v = hi[c];
                uint4 l;
                l.x = __float_as_uint(__uint_as_float(v.x) - __uint_as_float(v.x & 0xffffe000u)) + 0x1000u;
                l.y = __float_as_uint(__uint_as_float(v.y) - __uint_as_float(v.y & 0xffffe000u)) + 0x1000u;
                l.z = __float_as_uint(__uint_as_float(v.z) - __uint_as_float(v.z & 0xffffe000u)) + 0x1000u;
                l.w = __float_as_uint(__uint_as_float(v.w) - __uint_as_float(v.w & 0xffffe000u)) + 0x1000u;
                lo[c] = l;
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_arrive(&bar_conv[s]);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(32u) : "memory");
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                             CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
    const int panel_rows = 4 * 32768;                                  // [131072][128] floats = 64 MB, L2 resident
    float* w;
    cudaMalloc(&w, (size_t)panel_rows * MO * 4);
    std::vector<float> hw((size_t)1024 * MO);
    for (size_t i = 0; i < hw.size(); ++i) hw[i] = 0.02f * (float)((int)((i * 2654435761u) >> 20) % 41 - 20) / 20.f + 1e-4f * (float)(i % 7);
    cudaMemset(w, 0, (size_t)panel_rows * MO * 4);
    cudaMemcpy(w, hw.data(), hw.size() * 4, cudaMemcpyHostToDevice);
    for (int p = 1; p < panel_rows / 1024; ++p) cudaMemcpy(w + (size_t)p * 1024 * MO, w, 1024 * MO * 4, cudaMemcpyDeviceToDevice);
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qr;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr);
    CUtensorMap tm;
    cuuint64_t dims[3] = {(cuuint64_t)MO, 1, (cuuint64_t)panel_rows};
    cuuint64_t strides[2] = {(cuuint64_t)MO * 4, (cuuint64_t)MO * 4};
    cuuint32_t box[3] = {32, 1, RK}, es[3] = {1, 1, 1};
    if (((EncodeFn)fn)(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, w, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B,
                       CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) {
        printf("tensor map failed\n");
        return 1;
    }
    float* out;
    long long* cyc;
    cudaMalloc(&out, (size_t)148 * MO * 32 * 4);
    cudaMallocManaged(&cyc, 148 * 8);
    cudaFuncSetAttribute(dense_chain, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    auto smem_for = [](int K, int N, int stages) { return (size_t)stages * 2 * 4 * RK * 128 + 2 * (size_t)(K / 32) * N * 128 + 1024; };
    // ---- correctness of one layer against the host (double)
    for (int N : {16, 32}) {
        Args a{256, N, 1, 4, panel_rows, out, cyc, 1, 0};
        dense_chain<<<1, THREADS, smem_for(256, N, 4)>>>(tm, a);
        cudaError_t e = cudaDeviceSynchronize();
        std::vector<float> ho((size_t)MO * N);
        cudaMemcpy(ho.data(), out, ho.size() * 4, cudaMemcpyDeviceToHost);
        double maxerr = 0, maxv = 0;
        for (int f = 0; f < MO; ++f)
            for (int r = 0; r < N; ++r) {
                double acc = 0;
                for (int k = 0; k < 256; ++k) acc += (double)hw[(size_t)k * MO + f] * (0.01 * (double)((k * 7 + r * 13) % 31 - 15));
                const double y = acc > 0 ? acc : expm1(acc);
                maxerr = fmax(maxerr, fabs(y - (double)ho[(size_t)f * N + r]));
                maxv = fmax(maxv, fabs(y));
            }
        printf("verify K=256 N=%d: max |err| %.3e of max |y| %.3e  (%s)\n", N, maxerr, maxv, e == cudaSuccess ? "ok" : cudaGetErrorString(e));
    }
    // ---- chains of dependent calls
    for (int grid : {1, 128}) {
        for (int K : {256, 672}) {
            for (int N : {16, 32}) {
                for (int stages : {3, 4, 8}) {
                  for (int nodep : {0, 1}) {
                    if (smem_for(K, N, stages) > 220 * 1024) continue;
                    const int layers = 200;
                    Args a{K, N, layers, stages, panel_rows, out, cyc, 0, nodep};
                    dense_chain<<<grid, THREADS, smem_for(K, N, stages)>>>(tm, a);
                    cudaDeviceSynchronize();
                    cudaEvent_t e0, e1;
                    cudaEventCreate(&e0); cudaEventCreate(&e1);
                    cudaEventRecord(e0);
                    dense_chain<<<grid, THREADS, smem_for(K, N, stages)>>>(tm, a);
                    cudaEventRecord(e1);
                    cudaError_t e = cudaDeviceSynchronize();
                    float ms;
                    cudaEventElapsedTime(&ms, e0, e1);
                    const double us = ms * 1e3 / layers, bytes = (double)K * MO * 4;
                    printf("grid %3d  K=%3d out=128 rows N=%2d stages %d %s: %6.2f us per call (%5.0f cycles in-kernel), %6.1f GB/s of weights per SM, "
                           "%7.1f GB/s aggregate  %s\n", grid, K, N, stages, nodep ? "independent calls" : "dependent chain  ", us, (double)cyc[0], bytes / us / 1e3,
                           bytes * grid / us / 1e3, e == cudaSuccess ? "" : cudaGetErrorString(e));
                  }
                }
            }
        }
    }
    return 0;
}
