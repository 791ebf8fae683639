// sqair_wgrad_tc.cu -- the weight-gradient GEMM of the backward pass on Blackwell's 5th-generation tensor cores.
//
//   dW[k, n] += sum_m X[m, k] * dY[m, n]          M = frames x rows x slots (6 400 at BASELINE configs[1]), K, N <= 672
//
// (reference: the matmul gradients inside `opt.compute_gradients(target)`, sqair/model.py:160; cuBLAS SGEMM under TF1.)
// Both operands are "MN-major" for this product -- the reduction index m is the slow index of X and dY in memory -- which
// tcgen05 supports for tf32 (instruction-descriptor bits a_major / b_major).  One CTA owns a 128 x N tile of dW and a
// slice of the reduction (split-M, partial tiles meet through red.global.add):
//   * warp 0 (one elected lane): TMA producer.  A stage holds RK reduction rows: 4 + N/32 boxes of [RK rows][32 floats]
//     (cp.async.bulk.tensor.3d, SWIZZLE_128B_ATOM_32B) -- exactly the canonical MN-major tf32 UMMA layout, no reshuffling.
//   * warps 2-5: fp32 fidelity.  kind::tf32 reads only the upper 19 bits of every word (measured: feeding the raw fp32
//     words or words with the low 13 bits cleared gives the same result), so the TMA-written tile IS the hi operand and
//     the warps only add a second copy lo = x - hi; three products per k-step (lo.hi + hi.lo + hi.hi, the 3xTF32 scheme)
//     recover fp32 accuracy (one product alone: 8e-4 relative error).  The same warps drain the accumulator at the end
//     (tcgen05.ld 32 lanes x 32 columns per warp) and add it to dW.
//   * warp 1 (one elected lane): tcgen05.mma.cta_group::1.kind::tf32, M = 128, N <= 256, K = 8 per instruction,
//     accumulator in TMEM; tcgen05.commit releases the stage to the producer and finally signals the epilogue.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include <mutex>
#include <utility>
#include <vector>

#include "sqair_internal.h"

using sqi::fail;

namespace {

constexpr int TC_BM = 128;            // rows of dW per CTA (= UMMA M)
constexpr int TC_MAX_STAGES = 8;     // pipeline depth is chosen per launch from the shared memory a stage needs
constexpr int TC_THREADS = 192;       // warp 0: TMA, warp 1: MMA + TMEM owner, warps 2-5: split + epilogue
constexpr int TC_CONV_THREADS = 128;
constexpr int TC_MAX_N = 256;
constexpr int TC_MAX_RK = 32;
constexpr int TC_SMEM_BUDGET = 216 * 1024;

struct TcArgs {
    float* dw;
    int ldw, K, N;
    int nkb;            // number of RK-row blocks of the reduction
    int kb_per_cta;
    int RK;             // reduction rows per stage (multiple of 8)
    int bz;             // box extent along the outer row index (RK / by)
    int ntile;          // UMMA N of this launch (multiple of 16, <= 256)
    int nblk_b;         // 32-column blocks of the dY tile
    int tmem_cols;
    int stages;
    int vec4;           // dW rows are 16-byte aligned and N is a multiple of 4: vector reductions
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra WAIT_DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "WAIT_DONE:\n\t"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst),
                 "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
}
// Shared-memory matrix descriptor (cute/arch/mma_sm100_desc.hpp: SmemDescriptor) for an MN-major tf32 operand.  The only
// layout the tensor core accepts for that case is SWIZZLE_128B_BASE32B (layout type 1; with plain SWIZZLE_128B the MMA
// returns zeros): rows of 128 bytes = 32 elements along M / N, the four 32-byte chunks of a row XOR-ed with (row mod 4),
// atoms of 4 reduction rows.  Leading byte offset = distance between 32-element blocks along M / N, stride byte offset =
// distance between consecutive 4-row atoms (512: rows are contiguous), both in 16-byte units; version 1 (Blackwell).
// TMA writes exactly this with CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B.
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout_type = 1u) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | (1ull << 46) | ((uint64_t)layout_type << 61);
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}


__global__ void __launch_bounds__(TC_THREADS, 1)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap tmx, const __grid_constant__ CUtensorMap tmy, const __grid_constant__ TcArgs A) {
    extern __shared__ __align__(1024) uint8_t tc_smem_raw[];
    __shared__ uint64_t bar_full[TC_MAX_STAGES], bar_conv[TC_MAX_STAGES], bar_empty[TC_MAX_STAGES], bar_accum;
    __shared__ uint32_t tmem_base_s;
    // dynamic shared memory is only guaranteed 16-byte aligned: round up to the 1024 bytes the swizzle atoms need
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(tc_smem_raw) + 1023) & ~(uintptr_t)1023);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int k0 = blockIdx.x * TC_BM, n0 = blockIdx.y * A.ntile;
    const int kb0 = blockIdx.z * A.kb_per_cta;
    const int nk = min(A.kb_per_cta, A.nkb - kb0);
    const uint32_t blk_bytes = (uint32_t)A.RK * 128u;                    // one [RK][32 floats] box
    const uint32_t hi_bytes = (4u + (uint32_t)A.nblk_b) * blk_bytes;    // X tile (4 blocks) + dY tile
    const uint32_t stage_bytes = 2u * hi_bytes;                          // + the lo copies

    if (threadIdx.x == 0) {
        for (int s = 0; s < A.stages; ++s) {
            mbar_init(&bar_full[s], 1);
            mbar_init(&bar_conv[s], TC_CONV_THREADS);
            mbar_init(&bar_empty[s], 1);
        }
        mbar_init(&bar_accum, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmx)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmy)) : "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"((uint32_t)A.tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_s;

    if (warp == 0) {
        if (lane == 0) {
            for (int it = 0; it < nk; ++it) {
                const int s = it % A.stages;
                const uint32_t ph = (uint32_t)(it / A.stages) & 1u;
                mbar_wait(&bar_empty[s], ph ^ 1u);
                mbar_expect_tx(&bar_full[s], hi_bytes);
                const uint32_t base = smem_u32(smem) + (uint32_t)s * stage_bytes;
                const int cz = (kb0 + it) * A.bz;
                for (int b = 0; b < 4; ++b) tma_load_3d(base + (uint32_t)b * blk_bytes, &tmx, k0 + 32 * b, 0, cz, &bar_full[s]);
                for (int b = 0; b < A.nblk_b; ++b) tma_load_3d(base + (uint32_t)(4 + b) * blk_bytes, &tmy, n0 + 32 * b, 0, cz, &bar_full[s]);
            }
        }
    } else if (warp == 1) {
        // instruction descriptor (cute/arch/mma_sm100_desc.hpp: InstrDescriptor): D = f32, A = B = tf32, both MN-major
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(A.ntile >> 3) << 17) |
                         ((uint32_t)(TC_BM >> 4) << 24);
        const uint32_t lbo = blk_bytes, sbo = 512u;
        for (int it = 0; it < nk; ++it) {
            const int s = it % A.stages;
            const uint32_t ph = (uint32_t)(it / A.stages) & 1u;
            mbar_wait(&bar_full[s], ph);
            mbar_wait(&bar_conv[s], ph);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (lane == 0) {
                const uint32_t a_hi = smem_u32(smem) + (uint32_t)s * stage_bytes, b_hi = a_hi + 4u * blk_bytes;
                const uint32_t a_lo = a_hi + hi_bytes, b_lo = b_hi + hi_bytes;
                for (int j = 0; j < A.RK / 8; ++j) {
                    const uint32_t off = (uint32_t)j * 1024u;             // 8 reduction rows of 128 bytes
                    const uint64_t dah = umma_desc(a_hi + off, lbo, sbo), dal = umma_desc(a_lo + off, lbo, sbo);
                    const uint64_t dbh = umma_desc(b_hi + off, lbo, sbo), dbl = umma_desc(b_lo + off, lbo, sbo);
                    umma_tf32(tmem_base, dal, dbh, idesc, (it > 0 || j > 0) ? 1u : 0u);      // small terms first
                    umma_tf32(tmem_base, dah, dbl, idesc, 1u);
                    umma_tf32(tmem_base, dah, dbh, idesc, 1u);
                }
                umma_commit(&bar_empty[s]);                               // the stage is free once these MMAs have read it
                if (it == nk - 1) umma_commit(&bar_accum);
            }
            __syncwarp();
        }
    } else {
        const int ct = threadIdx.x - 64;                                  // 0 .. 127
        for (int it = 0; it < nk; ++it) {
            const int s = it % A.stages;
            const uint32_t ph = (uint32_t)(it / A.stages) & 1u;
            mbar_wait(&bar_full[s], ph);
            uint4* hi = reinterpret_cast<uint4*>(smem + (size_t)s * stage_bytes);
            uint4* lo = reinterpret_cast<uint4*>(smem + (size_t)s * stage_bytes + hi_bytes);
            const int chunks = (int)(hi_bytes >> 4);
#pragma unroll 4
            for (int c = ct; c < chunks; c += TC_CONV_THREADS) {
                const uint4 v = hi[c];
                uint4 h, l;
                h.x = v.x & 0xffffe000u; h.y = v.y & 0xffffe000u; h.z = v.z & 0xffffe000u; h.w = v.w & 0xffffe000u;
                // lo = x - hi is exact in fp32; + 0x1000 rounds it to tf32 precision (the tensor core truncates)
                l.x = __float_as_uint(__uint_as_float(v.x) - __uint_as_float(h.x)) + 0x1000u;
                l.y = __float_as_uint(__uint_as_float(v.y) - __uint_as_float(h.y)) + 0x1000u;
                l.z = __float_as_uint(__uint_as_float(v.z) - __uint_as_float(h.z)) + 0x1000u;
                l.w = __float_as_uint(__uint_as_float(v.w) - __uint_as_float(h.w)) + 0x1000u;
                lo[c] = l;
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the MMA's reads
            mbar_arrive(&bar_conv[s]);
        }
        // epilogue: TMEM -> registers -> dW.  A warp may only touch the 32 TMEM lanes of its quarter (warp id mod 4).
        if (nk > 0) {
            mbar_wait(&bar_accum, 0);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const int q = warp & 3;
            for (int cb = 0; cb < A.ntile; cb += 32) {
                uint32_t r[32];
                const uint32_t taddr = tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)cb;
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, "
                    "%19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                    : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                      "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
                      "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
                      "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                    : "r"(taddr)
                    : "memory");
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                // transpose the warp's 32 x 32 block through shared memory (the pipeline stages are idle now) so that one
                // instruction adds 32 consecutive floats of ONE row of dW: 4 sectors per warp instruction instead of 32
                float* tile = reinterpret_cast<float*>(smem) + (warp - 2) * (32 * 33);
#pragma unroll
                for (int i = 0; i < 32; ++i) tile[lane * 33 + i] = __uint_as_float(r[i]);
                __syncwarp();
                const int rbase = k0 + 32 * q;
                if (A.vec4) {
                    // 8 lanes x 16 bytes cover the 32 columns of a row, the warp covers 4 rows per instruction
                    const int c4 = (lane & 7) * 4, r4 = lane >> 3;
                    const int col = n0 + cb + c4;
                    const bool col_ok = cb + c4 < A.ntile && col < A.N;          // (ntile and N are multiples of 4 here)
#pragma unroll
                    for (int rr = 0; rr < 32; rr += 4) {
                        const int row = rbase + rr + r4;
                        if (col_ok && row < A.K) {
                            const float* t = tile + (rr + r4) * 33 + c4;
                            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(A.dw + (size_t)row * A.ldw + col), "f"(t[0]), "f"(t[1]),
                                         "f"(t[2]), "f"(t[3])
                                         : "memory");
                        }
                    }
                } else {
                    const int col = n0 + cb + lane;
                    const bool col_ok = cb + lane < A.ntile && col < A.N;
#pragma unroll 8
                    for (int rr = 0; rr < 32; ++rr)
                        if (col_ok && rbase + rr < A.K) atomicAdd(A.dw + (size_t)(rbase + rr) * A.ldw + col, tile[rr * 33 + lane]);
                }
                __syncwarp();
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)A.tmem_cols) : "memory");
    }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qr;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qr) == cudaSuccess && qr == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    });
    return fn;
}

int lcm8(int ny) {
    int r = ny;
    while (r % 8) r += ny;
    return r;
}

// rows m = z * ny + y at p + z * outer + y * inner (floats): a 3-D tensor {width, ny, M / ny}; box {32, ny, bz}.
// Encoding a tensor map costs a few microseconds of host time and a training step asks for the same ~80 maps every
// iteration (the buffers do not move), so the last maps are kept.
struct MapKey {
    const float* p;
    int64_t outer, inner;
    int ny, width, M, bz;
    bool operator==(const MapKey& o) const {
        return p == o.p && outer == o.outer && inner == o.inner && ny == o.ny && width == o.width && M == o.M && bz == o.bz;
    }
};
std::mutex g_map_mutex;
std::vector<std::pair<MapKey, CUtensorMap>> g_maps;

bool make_map(CUtensorMap* map, const sqi::TcOperand& op, int width, int M, int bz) {
    const int ny = op.ny > 1 ? op.ny : 1;
    const MapKey key{op.p, op.outer, ny > 1 ? op.inner : 0, ny, width, M, bz};
    {
        std::lock_guard<std::mutex> lock(g_map_mutex);
        for (const auto& kv : g_maps)
            if (kv.first == key) { *map = kv.second; return true; }
    }
    cuuint64_t dims[3] = {(cuuint64_t)width, (cuuint64_t)ny, (cuuint64_t)(M / ny)};
    cuuint64_t strides[2] = {(cuuint64_t)(ny > 1 ? op.inner : op.outer) * 4u, (cuuint64_t)op.outer * 4u};
    cuuint32_t box[3] = {32u, (cuuint32_t)ny, (cuuint32_t)bz};
    cuuint32_t estr[3] = {1u, 1u, 1u};
    if (encode_fn()(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(op.p), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return false;
    std::lock_guard<std::mutex> lock(g_map_mutex);
    if (g_maps.size() >= 512) g_maps.erase(g_maps.begin(), g_maps.begin() + 256);
    g_maps.emplace_back(key, *map);
    return true;
}

bool operand_ok(const sqi::TcOperand& op, int M) {
    const int ny = op.ny > 1 ? op.ny : 1;
    if ((reinterpret_cast<uintptr_t>(op.p) & 15u) != 0 || op.outer <= 0 || (op.outer & 3) != 0 || M % ny != 0) return false;
    if (ny > 1 && (op.inner <= 0 || (op.inner & 3) != 0)) return false;
    return true;
}

}  // namespace


bool sqi::wgrad_tc_supported(const TcOperand& x, const TcOperand& dy, int M, int K, int N) {
    if (sqi::env_int("SQAIR_NO_TC")) return false;
    if (!encode_fn()) return false;
    const int ny = x.ny > 1 ? x.ny : 1;
    if ((dy.ny > 1 ? dy.ny : 1) != ny) return false;
    if (K < 32 || N < 32 || M < 256) return false;                 // tiny layers: the register-fed mma.sync kernel
    if (lcm8(ny) > TC_MAX_RK) return false;
    return operand_ok(x, M) && operand_ok(dy, M);
}

int sqi::wgrad_tc(const TcOperand& x, const TcOperand& dy, float* dw, int ldw, int M, int K, int N, cudaStream_t st) {
    const int ny = x.ny > 1 ? x.ny : 1;
    int RK = lcm8(ny);
    while (RK < 16) RK *= 2;                      // short stages, deep pipeline: the loop is bound by load latency, not by issue
    TcArgs A;
    memset(&A, 0, sizeof(A));
    A.dw = dw; A.ldw = ldw; A.K = K; A.N = N;
    A.RK = RK; A.bz = RK / ny;
    A.nkb = (M + RK - 1) / RK;
    const int ntiles = (N + TC_MAX_N - 1) / TC_MAX_N;
    const int nper = (N + ntiles - 1) / ntiles;
    A.ntile = (nper + 15) / 16 * 16;
    A.nblk_b = (A.ntile + 31) / 32;
    A.vec4 = ((reinterpret_cast<uintptr_t>(dw) & 15u) == 0 && ldw % 4 == 0 && N % 4 == 0) ? 1 : 0;
    A.tmem_cols = A.ntile <= 32 ? 32 : A.ntile <= 64 ? 64 : A.ntile <= 128 ? 128 : 256;
    const int ktiles = (K + TC_BM - 1) / TC_BM;
    int device = 0, sms = 148;
    cudaGetDevice(&device);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    int splits = sms / (ktiles * ntiles);
    if (splits < 1) splits = 1;
    if (splits > A.nkb) splits = A.nkb;
    A.kb_per_cta = (A.nkb + splits - 1) / splits;
    splits = (A.nkb + A.kb_per_cta - 1) / A.kb_per_cta;
    CUtensorMap tmx, tmy;
    if (!make_map(&tmx, x, K, M, A.bz) || !make_map(&tmy, dy, N, M, A.bz)) return fail(SQAIR_EINVAL, "cuTensorMapEncodeTiled failed for the weight-gradient operands");
    const size_t stage_bytes = 2u * (4u + (size_t)A.nblk_b) * (size_t)RK * 128u;
    A.stages = (int)((TC_SMEM_BUDGET - 1024) / stage_bytes);
    if (A.stages > TC_MAX_STAGES) A.stages = TC_MAX_STAGES;
    if (A.stages > A.kb_per_cta) A.stages = A.kb_per_cta;
    if (A.stages < 1) return fail(SQAIR_EUNSUPPORTED, "weight-gradient tile does not fit shared memory");
    size_t smem = (size_t)A.stages * stage_bytes + 1024u;
    if (smem < 4 * 32 * 33 * sizeof(float) + 1024u) smem = 4 * 32 * 33 * sizeof(float) + 1024u;       // epilogue transpose tiles
    static std::once_flag once;
    std::call_once(once, [] { cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BUDGET); });
    wgrad_tc_kernel<<<dim3(ktiles, ntiles, splits), TC_THREADS, smem, st>>>(tmx, tmy, A);
    CUDA_TRY(cudaGetLastError());
    return SQAIR_OK;
}
