// sqair_train.cu -- CUDA backend of the backward pass (sqair_backward.h) and its C ABI.
//
// The frame recursion of the reverse program runs as ONE persistent cluster kernel (`bwd_program_kernel`): the host records
// the operations the driver issues (row stages = the hand-written adjoints of the element-wise stages, dgrad products
// dX = dY . W^T on the tensor cores, clears) into a table, every cluster of thread blocks interprets it for the rows it owns,
// cluster barriers separate dependent operations.  The weight gradients follow as GEMMs over M = T x rows x slots: tcgen05
// (sqair_wgrad_tc.cu) where TMA can describe the operands, `wgrad_addr_kernel` (mma.sync, fp32-faithful tf32 split) otherwise,
// column sums for the biases, all on a few side streams; then the scatter back to the reference's variables.
// Kept for A/B measurements and tests (SQAIR_BWD_LAUNCHES=1): the same program as one launch per operation
// (`bwd_stage_kernel<STAGE>`, `dgrad_kernel`: fp32 FFMA, programmatic dependent launch).  Also here: the fused optimiser
// update and the sprite renderer of the data path.
#include <cuda_runtime.h>
#include <algorithm>
#include <math.h>
#include <stddef.h>
#include <stdio.h>
#include <string.h>
#include <mutex>
#include <string>
#include <utility>
#include <vector>

#include "sqair_backward.h"
#include "sqair_internal.h"

using namespace sq;
using sqi::Shape;
using sqi::fail;

// ---------------------------------------------------------------------------------------------
// Programmatic dependent launch.  The reverse program is a chain of ~1 450 short kernels; with the stream-serialisation
// attribute a kernel is scheduled while its predecessor still runs.  Every kernel launched this way releases its
// successor immediately (pdl_trigger) and blocks at pdl_wait until the predecessor has completed and flushed its
// writes -- nothing that depends on earlier kernels may be touched before the wait.  What may: kernel arguments and the
// weight copies (constant during a backward call), which `dgrad_kernel` fetches ahead of the wait.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

template <class... KArgs, class... Args>
static cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

// ---------------------------------------------------------------------------------------------
// row stages
// ---------------------------------------------------------------------------------------------
struct DevEx {
    int tid, nt;
    float* scratch;
    float* red;          // [4][32]
    int bar_id;          // 0: the row owns the whole thread block; else the named barrier of the row's `nt` threads
    __device__ __forceinline__ void sync() {
        if (bar_id == 0) __syncthreads();
        else asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "r"(nt) : "memory");
    }
    // dst[0..3] += the warp's sums of v[0..3] (all 32 lanes must call; one shared-memory atomic per warp and value)
    __device__ __forceinline__ void warp_add4(float* dst, const float (&v)[4]) {
        // most warps of the canvas stage lie outside a glimpse: nothing to add (exact zeros), skip the 20 shuffles
        if (!__any_sync(0xffffffffu, v[0] != 0.f || v[1] != 0.f || v[2] != 0.f || v[3] != 0.f)) return;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            float x = v[q];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
            if ((tid & 31) == 0) atomicAdd(dst + q, x);
        }
    }
    __device__ __forceinline__ void sum4(float (&v)[4]) {
        const int lane = tid & 31, warp = tid >> 5, nw = (nt + 31) >> 5;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            float x = v[q];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
            if (lane == 0) red[q * 32 + warp] = x;
        }
        sync();
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            float x = 0.f;
            for (int w = 0; w < nw; ++w) x += red[q * 32 + w];
            v[q] = x;
        }
        sync();
    }
};

template <int STAGE>
__global__ void bwd_stage_kernel(const __grid_constant__ BwdCtx c, int t, int s) {
    extern __shared__ float bw_smem[];
    __shared__ float red[4 * 32];
    pdl_trigger();
    pdl_wait();
    DevEx ex;
    ex.tid = threadIdx.x; ex.nt = blockDim.x; ex.scratch = bw_smem; ex.red = red; ex.bar_id = 0;
    bw_stage<STAGE>(c, ex, t, s, (int)blockIdx.x);
}

// ---------------------------------------------------------------------------------------------
// dgrad: dX_seg[m, k - k0] (=|+=) sum_n A[m, n] W[k, n],  A = a * act'(y)   (fp32 FFMA)
// M is 160 .. 640 rows and the call sits on the dependent chain of the reverse program, so the kernel is built for
// latency, not throughput.  A block owns 32 rows x 32 input features.  N is walked in phases of 256: all threads first
// fetch the phase's operand tiles with independent loads (A[32][256] with the activation derivative applied, and the
// slice Wt[256][32] of the transposed matrix, 16-byte loads) into shared memory -- one memory round trip --, then the
// eight warps split the 256 n (a lane owns one row and 32 accumulators; weights are shared-memory broadcasts).
// Partial sums of the warps meet in shared memory (aliasing the tiles).
// ---------------------------------------------------------------------------------------------
constexpr int DG_BM = 32, DG_BK = 32, DG_WARPS = 16, DG_THREADS = 32 * DG_WARPS, DG_NP = 256, DG_LDA = DG_BM + 4;
constexpr int DG_SMEM_FLOATS = DG_NP * DG_LDA + DG_NP * DG_BK;             // As[n][36] + Ws[n][32]  (69.6 KB)
static_assert(DG_WARPS * DG_BM * (DG_BK + 1) <= DG_SMEM_FLOATS, "partial sums alias the operand tiles");

struct DgradDev {
    DgradArgs a;
    const float* wt;       // [N][ldt] transposed virtual matrix, rows padded to 16 bytes
    int ldt;
};

__device__ __forceinline__ const float* addr_row(const Addr& a, int m, int ny) {
    return a.p + (size_t)(m / ny) * a.outer + (size_t)(m % ny) * a.inner;
}

// The kernel runs once per call on a cold instruction cache, so it is written for a SMALL code footprint (the first,
// fully unrolled version spent 5 of 7 issue-stall cycles on instruction fetch, ncu `stalled_no_instruction`): rolled
// loops around short unrolled bodies, the activation derivative resolved at compile time.
template <int ACT>
__device__ __forceinline__ float act_deriv_t(float y, float scale, float add) {
    if (ACT == ACT_ELU) return y > 0.f ? 1.f : y + 1.f;
    if (ACT == ACT_TANH) return 1.f - y * y;
    if (ACT == ACT_SIGMOID) { const float s_ = y / scale; return scale * s_ * (1.f - s_); }
    if (ACT == ACT_SOFTPLUS) return -expm1f(-(y - add));
    return 1.f;
}

// One 32 x 32 tile (bx = feature tile, by = row tile).  PDL: the stand-alone kernel releases its successor at once and
// waits for its predecessor after the weight loads are in flight; inside the persistent reverse-program kernel the
// grid barrier has already ordered everything.
template <int ACT, bool PDL>
__device__ __forceinline__ void dgrad_tile(const DgradDev& D, const int bx, const int by, float* dg_smem) {
    const DgradArgs& A = D.a;
    float* As = dg_smem;                                  // [DG_NP][DG_LDA]  (rows of the tile contiguous: 16-byte loads of 4 rows)
    float* Ws = dg_smem + DG_NP * DG_LDA;                 // [DG_NP][DG_BK]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int m0 = by * DG_BM, kb = bx * DG_BK;
    const bool has_k = A.nseg > 0 && kb < A.K;
    if (PDL) pdl_trigger();
    if (!has_k && bx != 0) return;
    const bool store_dy = A.dy.p != nullptr && bx == 0;
    const int nq = has_k ? min(DG_BK / 4, (D.ldt - kb) / 4) : 0;          // 16-byte groups of this tile inside the padded row
    const int rg = lane >> 2, fq = lane & 3;                 // compute: rows 4 rg .. 4 rg + 3, features 8 fq .. 8 fq + 7
    float acc[4][8];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
    for (int nb = 0; nb < A.N; nb += DG_NP) {
        const int np = min(DG_NP, A.N - nb);
        if (nb) __syncthreads();
        // operand fetch: every thread issues a batch of independent loads before it touches the first result
        if (has_k) {                                             // weight slice Wt[nb .. nb+np)[kb .. kb+32): 4 x 16 bytes per thread
            constexpr int NW = DG_NP * (DG_BK / 4) / DG_THREADS;
            float4 wv[NW];
#pragma unroll
            for (int j = 0; j < NW; ++j) {
                const int i = tid + j * DG_THREADS, n = i >> 3, q = i & 7;
                wv[j] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (n < np && q < nq) wv[j] = __ldg(reinterpret_cast<const float4*>(D.wt + (size_t)(nb + n) * D.ldt + kb) + q);
            }
            if (PDL && nb == 0) pdl_wait();                      // the weight loads above are in flight across the wait
#pragma unroll
            for (int j = 0; j < NW; ++j) {
                const int i = tid + j * DG_THREADS, n = i >> 3, q = i & 7;
                if (n < np) *reinterpret_cast<float4*>(Ws + n * DG_BK + 4 * q) = wv[j];
            }
        } else if (PDL && nb == 0) {
            pdl_wait();
        }
        // A tile: warp w owns rows 2w, 2w + 1, lane l the columns l, l + 32, ... (coalesced); one row per iteration
#pragma unroll 1
        for (int i = 0; i < DG_BM / DG_WARPS; ++i) {
            const int r = (DG_BM / DG_WARPS) * warp + i, m = m0 + r;
            const bool okm = m < A.M;
            const float* arow = okm ? addr_row(A.a, m, A.ny) + nb : nullptr;
            const float* yrow = (okm && ACT != ACT_NONE) ? addr_row(A.y, m, A.ny) + nb : nullptr;
            float* drow = (okm && store_dy) ? const_cast<float*>(addr_row(A.dy, m, A.ny)) + nb : nullptr;
            float av[DG_NP / 32], yv[DG_NP / 32];
#pragma unroll
            for (int j = 0; j < DG_NP / 32; ++j) {
                const int n = lane + 32 * j;
                const bool ok = okm && n < np;
                av[j] = ok ? arow[n] : 0.f;
                yv[j] = (ok && ACT != ACT_NONE) ? yrow[n] : 0.f;
            }
#pragma unroll
            for (int j = 0; j < DG_NP / 32; ++j) {
                const int n = lane + 32 * j;
                if (n < np) {
                    const float v = av[j] * act_deriv_t<ACT>(yv[j], A.act_scale, A.act_add);
                    if (drow) drow[n] = v;
                    As[n * DG_LDA + r] = v;
                }
            }
        }
        __syncthreads();
        if (has_k) {
#pragma unroll 2
            for (int n = warp; n < np; n += DG_WARPS) {
                const float4 a4 = *reinterpret_cast<const float4*>(As + n * DG_LDA + 4 * rg);
                const float4 w0 = *reinterpret_cast<const float4*>(Ws + n * DG_BK + 8 * fq);
                const float4 w1 = *reinterpret_cast<const float4*>(Ws + n * DG_BK + 8 * fq + 4);
                const float av4[4] = {a4.x, a4.y, a4.z, a4.w};
                const float wv8[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 8; ++j) acc[i][j] += av4[i] * wv8[j];
            }
        }
    }
    if (!has_k) return;
    __syncthreads();
    float* part = dg_smem;                                // [DG_WARPS][DG_BM][DG_BK + 1]
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) part[(warp * DG_BM + 4 * rg + i) * (DG_BK + 1) + 8 * fq + j] = acc[i][j];
    __syncthreads();
    const int mi = tid >> 4, kq = (tid & 15) * 2;
    const int mo = m0 + mi;
    if (mo >= A.M) return;
#pragma unroll 1
    for (int j = 0; j < 2; ++j) {
        const int k = kb + kq + j;
        if (k >= A.K) break;
        float sum = 0.f;
#pragma unroll
        for (int w = 0; w < DG_WARPS; ++w) sum += part[(w * DG_BM + mi) * (DG_BK + 1) + kq + j];
        int si = -1;
        for (int q = 0; q < A.nseg; ++q)
            if (k >= A.seg[q].k0 && k < A.seg[q].k1) si = q;
        if (si < 0) continue;
        const DgradArgs::Seg& S = A.seg[si];
        float* d = const_cast<float*>(addr_row(S.d, mo, A.ny)) + (k - S.k0);
        if (S.mode == SEGM_STORE) *d = sum; else *d += sum;
    }
}

template <int ACT>
__global__ void __launch_bounds__(DG_THREADS, 1) dgrad_kernel(const __grid_constant__ DgradDev D) {
    extern __shared__ __align__(16) float dg_smem[];
    dgrad_tile<ACT, true>(D, (int)blockIdx.x, (int)blockIdx.y, dg_smem);
}

// ---------------------------------------------------------------------------------------------
// wgrad: dW[k, n] += sum_m X[m, k] dY[m, n]   (tensor cores, tf32 hi/lo split of both operands, four products,
// fp32 accumulation outside the tensor core -- same arithmetic as the forward layers)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void split_tf32_(float x, uint32_t& hi, uint32_t& lo) {
    hi = __float_as_uint(x) & 0xffffe000u;
    lo = __float_as_uint(x - __uint_as_float(hi)) + 0x1000u;
}
__device__ __forceinline__ void mma_tf32_zero_(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%10,%10,%10};"
                 : "=f"(d[0]), "=f"(d[1]), "=f"(d[2]), "=f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1), "f"(0.f));
}
__device__ __forceinline__ void mma_tf32_(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

constexpr int WG_T = 64, WG_MC = 32, WG_LD = 72;
__global__ void __launch_bounds__(128) wgrad_addr_kernel(const __grid_constant__ WgradArgs A, int m_per_block) {
    __shared__ float Xs[WG_MC * WG_LD], Ys[WG_MC * WG_LD];
    const int n0 = blockIdx.x * WG_T, k0 = blockIdx.y * WG_T;
    const int m_begin = blockIdx.z * m_per_block, m_end = min(A.M, m_begin + m_per_block);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const int wk = (warp >> 1) * 32, wn = (warp & 1) * 32;
    const int cc = threadIdx.x & 63, mm0 = threadIdx.x >> 6;          // loader: column cc, rows mm0, mm0 + 2, ...
    float acc[2][4][4];
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b)
#pragma unroll
            for (int q = 0; q < 4; ++q) acc[a][b][q] = 0.f;
    for (int m0 = m_begin; m0 < m_end; m0 += WG_MC) {
        __syncthreads();
#pragma unroll
        for (int mm = mm0; mm < WG_MC; mm += 2) {
            const int m = m0 + mm;
            float xv = 0.f, yv = 0.f;
            if (m < m_end) {
                if (k0 + cc < A.K) xv = __ldg(addr_row(A.x, m, A.ny) + k0 + cc);
                if (n0 + cc < A.N) yv = __ldg(addr_row(A.dy, m, A.ny) + n0 + cc);
            }
            Xs[mm * WG_LD + cc] = xv;
            Ys[mm * WG_LD + cc] = yv;
        }
        __syncthreads();
#pragma unroll
        for (int ms = 0; ms < WG_MC; ms += 8) {
            uint32_t ah[2][4], al[2][4], bh[4][2], bl[4][2];
#pragma unroll
            for (int a = 0; a < 2; ++a) {
                const float* p = Xs + (ms + t) * WG_LD + wk + a * 16 + g;
                split_tf32_(p[0], ah[a][0], al[a][0]);
                split_tf32_(p[8], ah[a][1], al[a][1]);
                split_tf32_(p[4 * WG_LD], ah[a][2], al[a][2]);
                split_tf32_(p[4 * WG_LD + 8], ah[a][3], al[a][3]);
            }
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                const float* p = Ys + (ms + t) * WG_LD + wn + b * 8 + g;
                split_tf32_(p[0], bh[b][0], bl[b][0]);
                split_tf32_(p[4 * WG_LD], bh[b][1], bl[b][1]);
            }
#pragma unroll
            for (int a = 0; a < 2; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                    float d[4];
                    mma_tf32_zero_(d, al[a], bl[b][0], bl[b][1]);
                    mma_tf32_(d, al[a], bh[b][0], bh[b][1]);
                    mma_tf32_(d, ah[a], bl[b][0], bl[b][1]);
                    mma_tf32_(d, ah[a], bh[b][0], bh[b][1]);
#pragma unroll
                    for (int q = 0; q < 4; ++q) acc[a][b][q] += d[q];
                }
        }
    }
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b)
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int k = k0 + wk + a * 16 + g + (q >> 1) * 8, n = n0 + wn + b * 8 + 2 * t + (q & 1);
                if (k < A.K && n < A.N) atomicAdd(A.dw + (size_t)k * A.ldw + n, acc[a][b][q]);
            }
}

// out[n] += sum_m dY[m, n]: 32 columns x 8 row lanes per block, grid.y splits M
__global__ void __launch_bounds__(256) colsum_kernel(const __grid_constant__ ColsumArgs A, int m_per_block) {
    __shared__ float red[8][33];
    pdl_trigger();
    pdl_wait();
    const int col = threadIdx.x & 31, rl = threadIdx.x >> 5, n = blockIdx.x * 32 + col;
    const int m_begin = blockIdx.y * m_per_block, m_end = min(A.M, m_begin + m_per_block);
    float a = 0.f;
    if (n < A.N)
        for (int m = m_begin + rl; m < m_end; m += 8) a += __ldg(addr_row(A.dy, m, A.ny) + n);
    red[rl][col] = a;
    __syncthreads();
    if (rl == 0 && n < A.N) {
        float s = 0.f;
#pragma unroll
        for (int q = 0; q < 8; ++q) s += red[q][col];
        atomicAdd(A.out + n, s);
    }
}

__global__ void img_reduce_kernel(const float* __restrict__ dy, float* __restrict__ out, int TB, int K, int nh) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < TB * nh; i += gridDim.x * blockDim.x) {
        const int b = i / nh, j = i - b * nh;
        float a = 0.f;
        for (int k = 0; k < K; ++k) a += dy[((size_t)b * K + k) * nh + j];
        out[i] = a;
    }
}

// pieces of the reference's variables <-> per-layer virtual matrices ([KU + 1][NU] row-major, and transposed)
struct BPiece {
    int K, N, src_off, src_ld, NU, ldt;
    long long dst, dst_t;       // offsets of element (0, 0) of the piece in the row-major / transposed copy
};
struct BPieceTab {
    int n;
    BPiece p[160];
};

__global__ void pack_backward_kernel(const __grid_constant__ BPieceTab tab, const float* __restrict__ src, float* __restrict__ bw) {
    const BPiece& pc = tab.p[blockIdx.y];
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < pc.K * pc.N; i += gridDim.x * blockDim.x) {
        const int k = i / pc.N, n = i - k * pc.N;
        const float v = src[(size_t)pc.src_off + (size_t)k * pc.src_ld + n];
        atomicAdd(bw + pc.dst + (long long)k * pc.NU + n, v);            // (a bias row may be the sum of two bias vectors)
        atomicAdd(bw + pc.dst_t + (long long)n * pc.ldt + k, v);
    }
}

__global__ void unpack_backward_kernel(const __grid_constant__ BPieceTab tab, const float* __restrict__ dwv, float* __restrict__ d_params) {
    const BPiece& pc = tab.p[blockIdx.y];
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < pc.K * pc.N; i += gridDim.x * blockDim.x) {
        const int k = i / pc.N, n = i - k * pc.N;
        atomicAdd(d_params + (size_t)pc.src_off + (size_t)k * pc.src_ld + n, dwv[pc.dst + (long long)k * pc.NU + n]);
    }
}

__global__ void small_to_params_kernel(const float* __restrict__ small, float* __restrict__ d_params, POff po) {
    const int i = threadIdx.x;
    if (i < 10) d_params[po.cholesky + i] += small[SM_CHOL + i];
    if (i == 10) d_params[po.d_scale_offset] += small[SM_DSO];
    if (i == 11) d_params[po.p_scale_offset] += small[SM_PSO];
    if (i == 12) d_params[po.output_scale] += small[SM_OUTSCALE];
}


// ---------------------------------------------------------------------------------------------
// optimiser step on the flat buffers (scripts/experiment.py:138-146,155: opt.apply_gradients).  TF 1.x update rules:
//   rmsprop : ms += (1 - decay)(g^2 - ms);  mom = momentum mom + lr g / sqrt(ms + eps);  w -= mom   (slots: ms = 1, mom = 0)
//   adam    : m += (1 - b1)(g - m);  v += (1 - b2)(g^2 - v);  w -= lr_t m / (sqrt(v) + eps),  lr_t = lr sqrt(1 - b2^t) / (1 - b1^t)
//   momentum: acc = momentum acc + g;  w -= lr acc          sgd: w -= lr g
// with g = grad_scale * grad + l2 * w  (targets.l2_reg: weight * sum_v l2_loss(v), gradient weight * v).
// ---------------------------------------------------------------------------------------------
template <int KIND>
__global__ void __launch_bounds__(256) optimizer_kernel(float* __restrict__ w, const float* __restrict__ grad, float* __restrict__ s0,
                                                        float* __restrict__ s1, int64_t n, float lr, float a, float b, float eps,
                                                        float grad_scale, float l2) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float wv = w[i];
        const float g = grad_scale * grad[i] + l2 * wv;
        if (KIND == SQAIR_OPT_RMSPROP) {
            float ms = s0[i], mom = s1[i];
            ms += (g * g - ms) * (1.f - a);
            mom = mom * b + (g * lr) / sqrtf(ms + eps);
            s0[i] = ms; s1[i] = mom;
            w[i] = wv - mom;
        } else if (KIND == SQAIR_OPT_ADAM) {
            float m = s0[i], v = s1[i];
            m += (g - m) * (1.f - a);
            v += (g * g - v) * (1.f - b);
            s0[i] = m; s1[i] = v;
            w[i] = wv - (m * lr) / (sqrtf(v) + eps);
        } else if (KIND == SQAIR_OPT_MOMENTUM) {
            const float acc = s0[i] * a + g;
            s0[i] = acc;
            w[i] = wv - lr * acc;
        } else {
            w[i] = wv - lr * g;
        }
    }
}


// ---------------------------------------------------------------------------------------------
// data path: frames of moving sprites rendered on the device (data/template.py:69-104 `TemplateDataset`: every object
// is pasted at its rounded position with a max blend; uint8 -> float32 / 255 as data/data.py:199 does)
// ---------------------------------------------------------------------------------------------
__global__ void render_sprites_kernel(const uint8_t* __restrict__ atlas, const int32_t* __restrict__ atlas_hw, const int32_t* __restrict__ pos,
                                      const int32_t* __restrict__ sprite, float* __restrict__ frames, int T, int B, int n, int H, int W,
                                      int S, int cell) {
    const int64_t total = (int64_t)T * B * H * W;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int x = (int)(i % W), y = (int)((i / W) % H);
        const int64_t tb = i / ((int64_t)H * W);
        const int b = (int)(tb % B);
        int v = 0;
        for (int j = 0; j < n; ++j) {
            const int sp = sprite[b * n + j];
            if (sp < 0 || sp >= S) continue;
            const int y0 = pos[(tb * n + j) * 2], x0 = pos[(tb * n + j) * 2 + 1];
            const int dy = y - y0, dx = x - x0;
            if (dy < 0 || dx < 0 || dy >= atlas_hw[sp * 2] || dx >= atlas_hw[sp * 2 + 1]) continue;
            v = max(v, (int)atlas[((int64_t)sp * cell + dy) * cell + dx]);
        }
        frames[i] = (float)v / 255.f;
    }
}

static void fill_piece_tab(const Shape& sh, BPieceTab& bt) {
    memset(&bt, 0, sizeof(bt));
    bt.n = (int)sh.pieces.size();
    for (size_t i = 0; i < sh.pieces.size(); ++i) {
        const Piece& p = sh.pieces[i];
        const LayerB& LB = sh.plan.LB[p.layer];
        BPiece& b = bt.p[i];
        b.K = p.K; b.N = p.N; b.src_off = (int)p.src_off; b.src_ld = p.src_ld; b.NU = LB.NU; b.ldt = LB.ldt;
        b.dst = LB.bw_off + (long long)p.urow0 * LB.NU + p.ucol0;
        b.dst_t = sh.plan.bw_total + LB.bwt_off + (long long)p.ucol0 * LB.ldt + p.urow0;
    }
}

// ---------------------------------------------------------------------------------------------
// The reverse program as ONE persistent cluster kernel.  The frame recursion of the backward pass is a dependent chain
// of ~1 400 short operations (row stages, dgrad products, buffer clears); as separate launches every link pays a kernel
// boundary on a cold instruction cache.  Rows (b, k) never interact inside that chain, so -- like the forward kernel -- a
// cluster of C thread blocks owns R rows for the whole reverse program: the host records the operations into a table
// (`BwdOp`, built by the same `BwdDriver`), every cluster interprets the table for ITS rows, and a cluster barrier
// (release / acquire, which also orders the global-memory traffic between the blocks of the cluster) separates
// dependent operations instead of a kernel boundary.  No grid-wide synchronisation exists.
//   * row stage: the cluster's rows are dealt to (block, 128-thread group) pairs, each with its own named barrier;
//   * dgrad: dX = (dY act'(y)) W^T with out-features as the MMA M dimension and the cluster's rows as N (swap-AB, as
//     in the forward): the block stages A = dY act'(y) of its cluster's rows in shared memory, streams ITS column panel
//     of W^T -- stored in mma.m16n8k8 A-fragment order, pre-split into tf32 (hi, lo), by sqair_pack_backward -- through
//     the tensor cores with the forward's 4-product fp32-faithful k-step, and scatters / accumulates its slice of dX
//     straight into the gradient buffers;
//   * clear: every cluster clears the rows it owns.
// ---------------------------------------------------------------------------------------------
enum { OP_STAGE = 0, OP_DGRAD = 1, OP_ZERO = 2 };
constexpr int BP_THREADS = 512, BP_WARPS = BP_THREADS / 32;
constexpr int BP_XLD = 8;                      // staged operand: X[n][8] (rows of the cluster, padded to the MMA N dimension)
constexpr int BP_RED_FLOATS = 12288;           // k-slice partial sums [ksplit][Nc][8]
constexpr int BP_XMAX_N = 768;                 // widest layer output the staged operand holds
constexpr int BP_NTILE_MAX = 3;                // MMA n-tiles (8 operand rows each) per pass

// W^T of one layer as the kernel reads it: panel p (the out-features [p Nc, (p + 1) Nc) of dX) at w_off + p panel_floats,
// [m-tile][k-step][hi | lo][lane][4] (unsplit fp32 panels with the split in the loop and four k-steps in flight measured
// 1.5 % slower: the k-step is not bound by the bytes in flight)
struct TLayer {
    int ksteps, nmt, Nc, ksplit, kper, panel_floats;
    long long w_off;         // floats from the start of the backward parameter buffer
};

struct BwdOp {
    int kind, stage, t, s;
    int act;                 // dgrad: activation whose derivative is applied (ACT_NONE when there is none)
    int zstride;             // zero: floats per row
    int nobar, pad0;         // 1: the next operation does not depend on this one (a clear followed by a clear)
    float* carry[6];         // stage: gZc gTc gPc gZo gTo gPo of the frame (the only BwdCtx fields that change)
    float* zp;               // zero: [rows, zstride]
    TLayer tl;
    DgradArgs d;
};
struct BwdProgramHdr {
    BwdCtx ctx;
    int nops, R, pad[2];
};
constexpr int OP_WORDS = (int)(sizeof(BwdOp) / 4);
static_assert(sizeof(BwdOp) % 8 == 0 && OP_WORDS <= BP_THREADS, "an operation is staged with one thread per word");
static_assert(sizeof(BwdProgramHdr) % 16 == 0, "operations follow the header 16-byte aligned");
static_assert(offsetof(BwdCtx, gPo) - offsetof(BwdCtx, gZc) == 5 * sizeof(float*), "the six carried-gradient pointers are patched as an array");

__device__ __forceinline__ void cluster_sync_rel_acq() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ int cluster_rank() { uint32_t r; asm("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return (int)r; }
__device__ __forceinline__ int cluster_size() { uint32_t r; asm("mov.u32 %0, %%cluster_nctarank;" : "=r"(r)); return (int)r; }
__device__ __forceinline__ int cluster_id() { uint32_t r; asm("mov.u32 %0, %%clusterid.x;" : "=r"(r)); return (int)r; }

struct AFragB {
    float4 hi, lo;
};
__device__ __forceinline__ AFragB ldg_afrag_b(const float4* p) {
    AFragB a;
    asm volatile("ld.global.nc.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(a.hi.x), "=f"(a.hi.y), "=f"(a.hi.z), "=f"(a.hi.w) : "l"(p));
    asm volatile("ld.global.nc.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(a.lo.x), "=f"(a.lo.y), "=f"(a.lo.z), "=f"(a.lo.w) : "l"(p + 32));
    return a;
}
// one k-step, the forward's arithmetic (sqair_device.cuh: mma_kstep): four tf32 partial products summed on the tensor
// core starting from zero, joined to the running sum with round-to-nearest FADDs
__device__ __forceinline__ void mma_kstep_b(float (&acc)[4], const uint32_t (&ah)[4], const uint32_t (&al)[4], float b0f, float b1f) {
    uint32_t b0h, b0l, b1h, b1l;
    split_tf32_(b0f, b0h, b0l); split_tf32_(b1f, b1h, b1l);
    float d[4], e[4];
    mma_tf32_zero_(d, al, b0l, b1l);
    mma_tf32_zero_(e, ah, b0l, b1l);
    mma_tf32_(d, al, b0h, b1h);
    mma_tf32_(e, ah, b0h, b1h);
    acc[0] += d[0] + e[0]; acc[1] += d[1] + e[1]; acc[2] += d[2] + e[2]; acc[3] += d[3] + e[3];
}

#ifdef SQAIR_PROG_PROFILE
__device__ long long g_prog_prof[64][3];      // per operation class: cycles in the body, cycles in the barrier, count (block 0)
#define PROG_TICK(cls) do { if (blockIdx.x == 0 && threadIdx.x == 0) { const long long t_ = clock64(); g_prog_prof[cls][0] += t_ - ptick; g_prog_prof[cls][2] += 1; ptick = t_; } } while (0)
#else
#define PROG_TICK(cls) do { } while (0)
#endif

// dX slice of this block for the cluster's rows [row0, row0 + R).  The operand rows m = (row0 + r) ny + slot of the
// cluster (R ny of them: the slots of a batched operand ride along as extra MMA columns, so W^T is streamed once) are
// processed in passes of up to 8 NTILE columns.
template <int ACT, int NTILE>
__device__ __forceinline__ void program_dgrad(const BwdOp& op, float* smem, int row0, int R, int rows) {
    const DgradArgs& A = op.d;
    const TLayer& TL = op.tl;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t4 = lane & 3;
    const int rank = cluster_rank();
    const int nr = min(R, rows - row0);
    const int ncol = nr * A.ny, m_base = row0 * A.ny;
    const int kcol0 = rank * TL.Nc;
    const bool has_k = A.nseg > 0 && kcol0 < A.K;
    const float4* panel = reinterpret_cast<const float4*>(A.w + (long long)rank * TL.panel_floats);   // A.w = panel 0 of W^T here
    const int npad = TL.ksteps * 8, xtile = npad * BP_XLD;
    float* X = smem;                                   // [NTILE][npad][BP_XLD]
    float* red = smem + BP_NTILE_MAX * BP_XMAX_N * BP_XLD;   // [ksplit][Nc][8 NTILE]
    // row offsets of the pass's operand rows, computed once (the (m / ny, m % ny) split costs two integer divisions):
    // [0] a, [1] y, [2] dy, [3 + segment] destination of the segment
    long long* rowoff = reinterpret_cast<long long*>(red + BP_RED_FLOATS);     // [3 + BW_MAXSEG][8 NTILE_MAX]
    constexpr int RO = 8 * BP_NTILE_MAX;
    constexpr int NCOLP = 8 * NTILE, PF = 2;           // PF k-steps of weights in flight per warp (register budget)
#ifdef SQAIR_PROG_PROFILE
    long long ptick = clock64();
#endif
    for (int c0 = 0; c0 < ncol; c0 += NCOLP) {
        const int ncp = min(NCOLP, ncol - c0);
        if (c0) __syncthreads();
        const int nunits = has_k ? TL.nmt * TL.ksplit : 0;
        if (tid < (3 + A.nseg) * RO) {
            const int which = tid / RO, c = tid - which * RO;
            if (c < ncp) {
                const int m = m_base + c0 + c;
                const Addr& ad = which == 0 ? A.a : which == 1 ? A.y : which == 2 ? A.dy : A.seg[which - 3].d;
                rowoff[tid] = (long long)(m / A.ny) * ad.outer + (long long)(m % A.ny) * ad.inner;
            }
        }
        __syncthreads();
        // operand staging: X[c / 8][n][c % 8] = a[m][n] act'(y[m][n]), also stored to dy by the first block.  A thread owns
        // feature n and walks the columns of one n-tile: up to 16 independent loads, one memory round trip per tile.
        for (int n = tid; n < npad; n += BP_THREADS) {
            const bool okn = n < A.N;
#pragma unroll
            for (int q = 0; q < NTILE; ++q) {
                if (q * 8 >= ncp) break;
                float av[8], yv[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int c = q * 8 + j;
                    av[j] = 0.f; yv[j] = 0.f;
                    if (okn && c < ncp) {
                        av[j] = A.a.p[rowoff[c] + n];
                        if (ACT != ACT_NONE) yv[j] = A.y.p[rowoff[RO + c] + n];
                    }
                }
                float v[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int c = q * 8 + j;
                    v[j] = av[j];
                    if (ACT != ACT_NONE) v[j] *= act_deriv_t<ACT>(yv[j], A.act_scale, A.act_add);
                    if (rank == 0 && A.dy.p != nullptr && okn && c < ncp) A.dy.p[rowoff[2 * RO + c] + n] = v[j];
                }
                float4* xd = reinterpret_cast<float4*>(X + q * xtile + n * BP_XLD);
                xd[0] = make_float4(v[0], v[1], v[2], v[3]);
                xd[1] = make_float4(v[4], v[5], v[6], v[7]);
            }
        }
        __syncthreads();
        PROG_TICK(56);
        for (int u = warp; u < nunits; u += BP_WARPS) {
            const int sl = u / TL.nmt, mt = u - sl * TL.nmt;
            const int k0 = sl * TL.kper, k1 = min(k0 + TL.kper, TL.ksteps);
            const float4* wp = panel + ((size_t)mt * TL.ksteps + k0) * 64 + lane;
            float acc[NTILE][4];
#pragma unroll
            for (int q = 0; q < NTILE; ++q) acc[q][0] = acc[q][1] = acc[q][2] = acc[q][3] = 0.f;
            const float* xp = X + (k0 * 8 + t4) * BP_XLD + g;
            // PF k-steps of weights in flight per warp (named registers: an indexed buffer ends up in local memory)
            AFragB b0 = ldg_afrag_b(wp), b1, b2, b3;
            if (k0 + 1 < k1) b1 = ldg_afrag_b(wp + 64);
            if (PF == 4) {
                if (k0 + 2 < k1) b2 = ldg_afrag_b(wp + 128);
                if (k0 + 3 < k1) b3 = ldg_afrag_b(wp + 192);
            }
            auto step = [&](AFragB& slot, int ks) {
                const uint32_t ah[4] = {__float_as_uint(slot.hi.x), __float_as_uint(slot.hi.y), __float_as_uint(slot.hi.z), __float_as_uint(slot.hi.w)};
                const uint32_t al[4] = {__float_as_uint(slot.lo.x), __float_as_uint(slot.lo.y), __float_as_uint(slot.lo.z), __float_as_uint(slot.lo.w)};
                if (ks + PF < k1) slot = ldg_afrag_b(wp + PF * 64);
                wp += 64;
#pragma unroll
                for (int q = 0; q < NTILE; ++q)
                    if (q * 8 < ncp) mma_kstep_b(acc[q], ah, al, xp[q * xtile], xp[q * xtile + 4 * BP_XLD]);
                xp += 8 * BP_XLD;
            };
            for (int kk = k0; kk < k1; kk += PF) {
                step(b0, kk);
                if (kk + 1 < k1) step(b1, kk + 1);
                if (PF == 4) {
                    if (kk + 2 < k1) step(b2, kk + 2);
                    if (kk + 3 < k1) step(b3, kk + 3);
                }
            }
            // C fragment: acc[0] = (col g, row 2 t4), acc[1] = (g, 2 t4 + 1), acc[2] = (g + 8, 2 t4), acc[3] = (g + 8, 2 t4 + 1)
#pragma unroll
            for (int q = 0; q < NTILE; ++q) {
                float* rp = red + ((size_t)sl * TL.Nc + mt * 16 + g) * NCOLP + q * 8 + 2 * t4;
                rp[0] = acc[q][0]; rp[1] = acc[q][1]; rp[8 * NCOLP] = acc[q][2]; rp[8 * NCOLP + 1] = acc[q][3];
            }
        }
        PROG_TICK(57);
        __syncthreads();
        PROG_TICK(58);
        if (has_k) {
            for (int o = tid; o < TL.Nc * NCOLP; o += BP_THREADS) {
                const int col = o / NCOLP, c = o - col * NCOLP, k = kcol0 + col;     // NCOLP is a compile-time constant
                if (c >= ncp || k >= A.K) continue;
                int si = -1;
                for (int q = 0; q < A.nseg; ++q)
                    if (k >= A.seg[q].k0 && k < A.seg[q].k1) si = q;
                if (si < 0) continue;
                float sum = 0.f;
                for (int s2 = 0; s2 < TL.ksplit; ++s2) sum += red[((size_t)s2 * TL.Nc + col) * NCOLP + c];
                const DgradArgs::Seg& S = A.seg[si];
                float* d = S.d.p + rowoff[(3 + si) * RO + c] + (k - S.k0);
                if (S.mode == SEGM_STORE) *d = sum;
                else atomicAdd(d, sum);                // one writer per element: a fire-and-forget RED instead of a load round trip
            }
        }
        PROG_TICK(59);
    }
}

// rows of the cluster -> (block, thread group): the smallest number of groups per block that covers R rows in one round
template <int STAGE>
__device__ __forceinline__ void program_stage(const BwdCtx& c, float* smem, float* red, int scratch_floats, int t, int s, int row0,
                                              int R) {
    const int C = cluster_size();
    const int nsub = R <= C ? 1 : R <= 2 * C ? 2 : 4, nts = BP_THREADS / nsub;
    const int sub = threadIdx.x / nts;
    DevEx ex;
    ex.tid = threadIdx.x - sub * nts; ex.nt = nts; ex.scratch = smem + (size_t)sub * scratch_floats; ex.red = red + sub * 128;
    ex.bar_id = 1 + sub;
    for (int i = cluster_rank() + sub * C; i < R && row0 + i < c.rows; i += nsub * C) {
        bw_stage<STAGE>(c, ex, t, s, row0 + i);
        ex.sync();                                   // the group's scratch is reused by its next row
    }
}

__global__ void __launch_bounds__(BP_THREADS, 1) bwd_program_kernel(const BwdProgramHdr* __restrict__ hdr, int scratch_floats) {
    extern __shared__ __align__(16) float bp_smem[];
    __shared__ __align__(16) BwdCtx sc;
    __shared__ __align__(16) BwdOp sop;
    __shared__ float red[4 * 128];
    const int tid = threadIdx.x;
    {
        const int* src = reinterpret_cast<const int*>(&hdr->ctx);
        int* dst = reinterpret_cast<int*>(&sc);
        for (int i = tid; i < (int)(sizeof(BwdCtx) / 4); i += BP_THREADS) dst[i] = __ldg(src + i);
    }
    const int nops = hdr->nops, R = hdr->R;
    const int row0 = cluster_id() * R;
    const int* ops = reinterpret_cast<const int*>(hdr + 1);
    int next_word = (tid < OP_WORDS && nops > 0) ? __ldg(ops + tid) : 0;
    for (int i = 0; i < nops; ++i) {
        if (tid < OP_WORDS) reinterpret_cast<int*>(&sop)[tid] = next_word;
        __syncthreads();
        if (tid < OP_WORDS && i + 1 < nops) next_word = __ldg(ops + (size_t)(i + 1) * OP_WORDS + tid);
        const int kind = sop.kind;
#ifdef SQAIR_PROG_PROFILE
        const long long pt0 = clock64();
        const int pclass = kind == OP_STAGE ? sop.stage : kind == OP_DGRAD ? 32 + (sop.d.ny > 1 ? 8 : 0) + sop.act : kind == OP_ZERO ? 48 : 49;
#endif
        if (kind == OP_STAGE) {
            if (tid < 6) (&sc.gZc)[tid] = sop.carry[tid];
            __syncthreads();
            const int t = sop.t, s = sop.s;
            switch (sop.stage) {
#define SQ_STAGE_CASE(ID) case ID: program_stage<ID>(sc, bp_smem, red, scratch_floats, t, s, row0, R); break;
                SQ_STAGE_CASE(BS_CANVAS) SQ_STAGE_CASE(BS_COMPACT) SQ_STAGE_CASE(BS_DISC_POST) SQ_STAGE_CASE(BS_DISC_A)
                SQ_STAGE_CASE(BS_DISC_B) SQ_STAGE_CASE(BS_DISC_C) SQ_STAGE_CASE(BS_LAT_PRE) SQ_STAGE_CASE(BS_PROP_A)
                SQ_STAGE_CASE(BS_PROP_B) SQ_STAGE_CASE(BS_PROP_C) SQ_STAGE_CASE(BS_PROP_D) SQ_STAGE_CASE(BS_PROP_E)
                SQ_STAGE_CASE(BS_PROP_F) SQ_STAGE_CASE(BS_STN1) SQ_STAGE_CASE(BS_PRIOR_PRE) SQ_STAGE_CASE(BS_PGRU_A)
                SQ_STAGE_CASE(BS_PGRU_B) SQ_STAGE_CASE(BS_FINAL_STATES)
#undef SQ_STAGE_CASE
                default: break;
            }
        } else if (kind == OP_DGRAD) {
            if (sop.d.ny * R <= 8) {                // at most 8 operand rows per cluster: one MMA n-tile
                switch (sop.act) {
                    case ACT_ELU: program_dgrad<ACT_ELU, 1>(sop, bp_smem, row0, R, sc.rows); break;
                    case ACT_TANH: program_dgrad<ACT_TANH, 1>(sop, bp_smem, row0, R, sc.rows); break;
                    case ACT_SIGMOID: program_dgrad<ACT_SIGMOID, 1>(sop, bp_smem, row0, R, sc.rows); break;
                    default: program_dgrad<ACT_NONE, 1>(sop, bp_smem, row0, R, sc.rows); break;
                }
            } else {                                // operand batched over slots
                switch (sop.act) {
                    case ACT_ELU: program_dgrad<ACT_ELU, BP_NTILE_MAX>(sop, bp_smem, row0, R, sc.rows); break;
                    case ACT_TANH: program_dgrad<ACT_TANH, BP_NTILE_MAX>(sop, bp_smem, row0, R, sc.rows); break;
                    case ACT_SIGMOID: program_dgrad<ACT_SIGMOID, BP_NTILE_MAX>(sop, bp_smem, row0, R, sc.rows); break;
                    default: program_dgrad<ACT_NONE, BP_NTILE_MAX>(sop, bp_smem, row0, R, sc.rows); break;
                }
            }
        } else {
            const int nr = min(R, sc.rows - row0);
            float* p = sop.zp + (size_t)row0 * sop.zstride;
            const long long n = (long long)nr * sop.zstride;
            const int ctid = cluster_rank() * BP_THREADS + tid, cn = cluster_size() * BP_THREADS;
            for (long long j = ctid; j < n; j += cn) p[j] = 0.f;
        }
#ifdef SQAIR_PROG_PROFILE
        __syncthreads();
        const long long pt1 = clock64();
#endif
        if (sop.nobar) __syncthreads(); else cluster_sync_rel_acq();
#ifdef SQAIR_PROG_PROFILE
        if (blockIdx.x == 0 && tid == 0) {
            g_prog_prof[pclass][0] += pt1 - pt0; g_prog_prof[pclass][1] += clock64() - pt1; g_prog_prof[pclass][2] += 1;
        }
#endif
    }
}

// W^T panels in fragment order from the row-major virtual matrix [KU + 1][NU] of the backward parameter buffer
struct TPackTab {
    int n;
    struct E { long long src, dst; int KU, NU, Nc, ksteps, panel_floats; } e[L_COUNT];
};
__global__ void pack_tfrag_kernel(const __grid_constant__ TPackTab tab, float* __restrict__ bw) {
    const TPackTab::E& E = tab.e[blockIdx.y];
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < E.KU * E.NU; i += gridDim.x * blockDim.x) {
        const int k = i / E.NU, n = i - k * E.NU;
        const int panel = k / E.Nc, cc = k - panel * E.Nc;
        const int fo = frag_off(E.ksteps, cc >> 4, n >> 3, cc & 15, n & 7);
        float hi, lo;
        split_weight(bw[E.src + i], hi, lo);
        float* d = bw + E.dst + (long long)panel * E.panel_floats + (long long)(fo >> 7) * 256 + (fo & 127);
        d[0] = hi; d[128] = lo;
    }
}

// layers that have a dgrad product in the reverse program get W^T panels for a cluster of C blocks
static void build_tlayers(const Plan& plan, int C, TLayer* tl, int64_t* total, int ncolp = 8 * BP_NTILE_MAX) {
    int64_t cur = (plan.bw_total + plan.bwt_total + 31) / 32 * 32;
    for (int l = 0; l < L_COUNT; ++l) {
        TLayer& T = tl[l];
        memset(&T, 0, sizeof(T));
        if (plan.L[l].nhead == 0) continue;
        const LayerB& LB = plan.LB[l];
        T.ksteps = (LB.NU + 7) / 8;
        T.Nc = ((LB.KU + C - 1) / C + 15) / 16 * 16;
        T.nmt = T.Nc / 16;
        T.panel_floats = T.nmt * T.ksteps * 256;
        T.w_off = cur;
        cur += (int64_t)C * T.panel_floats;
        int best = 1;
        double best_cost = 1e30;
        for (int ks = 1; ks <= BP_WARPS; ++ks) {
            const int kper = (T.ksteps + ks - 1) / ks;
            if (ks > 1 && (kper < 2 || (ks - 1) * kper >= T.ksteps)) continue;
            if ((int64_t)ks * T.Nc * ncolp > BP_RED_FLOATS) continue;
            const int rounds = (T.nmt * ks + BP_WARPS - 1) / BP_WARPS;
            const double cost = (double)rounds * (kper + 3.0) + 0.5 * ks;
            if (cost < best_cost) { best_cost = cost; best = ks; }
        }
        T.ksplit = best;
        T.kper = (T.ksteps + best - 1) / best;
    }
    if (total) *total = cur;
}

// device copies of recorded programs, keyed by content (a backward call of the same shape on the same buffers records
// the same bytes: the upload happens once; a CUDA graph of the call keeps pointing at the cached copy)
struct ProgramEntry {
    int device;
    std::vector<char> host;
    void* dev;
};
static std::mutex g_prog_mutex;
static std::vector<ProgramEntry*> g_programs;

static cudaError_t get_program(const std::vector<char>& bytes, cudaStream_t st, const void** out) {
    int device = 0;
    cudaError_t e = cudaGetDevice(&device);
    if (e != cudaSuccess) return e;
    std::lock_guard<std::mutex> lock(g_prog_mutex);
    for (size_t i = g_programs.size(); i-- > 0;) {          // most recently used last: the common case is one comparison
        ProgramEntry* pe = g_programs[i];
        if (pe->device == device && pe->host.size() == bytes.size() && memcmp(pe->host.data(), bytes.data(), bytes.size()) == 0) {
            if (i + 1 != g_programs.size()) { g_programs.erase(g_programs.begin() + i); g_programs.push_back(pe); }
            *out = pe->dev;
            return cudaSuccess;
        }
    }
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(st, &cs) != cudaSuccess) { cudaGetLastError(); cs = cudaStreamCaptureStatusNone; }
    if (cs != cudaStreamCaptureStatusNone) return cudaErrorStreamCaptureUnsupported;   // run the shape eagerly once before capturing it
    if (g_programs.size() >= 64) {              // rare: many distinct (shape, buffer) combinations; drop the oldest once idle
        cudaDeviceSynchronize();
        cudaFree(g_programs.front()->dev);
        delete g_programs.front();
        g_programs.erase(g_programs.begin());
    }
    ProgramEntry* pe = new ProgramEntry{device, bytes, nullptr};
    e = cudaMalloc(&pe->dev, bytes.size());
    if (e == cudaSuccess) e = cudaMemcpyAsync(pe->dev, pe->host.data(), bytes.size(), cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);      // once per program: every later call finds the table resident
    if (e != cudaSuccess) { if (pe->dev) cudaFree(pe->dev); delete pe; return e; }
    g_programs.push_back(pe);
    *out = pe->dev;
    return cudaSuccess;
}

// ---------------------------------------------------------------------------------------------
// backend
// ---------------------------------------------------------------------------------------------
// The weight-gradient phase is ~60 independent GEMMs and column sums, a third of them small (K or N below 32): issued
// round-robin on a few side streams (fork / join with events, which a stream capture follows) they overlap instead of
// queueing behind each other.
constexpr int WG_STREAMS = 8;          // upper bound; SQAIR_WGRAD_STREAMS picks how many are used (measured: 2 -> 11.00, 4 -> 10.71, 6 -> 10.53, 8 -> 10.53 ms per backward)
static cudaError_t side_streams(cudaStream_t (&out)[WG_STREAMS]) {
    static std::mutex mu;
    static cudaStream_t cache[64][WG_STREAMS];
    static bool have[64];
    int device = 0;
    cudaError_t e = cudaGetDevice(&device);
    if (e != cudaSuccess) return e;
    if (device < 0 || device >= 64) return cudaErrorInvalidDevice;
    std::lock_guard<std::mutex> lock(mu);
    if (!have[device]) {
        for (int i = 0; i < WG_STREAMS; ++i) {
            e = cudaStreamCreateWithFlags(&cache[device][i], cudaStreamNonBlocking);
            if (e != cudaSuccess) return e;
        }
        have[device] = true;
    }
    for (int i = 0; i < WG_STREAMS; ++i) out[i] = cache[device][i];
    return cudaSuccess;
}

struct CudaBackend {
    cudaStream_t st;
    const Shape* sh;
    int rows;
    int scratch_bytes;
    cudaError_t err = cudaSuccess;
    long launches = 0, tc_launches = 0;
    bool verbose = false;
    // fork / join of the weight-gradient phase
    bool use_side = true, forked = false;
    int nside = 8;
    cudaStream_t side[WG_STREAMS];
    cudaStream_t cur = nullptr;
    int rr = 0, sticky = 0;

    void check() {
        if (err == cudaSuccess) err = cudaGetLastError();
        ++launches;
    }
    void note(cudaError_t e) { if (err == cudaSuccess) err = e; }
    // stream of the next independent operation of the weight-gradient phase (`chain` more operations follow on the same one)
    cudaStream_t pick(int chain = 0) {
        if (!use_side) return st;
        if (!forked) {
            cudaEvent_t ev = nullptr;
            note(side_streams(side));
            note(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
            if (err != cudaSuccess) { use_side = false; return st; }
            note(cudaEventRecord(ev, st));
            for (int i = 0; i < nside; ++i) note(cudaStreamWaitEvent(side[i], ev, 0));
            note(cudaEventDestroy(ev));
            forked = true;
        }
        if (sticky > 0) { --sticky; return cur; }
        cur = side[rr++ % nside];
        sticky = chain;
        return cur;
    }
    void join() {
        if (!forked) return;
        for (int i = 0; i < nside; ++i) {
            cudaEvent_t ev = nullptr;
            note(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
            if (ev) { note(cudaEventRecord(ev, side[i])); note(cudaStreamWaitEvent(st, ev, 0)); note(cudaEventDestroy(ev)); }
        }
        forked = false; sticky = 0;
    }
    template <int STAGE>
    void stage(const BwdCtx& c, int t, int s) {
        const int threads = STAGE == BS_CANVAS ? 256 : 128;
        cudaError_t e = launch_pdl(bwd_stage_kernel<STAGE>, dim3(rows), dim3(threads), (size_t)scratch_bytes, st, c, t, s);
        if (err == cudaSuccess) err = e;
        check();
    }
    void dgrad(const DgradArgs& A) {
        DgradDev D;
        D.a = A;
        const LayerB& LB = sh->plan.LB[A.layer];
        D.wt = A.w - LB.bw_off + sh->plan.bw_total + LB.bwt_off;      // transposed copies follow the row-major matrices
        D.ldt = LB.ldt;
        int gx = A.nseg > 0 ? (A.K + DG_BK - 1) / DG_BK : 1;
        const dim3 grid(gx, (A.M + DG_BM - 1) / DG_BM);
        const size_t smem = DG_SMEM_FLOATS * sizeof(float);
        cudaError_t e;
        switch (A.y.p ? A.act : ACT_NONE) {
            case ACT_ELU: e = launch_pdl(dgrad_kernel<ACT_ELU>, grid, dim3(DG_THREADS), smem, st, D); break;
            case ACT_TANH: e = launch_pdl(dgrad_kernel<ACT_TANH>, grid, dim3(DG_THREADS), smem, st, D); break;
            case ACT_SIGMOID: e = launch_pdl(dgrad_kernel<ACT_SIGMOID>, grid, dim3(DG_THREADS), smem, st, D); break;
            case ACT_SOFTPLUS: e = launch_pdl(dgrad_kernel<ACT_SOFTPLUS>, grid, dim3(DG_THREADS), smem, st, D); break;
            default: e = launch_pdl(dgrad_kernel<ACT_NONE>, grid, dim3(DG_THREADS), smem, st, D); break;
        }
        if (err == cudaSuccess) err = e;
        check();
    }
    void wgrad(const WgradArgs& A) {
        const cudaStream_t ws_ = pick();
        {   // Blackwell path (sqair_wgrad_tc.cu): TMA + tcgen05.mma + TMEM, whenever TMA can describe both operands
            const sqi::TcOperand ox{A.x.p, A.x.outer, A.x.inner, A.ny}, oy{A.dy.p, A.dy.outer, A.dy.inner, A.ny};
            if (sqi::wgrad_tc_supported(ox, oy, A.M, A.K, A.N)) {
                if (sqi::wgrad_tc(ox, oy, A.dw, A.ldw, A.M, A.K, A.N, ws_) != SQAIR_OK && err == cudaSuccess) err = cudaErrorInvalidValue;
                ++launches; ++tc_launches;
                return;
            }
        }
        if (verbose)
            fprintf(stderr, "  wgrad on the mma.sync kernel: M %d K %d N %d ny %d  x base%%4 %d outer %d inner %d | dy base%%4 %d outer %d inner %d\n", A.M,
                    A.K, A.N, A.ny, (int)((reinterpret_cast<uintptr_t>(A.x.p) / 4) % 4), A.x.outer, A.x.inner,
                    (int)((reinterpret_cast<uintptr_t>(A.dy.p) / 4) % 4), A.dy.outer, A.dy.inner);
        const int tiles = ((A.N + WG_T - 1) / WG_T) * ((A.K + WG_T - 1) / WG_T);
        int msplit = (4 * 148 + tiles - 1) / tiles;
        // layers with few output tiles are bound by the latency of a block's row loop, not by the reductions: one 32-row
        // step per block (measured: 4 steps per block cost 50 us per launch whatever the layer)
        const int steps_min = tiles <= 8 ? 1 : 4;
        const int max_split = (A.M + steps_min * WG_MC - 1) / (steps_min * WG_MC);
        if (msplit > max_split) msplit = max_split;
        if (msplit < 1) msplit = 1;
        const int m_per_block = ((A.M + msplit - 1) / msplit + WG_MC - 1) / WG_MC * WG_MC;
        msplit = (A.M + m_per_block - 1) / m_per_block;
        wgrad_addr_kernel<<<dim3((A.N + WG_T - 1) / WG_T, (A.K + WG_T - 1) / WG_T, msplit), 128, 0, ws_>>>(A, m_per_block);
        check();
    }
    void colsum(const ColsumArgs& A) {
        const int gx = (A.N + 31) / 32;
        int msplit = (2 * 148 + gx - 1) / gx;
        const int max_split = (A.M + 63) / 64;
        if (msplit > max_split) msplit = max_split;
        if (msplit < 1) msplit = 1;
        const int m_per_block = (A.M + msplit - 1) / msplit;
        cudaError_t e = launch_pdl(colsum_kernel, dim3(gx, (A.M + m_per_block - 1) / m_per_block), dim3(256), 0, pick(), A, m_per_block);
        if (err == cudaSuccess) err = e;
        check();
    }
    void zero(float* p, int64_t n) {
        if (n <= 0) return;
        cudaError_t e = cudaMemsetAsync(p, 0, (size_t)n * sizeof(float), st);
        if (err == cudaSuccess) err = e;
        ++launches;
    }
    void img_reduce(const float* dy, float* out, int TB, int K, int nh) {
        // its consumers follow on the same stream: the image encoder's GEMM and column sum, or the GEMM of a slot-shared operand
        img_reduce_kernel<<<(TB * nh + 255) / 256, 256, 0, pick(2)>>>(dy, out, TB, K, nh);
        check();
    }
    void unpack(const float* dwv, float* d_params) {
        BPieceTab bt;
        fill_piece_tab(*sh, bt);
        join();                                      // every virtual-matrix gradient is complete
        unpack_backward_kernel<<<dim3(32, bt.n), 256, 0, st>>>(bt, dwv, d_params);
        check();
    }
    void small_to_params(const float* small, float* d_params, const POff& po) {
        join();
        small_to_params_kernel<<<1, 32, 0, st>>>(small, d_params, po);
        check();
    }
};

// Records the frame recursion (row stages, dgrad products, clears) into a BwdOp table and runs it as one persistent
// cluster kernel as soon as an operation arrives that is not part of it (the weight-gradient GEMMs at the end, which
// stay separate launches on the same stream).
struct ProgramBackend {
    CudaBackend& cb;
    const BwdLayout& BL;
    float* ws;
    const float* bw;             // backward parameter buffer (the W^T panels follow the matrices)
    int R, C, rows, scratch_floats;
    TLayer tl[L_COUNT];
    BwdProgramHdr hdr;
    bool have_ctx = false, bad = false;
    std::vector<BwdOp> ops;
    long program_launches = 0;

    ProgramBackend(CudaBackend& c, const BwdLayout& bl, float* workspace, const float* bwp, int r, int cl, int nrows, int scratch)
        : cb(c), BL(bl), ws(workspace), bw(bwp), R(r), C(cl), rows(nrows), scratch_floats(scratch) {
        memset((void*)&hdr, 0, sizeof(hdr));
        build_tlayers(cb.sh->plan, C, tl, nullptr);
    }
    BwdOp& push(int kind) {
        BwdOp op;
        memset((void*)&op, 0, sizeof(op));
        op.kind = kind;
        ops.push_back(op);
        return ops.back();
    }
    bool recording() const { return have_ctx; }      // the program starts with the first row stage
    template <int STAGE>
    void stage(const BwdCtx& c, int t, int s) {
        if (!have_ctx) { hdr.ctx = c; have_ctx = true; }
        BwdOp& op = push(OP_STAGE);
        op.stage = STAGE; op.t = t; op.s = s;
        float* const cy[6] = {c.gZc, c.gTc, c.gPc, c.gZo, c.gTo, c.gPo};
        for (int i = 0; i < 6; ++i) op.carry[i] = cy[i];
    }
    void dgrad(const DgradArgs& A) {
        if (!recording()) { bad = true; return; }
        if (A.N > BP_XMAX_N) { bad = true; return; }
        const int act = A.y.p ? A.act : ACT_NONE;
        if (act == ACT_SOFTPLUS) { bad = true; return; }       // not instantiated (no such product in the reverse program)
        BwdOp& op = push(OP_DGRAD);
        op.d = A;
        op.tl = tl[A.layer];
        op.d.w = bw + op.tl.w_off;                 // panel 0 of W^T (the row-major matrix is not used by this kernel)
        op.act = A.y.p ? A.act : ACT_NONE;
    }
    void zero_rows(float* p, int stride) {
        if (!ops.empty() && ops.back().kind == OP_ZERO) ops.back().nobar = 1;
        BwdOp& op = push(OP_ZERO);
        op.zp = p; op.zstride = stride;
    }
    void zero(float* p, int64_t n) {
        if (n <= 0) return;
        if (!recording()) { cb.zero(p, n); return; }            // clears ahead of the program: plain memsets, in stream order
        if (p == ws + BL.frame_begin && n == BL.frame_end - BL.frame_begin) {
            for (int i = 0; i < BL.nframe; ++i) zero_rows(ws + BL.frame_off[i], BL.frame_stride[i]);
        } else if (n % rows == 0) {
            zero_rows(p, (int)(n / rows));                       // a carried state-gradient set: [rows, n * width]
        } else {
            bad = true;
        }
    }
    void flush() {
        if (ops.empty() || cb.err != cudaSuccess) { ops.clear(); return; }
        if (bad) { cb.err = cudaErrorInvalidValue; ops.clear(); return; }
        // the carry fields of the header are patched per operation; keep the recorded bytes independent of the last frame
        hdr.ctx.gZc = hdr.ctx.gTc = hdr.ctx.gPc = hdr.ctx.gZo = hdr.ctx.gTo = hdr.ctx.gPo = nullptr;
        hdr.nops = (int)ops.size();
        hdr.R = R;
        std::vector<char> bytes(sizeof(hdr) + ops.size() * sizeof(BwdOp));
        memcpy(bytes.data(), &hdr, sizeof(hdr));
        memcpy(bytes.data() + sizeof(hdr), ops.data(), ops.size() * sizeof(BwdOp));
        ops.clear();
        have_ctx = false;
        const void* dev = nullptr;
        cudaError_t e = get_program(bytes, cb.st, &dev);
        const int smem_floats = std::max(BP_NTILE_MAX * BP_XMAX_N * BP_XLD + BP_RED_FLOATS + 2 * (3 + BW_MAXSEG) * 8 * BP_NTILE_MAX, 4 * scratch_floats);
        const int smem_bytes = smem_floats * (int)sizeof(float);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(bwd_program_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
        if (e == cudaSuccess) {
            cudaLaunchConfig_t cfg;
            memset(&cfg, 0, sizeof(cfg));
            const int ncl = (rows + R - 1) / R;
            cfg.gridDim = dim3(ncl * C); cfg.blockDim = dim3(BP_THREADS); cfg.dynamicSmemBytes = (size_t)smem_bytes; cfg.stream = cb.st;
            cudaLaunchAttribute attr[1];
            attr[0].id = cudaLaunchAttributeClusterDimension;
            attr[0].val.clusterDim.x = C; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
            cfg.attrs = attr; cfg.numAttrs = C > 1 ? 1 : 0;
            e = cudaLaunchKernelEx(&cfg, bwd_program_kernel, (const BwdProgramHdr*)dev, scratch_floats);
        }
        if (cb.err == cudaSuccess) cb.err = e;
        ++cb.launches; ++program_launches;
#ifdef SQAIR_PROG_PROFILE
        if (sqi::env_int("SQAIR_PROG_PRINT")) {
            long long h[64][3];
            cudaStreamSynchronize(cb.st);
            cudaMemcpyFromSymbol(h, g_prog_prof, sizeof(h));
            long long tb = 0, tw = 0;
            for (int i = 0; i < 64; ++i) { tb += h[i][0]; tw += h[i][1]; }
            fprintf(stderr, "reverse program, block 0 (classes: 0-17 stage id, 32+act dgrad over rows, 40+act dgrad over rows x slots, 48 clear)\n");
            for (int i = 0; i < 64; ++i)
                if (h[i][2]) fprintf(stderr, "  class %2d: n %6lld  body %8.0f cyc/op  barrier %8.0f cyc/op  share %.3f\n", i, h[i][2], (double)h[i][0] / h[i][2],
                                     (double)h[i][1] / h[i][2], (double)(h[i][0] + h[i][1]) / (double)(tb + tw));
            fprintf(stderr, "  total body %.3f Mcyc, barrier %.3f Mcyc\n", tb * 1e-6, tw * 1e-6);
            memset(h, 0, sizeof(h));
            cudaMemcpyToSymbol(g_prog_prof, h, sizeof(h));
        }
#endif
    }
    void wgrad(const WgradArgs& A) { flush(); cb.wgrad(A); }
    void colsum(const ColsumArgs& A) { flush(); cb.colsum(A); }
    void img_reduce(const float* dy, float* out, int TB, int K, int nh) { flush(); cb.img_reduce(dy, out, TB, K, nh); }
    void unpack(const float* dwv, float* d_params) { flush(); cb.unpack(dwv, d_params); }
    void small_to_params(const float* small, float* d_params, const POff& po) { flush(); cb.small_to_params(small, d_params, po); }
};

// Launch shape of the reverse-program kernel: rows per cluster and cluster size.  Default: the forward's (R, C) -- except
// that the forward's shared-memory budget sometimes forces clusters of fewer than 4 blocks (BASELINE configs[3]: R=4, C=3),
// while this kernel's shared memory does not depend on the rows: then C = 4 with as many rows per cluster as one wave of
// resident clusters needs (configs[3]: 17.1 -> 15.6 ms per backward).  SQAIR_BWD_ROWS / SQAIR_BWD_CLUSTER override (tuning).
static int program_max_clusters(int C, int smem_bytes) {
    static std::mutex mu;
    static int cached[64][9];
    int device = 0;
    if (cudaGetDevice(&device) != cudaSuccess || device < 0 || device >= 64 || C < 1 || C > 8) { cudaGetLastError(); return 32; }
    std::lock_guard<std::mutex> lock(mu);
    if (cached[device][C] == 0) {
        int n = 0;
        cudaLaunchConfig_t cfg;
        memset(&cfg, 0, sizeof(cfg));
        cfg.blockDim = dim3(BP_THREADS); cfg.gridDim = dim3(C * 64); cfg.dynamicSmemBytes = (size_t)smem_bytes;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = C; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        cudaError_t e = cudaFuncSetAttribute(bwd_program_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
        if (e == cudaSuccess) e = cudaOccupancyMaxActiveClusters(&n, bwd_program_kernel, &cfg);
        if (e != cudaSuccess || n <= 0) { cudaGetLastError(); n = 128 / C; }      // what a 148-SM B200 answers for C = 4: 33
        cached[device][C] = n;
    }
    return cached[device][C];
}
static int program_smem_bytes(const sqair_cfg& cfg) {
    return std::max(BP_NTILE_MAX * BP_XMAX_N * BP_XLD + BP_RED_FLOATS + 2 * (3 + BW_MAXSEG) * 8 * BP_NTILE_MAX, 4 * bw_stage_scratch_floats(cfg)) *
           (int)sizeof(float);
}
static void program_shape(const sqair_cfg& cfg, const Shape& sh, int& R, int& C) {
    R = sh.R; C = sh.C;
    const int rows = cfg.B * cfg.K;
    if (C < 4 && rows > 4 * R) {
        C = 4;
        const int maxcl = program_max_clusters(4, program_smem_bytes(cfg));
        R = std::max(R, (rows + maxcl - 1) / maxcl);
    }
    const int er = sqi::env_int("SQAIR_BWD_ROWS"), ec = sqi::env_int("SQAIR_BWD_CLUSTER");
    if (er >= 1 && er <= 24) R = er;
    if (ec >= 1 && ec <= 8) C = ec;
}

static std::string prepare(const sqair_cfg* cfg, Shape& sh, std::vector<ParamEntry>& tab) {
    std::string e = validate_cfg(*cfg);
    if (!e.empty()) return e;
    tab = param_table(*cfg);
    return sqi::choose_shape(*cfg, tab, sh);
}

extern "C" {

int sqair_query_train_sizes(const sqair_cfg* cfg, sqair_train_sizes* out) {
    if (!cfg || !out) return fail(SQAIR_EINVAL, "null argument");
    Shape sh;
    std::vector<ParamEntry> tab;
    std::string e = prepare(cfg, sh, tab);
    if (!e.empty()) return fail(SQAIR_EINVAL, e);
    const StashLayout SL = build_stash(*cfg);
    if (SL.total < 0) return fail(SQAIR_EUNSUPPORTED, "training stash exceeds 2^31 floats");
    const BwdLayout BL = build_bwd_layout(*cfg, sh.plan);
    out->stash_floats = SL.total;
    out->workspace_floats = BL.total;
    {
        TLayer tl[L_COUNT];
        int64_t total = 0;
        int pr, pc;
        program_shape(*cfg, sh, pr, pc);
        build_tlayers(sh.plan, pc, tl, &total);
        out->backward_param_floats = total;             // matrices, transposed copies, W^T fragment panels
    }
    return SQAIR_OK;
}

int sqair_pack_backward(const sqair_cfg* cfg, const float* params, float* bw_params, void* stream) {
    if (!cfg || !params || !bw_params) return fail(SQAIR_EINVAL, "null argument");
    Shape sh;
    std::vector<ParamEntry> tab;
    std::string e = prepare(cfg, sh, tab);
    if (!e.empty()) return fail(SQAIR_EINVAL, e);
    if (sh.pieces.size() > 160) return fail(SQAIR_EUNSUPPORTED, "too many variables");
    cudaStream_t st = (cudaStream_t)stream;
    BPieceTab bt;
    fill_piece_tab(sh, bt);
    TPackTab tt;
    memset(&tt, 0, sizeof(tt));
    int64_t bw_floats = 0;
    {
        TLayer tl[L_COUNT];
        int pr, pc;
        program_shape(*cfg, sh, pr, pc);
        build_tlayers(sh.plan, pc, tl, &bw_floats);
        for (int l = 0; l < L_COUNT; ++l) {
            if (sh.plan.L[l].nhead == 0) continue;
            const LayerB& LB = sh.plan.LB[l];
            TPackTab::E& E = tt.e[tt.n++];
            E.src = LB.bw_off; E.dst = tl[l].w_off; E.KU = LB.KU; E.NU = LB.NU; E.Nc = tl[l].Nc; E.ksteps = tl[l].ksteps;
            E.panel_floats = tl[l].panel_floats;
        }
    }
    CUDA_TRY(cudaMemsetAsync(bw_params, 0, (size_t)bw_floats * sizeof(float), st));
    pack_backward_kernel<<<dim3(32, bt.n), 256, 0, st>>>(bt, params, bw_params);
    CUDA_TRY(cudaGetLastError());
    pack_tfrag_kernel<<<dim3(32, tt.n), 256, 0, st>>>(tt, bw_params);       // from the matrices the kernel above assembled
    CUDA_TRY(cudaGetLastError());
    return SQAIR_OK;
}

int sqair_backward(const sqair_cfg* cfg, const float* params, const float* bw_params, const float* obs, const float* eps_where,
                   const float* eps_what, const float* stash, const float* d_log_weights, const float* d_discrete_log_prob,
                   float* workspace, float* d_params, int32_t* n_launches, void* stream) {
    if (!cfg || !params || !bw_params || !obs || !eps_where || !eps_what || !stash || !d_log_weights || !workspace || !d_params)
        return fail(SQAIR_EINVAL, "null argument");
    Shape sh;
    std::vector<ParamEntry> tab;
    std::string e = prepare(cfg, sh, tab);
    if (!e.empty()) return fail(SQAIR_EINVAL, e);
    const BwdLayout BL = build_bwd_layout(*cfg, sh.plan);
    BwdInputs in;
    in.params = params; in.bw = bw_params; in.obs = obs; in.eps_where = eps_where; in.eps_what = eps_what; in.stash = stash;
    in.d_log_w = d_log_weights; in.d_disc_lp = d_discrete_log_prob; in.ws = workspace; in.d_params = d_params; in.vimco = 1;
    const int dg_smem_bytes = DG_SMEM_FLOATS * (int)sizeof(float);
    CUDA_TRY(cudaFuncSetAttribute(dgrad_kernel<ACT_NONE>, cudaFuncAttributeMaxDynamicSharedMemorySize, dg_smem_bytes));
    CUDA_TRY(cudaFuncSetAttribute(dgrad_kernel<ACT_ELU>, cudaFuncAttributeMaxDynamicSharedMemorySize, dg_smem_bytes));
    CUDA_TRY(cudaFuncSetAttribute(dgrad_kernel<ACT_TANH>, cudaFuncAttributeMaxDynamicSharedMemorySize, dg_smem_bytes));
    CUDA_TRY(cudaFuncSetAttribute(dgrad_kernel<ACT_SIGMOID>, cudaFuncAttributeMaxDynamicSharedMemorySize, dg_smem_bytes));
    CUDA_TRY(cudaFuncSetAttribute(dgrad_kernel<ACT_SOFTPLUS>, cudaFuncAttributeMaxDynamicSharedMemorySize, dg_smem_bytes));
    CudaBackend be;
    be.st = (cudaStream_t)stream; be.sh = &sh; be.rows = cfg->B * cfg->K;
    be.verbose = sqi::env_int("SQAIR_VERBOSE") > 1;
    be.use_side = !sqi::env_int("SQAIR_WGRAD_ONE_STREAM");
    if (sqi::env_int("SQAIR_WGRAD_STREAMS") >= 1 && sqi::env_int("SQAIR_WGRAD_STREAMS") <= WG_STREAMS) be.nside = sqi::env_int("SQAIR_WGRAD_STREAMS");
    be.scratch_bytes = bw_stage_scratch_floats(*cfg) * (int)sizeof(float);
    if (be.scratch_bytes > 48 * 1024) return fail(SQAIR_EUNSUPPORTED, "glimpses do not fit the shared memory of the canvas stage");
    static_assert(sizeof(BwdCtx) <= 4000, "BwdCtx must fit the kernel parameter space");
    const int64_t n_params = tab.back().offset + tab.back().count;
    if (sqi::env_int("SQAIR_BWD_LAUNCHES")) {           // one launch per operation (the round-2 path; kept for A/B timing and tests)
        BwdDriver<CudaBackend> drv(be, *cfg, sh.plan, sh.plan.poc, BL, in);
        drv.param_count_ = n_params;
        drv.run(d_params);
    } else {
        const int scratch_floats = bw_stage_scratch_floats(*cfg);
        if ((int64_t)std::max(BP_NTILE_MAX * BP_XMAX_N * BP_XLD + BP_RED_FLOATS + 2 * (3 + BW_MAXSEG) * 8 * BP_NTILE_MAX, 4 * scratch_floats) * 4 > 200 * 1024)
            return fail(SQAIR_EUNSUPPORTED, "glimpses do not fit the shared memory of the reverse-program kernel");
        int pr, pc;
        program_shape(*cfg, sh, pr, pc);
        ProgramBackend pb(be, BL, workspace, bw_params, pr, pc, cfg->B * cfg->K, scratch_floats);
        BwdDriver<ProgramBackend> drv(pb, *cfg, sh.plan, sh.plan.poc, BL, in);
        drv.param_count_ = n_params;
        drv.run(d_params);
        pb.flush();
    }
    be.join();
    if (be.err != cudaSuccess) return sqi::cuda_fail(be.err, "sqair_backward");
    if (n_launches) *n_launches = (int32_t)be.launches;
    if (sqi::env_int("SQAIR_VERBOSE")) fprintf(stderr, "sqair_backward: %ld launches, %ld of them tcgen05 weight-gradient GEMMs\n", be.launches, be.tc_launches);
    return SQAIR_OK;
}

int sqair_optimizer_update(int32_t kind, float* params, const float* grad, float* slot0, float* slot1, int64_t n, float lr,
                           float hyper_a, float hyper_b, float epsilon, float grad_scale, float l2_weight, void* stream) {
    if (!params || !grad || n < 0) return fail(SQAIR_EINVAL, "null argument");
    if ((kind == SQAIR_OPT_RMSPROP || kind == SQAIR_OPT_ADAM) && (!slot0 || !slot1)) return fail(SQAIR_EINVAL, "optimiser slots missing");
    if (kind == SQAIR_OPT_MOMENTUM && !slot0) return fail(SQAIR_EINVAL, "optimiser slots missing");
    if (n == 0) return SQAIR_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const int blocks = (int)std::min<int64_t>((n + 255) / 256, 148 * 8);
    switch (kind) {
        case SQAIR_OPT_RMSPROP: optimizer_kernel<SQAIR_OPT_RMSPROP><<<blocks, 256, 0, st>>>(params, grad, slot0, slot1, n, lr, hyper_a, hyper_b, epsilon, grad_scale, l2_weight); break;
        case SQAIR_OPT_ADAM: optimizer_kernel<SQAIR_OPT_ADAM><<<blocks, 256, 0, st>>>(params, grad, slot0, slot1, n, lr, hyper_a, hyper_b, epsilon, grad_scale, l2_weight); break;
        case SQAIR_OPT_MOMENTUM: optimizer_kernel<SQAIR_OPT_MOMENTUM><<<blocks, 256, 0, st>>>(params, grad, slot0, slot1, n, lr, hyper_a, hyper_b, epsilon, grad_scale, l2_weight); break;
        case SQAIR_OPT_SGD: optimizer_kernel<SQAIR_OPT_SGD><<<blocks, 256, 0, st>>>(params, grad, slot0, slot1, n, lr, hyper_a, hyper_b, epsilon, grad_scale, l2_weight); break;
        default: return fail(SQAIR_EINVAL, "unknown optimiser kind");
    }
    CUDA_TRY(cudaGetLastError());
    return SQAIR_OK;
}

int sqair_render_sprites(const uint8_t* atlas, const int32_t* atlas_hw, const int32_t* pos, const int32_t* sprite, float* frames, int32_t T,
                         int32_t B, int32_t n, int32_t H, int32_t W, int32_t S, int32_t cell, void* stream) {
    if (!atlas || !atlas_hw || !pos || !sprite || !frames || T < 1 || B < 1 || n < 0 || H < 1 || W < 1 || S < 1 || cell < 1)
        return fail(SQAIR_EINVAL, "bad argument");
    const int64_t total = (int64_t)T * B * H * W;
    const int blocks = (int)std::min<int64_t>((total + 255) / 256, 148 * 16);
    render_sprites_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(atlas, atlas_hw, pos, sprite, frames, T, B, n, H, W, S, cell);
    CUDA_TRY(cudaGetLastError());
    return SQAIR_OK;
}

}  // extern "C"
