// sqair_train.cu -- CUDA backend of the backward pass (sqair_backward.h) and its C ABI.
//
// Kernels: `bwd_stage_kernel<STAGE>` (one thread block per row: the hand-written adjoints of the element-wise stages),
// `dgrad_kernel` (dX = dY . W^T for M = rows or rows x slots, fp32 FFMA, activation derivative fused into the operand
// load, input-segment scatter / accumulate fused into the epilogue), `wgrad_kernel` (dW += X^T . dY over M = T x rows x
// slots on the tensor cores, fp32-faithful tf32 split), column sums, and the packing / unpacking between the reference's
// variables and the per-layer virtual matrices.
#include <cuda_runtime.h>
#include <algorithm>
#include <math.h>
#include <stdio.h>
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
    __device__ __forceinline__ void sync() { __syncthreads(); }
    // dst[0..3] += the warp's sums of v[0..3] (all 32 lanes must call; one shared-memory atomic per warp and value)
    __device__ __forceinline__ void warp_add4(float* dst, const float (&v)[4]) {
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
        __syncthreads();
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            float x = 0.f;
            for (int w = 0; w < nw; ++w) x += red[q * 32 + w];
            v[q] = x;
        }
        __syncthreads();
    }
};

template <int STAGE>
__global__ void bwd_stage_kernel(const __grid_constant__ BwdCtx c, int t, int s) {
    extern __shared__ float bw_smem[];
    __shared__ float red[4 * 32];
    pdl_trigger();
    pdl_wait();
    DevEx ex;
    ex.tid = threadIdx.x; ex.nt = blockDim.x; ex.scratch = bw_smem; ex.red = red;
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

template <int ACT>
__global__ void __launch_bounds__(DG_THREADS, 1) dgrad_kernel(const __grid_constant__ DgradDev D) {
    const DgradArgs& A = D.a;
    extern __shared__ __align__(16) float dg_smem[];
    float* As = dg_smem;                                  // [DG_NP][DG_LDA]  (rows of the tile contiguous: 16-byte loads of 4 rows)
    float* Ws = dg_smem + DG_NP * DG_LDA;                 // [DG_NP][DG_BK]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int m0 = blockIdx.y * DG_BM, kb = blockIdx.x * DG_BK;
    const bool has_k = A.nseg > 0 && kb < A.K;
    pdl_trigger();
    if (!has_k && blockIdx.x != 0) return;
    const bool store_dy = A.dy.p != nullptr && blockIdx.x == 0;
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
            if (nb == 0) pdl_wait();                             // the weight loads above are in flight across the wait
#pragma unroll
            for (int j = 0; j < NW; ++j) {
                const int i = tid + j * DG_THREADS, n = i >> 3, q = i & 7;
                if (n < np) *reinterpret_cast<float4*>(Ws + n * DG_BK + 4 * q) = wv[j];
            }
        } else if (nb == 0) {
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
#pragma unroll 4
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
// backend
// ---------------------------------------------------------------------------------------------
struct CudaBackend {
    cudaStream_t st;
    const Shape* sh;
    int rows;
    int scratch_bytes;
    cudaError_t err = cudaSuccess;
    long launches = 0, tc_launches = 0;

    void check() {
        if (err == cudaSuccess) err = cudaGetLastError();
        ++launches;
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
        {   // Blackwell path (sqair_wgrad_tc.cu): TMA + tcgen05.mma + TMEM, whenever TMA can describe both operands
            const sqi::TcOperand ox{A.x.p, A.x.outer, A.x.inner, A.ny}, oy{A.dy.p, A.dy.outer, A.dy.inner, A.ny};
            if (sqi::wgrad_tc_supported(ox, oy, A.M, A.K, A.N)) {
                if (sqi::wgrad_tc(ox, oy, A.dw, A.ldw, A.M, A.K, A.N, st) != SQAIR_OK && err == cudaSuccess) err = cudaErrorInvalidValue;
                ++launches; ++tc_launches;
                return;
            }
        }
        const int tiles = ((A.N + WG_T - 1) / WG_T) * ((A.K + WG_T - 1) / WG_T);
        int msplit = (4 * 148 + tiles - 1) / tiles;
        const int max_split = (A.M + 4 * WG_MC - 1) / (4 * WG_MC);
        if (msplit > max_split) msplit = max_split;
        if (msplit < 1) msplit = 1;
        const int m_per_block = ((A.M + msplit - 1) / msplit + WG_MC - 1) / WG_MC * WG_MC;
        msplit = (A.M + m_per_block - 1) / m_per_block;
        wgrad_addr_kernel<<<dim3((A.N + WG_T - 1) / WG_T, (A.K + WG_T - 1) / WG_T, msplit), 128, 0, st>>>(A, m_per_block);
        check();
    }
    void colsum(const ColsumArgs& A) {
        const int gx = (A.N + 31) / 32;
        int msplit = (2 * 148 + gx - 1) / gx;
        const int max_split = (A.M + 63) / 64;
        if (msplit > max_split) msplit = max_split;
        if (msplit < 1) msplit = 1;
        const int m_per_block = (A.M + msplit - 1) / msplit;
        cudaError_t e = launch_pdl(colsum_kernel, dim3(gx, (A.M + m_per_block - 1) / m_per_block), dim3(256), 0, st, A, m_per_block);
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
        img_reduce_kernel<<<(TB * nh + 255) / 256, 256, 0, st>>>(dy, out, TB, K, nh);
        check();
    }
    void unpack(const float* dwv, float* d_params) {
        BPieceTab bt;
        fill_piece_tab(*sh, bt);
        unpack_backward_kernel<<<dim3(32, bt.n), 256, 0, st>>>(bt, dwv, d_params);
        check();
    }
    void small_to_params(const float* small, float* d_params, const POff& po) {
        small_to_params_kernel<<<1, 32, 0, st>>>(small, d_params, po);
        check();
    }
};

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
    out->backward_param_floats = sh.plan.bw_total + sh.plan.bwt_total;
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
    CUDA_TRY(cudaMemsetAsync(bw_params, 0, (size_t)(sh.plan.bw_total + sh.plan.bwt_total) * sizeof(float), st));
    pack_backward_kernel<<<dim3(32, bt.n), 256, 0, st>>>(bt, params, bw_params);
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
    be.scratch_bytes = bw_stage_scratch_floats(*cfg) * (int)sizeof(float);
    if (be.scratch_bytes > 48 * 1024) return fail(SQAIR_EUNSUPPORTED, "glimpses do not fit the shared memory of the canvas stage");
    static_assert(sizeof(BwdCtx) <= 4000, "BwdCtx must fit the kernel parameter space");
    BwdDriver<CudaBackend> drv(be, *cfg, sh.plan, sh.plan.poc, BL, in);
    drv.param_count_ = tab.back().offset + tab.back().count;
    drv.run(d_params);
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
