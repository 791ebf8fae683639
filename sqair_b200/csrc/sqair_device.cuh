// sqair_device.cuh -- the SQAIR per-frame step for one block of R rows.
//
// Single source: compiled by nvcc as device code for the persistent sequence kernel
// (sqair_api.cu) and, with SQAIR_HOST_EMU defined, by g++ as a one-thread-per-block emulation used
// ONLY by tests/host_emu to debug the kernel logic against the oracle without a GPU.
//
// Reference walk (all under /root/reference/sqair): seq.py:181-276 (frame body) ->
// sqair_modules.py:446-582 (SQAIRTimestep) -> sqair_modules.py:250-329 + propagate.py:68-184 +
// core.py:280-359 (propagation) -> sqair_modules.py:368-385 (latent summary) ->
// sqair_modules.py:94-229 + core.py:164-227 (discovery) -> modules.py:548-607, prior.py:61-102
// (discovery priors / number-of-steps posterior) -> index.py:132-221 (slot compaction, ids) ->
// modules.py:435-467 (decoder, canvas) -> seq.py:271-276 (log weights).
#pragma once
#include <math.h>
#include "sqair_core.h"

#ifdef SQAIR_HOST_EMU
#include <atomic>
#include <cstdlib>
#define SQ_DEV inline
#define SQ_DEVNI inline
#define SQ_LDG(p) (*(p))
#define SQ_STCS(p, v) (*(p) = (v))
#define SQ_RESTRICT
#else
#define SQ_DEV __device__ __forceinline__
#define SQ_DEVNI __device__ __noinline__
#define SQ_LDG(p) __ldg(p)
#define SQ_STCS(p, v) __stcs((p), (v))      // streaming store (evict first): the stash is written once, read a pass later, and larger than L2
#define SQ_RESTRICT __restrict__
#endif

namespace sq {

// Address spaces.  On the device the dynamic shared memory and the plan are referenced through their
// own symbols so that the compiler emits LDS/STS and constant-bank loads (a pointer carried in a struct
// would degrade every access to a generic load with 64-bit address arithmetic).
#ifdef SQAIR_HOST_EMU
#define SQ_SM c.sm
#define SQ_PLAN (*c.plan)
#define SQ_PLAN_OF(ctx) (*(ctx).plan)
#else
extern __shared__ __align__(128) float g_smem[];
__constant__ PlanHdr c_plan;
#define SQ_SM g_smem
#define SQ_PLAN c_plan
#define SQ_PLAN_OF(ctx) c_plan
#endif

#ifdef SQAIR_HOST_EMU
// split-phase reusable barrier for the emulated cluster (one host thread per block)
struct EmuClusterBarrier {
    std::atomic<long> count{0};
    int n = 1;
};
#endif

// Per-thread view of the block: indices, shared memory, position in the cluster and the state of
// the call sequence (identical in all threads of a block; advanced deterministically).
struct Ctx {
#ifdef SQAIR_HOST_EMU
    int tid_, nthreads_, lane_, nlanes_, warp_, nwarps_, ncompute_, rank_, ncta_;
    float* sm;               // this block's "shared memory"
    const Plan* plan;
    float** peers;           // shared memory of every block of the cluster
    EmuClusterBarrier* cb;
    long cb_gen;
    int tid() const { return tid_; }
    int nthreads() const { return nthreads_; }
    int lane() const { return lane_; }
    int nlanes() const { return nlanes_; }
    int warp() const { return warp_; }
    int nwarps() const { return nwarps_; }
    int ncompute() const { return ncompute_; }
    int rank() const { return rank_; }
    int ncta() const { return ncta_; }
#else
    // Immutable coordinates come from special registers (this struct's address escapes into the non-inlined
    // dense(), so anything stored here lives in local memory).
    SQ_DEV int tid() const { return (int)threadIdx.x; }
    SQ_DEV int nthreads() const { return (int)blockDim.x; }
    SQ_DEV int lane() const { return (int)(threadIdx.x & 31); }
    SQ_DEV int nlanes() const { return 32; }
    SQ_DEV int warp() const { return (int)(threadIdx.x >> 5); }
    SQ_DEV int nwarps() const { return (int)(blockDim.x >> 5); }
    SQ_DEV int ncompute() const { return NT; }
    SQ_DEV int rank() const { uint32_t r; asm("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return (int)r; }
    SQ_DEV int ncta() const { uint32_t r; asm("mov.u32 %0, %%cluster_nctarank;" : "=r"(r)); return (int)r; }
#endif
    // Host emulation only: position in Plan::seq (checked against the layer id of every dense call).  On the device the
    // position and the descriptor slot live in registers of the per-block program (Block::call_idx_, desc_cur_).
    int call_idx;            // dense calls executed in the current frame
#if defined(SQAIR_PROFILE)
    long long prof[8];       // cycle counters
    long long t_last;
#endif
    SQ_DEV void sync() const {
#ifndef SQAIR_HOST_EMU
        __syncthreads();
#endif
    }
};

#ifndef SQAIR_HOST_EMU
// ---- PTX wrappers: tensor-core MMA, streaming load, cluster barrier, distributed shared memory ----
SQ_DEV uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
SQ_DEV void cluster_arrive_() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
SQ_DEV void cluster_wait_() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
SQ_DEV void st_cluster_f32(uint32_t local_addr, int rank, float v) {
    uint32_t ra;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(local_addr), "r"(rank));
    asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(ra), "f"(v) : "memory");
}
// mbarrier + bulk asynchronous copy (TMA, cp.async.bulk): the frame copy at the start of every frame
SQ_DEV void mbar_init(uint32_t bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
SQ_DEV void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
SQ_DEV bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
SQ_DEV void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
SQ_DEV void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
// D = A(16x8, tf32) * B(8x8, tf32)
SQ_DEV void mma_tf32_zero(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%10,%10,%10};"
                 : "=f"(d[0]), "=f"(d[1]), "=f"(d[2]), "=f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1), "f"(0.f));
}
// D += A(16x8, tf32) * B(8x8, tf32), fp32 accumulate
SQ_DEV void mma_tf32(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// plain read-only loads: when streaming they sustain 117 GB/s per SM, L1::no_allocate only 74 (tools/ldg_stream.cu)
#ifndef SQAIR_LDG_KIND
#define SQAIR_LDG_KIND 0
#endif
SQ_DEV float4 ldg_stream(const float4* p) {
    float4 v;
#if SQAIR_LDG_KIND == 1
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
#elif SQAIR_LDG_KIND == 2
    asm volatile("ld.global.cg.v4.f32 {%0,%1,%2,%3}, [%4];"
#else
    asm volatile("ld.global.nc.v4.f32 {%0,%1,%2,%3}, [%4];"
#endif
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}
// activations: fp32 -> (hi, lo): hi = x truncated to the tf32 mantissa (so hi + r == x exactly), lo = r rounded to tf32
SQ_DEV void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
    hi = __float_as_uint(x) & 0xffffe000u;
    const float r = x - __uint_as_float(hi);
#ifdef SQAIR_TF32_TRUNC_LO
    lo = __float_as_uint(r);                     // the tensor core ignores the low 13 bits
#else
    lo = __float_as_uint(r) + 0x1000u;           // round to nearest (ties away) before the tensor core drops the low 13 bits
#endif
}
#endif

#if defined(SQAIR_PROFILE) && !defined(SQAIR_HOST_EMU)
__device__ long long g_trace[3][L_COUNT];      // per layer id: cycles inside dense, cycles since previous dense exit, calls
__device__ long long g_trace_last;
#endif


// cluster barrier, split phase: arrive (release) ... wait (acquire)
SQ_DEV void cluster_arrive(Ctx& c) {
#ifdef SQAIR_HOST_EMU
    c.cb->count.fetch_add(1, std::memory_order_acq_rel);
    c.cb_gen += 1;
#else
    cluster_arrive_();
#endif
}
SQ_DEV void cluster_wait(Ctx& c) {
#ifdef SQAIR_HOST_EMU
    while (c.cb->count.load(std::memory_order_acquire) < c.cb_gen * c.cb->n) {}
#else
    cluster_wait_();
#endif
}
// store one float at the same shared-memory offset of block `rank` of the cluster
SQ_DEV void store_peer(Ctx& c, int off, int rank, float v) {
#ifdef SQAIR_HOST_EMU
    c.peers[rank][off] = v;
#else
    st_cluster_f32(smem_u32(SQ_SM + off), rank, v);
#endif
}

#if defined(SQAIR_PROFILE) && !defined(SQAIR_HOST_EMU)
#define SQ_TICK(c, slot) do { long long t_ = clock64(); (c).prof[slot] += t_ - (c).t_last; (c).t_last = t_; } while (0)
#else
#define SQ_TICK(c, slot) do { } while (0)
#endif

SQ_DEV float warp_sum(float v) {
#ifndef SQAIR_HOST_EMU
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
#endif
    return v;
}

// ---- scalar math (fp32, accurate libm variants: parity needs fp32-faithful logits) -----------
SQ_DEV float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }
SQ_DEV float eluf_(float x) { return x > 0.f ? x : expm1f(x); }
SQ_DEV float softplusf_(float x) { return x > 20.f ? x : log1pf(expf(x)); }
SQ_DEV float actf(int act, float v) {
    switch (act) {
        case ACT_ELU: return eluf_(v);
        case ACT_SIGMOID: return sigmoidf_(v);
        case ACT_TANH: return tanhf(v);
        case ACT_SOFTPLUS: return softplusf_(v);
        default: return v;
    }
}
#define SQ_LOG_2PI 1.8378770664093453f
// tfd.Normal.log_prob
SQ_DEV float normal_lp(float x, float loc, float scale) {
    float z = (x - loc) / scale;
    return -0.5f * z * z - logf(scale) - 0.5f * SQ_LOG_2PI;
}
// tfd.Bernoulli(logits).log_prob(x) = -sigmoid_cross_entropy_with_logits(labels=x, logits)
SQ_DEV float bernoulli_lp(float x, float l) {
    return -(fmaxf(l, 0.f) - l * x + log1pf(expf(-fabsf(l))));
}

// tf.contrib.resampler bilinear sample with zero padding.  `F(ix, iy)` fetches
// an in-range texel.
template <class Fetch>
SQ_DEV float bilinear_zero_pad(float x, float y, int w, int h, Fetch fetch) {
    if (!(x > -1.f && y > -1.f && x < (float)w && y < (float)h)) return 0.f;
    float fx = floorf(x), fy = floorf(y);
    float cx = fx + 1.f, cy = fy + 1.f;
    float dx = cx - x, dy = cy - y;
    int ifx = (int)fx, ify = (int)fy, icx = ifx + 1, icy = ify + 1;
    bool fx_ok = ifx >= 0 && ifx <= w - 1, cx_ok = icx >= 0 && icx <= w - 1;
    bool fy_ok = ify >= 0 && ify <= h - 1, cy_ok = icy >= 0 && icy <= h - 1;
    float v00 = (fx_ok && fy_ok) ? fetch(ifx, ify) : 0.f;
    float v11 = (cx_ok && cy_ok) ? fetch(icx, icy) : 0.f;
    float v01 = (fx_ok && cy_ok) ? fetch(ifx, icy) : 0.f;
    float v10 = (cx_ok && fy_ok) ? fetch(icx, ify) : 0.f;
    return dx * dy * v00 + (1.f - dx) * (1.f - dy) * v11 + dx * (1.f - dy) * v01 + (1.f - dx) * dy * v10;
}
// i-th point of linspace(-1, 1, n) (symmetric evaluation, as torch/numpy do)
SQ_DEV float lin11(int i, int n) {
    float step = 2.f / (float)(n - 1);
    return (i < n / 2) ? (-1.f + step * (float)i) : (1.f - step * (float)(n - 1 - i));
}

// Everything a block needs for one call.
struct Job {
    const float* prm;        // packed parameters
    const float* obs;        // [T][B][P]
    const float* eps_where;  // [T][rows][2n][4]
    const float* eps_what;   // [T][rows][2n][nw]
    const float* u_pres;     // [T][rows][2n]
    sqair_outputs out;
    int debug_flags;         // tuning experiments only: 1 = skip the MMA math (garbage results), 4/8/16/32/64 skip other stages, 128 / 256 skip the stash copies of the element-wise stages / of the dense epilogue
    const float* ltab;       // layer table of this launch shape: L_COUNT x DESC_WORDS words (library-owned device buffer)
    float* stash;            // training stash (build_stash layout) or nullptr (inference)
    // generation (seq.py:46,198-203 `sample_from_prior` / `generate_after`): a second noise set for the draws from the
    // priors (same shapes; the propagation slots 0 .. n-1 are used) and the last frame that keeps the posterior samples
    const float* eps_where_prior;
    const float* eps_what_prior;
    const float* u_pres_prior;
    int generate_after;
};

SQ_DEV bool layer_has_work(const Layer& L, int rank) { return !L.split || rank < L.npanel; }
SQ_DEV const float* panel_ptr(const Layer& L, const float* prm, int rank) {
    return prm + L.w_off + (size_t)(L.split ? rank : 0) * L.panel_floats;
}

SQ_DEV void calls_init(Ctx& c, const float* ltab) {
    const auto& P = SQ_PLAN;
#ifdef SQAIR_HOST_EMU
    c.call_idx = 0;
    (void)ltab;
#else
    if (c.tid() < DESC_WORDS) SQ_SM[P.sm.Desc + c.tid()] = SQ_LDG(ltab + (int)P.seq[0] * DESC_WORDS + c.tid());
    c.sync();
#endif
}

// head that owns virtual column vc (heads start at multiples of 4); returns -1 for padding columns
SQ_DEV int head_of(const Layer& L, int vc, int& j) {
    for (int h = 0; h < L.nhead; ++h) {
        j = vc - L.head[h].col0;
        if (j >= 0 && j < L.head[h].N) return h;
    }
    return -1;
}

#ifndef SQAIR_HOST_EMU
#ifndef SQAIR_MMA_U
#define SQAIR_MMA_U 2        // k-steps in flight (2 LDG.128 each)
#endif
constexpr int MMA_U = SQAIR_MMA_U;   // k-steps (LDG.128 per lane) in flight per warp

// A fragment of one k-step as stored: tf32 hi and lo parts of the four weights of this lane
struct AFrag {
    float4 hi, lo;
};
SQ_DEV AFrag ldg_afrag(const float4* p) {
    AFrag a;
    a.hi = ldg_stream(p);
    a.lo = ldg_stream(p + 32);
    return a;
}
SQ_DEV void afrag_bits(const AFrag& a, uint32_t (&ah)[4], uint32_t (&al)[4]) {
    ah[0] = __float_as_uint(a.hi.x); ah[1] = __float_as_uint(a.hi.y); ah[2] = __float_as_uint(a.hi.z); ah[3] = __float_as_uint(a.hi.w);
    al[0] = __float_as_uint(a.lo.x); al[1] = __float_as_uint(a.lo.y); al[2] = __float_as_uint(a.lo.z); al[3] = __float_as_uint(a.lo.w);
}

// One k-step: acc += A * B in (almost) fp32.  The weights arrive pre-split into tf32 (hi, lo) (pack time), the
// activations are split here; all four partial products of the 8 k values are summed on the tensor core, smallest
// first, starting from zero; the k-step's sum then joins the running sum with a round-to-nearest FADD.  (The tensor
// core truncates when it accumulates: chaining one accumulator through hundreds of MMAs biases the result by ~1e-5
// relative, ten times the fp32 FMA-chain error -- measured as canvas parity failures on the c2 workload.)
SQ_DEV void mma_kstep(float (&acc)[4], const AFrag& a, float b0f, float b1f) {
    uint32_t ah[4], al[4], b0h, b0l, b1h, b1l;
    afrag_bits(a, ah, al);
    split_tf32(b0f, b0h, b0l); split_tf32(b1f, b1h, b1l);
    float d[4], e[4];                               // two independent chains (an MMA has ~21 cycles of latency)
    mma_tf32_zero(d, al, b0l, b1l);
    mma_tf32_zero(e, ah, b0l, b1l);
    mma_tf32(d, al, b0h, b1h);
    mma_tf32(e, ah, b0h, b1h);
    acc[0] += d[0] + e[0]; acc[1] += d[1] + e[1]; acc[2] += d[2] + e[2]; acc[3] += d[3] + e[3];
}

// Two k-steps at once: the MMA chains of the pair are independent (more work per dependent-issue slot; the loop is
// bound by in-order issue latency with 3 warps per scheduler).
SQ_DEV void mma_kstep2(float (&acc)[4], const AFrag& a0, float p0, float p1, const AFrag& a1, float q0, float q1) {
    uint32_t ah[4], al[4], ch[4], cl[4], p0h, p0l, p1h, p1l, q0h, q0l, q1h, q1l;
    afrag_bits(a0, ah, al);
    afrag_bits(a1, ch, cl);
    split_tf32(p0, p0h, p0l); split_tf32(p1, p1h, p1l); split_tf32(q0, q0h, q0l); split_tf32(q1, q1h, q1l);
    float d[4], e[4];
    mma_tf32_zero(d, al, p0l, p1l);
    mma_tf32_zero(e, cl, q0l, q1l);
    mma_tf32(d, al, p0h, p1h);
    mma_tf32(e, cl, q0h, q1h);
    mma_tf32(d, ah, p0l, p1l);
    mma_tf32(e, ch, q0l, q1l);
    mma_tf32(d, ah, p0h, p1h);
    mma_tf32(e, ch, q0h, q1h);
    acc[0] += d[0]; acc[1] += d[1]; acc[2] += d[2]; acc[3] += d[3];
    acc[0] += e[0]; acc[1] += e[1]; acc[2] += e[2]; acc[3] += e[3];
}

// One work unit: m-tile `mt` (16 output columns) x k-steps [k0, k1) of this block's panel.  The A fragments stream
// from global memory (fragment order, 512 contiguous bytes per k-step, MMA_U k-steps in flight), the B fragments
// (activations, x[k][row]) come from shared memory -- or from the frame for SEG_IMAGE.  The last k-step of a
// segment may read up to 7 feature rows past the segment: the matching weight rows are zero and shared memory only
// ever holds finite values (it is cleared at kernel start), so those products vanish.
#ifdef SQAIR_UNIT_NOINLINE
#define SQ_UNIT SQ_DEVNI
#else
#define SQ_UNIT SQ_DEV
#endif
template <int R, bool IMAGE>
SQ_UNIT void mma_unit(const Layer& L, const float4* SQ_RESTRICT wp, int k0, int k1, int slot, const float* SQ_RESTRICT img_g,
                     int lane, float (&acc)[4], bool wait_first) {
    const int g = lane >> 2, t = lane & 3, gr = g < R ? g : R - 1;
    AFrag buf[MMA_U];
#pragma unroll
    for (int j = 0; j < MMA_U; ++j)
        if (k0 + j < k1) buf[j] = ldg_afrag(wp + j * 64);
    wp += MMA_U * 64;
    // The previous dense call left its cluster barrier open (its outputs are this call's inputs): the weight loads above
    // do not depend on them, so their L2 latency overlaps the wait.
    if (wait_first) cluster_wait_();
    int si = 0;                                        // segment that holds k-step k0
    while (si + 1 < L.nseg && L.seg[si + 1].ks0 <= k0) ++si;
    int seg_end = 0, ld4 = 0, step = 0, K = 0, kloc = 0;
    const float* xp = nullptr;
    bool image = false;
    auto enter_segment = [&](int ks) {
        const Seg S = L.seg[si];
        seg_end = (si + 1 < L.nseg) ? L.seg[si + 1].ks0 : L.ksteps;
        image = S.kind == SEG_IMAGE;
        K = S.K;
        kloc = (ks - S.ks0) * 8 + t;
        ld4 = 4 * S.ld; step = 8 * S.ld;
        xp = SQ_SM + S.x_off + slot * S.x_sstride + gr + kloc * S.ld;
    };
    enter_segment(k0);
    // B fragment of k-step ks (advances the segment cursor)
    auto fetch_b = [&](int ks, float& b0f, float& b1f) {
        if (ks == seg_end) { ++si; enter_segment(ks); }
        if (!IMAGE || !image) {
            b0f = xp[0]; b1f = xp[ld4];
            xp += step;
        } else {
            int ka = kloc, kb = kloc + 4;
            if (kb >= K) { ka = ka < K ? ka : K - 1; kb = K - 1; }
            b0f = img_g[ka]; b1f = img_g[kb];            // generic loads: the frame is in shared memory when staged
            kloc += 8;
        }
    };
    auto kstep = [&](int ks, AFrag& slot_buf, bool refill) {
        const AFrag a = slot_buf;
        float b0f, b1f;
        fetch_b(ks, b0f, b1f);
        mma_kstep(acc, a, b0f, b1f);
        if (refill) slot_buf = ldg_afrag(wp);          // issued after the MMAs that consumed this slot (both are volatile)
        wp += 64;
    };
    static_assert(MMA_U % 2 == 0, "k-steps are processed in pairs");
    int kk = k0;
    for (; kk + 2 * MMA_U <= k1; kk += MMA_U) {        // steady state: pairs of k-steps, every slot is refilled
#pragma unroll
        for (int j = 0; j < MMA_U; j += 2) {
            float p0, p1, q0, q1;
            fetch_b(kk + j, p0, p1);
            fetch_b(kk + j + 1, q0, q1);
            mma_kstep2(acc, buf[j], p0, p1, buf[j + 1], q0, q1);
            buf[j] = ldg_afrag(wp);
            buf[j + 1] = ldg_afrag(wp + 64);
            wp += 128;
        }
    }
    for (; kk < k1; kk += MMA_U) {                     // last one or two groups
#pragma unroll
        for (int j = 0; j < MMA_U; ++j)
            if (kk + j < k1) kstep(kk + j, buf[j], kk + j + MMA_U < k1);
    }
}
#endif

// ---------------------------------------------------------------------------------------------
// Dense layer (snt.Linear / Nonlinear, neural.py:34-47; VanillaRNN / GRU gate pre-activations):
//   out[col][r] = act( sum_seg sum_k W[k][col] * x_seg[k][r] + b[col] (+ b2[col]) ) * scale + add
// Block `rank` owns the virtual columns [rank*Nc, rank*Nc + Nc) when the layer is split.  Work units =
// (m-tile of 16 columns) x (k-slice), dealt round-robin to the warps; the slices' partial sums meet in shared
// memory; all threads then finish one output each (activation) and store it into every block of the cluster.
// ---------------------------------------------------------------------------------------------
// flags: bit 0 = the previous dense call deferred its cluster barrier wait (do it here, after the first weight loads are
// in flight); bit 1 = the caller promises that the next operation is another dense call, so this call may leave its
// own barrier open.  Returns true when it did.  call_idx / desc_cur: position in Plan::seq and descriptor slot (kept in
// registers by the caller).
// Training stash: `stash` != nullptr makes every finished output also go to global memory (Head::st_off), at frame
// st_t, rows row0 + r, entry st_entry of the head's signal.
template <int R, bool TR>
SQ_DEVNI bool dense(Ctx& c, const float* SQ_RESTRICT prm, const float* SQ_RESTRICT ltab, int layer_id, int slot,
                    const float* const* imgrow, const float* SQ_RESTRICT img_g, int dbg, int flags, int call_idx, int desc_cur,
                    float* SQ_RESTRICT stash, int st_t, int st_entry, int row0) {
    const auto& P = SQ_PLAN;
#ifdef SQAIR_HOST_EMU
    const Layer& L = P.L[layer_id];
    if (P.seq[c.call_idx] != layer_id) {
        fprintf(stderr, "emu: dense call %d is layer %d but Plan::seq says %d\n", c.call_idx, layer_id, P.seq[c.call_idx]);
        abort();
    }
    if (++c.call_idx >= P.nseq) c.call_idx = 0;
    (void)img_g; (void)call_idx; (void)desc_cur; (void)ltab;
#else
    if (++call_idx >= P.nseq) call_idx = 0;
    // the descriptor of this call was staged in shared memory during the previous call (before that call's block
    // barrier, so it is visible even if the cluster barrier of that call is still open); start fetching the next one
    const Layer& L = *reinterpret_cast<const Layer*>(SQ_SM + P.sm.Desc + desc_cur * DESC_WORDS);
    float next_desc_word = 0.f;
    if (c.tid() < DESC_WORDS) next_desc_word = SQ_LDG(ltab + (int)P.seq[call_idx] * DESC_WORDS + c.tid());
    (void)imgrow;
#endif
    SQ_TICK(c, 5);                               // time since the previous dense call (element-wise stages)
#if defined(SQAIR_PROFILE) && !defined(SQAIR_HOST_EMU)
    long long t_enter = 0;
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        t_enter = clock64();
        g_trace[1][layer_id] += t_enter - g_trace_last;
        g_trace[2][layer_id] += 1;
    }
#endif
    const bool exchange = L.split && c.ncta() > 1;
    const bool work = layer_has_work(L, c.rank());
    const int Nc = L.Nc;
    // Phase A ("nobody writes my buffers before I have entered this layer") is only needed if an element-wise
    // stage could still touch the output buffer of the dense call that follows it; the frame program never does
    // (audited in DESIGN.md), so it is compiled in only for debugging.
#ifdef SQAIR_SAFE_EXCHANGE
    if (exchange) cluster_arrive(c);
#endif
    float* red = SQ_SM + P.sm.Red;
    int ks = 1;
    bool waited = !(flags & 1);                  // false: the previous call's cluster barrier is still open
#ifdef SQAIR_HOST_EMU
    std::vector<float> redv;
#endif
    if (work) {
        const float* panel = panel_ptr(L, prm, c.rank());
#ifdef SQAIR_HOST_EMU
        if (!waited) { cluster_wait(c); waited = true; }
        // one sequential thread: plain fp32 dot products over the same fragment-ordered panel and segment table
        redv.assign((size_t)Nc * R, 0.f);
        for (int col = 0; col < Nc; ++col)
            for (int si = 0; si < L.nseg; ++si) {
                const Seg& S = L.seg[si];
                for (int k = 0; k < S.K; ++k) {
                    const int fo = frag_off(L.ksteps, col >> 4, S.ks0 + (k >> 3), col & 15, k & 7);
                    const float w = panel[(fo >> 7) * 256 + (fo & 127)] + panel[(fo >> 7) * 256 + 128 + (fo & 127)];   // hi + lo
                    for (int r = 0; r < R; ++r) {
                        const float x = S.kind == SEG_IMAGE ? imgrow[r][k] : SQ_SM[S.x_off + slot * S.x_sstride + k * S.ld + r];
                        redv[(size_t)col * R + r] += w * x;
                    }
                }
                // the zero padding of the packed panel must really be zero
                for (int k = S.K; k < (S.K + 7) / 8 * 8; ++k)
                    if (panel[(frag_off(L.ksteps, col >> 4, S.ks0 + (k >> 3), col & 15, k & 7) >> 7) * 256 +
                              (frag_off(L.ksteps, col >> 4, S.ks0 + (k >> 3), col & 15, k & 7) & 127)] != 0.f) {
                        fprintf(stderr, "emu: non-zero padding weight (layer %d)\n", layer_id);
                        abort();
                    }
            }
        red = redv.data();
#else
        ks = L.ksplit;
        const int nmt = L.nmt, kper = L.kper, ksteps = L.ksteps, nunits = nmt * ks;
        const int lane = c.lane(), g = lane >> 2, t = lane & 3;
        SQ_TICK(c, 0);                           // call prologue: counters, descriptor
        for (int u = c.warp(); u < ((dbg & 1) ? 0 : nunits); u += NWARP) {
            const int sl = u / nmt, mt = u - sl * nmt;
            const int k0 = sl * kper, k1 = (k0 + kper < ksteps) ? (k0 + kper) : ksteps;
#if defined(SQAIR_PROFILE)
            c.prof[6] += k1 - k0; c.prof[7] += 1;
#endif
            float acc[4] = {0.f, 0.f, 0.f, 0.f};
            const float4* wp = reinterpret_cast<const float4*>(panel) + ((size_t)mt * ksteps + k0) * 64 + lane;
            if (L.seg[0].kind == SEG_IMAGE) mma_unit<R, true>(L, wp, k0, k1, slot, img_g, lane, acc, !waited);
            else mma_unit<R, false>(L, wp, k0, k1, slot, img_g, lane, acc, !waited);
            waited = true;
            // C fragment: acc[0] = (col g, row 2t), acc[1] = (g, 2t+1), acc[2] = (g+8, 2t), acc[3] = (g+8, 2t+1)
            float* rp = red + ((size_t)sl * Nc + mt * 16 + g) * R;
            if (2 * t < R) { rp[2 * t] = acc[0]; rp[8 * R + 2 * t] = acc[2]; }
            if (2 * t + 1 < R) { rp[2 * t + 1] = acc[1]; rp[8 * R + 2 * t + 1] = acc[3]; }
        }
        SQ_TICK(c, 1);
#endif
    }
    if (!waited) cluster_wait(c);                // warps without a unit in this layer (and the emulator)
#ifndef SQAIR_HOST_EMU
    if (c.tid() < DESC_WORDS) SQ_SM[P.sm.Desc + (desc_cur ^ 1) * DESC_WORDS + c.tid()] = next_desc_word;
#endif
    c.sync();
    SQ_TICK(c, 2);
#ifdef SQAIR_SAFE_EXCHANGE
    if (exchange) cluster_wait(c);               // phase A complete: every block has entered this layer
#endif
    if (work) {
        const int vbase = L.split ? c.rank() * Nc : 0;
        for (int o = c.tid(); o < ((dbg & 16) ? 0 : Nc * R); o += c.nthreads()) {
            const int col = o / R, r = o - col * R;
            int j;
            const int h = head_of(L, vbase + col, j);
            if (h < 0) continue;
            const Head& H = L.head[h];
            float v = 0.f;
            for (int s2 = 0; s2 < ks; ++s2) v += red[((size_t)s2 * Nc + col) * R + r];
            v = actf(H.act, v) * H.scale + H.add;                     // the bias arrived through the product (constant-1 input)
            if (H.scale_p_off >= 0) v *= SQ_LDG(prm + H.scale_p_off);
            const int off = H.out_off + slot * H.out_sstride + j * H.out_ld + r;
            if (exchange) {
                for (int q = 0; q < c.ncta(); ++q) store_peer(c, off, q, v);
            } else {
                SQ_SM[off] = v;
            }
            if (TR && stash != nullptr && !(dbg & 256) && H.st_off >= 0 && row0 + r < P.rows && (L.split || c.rank() == 0))
                SQ_STCS(&stash[(size_t)H.st_off + ((size_t)(st_t * P.rows + row0 + r) * H.st_entries + st_entry) * H.st_width + j], v);
        }
    }
    SQ_TICK(c, 3);
    bool deferred = false;
    if (exchange) {                              // phase B: all slices have landed everywhere
        cluster_arrive(c);
#ifndef SQAIR_SAFE_EXCHANGE
        deferred = (flags & 2) != 0;
#endif
        if (!deferred) cluster_wait(c);
    } else {
        c.sync();
    }
    SQ_TICK(c, 4);
#if defined(SQAIR_PROFILE) && !defined(SQAIR_HOST_EMU)
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        const long long t_exit = clock64();
        g_trace[0][layer_id] += t_exit - t_enter;
        g_trace_last = t_exit;
    }
#endif
    return deferred;
}

// ---------------------------------------------------------------------------------------------
// The per-block state machine
// ---------------------------------------------------------------------------------------------
#ifndef SQAIR_HOST_EMU
#define P c_plan
#endif
// TR = false compiles the training stash out (the inference kernel); GEN = true is the `sample_from_prior` variant
template <int R, bool TR = true, bool GEN = false>
struct Block {
    Ctx& c;
#ifdef SQAIR_HOST_EMU
    const Plan& P;
#else
    // no member on the device: `P` is the __constant__ symbol itself (see the macro above the struct), so that every
    // plan field is a constant-bank load; through a reference member the compiler emitted ~2 000 generic LD.E
#endif
    const Job& J;            // the __grid_constant__ kernel parameter; only the output table is read through it
    const float* prm_;       // hot fields of the job, copied once (through the reference they were generic loads)
    const float* obs_;
    const float* eps_where_;
    const float* eps_what_;
    const float* u_pres_;
    int dbg_;
    const float* ltab_;
    float* stash_;                       // training stash or nullptr
    int t_;                              // current frame
    mutable int call_idx_, desc_cur_;    // position in Plan::seq / descriptor slot of the next dense call
    mutable bool pend_;                  // the last dense call left its cluster barrier open
    int row0;                       // first global row of this block
    const float* imgrow[R];         // frame of each row for the current t
    const float* img_g;             // frame of the row this lane feeds to the MMA B fragment (row min(lane / 4, R - 1))
    uint32_t frame_parity;          // mbarrier phase of the current frame's TMA copy

#ifdef SQAIR_HOST_EMU
    SQ_DEV Block(Ctx& c_, const Job& J_, int row0_) : c(c_), P(SQ_PLAN_OF(c_)), J(J_), row0(row0_) {
#else
    SQ_DEV Block(Ctx& c_, const Job& J_, int row0_) : c(c_), J(J_), row0(row0_) {
#endif
        call_idx_ = 0; desc_cur_ = 0; pend_ = false;
        prm_ = J_.prm; obs_ = J_.obs; eps_where_ = J_.eps_where; eps_what_ = J_.eps_what; u_pres_ = J_.u_pres; dbg_ = J_.debug_flags;
        ltab_ = J_.ltab; stash_ = J_.stash; t_ = 0;
#pragma unroll
        for (int r = 0; r < R; ++r) {
            imgrow[r] = nullptr;
        }
        img_g = nullptr;
        frame_parity = 0;
    }
    // accessors ------------------------------------------------------------------------------
    SQ_DEV float* sm() const { return SQ_SM; }
    SQ_DEV int LDS() const { return P.NS * R; }
    SQ_DEV int LDE() const { return (P.NS + 1) * R; }
    SQ_DEV float& Z(int f, int s, int r) const { return SQ_SM[P.sm.Z + f * LDS() + s * R + r]; }
    SQ_DEV float& rec(int base, int e, int f, int r) const { return SQ_SM[base + f * LDE() + e * R + r]; }
    SQ_DEV float& pri(int f, int s, int r) const { return SQ_SM[P.sm.Pri + f * LDS() + s * R + r]; }
    SQ_DEV float& lp(int k, int s, int r) const { return SQ_SM[P.sm.Lp + k * LDS() + s * R + r]; }
    SQ_DEV float& rowacc(int k, int r) const { return SQ_SM[P.sm.RowAcc + k * R + r]; }
    SQ_DEV float prm(int off) const { return SQ_LDG(prm_ + off); }
    // `next_is_dense`: nothing but another lin() follows (no element-wise stage reads or writes shared memory in
    // between), so the cluster barrier of this call may be completed inside the next one, behind its first weight loads
    // `st_entry`: entry of the training-stash signals this call's outputs belong to (default: the slot)
    SQ_DEV void lin(int id, int slot = 0, bool next_is_dense = false, int st_entry = -1) const {
        if (dbg_ & 64) return;
        const int flags = (pend_ ? 1 : 0) | (next_is_dense ? 2 : 0);
        if (st_entry < 0) st_entry = slot;
#ifdef SQAIR_HOST_EMU
        pend_ = dense<R, TR>(c, prm_, ltab_, id, slot, imgrow, img_g, dbg_, flags, call_idx_, desc_cur_, stash_, t_, st_entry, row0);
#else
        pend_ = dense<R, TR>(c, prm_, ltab_, id, slot, nullptr, img_g, dbg_, flags, call_idx_, desc_cur_, stash_, t_, st_entry, row0);   // (passing imgrow would pin it to local memory)
#endif
        if (++call_idx_ >= P.nseq) call_idx_ = 0;
        desc_cur_ ^= 1;
    }
    SQ_DEV size_t nidx(int t, int r, int slot2) const {      // noise index of (t, row, slot in [0,2n))
        return ((size_t)t * P.rows + grow_of(r)) * (2 * P.NS) + slot2;
    }
    // Training stash of a signal that an element-wise stage produced: nfeat features of entry `entry`, read from
    // shared memory at smem_off + f * fstride + r.  The (replicated) state is split by features across the cluster's
    // blocks; consecutive threads write consecutive features of one row.
    SQ_DEV void stash_copy(int sig, int t, int entry, int smem_off, int fstride, int nfeat, int col0 = 0) const {
        if (!TR || stash_ == nullptr || (dbg_ & 128)) return;
        const Sig g = P.st[sig];
        const int per = (nfeat + c.ncta() - 1) / c.ncta(), f0 = c.rank() * per;
        const int cnt = (f0 + per < nfeat ? f0 + per : nfeat) - f0;
        for (int i = c.tid(); i < cnt * R; i += c.nthreads()) {
            const int r = i / cnt, f = f0 + i - r * cnt;
            if (valid_row(r))
                SQ_STCS(&stash_[(size_t)g.off + ((size_t)(t * P.rows + row0 + r) * g.entries + entry) * g.width + col0 + f],
                        SQ_SM[smem_off + f * fstride + r]);
        }
    }
    enum { RA_QPRES = 0, RA_PPRES = 1, RA_NPROP = 2, RA_NDISC = 3, RA_QNUM = 4, RA_PNUM = 5, RA_LL = 6 };
    enum { LP_PQWHAT = 0, LP_PQWHERE = 1, LP_PPWHAT = 2, LP_PPWHERE = 3, LP_PROB = 4,
           LP_DQWHAT = 5, LP_DQWHERE = 6, LP_DPWHAT = 7, LP_DPWHERE = 8 };

    // ------------------------------------------------------------------------------------------
    // sequence start: seq.py:86-104 (initial z = 0, ids = -1, trainable initial GRU states)
    // ------------------------------------------------------------------------------------------
    SQ_DEV void init_sequence() const {
        const Smem& m = P.sm;
        const int nh = P.nh, nw = P.nw, NS = P.NS;
#ifndef SQAIR_HOST_EMU
        // every shared-memory word must be finite: the MMA B fragments may over-read into neighbouring buffers
        // (against zero weights).  The staged descriptor and the counters written by calls_init() stay.
        for (int i = m.Desc + 2 * DESC_WORDS + c.tid(); i < m.total; i += c.nthreads()) SQ_SM[i] = 0.f;
        c.sync();
        if (c.tid() == 0) { mbar_init(smem_u32(SQ_SM + m.ImgBar), 1); fence_barrier_init(); }
        c.sync();
#endif
        for (int i = c.tid(); i < (nw + 6) * LDS(); i += c.nthreads()) SQ_SM[m.Z + i] = 0.f;
        for (int i = c.tid(); i < LDS(); i += c.nthreads()) SQ_SM[m.Ids + i] = -1.f;
        for (int i = c.tid(); i < R; i += c.nthreads()) { SQ_SM[m.LastId + i] = -1.f; SQ_SM[m.Ones + i] = 1.f; }
        for (int i = c.tid(); i < nh * LDS(); i += c.nthreads()) {
            int f = i / LDS();
            SQ_SM[m.Tst + i] = prm(P.po.temporal_h0 + f);
            SQ_SM[m.Pst + i] = prm(P.po.prior_h0 + f);
        }
        // entry 0 of the slot records = RNN-core initial state (core.py:132-139,153,238)
        for (int i = c.tid(); i < (nw + 5) * R; i += c.nthreads()) {
            int f = i / R, r = i % R;
            rec(m.PropOut, 0, f, r) = 0.f;
            rec(m.DiscOut, 0, f, r) = (f == P.rec.pres) ? 1.f : 0.f;
        }
        if (P.cfg.rec_where_prior)
            for (int i = c.tid(); i < 4 * R; i += c.nthreads()) SQ_SM[m.RnInit + i] = prm(P.po.rn_init_state + i / R);
        c.sync();
        for (int s = 0; s < NS; ++s) stash_copy(S_Z, 0, s, m.Z + s * R, LDS(), nw + 6);
    }

    // the frame's bulk copy has landed (returns at once when it already has; no-op when frames are read from global)
    SQ_DEV void wait_frame() const {
#ifndef SQAIR_HOST_EMU
        if (P.sm.img_n > 0) {
            const uint32_t bar = smem_u32(SQ_SM + P.sm.ImgBar);
            while (!mbar_try_wait(bar, frame_parity)) {}
        }
#endif
    }
    // forward spatial transformer (modules.py:165-172,204-218) at Coords -> Glm (* Mask)
    SQ_DEV void extract_glimpse(bool use_mask, int st_entry) const {
        const Smem& m = P.sm;
        const int G = P.cfg.G, W = P.cfg.W, H = P.cfg.H;
        const float hw = 0.5f * (float)(W - 1), hh = 0.5f * (float)(H - 1);
        wait_frame();
        for (int i = c.tid(); i < ((dbg_ & 8) ? 0 : P.g * R); i += c.nthreads()) {
            const int r = i % R, j = i / R;
            const int gy = j / G, gx = j % G;
            const float sx = SQ_SM[m.Coords + 0 * R + r], sy = SQ_SM[m.Coords + 1 * R + r];
            const float tx = SQ_SM[m.Coords + 2 * R + r], ty = SQ_SM[m.Coords + 3 * R + r];
            const float x = hw * (sx * lin11(gx, G) + tx) + hw;
            const float y = hh * (sy * lin11(gy, G) + ty) + hh;
            const float* img = imgrow[0];
#pragma unroll
            for (int q = 1; q < R; ++q) if (r == q) img = imgrow[q];
            float v = bilinear_zero_pad(x, y, W, H, [&](int ix, int iy) { return img[iy * W + ix]; });
            if (use_mask) v *= SQ_SM[m.Mask + i];
            SQ_SM[m.Glm + i] = v;
        }
        c.sync();
        stash_copy(S_GLM, t_, st_entry, m.Glm, R, P.g);
    }
    // to_coords (modules.py:220-227) + clip_preserve(scale, 1e-4) (modules.py:206)
    SQ_DEV void set_coords(float w0, float w1, float w2, float w3, int r) const {
        const Smem& m = P.sm;
        SQ_SM[m.Coords + 0 * R + r] = fmaxf(sigmoidf_(w0), 1e-4f);
        SQ_SM[m.Coords + 1 * R + r] = fmaxf(sigmoidf_(w1), 1e-4f);
        SQ_SM[m.Coords + 2 * R + r] = tanhf(w2);
        SQ_SM[m.Coords + 3 * R + r] = tanhf(w3);
    }
    SQ_DEV void encode_glimpse(int last_layer, int st_entry, bool next_is_dense = false) const {
        lin(L_ENC1, 0, true, st_entry); lin(L_ENC2, 0, true, st_entry); lin(last_layer, 0, next_is_dense, st_entry);
    }
    // snt.GRU gate algebra (Appendix B) around the two dense calls: Gr <- r*h, then state update.
    SQ_DEV void gru_mul_r(int state_off, int s, int sig_rh) const {
        const Smem& m = P.sm;
        for (int i = c.tid(); i < P.nh * R; i += c.nthreads()) {
            int f = i / R, r = i % R;
            SQ_SM[m.Gr + i] *= SQ_SM[state_off + f * LDS() + s * R + r];
        }
        c.sync();
        stash_copy(sig_rh, t_, s, m.Gr, R, P.nh);
    }

    // ------------------------------------------------------------------------------------------
    // propagation prior for slot s (propagate.py:68-98,123-158); updates Pst in place
    // ------------------------------------------------------------------------------------------
    SQ_DEV void prop_prior(int s) const {
        const Smem& m = P.sm;
        const int nh = P.nh, nw = P.nw;
        lin(L_PGRU_ZR, s);
        gru_mul_r(m.Pst, s, S_PGRH);
        lin(L_PGRU_C, s);
        for (int i = c.tid(); i < nh * R; i += c.nthreads()) {
            int f = i / R, r = i % R;
            float& h = SQ_SM[m.Pst + f * LDS() + s * R + r];
            float z = SQ_SM[m.Gz + i];
            h = (1.f - z) * h + z * SQ_SM[m.Gc + i];
        }
        c.sync();
        stash_copy(S_PSTNEW, t_, s, m.Pst + s * R, LDS(), nh);
        lin(L_PLIN, s);
        // stats post-processing: rows 0 logit | 1..4 where_loc | 5..4+nw what_loc | where_scale(4) | what_scale(nw)
        const int nstat = 2 * (4 + nw) + 1;
        for (int i = c.tid(); i < nstat * R; i += c.nthreads()) {
            int f = i / R, r = i % R;
            float v = pri(f, s, r);
            const float ptm1 = Z(nw + 4, s, r);
            if (f == 0) {
                v += P.cfg.prop_prior_step_bias;
                v = ptm1 * v + (ptm1 - 1.f) * 88.f;
                if (P.cfg.prior_type != SQAIR_PRIOR_RNN) v = Z(nw + 5, s, r) + 0.1f * v;
            } else if (f <= 4 + nw) {
                // locs: where_loc (4) then what_loc (nw); Z rows: what 0..nw-1, where nw..nw+3
                const int zf = (f <= 4) ? (nw + f - 1) : (f - 5);
                if (P.cfg.prior_type == SQAIR_PRIOR_RW) v = Z(zf, s, r);
                else if (P.cfg.prior_type == SQAIR_PRIOR_GUIDED) v = Z(zf, s, r) + 0.1f * v;
            } else {
                v = softplusf_(v) + 1e-2f;
            }
            pri(f, s, r) = v;
        }
        c.sync();
        stash_copy(S_PRI, t_, s, m.Pri + s * R, LDS(), nstat);
    }

    // ------------------------------------------------------------------------------------------
    // one propagation slot (core.py:280-359) + its log-prob terms (sqair_modules.py:281-326)
    // ------------------------------------------------------------------------------------------
    SQ_DEV void prop_slot(int t, int s) const {
        const Smem& m = P.sm;
        const RecF& F = P.rec;
        const int nh = P.nh, nw = P.nw, e = s + 1;
        const bool masked = P.cfg.masked_glimpse != 0;
        prop_prior(s);
        // where_bias MLP and glimpse mask MLP on the slot's temporal state (core.py:291; modules.py:350-356)
        lin(L_WBMK1, s, true);
        lin(L_WB2, 0, masked, s);
        if (masked) lin(L_MK2, 0, false, s);
        for (int r = c.tid(); r < R; r += c.nthreads())
            set_coords(Z(nw + 0, s, r) + SQ_SM[m.Wb + 0 * R + r], Z(nw + 1, s, r) + SQ_SM[m.Wb + 1 * R + r],
                       Z(nw + 2, s, r) + SQ_SM[m.Wb + 2 * R + r], Z(nw + 3, s, r) + SQ_SM[m.Wb + 3 * R + r], r);
        c.sync();
        extract_glimpse(masked, s);
        encode_glimpse(L_ENC3_LOC, s, true);                      // -> Loc1 (core.py:292-293)
        lin(L_PRNN, s, true, s + 1);                              // core.py:295-302 -> Hrnn[1]
        lin(L_PT1, s, true); lin(L_PT2, 0, true, s); lin(L_PT3, 0, false, s);     // core.py:323-324 -> Tp
        // where ~ MVN_TriL(where_tm1 + us*loc, L) (core.py:326-330; modules.py:535-545)
        for (int r = c.tid(); r < R; r += c.nthreads()) {
            float loc[4], sc[4], eps[4], L[4][4], wh[4];
            const float so = prm(P.po.p_scale_offset);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                loc[i] = Z(nw + i, s, r) + P.cfg.where_update_scale * SQ_SM[m.Tp + i * R + r];
                sc[i] = softplusf_(SQ_SM[m.Tp + (4 + i) * R + r] + so - 1.f) + 1e-2f;
                eps[i] = eps_where_[nidx(t, r, s) * 4 + i];
            }
            tril_from_scale(sc, L);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                float a = 0.f;
#pragma unroll
                for (int j = 0; j < 4; ++j) if (j <= i) a += L[i][j] * eps[j];
                wh[i] = loc[i] + a;
                rec(m.PropOut, e, F.where + i, r) = wh[i];
                rec(m.PropOut, e, F.where_loc + i, r) = loc[i];
                rec(m.PropOut, e, F.where_scale + i, r) = sc[i];
            }
            set_coords(wh[0], wh[1], wh[2], wh[3], r);
        }
        c.sync();
        extract_glimpse(masked, P.NS + s);
        encode_glimpse(L_ENC3, P.NS + s);                         // -> Enc = (loc2, scale2) (core.py:336-337)
        // temporal GRU (core.py:339-340); new state left in Gc, Tst updated at the end of the slot
        lin(L_TGRU_ZR, s);
        gru_mul_r(m.Tst, s, S_TGRH);
        lin(L_TGRU_C, s);
        for (int i = c.tid(); i < nh * R; i += c.nthreads()) {
            int f = i / R, r = i % R;
            float h = SQ_SM[m.Tst + f * LDS() + s * R + r], z = SQ_SM[m.Gz + i];
            SQ_SM[m.Gc + i] = (1.f - z) * h + z * SQ_SM[m.Gc + i];
        }
        c.sync();
        stash_copy(S_TSTNEW, t_, s, m.Gc, R, nh);
        lin(L_PHEADS, 0, false, s);                               // core.py:343-349 -> Tg, Gt
        for (int i = c.tid(); i < nw * R; i += c.nthreads()) {        // core.py:351-357
            int j = i / R, r = i % R;
            float fg = SQ_SM[m.Gt + i], ig = SQ_SM[m.Gt + nw * R + i], tg = SQ_SM[m.Gt + 2 * nw * R + i];
            float loc2 = SQ_SM[m.Enc + i], sc2 = SQ_SM[m.Enc + nw * R + i];
            float loct = SQ_SM[m.Tg + i], sct = SQ_SM[m.Tg + nw * R + i];
            float wl = fg * Z(j, s, r) + (1.f - ig) * loc2 + (1.f - tg) * loct;
            float ws = (1.f - ig) * sc2 + (1.f - tg) * sct;
            float what = wl + ws * eps_what_[nidx(t, r, s) * nw + j];
            rec(m.PropOut, e, F.what + j, r) = what;
            rec(m.PropOut, e, F.what_loc + j, r) = wl;
            rec(m.PropOut, e, F.what_scale + j, r) = ws;
        }
        c.sync();
        lin(L_PST1, s, true); lin(L_PST2);                        // modules.py:506-513
        for (int r = c.tid(); r < R; r += c.nthreads()) {             // core.py:141-144
            const float ptm1 = Z(nw + 4, s, r);
            float logit = ptm1 * SQ_SM[m.Lg + r] + (ptm1 - 1.f) * 88.f;
            float prob = sigmoidf_(logit);
            float pres = (u_pres_[nidx(t, r, s)] < prob ? 1.f : 0.f) * ptm1;
            rec(m.PropOut, e, F.logit, r) = logit;
            rec(m.PropOut, e, F.prob, r) = prob;
            rec(m.PropOut, e, F.pres, r) = pres;
        }
        c.sync();
        if (GEN) {      // sample_from_prior: q is evaluated at draws from the prior, after all slots ran (prop_generate)
            for (int r = c.tid(); r < R; r += c.nthreads()) rowacc(RA_NPROP, r) += rec(m.PropOut, e, F.pres, r);
        } else
        {
        // log-probs under q and p (sqair_modules.py:290-317): one warp per row, lanes over dims
        for (int r = c.warp(); r < R; r += c.nwarps()) {
            float qw = 0.f, pw = 0.f;
            for (int j = c.lane(); j < nw; j += c.nlanes()) {
                float x = rec(m.PropOut, e, F.what + j, r);
                qw += normal_lp(x, rec(m.PropOut, e, F.what_loc + j, r), rec(m.PropOut, e, F.what_scale + j, r));
                pw += normal_lp(x, pri(5 + j, s, r), pri(9 + nw + j, s, r));
            }
            qw = warp_sum(qw);
            pw = warp_sum(pw);
            if (c.lane() == 0) {
                float x[4], loc[4], sc[4], L[4][4], y[4];
                float pwh = 0.f;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    x[i] = rec(m.PropOut, e, F.where + i, r);
                    loc[i] = rec(m.PropOut, e, F.where_loc + i, r);
                    sc[i] = rec(m.PropOut, e, F.where_scale + i, r);
                    pwh += normal_lp(x[i], pri(1 + i, s, r), pri(5 + nw + i, s, r));
                }
                tril_from_scale(sc, L);
                float qwh = -2.f * SQ_LOG_2PI, ss = 0.f;
#pragma unroll
                for (int i = 0; i < 4; ++i) {               // forward substitution L y = x - loc
                    float a = x[i] - loc[i];
#pragma unroll
                    for (int j = 0; j < 4; ++j) if (j < i) a -= L[i][j] * y[j];
                    y[i] = a / L[i][i];
                    ss += y[i] * y[i];
                    qwh -= logf(fabsf(L[i][i]));
                }
                qwh += -0.5f * ss;
                const float ptm1 = Z(nw + 4, s, r), pres = rec(m.PropOut, e, F.pres, r);
                const float qp = bernoulli_lp(pres, rec(m.PropOut, e, F.logit, r));
                const float pp = bernoulli_lp(pres, pri(0, s, r));
                const float mk = ptm1 * pres;
                lp(LP_PQWHAT, s, r) = qw * mk;
                lp(LP_PQWHERE, s, r) = qwh * mk;
                lp(LP_PPWHAT, s, r) = pw * mk;
                lp(LP_PPWHERE, s, r) = pwh * mk;
                lp(LP_PROB, s, r) = expf(qp) * ptm1;
                rowacc(RA_QPRES, r) += qp * ptm1;
                rowacc(RA_PPRES, r) += pp * ptm1;
                rowacc(RA_NPROP, r) += pres;
            }
        }
        }
        // commit the slot: temporal state, RNN hidden
        for (int i = c.tid(); i < nh * R; i += c.nthreads()) {
            int f = i / R, r = i % R;
            SQ_SM[m.Tst + f * LDS() + s * R + r] = SQ_SM[m.Gc + i];
            SQ_SM[m.Hrnn + i] = SQ_SM[m.Hrnn + nh * R + i];
        }
        c.sync();
        stash_copy(S_PROPREC, t_, e, m.PropOut + e * R, LDE(), F.size);
    }

    // ------------------------------------------------------------------------------------------
    // sample_from_prior / generate_after (sqair_modules.py:281-326, after the SSM ran all slots): draws from the
    // propagation prior; the posterior is evaluated at THOSE draws; in generated frames (do_generate) they replace what /
    // where / presence of the slot.  Masks and the prior's presence term keep the posterior's presence (:286,:306).
    // ------------------------------------------------------------------------------------------
    SQ_DEV void prop_generate(int t, bool do_generate) const {
        const Smem& m = P.sm;
        const RecF& F = P.rec;
        const int nw = P.nw;
        for (int s = 0; s < P.NS; ++s) {
            const int e = s + 1;
            for (int r = c.warp(); r < R; r += c.nwarps()) {
                const size_t ni = nidx(t, r, s);
                float qw = 0.f, pw = 0.f;
                for (int j = c.lane(); j < nw; j += c.nlanes()) {
                    const float x = pri(5 + j, s, r) + pri(9 + nw + j, s, r) * J.eps_what_prior[ni * nw + j];
                    qw += normal_lp(x, rec(m.PropOut, e, F.what_loc + j, r), rec(m.PropOut, e, F.what_scale + j, r));
                    const float xp = do_generate ? x : rec(m.PropOut, e, F.what + j, r);
                    pw += normal_lp(xp, pri(5 + j, s, r), pri(9 + nw + j, s, r));
                    if (do_generate) rec(m.PropOut, e, F.what + j, r) = x;
                }
                qw = warp_sum(qw);
                pw = warp_sum(pw);
                if (c.lane() == 0) {
                    float x[4], loc[4], sc[4], L[4][4], y[4];
                    float pwh = 0.f;
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        x[i] = pri(1 + i, s, r) + pri(5 + nw + i, s, r) * J.eps_where_prior[ni * 4 + i];
                        loc[i] = rec(m.PropOut, e, F.where_loc + i, r);
                        sc[i] = rec(m.PropOut, e, F.where_scale + i, r);
                        const float xp = do_generate ? x[i] : rec(m.PropOut, e, F.where + i, r);
                        pwh += normal_lp(xp, pri(1 + i, s, r), pri(5 + nw + i, s, r));
                        if (do_generate) rec(m.PropOut, e, F.where + i, r) = x[i];
                    }
                    tril_from_scale(sc, L);
                    float qwh = -2.f * SQ_LOG_2PI, ss = 0.f;
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        float a = x[i] - loc[i];
#pragma unroll
                        for (int j = 0; j < 4; ++j) if (j < i) a -= L[i][j] * y[j];
                        y[i] = a / L[i][i];
                        ss += y[i] * y[i];
                        qwh -= logf(fabsf(L[i][i]));
                    }
                    qwh += -0.5f * ss;
                    const float ptm1 = Z(nw + 4, s, r), pres = rec(m.PropOut, e, F.pres, r);
                    const float sp = J.u_pres_prior[ni] < sigmoidf_(pri(0, s, r)) ? 1.f : 0.f;
                    const float qp = bernoulli_lp(sp, rec(m.PropOut, e, F.logit, r));
                    const float pp = bernoulli_lp(pres, pri(0, s, r));
                    const float mk = ptm1 * pres;
                    lp(LP_PQWHAT, s, r) = qw * mk;
                    lp(LP_PQWHERE, s, r) = qwh * mk;
                    lp(LP_PPWHAT, s, r) = pw * mk;
                    lp(LP_PPWHERE, s, r) = pwh * mk;
                    lp(LP_PROB, s, r) = expf(qp) * ptm1;
                    rowacc(RA_QPRES, r) += qp * ptm1;
                    rowacc(RA_PPRES, r) += pp * ptm1;
                    if (do_generate) rec(m.PropOut, e, F.pres, r) = sp;
                }
            }
        }
        c.sync();
    }

    // Generated frames discover nothing (sqair_modules.py:161-170: presence = pres_sample * 0): the discovery slots keep
    // their posterior statistics and the count terms (evaluated at the posterior's count, :145), every presence-masked
    // term vanishes.  The prior draws of what / where would only fill absent slots that never reach z_t.
    SQ_DEV void disc_generate() const {
        const Smem& m = P.sm;
        for (int i = c.tid(); i < LDS(); i += c.nthreads()) {
            const int s = i / R, r = i % R;
            rec(m.DiscOut, s + 1, P.rec.pres, r) = 0.f;
            lp(LP_DQWHAT, s, r) = 0.f; lp(LP_DQWHERE, s, r) = 0.f; lp(LP_DPWHAT, s, r) = 0.f;
        }
        c.sync();
    }

    // L = fill_triangular(cholesky_scale) * scale[:, None] + diag(scale) (modules.py:535-545).
    // TF fill_triangular order for a 10-vector x: row0 [x4], row1 [x8 x9], row2 [x7 x6 x5], row3 [x3 x2 x1 x0].
    SQ_DEV void tril_from_scale(const float (&sc)[4], float (&L)[4][4]) const {
        const int co = P.po.cholesky;
        const int idx[4][4] = {{4, -1, -1, -1}, {8, 9, -1, -1}, {7, 6, 5, -1}, {3, 2, 1, 0}};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float v = 0.f;
                if (j <= i) v = prm(co + idx[i][j]) * sc[i];
                if (i == j) v += sc[i];
                L[i][j] = v;
            }
    }

    // ------------------------------------------------------------------------------------------
    // one discovery slot (core.py:192-227) + posterior / N(0,1) prior terms (sqair_modules.py:177-186)
    // ------------------------------------------------------------------------------------------
    SQ_DEV void disc_slot(int t, int s) const {
        const Smem& m = P.sm;
        const RecF& F = P.rec;
        const int nh = P.nh, nw = P.nw, e = s + 1, ns2 = P.NS + s;
        lin(L_DRNN, s, true, s + 1);
        lin(L_DT1, 0, true, s); lin(L_DT2, 0, true, s); lin(L_DT3, 0, false, s);
        for (int r = c.tid(); r < R; r += c.nthreads()) {             // core.py:220-227
            const float so = prm(P.po.d_scale_offset);
            float wh[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                float loc = SQ_SM[m.Tp + i * R + r];
                float sc = softplusf_(SQ_SM[m.Tp + (4 + i) * R + r] + so) + 1e-2f;
                wh[i] = loc + sc * eps_where_[nidx(t, r, ns2) * 4 + i];
                rec(m.DiscOut, e, F.where + i, r) = wh[i];
                rec(m.DiscOut, e, F.where_loc + i, r) = loc;
                rec(m.DiscOut, e, F.where_scale + i, r) = sc;
            }
            set_coords(wh[0], wh[1], wh[2], wh[3], r);
        }
        // the new hidden state becomes the next slot's old one.  Done here (not at the end of the slot) so that no
        // element-wise stage reads Hrnn[1] right before the next L_DRNN writes it from a peer block.
        for (int i = c.tid(); i < nh * R; i += c.nthreads()) SQ_SM[m.Hrnn + i] = SQ_SM[m.Hrnn + nh * R + i];
        c.sync();
        extract_glimpse(false, 2 * P.NS + s);                     // discovery passes no mask (core.py:217)
        encode_glimpse(L_ENC3, 2 * P.NS + s);
        for (int i = c.tid(); i < nw * R; i += c.nthreads()) {        // core.py:216-218
            int j = i / R, r = i % R;
            float wl = SQ_SM[m.Enc + i], ws = SQ_SM[m.Enc + nw * R + i];
            rec(m.DiscOut, e, F.what + j, r) = wl + ws * eps_what_[nidx(t, r, ns2) * nw + j];
            rec(m.DiscOut, e, F.what_loc + j, r) = wl;
            rec(m.DiscOut, e, F.what_scale + j, r) = ws;
        }
        c.sync();
        lin(L_DST1, s, true); lin(L_DST2);
        for (int r = c.tid(); r < R; r += c.nthreads()) {
            const float pkm1 = rec(m.DiscOut, s, F.pres, r);      // entry 0 holds the initial 1 (core.py:153)
            float logit = pkm1 * SQ_SM[m.Lg + r] + (pkm1 - 1.f) * 88.f;
            float prob = sigmoidf_(logit);
            float pres = (u_pres_[nidx(t, r, ns2)] < prob ? 1.f : 0.f) * pkm1;
            rec(m.DiscOut, e, F.logit, r) = logit;
            rec(m.DiscOut, e, F.prob, r) = prob;
            rec(m.DiscOut, e, F.pres, r) = pres;
        }
        c.sync();
        for (int r = c.warp(); r < R; r += c.nwarps()) {
            float qw = 0.f, pw = 0.f;
            for (int j = c.lane(); j < nw; j += c.nlanes()) {
                float x = rec(m.DiscOut, e, F.what + j, r);
                qw += normal_lp(x, rec(m.DiscOut, e, F.what_loc + j, r), rec(m.DiscOut, e, F.what_scale + j, r));
                pw += normal_lp(x, 0.f, 1.f);
            }
            qw = warp_sum(qw);
            pw = warp_sum(pw);
            if (c.lane() == 0) {
                float qwh = 0.f;
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    qwh += normal_lp(rec(m.DiscOut, e, F.where + i, r), rec(m.DiscOut, e, F.where_loc + i, r),
                                     rec(m.DiscOut, e, F.where_scale + i, r));
                const float pres = rec(m.DiscOut, e, F.pres, r);
                lp(LP_DQWHAT, s, r) = qw * pres;
                lp(LP_DQWHERE, s, r) = qwh * pres;
                lp(LP_DPWHAT, s, r) = pw * pres;
                rowacc(RA_NDISC, r) += pres;
            }
        }
        c.sync();
        stash_copy(S_DISCREC, t_, e, m.DiscOut + e * R, LDE(), F.size);
    }

    // discovery priors and the number-of-steps posterior (sqair_modules.py:149-226; modules.py:548-607;
    // prior.py:61-102)
    SQ_DEV void disc_priors(int t) const {
        const Smem& m = P.sm;
        const RecF& F = P.rec;
        const int NS = P.NS;
        if (P.cfg.rec_where_prior) {
            // previous-sample inputs of the autoregressive prior: init_sample, then where_{s-1}
            for (int i = c.tid(); i < 4 * LDS(); i += c.nthreads()) {
                int f = i / LDS(), s = (i / R) % NS, r = i % R;
                SQ_SM[m.RnPrev0 + i] = (s == 0) ? prm(P.po.rn_init_sample + f) : rec(m.DiscOut, s, F.where + f, r);
            }
            c.sync();
            for (int s = 0; s < NS; ++s) stash_copy(S_RNPREV, t_, s, m.RnPrev0 + s * R, LDS(), 4);
            lin(L_RN1);
            for (int s = 0; s < NS; ++s) {
                lin(L_RN2, s, true); lin(L_RN3, 0, false, s);
                for (int r = c.tid(); r < R; r += c.nthreads()) {
                    float a = 0.f;
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        a += normal_lp(rec(m.DiscOut, s + 1, F.where + i, r), SQ_SM[m.Rns + i * R + r],
                                       SQ_SM[m.Rns + (4 + i) * R + r]);
                    lp(LP_DPWHERE, s, r) = a * rec(m.DiscOut, s + 1, F.pres, r);
                }
                c.sync();
            }
        } else {
            for (int i = c.tid(); i < LDS(); i += c.nthreads()) {
                int s = i / R, r = i % R;
                float a = 0.f;
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    a += normal_lp(rec(m.DiscOut, s + 1, F.where + k, r), P.cfg.where_mean[k], P.cfg.where_std[k]);
                lp(LP_DPWHERE, s, r) = a * rec(m.DiscOut, s + 1, F.pres, r);
            }
            c.sync();
        }
        if (P.cfg.disc_prior_type == SQAIR_DISC_PRIOR_CAT) { lin(L_SP1, 0, true); lin(L_SP2); }
        for (int r = c.tid(); r < R; r += c.nthreads()) {
            const int num = (int)rowacc(RA_NDISC, r);
            // p(N): Categorical(elu(bias + (t>0) tbias + MLP(E[n_prop]))) (sqair_modules.py:208-221)
            float pnum;
            if (P.cfg.disc_prior_type == SQAIR_DISC_PRIOR_CAT) {
                float lg[MAX_SLOTS + 1], mx = -INFINITY;
                for (int k = 0; k <= NS; ++k) {
                    float v = prm(P.po.step_prior_bias + k) + (t == 0 ? 0.f : 1.f) * prm(P.po.step_prior_tbias + k);
                    v = eluf_(v + SQ_SM[m.Spl + k * R + r]);
                    lg[k] = v;
                    mx = fmaxf(mx, v);
                }
                float se = 0.f, sel = 0.f;
                for (int k = 0; k <= NS; ++k) { se += expf(lg[k] - mx); if (k == num) sel = lg[k]; }
                pnum = sel - mx - logf(se);
            } else {                                              // tfd.Geometric(probs = 1 - success)
                const float pr = 1.f - P.cfg.step_success_prob;
                pnum = (float)num * log1pf(-pr) + logf(pr);
            }
            // q(N): modified geometric from the Bernoulli chain, float64 (prior.py:61-67,95-102)
            double pp[MAX_SLOTS], mod[MAX_SLOTS + 1], cum = 1.0, tot = 0.0;
            for (int k = 0; k < NS; ++k) pp[k] = (double)rec(m.DiscOut, k + 1, F.prob, r);
            for (int k = 0; k < NS; ++k) {
                mod[k] = (1.0 - pp[k]) * cum;
                cum *= pp[k];
                tot += mod[k];
            }
            mod[NS] = cum;
            tot += cum;
            float qsel = 0.f;
            for (int k = 0; k <= NS; ++k) {
                float v = (float)(mod[k] / tot);
                SQ_SM[m.Spl + k * R + r] = v;                      // Spl now holds disc_prob
                if (k == num) qsel = v;
            }
            rowacc(RA_QNUM, r) = logf(fminf(fmaxf(qsel, 1e-16f), 1.f));
            rowacc(RA_PNUM, r) = pnum;
        }
        c.sync();
    }

    // ------------------------------------------------------------------------------------------
    // slot compaction + object ids (sqair_modules.py:514-582; index.py:132-221)
    // ------------------------------------------------------------------------------------------
    SQ_DEV void choose_latents(int t) const {
        const Smem& m = P.sm;
        const RecF& F = P.rec;
        const int NS = P.NS, nw = P.nw, nh = P.nh;
        const sqair_outputs& o = J.out;
        for (int r = c.tid(); r < R; r += c.nthreads()) {
            // stable partition of the 2n candidates, present first
            int order[2 * MAX_SLOTS], cnt = 0;
            for (int k = 0; k < 2 * NS; ++k) {
                float p = k < NS ? rec(m.PropOut, k + 1, F.pres, r) : rec(m.DiscOut, k - NS + 1, F.pres, r);
                if (p != 0.f) order[cnt++] = k;
            }
            for (int k = 0; k < 2 * NS; ++k) {
                float p = k < NS ? rec(m.PropOut, k + 1, F.pres, r) : rec(m.DiscOut, k - NS + 1, F.pres, r);
                if (p == 0.f) order[cnt++] = k;
            }
            // ids (index.py:198-221)
            float ids[2 * MAX_SLOTS];
            float last = SQ_SM[m.LastId + r], inc = 0.f;
            for (int k = 0; k < NS; ++k) {
                float pp = rec(m.PropOut, k + 1, F.pres, r);
                ids[k] = SQ_SM[m.Ids + k * R + r] * pp - (1.f - pp);
            }
            for (int k = 0; k < NS; ++k) {
                float dp = rec(m.DiscOut, k + 1, F.pres, r);
                inc += dp;
                ids[NS + k] = (inc + last) * dp - (1.f - dp);
            }
            SQ_SM[m.LastId + r] = last + inc;
            float nsteps = 0.f;
            for (int j = 0; j < NS; ++j) {
                SQ_SM[m.Perm + j * R + r] = (float)order[j];
                SQ_SM[m.Ids + j * R + r] = ids[order[j]];
                int k = order[j];
                nsteps += k < NS ? rec(m.PropOut, k + 1, F.pres, r) : rec(m.DiscOut, k - NS + 1, F.pres, r);
            }
            rowacc(7, r) = nsteps;
        }
        c.sync();
        stash_copy(S_PERM, t_, 0, m.Perm, R, NS);
        // new z_t and the 9 compacted heads
        const size_t trow = (size_t)t * P.rows;
        for (int i = c.tid(); i < F.size * LDS(); i += c.nthreads()) {
            const int r = i % R, j = (i / R) % NS, f = i / LDS();
            const int k = (int)SQ_SM[m.Perm + j * R + r];
            const float v = k < NS ? rec(m.PropOut, k + 1, f, r) : rec(m.DiscOut, k - NS + 1, f, r);
            float* dst = nullptr;
            int ff = 0, width = 1;
            if (f < F.where) { Z(f, j, r) = v; dst = o.what; ff = f; width = nw; }
            else if (f < F.pres) { Z(f, j, r) = v; dst = o.where; ff = f - F.where; width = 4; }
            else if (f == F.pres) { Z(nw + 4, j, r) = v; dst = o.presence; }
            else if (f < F.what_scale) { dst = o.what_loc; ff = f - F.what_loc; width = nw; }
            else if (f < F.where_loc) { dst = o.what_scale; ff = f - F.what_scale; width = nw; }
            else if (f < F.where_scale) { dst = o.where_loc; ff = f - F.where_loc; width = 4; }
            else if (f < F.prob) { dst = o.where_scale; ff = f - F.where_scale; width = 4; }
            else if (f == F.prob) { dst = o.presence_prob; }
            else { Z(nw + 5, j, r) = v; dst = o.presence_logit; }
            if (dst && c.rank() == 0 && valid_row(r)) dst[((trow + grow_of(r)) * NS + j) * width + ff] = v;
        }
        // GRU states travel with their slots; discovered objects start from the trainable initial states
        for (int i = c.tid(); i < nh * R; i += c.nthreads()) {
            const int f = i / R, r = i % R;
            float tv[MAX_SLOTS], pv[MAX_SLOTS];
            const float t0 = prm(P.po.temporal_h0 + f), p0 = prm(P.po.prior_h0 + f);
            for (int j = 0; j < NS; ++j) {
                const int k = (int)SQ_SM[m.Perm + j * R + r];
                tv[j] = k < NS ? SQ_SM[m.Tst + f * LDS() + k * R + r] : t0;
                pv[j] = k < NS ? SQ_SM[m.Pst + f * LDS() + k * R + r] : p0;
            }
            for (int j = 0; j < NS; ++j) {
                SQ_SM[m.Tst + f * LDS() + j * R + r] = tv[j];
                SQ_SM[m.Pst + f * LDS() + j * R + r] = pv[j];
            }
        }
        c.sync();
        for (int j = 0; j < NS; ++j) stash_copy(S_Z, t + 1, j, m.Z + j * R, LDS(), nw + 6);
    }
    // local row r of this block: is it a real row, and its global index (clamped: padding rows recompute the last row).
    // Plain arithmetic on purpose: member arrays indexed with a run-time r would pin the whole object to local memory.
    SQ_DEV bool valid_row(int r) const { return row0 + r < P.rows; }
    SQ_DEV int grow_of(int r) const { return row0 + r < P.rows ? row0 + r : P.rows - 1; }

    // ------------------------------------------------------------------------------------------
    // decoder + canvas + pixel likelihood (modules.py:435-467; seq.py:271-273)
    // ------------------------------------------------------------------------------------------
    SQ_DEV void decode_and_score(int t) const {
        const Smem& m = P.sm;
        const int NS = P.NS, nw = P.nw, G = P.cfg.G, W = P.cfg.W, H = P.cfg.H, g = P.g, PX = P.PX;
        const sqair_outputs& o = J.out;
        const size_t trow = (size_t)t * P.rows;
        for (int s = 0; s < NS; ++s) { lin(L_DEC1, s, true); lin(L_DEC2, 0, true, s); lin(L_DEC3, s, s + 1 < NS); }
        if (o.glimpse && c.rank() == 0)
            for (int i = c.tid(); i < R * NS * g; i += c.nthreads()) {
                const int px = i % g, s = (i / g) % NS, r = i / (g * NS);
                if (valid_row(r)) o.glimpse[((trow + grow_of(r)) * NS + s) * g + px] = SQ_SM[m.Dgl + px * LDS() + s * R + r];
            }
        // inverse-transformer coordinates of every slot (Pri is dead by now and reused as a table)
        float* cc = SQ_SM + m.Pri;    // [4][NS][R]
        for (int i = c.tid(); i < LDS(); i += c.nthreads()) {
            const int s = i / R, r = i % R;
            cc[0 * LDS() + i] = fmaxf(sigmoidf_(Z(nw + 0, s, r)), 1e-4f);
            cc[1 * LDS() + i] = fmaxf(sigmoidf_(Z(nw + 1, s, r)), 1e-4f);
            cc[2 * LDS() + i] = tanhf(Z(nw + 2, s, r));
            cc[3 * LDS() + i] = tanhf(Z(nw + 3, s, r));
        }
        c.sync();
        const float hg = 0.5f * (float)(G - 1);
        // p(x|z) stds as the reference builds them: (sqrt(std))^2 (modules.py:419-422)
        const float sf0 = sqrtf(P.cfg.output_std), sb0 = sqrtf(P.cfg.bg_std);
        const float sf = sf0 * sf0, sb = sb0 * sb0;
        float ll[R];
#pragma unroll
        for (int r = 0; r < R; ++r) ll[r] = 0.f;
        // every block of the cluster composes its share of the pixels
        const int px_per = (PX + c.ncta() - 1) / c.ncta();
        const int px0 = c.rank() * px_per, px1 = (px0 + px_per < PX) ? (px0 + px_per) : PX;
#pragma unroll
        for (int r = 0; r < R; ++r) {
            for (int px = px0 + c.tid(); px < ((dbg_ & 4) ? 0 : px1); px += c.nthreads()) {
                const int iy = px / W, ix = px % W;
                const float u = lin11(ix, W), v = lin11(iy, H);
                float canvas = 0.f, nz = 0.f;
                for (int s = 0; s < NS; ++s) {
                    const float pres = Z(nw + 4, s, r);
                    if (pres == 0.f) continue;
                    const float sx = cc[0 * LDS() + s * R + r], sy = cc[1 * LDS() + s * R + r];
                    const float tx = cc[2 * LDS() + s * R + r], ty = cc[3 * LDS() + s * R + r];
                    const float xg = hg * ((u - tx) / sx) + hg;
                    const float yg = hg * ((v - ty) / sy) + hg;
                    const float* gl = SQ_SM + m.Dgl + s * R + r;
                    const int lds = LDS();
                    canvas += pres * bilinear_zero_pad(xg, yg, G, G, [&](int gx, int gy) { return gl[(gy * G + gx) * lds]; });
                    nz += pres * bilinear_zero_pad(xg, yg, G, G, [&](int, int) { return 1.f; });
                }
                const float mask = sigmoidf_(-10.f + nz * 20.f);                    // modules.py:462
                canvas += prm(P.po.mean_img + px) * mask;                           // modules.py:465
                const float std = mask * sf + (1.f - mask) * sb;                    // modules.py:453
                ll[r] += normal_lp(imgrow[r][px], canvas, std);
                if (o.canvas && valid_row(r)) o.canvas[(trow + grow_of(r)) * PX + px] = canvas;
            }
        }
        // block reduction of the R partial sums, then the cluster's partial sums meet in every block
        float* red = SQ_SM + m.Red;
#pragma unroll
        for (int r = 0; r < R; ++r) {
            float v = warp_sum(ll[r]);
            if (c.lane() == 0) red[c.warp() * R + r] = v;
        }
        c.sync();
        for (int r = c.tid(); r < R; r += c.nthreads()) {
            float a = 0.f;
            for (int w = 0; w < c.nwarps(); ++w) a += red[w * R + r];
            if (c.ncta() > 1) {
                for (int q = 0; q < c.ncta(); ++q) store_peer(c, m.RowAcc + (8 + c.rank()) * R + r, q, a);
            } else {
                rowacc(RA_LL, r) = a;
            }
        }
        if (c.ncta() > 1) {
            cluster_arrive(c);
            cluster_wait(c);
            for (int r = c.tid(); r < R; r += c.nthreads()) {
                float a = 0.f;
                for (int q = 0; q < c.ncta(); ++q) a += rowacc(8 + q, r);
                rowacc(RA_LL, r) = a;
            }
        }
        c.sync();
    }

    // per-frame scalars and the remaining outputs (seq.py:215-257,271-276; sqair_modules.py:473-488)
    SQ_DEV void write_frame_outputs(int t) const {
        const Smem& m = P.sm;
        const RecF& F = P.rec;
        const int NS = P.NS;
        const sqair_outputs& o = J.out;
        const size_t trow = (size_t)t * P.rows;
        if (c.rank() != 0 || (dbg_ & 32)) { c.sync(); return; }                  // replicas hold identical values
        for (int i = c.tid(); i < LDS(); i += c.nthreads()) {
            const int s = i / R, r = i % R;
            if (!valid_row(r)) continue;
            const size_t b = (trow + grow_of(r)) * NS + s;
            if (o.obj_id) o.obj_id[b] = SQ_SM[m.Ids + s * R + r];
            if (o.disc_what_log_prob) o.disc_what_log_prob[b] = lp(LP_DQWHAT, s, r);
            if (o.disc_where_log_prob) o.disc_where_log_prob[b] = lp(LP_DQWHERE, s, r);
            if (o.disc_what_prior_log_prob) o.disc_what_prior_log_prob[b] = lp(LP_DPWHAT, s, r);
            if (o.disc_where_prior_log_prob) o.disc_where_prior_log_prob[b] = lp(LP_DPWHERE, s, r);
            if (o.prop_what_log_prob) o.prop_what_log_prob[b] = lp(LP_PQWHAT, s, r);
            if (o.prop_where_log_prob) o.prop_where_log_prob[b] = lp(LP_PQWHERE, s, r);
            if (o.prop_what_prior_log_prob) o.prop_what_prior_log_prob[b] = lp(LP_PPWHAT, s, r);
            if (o.prop_where_prior_log_prob) o.prop_where_prior_log_prob[b] = lp(LP_PPWHERE, s, r);
            if (o.prop_prob) o.prop_prob[b] = lp(LP_PROB, s, r);
            if (o.prop_pres) o.prop_pres[b] = rec(m.PropOut, s + 1, F.pres, r);
            if (o.disc_pres) o.disc_pres[b] = rec(m.DiscOut, s + 1, F.pres, r);
        }
        if (o.disc_prob)
            for (int i = c.tid(); i < (NS + 1) * R; i += c.nthreads()) {
                const int k = i / R, r = i % R;
                if (valid_row(r)) o.disc_prob[(trow + grow_of(r)) * (NS + 1) + k] = SQ_SM[m.Spl + i];
            }
        for (int r = c.tid(); r < R; r += c.nthreads()) {
            if (!valid_row(r)) continue;
            // q, p in the reference's summation order: sum_s(what_s + where_s) + discrete term, disc + prop
            float pq = 0.f, pp = 0.f, dq = 0.f, dp = 0.f;
            for (int s = 0; s < NS; ++s) {
                pq += lp(LP_PQWHAT, s, r) + lp(LP_PQWHERE, s, r);
                pp += lp(LP_PPWHAT, s, r) + lp(LP_PPWHERE, s, r);
                dq += lp(LP_DQWHAT, s, r) + lp(LP_DQWHERE, s, r);
                dp += lp(LP_DPWHAT, s, r) + lp(LP_DPWHERE, s, r);
            }
            const float prop_lp = rowacc(RA_QPRES, r), prop_plp = rowacc(RA_PPRES, r);
            const float qnum = rowacc(RA_QNUM, r), pnum = rowacc(RA_PNUM, r);
            const float q = (dq + qnum) + (pq + prop_lp);
            const float p = (dp + pnum) + (pp + prop_plp);
            const float ll = rowacc(RA_LL, r), kl = q - p;
            const size_t b = trow + grow_of(r);
            if (o.step_log_prob) o.step_log_prob[b] = prop_lp + qnum;
            if (o.discrete_log_prob) o.discrete_log_prob[b] = prop_lp + qnum;
            if (o.disc_log_prob) o.disc_log_prob[b] = qnum;
            if (o.disc_prior_log_prob) o.disc_prior_log_prob[b] = pnum;
            if (o.prop_log_prob) o.prop_log_prob[b] = prop_lp;
            if (o.prop_prior_log_prob) o.prop_prior_log_prob[b] = prop_plp;
            if (o.num_prop_steps_per_sample) o.num_prop_steps_per_sample[b] = rowacc(RA_NPROP, r);
            if (o.num_disc_steps_per_sample) o.num_disc_steps_per_sample[b] = rowacc(RA_NDISC, r);
            if (o.num_steps_per_sample) o.num_steps_per_sample[b] = rowacc(7, r);
            if (o.data_ll_per_sample) o.data_ll_per_sample[b] = ll;
            if (o.kl_per_sample) o.kl_per_sample[b] = kl;
            if (o.log_q_z_given_x_per_sample) o.log_q_z_given_x_per_sample[b] = q;
            if (o.log_p_z_per_sample) o.log_p_z_per_sample[b] = p;
            if (o.log_weights_per_timestep) o.log_weights_per_timestep[b] = ll - kl;
        }
        c.sync();
    }

    // ------------------------------------------------------------------------------------------
    // one frame (seq.py:181-269 / sqair_modules.py:446-512)
    // ------------------------------------------------------------------------------------------
    SQ_DEV void frame(int t) {
        const Smem& m = P.sm;
        const int NS = P.NS, nh = P.nh;
#pragma unroll
        for (int r = 0; r < R; ++r) imgrow[r] = obs_ + ((size_t)t * P.cfg.B + grow_of(r) / P.cfg.K) * P.PX;
#ifndef SQAIR_HOST_EMU
        if (m.img_n > 0) {
            // TMA: one bulk copy per sequence of this block's rows, completion on an mbarrier whose phase is the frame
            // index; the first glimpse extraction of the frame waits for it (a few dense layers later).  Every thread is
            // past the barrier that ended the previous frame, so nobody still reads the old frame.
            const int b0 = grow_of(0) / P.cfg.K, nimg = grow_of(R - 1) / P.cfg.K - b0 + 1;
            if (c.tid() == 0) {
                const uint32_t bar = smem_u32(SQ_SM + m.ImgBar);
                mbar_expect_tx(bar, (uint32_t)(nimg * P.PX) * 4u);
                for (int i = 0; i < nimg; ++i)
                    bulk_g2s(smem_u32(SQ_SM + m.Img + i * P.PX), obs_ + ((size_t)t * P.cfg.B + b0 + i) * P.PX, (uint32_t)P.PX * 4u, bar);
            }
#pragma unroll
            for (int r = 0; r < R; ++r) imgrow[r] = SQ_SM + m.Img + (grow_of(r) / P.cfg.K - b0) * P.PX;
            frame_parity = (uint32_t)(t & 1);
        }
#endif
        {
            const int g = c.lane() >> 2;
            img_g = imgrow[R - 1];
#pragma unroll
            for (int r = 0; r < R - 1; ++r) if (g == r) img_g = imgrow[r];
        }
        for (int i = c.tid(); i < 16 * R; i += c.nthreads()) SQ_SM[m.RowAcc + i] = 0.f;
        for (int i = c.tid(); i < nh * R; i += c.nthreads()) {
            SQ_SM[m.Hrnn + i] = prm(P.po.prop_h0 + i / R);          // propagate.py:170 / core.py:130
            SQ_SM[m.DIn + nh * R + i] = 0.f;                         // conditioning accumulator
        }
        c.sync();
        t_ = t;
        if (TR && stash_ != nullptr) {                            // state entering the frame, initial slot records
            for (int s = 0; s < NS; ++s) {
                stash_copy(S_TST, t, s, m.Tst + s * R, LDS(), nh);
                stash_copy(S_PST, t, s, m.Pst + s * R, LDS(), nh);
            }
            stash_copy(S_PH, t, 0, m.Hrnn, R, nh);
            stash_copy(S_PROPREC, t, 0, m.PropOut, LDE(), P.rec.size);
            stash_copy(S_DISCREC, t, 0, m.DiscOut, LDE(), P.rec.size);
        }
        for (int s = 0; s < NS; ++s) prop_slot(t, s);
        const bool do_generate = GEN && J.generate_after > 0 && t > J.generate_after;          // seq.py:198-203
        if (GEN) prop_generate(t, do_generate);
        // latent summary: sum_s pres_s * MLP([what_s, where_s]) (sqair_modules.py:368-385,501)
        for (int s = 0; s < NS; ++s) {
            lin(L_LAT1, s, true); lin(L_LAT2, 0, false, s);
            for (int i = c.tid(); i < nh * R; i += c.nthreads())
                SQ_SM[m.DIn + nh * R + i] += SQ_SM[m.A1 + i] * rec(m.PropOut, s + 1, P.rec.pres, i % R);
            c.sync();
        }
        for (int r = c.tid(); r < R; r += c.nthreads()) {               // sqair_modules.py:505-507
            float a = 0.f;
            for (int s = 0; s < NS; ++s) a += (sigmoidf_(pri(0, s, r)) - 0.5f) / (float)NS;
            SQ_SM[m.Exp + r] = a;
        }
        stash_copy(S_DIN, t, 0, m.DIn + nh * R, R, nh, nh);          // conditioning (the image encoding follows below)
        lin(L_IMG1, 0, true); lin(L_IMG2);                           // core.py:165, hoisted out of the slot loop
        for (int i = c.tid(); i < nh * R; i += c.nthreads()) SQ_SM[m.Hrnn + i] = prm(P.po.disc_h0 + i / R);
        c.sync();
        stash_copy(S_EXP, t, 0, m.Exp, R, 1);
        stash_copy(S_DH, t, 0, m.Hrnn, R, nh);
        for (int s = 0; s < NS; ++s) disc_slot(t, s);
        if (do_generate) disc_generate();
        disc_priors(t);
        choose_latents(t);
        decode_and_score(t);
        write_frame_outputs(t);
    }

    SQ_DEV void run() {
        calls_init(c, ltab_);
        init_sequence();
        for (int t = 0; t < P.cfg.T; ++t) frame(t);
    }
};
#ifndef SQAIR_HOST_EMU
#undef P
#endif

}  // namespace sq
