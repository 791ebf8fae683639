// sqair_api.cu -- CUDA kernels (sm_100a) and the C ABI of include/sqair_b200.h.
//
// The hot path is ONE persistent kernel per call: `sqair_sequence_kernel<R>` runs the whole
// T-frame Discover/Propagate recursion for R rows per thread block with all recurrent state in
// shared memory and every dense layer on the tensor cores (sqair_device.cuh).  Small auxiliary kernels: parameter packing, counter-based
// noise, the particle objective, and stand-alone entry points for the two bandwidth-shaped ops
// (glimpse sampler, canvas compose + likelihood).
#include <cuda_runtime.h>
#include <math.h>
#include <stdlib.h>
#include <memory>
#include <mutex>
#include <string>
#include <utility>
#include <vector>

#include "sqair_device.cuh"
#include "sqair_internal.h"

using namespace sq;
using sqi::Shape;
using sqi::choose_shape;
using sqi::env_int;
using sqi::fail;
using sqi::cuda_fail;

static thread_local std::string g_err;

int sqi::fail(int code, const std::string& msg) {
    g_err = msg;
    return code;
}
int sqi::cuda_fail(cudaError_t e, const char* what) {
    return sqi::fail(SQAIR_ECUDA, std::string(what) + ": " + cudaGetErrorString(e));
}

// ---------------------------------------------------------------------------------------------
// persistent sequence kernel: one cluster of C blocks per R rows, the whole T-frame recursion
// ---------------------------------------------------------------------------------------------
template <int R, bool TR, bool GEN>
__global__ void __launch_bounds__(NT_LAUNCH) sqair_sequence_kernel(const __grid_constant__ Job job) {
    const PlanHdr& plan = c_plan;
    Ctx c;
#ifdef SQAIR_PROFILE
    for (int i = 0; i < 8; ++i) c.prof[i] = 0;
    c.t_last = clock64();
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        for (int i = 0; i < L_COUNT; ++i) g_trace[0][i] = g_trace[1][i] = g_trace[2][i] = 0;
        g_trace_last = clock64();
    }
#endif
    // the packed parameters carry a layout header; a buffer packed for another cluster size / model must not be read
    {
        const uint32_t* h = reinterpret_cast<const uint32_t*>(job.prm + plan.phdr_off);
        if (h[0] != PACK_MAGIC || h[1] != (uint32_t)plan.C || h[2] != (uint32_t)plan.NS || h[3] != (uint32_t)plan.PX) __trap();
    }
    // no block writes into a peer's shared memory before every block of the cluster has started (racecheck: "address is
    // located in a block that might not have entered yet" on the first exchange)
    if (plan.C > 1) { cluster_arrive(c); cluster_wait(c); }
    Block<R, TR, GEN> blk(c, job, (int)(blockIdx.x / plan.C) * R);
    blk.run();
    if (plan.C > 1) { cluster_arrive(c); cluster_wait(c); }      // no block exits while peers may still write to it
#ifdef SQAIR_PROFILE
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        long long tin = 0, tgap = 0, n = 0;
        for (int i = 0; i < L_COUNT; ++i) {
            if (g_trace[2][i] == 0) continue;
            printf("[trace] layer %2d calls %5lld  in-dense %8.0f cyc/call  gap-before %8.0f cyc/call\n", i, g_trace[2][i],
                   (double)g_trace[0][i] / g_trace[2][i], (double)g_trace[1][i] / g_trace[2][i]);
            tin += g_trace[0][i]; tgap += g_trace[1][i]; n += g_trace[2][i];
        }
        printf("[trace] total calls %lld  in-dense %lld cyc  gaps %lld cyc\n", n, tin, tgap);
    }
    if (blockIdx.x == 0 && (threadIdx.x == 0 || threadIdx.x == 32 * 11)) {
        printf("[profile tid %d] cycles: call-prologue %lld  mma-units %lld (%lld units, %lld k-steps)  partials+sync %lld  finish+stores %lld  barrierB %lld  between-dense %lld\n",
               (int)threadIdx.x, c.prof[0], c.prof[1], c.prof[7], c.prof[6], c.prof[2], c.prof[3], c.prof[4], c.prof[5]);
    }
#endif
}

static const int kRowChoices[] = {1, 2, 3, 4, 5, 6};
static const int kClusterChoices[] = {1, 2, 3, 4, 5, 6, 7, 8};
static const int kSmemLimit = 232448;    // 227 KB opt-in shared memory per block on sm_100

// The plan header lives in __constant__ memory (descriptor reads are constant-bank loads); every device has its own
// copy of the symbol, so the "what is uploaded" state is kept per device.  The header is re-uploaded only when it
// changes, after every launch that may still be reading the old one has finished (one event per stream that launched
// since the last upload).
struct DevicePlanState {
    PlanHdr uploaded;
    bool valid = false;
    std::vector<std::pair<cudaStream_t, cudaEvent_t>> inflight;
};
static std::mutex g_plan_mutex;
static DevicePlanState g_plan_state[64];

static bool stream_is_capturing(cudaStream_t st) {
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(st, &cs) != cudaSuccess) { cudaGetLastError(); return false; }
    return cs != cudaStreamCaptureStatusNone;
}

static int upload_plan(const Plan& plan, cudaStream_t st, DevicePlanState& ds) {
    const PlanHdr& hdr = plan;
    if (ds.valid && memcmp(&ds.uploaded, &hdr, sizeof(PlanHdr)) == 0) return SQAIR_OK;
    if (stream_is_capturing(st))
        return fail(SQAIR_EINVAL, "the launch plan of this configuration is not resident: run the call once before capturing it in a CUDA graph");
    for (auto& se : ds.inflight) CUDA_TRY(cudaEventSynchronize(se.second));
    ds.valid = false;
    ds.uploaded = hdr;                          // the copy source must outlive the (possibly staged) transfer
    CUDA_TRY(cudaMemcpyToSymbolAsync(c_plan, &ds.uploaded, sizeof(PlanHdr), 0, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    ds.valid = true;
    return SQAIR_OK;
}

// Layer table of a launch shape (L_COUNT x DESC_WORDS words, staged into shared memory one call ahead by the kernel).
// It depends on everything the plan depends on (rows per cluster, frame staging, stash offsets), so it is NOT part of
// the packed parameters: the library keeps one small device copy per (device, table contents), for the process
// lifetime -- the only device memory it ever allocates.
struct DevTable {
    int device;
    uint64_t hash;
    std::vector<int32_t> host;
    float* dev;
};
static std::mutex g_ltab_mutex;
static std::vector<DevTable*> g_ltabs;

static int get_layer_table(const Plan& plan, int device, cudaStream_t st, const float** out) {
    std::vector<int32_t> tab((size_t)L_COUNT * DESC_WORDS, 0);
    for (int i = 0; i < L_COUNT; ++i) memcpy(&tab[(size_t)i * DESC_WORDS], &plan.L[i], sizeof(Layer));
    uint64_t h = 1469598103934665603ull;
    for (int32_t w : tab) { h ^= (uint32_t)w; h *= 1099511628211ull; }
    std::lock_guard<std::mutex> lock(g_ltab_mutex);
    for (DevTable* t : g_ltabs)
        if (t->device == device && t->hash == h && t->host == tab) { *out = t->dev; return SQAIR_OK; }
    if (stream_is_capturing(st))
        return fail(SQAIR_EINVAL, "the layer table of this configuration is not resident: run the call once before capturing it in a CUDA graph");
    DevTable* t = new DevTable{device, h, std::move(tab), nullptr};
    CUDA_TRY(cudaMalloc((void**)&t->dev, t->host.size() * sizeof(int32_t)));
    CUDA_TRY(cudaMemcpyAsync(t->dev, t->host.data(), t->host.size() * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    g_ltabs.push_back(t);
    *out = t->dev;
    return SQAIR_OK;
}

template <int R, bool TR, bool GEN>
static int launch_sequence_impl(const Plan& plan, Job job, cudaStream_t st) {
    int device = 0;
    CUDA_TRY(cudaGetDevice(&device));
    if (device < 0 || device >= 64) return fail(SQAIR_EUNSUPPORTED, "device ordinal out of range");
    int rc = get_layer_table(plan, device, st, &job.ltab);
    if (rc != SQAIR_OK) return rc;
    std::lock_guard<std::mutex> lock(g_plan_mutex);
    DevicePlanState& ds = g_plan_state[device];
    rc = upload_plan(plan, st, ds);
    if (rc != SQAIR_OK) return rc;
    const int smem_bytes = plan.sm.total * (int)sizeof(float);
    CUDA_TRY(cudaFuncSetAttribute(sqair_sequence_kernel<R, TR, GEN>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    const int ncl = (plan.rows + R - 1) / R;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(ncl * plan.C);
    cfg.blockDim = dim3(NT_LAUNCH);
    cfg.dynamicSmemBytes = smem_bytes;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = plan.C;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = plan.C > 1 ? 1 : 0;
    CUDA_TRY(cudaLaunchKernelEx(&cfg, sqair_sequence_kernel<R, TR, GEN>, job));
    if (!stream_is_capturing(st)) {           // (a captured launch replays with the plan that was resident at capture time)
        cudaEvent_t ev = nullptr;
        for (auto& se : ds.inflight) if (se.first == st) ev = se.second;
        if (!ev) {
            if (ds.inflight.size() >= 32) {
                for (auto& se : ds.inflight) { cudaEventSynchronize(se.second); cudaEventDestroy(se.second); }
                ds.inflight.clear();
            }
            CUDA_TRY(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
            ds.inflight.emplace_back(st, ev);
        }
        CUDA_TRY(cudaEventRecord(ev, st));
    }
    return SQAIR_OK;
}

// inference (no stash: the stash code is compiled out), training, or generation (sample_from_prior) instantiation
template <int R>
static int launch_sequence(const Plan& plan, const Job& job, cudaStream_t st) {
    if (job.eps_where_prior) return launch_sequence_impl<R, false, true>(plan, job, st);
    if (env_int("SQAIR_FORCE_TRAIN_KERNEL")) return launch_sequence_impl<R, true, false>(plan, job, st);      // tuning: the training instantiation without a stash
    return job.stash ? launch_sequence_impl<R, true, false>(plan, job, st) : launch_sequence_impl<R, false, false>(plan, job, st);
}

// Launch shape.  R rows per cluster (more rows = fewer re-reads of the weights from L2), C blocks per
// cluster (more blocks = less weight traffic and math per SM, more exchange).  Default: the (R, C) with
// the lowest modelled frame time among those that fit shared memory and run as a single wave on 148 SMs;
// SQAIR_ROWS_PER_CTA / SQAIR_CLUSTER override (tuning sweeps).  Packing depends on C only.
int sqi::env_int(const char* name) {
    const char* e = getenv(name);
    return e ? atoi(e) : 0;
}

static std::string choose_shape_uncached(const sqair_cfg& c, const std::vector<ParamEntry>& tab, Shape& out);

// choose_shape is called by every entry point; its result only depends on the configuration and the tuning
// environment variables, so the last few results are cached (building ~25 candidate plans costs ~1 ms of host time).
struct ShapeKey {
    sqair_cfg cfg;
    int env[5];
    bool operator==(const ShapeKey& o) const { return memcmp(this, &o, sizeof(ShapeKey)) == 0; }
};
static std::mutex g_shape_mutex;
static std::vector<std::pair<ShapeKey, std::shared_ptr<Shape>>> g_shape_cache;

std::string sqi::choose_shape(const sqair_cfg& c, const std::vector<ParamEntry>& tab, Shape& out) {
    ShapeKey key;
    memset(&key, 0, sizeof(key));
    key.cfg = c;
    key.env[0] = env_int("SQAIR_ROWS_PER_CTA"); key.env[1] = env_int("SQAIR_CLUSTER"); key.env[2] = env_int("SQAIR_NO_FRAME_STAGING");
    key.env[3] = env_int("SQAIR_UNIT_COST"); key.env[4] = env_int("SQAIR_SLICE_COST");
    std::lock_guard<std::mutex> lock(g_shape_mutex);
    for (auto& kv : g_shape_cache)
        if (kv.first == key) { out = *kv.second; return ""; }
    std::string e = choose_shape_uncached(c, tab, out);
    if (!e.empty()) return e;
    if (g_shape_cache.size() >= 16) g_shape_cache.erase(g_shape_cache.begin());
    g_shape_cache.emplace_back(key, std::make_shared<Shape>(out));
    return "";
}

// Blocks of a cluster of size C that can be resident at once.  Asked of the device the call runs on
// (cudaOccupancyMaxActiveClusters for the sequence kernel at its largest shared-memory footprint: GPCs hold 16-18 SMs, so
// odd cluster sizes leave SMs unused); the table is what a 148-SM B200 answers and serves hosts without a device
// (sqair_query_sizes in CPU-only tooling).
static int max_resident_blocks(int C) {
    static const int tab[9] = {0, 148, 148, 132, 132, 120, 120, 112, 128};
    static std::mutex mu;
    static int cached[64][9];
    int device = 0;
    if (cudaGetDevice(&device) != cudaSuccess || device < 0 || device >= 64) { cudaGetLastError(); return tab[C]; }
    std::lock_guard<std::mutex> lock(mu);
    if (cached[device][C] == 0) {
        int n = 0;
        cudaLaunchConfig_t cfg;
        memset(&cfg, 0, sizeof(cfg));
        cfg.blockDim = dim3(NT_LAUNCH);
        cfg.dynamicSmemBytes = 220 * 1024;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = C; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        cfg.gridDim = dim3(C * 64);
        cudaError_t e = cudaFuncSetAttribute(sqair_sequence_kernel<1, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
        if (e == cudaSuccess) e = cudaOccupancyMaxActiveClusters(&n, sqair_sequence_kernel<1, false, false>, &cfg);
        if (e != cudaSuccess || n <= 0) { cudaGetLastError(); cached[device][C] = tab[C]; }
        else cached[device][C] = n * C;
    }
    return cached[device][C];
}

static std::string choose_shape_uncached(const sqair_cfg& c, const std::vector<ParamEntry>& tab, Shape& out) {
    const int rows = c.B * c.K;
    const int fR = env_int("SQAIR_ROWS_PER_CTA"), fC = env_int("SQAIR_CLUSTER");
    // Cost model (B200 measurements in profiles/): a dense call has a fixed cost (barriers, descriptor, first-load
    // latency) and streams its weight panel through the tensor-core loop at `stream_bw` per block; a frame is
    // `calls` dense calls and touches W bytes of weights (the glimpse encoder runs three times per slot).
    const double W = 11.1e6 * c.n * 4.0, stream_bw = 60e9, fixed = 1.5e-6;
    const double calls = 38.0 * c.n + 1.0;
    double best = 1e30;
    std::string err = "configuration does not fit shared memory";
    out.R = 0;
    for (int C : kClusterChoices) {
        if (fC && C != fC) continue;
        for (int R : kRowChoices) {
            if (fR && R != fR) continue;
            if (R > rows && R != 1) continue;
            Shape s;
            bool ok = false;
            for (int stage = env_int("SQAIR_NO_FRAME_STAGING") ? 0 : 1; stage >= 0 && !ok; --stage) {   // frames in shared memory if they fit
                // SQAIR_UNIT_COST / SQAIR_SLICE_COST (tenths of a k-step): k-slicing cost model overrides for tuning sweeps
                const double uc = env_int("SQAIR_UNIT_COST") ? env_int("SQAIR_UNIT_COST") * 0.1 : 3.0;
                const double sc = env_int("SQAIR_SLICE_COST") ? env_int("SQAIR_SLICE_COST") * 0.1 : 0.5;
                std::string e = build_plan(c, R, C, s.plan, tab, s.pieces, &s.packed_total, stage != 0, uc, sc);
                if (!e.empty()) { err = e; break; }
                ok = s.plan.sm.total * (int)sizeof(float) <= kSmemLimit;
            }
            if (!ok) continue;
            const int ncl = (rows + R - 1) / R;
            const int max_blocks = max_resident_blocks(C);
            const double waves = (double)((ncl * C + max_blocks - 1) / max_blocks);
            const double t = waves * (calls * fixed + W / C / stream_bw + (C > 1 ? calls * 0.4e-6 : 0.0));
            if (t < best) { best = t; s.R = R; s.C = C; out = s; }
        }
    }
    if (out.R == 0) return (fR || fC) ? "SQAIR_ROWS_PER_CTA / SQAIR_CLUSTER value unsupported or does not fit shared memory" : err;
    return "";
}

// The packed layout depends on the launch shape (cluster size).  sqair_pack_params records the
// layout tag of every buffer it fills; sqair_forward refuses a buffer that was packed for a different shape.
struct LayoutTag {
    int C;
    int64_t packed_total;
    bool operator==(const LayoutTag& o) const { return C == o.C && packed_total == o.packed_total; }
};
static std::mutex g_tag_mutex;
static std::vector<std::pair<const void*, LayoutTag>> g_tags;
static LayoutTag tag_of(const Shape& sh) { return LayoutTag{sh.C, sh.packed_total}; }

// ---------------------------------------------------------------------------------------------
// auxiliary kernels
// ---------------------------------------------------------------------------------------------
struct PackTab {
    int n;
    int src[128], dst[128], cnt[128];
};

__global__ void pack_kernel(const __grid_constant__ PackTab tab, const float* __restrict__ src, float* __restrict__ dst,
                            int total) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        int lo = 0, hi = tab.n - 1;
        while (lo < hi) {
            int mid = (lo + hi + 1) >> 1;
            if (tab.src[mid] <= i) lo = mid; else hi = mid - 1;
        }
        dst[tab.dst[lo] + (i - tab.src[lo])] = src[i];
    }
}

// canonical matrices -> per-layer column panels in MMA fragment order (sqair_core.h: frag_off)
struct PieceDev {
    int vrow0, vcol0, K, N, src_off, src_ld, w_off, ksteps, Nc, panel_floats;
};
struct PieceTab {
    int n;
    PieceDev p[160];
};

__global__ void pack_panels_kernel(const __grid_constant__ PieceTab tab, const float* __restrict__ src,
                                   float* __restrict__ dst) {
    const PieceDev& pc = tab.p[blockIdx.y];
    const int total = pc.K * pc.N;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int k = i / pc.N, n = i - k * pc.N;
        const int vrow = pc.vrow0 + k, vcol = pc.vcol0 + n, panel = vcol / pc.Nc;
        // accumulate (the buffer was zeroed): a bias row may be the sum of two bias vectors
        const int cc = vcol - panel * pc.Nc;
        atomicAdd(&dst[(size_t)pc.w_off + (size_t)panel * pc.panel_floats + frag_off(pc.ksteps, cc >> 4, vrow >> 3, cc & 15, vrow & 7)],
                  src[(size_t)pc.src_off + (size_t)k * pc.src_ld + n]);
    }
}

// plain fragment order (fp32 sums of the packed pieces) -> (hi, lo) fragment order read by the sequence kernel
struct SplitTab {
    int n;
    int w1_off[L_COUNT], w_off[L_COUNT], floats[L_COUNT];      // floats = npanel * panel_floats / 2
};
__global__ void split_panels_kernel(const __grid_constant__ SplitTab tab, float* __restrict__ buf) {
    const int l = blockIdx.y;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < tab.floats[l]; i += gridDim.x * blockDim.x) {
        float hi, lo;
        split_weight(buf[(size_t)tab.w1_off[l] + i], hi, lo);
        const size_t d = (size_t)tab.w_off[l] + (size_t)(i >> 7) * 256 + (i & 127);
        buf[d] = hi;
        buf[d + 128] = lo;
    }
}

__global__ void pack_header_kernel(uint32_t* hdr, uint32_t C, uint32_t n, uint32_t px) {
    if (threadIdx.x == 0) { hdr[0] = PACK_MAGIC; hdr[1] = C; hdr[2] = n; hdr[3] = px; }
}

// Philox4x32-10 (Salmon et al. 2011), counter = (row, frame, slot, block), key = seed.
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                              uint32_t k1, uint32_t (&out)[4]) {
#pragma unroll
    for (int i = 0; i < 10; ++i) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
__device__ __forceinline__ float u01(uint32_t x) { return (float)(x >> 8) * 5.9604644775390625e-8f; }
__device__ __forceinline__ void box_muller(uint32_t a, uint32_t b, float& z0, float& z1) {
    const float u1 = ((float)(a >> 8) + 1.0f) * 5.9604644775390625e-8f;    // (0, 1]
    const float r = sqrtf(-2.0f * logf(u1));
    const float th = 6.283185307179586f * u01(b);
    z0 = r * cosf(th);
    z1 = r * sinf(th);
}

__global__ void noise_kernel(int T, int rows, int n2, int nw, uint32_t k0, uint32_t k1, int row_offset,
                             float* __restrict__ eps_where, float* __restrict__ eps_what, float* __restrict__ u_pres) {
    const int nblk = (nw + 3) / 4;          // what blocks; block 0 = where, block nblk+1 = presence
    const long long total = (long long)T * rows * n2 * (nblk + 2);
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int j = (int)(i % (nblk + 2));
        const long long site = i / (nblk + 2);            // (t, row, slot)
        const int s = (int)(site % n2);
        const int row = (int)((site / n2) % rows);
        const int t = (int)(site / ((long long)n2 * rows));
        uint32_t x[4];
        if (j == nblk + 1) {
            philox4x32_10((uint32_t)(row + row_offset), (uint32_t)t, (uint32_t)s, 63u, k0, k1, x);
            u_pres[site] = u01(x[0]);
        } else {
            philox4x32_10((uint32_t)(row + row_offset), (uint32_t)t, (uint32_t)s, (uint32_t)j, k0, k1, x);
            float z[4];
            box_muller(x[0], x[1], z[0], z[1]);
            box_muller(x[2], x[3], z[2], z[3]);
            if (j == 0) {
#pragma unroll
                for (int q = 0; q < 4; ++q) eps_where[site * 4 + q] = z[q];
            } else {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int d = (j - 1) * 4 + q;
                    if (d < nw) eps_what[site * nw + d] = z[q];
                }
            }
        }
    }
}

// Particle reductions (model.py:88-103,150-158; targets.py:38-75; ops.py:52-59).  One block; thread b
// owns sample b (K is small), then a block reduction for the batch means.
__global__ void objective_kernel(const float* __restrict__ lw_t, const float* __restrict__ lp_t, int T, int B, int K,
                                 float* __restrict__ log_weights, float* __restrict__ iwae_pe,
                                 float* __restrict__ imp_w, float* __restrict__ scalars) {
    __shared__ float red[4][32];
    float acc[4] = {0.f, 0.f, 0.f, 0.f};      // sum lw, sum iwae, sum ess, sum vimco proxy
    const float logK = logf((float)K);
    for (int b = threadIdx.x; b < B; b += blockDim.x) {
        float S = 0.f, mx = -INFINITY;
        for (int k = 0; k < K; ++k) {
            float a = 0.f;
            for (int t = 0; t < T; ++t) a += lw_t[(size_t)t * B * K + b * K + k];
            if (log_weights) log_weights[b * K + k] = a;
            S += a;
            mx = fmaxf(mx, a);
        }
        // second pass re-sums from global (K*T small) to avoid a K-sized local array
        float se = 0.f;
        for (int k = 0; k < K; ++k) {
            float a = 0.f;
            for (int t = 0; t < T; ++t) a += lw_t[(size_t)t * B * K + b * K + k];
            se += expf(a - mx);
        }
        const float lse = mx + logf(se);
        const float iw = lse - logK;                                        // targets.py:38-43
        if (iwae_pe) iwae_pe[b] = iw;
        float sw = 0.f, sw2 = 0.f, proxy = 0.f;
        for (int j = 0; j < K; ++j) {
            float lwj = 0.f, lpj = 0.f;
            for (int t = 0; t < T; ++t) {
                lwj += lw_t[(size_t)t * B * K + b * K + j];
                if (lp_t) lpj += lp_t[(size_t)t * B * K + b * K + j];
            }
            const float w = expf(lwj - lse);                                // softmax, model.py:100
            if (imp_w) imp_w[b * K + j] = w;
            sw += w;
            sw2 += w * w;
            // VIMCO control variate (targets.py:46-59): logsumexp over i of (i == j ? mean_{i != j} lw : lw_i) - log K
            const float abo = (S - lwj) / ((float)K - 1.f);
            float m2 = abo;
            for (int i = 0; i < K; ++i) {
                if (i == j) continue;
                float a = 0.f;
                for (int t = 0; t < T; ++t) a += lw_t[(size_t)t * B * K + b * K + i];
                m2 = fmaxf(m2, a);
            }
            float s2 = expf(abo - m2);
            for (int i = 0; i < K; ++i) {
                if (i == j) continue;
                float a = 0.f;
                for (int t = 0; t < T; ++t) a += lw_t[(size_t)t * B * K + b * K + i];
                s2 += expf(a - m2);
            }
            const float cv = m2 + logf(s2) - logK;
            proxy += -iw - (lwj - cv) * lpj;                                // targets.py:62-75
        }
        acc[0] += S;
        acc[1] += iw;
        acc[2] += sw * sw / sw2;                                            // ops.py:52-59
        acc[3] += proxy;
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        float v = warp_sum(acc[q]);
        if (lane == 0) red[q][warp] = v;
    }
    __syncthreads();
    if (threadIdx.x == 0 && scalars) {
        float tot[4] = {0.f, 0.f, 0.f, 0.f};
        for (int q = 0; q < 4; ++q)
            for (int w = 0; w < (int)(blockDim.x >> 5); ++w) tot[q] += red[q][w];
        for (int q = 0; q < SQAIR_OBJ_N; ++q) scalars[q] = 0.f;
        scalars[SQAIR_OBJ_ELBO_VAE] = tot[0] / (float)(B * K);
        scalars[SQAIR_OBJ_ELBO_IWAE] = tot[1] / (float)B;
        scalars[SQAIR_OBJ_ESS] = tot[2] / (float)B;
        scalars[SQAIR_OBJ_VIMCO_TARGET] = tot[3] / (float)(B * K) / (float)T;   // model.py:158
        scalars[SQAIR_OBJ_IWAE_TARGET] = -(tot[1] / (float)B) / (float)T;
    }
}

// Gradient of the VIMCO target w.r.t. the rows' summed log weights and discrete log-probs (targets.py:46-75;
// model.py:150-158): one thread per sample, K is small.
__global__ void objective_grad_kernel(const float* __restrict__ lw_t, const float* __restrict__ lp_t, int T, int B, int K,
                                      float* __restrict__ d_lw, float* __restrict__ d_lp) {
    const float logK = logf((float)K);
    for (int b = blockIdx.x * blockDim.x + threadIdx.x; b < B; b += gridDim.x * blockDim.x) {
        auto lw = [&](int k) {
            float a = 0.f;
            for (int t = 0; t < T; ++t) a += lw_t[(size_t)t * B * K + b * K + k];
            return a;
        };
        float S = 0.f, mx = -INFINITY;
        for (int k = 0; k < K; ++k) { const float a = lw(k); S += a; mx = fmaxf(mx, a); }
        float se = 0.f;
        for (int k = 0; k < K; ++k) se += expf(lw(k) - mx);
        const float lse = mx + logf(se);
        for (int j = 0; j < K; ++j) {
            const float lwj = lw(j);
            if (d_lw) d_lw[b * K + j] = -expf(lwj - lse) / ((float)B * (float)T);
            if (d_lp) {
                // control variate: logsumexp_i(i == j ? mean_{i != j} lw_i : lw_i) - log K
                const float abo = (S - lwj) / ((float)K - 1.f);
                float m2 = abo;
                for (int i = 0; i < K; ++i) if (i != j) m2 = fmaxf(m2, lw(i));
                float s2 = expf(abo - m2);
                for (int i = 0; i < K; ++i) if (i != j) s2 += expf(lw(i) - m2);
                const float cv = m2 + logf(s2) - logK;
                d_lp[b * K + j] = -(lwj - cv) / ((float)B * (float)K * (float)T);
            }
        }
    }
    (void)lp_t;
}

// SpatialTransformer forward (modules.py:165-172,204-218): one thread per glimpse texel.
__global__ void stn_glimpse_kernel(const float* __restrict__ img, const float* __restrict__ where,
                                   float* __restrict__ glimpse, int N, int H, int W, int G) {
    const long long total = (long long)N * G * G;
    const float hw = 0.5f * (float)(W - 1), hh = 0.5f * (float)(H - 1);
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int gx = (int)(i % G), gy = (int)((i / G) % G);
        const int nidx = (int)(i / ((long long)G * G));
        const float* wl = where + (size_t)nidx * 4;
        const float sx = fmaxf(sigmoidf_(__ldg(wl)), 1e-4f), sy = fmaxf(sigmoidf_(__ldg(wl + 1)), 1e-4f);
        const float tx = tanhf(__ldg(wl + 2)), ty = tanhf(__ldg(wl + 3));
        const float x = hw * (sx * lin11(gx, G) + tx) + hw;
        const float y = hh * (sy * lin11(gy, G) + ty) + hh;
        const float* im = img + (size_t)nidx * H * W;
        glimpse[i] = bilinear_zero_pad(x, y, W, H, [&](int ix, int iy) { return __ldg(im + iy * W + ix); });
    }
}

// Backward of the glimpse sampler w.r.t. the where-logits: one block per glimpse, threads over texels, block reduction
// of the four coordinate-space sums, chain rule through to_coords in thread 0.
__global__ void stn_glimpse_grad_kernel(const float* __restrict__ img, const float* __restrict__ where,
                                        const float* __restrict__ d_glimpse, float* __restrict__ d_where, int H, int W, int G) {
    const int nidx = blockIdx.x;
    const float* wl = where + (size_t)nidx * 4;
    const float s0 = sigmoidf_(__ldg(wl)), s1 = sigmoidf_(__ldg(wl + 1));
    const float sx = fmaxf(s0, 1e-4f), sy = fmaxf(s1, 1e-4f);
    const float tx = tanhf(__ldg(wl + 2)), ty = tanhf(__ldg(wl + 3));
    const float hw = 0.5f * (float)(W - 1), hh = 0.5f * (float)(H - 1);
    const float* im = img + (size_t)nidx * H * W;
    float a_sx = 0.f, a_sy = 0.f, a_tx = 0.f, a_ty = 0.f;
    for (int i = threadIdx.x; i < G * G; i += blockDim.x) {
        const int gx = i % G, gy = i / G;
        const float u = lin11(gx, G), v = lin11(gy, G);
        const float x = hw * (sx * u + tx) + hw, y = hh * (sy * v + ty) + hh;
        if (!(x > -1.f && y > -1.f && x < (float)W && y < (float)H)) continue;      // sample is 0 there, so is its gradient
        const float fx = floorf(x), fy = floorf(y);
        const float dx = fx + 1.f - x, dy = fy + 1.f - y;
        const int ifx = (int)fx, ify = (int)fy, icx = ifx + 1, icy = ify + 1;
        const bool fx_ok = ifx >= 0 && ifx <= W - 1, cx_ok = icx >= 0 && icx <= W - 1;
        const bool fy_ok = ify >= 0 && ify <= H - 1, cy_ok = icy >= 0 && icy <= H - 1;
        const float v00 = (fx_ok && fy_ok) ? __ldg(im + ify * W + ifx) : 0.f;
        const float v11 = (cx_ok && cy_ok) ? __ldg(im + icy * W + icx) : 0.f;
        const float v01 = (fx_ok && cy_ok) ? __ldg(im + icy * W + ifx) : 0.f;
        const float v10 = (cx_ok && fy_ok) ? __ldg(im + ify * W + icx) : 0.f;
        const float gxv = dy * (v10 - v00) + (1.f - dy) * (v11 - v01);             // d sample / d x
        const float gyv = dx * (v01 - v00) + (1.f - dx) * (v11 - v10);             // d sample / d y
        const float dg = __ldg(d_glimpse + (size_t)nidx * G * G + i);
        a_sx += dg * gxv * hw * u; a_tx += dg * gxv * hw;
        a_sy += dg * gyv * hh * v; a_ty += dg * gyv * hh;
    }
    __shared__ float red[4][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float acc[4] = {a_sx, a_sy, a_tx, a_ty};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const float r = warp_sum(acc[q]);
        if (lane == 0) red[q][warp] = r;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        float t[4] = {0.f, 0.f, 0.f, 0.f};
        for (int q = 0; q < 4; ++q)
            for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t[q] += red[q][w];
        float* o = d_where + (size_t)nidx * 4;
        o[0] = t[0] * s0 * (1.f - s0);          // straight-through clip: the gradient ignores the 1e-4 floor
        o[1] = t[1] * s1 * (1.f - s1);
        o[2] = t[2] * (1.f - tx * tx);
        o[3] = t[3] * (1.f - ty * ty);
    }
}

// AIRDecoder._decode/_add_mean_image + Normal(canvas, std).log_prob(img) (modules.py:435-467; seq.py:272-273).
// One block per image; the n decoded glimpses are staged in shared memory.
__global__ void canvas_ll_kernel(const float* __restrict__ glimpse, const float* __restrict__ where,
                                 const float* __restrict__ presence, const float* __restrict__ mean_img,
                                 const float* __restrict__ img, float* __restrict__ canvas, float* __restrict__ data_ll,
                                 int n, int H, int W, int G, float output_std, float bg_std) {
    extern __shared__ float sg[];              // [n][G*G] then coords [n][5]
    const int b = blockIdx.x, g = G * G;
    float* cc = sg + n * g;
    for (int i = threadIdx.x; i < n * g; i += blockDim.x) sg[i] = glimpse[(size_t)b * n * g + i];
    for (int s = threadIdx.x; s < n; s += blockDim.x) {
        const float* wl = where + ((size_t)b * n + s) * 4;
        cc[s * 5 + 0] = fmaxf(sigmoidf_(wl[0]), 1e-4f);
        cc[s * 5 + 1] = fmaxf(sigmoidf_(wl[1]), 1e-4f);
        cc[s * 5 + 2] = tanhf(wl[2]);
        cc[s * 5 + 3] = tanhf(wl[3]);
        cc[s * 5 + 4] = presence[(size_t)b * n + s];
    }
    __syncthreads();
    const float hg = 0.5f * (float)(G - 1);
    const float sf0 = sqrtf(output_std), sb0 = sqrtf(bg_std);
    const float sf = sf0 * sf0, sb = sb0 * sb0;
    float ll = 0.f;
    for (int px = threadIdx.x; px < H * W; px += blockDim.x) {
        const int iy = px / W, ix = px % W;
        const float u = lin11(ix, W), v = lin11(iy, H);
        float cv = 0.f, nz = 0.f;
        for (int s = 0; s < n; ++s) {
            const float pres = cc[s * 5 + 4];
            if (pres == 0.f) continue;
            const float xg = hg * ((u - cc[s * 5 + 2]) / cc[s * 5 + 0]) + hg;
            const float yg = hg * ((v - cc[s * 5 + 3]) / cc[s * 5 + 1]) + hg;
            const float* gl = sg + s * g;
            cv += pres * bilinear_zero_pad(xg, yg, G, G, [&](int gx, int gy) { return gl[gy * G + gx]; });
            nz += pres * bilinear_zero_pad(xg, yg, G, G, [&](int, int) { return 1.f; });
        }
        const float mask = sigmoidf_(-10.f + nz * 20.f);
        cv += __ldg(mean_img + px) * mask;
        const float std = mask * sf + (1.f - mask) * sb;
        ll += normal_lp(__ldg(img + (size_t)b * H * W + px), cv, std);
        canvas[(size_t)b * H * W + px] = cv;
    }
    __shared__ float red[32];
    ll = warp_sum(ll);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ll;
    __syncthreads();
    if (threadIdx.x == 0) {
        float a = 0.f;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) a += red[w];
        data_ll[b] = a;
    }
}

// Weight gradient dW[K,N] (+)= X[M,K]^T dY[M,N].  Block tile 64 (K) x 64 (N), 4 warps of 32 x 32 (2 m16 x 4 n8 MMA tiles),
// M walked in chunks of 32 rows staged in shared memory (row stride 72 floats: conflict-free fragment reads), grid.z
// splits M (partial tiles meet by atomicAdd).  Same arithmetic as the forward layers: tf32 (hi, lo) of both operands,
// four products per 8 reduction steps summed from zero on the tensor core, fp32 round-to-nearest accumulation outside.
constexpr int WG_T = 64, WG_MC = 32, WG_LD = 72;
__global__ void __launch_bounds__(128) wgrad_kernel(const float* __restrict__ X, const float* __restrict__ dY, float* __restrict__ dW,
                                                    int M, int K, int N, int m_per_block, int atomic) {
    __shared__ float Xs[WG_MC * WG_LD], Ys[WG_MC * WG_LD];
    const int n0 = blockIdx.x * WG_T, k0 = blockIdx.y * WG_T;
    const int m_begin = blockIdx.z * m_per_block, m_end = min(M, m_begin + m_per_block);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const int wk = (warp >> 1) * 32, wn = (warp & 1) * 32;          // warp tile origin inside the block tile
    float acc[2][4][4];
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b)
#pragma unroll
            for (int q = 0; q < 4; ++q) acc[a][b][q] = 0.f;
    for (int m0 = m_begin; m0 < m_end; m0 += WG_MC) {
        __syncthreads();
        for (int i = threadIdx.x; i < WG_MC * WG_T; i += blockDim.x) {     // zero-filled edges
            const int mm = i / WG_T, cc = i - mm * WG_T, m = m0 + mm;
            Xs[mm * WG_LD + cc] = (m < m_end && k0 + cc < K) ? __ldg(X + (size_t)m * K + k0 + cc) : 0.f;
            Ys[mm * WG_LD + cc] = (m < m_end && n0 + cc < N) ? __ldg(dY + (size_t)m * N + n0 + cc) : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int ms = 0; ms < WG_MC; ms += 8) {
            uint32_t ah[2][4], al[2][4], bh[4][2], bl[4][2];
#pragma unroll
            for (int a = 0; a < 2; ++a) {       // A = X^T: element (row k, col m) = Xs[m][k]
                const float* p = Xs + (ms + t) * WG_LD + wk + a * 16 + g;
                split_tf32(p[0], ah[a][0], al[a][0]);
                split_tf32(p[8], ah[a][1], al[a][1]);
                split_tf32(p[4 * WG_LD], ah[a][2], al[a][2]);
                split_tf32(p[4 * WG_LD + 8], ah[a][3], al[a][3]);
            }
#pragma unroll
            for (int b = 0; b < 4; ++b) {       // B = dY: element (row m, col n)
                const float* p = Ys + (ms + t) * WG_LD + wn + b * 8 + g;
                split_tf32(p[0], bh[b][0], bl[b][0]);
                split_tf32(p[4 * WG_LD], bh[b][1], bl[b][1]);
            }
#pragma unroll
            for (int a = 0; a < 2; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                    float d[4];
                    mma_tf32_zero(d, al[a], bl[b][0], bl[b][1]);
                    mma_tf32(d, al[a], bh[b][0], bh[b][1]);
                    mma_tf32(d, ah[a], bl[b][0], bl[b][1]);
                    mma_tf32(d, ah[a], bh[b][0], bh[b][1]);
#pragma unroll
                    for (int q = 0; q < 4; ++q) acc[a][b][q] += d[q];
                }
        }
    }
    // C fragment: d[0] = (row g, col 2t), d[1] = (g, 2t+1), d[2] = (g+8, 2t), d[3] = (g+8, 2t+1)
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b)
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int k = k0 + wk + a * 16 + g + (q >> 1) * 8, n = n0 + wn + b * 8 + 2 * t + (q & 1);
                if (k < K && n < N) {
                    if (atomic) atomicAdd(dW + (size_t)k * N + n, acc[a][b][q]);
                    else dW[(size_t)k * N + n] = acc[a][b][q];
                }
            }
}

// Input gradient dX[M,K] = dY[M,N] . W[K,N]^T.  Block tile 64 (M) x 64 (K), 4 warps of 32 x 32, N walked in chunks of 32
// staged in shared memory (row stride 36 floats: conflict-free fragment reads); arithmetic as in wgrad_kernel.
constexpr int DG_T = 64, DG_NC = 32, DG_LD = 36;
__global__ void __launch_bounds__(128) dgrad_kernel(const float* __restrict__ dY, const float* __restrict__ Wm, float* __restrict__ dX,
                                                    int M, int K, int N) {
    __shared__ float As[DG_T * DG_LD], Bs[DG_T * DG_LD];
    const int k0 = blockIdx.x * DG_T, m0 = blockIdx.y * DG_T;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const int wm = (warp >> 1) * 32, wk = (warp & 1) * 32;
    float acc[2][4][4];
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b)
#pragma unroll
            for (int q = 0; q < 4; ++q) acc[a][b][q] = 0.f;
    for (int n0 = 0; n0 < N; n0 += DG_NC) {
        __syncthreads();
        for (int i = threadIdx.x; i < DG_T * DG_NC; i += blockDim.x) {     // zero-filled edges
            const int rr = i / DG_NC, cc = i - rr * DG_NC;
            As[rr * DG_LD + cc] = (m0 + rr < M && n0 + cc < N) ? __ldg(dY + (size_t)(m0 + rr) * N + n0 + cc) : 0.f;
            Bs[rr * DG_LD + cc] = (k0 + rr < K && n0 + cc < N) ? __ldg(Wm + (size_t)(k0 + rr) * N + n0 + cc) : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int ns = 0; ns < DG_NC; ns += 8) {
            uint32_t ah[2][4], al[2][4], bh[4][2], bl[4][2];
#pragma unroll
            for (int a = 0; a < 2; ++a) {       // A = dY: element (row m, col n)
                const float* p = As + (wm + a * 16 + g) * DG_LD + ns + t;
                split_tf32(p[0], ah[a][0], al[a][0]);
                split_tf32(p[8 * DG_LD], ah[a][1], al[a][1]);
                split_tf32(p[4], ah[a][2], al[a][2]);
                split_tf32(p[8 * DG_LD + 4], ah[a][3], al[a][3]);
            }
#pragma unroll
            for (int b = 0; b < 4; ++b) {       // B = W^T: element (row n, col k) = Bs[k][n]
                const float* p = Bs + (wk + b * 8 + g) * DG_LD + ns + t;
                split_tf32(p[0], bh[b][0], bl[b][0]);
                split_tf32(p[4], bh[b][1], bl[b][1]);
            }
#pragma unroll
            for (int a = 0; a < 2; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                    float d[4];
                    mma_tf32_zero(d, al[a], bl[b][0], bl[b][1]);
                    mma_tf32(d, al[a], bh[b][0], bh[b][1]);
                    mma_tf32(d, ah[a], bl[b][0], bl[b][1]);
                    mma_tf32(d, ah[a], bh[b][0], bh[b][1]);
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
                const int m = m0 + wm + a * 16 + g + (q >> 1) * 8, k = k0 + wk + b * 8 + 2 * t + (q & 1);
                if (m < M && k < K) dX[(size_t)m * K + k] = acc[a][b][q];
            }
}

// bilinear sample with zero padding plus its derivatives w.r.t. the sample position; `w4`/`idx4` receive the four texel
// weights / linear indices (-1 = outside) for the scatter of the data gradient.
struct Bilin {
    float val, ddx, ddy;
    float w4[4];
    int idx4[4];
};
template <class Fetch>
__device__ __forceinline__ Bilin bilinear_zero_pad_grad(float x, float y, int w, int h, Fetch fetch) {
    Bilin r;
    r.val = r.ddx = r.ddy = 0.f;
#pragma unroll
    for (int q = 0; q < 4; ++q) { r.w4[q] = 0.f; r.idx4[q] = -1; }
    if (!(x > -1.f && y > -1.f && x < (float)w && y < (float)h)) return r;
    const float fx = floorf(x), fy = floorf(y);
    const float dx = fx + 1.f - x, dy = fy + 1.f - y;
    const int ifx = (int)fx, ify = (int)fy, icx = ifx + 1, icy = ify + 1;
    const bool fx_ok = ifx >= 0 && ifx <= w - 1, cx_ok = icx >= 0 && icx <= w - 1;
    const bool fy_ok = ify >= 0 && ify <= h - 1, cy_ok = icy >= 0 && icy <= h - 1;
    const float v00 = (fx_ok && fy_ok) ? fetch(ifx, ify) : 0.f;
    const float v11 = (cx_ok && cy_ok) ? fetch(icx, icy) : 0.f;
    const float v01 = (fx_ok && cy_ok) ? fetch(ifx, icy) : 0.f;
    const float v10 = (cx_ok && fy_ok) ? fetch(icx, ify) : 0.f;
    r.val = dx * dy * v00 + (1.f - dx) * (1.f - dy) * v11 + dx * (1.f - dy) * v01 + (1.f - dx) * dy * v10;
    r.ddx = dy * (v10 - v00) + (1.f - dy) * (v11 - v01);
    r.ddy = dx * (v01 - v00) + (1.f - dx) * (v11 - v10);
    if (fx_ok && fy_ok) { r.w4[0] = dx * dy; r.idx4[0] = ify * w + ifx; }
    if (cx_ok && cy_ok) { r.w4[1] = (1.f - dx) * (1.f - dy); r.idx4[1] = icy * w + icx; }
    if (fx_ok && cy_ok) { r.w4[2] = dx * (1.f - dy); r.idx4[2] = icy * w + ifx; }
    if (cx_ok && fy_ok) { r.w4[3] = (1.f - dx) * dy; r.idx4[3] = ify * w + icx; }
    return r;
}

// Backward of canvas_ll_kernel.  One block per image; the n glimpses and their gradient accumulators live in shared
// memory (the data gradient is a scatter-add: shared-memory atomics), the where-gradient sums are block-reduced.
__global__ void canvas_ll_grad_kernel(const float* __restrict__ glimpse, const float* __restrict__ where,
                                      const float* __restrict__ presence, const float* __restrict__ mean_img,
                                      const float* __restrict__ img, const float* __restrict__ d_ll,
                                      float* __restrict__ d_glimpse, float* __restrict__ d_where, float* __restrict__ d_mean_img,
                                      int n, int H, int W, int G, float output_std, float bg_std) {
    extern __shared__ float sg[];              // glimpses [n][g], gradient accumulators [n][g], coords [n][7], where sums [n][4]
    const int b = blockIdx.x, g = G * G;
    float* dgl = sg + n * g;
    float* cc = dgl + n * g;
    float* wsum = cc + n * 7;
    for (int i = threadIdx.x; i < n * g; i += blockDim.x) { sg[i] = glimpse[(size_t)b * n * g + i]; dgl[i] = 0.f; }
    for (int s = threadIdx.x; s < n; s += blockDim.x) {
        const float* wl = where + ((size_t)b * n + s) * 4;
        const float s0 = sigmoidf_(wl[0]), s1 = sigmoidf_(wl[1]);
        cc[s * 7 + 0] = fmaxf(s0, 1e-4f);
        cc[s * 7 + 1] = fmaxf(s1, 1e-4f);
        cc[s * 7 + 2] = tanhf(wl[2]);
        cc[s * 7 + 3] = tanhf(wl[3]);
        cc[s * 7 + 4] = presence[(size_t)b * n + s];
        cc[s * 7 + 5] = s0 * (1.f - s0);        // d sigmoid (straight-through clip: the 1e-4 floor is ignored)
        cc[s * 7 + 6] = s1 * (1.f - s1);
    }
    for (int i = threadIdx.x; i < n * 4; i += blockDim.x) wsum[i] = 0.f;
    __syncthreads();
    const float hg = 0.5f * (float)(G - 1);
    const float sf0 = sqrtf(output_std), sb0 = sqrtf(bg_std);
    const float sf = sf0 * sf0, sb = sb0 * sb0;
    const float up = d_ll[b];
    for (int px = threadIdx.x; px < H * W; px += blockDim.x) {
        const int iy = px / W, ix = px % W;
        const float u = lin11(ix, W), v = lin11(iy, H);
        // forward values of this pixel
        float cv = 0.f, nz = 0.f;
        for (int s = 0; s < n; ++s) {
            const float pres = cc[s * 7 + 4];
            if (pres == 0.f) continue;
            const float xg = hg * ((u - cc[s * 7 + 2]) / cc[s * 7 + 0]) + hg;
            const float yg = hg * ((v - cc[s * 7 + 3]) / cc[s * 7 + 1]) + hg;
            const float* gl = sg + s * g;
            cv += pres * bilinear_zero_pad(xg, yg, G, G, [&](int gx, int gy) { return gl[gy * G + gx]; });
            nz += pres * bilinear_zero_pad(xg, yg, G, G, [&](int, int) { return 1.f; });
        }
        const float mask = sigmoidf_(-10.f + nz * 20.f);
        const float mi = __ldg(mean_img + px);
        cv += mi * mask;
        const float std = mask * sf + (1.f - mask) * sb;
        const float z = (__ldg(img + (size_t)b * H * W + px) - cv) / std;
        // d ll / d canvas, d ll / d std  (tfd.Normal.log_prob)
        const float d_cv = up * z / std;
        const float d_std = up * (z * z - 1.f) / std;
        const float d_mask = d_cv * mi + d_std * (sf - sb);
        const float d_nz = d_mask * 20.f * mask * (1.f - mask);
        atomicAdd(d_mean_img + px, d_cv * mask);
        for (int s = 0; s < n; ++s) {
            const float pres = cc[s * 7 + 4];
            if (pres == 0.f) continue;
            const float sx = cc[s * 7 + 0], sy = cc[s * 7 + 1], tx = cc[s * 7 + 2], ty = cc[s * 7 + 3];
            const float xg = hg * ((u - tx) / sx) + hg, yg = hg * ((v - ty) / sy) + hg;
            const float* gl = sg + s * g;
            const Bilin bg = bilinear_zero_pad_grad(xg, yg, G, G, [&](int gx, int gy) { return gl[gy * G + gx]; });
            const Bilin bo = bilinear_zero_pad_grad(xg, yg, G, G, [&](int, int) { return 1.f; });
#pragma unroll
            for (int q = 0; q < 4; ++q)
                if (bg.idx4[q] >= 0) atomicAdd(dgl + s * g + bg.idx4[q], d_cv * pres * bg.w4[q]);
            // position gradients of both samplers, then the inverse warp x_g = hg ((u - tx) / sx) + hg
            const float d_xg = pres * (d_cv * bg.ddx + d_nz * bo.ddx), d_yg = pres * (d_cv * bg.ddy + d_nz * bo.ddy);
            atomicAdd(wsum + s * 4 + 0, d_xg * (-hg * (u - tx) / (sx * sx)));
            atomicAdd(wsum + s * 4 + 1, d_yg * (-hg * (v - ty) / (sy * sy)));
            atomicAdd(wsum + s * 4 + 2, d_xg * (-hg / sx));
            atomicAdd(wsum + s * 4 + 3, d_yg * (-hg / sy));
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n * g; i += blockDim.x) d_glimpse[(size_t)b * n * g + i] = dgl[i];
    for (int i = threadIdx.x; i < n * 4; i += blockDim.x) {
        const int s = i / 4, q = i % 4;
        const float chain = q == 0 ? cc[s * 7 + 5] : q == 1 ? cc[s * 7 + 6] : q == 2 ? (1.f - cc[s * 7 + 2] * cc[s * 7 + 2])
                                                                                     : (1.f - cc[s * 7 + 3] * cc[s * 7 + 3]);
        d_where[((size_t)b * n + s) * 4 + q] = wsum[i] * chain;
    }
}

// ---------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------
extern "C" {

const char* sqair_last_error(void) { return g_err.c_str(); }
int sqair_version(void) { return 100; }

int sqair_query_sizes(const sqair_cfg* cfg, sqair_sizes* out) {
    if (!cfg || !out) return fail(SQAIR_EINVAL, "null argument");
    std::string e = validate_cfg(*cfg);
    if (!e.empty()) return fail(SQAIR_EINVAL, e);
    auto tab = param_table(*cfg);
    Shape sh;
    e = choose_shape(*cfg, tab, sh);
    if (!e.empty()) return fail(SQAIR_EUNSUPPORTED, e);
    const int64_t rows = (int64_t)cfg->B * cfg->K, sites = (int64_t)cfg->T * rows * 2 * cfg->n;
    out->param_count = tab.back().offset + tab.back().count;
    out->packed_floats = sh.packed_total;
    out->eps_where_floats = sites * 4;
    out->eps_what_floats = sites * cfg->n_what;
    out->u_pres_floats = sites;
    out->rows = (int32_t)rows;
    out->rows_per_cta = sh.R;
    out->cluster_size = sh.C;
    out->n_ctas = (int32_t)((rows + sh.R - 1) / sh.R) * sh.C;
    out->smem_bytes = sh.plan.sm.total * (int)sizeof(float);
    out->n_layers = sh.plan.nseq;
    return SQAIR_OK;
}

int sqair_param_layout(const sqair_cfg* cfg, sqair_param_desc* descs, int32_t* n) {
    if (!cfg || !n) return fail(SQAIR_EINVAL, "null argument");
    std::string e = validate_cfg(*cfg);
    if (!e.empty()) return fail(SQAIR_EINVAL, e);
    auto tab = param_table(*cfg);
    if (!descs) { *n = (int32_t)tab.size(); return SQAIR_OK; }
    if (*n < (int32_t)tab.size()) return fail(SQAIR_EINVAL, "descriptor array too small");
    for (size_t i = 0; i < tab.size(); ++i) {
        sqair_param_desc& d = descs[i];
        memset(&d, 0, sizeof(d));
        snprintf(d.name, sizeof(d.name), "%s", tab[i].name.c_str());
        d.ndim = tab[i].ndim;
        for (int k = 0; k < 3; ++k) d.shape[k] = tab[i].shape[k];
        d.offset = tab[i].offset;
        d.packed_offset = tab[i].packed_offset;
    }
    *n = (int32_t)tab.size();
    return SQAIR_OK;
}

int sqair_pack_params(const sqair_cfg* cfg, const float* params, float* packed, void* stream) {
    if (!cfg || !params || !packed) return fail(SQAIR_EINVAL, "null argument");
    std::string e = validate_cfg(*cfg);
    if (!e.empty()) return fail(SQAIR_EINVAL, e);
    auto tab = param_table(*cfg);
    Shape sh;
    e = choose_shape(*cfg, tab, sh);
    if (!e.empty()) return fail(SQAIR_EUNSUPPORTED, e);
    if (tab.size() > 128 || sh.pieces.size() > 160) return fail(SQAIR_EUNSUPPORTED, "too many variables");
    PackTab pt;
    memset(&pt, 0, sizeof(pt));
    pt.n = (int)tab.size();
    for (size_t i = 0; i < tab.size(); ++i) {
        pt.src[i] = (int)tab[i].offset; pt.dst[i] = (int)tab[i].packed_offset; pt.cnt[i] = (int)tab[i].count;
    }
    PieceTab qt;
    memset(&qt, 0, sizeof(qt));
    qt.n = (int)sh.pieces.size();
    for (size_t i = 0; i < sh.pieces.size(); ++i) {
        const Piece& p = sh.pieces[i];
        const Layer& L = sh.plan.L[p.layer];
        qt.p[i] = PieceDev{p.vrow0, p.vcol0, p.K, p.N, (int)p.src_off, p.src_ld, L.w1_off, L.ksteps, L.Nc, L.panel_floats / 2};
    }
    cudaStream_t st = (cudaStream_t)stream;
    const int total = (int)(tab.back().offset + tab.back().count);
    CUDA_TRY(cudaMemsetAsync(packed, 0, sh.packed_total * sizeof(float), st));
    pack_kernel<<<592, 256, 0, st>>>(pt, params, packed, total);
    CUDA_TRY(cudaGetLastError());
    pack_panels_kernel<<<dim3(64, qt.n), 256, 0, st>>>(qt, params, packed);
    CUDA_TRY(cudaGetLastError());
    SplitTab stab;
    memset(&stab, 0, sizeof(stab));
    for (int i = 0; i < L_COUNT; ++i) {
        const Layer& L = sh.plan.L[i];
        if (L.nhead == 0) continue;
        stab.w1_off[stab.n] = L.w1_off; stab.w_off[stab.n] = L.w_off; stab.floats[stab.n] = L.npanel * (L.panel_floats / 2);
        ++stab.n;
    }
    split_panels_kernel<<<dim3(32, stab.n), 256, 0, st>>>(stab, packed);
    CUDA_TRY(cudaGetLastError());
    // layout header: the sequence kernel refuses (traps on) a buffer packed for another cluster size / model
    pack_header_kernel<<<1, 32, 0, st>>>(reinterpret_cast<uint32_t*>(packed + sh.plan.phdr_off), (uint32_t)sh.C, (uint32_t)cfg->n,
                                         (uint32_t)(cfg->H * cfg->W));
    CUDA_TRY(cudaGetLastError());
    {
        std::lock_guard<std::mutex> lock(g_tag_mutex);
        bool found = false;
        for (auto& kv : g_tags) if (kv.first == packed) { kv.second = tag_of(sh); found = true; }
        if (!found) {
            if (g_tags.size() >= 64) g_tags.erase(g_tags.begin());
            g_tags.emplace_back((const void*)packed, tag_of(sh));
        }
    }
    return SQAIR_OK;
}

int sqair_fill_noise(const sqair_cfg* cfg, uint64_t seed, int32_t row_offset, float* eps_where, float* eps_what,
                     float* u_pres, void* stream) {
    if (!cfg || !eps_where || !eps_what || !u_pres) return fail(SQAIR_EINVAL, "null argument");
    std::string e = validate_cfg(*cfg);
    if (!e.empty()) return fail(SQAIR_EINVAL, e);
    const int rows = cfg->B * cfg->K;
    noise_kernel<<<592, 256, 0, (cudaStream_t)stream>>>(cfg->T, rows, 2 * cfg->n, cfg->n_what, (uint32_t)(seed & 0xffffffffu),
                                                        (uint32_t)(seed >> 32), row_offset, eps_where, eps_what, u_pres);
    CUDA_TRY(cudaGetLastError());
    return SQAIR_OK;
}

static int forward_impl(const sqair_cfg* cfg, const float* packed_params, const float* obs, const float* eps_where,
                        const float* eps_what, const float* u_pres, const sqair_outputs* out, float* stash, void* stream,
                        const float* eps_where_prior = nullptr, const float* eps_what_prior = nullptr,
                        const float* u_pres_prior = nullptr, int generate_after = -1) {
    if (!cfg || !packed_params || !obs || !eps_where || !eps_what || !u_pres || !out)
        return fail(SQAIR_EINVAL, "null argument");
    if ((reinterpret_cast<uintptr_t>(obs) | reinterpret_cast<uintptr_t>(packed_params)) & 15)
        return fail(SQAIR_EINVAL, "obs and packed_params must be 16-byte aligned (bulk frame copies / 128-bit weight loads)");
    std::string e = validate_cfg(*cfg);
    if (!e.empty()) return fail(SQAIR_EINVAL, e);
    auto tab = param_table(*cfg);
    Shape sh;
    e = choose_shape(*cfg, tab, sh);
    if (!e.empty()) return fail(SQAIR_EUNSUPPORTED, e);
    {
        std::lock_guard<std::mutex> lock(g_tag_mutex);
        for (auto& kv : g_tags)
            if (kv.first == packed_params && !(kv.second == tag_of(sh)))
                return fail(SQAIR_EINVAL, "packed_params were packed for a different launch shape (cluster size); "
                                          "call sqair_pack_params with the configuration of this call");
    }
    Job job{packed_params, obs, eps_where, eps_what, u_pres, *out, env_int("SQAIR_DEBUG_FLAGS"), nullptr, stash,
            eps_where_prior, eps_what_prior, u_pres_prior, generate_after};
    cudaStream_t st = (cudaStream_t)stream;
#ifdef SQAIR_ONLY_R              // tuning builds: one instantiation compiles in seconds
    if (sh.R == SQAIR_ONLY_R) return launch_sequence<SQAIR_ONLY_R>(sh.plan, job, st);
#else
    switch (sh.R) {
        case 1: return launch_sequence<1>(sh.plan, job, st);
        case 2: return launch_sequence<2>(sh.plan, job, st);
        case 3: return launch_sequence<3>(sh.plan, job, st);
        case 4: return launch_sequence<4>(sh.plan, job, st);
        case 5: return launch_sequence<5>(sh.plan, job, st);
        case 6: return launch_sequence<6>(sh.plan, job, st);
    }
#endif
    return fail(SQAIR_EUNSUPPORTED, "unsupported rows per block");
}

int sqair_forward(const sqair_cfg* cfg, const float* packed_params, const float* obs, const float* eps_where,
                  const float* eps_what, const float* u_pres, const sqair_outputs* out, void* stream) {
    return forward_impl(cfg, packed_params, obs, eps_where, eps_what, u_pres, out, nullptr, stream);
}

int sqair_forward_train(const sqair_cfg* cfg, const float* packed_params, const float* obs, const float* eps_where,
                        const float* eps_what, const float* u_pres, const sqair_outputs* out, float* stash, void* stream) {
    if (!stash) return fail(SQAIR_EINVAL, "null stash (sqair_query_sizes: stash_floats)");
    return forward_impl(cfg, packed_params, obs, eps_where, eps_what, u_pres, out, stash, stream);
}

int sqair_forward_generate(const sqair_cfg* cfg, const float* packed_params, const float* obs, const float* eps_where,
                           const float* eps_what, const float* u_pres, const float* eps_where_prior, const float* eps_what_prior,
                           const float* u_pres_prior, int32_t generate_after, const sqair_outputs* out, void* stream) {
    if (!eps_where_prior || !eps_what_prior || !u_pres_prior) return fail(SQAIR_EINVAL, "null prior-noise argument");
    return forward_impl(cfg, packed_params, obs, eps_where, eps_what, u_pres, out, nullptr, stream, eps_where_prior, eps_what_prior,
                        u_pres_prior, generate_after);
}

int sqair_objective(const float* log_w_t, const float* disc_lp_t, int32_t T, int32_t B, int32_t K, float* log_weights,
                    float* elbo_iwae_per_example, float* importance_weights, float* scalars, void* stream) {
    if (!log_w_t || T < 1 || B < 1 || K < 1) return fail(SQAIR_EINVAL, "bad argument");
    objective_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(log_w_t, disc_lp_t, T, B, K, log_weights, elbo_iwae_per_example,
                                                           importance_weights, scalars);
    CUDA_TRY(cudaGetLastError());
    return SQAIR_OK;
}

int sqair_objective_grad(const float* log_w_t, const float* disc_lp_t, int32_t T, int32_t B, int32_t K, float* d_log_weights,
                         float* d_discrete_log_prob, void* stream) {
    if (!log_w_t || T < 1 || B < 1 || K < 1) return fail(SQAIR_EINVAL, "bad argument");
    objective_grad_kernel<<<(B + 127) / 128, 128, 0, (cudaStream_t)stream>>>(log_w_t, disc_lp_t, T, B, K, d_log_weights,
                                                                               d_discrete_log_prob);
    CUDA_TRY(cudaGetLastError());
    return SQAIR_OK;
}

int sqair_stn_glimpse(const float* img, const float* where, float* glimpse, int32_t N, int32_t H, int32_t W, int32_t G,
                      void* stream) {
    if (!img || !where || !glimpse || N < 1 || H < 2 || W < 2 || G < 2) return fail(SQAIR_EINVAL, "bad argument");
    const long long total = (long long)N * G * G;
    int blocks = (int)((total + 255) / 256);
    if (blocks > 148 * 16) blocks = 148 * 16;
    stn_glimpse_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(img, where, glimpse, N, H, W, G);
    CUDA_TRY(cudaGetLastError());
    return SQAIR_OK;
}

int sqair_wgrad(const float* x, const float* dy, float* dw, int32_t M, int32_t K, int32_t N, int32_t accumulate, void* stream) {
    if (!x || !dy || !dw || M < 1 || K < 1 || N < 1) return fail(SQAIR_EINVAL, "bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    {   // Blackwell path: TMA-fed tcgen05.mma (3xTF32) with the accumulator in TMEM, when TMA can describe the operands
        const sqi::TcOperand ox{x, (int64_t)K, 0, 1}, oy{dy, (int64_t)N, 0, 1};
        if (sqi::wgrad_tc_supported(ox, oy, M, K, N)) {
            if (!accumulate) CUDA_TRY(cudaMemsetAsync(dw, 0, (size_t)K * N * sizeof(float), st));
            return sqi::wgrad_tc(ox, oy, dw, N, M, K, N, st);
        }
    }
    const int tiles = ((N + WG_T - 1) / WG_T) * ((K + WG_T - 1) / WG_T);
    // split M so that the grid covers the 148 SMs a few times over; partial tiles then meet by atomicAdd
    int msplit = (4 * 148 + tiles - 1) / tiles;
    const int max_split = (M + 4 * WG_MC - 1) / (4 * WG_MC);
    if (msplit > max_split) msplit = max_split;
    if (msplit < 1) msplit = 1;
    int m_per_block = ((M + msplit - 1) / msplit + WG_MC - 1) / WG_MC * WG_MC;
    msplit = (M + m_per_block - 1) / m_per_block;
    const int atomic = (msplit > 1 || accumulate) ? 1 : 0;
    if (atomic && !accumulate) CUDA_TRY(cudaMemsetAsync(dw, 0, (size_t)K * N * sizeof(float), st));
    wgrad_kernel<<<dim3((N + WG_T - 1) / WG_T, (K + WG_T - 1) / WG_T, msplit), 128, 0, st>>>(x, dy, dw, M, K, N, m_per_block, atomic);
    CUDA_TRY(cudaGetLastError());
    return SQAIR_OK;
}

int sqair_dgrad(const float* dy, const float* w, float* dx, int32_t M, int32_t K, int32_t N, void* stream) {
    if (!dy || !w || !dx || M < 1 || K < 1 || N < 1) return fail(SQAIR_EINVAL, "bad argument");
    dgrad_kernel<<<dim3((K + DG_T - 1) / DG_T, (M + DG_T - 1) / DG_T), 128, 0, (cudaStream_t)stream>>>(dy, w, dx, M, K, N);
    CUDA_TRY(cudaGetLastError());
    return SQAIR_OK;
}

int sqair_stn_glimpse_grad(const float* img, const float* where, const float* d_glimpse, float* d_where, int32_t N, int32_t H,
                           int32_t W, int32_t G, void* stream) {
    if (!img || !where || !d_glimpse || !d_where || N < 1 || H < 2 || W < 2 || G < 2) return fail(SQAIR_EINVAL, "bad argument");
    stn_glimpse_grad_kernel<<<N, 128, 0, (cudaStream_t)stream>>>(img, where, d_glimpse, d_where, H, W, G);
    CUDA_TRY(cudaGetLastError());
    return SQAIR_OK;
}

int sqair_canvas_ll(const float* glimpse, const float* where, const float* presence, const float* mean_img,
                    const float* img, float* canvas, float* data_ll, int32_t N, int32_t n, int32_t H, int32_t W,
                    int32_t G, float output_std, float bg_std, void* stream) {
    if (!glimpse || !where || !presence || !mean_img || !img || !canvas || !data_ll || N < 1 || n < 1)
        return fail(SQAIR_EINVAL, "bad argument");
    const int smem = (n * G * G + n * 5) * (int)sizeof(float);
    if (smem > 48 * 1024)
        CUDA_TRY(cudaFuncSetAttribute(canvas_ll_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    canvas_ll_kernel<<<N, 256, smem, (cudaStream_t)stream>>>(glimpse, where, presence, mean_img, img, canvas, data_ll, n, H,
                                                             W, G, output_std, bg_std);
    CUDA_TRY(cudaGetLastError());
    return SQAIR_OK;
}

int sqair_canvas_ll_grad(const float* glimpse, const float* where, const float* presence, const float* mean_img,
                         const float* img, const float* d_ll, float* d_glimpse, float* d_where, float* d_mean_img, int32_t N,
                         int32_t n, int32_t H, int32_t W, int32_t G, float output_std, float bg_std, void* stream) {
    if (!glimpse || !where || !presence || !mean_img || !img || !d_ll || !d_glimpse || !d_where || !d_mean_img || N < 1 || n < 1)
        return fail(SQAIR_EINVAL, "bad argument");
    const int smem = (2 * n * G * G + n * 11) * (int)sizeof(float);
    if (smem > kSmemLimit) return fail(SQAIR_EUNSUPPORTED, "glimpses do not fit shared memory");
    if (smem > 48 * 1024)
        CUDA_TRY(cudaFuncSetAttribute(canvas_ll_grad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    canvas_ll_grad_kernel<<<N, 256, smem, (cudaStream_t)stream>>>(glimpse, where, presence, mean_img, img, d_ll, d_glimpse,
                                                                  d_where, d_mean_img, n, H, W, G, output_std, bg_std);
    CUDA_TRY(cudaGetLastError());
    return SQAIR_OK;
}

}  // extern "C"
