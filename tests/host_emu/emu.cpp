// Host emulation of the CUDA kernel's per-block program (TEST INFRASTRUCTURE ONLY).
//
// Compiles sqair_b200/csrc/sqair_device.cuh with SQAIR_HOST_EMU: every thread block becomes one
// sequential host thread (tid 0 of 1), block barriers are no-ops, a cluster of C blocks becomes C
// host threads that exchange layer outputs through each other's "shared memory" arrays and meet at
// a split-phase barrier, and the tensor-core product is replaced by plain fp32 dot products over the
// same fragment-ordered packed panels and segment tables.  This lets the kernel's arithmetic, indexing, packing tables,
// column split and slot bookkeeping be checked against the oracle on a machine without a GPU.  It
// is never loaded by the sqair_b200 package; the product path is the CUDA library only.
#define SQAIR_HOST_EMU 1
#include <stdlib.h>
#include <algorithm>
#include <thread>
#include <vector>
struct float4 { float x, y, z, w; };
#include "../../sqair_b200/csrc/sqair_device.cuh"
#include "../../sqair_b200/csrc/sqair_backward.h"

using namespace sq;

static void pack_host(const Plan& plan, const std::vector<ParamEntry>& tab, const std::vector<Piece>& pieces,
                      const float* params, std::vector<float>& packed) {
    for (const auto& p : tab)
        for (int64_t i = 0; i < p.count; ++i) packed[p.packed_offset + i] = params[p.offset + i];
    for (const auto& pc : pieces) {
        const Layer& L = plan.L[pc.layer];
        for (int k = 0; k < pc.K; ++k)
            for (int n = 0; n < pc.N; ++n) {
                const int vrow = pc.vrow0 + k, vcol = pc.vcol0 + n, panel = vcol / L.Nc, cc = vcol - panel * L.Nc;
                packed[(size_t)L.w1_off + (size_t)panel * (L.panel_floats / 2) + frag_off(L.ksteps, cc >> 4, vrow >> 3, cc & 15, vrow & 7)] +=
                    params[pc.src_off + (int64_t)k * pc.src_ld + n];
            }
    }
    for (int id = 0; id < L_COUNT; ++id) {          // same split as split_panels_kernel
        const Layer& L = plan.L[id];
        if (L.nhead == 0) continue;
        for (int i = 0; i < L.npanel * (L.panel_floats / 2); ++i) {
            float hi, lo;
            split_weight(packed[(size_t)L.w1_off + i], hi, lo);
            const size_t d = (size_t)L.w_off + (size_t)(i >> 7) * 256 + (i & 127);
            packed[d] = hi;
            packed[d + 128] = lo;
        }
    }
}

template <int R>
static void run_clusters(const Plan& plan, const Job& job) {
    const int C = plan.C;
    const int ncl = (plan.rows + R - 1) / R;
    for (int cl = 0; cl < ncl; ++cl) {
        std::vector<std::vector<float>> smem(C, std::vector<float>(plan.sm.total + 64, 1e30f));   // poison
        std::vector<float*> peers(C);
        for (int q = 0; q < C; ++q) peers[q] = smem[q].data();
        EmuClusterBarrier cb;
        cb.n = C;
        auto body = [&](int rank) {
            Ctx c{};
            c.tid_ = 0; c.nthreads_ = 1; c.lane_ = 0; c.nlanes_ = 1; c.warp_ = 0; c.nwarps_ = 1; c.ncompute_ = 1;
            c.sm = peers[rank]; c.rank_ = rank; c.ncta_ = C;
            c.peers = peers.data(); c.cb = &cb; c.cb_gen = 0; c.plan = &plan;
            if (job.eps_where_prior) {                  // generation variant (sample_from_prior)
                Block<R, true, true> blk(c, job, cl * R);
                blk.run();
            } else {
                Block<R> blk(c, job, cl * R);
                blk.run();
            }
        };
        if (C == 1) {
            body(0);
        } else {
            std::vector<std::thread> th;
            for (int q = 0; q < C; ++q) th.emplace_back(body, q);
            for (auto& t : th) t.join();
        }
    }
}

extern "C" int emu_smem_floats(const sqair_cfg* cfg, int R, int C) {
    auto tab = param_table(*cfg);
    Plan plan;
    std::vector<Piece> pieces;
    int64_t total;
    if (!build_plan(*cfg, R, C, plan, tab, pieces, &total).empty()) return -1;
    return plan.sm.total;
}

static int emu_forward_impl(const sqair_cfg* cfg, const float* params, const float* obs,
                            const float* eps_where, const float* eps_what, const float* u_pres,
                            const sqair_outputs* out, int R, int C, float* stash, const float* const* prior_noise = nullptr,
                            int generate_after = -1) {
    std::string e = validate_cfg(*cfg);
    if (!e.empty()) { fprintf(stderr, "emu: %s\n", e.c_str()); return -1; }
    auto tab = param_table(*cfg);
    static Plan plan;
    std::vector<Piece> pieces;
    int64_t total = 0;
    e = build_plan(*cfg, R, C, plan, tab, pieces, &total);
    if (!e.empty()) { fprintf(stderr, "emu: %s\n", e.c_str()); return -1; }
    std::vector<float> packed(total, 0.f);
    pack_host(plan, tab, pieces, params, packed);
    Job job{packed.data(), obs, eps_where, eps_what, u_pres, *out, 0, nullptr, stash, prior_noise ? prior_noise[0] : nullptr,
            prior_noise ? prior_noise[1] : nullptr, prior_noise ? prior_noise[2] : nullptr, generate_after};
    switch (R) {
        case 1: run_clusters<1>(plan, job); break;
        case 2: run_clusters<2>(plan, job); break;
        case 3: run_clusters<3>(plan, job); break;
        case 4: run_clusters<4>(plan, job); break;
        case 5: run_clusters<5>(plan, job); break;
        case 6: run_clusters<6>(plan, job); break;
        default: fprintf(stderr, "emu: unsupported R=%d\n", R); return -2;
    }
    return 0;
}

extern "C" int emu_forward_generate(const sqair_cfg* cfg, const float* params, const float* obs, const float* eps_where,
                                    const float* eps_what, const float* u_pres, const float* eps_where_prior,
                                    const float* eps_what_prior, const float* u_pres_prior, int generate_after,
                                    const sqair_outputs* out, int R, int C) {
    const float* pn[3] = {eps_where_prior, eps_what_prior, u_pres_prior};
    return emu_forward_impl(cfg, params, obs, eps_where, eps_what, u_pres, out, R, C, nullptr, pn, generate_after);
}

extern "C" int emu_forward(const sqair_cfg* cfg, const float* params, const float* obs,
                           const float* eps_where, const float* eps_what, const float* u_pres,
                           const sqair_outputs* out, int R, int C) {
    return emu_forward_impl(cfg, params, obs, eps_where, eps_what, u_pres, out, R, C, nullptr);
}

// ---------------------------------------------------------------------------------------------
// backward pass on the host: the same stage functions and driver as the CUDA library (sqair_backward.h) with
// sequential loops in place of kernels
// ---------------------------------------------------------------------------------------------
struct HostEx {
    int tid = 0, nt = 1;
    float* scratch;
    void sync() {}
    void sum4(float (&)[4]) {}
    void warp_add4(float* dst, const float (&v)[4]) { for (int q = 0; q < 4; ++q) dst[q] += v[q]; }
};

struct HostBackend {
    const Plan* plan;
    const std::vector<Piece>* pieces;
    int rows;
    std::vector<float> scratch;
    long n_dgrad = 0, n_stage = 0, n_wgrad = 0;
    bool warned = false;

    const float* frame_gw = nullptr;   // debugging aid: per-frame upstream gradients [T][rows] (SQAIR_EMU_FRAME_GW)
    const float* frame_gp = nullptr;
    template <int STAGE>
    void stage(const BwdCtx& c, int t, int s) {
        ++n_stage;
        if (frame_gw) {
            BwdCtx& cc = const_cast<BwdCtx&>(c);
            cc.gw = frame_gw + (size_t)t * rows;
            cc.gp = frame_gp + (size_t)t * rows;
        }
        if (STAGE == BS_COMPACT && getenv("SQAIR_EMU_DUMP")) {        // debugging aid: d target / d z_t after the decoder
            char name[256];
            snprintf(name, sizeof(name), "%s.gz%d.bin", getenv("SQAIR_EMU_DUMP"), t);
            FILE* f = fopen(name, "wb");
            if (f) { fwrite(c.gZc, sizeof(float), (size_t)rows * c.n * c.zw, f); fclose(f); }
        }
        HostEx ex;
        ex.scratch = scratch.data();
        for (int row = 0; row < rows; ++row) bw_stage<STAGE>(c, ex, t, s, row);
    }
    static float* at(const Addr& a, int m, int ny) { return a.p + (size_t)(m / ny) * a.outer + (size_t)(m % ny) * a.inner; }
    void dgrad(const DgradArgs& A) {
        ++n_dgrad;
        std::vector<float> av(A.N);
        for (int m = 0; m < A.M; ++m) {
            const float* a = at(A.a, m, A.ny);
            const float* y = A.y.p ? at(A.y, m, A.ny) : nullptr;
            for (int j = 0; j < A.N; ++j) {
                av[j] = a[j] * (y ? act_deriv(A.act, y[j], A.act_scale, A.act_add) : 1.f);
                if (av[j] != av[j] && !warned) {
                    warned = true;
                    fprintf(stderr, "emu backward: NaN input of dgrad #%ld (layer %d, row %d, col %d; a %g y %g) after %ld stages\n", n_dgrad,
                            A.layer, m, j, a[j], y ? y[j] : 0.f, n_stage);
                }
            }
            if (A.dy.p) { float* d = at(A.dy, m, A.ny); for (int j = 0; j < A.N; ++j) d[j] = av[j]; }
            for (int si = 0; si < A.nseg; ++si) {
                const DgradArgs::Seg& S = A.seg[si];
                float* d = at(S.d, m, A.ny);
                for (int k = S.k0; k < S.k1; ++k) {
                    const float* w = A.w + (size_t)k * A.N;
                    double acc = 0.0;
                    for (int j = 0; j < A.N; ++j) acc += (double)av[j] * w[j];
                    if (S.mode == SEGM_STORE) d[k - S.k0] = (float)acc; else d[k - S.k0] += (float)acc;
                }
            }
        }
    }
    void wgrad(const WgradArgs& A) {
        ++n_wgrad;
        std::vector<double> acc((size_t)A.K * A.N, 0.0);
        for (int m = 0; m < A.M; ++m) {
            const float* x = at(A.x, m, A.ny);
            const float* d = at(A.dy, m, A.ny);
            for (int k = 0; k < A.K; ++k) {
                const double xv = x[k];
                if (xv == 0.0) continue;
                double* a = &acc[(size_t)k * A.N];
                for (int j = 0; j < A.N; ++j) a[j] += xv * d[j];
            }
        }
        for (int k = 0; k < A.K; ++k)
            for (int j = 0; j < A.N; ++j) A.dw[(size_t)k * A.ldw + j] += (float)acc[(size_t)k * A.N + j];
    }
    void colsum(const ColsumArgs& A) {
        std::vector<double> acc(A.N, 0.0);
        for (int m = 0; m < A.M; ++m) {
            const float* d = at(A.dy, m, A.ny);
            for (int j = 0; j < A.N; ++j) acc[j] += d[j];
        }
        for (int j = 0; j < A.N; ++j) A.out[j] += (float)acc[j];
    }
    void zero(float* p, int64_t n) { memset(p, 0, (size_t)n * sizeof(float)); }
    void img_reduce(const float* dy, float* out, int TB, int K, int nh) {
        for (int i = 0; i < TB; ++i)
            for (int j = 0; j < nh; ++j) {
                float a = 0.f;
                for (int k = 0; k < K; ++k) a += dy[((size_t)i * K + k) * nh + j];
                out[(size_t)i * nh + j] = a;
            }
    }
    void unpack(const float* dwv, float* d_params) {
        for (const auto& pc : *pieces) {
            const LayerB& LB = plan->LB[pc.layer];
            for (int k = 0; k < pc.K; ++k)
                for (int j = 0; j < pc.N; ++j)
                    d_params[pc.src_off + (int64_t)k * pc.src_ld + j] += dwv[LB.bw_off + (int64_t)(pc.urow0 + k) * LB.NU + pc.ucol0 + j];
        }
    }
    void small_to_params(const float* small, float* d_params, const POff& po) {
        for (int i = 0; i < 10; ++i) d_params[po.cholesky + i] += small[SM_CHOL + i];
        d_params[po.d_scale_offset] += small[SM_DSO];
        d_params[po.p_scale_offset] += small[SM_PSO];
        d_params[po.output_scale] += small[SM_OUTSCALE];
    }
};

// canonical parameters -> backward parameter buffer (virtual matrices); same table as sqair_pack_backward
static void pack_backward_host(const Plan& plan, const std::vector<Piece>& pieces, const float* params, std::vector<float>& bw) {
    bw.assign((size_t)plan.bw_total, 0.f);
    for (const auto& pc : pieces) {
        const LayerB& LB = plan.LB[pc.layer];
        for (int k = 0; k < pc.K; ++k)
            for (int j = 0; j < pc.N; ++j)
                bw[LB.bw_off + (int64_t)(pc.urow0 + k) * LB.NU + pc.ucol0 + j] += params[pc.src_off + (int64_t)k * pc.src_ld + j];
    }
}

extern "C" int64_t emu_stash_floats(const sqair_cfg* cfg) { return build_stash(*cfg).total; }

// forward (with stash) + backward on the host.  d_log_w / d_disc_lp: [rows] upstream gradients of the objective.
extern "C" int emu_forward_backward(const sqair_cfg* cfg, const float* params, const float* obs, const float* eps_where,
                                    const float* eps_what, const float* u_pres, const sqair_outputs* out, int R, int C,
                                    const float* d_log_w, const float* d_disc_lp, float* d_params) {
    const StashLayout SL = build_stash(*cfg);
    if (SL.total < 0) return -3;
    std::vector<float> stash((size_t)SL.total, NAN);
    int rc = emu_forward_impl(cfg, params, obs, eps_where, eps_what, u_pres, out, R, C, stash.data());
    if (rc) return rc;
    auto tab = param_table(*cfg);
    static Plan plan;
    std::vector<Piece> pieces;
    int64_t total = 0;
    std::string e = build_plan(*cfg, 1, 1, plan, tab, pieces, &total);
    if (!e.empty()) { fprintf(stderr, "emu: %s\n", e.c_str()); return -1; }
    std::vector<float> bw;
    pack_backward_host(plan, pieces, params, bw);
    const BwdLayout BL = build_bwd_layout(*cfg, plan);
    std::vector<float> ws((size_t)BL.total, NAN);
    BwdInputs in;
    in.params = params; in.bw = bw.data(); in.obs = obs; in.eps_where = eps_where; in.eps_what = eps_what;
    in.stash = stash.data(); in.d_log_w = d_log_w; in.d_disc_lp = d_disc_lp; in.ws = ws.data(); in.d_params = d_params; in.vimco = 1;
    HostBackend be;
    be.plan = &plan; be.pieces = &pieces; be.rows = cfg->B * cfg->K;
    be.scratch.assign(bw_stage_scratch_floats(*cfg), 0.f);
    if (getenv("SQAIR_EMU_FRAME_GW")) { be.frame_gw = d_log_w; be.frame_gp = d_disc_lp; }
    BwdDriver<HostBackend> drv(be, *cfg, plan, plan.poc, BL, in);
    drv.param_count_ = tab.back().offset + tab.back().count;
    drv.run(d_params);
    if (getenv("SQAIR_EMU_VERBOSE"))
        fprintf(stderr, "emu backward: %ld stages, %ld dgrad, %ld wgrad calls; stash %.1f MB, workspace %.1f MB\n", be.n_stage, be.n_dgrad,
                be.n_wgrad, SL.total * 4e-6, BL.total * 4e-6);
    return 0;
}

// plan dump for tuning scripts: per layer id {ksteps, Nc, nmt, ksplit, kper, npanel, split}
extern "C" int emu_dump_plan(const sqair_cfg* cfg, int R, int C, int* out /* L_COUNT x 8 */) {
    auto tab = param_table(*cfg);
    static Plan plan;
    std::vector<Piece> pieces;
    int64_t total;
    if (!build_plan(*cfg, R, C, plan, tab, pieces, &total).empty()) return -1;
    for (int i = 0; i < L_COUNT; ++i) {
        const Layer& L = plan.L[i];
        int* o = out + 8 * i;
        o[0] = L.ksteps; o[1] = L.Nc; o[2] = L.nmt; o[3] = L.ksplit; o[4] = L.kper; o[5] = L.npanel; o[6] = L.split; o[7] = L.Ntot;
    }
    return L_COUNT;
}
