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
            Block<R> blk(c, job, cl * R);
            blk.run();
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

extern "C" int emu_forward(const sqair_cfg* cfg, const float* params, const float* obs,
                           const float* eps_where, const float* eps_what, const float* u_pres,
                           const sqair_outputs* out, int R, int C) {
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
    Job job{packed.data(), obs, eps_where, eps_what, u_pres, *out, 0};
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
